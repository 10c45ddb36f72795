"""In-tree build of the CUDA extension (sm_100a only) with plain nvcc.

    python -m bayesembler_b200._build        # -> bayesembler_b200/libbayesembler_b200.so

The shared library carries the C ABI of include/bayesembler_b200.h.  It is built in-tree so that it travels to
the GPU box with the repo snapshot; nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbayesembler_b200.so")
CLI = os.path.join(HERE, "bayesembler_b200_cli")
CLI_SOURCES = [os.path.join(HERE, "host", f) for f in ("main.cpp", "assembler.hpp", "bayesembler_b200.hpp")]
SOURCES = ["capi.cu", "gibbs_kernel.cu", "gibbs_fast.cu", "gibbs_rows.cu", "host_steps.cpp", "pack.cpp", "likelihood.cpp", "allpaths.cpp", "bamfront.cpp"]
HEADERS = ["internal.h", "layout.h", "rng.cuh", "gibbs_common.cuh", os.path.join("..", "..", "include", "bayesembler_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def nvcc_command(out=LIB, extra=()):
    # -fmad=false: keep the sampler's own double arithmetic un-contracted so that it rounds like the CPU oracle
    # (the CUDA math library's log/exp/cos use explicit fma internally and are unaffected).
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-ffp-contract=off", "-shared", "-o", out]
    if os.environ.get("BAYES_NVCC_FLAGS"):  # experiment knob, e.g. -DBAYES_BINV_NQ=24.0
        cmd += os.environ["BAYES_NVCC_FLAGS"].split()
    cmd += list(extra)
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lz"]  # the BAM front end inflates BGZF blocks with zlib
    return cmd


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = nvcc_command(extra=("-Xptxas", "-v") if verbose else ())
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % LIB)
    return LIB


def build_cli(force=False):
    """The command line of the host driver (host/main.cpp), linked against the library next to it."""
    build()
    deps = CLI_SOURCES + [LIB, os.path.join(HERE, "..", "include", "bayesembler_b200.h")]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    cmd = ["/usr/bin/g++", "-O2", "-std=gnu++17", "-Wall", "-Werror", "-o", CLI, CLI_SOURCES[0], "-L" + HERE, "-lbayesembler_b200",
           "-pthread", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("g++ failed building %s" % CLI)
    return CLI


# development builds of the same sources, loaded instead of the product library through BAYES_B200_LIB=<path>
VARIANTS = {"debug": ("-DBAYES_DEBUG_BOUNDS",),                       # every index derived from packed data is checked on the device
            "phases": ("-DBAYES_ROWS_PHASES", "-DBAYES_TILE_PHASES")}  # per-phase clocks of an iteration, printed when a unit ends


def build_variant(name):
    out = os.path.join(HERE, "libbayes_%s.so" % name)
    res = subprocess.run(nvcc_command(out=out, extra=VARIANTS[name]), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building %s" % out)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2]))
    else:
        build(force=True, verbose=True)
        build_cli(force=True)
        print(LIB)
