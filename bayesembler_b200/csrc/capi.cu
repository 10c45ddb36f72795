// C ABI of the B200 Gibbs sampler (include/bayesembler_b200.h): context, host packing on a thread pool, device
// buffers, launches.  Everything that computes on the device is in gibbs_kernel.cu; nothing here falls back to the CPU.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <numeric>
#include <thread>

#include "internal.h"

namespace bayes {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

namespace {

int cuda_rc(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BAYES_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return BAYES_E_CUDA;
}
#define CU(call)                                   \
    do {                                           \
        int _rc = cuda_rc((call), #call);          \
        if (_rc != BAYES_OK) return _rc;           \
    } while (0)

// growable device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return BAYES_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(n + n / 8, 256);
        CU(cudaMalloc((void **)&p, want * sizeof(T)));
        cap = want;
        return BAYES_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// growable pinned host buffer
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0, n = 0;
    int resize(size_t want) {
        if (want > cap) {
            T *q = nullptr;
            const size_t c = std::max<size_t>(want + want / 4, 256);
            CU(cudaMallocHost((void **)&q, c * sizeof(T)));
            if (p) {
                memcpy(q, p, n * sizeof(T));
                cudaFreeHost(p);
            }
            p = q;
            cap = c;
        }
        n = want;
        return BAYES_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = n = 0;
    }
};

struct LocusHost {
    int32_t T, n_chains, samples, cover, folded;
    double pi;
    int64_t first_job;
    // T == 1 closed form (gibbsSampler.cpp:57-67): no device work
    bool closed_form;
    double single_fpkm;
    int32_t n_frag;
};

struct Launch {
    int64_t job_base, n_jobs;
    int block;
    bool mat_smem;
    size_t smem;
};

}  // namespace
}  // namespace bayes

using namespace bayes;

struct bayes_gibbs_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int host_threads = 0;
    bool uploaded = false, ran = false, downloaded = false;

    // staging (pinned), concatenated over loci
    PinBuf<LocusDev> h_loci;
    PinBuf<JobDev> h_jobs;
    PinBuf<double> h_val, h_efflen, h_tab_val;
    PinBuf<uint16_t> h_col;
    PinBuf<int32_t> h_slab, h_row_n, h_row_nnz, h_row_id, h_base, h_tab_row;
    // results (pinned)
    PinBuf<int32_t> h_occ, h_counts;
    PinBuf<double> h_mean, h_m2;

    DevBuf<LocusDev> d_loci;
    DevBuf<JobDev> d_jobs;
    DevBuf<double> d_val, d_efflen, d_tab_val, d_mean, d_m2, d_tr_theta;
    DevBuf<uint16_t> d_col;
    DevBuf<int32_t> d_slab, d_row_n, d_row_nnz, d_row_id, d_base, d_tab_row, d_occ, d_counts, d_tr_counts, d_tr_b, d_tr_init;

    std::vector<LocusHost> loci;       // per submitted locus
    std::vector<int32_t> dev_index;    // locus -> index in h_loci (-1 for closed form)
    std::vector<double> job_cost;      // per job, submission order
    std::vector<int32_t> job_block;
    std::vector<Launch> launches;
    int64_t out_len = 0;               // total T over jobs
    int32_t last_launches = 0;

    int64_t trace_locus = -1;
    int32_t trace_chain = 0, trace_iters = 0, trace_T = 0;
    int64_t trace_job_sorted = -1;
};

namespace {

int use_device(bayes_gibbs_ctx *ctx) { return cuda_rc(cudaSetDevice(ctx->device), "cudaSetDevice"); }

template <typename T>
int append(PinBuf<T> &dst, const std::vector<T> &src) {
    const size_t at = dst.n;
    int rc = dst.resize(at + src.size());
    if (rc) return rc;
    if (!src.empty()) memcpy(dst.p + at, src.data(), src.size() * sizeof(T));
    return BAYES_OK;
}

template <typename T>
int h2d(DevBuf<T> &d, const PinBuf<T> &h, cudaStream_t st) {
    int rc = d.reserve(h.n);
    if (rc) return rc;
    if (h.n) CU(cudaMemcpyAsync(d.p, h.p, h.n * sizeof(T), cudaMemcpyHostToDevice, st));
    return BAYES_OK;
}

// Order jobs so that each launch is one (block size, matrix placement) class with its most expensive jobs first: the
// hardware block scheduler hands blocks out in index order, which makes this the largest-first queue of the reference
// (assembler.cpp:1886-1916).
int plan_launches(bayes_gibbs_ctx *ctx) {
    const int64_t n = (int64_t)ctx->h_jobs.n;
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    JobDev *jobs = ctx->h_jobs.p;
    const LocusDev *loci = ctx->h_loci.p;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const LocusDev &la = loci[jobs[a].locus], &lb = loci[jobs[b].locus];
        if (la.mat_smem != lb.mat_smem) return la.mat_smem < lb.mat_smem;  // streaming (largest) loci first
        if (ctx->job_block[a] != ctx->job_block[b]) return ctx->job_block[a] > ctx->job_block[b];
        return ctx->job_cost[a] > ctx->job_cost[b];
    });
    std::vector<JobDev> sorted(n);
    std::vector<double> cost(n);
    std::vector<int32_t> block(n);
    for (int64_t i = 0; i < n; i++) {
        sorted[i] = jobs[order[i]];
        cost[i] = ctx->job_cost[order[i]];
        block[i] = ctx->job_block[order[i]];
    }
    memcpy(jobs, sorted.data(), n * sizeof(JobDev));
    ctx->job_cost.swap(cost);
    ctx->job_block.swap(block);
    ctx->launches.clear();
    for (int64_t i = 0; i < n;) {
        const LocusDev &l0 = loci[jobs[i].locus];
        int64_t j = i;
        size_t smem = 0;
        while (j < n && ctx->job_block[j] == ctx->job_block[i] && loci[jobs[j].locus].mat_smem == l0.mat_smem) {
            smem = std::max(smem, (size_t)loci[jobs[j].locus].smem_bytes);
            j++;
        }
        ctx->launches.push_back(Launch{i, j - i, ctx->job_block[i], l0.mat_smem != 0, smem});
        i = j;
    }
    // first_job / trace bookkeeping refer to (locus, chain): rebuild the lookup
    ctx->trace_job_sorted = -1;
    for (auto &l : ctx->loci) l.first_job = -1;
    return BAYES_OK;
}

KernelArgs kernel_args(bayes_gibbs_ctx *ctx) {
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.loci = ctx->d_loci.p;
    a.jobs = ctx->d_jobs.p;
    a.val = ctx->d_val.p;
    a.col = ctx->d_col.p;
    a.slab_cell_off = ctx->d_slab.p;
    a.row_n = ctx->d_row_n.p;
    a.row_nnz = ctx->d_row_nnz.p;
    a.row_id = ctx->d_row_id.p;
    a.efflen = ctx->d_efflen.p;
    a.base_counts = ctx->d_base.p;
    a.tab_val = ctx->d_tab_val.p;
    a.tab_row = ctx->d_tab_row.p;
    a.out_occ = ctx->d_occ.p;
    a.out_mean = ctx->d_mean.p;
    a.out_m2 = ctx->d_m2.p;
    a.out_counts = ctx->d_counts.p;
    a.trace_job = ctx->trace_job_sorted;
    a.trace_iters = ctx->trace_iters;
    a.tr_theta = ctx->d_tr_theta.p;
    a.tr_counts = ctx->d_tr_counts.p;
    a.tr_b = ctx->d_tr_b.p;
    a.tr_init = ctx->d_tr_init.p;
    return a;
}

}  // namespace

extern "C" {

const char *bayes_last_error(void) { return g_last_error.c_str(); }

int bayes_gibbs_create(int device, bayes_gibbs_ctx **out) {
    if (!out) {
        set_error("bayes_gibbs_create: null out pointer");
        return BAYES_E_INVALID;
    }
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev < 1) {
        set_error(std::string("bayes_gibbs_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
        return BAYES_E_NODEVICE;
    }
    if (device < 0 || device >= n_dev) {
        set_error("bayes_gibbs_create: device index out of range");
        return BAYES_E_INVALID;
    }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("bayes_gibbs_create: kernels are built for sm_100a only (found sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        return BAYES_E_NODEVICE;
    }
    int rc = configure_kernels();
    if (rc) return rc;
    bayes_gibbs_ctx *ctx = new (std::nothrow) bayes_gibbs_ctx();
    if (!ctx) return BAYES_E_NOMEM;
    ctx->device = device;
    rc = cuda_rc(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (rc) {
        delete ctx;
        return rc;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return BAYES_OK;
}

void bayes_gibbs_destroy(bayes_gibbs_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->h_loci.release(); ctx->h_jobs.release(); ctx->h_val.release(); ctx->h_efflen.release(); ctx->h_tab_val.release();
    ctx->h_col.release(); ctx->h_slab.release(); ctx->h_row_n.release(); ctx->h_row_nnz.release(); ctx->h_row_id.release();
    ctx->h_base.release(); ctx->h_tab_row.release(); ctx->h_occ.release(); ctx->h_counts.release(); ctx->h_mean.release();
    ctx->h_m2.release();
    ctx->d_loci.release(); ctx->d_jobs.release(); ctx->d_val.release(); ctx->d_efflen.release(); ctx->d_tab_val.release();
    ctx->d_mean.release(); ctx->d_m2.release(); ctx->d_tr_theta.release(); ctx->d_col.release(); ctx->d_slab.release();
    ctx->d_row_n.release(); ctx->d_row_nnz.release(); ctx->d_row_id.release(); ctx->d_base.release(); ctx->d_tab_row.release();
    ctx->d_occ.release(); ctx->d_counts.release(); ctx->d_tr_counts.release(); ctx->d_tr_b.release(); ctx->d_tr_init.release();
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int bayes_gibbs_set_stream(bayes_gibbs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return BAYES_E_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return BAYES_OK;
}

void *bayes_gibbs_get_stream(bayes_gibbs_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int bayes_gibbs_set_host_threads(bayes_gibbs_ctx *ctx, int n_threads) {
    if (!ctx || n_threads < 0) return BAYES_E_INVALID;
    ctx->host_threads = n_threads;
    return BAYES_OK;
}

int bayes_gibbs_clear(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    int rc = use_device(ctx);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->h_loci.n = ctx->h_jobs.n = ctx->h_val.n = ctx->h_efflen.n = ctx->h_tab_val.n = ctx->h_col.n = 0;
    ctx->h_slab.n = ctx->h_row_n.n = ctx->h_row_nnz.n = ctx->h_row_id.n = ctx->h_base.n = ctx->h_tab_row.n = 0;
    ctx->loci.clear();
    ctx->dev_index.clear();
    ctx->job_cost.clear();
    ctx->job_block.clear();
    ctx->launches.clear();
    ctx->out_len = 0;
    ctx->uploaded = ctx->ran = ctx->downloaded = false;
    ctx->trace_locus = -1;
    ctx->trace_job_sorted = -1;
    ctx->trace_iters = 0;
    return BAYES_OK;
}

int bayes_gibbs_submit(bayes_gibbs_ctx *ctx, int32_t n_loci, const bayes_locus_desc *loci, int64_t *first_id) {
    if (!ctx || n_loci < 0 || (n_loci > 0 && !loci)) {
        set_error("bayes_gibbs_submit: bad arguments");
        return BAYES_E_INVALID;
    }
    if (ctx->uploaded) {
        set_error("bayes_gibbs_submit: batch already uploaded; call bayes_gibbs_clear first");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    if (first_id) *first_id = (int64_t)ctx->loci.size();

    // pack on a thread pool (set cover + CDF table + SELL packing are independent per locus)
    std::vector<PackedLocus> packed(n_loci);
    std::vector<int> prc(n_loci, 0);
    std::vector<std::string> perr(n_loci);
    int n_thr = ctx->host_threads > 0 ? ctx->host_threads : (int)std::thread::hardware_concurrency();
    n_thr = std::max(1, std::min(n_thr, (int)n_loci));
    std::atomic<int32_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int32_t i = next.fetch_add(1);
            if (i >= n_loci) return;
            prc[i] = pack_locus(loci[i], &packed[i], &perr[i]);
        }
    };
    if (n_thr == 1) work();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; t++) pool.emplace_back(work);
        for (auto &t : pool) t.join();
    }
    for (int32_t i = 0; i < n_loci; i++)
        if (prc[i] != BAYES_OK) {
            set_error("bayes_gibbs_submit: locus " + std::to_string(i) + ": " + perr[i]);
            return prc[i];
        }

    for (int32_t i = 0; i < n_loci; i++) {
        PackedLocus &P = packed[i];
        LocusHost lh;
        lh.T = P.d.T;
        lh.n_chains = P.n_chains;
        lh.samples = P.d.samples;
        lh.cover = P.cover;
        lh.folded = P.folded;
        lh.pi = P.pi;
        lh.first_job = -1;
        lh.n_frag = P.d.N_f;
        lh.closed_form = P.d.T == 1;
        lh.single_fpkm = 0;
        if (lh.closed_form) {
            // gibbsSampler.cpp:57-67: (num_fragments * 1e9) / (total_num_fragments * effective_length)
            lh.single_fpkm = (P.d.N_f * 1e9) / (P.d.N_total * P.efflen[0]);
            ctx->dev_index.push_back(-1);
            ctx->loci.push_back(lh);
            continue;
        }
        P.d.slab_off = (int64_t)ctx->h_slab.n;
        P.d.cell_off = (int64_t)ctx->h_val.n;
        P.d.row_off = (int64_t)ctx->h_row_n.n;
        P.d.t_off = (int64_t)ctx->h_efflen.n;
        P.d.tab_off = (int64_t)ctx->h_tab_val.n;
        P.d.tabrow_off = (int64_t)ctx->h_tab_row.n;
        if ((rc = append(ctx->h_slab, P.slab_cell_off)) || (rc = append(ctx->h_val, P.val)) || (rc = append(ctx->h_col, P.col)) ||
            (rc = append(ctx->h_row_n, P.row_n)) || (rc = append(ctx->h_row_nnz, P.row_nnz)) || (rc = append(ctx->h_row_id, P.row_id)) ||
            (rc = append(ctx->h_efflen, P.efflen)) || (rc = append(ctx->h_base, P.base_counts)) ||
            (rc = append(ctx->h_tab_val, P.tab_val)) || (rc = append(ctx->h_tab_row, P.tab_row)))
            return rc;
        const int32_t li = (int32_t)ctx->h_loci.n;
        if ((rc = ctx->h_loci.resize(li + 1))) return rc;
        ctx->h_loci.p[li] = P.d;
        ctx->dev_index.push_back(li);
        const int64_t locus_id = (int64_t)ctx->loci.size();
        for (int32_t c = 0; c < P.n_chains; c++) {
            const size_t j = ctx->h_jobs.n;
            if ((rc = ctx->h_jobs.resize(j + 1))) return rc;
            JobDev jd;
            jd.locus = li;
            jd.chain = c;
            jd.key = bayes_chain_key(P.seed, c);
            jd.out_off = ctx->out_len;
            ctx->h_jobs.p[j] = jd;
            ctx->out_len += P.d.T;
            ctx->job_cost.push_back(P.cost);
            ctx->job_block.push_back(P.block);
        }
        (void)locus_id;
        ctx->loci.push_back(lh);
    }
    return BAYES_OK;
}

int bayes_gibbs_upload(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    int rc = use_device(ctx);
    if (rc) return rc;
    if ((rc = plan_launches(ctx))) return rc;
    // (locus, chain) -> sorted job index: out_off is stable, so fetch looks results up through the host copy of jobs
    const int64_t n_jobs = (int64_t)ctx->h_jobs.n;
    std::vector<int64_t> first(ctx->h_loci.n, -1);
    for (int64_t j = 0; j < n_jobs; j++) {
        const JobDev &jd = ctx->h_jobs.p[j];
        if (jd.chain == 0) first[jd.locus] = jd.out_off;
    }
    for (size_t l = 0; l < ctx->loci.size(); l++)
        if (ctx->dev_index[l] >= 0) ctx->loci[l].first_job = first[ctx->dev_index[l]];
    if (ctx->trace_locus >= 0) {
        const int32_t li = ctx->dev_index[ctx->trace_locus];
        for (int64_t j = 0; j < n_jobs; j++)
            if (ctx->h_jobs.p[j].locus == li && ctx->h_jobs.p[j].chain == ctx->trace_chain) ctx->trace_job_sorted = j;
        const size_t n = (size_t)ctx->trace_iters * ctx->trace_T;
        if ((rc = ctx->d_tr_theta.reserve(n)) || (rc = ctx->d_tr_counts.reserve(n)) || (rc = ctx->d_tr_b.reserve(ctx->trace_iters)) ||
            (rc = ctx->d_tr_init.reserve(ctx->trace_T)))
            return rc;
    }
    cudaStream_t st = ctx->stream;
    if ((rc = h2d(ctx->d_loci, ctx->h_loci, st)) || (rc = h2d(ctx->d_jobs, ctx->h_jobs, st)) || (rc = h2d(ctx->d_val, ctx->h_val, st)) ||
        (rc = h2d(ctx->d_col, ctx->h_col, st)) || (rc = h2d(ctx->d_slab, ctx->h_slab, st)) || (rc = h2d(ctx->d_row_n, ctx->h_row_n, st)) ||
        (rc = h2d(ctx->d_row_nnz, ctx->h_row_nnz, st)) || (rc = h2d(ctx->d_row_id, ctx->h_row_id, st)) ||
        (rc = h2d(ctx->d_efflen, ctx->h_efflen, st)) || (rc = h2d(ctx->d_base, ctx->h_base, st)) ||
        (rc = h2d(ctx->d_tab_val, ctx->h_tab_val, st)) || (rc = h2d(ctx->d_tab_row, ctx->h_tab_row, st)))
        return rc;
    const size_t n_out = (size_t)ctx->out_len;
    if ((rc = ctx->d_occ.reserve(n_out)) || (rc = ctx->d_counts.reserve(n_out)) || (rc = ctx->d_mean.reserve(n_out)) ||
        (rc = ctx->d_m2.reserve(n_out)))
        return rc;
    if ((rc = ctx->h_occ.resize(n_out)) || (rc = ctx->h_counts.resize(n_out)) || (rc = ctx->h_mean.resize(n_out)) ||
        (rc = ctx->h_m2.resize(n_out)))
        return rc;
    ctx->uploaded = true;
    ctx->ran = ctx->downloaded = false;
    return BAYES_OK;
}

int bayes_gibbs_run(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    if (!ctx->uploaded) {
        set_error("bayes_gibbs_run: call bayes_gibbs_upload first");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    const KernelArgs args = kernel_args(ctx);
    ctx->last_launches = 0;
    for (const Launch &l : ctx->launches) {
        if ((rc = launch_gibbs(args, l.job_base, l.n_jobs, l.block, l.mat_smem, l.smem, ctx->stream))) return rc;
        ctx->last_launches++;
    }
    ctx->ran = true;
    ctx->downloaded = false;
    return BAYES_OK;
}

int bayes_gibbs_download(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    if (!ctx->ran) {
        set_error("bayes_gibbs_download: call bayes_gibbs_run first");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    const size_t n = (size_t)ctx->out_len;
    if (n) {
        CU(cudaMemcpyAsync(ctx->h_occ.p, ctx->d_occ.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_mean.p, ctx->d_mean.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_m2.p, ctx->d_m2.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_counts.p, ctx->d_counts.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->downloaded = true;
    return BAYES_OK;
}

int bayes_gibbs_sync(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    int rc = use_device(ctx);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return BAYES_OK;
}

int bayes_gibbs_run_batch(bayes_gibbs_ctx *ctx, int32_t n_loci, const bayes_locus_desc *loci, int64_t *first_id) {
    int rc;
    if ((rc = bayes_gibbs_submit(ctx, n_loci, loci, first_id))) return rc;
    if ((rc = bayes_gibbs_upload(ctx))) return rc;
    if ((rc = bayes_gibbs_run(ctx))) return rc;
    if ((rc = bayes_gibbs_download(ctx))) return rc;
    return bayes_gibbs_sync(ctx);
}

int bayes_gibbs_fetch(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, int32_t *occ, double *mean, double *m2,
                      int32_t *final_counts) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size()) {
        set_error("bayes_gibbs_fetch: bad locus id");
        return BAYES_E_INVALID;
    }
    const LocusHost &lh = ctx->loci[locus_id];
    if (chain < 0 || chain >= lh.n_chains) {
        set_error("bayes_gibbs_fetch: bad chain");
        return BAYES_E_INVALID;
    }
    if (lh.closed_form) {
        // every one of the `samples` identical samples is accumulated (gibbsSampler.cpp:63-66, gibbsOutput.cpp:51-75)
        if (occ) occ[0] = lh.samples;
        if (mean) mean[0] = lh.single_fpkm;
        if (m2) m2[0] = 0.0;
        if (final_counts) final_counts[0] = lh.n_frag;
        return BAYES_OK;
    }
    if (!ctx->downloaded) {
        set_error("bayes_gibbs_fetch: results not downloaded (run, download, sync first)");
        return BAYES_E_STATE;
    }
    const int64_t off = lh.first_job + (int64_t)chain * lh.T;  // chains of a locus are contiguous in the outputs
    if (occ) memcpy(occ, ctx->h_occ.p + off, lh.T * sizeof(int32_t));
    if (mean) memcpy(mean, ctx->h_mean.p + off, lh.T * sizeof(double));
    if (m2) memcpy(m2, ctx->h_m2.p + off, lh.T * sizeof(double));
    if (final_counts) memcpy(final_counts, ctx->h_counts.p + off, lh.T * sizeof(int32_t));
    return BAYES_OK;
}

int bayes_gibbs_locus_info(bayes_gibbs_ctx *ctx, int64_t locus_id, double *pi, int32_t *cover_size, int32_t *rows_folded) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size()) {
        set_error("bayes_gibbs_locus_info: bad locus id");
        return BAYES_E_INVALID;
    }
    const LocusHost &lh = ctx->loci[locus_id];
    if (pi) *pi = lh.pi;
    if (cover_size) *cover_size = lh.cover;
    if (rows_folded) *rows_folded = lh.folded;
    return BAYES_OK;
}

int64_t bayes_gibbs_num_loci(bayes_gibbs_ctx *ctx) { return ctx ? (int64_t)ctx->loci.size() : 0; }
int64_t bayes_gibbs_num_jobs(bayes_gibbs_ctx *ctx) { return ctx ? (int64_t)ctx->h_jobs.n : 0; }
int32_t bayes_gibbs_last_launches(bayes_gibbs_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

int bayes_gibbs_set_trace(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, int32_t trace_iters) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size() || trace_iters < 1) {
        set_error("bayes_gibbs_set_trace: bad arguments");
        return BAYES_E_INVALID;
    }
    if (ctx->uploaded) {
        set_error("bayes_gibbs_set_trace: must be called before bayes_gibbs_upload");
        return BAYES_E_STATE;
    }
    const LocusHost &lh = ctx->loci[locus_id];
    if (lh.closed_form || chain < 0 || chain >= lh.n_chains) {
        set_error("bayes_gibbs_set_trace: locus has no device job for that chain");
        return BAYES_E_INVALID;
    }
    ctx->trace_locus = locus_id;
    ctx->trace_chain = chain;
    ctx->trace_iters = trace_iters;
    ctx->trace_T = lh.T;
    return BAYES_OK;
}

int bayes_gibbs_fetch_trace(bayes_gibbs_ctx *ctx, double *theta_trace, int32_t *counts_trace, int32_t *b_trace, int32_t *init_counts) {
    if (!ctx || ctx->trace_locus < 0 || !ctx->ran) {
        set_error("bayes_gibbs_fetch_trace: no traced run");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)ctx->trace_iters * ctx->trace_T;
    if (theta_trace) CU(cudaMemcpy(theta_trace, ctx->d_tr_theta.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (counts_trace) CU(cudaMemcpy(counts_trace, ctx->d_tr_counts.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (b_trace) CU(cudaMemcpy(b_trace, ctx->d_tr_b.p, ctx->trace_iters * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (init_counts) CU(cudaMemcpy(init_counts, ctx->d_tr_init.p, ctx->trace_T * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return BAYES_OK;
}

// ---- single steps --------------------------------------------------------------------------------------------------
static int step_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key, uint32_t iteration,
                           int32_t *counts_out, bool init) {
    if (!locus || !counts_out || (!init && !theta)) {
        set_error("bayes_step_*_assignment: null argument");
        return BAYES_E_INVALID;
    }
    bayes_gibbs_ctx *ctx = nullptr;
    int rc = bayes_gibbs_create(device, &ctx);
    if (rc) return rc;
    bayes_locus_desc d = *locus;
    if (d.pi < 0) d.pi = 0.5;  // the cover is not part of this step
    d.n_chains = 1;
    if (d.burn_in < 0) d.burn_in = 0;
    if (d.samples < 1) d.samples = 1;
    PackedLocus P;
    std::string err;
    if ((rc = pack_locus(d, &P, &err))) {
        set_error("bayes_step_*_assignment: " + err);
        bayes_gibbs_destroy(ctx);
        return rc;
    }
    const int T = d.T;
    auto fail = [&](int code) { bayes_gibbs_destroy(ctx); return code; };
    P.d.slab_off = P.d.cell_off = P.d.row_off = P.d.t_off = P.d.tab_off = P.d.tabrow_off = 0;
    JobDev jd;
    jd.locus = 0; jd.chain = 0; jd.key = key; jd.out_off = 0;
    if ((rc = ctx->h_loci.resize(1)) || (rc = ctx->h_jobs.resize(1))) return fail(rc);
    ctx->h_loci.p[0] = P.d;
    ctx->h_jobs.p[0] = jd;
    if ((rc = append(ctx->h_slab, P.slab_cell_off)) || (rc = append(ctx->h_val, P.val)) || (rc = append(ctx->h_col, P.col)) ||
        (rc = append(ctx->h_row_n, P.row_n)) || (rc = append(ctx->h_row_nnz, P.row_nnz)) || (rc = append(ctx->h_row_id, P.row_id)) ||
        (rc = append(ctx->h_base, P.base_counts)))
        return fail(rc);
    cudaStream_t st = ctx->stream;
    if ((rc = h2d(ctx->d_loci, ctx->h_loci, st)) || (rc = h2d(ctx->d_jobs, ctx->h_jobs, st)) || (rc = h2d(ctx->d_val, ctx->h_val, st)) ||
        (rc = h2d(ctx->d_col, ctx->h_col, st)) || (rc = h2d(ctx->d_slab, ctx->h_slab, st)) || (rc = h2d(ctx->d_row_n, ctx->h_row_n, st)) ||
        (rc = h2d(ctx->d_row_nnz, ctx->h_row_nnz, st)) || (rc = h2d(ctx->d_row_id, ctx->h_row_id, st)) ||
        (rc = h2d(ctx->d_base, ctx->h_base, st)))
        return fail(rc);
    if ((rc = ctx->d_mean.reserve(T)) || (rc = ctx->d_occ.reserve(T))) return fail(rc);
    if (!init && (rc = cuda_rc(cudaMemcpyAsync(ctx->d_mean.p, theta, T * sizeof(double), cudaMemcpyHostToDevice, st), "theta H2D")))
        return fail(rc);
    const KernelArgs args = kernel_args(ctx);
    if ((rc = launch_step_assignment(args, init, ctx->d_mean.p, iteration, ctx->d_occ.p, 12 * (size_t)T + 16, st))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(counts_out, ctx->d_occ.p, T * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "counts D2H")))
        return fail(rc);
    if ((rc = cuda_rc(cudaStreamSynchronize(st), "step sync"))) return fail(rc);
    bayes_gibbs_destroy(ctx);
    return BAYES_OK;
}

int bayes_step_generate_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key,
                                   uint32_t iteration, int32_t *counts_out) {
    return step_assignment(device, locus, theta, key, iteration, counts_out, false);
}

int bayes_step_init_assignment(int device, const bayes_locus_desc *locus, uint64_t key, int32_t *counts_out) {
    return step_assignment(device, locus, nullptr, key, 0, counts_out, true);
}

int bayes_step_generate_expression(int device, int32_t T, const int32_t *counts, int32_t b, double gamma, uint64_t key,
                                   uint32_t iteration, double *theta_out) {
    if (T < 1 || T > kMaxT || !counts || !theta_out || !(gamma > 0)) {
        set_error("bayes_step_generate_expression: bad arguments");
        return BAYES_E_INVALID;
    }
    int n_plus = 0;
    for (int t = 0; t < T; t++) n_plus += counts[t] > 0;
    if (n_plus < 1 || b < n_plus || b > T) {
        set_error("bayes_step_generate_expression: need 1 <= |S+| <= b <= T");
        return BAYES_E_INVALID;
    }
    bayes_gibbs_ctx *ctx = nullptr;
    int rc = bayes_gibbs_create(device, &ctx);
    if (rc) return rc;
    auto fail = [&](int code) { bayes_gibbs_destroy(ctx); return code; };
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->d_occ.reserve(T)) || (rc = ctx->d_mean.reserve(T))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(ctx->d_occ.p, counts, T * sizeof(int32_t), cudaMemcpyHostToDevice, st), "counts H2D"))) return fail(rc);
    if ((rc = launch_step_expression(T, ctx->d_occ.p, b, gamma, key, iteration, ctx->d_mean.p, st))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(theta_out, ctx->d_mean.p, T * sizeof(double), cudaMemcpyDeviceToHost, st), "theta D2H"))) return fail(rc);
    if ((rc = cuda_rc(cudaStreamSynchronize(st), "step sync"))) return fail(rc);
    bayes_gibbs_destroy(ctx);
    return BAYES_OK;
}

int bayes_step_variates(int device, int kind, double a, double p, uint64_t key, int32_t n, double *out) {
    if (n < 1 || !out || kind < 0 || kind > 2) {
        set_error("bayes_step_variates: bad arguments");
        return BAYES_E_INVALID;
    }
    bayes_gibbs_ctx *ctx = nullptr;
    int rc = bayes_gibbs_create(device, &ctx);
    if (rc) return rc;
    auto fail = [&](int code) { bayes_gibbs_destroy(ctx); return code; };
    if ((rc = ctx->d_mean.reserve(n))) return fail(rc);
    if ((rc = launch_step_variates(kind, a, p, key, n, ctx->d_mean.p, ctx->stream))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(out, ctx->d_mean.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream), "variates D2H"))) return fail(rc);
    if ((rc = cuda_rc(cudaStreamSynchronize(ctx->stream), "step sync"))) return fail(rc);
    bayes_gibbs_destroy(ctx);
    return BAYES_OK;
}

}  // extern "C"
