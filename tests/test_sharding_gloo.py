"""N > 1 host path on CPU: two gloo ranks agree on a cost-balanced, disjoint, complete sharding of the loci
(bayesembler_b200/shard.py), exactly the exchange bench.py does over NCCL before each rank runs its own shard."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_rank, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from bayesembler_b200 import shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    block = synth.SynthBatch(2, n_loci=per_rank, first=rank * per_rank, n_threads=2)
    cost_local = shard.desc_costs(block.desc_array())
    cost_all = shard.gather_costs(cost_local, dist)
    mine, load = shard.shard_loci(cost_all, world, rank)
    batch = synth.SynthBatch(2, ids=mine, n_threads=2)
    cost_mine = shard.desc_costs(batch.desc_array())  # the regenerated loci are the ones that were costed
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mine=mine, load=load, cost_all=cost_all, cost_mine=cost_mine,
             n_frag=batch.n_frag)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding(lib, tmp_path):
    import torch.multiprocessing as mp
    from bayesembler_b200 import synth
    synth.build()
    world, per_rank = 2, 150
    mp.spawn(_worker, args=(world, _free_port(), per_rank, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(str(tmp_path / ("rank%d.npz" % k))) for k in range(world)]
    assert np.array_equal(r[0]["cost_all"], r[1]["cost_all"]) and np.array_equal(r[0]["load"], r[1]["load"])
    both = np.concatenate([r[0]["mine"], r[1]["mine"]])
    assert np.array_equal(np.sort(both), np.arange(world * per_rank))  # disjoint and complete
    for k in range(world):
        assert np.allclose(r[k]["cost_mine"], r[k]["cost_all"][r[k]["mine"]])
        assert np.isclose(r[k]["cost_mine"].sum(), r[k]["load"][k])
    load = r[0]["load"]
    assert load.max() / load.mean() < 1.05  # cost-balanced
    # same partition as a single process would compute
    whole = synth.SynthBatch(2, n_loci=world * per_rank, n_threads=2)
    from bayesembler_b200 import shard
    assert np.allclose(shard.desc_costs(whole.desc_array()), r[0]["cost_all"])
