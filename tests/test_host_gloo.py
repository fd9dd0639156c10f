"""world_size-2 gloo test of the multi-GPU host logic on CPU: contiguous
column shards, no data-path collective, one all-reduce of the eight domain diagnostics.  The
oracle stands in for the per-rank column step (it is the checker here, not the product)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NCOL, NZ, DT = 512, 60, 10.0


def _worker(rank, world, initfile, outdir):
    from kid_b200 import synth
    from kid_b200.shard import shard_range, diag_from_state, allreduce_diag
    from oracle.oracle import Oracle, FIELDS
    dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=world)
    c0, c1 = shard_range(NCOL, rank, world)
    st, p, dz = synth.make_domain(c1 - c0, nz=NZ, col0=c0, nx=1024, cloudy_fraction=1.0, coherent=False)
    state = {k: v.numpy().copy() for k, v in st.items()}
    o = Oracle(nthreads=2)
    ppt = o.step(DT, state, p.numpy(), dz.numpy())
    local = diag_from_state(state, p.numpy(), dz.numpy(), ppt)
    total = allreduce_diag(local)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), ppt=ppt, diag=total, local=local, c0=c0, c1=c1,
             **{k: state[k] for k in FIELDS})
    o.close()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from kid_b200 import synth
    from kid_b200.shard import shard_range, diag_from_state
    from oracle.oracle import Oracle, FIELDS
    world = 2
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, initfile, d), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
        st, p, dz = synth.make_domain(NCOL, nz=NZ, nx=1024, cloudy_fraction=1.0, coherent=False)
        state = {k: v.numpy().copy() for k, v in st.items()}
        o = Oracle(nthreads=2)
        ppt = o.step(DT, state, p.numpy(), dz.numpy())
        o.close()
        # shards tile the domain
        assert int(parts[0]["c0"]) == 0 and int(parts[0]["c1"]) == int(parts[1]["c0"]) and int(parts[1]["c1"]) == NCOL
        # per-column results are bitwise independent of the sharding
        for k in FIELDS:
            assert np.array_equal(np.concatenate([q[k] for q in parts], axis=1), state[k]), k
        assert np.array_equal(np.concatenate([q["ppt"] for q in parts], axis=1), ppt)
        # every rank holds the same reduced diagnostics = the single-process sums (f64 order only)
        whole = diag_from_state(state, p.numpy(), dz.numpy(), ppt)
        assert np.array_equal(parts[0]["diag"], parts[1]["diag"])
        np.testing.assert_allclose(parts[0]["diag"], whole, rtol=1e-12)
        assert whole[7] == NCOL


def test_shard_range_edges():
    from kid_b200.shard import shard_range
    for n, w in ((10, 3), (7, 8), (1048576, 8), (5, 1)):
        r = [shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_numa_helpers_parse_and_report():
    """kid_b200/shard.py: the cpulist parser, and the binding helper reports instead of raising on a box without GPUs."""
    from kid_b200 import shard
    assert shard._parse_list("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert shard._parse_list("") == []
    assert shard.gpu_numa_node("ffff:ff:1f.0") == -1
    before = os.sched_getaffinity(0)
    info = shard.bind_to_gpu_numa(0)
    assert info["gpu"] == 0 and "node" in info and "mem_policy" in info
    if info["node"] < 0:
        assert os.sched_getaffinity(0) == before
