"""CPU, 2 processes over gloo: the view-shard + one-all-reduce scheme of semantic_meshes.distributed gives the same
accumulator as one process adding every view (the per-rank fusion is done by the CPU oracle here; the N>1 GPU path runs
the same host logic with NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

N_VIEWS, W, H, C, P = 7, 24, 20, 5, 60


def make_views():
    rng = np.random.default_rng(99)
    views = []
    for v in range(N_VIEWS):
        base = rng.integers(0, P, size=(W // 2, H // 2))
        ids = np.repeat(np.repeat(base, 2, 0), 2, 1).astype(np.uint32)
        ids[rng.random((W, H)) < 0.15] = 0xFFFFFFFF
        probs = rng.dirichlet(np.ones(C), size=(W, H)).astype(np.float32)
        probs[rng.random((W, H)) < 0.05] = 0
        probs[rng.random((W, H, C)) < 0.03] = 0
        views.append((ids, probs))
    return views


def worker(rank, world, port, kind, out_dir):
    for p in (ROOT, os.path.join(ROOT, "semantic-meshes_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    from semantic_meshes.distributed import allreduce_accumulator, reduce_scatter_rows, shard_views
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    views = make_views()
    agg = oracle.Aggregator(P, C, kind, 0.5)
    mine = shard_views(N_VIEWS, rank, world)
    for v in mine:
        agg.add(*views[v])
    acc = torch.from_numpy(agg.acc)
    (first, last), part = reduce_scatter_rows(acc)   # rows of this rank's slice, summed over the ranks
    allreduce_accumulator(acc)
    assert last - first == part.shape[0] and torch.equal(part, acc[first:last])
    assert (first, last) == ((0, P // 2) if rank == 0 else (P // 2, P))
    np.save(os.path.join(out_dir, f"acc_{kind}_{rank}.npy"), acc.numpy())
    np.save(os.path.join(out_dir, f"views_{kind}_{rank}.npy"), np.array(mine))
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind", ["sum", "mul"])
def test_two_rank_view_shard_matches_single_process(tmp_path, kind):
    import oracle
    world = 2
    mp.spawn(worker, args=(world, free_port(), kind, str(tmp_path)), nprocs=world, join=True)
    single = oracle.Aggregator(P, C, kind, 0.5)
    for ids, probs in make_views():
        single.add(ids, probs)
    accs = [np.load(tmp_path / f"acc_{kind}_{r}.npy") for r in range(world)]
    seen = sorted(int(v) for r in range(world) for v in np.load(tmp_path / f"views_{kind}_{r}.npy"))
    assert seen == list(range(N_VIEWS))                      # every view exactly once
    assert np.array_equal(accs[0], accs[1])                  # every rank ends with the same accumulator
    assert np.array_equal(np.isinf(accs[0]), np.isinf(single.acc))
    fin = np.isfinite(single.acc)
    np.testing.assert_allclose(accs[0][fin], single.acc[fin], rtol=1e-5, atol=1e-6)
    # and therefore the same per-face distribution
    out = np.empty_like(single.acc)
    oracle.lib().oracle_fuse_get(single.kind, np.ascontiguousarray(accs[0]).ctypes.data, P, C, out.ctypes.data)
    np.testing.assert_allclose(out, single.get(), rtol=5e-4 if kind == "mul" else 1e-5, atol=1e-6)


def test_shard_views_partition():
    from semantic_meshes.distributed import shard_views
    for n in (0, 1, 7, 500):
        for world in (1, 2, 4, 8):
            parts = [shard_views(n, r, world) for r in range(world)]
            assert sorted(v for p in parts for v in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)
