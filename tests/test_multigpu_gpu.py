"""GPU, >= 2 devices (skipped on a single-GPU box): the view-sharded N-way result on real hardware.

Every rank renders and fuses its round-robin share of the views with the CUDA path, one NCCL all-reduce sums the
accumulators (MeshAggregator.allreduce), and the result must equal the accumulator of ALL views added on one GPU within 1e-5
 - for every aggregator kind; `reduce_scatter` + `get(rows=...)` must give the same rows of get()."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "semantic-meshes_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import semantic_meshes
    from semantic_meshes import synthetic
    from semantic_meshes.distributed import shard_views
    W, H, C, F, n_views = 320, 256, 19, 30000, 10
    mesh = synthetic.mesh("terrain", F, seed=3)
    cams = synthetic.terrain_cameras(n_views, W, H, F, tris_per_view=9000, seed=21)
    renderer = semantic_meshes.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    result = {}
    for kind in ("sum", "summax", "mul"):
        agg = semantic_meshes.fusion.MeshAggregator(P, C, kind)
        for v in shard_views(n_views, rank, world):
            idx, _ = renderer.render(cams[v])
            agg.add(idx, synthetic.predictions_torch(W, H, C, seed=500 + v, device="cuda"))
        shard = semantic_meshes.fusion.MeshAggregator(P, C, kind)
        shard.load_state(agg.state())
        agg.allreduce()
        rows, dist_rows = shard.reduce_scatter_get()
        if rank == 0:
            single = semantic_meshes.fusion.MeshAggregator(P, C, kind)
            for v in range(n_views):
                idx, _ = renderer.render(cams[v])
                single.add(idx, synthetic.predictions_torch(W, H, C, seed=500 + v, device="cuda"))
            result[kind] = (agg.state().cpu().numpy(), single.state().cpu().numpy(), agg.get(), single.get())
        full = agg.get(device=True)
        torch.testing.assert_close(dist_rows, full[rows[0]:rows[1]], rtol=1e-6, atol=1e-7)
    if rank == 0:
        np.savez(os.path.join(out_dir, "multi.npz"), **{f"{k}_{i}": a for k, v in result.items() for i, a in enumerate(v)})
    dist.barrier()
    dist.destroy_process_group()


def test_view_shard_allreduce_matches_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    world = 2 if torch.cuda.device_count() < 4 else 4
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    data = np.load(tmp_path / "multi.npz")
    for kind in ("sum", "summax", "mul"):
        multi, single, g_multi, g_single = (data[f"{kind}_{i}"] for i in range(4))
        assert np.array_equal(np.isinf(multi), np.isinf(single))
        fin = np.isfinite(single)
        scale = float(np.abs(single[fin]).max())
        assert scale > 0
        assert float(np.abs(multi[fin] - single[fin]).max()) <= 1e-5 * scale, kind
        np.testing.assert_allclose(g_multi, g_single, rtol=5e-4 if kind == "mul" else 1e-5, atol=5e-6 if kind == "mul" else 1e-7)
