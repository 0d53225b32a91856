"""GPU box: time the scatter kernel alone on cfg3's index images for a list of class counts (run once as is and once with
SMESH_NO_PAIR=1 to compare the two-pixels-per-lane kernel with the per-pixel ring kernel).
usage: python tools/time_scatter_classes.py 4,8,13,16,20 [views] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import numpy as np, torch
import semantic_meshes
from semantic_meshes import synthetic, _lib
import bench
classes = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "4,8,13,16,20").split(",")]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = bench.CONFIGS["cfg3"]
W, H = cfg["W"], cfg["H"]
mesh, cams = bench.build_scene(cfg, 0, B)
renderer = semantic_meshes.render.triangles(mesh)
P = renderer.getPrimitivesNum()
ids = [renderer.render(cams[b])[0] for b in range(B)]
stream = torch.cuda.current_stream().cuda_stream
tag = "ring" if os.environ.get("SMESH_NO_PAIR") else "pair"
for C in classes:
    agg = semantic_meshes.fusion.MeshAggregator(P, C)
    probs = [synthetic.predictions_torch(W, H, C, seed=b, device="cuda") for b in range(B)]
    agg.restart_epochs()
    counts = torch.zeros((B, P), dtype=torch.int32, device="cuda")
    for b in range(B):
        _lib.check(_lib.lib.smesh_fuse_count(ids[b].data_ptr(), _lib.ID_I32, H, 1, W, H, P, counts[b].data_ptr(), b + 1, None, stream))

    def scatter_all():
        s = torch.cuda.current_stream().cuda_stream
        for b in range(B):
            _lib.check(_lib.lib.smesh_fuse_scatter(0, ids[b].data_ptr(), probs[b].data_ptr(), None, W * H, C, P, 0.5, counts[b].data_ptr(), b + 1, agg._acc.data_ptr(), s))

    t = bench.timed_graph(torch, scatter_all, reps) / B * 1e3   # launches back to back in a CUDA graph
    wide = os.environ.get("SMESH_PAIR_WIDE", "default")
    print(f"C={C:3d} {tag} wide={wide}: scatter {t:.1f} us per launch; input {(4 * W * H * (C + 1)) / t / 1e3:.0f} GB/s", flush=True)
    del counts
    del agg, probs
