"""GPU box: the count stage alone and the whole add (batch, overlapped / not) for the count-kernel variants
(SMESH_COUNT_VARIANT, read per call by the library). usage: python tools/time_count.py [cfg3] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import numpy as np, torch
import semantic_meshes
from semantic_meshes import _lib
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cfg = dict(bench.CONFIGS[name])
cfg["B"] = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
scene = bench.Scene(cfg, 0, 1, dev)
B, W, H, C, P = scene.B, scene.W, scene.H, scene.C, scene.P
agg = semantic_meshes.fusion.MeshAggregator(P, C)
counts = torch.zeros((B, P), dtype=torch.int32, device=dev)
ids = scene.ids
flat = ids.reshape(B, -1)
runs = int(((flat[:, 1:] != flat[:, :-1]) | (torch.arange(1, W * H, device=dev) % 32 == 0)).sum().item() / B)
print(f"{name}: {W}x{H}, runs inside 32-pixel groups per view: {runs}", flush=True)


def count_all():
    s = torch.cuda.current_stream().cuda_stream
    for b in range(B):
        _lib.check(_lib.lib.smesh_fuse_count(ids[b].data_ptr(), _lib.ID_I32, H, 1, W, H, P, counts[b].data_ptr(), 1 + b, None, s))


def add_all():
    agg.restart_epochs()
    agg.add_batch(scene.ids, scene.probs)


ref = None
for variant in (0, 1, 10):
    os.environ["SMESH_COUNT_VARIANT"] = str(variant)
    counts.zero_()
    count_all()
    torch.cuda.synchronize()
    c = (counts & 0xFFFFFF).clone()
    if variant == 0:
        ref = c
    elif variant < 10:
        assert torch.equal(c, ref), f"variant {variant} counts differ"
    if os.environ.get("NCU") == "1":  # under ncu: the eager launches above are all that is needed
        continue
    t_count = bench.timed_graph(torch, count_all, 10) / B * 1e3
    line = f"variant {variant:2d}: count {t_count:6.2f} us"
    if variant < 10:
        os.environ.pop("SMESH_NO_BATCH_OVERLAP", None)
        t_add = bench.timed_graph(torch, add_all, 10) / B * 1e3
        line += f"   add_batch (overlapped) {t_add:6.2f} us"
    print(line, flush=True)
