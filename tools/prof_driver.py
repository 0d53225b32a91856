"""Small driver for ncu: a few eager views of one bench config (render + add), then get() - nothing else on the GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import torch
import semantic_meshes
from semantic_meshes import synthetic
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
views = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = bench.CONFIGS[name]
W, H, C = cfg["W"], cfg["H"], cfg["C"]
mesh, cams = bench.build_scene(cfg, 0, views)
renderer = semantic_meshes.render.triangles(mesh)
agg = semantic_meshes.fusion.MeshAggregator(renderer.getPrimitivesNum(), C)
probs = [synthetic.predictions_torch(W, H, C, seed=b, device="cuda") for b in range(views)]
torch.cuda.synchronize()
for rep in range(2):
    for b in range(views):
        idx, _ = renderer.render(cams[b])
        agg.add(idx, probs[b])
out = agg.get(device=True)
torch.cuda.synchronize()
print("done", float(out.sum()))
