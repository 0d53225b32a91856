"""GPU box: cluster statistics of a bench config: candidate clusters per view, big-queue entries."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import numpy as np, torch
import semantic_meshes, bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
views = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = bench.CONFIGS[name]
mesh, cams = bench.build_scene(cfg, 0, views)
r = semantic_meshes.render.triangles(mesh)
V, F = r._V, r._F
nc = (F + 127) // 128
al = lambda v: (v + 255) // 256 * 256
off_faces = al(max(V, 1) * 16)
off_clusters = off_faces + al(nc * 128 * 16)
cl = r._mesh[off_clusters:off_clusters + nc * 16].view(torch.float32).view(-1, 4).cpu().numpy()
rad = np.abs(cl[:, 3]); notwell = np.signbit(cl[:, 3])
print(f"{name}: F={F} clusters={nc} radius median {np.median(rad):.2f} p99 {np.percentile(rad, 99):.2f}; not-all-well-shaped {notwell.mean():.3f}")
for v, cam in enumerate(cams):
    idx, _ = r.render(cam)
    torch.cuda.synchronize()
    cnt = r._workspace[:32].view(torch.int32).cpu().numpy()
    vis = torch.unique(idx[idx >= 0]).numel()
    print(f"view {v}: candidate clusters {cnt[0]} ({cnt[0]*128} faces; visible faces {vis}); big queue entries {cnt[3]} chunks {cnt[2]}")
