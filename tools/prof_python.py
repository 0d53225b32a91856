"""GPU box: where the HOST time of the eager API path goes (cProfile of ViewPipeline.run over a few hundred views)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import torch
import semantic_meshes
from semantic_meshes.pipeline import ViewPipeline
import bench
cfg = dict(bench.CONFIGS["cfg3"]); cfg["B"] = 8
sc = bench.Scene(cfg, 0, 1, torch.device("cuda", 0))
agg = semantic_meshes.fusion.MeshAggregator(sc.P, sc.C)
pipe = ViewPipeline(sc.renderer, agg)
cams = sc.cams * 32
preds = [sc.probs[b % 8] for b in range(len(cams))]
pipe.run(cams, preds); torch.cuda.synchronize()
t0 = time.perf_counter(); pipe.run(cams, preds); t_issue = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"{len(cams)} views: host issue {t_issue*1e6/len(cams):.1f} us/view, wall {t_all*1e6/len(cams):.1f} us/view")
pr = cProfile.Profile(); pr.enable(); pipe.run(cams, preds); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
