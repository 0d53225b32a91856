"""GPU box: add_batch time per view of the `mul` (and `sum`) aggregator of bench configs.
usage: python tools/time_mul.py cfg3 cfg2 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import argparse, torch
import semantic_meshes
import bench
for name in sys.argv[1:] or ["cfg3"]:
    sc = bench.Scene(dict(bench.CONFIGS[name]), 0, 1, torch.device("cuda", 0))
    agg = semantic_meshes.fusion.MeshAggregator(sc.P, sc.C)
    st = bench.stage_timings(argparse.Namespace(steps=10, quick=True), sc, agg, True, kinds=("sum", "mul"), with_get=False)
    tag = " ".join(f"{k[6:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SMESH_"))
    print(f"{name} [{tag}]: add_batch sum {st['add_ms_per_view']*1e3:.1f} mul {st['add_ms_per_view_mul']*1e3:.1f} us/view", flush=True)
