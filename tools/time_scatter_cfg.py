"""GPU box: count / scatter / add_batch times of one bench config (with the SMESH_* tuning variables of the environment).
usage: python tools/time_scatter_cfg.py cfg2 [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import argparse, torch
import semantic_meshes
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cfg = dict(bench.CONFIGS[name])
if len(sys.argv) > 2:
    cfg["B"] = int(sys.argv[2])
sc = bench.Scene(cfg, 0, 1, torch.device("cuda", 0))
agg = semantic_meshes.fusion.MeshAggregator(sc.P, sc.C)
st = bench.stage_timings(argparse.Namespace(steps=10, quick=True), sc, agg, True, with_get=False)
tag = " ".join(f"{k[6:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SMESH_"))
print(f"{name} [{tag}]: count {st['count_kernel_ms']*1e3:.1f} scatter {st['scatter_kernel_ms']*1e3:.1f} add_batch {st['add_ms_per_view']*1e3:.1f} "
      f"add_serial {st['add_ms_per_view_serial']*1e3:.1f} render {st['render_ms_per_view']*1e3:.1f} us; add frac {sc.bytes_add/(st['add_ms_per_view']*1e-3)/1e9/bench.measured_peak_gbs()[0]:.3f}", flush=True)
