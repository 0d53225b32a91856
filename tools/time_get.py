"""GPU box: MeshAggregator.get(device=True) time and roofline fraction for (P, C) pairs, sum and mul.
usage: python tools/time_get.py [P:C ...]     (default 2000000:19 5000000:19 500000:40 1000000:150)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import torch
import semantic_meshes
import bench
dev = torch.device("cuda", 0)
peak = bench.measured_peak_gbs()[0]
tag = " ".join(f"{k[6:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SMESH_"))
for spec in sys.argv[1:] or ["2000000:19", "5000000:19", "500000:40", "1000000:150"]:
    P, C = (int(x) for x in spec.split(":"))
    g = torch.Generator(device=dev).manual_seed(1)
    W, H = 1024, 512
    for kind in ("sum", "mul"):
        agg = semantic_meshes.fusion.MeshAggregator(P, C, aggregator=kind)
        for _ in range(3):
            ids = torch.randint(0, P, (W, H), device=dev, dtype=torch.int32, generator=g)
            probs = torch.softmax(3 * torch.randn((W, H, C), device=dev, generator=g), -1)
            agg.add(ids, probs)
        ms = bench.timed_graph(torch, lambda: agg.get(device=True), 10, True)
        frac = 4.0 * P * (agg._cpad + C) / (ms * 1e-3) / 1e9 / peak
        print(f"get P={P} C={C} {kind} [{tag}]: {ms*1e3:.1f} us = {frac:.3f} of the roofline", flush=True)
