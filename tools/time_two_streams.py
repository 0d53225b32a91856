"""GPU box experiment: do two half-grid scatter pipelines on two streams beat one full-grid pipeline for small views?
Two aggregators (so that no counter array is shared), views dealt alternately, one add() per view on each stream, all
captured into one CUDA graph (fork / join); SMESH_PAIR_CTAS / SMESH_SCATTER_CTAS of the environment cap the CTAs per SM.
usage: [SMESH_PAIR_CTAS=2] python tools/time_two_streams.py cfg5 [one|two]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200")]
import torch
import semantic_meshes
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
mode = sys.argv[2] if len(sys.argv) > 2 else "two"
dev = torch.device("cuda", 0)
sc = bench.Scene(dict(bench.CONFIGS[name]), 0, 1, dev)
aggs = [semantic_meshes.fusion.MeshAggregator(sc.P, sc.C) for _ in range(2)]
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]


def one():
    for b in range(sc.B):
        aggs[0].add(sc.ids[b], sc.probs[b])


def two():
    cur = torch.cuda.current_stream()
    for k in range(2):
        streams[k].wait_stream(cur)
        with torch.cuda.stream(streams[k]):
            for b in range(k, sc.B, 2):
                aggs[k].add(sc.ids[b], sc.probs[b])
    for k in range(2):
        cur.wait_stream(streams[k])


ms = bench.timed_graph(torch, one if mode == "one" else two, 10, True) / sc.B
tag = " ".join(f"{k[6:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SMESH_"))
print(f"{name} {mode} [{tag}]: add {ms*1e3:.1f} us per view = {sc.bytes_add/(ms*1e-3)/1e9/bench.measured_peak_gbs()[0]:.3f}", flush=True)
