#!/usr/bin/env python
"""bench.py - the headline measurement of the two hot paths (render.triangles(...).render + MeshAggregator.add).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]

One STEP = one pass over a batch of B synthetic views: for every view, render the mesh into the camera (primitive index
+ depth image) and fuse that view's (W, H, C) prediction into the per-face accumulator. `value` = views/s over all ranks
with every input already resident in HBM (the step is replayed as a CUDA graph); `e2e` = the same loop through the
public Python API with the predictions in pinned HOST memory (H2D copy of every view and a D2H read of the last view's
render result inside the timed region). `roofline` is the scatter kernel of MeshAggregator.add (the dominant kernel
of the HBM-bound path; timed as the B views' launches back to back in a CUDA graph) against the measured HBM peak; `cpu_baseline` / `--impl reference` time the GENUINE reference (oracle/_ref, compiled from
the reference's sources) on the host cores of the same box.

Multi-GPU (torchrun, one rank per GPU): views shard across ranks (weak scaling: every rank runs the same number of
views), each rank owns a private accumulator, ONE NCCL all-reduce of it closes the timed region.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "semantic-meshes_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# BASELINE.json `configs`, shapes of SURVEY.md section 8. tris_per_view sets the camera height (how much of the mesh one
# view sees); views = size of the whole job in the reference configuration (the bench cycles through B distinct views).
CONFIGS = {
    "cfg1": dict(name="icosphere 1280 tris, 256x256, 19 classes", mesh="icosphere", F=1280, W=256, H=256, C=19, views=4,
                 B=4),
    "cfg2": dict(name="ScanNet-scale 500k tris, 640x480, 40 classes", mesh="terrain", F=500_000, W=640, H=480, C=40,
                 views=200, tris_per_view=30_000, B=16),
    "cfg3": dict(name="Cityscapes-scale 2M tris, 2048x1024 (1024x2048 images), 19 classes", mesh="terrain", F=2_000_000,
                 W=2048, H=1024, C=19, views=500, tris_per_view=150_000, B=16),
    "cfg4": dict(name="wide-C 1M tris, 1920x1080, 150 classes", mesh="terrain", F=1_000_000, W=1920, H=1080, C=150,
                 views=100, tris_per_view=120_000, B=4),
    "cfg5": dict(name="dense sweep 5M tris, 1280x720, 19 classes", mesh="terrain", F=5_000_000, W=1280, H=720, C=19,
                 views=2000, tris_per_view=80_000, B=16),
}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report it rather than fail the bench
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)

    def summary(self, t0, t1):
        inside = [m for (t, m) in self.samples if t0 <= t <= t1] or [m for (_, m) in self.samples]
        out = {"sm_mhz": float(np.median(inside)) if inside else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(inside)}
        if self.error:
            out["error"] = self.error
        return out


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


def build_scene(cfg, rank, n_distinct):
    from semantic_meshes import synthetic
    if cfg["mesh"] == "icosphere":
        mesh = synthetic.mesh("icosphere")
        cams = synthetic.orbit_cameras(n_distinct, cfg["W"], cfg["H"], (0, 0, 0), 3.0, seed=100 + rank, tilt_deg=(0, 180))
    else:
        mesh = synthetic.mesh("terrain", cfg["F"], seed=1234)
        cams = synthetic.terrain_cameras(n_distinct, cfg["W"], cfg["H"], cfg["F"], cfg["tris_per_view"], seed=100 + rank)
    return mesh, cams


def run_ours(args, cfg):
    import torch
    import semantic_meshes
    from semantic_meshes import _lib, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    W, H, C, B = cfg["W"], cfg["H"], cfg["C"], cfg["B"]
    npix = W * H
    mesh, cams = build_scene(cfg, rank, B)
    renderer = semantic_meshes.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    agg = semantic_meshes.fusion.MeshAggregator(primitives=P, classes=C)

    # B distinct views resident in HBM; one view (>= 49 MB, 160 MB at cfg3) is re-read only after B-1 others, so with
    # B * view bytes >> 126 MB of L2 nothing is served from cache between timed iterations
    probs = torch.empty((B, W, H, C), dtype=torch.float32, device=dev)
    for b in range(B):
        synthetic.predictions_torch(W, H, C, seed=1000 * rank + b, device=dev, out=probs[b])
    ids_all = torch.empty((B, W, H), dtype=torch.int32, device=dev)

    from semantic_meshes.pipeline import ViewPipeline
    pipe = ViewPipeline(renderer, agg)

    def step():
        agg.restart_epochs()  # the step is replayed as a graph: every replay must see the same count epochs (smesh.h)
        if args.no_overlap:
            for b in range(B):
                idx, _ = renderer.render(cams[b])
                agg.add(idx, probs[b])
        else:
            pipe.run(cams, probs)  # same work, render of view b+1 overlapped with the fusion of view b (two streams)

    # per-view statistics (outside any timed region): accepted pixels and touched faces
    accepted, touched, covered = [], [], []
    for b in range(B):
        idx, _ = renderer.render(cams[b])
        ids_all[b] = idx
        valid = idx >= 0
        ok = valid & (probs[b].sum(-1) > 0.5)
        accepted.append(int(ok.sum().item()))
        touched.append(int(torch.unique(idx[ok]).numel()))
        covered.append(float(valid.float().mean().item()))
    torch.cuda.synchronize()

    # ---- device-resident throughput: the step as a CUDA graph ----
    use_graph = not args.no_graph
    step()  # warm the library's per-kernel configuration and the allocator before capture
    torch.cuda.synchronize()
    graph = None
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"CUDA graph capture failed ({e!r}); timing eager launches\n")
            graph = None
            torch.cuda.synchronize()
    run_step = graph.replay if graph is not None else step

    agg.reset()
    for _ in range(args.warmup):
        run_step()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    start, stop, ar_start = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    profile_range = os.environ.get("SMESH_PROFILE_RANGE") == "1"  # ncu --profile-from-start off: timed region only
    if profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    start.record()
    for _ in range(args.steps):
        run_step()
    ar_start.record()
    agg.allreduce()
    stop.record()
    torch.cuda.synchronize()
    if profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    allreduce_ms = ar_start.elapsed_time(stop)
    if dist is not None:
        t = torch.tensor([elapsed_ms, allreduce_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, allreduce_ms = float(t[0]), float(t[1])
    views_total = args.steps * B * world
    value = views_total / (elapsed_ms * 1e-3)
    clocks = sampler.summary(t0, t1)

    # ---- stage timing on the same inputs: render alone, add alone, and the scatter kernel alone (roofline) ----
    lib = _lib.lib
    stream = torch.cuda.current_stream().cuda_stream
    kind = _lib.KIND["sum"]
    reps = max(1, min(args.steps, 20))

    def timed(fn, n):
        """Mean device time of fn(): captured once into a CUDA graph and replayed n times (eager launches from Python
        would measure the host, not the kernels); falls back to eager launches if capture fails."""
        fn()
        torch.cuda.synchronize()
        run = fn
        if use_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                run = g.replay
            except Exception:
                torch.cuda.synchronize()
        run()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def add_all():
        agg.restart_epochs()
        agg.add_batch(ids_all, probs)

    render_ms = timed(lambda: [renderer.render(cams[b]) for b in range(B)], reps) / B
    add_ms = timed(add_all, reps) / B
    # the scatter kernel alone: every view's per-face counts are prepared in an array of its own (untimed), then the B
    # scatter launches run back to back - as a CUDA graph, so that the interval between the two events holds kernels and
    # nothing of the host's launch path - and the interval is divided by the number of launches
    counts_b = torch.zeros((B, max(P, 1)), dtype=torch.int32, device=dev)
    for b in range(B):
        _lib.check(lib.smesh_fuse_count(ids_all[b].data_ptr(), _lib.ID_I32, H, 1, W, H, P, counts_b[b].data_ptr(), 1, None,
                                        stream))

    def scatter_all():
        s = torch.cuda.current_stream().cuda_stream
        for b in range(B):
            _lib.check(lib.smesh_fuse_scatter(kind, ids_all[b].data_ptr(), probs[b].data_ptr(), None, npix, C, P,
                                              agg.images_equal_weight, counts_b[b].data_ptr(), 1, agg._acc.data_ptr(), s))

    scatter_ms = timed(scatter_all, reps) / B
    # for comparison: every launch bracketed by its own pair of events on an otherwise idle stream (this interval also
    # holds the launch latency of one kernel, ~2-4 us)
    pairs = []
    torch.cuda.synchronize()
    for b in range(B):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        _lib.check(lib.smesh_fuse_scatter(kind, ids_all[b].data_ptr(), probs[b].data_ptr(), None, npix, C, P,
                                          agg.images_equal_weight, counts_b[b].data_ptr(), 1, agg._acc.data_ptr(), stream))
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    scatter_single_ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
    del counts_b
    peak, peak_src = measured_peak_gbs()
    bytes_inputs = 4.0 * npix * C + 4.0 * npix
    bytes_alg = bytes_inputs + 8.0 * C * float(np.mean(touched))  # probs + ids once, touched rows r+w
    achieved = bytes_alg / (scatter_ms * 1e-3) / 1e9

    # ---- end to end through the public API: predictions in pinned host memory ----
    n_host = min(B, 4)
    host_probs = [torch.empty((W, H, C), dtype=torch.float32, pin_memory=True) for _ in range(n_host)]
    for b in range(n_host):
        host_probs[b].copy_(probs[b])
    host_idx = torch.empty((W, H), dtype=torch.int32, pin_memory=True)
    host_depth = torch.empty((W, H), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(1, min(args.steps, 4))

    def e2e_step():
        for b in range(B):
            idx, depth = renderer.render(cams[b])
            agg.add(idx, host_probs[b % n_host])  # H2D of the (W, H, C) prediction happens inside add()
        host_idx.copy_(idx, non_blocking=True)
        host_depth.copy_(depth, non_blocking=True)

    e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    agg.allreduce()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = e2e_steps * B * world / (e2e_ms * 1e-3)

    # kernel launches of ours inside the timed region: per view 4 (render: begin/cull, clusters, big, resolve) + 2 (add:
    # count, scatter)
    gpu_launches = args.steps * B * 6

    line = {
        "metric": "views/s (render + MeshAggregator.add per view), whole job",
        "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['name']}", "views_per_step_per_gpu": B, "triangles": P, "W": W, "H": H,
                   "classes": C, "aggregator": "sum", "images_equal_weight": 0.5, "parallelism": f"view-shard x{world}",
                   "pixels_covered": float(np.mean(covered)), "faces_touched_per_view": float(np.mean(touched)),
                   "cache": f"{B} distinct views of {bytes_inputs / 1e6:.0f} MB cycled per step (>> 126 MB L2)",
                   "cuda_graph": graph is not None, "allreduce_in_timed_region": world > 1,
                   "render_add_overlap": not args.no_overlap},
        "mpixel_face_scatters_per_s": value * float(np.mean(accepted)) / 1e6,
        "stages": {"render_ms_per_view": render_ms, "add_ms_per_view": add_ms, "scatter_kernel_ms": scatter_ms,
                   "scatter_kernel_ms_single_launch_events": scatter_single_ms,
                   "render_views_per_s": 1e3 / render_ms, "add_views_per_s": 1e3 / add_ms, "allreduce_ms": allreduce_ms},
        "roofline": {"bound": "hbm", "kernel": "smesh::fuse::scatter_pair_kernel" if C == 19 else "smesh::fuse::scatter kernel of this C", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "timing": f"{B} launches back to back in a CUDA graph, {reps} replays, CUDA events / launches",
                     "algorithmic_bytes_per_launch": bytes_alg, "input_only_frac": bytes_inputs / (scatter_ms * 1e-3) / 1e9 / peak},
        "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": int(B * npix * C * 4),
                "d2h_bytes_per_step": int(npix * 8), "steps": e2e_steps},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
    }
    traffic_file = os.path.join(ROOT, "profiles", "scatter_traffic.json")
    if os.path.exists(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file)).get(args.config)
        except Exception:
            pass

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, P, ids_all, probs, n_views=args.cpu_views)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def cpu_baseline(cfg, P, ids_all, probs, n_views=3):
    """The GENUINE reference aggregator (oracle/_ref/libref_fusion.so = include/semantic_meshes/fusion/Mesh.h compiled
    from the reference's sources) on the host cores, same ids and predictions, a bounded sample of views."""
    import oracle
    W, H, C = cfg["W"], cfg["H"], cfg["C"]
    cores = os.cpu_count() or 1
    n_views = min(n_views, ids_all.shape[0])
    use_ref = os.path.exists(oracle.ref_fusion_path()) and oracle.ref_fusion_lib().ref_fusion_has_classes(C)
    agg = oracle.RefAggregator(P, C) if use_ref else oracle.Aggregator(P, C)
    host = [(ids_all[b].cpu().numpy().view(np.uint32), probs[b].cpu().numpy()) for b in range(n_views)]
    t0 = time.perf_counter()
    for ids, pr in host:
        agg.add(ids, pr)
    dt = time.perf_counter() - t0
    return {"value": n_views / dt, "unit": "views/s (MeshAggregator.add only, inputs already in host arrays)",
            "cores": cores if use_ref else 1, "kind": "reference" if use_ref else "port",
            "sample": f"{n_views} views of this workload, add() only ({dt:.2f} s)"}


def run_reference(args, cfg):
    """--impl reference: the reference's own implementation of the path on this box: its CUDA rasterizer (genuine kernel,
    oracle/_ref/libref_raster.so; falls back to the CPU restatement if it cannot run) + its CPU/OpenMP aggregator with
    all host threads, one view per step, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle
    from oracle import write_plain_ply
    W, H, C = cfg["W"], cfg["H"], cfg["C"]
    n_distinct = min(cfg["B"], 4)
    mesh, cams = build_scene(cfg, 0, n_distinct)
    P = mesh.faces.shape[0]
    from semantic_meshes import synthetic
    preds = [synthetic.predictions_numpy(W, H, C, seed=b) for b in range(n_distinct)]
    cores = os.cpu_count() or 1
    use_ref_fusion = os.path.exists(oracle.ref_fusion_path()) and oracle.ref_fusion_lib().ref_fusion_has_classes(C)
    agg = oracle.RefAggregator(P, C) if use_ref_fusion else oracle.Aggregator(P, C)
    ref_renderer, raster_kind = None, "port (CPU restatement)"
    tmp = tempfile.TemporaryDirectory()
    if os.path.exists(oracle.ref_raster_path()):
        try:
            ply = os.path.join(tmp.name, "mesh.ply")
            write_plain_ply(ply, mesh.vertices, mesh.faces)
            ref_renderer = oracle.RefRenderer(ply)
            raster_kind = "reference (genuine CUDA kernel)"
        except Exception as e:
            sys.stderr.write(f"reference rasterizer unavailable ({e!r}); using the CPU restatement\n")

    def render(cam):
        if ref_renderer is not None:
            return ref_renderer.render(cam.rotation, cam.translation, cam.focal_lengths, cam.principal_point, W, H)[0]
        return oracle.raster_render(mesh.vertices, mesh.faces, cam.rotation, cam.translation, cam.focal_lengths,
                                    cam.principal_point, W, H)[0]

    def step(i):
        ids = render(cams[i % n_distinct])
        agg.add(ids, preds[i % n_distinct])

    budget = float(os.environ.get("SMESH_REF_BUDGET_S", "240"))
    t_begin = time.perf_counter()
    for i in range(args.warmup):
        step(i)
        if time.perf_counter() - t_begin > budget / 4:
            break
    done = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
        done += 1
        if time.perf_counter() - t_begin > budget:
            break
    dt = time.perf_counter() - t0
    value = done / dt
    kind = "reference" if use_ref_fusion else "port"
    line = {
        "impl": "reference", "metric": "views/s (render + MeshAggregator.add per view), whole job", "value": value,
        "unit": "views/s", "n_gpus": world, "steps": done, "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['name']}", "views_per_step_per_gpu": 1, "triangles": P, "W": W, "H": H,
                   "classes": C, "aggregator": "sum", "images_equal_weight": 0.5, "rasterizer": raster_kind,
                   "note": "rank 0 only; one view per step; stops early after SMESH_REF_BUDGET_S seconds"},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores if use_ref_fusion else 1, "kind": kind,
                         "sample": f"{done} views of this workload, render + add"},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    tmp.cleanup()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="render and add strictly one after the other on one stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-views", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
