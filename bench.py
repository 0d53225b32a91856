#!/usr/bin/env python
"""bench.py - the headline measurement of the two hot paths (render.triangles(...).render + MeshAggregator.add).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference] [--scaling weak|strong]

One STEP = `cycles` passes over a batch of B synthetic views: for every view, render the mesh into the camera (primitive
index + depth image) and fuse that view's (W, H, C) prediction into the per-face accumulator (cfg3: 24 x 16 = 384 views,
about 30 ms, so that the driver's K = 20 steps give a timed region of > 0.5 s).
  value      views/s over all ranks with every input already resident in HBM (the step is replayed as a CUDA graph)
  e2e        the same loop through the public Python API with the predictions in pinned HOST memory (H2D copy of every
             view and a D2H read of the last view's render result inside the timed region)
  roofline   the WHOLE MeshAggregator.add (count stage + scatter stage, SURVEY.md 8d: B_add / t_add) against the
             measured HBM peak, with the scatter kernel alone (back to back, and launch by launch) beside it
  parity     checked BEFORE anything is timed: view 0 rendered by the CUDA path is bit-identical to the CPU oracle's and
             its fusion is within 1e-5 of the oracle's; under torchrun also: the all-reduced N-way accumulator equals
             the accumulator of the same N x B views added on one GPU within 1e-5. A failed check aborts the run.
  cpu_baseline / --impl reference   the GENUINE reference (oracle/_ref, compiled from the reference's sources) on the
             host cores of the same box.

Multi-GPU (torchrun, one rank per GPU): views shard round-robin across ranks, each rank owns a private accumulator, ONE
NCCL all-reduce of it closes the timed region (weak scaling: every rank runs B x cycles views per step; over the cycles
of a step a rank walks through the shards of all ranks, so no rank is stuck with the expensive cameras). `--scaling
strong` times the config's own job instead (cfg3: 500 views dealt over the ranks, all-reduce after every job).
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
# view sees); views = size of the whole job in the reference configuration; B = distinct views resident per GPU;
# cycles = passes over them per step.
CONFIGS = {
    "cfg1": dict(name="icosphere 1280 tris, 256x256, 19 classes", mesh="icosphere", F=1280, W=256, H=256, C=19, views=4,
                 B=4, cycles=256),
    "cfg2": dict(name="ScanNet-scale 500k tris, 640x480, 40 classes", mesh="terrain", F=500_000, W=640, H=480, C=40,
                 views=200, tris_per_view=30_000, B=16, cycles=48),
    "cfg3": dict(name="Cityscapes-scale 2M tris, 2048x1024 (1024x2048 images), 19 classes", mesh="terrain", F=2_000_000,
                 W=2048, H=1024, C=19, views=500, tris_per_view=150_000, B=16, cycles=24),
    "cfg4": dict(name="wide-C 1M tris, 1920x1080, 150 classes", mesh="terrain", F=1_000_000, W=1920, H=1080, C=150,
                 views=100, tris_per_view=120_000, B=4, cycles=24),
    "cfg5": dict(name="dense sweep 5M tris, 1280x720, 19 classes", mesh="terrain", F=5_000_000, W=1280, H=720, C=19,
                 views=2000, tris_per_view=80_000, B=16, cycles=48),
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


def bind_to_gpu_numa_node(physical_index):
    """Run this rank's host threads on the CPUs next to its GPU (NVML's ideal CPU affinity) BEFORE any pinned buffer is
    allocated: pinned pages are then first touched - and so placed - on the GPU's NUMA node, and the staging copies do not
    cross the socket interconnect. Round 1's 8-GPU end-to-end number (0.41 efficiency) had every rank's buffers wherever
    the scheduler happened to start it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return {"cpus": len(cpus), "first": min(cpus), "last": max(cpus), "of": ncpu}
    except Exception as e:
        return {"error": repr(e)}


def reference_arm_imports():
    """The reference arm must not map the product library: `semantic_meshes.synthetic` / `.data` are numpy-only, but
    importing them through the package would run its __init__, which dlopens libsmesh_b200.so. A bare namespace stands in
    for the package in this (separate) process."""
    import types
    if "semantic_meshes" not in sys.modules:
        pkg = types.ModuleType("semantic_meshes")
        pkg.__path__ = [os.path.join(ROOT, "semantic-meshes_b200", "semantic_meshes")]
        sys.modules["semantic_meshes"] = pkg


def product_lib_mapped():
    try:
        with open("/proc/self/maps") as fh:
            return "libsmesh_b200" in fh.read()
    except OSError:
        return None


def build_scene(cfg, rank, n_distinct):
    from semantic_meshes import synthetic
    if cfg["mesh"] == "icosphere":
        mesh = synthetic.mesh("icosphere")
        cams = synthetic.orbit_cameras(n_distinct, cfg["W"], cfg["H"], (0, 0, 0), 3.0, seed=100 + rank, tilt_deg=(0, 180))
    else:
        mesh = synthetic.mesh("terrain", cfg["F"], seed=1234)
        cams = synthetic.terrain_cameras(n_distinct, cfg["W"], cfg["H"], cfg["F"], cfg["tris_per_view"], seed=100 + rank)
    return mesh, cams


def timed_graph(torch, fn, n, use_graph=True):
    """Mean device time of fn() in ms: captured once into a CUDA graph and replayed n times (eager launches from Python
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
        except Exception as e:
            sys.stderr.write(f"graph capture failed ({e!r}); eager timing\n")
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


class Scene:
    """Everything of one config resident on one GPU: prepared mesh, B views' predictions, their cameras (one set per rank
    shard), the rendered index images and per-view statistics."""

    def __init__(self, cfg, rank, world, dev):
        import torch
        import semantic_meshes
        from semantic_meshes import synthetic
        self.cfg = cfg
        W, H, C, B = cfg["W"], cfg["H"], cfg["C"], cfg["B"]
        self.W, self.H, self.C, self.B, self.npix = W, H, C, B, W * H
        # one global, seeded pool of B * world cameras dealt round-robin (distributed.shard_views): shard s = pool[s::world]
        self.mesh, pool = build_scene(cfg, 0, B * world)
        self.shards = [pool[s::world] for s in range(world)]
        self.cams = self.shards[rank]
        self.renderer = semantic_meshes.render.triangles(self.mesh)
        self.P = self.renderer.getPrimitivesNum()
        # B distinct views resident in HBM; one view (>= 49 MB, 160 MB at cfg3) is re-read only after B-1 others, so with
        # B * view bytes >> 126 MB of L2 nothing is served from cache between timed iterations
        self.probs = torch.empty((B, W, H, C), dtype=torch.float32, device=dev)
        for b in range(B):
            synthetic.predictions_torch(W, H, C, seed=1000 * rank + b, device=dev, out=self.probs[b])
        self.ids = torch.empty((B, W, H), dtype=torch.int32, device=dev)
        self.accepted, self.touched, self.covered = [], [], []
        for b in range(B):
            idx, _ = self.renderer.render(self.cams[b])
            self.ids[b] = idx
            valid = idx >= 0
            ok = valid & (self.probs[b].sum(-1) > 0.5)
            self.accepted.append(int(ok.sum().item()))
            self.touched.append(int(torch.unique(idx[ok]).numel()))
            self.covered.append(float(valid.float().mean().item()))
        torch.cuda.synchronize()
        self.bytes_inputs = 4.0 * self.npix * C + 4.0 * self.npix
        # probs + ids read once, touched accumulator rows read + written once
        self.bytes_add = self.bytes_inputs + 8.0 * C * float(np.mean(self.touched))


def check_parity(scene, kinds=("sum",)):
    """View 0 through the CUDA path against the CPU oracle (test infrastructure, used here only as the checker): index and
    depth bit-identical, fused distribution within 1e-5 (mul: 5e-4 after exp). Raises if not."""
    import oracle
    import semantic_meshes
    cam, W, H, C, P = scene.cams[0], scene.W, scene.H, scene.C, scene.P
    t0 = time.perf_counter()
    idx, depth = scene.renderer.render(cam)
    o_idx, o_depth = oracle.raster_render(scene.mesh.vertices, scene.mesh.faces, cam.rotation, cam.translation,
                                          cam.focal_lengths, cam.principal_point, W, H)
    g_idx = idx.cpu().numpy().view(np.uint32)
    n_idx = int((g_idx != o_idx).sum())
    n_depth = int((depth.cpu().numpy().view(np.uint32) != o_depth.view(np.uint32)).sum())
    out = {"raster_pixels_differing": n_idx + n_depth, "raster_bit_exact": n_idx == 0 and n_depth == 0, "fusion_max_err": {}}
    probs = scene.probs[0].cpu().numpy()
    ok = out["raster_bit_exact"]
    for kind in kinds:
        agg = semantic_meshes.fusion.MeshAggregator(P, C, kind)
        ref = oracle.Aggregator(P, C, kind)
        agg.add(idx, scene.probs[0])
        ref.add(o_idx, probs)
        got, exp = agg.get(), ref.get()
        tol = 5e-4 if kind == "mul" else 1e-5
        err = float(np.abs(got - exp).max())
        out["fusion_max_err"][kind] = err
        ok = ok and err <= tol and bool(np.isfinite(got).all())
    out["checked"] = "view 0 of this config vs oracle/ (CPU restatement of the reference), before timing"
    out["seconds"] = round(time.perf_counter() - t0, 2)
    out["ok"] = bool(ok)
    if not ok:
        raise SystemExit(f"bench.py: PARITY FAILED, nothing timed: {json.dumps(out)}")
    return out


def stage_timings(args, scene, agg, use_graph, kinds=("sum",), with_get=True):
    """Stage times on the scene's resident inputs (ms per view): render alone, add alone (count + scatter, batch call),
    count alone, scatter alone; optionally add for the other aggregator kinds and get()."""
    import torch
    import semantic_meshes
    from semantic_meshes import _lib
    lib = _lib.lib
    B, W, H, C, P, npix = scene.B, scene.W, scene.H, scene.C, scene.P, scene.npix
    reps = max(3, min(args.steps, 20))
    kind = _lib.KIND["sum"]
    out = {}

    def add_all(a):
        def f():
            a.restart_epochs()
            a.add_batch(scene.ids, scene.probs)
        return f

    out["render_ms_per_view"] = timed_graph(torch, lambda: [scene.renderer.render(scene.cams[b]) for b in range(B)], reps,
                                            use_graph) / B
    out["add_ms_per_view"] = timed_graph(torch, add_all(agg), reps, use_graph) / B

    def add_serial():
        agg.restart_epochs()
        for b in range(B):
            agg.add(scene.ids[b], scene.probs[b])

    out["add_ms_per_view_serial"] = timed_graph(torch, add_serial, reps, use_graph) / B
    # stages alone: every view's counts in an array of its own (prepared untimed for the scatter launches), the B launches
    # back to back in a CUDA graph (nothing of the host's launch path between the two events), divided by B
    counts_b = torch.zeros((B, max(P, 1)), dtype=torch.int32, device=scene.ids.device)

    def count_all():
        s = torch.cuda.current_stream().cuda_stream
        for b in range(B):
            _lib.check(lib.smesh_fuse_count(scene.ids[b].data_ptr(), _lib.ID_I32, H, 1, W, H, P, counts_b[b].data_ptr(),
                                            1 + (b % 200), None, s))

    def scatter_all():
        s = torch.cuda.current_stream().cuda_stream
        for b in range(B):
            _lib.check(lib.smesh_fuse_scatter(kind, scene.ids[b].data_ptr(), scene.probs[b].data_ptr(), None, npix, C, P,
                                              agg.images_equal_weight, counts_b[b].data_ptr(), 1 + (b % 200),
                                              agg._acc.data_ptr(), s))

    out["count_kernel_ms"] = timed_graph(torch, count_all, reps, use_graph) / B
    counts_b.zero_()
    count_all()
    out["scatter_kernel_ms"] = timed_graph(torch, scatter_all, reps, use_graph) / B
    # for comparison: every launch bracketed by its own pair of events on an otherwise idle stream (this interval also
    # holds the launch latency of one kernel, ~2-4 us)
    pairs = []
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    for b in range(B):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        _lib.check(lib.smesh_fuse_scatter(kind, scene.ids[b].data_ptr(), scene.probs[b].data_ptr(), None, npix, C, P,
                                          agg.images_equal_weight, counts_b[b].data_ptr(), 1 + (b % 200),
                                          agg._acc.data_ptr(), stream))
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    out["scatter_kernel_ms_single_launch_events"] = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
    del counts_b
    for k in kinds:
        if k == "sum":
            continue
        other = semantic_meshes.fusion.MeshAggregator(P, C, k)
        out[f"add_ms_per_view_{k}"] = timed_graph(torch, add_all(other), reps, use_graph) / B
        if with_get:
            out[f"get_ms_{k}"] = timed_graph(torch, lambda: other.get(device=True), 5, use_graph)
        del other
    if with_get:
        out["get_ms"] = timed_graph(torch, lambda: agg.get(device=True), 5, use_graph)
        out["get_roofline_frac"] = (4.0 * P * (agg._cpad + C)) / (out["get_ms"] * 1e-3) / 1e9 / measured_peak_gbs()[0]
    if with_get:
        # render.texels: host constructor (OpenMP over the triangles, like the reference's) + per-view render
        t0 = time.perf_counter()
        tex = semantic_meshes.render.texels(scene.mesh, scene.cams, 0.1)
        out["texels_prepare_s"] = time.perf_counter() - t0
        out["texels_primitives"] = tex.getPrimitivesNum()
        out["texels_render_ms_per_view"] = timed_graph(torch, lambda: [tex.render(scene.cams[b]) for b in range(B)], 5,
                                                       use_graph) / B
        del tex
    out["render_views_per_s"] = 1e3 / out["render_ms_per_view"]
    out["add_views_per_s"] = 1e3 / out["add_ms_per_view"]
    return out


def run_ours(args, cfg):
    import torch
    import semantic_meshes
    from semantic_meshes.pipeline import ViewPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_to_gpu_numa_node(physical_gpu_index(local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    scene = Scene(cfg, rank, world, dev)
    W, H, C, B, P, npix = scene.W, scene.H, scene.C, scene.B, scene.P, scene.npix
    renderer = scene.renderer
    parity = None
    if not args.no_parity:
        parity = check_parity(scene, kinds=("sum", "summax", "mul") if rank == 0 else ("sum",))
    agg = semantic_meshes.fusion.MeshAggregator(primitives=P, classes=C)
    pipe = ViewPipeline(renderer, agg, fused_count=args.fused_count, count_ahead=args.count_ahead, write_depth=True,
                        group=args.group, count_stream=args.count_stream, lanes=args.lanes)
    strong = args.scaling == "strong"
    if strong:
        # the config's own job: cfg["views"] views dealt round-robin; a rank cycles its B resident views to make up its share
        from semantic_meshes.distributed import shard_views
        my_views = len(shard_views(cfg["views"], rank, world))
        view_list = [(scene.cams[v % B], v % B) for v in range(my_views)]
        views_per_step_total = cfg["views"]
    else:
        # cycle c of a step walks the camera shard of rank (rank + c) % world: over a step every rank sees every shard
        cycles = max(1, args.cycles if args.cycles > 0 else cfg["cycles"])
        view_list = [(scene.shards[(rank + c) % world][b], b) for c in range(cycles) for b in range(B)]
        views_per_step_total = len(view_list) * world

    def step():
        agg.restart_epochs()  # the step is replayed as a graph: every replay must see the same count epochs (smesh.h)
        if args.no_overlap:
            for cam, b in view_list:
                idx, _ = renderer.render(cam)
                agg.add(idx, scene.probs[b])
        else:
            # same work, render of view v+1 overlapped with the fusion of view v (two streams)
            pipe.run([c for c, _ in view_list], [scene.probs[b] for _, b in view_list])

    # ---- device-resident throughput: the step's views as a CUDA graph ----
    use_graph = not args.no_graph
    step()  # warm the library's per-kernel configuration and the allocator before capture
    torch.cuda.synchronize()
    # the communicator's first collective of this size pays for its setup: do one before the timed region, and time it
    allreduce_cold_ms = allreduce_warm_ms = None
    if dist is not None:
        scratch = torch.zeros_like(agg._acc)
        e = [torch.cuda.Event(True) for _ in range(3)]
        dist.barrier()
        torch.cuda.synchronize()
        e[0].record()
        dist.all_reduce(scratch)
        e[1].record()
        dist.all_reduce(scratch)
        e[2].record()
        torch.cuda.synchronize()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_cold_ms, allreduce_warm_ms = float(t[0]), float(t[1])
        del scratch

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
    run_views = graph.replay if graph is not None else step

    def run_step():
        run_views()
        if strong:
            agg.allreduce()  # a job ends with its all-reduce (NCCL launched eagerly behind the graph replay)

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
    if not strong:
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
    views_total = args.steps * views_per_step_total
    value = views_total / (elapsed_ms * 1e-3)
    clocks = sampler.summary(t0, t1)

    # ---- multi-GPU result check: the all-reduced N-way accumulator vs the same N x B views added on ONE GPU ----
    if dist is not None and not args.no_parity:
        agg.reset()
        for b in range(B):
            idx, _ = renderer.render(scene.cams[b])
            agg.add(idx, scene.probs[b])
        agg.allreduce()
        multi = agg.state().clone()
        err = 0.0
        if rank == 0:
            from semantic_meshes import synthetic
            single = semantic_meshes.fusion.MeshAggregator(P, C)
            tmp = torch.empty((W, H, C), dtype=torch.float32, device=dev)
            for r in range(world):
                for b in range(B):
                    synthetic.predictions_torch(W, H, C, seed=1000 * r + b, device=dev, out=tmp)
                    idx, _ = renderer.render(scene.shards[r][b])
                    single.add(idx, tmp)
            ref = single.state()
            scale = float(ref.abs().max().item())
            err = float((multi - ref).abs().max().item()) / max(scale, 1e-30)
            parity["multi_gpu_vs_one_gpu_max_rel_err"] = err
            parity["multi_gpu_views"] = world * B
            parity["ok"] = parity["ok"] and err <= 1e-5
            del single, tmp
        flag = torch.tensor([0 if err <= 1e-5 else 1], device=dev)
        dist.all_reduce(flag)
        if int(flag.item()) != 0:
            raise SystemExit(f"bench.py: MULTI-GPU PARITY FAILED (max rel err {err})")
        del multi
        agg.reset()

    # ---- stage timing on the same inputs ----
    full = rank == 0 and not args.quick
    stages = stage_timings(args, scene, agg, use_graph, kinds=("sum", "summax", "mul") if full else ("sum",), with_get=full)
    stages["allreduce_ms"] = allreduce_ms
    if dist is not None:
        # the cheaper end of a sharded job: reduce-scatter + get() on this rank's rows (against all-reduce + full get())
        torch.cuda.synchronize()
        dist.barrier()
        ev = [torch.cuda.Event(True) for _ in range(3)]
        scratch = agg._acc.clone()
        agg.reduce_scatter_get()
        agg.get(device=True)
        torch.cuda.synchronize()
        ev[0].record()
        agg.reduce_scatter_get()
        ev[1].record()
        dist.all_reduce(scratch)
        agg.get(device=True)
        ev[2].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stages["reduce_scatter_get_ms"], stages["allreduce_plus_full_get_ms"] = float(t[0]), float(t[1])
        del scratch
    if allreduce_cold_ms is not None:
        nbytes = agg._acc.numel() * 4
        stages.update({"allreduce_ms_cold_first_call": allreduce_cold_ms, "allreduce_ms_warm": allreduce_warm_ms,
                       "allreduce_bus_gbs_warm": 2.0 * (world - 1) / world * nbytes / (allreduce_warm_ms * 1e-3) / 1e9,
                       "allreduce_bytes": nbytes})
    peak, peak_src = measured_peak_gbs()
    add_s = stages["add_ms_per_view"] * 1e-3
    achieved = scene.bytes_add / add_s / 1e9

    # ---- the public API driven eagerly from Python with device-resident predictions (what a GPU model feeding add() sees) ----
    eager_cams, eager_probs = [c for c, _ in view_list], [scene.probs[b] for _, b in view_list]

    def eager_step():
        pipe.run(eager_cams, eager_probs)  # the timed step's own view list, launched eagerly instead of replayed

    eager_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n_eager = max(2, min(args.steps, 5))
    e0.record()
    for _ in range(n_eager):
        eager_step()
    e1.record()
    torch.cuda.synchronize()
    stages["api_eager_views_per_s"] = n_eager * len(eager_cams) / (e0.elapsed_time(e1) * 1e-3)
    # the same loop driven view by view from Python (render() and add() calls, streams and events as torch objects)
    pipe_py = ViewPipeline(renderer, agg, write_depth=True, native=False)
    pipe_py.run(eager_cams, eager_probs)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n_eager):
        pipe_py.run(eager_cams, eager_probs)
    e1.record()
    torch.cuda.synchronize()
    stages["api_eager_python_loop_views_per_s"] = n_eager * len(eager_cams) / (e0.elapsed_time(e1) * 1e-3)

    # ---- end to end through the public API: predictions in pinned host memory ----
    n_host = min(B, 4)
    host_probs = [torch.empty((W, H, C), dtype=torch.float32, pin_memory=True) for _ in range(n_host)]
    for b in range(n_host):
        host_probs[b].copy_(scene.probs[b])
    host_idx = torch.empty((W, H), dtype=torch.int32, pin_memory=True)
    host_depth = torch.empty((W, H), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(1, min(args.steps, 4))
    # the platform's ceiling for this path: a plain pinned -> device copy of the same buffers, all ranks at once
    stage_buf = torch.empty((W, H, C), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    c0, c1 = torch.cuda.Event(True), torch.cuda.Event(True)
    c0.record()
    for k in range(8):
        stage_buf.copy_(host_probs[k % n_host], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 8 * npix * C * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del stage_buf
    agg.async_host_inputs = True  # the bench refills nothing: leave two uploads in flight (see MeshAggregator.add)

    def e2e_step():
        for b in range(B):
            idx, depth = renderer.render(scene.cams[b])
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
    h2d_min = h2d_gbs
    if dist is not None:
        t = torch.tensor([e2e_ms, -h2d_gbs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, h2d_min = float(t[0]), -float(t[1])
    e2e_value = e2e_steps * B * world / (e2e_ms * 1e-3)
    agg.async_host_inputs = False

    # kernel launches of ours inside the timed region: per view 4 (render: begin/cull, clusters, big, resolve) + 2 (add:
    # count, scatter)
    gpu_launches = args.steps * len(view_list) * 6

    line = {
        "metric": "views/s (render + MeshAggregator.add per view), whole job",
        "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['name']}", "views_per_step_per_gpu": len(view_list),
                   "distinct_views_per_gpu": B, "triangles": P, "W": W, "H": H,
                   "classes": C, "aggregator": "sum", "images_equal_weight": 0.5, "parallelism": f"view-shard x{world}",
                   "pixels_covered": float(np.mean(scene.covered)), "faces_touched_per_view": float(np.mean(scene.touched)),
                   "cache": f"{B} distinct views of {scene.bytes_inputs / 1e6:.0f} MB cycled (>> 126 MB L2)",
                   "cuda_graph": graph is not None, "allreduce_in_timed_region": world > 1,
                   "render_add_overlap": not args.no_overlap, "fused_count": bool(args.fused_count),
                   "views_per_add_batch": args.group, "fusion_lanes": args.lanes,
                   "timed_region_ms": elapsed_ms},
        "mpixel_face_scatters_per_s": value * float(np.mean(scene.accepted)) / 1e6,
        "stages": stages,
        "roofline": {"bound": "hbm", "kernel": "MeshAggregator.add = smesh::fuse::count_runs_kernel + scatter kernel "
                                               "(scatter_pair_kernel at C <= 20, scatter_rows_kernel at wide C)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src,
                     "timing": f"whole add: smesh_fuse_add_batch over {B} views in a CUDA graph, CUDA events / views "
                               "(views dealt to two lanes: the caller's stream and a side stream of the library, one counter array each)",
                     "algorithmic_bytes_per_launch": scene.bytes_add,
                     "input_only_frac": scene.bytes_inputs / add_s / 1e9 / peak,
                     "frac_add_serial": scene.bytes_add / (stages["add_ms_per_view_serial"] * 1e-3) / 1e9 / peak,
                     "frac_scatter_kernel_alone": scene.bytes_add / (stages["scatter_kernel_ms"] * 1e-3) / 1e9 / peak,
                     "frac_scatter_kernel_single_launch":
                         scene.bytes_add / (stages["scatter_kernel_ms_single_launch_events"] * 1e-3) / 1e9 / peak},
        "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": int(B * npix * C * 4),
                "d2h_bytes_per_step": int(npix * 8), "steps": e2e_steps,
                "h2d_gbs_plain_copy_per_gpu_min_over_ranks": h2d_min,
                "h2d_ceiling_views_per_s": world * h2d_min * 1e9 / (npix * C * 4),
                "cpu_affinity": affinity},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
        "parity": parity,
    }
    traffic_file = os.path.join(ROOT, "profiles", "scatter_traffic.json")
    if os.path.exists(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file)).get(args.config)
        except Exception:
            pass

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, P, scene.ids, scene.probs, n_views=args.cpu_views)
    # the other BASELINE.json shapes, compact (single GPU, default config only)
    if rank == 0 and world == 1 and args.also and args.config == "cfg3":
        del pipe, agg, scene, graph, host_probs, renderer
        torch.cuda.empty_cache()
        line["other_configs"] = {}
        for name in args.also.split(","):
            try:
                line["other_configs"][name] = compact_config(args, name, dev, peak)
            except BaseException as e:  # pragma: no cover  (SystemExit of a failed parity check included)
                line["other_configs"][name] = {"error": repr(e)}
            torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def compact_config(args, name, dev, peak):
    """One of the other BASELINE.json shapes in brief: views/s of the overlapped pipeline (CUDA graph), stage times and the
    whole-add roofline fraction."""
    import torch
    import semantic_meshes
    from semantic_meshes.pipeline import ViewPipeline
    cfg = CONFIGS[name]
    scene = Scene(cfg, 0, 1, dev)
    parity = check_parity(scene)
    agg = semantic_meshes.fusion.MeshAggregator(scene.P, scene.C)
    pipe = ViewPipeline(scene.renderer, agg, write_depth=True)

    def step():
        agg.restart_epochs()
        pipe.run(scene.cams, scene.probs)

    ms = timed_graph(torch, step, max(5, min(args.steps, 20)))
    sub = argparse.Namespace(steps=min(args.steps, 10), quick=True)
    st = stage_timings(sub, scene, agg, True, with_get=False)
    add_s = st["add_ms_per_view"] * 1e-3
    return {"workload": cfg["name"], "views_per_s": scene.B / (ms * 1e-3), "render_us_per_view": st["render_ms_per_view"] * 1e3,
            "add_us_per_view": st["add_ms_per_view"] * 1e3, "count_us": st["count_kernel_ms"] * 1e3,
            "scatter_us": st["scatter_kernel_ms"] * 1e3, "add_roofline_frac": scene.bytes_add / add_s / 1e9 / peak,
            "parity_ok": parity["ok"]}


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
    reference_arm_imports()
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
        "product_lib_mapped": product_lib_mapped(),
    }
    print(json.dumps(line))
    tmp.cleanup()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cycles", type=int, default=0, help="passes over the B resident views per step (0 = the config's)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="render and add strictly one after the other on one stream")
    ap.add_argument("--fused-count", action="store_true", help="count the view's pixels per face in the render pass (N2)")
    ap.add_argument("--count-ahead", action="store_true", help="pipeline: the count of view v+1 rides in the scatter of view v")
    ap.add_argument("--count-stream", action="store_true", help="pipeline: the count stage on a third stream")
    ap.add_argument("--lanes", type=int, default=1, help="pipeline: 2 = adds alternate between two fusion streams")
    ap.add_argument("--group", type=int, default=1, help="pipeline: views per add_batch call (1 = one add per view)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity checks (profiling runs only)")
    ap.add_argument("--quick", action="store_true", help="skip the per-kind / get() stage timings")
    ap.add_argument("--also", default="cfg2,cfg4,cfg5", help="other configs measured in brief after the default one ('' = none)")
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
