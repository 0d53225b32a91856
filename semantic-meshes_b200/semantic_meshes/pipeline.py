"""Render + fuse many views with the two hot paths overlapped (extension, not in the reference API).

`renderer.render` is ALU/latency bound and `aggregator.add` is HBM bound, so running view v+1's render while view v is
being fused keeps both halves of the GPU busy. Two CUDA streams, one event per view. With fused_count=True the render
also counts the view's pixels per face into the aggregator (`render(camera, count_into=aggregator)`, SURVEY 8f N2: the
index image is not read a second time for the histogram) and `add` is the scatter stage alone; measured on cfg3 this is
slower than counting on the fusion stream (10.2 k vs 10.8 k views/s) because the render stream is the longer of the two,
so it is off by default. With group=K > 1 the views go K at a time: K renders into one index buffer on the render
stream, then ONE `add_batch` of the group on the fusion stream (one count launch per group, the other counts ride in the
scatter launches) while the next group renders. The results are identical to the sequential README loop
(`idx, _ = renderer.render(cam); aggregator.add(idx, probs)`).
"""
import ctypes
import os

from . import _lib


class ViewPipeline:
    def __init__(self, renderer, aggregator, fused_count=False, count_ahead=False, write_depth=False, group=1,
                 count_stream=False, lanes=1, native=True):
        torch = _lib.require_cuda()
        self.fused_count = bool(fused_count)
        # count_ahead: the count stage of view v+1 rides in the scatter launch of view v (MeshAggregator.add(count_next=)).
        # Off by default: the fusion of view v then has to wait for the render of view v+1, the two streams fall into
        # lock step and the tail of every render runs alone - measured 11.1 k against 12.4 k views/s at config 3.
        self.count_ahead = bool(count_ahead)
        # write_depth: materialise the depth image of every view although the pipeline has no use for it (bench.py does,
        # so that a timed view is exactly one reference-style render() + add())
        self.write_depth = bool(write_depth)
        # count_stream: the count stage of every view on a THIRD stream (MeshAggregator.precount), between its render and
        # its scatter. Off by default: measured 12.3 - 12.4 k against 12.6 k views/s at config 3 (the step is bound by the
        # SMs' total work, not by the length of either stream)
        self.count_stream = bool(count_stream)
        # group: views per add_batch call (1 = one add per view); needs predictions that form a regular batch in memory
        self.group = max(1, int(group))
        # lanes: 2 = the adds alternate between the caller's stream and a second fusion stream (views v and v+1 use
        # different counter arrays of the aggregator, so their count + scatter stages are independent; the accumulator
        # updates are atomic and commute): one view's count stage, launch gaps and tail run under the other's scatter.
        # Off by default: that is what makes add_batch 9 % faster on its own, but under the renderer the SMs have no idle
        # time left to fill - measured 11.1 k against 12.6 k views/s at config 3 (12.3 k with a high-priority render
        # stream), profiles/r02ad_pipeline_lanes.txt
        self.lanes = 2 if int(lanes) == 2 and not (self.fused_count or self.count_ahead or self.count_stream) else 1
        # native: the loop is enqueued by the library (smesh_pipeline_views) whenever the inputs qualify (_run_native) and no
        # other mode is selected
        self.native = bool(native) and not (self.fused_count or self.count_ahead or self.count_stream) and self.group == 1 \
            and self.lanes == 1
        # index images in flight in the native loop (with 2 the renderer waits for the fusion of the view before last and
        # the two streams fall into lock step: 11.0 k against 12.6 k views/s at config 3)
        self.ring = max(1, min(8, int(os.environ.get("SMESH_PIPELINE_RING", "4"))))
        self._ring = self._depth_ring = None
        self._torch = torch
        self.renderer, self.aggregator = renderer, aggregator
        with torch.cuda.device(renderer.device):
            self._render_stream = torch.cuda.Stream()
            self._count_stream = torch.cuda.Stream() if self.count_stream else None
            self._fuse_stream2 = torch.cuda.Stream() if self.lanes == 2 else None

    def run(self, cameras, predictions, weights=None, keep_indices=False):
        """cameras: sequence of data.Camera; predictions: sequence (or batched tensor) of (W, H, C) float32 arrays, one per
        camera; weights: optional sequence of (W, H) float32. Returns the list of index images if keep_indices."""
        torch = self._torch
        if self.native and not keep_indices and len(cameras) > 0 and self._run_native(cameras, predictions, weights):
            return None
        if self.group > 1 and not self.fused_count and len(cameras) > 0:
            done = self._run_grouped(cameras, predictions, weights, keep_indices)
            if done is not False:
                return done
        main = torch.cuda.current_stream()
        rs = self._render_stream
        rs.wait_stream(main)  # the render stream starts after whatever produced the inputs
        kept = []
        pending = None  # (indices, event) of the view rendered ahead
        n = len(cameras)
        fused = self.fused_count and self.aggregator.primitives == self.renderer.getPrimitivesNum()
        added = []  # events: add of view v enqueued (the counter array of view v is free again after it)
        fuse = [main] if self._fuse_stream2 is None else [main, self._fuse_stream2]
        for f in fuse[1:]:
            f.wait_stream(main)
        for v in range(n + 1):
            nxt = None
            if v < n:
                if fused and v >= 2:
                    rs.wait_event(added[v - 2])  # two counter arrays: view v reuses the one of view v - 2
                with torch.cuda.stream(rs):
                    idx, _ = self.renderer.render(cameras[v], count_into=self.aggregator if fused else None,
                                                  depth=self.write_depth)
                    ev = torch.cuda.Event()
                    ev.record(rs)
                if self._count_stream is not None and not fused:
                    cs = self._count_stream
                    cs.wait_event(ev)
                    if v >= 2:
                        cs.wait_event(added[v - 2])  # two counter arrays: view v counts into the one view v - 2 used
                    with torch.cuda.stream(cs):
                        self.aggregator.precount(idx)
                        idx.record_stream(cs)
                        ev = torch.cuda.Event()
                        ev.record(cs)
                nxt = (idx, ev)
            if pending is not None:
                idx_prev, ev_prev = pending
                f = fuse[(v - 1) % len(fuse)]
                f.wait_event(ev_prev)
                ride = None
                if self.count_ahead and not fused and nxt is not None:
                    # the fusion of view v-1 waits for the render of view v as well and counts it on the way (the renderer
                    # is then two views ahead of the fusion instead of one: same throughput, one launch less per view)
                    f.wait_event(nxt[1])
                    ride = nxt[0]
                    ride.record_stream(f)
                with torch.cuda.stream(f):
                    self.aggregator.add(idx_prev, predictions[v - 1], None if weights is None else weights[v - 1],
                                        count_next=ride)
                idx_prev.record_stream(f)
                ev_add = torch.cuda.Event()
                ev_add.record(f)
                added.append(ev_add)
                if keep_indices:
                    kept.append(idx_prev)
            pending = nxt
        for f in fuse[1:]:
            main.wait_stream(f)
        rs.wait_stream(main)  # the renderer's workspace / outputs are not reused before the last add has been enqueued
        if self._count_stream is not None:
            main.wait_stream(self._count_stream)
        return kept if keep_indices else None

    def _run_native(self, cameras, predictions, weights):
        """The same loop enqueued by the library (smesh_pipeline_views: one C call per <= 200 views instead of ~10 Python
        level calls per view - the per-view host time of the loop above is ~90 us, more than the GPU needs). -> False if
        the inputs do not qualify: views of one resolution below 2^24 pixels, predictions (and weights) that are contiguous,
        16-byte aligned float32 tensors on the aggregator's device, a plain triangle renderer."""
        import numpy as np
        from .render import TriangleRenderer
        torch = self._torch
        ren, agg = self.renderer, self.aggregator
        if type(ren) is not TriangleRenderer or ren.device != agg.device or ren._F == 0 or agg.primitives == 0:
            return False
        n = len(cameras)
        W, H = cameras[0].resolution
        if W < 1 or H < 1 or W * H >= (1 << 24) or any(c.resolution != (W, H) for c in cameras):
            return False
        C = agg.classes
        if len(predictions) != n or (weights is not None and len(weights) != n):
            return False
        shape_p, shape_w, f32 = (W, H, C), (W, H), torch.float32
        pp = (ctypes.c_void_p * n)()
        for i in range(n):
            p = predictions[i]
            if (not isinstance(p, torch.Tensor) or p.dtype != f32 or tuple(p.shape) != shape_p or p.device != agg.device
                    or not p.is_contiguous()):
                return False
            pp[i] = p.data_ptr()
            if pp[i] % 16 != 0:
                return False
        wp = None
        if weights is not None:
            wp = (ctypes.c_void_p * n)()
            for i in range(n):
                w = weights[i]
                if (not isinstance(w, torch.Tensor) or w.dtype != f32 or tuple(w.shape) != shape_w or w.device != agg.device
                        or not w.is_contiguous()):
                    return False
                wp[i] = w.data_ptr()
        ptr = ctypes.sizeof(ctypes.c_void_p)
        R = np.ascontiguousarray(np.stack([np.asarray(c.rotation, dtype=np.float32).reshape(9) for c in cameras]))
        t = np.ascontiguousarray(np.stack([np.asarray(c.translation, dtype=np.float32).reshape(3) for c in cameras]))
        f = np.ascontiguousarray(np.stack([np.asarray(c.focal_lengths, dtype=np.float64).reshape(2) for c in cameras]))
        c0 = np.ascontiguousarray(np.stack([np.asarray(c.principal_point, dtype=np.float64).reshape(2) for c in cameras]))
        with _lib.on_device(torch, ren._dev_index):
            ws = ren._ensure_workspace(W, H)
            nring = self.ring
            if self._ring is None or tuple(self._ring.shape) != (nring, W, H):
                self._ring = torch.empty((nring, W, H), dtype=torch.int32, device=ren.device)
                self._depth_ring = torch.empty((nring, W, H), dtype=torch.float32, device=ren.device) if self.write_depth else None
            stream = _lib.raw_stream(torch, ren._dev_index)
            first = 0
            while first < n:
                k = min(200, n - first)
                epoch0 = agg._next_epochs(W * H, k)
                rc = _lib.lib.smesh_pipeline_views(
                    ren._mesh_ptr, ren._mesh_bytes, ren._V, ren._F, k, R[first:].ctypes.data, t[first:].ctypes.data,
                    f[first:].ctypes.data, c0[first:].ctypes.data, W, H, ws.data_ptr(), ws.numel(), nring, self._ring.data_ptr(),
                    self._depth_ring.data_ptr() if self._depth_ring is not None else None, agg._kind,
                    ctypes.addressof(pp) + first * ptr, ctypes.addressof(wp) + first * ptr if wp is not None else None,
                    C, agg.primitives, agg.images_equal_weight, agg._counts2.data_ptr(), epoch0, agg._acc.data_ptr(),
                    stream)
                _lib.check(rc)
                first += k
        return True

    @staticmethod
    def _as_batch(torch, items, first, k):
        """items[first : first + k] as ONE (k, ...) tensor without a copy, or None: a batched tensor is sliced; a sequence of
        tensors qualifies when its elements are equally shaped views of one storage at a constant distance."""
        if isinstance(items, torch.Tensor):
            return items[first:first + k]
        a = items[first]
        if not isinstance(a, torch.Tensor) or not a.is_cuda:
            return None
        if k == 1:
            return a.unsqueeze(0)
        step = items[first + 1].data_ptr() - a.data_ptr() if isinstance(items[first + 1], torch.Tensor) else 0
        if step <= 0 or step % a.element_size() != 0:
            return None
        for j in range(1, k):
            b = items[first + j]
            if (not isinstance(b, torch.Tensor) or b.shape != a.shape or b.stride() != a.stride() or b.dtype != a.dtype
                    or b.device != a.device or b.data_ptr() != a.data_ptr() + j * step
                    or b.untyped_storage().data_ptr() != a.untyped_storage().data_ptr()):
                return None
        return torch.as_strided(a, (k,) + tuple(a.shape), (step // a.element_size(),) + tuple(a.stride()))

    def _run_grouped(self, cameras, predictions, weights, keep_indices):
        """K views at a time (see the module docstring). -> False if the inputs do not form regular batches."""
        torch = self._torch
        n, K = len(cameras), self.group
        res = cameras[0].resolution
        if any(c.resolution != res for c in cameras):
            return False
        groups = [(g, min(K, n - g)) for g in range(0, n, K)]
        batches = []
        for g, k in groups:
            pb = self._as_batch(torch, predictions, g, k)
            wb = self._as_batch(torch, weights, g, k) if weights is not None else None
            if pb is None or (weights is not None and wb is None):
                return False
            batches.append((pb, wb))
        main = torch.cuda.current_stream()
        rs = self._render_stream
        rs.wait_stream(main)
        W, H = res
        kept = []
        pending = None
        for gi in range(len(groups) + 1):
            nxt = None
            if gi < len(groups):
                g, k = groups[gi]
                with torch.cuda.stream(rs):
                    buf = torch.empty((k, W, H), dtype=torch.int32, device=self.renderer.device)
                    for j in range(k):
                        self.renderer.render(cameras[g + j], depth=self.write_depth, out_indices=buf[j])
                    ev = torch.cuda.Event()
                    ev.record(rs)
                nxt = (buf, ev, gi)
            if pending is not None:
                buf_prev, ev_prev, gp = pending
                main.wait_event(ev_prev)
                self.aggregator.add_batch(buf_prev, batches[gp][0], batches[gp][1])
                buf_prev.record_stream(main)
                if keep_indices:
                    kept.extend(buf_prev[j] for j in range(buf_prev.shape[0]))
            pending = nxt
        rs.wait_stream(main)
        return kept if keep_indices else None
