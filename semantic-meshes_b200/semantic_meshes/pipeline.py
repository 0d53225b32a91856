"""Render + fuse many views with the two hot paths overlapped (extension, not in the reference API).

`renderer.render` is ALU/latency bound and `aggregator.add` is HBM bound, so running view v+1's render while view v is
being fused keeps both halves of the GPU busy. Two CUDA streams, one event per view. With fused_count=True the render
also counts the view's pixels per face into the aggregator (`render(camera, count_into=aggregator)`, SURVEY 8f N2: the
index image is not read a second time for the histogram) and `add` is the scatter stage alone; measured on cfg3 this is
slower than counting on the fusion stream (10.2 k vs 10.8 k views/s) because the render stream is the longer of the two,
so it is off by default. The results are identical to the sequential README loop
(`idx, _ = renderer.render(cam); aggregator.add(idx, probs)`).
"""
from . import _lib


class ViewPipeline:
    def __init__(self, renderer, aggregator, fused_count=False, count_ahead=False):
        torch = _lib.require_cuda()
        self.fused_count = bool(fused_count)
        # count_ahead: the count stage of view v+1 rides in the scatter launch of view v (MeshAggregator.add(count_next=)).
        # Off by default: the fusion of view v then has to wait for the render of view v+1, the two streams fall into
        # lock step and the tail of every render runs alone - measured 11.1 k against 12.4 k views/s at config 3.
        self.count_ahead = bool(count_ahead)
        self._torch = torch
        self.renderer, self.aggregator = renderer, aggregator
        with torch.cuda.device(renderer.device):
            self._render_stream = torch.cuda.Stream()

    def run(self, cameras, predictions, weights=None, keep_indices=False):
        """cameras: sequence of data.Camera; predictions: sequence (or batched tensor) of (W, H, C) float32 arrays, one per
        camera; weights: optional sequence of (W, H) float32. Returns the list of index images if keep_indices."""
        torch = self._torch
        main = torch.cuda.current_stream()
        rs = self._render_stream
        rs.wait_stream(main)  # the render stream starts after whatever produced the inputs
        kept = []
        pending = None  # (indices, event) of the view rendered ahead
        n = len(cameras)
        fused = self.fused_count and self.aggregator.primitives == self.renderer.getPrimitivesNum()
        added = []  # events: add of view v enqueued (the counter array of view v is free again after it)
        for v in range(n + 1):
            nxt = None
            if v < n:
                if fused and v >= 2:
                    rs.wait_event(added[v - 2])  # two counter arrays: view v reuses the one of view v - 2
                with torch.cuda.stream(rs):
                    idx, _ = self.renderer.render(cameras[v], count_into=self.aggregator if fused else None)
                    ev = torch.cuda.Event()
                    ev.record(rs)
                nxt = (idx, ev)
            if pending is not None:
                idx_prev, ev_prev = pending
                main.wait_event(ev_prev)
                ride = None
                if self.count_ahead and not fused and nxt is not None:
                    # the fusion of view v-1 waits for the render of view v as well and counts it on the way (the renderer
                    # is then two views ahead of the fusion instead of one: same throughput, one launch less per view)
                    main.wait_event(nxt[1])
                    ride = nxt[0]
                    ride.record_stream(main)
                self.aggregator.add(idx_prev, predictions[v - 1], None if weights is None else weights[v - 1], count_next=ride)
                idx_prev.record_stream(main)
                ev_add = torch.cuda.Event()
                ev_add.record(main)
                added.append(ev_add)
                if keep_indices:
                    kept.append(idx_prev)
            pending = nxt
        rs.wait_stream(main)  # the renderer's workspace / outputs are not reused before the last add has been enqueued
        return kept if keep_indices else None
