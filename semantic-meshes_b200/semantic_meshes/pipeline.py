"""Render + fuse many views with the two hot paths overlapped (extension, not in the reference API).

`renderer.render` is ALU/latency bound and `aggregator.add` is HBM bound, so running view v+1's render while view v is
being fused keeps both halves of the GPU busy. Two CUDA streams, one event per view; the results are identical to the
sequential README loop (`idx, _ = renderer.render(cam); aggregator.add(idx, probs)`).
"""
from . import _lib


class ViewPipeline:
    def __init__(self, renderer, aggregator):
        torch = _lib.require_cuda()
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
        for v in range(n + 1):
            nxt = None
            if v < n:
                with torch.cuda.stream(rs):
                    idx, _ = self.renderer.render(cameras[v])
                    ev = torch.cuda.Event()
                    ev.record(rs)
                nxt = (idx, ev)
            if pending is not None:
                idx_prev, ev_prev = pending
                main.wait_event(ev_prev)
                self.aggregator.add(idx_prev, predictions[v - 1], None if weights is None else weights[v - 1])
                idx_prev.record_stream(main)
                if keep_indices:
                    kept.append(idx_prev)
            pending = nxt
        rs.wait_stream(main)  # the renderer's workspace / outputs are not reused before the last add has been enqueued
        return kept if keep_indices else None
