"""semantic_meshes.fusion - per-primitive label fusion (python/semantic_meshes/src/Fusion.cu:140-151).

    aggregator = semantic_meshes.fusion.MeshAggregator(primitives=P, classes=C[, aggregator="sum"[, images_equal_weight=0.5]])
    aggregator.add(primitive_indices, probs[, weights])     # (W,H) ints, (W,H,C) float32, (W,H) float32
    annotations = aggregator.get()                          # (P, C) float32 numpy

Semantics of include/semantic_meshes/fusion/Mesh.h:57-133 with the aggregator chains of Fusion.cu:46-92 ("sum",
"summax", "mul"). The accumulator lives on the GPU as a torch tensor; `add` is asynchronous on the current CUDA stream
(host inputs have been read when it returns, like the reference's synchronous add; see `async_host_inputs`).
Views may be sharded over several processes / GPUs: every rank adds its views, then `allreduce()` sums the
accumulators (one NCCL all-reduce) before `get()`.
"""
import numpy as np

from . import _lib

_NONE_MATCHED = "None matched from [torch.Tensor, numpy.ndarray, DLPack capsule / __dlpack__]: "


def _torch_id_dtypes(torch):
    return {torch.uint32: _lib.ID_U32, torch.int32: _lib.ID_I32, torch.uint64: _lib.ID_U64, torch.int64: _lib.ID_I64}


class MeshAggregator:
    """One accumulator row of `classes` floats per primitive.

    aggregator: "sum" (default) | "summax" | "mul" - first letter case-insensitive like the reference (Fusion.cu:126).
    The reference only knows the class counts it was compiled for (CLASSES_NUMS, Fusion.cu:122-125); here `classes` is a
    run-time value.
    """

    def __init__(self, primitives, classes, aggregator="sum", images_equal_weight=0.5, device=None):
        torch = _lib.require_cuda()
        self._torch = torch
        primitives, classes = int(primitives), int(classes)
        if primitives < 0:
            raise ValueError("MeshAggregator: primitives must be >= 0")
        if classes < 1:
            raise ValueError(f"The project does not support the following number of classes: {classes}")
        name = str(aggregator)
        name = name[:1].lower() + name[1:]
        if name not in _lib.KIND:
            # the reference calls an empty std::function here -> std::bad_function_call -> RuntimeError
            raise RuntimeError(f"MeshAggregator: unknown aggregator '{aggregator}' (sum, summax, mul)")
        self.primitives, self.classes = primitives, classes
        self.aggregator = name
        self.images_equal_weight = float(images_equal_weight)
        self._kind = _lib.KIND[name]
        self._cpad = int(_lib.lib.smesh_fuse_padded_classes(classes))
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._id_dtypes = _torch_id_dtypes(torch)
        # raw accumulator [P, Cpad] (mul: -log p), per-view pixel counters [P], flat id scratch (grown on demand)
        self._acc = torch.zeros((primitives, self._cpad), dtype=torch.float32, device=self.device)
        # two counter arrays, used alternately (epoch parity): a renderer may already be counting the next view into one
        # (render(camera, count_into=...)) while this view's scatter still reads the other
        self._counts2 = torch.zeros((2, max(primitives, 1)), dtype=torch.int32, device=self.device)
        self._ids32 = torch.empty((0,), dtype=torch.int32, device=self.device)
        self._epoch = 0      # last count epoch handed out (see include/smesh.h: tagged per-view pixel counters)
        self._epoch_gen = 0  # bumped whenever the counters are zeroed: tokens of earlier counted renders become void
        self._array_epoch = [0, 0]  # epoch last counted into each of the two arrays (a token is valid only while its
        #                             array still holds its view)
        # streams that have touched the counters since they were last zeroed, and the event of that zeroing: a reset on
        # one stream is ordered against the users on the others (render(count_into=...) runs on a render stream)
        self._users, self._reset_event, self._reset_stream, self._reset_captured = set(), None, None, False
        # False (default): a pinned host tensor passed to add() has been copied when add() returns, like the reference's
        # synchronous add - the caller may refill it right away. True: the copy is left in flight (tensor.to(device,
        # non_blocking=True) semantics), the caller must not touch the buffer before the stream has caught up.
        self.async_host_inputs = False
        # upload path of host predictions (see _upload_probs)
        self._copy_stream, self._stage_bufs, self._stage_done, self._stage_next, self._stage_last = None, [None, None], [None, None], 0, None

    # ---------------------------------------------------------------------------------------------------------------
    def _as_tensor(self, obj, what):
        torch = self._torch
        if isinstance(obj, torch.Tensor):
            t = obj
        elif isinstance(obj, np.ndarray):
            if any(s < 0 for s in obj.strides) or not obj.flags.writeable:
                obj = np.ascontiguousarray(obj) if any(s < 0 for s in obj.strides) else obj
                if not obj.flags.writeable:
                    obj = obj.copy()
            t = torch.from_numpy(obj)
        elif hasattr(obj, "__dlpack__") or type(obj).__name__ == "PyCapsule":
            t = torch.utils.dlpack.from_dlpack(obj)
        else:
            raise ValueError(_NONE_MATCHED + f"{what} has type {type(obj).__name__}")
        if t.device != self.device:
            if what == "probs" and t.device.type == "cpu" and t.is_contiguous() and not torch.cuda.is_current_stream_capturing():
                return self._upload_probs(t)
            # (a pinned source stays in use until the copy has run: wait for it unless the caller opted out)
            t = t.to(self.device, non_blocking=self.async_host_inputs or not (t.device.type == "cpu" and t.is_pinned()))
        return t

    def _upload_probs(self, host):
        """Host predictions (159 MB per view at config 3) go up on a copy stream into one of two staging buffers, so the
        upload of view v+1 overlaps the kernels of view v; the compute stream waits for its view's upload, the copy stream
        for the kernels that last read the buffer it is about to overwrite (`_stage_done`, set at the end of add)."""
        torch = self._torch
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            k = self._stage_next
            self._stage_next ^= 1
            buf = self._stage_bufs[k]
            if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
                buf = self._stage_bufs[k] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
                self._stage_done[k] = None
            cs, main = self._copy_stream, torch.cuda.current_stream()
            cs.wait_stream(main) if self._stage_done[k] is None else cs.wait_event(self._stage_done[k])
            with torch.cuda.stream(cs):
                buf.copy_(host, non_blocking=True)
                up = torch.cuda.Event()
                up.record(cs)
            main.wait_event(up)
            self._stage_last = k
            if host.is_pinned() and not self.async_host_inputs:
                up.synchronize()  # pageable memory is staged by the driver before copy_ returns; pinned memory is not
        return buf

    def _release_stage(self):
        """Called after the kernels of an add have been enqueued: the staging buffer they read may be overwritten once
        they are done."""
        k = self._stage_last
        if k is not None:
            ev = self._torch.cuda.Event()
            ev.record(self._torch.cuda.current_stream())
            self._stage_done[k] = ev
            self._stage_last = None

    def _stage(self, primitive_indices, probs, weights):
        """Validate like Fusion.h:42-64 / Mesh.h:68-74 and move to the device. -> (ids, id_dtype, probs, weights)"""
        ids = self._as_tensor(primitive_indices, "primitive_indices")
        pr = self._as_tensor(probs, "probs")
        id_dtype = self._id_dtypes.get(ids.dtype)
        if id_dtype is None or ids.dim() != 2:
            raise ValueError(_NONE_MATCHED + f"primitive_indices must be a rank-2 uint32/int32/uint64/int64 array, got "
                             f"rank {ids.dim()} {ids.dtype}")
        if pr.dtype != self._torch.float32 or pr.dim() != 3:
            raise ValueError(_NONE_MATCHED + f"probs must be a rank-3 float32 array, got rank {pr.dim()} {pr.dtype}")
        wt = None
        if weights is not None:
            wt = self._as_tensor(weights, "weights")
            if wt.dtype != self._torch.float32 or wt.dim() != 2:
                raise ValueError(_NONE_MATCHED + f"weights must be a rank-2 float32 array, got rank {wt.dim()} {wt.dtype}")
        W, H = ids.shape
        if tuple(pr.shape[:2]) != (W, H) or (wt is not None and tuple(wt.shape) != (W, H)):
            raise ValueError(f"Primitive image {tuple(ids.shape)}, probs image {tuple(pr.shape[:2])} and weights image "
                             f"{tuple(wt.shape) if wt is not None else (W, H)} must have the same width and height")
        if pr.shape[2] != self.classes:
            raise ValueError(f"probs image has {pr.shape[2]} classes, aggregator was built for {self.classes}")
        return ids, id_dtype, pr, wt

    def _layout(self, ids, pr, wt):
        """Pick the pixel order in which the probability image is contiguous, so it is never copied (callers pass
        `transpose(pred, (1, 0, 2))` views of (H, W, C) network outputs, python/scripts/colorize_mesh.py:66)."""
        W, H = ids.shape
        C = self.classes
        sx, sy, sc = pr.stride()
        if W * H == 0:
            return None
        if (sc == 1 or C == 1) and (sy == C or H == 1) and (sx == H * C or W == 1):
            x_major = True
        elif (sc == 1 or C == 1) and (sx == C or W == 1) and (sy == W * C or H == 1):
            x_major = False
        else:
            pr = pr.contiguous()
            x_major = True
        if pr.data_ptr() % 16 != 0:
            pr = pr.clone(memory_format=self._torch.contiguous_format)
            x_major = True
        ix, iy = ids.stride()
        if x_major:
            n_outer, n_inner, ids_so, ids_si = W, H, ix, iy
        else:
            n_outer, n_inner, ids_so, ids_si = H, W, iy, ix
        w_so = w_si = 0
        if wt is not None:
            wx, wy = wt.stride()
            w_so, w_si = (wx, wy) if x_major else (wy, wx)
            if not ((w_si == 1 or n_inner == 1) and (w_so == n_inner or n_outer == 1)):
                wt = wt.contiguous() if x_major else wt.t().contiguous().t()
                w_so, w_si = n_inner, 1
        return pr, wt, n_outer, n_inner, ids_so, ids_si, w_so, w_si

    def restart_epochs(self):
        """Zero the per-view pixel counters and start the count epochs over. Called automatically when the 8-bit epoch
        wraps; call it yourself at the start of any region you capture into a CUDA graph, so every replay sees the same
        epochs on clean counters. The zero-fill runs on the current stream after everything the other streams have
        enqueued on the counters so far (counted renders run on a render stream), and later users on other streams wait
        for it."""
        torch = self._torch
        cur = torch.cuda.current_stream(self.device)
        capturing = torch.cuda.is_current_stream_capturing()
        others = [torch.cuda.default_stream(self.device) if h == 0 else torch.cuda.ExternalStream(h, device=self.device)
                  for h in self._users if h != cur.cuda_stream]
        others = [s for s in others if self._may_wait_for(s, capturing)]
        for s in others:
            cur.wait_stream(s)
        self._counts2.zero_()
        self._reset_event = None
        if others:
            self._reset_event = torch.cuda.Event()
            self._reset_event.record(cur)
        self._reset_stream, self._reset_captured = cur.cuda_stream, capturing
        self._users = {cur.cuda_stream}
        self._epoch = 0
        self._epoch_gen += 1
        self._array_epoch = [0, 0]

    def _may_wait_for(self, stream, capturing):
        """While the current stream is being captured into a CUDA graph only streams of the same capture may be waited
        for (a dependency on work outside the capture cannot be expressed; whoever captures starts from a quiet device)."""
        if not capturing:
            return True
        with self._torch.cuda.stream(stream):
            return self._torch.cuda.is_current_stream_capturing()

    def _enter(self):
        """Before anything that reads or writes the counters is enqueued on the current stream (streams are tracked by
        their raw handles: this runs once per view)."""
        torch = self._torch
        h = _lib.raw_stream(torch, self._dev_index)
        if h not in self._users:
            ev = self._reset_event
            # (an event recorded inside a capture means something only to streams of that capture)
            if ev is not None and h != self._reset_stream and self._reset_captured == torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(self.device).wait_event(ev)
            self._users.add(h)

    def _counts_for(self, epoch):
        return self._counts2[epoch & 1]

    @property
    def _counts(self):
        """The counter array of the NEXT epoch to be handed out (tools that drive the C ABI stage by stage use it with
        explicit epochs of one parity; `add` picks the array from the epoch itself)."""
        return self._counts2[0]

    def _next_epochs(self, npix, n=1):
        """-> first of n consecutive count epochs for views of npix pixels (0 = untagged mode for huge images)."""
        if npix >= (1 << 24):
            if self._epoch != 0:
                self.restart_epochs()
            self._enter()
            return 0
        if self._epoch + n > 255:
            self.restart_epochs()
        self._enter()
        first = self._epoch + 1
        self._epoch += n
        for e in range(max(first, self._epoch - 1), self._epoch + 1):
            self._array_epoch[e & 1] = e
        return first

    def _scratch(self, npix):
        if self._ids32.numel() < npix:
            self._ids32 = self._torch.empty((npix,), dtype=self._torch.int32, device=self.device)
        return self._ids32

    # ---------------------------------------------------------------------------------------------------------------
    def add(self, primitive_indices, probs, weights=None, count_next=None):
        """Fuse one view (ModelAggregator::add, Mesh.h:65-107). If `primitive_indices` comes from
        `renderer.render(camera, count_into=self)`, or was the `count_next` of the previous add, the per-face pixel counts
        are already in place and only the scatter stage runs.

        count_next (extension): the index image of the view that will be added NEXT (a device int32 tensor, e.g. from a
        renderer running one view ahead). Its count stage then rides in this view's scatter launch (one extra warp per
        CTA, smesh_fuse_scatter_count_next) instead of being a launch of its own in front of the next scatter.

        The kernels are asynchronous on the current CUDA stream (the reference's add is synchronous); host inputs have
        been read when the call returns, so the caller may reuse its buffers like with the reference (set
        `async_host_inputs = True` to leave the copy of a PINNED host tensor in flight instead)."""
        torch = self._torch
        ids, id_dtype, pr, wt = self._stage(primitive_indices, probs, weights)
        lay = self._layout(ids, pr, wt)
        if lay is None or self.primitives == 0:
            return
        pr, wt, n_outer, n_inner, ids_so, ids_si, w_so, w_si = lay
        npix = n_outer * n_inner
        stream = _lib.raw_stream(torch, self._dev_index)
        flat32 = ((ids.dtype == torch.int32 and self.primitives <= 0x7FFFFFFF) or ids.dtype == torch.uint32) \
            and (ids_si == 1 or n_inner == 1) and (ids_so == n_inner or n_outer == 1)
        token = getattr(primitive_indices, "_smesh_counted", None)
        counted = (token is not None and token[0] == id(self) and token[1] == self._epoch_gen and flat32
                   and self._array_epoch[token[2] & 1] == token[2])  # no later view has counted into that array since
        nxt = self._rider_candidate(count_next) if count_next is not None else None
        # this view's scatter can carry the next view's count stage: two epochs without a wrap, and the next view's counter
        # array (epoch parity) is not the one this view's counts sit in
        if (nxt is not None and flat32 and npix < (1 << 24) and self._epoch + 2 <= 255
                and (not counted or ((self._epoch + 1) ^ token[2]) & 1)):
            epoch = token[2] if counted else self._next_epochs(npix)
            self._enter()
            epoch2 = self._next_epochs(nxt.numel())
            with _lib.on_device(torch, self._dev_index):
                rc = _lib.lib.smesh_fuse_scatter_count_next(
                    self._kind, ids.data_ptr(), pr.data_ptr(), wt.data_ptr() if wt is not None else None, npix,
                    self.classes, self.primitives, self.images_equal_weight, self._counts_for(epoch).data_ptr(), epoch,
                    1 if counted else 0, nxt.data_ptr(), nxt.numel(), self._counts_for(epoch2).data_ptr(), epoch2,
                    self._acc.data_ptr(), stream)
            self._release_stage()
            _lib.check(rc)
            count_next._smesh_counted = (id(self), self._epoch_gen, epoch2)
            return
        if counted:
            epoch = token[2]
            self._enter()
            with _lib.on_device(torch, self._dev_index):
                rc = _lib.lib.smesh_fuse_scatter(self._kind, ids.data_ptr(), pr.data_ptr(),
                                                 wt.data_ptr() if wt is not None else None, npix, self.classes,
                                                 self.primitives, self.images_equal_weight,
                                                 self._counts_for(epoch).data_ptr(), epoch, self._acc.data_ptr(), stream)
            self._release_stage()
            _lib.check(rc)
            return
        epoch = self._next_epochs(npix)
        with _lib.on_device(torch, self._dev_index):
            rc = _lib.lib.smesh_fuse_add(self._kind, ids.data_ptr(), id_dtype, ids_so, ids_si, pr.data_ptr(),
                                         wt.data_ptr() if wt is not None else None, w_so, w_si, n_outer, n_inner,
                                         self.classes, self.primitives, self.images_equal_weight,
                                         self._counts_for(epoch).data_ptr(), epoch,
                                         self._scratch(npix).data_ptr(), self._acc.data_ptr(), stream)
        self._release_stage()
        _lib.check(rc)

    def precount(self, primitive_indices):
        """Extension: take the per-face pixel counts of an index image NOW, on the current stream (the count stage of
        `add` as a call of its own); the `add` of that image then runs its scatter stage only. For callers that can spare a
        stream: the count stage is bound by latency, not by throughput, and overlaps with anything. Returns False (and
        does nothing) if the image cannot be counted ahead (not a flat int32 / uint32 device image below 2^24 pixels, or
        the 8-bit epoch is about to wrap)."""
        torch = self._torch
        t = self._rider_candidate(primitive_indices)
        if t is None or self.primitives == 0 or self._epoch + 1 > 255:
            return False
        epoch = self._next_epochs(t.numel())
        with _lib.on_device(torch, self._dev_index):
            rc = _lib.lib.smesh_fuse_count(t.data_ptr(), _lib.ID_U32 if t.dtype == torch.uint32 else _lib.ID_I32, t.shape[1], 1,
                                           t.shape[0], t.shape[1], self.primitives, self._counts_for(epoch).data_ptr(), epoch,
                                           None, _lib.raw_stream(torch, self._dev_index))
        _lib.check(rc)
        primitive_indices._smesh_counted = (id(self), self._epoch_gen, epoch)
        return True

    def _rider_candidate(self, count_next):
        """The next view's index image if its count stage can ride in this view's scatter launch: a contiguous int32 /
        uint32 device tensor of fewer than 2^24 pixels (the histogram does not depend on the pixel order)."""
        torch = self._torch
        t = count_next
        if not isinstance(t, torch.Tensor) or t.device != self.device or t.dim() != 2 or not t.is_contiguous():
            return None
        if not (t.dtype == torch.uint32 or (t.dtype == torch.int32 and self.primitives <= 0x7FFFFFFF)):
            return None
        if t.numel() == 0 or t.numel() >= (1 << 24) or t.data_ptr() % 4 != 0:
            return None
        return t

    def add_batch(self, primitive_indices, probs, weights=None):
        """Extension: fuse B views held in batched device tensors (B,W,H) / (B,W,H,C) / (B,W,H) with one call; the same
        result as B `add` calls in order (up to the order of the float additions), without B trips through Python - and
        with the views dealt to two lanes (the current stream and a side stream of the library, one counter array each),
        so that one view's count stage, launch gaps and tail run under another view's scatter (smesh_fuse_add_batch)."""
        torch = self._torch
        ids = self._as_tensor(primitive_indices, "primitive_indices")
        pr = self._as_tensor(probs, "probs")
        wt = self._as_tensor(weights, "weights") if weights is not None else None
        if ids.dim() != 3 or pr.dim() != 4 or ids.shape[0] != pr.shape[0]:
            raise ValueError("add_batch expects (B,W,H) indices and (B,W,H,C) probs")
        B = ids.shape[0]
        if B == 0:
            return
        ids0, id_dtype, pr0, wt0 = self._stage(ids[0], pr[0], wt[0] if wt is not None else None)
        lay = self._layout(ids0, pr0, wt0)
        if lay is None or self.primitives == 0:
            return
        pr0c, wt0c, n_outer, n_inner, ids_so, ids_si, w_so, w_si = lay
        if pr0c.data_ptr() != pr[0].data_ptr() or (wt is not None and wt0c.data_ptr() != wt[0].data_ptr()) \
                or (pr.stride(0) * 4) % 16 != 0:
            # layouts that need a copy: go view by view. The whole batch may sit in ONE upload buffer: it is released
            # (may be overwritten by the next upload) only after the last view's kernels, not after the first's.
            staged, self._stage_last = self._stage_last, None
            for b in range(B):
                self.add(ids[b], pr[b], wt[b] if wt is not None else None)
            self._stage_last = staged
            self._release_stage()
            return
        npix = n_outer * n_inner
        done = 0
        while done < B:
            nb = min(B - done, 255)
            epoch0 = self._next_epochs(npix, nb)
            with torch.cuda.device(self.device):
                rc = _lib.lib.smesh_fuse_add_batch(
                    self._kind, nb, ids[done].data_ptr(), id_dtype, ids.stride(0), ids_so, ids_si, pr[done].data_ptr(),
                    pr.stride(0), wt[done].data_ptr() if wt is not None else None, wt.stride(0) if wt is not None else 0,
                    w_so, w_si, n_outer, n_inner, self.classes, self.primitives, self.images_equal_weight,
                    self._counts_for(0).data_ptr(), epoch0, self._scratch(npix).data_ptr(), self._acc.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
            done += nb
        self._release_stage()

    def reset(self):
        """ModelAggregator::reset (Mesh.h:119-122): every row back to the aggregator's zero (mul: -log 1 = 0)."""
        self._acc.zero_()

    def get(self, device=False, rows=None):
        """Per-primitive class distribution (P, C) (Fusion.h:72-76): the accumulator row (mul: exp(-(l - min l))),
        L1-normalised, NaN/Inf -> 0. numpy by default like the reference; device=True returns the torch CUDA tensor.
        rows=(first, last) (extension): only that slice of primitives."""
        torch = self._torch
        first, last = (0, self.primitives) if rows is None else (int(rows[0]), int(rows[1]))
        if not (0 <= first <= last <= self.primitives):
            raise ValueError("get(rows=...): row range outside the primitives")
        out = torch.empty((last - first, self.classes), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib.smesh_fuse_get(self._kind, self._acc[first:last].data_ptr() if last > first else None, last - first,
                                         self.classes, out.data_ptr() if last > first else None,
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        return out if device else out.cpu().numpy()

    # ---------------------------------------------------------------------------------------------------------------
    # what the reference's scripts do with get(), on the device
    def labels(self, dont_care_threshold=0.9, device=False):
        """Per-primitive class index (python/scripts/colorize_mesh.py:82-88): argmax of get(), -1 where the primitive
        received no annotation (its distribution sums to less than dont_care_threshold). int32 (P,), numpy by default."""
        torch = self._torch
        dist = self.get(device=True)
        out = torch.empty((self.primitives,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib.smesh_fuse_labels(dist.data_ptr(), self.primitives, self.classes, float(dont_care_threshold),
                                            out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        return out if device else out.cpu().numpy()

    def colors(self, class_to_color, dont_care_threshold=0.9, device=False):
        """Per-primitive colour for `data.Ply.save` (colorize_mesh.py:86-92): class_to_color[label], black where the
        primitive has no annotation. class_to_color: (classes, 3) uint8. uint8 (P, 3)."""
        torch = self._torch
        table = self._as_tensor(np.ascontiguousarray(class_to_color, dtype=np.uint8), "class_to_color")
        if tuple(table.shape) != (self.classes, 3):
            raise ValueError("class_to_color must have shape (classes, 3)")
        lab = self.labels(dont_care_threshold, device=True).long()
        col = table[lab.clamp(min=0)]
        col[lab < 0] = 0
        return col if device else col.cpu().numpy()

    def render(self, primitive_indices, annotations, background):
        """ModelRenderer::render (include/semantic_meshes/fusion/Mesh.h:24-43): the image of per-primitive annotations,
        out[x, y] = annotations[primitive_indices[x, y]] where that index is a primitive, else background.
        annotations: (P,) or (P, K) array of any 1-, 2-, 4-byte ... dtype (labels, colours, distributions); background:
        one element. Returns a torch CUDA tensor of shape (W, H) or (W, H, K)."""
        torch = self._torch
        ids = self._as_tensor(primitive_indices, "primitive_indices")
        id_dtype = self._id_dtypes.get(ids.dtype)
        if id_dtype is None or ids.dim() != 2:
            raise ValueError(_NONE_MATCHED + "primitive_indices must be a rank-2 uint32/int32/uint64/int64 array")
        ann = self._as_tensor(annotations, "annotations").contiguous()
        if ann.dim() not in (1, 2) or ann.shape[0] != self.primitives:
            raise ValueError("annotations must have one row per primitive")
        bg = torch.as_tensor(background, dtype=ann.dtype, device=self.device).reshape(ann.shape[1:]).contiguous()
        W, H = ids.shape
        if id_dtype in (_lib.ID_U64, _lib.ID_I64):
            ids32 = torch.where((ids >= 0) & (ids < self.primitives), ids, torch.full_like(ids, -1)).to(torch.int32)
        else:
            ids32 = ids.view(torch.int32) if ids.dtype != torch.int32 else ids
        ids32 = ids32.contiguous()
        out = torch.empty((W, H) + tuple(ann.shape[1:]), dtype=ann.dtype, device=self.device)
        elem_bytes = ann.element_size() * (ann.shape[1] if ann.dim() == 2 else 1)
        with torch.cuda.device(self.device):
            rc = _lib.lib.smesh_fuse_render(ann.data_ptr(), self.primitives, elem_bytes, ids32.data_ptr(), W * H,
                                            bg.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        return out

    # ---------------------------------------------------------------------------------------------------------------
    # extensions for view-sharded multi-GPU runs and checkpointing
    def allreduce(self, group=None):
        """Sum the raw accumulators of all ranks in place (one NCCL all-reduce of P x Cpad floats). Every aggregator
        kind accumulates by addition (mul adds -log p, +inf stays absorbing), so the sum of per-rank accumulators equals
        the accumulator of all views added on one GPU up to float reassociation."""
        from .distributed import allreduce_accumulator
        allreduce_accumulator(self._acc, group=group)

    def reduce_scatter_get(self, group=None):
        """Cheaper end of a view-sharded job when every rank only needs ITS slice of the result: one reduce-scatter (half
        the bytes of an all-reduce on the wire) leaves rank r with the summed accumulator rows [first, last) of its
        slice, and get() runs on those rows only. -> ((first, last), distribution rows as a torch CUDA tensor). The
        local accumulator is left untouched."""
        from .distributed import reduce_scatter_rows
        (first, last), part = reduce_scatter_rows(self._acc, group=group)
        torch = self._torch
        out = torch.empty((last - first, self.classes), dtype=torch.float32, device=self.device)
        if last > first:
            with torch.cuda.device(self.device):
                rc = _lib.lib.smesh_fuse_get(self._kind, part.data_ptr(), last - first, self.classes, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
        return (first, last), out

    def state(self):
        """Raw accumulator (P, C) as a torch CUDA tensor (a view without the alignment padding)."""
        return self._acc[:, :self.classes]

    def load_state(self, acc):
        acc = self._as_tensor(acc, "state")
        if tuple(acc.shape) != (self.primitives, self.classes) or acc.dtype != self._torch.float32:
            raise ValueError("load_state expects a (primitives, classes) float32 array")
        self._acc.zero_()
        self._acc[:, :self.classes] = acc
