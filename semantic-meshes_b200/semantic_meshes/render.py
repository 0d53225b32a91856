"""semantic_meshes.render - triangle rasterizer (python/semantic_meshes/src/Render.cu:3-24).

    renderer = semantic_meshes.render.triangles(mesh)
    primitive_indices, depth = renderer.render(camera)

Follows Renderer<T>::render (python/semantic_meshes/include/Renderer.h:25-43): both images are (W, H) device arrays,
pixel (x, y) at [x, y]; nothing hit = index 0xFFFFFFFF and depth +inf.
"""
import ctypes

import numpy as np

from . import _lib
from .data import Camera, Ply


class TriangleRenderer:
    """Device-resident mesh + scratch; one `render(camera)` per view (include/semantic_meshes/render/TriangleRenderer.h)."""

    def __init__(self, ply, device=None):
        torch = _lib.require_cuda()
        if not isinstance(ply, Ply):
            raise TypeError("render.triangles expects a semantic_meshes.data.Ply")
        self._torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        verts = np.ascontiguousarray(ply.vertices, dtype=np.float32)
        faces = np.ascontiguousarray(ply.faces, dtype=np.int32)
        if faces.size and (faces.min() < 0 or faces.max() >= verts.shape[0]):
            raise ValueError("render.triangles: face index out of range")
        self._V, self._F = int(verts.shape[0]), int(faces.shape[0])
        # what TriangleRenderer's ctor uploads (TriangleRenderer.h:30-39), turned into the prepared mesh once:
        # Morton-sorted faces in units of 32, each one contiguous block of vertices + a bounding sphere (include/smesh.h)
        mesh_bytes, temp_bytes = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _lib.check(_lib.lib.smesh_raster_mesh_bytes(self._V, self._F, ctypes.byref(mesh_bytes), ctypes.byref(temp_bytes)))
        with torch.cuda.device(self.device):
            verts_d = torch.from_numpy(verts).to(self.device)
            faces_d = torch.from_numpy(faces).to(self.device)
            self._mesh = torch.zeros(mesh_bytes.value, dtype=torch.uint8, device=self.device)
            temp = torch.empty(temp_bytes.value, dtype=torch.uint8, device=self.device)
            _lib.check(_lib.lib.smesh_raster_mesh_build(verts_d.data_ptr(), self._V, faces_d.data_ptr(), self._F,
                                                        self._mesh.data_ptr(), self._mesh.numel(), temp.data_ptr(),
                                                        temp.numel(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.current_stream().synchronize()  # temp / verts_d / faces_d are released on return
        self._mesh_ptr, self._mesh_bytes = self._mesh.data_ptr(), self._mesh.numel()
        self._workspace = None
        self._workspace_res = None

    def getPrimitivesNum(self):
        return self._F

    def face_flags(self):
        """Diagnostics: uint8 numpy (F,), 1 where the prepared mesh tagged the face "well shaped" (include/smesh.h)."""
        import numpy as np
        F = self._F
        nu = (F + 31) // 32
        # unit blocks come first in the prepared mesh: per face 3 float4 {v0, index bits}, {v1, flags}, {v2, 0}
        rec = self._mesh[:nu * 32 * 48].view(self._torch.int32).view(-1, 12).cpu().numpy()
        rec = rec[rec[:, 3] != -1]
        out = np.zeros(F, dtype=np.uint8)
        out[rec[:, 3].astype(np.int64)] = (rec[:, 7] & 1).astype(np.uint8)
        return out

    def _ensure_workspace(self, W, H):
        if self._workspace_res != (W, H):
            nbytes = ctypes.c_size_t(0)
            _lib.check(_lib.lib.smesh_raster_workspace_bytes(self._V, self._F, W, H, ctypes.byref(nbytes)))
            self._workspace = self._torch.zeros(nbytes.value, dtype=self._torch.uint8, device=self.device)
            self._workspace_res = (W, H)
        return self._workspace

    def render(self, camera, capsule=False, count_into=None, depth=True, out_indices=None):
        """-> (primitive_indices, depth): torch tensors on the GPU, shapes (W, H). depth=False (extension): the depth image
        is not written and None is returned for it (a caller that only fuses the view saves 4 bytes per pixel of traffic).
        out_indices (extension): a contiguous int32 (W, H) device tensor to write the index image into (a slice of a
        batch buffer for `MeshAggregator.add_batch`).

        primitive_indices is int32 holding the reference's uint32 bit pattern (background 0xFFFFFFFF reads as -1; torch
        has few uint32 ops) - `MeshAggregator.add` takes it as is. With capsule=True both are returned as DLPack
        capsules named "dltensor", exactly what the reference returns (Renderer.h:37-41), for
        `tf.experimental.dlpack.from_dlpack` style consumers.

        count_into=aggregator (extension): the pass that writes the index image also leaves the per-face pixel counts
        of this view in the aggregator's counters, and `aggregator.add(primitive_indices, probs)` of THIS index image then
        skips its count pass. The aggregator must be over the faces of this mesh; between such a render and its add at
        most one other counted render may be issued (two counter arrays).
        """
        if not isinstance(camera, Camera):
            raise TypeError("render expects a semantic_meshes.data.Camera")
        torch = self._torch
        W, H = camera.resolution
        if W < 1 or H < 1:
            raise ValueError("render: empty resolution")
        epoch = 0
        if count_into is not None:
            if count_into.primitives != self._F or count_into.device != self.device:
                raise ValueError("render(count_into=...): the aggregator must hold one row per face of this mesh, on this device")
            if W * H < (1 << 24) and self._F > 0 and not capsule:
                epoch = count_into._next_epochs(W * H)
        R, t, f, c = camera._pointers()
        with _lib.on_device(torch, self._dev_index):
            ws = self._ensure_workspace(W, H)
            if out_indices is None:
                idx = torch.empty((W, H), dtype=torch.int32, device=self.device)
            else:
                idx = out_indices
                if (idx.dtype != torch.int32 or tuple(idx.shape) != (W, H) or not idx.is_contiguous() or idx.device != self.device
                        or idx.data_ptr() % 8 != 0):
                    raise ValueError("render(out_indices=...): expected a contiguous, 8-byte aligned int32 (W, H) tensor on this device")
            depth = torch.empty((W, H), dtype=torch.float32, device=self.device) if (depth or capsule) else None
            depth_ptr = depth.data_ptr() if depth is not None else None
            stream = _lib.raw_stream(torch, self._dev_index)
            if epoch != 0:
                rc = _lib.lib.smesh_raster_render_counted(self._mesh_ptr, self._mesh_bytes, self._V, self._F, R, t, f, c, W, H,
                                                          ws.data_ptr(), ws.numel(), idx.data_ptr(), depth_ptr,
                                                          count_into._counts_for(epoch).data_ptr(), epoch, stream)
            else:
                rc = _lib.lib.smesh_raster_render(self._mesh_ptr, self._mesh_bytes, self._V, self._F, R, t, f, c, W, H,
                                                  ws.data_ptr(), ws.numel(), idx.data_ptr(), depth_ptr, stream)
        _lib.check(rc)
        if epoch != 0:
            idx._smesh_counted = (id(count_into), count_into._epoch_gen, epoch)
        if capsule:
            from torch.utils.dlpack import to_dlpack
            return to_dlpack(idx.view(torch.uint32)), to_dlpack(depth)
        return idx, depth


def triangles(ply, device=None):
    """render.triangles(ply) (Render.cu:24, python/semantic_meshes/include/Ply.h:121-124)."""
    return TriangleRenderer(ply, device=device)


class TexturedTriangleRenderer(TriangleRenderer):
    """render.texels(...): primitives are the texels of per-triangle textures
    (include/semantic_meshes/render/TexturedTriangleRenderer.h). The constructor sizes every triangle's texture from its
    largest projection over `cameras` and reorders each face's corners, on the host like the reference's; `render` is
    the triangle rasterizer with the texel-index shader."""

    def __init__(self, ply, cameras, texels_per_pixel=0.1, device=None):
        torch = _lib.require_cuda()
        if not isinstance(ply, Ply):
            raise TypeError("render.texels expects a semantic_meshes.data.Ply")
        cameras = list(cameras.getCameras()) if hasattr(cameras, "getCameras") else list(cameras)
        for cam in cameras:
            if not isinstance(cam, Camera):
                raise TypeError("render.texels expects data.Camera objects or a data.Colmap workspace")
        verts = np.ascontiguousarray(ply.vertices, dtype=np.float32)
        faces = np.array(ply.faces, dtype=np.int32, order="C", copy=True)
        F, n = faces.shape[0], len(cameras)
        R = np.ascontiguousarray([c.rotation for c in cameras], dtype=np.float32).reshape(n, 9)
        t = np.ascontiguousarray([c.translation for c in cameras], dtype=np.float32).reshape(n, 3)
        f = np.ascontiguousarray([c.focal_lengths for c in cameras], dtype=np.float64).reshape(n, 2)
        c = np.ascontiguousarray([c.principal_point for c in cameras], dtype=np.float64).reshape(n, 2)
        res = np.ascontiguousarray([c.resolution for c in cameras], dtype=np.int32).reshape(n, 2)
        tri_res = np.zeros(max(F, 1), dtype=np.uint32)
        first = np.zeros(max(F, 1), dtype=np.uint32)
        total = ctypes.c_uint64(0)
        _lib.check(_lib.lib.smesh_texels_prepare(verts.ctypes.data, verts.shape[0], faces.ctypes.data, F, n, R.ctypes.data,
                                                 t.ctypes.data, f.ctypes.data, c.ctypes.data, res.ctypes.data,
                                                 float(texels_per_pixel), tri_res.ctypes.data, first.ctypes.data,
                                                 ctypes.byref(total)))
        self._texels = int(total.value)
        self.faces = faces            # reordered like the reference reorders ply->getTinyplyFaces() (:133-150)
        self.triangle_resolutions = tri_res[:F]
        super().__init__(Ply.from_arrays(verts, faces), device=device)
        self._tri_res = torch.from_numpy(tri_res.view(np.int32)).to(self.device)
        self._first_texel = torch.from_numpy(first.view(np.int32)).to(self.device)

    def getPrimitivesNum(self):
        return self._texels

    def render(self, camera, capsule=False):
        """-> (texel_indices, depth), shapes (W, H); index 0xFFFFFFFF (-1) where nothing is hit."""
        if not isinstance(camera, Camera):
            raise TypeError("render expects a semantic_meshes.data.Camera")
        torch = self._torch
        W, H = camera.resolution
        if W < 1 or H < 1:
            raise ValueError("render: empty resolution")
        R, t, f, c = camera._pointers()
        with _lib.on_device(torch, self._dev_index):
            ws = self._ensure_workspace(W, H)
            idx = torch.empty((W, H), dtype=torch.int32, device=self.device)
            depth = torch.empty((W, H), dtype=torch.float32, device=self.device)
            rc = _lib.lib.smesh_raster_render_texels(self._mesh_ptr, self._mesh_bytes, self._V, self._F,
                                                     self._tri_res.data_ptr(), self._first_texel.data_ptr(), R, t, f, c, W, H,
                                                     ws.data_ptr(), ws.numel(), idx.data_ptr(), depth.data_ptr(),
                                                     _lib.raw_stream(torch, self._dev_index))
        _lib.check(rc)
        if capsule:
            from torch.utils.dlpack import to_dlpack
            return to_dlpack(idx.view(torch.uint32)), to_dlpack(depth)
        return idx, depth


def texels(ply, cameras, texels_per_pixel=0.1, device=None):
    """render.texels(ply, colmap | [cameras][, texels_per_pixel]) (Render.cu:20-23,
    python/semantic_meshes/include/Ply.h:54-119)."""
    return TexturedTriangleRenderer(ply, cameras, texels_per_pixel, device=device)
