"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the CPU oracle (oracle/liboracle.so, a plain-C restatement of the
reference hot paths, see smesh_oracle.c) and of the genuine reference builds under oracle/_ref/ (see Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product package (semantic-meshes_b200/semantic_meshes) never does; it fails loudly without its CUDA library.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = {"sum": 0, "summax": 1, "mul": 2}

_c_void_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int


def build(ref=False):
    """Compile oracle/liboracle.so (and, if ref=True and /root/reference exists, oracle/_ref/*.so)."""
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    if ref and os.path.isdir(os.environ.get("SMESH_REFERENCE", "/root/reference")):
        subprocess.run(["make", "-C", _HERE, "ref_fusion"], check=True, capture_output=True)
        subprocess.run(["make", "-C", _HERE, "ref_raster"], check=True, capture_output=True)
        subprocess.run(["make", "-C", _HERE, "ref_texels"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_raster_render.restype = _int
        L.oracle_raster_render.argtypes = [_c_void_p, _i64, _c_void_p, _i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                           _int, _int, _c_void_p, _c_void_p]
        L.oracle_raster_candidates.restype = _i64
        L.oracle_raster_candidates.argtypes = [_c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                               _int, _int]
        L.oracle_fuse_count.restype = None
        L.oracle_fuse_count.argtypes = [_c_void_p, _i64, _i64, _c_void_p]
        L.oracle_fuse_add.restype = None
        L.oracle_fuse_add.argtypes = [_int, _c_void_p, _c_void_p, _c_void_p, _i64, _int, _i64, ctypes.c_float, _c_void_p]
        L.oracle_fuse_get.restype = None
        L.oracle_fuse_get.argtypes = [_int, _c_void_p, _i64, _int, _c_void_p]
        L.oracle_fuse_stats.restype = None
        L.oracle_fuse_stats.argtypes = [_c_void_p, _c_void_p, _i64, _int, _i64, _c_void_p, _c_void_p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def camera_arrays(rotation, translation, focal_lengths, principal_point):
    """Round a camera the way the reference binding does (python/semantic_meshes/include/Camera.h:16-57): everything to
    float32 first, intrinsics then widened to double."""
    R = _f32(rotation).reshape(3, 3)
    t = _f32(translation).reshape(3)
    f = _f32(focal_lengths).reshape(2).astype(np.float64)
    c = _f32(principal_point).reshape(2).astype(np.float64)
    return R, t, f, c


def raster_render(verts, faces, R, t, f, c, W, H):
    """-> (idx uint32 [W,H], depth float32 [W,H]); R,t float32, f,c float64 (see camera_arrays)."""
    verts = _f32(verts).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    R = _f32(R).reshape(9)
    t = _f32(t).reshape(3)
    f = np.ascontiguousarray(f, dtype=np.float64).reshape(2)
    c = np.ascontiguousarray(c, dtype=np.float64).reshape(2)
    idx = np.empty((W, H), dtype=np.uint32)
    depth = np.empty((W, H), dtype=np.float32)
    rc = lib().oracle_raster_render(verts.ctypes.data, verts.shape[0], faces.ctypes.data, faces.shape[0], R.ctypes.data,
                                    t.ctypes.data, f.ctypes.data, c.ctypes.data, W, H, idx.ctypes.data, depth.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_raster_render failed")
    return idx, depth


def raster_candidates(verts, faces, R, t, f, c, W, H):
    verts = _f32(verts).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    R = _f32(R).reshape(9)
    t = _f32(t).reshape(3)
    f = np.ascontiguousarray(f, dtype=np.float64).reshape(2)
    c = np.ascontiguousarray(c, dtype=np.float64).reshape(2)
    return int(lib().oracle_raster_candidates(verts.ctypes.data, faces.ctypes.data, faces.shape[0], R.ctypes.data,
                                              t.ctypes.data, f.ctypes.data, c.ctypes.data, W, H))


class Aggregator:
    """CPU oracle twin of fusion.MeshAggregator (same constructor arguments, add/get/reset)."""

    def __init__(self, primitives, classes, aggregator="sum", images_equal_weight=0.5):
        self.P, self.C = int(primitives), int(classes)
        self.kind = KINDS[aggregator.lower()]
        self.iew = float(images_equal_weight)
        self.acc = np.zeros((self.P, self.C), dtype=np.float32)

    def reset(self):
        self.acc[:] = 0

    def add(self, ids, probs, weights=None):
        ids = np.ascontiguousarray(np.asarray(ids).astype(np.uint32, copy=False))
        probs = _f32(probs)
        assert probs.shape == ids.shape + (self.C,), (probs.shape, ids.shape)
        w_ptr = None
        if weights is not None:
            weights = _f32(weights)
            assert weights.shape == ids.shape
            w_ptr = weights.ctypes.data
        lib().oracle_fuse_add(self.kind, ids.ctypes.data, probs.ctypes.data, w_ptr, ids.size, self.C, self.P, self.iew,
                              self.acc.ctypes.data)

    def get(self):
        out = np.empty_like(self.acc)
        lib().oracle_fuse_get(self.kind, self.acc.ctypes.data, self.P, self.C, out.ctypes.data)
        return out


def fuse_count(ids, P):
    ids = np.ascontiguousarray(np.asarray(ids).astype(np.uint32, copy=False))
    counts = np.empty(P, dtype=np.uint32)
    lib().oracle_fuse_count(ids.ctypes.data, ids.size, P, counts.ctypes.data)
    return counts


def fuse_stats(ids, probs, P):
    """-> (accepted pixels, faces touched) of one view (the units of the roofline / scatter metrics)."""
    ids = np.ascontiguousarray(np.asarray(ids).astype(np.uint32, copy=False))
    probs = _f32(probs)
    a, t = _i64(0), _i64(0)
    lib().oracle_fuse_stats(ids.ctypes.data, probs.ctypes.data, ids.size, probs.shape[-1], P, ctypes.byref(a),
                            ctypes.byref(t))
    return a.value, t.value


def write_plain_ply(path, verts, faces):
    """Binary PLY with only what the reference loader accepts (src/data/Ply.cpp:9-15): float xyz vertices and one
    `list uchar int vertex_indices` face property. Used to hand synthetic meshes to the genuine reference renderer."""
    verts = np.ascontiguousarray(verts, dtype="<f4")
    faces = np.ascontiguousarray(faces, dtype="<i4")
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
              "property float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n"
              % (verts.shape[0], faces.shape[0]))
    rec = np.empty(faces.shape[0], dtype=np.dtype([("n", "u1"), ("v", "<i4", (3,))]))
    rec["n"] = 3
    rec["v"] = faces
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(verts.tobytes())
        fh.write(rec.tobytes())


# ---------------------------------------------------------------------------------------------------------------------
# Genuine reference builds (oracle/_ref). Present in the build container and shipped prebuilt to the GPU box.
# ---------------------------------------------------------------------------------------------------------------------

def ref_fusion_path():
    return os.path.join(_HERE, "_ref", "libref_fusion.so")


def ref_raster_path():
    return os.path.join(_HERE, "_ref", "libref_raster.so")


_ref_fusion = None


def ref_fusion_lib():
    global _ref_fusion
    if _ref_fusion is None:
        L = ctypes.CDLL(ref_fusion_path())
        L.ref_fusion_create.restype = _c_void_p
        L.ref_fusion_create.argtypes = [_int, _int, ctypes.c_uint64, ctypes.c_float]
        L.ref_fusion_add.restype = None
        L.ref_fusion_add.argtypes = [_c_void_p, _int, _int, _c_void_p, _c_void_p, _c_void_p]
        L.ref_fusion_get.restype = None
        L.ref_fusion_get.argtypes = [_c_void_p, _c_void_p]
        L.ref_fusion_reset.restype = None
        L.ref_fusion_reset.argtypes = [_c_void_p]
        L.ref_fusion_destroy.restype = None
        L.ref_fusion_destroy.argtypes = [_c_void_p]
        L.ref_fusion_has_classes.restype = _int
        L.ref_fusion_has_classes.argtypes = [_int]
        _ref_fusion = L
    return _ref_fusion


class RefAggregator:
    """The genuine semantic_meshes::ModelAggregator (include/semantic_meshes/fusion/Mesh.h) behind a C driver."""

    def __init__(self, primitives, classes, aggregator="sum", images_equal_weight=0.5):
        L = ref_fusion_lib()
        self.P, self.C = int(primitives), int(classes)
        if not L.ref_fusion_has_classes(self.C):
            raise ValueError(f"reference harness not instantiated for {classes} classes")
        self._L = L
        self._h = L.ref_fusion_create(KINDS[aggregator.lower()], self.C, self.P, float(images_equal_weight))
        if not self._h:
            raise RuntimeError("ref_fusion_create failed")

    def add(self, ids, probs, weights=None):
        ids = np.ascontiguousarray(np.asarray(ids).astype(np.uint32, copy=False))
        probs = _f32(probs)
        W, H = ids.shape
        assert probs.shape == (W, H, self.C)
        w_ptr = None
        if weights is not None:
            weights = _f32(weights)
            w_ptr = weights.ctypes.data
        self._L.ref_fusion_add(self._h, W, H, ids.ctypes.data, probs.ctypes.data, w_ptr)

    def get(self):
        out = np.empty((self.P, self.C), dtype=np.float32)
        self._L.ref_fusion_get(self._h, out.ctypes.data)
        return out

    def reset(self):
        self._L.ref_fusion_reset(self._h)

    def close(self):
        if self._h:
            self._L.ref_fusion_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ref_raster = None


def ref_raster_lib():
    global _ref_raster
    if _ref_raster is None:
        L = ctypes.CDLL(ref_raster_path())
        L.ref_raster_create.restype = _c_void_p
        L.ref_raster_create.argtypes = [ctypes.c_char_p]
        L.ref_raster_primitives.restype = ctypes.c_uint64
        L.ref_raster_primitives.argtypes = [_c_void_p]
        L.ref_raster_render.restype = _int
        L.ref_raster_render.argtypes = [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _int, _int, _c_void_p,
                                        _c_void_p]
        L.ref_raster_destroy.restype = None
        L.ref_raster_destroy.argtypes = [_c_void_p]
        L.ref_raster_last_error.restype = ctypes.c_char_p
        _ref_raster = L
    return _ref_raster


class RefRenderer:
    """The genuine semantic_meshes::render::TriangleRenderer (reference CUDA kernel; needs a GPU) behind a C driver."""

    def __init__(self, ply_path):
        L = ref_raster_lib()
        self._L = L
        self._h = L.ref_raster_create(str(ply_path).encode())
        if not self._h:
            raise RuntimeError("ref_raster_create: " + L.ref_raster_last_error().decode())

    def getPrimitivesNum(self):
        return int(self._L.ref_raster_primitives(self._h))

    def render(self, R, t, f, c, W, H):
        R = _f32(R).reshape(9)
        t = _f32(t).reshape(3)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(2)
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(2)
        idx = np.empty((W, H), dtype=np.uint32)
        depth = np.empty((W, H), dtype=np.float32)
        rc = self._L.ref_raster_render(self._h, R.ctypes.data, t.ctypes.data, f.ctypes.data, c.ctypes.data, W, H,
                                       idx.ctypes.data, depth.ctypes.data)
        if rc != 0:
            raise RuntimeError("ref_raster_render: " + self._L.ref_raster_last_error().decode())
        return idx, depth

    def close(self):
        if self._h:
            self._L.ref_raster_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------------
# texel renderer (TexturedTriangleRenderer): restatement and genuine reference
# ---------------------------------------------------------------------------------------------------------------------

def _camera_block(cameras):
    """cameras: sequence of objects with rotation / translation / focal_lengths / principal_point / resolution (already
    rounded like semantic_meshes.data.Camera does) -> flat arrays R [n,9], t [n,3], f [n,2], c [n,2], res [n,2]."""
    n = len(cameras)
    R = np.ascontiguousarray([_f32(c.rotation).reshape(9) for c in cameras], dtype=np.float32).reshape(n, 9)
    t = np.ascontiguousarray([_f32(c.translation).reshape(3) for c in cameras], dtype=np.float32).reshape(n, 3)
    f = np.ascontiguousarray([np.asarray(c.focal_lengths, dtype=np.float64).reshape(2) for c in cameras]).reshape(n, 2)
    pp = np.ascontiguousarray([np.asarray(c.principal_point, dtype=np.float64).reshape(2) for c in cameras]).reshape(n, 2)
    res = np.ascontiguousarray([np.asarray(c.resolution, dtype=np.int32).reshape(2) for c in cameras], dtype=np.int32).reshape(n, 2)
    return R, t, f, pp, res


def texels_prepare(verts, faces, cameras, texels_per_pixel=0.1):
    """The constructor of TexturedTriangleRenderer -> (reordered faces int32 [F,3], tri_res uint32 [F],
    first_texel uint32 [F], number of texels)."""
    L = lib()
    L.oracle_texels_prepare.restype = ctypes.c_uint64
    L.oracle_texels_prepare.argtypes = [_c_void_p, _c_void_p, _i64, _int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                        _c_void_p, ctypes.c_float, _c_void_p, _c_void_p]
    verts = _f32(verts).reshape(-1, 3)
    faces = np.array(faces, dtype=np.int32, order="C", copy=True).reshape(-1, 3)
    R, t, f, pp, res = _camera_block(cameras)
    F = faces.shape[0]
    tri_res = np.zeros(max(F, 1), dtype=np.uint32)
    first = np.zeros(max(F, 1), dtype=np.uint32)
    total = L.oracle_texels_prepare(verts.ctypes.data, faces.ctypes.data, F, len(cameras), R.ctypes.data, t.ctypes.data,
                                    f.ctypes.data, pp.ctypes.data, res.ctypes.data, float(texels_per_pixel),
                                    tri_res.ctypes.data, first.ctypes.data)
    return faces, tri_res[:F], first[:F], int(total)


def texels_render(verts, faces, tri_res, first_texel, R, t, f, c, W, H):
    """TexturedTriangleRenderer::render with the REORDERED faces -> (texel idx uint32 [W,H], depth float32 [W,H])."""
    L = lib()
    L.oracle_texels_render.restype = _int
    L.oracle_texels_render.argtypes = [_c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                       _c_void_p, _int, _int, _c_void_p, _c_void_p]
    verts = _f32(verts).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    tri_res = np.ascontiguousarray(tri_res, dtype=np.uint32)
    first_texel = np.ascontiguousarray(first_texel, dtype=np.uint32)
    R = _f32(R).reshape(9)
    t = _f32(t).reshape(3)
    f = np.ascontiguousarray(f, dtype=np.float64).reshape(2)
    c = np.ascontiguousarray(c, dtype=np.float64).reshape(2)
    idx = np.empty((W, H), dtype=np.uint32)
    depth = np.empty((W, H), dtype=np.float32)
    rc = L.oracle_texels_render(verts.ctypes.data, faces.ctypes.data, faces.shape[0], tri_res.ctypes.data,
                                first_texel.ctypes.data, R.ctypes.data, t.ctypes.data, f.ctypes.data, c.ctypes.data, W, H,
                                idx.ctypes.data, depth.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_texels_render failed")
    return idx, depth


def ref_texels_path():
    return os.path.join(_HERE, "_ref", "libref_texels.so")


class RefTexelRenderer:
    """The genuine semantic_meshes::render::TexturedTriangleRenderer (host constructor + reference CUDA kernel; needs a
    GPU) behind a C driver (oracle/ref_harness/ref_texels.cu)."""

    def __init__(self, ply_path, cameras, texels_per_pixel=0.1):
        L = ctypes.CDLL(ref_texels_path())
        L.ref_texels_create.restype = _c_void_p
        L.ref_texels_create.argtypes = [ctypes.c_char_p, _int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                        ctypes.c_float]
        L.ref_texels_primitives.restype = ctypes.c_uint64
        L.ref_texels_primitives.argtypes = [_c_void_p]
        L.ref_texels_faces.restype = None
        L.ref_texels_faces.argtypes = [_c_void_p, _c_void_p]
        L.ref_texels_render.restype = _int
        L.ref_texels_render.argtypes = [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _int, _int, _c_void_p, _c_void_p]
        L.ref_texels_destroy.restype = None
        L.ref_texels_destroy.argtypes = [_c_void_p]
        L.ref_texels_last_error.restype = ctypes.c_char_p
        self._L = L
        R, t, f, pp, res = _camera_block(cameras)
        self._h = L.ref_texels_create(str(ply_path).encode(), len(cameras), R.ctypes.data, t.ctypes.data, f.ctypes.data,
                                      pp.ctypes.data, res.ctypes.data, float(texels_per_pixel))
        if not self._h:
            raise RuntimeError("ref_texels_create: " + L.ref_texels_last_error().decode())

    def getPrimitivesNum(self):
        return int(self._L.ref_texels_primitives(self._h))

    def faces(self, F):
        out = np.empty((F, 3), dtype=np.int32)
        self._L.ref_texels_faces(self._h, out.ctypes.data)
        return out

    def render(self, R, t, f, c, W, H):
        R = _f32(R).reshape(9)
        t = _f32(t).reshape(3)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(2)
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(2)
        idx = np.empty((W, H), dtype=np.uint32)
        depth = np.empty((W, H), dtype=np.float32)
        rc = self._L.ref_texels_render(self._h, R.ctypes.data, t.ctypes.data, f.ctypes.data, c.ctypes.data, W, H,
                                       idx.ctypes.data, depth.ctypes.data)
        if rc != 0:
            raise RuntimeError("ref_texels_render: " + self._L.ref_texels_last_error().decode())
        return idx, depth

    def close(self):
        if self._h:
            self._L.ref_texels_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
