"""ctypes binding of libsmesh_b200.so (C ABI: include/smesh.h). There is no CPU or PyTorch fallback: if the library has
not been built the import fails, and every call needs a CUDA device."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMESH_LIB", os.path.join(_HERE, "libsmesh_b200.so"))

OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
KIND = {"sum": 0, "summax": 1, "mul": 2}
ID_U32, ID_I32, ID_U64, ID_I64 = 0, 1, 2, 3

_vp, _i64, _int, _f32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_u32 = ctypes.c_uint32

# every symbol include/smesh.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "smesh_last_error": (ctypes.c_char_p, []),
    "smesh_version": (ctypes.c_char_p, []),
    "smesh_raster_workspace_bytes": (_int, [_i64, _i64, _int, _int, ctypes.POINTER(_sz)]),
    "smesh_raster_mesh_bytes": (_int, [_i64, _i64, ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "smesh_raster_mesh_build": (_int, [_vp, _i64, _vp, _i64, _vp, _sz, _vp, _sz, _vp]),
    "smesh_raster_render": (_int, [_vp, _sz, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz, _vp, _vp, _vp]),
    "smesh_raster_render_counted": (_int, [_vp, _sz, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz, _vp, _vp, _vp, _u32,
                                           _vp]),
    "smesh_texels_prepare": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp,
                                    ctypes.POINTER(ctypes.c_uint64)]),
    "smesh_raster_render_texels": (_int, [_vp, _sz, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz, _vp, _vp,
                                          _vp]),
    "smesh_fuse_padded_classes": (_int, [_int]),
    "smesh_fuse_add": (_int, [_int, _vp, _int, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _int, _i64, _f32, _vp, _u32,
                              _vp, _vp, _vp]),
    "smesh_fuse_count": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _i64, _vp, _u32, _vp, _vp]),
    "smesh_fuse_scatter": (_int, [_int, _vp, _vp, _vp, _i64, _int, _i64, _f32, _vp, _u32, _vp, _vp]),
    "smesh_fuse_scatter_count_next": (_int, [_int, _vp, _vp, _vp, _i64, _int, _i64, _f32, _vp, _u32, _int, _vp, _i64, _vp, _u32,
                                             _vp, _vp]),
    "smesh_fuse_clear": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "smesh_fuse_add_batch": (_int, [_int, _i64, _vp, _int, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64,
                                    _int, _i64, _f32, _vp, _u32, _vp, _vp, _vp]),
    "smesh_fuse_get": (_int, [_int, _vp, _i64, _int, _vp, _vp]),
    "smesh_fuse_labels": (_int, [_vp, _i64, _int, _f32, _vp, _vp]),
    "smesh_fuse_render": (_int, [_vp, _i64, _int, _vp, _i64, _vp, _vp, _vp]),
    "smesh_selftest_inv_sqrt": (_int, [_u32, ctypes.c_uint64, _vp, _vp]),
    "smesh_pipeline_views": (_int, [_vp, _sz, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz, _int, _vp, _vp, _int,
                                    _vp, _vp, _int, _i64, _f32, _vp, _u32, _vp, _vp]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' or "
        "semantic-meshes_b200/csrc/build.sh). semantic_meshes has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc):
    """Translate an smesh_status into the exception the reference binding would raise (Boost.Python maps
    std::invalid_argument to ValueError and every other std::exception to RuntimeError)."""
    if rc == OK:
        return
    msg = lib.smesh_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise RuntimeError(msg)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(torch, index):
    """Context in which CUDA device `index` is current: nothing to do (and nothing paid) when it already is."""
    return _NO_GUARD if torch.cuda.current_device() == index else torch.cuda.device(index)


def raw_stream(torch, index):
    """cudaStream_t (as an int) of the current torch stream on device `index`, without building a Stream object."""
    try:
        return torch._C._cuda_getCurrentRawStream(index)
    except AttributeError:  # pragma: no cover
        return torch.cuda.current_stream(index).cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("semantic_meshes needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
