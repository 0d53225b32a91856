"""CPU: the C-ABI library loads and exports exactly what include/smesh.h declares; argument validation that needs no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "smesh.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smesh_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from semantic_meshes import _lib
    names = declared_symbols()
    assert "smesh_raster_render" in names and "smesh_fuse_add" in names and len(names) >= 8
    for name in names:
        assert hasattr(_lib.lib, name), f"{name} declared in include/smesh.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"


def test_version_and_padding():
    from semantic_meshes import _lib
    assert b"sm_100a" in _lib.lib.smesh_version()
    assert [_lib.lib.smesh_fuse_padded_classes(c) for c in (1, 3, 4, 19, 40, 150)] == [4, 4, 4, 20, 40, 152]


def test_workspace_bytes_and_errors():
    from semantic_meshes import _lib
    n = ctypes.c_size_t(0)
    assert _lib.lib.smesh_raster_workspace_bytes(642, 1280, 256, 256, ctypes.byref(n)) == 0
    assert n.value >= 256 * 256 * 8 + 1280 * 8
    assert _lib.lib.smesh_raster_workspace_bytes(10, 10, 0, 5, ctypes.byref(n)) == _lib.ERR_INVALID_ARGUMENT
    assert b"invalid argument" in _lib.lib.smesh_last_error()
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_INVALID_ARGUMENT)
    # argument checks happen before any CUDA call
    assert _lib.lib.smesh_fuse_add(7, None, 0, 0, 0, None, None, 0, 0, 4, 4, 3, 10, 0.5, None, 1, None, None, None) \
        == _lib.ERR_INVALID_ARGUMENT
    assert _lib.lib.smesh_fuse_get(0, None, 5, 0, None, None) == _lib.ERR_INVALID_ARGUMENT
    assert _lib.lib.smesh_raster_render(None, 0, 0, 0, None, None, None, None, 4, 4, None, 0, None, None, None) \
        == _lib.ERR_INVALID_ARGUMENT
    m, t = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert _lib.lib.smesh_raster_mesh_bytes(642, 1280, ctypes.byref(m), ctypes.byref(t)) == 0
    assert m.value >= 642 * 16 + 1280 * 16 + 10 * 16 and t.value >= 1280 * 24
    assert _lib.lib.smesh_raster_mesh_bytes(-1, 5, ctypes.byref(m), ctypes.byref(t)) == _lib.ERR_INVALID_ARGUMENT
    assert _lib.lib.smesh_raster_mesh_build(None, 3, None, 1, None, 0, None, 0, None) == _lib.ERR_INVALID_ARGUMENT
    # the one-call view loop: null buffers, a ring outside 1..8, count epochs that would leave 1..255
    dummy = ctypes.c_void_p(16)
    def views(B, ring, epoch0, buf):
        return _lib.lib.smesh_pipeline_views(buf, 64, 3, 1, B, buf, buf, buf, buf, 4, 4, buf, 64, ring, buf, None, 0, buf, None,
                                             3, 1, 0.5, buf, epoch0, buf, None)
    assert views(2, 4, 1, None) == _lib.ERR_INVALID_ARGUMENT
    assert views(2, 0, 1, dummy) == _lib.ERR_INVALID_ARGUMENT
    assert views(2, 9, 1, dummy) == _lib.ERR_INVALID_ARGUMENT
    assert views(2, 4, 0, dummy) == _lib.ERR_INVALID_ARGUMENT
    assert views(3, 4, 254, dummy) == _lib.ERR_INVALID_ARGUMENT and b"count epochs" in _lib.lib.smesh_last_error()
    assert views(0, 4, 1, dummy) == _lib.OK                      # an empty list of views is not an error
    assert _lib.lib.smesh_selftest_inv_sqrt(0x7F000000, 1 << 24, dummy, None) == _lib.ERR_INVALID_ARGUMENT
    assert _lib.lib.smesh_selftest_inv_sqrt(0x3F800000, 16, None, None) == _lib.ERR_INVALID_ARGUMENT


def test_no_cpu_fallback():
    import torch
    import semantic_meshes
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        semantic_meshes.fusion.MeshAggregator(primitives=4, classes=3)
    from semantic_meshes import synthetic
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        semantic_meshes.render.triangles(synthetic.mesh("icosphere"))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "semantic-meshes_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "_ref/" not in text, f
