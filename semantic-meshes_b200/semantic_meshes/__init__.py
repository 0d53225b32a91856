"""semantic_meshes - B200-native drop-in for the hot paths of fferflo/semantic-meshes.

Same module names and call signatures as the reference package (python/semantic_meshes/__init__.py:1-4):
    data.Ply, data.Colmap, data.Camera, render.triangles(mesh).render(camera), fusion.MeshAggregator(...).add/get/reset
Device buffers are torch tensors; the arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI
(include/smesh.h, libsmesh_b200.so). There is no CPU fallback.
"""
from . import data
from . import fusion
from . import render
from . import pipeline

__all__ = ["data", "fusion", "render", "pipeline"]
