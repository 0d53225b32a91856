"""CPU: properties of the oracle rasterizer + (once generated on a B200) the genuine reference kernel's golden images."""
import os

import numpy as np
import pytest

import oracle
from semantic_meshes import synthetic
from semantic_meshes.data import Camera

from conftest import GOLDEN

BG = 0xFFFFFFFF


def cam_args(cam):
    return cam.rotation, cam.translation, cam.focal_lengths, cam.principal_point


def test_single_triangle_plane_depth():
    """A big triangle in the plane z = 2 facing the camera: every covered pixel has depth 2 * uz / uz = 2 (+- rounding),
    index 0; the rest is background with depth +inf."""
    verts = np.array([[-10, -10, 2], [10, -10, 2], [0, 10, 2]], dtype=np.float32)
    faces = np.array([[0, 1, 2]], dtype=np.int32)
    W, H = 32, 24
    cam = Camera(np.eye(3), np.zeros(3), np.array([W, H]), np.array([20.0, 20.0]), np.array([W / 2, H / 2]))
    idx, depth = oracle.raster_render(verts, faces, *cam_args(cam), W, H)
    assert (idx == 0).all()
    np.testing.assert_allclose(depth, 2.0, rtol=1e-6)
    # behind the camera: culled (Triangle.h:107-110)
    verts[:, 2] = -2
    idx, depth = oracle.raster_render(verts, faces, *cam_args(cam), W, H)
    assert (idx == BG).all() and np.isinf(depth).all()


def test_nearest_wins_and_tie_rule():
    """Two coplanar copies of the same triangle: identical z at every pixel -> the lower index wins (contract for the
    reference's order-dependent tie); a nearer triangle wins regardless of index."""
    tri = np.array([[-10, -10, 3], [10, -10, 3], [0, 10, 3]], dtype=np.float32)
    near = tri.copy()
    near[:, 2] = 1.5
    verts = np.concatenate([tri, tri, near])
    W, H = 16, 16
    cam = Camera(np.eye(3), np.zeros(3), np.array([W, H]), np.array([10.0, 10.0]), np.array([8.0, 8.0]))
    idx, _ = oracle.raster_render(verts, np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32), *cam_args(cam), W, H)
    assert set(np.unique(idx)) <= {0, BG} and (idx == 0).any()
    idx, depth = oracle.raster_render(verts, np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]], dtype=np.int32), *cam_args(cam), W, H)
    assert (idx[depth < 2] == 2).all() and (idx == 2).any()


def test_box_coverage_like_reference_test():
    """extern/template-tensors/test/geometry/TestRender.h:87-133 renders a 12-triangle box from several angles at 640x480
    and asserts >= 5 % coverage; same scene, same bound."""
    c = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    faces = np.array([t for q in quads for t in ((q[0], q[1], q[2]), (q[0], q[2], q[3]))], dtype=np.int32)
    W, H = 640, 480
    for cam in synthetic.orbit_cameras(4, W, H, center=(0, 0, 0), distance=5.0, seed=3, tilt_deg=(20, 160)):
        idx, depth = oracle.raster_render(c, faces, *cam_args(cam), W, H)
        cov = (idx != BG).mean()
        assert cov >= 0.05
        assert np.isfinite(depth[idx != BG]).all() and np.isinf(depth[idx == BG]).all()
        assert idx[idx != BG].max() < 12


def test_icosphere_config1_properties():
    verts, faces = synthetic.icosphere(3)
    assert verts.shape == (642, 3) and faces.shape == (1280, 3)
    W = H = 256
    cam = synthetic.orbit_cameras(1, W, H, center=(0, 0, 0), distance=3.0, seed=1, tilt_deg=(0, 180))[0]
    idx, depth = oracle.raster_render(verts, faces, *cam_args(cam), W, H)
    hit = idx != BG
    assert 0.2 < hit.mean() < 0.9
    # a unit sphere seen from distance 3: nearest depth ~2, silhouette depth ~ sqrt(8)*cos(...) < 3
    assert 1.9 < depth[hit].min() < 2.1 and depth[hit].max() < 3.0
    # idempotent / deterministic
    idx2, depth2 = oracle.raster_render(verts, faces, *cam_args(cam), W, H)
    assert np.array_equal(idx, idx2) and np.array_equal(depth.view(np.uint32), depth2.view(np.uint32))


def test_partially_behind_camera_does_not_crash():
    """H3: a triangle with one vertex behind the camera projects to garbage; cvt.rzi saturation keeps the box on screen."""
    verts = np.array([[-1, -1, 2], [1, -1, 2], [0, 1, -0.5]], dtype=np.float32)
    faces = np.array([[0, 1, 2]], dtype=np.int32)
    W, H = 40, 30
    cam = Camera(np.eye(3), np.zeros(3), np.array([W, H]), np.array([30.0, 30.0]), np.array([20.0, 15.0]))
    idx, depth = oracle.raster_render(verts, faces, *cam_args(cam), W, H)
    assert idx.shape == (W, H)
    assert oracle.raster_candidates(verts, faces, *cam_args(cam), W, H) > 0


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "raster_ref.npz")), reason="raster golden not generated yet")
def test_golden_from_genuine_reference_kernel():
    """raster_ref.npz holds index/depth images rendered by the GENUINE reference CUDA kernel on a B200
    (tests/golden/make_raster_golden.py). Bit-exact except where two triangles have exactly the same depth (the
    reference's winner is order-dependent there; the contract picks the lowest index)."""
    data = np.load(os.path.join(GOLDEN, "raster_ref.npz"))
    for name in [str(n) for n in data["cases"]]:
        verts, faces = data[f"{name}_verts"], data[f"{name}_faces"]
        W, H = (int(v) for v in data[f"{name}_res"])
        idx, depth = oracle.raster_render(verts, faces, data[f"{name}_R"], data[f"{name}_t"], data[f"{name}_f"],
                                          data[f"{name}_c"], W, H)
        ref_idx, ref_depth = data[f"{name}_idx"], data[f"{name}_depth"]
        assert np.array_equal(depth.view(np.uint32), ref_depth.view(np.uint32)), name
        diff = idx != ref_idx
        assert diff.mean() <= 1e-3, name  # exact-depth ties only
        if diff.any():
            assert (idx[diff] < ref_idx[diff]).all(), name
