"""CPU: the texel renderer's constructor (TexturedTriangleRenderer.h:86-176). The product does it on the host like the
reference (smesh_texels_prepare in libsmesh_b200.so, no GPU needed); it must agree with the oracle's restatement exactly,
and both with what the constructor's definition implies."""
import ctypes

import numpy as np


def prepare_both(mesh, cams, tpp):
    import oracle
    from semantic_meshes import _lib
    of, ores, ofirst, ototal = oracle.texels_prepare(mesh.vertices, mesh.faces, cams, tpp)
    verts = np.ascontiguousarray(mesh.vertices, dtype=np.float32)
    faces = np.array(mesh.faces, dtype=np.int32, copy=True)
    R, t, f, c, res = oracle._camera_block(cams)
    F = faces.shape[0]
    tri, first, tot = np.zeros(F, np.uint32), np.zeros(F, np.uint32), ctypes.c_uint64(0)
    rc = _lib.lib.smesh_texels_prepare(verts.ctypes.data, verts.shape[0], faces.ctypes.data, F, len(cams), R.ctypes.data,
                                       t.ctypes.data, f.ctypes.data, c.ctypes.data, res.ctypes.data, tpp, tri.ctypes.data,
                                       first.ctypes.data, ctypes.byref(tot))
    assert rc == 0
    return (of, ores, ofirst, ototal), (faces, tri, first, int(tot.value))


def test_constructor_matches_oracle_and_definition():
    from semantic_meshes import synthetic
    cases = [(synthetic.mesh("icosphere"), synthetic.orbit_cameras(5, 320, 240, (0, 0, 0), 3.0, seed=1, tilt_deg=(0, 180)), 0.5),
             (synthetic.mesh("terrain", 20000, seed=3), synthetic.terrain_cameras(6, 320, 240, 20000, 5000, seed=5), 0.3),
             (synthetic.mesh("terrain", 800, seed=9), [], 0.1)]      # no camera: every triangle gets resolution 0
    for mesh, cams, tpp in cases:
        (of, ores, ofirst, ototal), (pf, pres, pfirst, ptotal) = prepare_both(mesh, cams, tpp)
        assert np.array_equal(of, pf) and np.array_equal(ores, pres) and np.array_equal(ofirst, pfirst) and ototal == ptotal
        # definition: texels per triangle r (r + 1) / 2, first texel = running sum, faces only permuted
        r = ores.astype(np.int64)
        assert ototal == int((r * (r + 1) // 2).sum())
        assert np.array_equal(ofirst.astype(np.int64), np.concatenate([[0], np.cumsum(r * (r + 1) // 2)[:-1]]))
        assert np.array_equal(np.sort(of, 1), np.sort(mesh.faces, 1))
        if not cams:
            assert ototal == 0
        # the first corner is the one whose angle is closest to 90 degrees
        v = mesh.vertices.astype(np.float64)
        def ang(f, j):
            a, b = v[f[:, (j + 1) % 3]] - v[f[:, j]], v[f[:, (j + 2) % 3]] - v[f[:, j]]
            cosv = (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
            return np.abs(np.arccos(np.clip(cosv, -1, 1)) - np.pi / 2)
        d = np.stack([ang(of, j) for j in range(3)], 1)
        assert (d[:, 0] <= d[:, 1] + 1e-5).all() and (d[:, 0] <= d[:, 2] + 1e-5).all() and (d[:, 1] <= d[:, 2] + 1e-5).all()


def test_texel_shader_against_its_definition():
    """TexturedTriangle::getTexelIndex (TexturedTriangleRenderer.h:32-41) on one fronto-parallel right triangle with a
    resolution of 8: for every pixel safely inside a texel, the oracle's index must be the lower-triangular index of
    trunc(uv * 8) computed here in double from the pixel's barycentric coordinates."""
    import oracle
    from semantic_meshes.data import Camera
    # corner 0 at the right angle: u runs along edge 0->1, v along edge 0->2 (uv = bc1 * (1,0) + bc2 * (0,1))
    verts = np.array([[-1.0, -1.0, 2.0], [1.0, -1.0, 2.0], [-1.0, 1.0, 2.0]], dtype=np.float32)
    faces = np.array([[0, 1, 2]], dtype=np.int32)
    W = H = 240
    f, c = 100.0, 120.0
    cam = Camera(np.eye(3), np.zeros(3), np.array([W, H]), np.array([f, f]), np.array([c, c]))
    res, first = np.array([8], dtype=np.uint32), np.array([5], dtype=np.uint32)
    idx, depth = oracle.texels_render(verts, faces, res, first, cam.rotation, cam.translation, cam.focal_lengths,
                                      cam.principal_point, W, H)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    # pixel (x, y) sees the plane z = 2 at ((x - c) / f * 2, (y - c) / f * 2): u = (X + 1) / 2, v = (Y + 1) / 2
    u = ((xs - c) / f * 2 + 1) / 2
    v = ((ys - c) / f * 2 + 1) / 2
    inside = (u > 0.01) & (v > 0.01) & (u + v < 0.99)
    safe = inside & (np.abs(u * 8 - np.round(u * 8)) > 0.02) & (np.abs(v * 8 - np.round(v * 8)) > 0.02)
    row, col = np.floor(u * 8).astype(np.int64), np.floor(v * 8).astype(np.int64)
    expect = 5 + np.where(row >= col, (row + 1) * row // 2 + col, (col + 1) * col // 2 + row)
    assert safe.sum() > 3000
    assert np.array_equal(idx[safe].astype(np.int64), expect[safe])
    assert (idx[inside] != 0xFFFFFFFF).all() and np.allclose(depth[inside], 2.0, rtol=1e-6)
    hit = idx != 0xFFFFFFFF
    assert set(np.unique(idx[hit]) - 5) <= set(range(8 * 9 // 2 + 8))      # inside the triangle's texel range (+ diagonal)
