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
