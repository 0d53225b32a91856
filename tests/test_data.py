"""CPU: data.Ply / data.Colmap / data.Camera (the reference's loader semantics, python/semantic_meshes/src/Data.cu)."""
import os
import struct

import numpy as np
import pytest

from semantic_meshes import data, synthetic
from conftest import GOLDEN
from oracle import write_plain_ply


def test_ply_roundtrip_binary_and_ascii(tmp_path):
    verts, faces = synthetic.icosphere(1)
    src = tmp_path / "in.ply"
    write_plain_ply(str(src), verts, faces)
    mesh = data.Ply(str(src))
    assert np.array_equal(mesh.vertices, verts) and np.array_equal(mesh.faces, faces)
    colors = (np.arange(faces.shape[0] * 3) % 251).astype(np.uint8).reshape(-1, 3)
    for binary in (True, False):
        dst = tmp_path / f"out_{binary}.ply"
        mesh.save(str(dst), colors, binary)
        head = open(dst, "rb").read(400).decode("ascii", "replace")
        assert "property list uchar int vertex_indices" in head and "property uchar red" in head
        assert ("binary_little_endian" in head) == binary
        # the saved file has extra face properties, which the reference loader (and ours) rejects as a face element
        with pytest.raises(IOError):
            data.Ply(str(dst))
    # read the binary file back by hand
    raw = open(tmp_path / "out_True.ply", "rb").read()
    body = raw[raw.index(b"end_header\n") + 11:]
    v = np.frombuffer(body[:verts.size * 4], dtype="<f4").reshape(-1, 3)
    assert np.array_equal(v, verts)
    rec = np.frombuffer(body[verts.size * 4:], dtype=np.dtype([("n", "u1"), ("v", "<i4", (3,)), ("c", "u1", (3,))]))
    assert np.array_equal(rec["v"], faces) and np.array_equal(rec["c"], colors) and (rec["n"] == 3).all()


def test_ply_ascii_with_extra_vertex_properties(tmp_path):
    p = tmp_path / "a.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment hi\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
                 "0 0 0 255\n1 0 0 255\n0 1 0.5 255\n3 0 1 2\n")
    mesh = data.Ply(str(p))
    assert mesh.vertices.dtype == np.float32 and mesh.vertices.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0.5]]
    assert mesh.faces.tolist() == [[0, 1, 2]]


def test_ply_type_checks_like_tinyply_adapter(tmp_path):
    # double vertices / uint face indices are rejected (tt/interface/tinyply/Tinyply.h:93-97 "Invalid scalar type")
    p = tmp_path / "d.ply"
    p.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty double x\nproperty double y\nproperty double z\n"
                 "element face 0\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n")
    with pytest.raises(IOError, match="Invalid scalar type"):
        data.Ply(str(p))
    p.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar uint vertex_indices\nend_header\n0 0 0\n3 0 0 0\n")
    with pytest.raises(IOError, match="Invalid scalar type"):
        data.Ply(str(p))
    mesh = data.Ply.from_arrays(np.zeros((3, 3)), [[0, 1, 2]])
    with pytest.raises(ValueError):
        mesh.save(str(tmp_path / "x.ply"), np.zeros((1, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        mesh.save(str(tmp_path / "x.ply"), np.zeros((2, 3), dtype=np.uint8))


def test_camera_rounds_to_float32_then_widens():
    R = np.eye(3) + 1e-9
    cam = data.Camera(R, [0.1, 0.2, 0.3], np.array([640, 480]), [500.123456789, 501.0], [320.00000001, 240.5])
    assert cam.rotation.dtype == np.float32 and cam.translation.dtype == np.float32
    assert cam.focal_lengths.dtype == np.float64 and cam.focal_lengths[0] == float(np.float32(500.123456789))
    assert cam.principal_point[0] == float(np.float32(320.00000001)) and cam.resolution == (640, 480)
    with pytest.raises(ValueError):
        data.Camera(np.eye(3, dtype=np.int32), [0, 0, 0.0], np.array([4, 4]), [1.0, 1.0], [0.0, 0.0])
    with pytest.raises(ValueError):
        data.Camera(np.eye(4), [0.0, 0, 0], np.array([4, 4]), [1.0, 1.0], [0.0, 0.0])


def colmap_workspace(tmp, binary):
    q = np.array([0.9238795, 0.0, 0.3826834, 0.0])  # 45 deg about y, wxyz
    if binary:
        with open(tmp / "cameras.bin", "wb") as fh:
            fh.write(struct.pack("<Q", 2))
            fh.write(struct.pack("<IIQQ", 1, 0, 640, 480) + struct.pack("<3d", 500.5, 320.25, 240.75))
            fh.write(struct.pack("<IIQQ", 2, 1, 800, 600) + struct.pack("<4d", 700.0, 710.0, 400.0, 300.0))
        with open(tmp / "images.bin", "wb") as fh:
            fh.write(struct.pack("<Q", 2))
            for image_id, name, cam in ((1, b"b.png", 2), (2, b"a.png", 1)):
                fh.write(struct.pack("<I", image_id) + struct.pack("<4d", *q) + struct.pack("<3d", 1.0, 2.0, 3.0))
                fh.write(struct.pack("<I", cam) + name + b"\0" + struct.pack("<Q", 1) + struct.pack("<ddQ", 1.0, 2.0, 7))
    else:
        (tmp / "cameras.txt").write_text("# comment\n1 SIMPLE_PINHOLE 640 480 500.5 320.25 240.75\n"
                                         "2 PINHOLE 800 600 700 710 400 300\n")
        (tmp / "images.txt").write_text("# comment\n1 %s 1 2 3 2 b.png\n1.0 2.0 7\n2 %s 1 2 3 1 a.png\n\n"
                                        % (" ".join(map(str, q)), " ".join(map(str, q))))


@pytest.mark.parametrize("binary", [True, False])
def test_colmap_workspace(tmp_path, binary):
    colmap_workspace(tmp_path, binary)
    ws = data.Colmap(str(tmp_path))
    assert ws.getImageNum() == 2
    cam_a = ws.getCamera(0)                    # images are sorted by name: a.png first
    assert cam_a.resolution == (640, 480) and cam_a.focal_lengths.tolist() == [500.5, 500.5]
    assert cam_a.principal_point.tolist() == [320.25, 240.75]
    cam_b = ws.getCamera("some/dir/b.png")
    assert cam_b.resolution == (800, 600) and cam_b.focal_lengths.tolist() == [700.0, 710.0]
    # quaternion -> matrix in float32 (tt/tensor/linear_algebra/Quaternion.h:10-28)
    c, s = np.float32(0.70710677), np.float32(0.70710677)
    np.testing.assert_allclose(cam_a.rotation, [[c, 0, s], [0, 1, 0], [-s, 0, c]], atol=1e-6)
    assert cam_a.rotation.dtype == np.float32 and cam_a.translation.tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(KeyError):
        ws.getCamera("missing.png")


def test_colmap_unsupported_model(tmp_path):
    (tmp_path / "cameras.txt").write_text("1 SIMPLE_RADIAL 640 480 500 320 240 0.1\n")
    (tmp_path / "images.txt").write_text("")
    with pytest.raises(IOError, match="not supported"):
        data.Colmap(str(tmp_path))
    with pytest.raises(IOError, match="could not be found"):
        data.Colmap(str(tmp_path / "nope"))


def test_ply_loader_pinned_to_the_genuine_reference_loader():
    """data.Ply against what the reference's own loader (src/data/Ply.cpp:9-15 + tt/interface/tinyply/Tinyply.h:93-97,
    195-230 + vendored tinyply, compiled into oracle/_ref/libref_ply.so and run by tests/golden/make_ply_golden.py) makes of
    the PLY fixtures the reference ships, extern/tinyply/assets/*.ply: the same files load, the same files are rejected
    (sofa.ply: a face element with more than the list property; elephant.ply: no face element), and a loaded file yields
    bit-identical float32 vertices and int32 faces (SHA-256 of the arrays)."""
    import hashlib
    import json
    assets = "/root/reference/extern/tinyply/assets"
    if not os.path.isdir(assets):
        pytest.skip("reference checkout absent (GPU box)")
    golden = json.load(open(os.path.join(GOLDEN, "ply_ref.json")))
    assert sorted(golden) == sorted(f for f in os.listdir(assets) if f.endswith(".ply"))
    assert sum(1 for g in golden.values() if g["loads"]) >= 4
    for name, g in golden.items():
        path = os.path.join(assets, name)
        if not g["loads"]:
            with pytest.raises((ValueError, RuntimeError, OSError)):
                data.Ply(path)
            continue
        ply = data.Ply(path)
        verts = np.ascontiguousarray(ply.vertices, dtype=np.float32)
        faces = np.ascontiguousarray(ply.faces, dtype=np.int32)
        assert verts.shape == (g["V"], 3) and faces.shape == (g["F"], 3), name
        assert hashlib.sha256(verts.tobytes()).hexdigest() == g["verts_sha256"], name
        assert hashlib.sha256(faces.tobytes()).hexdigest() == g["faces_sha256"], name
