"""Generates tests/golden/ply_ref.json: what the GENUINE reference PLY loader (oracle/_ref/libref_ply.so = the reference's
src/data/Ply.cpp + tt/interface/tinyply + vendored tinyply, compiled by `make -C oracle ref_ply`) makes of the PLY
fixtures the reference ships (extern/tinyply/assets/*.ply): loads or rejects, vertex / face counts,
SHA-256 of the float32 vertex array and the int32 face array, and a few sample rows. Run in the build container (needs
/root/reference); tests/test_data.py compares semantic_meshes.data.Ply against it."""
import ctypes
import glob
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ASSETS = "/root/reference/extern/tinyply/assets"


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_ply.so"))
    lib.ref_ply_load.restype = ctypes.c_void_p
    lib.ref_ply_load.argtypes = [ctypes.c_char_p]
    lib.ref_ply_last_error.restype = ctypes.c_char_p
    for name in ("ref_ply_vertices", "ref_ply_faces"):
        getattr(lib, name).restype = ctypes.c_uint64
        getattr(lib, name).argtypes = [ctypes.c_void_p]
    lib.ref_ply_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ref_ply_free.argtypes = [ctypes.c_void_p]
    out = {}
    for path in sorted(glob.glob(os.path.join(ASSETS, "*.ply"))):
        name = os.path.basename(path)
        h = lib.ref_ply_load(path.encode())
        if not h:
            out[name] = {"loads": False, "error": lib.ref_ply_last_error().decode("utf-8", "replace")[:300]}
            continue
        V, F = int(lib.ref_ply_vertices(h)), int(lib.ref_ply_faces(h))
        verts = np.zeros((V, 3), dtype=np.float32)
        faces = np.zeros((F, 3), dtype=np.int32)
        lib.ref_ply_copy(h, verts.ctypes.data, faces.ctypes.data)
        lib.ref_ply_free(h)
        out[name] = {"loads": True, "V": V, "F": F, "verts_sha256": hashlib.sha256(verts.tobytes()).hexdigest(),
                     "faces_sha256": hashlib.sha256(faces.tobytes()).hexdigest(),
                     "verts_head": verts[:3].tolist(), "verts_tail": verts[-2:].tolist(),
                     "faces_head": faces[:3].tolist(), "faces_tail": faces[-2:].tolist()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ply_ref.json"), "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, {a: b for a, b in v.items() if a in ("loads", "V", "F", "error")})


if __name__ == "__main__":
    sys.exit(main())
