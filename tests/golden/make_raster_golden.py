"""Generates tests/golden/raster_ref.npz with the GENUINE reference CUDA kernel (oracle/_ref/libref_raster.so, built from
/root/reference by oracle/Makefile). Needs a GPU: run on the B200 box,
    python tests/golden/make_raster_golden.py gpurun_out/raster_ref.npz
then copy the file to tests/golden/. Each scene is rendered 3 times; pixels whose winner changes between runs (exact
depth ties, order-dependent in the reference) are reported."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-meshes_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import write_plain_ply  # noqa: E402
from semantic_meshes import synthetic  # noqa: E402
from semantic_meshes.data import Camera  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "raster_ref.npz")
scenes = {}
ico = synthetic.mesh("icosphere")
for i, cam in enumerate(synthetic.orbit_cameras(2, 96, 64, (0, 0, 0), 3.0, seed=1, tilt_deg=(0, 180))):
    scenes[f"ico{i}"] = (ico, cam)
terr = synthetic.mesh("terrain", 3000, seed=9)
for i, cam in enumerate(synthetic.terrain_cameras(2, 120, 90, 3000, tris_per_view=800, seed=3)):
    scenes[f"terrain{i}"] = (terr, cam)
# near-plane stress: camera standing on the terrain
R, t = synthetic.look_at(np.array([20.0, 20.0, 1.5]), np.array([21.0, 20.5, 1.3]))
scenes["nearplane"] = (terr, Camera(R, t, np.array([100, 80]), np.array([90.0, 90.0]), np.array([50.0, 40.0])))

out, names = {}, []
with tempfile.TemporaryDirectory() as tmp:
    for name, (mesh, cam) in scenes.items():
        ply = os.path.join(tmp, name + ".ply")
        write_plain_ply(ply, mesh.vertices, mesh.faces)
        ref = oracle.RefRenderer(ply)
        W, H = cam.resolution
        runs = [ref.render(cam.rotation, cam.translation, cam.focal_lengths, cam.principal_point, W, H) for _ in range(3)]
        ref.close()
        unstable = sum(int((runs[0][0] != r[0]).sum()) for r in runs[1:])
        o_idx, o_depth = oracle.raster_render(mesh.vertices, mesh.faces, cam.rotation, cam.translation, cam.focal_lengths,
                                              cam.principal_point, W, H)
        print(f"{name}: {W}x{H}, covered {(runs[0][0] != 0xFFFFFFFF).mean():.3f}, run-to-run index changes {unstable}, "
              f"oracle idx diffs {(o_idx != runs[0][0]).sum()}, oracle depth diffs "
              f"{(o_depth.view(np.uint32) != runs[0][1].view(np.uint32)).sum()}")
        out[f"{name}_verts"], out[f"{name}_faces"] = mesh.vertices, mesh.faces
        out[f"{name}_R"], out[f"{name}_t"] = cam.rotation, cam.translation
        out[f"{name}_f"], out[f"{name}_c"] = cam.focal_lengths, cam.principal_point
        out[f"{name}_res"] = np.array([W, H])
        out[f"{name}_idx"], out[f"{name}_depth"] = runs[0]
        names.append(name)
out["cases"] = np.array(names)
os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
np.savez_compressed(out_path, **out)
print("wrote", out_path)
