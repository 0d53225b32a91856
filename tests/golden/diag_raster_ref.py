"""Diagnostic (GPU box): genuine reference kernel vs oracle on the scenes of test_against_genuine_reference_kernel."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-meshes_b200"), os.path.join(ROOT, "tests")]
import oracle
from conftest import write_plain_ply
from semantic_meshes import synthetic

scenes = [("ico", synthetic.mesh("icosphere"), synthetic.orbit_cameras(3, 256, 256, (0, 0, 0), 3.0, seed=1, tilt_deg=(0, 180)))]
terr = synthetic.mesh("terrain", 20000, seed=77)
scenes.append(("terr", terr, synthetic.terrain_cameras(3, 320, 240, 20000, tris_per_view=5000, seed=5)))
with tempfile.TemporaryDirectory() as tmp:
    for name, mesh, cams in scenes:
        ply = os.path.join(tmp, name + ".ply")
        write_plain_ply(ply, mesh.vertices, mesh.faces)
        ref = oracle.RefRenderer(ply)
        for ci, cam in enumerate(cams):
            W, H = cam.resolution
            args = (cam.rotation, cam.translation, cam.focal_lengths, cam.principal_point, W, H)
            runs = [ref.render(*args) for _ in range(4)]
            o_idx, o_depth = oracle.raster_render(mesh.vertices, mesh.faces, *args)
            for r, (ri, rd) in enumerate(runs):
                dd = rd.view(np.uint32) != o_depth.view(np.uint32)
                di = ri != o_idx
                print(f"{name} cam{ci} run{r}: depth diffs {dd.sum()} idx diffs {di.sum()} (of {W*H}); vs run0: depth {(rd.view(np.uint32) != runs[0][1].view(np.uint32)).sum()} idx {(ri != runs[0][0]).sum()}")
                if r == 0:
                    for (x, y) in list(zip(*np.nonzero(dd | di)))[:8]:
                        print(f"   px ({x},{y}): ref idx {ri[x,y]} z {rd[x,y]!r} bits {rd.view(np.uint32)[x,y]:08x} | oracle idx {o_idx[x,y]} z {o_depth[x,y]!r} bits {o_depth.view(np.uint32)[x,y]:08x}")
        ref.close()
