"""Seeded synthetic meshes, cameras and predictions of the benchmark shapes (BASELINE.json `configs`).
Used by tests/ and bench.py; not part of the reference API."""
import math

import numpy as np

from .data import Camera, Ply


def icosphere(level=3, radius=1.0):
    """Subdivided icosahedron: level 3 -> 642 vertices, 1280 triangles (config 1)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
             (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
             (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    return (np.array(verts, dtype=np.float64) * radius).astype(np.float32), np.array(faces, dtype=np.int32)


def terrain(n_triangles, seed=1234, cell=1.0):
    """'Delaunay-like' height field: jittered sqrt(F/2) x sqrt(F/2) grid, every cell split along a random diagonal,
    random winding (the reference does not cull back faces). Returns (verts float32 (V,3), faces int32 (F,3))."""
    rng = np.random.default_rng(seed)
    n = max(1, int(round(math.sqrt(n_triangles / 2.0))))
    gx, gy = np.meshgrid(np.arange(n + 1, dtype=np.float64), np.arange(n + 1, dtype=np.float64), indexing="ij")
    x = (gx + rng.uniform(-0.3, 0.3, gx.shape)) * cell
    y = (gy + rng.uniform(-0.3, 0.3, gy.shape)) * cell
    L = n * cell
    z = np.zeros_like(x)
    for k in range(4):
        fx, fy = rng.uniform(1, 6, 2) * 2 * math.pi / L
        z += rng.uniform(0.01, 0.03) * L / (k + 1) * np.sin(fx * x + rng.uniform(0, 6.28)) * np.cos(fy * y + rng.uniform(0, 6.28))
    z += rng.normal(0, 0.15 * cell, z.shape)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (i * (n + 1) + j).ravel()
    v10, v01, v11 = v00 + (n + 1), v00 + 1, v00 + (n + 1) + 1
    diag = rng.random(v00.shape) < 0.5
    t0 = np.where(diag[:, None], np.stack([v00, v10, v11], 1), np.stack([v00, v10, v01], 1))
    t1 = np.where(diag[:, None], np.stack([v00, v11, v01], 1), np.stack([v10, v11, v01], 1))
    faces = np.stack([t0, t1], axis=1).reshape(-1, 3)
    flip = rng.random(faces.shape[0]) < 0.5
    faces[flip] = faces[flip][:, ::-1]
    return verts, faces.astype(np.int32)


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """World->camera rotation (rows = camera x, y, z axes; z forward, y down) and translation."""
    eye, target, up = (np.asarray(a, dtype=np.float64) for a in (eye, target, up))
    zc = target - eye
    zc /= np.linalg.norm(zc)
    xc = np.cross(zc, up)
    if np.linalg.norm(xc) < 1e-9:
        xc = np.cross(zc, np.array([0.0, 1.0, 0.0]))
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    R = np.stack([xc, yc, zc], axis=0)
    return R, -R @ eye


def orbit_cameras(n_views, W, H, center, distance, seed=0, tilt_deg=(0.0, 35.0), focal_scale=0.9):
    """Cameras on a seeded orbit around `center` looking at it from `distance`, pitched tilt_deg off the vertical."""
    rng = np.random.default_rng(seed)
    cams = []
    center = np.asarray(center, dtype=np.float64)
    for v in range(n_views):
        yaw = rng.uniform(0, 2 * math.pi)
        tilt = math.radians(rng.uniform(*tilt_deg))
        d = np.array([math.sin(tilt) * math.cos(yaw), math.sin(tilt) * math.sin(yaw), math.cos(tilt)])
        eye = center + d * distance
        up = (math.cos(yaw + 1.0), math.sin(yaw + 1.0), 0.0)
        R, t = look_at(eye, center, up)
        cams.append(Camera(R, t, np.array([W, H]), np.array([focal_scale * W, focal_scale * W]),
                           np.array([W / 2.0, H / 2.0])))
    return cams


def terrain_cameras(n_views, W, H, n_triangles, tris_per_view, seed=0, cell=1.0, focal_scale=0.9):
    """Cameras over the terrain() mesh, each seeing roughly tris_per_view triangles: random look-at points on the
    terrain, height chosen from the footprint, up to 30 degrees off nadir."""
    rng = np.random.default_rng(seed)
    n = max(1, int(round(math.sqrt(n_triangles / 2.0))))
    L = n * cell
    # footprint (world units) of a nadir view from height h: (W/f) h by (H/f) h; cells seen = footprint / cell^2
    cells = max(tris_per_view / 2.0, 1.0)
    h = math.sqrt(cells * cell * cell * (focal_scale * W) ** 2 / (W * H))
    cams = []
    for v in range(n_views):
        margin = min(0.45 * L, 0.6 * max(W, H) / (focal_scale * W) * h)
        target = np.array([rng.uniform(margin, L - margin), rng.uniform(margin, L - margin), 0.0])
        yaw = rng.uniform(0, 2 * math.pi)
        tilt = math.radians(rng.uniform(0.0, 30.0))
        d = np.array([math.sin(tilt) * math.cos(yaw), math.sin(tilt) * math.sin(yaw), math.cos(tilt)])
        eye = target + d * h
        up = (math.cos(yaw + 1.0), math.sin(yaw + 1.0), 0.0)
        R, t = look_at(eye, target, up)
        cams.append(Camera(R, t, np.array([W, H]), np.array([focal_scale * W, focal_scale * W]),
                           np.array([W / 2.0, H / 2.0])))
    return cams


def predictions_numpy(W, H, C, seed, dont_care=0.02):
    """(W, H, C) float32 softmax of 3*N(0,1) logits, a `dont_care` fraction of pixels zeroed (fails the 0.5 gate)."""
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((W, H, C), dtype=np.float32) * np.float32(3.0))
    logits -= logits.max(axis=-1, keepdims=True)
    e = np.exp(logits)
    p = (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)
    p[rng.random((W, H)) < dont_care] = 0
    return p


def predictions_torch(W, H, C, seed, device, dont_care=0.02, out=None):
    """Same distribution as predictions_numpy, generated on the device (values differ from the numpy generator)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    logits = torch.randn((W, H, C), generator=g, device=device, dtype=torch.float32) * 3.0
    p = torch.softmax(logits, dim=-1)
    mask = torch.rand((W, H), generator=g, device=device) < dont_care
    p[mask] = 0
    if out is not None:
        out.copy_(p)
        return out
    return p


def mesh(kind, n_triangles=None, seed=1234):
    verts, faces = icosphere(3) if kind == "icosphere" else terrain(n_triangles, seed)
    return Ply.from_arrays(verts, faces)
