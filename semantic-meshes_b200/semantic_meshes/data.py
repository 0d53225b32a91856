"""semantic_meshes.data - mesh / camera containers of the reference API (python/semantic_meshes/src/Data.cu:10-19).

Not a hot path: plain numpy on the host. Behaviour follows the reference's loaders:
  Ply     src/data/Ply.cpp:9-15 + tt/interface/tinyply/Tinyply.h (float32 x,y,z vertices; faces = a uchar-counted list of
          int32 triangles); Ply.save python/semantic_meshes/include/Ply.h:17-51
  Colmap  src/data/Colmap.cpp:7-63 + tt/file/colmap/Metadata.h:62-329 (cameras/images .bin or .txt, SIMPLE_PINHOLE and
          PINHOLE only, images sorted by name)
  Camera  python/semantic_meshes/include/Camera.h:16-57
"""
import os
import struct

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1",
    "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


class Camera:
    """Pinhole camera: world->camera rotation (3,3) and translation (3,), resolution (W, H), focal lengths (fx, fy),
    principal point (cx, cy).

    Like the reference binding (Camera.h:19-54) rotation, translation, focal lengths and principal point are rounded to
    float32 first; the intrinsics are then widened to double for the projection arithmetic.
    """

    def __init__(self, rotation, translation, resolution, focal_lengths, principal_point):
        rotation = _as_host_array(rotation, (np.float32, np.float64), 2, "rotation")
        translation = _as_host_array(translation, (np.float32, np.float64), 1, "translation")
        resolution = _as_host_array(resolution, (np.int32, np.uint32, np.int64, np.uint64), 1, "resolution")
        focal_lengths = _as_host_array(focal_lengths, (np.float32, np.float64), 1, "focal_lengths")
        principal_point = _as_host_array(principal_point, (np.float32, np.float64), 1, "principal_point")
        if rotation.shape != (3, 3) or translation.shape != (3,) or resolution.shape != (2,) \
                or focal_lengths.shape != (2,) or principal_point.shape != (2,):
            raise ValueError("Camera: expected rotation (3,3), translation (3,), resolution (2,), focal_lengths (2,), "
                             "principal_point (2,)")
        self.rotation = np.ascontiguousarray(rotation.astype(np.float32))
        self.translation = np.ascontiguousarray(translation.astype(np.float32))
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.focal_lengths = np.ascontiguousarray(focal_lengths.astype(np.float32).astype(np.float64))
        self.principal_point = np.ascontiguousarray(principal_point.astype(np.float32).astype(np.float64))

    @classmethod
    def _from_exact(cls, rotation_f32, translation_f32, resolution, focal_f64, principal_f64):
        """Camera whose intrinsics keep full double precision (what Colmap.getCamera returns, Colmap.h:19-25)."""
        self = cls.__new__(cls)
        self.rotation = np.ascontiguousarray(rotation_f32, dtype=np.float32).reshape(3, 3)
        self.translation = np.ascontiguousarray(translation_f32, dtype=np.float32).reshape(3)
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.focal_lengths = np.ascontiguousarray(focal_f64, dtype=np.float64).reshape(2)
        self.principal_point = np.ascontiguousarray(principal_f64, dtype=np.float64).reshape(2)
        return self

    def _pointers(self):
        """Host addresses of (rotation, translation, focal_lengths, principal_point) for the C ABI; cached while the four
        arrays are the same objects (numpy's .ctypes is slow enough to show in a loop over views)."""
        key = (id(self.rotation), id(self.translation), id(self.focal_lengths), id(self.principal_point))
        cached = self.__dict__.get("_ptr_cache")
        if cached is None or cached[0] != key:
            cached = (key, (self.rotation.ctypes.data, self.translation.ctypes.data, self.focal_lengths.ctypes.data,
                            self.principal_point.ctypes.data))
            self.__dict__["_ptr_cache"] = cached
        return cached[1]

    def __repr__(self):
        return (f"Camera(resolution={self.resolution}, f={self.focal_lengths.tolist()}, "
                f"c={self.principal_point.tolist()})")


def _as_host_array(obj, dtypes, rank, name):
    # Camera.h:19-52 accepts host arrays of the listed element types only (FromTensor<..., mem::HOST>)
    if hasattr(obj, "detach") and hasattr(obj, "cpu"):
        if getattr(obj, "is_cuda", False):
            raise ValueError(f"Camera: {name} must be a host array")
        obj = obj.detach().numpy()
    arr = np.asarray(obj)
    if arr.dtype == object or arr.ndim != rank:
        raise ValueError(f"Camera: {name} must be a rank-{rank} numeric array")
    if arr.dtype not in [np.dtype(d) for d in dtypes]:
        # numpy turns Python lists of ints/floats into int64/float64, which the reference accepts as well
        raise ValueError(f"Camera: {name} has unsupported element type {arr.dtype}")
    return arr


class Ply:
    """Triangle mesh from a PLY file: float32 vertices (V,3) and int32 faces (F,3)."""

    def __init__(self, path):
        self.path = os.fspath(path)
        self.vertices, self.faces = _read_ply(self.path)

    @classmethod
    def from_arrays(cls, vertices, faces):
        """Extension (not in the reference): build a mesh from arrays, e.g. synthetic benchmarks."""
        self = cls.__new__(cls)
        self.path = None
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        return self

    def save(self, path, colors, bin=True):
        """Write vertices, faces and one (red, green, blue) uchar triple per face (Ply.h:17-35)."""
        if hasattr(colors, "detach"):
            colors = colors.detach().cpu().numpy()
        elif hasattr(colors, "numpy") and not isinstance(colors, np.ndarray):
            colors = colors.numpy()
        colors = np.asarray(colors)
        if colors.dtype != np.uint8 or colors.ndim != 2:
            raise ValueError("Ply.save: colors must be a rank-2 uint8 array")  # Common.h:32-40
        if colors.shape != (self.faces.shape[0], 3):
            raise ValueError(f"Ply.save: colors must have shape ({self.faces.shape[0]}, 3), got {colors.shape}")
        _write_ply(os.fspath(path), self.vertices, self.faces, np.ascontiguousarray(colors), bool(bin))


def _read_ply(path):
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise IOError(f"File {path} is not a ply file")
        fmt = None
        elements = []  # [name, count, [(prop name, dtype) | (prop name, count dtype, item dtype)]]
        while True:
            line = fh.readline()
            if not line:
                raise IOError(f"File {path}: unexpected end of header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]]))
                else:
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise IOError(f"File {path}: unsupported ply format {fmt}")
        vertices = faces = None
        endian = ">" if fmt == "binary_big_endian" else "<"
        for name, count, props in elements:
            is_list = any(len(p) == 3 for p in props)
            if name == "vertex":
                vertices = _read_vertex_element(fh, fmt, endian, count, props, path)
            elif name == "face":
                faces = _read_face_element(fh, fmt, endian, count, props, path)
            else:
                _skip_element(fh, fmt, endian, count, props, is_list)
    if vertices is None:
        raise IOError("Failed to request element vertex(x, y, z)")
    if faces is None:
        raise IOError("Element with key 'face' not found in ply file")
    return vertices, faces


def _read_vertex_element(fh, fmt, endian, count, props, path):
    names = [p[0] for p in props]
    if any(len(p) == 3 for p in props):
        raise IOError(f"File {path}: list properties in the vertex element are not supported")
    for key in ("x", "y", "z"):
        if key not in names:
            raise IOError("Failed to request element vertex(x, y, z)")
        if dict((p[0], p[1]) for p in props)[key] != "f4":
            raise IOError("Invalid scalar type of vertex(x, y, z)")  # Tinyply.h:93-97: must be float32
    if fmt == "ascii":
        rows = np.loadtxt(fh, max_rows=count, ndmin=2, dtype=np.float64) if count > 0 else np.zeros((0, len(names)))
        cols = [names.index(k) for k in ("x", "y", "z")]
        return np.ascontiguousarray(rows[:, cols].astype(np.float32))
    dt = np.dtype([(n, endian + t) for n, t in props])
    data = np.frombuffer(fh.read(dt.itemsize * count), dtype=dt, count=count)
    return np.ascontiguousarray(np.stack([data["x"], data["y"], data["z"]], axis=1).astype(np.float32))


def _read_face_element(fh, fmt, endian, count, props, path):
    # The reference requests ALL properties of the face element as one typed block (Tinyply.h:195-230), which only works
    # for the usual single `list uchar int vertex_indices` property; the items must be int32 (Tinyply.h:93-97).
    if len(props) != 1 or len(props[0]) != 3:
        raise IOError("Failed to request element face: expected a single list property of int32 triangles")
    _, count_t, item_t = props[0]
    if item_t != "i4":
        raise IOError("Invalid scalar type of the face list (must be int32)")
    if fmt == "ascii":
        faces = np.empty((count, 3), dtype=np.int32)
        for i in range(count):
            tok = fh.readline().split()
            if int(tok[0]) != 3:
                raise IOError(f"File {path}: face {i} is not a triangle")
            faces[i] = [int(tok[1]), int(tok[2]), int(tok[3])]
        return faces
    dt = np.dtype([("n", endian + count_t), ("v", endian + item_t, (3,))])
    data = np.frombuffer(fh.read(dt.itemsize * count), dtype=dt, count=count)
    if count > 0 and not np.all(data["n"] == 3):
        raise IOError(f"File {path}: only triangle faces are supported")
    return np.ascontiguousarray(data["v"].astype(np.int32))


def _skip_element(fh, fmt, endian, count, props, is_list):
    if fmt == "ascii":
        for _ in range(count):
            fh.readline()
        return
    if not is_list:
        fh.seek(np.dtype([(n, endian + t) for n, t in props]).itemsize * count, os.SEEK_CUR)
        return
    for _ in range(count):
        for p in props:
            if len(p) == 2:
                fh.seek(np.dtype(p[1]).itemsize, os.SEEK_CUR)
            else:
                n = int(np.frombuffer(fh.read(np.dtype(p[1]).itemsize), dtype=endian + p[1])[0])
                fh.seek(np.dtype(p[2]).itemsize * n, os.SEEK_CUR)


def _write_ply(path, vertices, faces, colors, binary):
    V, F = vertices.shape[0], faces.shape[0]
    header = ("ply\nformat {} 1.0\nelement vertex {}\nproperty float x\nproperty float y\nproperty float z\n"
              "element face {}\nproperty list uchar int vertex_indices\nproperty uchar red\nproperty uchar green\n"
              "property uchar blue\nend_header\n").format("binary_little_endian" if binary else "ascii", V, F)
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        if binary:
            fh.write(np.ascontiguousarray(vertices, dtype="<f4").tobytes())
            rec = np.empty(F, dtype=np.dtype([("n", "u1"), ("v", "<i4", (3,)), ("c", "u1", (3,))]))
            rec["n"] = 3
            rec["v"] = faces
            rec["c"] = colors
            fh.write(rec.tobytes())
        else:
            for v in vertices:
                fh.write(("%s %s %s\n" % (repr(float(v[0])), repr(float(v[1])), repr(float(v[2])))).encode("ascii"))
            for f, c in zip(faces, colors):
                fh.write(("3 %d %d %d %d %d %d\n" % (f[0], f[1], f[2], c[0], c[1], c[2])).encode("ascii"))


class Colmap:
    """COLMAP sparse workspace: `cameras.{bin,txt}` + `images.{bin,txt}`; images are ordered by file name."""

    def __init__(self, workspace_path):
        workspace_path = os.fspath(workspace_path)
        self._cameras = _read_colmap_cameras(_pick(workspace_path, "cameras"))
        images = _read_colmap_images(_pick(workspace_path, "images"))
        self._images = sorted(images.values(), key=lambda im: im["name"])  # Colmap.cpp:12-22

    def getImageNum(self):
        return len(self._images)

    def getCamera(self, image):
        """image: index into the name-sorted images, or an image path / file name (Colmap.h(py):15-23)."""
        if isinstance(image, (str, os.PathLike)):
            name = os.path.basename(os.fspath(image).rstrip("/"))
            for im in self._images:
                if im["name"] == name:
                    break
            else:
                # the reference prints and calls exit(-1) here (Colmap.cpp:60-61); a KeyError is the usable equivalent
                raise KeyError(f"Image with name {name} not found in colmap workspace")
        else:
            im = self._images[int(image)]
        cam = self._cameras[im["camera_id"]]
        return Camera._from_exact(im["rotation"], im["translation"], cam["resolution"], cam["f"], cam["c"])

    def getCameras(self):
        return [self.getCamera(i) for i in range(len(self._images))]


def _pick(workspace, stem):
    for ext in (".bin", ".txt"):  # Metadata.h:195-210: .bin first, then .txt
        p = os.path.join(workspace, stem + ext)
        if os.path.exists(p):
            return p
    raise IOError(f"File {os.path.join(workspace, stem)}.* could not be found")


def _camera_from_params(model, params, width, height):
    if model in (0, "SIMPLE_PINHOLE"):
        f = [params[0], params[0]]
        c = [params[1], params[2]]
    elif model in (1, "PINHOLE"):
        f = [params[0], params[1]]
        c = [params[2], params[3]]
    else:
        raise IOError(f"Camera model {model} not supported")  # Metadata.h:112-115
    return {"resolution": (int(width), int(height)), "f": np.array(f, dtype=np.float64),
            "c": np.array(c, dtype=np.float64)}


def _read_colmap_cameras(path):
    cameras = {}
    if path.endswith(".bin"):
        with open(path, "rb") as fh:
            (n,) = struct.unpack("<Q", fh.read(8))
            for _ in range(n):
                cam_id, model, w, h = struct.unpack("<IIQQ", fh.read(24))
                nparams = {0: 3, 1: 4}.get(model)
                if nparams is None:
                    raise IOError(f"Camera model {model} not supported")
                params = struct.unpack("<%dd" % nparams, fh.read(8 * nparams))
                cameras[cam_id] = _camera_from_params(model, params, w, h)
    else:
        with open(path, "r") as fh:
            for line in fh:
                line = line.strip()
                if not line or line[0] == "#":
                    continue
                tok = line.split()
                cameras[int(tok[0])] = _camera_from_params(tok[1], [float(t) for t in tok[4:]], tok[2], tok[3])
    return cameras


def _quat_to_matrix_f32(q):
    """tt/tensor/linear_algebra/Quaternion.h:10-28 evaluated in float32 (w, x, y, z)."""
    w, x, y, z = (np.float32(v) for v in q)
    two, one = np.float32(2), np.float32(1)
    R = np.empty((3, 3), dtype=np.float32)
    R[0, 1] = two * (x * y - w * z)
    R[0, 2] = two * (x * z + w * y)
    R[1, 0] = two * (x * y + w * z)
    R[1, 2] = two * (y * z - w * x)
    R[2, 0] = two * (x * z - w * y)
    R[2, 1] = two * (y * z + w * x)
    R[0, 0] = one - two * (y * y + z * z)
    R[1, 1] = one - two * (x * x + z * z)
    R[2, 2] = one - two * (x * x + y * y)
    return R


def _image_record(image_id, quat, trans, camera_id, name):
    quat = np.asarray(quat, dtype=np.float64).astype(np.float32)  # Metadata.h:233: Quaternion<float> = Quaternion<double>
    if abs(float(np.sqrt(np.sum(quat.astype(np.float64) ** 2))) - 1.0) > 1e-4:
        raise IOError(f"Invalid quaternion {quat.tolist()}")
    return {"id": image_id, "rotation": _quat_to_matrix_f32(quat),
            "translation": np.asarray(trans, dtype=np.float64).astype(np.float32), "camera_id": camera_id, "name": name}


def _read_colmap_images(path):
    images = {}
    if path.endswith(".bin"):
        with open(path, "rb") as fh:
            (n,) = struct.unpack("<Q", fh.read(8))
            for _ in range(n):
                (image_id,) = struct.unpack("<I", fh.read(4))
                quat = struct.unpack("<4d", fh.read(32))
                trans = struct.unpack("<3d", fh.read(24))
                (camera_id,) = struct.unpack("<I", fh.read(4))
                name = b""
                while True:
                    ch = fh.read(1)
                    if ch in (b"\0", b""):
                        break
                    name += ch
                (npts,) = struct.unpack("<Q", fh.read(8))
                fh.seek(24 * npts, os.SEEK_CUR)
                images[image_id] = _image_record(image_id, quat, trans, camera_id, name.decode("utf-8"))
    else:
        with open(path, "r") as fh:
            lines = iter(fh)
            for line in lines:
                line = line.strip()
                if not line or line[0] == "#":
                    continue
                tok = line.split()
                images[int(tok[0])] = _image_record(int(tok[0]), [float(t) for t in tok[1:5]], [float(t) for t in tok[5:8]],
                                                    int(tok[8]), tok[9])
                next(lines, None)  # the 2-D points line
    return images
