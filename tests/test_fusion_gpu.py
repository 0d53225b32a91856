"""GPU: MeshAggregator (CUDA, through the C ABI) against the CPU oracle on identical inputs.
Tolerances: ids / gates / counts are exact; the float accumulator is compared at 1e-5 relative (north_star) because the
order of float additions differs (atomics) exactly as it does between two runs of the reference itself; `mul` is
compared at 1e-5 in its log-domain accumulator and at 5e-4 after exp()/normalise (the reference's own run-to-run
spread there is 1e-4, see tests/test_oracle_fusion.py)."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

KINDS = ["sum", "summax", "mul"]


@pytest.fixture(scope="module")
def sm():
    import torch
    import semantic_meshes
    assert torch.cuda.is_available()
    return semantic_meshes


def make_view(rng, W, H, C, P, block=2, bg=0.1, oob=0.02, dont_care=0.05, zeros=0.02):
    bx, by = (W + block - 1) // block, (H + block - 1) // block
    base = rng.integers(0, max(P, 1), size=(bx, by), dtype=np.int64)
    ids = np.repeat(np.repeat(base, block, 0), block, 1)[:W, :H].astype(np.uint32)
    ids[rng.random((W, H)) < bg] = 0xFFFFFFFF
    ids[rng.random((W, H)) < oob] = P + 3
    logits = rng.normal(size=(W, H, C)).astype(np.float32) * 3
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    probs[rng.random((W, H)) < dont_care] = 0
    probs[rng.random((W, H, C)) < zeros] = 0
    return ids, probs


def assert_acc_close(kind, got_acc, exp_acc, rtol=1e-5):
    if kind == "mul":
        assert np.array_equal(np.isinf(got_acc), np.isinf(exp_acc))
        fin = ~np.isinf(exp_acc)
        np.testing.assert_allclose(got_acc[fin], exp_acc[fin], rtol=rtol, atol=1e-6)
    else:
        scale = max(float(np.abs(exp_acc).max()), 1e-30)
        np.testing.assert_allclose(got_acc, exp_acc, rtol=rtol, atol=1e-6 * scale)


def assert_get_close(kind, got, exp):
    np.testing.assert_allclose(got, exp, rtol=5e-4 if kind == "mul" else 1e-5, atol=5e-6 if kind == "mul" else 1e-7)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("C", [1, 3, 19, 40, 150])
@pytest.mark.parametrize("iew", [0.5, 0.0])
def test_parity_with_oracle(sm, kind, C, iew):
    import torch
    rng = np.random.default_rng(C * 100 + len(kind))
    W, H, P = 61, 47, 300  # ragged: 2867 pixels, not a multiple of the tile
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind, images_equal_weight=iew)
    ref = oracle.Aggregator(P, C, kind, iew)
    for v in range(3):
        ids, probs = make_view(rng, W, H, C, P, block=1 + v)
        wts = (rng.random((W, H)) * 2).astype(np.float32) if v == 1 else None
        agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(probs).cuda(),
                None if wts is None else torch.from_numpy(wts).cuda())
        ref.add(ids, probs, wts)
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)
    assert_get_close(kind, agg.get(), ref.get())
    agg.reset()
    assert not agg.state().any().item()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("iew", [0.5, 0.0, 1.0])
def test_known_answer_vectors(sm, kind, iew):
    from test_oracle_fusion import kat_inputs
    kat = json.load(open(os.path.join(GOLDEN, "fusion_kat.json")))
    ids, probs = kat_inputs(kat)
    agg = sm.fusion.MeshAggregator(kat["P"], kat["C"], kind, iew)
    agg.add(ids, probs)
    agg.add(ids, probs)
    exp = np.array(kat["expected"][f"{kind}_{iew}"], dtype=np.float32).reshape(kat["P"], kat["C"])
    np.testing.assert_allclose(agg.get(), exp, rtol=1e-5, atol=1e-9)


def test_golden_from_genuine_reference(sm):
    data = np.load(os.path.join(GOLDEN, "fusion_ref.npz"))
    for case in [str(c) for c in data["cases"]]:
        kind, C, iew = case.split("_")
        C, iew = int(C[1:]), float(iew[3:])
        exp = data[f"{case}_get"]
        agg = sm.fusion.MeshAggregator(exp.shape[0], C, kind, iew)
        for v in range(3):
            w = data[f"{case}_weights{v}"] if f"{case}_weights{v}" in data else None
            agg.add(data[f"{case}_ids{v}"], data[f"{case}_probs{v}"], w)
        assert_get_close(kind, agg.get(), exp)


@pytest.mark.parametrize("dtype", ["uint32", "int32", "uint64", "int64"])
def test_id_dtypes_and_background(sm, dtype):
    import torch
    rng = np.random.default_rng(11)
    W, H, C, P = 40, 33, 19, 200
    ids, probs = make_view(rng, W, H, C, P)
    ref = oracle.Aggregator(P, C)
    ref.add(ids, probs)
    wide = ids.astype(np.int64)
    wide[ids == 0xFFFFFFFF] = -1
    if dtype.endswith("64"):
        wide[0, 0] = 2 ** 40 + 5  # must not alias a valid id after truncation
    if dtype.startswith("u"):
        arr = np.where(wide < 0, np.iinfo(dtype).max, wide).astype(dtype)
    else:
        arr = wide.astype(dtype)
    if dtype.endswith("64"):
        ids_ref = ids.copy()
        ids_ref[0, 0] = 0xFFFFFFFF
        ref = oracle.Aggregator(P, C)
        ref.add(ids_ref, probs)
    for src in (arr, torch.from_numpy(arr.view(dtype.replace("u", ""))).view(getattr(torch, dtype)).cuda()):
        agg = sm.fusion.MeshAggregator(P, C)
        agg.add(src, probs)
        assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)


def test_strided_inputs_are_not_copied(sm):
    """Callers pass transpose(pred, (1, 0, 2)) views of (H, W, C) network outputs (colorize_mesh.py:66) and (H, W)
    index images: same result as the contiguous layout."""
    import torch
    rng = np.random.default_rng(3)
    W, H, C, P = 50, 36, 19, 150
    ids, probs = make_view(rng, W, H, C, P)
    wts = rng.random((W, H)).astype(np.float32)
    ref = oracle.Aggregator(P, C)
    ref.add(ids, probs, wts)
    pr_hw = torch.from_numpy(np.ascontiguousarray(probs.transpose(1, 0, 2))).cuda()     # (H, W, C)
    ids_hw = torch.from_numpy(np.ascontiguousarray(ids.view(np.int32).T)).cuda()        # (H, W)
    w_hw = torch.from_numpy(np.ascontiguousarray(wts.T)).cuda()
    for ids_t, w_t in ((ids_hw.t(), w_hw.t()), (ids_hw.t().contiguous(), w_hw.t().contiguous())):
        agg = sm.fusion.MeshAggregator(P, C)
        agg.add(ids_t, pr_hw.permute(1, 0, 2), w_t)
        assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)
    # class-strided probabilities (C, W, H) -> falls back to one contiguous copy, same numbers
    agg = sm.fusion.MeshAggregator(P, C)
    agg.add(ids, torch.from_numpy(np.ascontiguousarray(probs.transpose(2, 0, 1))).cuda().permute(1, 2, 0), wts)
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)
    # numpy transposed view on the host
    agg = sm.fusion.MeshAggregator(P, C)
    agg.add(np.ascontiguousarray(ids.T).T, np.ascontiguousarray(probs.transpose(1, 0, 2)).transpose(1, 0, 2), wts)
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)


def test_dlpack_capsule_input(sm):
    import torch
    from torch.utils.dlpack import to_dlpack
    rng = np.random.default_rng(4)
    W, H, C, P = 20, 30, 3, 50
    ids, probs = make_view(rng, W, H, C, P)
    ref = oracle.Aggregator(P, C)
    ref.add(ids, probs)
    agg = sm.fusion.MeshAggregator(P, C)
    agg.add(to_dlpack(torch.from_numpy(ids.view(np.int32)).cuda()), to_dlpack(torch.from_numpy(probs).cuda()))
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)


def test_edge_cases(sm):
    import torch
    C, P = 19, 10
    agg = sm.fusion.MeshAggregator(P, C)
    agg.add(np.zeros((0, 5), dtype=np.uint32), np.zeros((0, 5, C), dtype=np.float32))          # empty view
    agg.add(np.full((7, 5), 0xFFFFFFFF, dtype=np.uint32), np.full((7, 5, C), 1 / C, dtype=np.float32))  # all background
    agg.add(np.zeros((7, 5), dtype=np.uint32), np.zeros((7, 5, C), dtype=np.float32))          # all don't-care
    assert not agg.state().any().item()
    assert not agg.get().any()
    # a single face collecting every pixel of a view: n = W*H, weight = iew/n + (1-iew)
    W, H = 33, 9
    probs = np.full((W, H, C), 1 / C, dtype=np.float32)
    agg.add(np.full((W, H), 4, dtype=np.uint32), probs)
    ref = oracle.Aggregator(P, C)
    ref.add(np.full((W, H), 4, dtype=np.uint32), probs)
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)
    # no primitives at all
    empty = sm.fusion.MeshAggregator(0, C)
    empty.add(np.zeros((4, 4), dtype=np.uint32), np.zeros((4, 4, C), dtype=np.float32))
    assert empty.get().shape == (0, C)
    # very wide class vector -> direct kernel
    Cw = 1000
    rng = np.random.default_rng(9)
    ids, probs = make_view(rng, 12, 10, Cw, P)
    wide, ref = sm.fusion.MeshAggregator(P, Cw), oracle.Aggregator(P, Cw)
    wide.add(ids, probs)
    ref.add(ids, probs)
    assert_acc_close("sum", wide.state().cpu().numpy(), ref.acc)


def test_error_behaviour(sm):
    """Same exception types as the reference binding."""
    C, P = 3, 10
    agg = sm.fusion.MeshAggregator(P, C)
    ids = np.zeros((4, 5), dtype=np.uint32)
    with pytest.raises(ValueError, match="must have the same width and height"):
        agg.add(ids, np.zeros((5, 4, C), dtype=np.float32))
    with pytest.raises(ValueError, match="must have the same width and height"):
        agg.add(ids, np.zeros((4, 5, C), dtype=np.float32), np.zeros((4, 4), dtype=np.float32))
    with pytest.raises(ValueError, match="None matched"):
        agg.add(ids.astype(np.float32), np.zeros((4, 5, C), dtype=np.float32))
    with pytest.raises(ValueError, match="None matched"):
        agg.add(ids, np.zeros((4, 5, C), dtype=np.float64))
    with pytest.raises(ValueError, match="None matched"):
        agg.add(ids.reshape(-1), np.zeros((4, 5, C), dtype=np.float32))
    with pytest.raises(ValueError, match="None matched"):
        agg.add([[0]], np.zeros((1, 1, C), dtype=np.float32))
    with pytest.raises(RuntimeError):
        sm.fusion.MeshAggregator(P, C, aggregator="median")
    sm.fusion.MeshAggregator(P, C, aggregator="Sum")  # first letter is case-insensitive (Fusion.cu:126)
    sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator="mul", images_equal_weight=0.25)


@pytest.mark.parametrize("kind", ["sum", "mul"])
@pytest.mark.parametrize("shape", [(40, 64), (9, 320), (33, 516), (64, 20), (31, 33)])
def test_pair_kernel_shapes(sm, kind, shape):
    """C = 19 takes the two-pixels-per-lane kernel: tiles that end inside a column, columns shorter than a tile, even and
    odd pixel counts, weights, large faces with runs longer than a warp tile."""
    import torch
    W, H = shape
    C, P = 19, 120
    rng = np.random.default_rng(W * 1000 + H)
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind)
    ref = oracle.Aggregator(P, C, kind)
    for v in range(3):
        ids, probs = make_view(rng, W, H, C, P, block=(1, 3, 7)[v])
        wts = (rng.random((W, H)) * 2).astype(np.float32) if v == 2 else None
        agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(probs).cuda(),
                None if wts is None else torch.from_numpy(wts).cuda())
        ref.add(ids, probs, wts)
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)
    # the transposed layout sweeps along the other image axis
    agg2 = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind)
    ids, probs = make_view(rng, W, H, C, P, block=4)
    ref2 = oracle.Aggregator(P, C, kind)
    ref2.add(ids, probs)
    pr_hw = torch.from_numpy(np.ascontiguousarray(probs.transpose(1, 0, 2))).cuda()
    agg2.add(torch.from_numpy(ids.view(np.int32)).cuda(), pr_hw.permute(1, 0, 2))
    assert_acc_close(kind, agg2.state().cpu().numpy(), ref2.acc)


@pytest.mark.parametrize("kind", ["sum", "mul"])
@pytest.mark.parametrize("C", [2, 3, 4, 5, 7, 8, 12, 13, 16, 17, 18, 20, 21])
def test_narrow_class_vectors(sm, kind, C):
    """Every class count from 2 to 20 has its own instance of the two-pixels-per-lane kernel (odd and even C, one to five
    128-bit chunks per accumulator row, 2 to 8 ring stages); 21 is the first count that takes the per-pixel ring kernel.
    Ragged image (tiles end inside a column), faces of 1 to 36 pixels, weights, gate edge cases."""
    import torch
    W, H, P = 53, 131, 250
    rng = np.random.default_rng(C * 31 + len(kind))
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind)
    ref = oracle.Aggregator(P, C, kind)
    for v in range(3):
        ids, probs = make_view(rng, W, H, C, P, block=(1, 2, 6)[v])
        flat = probs.reshape(-1, C)
        pick = rng.choice(flat.shape[0], flat.shape[0] // 10, replace=False)
        for k, i in enumerate(pick):
            row = np.abs(rng.normal(size=C)).astype(np.float32) + np.float32(1e-3)
            scale = (0.5, 0.5 * (1 - 3e-4), 0.5 * (1 + 3e-4))[k % 3]            # at, just below, just above the gate
            flat[i] = row * (np.float32(scale) / row.sum(dtype=np.float32))
        wts = (rng.random((W, H)) * 2).astype(np.float32) if v == 1 else None
        agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(probs).cuda(),
                None if wts is None else torch.from_numpy(wts).cuda())
        ref.add(ids, probs, wts)
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)
    assert_get_close(kind, agg.get(), ref.get())


@pytest.mark.parametrize("kind", ["sum", "mul"])
@pytest.mark.parametrize("shape", [(40, 64), (23, 128), (50, 260), (17, 1080), (300, 256), (7, 2048), (64, 516)])
def test_tall_column_shapes(sm, kind, shape):
    """C = 19 images with tall columns (64 ... 2048 pixels, not always a multiple of the 256-pixel tile): faces from 1 pixel
    to wider than the image, faces that reappear in every third column, weights, and a face change at every pixel
    (block = 1: as many run sums per warp tile as there are pixels)."""
    import torch
    W, H = shape
    C, P = 19, 400
    rng = np.random.default_rng(W * 977 + H)
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind)
    ref = oracle.Aggregator(P, C, kind)
    for v, block in enumerate((1, 2, 5, 16, 64)):
        ids, probs = make_view(rng, W, H, C, P, block=block, bg=0.05 if block > 1 else 0.3)
        if block == 5:
            ids[::3] = ids[0]                       # the same faces in every third column: a gap, then again
        wts = (rng.random((W, H)) * 2).astype(np.float32) if v % 2 else None
        agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(probs).cuda(),
                None if wts is None else torch.from_numpy(wts).cuda())
        ref.add(ids, probs, wts)
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)
    assert_get_close(kind, agg.get(), ref.get())


def test_add_on_rendered_ids(sm):
    """MeshAggregator.add on real index images (coherent faces of a dozen pixels, occlusion edges) at a size of several
    hundred tiles per CTA, against the oracle."""
    import torch
    from semantic_meshes import synthetic
    W, H, C = 640, 512, 19
    mesh = synthetic.mesh("terrain", 60000, seed=5)
    renderer = sm.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C)
    ref = oracle.Aggregator(P, C)
    for v, cam in enumerate(synthetic.terrain_cameras(3, W, H, 60000, tris_per_view=25000, seed=12)):
        idx, _ = renderer.render(cam)
        pred = synthetic.predictions_torch(W, H, C, seed=50 + v, device="cuda")
        agg.add(idx, pred)
        ref.add(idx.cpu().numpy().view(np.uint32), pred.cpu().numpy())
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc)


@pytest.mark.parametrize("kind", ["sum", "mul"])
@pytest.mark.parametrize("C", [32, 33, 40, 64, 66, 150, 257, 512])
def test_wide_class_vectors(sm, kind, C):
    """C >= 32 takes scatter_rows_kernel (lanes across the classes of a pixel): sub-groups of 8 / 16 / 32 lanes, one to
    four chunks per lane, 128-bit / 64-bit / scalar rows (C % 4, C % 2, odd C, an image that starts 8 or 4 bytes off a
    16-byte boundary), weights, and the gate: rows whose sum sits exactly at, just below and just above 0.5, rows with
    negative entries and rows with a NaN must be accepted / rejected exactly as the sequential float sum decides."""
    import torch
    W, H, P = 37, 29, 300
    rng = np.random.default_rng(C)
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind)
    ref = oracle.Aggregator(P, C, kind)
    for v in range(3):
        ids, probs = make_view(rng, W, H, C, P, block=(1, 3, 6)[v])
        flat = probs.reshape(-1, C)
        # gate edge cases on a tenth of the pixels
        pick = rng.choice(flat.shape[0], flat.shape[0] // 10, replace=False)
        for k, i in enumerate(pick):
            row = np.abs(rng.normal(size=C)).astype(np.float32)
            mode = k % 5
            if mode == 0:
                row *= np.float32(0.5) / row.sum(dtype=np.float32)             # at the threshold (either side by rounding)
            elif mode == 1:
                row *= np.float32(0.5 * (1 - 3e-4)) / row.sum(dtype=np.float32)
            elif mode == 2:
                row *= np.float32(0.5 * (1 + 3e-4)) / row.sum(dtype=np.float32)
            elif mode == 3:
                row[rng.integers(0, C)] *= -1                                  # a negative entry
                row *= np.float32(0.7) / max(abs(float(row.sum())), 1e-3)
            elif kind == "sum":
                row[rng.integers(0, C)] = np.nan                               # NaN: the gate fails, the pixel is skipped
            flat[i] = row
        if kind == "mul":
            np.abs(flat, out=flat)                                             # log of a negative probability is not defined
        wts = (rng.random((W, H)) * 2).astype(np.float32) if v == 1 else None
        pr = torch.from_numpy(probs).cuda()
        if v == 2:                                                             # a view that starts 4 or 8 bytes off
            off = 1 if C % 2 else 2
            buf = torch.empty(probs.size + 4, dtype=torch.float32, device="cuda")
            buf[off:off + probs.size] = pr.reshape(-1)
            pr = buf[off:off + probs.size].view(W, H, C)
        agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), pr, None if wts is None else torch.from_numpy(wts).cuda())
        ref.add(ids, probs, wts)
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)


def test_labels_colors_and_gather_back(sm):
    """what the reference's scripts do after get() - labels / colours per primitive
    (python/scripts/colorize_mesh.py:82-92) and ModelRenderer::render (Mesh.h:24-43), the image of per-primitive
    annotations - against the same few lines of numpy."""
    import torch
    W, H, C, P = 53, 47, 7, 90
    rng = np.random.default_rng(5)
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C)
    ids, probs = make_view(rng, W, H, C, P // 2, block=3)       # the upper half of the primitives is never seen
    agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(probs).cuda())
    dist = agg.get()
    exp_labels = np.where(dist.sum(-1, dtype=np.float32) < 0.9, -1, dist.argmax(-1)).astype(np.int32)
    labels = agg.labels()
    assert labels.dtype == np.int32 and np.array_equal(labels, exp_labels) and (labels == -1).sum() >= P // 2 - 2
    table = rng.integers(1, 255, size=(C, 3)).astype(np.uint8)
    colors = agg.colors(table)
    exp_colors = np.where(exp_labels[:, None] < 0, 0, table[np.maximum(exp_labels, 0)]).astype(np.uint8)
    assert colors.dtype == np.uint8 and np.array_equal(colors, exp_colors)
    valid = ids < P
    for ann, bg in ((exp_labels, -7), (exp_colors, (9, 8, 7)), (dist, np.full(C, 0.25, np.float32)),
                    (rng.integers(0, 200, P).astype(np.uint8), 255)):
        img = agg.render(torch.from_numpy(ids.view(np.int32)).cuda(), ann, bg).cpu().numpy()
        exp = np.empty((W, H) + ann.shape[1:], dtype=ann.dtype)
        exp[valid] = ann[ids[valid]]
        exp[~valid] = np.asarray(bg, dtype=ann.dtype)
        assert img.dtype == ann.dtype and np.array_equal(img, exp)
    img64 = agg.render(torch.from_numpy(ids.astype(np.int64)).cuda(), exp_labels, -7).cpu().numpy()
    ids_wide = ids.astype(np.int64)
    exp = np.where(ids_wide < P, exp_labels[np.minimum(ids_wide, P - 1)], -7)
    assert np.array_equal(img64, exp)


def test_host_predictions_upload_overlapped(sm):
    """Predictions handed to add() in HOST memory (pinned torch tensors, plain numpy arrays, a transposed host view) are
    uploaded on a copy stream into two alternating staging buffers while the previous view is still being fused: many
    views in a row, mixed with device inputs and changing shapes, must give what the device path gives."""
    import torch
    C, P = 19, 500
    rng = np.random.default_rng(77)
    dev_agg, host_agg = sm.fusion.MeshAggregator(P, C), sm.fusion.MeshAggregator(P, C)
    for v in range(9):
        W, H = (96, 80) if v % 4 != 3 else (64, 72)          # a shape change reallocates the staging buffer
        ids, probs = make_view(rng, W, H, C, P, block=3)
        ids_d = torch.from_numpy(ids.view(np.int32)).cuda()
        dev_agg.add(ids_d, torch.from_numpy(probs).cuda())
        if v % 3 == 0:
            host = torch.from_numpy(probs).pin_memory()
        elif v % 3 == 1:
            host = probs                                      # numpy, pageable
        else:
            host = torch.from_numpy(np.ascontiguousarray(probs.transpose(1, 0, 2))).pin_memory().permute(1, 0, 2)
        host_agg.add(ids_d if v % 2 else ids.view(np.int32), host)
        if v == 4:
            host_agg.add(ids_d, torch.from_numpy(probs).cuda())   # a device input in between
            dev_agg.add(ids_d, torch.from_numpy(probs).cuda())
    torch.testing.assert_close(host_agg.state(), dev_agg.state(), rtol=1e-6, atol=1e-7)


def test_count_epoch_wraparound(sm):
    """The per-view pixel counters are tagged with an 8-bit epoch instead of being cleared (include/smesh.h); 600 views
    cross the wrap twice, and the face -> pixel-count mapping changes every view."""
    import torch
    rng = np.random.default_rng(33)
    W, H, C, P = 16, 12, 3, 20
    agg, ref = sm.fusion.MeshAggregator(P, C), oracle.Aggregator(P, C)
    views = [make_view(rng, W, H, C, P, block=1 + (v % 3)) for v in range(12)]
    dev = [(torch.from_numpy(i.view(np.int32)).cuda(), torch.from_numpy(p).cuda()) for i, p in views]
    for v in range(600):
        agg.add(*dev[v % 12])
        ref.add(*views[v % 12])
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc, rtol=2e-5)
    # batches are split at the wrap as well
    ids = torch.stack([d[0] for d in dev] * 30)
    probs = torch.stack([d[1] for d in dev] * 30)
    agg.add_batch(ids, probs)
    for v in range(360):
        ref.add(*views[v % 12])
    assert_acc_close("sum", agg.state().cpu().numpy(), ref.acc, rtol=2e-5)


def test_add_batch_equals_sequential(sm):
    import torch
    rng = np.random.default_rng(21)
    B, W, H, C, P = 4, 32, 24, 19, 90
    views = [make_view(rng, W, H, C, P) for _ in range(B)]
    ids = torch.from_numpy(np.stack([v[0] for v in views]).view(np.int32)).cuda()
    probs = torch.from_numpy(np.stack([v[1] for v in views])).cuda()
    a, b = sm.fusion.MeshAggregator(P, C), sm.fusion.MeshAggregator(P, C)
    a.add_batch(ids, probs)
    for i in range(B):
        b.add(ids[i], probs[i])
    ref = oracle.Aggregator(P, C)
    for i, p in views:
        ref.add(i, p)
    assert_acc_close("sum", a.state().cpu().numpy(), ref.acc)
    assert_acc_close("sum", b.state().cpu().numpy(), ref.acc)


@pytest.mark.parametrize("shape", [(2048, 1024, 19, 2_000_000), (640, 480, 40, 500_000), (1920, 1080, 150, 1_000_000)])
def test_full_size_properties(sm, shape):
    """Configs 3, 2 and 4 at full size (e.g. 2 M primitives, 2048x1024x19; 1920x1080x150 = 1.2 GB of predictions): too big
    for the oracle in seconds, so check size-independent properties against a plain torch fp32 statement of the same op
    (bincount + index_add_), plus linearity."""
    import torch
    W, H, C, P = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    # 4x4-pixel blocks of random faces, 10 % background
    base = torch.randint(0, P, (W // 4, H // 4), generator=g, device="cuda", dtype=torch.int32)
    ids = base.repeat_interleave(4, 0).repeat_interleave(4, 1).contiguous()
    ids[torch.rand((W, H), generator=g, device="cuda") < 0.1] = -1
    probs = torch.softmax(torch.randn((W, H, C), generator=g, device="cuda") * 3, -1)
    probs[torch.rand((W, H), generator=g, device="cuda") < 0.02] = 0
    agg = sm.fusion.MeshAggregator(P, C)
    agg.add(ids, probs)
    acc1 = agg.state().clone()

    flat = ids.reshape(-1).long()
    valid = flat >= 0
    n = torch.bincount(flat[valid], minlength=P).float()
    s = probs.reshape(-1, C).sum(-1)
    ok = valid & (s > 0.5)
    w = 0.5 * (1.0 / n[flat.clamp(min=0)]) + 0.5
    exp = torch.zeros((P, C), device="cuda")
    exp.index_add_(0, flat[ok], probs.reshape(-1, C)[ok] * w[ok, None])
    torch.testing.assert_close(acc1, exp, rtol=1e-5, atol=1e-6)
    # conservation: total mass = sum of accepted pixel weights (each class vector sums to ~1)
    assert abs(acc1.double().sum().item() - (w[ok].double() * s[ok].double()).sum().item()) < 1e-3 * ok.sum().item() ** 0.5
    # linearity: the same view again doubles every row
    agg.add(ids, probs)
    torch.testing.assert_close(agg.state(), 2 * acc1, rtol=1e-6, atol=1e-7)
    # get(): rows sum to 1 where touched, 0 elsewhere
    out = agg.get(device=True)
    rows = out.sum(-1)
    touched = torch.zeros(P, dtype=torch.bool, device="cuda")
    touched[flat[ok]] = True
    assert torch.allclose(rows[touched], torch.ones_like(rows[touched]), atol=1e-5)
    assert not out[~touched].any().item()


# ---------------------------------------------------------------------------------------------------------------------
# round 2: count stage (cross-column run folding), overlapped batches, token validity, get(), mul fast path, full size
# ---------------------------------------------------------------------------------------------------------------------

def _count(sm, ids_t, P, epoch=1, want_ids32=False):
    """smesh_fuse_count through the C ABI on a (W, H) torch tensor of any supported dtype / strides -> counts (P,)"""
    import torch
    from semantic_meshes import _lib
    dt = {torch.uint32: _lib.ID_U32, torch.int32: _lib.ID_I32, torch.uint64: _lib.ID_U64, torch.int64: _lib.ID_I64}[ids_t.dtype]
    W, H = ids_t.shape
    counts = torch.zeros(max(P, 1), dtype=torch.int32, device="cuda")
    ids32 = torch.full((W * H,), 12345, dtype=torch.int32, device="cuda") if want_ids32 else None
    _lib.check(_lib.lib.smesh_fuse_count(ids_t.data_ptr(), dt, ids_t.stride(0), ids_t.stride(1), W, H, P, counts.data_ptr(), epoch,
                                         ids32.data_ptr() if want_ids32 else None, torch.cuda.current_stream().cuda_stream))
    c = counts.cpu().numpy().view(np.uint32)
    if epoch:
        assert ((c >> 24 == epoch) | (c == 0)).all()
        c = c & 0xFFFFFF
    return c, (ids32.cpu().numpy().view(np.uint32) if want_ids32 else None)


@pytest.mark.parametrize("shape", [(1, 1), (3, 70), (70, 3), (8, 32), (9, 33), (64, 64), (61, 47), (200, 517), (517, 200)])
def test_count_stage_exact(sm, shape):
    """The per-face pixel histogram (Mesh.h:90-93) is exact: blobs of every size (faces spanning several 8-column x
    32-row blocks, faces split by an occluder so that two runs of one column see the same run of the next), background,
    out-of-range ids, every id dtype, transposed (strided) index images, tagged and untagged counters."""
    import torch
    W, H = shape
    P = 97
    rng = np.random.default_rng(W * 131 + H)
    for block in (1, 2, 3, 5, 16, 40):
        ids, _ = make_view(rng, W, H, 1, P, block=block, bg=0.15, oob=0.03)
        if block == 5 and W > 2:
            ids[1::3] = ids[0]                     # the same column pattern again and again with gaps
        if block == 3 and H > 8:
            ids[:, H // 2] = (ids[:, H // 2] + 1) % P   # an occluder line splitting every face in two runs
        exp = np.bincount(ids[ids < P].astype(np.int64), minlength=P)
        flat_exp = np.where(ids < P, ids, 0xFFFFFFFF).reshape(-1)
        t32 = torch.from_numpy(ids.view(np.int32)).cuda()
        got, _ = _count(sm, t32, P, epoch=1 + block)
        assert np.array_equal(got, exp), f"block {block}"
        got, _ = _count(sm, t32, P, epoch=0)
        assert np.array_equal(got, exp)
        wide = ids.astype(np.int64)
        wide[ids == 0xFFFFFFFF] = -1
        got, flat = _count(sm, torch.from_numpy(wide).cuda(), P, epoch=3, want_ids32=True)
        assert np.array_equal(got, exp) and np.array_equal(flat, flat_exp)
        # the transposed view: element (x, y) at y*W + x
        tr = torch.from_numpy(np.ascontiguousarray(ids.view(np.int32).T)).cuda().t()
        got, flat = _count(sm, tr, P, epoch=200, want_ids32=True)
        assert np.array_equal(got, exp) and np.array_equal(flat, flat_exp)


def test_count_stage_on_rendered_ids(sm):
    import torch
    from semantic_meshes import synthetic
    W, H = 640, 512
    mesh = synthetic.mesh("terrain", 60000, seed=5)
    renderer = sm.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    for cam in synthetic.terrain_cameras(2, W, H, 60000, tris_per_view=25000, seed=12):
        idx, _ = renderer.render(cam)
        ids = idx.cpu().numpy().view(np.uint32)
        got, _ = _count(sm, idx, P, epoch=7)
        assert np.array_equal(got, np.bincount(ids[ids < P].astype(np.int64), minlength=P))


@pytest.mark.parametrize("lanes", [True, False])
@pytest.mark.parametrize("kind", KINDS)
def test_add_batch_overlapped_and_captured(sm, kind, lanes, monkeypatch):
    """add_batch deals the views to two lanes (the caller's stream and a side stream of the library, one counter array
    each); without lanes (SMESH_NO_BATCH_LANES, or when the side stream would have to be created during a capture) the
    count stage of view b+1 rides in the scatter launch of view b. Either way: same accumulator as the sequential loop,
    also with weights, also when the call is captured into a CUDA graph and replayed, and for odd / even batch sizes."""
    import torch
    if not lanes:
        monkeypatch.setenv("SMESH_NO_BATCH_LANES", "1")
    rng = np.random.default_rng(55)
    W, H, C, P = 96, 130, 19, 400
    for B in (2, 5, 8):
        views = [make_view(rng, W, H, C, P, block=1 + (b % 4)) for b in range(B)]
        ids = torch.from_numpy(np.stack([v[0] for v in views]).view(np.int32)).cuda()
        probs = torch.from_numpy(np.stack([v[1] for v in views])).cuda()
        wts = torch.from_numpy((rng.random((B, W, H)) * 2).astype(np.float32)).cuda() if B == 5 else None
        ref = oracle.Aggregator(P, C, kind)
        for b, (i, p) in enumerate(views):
            ref.add(i, p, None if wts is None else wts[b].cpu().numpy())
        a = sm.fusion.MeshAggregator(P, C, kind)
        a.add_batch(ids, probs, wts)
        assert_acc_close(kind, a.state().cpu().numpy(), ref.acc)
        # captured: every replay adds the batch once more
        g_agg = sm.fusion.MeshAggregator(P, C, kind)
        g_agg.restart_epochs()
        g_agg.add_batch(ids, probs, wts)           # warm-up outside the capture (side stream, kernel attributes)
        torch.cuda.synchronize()
        g_agg.reset()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_agg.restart_epochs()
            g_agg.add_batch(ids, probs, wts)
        graph.replay()
        torch.cuda.synchronize()
        assert_acc_close(kind, g_agg.state().cpu().numpy(), ref.acc)
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
        exp3 = np.where(np.isinf(ref.acc), ref.acc, 3 * ref.acc)
        assert_acc_close(kind, g_agg.state().cpu().numpy(), exp3, rtol=2e-5)


def test_counted_render_token_is_voided_by_later_counts(sm):
    """render(camera, count_into=agg) leaves the view's counts in one of the aggregator's two counter arrays; the index
    image may only skip its count pass while that array still holds them. Two plain add() calls (or an add_batch) in
    between overwrite both arrays: the add of the counted image must then recount - silently using another view's counts
    would give wrong weights."""
    import torch
    from semantic_meshes import synthetic
    W, H, C = 160, 120, 19
    mesh = synthetic.mesh("terrain", 6000, seed=2)
    renderer = sm.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    cams = synthetic.terrain_cameras(3, W, H, 6000, tris_per_view=1500, seed=8)
    preds = [synthetic.predictions_torch(W, H, C, seed=v, device="cuda") for v in range(3)]
    plain = [renderer.render(c)[0] for c in cams]
    for between in ("two_adds", "batch", "nothing"):
        agg, ref = sm.fusion.MeshAggregator(P, C), sm.fusion.MeshAggregator(P, C)
        idx0, _ = renderer.render(cams[0], count_into=agg)
        assert getattr(idx0, "_smesh_counted", None) is not None
        if between == "two_adds":
            agg.add(plain[1], preds[1])
            agg.add(plain[2], preds[2])
        elif between == "batch":
            agg.add_batch(torch.stack(plain[1:]), torch.stack(preds[1:]))
        agg.add(idx0, preds[0])
        if between != "nothing":
            ref.add(plain[1], preds[1])
            ref.add(plain[2], preds[2])
        ref.add(plain[0], preds[0])
        torch.testing.assert_close(agg.state(), ref.state(), rtol=1e-5, atol=1e-7)


def test_fused_count_pipeline_across_epoch_wrap(sm):
    """ViewPipeline(fused_count=True) over more than 255 views: the 8-bit count epoch wraps inside the run, the counters
    are zeroed on the render stream and both streams are ordered around that reset."""
    import torch
    from semantic_meshes import synthetic
    from semantic_meshes.pipeline import ViewPipeline
    W, H, C = 96, 80, 19
    mesh = synthetic.mesh("terrain", 4000, seed=4)
    renderer = sm.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    base = synthetic.terrain_cameras(6, W, H, 4000, tris_per_view=1200, seed=3)
    n = 600
    cams = [base[v % 6] for v in range(n)]
    preds6 = torch.stack([synthetic.predictions_torch(W, H, C, seed=v, device="cuda") for v in range(6)])
    preds = [preds6[v % 6] for v in range(n)]
    seq = sm.fusion.MeshAggregator(P, C)
    for v in range(6):
        idx, _ = renderer.render(base[v])
        seq.add(idx, preds6[v])
    ovl = sm.fusion.MeshAggregator(P, C)
    ViewPipeline(renderer, ovl, fused_count=True).run(cams, preds)
    torch.cuda.synchronize()
    assert torch.isfinite(ovl.state()).all()
    torch.testing.assert_close(ovl.state(), (n // 6) * seq.state(), rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("C", [1, 3, 19, 40, 150, 700, 3000])
def test_get_kernel_shapes(sm, kind, C):
    """get(): rows staged through shared memory in blocks of 256 (fewer for wide class vectors, one thread per row straight
    from global memory beyond ~1500 classes), block boundaries at P = 1, 255, 256, 257, 1000; untouched rows, +inf entries
    (mul) and huge / tiny magnitudes; against the oracle's get() on the same raw accumulator."""
    rng = np.random.default_rng(C)
    for P in (1, 255, 256, 257, 1000):
        acc = np.abs(rng.normal(size=(P, C))).astype(np.float32) * np.float32(10.0) ** rng.integers(-3, 4, size=(P, 1)).astype(np.float32)
        acc[rng.random(P) < 0.2] = 0                      # faces never seen
        if kind == "mul":
            acc[rng.random((P, C)) < 0.05] = np.inf       # absorbing zero probability
            if P > 3:
                acc[3] = np.inf                            # every class impossible
        agg, ref = sm.fusion.MeshAggregator(P, C, kind), oracle.Aggregator(P, C, kind)
        agg.load_state(acc)
        ref.acc[...] = acc
        got = agg.get()
        assert got.shape == (P, C) and np.isfinite(got).all()
        assert_get_close(kind, got, ref.get())


@pytest.mark.parametrize("kind", ["sum", "mul"])
@pytest.mark.parametrize("C", [19, 6, 8, 150])
def test_get_stream_pipeline_many_blocks(sm, kind, C):
    """get() for C <= 200 streams blocks of 32 rows through a two-stage bulk-copy pipeline per warp: enough rows that every
    warp takes several blocks and reuses both stages (and a last block of 1 row), against the oracle's get(); also from
    an accumulator slice that starts at row 1."""
    rng = np.random.default_rng(C + 100)
    P = (400_001 if C < 100 else 60_001)
    acc = np.abs(rng.normal(size=(P, C))).astype(np.float32)
    acc[rng.random(P) < 0.2] = 0
    if kind == "mul":
        acc[rng.random((P, C)) < 0.05] = np.inf
    agg, ref = sm.fusion.MeshAggregator(P, C, kind), oracle.Aggregator(P, C, kind)
    agg.load_state(acc)
    ref.acc[...] = acc
    exp = ref.get()
    assert_get_close(kind, agg.get(), exp)
    part = agg.get(rows=(1, P))
    assert_get_close(kind, part, exp[1:])


def test_mul_direct_form_matches_reference_sequence(sm, monkeypatch):
    """mul accumulates -log(p^w). The kernels use -w log p where p^w stays a normal float and the reference's own
    powf + logf sequence elsewhere (underflow to the absorbing zero, p = 0, w = 0); SMESH_MUL_EXACT=1 forces the reference's
    sequence for every element. Both against the oracle and against each other, with probabilities down to 1e-38, weights
    from 0 to 40 and exact zeros."""
    import torch
    rng = np.random.default_rng(8)
    W, H, P = 64, 48, 150
    for C in (19, 40, 150):
        ids, probs = make_view(rng, W, H, C, P, block=2)
        tiny = rng.random((W, H, C)) < 0.03
        probs[tiny] = np.float32(10.0) ** rng.uniform(-38, -5, size=int(tiny.sum())).astype(np.float32)
        wts = (rng.random((W, H)) * np.where(rng.random((W, H)) < 0.1, 40.0, 2.0)).astype(np.float32)
        wts[rng.random((W, H)) < 0.05] = 0
        ref = oracle.Aggregator(P, C, "mul")
        ref.add(ids, probs, wts)
        states = []
        for exact in ("0", "1"):
            monkeypatch.setenv("SMESH_MUL_EXACT", exact)
            agg = sm.fusion.MeshAggregator(P, C, "mul")
            agg.add(ids, probs, wts)
            states.append(agg.state().cpu().numpy())
            assert_acc_close("mul", states[-1], ref.acc)
            assert_get_close("mul", agg.get(), ref.get())
        assert np.isinf(ref.acc).any()
        assert np.array_equal(np.isinf(states[0]), np.isinf(states[1]))


def test_summax_runs_of_equal_class_merge(sm):
    """summax folds consecutive pixels of one face with the same best class into one reduction: piecewise-constant
    predictions (long runs), ties between classes (first maximum wins), class changes inside a face."""
    import torch
    rng = np.random.default_rng(4)
    W, H, C, P = 40, 200, 19, 60
    ids, probs = make_view(rng, W, H, C, P, block=8)
    blocky = probs[::4, ::10].repeat(4, 0).repeat(10, 1)[:W, :H].copy()
    blocky[5:9, 20:60, 3] = blocky[5:9, 20:60].max(-1)     # ties: class 3 equals the maximum
    agg, ref = sm.fusion.MeshAggregator(P, C, "summax"), oracle.Aggregator(P, C, "summax")
    for pr in (blocky, probs):
        agg.add(ids, pr)
        ref.add(ids, pr)
    assert_acc_close("summax", agg.state().cpu().numpy(), ref.acc)


@pytest.mark.parametrize("name,kinds", [("cfg3", KINDS), ("cfg5", ["sum"])])
def test_full_size_against_genuine_reference_aggregator(sm, name, kinds):
    """FULL-SIZE views (config 3: 2 M faces, 2048x1024x19; config 5: 5 M faces, 1280x720x19) rendered by our rasterizer and
    fused by the GENUINE reference aggregator (oracle/_ref/libref_fusion.so, all host threads) and by ours, all three
    aggregator kinds at config 3: get() within 1e-5 (mul: 5e-4, the reference's own run-to-run spread after exp)."""
    import torch
    import bench
    from semantic_meshes import synthetic
    if not (os.path.exists(oracle.ref_fusion_path()) and oracle.ref_fusion_lib().ref_fusion_has_classes(19)):
        pytest.skip("genuine reference aggregator build absent")
    cfg = bench.CONFIGS[name]
    W, H, C = cfg["W"], cfg["H"], cfg["C"]
    mesh, cams = bench.build_scene(cfg, 0, 2)
    renderer = sm.render.triangles(mesh)
    P = renderer.getPrimitivesNum()
    views = []
    for v, cam in enumerate(cams):
        idx, _ = renderer.render(cam)
        views.append((idx, synthetic.predictions_torch(W, H, C, seed=900 + v, device="cuda")))
    host = [(i.cpu().numpy().view(np.uint32), p.cpu().numpy()) for i, p in views]
    for kind in kinds:
        ours, ref = sm.fusion.MeshAggregator(P, C, kind), oracle.RefAggregator(P, C, kind)
        for (i, p), (hi, hp) in zip(views, host):
            ours.add(i, p)
            ref.add(hi, hp)
        got, exp = ours.get(), ref.get()
        ref.close()
        assert (exp.sum(-1) > 0.5).sum() > 100000
        assert_get_close(kind, got, exp)


@pytest.mark.parametrize("kind,C", [("sum", 19), ("mul", 19), ("summax", 19), ("sum", 40), ("sum", 150), ("sum", 3)])
def test_add_with_count_next_chain(sm, kind, C):
    """add(ids_v, probs_v, count_next=ids_{v+1}): the next view's count stage rides in this view's scatter launch (or is a
    launch of its own where the scatter kernel has no spare warp: summax, wide C), and the next add finds its counts in
    place. A chain of views of changing size, a break in the chain (a plain add in between voids nothing it should not),
    an unused count_next, and more than 255 views (the epoch wraps: the chain restarts) - always the plain loop's result."""
    import torch
    rng = np.random.default_rng(C * 7 + len(kind))
    P = 300
    shapes = [(96, 130), (96, 130), (64, 72), (96, 130), (31, 33), (96, 130)]
    views = [make_view(rng, W, H, C, P, block=1 + (v % 4)) for v, (W, H) in enumerate(shapes)]
    dev = [(torch.from_numpy(i.view(np.int32)).cuda(), torch.from_numpy(p).cuda()) for i, p in views]
    ref = oracle.Aggregator(P, C, kind)
    agg = sm.fusion.MeshAggregator(P, C, kind)
    n = len(dev)
    for v in range(n):
        agg.add(dev[v][0], dev[v][1], count_next=dev[v + 1][0] if v + 1 < n else None)
        ref.add(*views[v])
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)
    # a break in the chain: view 1 is counted ahead, then two other views are added, then view 1
    agg.add(dev[0][0], dev[0][1], count_next=dev[1][0])
    agg.add(dev[2][0], dev[2][1])
    agg.add(dev[3][0], dev[3][1], count_next=dev[4][0])      # dev[4] is counted but never added right away
    agg.add(dev[1][0], dev[1][1])                            # its token is stale by now: recount
    agg.add(dev[4][0], dev[4][1])
    for v in (0, 2, 3, 1, 4):
        ref.add(*views[v])
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc, rtol=2e-5)
    # a long chain across the epoch wrap
    agg.reset()
    ref.reset()
    for v in range(300):
        agg.add(dev[v % n][0], dev[v % n][1], count_next=dev[(v + 1) % n][0])
    for v in range(n):
        ref.add(*views[v])
    exp = ref.acc * 50
    if kind == "mul":
        exp = np.where(np.isinf(ref.acc), ref.acc, exp)
    assert_acc_close(kind, agg.state().cpu().numpy(), exp, rtol=5e-5)


@pytest.mark.parametrize("kind,C", [("summax", 8), ("summax", 16), ("sum", 24), ("mul", 24), ("sum", 32), ("sum", 48), ("summax", 48),
                                    ("sum", 56), ("sum", 44)])
def test_ring_kernel_padded_layouts(sm, kind, C):
    """The ring kernel pads its shared-memory tile for class counts whose 128-bit row loads would collide in the banks
    (C = 8, 24, 40, 56: every 4 pixels; C = 16, 48: every 2): ragged images, tiles that end inside a group, batches with
    the riding count, against the oracle."""
    import torch
    W, H, P = 53, 77, 250
    rng = np.random.default_rng(C * 13 + len(kind))
    agg = sm.fusion.MeshAggregator(primitives=P, classes=C, aggregator=kind)
    ref = oracle.Aggregator(P, C, kind)
    views = []
    for v in range(3):
        ids, probs = make_view(rng, W, H, C, P, block=(1, 2, 6)[v])
        wts = (rng.random((W, H)) * 2).astype(np.float32) if v == 1 else None
        agg.add(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(probs).cuda(),
                None if wts is None else torch.from_numpy(wts).cuda())
        ref.add(ids, probs, wts)
        views.append((ids, probs))
    ids_b = torch.from_numpy(np.stack([v[0] for v in views]).view(np.int32)).cuda()
    probs_b = torch.from_numpy(np.stack([v[1] for v in views])).cuda()
    agg.add_batch(ids_b, probs_b)
    for i, p in views:
        ref.add(i, p)
    assert_acc_close(kind, agg.state().cpu().numpy(), ref.acc)
    assert_get_close(kind, agg.get(), ref.get())
