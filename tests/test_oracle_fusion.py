"""CPU: the oracle's label fusion against the reference's known answers (no GPU, no /root/reference needed)."""
import json
import os

import numpy as np
import pytest

import oracle

from conftest import GOLDEN


def kat_inputs(kat):
    W, H, C = kat["W"], kat["H"], kat["C"]
    ids = np.array(kat["ids"], dtype=np.uint32).reshape(W, H)
    raw = np.array([[float((i * 7 + c * 5) % 11 + 1) for c in range(C)] for i in range(W * H)], dtype=np.float32)
    s = (raw[:, 0] + raw[:, 1]) + raw[:, 2]
    probs = (raw / s[:, None]).astype(np.float32)
    probs[7, :] = 0
    return ids, probs.reshape(W, H, C)


@pytest.mark.parametrize("kind", ["sum", "summax", "mul"])
@pytest.mark.parametrize("iew", [0.5, 0.0, 1.0])
def test_known_answer_vectors(kind, iew):
    kat = json.load(open(os.path.join(GOLDEN, "fusion_kat.json")))
    ids, probs = kat_inputs(kat)
    agg = oracle.Aggregator(kat["P"], kat["C"], kind, iew)
    agg.add(ids, probs)
    agg.add(ids, probs)
    exp = np.array(kat["expected"][f"{kind}_{iew}"], dtype=np.float32).reshape(kat["P"], kat["C"])
    np.testing.assert_allclose(agg.get(), exp, rtol=1e-5, atol=1e-9)


def golden_cases():
    data = np.load(os.path.join(GOLDEN, "fusion_ref.npz"))
    return data, [str(c) for c in data["cases"]]


@pytest.mark.parametrize("case", golden_cases()[1])
def test_golden_from_genuine_reference(case):
    """fusion_ref.npz was produced by the genuine ModelAggregator (tests/golden/make_fusion_golden.py, one thread =
    flat pixel order): the restatement reproduces it bit for bit, all three aggregators."""
    data, _ = golden_cases()
    kind, C, iew = case.split("_")
    C, iew = int(C[1:]), float(iew[3:])
    exp = data[f"{case}_get"]
    agg = oracle.Aggregator(exp.shape[0], C, kind, iew)
    for v in range(3):
        w = data[f"{case}_weights{v}"] if f"{case}_weights{v}" in data else None
        agg.add(data[f"{case}_ids{v}"], data[f"{case}_probs{v}"], w)
    assert np.array_equal(agg.get().view(np.uint32), exp.view(np.uint32))


@pytest.mark.skipif(not os.path.exists(oracle.ref_fusion_path()), reason="genuine reference build absent")
@pytest.mark.parametrize("kind", ["sum", "summax", "mul"])
def test_against_live_reference(kind):
    """Multi-threaded genuine reference (its own pixel order is unspecified) vs oracle on fresh random input."""
    rng = np.random.default_rng(5)
    W, H, C, P = 48, 40, 19, 300
    o, r = oracle.Aggregator(P, C, kind, 0.5), oracle.RefAggregator(P, C, kind, 0.5)
    for v in range(2):
        base = rng.integers(0, P, size=(W // 2, H // 2))
        ids = np.repeat(np.repeat(base, 2, 0), 2, 1).astype(np.uint32)
        ids[rng.random((W, H)) < 0.1] = 0xFFFFFFFF
        probs = rng.dirichlet(np.ones(C), size=(W, H)).astype(np.float32)
        probs[rng.random((W, H)) < 0.03] = 0
        o.add(ids, probs)
        r.add(ids, probs)
    # the reference's own run-to-run spread for `mul` is ~1e-4 (sums of logs, then exp)
    np.testing.assert_allclose(o.get(), r.get(), rtol=2e-4 if kind == "mul" else 1e-5, atol=1e-7)


def test_quirks():
    """Untouched face: zeros for sum/summax, uniform for mul; a don't-care pixel still counts in n[id]; ids >= P and the
    background constant are ignored; images_equal_weight mixes 1/n and 1."""
    C, P = 4, 3
    ids = np.array([[0, 0, 5, 0xFFFFFFFF]], dtype=np.uint32)
    probs = np.zeros((1, 4, C), dtype=np.float32)
    probs[0, 0] = [0.7, 0.1, 0.1, 0.1]
    probs[0, 1] = 0          # gate fails, but n[0] = 2
    probs[0, 2] = 0.25
    probs[0, 3] = 0.25
    a = oracle.Aggregator(P, C, "sum", 0.5)
    a.add(ids, probs)
    w = np.float32(0.5) * (np.float32(1) / np.float32(2)) + np.float32(0.5)
    np.testing.assert_array_equal(a.acc[0], probs[0, 0] * w)
    assert not a.acc[1:].any()
    assert not a.get()[1:].any()
    m = oracle.Aggregator(P, C, "mul", 0.5)
    m.add(ids, probs)
    np.testing.assert_allclose(m.get()[1], 0.25)


def test_empty_and_stats():
    a = oracle.Aggregator(5, 3)
    a.add(np.zeros((0, 7), dtype=np.uint32), np.zeros((0, 7, 3), dtype=np.float32))
    assert not a.acc.any()
    ids = np.array([[0, 1, 1, 9]], dtype=np.uint32)
    probs = np.full((1, 4, 3), 1 / 3, dtype=np.float32)
    assert oracle.fuse_stats(ids, probs, 5) == (3, 2)
    np.testing.assert_array_equal(oracle.fuse_count(ids, 5), [1, 2, 0, 0, 0])
