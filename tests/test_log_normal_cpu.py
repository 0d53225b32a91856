"""CPU: the logarithm the `mul` kernels use for normal positive probabilities (log_normal(), semantic-meshes_b200/csrc/
smesh_fuse.cu), re-stated operation for operation in numpy float32 (an FMA = the double-precision product and sum of
float32 operands rounded once more to float32) and compared with log in double: every float32 of [2/3, 4/3) - the range of
its polynomial - and a million floats over all normal exponents. The coefficients are read from the CUDA source."""
import os
import re

import numpy as np

from conftest import ROOT


def shipped_coefficients():
    src = open(os.path.join(ROOT, "semantic-meshes_b200", "csrc", "smesh_fuse.cu")).read()
    body = src[src.index("float log_normal(float x)"):]
    body = body[:body.index("return fmaf(fe")]
    first = re.search(r"float q = (-?[0-9.]+)f;", body).group(1)
    rest = re.findall(r"q = fmaf\(q, f, (-?[0-9.]+)f\);", body)
    assert len(rest) == 7
    return [np.float32(first)] + [np.float32(v) for v in rest]


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def log_normal(x):
    c = shipped_coefficients()
    i = x.view(np.int32).astype(np.int64)
    e = (i - 0x3f2aaaab) & 0xff800000
    e = np.where(e >= 2 ** 31, e - 2 ** 32, e)          # the & of two's-complement ints
    m = (i - e).astype(np.int32).view(np.float32)
    f = (m - np.float32(1.0)).astype(np.float32)
    fe = (e >> 23).astype(np.float32)
    q = np.full_like(f, c[0])
    for k in c[1:]:
        q = fma(q, f, np.full_like(f, k))
    s = (f * f).astype(np.float32)
    r = fma(fma(q, f, np.full_like(f, np.float32(-0.5))), s, f)
    return fma(fe, np.full_like(f, np.float32(0.693147182)), r)


def ulp_error(got, x):
    true = np.log(x.astype(np.float64))
    ulp = np.spacing(np.abs(true).astype(np.float32)).astype(np.float64)
    return np.where(true == 0, np.abs(got), np.abs(got.astype(np.float64) - true) / ulp)


def test_every_mantissa_of_the_polynomial_range():
    lo, hi = np.float32(2 / 3).view(np.int32), np.float32(4 / 3).view(np.int32)
    x = np.arange(int(lo), int(hi), dtype=np.int64).astype(np.int32).view(np.float32)
    err = ulp_error(log_normal(x), x)
    assert err.max() < 1.0, err.max()


def test_all_normal_exponents():
    rng = np.random.default_rng(1)
    bits = rng.integers(0x00800000, 0x7F800000, size=1_000_000, dtype=np.int64).astype(np.int32)
    x = np.concatenate([bits.view(np.float32), np.array([1.17549435e-38, 3.4028235e38, 1.0, 0.5, 2.0, 1e-30, 0.999999, 1.000001],
                                                        dtype=np.float32)])
    err = ulp_error(log_normal(x), x)
    assert err.max() < 1.5, (err.max(), x[err.argmax()])
