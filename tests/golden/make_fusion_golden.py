"""Generates tests/golden/fusion_ref.npz by running the GENUINE reference aggregator (oracle/_ref/libref_fusion.so, built
from /root/reference by oracle/Makefile) on seeded inputs. Run in the build container (the reference is not on the GPU
box):  OMP_NUM_THREADS=1 python tests/golden/make_fusion_golden.py
One thread makes the reference's pixel order (flat index order) and therefore its float sums deterministic."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

assert os.environ.get("OMP_NUM_THREADS") == "1", "run with OMP_NUM_THREADS=1"


def make_view(rng, W, H, C, P, block):
    bx, by = (W + block - 1) // block, (H + block - 1) // block
    base = rng.integers(0, P, size=(bx, by), dtype=np.int64)
    ids = np.repeat(np.repeat(base, block, 0), block, 1)[:W, :H].astype(np.uint32)
    ids[rng.random((W, H)) < 0.1] = 0xFFFFFFFF
    ids[rng.random((W, H)) < 0.02] = P + 3  # out-of-range but not the background constant
    logits = rng.normal(size=(W, H, C)).astype(np.float32) * 3
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    probs[rng.random((W, H)) < 0.05] = 0            # don't-care pixels
    probs[rng.random((W, H, C)) < 0.02] = 0         # exact zeros inside vectors (mul: +inf absorbing)
    return ids, probs


out = {}
cases = []
rng = np.random.default_rng(20240607)
for C in (3, 19, 40):
    for kind in ("sum", "summax", "mul"):
        for iew in ((0.5, 0.0, 1.0) if C == 3 else (0.5, 0.0)):
            W, H, P = 16, 12, 40
            name = f"{kind}_C{C}_iew{iew}"
            ref = oracle.RefAggregator(P, C, kind, iew)
            views = []
            for v in range(3):
                ids, probs = make_view(rng, W, H, C, P, block=1 + v)
                wts = rng.random((W, H)).astype(np.float32) * 2 if v == 2 else None
                ref.add(ids, probs, wts)
                out[f"{name}_ids{v}"] = ids
                out[f"{name}_probs{v}"] = probs
                if wts is not None:
                    out[f"{name}_weights{v}"] = wts
            out[f"{name}_get"] = ref.get()
            cases.append(name)
out["cases"] = np.array(cases)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fusion_ref.npz"), **out)
print("wrote", len(cases), "cases")
