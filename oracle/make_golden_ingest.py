"""Generates tests/golden/ingest_cv2.npz with the REAL cv2 chain of the reference (test/datasets.py:50-72) on seeded random
frames — run where cv2 imports (this container: 4.13.0; the reference pins 4.10.0.84).  Small sizes keep the file small."""
import os

import cv2
import numpy as np

CASES = [(48, 64), (64, 48), (37, 100), (40, 40), (45, 80), (5, 3)]   # (H, W)
RES = 64


def reference_chain(frame, res):
    h, w = frame.shape[:2]
    if w > h:
        nw, nh = res, int((h / w) * res)
    else:
        nh, nw = res, int((w / h) * res)
    r = cv2.resize(frame, (nw, nh))
    canvas = cv2.copyMakeBorder(r, top=(res - nh) // 2, bottom=(res - nh + 1) // 2, left=(res - nw) // 2, right=(res - nw + 1) // 2,
                                borderType=cv2.BORDER_CONSTANT, value=(0, 0, 0))
    return np.transpose(cv2.cvtColor(canvas, cv2.COLOR_BGR2RGB), (2, 0, 1))


if __name__ == "__main__":
    rng = np.random.default_rng(20240607)
    out = {"res": np.array(RES), "cv2_version": np.array(cv2.__version__)}
    for i, (h, w) in enumerate(CASES):
        f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out[f"in_{i}"] = f
        out[f"out_{i}"] = reference_chain(f, RES)
    # one real-size case, stored as a checksum only
    f = np.random.default_rng(7).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    out["big_sum"] = np.array(int(reference_chain(f, 384).astype(np.int64).sum()))
    out["big_xor"] = np.array(int(np.bitwise_xor.reduce(reference_chain(f, 384).astype(np.int64).ravel() * (np.arange(3 * 384 * 384) % 251 + 1))))
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ingest_cv2.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes")
