"""Oracle of the frame-ingest row (f3) pinned against cv2: live when cv2 imports, and through the committed fixture."""
import os

import numpy as np
import pytest

from oracle import ingest as I

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ingest_cv2.npz")


def test_oracle_matches_cv2_fixture():
    g = np.load(GOLD)
    res = int(g["res"])
    n = sum(1 for k in g.files if k.startswith("in_"))
    assert n >= 6
    for i in range(n):
        assert np.array_equal(I.ingest_frame(g[f"in_{i}"], res), g[f"out_{i}"]), i
    big = I.ingest_frame(np.random.default_rng(7).integers(0, 256, (720, 1280, 3), dtype=np.uint8), 384)
    assert int(big.astype(np.int64).sum()) == int(g["big_sum"])
    assert int(np.bitwise_xor.reduce(big.astype(np.int64).ravel() * (np.arange(3 * 384 * 384) % 251 + 1))) == int(g["big_xor"])


def test_oracle_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for h, w in [(480, 640), (640, 360), (384, 384), (385, 383), (100, 37), (1080, 1920), (2, 3)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        nw, nh = I.target_size(w, h)
        assert np.array_equal(I.resize_linear_u8(img, nw, nh), cv2.resize(img, (nw, nh))), (h, w)


def test_geometry_and_padding():
    img = np.full((50, 100, 3), 200, np.uint8)
    img[..., 0] = 10                                    # B
    out = I.ingest_frame(img, 64)
    assert out.shape == (3, 64, 64)
    assert I.target_size(100, 50, 64) == (64, 32)
    assert (out[:, :16] == 0).all() and (out[:, 48:] == 0).all()      # centred: 16 rows of padding above and below
    assert (out[2, 16:48] == 10).all() and (out[0, 16:48] == 200).all()   # BGR -> RGB
    assert I.target_size(37, 100) == (142, 384)         # int() truncation of 142.08
