"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's per-frame ingest transform.

Reference: test/datasets.py:50-72 and demo/liveinfer.py:32-54 — aspect-preserving `cv2.resize` (INTER_LINEAR, the default) of
a BGR uint8 frame so that its longer side is 384, centred zero padding to 384x384 (`cv2.copyMakeBorder`, BORDER_CONSTANT),
BGR -> RGB and HWC -> CHW.  The arithmetic of `cv2.resize` lives in OpenCV (requirements.txt:23 pins opencv-python==4.10.0.84;
this container has 4.13.0): 8-bit INTER_LINEAR is a fixed-point separable filter (modules/imgproc/src/resize.cpp:
INTER_RESIZE_COEF_BITS = 11, HResizeLinear into int32, VResizeLinear<uchar> with the (b*(S>>4))>>16 form), restated here in
numpy.  Pinned against cv2 itself by tests/test_ingest_cpu.py (live when cv2 imports, and through tests/golden/ingest_*.npz
made by oracle/make_golden_ingest.py)."""
import numpy as np

COEF_BITS = 11
ONE = 1 << COEF_BITS


def _axis_tables(src, dst):
    """Per output index: first source index and the two 11-bit weights (resize.cpp, the dx / dy loops)."""
    scale = 1.0 / (float(dst) / float(src))                    # double, as cv::resize computes it from inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)           # (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _coef(f):
    """saturate_cast<short>(w * 2048): round half to even (cvRound)."""
    w1 = np.rint(f.astype(np.float32) * np.float32(ONE)).astype(np.int64)
    w0 = np.rint((np.float32(1.0) - f.astype(np.float32)) * np.float32(ONE)).astype(np.int64)
    return np.clip(w0, -32768, 32767), np.clip(w1, -32768, 32767)


def resize_linear_u8(img, new_w, new_h):
    """cv2.resize(img, (new_w, new_h)) for uint8 HxWxC, INTER_LINEAR."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    if (new_w, new_h) == (W, H):
        return img.copy()
    sx, fx = _axis_tables(W, new_w)
    lo, hi = sx < 0, sx >= W - 1                               # horizontal: clamp AND drop the fraction
    fx = np.where(lo | hi, np.float32(0), fx).astype(np.float32)
    sx = np.where(lo, 0, np.where(hi, W - 1, sx))
    a0, a1 = _coef(fx)
    sy, fy = _axis_tables(H, new_h)                            # vertical: rows are clamped, the fraction is kept
    b0, b1 = _coef(fy)
    y0, y1 = np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    x1 = np.minimum(sx + 1, W - 1)
    src = img.astype(np.int64)
    rows = src[:, sx, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]          # HResizeLinear -> int
    S0, S1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2   # VResizeLinear<uchar>
    return np.clip(out, 0, 255).astype(np.uint8)


def target_size(in_w, in_h, res=384):
    """test/datasets.py:51-58 (int() truncation as written there)."""
    if in_w > in_h:
        return res, int((in_h / in_w) * res)
    return int((in_w / in_h) * res), res


def ingest_frame(frame_bgr, res=384):
    """One decoded frame (HxWx3 uint8, BGR as cv2 yields it) -> [3, res, res] uint8 RGB, resized and zero-padded."""
    H, W, _ = frame_bgr.shape
    nw, nh = target_size(W, H, res)
    r = resize_linear_u8(frame_bgr, nw, nh)
    top, left = (res - nh) // 2, (res - nw) // 2
    canvas = np.zeros((res, res, 3), np.uint8)
    canvas[top:top + nh, left:left + nw] = r
    return np.ascontiguousarray(canvas[:, :, ::-1].transpose(2, 0, 1))
