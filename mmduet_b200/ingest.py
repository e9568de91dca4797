"""Frame ingest on the GPU: what the reference does per sampled frame on the CPU with cv2 —
aspect-preserving `cv2.resize` of the longer side to 384, centred zero padding, BGR->RGB, HWC->CHW
(test/datasets.py:50-72, demo/liveinfer.py:32-54) — as one CUDA kernel, bit-exact with cv2's 8-bit INTER_LINEAR.
Video decoding (cv2.VideoCapture) stays on the host; its frames are uploaded as they are (uint8 BGR HWC)."""
import torch

from . import _lib


def target_size(in_w, in_h, res=384):
    """(new_w, new_h) as the reference computes them (test/datasets.py:51-58)."""
    if in_w > in_h:
        return res, int((in_h / in_w) * res)
    return int((in_w / in_h) * res), res


def ingest_frames(frames_bgr, res=384, out=None):
    """frames_bgr: uint8 CUDA tensor [T, H, W, 3] (or [H, W, 3]) in cv2's BGR order -> uint8 [T, 3, res, res] RGB, the tensor
    `LiveInferForBenchmark.input_video_stream` / `visual_embed(normalize=True)` take."""
    if frames_bgr.dim() == 3:
        frames_bgr = frames_bgr[None]
    if frames_bgr.dtype != torch.uint8 or frames_bgr.dim() != 4 or frames_bgr.shape[-1] != 3 or not frames_bgr.is_cuda:
        raise _lib.MmdError("ingest_frames: expected a uint8 CUDA tensor [T, H, W, 3]")
    frames_bgr = frames_bgr.contiguous()
    T, H, W, _ = frames_bgr.shape
    if out is None:
        out = torch.empty(T, 3, res, res, dtype=torch.uint8, device=frames_bgr.device)
    elif out.shape != (T, 3, res, res) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise _lib.MmdError("ingest_frames: out must be a contiguous uint8 [T, 3, res, res] tensor")
    lib = _lib.load()
    _lib.check(lib.mmd_frame_ingest(frames_bgr.data_ptr(), T, H, W, out.data_ptr(), res, _lib.stream_ptr()), "mmd_frame_ingest")
    return out
