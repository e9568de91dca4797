"""GPU bring-up of the tcgen05 GEMM: diagnostic cases first (identity weights expose layout/swizzle bugs), then the
shapes the path uses, then a quick timing.  Writes gpurun_out/bringup_gemm.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import ops  # noqa: E402
from mmduet_b200._lib import ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, EPI_BF16, EPI_F32, EPI_RESID_F32  # noqa: E402

report = {"cases": []}
dev = "cuda"
torch.manual_seed(0)


def ref_mm(x, w):
    return x.float() @ w.float().t()


def record(name, got, want, extra=None):
    got = got.float()
    want = want.float()
    err = (got - want).abs()
    r = {"name": name, "max_abs": err.max().item(), "mean_abs": err.mean().item(), "ref_absmax": want.abs().max().item(),
         "nan": bool(torch.isnan(got).any().item())}
    if r["max_abs"] > 0.05 * max(1.0, r["ref_absmax"]):
        bad = err > 0.05 * max(1.0, r["ref_absmax"])
        rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        r["bad_frac"] = bad.float().mean().item()
        r["bad_rows_head"] = rows
        r["bad_cols_head"] = cols
        r["sample_got"] = got[:4, :8].tolist()
        r["sample_want"] = want[:4, :8].tolist()
    if extra:
        r.update(extra)
    report["cases"].append(r)
    print(json.dumps(r)[:600], flush=True)
    return r


def run_case(fn, name):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        report["cases"].append({"name": name, "exception": repr(e)})
        print("EXC", name, repr(e), flush=True)
        return False
    return True


def diag_identity():
    # D[m, n] = X[m, n] for n < K when W = [I; 0]
    M, N, K = 128, 128, 64
    x = torch.randn(M, K, device=dev).bfloat16()
    w = torch.zeros(N, K, device=dev).bfloat16()
    w[:K, :K] = torch.eye(K, device=dev).bfloat16()
    out = ops.gemm(x, w, epi=EPI_F32)
    torch.cuda.synchronize()
    record("identity_128x128x64", out, ref_mm(x, w))


def shape_case(M, N, K, epi=EPI_BF16, act=ACT_NONE, with_bias=True):
    def f():
        x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev) if with_bias else None
        want = ref_mm(x, w) + (b if b is not None else 0)
        if epi == EPI_RESID_F32:
            res = torch.randn(M, N, device=dev)
            want = want + res
            out = res.clone()
            ops.gemm(x, w, bias=b, out=out, epi=epi)
        else:
            if act == ACT_GELU_TANH:
                want = torch.nn.functional.gelu(want, approximate="tanh")
            elif act == ACT_GELU_ERF:
                want = torch.nn.functional.gelu(want)
            out = ops.gemm(x, w, bias=b, act=act, epi=epi)
        torch.cuda.synchronize()
        record(f"normal_M{M}_N{N}_K{K}_epi{epi}_act{act}", out, want)
    return f


def t_case(M, N, K, splits):
    def f():
        x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        parts = ops.gemm_t_partials(x, w, splits)
        torch.cuda.synchronize()
        record(f"T_f32_M{M}_N{N}_K{K}_s{splits}", parts.sum(0), ref_mm(x, w), {"planes": parts.shape[0]})
    return f


def swiglu_case(M, N, K):
    def f():
        x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        wg = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        wu = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        out = ops.gemm_t_swiglu(x, wg, wu)
        torch.cuda.synchronize()
        want = torch.nn.functional.silu(ref_mm(x, wg)) * ref_mm(x, wu)
        record(f"T_swiglu_M{M}_N{N}_K{K}", out, want)
    return f


def timing(M, N, K, epi=EPI_BF16, iters=20):
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(x, w, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(x, w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # cuBLAS for context
    for _ in range(3):
        torch.matmul(x, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(x, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    r = {"name": f"time_M{M}_N{N}_K{K}", "ms": ms, "tflops": 2 * M * N * K / ms / 1e9, "cublas_ms": ms_ref,
         "cublas_tflops": 2 * M * N * K / ms_ref / 1e9}
    report["cases"].append(r)
    print(json.dumps(r), flush=True)


def timing_t(M, N, K, splits, iters=20):
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    ws = [(torch.randn(N, K, device=dev) * 0.05).bfloat16() for _ in range(8)]  # rotate weights: > L2 in aggregate
    out = None
    for w in ws[:3]:
        out = ops.gemm_t_partials(x, w, splits, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.gemm_t_partials(x, ws[i % 8], splits, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    r = {"name": f"timeT_M{M}_N{N}_K{K}_s{splits}", "ms": ms, "weight_GBps": N * K * 2 / ms / 1e6}
    report["cases"].append(r)
    print(json.dumps(r), flush=True)


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    print(torch.cuda.get_device_name(0), flush=True)
    ok = run_case(diag_identity, "identity")
    if ok:
        cases = [
            ("n1", shape_case(128, 256, 64, epi=EPI_F32, with_bias=False)),
            ("n2", shape_case(128, 256, 256, epi=EPI_F32, with_bias=False)),
            ("n3", shape_case(300, 1152, 1152)),
            ("n4", shape_case(729 * 2, 3456, 1152)),
            ("n5", shape_case(729, 4304, 1152, act=ACT_GELU_TANH)),
            ("n6", shape_case(729, 1152, 4304, epi=EPI_RESID_F32)),
            ("n7", shape_case(729, 1152, 592, epi=EPI_F32)),
            ("n8", shape_case(169 * 3, 3584, 1152, act=ACT_GELU_ERF)),
            ("n9", shape_case(5000, 1152, 1152, epi=EPI_RESID_F32)),
            ("t1", t_case(49, 4608, 3584, 4)),
            ("t2", t_case(49, 3584, 3584, 5)),
            ("t3", t_case(81, 3584, 18944, 5)),
            ("t4", t_case(300, 3584, 3584, 2)),
            ("t5", t_case(1, 152064, 3584, 1)),
            ("s1", swiglu_case(49, 18944, 3584)),
            ("s2", swiglu_case(200, 18944, 3584)),
        ]
        for name, f in cases:
            if not run_case(f, name):
                break
        else:
            try:
                timing(729 * 32, 3456, 1152)
                timing(729 * 32, 4304, 1152)
                timing(729 * 32, 1152, 4304)
                timing(8192, 8192, 8192)
                timing_t(49, 4608, 3584, 4)
                timing_t(49, 3584, 18944, 5)
                timing_t(49, 3584, 3584, 5)
            except Exception as e:  # noqa: BLE001
                report["timing_exception"] = repr(e)
    with open("gpurun_out/bringup_gemm.json", "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
