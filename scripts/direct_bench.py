#!/usr/bin/env python
"""f2 evidence: the direct ("Abbe") solver E = A G A^T (reference imageformation.py:3-30) and the
DFT-as-GEMM small-grid variant, FP32 CUDA cores against tensor cores, MEASURED on the B200.

Measurement aid, not part of the product.  For pn in 64/128/256 and the BASELINE cfg1-style inputs it reports

  * ours_fp32        litho_direct_accumulate with LITHO_DIRECT_TC=0 (csrc/direct_kernels.h: FP32 CUDA-core tiles)
  * ours_tc          litho_direct_accumulate, default for pn >= 64 (csrc/direct_tc.cu: tcgen05.mma kind::tf32, 3xTF32,
                     fp32 accumulation in tensor memory), us per source point
  * ours_fft         the FFT-approximation path on the same grid (what the FFT path costs there), us per point
  * tc_tf32x1/x3     the same two complex products per source point as real block GEMMs on the tensor cores
                     (torch.bmm with TF32 enabled = cuBLAS tcgen05 kernels; the launch names are in the ncu list):
                     1 pass (accuracy ~1e-3: fails the 1e-5 bar) and the 3-pass hi/lo split that meets it
  * tc_bf16x?        not run: bf16 needs >= 6 passes for 1e-5, strictly worse than 3xTF32 at the same MMA rate/2
  * accuracy         rel-L2 of every variant's image against the float64 evaluation of the same operator

    python scripts/direct_bench.py [--pn 64 128 256] [--points 64]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lithographysimulator_b200 as L  # noqa: E402
from lithographysimulator_b200 import workloads as wl  # noqa: E402
from lithographysimulator_b200.direct import _operator  # noqa: E402
from lithographysimulator_b200.imaging import AbbeEngine, source_shifts  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us


def tf32_round(x):
    """Round-to-nearest-even to TF32 (10 explicit mantissa bits), as the tensor core's operand conversion."""
    i = x.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


def blk(z):
    """complex [.., m, k] -> real block form [[re, -im], [im, re]] of shape [.., 2m, 2k]."""
    re, im = z.real, z.imag
    return torch.cat([torch.cat([re, -im], -1), torch.cat([im, re], -1)], -2).contiguous()


def stack(z):
    """complex [.., k, n] -> [re; im] of shape [.., 2k, n]."""
    return torch.cat([z.real, z.imag], -2).contiguous()


def gemm_image(A, G, passes):
    """sum_s |A G_s A^T|^2 with real block GEMMs: T = A_blk [G_re; G_im], E^T = A_blk [T^T ...].  passes = 1 (plain
    TF32) or 3 (hi/lo split: a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32 accumulate)."""
    Ab = blk(A)                                    # [2pn, 2S]

    def mm(x, y):
        if passes == 1:
            return x @ y
        xh, yh = tf32_round(x), tf32_round(y)
        return xh @ yh + (x - xh) @ yh + xh @ (y - yh)

    T = mm(Ab, stack(G))                            # [B, 2pn, S]   = A G  (re; im)
    pn = A.shape[0]
    Tc = torch.complex(T[:, :pn], T[:, pn:])        # [B, pn, S]
    E = mm(Ab, stack(Tc.transpose(1, 2)))           # [B, 2pn, pn]  = A (A G)^T = (A G A^T)^T
    return (E[:, :pn] ** 2 + E[:, pn:] ** 2).sum(0).T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pn", type=int, nargs="*", default=[64, 128, 256])
    ap.add_argument("--points", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    eng = AbbeEngine.get(dev)
    rows = []
    for pn in args.pn:
        mask = L.Mask(torch.from_numpy(wl.line_space(pn)), 25, dev)
        mft = mask.fraunhofer(193.0, False)
        ls = L.LightSource(0.6, 0.9, pn, 0.7, 0, 0, dev).generateAnnular()
        stride = max(1, int(round(math.sqrt(float(ls.sum()) / args.points))))
        ls = ls * torch.from_numpy(wl.lattice(pn, stride)).to(dev)
        pf = L.Pupil(pn, 193.0, 0.7, torch.tensor([0, 0, 0, 0, 50], dtype=torch.float16, device=dev), dev).generatePupilFunction()
        shifts = source_shifts(ls, pn)
        n = int(shifts.shape[0])
        r0, r1, c0, c1 = eng.pupil_bbox(pf)[:4]
        # operands of the GEMM form: G_s = roll(P, s) * M restricted to ... the full grid (roll may move the window)
        G = torch.stack([torch.roll(pf, (int(a), int(b)), (0, 1)) * mft for a, b in shifts.cpu().numpy()])
        A = _operator(eng.lib, pn, 25, 193.0, -1, dev)
        torch.cuda.synchronize()
        ref64 = None
        A64, G64 = A.to(torch.complex128), G.to(torch.complex128)
        E64 = A64 @ G64 @ A64.T
        ref64 = (E64.real ** 2 + E64.imag ** 2).sum(0)
        del A64, G64, E64

        def rel(x):
            return float((x.double() - ref64).norm() / ref64.norm())

        out = {"pn": pn, "source_points": n, "pupil_window": [int(r1 - r0 + 1), int(c1 - c0 + 1)]}
        os.environ["LITHO_DIRECT_TC"] = "0"       # FP32 CUDA-core kernels (csrc/direct_kernels.h)
        img = L.abbeImage(mask, mft, pf, ls, 25, mask.deltaK, 193.0, False, dev)
        out["ours_fp32_us_per_point"] = timed(lambda: L.abbeImage(mask, mft, pf, ls, 25, mask.deltaK, 193.0, False, dev)) / n
        out["ours_fp32_rel_l2_vs_f64"] = rel(img)
        os.environ["LITHO_DIRECT_TC"] = "1"       # tcgen05 3xTF32 kernel (csrc/direct_tc.cu), the default for pn >= 64
        img = L.abbeImage(mask, mft, pf, ls, 25, mask.deltaK, 193.0, False, dev)
        out["ours_tc_us_per_point"] = timed(lambda: L.abbeImage(mask, mft, pf, ls, 25, mask.deltaK, 193.0, False, dev)) / n
        out["ours_tc_rel_l2_vs_f64"] = rel(img)
        st = C.c_int(0)
        eng.lib.check(eng.lib.litho_direct_status(C.byref(st), 0), "litho_direct_status")
        out["ours_tc_status"] = st.value
        from lithographysimulator_b200 import direct as D
        out["ours_tc_us_per_point_batch64"] = timed(
            lambda: D.direct_abbe_image(mft, pf, ls, 25, 193.0, dev, batch=64)) / n
        mft_f = mask.fraunhofer(193.0, True)
        out["ours_fft_path_us_per_point"] = timed(lambda: L.abbeImage(mask, mft_f, pf, ls, 25, mask.deltaK, 193.0, True, dev)) / n
        torch.backends.cuda.matmul.allow_tf32 = False
        if hasattr(torch.backends.cuda.matmul, "fp32_precision"):
            torch.backends.cuda.matmul.fp32_precision = "ieee"
        out["cublas_fp32_us_per_point"] = timed(lambda: gemm_image(A, G, 1)) / n
        out["cublas_fp32_rel_l2_vs_f64"] = rel(gemm_image(A, G, 1))
        torch.backends.cuda.matmul.allow_tf32 = True
        if hasattr(torch.backends.cuda.matmul, "fp32_precision"):
            torch.backends.cuda.matmul.fp32_precision = "tf32"
        out["tc_tf32x1_us_per_point"] = timed(lambda: gemm_image(A, G, 1)) / n
        out["tc_tf32x1_rel_l2_vs_f64"] = rel(gemm_image(A, G, 1))
        out["tc_tf32x3_us_per_point"] = timed(lambda: gemm_image(A, G, 3)) / n
        out["tc_tf32x3_rel_l2_vs_f64"] = rel(gemm_image(A, G, 3))
        # GEMMs alone (no split/convert/concat passes): the tensor-core floor of the 3-pass scheme
        Ab, Gs = blk(A), stack(G)
        t1 = timed(lambda: Ab @ Gs)
        Ts = torch.randn(n, 2 * pn, pn, device=dev)
        Ab2 = torch.randn(2 * pn, 2 * pn, device=dev)
        t2 = timed(lambda: Ab2 @ Ts)
        out["tc_tf32_gemm_only_us_per_point_x1"] = (t1 + t2) / n
        out["tc_tf32_gemm_only_us_per_point_x3"] = 3 * (t1 + t2) / n
        flop = 8.0 * pn * pn * pn + 8.0 * pn * pn * pn          # full-grid block GEMMs per source point
        out["tc_gemm_only_tflops_x1"] = flop / ((t1 + t2) / n * 1e-6) / 1e12
        rows.append(out)
        print(json.dumps(out), flush=True)
    print(json.dumps({"direct_bench": rows}), flush=True)


if __name__ == "__main__":
    main()
