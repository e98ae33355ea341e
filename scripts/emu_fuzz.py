#!/usr/bin/env python
"""One-off fuzz of the plan space on the CPU emulation of the kernels (test infrastructure): grids 64/96/128 px, pixel
sizes 12/25/50 nm, pupil windows biased towards the edges of the fast path (M-1 .. M+4 samples, sparse or missing rim
lines), sources inside and outside the no-wrap range, weights, batches -- each against the oracle.

    python scripts/emu_fuzz.py FIRST_SEED LAST_SEED      (LITHO_TMA=0 / LITHO_EMU_ORDER=random|reverse for variants)

End of round 1: 480 seeds, no mismatch above 1e-5."""
import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
from oracle import abbe_oracle as O
bad=[]
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(99000 + seed)
    pn = int(rng.choice([64, 96, 128])); ps = float(rng.choice([12, 25, 50]))
    _, N = O.calculate_epsilon_n(4 / pn, ps, 193.0)
    if N < pn: ps = 25.0; _, N = O.calculate_epsilon_n(4 / pn, ps, 193.0)
    # bias window sizes towards the interesting edges (M-1 .. M+4 for M = 16, 32, 64)
    def size():
        if rng.random() < 0.6:
            M = int(rng.choice([16, 32, 64])); v = M + int(rng.integers(-1, 5))
        else:
            v = int(rng.integers(1, pn // 2 + 8))
        return max(1, min(v, pn))
    wr, wc = size(), size()
    r0, c0 = int(rng.integers(0, pn - wr + 1)), int(rng.integers(0, pn - wc + 1))
    pup = np.zeros((pn, pn), np.complex64)
    pup[r0:r0 + wr, c0:c0 + wc] = rng.standard_normal((wr, wc)) + 1j * rng.standard_normal((wr, wc))
    mode = seed % 4
    if mode == 0 and wr > 2 and wc > 2:   # sparse rims
        pup[r0, c0 + wc // 3:] = 0; pup[r0 + wr - 1, :c0 + wc // 2] = 0
        pup[r0:r0 + wr // 2, c0 + wc - 1] = 0; pup[r0 + wr // 3:, c0] = 0
    if mode == 1 and wr > 4 and wc > 4:   # second line in from the edge empty
        pup[r0 + 1, :] = 0; pup[:, c0 + wc - 2] = 0
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    n_src = int(rng.integers(1, 5))
    nz = np.argwhere(pup != 0)
    if len(nz) == 0: continue
    br0, br1, bc0, bc1 = nz[:,0].min(), nz[:,0].max(), nz[:,1].min(), nz[:,1].max()
    if seed % 5 == 4:
        shifts = rng.integers(-pn // 2, pn // 2, (n_src, 2)).astype(np.int32)
    else:
        shifts = np.stack([rng.integers(-br0, pn - 1 - br1 + 1, n_src), rng.integers(-bc0, pn - 1 - bc1 + 1, n_src)], 1).astype(np.int32)
    w = rng.uniform(0.25, 2.0, n_src).astype(np.float32) if seed % 2 else None
    img, info = H.emu_abbe_fft(mft, pup, None, ps, 193.0, shifts=shifts, weights=w, postprocess=False, batch=int(rng.integers(0, 4)))
    ref = np.zeros((pn, pn))
    for i, (d0, d1) in enumerate(shifts):
        wi = 1.0 if w is None else float(w[i])
        ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pup, (int(d0), int(d1)), (0, 1)), mft, pn, N)) ** 2
    err = O.rel_l2(img, ref)
    if not (err < 1e-5):
        bad.append((seed, pn, ps, N, wr, wc, r0, c0, info["path"], info["M"], err)); print("BAD", bad[-1], flush=True)
print("done", sys.argv[1], sys.argv[2], "bad:", len(bad))
