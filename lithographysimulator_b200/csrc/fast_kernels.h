// Fast path of the Abbe FFT-approximation hot loop (reference imageformation.py:59-67).
//
// Preconditions (checked on the host, litho_abi.cu): no source point makes roll() wrap the pupil
// window around the grid, and the window fits S <= M+1 with M a power of two (M = pn/2 for the
// reference's pupils).  Then two facts make the loop ~4x cheaper than the literal algorithm:
//
//  1. Shift equivalence.  The window's position only multiplies the field by
//     unit-modulus phase ramps, which |E|^2 removes, so the inputs are indexed 0..S-1 for every
//     source point and all twiddles become per-thread constants held in shared-memory tables.
//
//  2. Coarse sampling.  |E_s|^2 is a trigonometric polynomial with frequencies |d| <= S-1 <= M
//     (in units of 1/N), so its samples on the Nc = 2M grid of spacing q = N/Nc determine it.
//     Each 1-D transform is then an Nc-point DFT of <= M+1 inputs = 2 length-M FFTs
//         X[2k+r] = FFT_M( x[u] * w_Nc^(r*u) folded modulo M )[k],   r = 0,1
//     with every output used, instead of R = N/M pruned ones.  The intensity is accumulated on the
//     coarse grid and interpolated exactly (spectrally) to the reference's pn centre pixels once
//     per image (litho_abi.cu: finalize).  The single aliased frequency line |d| = M, which exists
//     only when S = M+1 (the pupil's fp16 rim pixels), is carried separately by rim_body below.
//
// FFTs hold PPT = 32 points per thread (radix 32 x 32 for M = 1024: one shared-memory exchange, 128
// registers) or PPT = 16 (radix 16 x 16 x 4: two exchanges, 64 registers, twice the resident warps);
// the plan picks the variant per M from measurements.
#pragma once
#include "fft_core.h"

namespace litho {

LITHO_HD int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// The window may exceed the even fit S = M+1 by up to RIM_EXTRA inputs per axis (S <= M + 1 + RIM_EXTRA): the
// reference's fp16 pupil grid gives S = pn/2 + 3 at pn = 8192.  Inputs u = M .. S-1 fold onto slots u - M of the
// length-M transforms, and the frequency lines |d| = M .. S-1 of sum_s |E_s|^2, which alias on the coarse grid,
// are carried separately by rim_body (RIM_LINES of them per axis at most).
#ifndef LITHO_RIM_EXTRA
#define LITHO_RIM_EXTRA 2
#endif
constexpr int RIM_EXTRA = LITHO_RIM_EXTRA;
constexpr int RIM_LINES = RIM_EXTRA + 1;

template <int M, int PPT>
struct FastShape {
    using Sh = FftShape<M, PPT>;
    static constexpr int TG = Sh::TG;
    static constexpr int R1 = Sh::NP > 1 ? Sh::radix(1) : 1;
    static constexpr int R2 = Sh::NP > 2 ? Sh::radix(2) : 1;
    static constexpr int NS1 = PPT, NS2 = PPT * R1;
    // shared-memory twiddle tables: pre[u] = w_2M^u (u = 0..M), tw1[(t-1)*NS1+k], tw2[(t-1)*NS2+k]
    static constexpr int PRE_OFF = 0;
    static constexpr int TW1_OFF = M + 1;
    static constexpr int TW2_OFF = TW1_OFF + (R1 - 1) * NS1;
    static constexpr int NTAB = TW2_OFF + (R2 - 1) * NS2;
    static constexpr int NTAB_PAD = (NTAB + 1) & ~1;  // keep the exchange region 16-byte aligned
    // rows kernel: one FFT group per TG threads; groups inside a warp sync with __syncwarp, groups of
    // 2..4 warps with a named barrier, a group that is the whole CTA with __syncthreads
    static constexpr int ROW_THREADS = TG <= 256 ? 256 : TG;
    static constexpr int ROW_GROUPS = ROW_THREADS / TG;
    static constexpr int ROW_SYNC = TG <= 32 ? 0 : (ROW_GROUPS > 1 ? 1 : 2);  // 0 warp, 1 named barrier, 2 CTA
    // M >= 4096: the full tables (65 KB) + two exchange regions (68 KB) would leave one 256-thread CTA per SM.  The row
    // pass then uses the compact device table of the TMA column kernel (TmaShape layout: pre[0..M/2] mirrored about
    // M/2, tw1) in shared memory and reads the third-pass table tw2 from global memory through L1: 24 KB of tables,
    // 92 KB per CTA, two CTAs per SM.
    static constexpr bool ROW_COMPACT = (M >= 4096);
    static constexpr int CPRE_N = M / 2 + 1;
    static constexpr int CTW1_OFF = (CPRE_N + 1) & ~1;
    static constexpr int CTW2_OFF = CTW1_OFF + (R1 - 1) * NS1;          // (global memory only)
    static constexpr int CTAB_SMEM_PAD = (CTW2_OFF + 1) & ~1;
    // M = 1024 (one FFT per warp, two radix-32 passes): the residue-1 pre-twiddle w_2M^u, u = g + 32 e, is folded away.
    // Its register part w_64^e is a compile-time constant per register of the first pass; its thread part w_2M^g
    // reaches the second pass in register e' = g of every thread (the exchange is a transpose), where it merges with
    // the inter-pass twiddle: w_2M^e' * w_M^(e' g') = w_2M^(e' (2 g' + 1)).  The row pass then needs no pre table at
    // all, just a second copy of the inter-pass table: 32 fewer shared-memory loads per residue-1 FFT.
    static constexpr bool ROW_FOLD = (TG == 32 && Sh::NP == 2 && R1 == 32);
    static constexpr int FOLD_ODD_OFF = (R1 - 1) * NS1;            // fold table: [tw1][tw1_odd]
    static constexpr int FOLD_TAB_ELEMS = 2 * (R1 - 1) * NS1;
    static constexpr int ROW_TAB_ELEMS = ROW_COMPACT ? CTAB_SMEM_PAD : (ROW_FOLD ? FOLD_TAB_ELEMS : NTAB_PAD);
    static constexpr size_t ROW_SMEM = (size_t)(ROW_TAB_ELEMS + ROW_GROUPS * Sh::SMEM_ELEMS) * sizeof(cplx);
    // cols kernel: CB adjacent columns per CTA, column fastest in the thread index (256 threads per CTA
    // where possible; CB >= 4 keeps every global request at full 32-byte sectors)
// threads per column-pass CTA: 512 (16 columns = full 128-byte lines at M = 1024) measured 3% faster
// than 256 and 50% faster than 128 on B200 (profiles/README.md)
#ifndef LITHO_COL_THREADS
#define LITHO_COL_THREADS 512
#endif
    // COL_DUAL (M <= 1024): one CTA computes BOTH output-row residues of its tile -- threads [0,HALF) take
    // rr = 0, [HALF,2*HALF) rr = 1 -- so the two halves request the same T sectors at the same time and
    // the second request is served by L1 (T crosses L2/HBM once instead of twice).
#ifndef LITHO_COL_DUAL
#define LITHO_COL_DUAL 1
#endif
    static constexpr bool COL_DUAL = (LITHO_COL_DUAL != 0) && PPT == 32 && TG <= 32;
    static constexpr int CB_SINGLE = (LITHO_COL_THREADS / TG) >= 16 ? 16 : ((LITHO_COL_THREADS / TG) >= 4 ? (LITHO_COL_THREADS / TG) : 4);
    static constexpr int CB_DUAL = (256 / TG) >= 16 ? 16 : (256 / TG);
    static constexpr int CB = COL_DUAL ? CB_DUAL : CB_SINGLE;
    static constexpr int COL_HALF = CB * TG;                       // threads per residue
    static constexpr int COL_THREADS = COL_DUAL ? 2 * COL_HALF : COL_HALF;
    static constexpr int COL_GRID_Y = COL_DUAL ? 1 : 2;
    static constexpr size_t COL_SMEM = (size_t)(NTAB_PAD + (COL_DUAL ? 2 : 1) * CB * Sh::SMEM_ELEMS) * sizeof(cplx);
    // occupancy targets: 4 registers per FFT point held -> 128 regs (PPT 32) / 64 regs (PPT 16) per thread
    static constexpr int TARGET_THREADS = PPT == 32 ? 512 : 1024;
    static constexpr int COL_MIN_BLOCKS = (TARGET_THREADS / COL_THREADS) >= 1 ? (TARGET_THREADS / COL_THREADS) : 1;
    static constexpr int ROW_MIN_BLOCKS = (TARGET_THREADS / ROW_THREADS) >= 1 ? (TARGET_THREADS / ROW_THREADS) : 1;
};

template <int M, int PPT>
struct SmemTw {
    const cplx* tab;
    template <int PASS>
    LITHO_HD cplx get(int t, int k) const {
        using F = FastShape<M, PPT>;
        if constexpr (PASS == 1) return tab[F::TW1_OFF + (t - 1) * F::NS1 + k];
        else return tab[F::TW2_OFF + (t - 1) * F::NS2 + k];
    }
};

// Row-pass twiddles: full shared-memory tables, or (FastShape::ROW_COMPACT) compact shared-memory tables + tw2 from global
// w_64^e = exp(2 pi i e / 64), e = 0..31: the register part of the residue-1 pre-twiddle (FastShape::ROW_FOLD)
LITHO_HD cplx w64(int e) {   // (e is a compile-time constant after unrolling: the tables fold to immediates)
    const float W64_RE[32] = {1.f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f,
                              0.773010453f, 0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f,
                              0.290284677f, 0.195090322f, 0.0980171403f, 0.f, -0.0980171403f, -0.195090322f,
                              -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f, -0.707106781f,
                              -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f,
                              -0.995184727f};
    const float W64_IM[32] = {0.f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f,
                              0.634393284f, 0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f,
                              0.956940336f, 0.98078528f, 0.995184727f, 1.f, 0.995184727f, 0.98078528f, 0.956940336f,
                              0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f, 0.707106781f, 0.634393284f,
                              0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f};
    return mk(W64_RE[e], W64_IM[e]);
}

template <int M, int PPT>
struct RowTw {
    const cplx* tab;    // ROW_FOLD: already offset to the table of this item's residue
    const cplx* gtab;
    template <int PASS>
    LITHO_HD cplx get(int t, int k) const {
        using F = FastShape<M, PPT>;
        if constexpr (F::ROW_FOLD) {
            return tab[(t - 1) * F::NS1 + k];
        } else if constexpr (!F::ROW_COMPACT) {
            if constexpr (PASS == 1) return tab[F::TW1_OFF + (t - 1) * F::NS1 + k];
            else return tab[F::TW2_OFF + (t - 1) * F::NS2 + k];
        } else {
            if constexpr (PASS == 1) return tab[F::CTW1_OFF + (t - 1) * F::NS1 + k];
            else return ldg_c(gtab + F::CTW2_OFF + (t - 1) * F::NS2 + k);
        }
    }
};

// pre-twiddle w_2M^u of slot u = g + TG*e from the full table (pre[0..M]) or the compact one (pre[0..M/2], mirrored)
template <int M, int PPT>
LITHO_HD cplx row_pre(const cplx* tab, int g, int e) {
    using F = FastShape<M, PPT>;
    constexpr int TG = F::TG;
    if constexpr (!F::ROW_COMPACT) {
        return tab[F::PRE_OFF + g + TG * e];
    } else {
        if (TG * e + TG - 1 <= M / 2) return tab[g + TG * e];
        const cplx t = tab[M - TG * e - g];     // u >= M/2 (M/2 is a multiple of TG): w^u = -conj(w^(M-u))
        return mk(-t.x, t.y);
    }
}

// MODE 0: warp barrier, 1: named barrier `id` over `n` threads, 2: CTA barrier
template <class Ctx, int MODE>
struct GroupSync {
    const Ctx& ctx;
    int id, n;
    LITHO_HD void sync() const {
        if constexpr (MODE == 0) ctx.sync_warp();
        else if constexpr (MODE == 1) ctx.sync_named(id, n);
        else ctx.sync();
    }
};

struct FastRowsParams {
    const cplx* pupil;
    const cplx* mask;
    int pn;
    int pr0, pc0, Sr, Sc;
    const int2_* shifts;
    int s_begin, batch;
    const cplx* tables;
    const cplx* tables_c;   // compact tables (TmaShape layout), used by the row pass when FastShape::ROW_COMPACT
    const cplx* tables_r;   // [tw1][tw1_odd], used by the row pass when FastShape::ROW_FOLD
    cplx* T;  // [n_focus][batch][2][Sr][M]
    int* status;  // plan-owned device words: [0] set to 1 when a shift had to be clamped (contract violated)
    // focus batching (BASELINE cfg5): n_focus pupil planes (pupil + f*pupil_stride) share the mask
    // spectrum and the source points; work items of one (source point, window row) are adjacent for all focus
    // values and both residues, so the mask row is fetched from L2/HBM once and served by L1 to the others
    int n_focus;            // >= 1
    size_t pupil_stride;    // elements between consecutive pupil planes
    size_t t_focus_stride;  // elements between the T blocks of consecutive focus values (batch_alloc*2*Sr*M)
};

struct FastColsParams {
    const cplx* T;
    int batch, Sr;
    const float* weights;
    int s_begin;
    const cplx* tables;
    float* ic;  // [2][2][M][M] : ((rr*2 + rc)*M + kr)*M + kc, accumulated
    // TMA-staged variant: T seen as one 2-D tensor (rows of M complex elements) starting at tile.base;
    // row_begin = row of T[sl = 0][rc = 0][u = 0] of this launch in that tensor; nbox boxes per tile
    // TMA-staged variant (fast_cols_tma_body): use_tma = columns per tile (0: plain loads)
    int use_tma, nbox, rim;      // boxes per tile; rim: rows u = M .. Sr-1 of the tile (0 .. RIM_LINES), 1-D copies
    long long row_begin;
    const cplx* tables_c;        // compact twiddle tables (TmaShape layout)
    int* status;                 // plan-owned device words: [1] set to 1 when a tile copy never completed
    TileMap tile;
};

template <int M, int PPT, class Ctx>
LITHO_HD void fast_load_tables(const cplx* tables, cplx* tab, const Ctx& ctx) {
    for (int i = ctx.tid(); i < FastShape<M, PPT>::NTAB; i += ctx.bdim()) tab[i] = tables[i];
    ctx.sync();
}
// Asynchronous variant: start the copy (16-byte chunks; the device table is padded to NTAB_PAD elements),
// do other work, then fast_tables_wait() before the first use.
template <int M, int PPT, class Ctx>
LITHO_HD void fast_tables_begin(const cplx* tables, cplx* tab, const Ctx& ctx) {
    for (int i = ctx.tid(); i < FastShape<M, PPT>::NTAB_PAD / 2; i += ctx.bdim())
        ctx.cp_async16(tab + 2 * i, tables + 2 * i);
}
template <class Ctx>
LITHO_HD void fast_tables_wait(const Ctx& ctx) {
    ctx.cp_async_wait();
    ctx.sync();
}

// Inputs of one row FFT: G_s[line][u] = pupil * shifted mask spectrum, pre-twiddled for residue r and with
// the rim input folded onto slot 0.  Branch-free loads (index clamped, value masked afterwards) so that all
// loads of a half are in flight together instead of one load-use round trip per element.
template <int M, int PPT>
LITHO_HD void fast_row_load(cplx (&v)[PPT], const FastRowsParams& P, int s, int line, int r, int g, const cplx* tab,
                            int f = 0) {
    using F = FastShape<M, PPT>;
    constexpr int TG = F::TG;
    constexpr int H = PPT / 2;
    const int2_ sh = P.shifts[s];
    const cplx* prow = P.pupil + (size_t)f * P.pupil_stride + (size_t)(P.pr0 + line) * P.pn + P.pc0;
    // shifts are inside the plan's no-wrap range by contract; clamping keeps a violated contract
    // memory-safe (the result is then wrong, never out of bounds)
    const int mr = iclamp(P.pr0 + line + sh.x, 0, P.pn - 1);
    const int mc = iclamp(P.pc0 + sh.y, 0, P.pn - P.Sc);
    if ((mr != P.pr0 + line + sh.x || mc != P.pc0 + sh.y) && g == 0 && P.status) P.status[0] = 1;
    const cplx* mrow = P.mask + (size_t)mr * P.pn + mc;
    const int last = P.Sc - 1;
    const cplx* pg = prow + g;
    const cplx* mg = mrow + g;
    if (last >= M - 1) {  // common case: every slot has an input
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            cplx a[H], b[H];
#pragma unroll
            for (int i = 0; i < H; ++i) {
                a[i] = ldg_c(pg + TG * (H * h + i));
                b[i] = ldg_c(mg + TG * (H * h + i));
            }
#pragma unroll
            for (int i = 0; i < H; ++i) v[H * h + i] = cmul(a[i], b[i]);
        }
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            cplx a[H], b[H];
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const int u = g + TG * (H * h + i);
                const int uc = u < last ? u : last;
                a[i] = ldg_c(prow + uc);
                b[i] = ldg_c(mrow + uc);
            }
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const int u = g + TG * (H * h + i);
                const cplx x = cmul(a[i], b[i]);
                v[H * h + i] = mk(u <= last ? x.x : 0.f, u <= last ? x.y : 0.f);
            }
        }
    }
    // inputs u = M .. Sc-1 fold onto slots u - M with w_2M^(r*M) = (-1)^r; the slot's pre-twiddle follows
    if (P.Sc > M) {
#pragma unroll
        for (int k = 0; k <= RIM_EXTRA; ++k) {
            if (k < P.Sc - M && g == k % TG) {
                const cplx y = cmul(ldg_c(prow + M + k), ldg_c(mrow + M + k));
                v[k / TG] = r ? csub(v[k / TG], y) : cadd(v[k / TG], y);
            }
        }
    }
    if (r) {
        if constexpr (F::ROW_FOLD) {
#pragma unroll
            for (int e = 1; e < PPT; ++e) v[e] = cmul(v[e], w64(e));   // (the thread part sits in the pass-1 table)
        } else {
#pragma unroll
            for (int e = 0; e < PPT; ++e) v[e] = cmul(v[e], row_pre<M, PPT>(tab, g, e));
        }
    }
}

// Inputs of one column FFT: T[u][kc] for the thread's slots u = g + TG*e (column kc = src offset).
// CG selects L2-coherent loads (T written by other CTAs of the same launch).
template <int M, int PPT, bool CG>
LITHO_HD void fast_col_load_raw(cplx (&v)[PPT], const cplx* src, int Sr, int g) {
    using F = FastShape<M, PPT>;
    constexpr int TG = F::TG;
    const int last = Sr - 1;
    const cplx* srcg = src + (size_t)g * M;
    if (last >= M - 1) {  // common case Sr >= M: every slot has an input, no masking needed
#pragma unroll
        for (int e = 0; e < PPT; ++e) v[e] = CG ? ldcg_c(srcg + (size_t)e * (TG * M)) : ldg_c(srcg + (size_t)e * (TG * M));
    } else {
#pragma unroll
        for (int e = 0; e < PPT; ++e) {
            const int u = g + TG * e;
            const cplx* pa = (u <= last) ? srcg + (size_t)e * (TG * M) : src + (size_t)last * M;
            const cplx x = CG ? ldcg_c(pa) : ldg_c(pa);
            v[e] = mk(u <= last ? x.x : 0.f, u <= last ? x.y : 0.f);
        }
    }
}

template <int M, int PPT, bool CG>
LITHO_HD void fast_col_finish(cplx (&v)[PPT], const cplx* src, int Sr, int rr, int g, const cplx* tab) {
    using F = FastShape<M, PPT>;
    constexpr int TG = F::TG;
    if (Sr > M) {  // rows u = M .. Sr-1 fold onto slots u - M with (-1)^rr, before the slot's pre-twiddle
#pragma unroll
        for (int k = 0; k <= RIM_EXTRA; ++k) {
            if (k < Sr - M && g == k % TG) {
                const cplx y = CG ? ldcg_c(src + (size_t)(M + k) * M) : ldg_c(src + (size_t)(M + k) * M);
                v[k / TG] = rr ? csub(v[k / TG], y) : cadd(v[k / TG], y);
            }
        }
    }
    if (rr) {
#pragma unroll
        for (int e = 0; e < PPT; ++e) v[e] = cmul(v[e], tab[F::PRE_OFF + g + TG * e]);
    }
}

template <int M, int PPT, bool CG>
LITHO_HD void fast_col_load(cplx (&v)[PPT], const cplx* src, int Sr, int rr, int g, const cplx* tab) {
    fast_col_load_raw<M, PPT, CG>(v, src, Sr, g);
    fast_col_finish<M, PPT, CG>(v, src, Sr, rr, g, tab);
}

// grid.x = any (persistent over the batch*Sr*2 work items), block = ROW_THREADS
template <int M, int PPT, class Ctx>
LITHO_HD void fast_rows_body(const FastRowsParams& P, const Ctx& ctx, cplx* smem) {
    using F = FastShape<M, PPT>;
    using Sh = typename F::Sh;
    constexpr int TG = F::TG;
    cplx* tab = smem;
    if constexpr (F::ROW_FOLD) {
        for (int i = ctx.tid(); i < F::FOLD_TAB_ELEMS; i += ctx.bdim()) tab[i] = P.tables_r[i];
        ctx.sync();
    } else if constexpr (F::ROW_COMPACT) {
        for (int i = ctx.tid(); i < F::CTAB_SMEM_PAD; i += ctx.bdim()) tab[i] = P.tables_c[i];
        ctx.sync();
    } else {
        fast_load_tables<M, PPT>(P.tables, tab, ctx);
    }
    const int grp = ctx.tid() / TG;
    const int g = ctx.tid() - grp * TG;
    cplx* ex = smem + F::ROW_TAB_ELEMS + grp * Sh::SMEM_ELEMS;
    const int nf = P.n_focus > 1 ? P.n_focus : 1;
    const int total = P.batch * P.Sr * 2 * nf;
    const int stride = ctx.gdx() * F::ROW_GROUPS;
    const int rounds = (total + stride - 1) / stride;
    const GroupSync<Ctx, F::ROW_SYNC> gs{ctx, 1 + grp, TG};

    for (int it = 0; it < rounds; ++it) {
        const int item = it * stride + ctx.bx() * F::ROW_GROUPS + grp;
        const bool active = item < total;
        const RowTw<M, PPT> tw{tab + ((F::ROW_FOLD && (item & 1)) ? F::FOLD_ODD_OFF : 0), P.tables_c};
        // item = ((sl*Sr + line)*nf + f)*2 + r : residue fastest, then focus value, then window row, then source point
        const int r = item & 1;
        int li = item >> 1;
        int f = 0;
        if (nf > 1) {
            const int q = li / nf;
            f = active ? li - q * nf : 0;
            li = q;
        }
        const int sl = active ? li / P.Sr : 0;
        const int line = active ? li - sl * P.Sr : 0;
        cplx v[PPT];
        if (active) {
            fast_row_load<M, PPT>(v, P, P.s_begin + sl, line, r, g, tab, f);
        } else {
#pragma unroll
            for (int e = 0; e < PPT; ++e) v[e] = mk(0.f, 0.f);
        }
        fft_run<M, PPT, false>(v, ex, 1, g, tw, gs);
        if (active) {
            cplx* dst = P.T + (size_t)f * P.t_focus_stride + ((size_t)(sl * 2 + r) * P.Sr + line) * M + g;
#pragma unroll
            for (int e = 0; e < PPT; ++e) dst[TG * e] = v[e];
        }
    }
}

// grid.x = 2*M/CB column blocks (rc major), grid.y = 2 (rr), block = COL_THREADS
template <int M, int PPT, class Ctx>
LITHO_HD void fast_cols_body(const FastColsParams& P, const Ctx& ctx, cplx* smem) {
    using F = FastShape<M, PPT>;
    constexpr int TG = F::TG;
    constexpr int CB = F::CB;
    cplx* tab = smem;
    fast_tables_begin<M, PPT>(P.tables, tab, ctx);  // lands while the first inputs are in flight
    const int half = F::COL_DUAL ? ctx.tid() / F::COL_HALF : 0;
    const int th = ctx.tid() - half * F::COL_HALF;
    const int col = th % CB;
    const int g = th / CB;
    cplx* ex = smem + F::NTAB_PAD + (size_t)half * (CB * F::Sh::SMEM_ELEMS) + col;
    const int rr = F::COL_DUAL ? half : ctx.by();
    constexpr int NBLK = M / CB;
    const int rc = ctx.bx() / NBLK;
    const int kc = (ctx.bx() - rc * NBLK) * CB + col;
    const SmemTw<M, PPT> tw{tab};
    const GroupSync<Ctx, 2> gs{ctx, 0, 0};

    // The accumulators start from the current intensity values: the read half of the plane's
    // read-modify-write is issued here and hidden behind the first FFT instead of stalling the epilogue.
    float* dst = P.ic + ((size_t)(rr * 2 + rc) * M + g) * M + kc;
    float acc[PPT];
#pragma unroll
    for (int e = 0; e < PPT; ++e) acc[e] = dst[(size_t)(TG * e) * M];

    for (int sl = 0; sl < P.batch; ++sl) {
        const cplx* src = P.T + ((size_t)(sl * 2 + rc) * P.Sr) * M + kc;
        cplx v[PPT];
        if (sl == 0) {
            // raw loads first, tables second: both are in flight together
            fast_col_load_raw<M, PPT, false>(v, src, P.Sr, g);
            fast_tables_wait(ctx);
            fast_col_finish<M, PPT, false>(v, src, P.Sr, rr, g, tab);
        } else {
            fast_col_load<M, PPT, false>(v, src, P.Sr, rr, g, tab);
        }
        fft_run<M, PPT, false>(v, ex, CB, g, tw, gs);
        const float w = P.weights ? P.weights[P.s_begin + sl] : 1.f;
#pragma unroll
        for (int e = 0; e < PPT; ++e) acc[e] += w * cnorm2(v[e]);
    }
    if (P.batch == 0) fast_tables_wait(ctx);
#pragma unroll
    for (int e = 0; e < PPT; ++e) dst[(size_t)(TG * e) * M] = acc[e];
}

// ----------------------------------------------------------------------------- TMA-staged column pass
// Same arithmetic and thread mapping as fast_cols_body (dual-residue CTA: threads [0,HALF) compute output-row
// residue rr = 0, [HALF,2*HALF) rr = 1 of the same CBT-column tile), but the T tile of source point sl+1 is copied
// global -> shared by the TMA engine (one elected thread issues <= 4 cp.async.bulk.tensor boxes of M/4 rows plus
// one 1-D bulk copy for the rim row u = M; all complete on one mbarrier) while the FFT of source point sl runs,
// and both residues read the staged tile from shared memory: the load latency leaves the critical path and the
// LSU issues one 256-byte-contiguous LDS per warp instead of 64-byte global segments.  The tile buffer is
// single: every thread moves its inputs to registers first, and the CTA barrier that precedes the first exchange
// of the FFT is also the point from which the buffer may be overwritten.
//
// CBT = columns per tile.  "Wide" tiles (CBT = FastShape::CB_DUAL: 8 at M = 1024) fill an SM with one 512-thread
// CTA; "narrow" tiles (half of that) need half the shared memory, so two independent 256-thread CTAs share an SM
// and the shared-memory bursts of one overlap the butterflies of the other.
//
// Shared-memory twiddle tables are compact here: pre[u] = w_2M^u only for u = 0..M/2 (the upper half follows
// from w_2M^(M-u) = -conj(w_2M^u)), then tw1, tw2 as in FastShape.
template <int M, int PPT, int CBT>
struct TmaShape {
    using F = FastShape<M, PPT>;
    using Sh = typename F::Sh;
    static constexpr int TG = F::TG;
    static constexpr int HALF = CBT * TG;
    static constexpr int THREADS = 2 * HALF;
    static constexpr int PRE_N = M / 2 + 1;
    static constexpr int TW1_OFF = (PRE_N + 1) & ~1;
    static constexpr int TW2_OFF = TW1_OFF + (F::R1 - 1) * F::NS1;
    static constexpr int NTAB = TW2_OFF + (F::R2 - 1) * F::NS2;      // device table: pre, tw1, tw2
    static constexpr int NTAB_PAD = (NTAB + 1) & ~1;
    // M = 4096: exchange buffer (135 KB) + tile (66 KB) + all tables (49 KB) exceed the 227 KB of a CTA.  The
    // third-pass table tw2 (3 x 1024 entries, indexed by k = thread-varying) then stays in global memory and is
    // read through L1 (same LSU cost as a shared-memory read); pre and tw1 stay in shared memory.
    static constexpr bool TW2_GLOBAL = (M >= 4096);
    static constexpr int NTAB_SMEM = TW2_GLOBAL ? TW2_OFF : NTAB;
    static constexpr int NTAB_SMEM_PAD = (NTAB_SMEM + 1) & ~1;
    static constexpr int BOX_ROWS = M / 4 > 256 ? 256 : M / 4;     // rows per TMA box (boxes of <= 256 rows cover u < M)
    static constexpr size_t EX_BYTES = (size_t)2 * CBT * Sh::SMEM_ELEMS * sizeof(cplx);
    static constexpr size_t TILE_OFF = ((size_t)NTAB_SMEM_PAD * sizeof(cplx) + EX_BYTES + 127) / 128 * 128;
    static constexpr size_t TILE_BYTES = ((size_t)(M + 1 + RIM_EXTRA) * CBT * sizeof(cplx) + 15) / 16 * 16;
    static constexpr size_t BAR_OFF = TILE_OFF + TILE_BYTES;
    static constexpr size_t SMEM = BAR_OFF + 48;  // mbarrier (8, padded to 16) + TileCtl
    // a box must be a multiple of 128 bytes (TMA destination alignment); the halves only meet at CTA barriers, so
    // any THREADS that is a multiple of 32 works
    static constexpr bool OK = PPT == 32 && M >= 32 && M <= 4096 && CBT >= 2 && (M % CBT) == 0 &&
                               ((size_t)BOX_ROWS * CBT * sizeof(cplx)) % 128 == 0 && THREADS % 32 == 0 &&
                               THREADS <= 512 && SMEM <= 227 * 1024;
    // two CTAs per SM when registers (128/thread) and shared memory (+1 KB reserved per CTA) allow
    static constexpr int MIN_BLOCKS = (2 * THREADS * 128 <= 65536 && 2 * (SMEM + 1024) <= 228 * 1024) ? 2 : 1;
};

template <int M, int PPT, int CBT>
struct TmaTw {
    const cplx* tab;    // shared memory: pre, tw1 (and tw2 unless TW2_GLOBAL)
    const cplx* gtab;   // the same table in global memory (device copy of the plan)
    template <int PASS>
    LITHO_HD cplx get(int t, int k) const {
        using S = TmaShape<M, PPT, CBT>;
        using F = FastShape<M, PPT>;
        if constexpr (PASS == 1) return tab[S::TW1_OFF + (t - 1) * F::NS1 + k];
        else if constexpr (S::TW2_GLOBAL) return ldg_c(gtab + S::TW2_OFF + (t - 1) * F::NS2 + k);
        else return tab[S::TW2_OFF + (t - 1) * F::NS2 + k];
    }
};

// Loop state of the tile copies lives in shared memory (only thread 0 touches it), so that it costs no
// registers in the FFT, whose 128-register budget is full.
struct TileCtl {
    int next_row;       // first tensor row of the next tile to fetch
    int rows_per_tile;  // tensor rows between consecutive source points (2*Sr)
    int remaining;      // tiles still to be fetched
    int col;            // first column of this CTA's tile
};

template <class Ctx>
struct TileHook {
    const Ctx& ctx;
    unsigned char* smem_raw;
    const FastColsParams* P;
    size_t tile_off, bar_off;
    int rim_elem;   // element offset of the rim row inside the tile buffer (M * CBT)
    LITHO_HD void after_last_gather() const {}
    LITHO_HD void after_first_sync() const {
        if (ctx.tid() == 0) {
            TileCtl* ctl = reinterpret_cast<TileCtl*>(smem_raw + bar_off + 16);
            const int rem = ctl->remaining;
            if (rem > 0) {
                const int row = ctl->next_row;
                ctx.tile_load(smem_raw + tile_off, P->tile, row, ctl->col, P->nbox, P->rim, rim_elem,
                              reinterpret_cast<unsigned long long*>(smem_raw + bar_off));
                ctl->next_row = row + ctl->rows_per_tile;
                ctl->remaining = rem - 1;
            }
        }
    }
};

// grid.x = 2*M/CBT column blocks (rc major), block = TmaShape::THREADS
template <int M, int PPT, int CBT, class Ctx>
LITHO_HD void fast_cols_tma_body(const FastColsParams& P, const Ctx& ctx, unsigned char* smem_raw) {
    using F = FastShape<M, PPT>;
    using S = TmaShape<M, PPT, CBT>;
    constexpr int TG = F::TG;
    cplx* smem = reinterpret_cast<cplx*>(smem_raw);
    cplx* tab = smem;
    const cplx* tile = reinterpret_cast<const cplx*>(smem_raw + S::TILE_OFF);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + S::BAR_OFF);
    for (int i = ctx.tid(); i < S::NTAB_SMEM_PAD / 2; i += ctx.bdim()) ctx.cp_async16(tab + 2 * i, P.tables_c + 2 * i);
    constexpr int NBLK = M / CBT;
    const int rc = ctx.bx() / NBLK;
    const int kc0 = (ctx.bx() - rc * NBLK) * CBT;
    if (ctx.tid() == 0) {
        ctx.mbar_init(bar, 1);
        TileCtl* ctl = reinterpret_cast<TileCtl*>(smem_raw + S::BAR_OFF + 16);
        ctl->next_row = (int)P.row_begin + rc * P.Sr;
        ctl->rows_per_tile = 2 * P.Sr;
        ctl->remaining = P.batch;
        ctl->col = kc0;
    }
    const int half = ctx.tid() / S::HALF;
    const int th = ctx.tid() - half * S::HALF;
    const int col = th % CBT;
    const int g = th / CBT;
    cplx* ex = smem + S::NTAB_SMEM_PAD + (size_t)half * (CBT * F::Sh::SMEM_ELEMS) + col;
    const int rr = half;
    const TmaTw<M, PPT, CBT> tw{tab, P.tables_c};
    const GroupSync<Ctx, 2> gs{ctx, 0, 0};
    const TileHook<Ctx> hook{ctx, smem_raw, &P, S::TILE_OFF, S::BAR_OFF, M * CBT};

    float* dst = P.ic + ((size_t)(rr * 2 + rc) * M + g) * M + kc0 + col;
    float acc[PPT];
#pragma unroll
    for (int e = 0; e < PPT; ++e) acc[e] = dst[(size_t)(TG * e) * M];

    fast_tables_wait(ctx);      // tables landed, barrier + loop state initialised (CTA barrier inside)
    hook.after_first_sync();    // fetch the first tile
    const int last = P.Sr - 1;

    for (int sl = 0; sl < P.batch; ++sl) {
        if (!ctx.mbar_wait(bar, (unsigned)(sl & 1)) && P.status) P.status[1] = 1;   // results invalid: reported, never silent
        cplx v[PPT];
        if (last >= M - 1) {
#pragma unroll
            for (int e = 0; e < PPT; ++e) v[e] = tile[th + e * (TG * CBT)];
        } else {
#pragma unroll
            for (int e = 0; e < PPT; ++e) {  // rows past Sr hold stale (finite or not) data: select, never multiply
                const cplx x = tile[th + e * (TG * CBT)];
                const bool in = g + TG * e <= last;
                v[e] = mk(in ? x.x : 0.f, in ? x.y : 0.f);
            }
        }
        if (P.Sr > M) {  // rim rows u = M .. Sr-1 fold onto slots u - M with (-1)^rr, before the slot's pre-twiddle
#pragma unroll
            for (int k = 0; k <= RIM_EXTRA; ++k) {
                if (k < P.Sr - M && g == k % TG) {
                    const cplx y = tile[(M + k) * CBT + col];
                    v[k / TG] = rr ? csub(v[k / TG], y) : cadd(v[k / TG], y);
                }
            }
        }
        if (rr) {  // pre-twiddle w_2M^u, u = g + TG*e; u > M/2 from the mirrored entry: w^u = -conj(w^(M-u))
#pragma unroll
            for (int e = 0; e < PPT; ++e) {
                if (TG * e + TG - 1 <= M / 2) {
                    v[e] = cmul(v[e], tab[g + TG * e]);
                } else {  // u >= M/2 (M/2 is a multiple of TG): index M-u is in [0, M/2]
                    const cplx t = tab[M - TG * e - g];
                    v[e] = cmul(v[e], mk(-t.x, t.y));
                }
            }
        }
        if constexpr (F::Sh::NP == 1) {  // single-pass FFT: no exchange, hence no barrier inside fft_run
            ctx.sync();
            hook.after_first_sync();
        }
        fft_run<M, PPT, false>(v, ex, CBT, g, tw, gs, hook);
        const float w = P.weights ? P.weights[P.s_begin + sl] : 1.f;
#pragma unroll
        for (int e = 0; e < PPT; ++e) acc[e] += w * cnorm2(v[e]);
    }
#pragma unroll
    for (int e = 0; e < PPT; ++e) dst[(size_t)(TG * e) * M] = acc[e];
}

// ----------------------------------------------------------------------------- rim lines
// F[d1][d2] = sum_s w_s sum_{u1-u2=d1, v1-v2=d2} G_s[u1][v1] conj(G_s[u2][v2]) are the Fourier coefficients of
// sum_s |E_s|^2.  Those with |d1| >= M or |d2| >= M alias on the Nc = 2M coarse grid; they only involve the few
// window rows (columns) within er = Sr-1-M (ec = Sc-1-M) of the edges, so they are summed directly here:
//   frow[k][n + Sc-1] = F[M+k][n],  k = 0..er, n = -(Sc-1)..Sc-1   (row pairs u2 = t, u1 = t + M + k, t = 0..er-k)
//   fcol[k][m + Sr-1] = F[m][M+k],  k = 0..ec, m = -(Sr-1)..Sr-1   (column pairs likewise)
// Negative frequencies follow from F[-d] = conj(F[d]).  `ext` = non-zero extents of the pupil along each of
// those lines (window coordinates), so that only the populated stretch is correlated:
//   ext[0][k] top row k, ext[1][k] bottom row Sr-1-k, ext[2][k] left column k, ext[3][k] right column Sc-1-k.
struct RimParams {
    const cplx* pupil;
    const cplx* mask;
    int pn, pr0, pc0, Sr, Sc, M;
    const int2_* shifts;
    const float* weights;
    int n_src;
    int per_cta;    // source points per CTA (CTA c takes [c*per_cta, (c+1)*per_cta))
    int ext[4][RIM_LINES][2];
    int er, ec;     // -1: no rim lines on that axis
    float* frow;    // this launch's partial sums: CTA c owns frow + c*stride, [er+1][2*Sc-1] complex, interleaved
    float* fcol;    // likewise, [ec+1][2*Sr-1] complex
    size_t stride;  // floats between the slices of consecutive CTAs (0: all CTAs share one slice -- single CTA only)
};

// One CTA per chunk of source points, taken in order; every partial sum has exactly one owner thread and is added
// to the CTA's private slice with plain read-modify-writes, so the result does not depend on scheduling.
// rim_reduce_body then folds the slices into the plane in chunk order.  smem holds the two vectors of the pair
// being correlated.
template <class Ctx>
LITHO_HD void rim_body(const RimParams& P, const Ctx& ctx, cplx* smem) {
    const int s_lo = ctx.bx() * P.per_cta;
    const int s_hi = (s_lo + P.per_cta) < P.n_src ? (s_lo + P.per_cta) : P.n_src;
    float* const frow = P.frow + (size_t)ctx.bx() * P.stride;
    float* const fcol = P.fcol + (size_t)ctx.bx() * P.stride;
    for (int s = s_lo; s < s_hi; ++s) {
    const int2_ sh = P.shifts[s];
    const float w = P.weights ? P.weights[s] : 1.f;
    for (int axis = 0; axis < 2; ++axis) {
        const int e = axis == 0 ? P.er : P.ec;
        const int Sfix = axis == 0 ? P.Sr : P.Sc;      // size along the axis the pair is separated on
        const int Slag = axis == 0 ? P.Sc : P.Sr;      // size along the lines
        float* F = axis == 0 ? frow : fcol;
        for (int k = 0; k <= e; ++k) {                 // frequency M + k
            for (int t = 0; t + k <= e; ++t) {         // pair: "lo" line t, "hi" line t + M + k = Sfix-1-b
                const int b = e - k - t;
                const int ll = P.ext[axis * 2][t][0], lh = P.ext[axis * 2][t][1];
                const int hl = P.ext[axis * 2 + 1][b][0], hh = P.ext[axis * 2 + 1][b][1];
                const int nl = lh - ll + 1, nh = hh - hl + 1;
                if (nl <= 0 || nh <= 0) continue;
                cplx* gh = smem;
                cplx* gl = smem + nh;
                ctx.sync();
                for (int i = ctx.tid(); i < nh + nl; i += ctx.bdim()) {
                    const bool hi = i < nh;
                    const int along = hi ? hl + i : ll + (i - nh);     // position along the line
                    const int fix = hi ? Sfix - 1 - b : t;             // the line's index on the other axis
                    const int u = axis == 0 ? fix : along;             // window row
                    const int v = axis == 0 ? along : fix;             // window column
                    const cplx p = P.pupil[(size_t)(P.pr0 + u) * P.pn + P.pc0 + v];
                    const cplx m = P.mask[(size_t)iclamp(P.pr0 + u + sh.x, 0, P.pn - 1) * P.pn +
                                          iclamp(P.pc0 + v + sh.y, 0, P.pn - 1)];
                    smem[i] = cmul(p, m);
                }
                ctx.sync();
                // correlation lags n = (hl + a) - (ll + c), a in [0,nh), c in [0,nl)
                const int nlag = nh + nl - 1;
                float* Fk = F + (size_t)2 * k * (2 * Slag - 1);
                for (int li = ctx.tid(); li < nlag; li += ctx.bdim()) {
                    const int d = li - (nl - 1);  // a - c
                    const int c0 = d < 0 ? -d : 0;
                    const int c1 = (nh - d) < nl ? (nh - d) : nl;
                    cplx acc = mk(0.f, 0.f);
                    for (int c = c0; c < c1; ++c) acc = cadd(acc, cmul(gh[c + d], cconj(gl[c])));
                    const int n = (hl - ll) + d;
                    // sole owner of this entry within the CTA (pairs and source points are taken in sequence,
                    // separated by the CTA barriers above)
                    Fk[2 * (n + Slag - 1)] += w * acc.x;
                    Fk[2 * (n + Slag - 1) + 1] += w * acc.y;
                }
            }
        }
    }
    }
}

// plane[j] += slices[0][j] + slices[1][j] + ... in slice order (one thread per float)
struct RimReduceParams {
    const float* slices;
    size_t stride;
    int n_slices;
    int n;          // floats per slice that are summed
    float* plane;
};
LITHO_HD void rim_reduce_elem(const RimReduceParams& P, int j) {
    float acc = 0.f;
    for (int c = 0; c < P.n_slices; ++c) acc += P.slices[(size_t)c * P.stride + j];
    P.plane[j] += acc;
}

// ----------------------------------------------------------------------------- coarse -> fine
// Spectrum assembly.  fhat is the centred Nc x Nc DFT of the coarse intensity (index m+K, K = Nc/2 = M, m in
// [-K,K)): fhat[m][n] = sum of F over all (m', n') congruent to (m, n) modulo Nc.  out is the true spectrum on
// m in [-K-Er, K+Er], n in [-K-Ec, K+Ec] (Er = max(er,0)): entries with |m| >= K or |n| >= K come from the rim
// sums, interior entries are fhat minus their aliased partners (which are rim entries).
struct AssembleParams {
    const cplx* fhat;   // [Nc][Nc]
    const float* frow;  // [er+1][2*Sc-1]
    const float* fcol;  // [ec+1][2*Sr-1]
    int Nc, er, ec, Sr, Sc;
    cplx* out;          // [Nc+1+2*Er][Nc+1+2*Ec]
};

LITHO_HD cplx rim_get(const float* f, size_t idx) { return mk(f[2 * idx], f[2 * idx + 1]); }

// F[m][n] for |m| >= K or |n| >= K (zero outside the extent of the autocorrelation)
LITHO_HD cplx rim_value(const AssembleParams& P, int m, int n) {
    const int K = P.Nc / 2;
    const int am = m < 0 ? -m : m, an = n < 0 ? -n : n;
    if (am >= K) {
        const int k = am - K;
        if (k > P.er || an > P.Sc - 1) return mk(0.f, 0.f);
        const int nn = m > 0 ? n : -n;
        const cplx v = rim_get(P.frow, (size_t)k * (2 * P.Sc - 1) + (nn + P.Sc - 1));
        return m > 0 ? v : cconj(v);
    }
    const int k = an - K;
    if (k > P.ec || am > P.Sr - 1) return mk(0.f, 0.f);
    const int mm = n > 0 ? m : -m;
    const cplx v = rim_get(P.fcol, (size_t)k * (2 * P.Sr - 1) + (mm + P.Sr - 1));
    return n > 0 ? v : cconj(v);
}

LITHO_HD void assemble_elem(const AssembleParams& P, int i, int j) {
    const int K = P.Nc / 2;
    const int Er = P.er > 0 ? P.er : 0, Ec = P.ec > 0 ? P.ec : 0;
    const int m = i - K - Er, n = j - K - Ec;
    const int am = m < 0 ? -m : m, an = n < 0 ? -n : n;
    cplx val;
    if (am >= K || an >= K) {
        val = rim_value(P, m, n);
    } else {
        val = P.fhat[(size_t)(m + K) * P.Nc + (n + K)];
        // aliased partners: m +- Nc, n +- Nc inside the extent
        const int mp = (m + P.Nc <= K + P.er) ? m + P.Nc : ((m - P.Nc >= -K - P.er) ? m - P.Nc : m);
        const int np = (n + P.Nc <= K + P.ec) ? n + P.Nc : ((n - P.Nc >= -K - P.ec) ? n - P.Nc : n);
        if (mp != m) val = csub(val, rim_value(P, mp, n));
        if (np != n) val = csub(val, rim_value(P, m, np));
        if (mp != m && np != n) val = csub(val, rim_value(P, mp, np));
    }
    P.out[(size_t)i * (P.Nc + 1 + 2 * Ec) + j] = val;
}

}  // namespace litho
