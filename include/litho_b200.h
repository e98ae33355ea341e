/* litho_b200.h -- C ABI of the B200-native Abbe imaging hot path.
 *
 * Drop-in boundary for the partially coherent aerial-image path of
 * quarterwave0/LithographySimulator.  The reference has no FFI layer of its own: its
 * boundary is the Python call surface.  Every entry point below
 * names the reference interface it replaces (file:line relative to the reference root);
 * lithographysimulator_b200/imageformation.py binds them with ctypes and mirrors the
 * reference signatures (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers on the current CUDA device unless named *_host;
 *   - planes are row-major; complex64 is interleaved (re, im) float pairs;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*) unless stated;
 *   - the caller owns every buffer including the workspace; a plan owns only its twiddle table;
 *   - return 0 on success, non-zero on error with a message in litho_last_error();
 *   - thread safety: entry points are re-entrant for DIFFERENT plans; a plan carries per-call state of its
 *     own (its T-ring events, an auxiliary stream, error words) and must be used by one stream / one host
 *     thread at a time -- create one plan per concurrent stream.
 */
#ifndef LITHO_B200_H
#define LITHO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LITHO_OK 0
#define LITHO_ERR_ARG 1      /* invalid argument / unsupported configuration */
#define LITHO_ERR_CUDA 2     /* CUDA runtime error */
#define LITHO_ERR_WORKSPACE 3

#define LITHO_ABI_VERSION 1

typedef struct litho_plan litho_plan_t;

typedef struct litho_plan_info {
    int pn, N;            /* grid side, FFT-approximation length (mask.py:67-72) */
    int bbox[4];          /* pupil support r0,r1,c0,c1 (inclusive) */
    int L, M, R, Wr;      /* transform length, sub-FFT length, residues, per-residue pitch */
    int path;             /* 1 = generic fine-grid kernels (any source, any pupil);
                             2 = fast coarse-grid kernels (need every shift inside shift_range) */
    int default_batch;    /* source points per launch pair used when batch <= 0 */
    uint64_t intensity_elems; /* float elements of the plan's intensity accumulator */
    int shift_range[4];   /* d0 min,max, d1 min,max for which roll() does not wrap the pupil window */
} litho_plan_info_t;

/* litho_plan_create flags */
#define LITHO_PLAN_GENERIC 1  /* force path 1 (required when a source point wraps the pupil window) */

int litho_abi_version(void);
const char* litho_last_error(void);
/* 1 if this library was built for the GPU (sm_100a), 0 for the CPU emulation used by tests */
int litho_is_device_build(void);

/* Mask.calculateEpsilonN / _nearest2SqInt                           mask.py:63-72
 * host-only helper: beta = wavelength/(deltaK*pixelSize), N = nearest power of two, eps = N/beta */
int litho_epsilon_n(double deltaK, double pixelSize, double wavelength, double* eps, int* N);

/* Bounding box of the non-zero pupil samples (the support that roll(), imageformation.py:63,
 * moves around the grid).  Synchronises `stream`; bbox_host = {r0,r1,c0,c1}, {0,-1,0,-1} if empty. */
int litho_pupil_bbox(const void* pupil, int pn, int* bbox_host, void* stream);

/* Support analysis used to plan the fast path: support_host[0..3] = bbox as above, [4..11] = non-zero
 * extents {cmin,cmax} of the first and last bbox row and {rmin,rmax} of the first and last bbox column
 * (the pupil's rim pixels, which carry the one frequency line the coarse grid aliases).  Synchronises. */
int litho_pupil_support(const void* pupil, int pn, int* support_host, void* stream);
/* Same for the `lines` (1..LITHO_RIM_LINES) outermost bbox rows and columns: support_host[4 + 8k ..] = extents of
 * row r0+k, row r1-k, column c0+k, column c1-k (4 + 8*lines ints in all; an empty line has lo > hi).  The
 * reference's fp16 pupil grid makes the support pn/2+3 wide at pn = 8192: the fast path then folds the two extra
 * rows/columns and needs the extents of three lines per side to keep its rim sums cheap. */
#define LITHO_RIM_LINES 3
int litho_pupil_support_lines(const void* pupil, int pn, int lines, int* support_host, void* stream);

/* min/max of the source shifts: bounds_host = {d0 min, d0 max, d1 min, d1 max}.  Synchronises.
 * A fast plan (path 2) may only be used when these lie inside plan_info.shift_range. */
int litho_shift_bounds(const int32_t* shifts, int n_src, int* bounds_host, void* stream);

/* Source points of a light-source plane in two short launches         imageformation.py:59-60
 *   shifts[k] = (row, col) - pn//2 of the non-zero elements of `lightsource` (pn x pn elements of elem_size 1/2/4/8
 *   bytes; is_float: -0.0 counts as zero, as in torch.argwhere), in row-major order -- the reference's loop order.
 * rank/world select the interleaved shard (point number o belongs to rank o % world, stored at index o / world;
 * 0/1 for all points); at most `capacity` points are stored.  meta_host receives {n_all, n_mine, min d0, max d0,
 * min d1, max d1} (bounds over ALL points, the argument of a fast plan's no-wrap check).  Synchronises `stream`.
 * Replaces the ~10 small kernels of the torch op sequence when images are staged back to back. */
int litho_source_points(const void* lightsource, int elem_size, int is_float, int pn, int rank, int world,
                        int32_t* shifts, int capacity, int* meta_host, void* stream);

/* Plan for abbeImage(fft=True) on a pn x pn grid with FFT-approximation length N.
 * litho_plan_create takes the 4-int bbox; litho_plan_create_ex the 12-int support of
 * litho_pupil_support (cheaper rim sums).  flags: 0 or LITHO_PLAN_GENERIC.
 * litho_plan_create_lines takes the support of litho_pupil_support_lines (lines = 0: bbox only).
 * Path 2 is chosen when the window fits S <= M + LITHO_RIM_LINES, M a power of two <= 4096, and 2M <= N. */
int litho_plan_create(int pn, int N, const int* bbox, int flags, litho_plan_t** plan);
int litho_plan_create_ex(int pn, int N, const int* support, int flags, litho_plan_t** plan);
int litho_plan_create_lines(int pn, int N, const int* support, int lines, int flags, litho_plan_t** plan);
void litho_plan_destroy(litho_plan_t* plan);
int litho_plan_get_info(const litho_plan_t* plan, litho_plan_info_t* info);
size_t litho_plan_workspace_bytes(const litho_plan_t* plan, int batch);
/* Sticky error words of a fast plan, read and cleared (synchronises `stream`): status_host[0] != 0: a source shift lay
 * outside plan_info.shift_range and was clamped (the images accumulated since the last query are WRONG: use a
 * LITHO_PLAN_GENERIC plan for sources that wrap the pupil window); status_host[1] != 0: a TMA tile copy of the
 * column pass never completed (results invalid).  Both stay 0 in correct use. */
int litho_plan_status(const litho_plan_t* plan, int* status_host, void* stream);
/* Columns per tile of the TMA-staged column-pass kernel this plan launches (the T tile of the next source
 * point is copied global -> shared by cp.async.bulk.tensor while the current FFT runs); 0 when the plan
 * uses the plain-load column kernel (generic path, LITHO_TMA=0, or a driver without
 * cuTensorMapEncodeTiled). */
int litho_plan_column_tile(const litho_plan_t* plan);

/* The hot loop of abbeImage                                  imageformation.py:59-67
 *   intensity += sum_s w_s * | centred zoom IDFT_N { roll(pupil, shift_s) * maskFT } |^2
 * shifts: n_src (d0,d1) int32 pairs = argwhere(lightsource) - pn//2 (imageformation.py:59);
 * weights: n_src floats or NULL (the reference ignores source values: all ones);
 * intensity: plan.intensity_elems floats in the plan's residue-major order, accumulated into
 *            (zero it before the first call; partial planes from several GPUs may simply be summed). */
int litho_abbe_fft_accumulate(const litho_plan_t* plan, const void* maskFT, const void* pupil,
                              const int32_t* shifts, const float* weights, int n_src, int batch,
                              float* intensity, void* workspace, size_t workspace_bytes, void* stream);

/* Same as litho_abbe_fft_accumulate with a phase mask: bit 0 = row pass, bit 1 = column pass (3 = both).
 * Running one pass alone re-uses whatever T holds, so results are only meaningful with both bits set;
 * bench.py uses the single-pass forms to time each kernel live.
 * LITHO_PHASE_INPUTS_READY (bit 2): maskFT, pupil and shifts are already valid on the device when the call is
 * made (not produced by work still queued on `stream`).  The row pass of this call may then start while
 * earlier work on `stream` (typically the last column pass of the previous image on the same plan and
 * workspace) is still running; the column passes and the intensity plane stay ordered on `stream`.
 * The flag also promises that NOTHING ELSE has used `workspace` since this plan's previous accumulate call on it
 * (give every plan that is driven this way a workspace of its own, as AbbeEngine does): the row pass of this
 * call only waits for this plan's own column passes that last read each T-ring slot. */
#define LITHO_PHASE_INPUTS_READY 4
int litho_abbe_fft_accumulate_ex(const litho_plan_t* plan, const void* maskFT, const void* pupil,
                                 const int32_t* shifts, const float* weights, int n_src, int batch,
                                 float* intensity, void* workspace, size_t workspace_bytes, void* stream,
                                 int phases);

/* Focus batching (BASELINE cfg5: one mask, one source, 16 defocus values): the hot loop of abbeImage for
 * n_focus pupil functions at once -- what the reference does with n_focus calls of abbeImage (imageformation.py:47-77)
 * on pupils from Pupil.generatePupilFunction (pupil.py:32-35, defocus in aberrations[4], pupil.py:88-100).
 *   pupils:      n_focus planes of pn x pn complex64, pupil_stride ELEMENTS apart; every pupil's support must lie in
 *                the plan's window and share its rim extents (true for focus/aberration variants of one pupil:
 *                |P| = 1 on the same disc)
 *   intensities: n_focus planes of plan.intensity_elems floats, intensity_stride ELEMENTS apart, accumulated into
 * One row pass serves all focus values: the work items of a (source point, window row) are adjacent for every focus
 * value, so the shifted mask-spectrum row crosses HBM/L2 once and is multiplied by n_focus pupil rows; then one column
 * pass per focus value.  Generic plans fall back to one focus value at a time.  batch <= 0: plan default / n_focus.
 * Results are bit-identical to n_focus separate litho_abbe_fft_accumulate calls. */
size_t litho_plan_workspace_bytes_focus(const litho_plan_t* plan, int batch, int n_focus);
int litho_abbe_fft_accumulate_focus(const litho_plan_t* plan, const void* maskFT, const void* pupils, int n_focus,
                                    size_t pupil_stride, const int32_t* shifts, const float* weights, int n_src,
                                    int batch, float* intensities, size_t intensity_stride, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Post-processing of abbeImage(fft=True)                      imageformation.py:69-75
 * abs -> bilinear resample by 1/eps -> zero border; out has litho_fft_output_side(pn,eps)^2 floats. */
int litho_fft_output_side(int pn, double eps);
/* workspace for finalize / unpermute: staging of the coarse->fine spectral interpolation (path 2) */
size_t litho_plan_finalize_workspace_bytes(const litho_plan_t* plan);
int litho_abbe_fft_finalize(const litho_plan_t* plan, const float* intensity, double eps, float* out,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Raw accumulated intensity in natural row-major order (pn x pn), no resampling. */
int litho_abbe_fft_unpermute(const litho_plan_t* plan, const float* intensity, float* out, void* workspace,
                             size_t workspace_bytes, void* stream);

/* calculateFFTAerial(pf, maskFFFT, pixelNumber, N)              imageformation.py:32-45
 * complex field (pn x pn complex64) of one already-shifted pupil `pf`; the plan must have been
 * created from the bounding box of `pf`. */
int litho_fft_field(const litho_plan_t* plan, const void* pf, const void* maskFT, void* field,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Mask._ffFraunhofer(epsilon, N)                                    mask.py:74-90
 * mask spectrum by the FFT approximation: geometry (pn x pn int16) -> bilinear upsample by eps ->
 * centre-pad to N -> centred forward DFT -> crop; maskFT receives pn x pn complex64. */
size_t litho_mask_spectrum_workspace_bytes(int pn, double eps, int N);
int litho_mask_spectrum(const int16_t* geometry, int pn, double eps, int N, void* maskFT, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- direct ("Abbe") solver: E = A * G * A^T with the fp16-quantised phase table ----
 * litho_direct_operator builds A[a][c] = w[c] * exp(sign*i*(2*pi/lambda)*fp16(fp16(k[a])*fp16(x[c])))
 * (pn x pn complex64): sign = -1 for imaging (imageformation.py:52), +1 for the mask spectrum (mask.py:42). */
int litho_direct_operator(int pn, double pixelSize, double wavelength, int sign, void* A, void* stream);
/* source points per launch group used when batch <= 0 (FP32 kernels: 8; tensor-core kernels: enough to fill the SMs) */
int litho_direct_default_batch(int pn, const int* bbox);
size_t litho_direct_workspace_bytes(int pn, const int* bbox, int batch);
/* abbeImage(fft=False) hot loop                                 imageformation.py:59-65 with :3-30
 *   intensity[pn][pn] += sum_s w_s | A (roll(pupil, shift_s) * maskFT) A^T |^2   (natural row-major order)
 * pn >= 64: both products run on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 hi/lo split, fp32 accumulation in
 * tensor memory; measured 1e-6 rel-L2 against float64) -- LITHO_DIRECT_TC=0 selects the FP32 CUDA-core kernels,
 * which also serve smaller grids.  batch <= 0: litho_direct_default_batch source points per launch group. */
int litho_direct_accumulate(const void* A, const void* maskFT, const void* pupil, int pn, const int* bbox,
                            const int32_t* shifts, const float* weights, int n_src, int batch, float* intensity,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Sticky error word of the tensor-core path of litho_direct_accumulate, read and cleared (synchronises): non-zero
 * if an MMA-completion wait timed out (results invalid).  Stays 0 in correct use; always 0 in the FP32 path. */
int litho_direct_status(int* status_host, void* stream);
/* calculateAerial(pupil, maskFT, ...)                            imageformation.py:3-30 */
int litho_direct_field(const void* A, const void* pupil, const void* maskFT, int pn, const int* bbox, void* field,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Mask.fraunhofer(wavelength, fft=False)                         mask.py:41-61  (A built with sign = +1) */
int litho_direct_mask_spectrum(const void* Aplus, const int16_t* geometry, int pn, void* maskFT, void* workspace,
                               size_t workspace_bytes, void* stream);

/* LightSource.generateAnnular / generateQuasar                    lightsource.py:34-73
 * fp16 sigma grid replayed op by op; out = pn x pn int64 0/1.  quasar_count = 0 selects the annulus. */
int litho_source_build(int pn, double sigma_in, double sigma_out, double shift_x, double shift_y, int quasar_count,
                       double rotation, int64_t* out, void* stream);

/* generateWavefrontError + generatePhi                                pupil.py:46-111
 * aberrations_host: n_ab OSA-indexed coefficients (HOST floats holding fp16 values, AFTER the caller applied
 * the reference's in-place defocus rescale aberrations[4] *= NA^2/(4*lambda), pupil.py:91-92).
 * pupil / wavefront: pn x pn complex64 outputs, either may be NULL.  Synchronises `stream`. */
int litho_pupil_build(const float* aberrations_host, int n_ab, int pn, void* pupil, void* wavefront, void* stream);

/* ---- multi-GPU: sum of the partial intensity planes over peer memory (one process per GPU, one box) ----
 * The reference's source loop is additive (imageformation.py:62-67), so ranks that each accumulate a shard of the
 * source points hold partial planes whose sum is the image.  Instead of a collective that every GPU must co-schedule,
 * the rank that post-processes an image reads the other ranks' planes directly over NVLink (CUDA IPC mappings) and
 * sums them in rank order -- deterministic -- in one kernel.  Cross-process ordering uses 64-bit sequence flags that
 * live in the same peer-mapped buffers (st.release.sys / ld.acquire.sys).
 *   litho_peer_alloc   device buffer (zeroed) that other processes of this box can map; handle = LITHO_PEER_HANDLE_BYTES
 *                      opaque bytes to send them (e.g. torch.distributed.all_gather_object)
 *   litho_peer_open    map another process's buffer here; litho_peer_close unmaps; litho_peer_free frees an owned one
 *   litho_peer_signal  after the work queued on `stream`: *flags[i] = value for i < n (flags may be remote)
 *   litho_peer_wait    block `stream` until flags[i] >= value for all i < n (flags: n consecutive uint64 in LOCAL
 *                      memory).  Both use stream memory operations (cuStreamWriteValue64 / cuStreamWaitValue64: no
 *                      kernel, so nothing has to find a free SM between persistent compute kernels) and fall back to
 *                      one-warp kernels (LITHO_PEER_MEMOPS=0); the kernel form of the wait gives up after ~20 s and
 *                      sets *err (device int, may be NULL) to 1, the memory-operation form blocks until signalled
 *   litho_peer_sum     out[e] = planes[0][e] + ... + planes[n-1][e], e < elems (16-byte aligned pointers; planes may
 *                      be remote).  flags/value (optional): every CTA re-acquires flags[r] >= value before loading. */
#define LITHO_MAX_PEERS 16
#define LITHO_PEER_HANDLE_BYTES 64
int litho_peer_alloc(size_t bytes, void** ptr, unsigned char* handle);
int litho_peer_open(const unsigned char* handle, void** ptr);
int litho_peer_close(void* ptr);
int litho_peer_free(void* ptr);
int litho_peer_signal(void* const* flags, int n, uint64_t value, void* stream);
int litho_peer_wait(const void* flags, int n, uint64_t value, int* err, void* stream);
/* dst <- src (either may be a mapped peer buffer), `bytes` bytes, by the copy engines (no kernel) after the work
 * queued on `stream` */
int litho_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
int litho_peer_sum(float* out, const float* const* planes, int n, uint64_t elems, const void* flags, uint64_t value,
                   int* err, void* stream);
const char* litho_peer_last_error(void);

/* FP32 FMA throughput probe (roofline denominator measured in the same run): launches `blocks` CTAs
 * of 256 threads, each thread doing iters*16 dependent-chain FMAs; *flops receives the flop count. */
int litho_fp32_probe(float* out, int blocks, int iters, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LITHO_B200_H */
