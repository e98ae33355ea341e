// Multi-GPU sum of the partial intensity planes over peer memory (the reference's additive
// source loop, imageformation.py:62-67, is what makes the planes summable).
//
// One process per GPU.  Every rank keeps its partial planes in a buffer that the other ranks of the box map through
// CUDA IPC; the rank that post-processes image i (the "root" of that image) sums the planes of all ranks with plain
// loads over NVLink/NVSwitch in one kernel, in rank order (deterministic), straight into the buffer its
// post-processing reads.  No collective kernel has to be co-scheduled on every GPU: the other ranks only publish
// a sequence number when their accumulation is done and go on with the next image.  Ordering between processes is
// carried by 64-bit sequence flags in the same peer-mapped buffers:
//     producer:  (accumulation kernels) -> flag = seq     (cuStreamWriteValue64, or peer_signal_kernel: st.release.sys)
//     consumer:  wait flag >= seq                         (cuStreamWaitValue64, or peer_wait_kernel: ld.acquire.sys)
//                -> peer_sum_kernel (each CTA re-acquires the flags, then loads)
// The kernel form of a wait gives up after ~20 s and raises an error word instead of hanging the GPU.
//
// With -DLITHO_EMU the same entry points work between PROCESSES of one host through POSIX shared memory, so that
// the protocol is covered by the world-size-2 gloo tests (tests only).
#include "../../include/litho_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#if defined(LITHO_EMU)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#else
#include <cuda_runtime.h>
#endif

static thread_local std::string g_perr;
extern "C" const char* litho_peer_last_error(void) { return g_perr.c_str(); }
static int pfail(int code, const std::string& msg) {
    g_perr = msg;
    return code;
}

#if !defined(LITHO_EMU)
namespace {

struct PeerPtrs {
    void* p[LITHO_MAX_PEERS];
    int n;
};

__global__ void peer_signal_kernel(const __grid_constant__ PeerPtrs P, unsigned long long value) {
    const int i = threadIdx.x;
    if (i < P.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.p[i]), "l"(value) : "memory");
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one thread per flag; flags live in THIS GPU's memory (the producers write them remotely)
__global__ void peer_wait_kernel(const unsigned long long* flags, int n, unsigned long long value, int* err) {
    const int i = threadIdx.x;
    if (i >= n) return;
    unsigned long long v = ld_acquire_sys(flags + i);
    long spins = 0;
    while (v < value) {
        __nanosleep(spins < 2000 ? 100 : 1000);
        if (++spins > 20000000L) {   // ~20 s: never hang the GPU on a lost peer
            if (err) *err = 1;
            break;
        }
        v = ld_acquire_sys(flags + i);
    }
}

// out[i] = planes[0][i] + planes[1][i] + ... (rank order), 16-byte loads that bypass L1 (peer data)
__global__ void __launch_bounds__(512) peer_sum_kernel(float* __restrict__ out, const __grid_constant__ PeerPtrs P,
                                                       unsigned long long elems, const unsigned long long* flags,
                                                       unsigned long long value, int* err) {
    if (flags) {   // (re-)acquire in this CTA: everything the producers published before their release is visible
        if (threadIdx.x == 0) {
            int good = 1;
            for (int r = 0; r < P.n; ++r)
                if (ld_acquire_sys(flags + r) < value) good = 0;
            if (!good && err) *err = 2;   // the wait kernel timed out before us
        }
        __syncthreads();
    }
    const unsigned long long n4 = elems / 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 acc = __ldcg(reinterpret_cast<const float4*>(P.p[0]) + i);
#pragma unroll 4
        for (int r = 1; r < P.n; ++r) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(P.p[r]) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x < (elems & 3)) {
        const unsigned long long i = n4 * 4 + threadIdx.x;
        float acc = __ldcg(reinterpret_cast<const float*>(P.p[0]) + i);
        for (int r = 1; r < P.n; ++r) acc += __ldcg(reinterpret_cast<const float*>(P.p[r]) + i);
        out[i] = acc;
    }
}

}  // namespace

// Stream memory operations (cuStreamWriteValue64 / cuStreamWaitValue64): the flag is written / polled by the
// stream's front end, not by a kernel.  That matters here: the flags are signalled and awaited between persistent
// compute kernels that hold every register of every SM, where even a one-warp kernel has to wait for a CTA of
// theirs to retire (~0.2 ms per tiny kernel at cfg3).  LITHO_PEER_MEMOPS=0 selects the kernels.
#include <cuda.h>
typedef CUresult (*write_value64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
typedef CUresult (*wait_value64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
static write_value64_fn g_write64 = nullptr;
static wait_value64_fn g_wait64 = nullptr;
static bool memops_available() {
    static const bool ok = []() -> bool {
        const char* env = getenv("LITHO_PEER_MEMOPS");
        if (env && atoi(env) == 0) return false;
        void* p1 = nullptr;
        void* p2 = nullptr;
        cudaDriverEntryPointQueryResult q1, q2;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &p1, cudaEnableDefault, &q1) != cudaSuccess ||
            q1 != cudaDriverEntryPointSuccess || !p1)
            return false;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p2, cudaEnableDefault, &q2) != cudaSuccess ||
            q2 != cudaDriverEntryPointSuccess || !p2)
            return false;
        g_write64 = (write_value64_fn)p1;
        g_wait64 = (wait_value64_fn)p2;
        return true;
    }();
    return ok;
}

#define PCHECK(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return pfail(LITHO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#else
namespace {
struct EmuSeg {
    size_t bytes;
    std::string name;
    bool owner;
};
std::mutex g_seg_mutex;
std::map<void*, EmuSeg> g_segs;
int g_seg_counter = 0;
}  // namespace
#endif

extern "C" {

int litho_peer_alloc(size_t bytes, void** ptr, unsigned char* handle) {
    if (!ptr || !handle || bytes == 0) return pfail(LITHO_ERR_ARG, "peer_alloc: bad argument");
    memset(handle, 0, LITHO_PEER_HANDLE_BYTES);
#if defined(LITHO_EMU)
    char name[48];
    {
        std::lock_guard<std::mutex> lock(g_seg_mutex);
        snprintf(name, sizeof(name), "/litho_peer_%d_%d", (int)getpid(), g_seg_counter++);
    }
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return pfail(LITHO_ERR_CUDA, "peer_alloc: shm_open failed");
    if (ftruncate(fd, (off_t)bytes) != 0) {
        close(fd);
        shm_unlink(name);
        return pfail(LITHO_ERR_CUDA, "peer_alloc: ftruncate failed");
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        shm_unlink(name);
        return pfail(LITHO_ERR_CUDA, "peer_alloc: mmap failed");
    }
    memset(p, 0, bytes);
    memcpy(handle, &bytes, sizeof(size_t));
    strncpy((char*)handle + 8, name, LITHO_PEER_HANDLE_BYTES - 9);
    {
        std::lock_guard<std::mutex> lock(g_seg_mutex);
        g_segs[p] = EmuSeg{bytes, name, true};
    }
    *ptr = p;
#else
    void* p = nullptr;
    PCHECK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return pfail(LITHO_ERR_CUDA, std::string("peer_alloc: ") + cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) <= LITHO_PEER_HANDLE_BYTES, "IPC handle fits the ABI handle");
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
#endif
    return LITHO_OK;
}

int litho_peer_open(const unsigned char* handle, void** ptr) {
    if (!ptr || !handle) return pfail(LITHO_ERR_ARG, "peer_open: bad argument");
#if defined(LITHO_EMU)
    size_t bytes = 0;
    memcpy(&bytes, handle, sizeof(size_t));
    const char* name = (const char*)handle + 8;
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return pfail(LITHO_ERR_CUDA, "peer_open: shm_open failed");
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return pfail(LITHO_ERR_CUDA, "peer_open: mmap failed");
    {
        std::lock_guard<std::mutex> lock(g_seg_mutex);
        g_segs[p] = EmuSeg{bytes, name, false};
    }
    *ptr = p;
#else
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    PCHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
#endif
    return LITHO_OK;
}

int litho_peer_close(void* ptr) {
    if (!ptr) return LITHO_OK;
#if defined(LITHO_EMU)
    std::lock_guard<std::mutex> lock(g_seg_mutex);
    auto it = g_segs.find(ptr);
    if (it == g_segs.end() || it->second.owner) return pfail(LITHO_ERR_ARG, "peer_close: not an opened peer buffer");
    munmap(ptr, it->second.bytes);
    g_segs.erase(it);
#else
    PCHECK(cudaIpcCloseMemHandle(ptr));
#endif
    return LITHO_OK;
}

int litho_peer_free(void* ptr) {
    if (!ptr) return LITHO_OK;
#if defined(LITHO_EMU)
    std::lock_guard<std::mutex> lock(g_seg_mutex);
    auto it = g_segs.find(ptr);
    if (it == g_segs.end() || !it->second.owner) return pfail(LITHO_ERR_ARG, "peer_free: not an owned peer buffer");
    munmap(ptr, it->second.bytes);
    shm_unlink(it->second.name.c_str());
    g_segs.erase(it);
#else
    PCHECK(cudaFree(ptr));
#endif
    return LITHO_OK;
}

int litho_peer_signal(void* const* flags, int n, uint64_t value, void* stream) {
    if (!flags || n < 1 || n > LITHO_MAX_PEERS) return pfail(LITHO_ERR_ARG, "peer_signal: bad argument");
#if defined(LITHO_EMU)
    (void)stream;
    for (int i = 0; i < n; ++i) __atomic_store_n((uint64_t*)flags[i], value, __ATOMIC_RELEASE);
#else
    if (memops_available()) {
        bool all = true;
        for (int i = 0; i < n && all; ++i)   // default flags: ordered after prior work of the stream, with a memory barrier
            all = g_write64((CUstream)stream, (CUdeviceptr)(uintptr_t)flags[i], (cuuint64_t)value, 0) == CUDA_SUCCESS;
        if (all) return LITHO_OK;
        // (an address the driver refuses: fall through to the kernel, which rewrites every flag)
    }
    PeerPtrs P;
    memset(&P, 0, sizeof(P));
    P.n = n;
    for (int i = 0; i < n; ++i) P.p[i] = flags[i];
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, (unsigned long long)value);
    PCHECK(cudaGetLastError());
#endif
    return LITHO_OK;
}

int litho_peer_wait(const void* flags, int n, uint64_t value, int* err, void* stream) {
    if (!flags || n < 1 || n > LITHO_MAX_PEERS) return pfail(LITHO_ERR_ARG, "peer_wait: bad argument");
#if defined(LITHO_EMU)
    (void)stream;
    const uint64_t* f = (const uint64_t*)flags;
    for (int i = 0; i < n; ++i) {
        long spins = 0;
        while (__atomic_load_n(f + i, __ATOMIC_ACQUIRE) < value) {
            struct timespec ts = {0, 200000};
            nanosleep(&ts, nullptr);
            if (++spins > 100000L) {   // 20 s
                if (err) *err = 1;
                break;
            }
        }
    }
#else
    if (memops_available() && value > 0) {
        bool all = true;
        for (int i = 0; i < n && all; ++i)
            all = g_wait64((CUstream)stream, (CUdeviceptr)((uintptr_t)flags + 8u * (unsigned)i), (cuuint64_t)value,
                           CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
        if (all) return LITHO_OK;   // (no time-out in this form: a lost peer blocks the stream)
    }
    if (value == 0) return LITHO_OK;   // flags start at 0
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)flags, n, (unsigned long long)value, err);
    PCHECK(cudaGetLastError());
#endif
    return LITHO_OK;
}

int litho_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
    if (!dst || !src) return pfail(LITHO_ERR_ARG, "peer_copy: null argument");
    if (bytes == 0) return LITHO_OK;
#if defined(LITHO_EMU)
    (void)stream;
    memcpy(dst, src, bytes);
#else
    // device-to-device over NVLink by the copy engines: no kernel, nothing competes with compute for SMs
    PCHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
#endif
    return LITHO_OK;
}

int litho_peer_sum(float* out, const float* const* planes, int n, uint64_t elems, const void* flags, uint64_t value,
                   int* err, void* stream) {
    if (!out || !planes || n < 1 || n > LITHO_MAX_PEERS) return pfail(LITHO_ERR_ARG, "peer_sum: bad argument");
    for (int r = 0; r < n; ++r)
        if (!planes[r]) return pfail(LITHO_ERR_ARG, "peer_sum: null plane");
    if (elems == 0) return LITHO_OK;
#if defined(LITHO_EMU)
    (void)stream;
    if (flags) {
        const uint64_t* f = (const uint64_t*)flags;
        for (int r = 0; r < n; ++r)
            if (__atomic_load_n(f + r, __ATOMIC_ACQUIRE) < value && err) *err = 2;
    }
    for (uint64_t i = 0; i < elems; ++i) {
        float acc = planes[0][i];
        for (int r = 1; r < n; ++r) acc += planes[r][i];
        out[i] = acc;
    }
#else
    PeerPtrs P;
    memset(&P, 0, sizeof(P));
    P.n = n;
    for (int r = 0; r < n; ++r) {
        if (((uintptr_t)planes[r] & 15) != 0) return pfail(LITHO_ERR_ARG, "peer_sum: planes must be 16-byte aligned");
        P.p[r] = const_cast<float*>(planes[r]);
    }
    if (((uintptr_t)out & 15) != 0) return pfail(LITHO_ERR_ARG, "peer_sum: out must be 16-byte aligned");
    // Few, long-lived CTAs: the sum usually runs next to persistent compute kernels that hold every register of every
    // SM, so each of its CTAs has to wait for one of theirs to retire -- 64 CTAs x 512 threads with n x 16 bytes in
    // flight per thread saturate the NVLink ingress and need 64 slot hand-overs instead of ~600 (LITHO_PEER_SUM_CTAS).
    static const int cap_env = []() { const char* e = getenv("LITHO_PEER_SUM_CTAS"); return e ? atoi(e) : 64; }();
    unsigned long long want = (elems / 4 + 511) / 512;
    const unsigned long long cap = (unsigned long long)(cap_env > 0 ? cap_env : 64);
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    peer_sum_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(out, P, (unsigned long long)elems,
                                                            (const unsigned long long*)flags, (unsigned long long)value, err);
    PCHECK(cudaGetLastError());
#endif
    return LITHO_OK;
}

}  // extern "C"
