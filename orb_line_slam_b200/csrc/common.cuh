// Shared helpers for the sm_100a front-end kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/olf_abi.h"

namespace olf {

void set_last_error(const std::string& s);
void count_launches(int n);
void count_allocs(int n);          // cudaMalloc / cudaHostAlloc / cudaFree calls (each one synchronises the device: none may happen in steady state)
// Wait for a stream without burning a host core: blocking-sync event (thread sleeps) unless OLF_SYNC=spin.
// Many pipelines x 4 extraction threads wait concurrently; spinning would oversubscribe the host CPUs.
cudaError_t stream_sync(cudaStream_t s);
// the two halves of stream_sync: mark "everything enqueued so far" / wait for that mark
cudaError_t stream_record(cudaStream_t s, cudaEvent_t* out);
cudaError_t event_wait(cudaEvent_t ev, bool blocking = false);
// A "wait for everything enqueued so far" event owned by a handle (not by the calling thread: the reference's Frame::Frame
// spawns fresh std::threads every frame, src/Frame.cc:164-171, so nothing CUDA may live in thread-local storage of callers).
// blocking: the waiting thread sleeps until the driver's interrupt (batch rigs: many waiting threads, few host cores);
// otherwise it polls, yielding the core between polls (lowest latency for a single frame).
struct SyncEvent {
    cudaEvent_t ev = nullptr; bool blocking = false;
    cudaError_t create(bool blocking_) {
        blocking = blocking_;
        return cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | (blocking ? cudaEventBlockingSync : 0));
    }
    void destroy() { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    cudaError_t record(cudaStream_t s) { return cudaEventRecord(ev, s); }
    cudaError_t wait() { return event_wait(ev, blocking); }
    cudaError_t sync(cudaStream_t s) { const cudaError_t e = record(s); return e != cudaSuccess ? e : wait(); }
};
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define OLF_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ::olf::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
static inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// RAII-free tiny device buffer (handles own them and free in their destructors)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return OLF_OK;
        // cudaFree / cudaMalloc synchronise the whole device (they would stall every rig in flight): when an existing
        // buffer has to grow, grow it geometrically so that steady state never reallocates
        if (p) count = count + count / 2 + 4096;
        T* np = nullptr;
        count_allocs(1);
        cudaError_t e = cudaMalloc((void**)&np, count * sizeof(T));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);      // the old buffer stays valid
        if (p) cudaFree(p);
        p = np; n = count;
        return OLF_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
// pinned host buffer, optionally mapped into the device address space (zero-copy result hand-off)
template <typename T>
struct PinBuf {
    T* p = nullptr;      // host pointer
    T* d = nullptr;      // device alias (mapped)
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return OLF_OK;
        if (p) count = count + count / 2 + 4096;
        T* np = nullptr; T* nd = nullptr;
        count_allocs(1);
        cudaError_t e = cudaHostAlloc((void**)&np, count * sizeof(T), cudaHostAllocMapped);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);            // the old buffer stays valid
        e = cudaHostGetDevicePointer((void**)&nd, np, 0);
        if (e != cudaSuccess) { cudaFreeHost(np); return cuda_fail(e, "cudaHostGetDevicePointer", __FILE__, __LINE__); }
        if (p) cudaFreeHost(p);
        p = np; d = nd; n = count;
        return OLF_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; d = nullptr; n = 0; }
};

// ---- exact float helpers: every op a single IEEE operation (no FMA contraction), SURVEY Appendix C.4 ----
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// cv::fastAtan2 (degrees), SURVEY A.4.  Bit-identical to the oracle's fast_atan2_deg.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // 0.9997878412794807f * (float)(180/pi) etc.: float products evaluated once on the host, written as hex floats
    const float p1 = 0x1.ca44dep+5f;
    const float p3 = -0x1.2aaddcp+4f;
    const float p5 = 0x1.1d3f7ep+3f;
    const float p7 = -0x1.4515b2p+1f;
    const float eps = 0x1p-52f;       // (float)DBL_EPSILON
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = fdiv(ay, fadd(ax, eps));
        c2 = fmul(c, c);
        a = fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = fdiv(ax, fadd(ay, eps));
        c2 = fmul(c, c);
        a = fsub(90.f, fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = fsub(180.f, a);
    if (y < 0) a = fsub(360.f, a);
    return a;
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint32_t* b) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d += __popc(a[i] ^ b[i]);
    return d;
}

}  // namespace olf
