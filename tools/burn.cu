// Diagnostic load generator (NOT part of the product): a persistent kernel that keeps `chains` dependent integer multiply-add chains per warp
// busy on one warp per SM sub-partition for `seconds`, i.e. takes a known share of every scheduler's issue slots while bench.py runs
// (OLF_BENCH_BURN=chains,seconds).  If the front end's throughput drops by that share, the front end is issue-bound.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(128) k_burn(unsigned long long dur_ns, int chains, unsigned* out, unsigned long long* iters_out) {
    unsigned a = threadIdx.x * 2654435761u + blockIdx.x, b = a ^ 0x9e3779b9u, c = a + 77u, d = ~a;
    unsigned long long it = 0, t_end_ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end_ns)); t_end_ns += dur_ns;
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 256; ++k) {
            a = a * 1664525u + 1013904223u;
            if (chains > 1) b = b * 22695477u + 1u;
            if (chains > 2) c = c * 1103515245u + 12345u;
            if (chains > 3) d = d * 134775813u + 1u;
        }
        ++it;
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t > t_end_ns) break;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
    if (threadIdx.x == 0 && blockIdx.x == 0) *iters_out = it;
}
// mode 2: every lane of `warps` warps per SM chases random 32-byte sectors of a 2 GB buffer with L1-bypassing loads (one dependent chain per lane): a
// known rate of single-sector requests on every SM's path to the L2 / DRAM, no issue-slot pressure to speak of
__global__ void __launch_bounds__(1024) k_burn_mem(unsigned long long dur_ns, const unsigned* __restrict__ buf, unsigned mask, unsigned* out, unsigned long long* iters_out) {
    unsigned long long it = 0, t_end_ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end_ns)); t_end_ns += dur_ns;
    unsigned idx = (threadIdx.x * 2654435761u + blockIdx.x * 40503u) & mask;
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 64; ++k) { unsigned v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(buf + (size_t)idx * 8)); idx = (idx * 1664525u + 1013904223u + v) & mask; }
        ++it;
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t > t_end_ns) break;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *iters_out = it;
}
// mode 3: a host thread keeps launching grids of `ctas` trivial CTAs (each writes one word and exits): load on the CTA dispatcher / launch path only
__global__ void __launch_bounds__(64) k_burn_cta(unsigned* out) { if (threadIdx.x == 0) out[blockIdx.x & 1023] = blockIdx.x; }
#include <thread>
#include <atomic>
#include <chrono>
static std::thread g_thr; static std::atomic<unsigned long long> g_launches{0};
// mode 4: as mode 2, but with the operations the region-growing passes use on their claim words: strong (relaxed, gpu-scope) 64-bit loads and a
// fire-and-forget 64-bit atomic min every fourth access, at random 32-byte sectors
__global__ void __launch_bounds__(1024) k_burn_strong(unsigned long long dur_ns, unsigned long long* __restrict__ buf, unsigned mask, unsigned* out, unsigned long long* iters_out) {
    unsigned long long it = 0, t_end_ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end_ns)); t_end_ns += dur_ns;
    unsigned idx = (threadIdx.x * 2654435761u + blockIdx.x * 40503u) & mask;
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 64; ++k) {
            unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(buf + (size_t)idx * 4) : "memory");
            if ((k & 3) == 3) atomicMin(buf + (size_t)idx * 4, 0xFFFFFFFFFFFFFFFFull - k);
            idx = (idx * 1664525u + 1013904223u + (unsigned)v) & mask;
        }
        ++it;
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t > t_end_ns) break;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *iters_out = it;
}
static unsigned* g_buf = nullptr;
static cudaStream_t g_s = nullptr; static unsigned* g_out = nullptr; static unsigned long long* g_it = nullptr;
extern "C" int burn_start(int device, int chains, double seconds) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (!g_s) { cudaStreamCreateWithFlags(&g_s, cudaStreamNonBlocking); cudaMalloc(&g_out, (size_t)sms * 128 * 4); cudaMallocHost(&g_it, 8); }
    if (chains >= 1000) {         // CTA-dispatch mode: chains = CTAs per launch
        const int ctas = chains;
        if (g_thr.joinable()) g_thr.join();
        g_launches = 0;
        g_thr = std::thread([device, ctas, seconds] {
            cudaSetDevice(device);
            const auto t_end = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
            unsigned long long n = 0;
            while (std::chrono::steady_clock::now() < t_end) {
                for (int k = 0; k < 8; ++k) k_burn_cta<<<ctas, 64, 0, g_s>>>(g_out);
                cudaStreamSynchronize(g_s); n += 8;
            }
            g_launches = n;
        });
        return 0;
    }
    if (chains >= 200 && chains < 1000) {          // strong loads + atomics: chains - 200 = warps per SM
        const int warps = chains - 200;
        const size_t sectors = (size_t)1 << 26;
        if (!g_buf) { cudaMalloc(&g_buf, sectors * 32); cudaMemset(g_buf, 0xFF, sectors * 32); cudaFree(g_out); cudaMalloc(&g_out, (size_t)sms * 1024 * 4); }
        k_burn_strong<<<sms, 32 * warps, 0, g_s>>>((unsigned long long)(seconds * 1e9), (unsigned long long*)g_buf, (unsigned)(sectors - 1), g_out, g_it);
        return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
    if (chains >= 100) {          // memory mode: chains - 100 = warps per SM
        const int warps = chains - 100;
        const size_t sectors = (size_t)1 << 26;           // 2 GB of 32-byte sectors
        if (!g_buf) { cudaMalloc(&g_buf, sectors * 32); cudaMemset(g_buf, 0, sectors * 32); cudaFree(g_out); cudaMalloc(&g_out, (size_t)sms * 1024 * 4); }
        k_burn_mem<<<sms, 32 * warps, 0, g_s>>>((unsigned long long)(seconds * 1e9), g_buf, (unsigned)(sectors - 1), g_out, g_it);
    } else
    k_burn<<<sms, 128, 0, g_s>>>((unsigned long long)(seconds * 1e9), chains, g_out, g_it);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
extern "C" unsigned long long burn_wait() { if (g_thr.joinable()) { g_thr.join(); return g_launches.load(); } cudaStreamSynchronize(g_s); return *g_it; }
