// Shared helpers for the b200det kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/b200det.h"

namespace b200det {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B2_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::b200det::set_error(__VA_ARGS__);  \
            return B200DET_EINVAL;              \
        }                                       \
    } while (0)

#define B2_CHECK_LIMIT(cond, ...)               \
    do {                                        \
        if (!(cond)) {                          \
            ::b200det::set_error(__VA_ARGS__);  \
            return B200DET_ELIMIT;              \
        }                                       \
    } while (0)

#define B2_CUDA(expr)                                               \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return ::b200det::cuda_fail(_e, #expr); \
    } while (0)

#define B2_LAUNCH_CHECK(name)                                       \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return ::b200det::cuda_fail(_e, name); \
    } while (0)

// ---- constants -------------------------------------------------------------------------------
constexpr int kTile = B200DET_TILE;        // candidate tile (slots per K1 CTA)
constexpr int kTileShift = 9;
static_assert((1 << kTileShift) == kTile, "tile shift");
constexpr uint32_t kSlotBits = 20;
constexpr uint32_t kSlotMask = (1u << kSlotBits) - 1u;
constexpr uint32_t kNone = 0xFFFFFFFFu;

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per sort CTA
static_assert(kSortTile % kTile == 0, "sort tile must cover whole candidate tiles");

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Streaming (read-once) global loads: bypass L1 allocation so the 548 MB head stream does not evict
// the small reused tables.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Monotone key for a DESCENDING score order under an ASCENDING radix sort.
//  * -0.0 and +0.0 compare equal in torch's argsort -> canonicalised to +0.0
//  * NaN sorts last (torch puts NaN at the end of an ascending sort of -score)
__device__ __forceinline__ uint32_t score_sort_key(float s) {
    if (s != s) return 0xFFFFFFFFu;
    if (s == 0.0f) s = 0.0f;
    uint32_t u = __float_as_uint(s);
    uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// `warp_sums` is a shared array of >= 32 ints.  Returns the exclusive prefix; *total gets the sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    const unsigned nwarps = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < nwarps) ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= (unsigned)o) winc += t;
        }
        warp_sums[lane] = winc - w;  // exclusive warp offsets; lane 31 holds total - last
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    int res = warp_sums[warp] + inc - v;
    *total = warp_sums[32];
    __syncthreads();  // warp_sums may be reused by the caller right away
    return res;
}

}  // namespace b200det
