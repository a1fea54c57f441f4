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

// SSD / RetinaNet prior decode of one prior (model/SSD.py:253-258): loc = (dx, dy, dw, dh), prior = (cx, cy, w, h) -> corners.
// One definition for the decode+filter kernel (prior.cu) and the NMS epilogue (nms.cu), which re-derives the box and label of
// the few output rows that the reference gathers with score-filtered indices from its unfiltered arrays (SSD.py:303-307).
__device__ __forceinline__ float4 prior_decode_box(const float4 l, const float4 pr) {
    const float cx = __fadd_rn(__fmul_rn(l.x, pr.z), pr.x);          // SSD.py:256
    const float cy = __fadd_rn(__fmul_rn(l.y, pr.w), pr.y);
    const float w = __fmul_rn(expf(l.z), pr.z);                      // SSD.py:257
    const float h = __fmul_rn(expf(l.w), pr.w);
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}
// torch.max(dim) over one row of class logits: first maximal index, the first NaN wins over everything (aten TensorCompareKernel).
// The row is cold (it was last read by the decode kernel, a gigabyte ago): eight 128-bit loads are in flight at a time.
__device__ __forceinline__ int prior_row_argmax(const float* __restrict__ row, const int C) {
    float best = row[0];
    int besti = 0;
    int c = 1;
    if ((C & 3) == 0 && (((uintptr_t)row) & 15) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const int n4 = C >> 2;
        for (int c4 = 0; c4 < n4; c4 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = c4 + u < n4 ? __ldg(r4 + c4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c4 + u >= n4) break;
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int cc = ((c4 + u) << 2) + k;
                    if (cc > 0 && !(e[k] <= best) && (best == best)) { best = e[k]; besti = cc; }
                }
            }
        }
        return besti;
    }
    for (; c < C; ++c) {
        const float v = row[c];
        if (!(v <= best) && (best == best)) { best = v; besti = c; }
    }
    return besti;
}

}  // namespace b200det
