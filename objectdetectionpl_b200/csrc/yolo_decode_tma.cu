// K1 (TMA variant) — the fused YOLO decode + filter kernel as a persistent, warp-specialised pipeline:
// one producer lane streams [planes x 512 cells] stages from HBM into shared memory with bulk asynchronous
// copies (cp.async.bulk global -> shared, completion on an mbarrier), four consumer warps read the stages
// conflict-free (float4 per thread), keep the running first-max argmax in registers and run the same
// epilogue as the LDG kernel (decode, filter, ordered tile compaction, class histogram).
//
// Bulk copies are issued by one thread, do not occupy load/store-unit miss slots or registers and keep
// STAGES * 16 KB per CTA in flight regardless of occupancy; the pipeline never drains because the producer runs ahead
// across the tile boundaries of the CTA's persistent tile list.
// Status: bit-identical to the LDG kernel and selectable with B200DET_K1=tma, but NOT the default: on B200 both
// variants are limited by the same thing (the 58 MB of candidate writes interleaved with the 548 MB read stream,
// tools/ubench/readbw.cu), and the LDG kernel measured 110.7 us vs 117-119 us for this one.
#include "yolo_ws.cuh"
#include "yolo_k1.cuh"

namespace b200det {

constexpr int kTmaPlanes = 8;                       // planes per stage
constexpr int kStageFloats = kTmaPlanes * kTile;    // 4096 floats = 16 KB
constexpr int kMaxRuns = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int NC>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

// exclusive scan over the NC consumer threads (named barrier 1; the producer warp does not take part)
template <int NC>
__device__ __forceinline__ int consumer_exclusive_scan(int v, int* ws /*>=8 ints*/, int* total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    consumer_sync<NC>();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NC / 32; ++w) {
        const int s = ws[w];
        if ((unsigned)w < warp) base += s;
        tot += s;
    }
    *total = tot;
    consumer_sync<NC>();
    return base + inc - v;
}

struct Run {
    const float* src;   // plane 0 of the run
    int len;            // cells
    int dst;            // cell offset inside the 512-cell tile
    int gg;             // plane stride (floats)
};

// maximal (level, anchor)-contiguous runs of the orig-index range [tile*512, min(+512, N))
__device__ __forceinline__ int tile_runs(const K1Params& p, int b, int tile, Run* runs) {
    const int n0 = tile * kTile, n1 = min(n0 + kTile, p.N);
    const int F = 5 + p.C;
    int n = n0, r = 0;
    while (n < n1 && r < kMaxRuns) {
        int lvl = 0;
#pragma unroll
        for (int l = 1; l < B200DET_MAX_LEVELS; ++l)
            if (l < p.nlevels && n >= p.off[l]) lvl = l;
        const int GG = p.GG[lvl];
        const int rel = n - p.off[lvl];
        const int a = rel / GG, cell = rel - a * GG;
        const int len = min(GG - cell, n1 - n);
        runs[r].src = p.head[lvl] + ((size_t)(b * p.A + a) * F) * (size_t)GG + cell;
        runs[r].len = len; runs[r].dst = n - n0; runs[r].gg = GG;
        n += len;
        ++r;
    }
    return r;
}

template <int CPT> struct SmemVec;
template <> struct SmemVec<4> { static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) { const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; } };
template <> struct SmemVec<2> { static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) { const float2 q = *reinterpret_cast<const float2*>(p); v[0] = q.x; v[1] = q.y; } };

// CPT cells per consumer thread (kTile / CPT consumer threads), kTmaStages = STAGES stages of 8 planes x 512 cells
template <int MODE, int CPT, int STAGES, int CTAS>
__global__ void __launch_bounds__(kTile / CPT + 32, CTAS) yolo_decode_filter_tma_kernel(const K1Params p, const int total_tiles) {
    constexpr int kTmaStages = STAGES;
    constexpr int kTmaConsumers = kTile / CPT;
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ K1Stage s_cand;
    float* s_stage = reinterpret_cast<float*>(s_raw);                                   // [kTmaStages][kTmaPlanes][kTile]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_raw + sizeof(float) * kTmaStages * kStageFloats);
    uint64_t* s_empty = s_full + kTmaStages;
    int* s_scan = reinterpret_cast<int*>(s_empty + kTmaStages);                          // [8]
    int* s_hist = s_scan + 8;                                                            // [C]

    const int tid = threadIdx.x;
    const int F = 5 + p.C;
    const int nchunks = (F + kTmaPlanes - 1) / kTmaPlanes;

    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kTmaConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kTmaConsumers) {
        // ===================== producer warp (one elected lane) =====================
        if (tid == kTmaConsumers) {
            int stage = 0;
            uint32_t phase = 0;
            Run runs[kMaxRuns];
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int b = t / p.n_tiles, tile = t - b * p.n_tiles;
                const int nr = tile_runs(p, b, tile, runs);
                int cells = 0;
                for (int r = 0; r < nr; ++r) cells += runs[r].len;
                for (int c = 0; c < nchunks; ++c) {
                    const int f0 = c * kTmaPlanes, np = min(kTmaPlanes, F - f0);
                    mbar_wait(&s_empty[stage], phase ^ 1u);
                    mbar_expect_tx(&s_full[stage], (uint32_t)(np * cells * 4));
                    float* dst = s_stage + (size_t)stage * kStageFloats;
                    for (int pl = 0; pl < np; ++pl)
                        for (int r = 0; r < nr; ++r)
                            bulk_g2s(dst + pl * kTile + runs[r].dst, runs[r].src + (size_t)(f0 + pl) * runs[r].gg,
                                     (uint32_t)(runs[r].len * 4), &s_full[stage]);
                    if (++stage == kTmaStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / p.n_tiles, tile = t - b * p.n_tiles;
        const int n0 = tile * kTile + tid * CPT;
        const bool in_range = n0 < p.N;
        for (int c = tid; c < p.C; c += kTmaConsumers) s_hist[c] = 0;

        float tv[5][CPT];
        float best[CPT];
        int besti[CPT];
#pragma unroll
        for (int v = 0; v < CPT; ++v) { best[v] = 0.f; besti[v] = 0; }
        for (int c = 0; c < nchunks; ++c) {
            const int f0 = c * kTmaPlanes, np = min(kTmaPlanes, F - f0);
            mbar_wait(&s_full[stage], phase);
            const float* sp = s_stage + (size_t)stage * kStageFloats + tid * CPT;
            if (c == 0) {
#pragma unroll
                for (int pl = 0; pl < kTmaPlanes; ++pl) {
                    if (pl < np) {
                        float qq[CPT];
                        SmemVec<CPT>::ld(sp + pl * kTile, qq);
                        if (pl < 5) {
#pragma unroll
                            for (int v = 0; v < CPT; ++v) tv[pl][v] = qq[v];
                        } else if (pl == 5) {
#pragma unroll
                            for (int v = 0; v < CPT; ++v) { best[v] = qq[v]; besti[v] = 0; }
                        } else {
#pragma unroll
                            for (int v = 0; v < CPT; ++v) argmax_step(qq[v], pl - 5, best[v], besti[v]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int pl = 0; pl < kTmaPlanes; ++pl) {
                    if (pl < np) {
                        float qq[CPT];
                        SmemVec<CPT>::ld(sp + pl * kTile, qq);
#pragma unroll
                        for (int v = 0; v < CPT; ++v) argmax_step(qq[v], f0 + pl - 5, best[v], besti[v]);
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == kTmaStages) { stage = 0; phase ^= 1u; }
        }

        // ---- epilogue: decode, filter, ordered compaction (same arithmetic as the LDG kernel) ----
        float box[CPT][4], conf[CPT], ccf[CPT];
        int cls[CPT];
        bool keep[CPT];
#pragma unroll
        for (int v = 0; v < CPT; ++v) keep[v] = false;
        if (in_range) {
            int lvl = 0;
#pragma unroll
            for (int l = 1; l < B200DET_MAX_LEVELS; ++l)
                if (l < p.nlevels && n0 >= p.off[l]) lvl = l;
            const int GG = p.GG[lvl];
            const int rel = n0 - p.off[lvl];
            const int a = rel / GG, cell = rel - a * GG;
#pragma unroll
            for (int v = 0; v < CPT; ++v) {
                float t5[5] = {tv[0][v], tv[1][v], tv[2][v], tv[3][v], tv[4][v]};
                k1_finish<MODE>(p, lvl, a, cell + v, t5, best[v], box[v], conf[v], ccf[v]);
                cls[v] = besti[v];
                keep[v] = conf[v] >= p.conf_thres;
            }
        }
        int cnt = 0;
#pragma unroll
        for (int v = 0; v < CPT; ++v) cnt += keep[v] ? 1 : 0;
        int total;
        int ofs = consumer_exclusive_scan<kTmaConsumers>(cnt, s_scan, &total);       // also orders the s_hist zero-fill
        const size_t img = (size_t)b * p.n_pad;
#pragma unroll
        for (int v = 0; v < CPT; ++v) {
            if (keep[v]) {
                k1_stage_put(s_cand, ofs, box[v], conf[v], ccf[v], (uint32_t)(n0 + v), cls[v]);
                atomicAdd(&s_hist[cls[v]], 1);
                ++ofs;
            }
        }
        if (tid == 0) {
            p.tile_count[(size_t)b * p.n_tiles + tile] = (uint32_t)total;
            if (total && p.count) atomicAdd(&p.count[b], (uint32_t)total);
        }
        consumer_sync<kTmaConsumers>();
        k1_stage_flush<kTmaConsumers>(s_cand, p, img, tile, total, tid);
        if (p.cls_hist) for (int c = tid; c < p.C; c += kTmaConsumers) {
            const int h = s_hist[c];
            if (h) atomicAdd(&p.cls_hist[(size_t)b * p.C + c], (uint32_t)h);
        }
        consumer_sync<kTmaConsumers>();                                // s_hist is re-zeroed by the next tile
    }
}

bool k1_tma_supported(const K1Params& p) {
    for (int l = 0; l < p.nlevels; ++l)
        if (p.GG[l] % 4 != 0 || p.GG[l] < 128 || ((uintptr_t)p.head[l] & 15) != 0) return false;   // <= 5 runs per tile
    return p.C <= 8192;
}

template <int MODE, int CPT, int STAGES, int CTAS>
static int launch_cfg(const K1Params& p, int total_tiles, int sm_count, cudaStream_t st) {
    constexpr int ctas_per_sm = CTAS;
    const size_t smem = sizeof(float) * STAGES * kStageFloats + sizeof(uint64_t) * 2 * STAGES + sizeof(int) * 8 +
                        sizeof(int) * (size_t)p.C;
    B2_CUDA(cudaFuncSetAttribute(yolo_decode_filter_tma_kernel<MODE, CPT, STAGES, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    const int grid = total_tiles < ctas_per_sm * sm_count ? total_tiles : ctas_per_sm * sm_count;   // persistent CTAs
    yolo_decode_filter_tma_kernel<MODE, CPT, STAGES, CTAS><<<grid, kTile / CPT + 32, smem, st>>>(p, total_tiles);
    B2_LAUNCH_CHECK("yolo_decode_filter_tma_kernel");
    return 0;
}

template <int MODE>
static int launch_mode(const K1Params& p, int total_tiles, int sm_count, cudaStream_t st) {
    // 2 cells per consumer thread (8 consumer warps), 5 stages x 16 KB, 2 persistent CTAs per SM: the best of the
    // measured configurations (4/2 cells per thread x 3..6 stages x 2..4 CTAs per SM were all within 7%)
    return launch_cfg<MODE, 2, 5, 2>(p, total_tiles, sm_count, st);
}

int launch_k1_tma(const b200det_yolo_desc* d, const K1Params& p, cudaStream_t st) {
    const int total_tiles = p.n_tiles * d->batch;
    static int sm_counts[64] = {};      // per device ordinal
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    int sm_count = (dev >= 0 && dev < 64) ? sm_counts[dev] : 0;
    if (sm_count == 0) {
        B2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_counts[dev] = sm_count;
    }
    switch (d->decode_mode) {
        case B200DET_DECODE_NONE: return launch_mode<B200DET_DECODE_NONE>(p, total_tiles, sm_count, st);
        case B200DET_DECODE_YOLO_EXP: return launch_mode<B200DET_DECODE_YOLO_EXP>(p, total_tiles, sm_count, st);
        default: return launch_mode<B200DET_DECODE_YOLOV5>(p, total_tiles, sm_count, st);
    }
}

}  // namespace b200det
