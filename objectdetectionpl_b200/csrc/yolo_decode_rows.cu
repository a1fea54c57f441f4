// K1r — the fused decode + filter kernel for heads stored channels-last, [B, A, G, G, 5+C].
//
// The reference's YOLOv5 head really writes this layout (model/YOLOV5.py:96) while its NMS reads the bytes as planar
// (YOLOV5.py:178-183); `layout = B200DET_LAYOUT_CHANNELS_LAST` lets a corrected caller hand the tensor over as it is —
// same candidate order (level, anchor, row, column), no 548 MB permute copy, same results as the planar kernel on the
// permuted tensor (tests/test_gpu_layout.py).
//
// A row (one cell) is 5+C contiguous floats, so 32 consecutive rows are ONE contiguous 32*(5+C)*4-byte block (10.9 KB at
// C = 80): exactly what the TMA's 1-D bulk copy moves.  A warp owns 4 such chunks of its tile; one lane issues
// `cp.async.bulk` (SASS UBLKCP) for chunk k+1 into the warp's second buffer, completion on an mbarrier, while all lanes walk
// their own row of chunk k in shared memory (row stride 5+C words: conflict-free for odd 5+C).  Chunks that straddle a
// level boundary, run past N or are not 16-byte aligned take a scalar per-row path.
// Compaction is ordered as in the planar kernel (warp ballots + a prefix over the 16 (warp, chunk) counts of the tile).
#include "yolo_ws.cuh"
#include "yolo_k1.cuh"

namespace b200det {

constexpr int kRowsThreads = 128;                       // 4 warps x 4 chunks x 32 rows = one 512-candidate tile
constexpr int kRowsWarps = kRowsThreads / 32;
constexpr int kRowsChunks = kTile / 32 / kRowsWarps;    // chunks per warp

__device__ __forceinline__ uint32_t rows_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rows_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rows_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rows_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rows_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rows_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(rows_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void rows_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rows_smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(rows_smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kRowsThreads) yolo_decode_filter_rows_kernel(const K1Params p, const int rowbuf_floats) {
    extern __shared__ __align__(128) float s_dyn[];     // [kRowsWarps][2][rowbuf_floats] row buffers, then int s_hist[C]
    __shared__ K1Stage s_stage;
    __shared__ int s_cnt[kRowsWarps][kRowsChunks];
    __shared__ int s_base[kRowsWarps][kRowsChunks + 1];
    __shared__ __align__(8) uint64_t s_bar[kRowsWarps][2];

    const int b = blockIdx.y, tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = 5 + p.C;
    float* rowbuf = s_dyn + (size_t)warp * 2 * rowbuf_floats;
    int* s_hist = reinterpret_cast<int*>(s_dyn + (size_t)kRowsWarps * 2 * rowbuf_floats);
    if (p.cls_hist) for (int c = tid; c < p.C; c += kRowsThreads) s_hist[c] = 0;
    if (lane == 0) {
        rows_mbar_init(&s_bar[warp][0], 1);
        rows_mbar_init(&s_bar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // per chunk: level / anchor / cell of this lane's row, and whether the 32 rows are one aligned block of one level
    int c_lvl[kRowsChunks], c_rel[kRowsChunks];
    const float* c_row[kRowsChunks];
    unsigned uniform_mask = 0u, range_mask = 0u;
#pragma unroll
    for (int k = 0; k < kRowsChunks; ++k) {
        const int n = tile * kTile + (warp * kRowsChunks + k) * 32 + lane;
        int lvl = 0;
#pragma unroll
        for (int l = 1; l < B200DET_MAX_LEVELS; ++l)
            if (l < p.nlevels && n >= p.off[l]) lvl = l;
        const bool in_range = n < p.N;
        const int rel = in_range ? n - p.off[lvl] : 0;
        c_lvl[k] = lvl; c_rel[k] = rel;
        c_row[k] = p.head[lvl] + ((size_t)b * p.A * p.GG[lvl] + rel) * (size_t)F;   // rows of one (image, level) are contiguous
        const int lvl0 = __shfl_sync(0xFFFFFFFFu, lvl, 0);
        const unsigned long long blk = __shfl_sync(0xFFFFFFFFu, (unsigned long long)c_row[k], 0);
        if (__all_sync(0xFFFFFFFFu, in_range && lvl == lvl0) && (blk & 15ull) == 0ull) uniform_mask |= 1u << k;
        if (in_range) range_mask |= 1u << k;
    }
    const uint32_t chunk_bytes = (uint32_t)(32 * F * 4);
    auto issue = [&](int k) {       // lane 0: start the bulk copy of chunk k into buffer k & 1
        if (((uniform_mask >> k) & 1u) && lane == 0) {
            const unsigned long long blk = (unsigned long long)c_row[k];          // lane 0's row = first row of the block
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // earlier generic reads of the buffer are done
            rows_mbar_expect_tx(&s_bar[warp][k & 1], chunk_bytes);
            rows_bulk_g2s(rowbuf + (size_t)(k & 1) * rowbuf_floats, (const void*)blk, chunk_bytes, &s_bar[warp][k & 1]);
        }
    };
    issue(0);

    float box[kRowsChunks][4], conf[kRowsChunks], ccf[kRowsChunks];
    int cls[kRowsChunks];
    unsigned keepmask = 0u;

#pragma unroll
    for (int k = 0; k < kRowsChunks; ++k) {
        __syncwarp();                                   // every lane is done with the buffer chunk k + 1 goes into
        if (k + 1 < kRowsChunks) issue(k + 1);
        const bool in_range = (range_mask >> k) & 1u;
        const int lvl = c_lvl[k];
        const int GG = p.GG[lvl];
        const int a = c_rel[k] / GG;
        const int cell = c_rel[k] - a * GG;
        float t5[5];
        float best = 0.f;
        int besti = 0;
        if ((uniform_mask >> k) & 1u) {
            // phase parity of the buffer's barrier = number of bulk copies issued on it before this one
            const uint32_t parity = (k >= 2 && ((uniform_mask >> (k - 2)) & 1u)) ? 1u : 0u;
            rows_mbar_wait(&s_bar[warp][k & 1], parity);
            const float* r = rowbuf + (size_t)(k & 1) * rowbuf_floats + lane * F;
#pragma unroll
            for (int f = 0; f < 5; ++f) t5[f] = r[f];
            best = r[5];
            for (int c = 1; c < p.C; ++c) argmax_step(r[5 + c], c, best, besti);
        } else if (in_range) {
            const float* row_g = c_row[k];
#pragma unroll
            for (int f = 0; f < 5; ++f) t5[f] = ldg_stream1(row_g + f);
            best = ldg_stream1(row_g + 5);
            for (int c = 1; c < p.C; ++c) argmax_step(ldg_stream1(row_g + 5 + c), c, best, besti);
        }
        bool keep = false;
        if (in_range) {
            k1_finish<MODE>(p, lvl, a, cell, t5, best, box[k], conf[k], ccf[k]);
            cls[k] = besti;
            keep = conf[k] >= p.conf_thres;                                           // model/YOLOV3.py:310
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) keepmask |= 1u << k;
        if (lane == 0) s_cnt[warp][k] = __popc(bal);
        // in-chunk offset of this lane's survivor, kept in the (otherwise unused) high bits of keepmask
        keepmask |= (unsigned)__popc(bal & lanemask_lt()) << (8 + 6 * k);
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < kRowsWarps; ++w) {
            for (int k = 0; k < kRowsChunks; ++k) { s_base[w][k] = run; run += s_cnt[w][k]; }
            s_base[w][kRowsChunks] = run;
        }
    }
    __syncthreads();
    const int total = s_base[kRowsWarps - 1][kRowsChunks];
#pragma unroll
    for (int k = 0; k < kRowsChunks; ++k) {
        if ((keepmask >> k) & 1u) {
            const int ofs = s_base[warp][k] + (int)((keepmask >> (8 + 6 * k)) & 63u);
            const int n = tile * kTile + (warp * kRowsChunks + k) * 32 + lane;
            k1_stage_put(s_stage, ofs, box[k], conf[k], ccf[k], (uint32_t)n, cls[k]);
            if (p.cls_hist) atomicAdd(&s_hist[cls[k]], 1);
        }
    }
    if (tid == 0) {
        p.tile_count[(size_t)b * p.n_tiles + tile] = (uint32_t)total;
        if (total && p.count) atomicAdd(&p.count[b], (uint32_t)total);
    }
    __syncthreads();
    k1_stage_flush<kRowsThreads>(s_stage, p, (size_t)b * p.n_pad, tile, total, tid);
    if (p.cls_hist) for (int c = tid; c < p.C; c += kRowsThreads) {
        const int h = s_hist[c];
        if (h) atomicAdd(&p.cls_hist[(size_t)b * p.C + c], (uint32_t)h);
    }
}

int launch_k1_rows(const b200det_yolo_desc* d, const K1Params& p, cudaStream_t st) {
    const int F = 5 + d->num_classes;
    const int rowbuf_floats = (32 * F + 3) / 4 * 4;
    const size_t smem = (size_t)kRowsWarps * 2 * rowbuf_floats * sizeof(float) + (size_t)d->num_classes * sizeof(int);
    B2_CHECK_LIMIT(smem + sizeof(K1Stage) + 256 <= 200 * 1024, "channels-last layout: %d classes need %zu bytes of shared memory",
                   d->num_classes, smem);
    dim3 grid(p.n_tiles, d->batch);
    auto launch = [&](auto kernel) -> int {
        B2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, kRowsThreads, smem, st>>>(p, rowbuf_floats);
        B2_LAUNCH_CHECK("yolo_decode_filter_rows_kernel");
        return 0;
    };
    switch (d->decode_mode) {
        case B200DET_DECODE_NONE: return launch(yolo_decode_filter_rows_kernel<B200DET_DECODE_NONE>);
        case B200DET_DECODE_YOLO_EXP: return launch(yolo_decode_filter_rows_kernel<B200DET_DECODE_YOLO_EXP>);
        case B200DET_DECODE_YOLOV4_NORM: return launch(yolo_decode_filter_rows_kernel<B200DET_DECODE_YOLOV4_NORM>);
        default: return launch(yolo_decode_filter_rows_kernel<B200DET_DECODE_YOLOV5>);
    }
}

}  // namespace b200det
