// K1 — fused YOLO head decode + confidence filter + class argmax + score key + ordered tile compaction.
// K0 — decode_box: full decoded map (planar -> rows transpose with the decode formulas).
//
// Replaces the view/permute/contiguous/cat/xywh2xyxy/filter/max/score prologue of the reference NMS
// (model/YOLOV3.py:289-319, YOLOV5.py:173-202) and the inline decode formulas D1/D2
// (LightningFunc/accuracy.py:412-435,459-466; utils/YoloV5Utils.py:241-248).
//
// HBM-bound streaming kernel: every thread owns 4 consecutive cells of one level and walks the 5+C planes at
// stride G*G, so a warp reads 512 contiguous bytes per plane (128-bit loads in tiles of levels with G*G % 4 == 0,
// scalar loads otherwise — a per-tile choice), with 8 independent loads in flight per thread.
// Survivors are compacted IN ORDER inside the CTA's 512-slot tile (ballot-free block scan), so the
// slot order equals the reference's candidate order and the later stable radix sort breaks score ties
// by ascending candidate index without any atomically-ordered append.
#include <stdlib.h>

#include "yolo_ws.cuh"
#include "yolo_k1.cuh"

namespace b200det {

int launch_k1_rows(const b200det_yolo_desc* d, const K1Params& p, cudaStream_t st);
bool k1_tma_supported(const K1Params& p);
int launch_k1_tma(const b200det_yolo_desc* d, const K1Params& p, cudaStream_t st);

template <int VEC>
struct VecLoad;
template <>
struct VecLoad<4> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        float4 t = ldg_stream4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <>
struct VecLoad<1> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = ldg_stream1(p); }
};

// One CTA = one 512-candidate tile of one level; one thread = 4 consecutive cells.  Tiles of a level whose planes keep
// 4-cell groups aligned (G*G % 4 == 0, 16-byte aligned head) use 128-bit loads; the others (13x13 ...) read the same four
// cells with scalar loads — a per-tile, CTA-uniform choice, so a YOLOv3/v4 call (13, 26, 52) streams 94 % of its bytes
// through the vector path instead of dropping the whole launch to the scalar one (measured 100 us for 232 MB before).
template <int MODE, int U, int MINB>
__global__ void __launch_bounds__(kTile / 4, MINB)
yolo_decode_filter_kernel(const K1Params p) {
    constexpr int VEC = 4;
    constexpr int NT = kTile / VEC;
    extern __shared__ int s_hist[];          // [C]
    __shared__ int s_scan[33];
    __shared__ K1Stage s_stage;

    // Launch order.  Default: x = tile, y = image.  A head with an odd plane (13 x 13 ...) has a few scalar-load tiles that
    // each take several times as long as a vector tile; with the default order image b's scalar tile only starts after the
    // b * n_tiles CTAs before it, and the last ones are the launch's tail.  slow_first: x = image, y = tile rotated so that
    // the first scalar level's tiles are y = 0 — all of them start at once and the vector tiles fill in around them.
    const int b = p.slow_first ? blockIdx.x : blockIdx.y;
    int tile = p.slow_first ? (int)blockIdx.y + p.slow_first - 1 : (int)blockIdx.x;
    if (p.slow_first && tile >= p.n_tiles) tile -= p.n_tiles;
    const int tid = threadIdx.x;
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < B200DET_MAX_LEVELS; ++l)
        if (l < p.nlevels && tile >= p.tile_off[l]) lvl = l;
    const int GG = p.GG[lvl];
    const int lvl_n = p.A * GG;                                  // candidates of this level
    const int rel0 = (tile - p.tile_off[lvl]) * kTile + tid * VEC;
    const int n0 = p.off[lvl] + rel0;                            // candidate index of the thread's first cell
    const int F = 5 + p.C;
    const bool vec_tile = (GG & 3) == 0 && (((uintptr_t)p.head[lvl]) & 15) == 0;

    if (p.cls_hist) for (int c = tid; c < p.C; c += NT) s_hist[c] = 0;

    float box[VEC][4];
    float conf[VEC], ccf[VEC];
    int cls[VEC];
    bool keep[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) keep[v] = false;

    if (vec_tile) {
        if (rel0 < lvl_n) {                  // the 4 cells are all in range and in one plane row (lvl_n % 4 == 0)
            const int a = rel0 / GG;
            const int cell = rel0 - a * GG;
            const float* base = p.head[lvl] + ((size_t)(b * p.A + a) * F) * (size_t)GG + cell;

            float t[5][VEC];
#pragma unroll
            for (int f = 0; f < 5; ++f) VecLoad<VEC>::ld(base + (size_t)f * GG, t[f]);

            float best[VEC];
            int besti[VEC];
            const float* cp = base + (size_t)5 * GG;
            {
                float v0[VEC];
                VecLoad<VEC>::ld(cp, v0);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { best[v] = v0[v]; besti[v] = 0; }
            }
            int c = 1;
            for (; c + U <= p.C; c += U) {
                float buf[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) VecLoad<VEC>::ld(cp + (size_t)(c + u) * GG, buf[u]);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) argmax_step(buf[u][v], c + u, best[v], besti[v]);
            }
            for (; c < p.C; ++c) {
                float buf[VEC];
                VecLoad<VEC>::ld(cp + (size_t)c * GG, buf);
#pragma unroll
                for (int v = 0; v < VEC; ++v) argmax_step(buf[v], c, best[v], besti[v]);
            }

#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float t5[5] = {t[0][v], t[1][v], t[2][v], t[3][v], t[4][v]};
                k1_finish<MODE>(p, lvl, a, cell + v, t5, best[v], box[v], conf[v], ccf[v]);
                cls[v] = besti[v];
                keep[v] = conf[v] >= p.conf_thres;     // model/YOLOV3.py:310 (NaN conf is dropped, as there)
            }
        }
    } else {
        // scalar tile: every cell on its own (a 4-cell group may straddle anchors and the end of the level)
        const float* base[VEC];
        int av[VEC], cellv[VEC];
        bool ok[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int rel = rel0 + v;
            ok[v] = rel < lvl_n;
            av[v] = ok[v] ? rel / GG : 0;
            cellv[v] = ok[v] ? rel - av[v] * GG : 0;
            base[v] = p.head[lvl] + ((size_t)(b * p.A + av[v]) * F) * (size_t)GG + cellv[v];
        }
        float t[5][VEC];
#pragma unroll
        for (int f = 0; f < 5; ++f)
#pragma unroll
            for (int v = 0; v < VEC; ++v) t[f][v] = ok[v] ? ldg_stream1(base[v] + (size_t)f * GG) : 0.0f;
        float best[VEC];
        int besti[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { best[v] = ok[v] ? ldg_stream1(base[v] + (size_t)5 * GG) : 0.0f; besti[v] = 0; }
        // 3 planes x 4 cells = 12 independent scalar loads in flight per thread (2 planes measured 68 us for the YOLOv3-416 head:
        // the 64 scalar tiles of the 13 x 13 level were the launch's long pole, 40 dependent round trips each)
        constexpr int US = 3;
        for (int c = 1; c < p.C; c += US) {
            float buf[US][VEC];
#pragma unroll
            for (int u = 0; u < US; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    buf[u][v] = (ok[v] && c + u < p.C) ? ldg_stream1(base[v] + (size_t)(5 + c + u) * GG) : 0.0f;
#pragma unroll
            for (int u = 0; u < US; ++u)
                if (c + u < p.C)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) argmax_step(buf[u][v], c + u, best[v], besti[v]);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (ok[v]) {
                const float t5[5] = {t[0][v], t[1][v], t[2][v], t[3][v], t[4][v]};
                k1_finish<MODE>(p, lvl, av[v], cellv[v], t5, best[v], box[v], conf[v], ccf[v]);
                cls[v] = besti[v];
                keep[v] = conf[v] >= p.conf_thres;
            }
        }
    }

    int cnt = 0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) cnt += keep[v] ? 1 : 0;
    int total;
    int ofs = block_exclusive_scan(cnt, s_scan, &total);   // also orders the s_hist zero-fill

    const size_t img = (size_t)b * p.n_pad;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (keep[v]) {
            k1_stage_put(s_stage, ofs, box[v], conf[v], ccf[v], (uint32_t)(n0 + v), cls[v]);
            if (p.cls_hist) atomicAdd(&s_hist[cls[v]], 1);
            ++ofs;
        }
    }
    if (tid == 0) {
        p.tile_count[(size_t)b * p.n_tiles + tile] = (uint32_t)total;
        if (total && p.count) atomicAdd(&p.count[b], (uint32_t)total);
    }
    __syncthreads();
    k1_stage_flush<NT>(s_stage, p, img, tile, total, tid);
    if (p.cls_hist) for (int c = tid; c < p.C; c += NT) {
        int h = s_hist[c];
        if (h) atomicAdd(&p.cls_hist[(size_t)b * p.C + c], (uint32_t)h);
    }
}

// Exclusive scan of the (image, class) histogram -> class segment offsets.  One CTA per image.
__global__ void __launch_bounds__(256) seg_scan_kernel(const uint32_t* __restrict__ cls_hist,
                                                       uint32_t* __restrict__ seg_off, int C) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    int carry = 0;
    for (int c0 = 0; c0 < C; c0 += 256) {
        const int c = c0 + threadIdx.x;
        int v = c < C ? (int)cls_hist[(size_t)b * C + c] : 0;
        int total;
        int ex = block_exclusive_scan(v, s_scan, &total);
        if (c < C) seg_off[(size_t)b * (C + 1) + c] = (uint32_t)(carry + ex);
        carry += total;
    }
    if (threadIdx.x == 0) seg_off[(size_t)b * (C + 1) + C] = (uint32_t)carry;
}

int yolo_validate(const b200det_yolo_desc* d, const void* ws, size_t ws_bytes) {
    B2_CHECK_ARG(d != nullptr, "desc is null");
    B2_CHECK_ARG(d->batch > 0 && d->num_anchors > 0 && d->num_classes > 0, "batch/anchors/classes must be > 0");
    B2_CHECK_ARG(d->layout == B200DET_LAYOUT_PLANAR || d->layout == B200DET_LAYOUT_CHANNELS_LAST, "bad layout %d", d->layout);
    B2_CHECK_LIMIT(d->num_levels >= 1 && d->num_levels <= B200DET_MAX_LEVELS, "num_levels %d out of [1,%d]",
                   d->num_levels, B200DET_MAX_LEVELS);
    B2_CHECK_LIMIT(d->num_anchors <= B200DET_MAX_ANCHORS, "num_anchors %d > %d", d->num_anchors, B200DET_MAX_ANCHORS);
    B2_CHECK_LIMIT(d->num_classes <= B200DET_MAX_CLASSES, "num_classes %d > %d", d->num_classes, B200DET_MAX_CLASSES);
    B2_CHECK_LIMIT(d->batch <= 65535, "batch %d > 65535", d->batch);
    for (int l = 0; l < d->num_levels; ++l) {
        B2_CHECK_ARG(d->head[l] != nullptr, "head[%d] is null", l);
        B2_CHECK_ARG(d->grid[l] > 0, "grid[%d] must be > 0", l);
    }
    B2_CHECK_ARG(d->decode_mode >= B200DET_DECODE_NONE && d->decode_mode <= B200DET_DECODE_YOLOV4_NORM, "bad decode_mode %d",
                 d->decode_mode);
    int N, n_pad;
    B2_CHECK_LIMIT(yolo_counts(d, &N, &n_pad) == 0, "candidates per image out of (0, %d]", B200DET_MAX_CANDIDATES);
    B2_CHECK_ARG(ws != nullptr, "workspace is null");
    B2_CHECK_ARG(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
    YoloWs w;
    yolo_ws_layout(d, (void*)ws, &w);
    if (ws_bytes < w.total_bytes) {
        set_error("workspace too small: %zu < %zu", ws_bytes, w.total_bytes);
        return B200DET_EWORKSPACE;
    }
    return 0;
}

static int launch_k1(const b200det_yolo_desc* d, const K1Params& p_in, cudaStream_t st) {
    K1Params p = p_in;
    dim3 grid(p.n_tiles, d->batch);
    p.slow_first = 0;
    for (int l = 0; l < p.nlevels && !p.slow_first; ++l)
        if ((p.GG[l] & 3) != 0 || (((uintptr_t)p.head[l]) & 15) != 0) p.slow_first = p.tile_off[l] + 1;   // first scalar-tile level
    if (p.slow_first && p.n_tiles <= 65535) grid = dim3(d->batch, p.n_tiles);
    else p.slow_first = 0;
    const size_t smem = (size_t)d->num_classes * sizeof(int);
    constexpr int VEC = 4;
    // U = 8 loads in flight per thread, <= 80 registers (6 CTAs/SM): best of the measured (U, occupancy, cache-hint,
    // 128/256-bit) variants on B200, all of which sit within 4% of each other (see DESIGN.md, K1 tuning).
    switch (d->decode_mode) {
        case B200DET_DECODE_NONE:
            yolo_decode_filter_kernel<B200DET_DECODE_NONE, 8, 6><<<grid, kTile / VEC, smem, st>>>(p);
            break;
        case B200DET_DECODE_YOLO_EXP:
            yolo_decode_filter_kernel<B200DET_DECODE_YOLO_EXP, 8, 5><<<grid, kTile / VEC, smem, st>>>(p);
            break;
        case B200DET_DECODE_YOLOV4_NORM:
            yolo_decode_filter_kernel<B200DET_DECODE_YOLOV4_NORM, 8, 5><<<grid, kTile / VEC, smem, st>>>(p);
            break;
        default:
            yolo_decode_filter_kernel<B200DET_DECODE_YOLOV5, 8, 5><<<grid, kTile / VEC, smem, st>>>(p);
            break;
    }
    B2_LAUNCH_CHECK("yolo_decode_filter_kernel");
    return 0;
}

// Zero the counter header of the workspace.  A plain kernel: cudaMemsetAsync of these ~0.4 MB measured 13-15 us
// between CUDA events on B200, this takes ~3.
__global__ void __launch_bounds__(256) zero_fill_kernel(uint4* __restrict__ p, size_t n16) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

int zero_fill_launch(void* p, size_t bytes, cudaStream_t st) {      // bytes is a multiple of 256 (workspace layout)
    const size_t n16 = bytes / 16;
    zero_fill_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((uint4*)p, n16);
    B2_LAUNCH_CHECK("zero_fill_kernel");
    return 0;
}

int yolo_stage_reset(const b200det_yolo_desc* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = yolo_validate(d, ws, ws_bytes);
    if (rc) return rc;
    YoloWs w;
    yolo_ws_layout(d, ws, &w);
    if (yolo_fast_path(w)) return 0;          // the cluster sort produces count / seg_off / zeroed chunk counters itself
    return zero_fill_launch(w.count, w.zero_bytes, st);
}

int yolo_stage_decode(const b200det_yolo_desc* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = yolo_validate(d, ws, ws_bytes);
    if (rc) return rc;
    YoloWs w;
    yolo_ws_layout(d, ws, &w);

    K1Params p;
    memset(&p, 0, sizeof(p));
    bool vec4 = true;
    int off = 0, toff = 0;
    for (int l = 0; l < d->num_levels; ++l) {
        p.head[l] = d->head[l];
        p.G[l] = d->grid[l];
        p.GG[l] = d->grid[l] * d->grid[l];
        p.off[l] = off;
        p.tile_off[l] = toff;
        off += d->num_anchors * p.GG[l];
        toff += ceil_div(d->num_anchors * p.GG[l], kTile);
        p.stride[l] = d->stride[l];
        for (int a = 0; a < d->num_anchors; ++a) {
            p.anc[l][a][0] = d->anchors[l][a][0];
            p.anc[l][a][1] = d->anchors[l][a][1];
        }
        if (p.GG[l] % 4 != 0 || ((uintptr_t)d->head[l] & 15) != 0) vec4 = false;
    }
    p.off[d->num_levels] = off;
    p.tile_off[d->num_levels] = toff;
    p.nlevels = d->num_levels; p.A = d->num_anchors; p.C = d->num_classes;
    p.N = w.N; p.n_pad = w.n_pad; p.n_tiles = w.n_tiles;
    p.conf_thres = d->conf_thres;
    p.sxy = d->scale_x_y != 0.0f ? d->scale_x_y : 1.0f;
    p.soff = (float)(0.5 * ((double)p.sxy - 1.0));
    p.box4 = w.box4; p.cc2 = w.cc2; p.orig = w.orig; p.key = w.key[0]; p.pay = w.pay[0];
    p.tile_count = w.tile_count; p.count = w.count; p.cls_hist = w.cls_hist;
    if (yolo_fast_path(w)) { p.count = nullptr; p.cls_hist = nullptr; }      // see yolo_stage_reset

    if (d->layout == B200DET_LAYOUT_CHANNELS_LAST) return launch_k1_rows(d, p, st);
    // default: the register-staged LDG kernel; B200DET_K1=tma selects the bulk-async (TMA) pipeline, which is
    // bit-identical and measured 6% slower on B200 (see yolo_decode_tma.cu)
    const char* k1 = getenv("B200DET_K1");
    if (vec4 && k1 && strcmp(k1, "tma") == 0 && k1_tma_supported(p) && d->decode_mode != B200DET_DECODE_YOLOV4_NORM)
        return launch_k1_tma(d, p, st);
    return launch_k1(d, p, st);
}

// class segment offsets; launched at the head of the sort stage (only the NMS needs them)
int seg_scan_launch(const uint32_t* cls_hist, uint32_t* seg_off, int C, int batch, cudaStream_t st) {
    seg_scan_kernel<<<batch, 256, 0, st>>>(cls_hist, seg_off, C);
    B2_LAUNCH_CHECK("seg_scan_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K0 — decode_box: [B, A*(5+C), G, G] planar -> [B, A*G*G, 5+C] rows, decoded.
// 32 cells x 64 planes per CTA through a padded shared tile: coalesced reads along cells,
// coalesced writes along fields.
// ---------------------------------------------------------------------------------------------
constexpr int kDecCells = 32;
constexpr int kDecPlanes = 64;

template <int MODE>
__global__ void __launch_bounds__(256)
decode_box_kernel(const float* __restrict__ head, float* __restrict__ out, int A, int C, int G, float stride,
                  const float* __restrict__ anchors /*[A,2] device, may be null for MODE NONE*/) {
    __shared__ float tile[kDecPlanes][kDecCells + 1];
    const int GG = G * G, F = 5 + C;
    const int ba = blockIdx.y;                       // b*A + a
    const int a = ba % A;
    const int cell0 = blockIdx.x * kDecCells;
    const int f0 = blockIdx.z * kDecPlanes;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 32 x 8
    const int cell = cell0 + lx;
    const float* src = head + (size_t)ba * F * GG;

    for (int fp = ly; fp < kDecPlanes; fp += 8) {
        const int f = f0 + fp;
        float v = 0.f;
        if (f < F && cell < GG) {
            v = ldg_stream1(src + (size_t)f * GG + cell);
            if (MODE != B200DET_DECODE_NONE) {
                const int gy = cell / G;
                const float gx = (float)(cell - gy * G);
                if (MODE == B200DET_DECODE_YOLO_EXP) {
                    if (f == 0) v = __fmul_rn(__fadd_rn(sigmoidf_acc(v), gx), stride);
                    else if (f == 1) v = __fmul_rn(__fadd_rn(sigmoidf_acc(v), (float)gy), stride);
                    else if (f < 4) v = __fmul_rn(__fmul_rn(expf(v), anchors[a * 2 + (f - 2)]), stride);
                    else v = sigmoidf_acc(v);
                } else {
                    float s = sigmoidf_acc(v);
                    if (f == 0) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.0f), 0.5f), gx), stride);
                    else if (f == 1) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.0f), 0.5f), (float)gy), stride);
                    else if (f < 4) { float s2 = __fmul_rn(s, 2.0f); v = __fmul_rn(__fmul_rn(s2, s2), anchors[a * 2 + (f - 2)]); }
                    else v = s;
                }
            }
        }
        tile[fp][lx] = v;
    }
    __syncthreads();
    const int nf = min(kDecPlanes, F - f0);          // fields of this chunk
    const int ncell = min(kDecCells, GG - cell0);
    float* dst = out + ((size_t)ba * GG + cell0) * F + f0;
    for (int i = threadIdx.x; i < ncell * nf; i += 256) {
        const int cl = i / nf, f = i - cl * nf;
        dst[(size_t)cl * F + f] = tile[f][cl];
    }
}

// K0t — the same transpose, one CTA per 64 consecutive cells of one (image, anchor) slab with ALL their 5+C fields.  The
// planes are read along the cells — 128-bit loads (16 lanes = 256 contiguous bytes per plane) when the plane size is a
// multiple of 4 and the head is 16-byte aligned (VEC: every YOLO grid but the odd ones, 13 x 13, 19 x 19), 32-bit loads
// otherwise — decoded, and stored into shared memory in the OUTPUT order [cell][field], which for 64 consecutive cells is one
// contiguous 64*(5+C)*4-byte block of the result: the write-out is a linear copy, 128 bits at a time (the tile sits in
// shared memory at the same offset mod 16 bytes as its destination, so both sides of the copy are aligned whatever the
// plane size).  (The 32 x 64 tile kernel above: scalar loads, an integer division per element on the way out, 1.5 TB/s;
// this one streams the 836 MB of a 640-pixel head level at 5.3 TB/s.  The old kernel remains for rows wider than 128.)
constexpr int kDecTileCells = 64;
constexpr int kDecTileMaxF = 128;

// one field of one cell; AUX: the box before the stride multiply (grid units) and the objectness go to aux5[0..4]
template <int MODE, bool AUX>
__device__ __forceinline__ float decode_field(float x, int f, int cell, int G, float aw, float ah, float stride, float* aux5) {
    if (MODE == B200DET_DECODE_NONE) return x;
    if (f < 2) {
        const int gy = cell / G;
        const float gxy = f == 0 ? (float)(cell - gy * G) : (float)gy;
        const float sg = sigmoidf_acc(x);
        if (MODE == B200DET_DECODE_YOLO_EXP) {
            const float gu = __fadd_rn(sg, gxy);
            if (AUX) aux5[f] = gu;
            return __fmul_rn(gu, stride);
        }
        return __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sg, 2.0f), 0.5f), gxy), stride);
    }
    if (f < 4) {
        const float an = f == 2 ? aw : ah;
        if (MODE == B200DET_DECODE_YOLO_EXP) {
            const float gu = __fmul_rn(expf(x), an);
            if (AUX) aux5[f] = gu;
            return __fmul_rn(gu, stride);
        }
        const float s2 = __fmul_rn(sigmoidf_acc(x), 2.0f);
        return __fmul_rn(__fmul_rn(s2, s2), an);
    }
    const float sg = sigmoidf_acc(x);
    if (AUX && f == 4) aux5[4] = sg;
    return sg;
}

// AUX (the metrics path, M3): also emits a compact [cell][5] side table with the box BEFORE the stride multiply (grid
// units, what build_targets matches on) and the objectness, so that nobody has to re-read or rescale the wide rows.
template <int MODE, bool AUX, bool VEC>
__global__ void __launch_bounds__(256, 6)
decode_box_tile_kernel(const float* __restrict__ head, float* __restrict__ out, int A, int C, int G, float stride,
                       const float* __restrict__ anchors, float* __restrict__ aux) {
    extern __shared__ __align__(16) float s_raw[];            // 4 floats of slack, [ncell][F], then [ncell][5] with AUX
    const int GG = G * G, F = 5 + C;
    const int ba = blockIdx.y;
    const int a = ba % A;
    const int cell0 = blockIdx.x * kDecTileCells;
    const int ncell = min(kDecTileCells, GG - cell0);
    const float* src = head + (size_t)ba * F * GG + cell0;
    float* dst = out + ((size_t)ba * GG + cell0) * F;
    const int mis = (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);       // 0 whenever VEC and `out` is 16-byte aligned
    float* s_tile = s_raw + mis;
    float* s_aux = s_raw + 4 + kDecTileCells * F;
    float aw = 0.f, ah = 0.f;
    if (MODE != B200DET_DECODE_NONE) { aw = anchors[a * 2]; ah = anchors[a * 2 + 1]; }
    if (VEC) {
        // thread -> (cell group g, planes f0, f0 + 16, ...): three independent 128-bit loads in flight before the first is used
        const int g = threadIdx.x & 15;
        if (g < (ncell >> 2)) {
            for (int f0 = threadIdx.x >> 4; f0 < F; f0 += 48) {
                float4 ld[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + 16 * u;
                    if (f < F) ld[u] = ldg_stream4(src + (size_t)f * GG + (g << 2));
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + 16 * u;
                    if (f >= F) continue;
                    const float v[4] = {ld[u].x, ld[u].y, ld[u].z, ld[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = (g << 2) + k;
                        s_tile[c * F + f] = decode_field<MODE, AUX>(v[k], f, cell0 + c, G, aw, ah, stride, s_aux + c * 5);
                    }
                }
            }
        }
    } else {
        // thread -> (cell c, planes f0, f0 + 4, ...): a warp reads 32 consecutive cells of one plane
        const int c = threadIdx.x & 63;
        if (c < ncell) {
            for (int f0 = threadIdx.x >> 6; f0 < F; f0 += 12) {
                float ld[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + 4 * u;
                    if (f < F) ld[u] = ldg_stream1(src + (size_t)f * GG + c);
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + 4 * u;
                    if (f >= F) continue;
                    s_tile[c * F + f] = decode_field<MODE, AUX>(ld[u], f, cell0 + c, G, aw, ah, stride, s_aux + c * 5);
                }
            }
        }
    }
    __syncthreads();
    // linear copy-out: scalars up to the first 16-byte boundary of the destination, 128-bit body, scalar tail
    const int total = ncell * F;
    const int first = min((4 - mis) & 3, total);
    const int n4 = (total - first) >> 2;
    if ((int)threadIdx.x < first) dst[threadIdx.x] = s_tile[threadIdx.x];
    {
        float4* d4 = reinterpret_cast<float4*>(dst + first);
        const float4* s4 = reinterpret_cast<const float4*>(s_tile + first);
        for (int i = threadIdx.x; i < n4; i += 256) d4[i] = s4[i];
        const int done = first + (n4 << 2);
        if ((int)threadIdx.x < total - done) dst[done + threadIdx.x] = s_tile[done + threadIdx.x];
    }
    if (AUX) {                                                 // ncell * 5 floats, contiguous
        float* adst = aux + ((size_t)ba * GG + cell0) * 5;
        if (VEC) {                                             // ncell % 4 == 0 and cell0 % 64 == 0: 16-byte aligned
            const float4* a4 = reinterpret_cast<const float4*>(s_aux);
            for (int i = threadIdx.x; i < (ncell * 5) >> 2; i += 256) reinterpret_cast<float4*>(adst)[i] = a4[i];
        } else {
            for (int i = threadIdx.x; i < ncell * 5; i += 256) adst[i] = s_aux[i];
        }
    }
}

bool decode_box_tileable(int G, int F, const void* head, const void* out) {
    (void)G; (void)head; (void)out;
    return F <= kDecTileMaxF;
}

template <int MODE, bool AUX>
static void decode_box_tile_launch(const float* head, int B, int A, int C, int G, const float* anchors_dev, float stride,
                                   float* out, float* aux, cudaStream_t st) {
    const int GG = G * G, F = 5 + C;
    dim3 grid(ceil_div(GG, kDecTileCells), B * A);
    const size_t smem = (size_t)(kDecTileCells * (F + (AUX ? 5 : 0)) + 4) * sizeof(float);
    const bool vec = (GG & 3) == 0 && (((uintptr_t)head | (uintptr_t)out | (uintptr_t)aux) & 15) == 0;
    if (vec) decode_box_tile_kernel<MODE, AUX, true><<<grid, 256, smem, st>>>(head, out, A, C, G, stride, anchors_dev, aux);
    else decode_box_tile_kernel<MODE, AUX, false><<<grid, 256, smem, st>>>(head, out, A, C, G, stride, anchors_dev, aux);
}

// M3's decode: rows in pixels plus the [cells][5] grid-unit side table; rows up to 128 wide (the caller checks)
int decode_box_aux_launch(const float* head, int B, int A, int C, int G, const float* anchors_dev, float stride, float* out,
                          float* aux, cudaStream_t st) {
    decode_box_tile_launch<B200DET_DECODE_YOLO_EXP, true>(head, B, A, C, G, anchors_dev, stride, out, aux, st);
    B2_LAUNCH_CHECK("decode_box_tile_kernel<aux>");
    return 0;
}

int decode_box_launch(const float* head, int B, int A, int C, int G, int mode, const float* anchors_dev, float stride,
                      float* out, cudaStream_t st) {
    const int GG = G * G, F = 5 + C;
    if (decode_box_tileable(G, F, head, out)) {
        if (mode == B200DET_DECODE_NONE) decode_box_tile_launch<B200DET_DECODE_NONE, false>(head, B, A, C, G, anchors_dev, stride, out, nullptr, st);
        else if (mode == B200DET_DECODE_YOLO_EXP) decode_box_tile_launch<B200DET_DECODE_YOLO_EXP, false>(head, B, A, C, G, anchors_dev, stride, out, nullptr, st);
        else decode_box_tile_launch<B200DET_DECODE_YOLOV5, false>(head, B, A, C, G, anchors_dev, stride, out, nullptr, st);
        B2_LAUNCH_CHECK("decode_box_tile_kernel");
        return 0;
    }
    dim3 grid(ceil_div(GG, kDecCells), B * A, ceil_div(F, kDecPlanes));
    if (mode == B200DET_DECODE_NONE) decode_box_kernel<B200DET_DECODE_NONE><<<grid, 256, 0, st>>>(head, out, A, C, G, stride, anchors_dev);
    else if (mode == B200DET_DECODE_YOLO_EXP) decode_box_kernel<B200DET_DECODE_YOLO_EXP><<<grid, 256, 0, st>>>(head, out, A, C, G, stride, anchors_dev);
    else decode_box_kernel<B200DET_DECODE_YOLOV5><<<grid, 256, 0, st>>>(head, out, A, C, G, stride, anchors_dev);
    B2_LAUNCH_CHECK("decode_box_kernel");
    return 0;
}

}  // namespace b200det
