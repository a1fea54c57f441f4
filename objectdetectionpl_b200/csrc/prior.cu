// K1' — SSD / RetinaNet prior decode + sigmoid-argmax + score filter, and the prior NMS pipeline.
//
// Replaces model/SSD.py:249-310 == model/RetinaNet.py:117-178:
//   xy = loc_xy * p_wh + p_xy ; wh = exp(loc_wh) * p_wh ; box = [xy - wh/2, xy + wh/2]     (SSD.py:253-258)
//   score, label = sigmoid(cls).max(1) ; candidates = score > class_thresh                   (SSD.py:260-262)
//   sort by score, keep top-k, class-agnostic greedy NMS                                     (SSD.py:270-302)
//
// Class logits are row-major [P, C]: a CTA streams a contiguous block of 128 rows through shared
// memory with fully coalesced 128-bit loads, then each thread scans its own row with a per-row
// rotation so that the 32 lanes of a warp hit 32 different banks (row stride C would otherwise put
// C%32==0 heads on 1-2 banks).  sigmoid is monotone, so the argmax is taken on the logits and only
// the winner is squashed (first maximal index on ties, like torch.max).
#include <stdlib.h>

#include "yolo_ws.cuh"

namespace b200det {

constexpr int kPriRows = 128;      // rows (priors) per staging round == threads per CTA
constexpr int kPriMaxCC = 128;     // classes staged per round

struct PriorWs {
    uint32_t* count;        // [B]   (zeroed)
    uint32_t* digit_hist;   // [B][kMaxPasses][256] (zeroed)
    uint32_t* ticket;       // [kMaxPasses][B] (zeroed)
    size_t zero_bytes;
    uint32_t* status;       // [kMaxPasses][B][sort_tiles][256]
    uint32_t* tile_count;   // [B][n_tiles]
    uint32_t* tile_prefix;  // [B][n_tiles]
    float4* box4;           // [B][n_pad]
    float2* cc2;            // [B][n_pad]  (0, score)
    uint32_t* orig;         // [B][n_pad]
    uint32_t* key[2];
    uint32_t* pay[2];
    float4* kbox;           // [B][n_pad]
    uint32_t* kpos;         // [B][n_pad]
    size_t total_bytes;
    int n_pad, n_tiles;
};

static void prior_ws_layout(const b200det_prior_desc* d, void* base, PriorWs* w) {
    const size_t B = (size_t)d->batch;
    w->n_pad = (int)align_up((size_t)d->num_priors, kTile);
    w->n_tiles = w->n_pad / kTile;
    const size_t P = (size_t)w->n_pad;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes, 256); return r; };
    w->count = (uint32_t*)take(B * 4);
    w->digit_hist = (uint32_t*)take(B * kMaxPasses * 256 * 4);
    w->ticket = (uint32_t*)take((size_t)kMaxPasses * B * 4);
    w->zero_bytes = off;
    w->status = (uint32_t*)take((size_t)kMaxPasses * B * ceil_div(w->n_pad, kSortTile) * 256 * 4);
    w->tile_count = (uint32_t*)take(B * (size_t)w->n_tiles * 4);
    w->tile_prefix = (uint32_t*)take(B * (size_t)w->n_tiles * 4);
    w->box4 = (float4*)take(B * P * 16);
    w->cc2 = (float2*)take(B * P * 8);
    w->orig = (uint32_t*)take(B * P * 4);
    for (int i = 0; i < 2; ++i) w->key[i] = (uint32_t*)take(B * P * 4);
    for (int i = 0; i < 2; ++i) w->pay[i] = (uint32_t*)take(B * P * 4);
    w->kbox = (float4*)take(B * P * 16);
    w->kpos = (uint32_t*)take(B * P * 4);
    w->total_bytes = off;
}

struct K1pParams {
    const float* loc;
    const float* cls;
    const float* priors;
    int P, C, n_pad, n_tiles;
    float class_thresh;
    float4* box4;
    float2* cc2;
    uint32_t* orig;
    uint32_t* key;
    uint32_t* pay;
    uint32_t* tile_count;
    uint32_t* count;
};

// argmax step for a rotated scan order: first maximal index wins, first NaN wins over everything
__device__ __forceinline__ void argmax_rot(float v, int c, float& best, int& besti) {
    const bool vn = v != v, bn = best != best;
    bool take;
    if (vn) take = !bn || c < besti;
    else if (bn) take = false;
    else take = v > best || (v == best && c < besti);
    if (take) { best = v; besti = c; }
}

// DIRECT: every thread reads its own row straight from global memory with 128-bit loads (C % 4 == 0).  A warp load
// touches 32 rows, i.e. 32 different sectors, half of each; the other half is used by the thread's next load and is
// served by L1.  No shared-memory staging, no barrier and the plain sequential first-max argmax (5 instructions per
// element instead of ~14 for the staged, rotated scan).
template <bool VEC4, bool DIRECT>
__global__ void __launch_bounds__(kPriRows) prior_decode_filter_kernel(const K1pParams p) {
    extern __shared__ float s_cls[];           // [kPriRows][cc] (staged variants)
    __shared__ int s_scan[33];
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int C = p.C;
    int run = 0;                                // survivors written so far in this tile
    const size_t img = (size_t)b * p.n_pad;

    for (int round = 0; round < kTile / kPriRows; ++round) {
        const int p0 = tile * kTile + round * kPriRows;
        const int nrows = min(kPriRows, p.P - p0);           // may be <= 0 (uniform)
        float best = 0.f;
        int besti = 0;
        bool first = true;
        if (DIRECT) {
            if (tid < nrows) {
                const float4* row = reinterpret_cast<const float4*>(p.cls + ((size_t)b * p.P + p0 + tid) * C);
                const int n4 = C >> 2;
                constexpr int U = 5;
                int c4 = 0;
                {
                    const float4 v = __ldg(row);
                    best = v.x; besti = 0;
                    if (!(v.y <= best) && (best == best)) { best = v.y; besti = 1; }
                    if (!(v.z <= best) && (best == best)) { best = v.z; besti = 2; }
                    if (!(v.w <= best) && (best == best)) { best = v.w; besti = 3; }
                    c4 = 1;
                }
                for (; c4 + U <= n4; c4 += U) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = __ldg(row + c4 + u);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int c = (c4 + u) << 2;
                        if (!(v[u].x <= best) && (best == best)) { best = v[u].x; besti = c; }
                        if (!(v[u].y <= best) && (best == best)) { best = v[u].y; besti = c + 1; }
                        if (!(v[u].z <= best) && (best == best)) { best = v[u].z; besti = c + 2; }
                        if (!(v[u].w <= best) && (best == best)) { best = v[u].w; besti = c + 3; }
                    }
                }
                for (; c4 < n4; ++c4) {
                    const float4 v = __ldg(row + c4);
                    const int c = c4 << 2;
                    if (!(v.x <= best) && (best == best)) { best = v.x; besti = c; }
                    if (!(v.y <= best) && (best == best)) { best = v.y; besti = c + 1; }
                    if (!(v.z <= best) && (best == best)) { best = v.z; besti = c + 2; }
                    if (!(v.w <= best) && (best == best)) { best = v.w; besti = c + 3; }
                }
            }
        } else
        for (int c0 = 0; c0 < C; c0 += kPriMaxCC) {
            const int cc = min(kPriMaxCC, C - c0);
            __syncthreads();                                  // staging buffer free
            if (nrows > 0) {
                const float* src = p.cls + ((size_t)b * p.P + p0) * C;
                if (VEC4 && cc == C) {
                    // one contiguous block of nrows*C floats
                    const int nvec = nrows * C / 4;
                    for (int i = tid; i < nvec; i += kPriRows) {
                        const float4 v = ldg_stream4(src + (size_t)i * 4);
                        reinterpret_cast<float4*>(s_cls)[i] = v;
                    }
                } else {
                    const int tot = nrows * cc;
                    for (int i = tid; i < tot; i += kPriRows) {
                        const int r = i / cc, c = i - r * cc;
                        s_cls[i] = ldg_stream1(src + (size_t)r * C + c0 + c);
                    }
                }
            }
            __syncthreads();
            if (tid < nrows) {
                const float* row = s_cls + (size_t)tid * cc;
                int k = (cc & 1) ? 0 : (tid % cc);           // rotation only needed for even strides
                for (int it = 0; it < cc; ++it) {
                    const float v = row[k];
                    if (first) { best = v; besti = c0 + k; first = false; }
                    else argmax_rot(v, c0 + k, best, besti);
                    if (++k == cc) k = 0;
                }
            }
        }

        bool keep = false;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f;
        const int pi = p0 + tid;
        if (tid < nrows) {
            const float4 l = *reinterpret_cast<const float4*>(p.loc + ((size_t)b * p.P + pi) * 4);
            const float4 pr = *reinterpret_cast<const float4*>(p.priors + (size_t)pi * 4);
            bx = prior_decode_box(l, pr);
            score = sigmoidf_acc(best);                                       // SSD.py:260
            keep = score > p.class_thresh;                                    // SSD.py:261
        }
        int total;
        const int ex = block_exclusive_scan(keep ? 1 : 0, s_scan, &total);
        if (keep) {
            const uint32_t slot = (uint32_t)(tile * kTile + run + ex);
            const size_t i = img + slot;
            p.box4[i] = bx;
            p.cc2[i] = make_float2(0.0f, score);
            p.orig[i] = (uint32_t)pi;
            p.key[i] = score_sort_key(score);
            p.pay[i] = ((uint32_t)besti << kSlotBits) | slot;
        }
        run += total;
    }
    if (tid == 0) {
        p.tile_count[(size_t)b * p.n_tiles + tile] = (uint32_t)run;
        if (run && p.count) atomicAdd(&p.count[b], (uint32_t)run);
    }
}

// exclusive scan of tile_count per image -> filtered-space index of the first slot of every tile
__global__ void __launch_bounds__(256) tile_prefix_kernel(const uint32_t* __restrict__ tile_count,
                                                          uint32_t* __restrict__ tile_prefix, int n_tiles) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    int carry = 0;
    for (int t0 = 0; t0 < n_tiles; t0 += 256) {
        const int t = t0 + threadIdx.x;
        int v = t < n_tiles ? (int)tile_count[(size_t)b * n_tiles + t] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_scan, &total);
        if (t < n_tiles) tile_prefix[(size_t)b * n_tiles + t] = (uint32_t)(carry + ex);
        carry += total;
    }
}

// from yolo_decode.cu / segsort.cu / nms.cu
int zero_fill_launch(void* p, size_t bytes, cudaStream_t st);
bool topk_select_supported(int k);
int topk_select_launch(const uint32_t* tile_count, const uint32_t* key, const uint32_t* pay, uint32_t* pay_out, int n_pad,
                       int n_tiles, int k, int batch, cudaStream_t st, uint32_t* count_out, uint32_t* tile_prefix_out);
int score_sort_launch(const uint32_t* tile_count, const uint32_t* count, uint32_t* digit_hist, uint32_t* ticket,
                      uint32_t* status, uint32_t* key[2], uint32_t* pay[2], int n_pad, int n_tiles, int batch,
                      cudaStream_t st);
struct NmsParams;
int prior_nms_launch_raw(const uint32_t* count, const uint32_t* spay, const float4* box4, const float2* cc2,
                         float4* kbox, uint32_t* kpos, int n_pad, float thr, int topk, int compat,
                         const uint32_t* tile_prefix, int n_tiles, const uint32_t* orig, const float* loc,
                         const float* cls, const float* priors, int P, int C, float* out_rows, int32_t* out_index, int32_t* out_count,
                         int batch, int mode_min, cudaStream_t st);

static int prior_validate(const b200det_prior_desc* d) {
    B2_CHECK_ARG(d != nullptr, "desc is null");
    B2_CHECK_ARG(d->batch > 0 && d->num_priors > 0 && d->num_classes > 0, "batch/priors/classes must be > 0");
    B2_CHECK_LIMIT(d->batch <= 65535, "batch %d > 65535", d->batch);
    B2_CHECK_LIMIT(d->num_priors <= B200DET_MAX_CANDIDATES, "num_priors %d > %d", d->num_priors, B200DET_MAX_CANDIDATES);
    B2_CHECK_LIMIT(d->num_classes <= B200DET_MAX_CLASSES, "num_classes %d > %d", d->num_classes, B200DET_MAX_CLASSES);
    B2_CHECK_ARG(d->loc && d->cls && d->priors, "loc/cls/priors is null");
    B2_CHECK_ARG((((uintptr_t)d->loc | (uintptr_t)d->priors) & 15) == 0, "loc and priors must be 16-byte aligned");
    B2_CHECK_ARG(d->topk > 0, "topk must be > 0");
    return 0;
}

size_t prior_workspace_bytes(const b200det_prior_desc* d) {
    PriorWs w;
    prior_ws_layout(d, nullptr, &w);
    return w.total_bytes;
}

static bool prior_use_select(const b200det_prior_desc* d) {
    const char* tk = getenv("B200DET_TOPK");
    return topk_select_supported(d->topk) && !(tk && strcmp(tk, "sort") == 0);
}

static int prior_check_ws(const b200det_prior_desc* d, void* ws, size_t ws_bytes, PriorWs* w) {
    int rc = prior_validate(d);
    if (rc) return rc;
    B2_CHECK_ARG(ws != nullptr && ((uintptr_t)ws & 255) == 0, "workspace must be non-null and 256-byte aligned");
    prior_ws_layout(d, ws, w);
    if (ws_bytes < w->total_bytes) {
        set_error("workspace too small: %zu < %zu", ws_bytes, w->total_bytes);
        return B200DET_EWORKSPACE;
    }
    return 0;
}

// stage 1 of the pipeline: K1' (prior decode + sigmoid-argmax + score filter + ordered tile compaction)
int prior_stage_decode(const b200det_prior_desc* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    PriorWs w;
    int rc = prior_check_ws(d, ws, ws_bytes, &w);
    if (rc) return rc;
    // with the radix select (the default) the select kernel derives the per-image count and the tile prefix itself: three
    // launches per step (decode+filter, select, NMS); the full-sort path needs the zeroed counters and the prefix kernel
    const bool use_select = prior_use_select(d);
    if (!use_select) rc = zero_fill_launch(w.count, w.zero_bytes, st);
    if (rc) return rc;

    K1pParams p;
    memset(&p, 0, sizeof(p));
    p.loc = d->loc; p.cls = d->cls; p.priors = d->priors;
    p.P = d->num_priors; p.C = d->num_classes; p.n_pad = w.n_pad; p.n_tiles = w.n_tiles;
    p.class_thresh = d->class_thresh;
    p.box4 = w.box4; p.cc2 = w.cc2; p.orig = w.orig; p.key = w.key[0]; p.pay = w.pay[0];
    p.tile_count = w.tile_count; p.count = use_select ? nullptr : w.count;
    const int cc = d->num_classes < kPriMaxCC ? d->num_classes : kPriMaxCC;
    const size_t smem = (size_t)kPriRows * cc * sizeof(float);
    dim3 grid(w.n_tiles, d->batch);
    // 128-bit staging needs every 128-row block to start 16-byte aligned and hold a multiple of 4 floats
    const bool vec4 = d->num_classes <= kPriMaxCC && d->num_classes % 4 == 0 && ((uintptr_t)d->cls & 15) == 0;
    const bool direct = d->num_classes % 4 == 0 && ((uintptr_t)d->cls & 15) == 0 &&
                        !(getenv("B200DET_PRIOR_K1") && strcmp(getenv("B200DET_PRIOR_K1"), "staged") == 0);
    if (direct) {
        prior_decode_filter_kernel<true, true><<<grid, kPriRows, 0, st>>>(p);
    } else if (vec4) {
        B2_CUDA(cudaFuncSetAttribute(prior_decode_filter_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prior_decode_filter_kernel<true, false><<<grid, kPriRows, smem, st>>>(p);
    } else {
        B2_CUDA(cudaFuncSetAttribute(prior_decode_filter_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prior_decode_filter_kernel<false, false><<<grid, kPriRows, smem, st>>>(p);
    }
    B2_LAUNCH_CHECK("prior_decode_filter_kernel");
    if (!use_select) {
        tile_prefix_kernel<<<d->batch, 256, 0, st>>>(w.tile_count, w.tile_prefix, w.n_tiles);
        B2_LAUNCH_CHECK("tile_prefix_kernel");
    }
    return 0;
}

// stages 2 + 3: top-k selection and the class-agnostic greedy NMS on the selected rows
int prior_stage_select_nms(const b200det_prior_desc* d, void* ws, size_t ws_bytes, float* out_rows, int32_t* out_index,
                           int32_t* out_count, int32_t* cand_count, cudaStream_t st) {
    PriorWs w;
    int rc = prior_check_ws(d, ws, ws_bytes, &w);
    if (rc) return rc;
    B2_CHECK_ARG(out_rows && out_count, "out_rows / out_count is null");
    const bool use_select = prior_use_select(d);
    // only the topk best-scoring candidates reach the NMS (SSD.py:273): radix select instead of a full sort
    // (B200DET_TOPK=sort keeps the full sort, the A/B reference of the tests)
    if (use_select)
        rc = topk_select_launch(w.tile_count, w.key[0], w.pay[0], w.pay[0], w.n_pad, w.n_tiles, d->topk, d->batch, st, w.count,
                                w.tile_prefix);
    else
        rc = score_sort_launch(w.tile_count, w.count, w.digit_hist, w.ticket, w.status, w.key, w.pay, w.n_pad, w.n_tiles, d->batch, st);
    if (rc) return rc;
    rc = prior_nms_launch_raw(w.count, w.pay[0], w.box4, w.cc2, w.kbox, w.kpos, w.n_pad, d->nms_thresh, d->topk,
                              d->compat, w.tile_prefix, w.n_tiles, w.orig, d->loc, d->cls, d->priors, d->num_priors, d->num_classes,
                              out_rows, out_index, out_count, d->batch, d->mode_min, st);
    if (rc) return rc;
    if (cand_count)
        B2_CUDA(cudaMemcpyAsync(cand_count, w.count, (size_t)d->batch * 4, cudaMemcpyDefault, st));   // cand_count may be mapped host memory
    return 0;
}

int prior_nms_pipeline(const b200det_prior_desc* d, void* ws, size_t ws_bytes, float* out_rows, int32_t* out_index,
                       int32_t* out_count, int32_t* cand_count, cudaStream_t st) {
    B2_CHECK_ARG(out_rows && out_count, "out_rows / out_count is null");
    int rc = prior_stage_decode(d, ws, ws_bytes, st);
    if (rc) return rc;
    return prior_stage_select_nms(d, ws, ws_bytes, out_rows, out_index, out_count, cand_count, st);
}

}  // namespace b200det
