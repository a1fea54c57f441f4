// K3 — greedy NMS over score-sorted segments, one CTA per segment.
//
//   VARIANT 0: YOLO class-aware merge-NMS (model/YOLOV3.py:320-333): segment = one (image, class);
//              IoU_+1 with +1e-16 (accuracy.py:54-68) > nms_thres removes a row and attributes it
//              to the FIRST keeper that hit it; the keeper's box becomes the conf-weighted mean of
//              its cluster (YOLOV3.py:327-329).
//   VARIANT 1/2: SSD / RetinaNet class-agnostic greedy NMS on the top-k rows of an image
//              (model/SSD.py:270-302), 'union' / 'min' overlap, survive when ovr <= thresh.
//
// A segment is processed in chunks of 512 score-ordered rows:
//   phase A  rows of the chunk vs. the keepers of EARLIER chunks (first hit = owner, early exit);
//   phase B  lower-triangular 64-bit overlap masks inside the chunk (each (row, word) item = 64 IoUs,
//            keeper box broadcast from shared memory, row box in registers);
//   sweep    one warp resolves the chunk serially-exact: 32 rows per step, prior words tested in
//            parallel per lane, the 32x32 diagonal resolved with 32 ballots;
//   owners   first kept bit of (mask & kept) -> cluster owner;  merge sums pulled per keeper in row
//            order (deterministic fp32 order: keeper first, then members by descending score).
// IoU arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) and IEEE division so
// the keep decisions are bit-identical to the reference's fp32 CPU path.
#include "yolo_ws.cuh"

namespace b200det {

constexpr int kNmsT = 512;                 // rows per chunk
constexpr int kNmsThreads = 256;
constexpr int kNmsW = kNmsT / 64;          // mask words per full row
constexpr int kNmsTriWords = 32 * kNmsW * (kNmsW + 1);   // packed lower-triangular rows
constexpr int kNmsStage = 256;             // earlier keepers staged per phase-A round

struct NmsParams {
    const uint32_t* seg_off;    // [B][C+1]     VARIANT 0
    const uint32_t* count;      // [B]
    const uint32_t* spay;       // [B][n_pad]   sorted payload (class << 20 | slot)
    const uint32_t* srank;      // [B][n_pad]   score rank of a sorted position (VARIANT 0)
    const float4* box4;         // [B][n_pad]   by slot
    const float2* cc2;          // [B][n_pad]   by slot
    uint32_t* kpay;             // [B][n_pad]   by rank (VARIANT 0)
    float4* mbox;               // [B][n_pad]   by rank (VARIANT 0)
    float4* kbox;               // [B][n_pad]   scratch: keepers' original boxes, per segment
    float* kacc;                // [B][n_pad][5]
    uint32_t* kpos;             // [B][n_pad]
    int n_pad, C;
    float thr;
    // VARIANT 1/2 (prior NMS)
    int topk, compat;
    const uint32_t* tile_prefix;   // [B][n_tiles] exclusive scan of tile_count (filtered index of a slot)
    int n_tiles;
    const uint32_t* orig;          // [B][n_pad] prior index of a slot
    const float4* dense_box;       // [B][P] decoded boxes of ALL priors (quirk ii)
    const int32_t* dense_label;    // [B][P]
    int P;
    float* out_rows;               // [B][topk][7]
    int32_t* out_index;            // [B][topk] or null
    int32_t* out_count;            // [B]
};

__device__ __forceinline__ int tri_off(int j) {
    const int q = j >> 6;
    return 32 * q * (q + 1) + (j & 63) * (q + 1);
}

__device__ __forceinline__ float box_area_plus1(const float4 b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// true when the earlier (higher-score) box `a` removes box `b`
template <int VARIANT>
__device__ __forceinline__ bool removes(const float4 a, const float aa, const float4 b, const float ab, const float thr) {
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
    const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float iw = fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.0f), 0.0f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.0f), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    if (VARIANT == 0) {
        const float uni = __fadd_rn(__fsub_rn(__fadd_rn(aa, ab), inter), 1e-16f);   // accuracy.py:66
        return __fdiv_rn(inter, uni) > thr;                                        // YOLOV3.py:323
    } else if (VARIANT == 1) {
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));    // SSD.py:292
        return !(ovr <= thr);                                                      // SSD.py:298
    } else {
        const float ovr = __fdiv_rn(inter, fminf(ab, aa));                          // SSD.py:294
        return !(ovr <= thr);
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(kNmsThreads) nms_segment_kernel(const NmsParams p) {
    __shared__ float4 s_box[kNmsT];
    __shared__ float s_area[kNmsT];
    __shared__ float s_conf[kNmsT];
    __shared__ unsigned long long s_L[kNmsTriWords];
    __shared__ unsigned long long s_kept[kNmsW];
    __shared__ int s_wpre[kNmsW + 1];
    __shared__ int s_own[kNmsT];
    __shared__ int s_pre[kNmsT];
    __shared__ float4 s_kb[kNmsStage];
    __shared__ float s_ka[kNmsStage];
    __shared__ int s_last_members;

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const size_t img = (size_t)b * p.n_pad;
    int s, e;
    if (VARIANT == 0) {
        const uint32_t* so = p.seg_off + (size_t)b * (p.C + 1) + blockIdx.x;
        s = (int)so[0]; e = (int)so[1];
    } else {
        s = 0; e = min((int)p.count[b], p.topk);                 // SSD.py:273
    }
    const int n = e - s;
    if (VARIANT == 0 && n <= 0) return;
    const bool single = n <= kNmsT;
    const float thr = p.thr;

    int Kprev = 0;          // keepers found in earlier chunks
    int last_k = -1;        // VARIANT 1/2: index of the last keeper so far
    if (tid == 0) s_last_members = 0;

    for (int c0 = s; c0 < e; c0 += kNmsT) {
        const int nc = min(kNmsT, e - c0);
        const int Wc = (nc + 63) >> 6;

        // ---- load the chunk ----------------------------------------------------------------
        for (int j = tid; j < nc; j += kNmsThreads) {
            const uint32_t pay = p.spay[img + c0 + j];
            const uint32_t slot = pay & kSlotMask;
            const float4 bx = p.box4[img + slot];
            s_box[j] = bx;
            s_area[j] = box_area_plus1(bx);
            s_conf[j] = p.cc2[img + slot].x;
            s_pre[j] = -1;
        }
        __syncthreads();

        // ---- phase A: against keepers of earlier chunks ----------------------------------------
        for (int kt = 0; kt < Kprev; kt += kNmsStage) {
            const int nk = min(kNmsStage, Kprev - kt);
            if (tid < nk) {
                const float4 kb = p.kbox[img + s + kt + tid];
                s_kb[tid] = kb;
                s_ka[tid] = box_area_plus1(kb);
            }
            __syncthreads();
            for (int j = tid; j < nc; j += kNmsThreads) {
                if (s_pre[j] >= 0) continue;
                const float4 bj = s_box[j];
                const float aj = s_area[j];
                for (int k = 0; k < nk; ++k) {
                    if (removes<VARIANT>(s_kb[k], s_ka[k], bj, aj, thr)) { s_pre[j] = kt + k; break; }
                }
            }
            __syncthreads();
        }

        // ---- phase B: lower-triangular overlap masks inside the chunk --------------------------
        for (int w = 0; w < Wc; ++w) {
            const int i0 = w << 6;
            for (int j = i0 + tid; j < nc; j += kNmsThreads) {
                unsigned long long bits = 0ull;
                if (s_pre[j] < 0) {
                    const float4 bj = s_box[j];
                    const float aj = s_area[j];
                    const int iend = min(i0 + 64, j);
                    for (int i = i0; i < iend; ++i) {
                        if (removes<VARIANT>(s_box[i], s_area[i], bj, aj, thr)) bits |= 1ull << (i - i0);
                    }
                }
                s_L[tri_off(j) + w] = bits;
            }
        }
        __syncthreads();

        // ---- sweep: exact greedy resolution by warp 0, 32 rows per step -------------------------
        if (tid < 32) {
            const int lane = tid;
            if (lane < kNmsW) s_kept[lane] = 0ull;
            __syncwarp();
            const int ngroups = (nc + 31) >> 5;
            for (int g = 0; g < ngroups; ++g) {
                const int j = (g << 5) + lane;
                const int wl = g >> 1;                       // word holding this group
                const bool valid = j < nc && s_pre[j] < 0;
                bool hit = false;
                unsigned m = 0;
                if (valid) {
                    const unsigned long long* row = &s_L[tri_off(j)];
                    for (int w = 0; w < wl; ++w) hit |= (row[w] & s_kept[w]) != 0ull;
                    const unsigned long long lw = row[wl];
                    if (g & 1) {
                        hit |= (lw & s_kept[wl] & 0xFFFFFFFFull) != 0ull;
                        m = (unsigned)(lw >> 32);
                    } else {
                        m = (unsigned)lw;
                    }
                }
                const bool pre = valid && !hit;
                const unsigned cand = __ballot_sync(0xFFFFFFFFu, pre);
                unsigned kg = 0;
                if (cand) {
                    const int top = 31 - __clz(cand);
                    for (int sft = 0; sft <= top; ++sft) {
                        const bool bit = pre && ((m & kg) == 0u);
                        const unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
                        kg = bal & (sft == 31 ? 0xFFFFFFFFu : ((2u << sft) - 1u));
                    }
                }
                if (lane == 0 && kg) s_kept[wl] |= (unsigned long long)kg << ((g & 1) * 32);
                __syncwarp();
            }
            if (lane == 0) {
                int run = 0;
                for (int w = 0; w < kNmsW; ++w) { s_wpre[w] = run; run += (w < Wc) ? __popcll(s_kept[w]) : 0; }
                s_wpre[kNmsW] = run;
            }
        }
        __syncthreads();
        const int Kc = s_wpre[kNmsW];

        // ---- owners ------------------------------------------------------------------------------
        bool any_cross_local = false;
        for (int j = tid; j < nc; j += kNmsThreads) {
            int own;
            bool is_keeper = false;
            if (s_pre[j] >= 0) {
                own = s_pre[j];
                any_cross_local = true;
            } else {
                const int wj = j >> 6;
                const unsigned long long below = (1ull << (j & 63)) - 1ull;
                if ((s_kept[wj] >> (j & 63)) & 1ull) {
                    own = Kprev + s_wpre[wj] + __popcll(s_kept[wj] & below);
                    is_keeper = true;
                } else {
                    const unsigned long long* row = &s_L[tri_off(j)];
                    int i = -1;
                    for (int w = 0; w <= wj; ++w) {
                        const unsigned long long h = row[w] & s_kept[w];
                        if (h) { i = (w << 6) + __ffsll((long long)h) - 1; break; }
                    }
                    // i >= 0 always: a non-kept, non-presuppressed row was hit by a kept row
                    if (i >= 0) {
                        const int wi = i >> 6;
                        own = Kprev + s_wpre[wi] + __popcll(s_kept[wi] & ((1ull << (i & 63)) - 1ull));
                    } else {
                        own = -2;
                    }
                }
            }
            s_own[j] = own;
            if (VARIANT == 0) {
                if (!is_keeper) p.kpay[img + p.srank[img + c0 + j]] = kNone;
            }
            if (is_keeper && !(VARIANT == 0 && single)) {
                p.kbox[img + s + own] = s_box[j];
                p.kpos[img + s + own] = (uint32_t)(c0 + j);
            }
        }
        const int any_cross = __syncthreads_or(any_cross_local ? 1 : 0);

        if (VARIANT == 0) {
            // ---- merge sums, pulled per keeper in row order ----------------------------------------
            for (int j = tid; j < nc; j += kNmsThreads) {
                const int wj = j >> 6;
                if (s_pre[j] >= 0 || !((s_kept[wj] >> (j & 63)) & 1ull)) continue;
                const int kidx = s_own[j];
                const float4 bj = s_box[j];
                const float w0 = s_conf[j];
                float ax = __fmul_rn(w0, bj.x), ay = __fmul_rn(w0, bj.y);
                float az = __fmul_rn(w0, bj.z), aw = __fmul_rn(w0, bj.w), ws = w0;
                for (int m = j + 1; m < nc; ++m) {
                    if (s_own[m] == kidx) {
                        const float4 bm = s_box[m];
                        const float wm = s_conf[m];
                        ax = __fadd_rn(ax, __fmul_rn(wm, bm.x));
                        ay = __fadd_rn(ay, __fmul_rn(wm, bm.y));
                        az = __fadd_rn(az, __fmul_rn(wm, bm.z));
                        aw = __fadd_rn(aw, __fmul_rn(wm, bm.w));
                        ws = __fadd_rn(ws, wm);
                    }
                }
                if (single) {
                    const uint32_t r = p.srank[img + c0 + j];
                    p.kpay[img + r] = p.spay[img + c0 + j];
                    p.mbox[img + r] = make_float4(__fdiv_rn(ax, ws), __fdiv_rn(ay, ws), __fdiv_rn(az, ws), __fdiv_rn(aw, ws));
                } else {
                    float* acc = p.kacc + (img + s + kidx) * 5;
                    acc[0] = ax; acc[1] = ay; acc[2] = az; acc[3] = aw; acc[4] = ws;
                }
            }
            if (any_cross) {
                // rows of this chunk owned by keepers of earlier chunks
                for (int kidx = tid; kidx < Kprev; kidx += kNmsThreads) {
                    float* acc = p.kacc + (img + s + kidx) * 5;
                    float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f, ws = 0.f;
                    bool loaded = false;
                    for (int m = 0; m < nc; ++m) {
                        if (s_own[m] == kidx) {
                            if (!loaded) { ax = acc[0]; ay = acc[1]; az = acc[2]; aw = acc[3]; ws = acc[4]; loaded = true; }
                            const float4 bm = s_box[m];
                            const float wm = s_conf[m];
                            ax = __fadd_rn(ax, __fmul_rn(wm, bm.x));
                            ay = __fadd_rn(ay, __fmul_rn(wm, bm.y));
                            az = __fadd_rn(az, __fmul_rn(wm, bm.z));
                            aw = __fadd_rn(aw, __fmul_rn(wm, bm.w));
                            ws = __fadd_rn(ws, wm);
                        }
                    }
                    if (loaded) { acc[0] = ax; acc[1] = ay; acc[2] = az; acc[3] = aw; acc[4] = ws; }
                }
            }
        } else {
            // ---- members of the (so far) last keeper, for the drop-last quirk (SSD.py:277-278) -----
            if (Kc > 0) {
                last_k = Kprev + Kc - 1;
                __syncthreads();
                if (tid == 0) s_last_members = 0;
                __syncthreads();
            }
            if (last_k >= 0) {
                int cnt = 0;
                for (int j = tid; j < nc; j += kNmsThreads) {
                    const bool is_keeper = s_pre[j] < 0 && ((s_kept[j >> 6] >> (j & 63)) & 1ull);
                    if (!is_keeper && s_own[j] == last_k) ++cnt;
                }
                if (cnt) atomicAdd(&s_last_members, cnt);
            }
        }
        Kprev += Kc;
        __syncthreads();   // kacc / kbox writes of this chunk are visible to the next one (same CTA)
    }

    if (VARIANT == 0) {
        if (!single) {
            for (int kidx = tid; kidx < Kprev; kidx += kNmsThreads) {
                const float* acc = p.kacc + (img + s + kidx) * 5;
                const uint32_t pos = p.kpos[img + s + kidx];
                const uint32_t r = p.srank[img + pos];
                const float ws = acc[4];
                p.kpay[img + r] = p.spay[img + pos];
                p.mbox[img + r] = make_float4(__fdiv_rn(acc[0], ws), __fdiv_rn(acc[1], ws), __fdiv_rn(acc[2], ws),
                                              __fdiv_rn(acc[3], ws));
            }
        }
    } else {
        int K = Kprev;
        if (p.compat && K > 0 && s_last_members == 0) K -= 1;          // SSD.py:277-278
        if (tid == 0) p.out_count[b] = K;
        const uint32_t* tp = p.tile_prefix + (size_t)b * p.n_tiles;
        for (int kidx = tid; kidx < K; kidx += kNmsThreads) {
            const uint32_t pos = p.kpos[img + kidx];
            const uint32_t pay = p.spay[img + pos];
            const uint32_t slot = pay & kSlotMask;
            float4 bx = p.box4[img + slot];
            int label = (int)(pay >> kSlotBits);
            const float score = p.cc2[img + slot].y;
            int src = (int)p.orig[img + slot];
            if (p.compat) {
                // quirk (ii): `keep` indexes the score-filtered set, but boxes/labels are gathered from the
                // UNFILTERED arrays with it (SSD.py:303-307)
                const int f = (int)(tp[slot >> kTileShift] + (slot & (kTile - 1)));
                bx = p.dense_box[(size_t)b * p.P + f];
                label = p.dense_label[(size_t)b * p.P + f];
                src = f;
            }
            float* o = p.out_rows + ((size_t)b * p.topk + kidx) * 7;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = 0.0f; o[5] = score; o[6] = (float)label;
            if (p.out_index) p.out_index[(size_t)b * p.topk + kidx] = src;   // prior whose box/label the row carries
        }
    }
}

// K3b — ordered emit: compact the kept rows of every image in ascending score rank.
struct EmitParams {
    const uint32_t* count;
    const uint32_t* kpay;
    const float4* mbox;
    const float2* cc2;
    const uint32_t* orig;
    float* out_rows;       // [B][n_pad][7]
    int32_t* out_index;    // [B][n_pad] or null
    int32_t* out_count;    // [B]
    int n_pad;
};

__global__ void __launch_bounds__(1024) yolo_emit_kernel(const EmitParams p) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    const size_t img = (size_t)b * p.n_pad;
    const int n = (int)p.count[b];
    int base = 0;
    for (int r0 = 0; r0 < n; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        uint32_t pay = kNone;
        if (r < n) pay = p.kpay[img + r];
        const int flag = pay != kNone ? 1 : 0;
        int total;
        const int ex = block_exclusive_scan(flag, s_scan, &total);
        if (flag) {
            const uint32_t slot = pay & kSlotMask;
            const float4 mb = p.mbox[img + r];
            const float2 cc = p.cc2[img + slot];
            float* o = p.out_rows + (img + base + ex) * 7;
            o[0] = mb.x; o[1] = mb.y; o[2] = mb.z; o[3] = mb.w;
            o[4] = cc.x; o[5] = cc.y; o[6] = (float)(pay >> kSlotBits);     // YOLOV3.py:318-319
            if (p.out_index) p.out_index[img + base + ex] = (int32_t)p.orig[img + slot];
        }
        base += total;
    }
    if (threadIdx.x == 0) p.out_count[b] = base;
}

int yolo_stage_nms(const b200det_yolo_desc* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = yolo_validate(d, ws, ws_bytes);
    if (rc) return rc;
    YoloWs w;
    yolo_ws_layout(d, ws, &w);
    NmsParams p;
    memset(&p, 0, sizeof(p));
    p.seg_off = w.seg_off; p.count = w.count;
    p.spay = yolo_sorted_pay(w); p.srank = yolo_sorted_rank(w);
    p.box4 = w.box4; p.cc2 = w.cc2; p.kpay = w.kpay; p.mbox = w.mbox;
    p.kbox = w.kbox; p.kacc = w.kacc; p.kpos = w.kpos;
    p.n_pad = w.n_pad; p.C = w.C; p.thr = d->nms_thres;
    dim3 grid(d->num_classes, d->batch);
    nms_segment_kernel<0><<<grid, kNmsThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("nms_segment_kernel<0>");
    return 0;
}

int yolo_stage_emit(const b200det_yolo_desc* d, void* ws, size_t ws_bytes, float* out_rows, int32_t* out_index,
                    int32_t* out_count, cudaStream_t st) {
    int rc = yolo_validate(d, ws, ws_bytes);
    if (rc) return rc;
    B2_CHECK_ARG(out_rows && out_count, "out_rows / out_count is null");
    YoloWs w;
    yolo_ws_layout(d, ws, &w);
    EmitParams p;
    p.count = w.count; p.kpay = w.kpay; p.mbox = w.mbox; p.cc2 = w.cc2; p.orig = w.orig;
    p.out_rows = out_rows; p.out_index = out_index; p.out_count = out_count; p.n_pad = w.n_pad;
    yolo_emit_kernel<<<d->batch, 1024, 0, st>>>(p);
    B2_LAUNCH_CHECK("yolo_emit_kernel");
    return 0;
}

int prior_nms_launch_raw(const uint32_t* count, const uint32_t* spay, const float4* box4, const float2* cc2,
                         float4* kbox, uint32_t* kpos, int n_pad, float thr, int topk, int compat,
                         const uint32_t* tile_prefix, int n_tiles, const uint32_t* orig, const float4* dense_box,
                         const int32_t* dense_label, int P, float* out_rows, int32_t* out_index, int32_t* out_count,
                         int batch, int mode_min, cudaStream_t st) {
    NmsParams p;
    memset(&p, 0, sizeof(p));
    p.count = count; p.spay = spay; p.box4 = box4; p.cc2 = cc2; p.kbox = kbox; p.kpos = kpos;
    p.n_pad = n_pad; p.thr = thr; p.topk = topk; p.compat = compat; p.tile_prefix = tile_prefix; p.n_tiles = n_tiles;
    p.orig = orig; p.dense_box = dense_box; p.dense_label = dense_label; p.P = P;
    p.out_rows = out_rows; p.out_index = out_index; p.out_count = out_count;
    dim3 grid(1, batch);
    if (mode_min) nms_segment_kernel<2><<<grid, kNmsThreads, 0, st>>>(p);
    else nms_segment_kernel<1><<<grid, kNmsThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("nms_segment_kernel<prior>");
    return 0;
}

}  // namespace b200det
