// K4 — loss-side target assignment.
//   build_targets_v5        LightningFunc/accuracy.py:472-521  (ordered, scan-based compaction)
//   v5 matched-row forward  LightningFunc/losses.py:105-123    (gather + D2 decode + GIoU + tobj scatter) + backward
//   build_targets (v2-v4)   LightningFunc/accuracy.py:305-380  (grid scatter, "highest target row wins")
//   SSDLoss.match           LightningFunc/losses.py:199-218
//   RetinaNet assignment    LightningFunc/losses.py:375-403, 423-443
#include "common.cuh"
#include "boxmath.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#include "targets.cuh"

namespace b200det {

// ================================================================================================
// T2 — build_targets_v5, one level.  Element e = a*nt + t (anchor-major, target-minor: the order of
// `t.repeat(na,1,1)[j]`, accuracy.py:490).  Five flags per element: base row kept, and the four
// neighbour copies (x-left j, y-up k, x-right l, y-down m; accuracy.py:503-506).  Output row of a
// flagged element = block base + its rank inside the block, blocks laid out [base | j | k | l | m].
// One 1024-thread CTA makes two passes over the elements: totals, then ranks + writes.
// ================================================================================================
struct Tv5Params {
    const float* targets;   // [nt,6]
    int nt, na, nx, ny;
    float anc[B200DET_MAX_ANCHORS][2];
    int32_t *ob, *oa, *ogj, *ogi, *ocls;
    float* otbox;
    float* oanch;
    int32_t* ocount;
};

__device__ __forceinline__ float pymod1(float x) {      // torch `x % 1.` (sign of the divisor)
    float r = fmodf(x, 1.0f);
    if (r < 0.0f) r += 1.0f;
    return r;
}

__device__ __forceinline__ unsigned tv5_flags(const Tv5Params& p, int e, float& gx, float& gy, float& gw, float& gh,
                                              float& tb, float& tc) {
    const int a = e / p.nt, t = e - a * p.nt;
    const float* r = p.targets + (size_t)t * 6;
    tb = r[0]; tc = r[1];
    gx = __fmul_rn(r[2], (float)p.nx); gy = __fmul_rn(r[3], (float)p.ny);      // accuracy.py:483,486
    gw = __fmul_rn(r[4], (float)p.nx); gh = __fmul_rn(r[5], (float)p.ny);
    const float rw = __fdiv_rn(gw, p.anc[a][0]), rh = __fdiv_rn(gh, p.anc[a][1]);   // accuracy.py:488
    const float mw = fmaxf(rw, __fdiv_rn(1.0f, rw)), mh = fmaxf(rh, __fdiv_rn(1.0f, rh));
    if (!(fmaxf(mw, mh) < 4.0f)) return 0u;                                     // accuracy.py:489
    unsigned f = 1u;
    const float fx = pymod1(gx), fy = pymod1(gy);
    if (fx < 0.5f && gx > 1.0f) f |= 2u;                                        // j
    if (fy < 0.5f && gy > 1.0f) f |= 4u;                                        // k
    if (fx > 0.5f && gx < __fsub_rn((float)p.nx, 1.0f)) f |= 8u;                // l
    if (fy > 0.5f && gy < __fsub_rn((float)p.ny, 1.0f)) f |= 16u;               // m
    return f;
}

struct Tv5Multi {
    Tv5Params lvl[B200DET_MAX_LEVELS];
};

// One level.  CL = 1: one CTA walks all na * nt (anchor, target) pairs.  CL > 1: a thread-block cluster per level, CTA `crank`
// owns the pair range [crank * per, (crank + 1) * per); the five flag totals of every CTA are exchanged over distributed
// shared memory, so a CTA knows where its rows of each of the five concatenated groups start (accuracy.py:503: the groups are
// concatenated, each in (anchor, target) order) without a second launch.  (One 1024-thread CTA per level took 59 us at
// BASELINE config 4: ten rounds of IEEE divisions, fmodf and five block scans on one SM.)
template <int CL>
__device__ __forceinline__ void build_targets_v5_body(const Tv5Params& p, const int crank) {
    __shared__ int s_wtot[5][32];
    __shared__ int s_woff[5][32];
    __shared__ int s_carry[5];
    __shared__ int s_base[5];
    __shared__ int s_tot[5];
    const int E = p.na * p.nt;
    const int per = CL == 1 ? E : ((E + CL - 1) / CL + 1023) / 1024 * 1024;
    const int e_lo = min(E, crank * per), e_hi = min(E, e_lo + per);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 5) s_carry[tid] = 0;
    __syncthreads();

    // phase 0: totals of the five flags over the CTA's range (per-thread counts, one reduction)
    {
        int cnt[5] = {0, 0, 0, 0, 0};
        for (int e = e_lo + tid; e < e_hi; e += 1024) {
            float gx, gy, gw, gh, tb, tc;
            const unsigned f = tv5_flags(p, e, gx, gy, gw, gh, tb, tc);
#pragma unroll
            for (int k = 0; k < 5; ++k) cnt[k] += (f >> k) & 1u;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            int v = cnt[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0) s_wtot[k][warp] = v;
        }
        __syncthreads();
        if (tid < 5) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += s_wtot[tid][w];
            s_tot[tid] = tot;
        }
        if (CL == 1) {
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int k = 0; k < 5; ++k) { s_base[k] = run; run += s_tot[k]; }
                p.ocount[0] = run;
            }
        } else {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();                                   // every CTA's s_tot is written
            if (tid == 0) {
                int run = 0;
                for (int k = 0; k < 5; ++k) {
                    int before = 0, all = 0;
                    for (int c = 0; c < CL; ++c) {
                        const int v = *cluster.map_shared_rank(&s_tot[k], c);
                        all += v;
                        if (c < crank) before += v;
                    }
                    s_base[k] = run + before;
                    run += all;
                }
                if (crank == 0) p.ocount[0] = run;
            }
            cluster.sync();                                   // nobody leaves (or overwrites s_tot) while a peer still reads it
        }
        __syncthreads();
    }
    // phase 1: ordered ranks + writes
    for (int e0 = e_lo; e0 < e_hi; e0 += 1024) {
        const int e = e0 + tid;
        float gx = 0, gy = 0, gw = 0, gh = 0, tb = 0, tc = 0;
        const unsigned f = e < e_hi ? tv5_flags(p, e, gx, gy, gw, gh, tb, tc) : 0u;
        unsigned bal[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) bal[k] = __ballot_sync(0xFFFFFFFFu, (f >> k) & 1u);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) s_wtot[k][warp] = __popc(bal[k]);
        }
        __syncthreads();
        if (warp < 5) {
            const int v = s_wtot[warp][lane];
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += u;
            }
            s_woff[warp][lane] = s_carry[warp] + inc - v;
            __syncwarp();
            if (lane == 31) s_carry[warp] += inc;
        }
        __syncthreads();
        if (f) {
            const int a = e / p.nt;
            const int ib = (int)tb, ic = (int)tc;                       // .long() truncation, accuracy.py:509
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                if ((f >> k) & 1u) {
                    const int row = s_base[k] + s_woff[k][warp] + __popc(bal[k] & lanemask_lt());
                    const float ox = k == 1 ? 0.5f : (k == 3 ? -0.5f : 0.0f);   // off * g, accuracy.py:506
                    const float oy = k == 2 ? 0.5f : (k == 4 ? -0.5f : 0.0f);
                    const int gi = (int)__fsub_rn(gx, ox), gj = (int)__fsub_rn(gy, oy);   // accuracy.py:512
                    p.ob[row] = ib; p.oa[row] = a; p.ogj[row] = gj; p.ogi[row] = gi; p.ocls[row] = ic;
                    float* tbx = p.otbox + (size_t)row * 4;
                    tbx[0] = __fsub_rn(gx, (float)gi); tbx[1] = __fsub_rn(gy, (float)gj);   // accuracy.py:517
                    tbx[2] = gw; tbx[3] = gh;
                    p.oanch[(size_t)row * 2] = p.anc[a][0];
                    p.oanch[(size_t)row * 2 + 1] = p.anc[a][1];
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) build_targets_v5_kernel(const Tv5Params p) { build_targets_v5_body<1>(p, 0); }

// all levels in one launch: one CTA per level (the levels are independent, accuracy.py:482)
constexpr int kTv5Cluster = 8;
__global__ void __launch_bounds__(1024) build_targets_v5_multi_kernel(const Tv5Multi m) {
    build_targets_v5_body<kTv5Cluster>(m.lvl[blockIdx.y], (int)blockIdx.x);       // grid (cluster, levels), cluster dims (8, 1, 1)
}

// ================================================================================================
// T4 — matched rows of one level: ps = pi[b,a,gj,gi] ; pxy = sigmoid*2-0.5 ; pwh = (sigmoid*2)^2*anch ;
// giou = bbox_iou_v5(pbox, tbox, xywh, GIoU) ; tobj[cell] = clamp(giou, 0), last row wins.
// ================================================================================================
__device__ __forceinline__ int match_rows(const MatchParams& p) { return p.m_dev ? min(p.m, (int)*p.m_dev) : p.m; }

__device__ __forceinline__ long long match_cell(const MatchParams& p, int i) {
    return (((long long)p.b[i] * p.na + p.a[i]) * p.ny + p.gj[i]) * p.nx + p.gi[i];
}

__device__ __forceinline__ void v5_match_fwd_body(const MatchParams& p, float* __restrict__ giou, int* __restrict__ tobj_as_int,
                                                  const int bid) {
    const int i = bid * blockDim.x + threadIdx.x;
    if (i >= match_rows(p)) return;
    const long long cell = match_cell(p, i);
    const float* ps = p.pi + cell * p.F;
    const float sx = sigmoidf_acc(ps[0]), sy = sigmoidf_acc(ps[1]);
    const float sw = sigmoidf_acc(ps[2]), sh = sigmoidf_acc(ps[3]);
    float4 pb;
    pb.x = __fsub_rn(__fmul_rn(sx, 2.0f), 0.5f);                                    // losses.py:115
    pb.y = __fsub_rn(__fmul_rn(sy, 2.0f), 0.5f);
    const float w2 = __fmul_rn(sw, 2.0f), h2 = __fmul_rn(sh, 2.0f);
    pb.z = __fmul_rn(__fmul_rn(w2, w2), p.anch[(size_t)i * 2]);                     // losses.py:116
    pb.w = __fmul_rn(__fmul_rn(h2, h2), p.anch[(size_t)i * 2 + 1]);
    const float4 tb = *reinterpret_cast<const float4*>(p.tbox + (size_t)i * 4);
    giou[i] = iou_v5_forward(pb, tb, false, B200DET_GIOU);                          // losses.py:118
    atomicMax(&tobj_as_int[cell], i + 1);                                           // winner = highest row
}
__global__ void v5_match_fwd_kernel(const MatchParams p, float* __restrict__ giou, int* __restrict__ tobj_as_int) {
    v5_match_fwd_body(p, giou, tobj_as_int, (int)blockIdx.x);
}

__device__ __forceinline__ void v5_match_tobj_body(const MatchParams& p, const float* __restrict__ giou, float* __restrict__ tobj,
                                                   const int bid) {
    const int i = bid * blockDim.x + threadIdx.x;
    if (i >= match_rows(p)) return;
    const long long cell = match_cell(p, i);
    if (reinterpret_cast<const int*>(tobj)[cell] == i + 1) tobj[cell] = fmaxf(giou[i], 0.0f);   // losses.py:123
}
__global__ void v5_match_tobj_kernel(const MatchParams p, const float* __restrict__ giou, float* __restrict__ tobj) {
    v5_match_tobj_body(p, giou, tobj, (int)blockIdx.x);
}

__global__ void v5_match_bwd_kernel(const MatchParams p, const float* __restrict__ ggiou, float* __restrict__ gpi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= match_rows(p)) return;
    const long long cell = match_cell(p, i);
    const float* ps = p.pi + cell * p.F;
    const float aw = p.anch[(size_t)i * 2], ah = p.anch[(size_t)i * 2 + 1];
    float s[4];
    for (int k = 0; k < 4; ++k) s[k] = sigmoidf_acc(ps[k]);
    float4 pb = make_float4(s[0] * 2.0f - 0.5f, s[1] * 2.0f - 0.5f, (s[2] * 2.0f) * (s[2] * 2.0f) * aw,
                            (s[3] * 2.0f) * (s[3] * 2.0f) * ah);
    const float4 tb = *reinterpret_cast<const float4*>(p.tbox + (size_t)i * 4);
    float g[4];
    iou_v5_backward(pb, tb, false, B200DET_GIOU, ggiou[i], g);
    float* gp = gpi + cell * p.F;
    atomicAdd(gp + 0, g[0] * 2.0f * s[0] * (1.0f - s[0]));
    atomicAdd(gp + 1, g[1] * 2.0f * s[1] * (1.0f - s[1]));
    atomicAdd(gp + 2, g[2] * 8.0f * s[2] * s[2] * (1.0f - s[2]) * aw);
    atomicAdd(gp + 3, g[3] * 8.0f * s[3] * s[3] * (1.0f - s[3]) * ah);
}

// ================================================================================================
// T1 — build_targets (YOLOv2..v4).
// ================================================================================================
struct BtParams {
    const float* pred_boxes;   // [B,A,G,G,4]   (cell stride box_ld floats; 4 = dense)
    const float* pred_cls;     // [B,A,G,G,C]   (cell stride cls_ld floats; C = dense)
    int box_ld, cls_ld;
    const float* target;       // [nt,6]
    const float* anchors;      // [A,2]
    int B, A, G, C, nt;
    float ignore_thres;
    int* winner;               // [B*A*G*G], -1
    int* tinfo;                // [nt][4]: b, best_n, gj, gi   (wrapped) or -1 when unusable
    float *iou_scores, *class_mask, *tx, *ty, *tw, *th, *tcls;
    uint8_t *obj, *noobj;
    int32_t* status;           // bit0 index guard, bit1 label guard, bit2 negative index out of range
};

__device__ __forceinline__ bool wrap_index(int& i, int dim) {   // torch advanced indexing wraps negatives
    if (i < 0) i += dim;
    return i >= 0;
}

// all output fills of build_targets (accuracy.py:316-324): winner = -1, obj = 0, noobj = 1, the fp32 maps = 0
__global__ void __launch_bounds__(256) bt_fill_kernel(const BtParams p, const size_t cells) {
    const size_t stride = (size_t)gridDim.x * 256, i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (size_t i = i0; i < cells; i += stride) {
        p.winner[i] = -1;
        p.obj[i] = 0; p.noobj[i] = 1;
        p.class_mask[i] = 0.0f; p.iou_scores[i] = 0.0f;
        if (p.tx) { p.tx[i] = 0.0f; p.ty[i] = 0.0f; p.tw[i] = 0.0f; p.th[i] = 0.0f; }
    }
    if (p.tcls) {
        const size_t n = cells * (size_t)p.C;
        if ((((uintptr_t)p.tcls) & 15) == 0) {
            float4* t4 = reinterpret_cast<float4*>(p.tcls);
            for (size_t i = i0; i < n / 4; i += stride) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (size_t i = (n / 4) * 4 + i0; i < n; i += stride) p.tcls[i] = 0.0f;
        } else {
            for (size_t i = i0; i < n; i += stride) p.tcls[i] = 0.0f;
        }
    }
    if (i0 == 0) *p.status = 0;
}

// pass 1: per target best anchor, guards, winner election, noobj ignore-threshold clearing
__global__ void build_targets_pass1(const BtParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.nt) return;
    const float* r = p.target + (size_t)t * 6;
    const float G = (float)p.G;
    const float gx = __fmul_rn(r[2], G), gy = __fmul_rn(r[3], G), gw = __fmul_rn(r[4], G), gh = __fmul_rn(r[5], G);
    int b = (int)r[0], lab = (int)r[1];
    int gi = (int)gx, gj = (int)gy;                                  // .long(), accuracy.py:337
    float best = 0.f;
    int best_n = 0;
    for (int a = 0; a < p.A; ++a) {                                  // bbox_wh_iou, accuracy.py:297-303
        const float aw = p.anchors[a * 2], ah = p.anchors[a * 2 + 1];
        const float inter = __fmul_rn(fminf(aw, gw), fminf(ah, gh));
        const float uni = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, ah), 1e-16f), __fmul_rn(gw, gh)), inter);
        const float v = __fdiv_rn(inter, uni);
        if (a == 0) { best = v; best_n = 0; }
        else if (!(v <= best) && (best == best)) { best = v; best_n = a; }
    }
    int st = 0;
    if (b >= p.B || gj >= p.G || gi >= p.G) st |= 1;                 // accuracy.py:340-344 (best_n < A always)
    if (lab >= p.C) st |= 2;                                         // accuracy.py:365
    int bw = b, gjw = gj, giw = gi, labw = lab;
    const bool okb = wrap_index(bw, p.B), okj = wrap_index(gjw, p.G), oki = wrap_index(giw, p.G);
    const bool okl = wrap_index(labw, p.C);
    if (!(okb && okj && oki && okl)) st |= 4;
    if (st) atomicOr(p.status, st);
    int* ti = p.tinfo + (size_t)t * 4;
    const bool usable = !(st & 1) && okb && okj && oki;
    ti[0] = usable ? bw : -1; ti[1] = best_n; ti[2] = gjw; ti[3] = giw;
    if (usable) {
        const long long cell = (((long long)bw * p.A + best_n) * p.G + gjw) * p.G + giw;
        atomicMax(&p.winner[cell], t);
        // accuracy.py:349-358 — per-target guard only (b, gj, gi upper bounds)
        for (int a = 0; a < p.A; ++a) {
            const float aw = p.anchors[a * 2], ah = p.anchors[a * 2 + 1];
            const float inter = __fmul_rn(fminf(aw, gw), fminf(ah, gh));
            const float uni = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, ah), 1e-16f), __fmul_rn(gw, gh)), inter);
            if (__fdiv_rn(inter, uni) > p.ignore_thres)
                p.noobj[(((long long)bw * p.A + a) * p.G + gjw) * p.G + giw] = 0;
        }
    }
}

// pass 2: scatter (all-or-nothing on the global guards)
__global__ void build_targets_pass2(const BtParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.nt) return;
    const int st = *p.status;
    if (st & (1 | 4)) return;                                         // accuracy.py:344: skip every scatter
    const int* ti = p.tinfo + (size_t)t * 4;
    const int b = ti[0], n = ti[1], gj = ti[2], gi = ti[3];
    const long long cell = (((long long)b * p.A + n) * p.G + gj) * p.G + gi;
    p.obj[cell] = 1;                                                  // accuracy.py:345-346
    p.noobj[cell] = 0;
    if (st & 2) return;                                               // accuracy.py:367
    const float* r = p.target + (size_t)t * 6;
    const float G = (float)p.G;
    const float gx = __fmul_rn(r[2], G), gy = __fmul_rn(r[3], G), gw = __fmul_rn(r[4], G), gh = __fmul_rn(r[5], G);
    int lab = (int)r[1];
    wrap_index(lab, p.C);
    if (p.tcls) p.tcls[cell * p.C + lab] = 1.0f;                      // accuracy.py:374 (multi-hot on duplicates)
    if (p.winner[cell] != t) return;                                  // highest target row wins
    if (p.tx) {                                                       // (the statistics path does not need the regression targets)
        p.tx[cell] = __fsub_rn(gx, floorf(gx));                       // accuracy.py:368-372
        p.ty[cell] = __fsub_rn(gy, floorf(gy));
        p.tw[cell] = logf(__fadd_rn(__fdiv_rn(gw, p.anchors[n * 2]), 1e-16f));
        p.th[cell] = logf(__fadd_rn(__fdiv_rn(gh, p.anchors[n * 2 + 1]), 1e-16f));
    }
    const float* pc = p.pred_cls + cell * p.cls_ld;                   // accuracy.py:376
    float best = pc[0];
    int besti = 0;
    for (int c = 1; c < p.C; ++c) {
        const float v = pc[c];
        if (!(v <= best) && (best == best)) { best = v; besti = c; }
    }
    p.class_mask[cell] = besti == lab ? 1.0f : 0.0f;
    float4 pb;                                                        // accuracy.py:377
    if (p.box_ld == 4) {
        pb = *reinterpret_cast<const float4*>(p.pred_boxes + cell * 4);
    } else {
        const float* q = p.pred_boxes + cell * p.box_ld;
        pb = make_float4(q[0], q[1], q[2], q[3]);
    }
    p.iou_scores[cell] = iou_plus1_eps(cxcywh_to_corners(pb), cxcywh_to_corners(make_float4(gx, gy, gw, gh)));
}

// ================================================================================================
// T5 — SSD matching.
// ================================================================================================
__device__ __forceinline__ float4 center_to_points(const float4 c) {          // losses.py:172-185
    return make_float4(fmaxf(__fsub_rn(c.x, __fmul_rn(c.z, 0.5f)), 0.0f), fmaxf(__fsub_rn(c.y, __fmul_rn(c.w, 0.5f)), 0.0f),
                       fminf(__fadd_rn(c.x, __fmul_rn(c.z, 0.5f)), 1.0f), fminf(__fadd_rn(c.y, __fmul_rn(c.w, 0.5f)), 1.0f));
}

// per prior: best GT (first max over m) ; per GT: best prior via packed 64-bit atomicMax (iou bits, ~p)
__global__ void ssd_match_kernel(const float4* __restrict__ priors, int P, const float4* __restrict__ gt, int M,
                                 float thresh, unsigned long long* __restrict__ gt_best, int32_t* __restrict__ idx,
                                 uint8_t* __restrict__ matched, int* __restrict__ forced) {
    extern __shared__ float4 s_gt[];
    for (int m = threadIdx.x; m < M; m += blockDim.x) s_gt[m] = center_to_points(gt[m]);
    __syncthreads();
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pidx < P;
    const float4 d = center_to_points(priors[live ? pidx : 0]);
    float best = 0.f;
    int bi = 0;
    for (int m = 0; m < M; ++m) {
        const float v = iou_plain(d, s_gt[m]);                                 // losses.py:209
        if (m == 0) { best = v; bi = 0; }
        else if (!(v <= best) && (best == best)) { best = v; bi = m; }        // losses.py:214 (max over GTs)
        // losses.py:211 (max over priors, first prior on ties) = max of (iou bits, ~prior index): reduced over the warp
        // first (two `redux.sync`), one atomic per warp — 175 k atomics on 20 addresses were most of this kernel's time
        const bool ok = live && v == v;
        const unsigned vb = ok ? __float_as_uint(v) : 0u;
        const unsigned vmax = __reduce_max_sync(0xFFFFFFFFu, vb);
        const unsigned lo = __reduce_max_sync(0xFFFFFFFFu, (ok && vb == vmax) ? 0xFFFFFFFFu - (unsigned)pidx : 0u);
        const unsigned any_ok = __ballot_sync(0xFFFFFFFFu, ok);
        if ((threadIdx.x & 31) == 0 && any_ok != 0u)
            atomicMax(&gt_best[m], ((unsigned long long)vmax << 32) | lo);
    }
    if (!live) return;
    idx[pidx] = bi;
    matched[pidx] = best >= thresh ? 1 : 0;                                    // losses.py:215
    forced[pidx] = -1;
}
__global__ void ssd_force_kernel(const unsigned long long* __restrict__ gt_best, int M, int* __restrict__ forced) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned pidx = 0xFFFFFFFFu - (unsigned)(gt_best[m] & 0xFFFFFFFFull);
    atomicMax(&forced[pidx], m);                                               // losses.py:216-217: last GT wins
}
__global__ void ssd_apply_kernel(const int* __restrict__ forced, int P, int32_t* __restrict__ idx, uint8_t* __restrict__ matched) {
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= P) return;
    const int f = forced[pidx];
    if (f >= 0) { idx[pidx] = f; matched[pidx] = 1; }
}

// ================================================================================================
// T6 — RetinaNet assignment.  Per image: ordered list of its target rows, then one thread per anchor.
// ================================================================================================
// Per image: its target rows in their original order, already in the form the assignment reads — pixel cx,cy,w,h
// (losses.py:425), corners ('xywh2xyxy', :373), +1-area, label — so that the anchor CTAs fetch them with one load each.
struct RetinaStaged {
    float4* xywh;     // [B][nt]
    float4* corners;  // [B][nt]
    float2* area_lab; // [B][nt]  (area, label bits)
    int* cnt;         // [B]
};
__global__ void __launch_bounds__(1024) group_targets_kernel(const float* __restrict__ targets, int nt, float img_size,
                                                             const RetinaStaged o) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    int base = 0;
    for (int t0 = 0; t0 < nt; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        const int f = (t < nt && targets[(size_t)t * 6] == (float)b) ? 1 : 0;    // losses.py:425 (targets[:,0]==bid)
        int total;
        const int ex = block_exclusive_scan(f, s_scan, &total);
        if (f) {
            const float* r = targets + (size_t)t * 6;
            const float4 x = make_float4(__fmul_rn(r[2], img_size), __fmul_rn(r[3], img_size), __fmul_rn(r[4], img_size),
                                         __fmul_rn(r[5], img_size));
            const float4 c = cxcywh_to_corners(x);
            const float ta = __fmul_rn(__fadd_rn(__fsub_rn(c.z, c.x), 1.0f), __fadd_rn(__fsub_rn(c.w, c.y), 1.0f));
            const size_t at = (size_t)b * nt + base + ex;
            o.xywh[at] = x; o.corners[at] = c;
            o.area_lab[at] = make_float2(ta, __int_as_float((int)r[1]));
        }
        base += total;
    }
    if (threadIdx.x == 0) o.cnt[b] = base;
}

// A warp's 32 consecutive anchors sit in three or four neighbouring cells; a target whose box cannot touch the hull of those
// anchors has IoU exactly 0 with every one of them and (after the image's first target, which seeds the running maximum
// even at 0) can never win `v > best`.  Each warp therefore tests 32 staged targets at a time against its hull (one lane
// per target, one ballot) and walks only the survivors, in order — about one target in twenty on the 600-pixel
// configuration.  The hull test repeats the pair test's own rounded operations on bounds (fsub/fadd are monotone), so it
// never drops a pair the full loop would score above 0; any non-finite or non-positive-area box in the CTA or the image
// so far switches the filter off.
__device__ __forceinline__ bool finite4(const float4 v) {
    return fabsf(v.x) < INFINITY && fabsf(v.y) < INFINITY && fabsf(v.z) < INFINITY && fabsf(v.w) < INFINITY;
}

__global__ void __launch_bounds__(256) retina_assign_kernel(const float4* __restrict__ anchors, int A,
                                                            const RetinaStaged tg, int nt,
                                                            float4* __restrict__ loc, int32_t* __restrict__ cls) {
    __shared__ float4 s_c[256];     // target corners
    __shared__ float4 s_x[256];     // target cx,cy,w,h (pixels)
    __shared__ float s_a[256];      // target area (+1)
    __shared__ int s_l[256];        // label
    const int b = blockIdx.y;
    const int ai = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // the first chunk of staged targets is fetched before the image's count is known (slots beyond it are never used)
    const size_t trow = (size_t)b * nt;
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f), px = pc;
    float2 pal = make_float2(0.f, 0.f);
    if ((int)threadIdx.x < nt) { pc = tg.corners[trow + threadIdx.x]; px = tg.xywh[trow + threadIdx.x]; pal = tg.area_lab[trow + threadIdx.x]; }
    const int M = tg.cnt[b];
    float4 an = make_float4(0.f, 0.f, 1.f, 1.f);
    if (ai < A) an = anchors[ai];
    const float4 ac = cxcywh_to_corners(an);                                     // losses.py:373 ('xywh2xyxy')
    const float aa = __fmul_rn(__fadd_rn(__fsub_rn(ac.z, ac.x), 1.0f), __fadd_rn(__fsub_rn(ac.w, ac.y), 1.0f));
    // hull of the warp's anchors, and whether every anchor of the CTA is an ordinary box
    float hx0 = INFINITY, hy0 = INFINITY, hx1 = -INFINITY, hy1 = -INFINITY;
    bool odd = false;
    if (ai < A) {
        hx0 = ac.x; hy0 = ac.y; hx1 = ac.z; hy1 = ac.w;
        odd = !(finite4(ac) && aa > 0.0f && aa < INFINITY);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hx0 = fminf(hx0, __shfl_xor_sync(0xFFFFFFFFu, hx0, o)); hy0 = fminf(hy0, __shfl_xor_sync(0xFFFFFFFFu, hy0, o));
        hx1 = fmaxf(hx1, __shfl_xor_sync(0xFFFFFFFFu, hx1, o)); hy1 = fmaxf(hy1, __shfl_xor_sync(0xFFFFFFFFu, hy1, o));
    }
    bool unfiltered = __syncthreads_or(odd) != 0;
    float best = 0.f;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    int bl = 0;
    bool have = false;
    for (int m0 = 0; m0 < M; m0 += 256) {
        const int mm = min(256, M - m0);
        __syncthreads();
        bool t_odd = false;
        if (m0 > 0 && (int)threadIdx.x < mm) {
            pc = tg.corners[trow + m0 + threadIdx.x]; px = tg.xywh[trow + m0 + threadIdx.x]; pal = tg.area_lab[trow + m0 + threadIdx.x];
        }
        if ((int)threadIdx.x < mm) {
            s_x[threadIdx.x] = px; s_c[threadIdx.x] = pc;
            s_a[threadIdx.x] = pal.x;
            s_l[threadIdx.x] = __float_as_int(pal.y);
            t_odd = !(finite4(pc) && pal.x > 0.0f && pal.x < INFINITY);
        }
        unfiltered = (__syncthreads_or(t_odd) != 0) || unfiltered;                // also publishes the staged targets
        for (int t0 = 0; t0 < mm; t0 += 32) {
            bool keep = false;
            if (t0 + lane < mm) {
                const float4 c = s_c[t0 + lane];
                // the pair test on the hull: extents that cannot be positive for any anchor of this warp
                const bool apart = __fadd_rn(__fsub_rn(c.z, hx0), 1.0f) <= 0.0f || __fadd_rn(__fsub_rn(hx1, c.x), 1.0f) <= 0.0f ||
                                   __fadd_rn(__fsub_rn(c.w, hy0), 1.0f) <= 0.0f || __fadd_rn(__fsub_rn(hy1, c.y), 1.0f) <= 0.0f;
                keep = unfiltered || !apart || (m0 + t0 + lane == 0);            // the first target seeds best/bx even at IoU 0
            }
            for (unsigned todo = __ballot_sync(0xFFFFFFFFu, keep); todo; todo &= todo - 1u) {
                const int m = t0 + __ffs((int)todo) - 1;
                const float4 c = s_c[m];
                const float iw = fmaxf(__fadd_rn(__fsub_rn(fminf(ac.z, c.z), fmaxf(ac.x, c.x)), 1.0f), 0.0f);   // losses.py:393-397
                const float ih = fmaxf(__fadd_rn(__fsub_rn(fminf(ac.w, c.w), fmaxf(ac.y, c.y)), 1.0f), 0.0f);
                const float inter = __fmul_rn(iw, ih);
                const float uni = __fsub_rn(__fadd_rn(aa, s_a[m]), inter);
                // losses.py:401.  Pairs that do not touch: 0 / positive is 0 without the IEEE division
                const float v = (inter == 0.0f && uni > 0.0f) ? 0.0f : __fdiv_rn(inter, uni);
                if (!have || (!(v <= best) && (best == best))) { best = v; bx = s_x[m]; bl = s_l[m]; have = true; }   // :431
            }
        }
    }
    if (ai >= A) return;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0;
    if (have) {
        o.x = __fdiv_rn(__fsub_rn(bx.x, an.x), an.z);                                                        // losses.py:434
        o.y = __fdiv_rn(__fsub_rn(bx.y, an.y), an.w);
        o.z = logf(__fdiv_rn(bx.z, an.z));                                                                    // losses.py:435
        o.w = logf(__fdiv_rn(bx.w, an.w));
        c = 1 + bl;                                                                                           // losses.py:437
        if (best < 0.5f) c = 0;                                                                               // losses.py:439
        if (best > 0.4f && best < 0.5f) c = -1;                                                               // losses.py:440-441
    }
    loc[(size_t)b * A + ai] = o;
    cls[(size_t)b * A + ai] = c;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int build_targets_v5_launch(const float* targets, int nt, const float* anchors_host, int na, int nx, int ny,
                            int32_t* ob, int32_t* oa, int32_t* ogj, int32_t* ogi, int32_t* ocls, float* otbox,
                            float* oanch, int32_t* ocount, cudaStream_t st) {
    Tv5Params p;
    memset(&p, 0, sizeof(p));
    p.targets = targets; p.nt = nt; p.na = na; p.nx = nx; p.ny = ny;
    for (int a = 0; a < na; ++a) { p.anc[a][0] = anchors_host[a * 2]; p.anc[a][1] = anchors_host[a * 2 + 1]; }
    p.ob = ob; p.oa = oa; p.ogj = ogj; p.ogi = ogi; p.ocls = ocls; p.otbox = otbox; p.oanch = oanch; p.ocount = ocount;
    build_targets_v5_kernel<<<1, 1024, 0, st>>>(p);
    B2_LAUNCH_CHECK("build_targets_v5_kernel");
    return 0;
}

// all levels in one launch; per-level arrays of nl entries (host): grids, anchors [nl][na][2], output pointers
int build_targets_v5_multi_launch(const float* targets, int nt, int nl, const float* anchors_host, int na, const int32_t* nx,
                                  const int32_t* ny, int32_t* const* ob, int32_t* const* oa, int32_t* const* ogj,
                                  int32_t* const* ogi, int32_t* const* ocls, float* const* otbox, float* const* oanch,
                                  int32_t* ocount, cudaStream_t st) {
    Tv5Multi m;
    memset(&m, 0, sizeof(m));
    for (int l = 0; l < nl; ++l) {
        Tv5Params& p = m.lvl[l];
        p.targets = targets; p.nt = nt; p.na = na; p.nx = nx[l]; p.ny = ny[l];
        for (int a = 0; a < na; ++a) {
            p.anc[a][0] = anchors_host[((size_t)l * na + a) * 2];
            p.anc[a][1] = anchors_host[((size_t)l * na + a) * 2 + 1];
        }
        p.ob = ob[l]; p.oa = oa[l]; p.ogj = ogj[l]; p.ogi = ogi[l]; p.ocls = ocls[l]; p.otbox = otbox[l]; p.oanch = oanch[l];
        p.ocount = ocount + l;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kTv5Cluster, nl, 1);
    cfg.blockDim = dim3(1024, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kTv5Cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2_CUDA(cudaLaunchKernelEx(&cfg, build_targets_v5_multi_kernel, m));
    return 0;
}

int v5_match_fwd_launch(const float* pi, int B, int na, int ny, int nx, int F, const int32_t* b, const int32_t* a,
                        const int32_t* gj, const int32_t* gi, const float* tbox, const float* anch, int m, float* giou,
                        float* tobj, cudaStream_t st) {
    if (m == 0) return 0;
    const int32_t* m_dev = nullptr;
    MatchParams p{pi, B, na, ny, nx, F, b, a, gj, gi, tbox, anch, m, m_dev};
    const int blocks = ceil_div(m, 256);
    v5_match_fwd_kernel<<<blocks, 256, 0, st>>>(p, giou, reinterpret_cast<int*>(tobj));
    B2_LAUNCH_CHECK("v5_match_fwd_kernel");
    v5_match_tobj_kernel<<<blocks, 256, 0, st>>>(p, giou, tobj);
    B2_LAUNCH_CHECK("v5_match_tobj_kernel");
    return 0;
}

int v5_match_bwd_launch(const float* pi, int B, int na, int ny, int nx, int F, const int32_t* b, const int32_t* a,
                        const int32_t* gj, const int32_t* gi, const float* tbox, const float* anch, int m,
                        const float* ggiou, float* gpi, cudaStream_t st) {
    if (m == 0) return 0;
    const int32_t* m_dev = nullptr;
    MatchParams p{pi, B, na, ny, nx, F, b, a, gj, gi, tbox, anch, m, m_dev};
    v5_match_bwd_kernel<<<ceil_div(m, 256), 256, 0, st>>>(p, ggiou, gpi);
    B2_LAUNCH_CHECK("v5_match_bwd_kernel");
    return 0;
}

// ================================================================================================
// T4' — the three per-level loss terms of MultiScaleRegionLoss_v5.forward (losses.py:105-137) in one op:
//   lbox = mean(1 - giou)                              (:119)
//   lobj = mean(FL(pi[..., 4], tobj))                   (:137, FocalLoss(BCEWithLogits) :37-64, gamma 1.5, alpha 0.25)
//   lcls = mean(FL(ps[:, 5:], onehot(tcls) cp / cn))    (:131-133)
// Forward leaves sums[3] (fp64): sum(1 - giou), sum FL_obj over all cells, sum FL_cls over the m x C matched logits;
// backward takes the three upstream scalars already divided by the element counts.
// ================================================================================================
// The focal terms are evaluated 1.6 M (objectness) + m x C (classes) times per level and direction, and the kernels that do it
// were bound by this arithmetic, not by memory (two expf, log1pf, a division and the generic powf: ~400 instructions per logit).
// One expf(-|x|) now serves the log-sigmoid and the sigmoid, and (1 - p_t)^gamma is q * sqrt(q) for the criterion's gamma = 1.5
// (any other gamma takes powf).  Same formulas as losses.py:37-64 up to the last-ulp rounding of the elementary functions
// (parity tolerance: 1e-5 relative on the loss terms, tests/test_gpu_targets.py).
__device__ __forceinline__ void bce_and_sigmoid(const float x, const float t, float& bce, float& pr) {
    const float e = expf(-fabsf(x));                                               // (0, 1]
    const float ls = fminf(x, 0.0f) - log1pf(e);                                   // log_sigmoid(x)
    bce = (1.0f - t) * x - ls;                                                     // BCEWithLogits, pos_weight = 1
    const float r = 1.0f / (1.0f + e);
    pr = x >= 0.0f ? r : e * r;                                                    // sigmoid(x)
}
__device__ __forceinline__ float focal_bce(const float x, const float t, const float gamma, const float alpha) {
    float bce, pr;
    bce_and_sigmoid(x, t, bce, pr);
    const float p_t = t * pr + (1.0f - t) * (1.0f - pr);                           // losses.py:54
    const float af = t * alpha + (1.0f - t) * (1.0f - alpha);                      // :55
    const float q = 1.0f - p_t;
    const float mf = gamma == 1.5f ? q * sqrtf(q) : powf(q, gamma);
    return bce * (af * mf);                                                        // :56-57
}
__device__ __forceinline__ float focal_bce_grad(const float x, const float t, const float gamma, const float alpha) {
    float bce, pr;
    bce_and_sigmoid(x, t, bce, pr);
    const float p_t = t * pr + (1.0f - t) * (1.0f - pr);
    const float af = t * alpha + (1.0f - t) * (1.0f - alpha);
    const float q = 1.0f - p_t;
    float mf, pw1;                                                                 // q^gamma, q^(gamma - 1)
    if (gamma == 1.5f) { pw1 = sqrtf(q); mf = q * pw1; }
    else { mf = powf(q, gamma); pw1 = powf(q, gamma - 1.0f); }
    const float dpt = (2.0f * t - 1.0f) * pr * (1.0f - pr);
    const float dmf = q > 0.0f ? -gamma * pw1 * dpt : 0.0f;
    return af * ((pr - t) * mf + bce * dmf);
}

__device__ __forceinline__ void block_add_double(double v, double* target) {
    __shared__ double s_part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        atomicAdd(target, t);
    }
    __syncthreads();
}

// One WARP per matched row: the 5+C fields of a row are contiguous, so the lanes read (and, backward, update) them
// coalesced; one thread per row walks 340-byte rows with a 32-way scattered access pattern (measured 182 / 214 us for the
// three levels of config 4 forward / backward).
__device__ __forceinline__ void v5_loss_rows_fwd_body(const MatchParams& p, const int32_t* __restrict__ tcls,
                                                      const float* __restrict__ giou, float cp, float cn, float gamma,
                                                      float alpha, int with_cls, double* __restrict__ sums, const int bid) {
    const int lane = threadIdx.x & 31;
    const int i = bid * 8 + (threadIdx.x >> 5);
    double box = 0.0, cls = 0.0;
    if (i < match_rows(p)) {
        if (lane == 0) box = (double)(1.0f - giou[i]);
        if (with_cls) {
            const float* ps = p.pi + match_cell(p, i) * p.F + 5;
            const int lab = tcls[i];
            for (int c = lane; c < p.F - 5; c += 32) cls += (double)focal_bce(ps[c], c == lab ? cp : cn, gamma, alpha);
        }
    }
    block_add_double(box, sums + 0);
    block_add_double(cls, sums + 2);
}
__global__ void __launch_bounds__(256) v5_loss_rows_fwd_kernel(const MatchParams p, const int32_t* __restrict__ tcls,
                                                               const float* __restrict__ giou, float cp, float cn,
                                                               float gamma, float alpha, int with_cls, double* __restrict__ sums) {
    v5_loss_rows_fwd_body(p, tcls, giou, cp, cn, gamma, alpha, with_cls, sums, (int)blockIdx.x);
}

// obj_grad (may be null): d FL / d logit of every cell, for the backward pass.  The backward would otherwise re-read column 4
// of pi with a 340-byte stride WHILE it streams the gradient tensor out, and those reads are what it then waits for (level 0 of
// the headline: 418 MB written; fill alone 62-69 us, with the strided reads 96-109 us — tools/ubench/gradfill.cu); here the
// logit is in a register anyway, and the backward reads 4 contiguous bytes per cell instead.
__device__ __forceinline__ void v5_loss_obj_fwd_body(const float* __restrict__ pi, int F, long long cells,
                                                     const float* __restrict__ tobj, float gamma, float alpha,
                                                     double* __restrict__ sums, float* __restrict__ obj_grad, const int bid,
                                                     const int nblk) {
    double acc = 0.0;
    const long long stride = (long long)nblk * 256;
    // four cells of the thread in flight at once: the 340-byte-strided logit reads are one L2 / DRAM round trip each
    for (long long c = (long long)bid * 256 + threadIdx.x; c < cells; c += 4 * stride) {
        float x[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long ck = c + k * stride;
            x[k] = ck < cells ? pi[ck * F + 4] : 0.0f;
            t[k] = ck < cells ? tobj[ck] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long ck = c + k * stride;
            if (ck < cells) {
                acc += (double)focal_bce(x[k], t[k], gamma, alpha);
                if (obj_grad) obj_grad[ck] = focal_bce_grad(x[k], t[k], gamma, alpha);
            }
        }
    }
    block_add_double(acc, sums + 1);
}
__global__ void __launch_bounds__(256) v5_loss_obj_fwd_kernel(const float* __restrict__ pi, int F, long long cells,
                                                              const float* __restrict__ tobj, float gamma, float alpha,
                                                              double* __restrict__ sums, float* __restrict__ obj_grad) {
    v5_loss_obj_fwd_body(pi, F, cells, tobj, gamma, alpha, sums, obj_grad, (int)blockIdx.x, (int)gridDim.x);
}

__global__ void __launch_bounds__(256) v5_loss_obj_bwd_kernel(const float* __restrict__ pi, int F, long long cells,
                                                              const float* __restrict__ tobj, float gamma, float alpha,
                                                              const float* __restrict__ g3, float inv_cells,
                                                              float* __restrict__ gpi) {
    const float g_obj = g3[1] * inv_cells;
    for (long long c = (long long)blockIdx.x * 256 + threadIdx.x; c < cells; c += (long long)gridDim.x * 256)
        gpi[c * F + 4] = g_obj * focal_bce_grad(pi[c * F + 4], tobj[c], gamma, alpha);   // the only writer of column 4
}

// The same objectness gradient, but the kernel writes the WHOLE gradient tensor of its cells — zeros everywhere except
// column 4 — so the caller needs no zero-fill of the 548 MB gradient before it (a separate memset pass plus the
// read-modify-write of one 4-byte field per 340-byte row afterwards).  A CTA takes 256 cells per step, 64 at a time: the
// 64 x F floats (contiguous, 16-byte aligned) are zero-filled with plain float4 stores and, after a CTA barrier (which orders
// the two stores to the same words), the 64 owners store their cell's column 4 into lines that are still in L2 — 26 MB are
// in flight over the whole GPU.  The per-cell factor d FL / d logit comes from the forward pass (obj_grad, 4 contiguous bytes
// per cell) and the NEXT step's value is loaded before the fill.  History (level 0 = 418 MB; plain memset of it: 58-60 us):
// an index walk composing every float4 inside the store loop 105 us; fill + patch with the logit re-read from pi 102-104 us —
// tools/ubench/gradfill.cu shows that it is those 340-byte-strided reads next to the write stream that cost the 35-40 us, not
// the composition or the patch (DRAM reads 163 MB for 39 MB of sectors asked for).
// The matched-row kernel then adds its terms on top, as before.
__device__ __forceinline__ void v5_loss_obj_bwd_full_body(const float* __restrict__ pi, int F, long long cells,
                                                          const float* __restrict__ tobj, const float* __restrict__ obj_grad,
                                                          float gamma, float alpha, const float* __restrict__ g3,
                                                          float inv_cells, float* __restrict__ gpi, const int bid,
                                                          const int nblk) {
    const float g_obj = g3[1] * inv_cells;
    const int tid = threadIdx.x;
    const long long stride = (long long)nblk * 256;
    long long c0 = (long long)bid * 256;
    // d FL / d logit of this thread's cell: from the forward pass (obj_grad), else recomputed from the strided logit
    auto cell_grad = [&](const long long c) -> float {
        if (c >= cells) return 0.0f;
        return obj_grad ? obj_grad[c] : focal_bce_grad(pi[c * F + 4], tobj[c], gamma, alpha);
    };
    float h = cell_grad(c0 + tid);
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (; c0 < cells; c0 += stride) {
        const float g = g_obj * h;
        h = cell_grad(c0 + stride + tid);                                       // next step's cell of this thread
        const int ncell = (int)min((long long)256, cells - c0);
        float* out = gpi + c0 * F;
        for (int cb = 0; cb < ncell; cb += 64) {
            const int nsub = min(64, ncell - cb);
            float* o = out + cb * F;                                            // 64 * F * 4 bytes per sub-block: 16-byte aligned
            const int n = nsub * F, n4 = n >> 2;
            for (int v = tid; v < n4; v += 256) reinterpret_cast<float4*>(o)[v] = z;
            for (int idx = (n4 << 2) + tid; idx < n; idx += 256) o[idx] = 0.0f; // tail of the tensor's last sub-block
            __syncthreads();
            if (tid >= cb && tid < cb + nsub) out[tid * F + 4] = g;
        }
    }
}
__global__ void __launch_bounds__(256) v5_loss_obj_bwd_full_kernel(const float* __restrict__ pi, int F, long long cells,
                                                                   const float* __restrict__ tobj,
                                                                   const float* __restrict__ obj_grad, float gamma, float alpha,
                                                                   const float* __restrict__ g3, float inv_cells,
                                                                   float* __restrict__ gpi) {
    v5_loss_obj_bwd_full_body(pi, F, cells, tobj, obj_grad, gamma, alpha, g3, inv_cells, gpi, (int)blockIdx.x, (int)gridDim.x);
}

__device__ __forceinline__ void v5_loss_rows_bwd_body(const MatchParams& p, const int32_t* __restrict__ tcls, float cp, float cn,
                                                      float gamma, float alpha, int with_cls, const float* __restrict__ g3,
                                                      float inv_nbox, float inv_ncls, float* __restrict__ gpi, const int bid) {
    const int lane = threadIdx.x & 31;
    const int i = bid * 8 + (threadIdx.x >> 5);
    if (i >= match_rows(p)) return;
    if (p.m_dev) {                                                  // the means' divisors from the device-side row count
        const long long mm = max(match_rows(p), 1);
        inv_nbox = (float)(1.0 / (double)mm);
        inv_ncls = (float)(1.0 / (double)max(mm * (p.F - 5), 1ll));
    }
    const float g_box = g3[0] * inv_nbox, g_cls = g3[2] * inv_ncls;
    const long long cell = match_cell(p, i);
    const float* ps = p.pi + cell * p.F;
    float* gp = gpi + cell * p.F;
    if (lane == 0) {
        const float aw = p.anch[(size_t)i * 2], ah = p.anch[(size_t)i * 2 + 1];
        float s[4];
        for (int k = 0; k < 4; ++k) s[k] = sigmoidf_acc(ps[k]);
        const float4 pb = make_float4(s[0] * 2.0f - 0.5f, s[1] * 2.0f - 0.5f, (s[2] * 2.0f) * (s[2] * 2.0f) * aw,
                                      (s[3] * 2.0f) * (s[3] * 2.0f) * ah);
        const float4 tb = *reinterpret_cast<const float4*>(p.tbox + (size_t)i * 4);
        float g[4];
        iou_v5_backward(pb, tb, false, B200DET_GIOU, -g_box, g);                    // d(1 - giou) = -d giou
        atomicAdd(gp + 0, g[0] * 2.0f * s[0] * (1.0f - s[0]));
        atomicAdd(gp + 1, g[1] * 2.0f * s[1] * (1.0f - s[1]));
        atomicAdd(gp + 2, g[2] * 8.0f * s[2] * s[2] * (1.0f - s[2]) * aw);
        atomicAdd(gp + 3, g[3] * 8.0f * s[3] * s[3] * (1.0f - s[3]) * ah);
    }
    if (with_cls) {
        const int lab = tcls[i];
        for (int c = lane; c < p.F - 5; c += 32)
            atomicAdd(gp + 5 + c, g_cls * focal_bce_grad(ps[5 + c], c == lab ? cp : cn, gamma, alpha));
    }
}
__global__ void __launch_bounds__(256) v5_loss_rows_bwd_kernel(const MatchParams p, const int32_t* __restrict__ tcls, float cp,
                                                               float cn, float gamma, float alpha, int with_cls,
                                                               const float* __restrict__ g3, float inv_nbox, float inv_ncls,
                                                               float* __restrict__ gpi) {
    v5_loss_rows_bwd_body(p, tcls, cp, cn, gamma, alpha, with_cls, g3, inv_nbox, inv_ncls, gpi, (int)blockIdx.x);
}

__global__ void v5_loss_means_kernel(double* __restrict__ sums, double n_box, double n_cells, double n_cls,
                                     const int32_t* __restrict__ m_dev, int cap, int C) {
    if (m_dev) {
        const long long mm = min((long long)cap, (long long)*m_dev);
        n_box = (double)max(mm, 1ll);
        n_cls = (double)max(mm * C, 1ll);
    }
    if (threadIdx.x < 3) sums[threadIdx.x] = sums[threadIdx.x] / (threadIdx.x == 0 ? n_box : threadIdx.x == 1 ? n_cells : n_cls);
}

int v5_loss_fwd_launch(const float* pi, int B, int na, int ny, int nx, int F, const int32_t* b, const int32_t* a,
                       const int32_t* gj, const int32_t* gi, const int32_t* tcls, const float* tbox, const float* anch, int m,
                       float cp, float cn, float gamma, float alpha, int with_cls, float* giou, float* tobj, double* sums,
                       const int32_t* m_dev, float* obj_grad, cudaStream_t st) {
    const long long cells = (long long)B * na * ny * nx;
    B2_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
    B2_CUDA(cudaMemsetAsync(tobj, 0, (size_t)cells * 4, st));                      // torch.zeros_like(pi[..., 0])  (:107)
    MatchParams p{pi, B, na, ny, nx, F, b, a, gj, gi, tbox, anch, m, m_dev};
    if (m > 0) {
        const int blocks = ceil_div(m, 256);
        v5_match_fwd_kernel<<<blocks, 256, 0, st>>>(p, giou, reinterpret_cast<int*>(tobj));
        B2_LAUNCH_CHECK("v5_match_fwd_kernel");
        v5_match_tobj_kernel<<<blocks, 256, 0, st>>>(p, giou, tobj);
        B2_LAUNCH_CHECK("v5_match_tobj_kernel");
        v5_loss_rows_fwd_kernel<<<ceil_div(m, 8), 256, 0, st>>>(p, tcls, giou, cp, cn, gamma, alpha, with_cls, sums);
        B2_LAUNCH_CHECK("v5_loss_rows_fwd_kernel");
    }
    const int grid = (int)((cells + 255) / 256 < 148 * 8 ? (cells + 255) / 256 : 148 * 8);
    v5_loss_obj_fwd_kernel<<<grid, 256, 0, st>>>(pi, F, cells, tobj, gamma, alpha, sums, obj_grad);
    B2_LAUNCH_CHECK("v5_loss_obj_fwd_kernel");
    // sums -> means in place: box / max(m, 1), obj / cells, cls / max(m * C, 1)   (reduction 'mean', losses.py:119-137)
    v5_loss_means_kernel<<<1, 32, 0, st>>>(sums, (double)(m > 0 ? m : 1), (double)cells,
                                           (double)((long long)m * (F - 5) > 0 ? (long long)m * (F - 5) : 1), m_dev, m, F - 5);
    B2_LAUNCH_CHECK("v5_loss_means_kernel");
    return 0;
}

// losses.py:139-152 on the device: the per-level means (fp64 -> fp32) are added in level order to fp32 zeros, scaled by the
// three gains, summed: out = (loss, Localization, Classification, Conf_obj).  Levels without matched rows contribute 0.
__global__ void v5_loss_combine_kernel(const double* __restrict__ means, int nl, float wbox, float wobj, float wcls,
                                       float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    float lbox = 0.f, lobj = 0.f, lcls = 0.f;
    for (int i = 0; i < nl; ++i) {
        lbox = __fadd_rn(lbox, (float)means[i * 3 + 0]);
        lobj = __fadd_rn(lobj, (float)means[i * 3 + 1]);
        lcls = __fadd_rn(lcls, (float)means[i * 3 + 2]);
    }
    lbox = __fmul_rn(lbox, wbox); lobj = __fmul_rn(lobj, wobj); lcls = __fmul_rn(lcls, wcls);
    out[0] = __fadd_rn(__fadd_rn(lbox, lobj), lcls);
    out[1] = lbox; out[2] = lcls; out[3] = lobj;
}
// backward of the combination: g3 = (d/d mean_box, d/d mean_obj, d/d mean_cls), the same for every level
__global__ void v5_loss_combine_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ g_box,
                                           const float* __restrict__ g_cls, const float* __restrict__ g_obj, float wbox,
                                           float wobj, float wcls, float* __restrict__ g3) {
    if (threadIdx.x != 0) return;
    const float gl = g_loss ? g_loss[0] : 0.f;
    g3[0] = __fmul_rn(__fadd_rn(gl, g_box ? g_box[0] : 0.f), wbox);
    g3[1] = __fmul_rn(__fadd_rn(gl, g_obj ? g_obj[0] : 0.f), wobj);
    g3[2] = __fmul_rn(__fadd_rn(gl, g_cls ? g_cls[0] : 0.f), wcls);
}
int v5_loss_combine_launch(const double* means, int nl, float wbox, float wobj, float wcls, float* out, cudaStream_t st) {
    v5_loss_combine_kernel<<<1, 32, 0, st>>>(means, nl, wbox, wobj, wcls, out);
    B2_LAUNCH_CHECK("v5_loss_combine_kernel");
    return 0;
}
int v5_loss_combine_bwd_launch(const float* g_loss, const float* g_box, const float* g_cls, const float* g_obj, float wbox,
                               float wobj, float wcls, float* g3, cudaStream_t st) {
    v5_loss_combine_bwd_kernel<<<1, 32, 0, st>>>(g_loss, g_box, g_cls, g_obj, wbox, wobj, wcls, g3);
    B2_LAUNCH_CHECK("v5_loss_combine_bwd_kernel");
    return 0;
}

// gpi must be zero-filled by the caller; g3 (device) = upstream gradients of the three means, inv_* = 1 / their counts
int v5_loss_bwd_launch(const float* pi, int B, int na, int ny, int nx, int F, const int32_t* b, const int32_t* a,
                       const int32_t* gj, const int32_t* gi, const int32_t* tcls, const float* tbox, const float* anch, int m,
                       float cp, float cn, float gamma, float alpha, int with_cls, const float* tobj, const float* g3,
                       float inv_nbox, float inv_cells, float inv_ncls, float* gpi, int fill, const int32_t* m_dev,
                       const float* obj_grad, cudaStream_t st) {
    const long long cells = (long long)B * na * ny * nx;
    const int grid = (int)((cells + 255) / 256 < 148 * 8 ? (cells + 255) / 256 : 148 * 8);
    if (fill && (((uintptr_t)gpi) & 15) == 0) {
        v5_loss_obj_bwd_full_kernel<<<grid, 256, 0, st>>>(pi, F, cells, tobj, obj_grad, gamma, alpha, g3, inv_cells, gpi);
        B2_LAUNCH_CHECK("v5_loss_obj_bwd_full_kernel");
    } else {
        if (fill) B2_CUDA(cudaMemsetAsync(gpi, 0, (size_t)cells * F * sizeof(float), st));
        v5_loss_obj_bwd_kernel<<<grid, 256, 0, st>>>(pi, F, cells, tobj, gamma, alpha, g3, inv_cells, gpi);
        B2_LAUNCH_CHECK("v5_loss_obj_bwd_kernel");
    }
    if (m > 0) {
        MatchParams p{pi, B, na, ny, nx, F, b, a, gj, gi, tbox, anch, m, m_dev};
        v5_loss_rows_bwd_kernel<<<ceil_div(m, 8), 256, 0, st>>>(p, tcls, cp, cn, gamma, alpha, with_cls, g3, inv_nbox, inv_ncls,
                                                                gpi);
        B2_LAUNCH_CHECK("v5_loss_rows_bwd_kernel");
    }
    return 0;
}

// ================================================================================================
// All levels of the criterion in one launch per stage (blockIdx.y = level): the per-level kernels above are 5-45 us each and a
// level-by-level step is ~30 launches whose ramps and tails — the stride-32 level has 300 CTAs of work — add up to a quarter of
// it.  Same bodies, same arithmetic, same results.
// ================================================================================================
__global__ void v5_match_fwd_multi_kernel(const __grid_constant__ V5Multi mp) {
    const V5Level& L = mp.lv[blockIdx.y];
    if ((int)blockIdx.x * 256 >= L.p.m) return;
    v5_match_fwd_body(L.p, L.giou, reinterpret_cast<int*>(L.tobj), (int)blockIdx.x);
}
__global__ void v5_match_tobj_multi_kernel(const __grid_constant__ V5Multi mp) {
    const V5Level& L = mp.lv[blockIdx.y];
    if ((int)blockIdx.x * 256 >= L.p.m) return;
    v5_match_tobj_body(L.p, L.giou, L.tobj, (int)blockIdx.x);
}
__global__ void __launch_bounds__(256) v5_loss_rows_fwd_multi_kernel(const __grid_constant__ V5Multi mp) {
    const V5Level& L = mp.lv[blockIdx.y];
    if ((int)blockIdx.x * 8 >= match_rows(L.p)) return;                         // CTA-uniform: nothing to add to the sums
    v5_loss_rows_fwd_body(L.p, L.tcls, L.giou, mp.cp, mp.cn, mp.gamma, mp.alpha, mp.with_cls, L.sums, (int)blockIdx.x);
}
__global__ void __launch_bounds__(256) v5_loss_obj_fwd_multi_kernel(const __grid_constant__ V5Multi mp) {
    const V5Level& L = mp.lv[blockIdx.y];
    if ((int)blockIdx.x >= L.obj_blocks) return;
    v5_loss_obj_fwd_body(L.p.pi, L.p.F, L.cells, L.tobj, mp.gamma, mp.alpha, L.sums, L.obj_grad, (int)blockIdx.x, L.obj_blocks);
}
__global__ void __launch_bounds__(256) v5_loss_obj_bwd_full_multi_kernel(const __grid_constant__ V5Multi mp) {
    const V5Level& L = mp.lv[blockIdx.y];
    if ((int)blockIdx.x >= L.obj_blocks) return;
    v5_loss_obj_bwd_full_body(L.p.pi, L.p.F, L.cells, L.tobj, L.obj_grad, mp.gamma, mp.alpha, mp.g3, L.inv_cells, L.gpi,
                              (int)blockIdx.x, L.obj_blocks);
}
__global__ void __launch_bounds__(256) v5_loss_rows_bwd_multi_kernel(const __grid_constant__ V5Multi mp) {
    const V5Level& L = mp.lv[blockIdx.y];
    if ((int)blockIdx.x * 8 >= match_rows(L.p)) return;
    v5_loss_rows_bwd_body(L.p, L.tcls, mp.cp, mp.cn, mp.gamma, mp.alpha, mp.with_cls, mp.g3, 0.0f, 0.0f, L.gpi, (int)blockIdx.x);
}
// sums -> means per level (reduction 'mean', losses.py:119-137: box / max(m, 1), obj / cells, cls / max(m * C, 1)), then the
// gain-weighted combination of losses.py:139-152 in level order — v5_loss_means_kernel x nl + v5_loss_combine_kernel in one thread
__global__ void v5_loss_means_combine_multi_kernel(const __grid_constant__ V5Multi mp, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lbox = 0.f, lobj = 0.f, lcls = 0.f;
    for (int i = 0; i < mp.nl; ++i) {
        const V5Level& L = mp.lv[i];
        const long long mm = match_rows(L.p);
        const double n_box = (double)max(mm, 1ll), n_cls = (double)max(mm * (L.p.F - 5), 1ll);
        const double m0 = L.sums[0] / n_box, m1 = L.sums[1] / (double)L.cells, m2 = L.sums[2] / n_cls;
        L.sums[0] = m0; L.sums[1] = m1; L.sums[2] = m2;
        lbox = __fadd_rn(lbox, (float)m0);
        lobj = __fadd_rn(lobj, (float)m1);
        lcls = __fadd_rn(lcls, (float)m2);
    }
    lbox = __fmul_rn(lbox, mp.wbox); lobj = __fmul_rn(lobj, mp.wobj); lcls = __fmul_rn(lcls, mp.wcls);
    out[0] = __fadd_rn(__fadd_rn(lbox, lobj), lcls);
    out[1] = lbox; out[2] = lcls; out[3] = lobj;
}

static int v5_obj_blocks(long long cells) { return (int)((cells + 255) / 256 < 148 * 8 ? (cells + 255) / 256 : 148 * 8); }

// tobj / obj_grad: the levels back to back ([sum of cells]); giou: [nl][cap]; sums: [nl][3]
int v5_loss_fwd_all_launch(V5Multi& mp, int cap, float* giou, float* tobj, float* obj_grad, double* sums, float* out4,
                           cudaStream_t st) {
    long long total = 0;
    int max_obj = 0;
    for (int i = 0; i < mp.nl; ++i) {
        V5Level& L = mp.lv[i];
        L.giou = giou + (size_t)i * cap;
        L.tobj = tobj + total;
        L.obj_grad = obj_grad ? obj_grad + total : nullptr;
        L.sums = sums + 3 * i;
        L.obj_blocks = v5_obj_blocks(L.cells);
        max_obj = max(max_obj, L.obj_blocks);
        total += L.cells;
    }
    B2_CUDA(cudaMemsetAsync(sums, 0, (size_t)mp.nl * 3 * sizeof(double), st));
    B2_CUDA(cudaMemsetAsync(tobj, 0, (size_t)total * 4, st));                      // torch.zeros_like(pi[..., 0])  (:107)
    const dim3 grows(ceil_div(cap, 256), mp.nl);
    v5_match_fwd_multi_kernel<<<grows, 256, 0, st>>>(mp);
    B2_LAUNCH_CHECK("v5_match_fwd_multi_kernel");
    v5_match_tobj_multi_kernel<<<grows, 256, 0, st>>>(mp);
    B2_LAUNCH_CHECK("v5_match_tobj_multi_kernel");
    v5_loss_rows_fwd_multi_kernel<<<dim3(ceil_div(cap, 8), mp.nl), 256, 0, st>>>(mp);
    B2_LAUNCH_CHECK("v5_loss_rows_fwd_multi_kernel");
    v5_loss_obj_fwd_multi_kernel<<<dim3(max_obj, mp.nl), 256, 0, st>>>(mp);
    B2_LAUNCH_CHECK("v5_loss_obj_fwd_multi_kernel");
    v5_loss_means_combine_multi_kernel<<<1, 32, 0, st>>>(mp, out4);
    B2_LAUNCH_CHECK("v5_loss_means_combine_multi_kernel");
    return 0;
}

// g3 (device, [3]) is scratch: the upstream gradients of the four outputs -> d/d mean_box, d/d mean_obj, d/d mean_cls
int v5_loss_bwd_all_launch(V5Multi& mp, int cap, const float* obj_grad, const float* g_loss, const float* g_box,
                           const float* g_cls, const float* g_obj, float* g3, cudaStream_t st) {
    long long total = 0;
    int max_obj = 0;
    for (int i = 0; i < mp.nl; ++i) {
        V5Level& L = mp.lv[i];
        L.obj_grad = const_cast<float*>(obj_grad) + total;
        L.tobj = nullptr;
        L.obj_blocks = v5_obj_blocks(L.cells);
        L.inv_cells = (float)(1.0 / (double)L.cells);
        max_obj = max(max_obj, L.obj_blocks);
        total += L.cells;
    }
    mp.g3 = g3;
    v5_loss_combine_bwd_kernel<<<1, 32, 0, st>>>(g_loss, g_box, g_cls, g_obj, mp.wbox, mp.wobj, mp.wcls, g3);
    B2_LAUNCH_CHECK("v5_loss_combine_bwd_kernel");
    v5_loss_obj_bwd_full_multi_kernel<<<dim3(max_obj, mp.nl), 256, 0, st>>>(mp);
    B2_LAUNCH_CHECK("v5_loss_obj_bwd_full_multi_kernel");
    v5_loss_rows_bwd_multi_kernel<<<dim3(ceil_div(cap, 8), mp.nl), 256, 0, st>>>(mp);
    B2_LAUNCH_CHECK("v5_loss_rows_bwd_multi_kernel");
    return 0;
}

size_t build_targets_ws_bytes(int B, int A, int G, int nt) {
    return align_up((size_t)B * A * G * G * 4, 256) + align_up((size_t)(nt > 0 ? nt : 1) * 16, 256);
}

int build_targets_launch_ex(const float* pred_boxes, int box_ld, const float* pred_cls, int cls_ld, const float* target,
                            const float* anchors, int B, int A, int G, int C, int nt, float ignore_thres, void* ws,
                            float* iou_scores, float* class_mask, uint8_t* obj, uint8_t* noobj, float* tx, float* ty,
                            float* tw, float* th, float* tcls, int32_t* status, cudaStream_t st);

int build_targets_launch(const float* pred_boxes, const float* pred_cls, const float* target, const float* anchors,
                         int B, int A, int G, int C, int nt, float ignore_thres, void* ws, float* iou_scores,
                         float* class_mask, uint8_t* obj, uint8_t* noobj, float* tx, float* ty, float* tw, float* th,
                         float* tcls, int32_t* status, cudaStream_t st) {
    return build_targets_launch_ex(pred_boxes, 4, pred_cls, C, target, anchors, B, A, G, C, nt, ignore_thres, ws, iou_scores,
                                   class_mask, obj, noobj, tx, ty, tw, th, tcls, status, st);
}

// box_ld / cls_ld: floats between consecutive cells of pred_boxes / pred_cls (so both can be columns of one decoded
// [cells, 5+C] row array); tx, ty, tw, th (all four or none) and tcls may be null when the caller does not need them.
int build_targets_launch_ex(const float* pred_boxes, int box_ld, const float* pred_cls, int cls_ld, const float* target,
                            const float* anchors, int B, int A, int G, int C, int nt, float ignore_thres, void* ws,
                            float* iou_scores, float* class_mask, uint8_t* obj, uint8_t* noobj, float* tx, float* ty,
                            float* tw, float* th, float* tcls, int32_t* status, cudaStream_t st) {
    const size_t cells = (size_t)B * A * G * G;
    BtParams p;
    p.pred_boxes = pred_boxes; p.pred_cls = pred_cls; p.target = target; p.anchors = anchors;
    p.box_ld = box_ld; p.cls_ld = cls_ld;
    p.B = B; p.A = A; p.G = G; p.C = C; p.nt = nt; p.ignore_thres = ignore_thres;
    p.winner = (int*)ws;
    p.tinfo = (int*)((char*)ws + align_up(cells * 4, 256));
    p.iou_scores = iou_scores; p.class_mask = class_mask; p.tx = tx; p.ty = ty; p.tw = tw; p.th = th; p.tcls = tcls;
    p.obj = obj; p.noobj = noobj; p.status = status;
    // accuracy.py:316-324 fills, in one launch (eleven cudaMemsetAsync calls cost ~60 us of launch latency at G = 13)
    {
        const size_t n4 = (tcls ? cells * (size_t)C : cells);
        const int grid = (int)((n4 / 4 + 255) / 256 < 148 * 8 ? (n4 / 4 + 255) / 256 + 1 : 148 * 8);
        bt_fill_kernel<<<grid, 256, 0, st>>>(p, cells);
        B2_LAUNCH_CHECK("bt_fill_kernel");
    }
    if (nt == 0) return 0;
    build_targets_pass1<<<ceil_div(nt, 128), 128, 0, st>>>(p);
    B2_LAUNCH_CHECK("build_targets_pass1");
    build_targets_pass2<<<ceil_div(nt, 128), 128, 0, st>>>(p);
    B2_LAUNCH_CHECK("build_targets_pass2");
    return 0;
}

size_t ssd_match_ws_bytes(int P, int M) {
    return align_up((size_t)(M > 0 ? M : 1) * 8, 256) + align_up((size_t)P * 4, 256);
}

int ssd_match_launch(const float* priors, int P, const float* gt, int M, float thresh, void* ws, int32_t* idx,
                     uint8_t* matched, cudaStream_t st) {
    unsigned long long* gt_best = (unsigned long long*)ws;
    int* forced = (int*)((char*)ws + align_up((size_t)(M > 0 ? M : 1) * 8, 256));
    B2_CUDA(cudaMemsetAsync(gt_best, 0, (size_t)(M > 0 ? M : 1) * 8, st));
    const size_t smem = (size_t)M * sizeof(float4);
    if (smem > 48 * 1024) {
        set_error("ssd_match: too many ground-truth boxes (%d > 3072)", M);
        return B200DET_ELIMIT;
    }
    ssd_match_kernel<<<ceil_div(P, 256), 256, smem, st>>>((const float4*)priors, P, (const float4*)gt, M, thresh, gt_best,
                                                          idx, matched, forced);
    B2_LAUNCH_CHECK("ssd_match_kernel");
    if (M > 0) {
        ssd_force_kernel<<<ceil_div(M, 128), 128, 0, st>>>(gt_best, M, forced);
        B2_LAUNCH_CHECK("ssd_force_kernel");
        ssd_apply_kernel<<<ceil_div(P, 256), 256, 0, st>>>(forced, P, idx, matched);
        B2_LAUNCH_CHECK("ssd_apply_kernel");
    }
    return 0;
}

static RetinaStaged retina_layout(void* ws, int B, int nt) {
    const size_t slots = (size_t)B * (nt > 0 ? nt : 1);
    char* p = (char*)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes, 256); return r; };
    RetinaStaged o;
    o.xywh = (float4*)take(slots * 16);
    o.corners = (float4*)take(slots * 16);
    o.area_lab = (float2*)take(slots * 8);
    o.cnt = (int*)take((size_t)B * 4);
    return o;
}
size_t retina_assign_ws_bytes(int B, int nt) {
    const size_t slots = (size_t)B * (nt > 0 ? nt : 1);
    return align_up(slots * 16, 256) * 2 + align_up(slots * 8, 256) + align_up((size_t)B * 4, 256);
}

int retina_assign_launch(const float* anchors, int A, const float* targets, int nt, int B, float img_size, void* ws,
                         float* loc, int32_t* cls, cudaStream_t st) {
    const RetinaStaged tg = retina_layout(ws, B, nt);
    group_targets_kernel<<<B, 1024, 0, st>>>(targets, nt, img_size, tg);
    B2_LAUNCH_CHECK("group_targets_kernel");
    dim3 grid(ceil_div(A, 256), B);
    retina_assign_kernel<<<grid, 256, 0, st>>>((const float4*)anchors, A, tg, nt, (float4*)loc, cls);
    B2_LAUNCH_CHECK("retina_assign_kernel");
    return 0;
}

}  // namespace b200det
