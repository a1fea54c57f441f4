// K3 — greedy NMS over score-sorted segments, one CTA per segment.
//
//   VARIANT 0: YOLO class-aware merge-NMS (model/YOLOV3.py:320-333): segment = one (image, class);
//              IoU_+1 with +1e-16 (accuracy.py:54-68) > nms_thres removes a row and attributes it
//              to the FIRST keeper that hit it; the keeper's box becomes the conf-weighted mean of
//              its cluster (YOLOV3.py:327-329).
//   VARIANT 1/2: SSD / RetinaNet class-agnostic greedy NMS on the top-k rows of an image
//              (model/SSD.py:270-302), 'union' / 'min' overlap, survive when ovr <= thresh.
//
// A segment is processed in chunks of CT (384, or 192 for short segments) score-ordered rows:
//   phase A  rows of the chunk vs. the keepers of EARLIER chunks, as 32 x 32 warp tasks through the vectorised half2
//            bound test + exact test; owner = lowest keeper index that removes the row (atomicMin);
//   phase B  lower-triangular 64-bit overlap masks inside the chunk: warp tasks (mask word x 32-row group), column
//            bounds broadcast from shared memory, row box in registers, verdict bits accumulated on the FMA pipe;
//   resolve  greedy keep / remove as a parallel fixed point over all warps (a row is decided once a kept row hits it or
//            every row in its mask is decided and not kept); a serial 32-ballot sweep finishes adversarial chains;
//   owners   first kept bit of (mask & kept) -> cluster owner; members compacted in row order and chained per owner;
//   merge    every keeper walks its chain (deterministic fp32 order: keeper first, then members by descending score);
//            rows owned by keepers of earlier chunks are bitonic-sorted by (owner, row) and summed per owner run.
// IoU arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) and IEEE division so
// the keep decisions are bit-identical to the reference's fp32 CPU path.
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>

#include "yolo_ws.cuh"

namespace b200det {

constexpr int kNmsT = 384;                 // rows per chunk (384: 6 CTAs/SM; 512 measured 320 vs 295 us at the headline)
constexpr int kNmsThreads = 256;
constexpr int kNmsRounds = 12;             // parallel fixed-point rounds before the serial sweep takes over
constexpr int kNmsLevels = 32;             // TAB: quantisation levels of the per-axis bound tables (one lane per level in the scans)
constexpr int kNmsQCap = 128;              // TAB: candidate-pair queue entries per warp

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
    uint32_t* chunk_cnt;        // [B][n_chunks] kept rows per 1024 score ranks (VARIANT 0)
    int n_chunks;
    int n_pad, C;
    float thr;
    // VARIANT 1/2 (prior NMS)
    int topk, compat;
    const uint32_t* tile_prefix;   // [B][n_tiles] exclusive scan of tile_count (filtered index of a slot)
    int n_tiles;
    const uint32_t* orig;          // [B][n_pad] prior index of a slot
    const float* loc;              // [B][P][4]  the call's inputs: box and label of ANY prior are re-derived from them for
    const float* pcls;             // [B][P][C]  the few output rows that need it (quirk ii)
    const float* priors;           // [P][4]
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

// true when the earlier (higher-score) box `a` removes box `b`.
// The decision is bit-identical to the reference's fp32 `inter / union > thr` (resp. `!(ovr <= thr)`), but the
// IEEE division is only executed inside a +-2^-20 guard band around the threshold: outside it the comparison
// `inter <> thr * den` is provably equivalent (rounding errors of the two products are < 2^-22 relative), and
// inter == 0 (no overlap: ~90% of all pairs) is decided without touching the divider at all (0/den = +-0).
template <int VARIANT>
__device__ __forceinline__ bool removes(const float4 a, const float4 b, const float thr) {
    const float aa = box_area_plus1(a), ab = box_area_plus1(b);
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
    const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float iw = fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.0f), 0.0f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.0f), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    float den;
    if (VARIANT == 0) den = __fadd_rn(__fsub_rn(__fadd_rn(aa, ab), inter), 1e-16f);   // accuracy.py:66
    else if (VARIANT == 1) den = __fsub_rn(__fadd_rn(aa, ab), inter);                 // SSD.py:292
    else den = fminf(ab, aa);                                                         // SSD.py:294
    const bool den_ok = den == den && den != 0.0f;
    if (inter == 0.0f) {
        // quotient is +-0 (or NaN when den is 0/NaN)
        if (VARIANT == 0) return den_ok && thr < 0.0f;          // 0 > thr            (YOLOV3.py:323)
        return !(den_ok && thr >= 0.0f);                        // !(0 <= thr)        (SSD.py:298)
    }
    if (den > 0.0f && thr > 0.0f) {
        const float q = __fmul_rn(thr, den);
        if (inter > __fmul_rn(q, 1.000001f)) return true;       // ratio > thr*(1+2^-21)  =>  fl(ratio) > thr
        if (inter < __fmul_rn(q, 0.999999f)) return false;      // ratio < thr*(1-2^-21)  =>  fl(ratio) < thr
    }
    const float ratio = __fdiv_rn(inter, den);
    if (VARIANT == 0) return ratio > thr;
    return !(ratio <= thr);
}

// The same decision for VARIANT 0 with nms_thres >= 0 and the two +1-areas precomputed (identical fp32 values: box_area_plus1
// of the same boxes), for the pair loops of the FAST path: inter == 0 can never exceed a non-negative threshold.
__device__ __forceinline__ bool removes_v0_areas(const float4 a, const float aa, const float4 b, const float ab, const float thr) {
    const float iw = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.0f), 0.0f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.0f), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    if (inter == 0.0f) return false;                                    // quotient +-0 or NaN: never > thr >= 0
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(aa, ab), inter), 1e-16f);   // accuracy.py:66
    if (den > 0.0f && thr > 0.0f) {
        const float q = __fmul_rn(thr, den);
        if (inter > __fmul_rn(q, 1.000001f)) return true;
        if (inter < __fmul_rn(q, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, den) > thr;
}

// Conservative half2 summary of a corner box for the pair pre-filter (one 8-byte shared-memory word per row):
//   lo = (x1, y1) rounded DOWN,  u = (x2 + 1, y2 + 1) - thr' * (w + 1, h + 1) rounded UP,  thr' = thr * (1 - 2^-7).
// A pair can only reach IoU_+1 > thr if  min(u_a, u_b) > max(lo_a, lo_b)  in BOTH axes, because
//   IoU <= inter / max(area_a, area_b) <= iw / max(w_a + 1, w_b + 1)   (and likewise for ih), i.e.
//   iw = min(x2_a, x2_b) + 1 - max(x1_a, x1_b) > thr * (w_a + 1)  and  > thr * (w_b + 1), which gives
//   (x2_a + 1) - thr (w_a + 1) > max(x1)  and the same for b.   (All roundings go the passing way; the 2^-7 covers the fp32
//   rounding of the reference's own ratio.)  The test never rejects a pair the reference would remove (for thr >= 0:
//   degenerate boxes with w + 1 <= 0 only ADD candidates, which the exact test then rejects).  It is weaker than testing
//   min(hi) - max(lo) against max(t_a, t_b) — it lets a small box inside a large one through — but costs 3 half2
//   instructions + the mask accumulate instead of 5, and one 64-bit shared load instead of 64 + 32: on the headline data
//   it passes 2.3 % of the pairs instead of 0.9 % (2.8 instead of 1.8 exact tests per lane and 32 x 32 task).
__device__ __forceinline__ uint2 box_bounds_h2(const float4 b, const float thr) {
    const __half2 lo = __halves2half2(__float2half_rd(b.x), __float2half_rd(b.y));
    const float thr_lo = __fmul_rd(thr, 1.0f - 0.0078125f);
    const float tx = __fmul_rd(thr_lo, __fadd_rd(__fsub_rd(b.z, b.x), 1.0f));
    const float ty = __fmul_rd(thr_lo, __fadd_rd(__fsub_rd(b.w, b.y), 1.0f));
    const __half2 u = __halves2half2(__float2half_ru(__fsub_ru(__fadd_ru(b.z, 1.0f), tx)),
                                     __float2half_ru(__fsub_ru(__fadd_ru(b.w, 1.0f), ty)));
    uint2 r;
    r.x = *reinterpret_cast<const unsigned*>(&lo);
    r.y = *reinterpret_cast<const unsigned*>(&u);
    return r;
}
__device__ __forceinline__ bool may_remove(const uint2 qa, const uint2 qb) {
    const __half2 lo = __hmax2(*reinterpret_cast<const __half2*>(&qa.x), *reinterpret_cast<const __half2*>(&qb.x));
    const __half2 u = __hmin2(*reinterpret_cast<const __half2*>(&qa.y), *reinterpret_cast<const __half2*>(&qb.y));
    return __hbgt2(u, lo);
}

// The same test for 32 consecutive rows of shared memory against one box, as a 32-bit mask.  The ALU pipe is the
// bottleneck of this kernel (ncu: alu 70 %, fma 16 %), so the per-pair verdicts are not turned into predicates and
// bit-inserted (HSET2 + ISETP + SEL + IADD3 on the ALU pipe) but accumulated on the FMA pipe: HSET2.BF yields 1.0 / 0.0
// per axis, `acc = verdict * 2^k + acc` (HFMA2) collects 8 pairs per half in the mantissa of 1024 + v (exact, v < 256),
// and three byte permutes + one AND per 32 pairs combine the two axes.  Per pair: 2 HMNMX2 + HSET2 (ALU), HFMA2 (FMA).
// The ALU pipe (half rate: HMNMX2, HSET2, integer / logic) is what this kernel saturates, so the test is written without the two
// min / max:  min(u_a, u_b) > max(lo_a, lo_b)  <=>  u_a > lo_b  and  u_b > lo_a  (and the two self-conditions u > lo of a box
// with itself, which hold for every box with w + 1 > 0 — a box with w + 1 <= 0 has an empty +1-intersection with everything
// and is rejected by the exact test).  Two HSET2.BF per pair (ALU) instead of two HMNMX2 + HSET2, and the two verdict streams
// are accumulated separately on the FMA pipe (two HFMA2) and ANDed as bit masks at the end.
// Measured (NMS kernel, headline / dense-crowd shard): 3 ALU + 1 FMA per pair (HMNMX2 x2, HSET2, HFMA2) 225 / 343 us; this form
// (2 ALU + 2 FMA) 219 / 311 us; the compares moved to the FMA pipe as well — bounds stored as quarter-pixel integers so that
// [u > lo] = saturate(4u - 4lo) is one HADD2.SAT — 1 ALU + 3 FMA 220 / 328 us, 0 ALU + 4 FMA 219 / 356 us: the fp16x2 FMA-pipe
// instructions are no cheaper than HSET2, the split across both pipes is what pays.
__device__ __forceinline__ unsigned may_remove_mask32(const uint2* __restrict__ q, const uint2 qj) {
    const __half2 lo_j = *reinterpret_cast<const __half2*>(&qj.x);
    const __half2 u_j = *reinterpret_cast<const __half2*>(&qj.y);
    unsigned u1[4], u2[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        __half2 acc1 = __floats2half2_rn(1024.0f, 1024.0f), acc2 = acc1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint2 qa = q[g * 8 + k];
            const __half2 v1 = __hgt2(*reinterpret_cast<const __half2*>(&qa.y), lo_j);       // u_a > lo_j, per axis: 1.0 / 0.0
            const __half2 v2 = __hgt2(u_j, *reinterpret_cast<const __half2*>(&qa.x));        // u_j > lo_a
            const float w = (float)(1 << k);
            const __half2 w2 = __floats2half2_rn(w, w);
            acc1 = __hfma2(v1, w2, acc1);
            acc2 = __hfma2(v2, w2, acc2);
        }
        u1[g] = *reinterpret_cast<const unsigned*>(&acc1);
        u2[g] = *reinterpret_cast<const unsigned*>(&acc2);
    }
    // byte 0 / byte 2 of u[g] = the 8 verdicts of the x / y axis
    const unsigned ra1 = __byte_perm(u1[0], u1[1], 0x6240), rb1 = __byte_perm(u1[2], u1[3], 0x6240);
    const unsigned ra2 = __byte_perm(u2[0], u2[1], 0x6240), rb2 = __byte_perm(u2[2], u2[3], 0x6240);
    const unsigned ra = ra1 & ra2, rb = rb1 & rb2;           // x bits 0..15 | y bits 0..15 << 16 ; x 16..31 | y 16..31 << 16
    return __byte_perm(ra, rb, 0x5410) & __byte_perm(ra, rb, 0x7632);
}

// ---- TAB pre-filter -------------------------------------------------------------------------------------------------
// The same necessary condition as box_bounds_h2 / may_remove (per axis: u_a > lo_b and u_b > lo_a with lo = x1,
// u = x2 + 1 - thr' (w + 1)), but evaluated for 64 columns at once by table lookup instead of pair by pair:
// every bound is quantised to one of kNmsLevels levels with a MONOTONE map (level(v) = clamp(floor((v - base) * inv)); base / inv
// from the min lo and max u of the chunk), so lo_i < u_j implies level(lo_i) <= level(u_j).  Per axis two tables of
// column bitmaps:  A[q] = { columns i : level(lo_i) <= q }  and  B[q] = { i : level(u_i) >= q };  the columns that can
// remove / be removed by row j are  A_x[level(u_x_j)] & B_x[level(lo_x_j)] & A_y[level(u_y_j)] & B_y[level(lo_y_j)]:
// four 64-bit shared loads and three ANDs per (row, 64 columns) instead of 64 x (load + 4 half2 instructions).  It lets
// about twice as many pairs through as the half2 test (quantisation slack); those pairs are not tested lane by lane in a
// divergent loop but pushed into a per-warp queue and tested 32 at a time with every lane busy.  Rows with a non-finite
// bound never remove anything and are never removed (their IoU is 0 or NaN), so whatever level they land on is fine.
__device__ __forceinline__ void box_bounds_f32(const float4 b, const float thr, float& lox, float& loy, float& ux, float& uy) {
    const float thr_lo = __fmul_rd(thr, 1.0f - 0.0078125f);
    lox = b.x;
    loy = b.y;
    ux = __fsub_ru(__fadd_ru(b.z, 1.0f), __fmul_rd(thr_lo, __fadd_rd(__fsub_rd(b.z, b.x), 1.0f)));
    uy = __fsub_ru(__fadd_ru(b.w, 1.0f), __fmul_rd(thr_lo, __fadd_rd(__fsub_rd(b.w, b.y), 1.0f)));
}
__device__ __forceinline__ int ordered_int(float f) {            // signed-int order == float order (finite values)
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }
__device__ __forceinline__ bool finite_f(float f) { return (__float_as_uint(f) & 0x7F800000u) != 0x7F800000u; }
__device__ __forceinline__ unsigned bound_level(float v, float base, float inv) {
    const int l = __float2int_rd(__fmul_rn(__fsub_rn(v, base), inv));           // NaN -> 0, +-inf -> saturates
    return (unsigned)min(kNmsLevels - 1, max(0, l));
}

#ifdef B200DET_NMS_DEBUG
__device__ unsigned long long g_nms_dbg[8];
#endif
// Development aid: phase timestamps of the first chunk of every segment CTA (see tools/stage_timing.py --trace-nms).
__device__ unsigned long long* g_nms_trace = nullptr;
__device__ __forceinline__ void nms_stamp(unsigned long long* tr, int point) {
#ifdef B200DET_NMS_TRACE          // compiled out by default: the extra live pointer costs ~15 % of the kernel
    if (tr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tr[point] = t;
    }
#endif
}

// FAST: the pre-filter is sound when "no overlap" implies "not removed", i.e. VARIANT 0 with nms_thres >= 0.
// NT threads per CTA: 256 for the usual ~300-row segments (6 CTAs/SM); 512 when the average segment is longer than a chunk
// (few classes, dense crowds): there are then few CTAs and the phase-A / phase-B warp tasks of a segment are the parallelism.
// CT rows per chunk: 384 (the triangle of 6 mask words per row) — or 192 with 128 threads when the average segment is
// short (YOLOv3-416: 133 rows): a CTA then needs 14 KB instead of 32 KB of shared memory, 12 of them fit an SM and twice as
// many segments hide each other's fixed latencies (loads, barriers, global stores).
template <int VARIANT, bool FAST, int NT, int CT, bool TAB = false>
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : (TAB ? (NT == 128 ? 10 : (NT == 256 ? 5 : 2)) : (NT == 128 ? 12 : (NT == 256 ? 6 : 3))))
nms_segment_kernel(const NmsParams p) {
    static_assert(!TAB || (VARIANT == 0 && FAST), "the table pre-filter is built for the YOLO class-aware NMS with nms_thres >= 0");
    constexpr int NW = CT / 64;                       // mask words per full row
    constexpr int TRI = 32 * NW * (NW + 1);           // packed lower-triangular rows
    constexpr int STAGE = (TRI * 8 / 24) / 32 * 32;   // earlier keepers staged per phase-A round, in the aliased mask triangle
    __shared__ float4 s_box[CT];
    __shared__ float s_conf[CT];
    __shared__ uint2 s_q[CT];        // half2 lo, u   (see box_bounds_h2)
    __shared__ unsigned long long s_L[TRI];
    __shared__ unsigned long long s_kept[NW];
    __shared__ __align__(16) uint2 s_state[CT / 32];   // per 32-row group: x = kept rows, y = decided rows
    __shared__ unsigned long long s_member[NW];
    __shared__ int s_wpre[NT / 32][NW + 1];              // per warp: exclusive prefix of the kept-row counts per mask word
    __shared__ unsigned long long s_keptw[NT / 32][NW];  // per warp: its own copy of the kept rows (built without a CTA barrier)
    __shared__ int s_own[CT];
    __shared__ int s_pre[CT];
    // phase-A staging (keepers of earlier chunks) lives in the mask triangle: phase A of a chunk is over before phase B
    // writes the triangle, and the triangle of the previous chunk is dead by then.  (Shared memory per CTA decides the
    // carve-out: 5 CTAs x <= 39 KB fit the 196 KB setting and leave 32 KB of L1 for the box gathers; 42 KB per CTA measured
    // 248 vs 238 us.)
    static_assert(STAGE * (16 + 8) <= TRI * 8, "phase-A staging (box + half2 summary) must fit the mask triangle");
    float4* const s_kb = reinterpret_cast<float4*>(s_L);
    uint2* const s_kq = reinterpret_cast<uint2*>(s_L + STAGE * 2);
    __shared__ int s_last_members;
    __shared__ uint32_t s_mlist[CT];     // members of in-chunk clusters in ascending row order: owner index << 10 | row
    __shared__ __align__(4) uint8_t s_nzw[CT];   // per row: which of its mask words are non-zero (most rows: none)
    static_assert(NW <= 8, "one byte of non-zero-word flags per row");
    __shared__ uint32_t s_rank[CT];      // score rank of the row, fetched with the boxes so that the owner / merge phases
                                            // do not wait on L2 for it again (VARIANT 0)
    __shared__ int16_t s_next[CT];       // member list position of the next member of the same cluster, or -1
    __shared__ int16_t s_first[CT];      // per in-chunk keeper ordinal: list position of its first member, or -1
    // ... of its last member so far, while the chains are built (the shared-memory budget of 6 CTAs per SM inside the
    // 196 KB carve-out is 32.6 KB per CTA)
    __shared__ int16_t s_last[CT];
    __shared__ int s_mpre[CT / 32 + 1];  // exclusive prefix of the member bitmap's popcounts
    // +1-areas of the chunk's rows for the exact pair tests of phase B (FAST): dead before the merge phase builds s_mlist
    float* const s_area = reinterpret_cast<float*>(s_mlist);
    __shared__ uint16_t s_task[NW * (NW + 1) + 2];       // phase-B task -> mask word << 8 | row group inside the word's range
    // TAB pre-filter (see box_bounds_f32): bound tables, per-row levels, per-warp candidate queues, chunk range
    __shared__ __align__(16) unsigned long long s_tab[TAB ? 4 : 1][TAB ? kNmsLevels : 1][TAB ? NW : 1];
    __shared__ uchar4 s_lvl[TAB ? CT : 1];           // level(u_x), level(lo_x), level(u_y), level(lo_y) of the row
    __shared__ uint32_t s_queue[TAB ? NT / 32 : 1][TAB ? kNmsQCap : 1];
    __shared__ int s_rng[4];                         // ordered-int min lo_x, max u_x, min lo_y, max u_y of the chunk

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
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
    const bool single = n <= CT;
    const float thr = p.thr;

#ifdef B200DET_NMS_TRACE
    unsigned long long* tr = (VARIANT == 0 && g_nms_trace) ? g_nms_trace + ((size_t)b * gridDim.x + blockIdx.x) * 8 : nullptr;
#else
    unsigned long long* tr = nullptr;
#endif
    nms_stamp(tr, 0);
    int Kprev = 0;          // keepers found in earlier chunks
    int last_k = -1;        // VARIANT 1/2: index of the last keeper so far
    if (tid == 0) s_last_members = 0;
    if (TAB) {
        if (tid < 4) s_rng[tid] = (tid & 1) ? INT_MIN : INT_MAX;
        __syncthreads();
    }

    for (int c0 = s; c0 < e; c0 += CT) {
        const int nc = min(CT, e - c0);
        const int Wc = (nc + 63) >> 6;
        if (TAB) {
            uint4* z = reinterpret_cast<uint4*>(&s_tab[0][0][0]);
            for (int i = tid; i < 4 * kNmsLevels * NW / 2; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }

        // ---- load the chunk ----------------------------------------------------------------
        // both rows of a thread are fetched together: payload -> slot -> box is a chain of two L2 round trips, and a
        // row-at-a-time loop would walk it twice back to back
        {
            static_assert(CT <= 2 * NT, "a thread loads at most two rows of a chunk");
            const int j0 = tid, j1 = tid + NT;
            uint32_t slot0 = 0, slot1 = 0;
            if (j0 < nc) slot0 = p.spay[img + c0 + j0] & kSlotMask;
            if (j1 < nc) slot1 = p.spay[img + c0 + j1] & kSlotMask;
            if (VARIANT == 0) {
                if (j0 < nc) s_rank[j0] = p.srank[img + c0 + j0];
                if (j1 < nc) s_rank[j1] = p.srank[img + c0 + j1];
            }
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            float f0 = 0.f, f1 = 0.f;
            if (j0 < nc) { b0 = p.box4[img + slot0]; f0 = p.cc2[img + slot0].x; }
            if (j1 < nc) { b1 = p.box4[img + slot1]; f1 = p.cc2[img + slot1].x; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = h ? j1 : j0;
                if (j >= nc) continue;
                const float4 bx = h ? b1 : b0;
                s_box[j] = bx;
                s_conf[j] = h ? f1 : f0;
                if (FAST) s_q[j] = box_bounds_h2(bx, thr);
                if (FAST && VARIANT == 0) s_area[j] = box_area_plus1(bx);
                s_pre[j] = -1;
                s_first[j] = -1;
                s_nzw[j] = 0u;
            }
            if (TAB) {
                // range of the chunk's bounds (finite ones): per-warp redux, one shared atomic per warp and value
                int mn_x = INT_MAX, mx_x = INT_MIN, mn_y = INT_MAX, mx_y = INT_MIN;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if ((h ? j1 : j0) >= nc) continue;
                    float lox, loy, ux, uy;
                    box_bounds_f32(h ? b1 : b0, thr, lox, loy, ux, uy);
                    if (finite_f(lox) && finite_f(ux)) { mn_x = min(mn_x, ordered_int(lox)); mx_x = max(mx_x, ordered_int(ux)); }
                    if (finite_f(loy) && finite_f(uy)) { mn_y = min(mn_y, ordered_int(loy)); mx_y = max(mx_y, ordered_int(uy)); }
                }
                mn_x = __reduce_min_sync(0xFFFFFFFFu, mn_x); mx_x = __reduce_max_sync(0xFFFFFFFFu, mx_x);
                mn_y = __reduce_min_sync(0xFFFFFFFFu, mn_y); mx_y = __reduce_max_sync(0xFFFFFFFFu, mx_y);
                if (lane == 0) {
                    atomicMin(&s_rng[0], mn_x); atomicMax(&s_rng[1], mx_x);
                    atomicMin(&s_rng[2], mn_y); atomicMax(&s_rng[3], mx_y);
                }
            }
        }
        if (tid < NW) s_member[tid] = 0ull;      // 32-row groups past nc are never written by the ballots below
        if (FAST && !TAB) {
            // phase-B task table: task t -> (mask word w, row group 2w + rem); one thread per task resolves its own index once
            const int ngr = (nc + 31) >> 5;
            int nt = 0;
            for (int w = 0; w < Wc; ++w) nt += ngr - 2 * w;
            if (tid < nt) {
                int w = 0, rem = tid;
                while (rem >= ngr - 2 * w) { rem -= ngr - 2 * w; ++w; }
                s_task[tid] = (uint16_t)((w << 8) | rem);
            }
        }
        __syncthreads();

        // ---- phase A: against keepers of earlier chunks ----------------------------------------
        // FAST: one task = 32 rows x 32 staged keepers, through the same vectorised bound test as phase B, dealt round-robin
        // to the warps; the owner of a row is the LOWEST keeper index that removes it (atomicMin), rows already owned after
        // an earlier stage are skipped.  (The row-per-thread loop over all keepers it replaces made large segments — few
        // classes, dense crowds — crawl: 930 us for the config-5 shard.)
        if (FAST && Kprev > 0) {
            for (int j = tid; j < nc; j += NT) s_own[j] = INT_MAX;      // s_own doubles as the hit slot until the owners phase
        }
        for (int kt = 0; kt < Kprev; kt += STAGE) {
            const int nk = min(STAGE, Kprev - kt);
            for (int i = tid; i < nk; i += NT) {
                const float4 kb = p.kbox[img + s + kt + i];
                s_kb[i] = kb;
                if (FAST) s_kq[i] = box_bounds_h2(kb, thr);
            }
            __syncthreads();
            if (FAST) {
                const int nkb = (nk + 31) >> 5, ngr = (nc + 31) >> 5;
                // task index -> (row group, keeper block) without an integer division (nkb <= 14, t < 2^11: the float
                // reciprocal with the +0.5 nudge is exact; the division was 6 % of the kernel on the dense-crowd shard)
                const float inv_nkb = __frcp_rn((float)nkb);
                for (int t = tid >> 5; t < ngr * nkb; t += NT / 32) {
                    const int g = __float2int_rz(__fmul_rn((float)t + 0.5f, inv_nkb)), kb0 = (t - g * nkb) << 5;
                    const int j = (g << 5) + lane;
                    const bool active = j < nc && s_pre[j] < 0;
                    if (!__any_sync(0xFFFFFFFFu, active)) continue;
                    if (!active) continue;
                    const float4 bj = s_box[j];
                    const uint2 qj = s_q[j];
                    unsigned cand = may_remove_mask32(s_kq + kb0, qj);
                    if (nk - kb0 < 32) cand &= (1u << (nk - kb0)) - 1u;                   // stale entries past the stage
                    while (cand) {
                        const int k = __ffs((int)cand) - 1;
                        cand &= cand - 1u;
                        if (removes<VARIANT>(s_kb[kb0 + k], bj, thr)) { atomicMin(&s_own[j], kt + kb0 + k); break; }
                    }
                }
                __syncthreads();
                for (int j = tid; j < nc; j += NT)
                    if (s_pre[j] < 0 && s_own[j] != INT_MAX) s_pre[j] = s_own[j];
            } else {
                for (int j = tid; j < nc; j += NT) {
                    if (s_pre[j] >= 0) continue;
                    const float4 bj = s_box[j];
                    for (int k = 0; k < nk; ++k) {
                        if (removes<VARIANT>(s_kb[k], bj, thr)) { s_pre[j] = kt + k; break; }
                    }
                }
            }
            __syncthreads();
        }

        if (c0 == s) nms_stamp(tr, 1);
        if (TAB) {
            // ---- bound tables of the chunk's live rows (see box_bounds_f32) ----------------------------------------
            unsigned* const tab32 = reinterpret_cast<unsigned*>(&s_tab[0][0][0]);       // [4][levels][2 * NW] 32-bit halves
            const float base_x = ordered_float(s_rng[0]), top_x = ordered_float(s_rng[1]);
            const float base_y = ordered_float(s_rng[2]), top_y = ordered_float(s_rng[3]);
            float inv_x = top_x > base_x ? __fdiv_rn((float)kNmsLevels, __fsub_rn(top_x, base_x)) : 0.0f;
            float inv_y = top_y > base_y ? __fdiv_rn((float)kNmsLevels, __fsub_rn(top_y, base_y)) : 0.0f;
            if (!finite_f(inv_x)) inv_x = 0.0f;
            if (!finite_f(inv_y)) inv_y = 0.0f;
            for (int j = tid; j < nc; j += NT) {
                if (s_pre[j] >= 0) continue;                                 // removed by an earlier chunk: not a column
                float lox, loy, ux, uy;
                box_bounds_f32(s_box[j], thr, lox, loy, ux, uy);
                const unsigned l_ux = bound_level(ux, base_x, inv_x), l_lox = bound_level(lox, base_x, inv_x);
                const unsigned l_uy = bound_level(uy, base_y, inv_y), l_loy = bound_level(loy, base_y, inv_y);
                s_lvl[j] = make_uchar4((unsigned char)l_ux, (unsigned char)l_lox, (unsigned char)l_uy, (unsigned char)l_loy);
                const unsigned bit = 1u << (j & 31);
                const int h = j >> 5;
                atomicOr(&tab32[((0 * kNmsLevels + l_lox) * 2 * NW) + h], bit);           // A_x: exact level of lo_x
                atomicOr(&tab32[((1 * kNmsLevels + l_ux) * 2 * NW) + h], bit);            // B_x: exact level of u_x
                atomicOr(&tab32[((2 * kNmsLevels + l_loy) * 2 * NW) + h], bit);
                atomicOr(&tab32[((3 * kNmsLevels + l_uy) * 2 * NW) + h], bit);
            }
            __syncthreads();
            if (tid < 4) s_rng[tid] = (tid & 1) ? INT_MIN : INT_MAX;          // everybody has read the range: reset for the next chunk
            // A[q] = OR of the exact levels <= q (prefix over the lanes = levels), B[q] = OR of the levels >= q (suffix)
            static_assert(kNmsLevels == 32, "one lane per level");
            for (int c = tid >> 5; c < 4 * 2 * NW; c += NT / 32) {
                const int t = c / (2 * NW), h = c - t * 2 * NW;
                unsigned* cell = &tab32[((t * kNmsLevels + lane) * 2 * NW) + h];
                unsigned v = *cell;
                if (t & 1) {
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned u = __shfl_down_sync(0xFFFFFFFFu, v, o);
                        if (lane + o < 32) v |= u;
                    }
                } else {
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned u = __shfl_up_sync(0xFFFFFFFFu, v, o);
                        if (lane >= o) v |= u;
                    }
                }
                *cell = v;
            }
            __syncthreads();
        }
        // ---- phase B: lower-triangular overlap masks inside the chunk --------------------------
        // One task = (mask word w, 32-row group g >= 2w).  The tasks are dealt round-robin to the warps: looping over
        // the words with rows striped over the CTA leaves the high warps idle for the later words (6 task slots for
        // warps 0-1 vs 2 for warps 6-7 at ~315 rows), which showed up as barrier stalls.
        const int ngroups_b = (nc + 31) >> 5;
        int n_tasks = 0;
        for (int w = 0; w < Wc; ++w) n_tasks += ngroups_b - 2 * w;
        if (TAB) {
            // candidate columns by table lookup; candidate pairs go through the warp's queue and are tested 32 at a time
            uint32_t* const qw = s_queue[tid >> 5];
            int fill = 0;
            auto test_pair = [&](const uint32_t ent) {
                const int j = (int)(ent >> 10), i = (int)(ent & 1023u);
                if (removes<VARIANT>(s_box[i], s_box[j], thr)) {
                    atomicOr(reinterpret_cast<unsigned*>(&s_L[tri_off(j) + (i >> 6)]) + ((i >> 5) & 1), 1u << (i & 31));
                    atomicOr(reinterpret_cast<unsigned*>(s_nzw) + (j >> 2), (1u << (i >> 6)) << (8 * (j & 3)));
                }
            };
            for (int t = tid >> 5; t < n_tasks; t += NT / 32) {
                int w = 0, rem = t;
                while (rem >= ngroups_b - 2 * w) { rem -= ngroups_b - 2 * w; ++w; }
                const int i0 = w << 6;
                const int j = ((2 * w + rem) << 5) + lane;
                unsigned long long cand = 0ull;
                if (j < nc) {
                    if (s_pre[j] < 0) {
                        const uchar4 lv = s_lvl[j];
                        cand = s_tab[0][lv.x][w] & s_tab[1][lv.y][w] & s_tab[2][lv.z][w] & s_tab[3][lv.w][w];
                        const int ni = j - i0;                                     // >= 0: columns of the word before row j
                        if (ni < 64) cand &= (1ull << ni) - 1ull;
#ifdef B200DET_NMS_DEBUG
                        {
                            const unsigned long long tri = ni < 64 ? (1ull << ni) - 1ull : ~0ull;
                            atomicAdd(&g_nms_dbg[0], (unsigned long long)__popcll(tri));
                            atomicAdd(&g_nms_dbg[1], (unsigned long long)__popcll(tri & s_tab[0][lv.x][w]));
                            atomicAdd(&g_nms_dbg[2], (unsigned long long)__popcll(tri & s_tab[1][lv.y][w]));
                            atomicAdd(&g_nms_dbg[3], (unsigned long long)__popcll(tri & s_tab[2][lv.z][w]));
                            atomicAdd(&g_nms_dbg[4], (unsigned long long)__popcll(tri & s_tab[3][lv.w][w]));
                            atomicAdd(&g_nms_dbg[5], (unsigned long long)__popcll(cand));
                            atomicAdd(&g_nms_dbg[6], (unsigned long long)(lv.x + 256 * lv.y));
                            atomicAdd(&g_nms_dbg[7], (unsigned long long)(lv.z + 256 * lv.w));
                        }
#endif
                    }
                    s_L[tri_off(j) + w] = 0ull;                                    // hits are OR-ed in by test_pair
                }
                for (;;) {
                    const int cnt = __popcll(cand);
                    int inc = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                        if (lane >= o) inc += u;
                    }
                    const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
                    if (total == 0) break;
                    int pos = fill + inc - cnt;
                    while (cand && pos < kNmsQCap) {
                        const int k = __ffsll((long long)cand) - 1;
                        cand &= cand - 1ull;
                        qw[pos++] = ((uint32_t)j << 10) | (uint32_t)(i0 + k);
                    }
                    fill = min(kNmsQCap, fill + total);
                    __syncwarp();
                    int done = 0;
                    for (; fill - done >= 32; done += 32) test_pair(qw[done + lane]);
                    const int r = fill - done;                                     // < 32 entries carried to the next round
                    const uint32_t carry = lane < r ? qw[done + lane] : 0u;
                    __syncwarp();
                    if (lane < r) qw[lane] = carry;
                    __syncwarp();
                    fill = r;
                }
            }
            if (lane < fill) test_pair(qw[lane]);
        }
        for (int t = tid >> 5; !TAB && t < n_tasks; t += NT / 32) {
            int w = 0, rem = t;
            if (FAST) {
                const unsigned tk = s_task[t];
                w = (int)(tk >> 8); rem = (int)(tk & 255u);
            } else {
                while (rem >= ngroups_b - 2 * w) { rem -= ngroups_b - 2 * w; ++w; }
            }
            const int i0 = w << 6;
            {
                const int j = ((2 * w + rem) << 5) + lane;
                if (j >= nc) continue;
                unsigned long long bits = 0ull;
                if (s_pre[j] < 0) {
                    const float4 bj = s_box[j];
                        const int ni = min(64, j - i0);
                    if (FAST) {
                        // pass 1: half2 bounding test, branch-free; pass 2: exact test on the few candidates
                        const uint2 qj = s_q[j];
                        unsigned c_lo = 0u, c_hi = 0u;
                        // all 64 columns of the word are tested (rows >= j hold valid or stale-but-harmless bounds)
                        // and the columns >= ni are masked off afterwards: no variable-trip-count loop on the diagonal
                        c_lo = may_remove_mask32(s_q + i0, qj);
                        if (ni > 32) c_hi = may_remove_mask32(s_q + i0 + 32, qj);
                        if (ni < 32) c_lo &= (1u << ni) - 1u;
                        else if (ni < 64) c_hi &= (1u << (ni - 32)) - 1u;
                        unsigned long long cand = ((unsigned long long)c_hi << 32) | c_lo;
                        if (VARIANT == 0) {
                            const float aj = s_area[j];
                            while (cand) {
                                const int k = __ffsll((long long)cand) - 1;
                                cand &= cand - 1ull;
                                if (removes_v0_areas(s_box[i0 + k], s_area[i0 + k], bj, aj, thr)) bits |= 1ull << k;
                            }
                        } else {
                            while (cand) {
                                const int k = __ffsll((long long)cand) - 1;
                                cand &= cand - 1ull;
                                if (removes<VARIANT>(s_box[i0 + k], bj, thr)) bits |= 1ull << k;
                            }
                        }
                    } else {
                        for (int k = 0; k < ni; ++k) {
                            if (removes<VARIANT>(s_box[i0 + k], bj, thr)) bits |= 1ull << k;
                        }
                    }
                }
                s_L[tri_off(j) + w] = bits;
                if (bits) atomicOr(reinterpret_cast<unsigned*>(s_nzw) + (j >> 2), (1u << w) << (8 * (j & 3)));
            }
        }
        __syncthreads();

        if (c0 == s) nms_stamp(tr, 2);
        // ---- greedy resolution ------------------------------------------------------------------------------
        // Row j is kept iff no KEPT earlier row has a bit in its mask.  The masks are lower-triangular, so the solution
        // is unique and can be reached as a fixed point in parallel: a row is decided once one kept row hits it
        // (removed) or every row in its mask is decided and not kept (kept).  Dependency chains are short on real
        // inputs (a few rounds for the whole CTA, instead of one warp walking the rows 32 at a time while seven wait);
        // kept/decided bits of a 32-row group live in one 64-bit word so that a reader sees a consistent pair, and the
        // serial sweep below finishes adversarial chains after kNmsRounds rounds.
        {
            for (int g = tid >> 5; g < CT / 32; g += NT / 32) {
                const int j = (g << 5) + lane;
                const bool live = j < nc && s_pre[j] < 0;
                const bool free_row = live && s_nzw[j] == 0u;                  // overlaps no earlier row: kept
                const unsigned kept0 = __ballot_sync(0xFFFFFFFFu, free_row);
                const unsigned dec0 = __ballot_sync(0xFFFFFFFFu, !live || free_row);
                if (lane == 0) s_state[g] = make_uint2(kept0, dec0);
            }
            __syncthreads();
            int undecided = 1;
            for (int round = 0; round < kNmsRounds && undecided; ++round) {
                bool left = false;
                for (int g = tid >> 5; g < ngroups_b; g += NT / 32) {
                    const int j = (g << 5) + lane;
                    const uint2 mine = s_state[g];
                    if (mine.y == 0xFFFFFFFFu) continue;                       // warp-uniform: group fully decided
                    bool now_kept = false, now_dec = false;
                    if (!((mine.y >> lane) & 1u)) {
                        const unsigned long long* row = &s_L[tri_off(j)];
                        bool hit = false, pend = false;
                        for (unsigned nz = s_nzw[j]; nz; nz &= nz - 1u) {
                            const int w = __ffs((int)nz) - 1;
                            const unsigned long long m = row[w];
                            const uint4 st = *reinterpret_cast<const uint4*>(&s_state[2 * w]);   // {kept, dec} x 2 groups
                            const unsigned long long kept = ((unsigned long long)st.z << 32) | st.x;
                            const unsigned long long dec = ((unsigned long long)st.w << 32) | st.y;
                            hit |= (m & kept) != 0ull;
                            pend |= (m & ~dec) != 0ull;
                        }
                        now_kept = !hit && !pend;
                        now_dec = hit || !pend;
                    }
                    const unsigned kb = __ballot_sync(0xFFFFFFFFu, now_kept);
                    const unsigned db = __ballot_sync(0xFFFFFFFFu, now_dec);
                    if (lane == 0 && db) s_state[g] = make_uint2(mine.x | kb, mine.y | db);
                    left |= (mine.y | db) != 0xFFFFFFFFu;
                }
                undecided = __syncthreads_or(left ? 1 : 0);
            }
            if (undecided) {                              // CTA-uniform: an adversarial chain outlasted the rounds
                if (tid < 32) {
                    if (lane < NW) s_kept[lane] = 0ull;
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
                        // lanes whose in-group mask is empty are decided already; only the others need the serial steps
                        const unsigned dep = __ballot_sync(0xFFFFFFFFu, pre && m != 0u);
                        unsigned kg = cand & ~dep;
                        for (unsigned dd = dep; dd; dd &= dd - 1u) {
                            // every lane below the lowest pending one is final in kg -> that lane can be decided now
                            const bool bit = pre && ((m & kg) == 0u);
                            const unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
                            kg |= bal & (dd & (0u - dd));
                        }
                        if (lane == 0 && kg) s_kept[wl] |= (unsigned long long)kg << ((g & 1) * 32);
                        __syncwarp();
                    }
                }
                __syncthreads();
                if (tid < CT / 32) s_state[tid].x = (unsigned)(s_kept[tid >> 1] >> ((tid & 1) * 32));
                __syncthreads();
            }
        }
        // every warp builds its own copy of the kept words and their prefix from s_state: no CTA barrier, no single thread
        const int wq = tid >> 5;
        {
            unsigned long long kw = 0ull;
            if (lane < NW) kw = ((unsigned long long)s_state[2 * lane + 1].x << 32) | s_state[2 * lane].x;
            const int c = lane < Wc ? __popcll(kw) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane < NW) { s_keptw[wq][lane] = kw; s_wpre[wq][lane] = inc - c; }
            if (lane == NW - 1) s_wpre[wq][NW] = inc;
            __syncwarp();
        }
        const int Kc = s_wpre[wq][NW];

        if (c0 == s) nms_stamp(tr, 3);
        // ---- owners (warp-uniform trip count so that the member bitmap can be built with ballots) ------
        bool any_cross_local = false;
        for (int jb = (tid & ~31); jb < nc; jb += NT) {
            const int j = jb + lane;
            int own = -2;
            bool is_keeper = false, is_member = false;
            if (j < nc) {
                if (s_pre[j] >= 0) {
                    own = s_pre[j];
                    any_cross_local = true;
                } else {
                    const int wj = j >> 6;
                    const unsigned long long below = (1ull << (j & 63)) - 1ull;
                    if ((s_keptw[wq][wj] >> (j & 63)) & 1ull) {
                        own = Kprev + s_wpre[wq][wj] + __popcll(s_keptw[wq][wj] & below);
                        is_keeper = true;
                    } else {
                        const unsigned long long* row = &s_L[tri_off(j)];
                        int i = -1;
                        for (unsigned nz = s_nzw[j]; nz; nz &= nz - 1u) {
                            const int w = __ffs((int)nz) - 1;
                            const unsigned long long h = row[w] & s_keptw[wq][w];
                            if (h) { i = (w << 6) + __ffsll((long long)h) - 1; break; }
                        }
                        if (i >= 0) {   // always: a non-kept, non-presuppressed row was hit by a kept row
                            const int wi = i >> 6;
                            own = Kprev + s_wpre[wq][wi] + __popcll(s_keptw[wq][wi] & ((1ull << (i & 63)) - 1ull));
                            is_member = true;
                        }
                    }
                }
                s_own[j] = own;
                if (VARIANT == 0) {
                    if (!is_keeper) p.kpay[img + s_rank[j]] = kNone;
                }
                if (is_keeper && !(VARIANT == 0 && single)) {
                    p.kbox[img + s + own] = s_box[j];
                    p.kpos[img + s + own] = (uint32_t)(c0 + j);
                }
            }
            const unsigned mb = __ballot_sync(0xFFFFFFFFu, is_member);
            if (lane == 0) reinterpret_cast<unsigned*>(s_member)[jb >> 5] = mb;
        }
        const int any_cross = __syncthreads_or(any_cross_local ? 1 : 0);

        if (VARIANT == 0) {
            if (c0 == s) nms_stamp(tr, 4);
            // ---- merge sums (YOLOV3.py:327-329): keeper first, then its members by descending score ----------
            // Most keepers own nothing and finish at once; the keepers that own members are compacted so that a few
            // full warps walk the (short, ordered) member list instead of every keeper scanning the member bitmap.
            const unsigned* mem32 = reinterpret_cast<const unsigned*>(s_member);
            const int nh = (nc + 31) >> 5;
            if (tid < 32) {
                const int c = lane < nh ? __popc(mem32[lane]) : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane <= CT / 32) s_mpre[lane] = inc - c;
            }
            __syncthreads();
            auto finish = [&](const int j, const int kidx, const float ax, const float ay, const float az, const float aw,
                              const float ws) {
                if (single) {
                    const uint32_t r = s_rank[j];
                    p.kpay[img + r] = p.spay[img + c0 + j];
                    p.mbox[img + r] = make_float4(__fdiv_rn(ax, ws), __fdiv_rn(ay, ws), __fdiv_rn(az, ws), __fdiv_rn(aw, ws));
                    atomicAdd(&p.chunk_cnt[(size_t)b * p.n_chunks + (r >> kEmitShift)], 1u);
                } else {
                    float* acc = p.kacc + (img + s + kidx) * 5;
                    acc[0] = ax; acc[1] = ay; acc[2] = az; acc[3] = aw; acc[4] = ws;
                }
            };
            for (int j = tid; j < nc; j += NT) {
                s_last[j] = -1;
                const unsigned mw = mem32[j >> 5];
                if ((mw >> (j & 31)) & 1u)
                    s_mlist[s_mpre[j >> 5] + __popc(mw & ((1u << (j & 31)) - 1u))] = ((uint32_t)s_own[j] << 10) | (uint32_t)j;
            }
            __syncthreads();
            const int M = s_mpre[nh];
            // chain the members of every cluster in list (= row) order: one warp walks the list 32 entries at a time;
            // the predecessor of an entry is the closest lower lane with the same owner, else the owner's last entry
            // of the earlier steps (s_last)
            if (tid < 32) {
                for (int t0 = 0; t0 < M; t0 += 32) {
                    const int t = t0 + lane;
                    const bool valid = t < M;
                    const int o = valid ? (int)(s_mlist[t] >> 10) - Kprev : -1 - lane;     // distinct dummies
                    if (valid) s_next[t] = -1;
                    const unsigned peers = __match_any_sync(0xFFFFFFFFu, o);
                    const unsigned lower = peers & lanemask_lt();
                    __syncwarp();
                    if (valid) {
                        const int prev = lower ? t0 + 31 - __clz((int)lower) : (int)s_last[o];
                        if (prev >= 0) s_next[prev] = (int16_t)t;
                        else s_first[o] = (int16_t)t;
                        if ((peers >> lane) <= 1u) s_last[o] = (int16_t)t;                // highest lane of the group
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            for (int j = tid; j < nc; j += NT) {
                if (s_pre[j] >= 0 || !((s_keptw[wq][j >> 6] >> (j & 63)) & 1ull)) continue;
                const int kidx = s_own[j];
                const float4 bj = s_box[j];
                const float w0 = s_conf[j];
                float ax = __fmul_rn(w0, bj.x), ay = __fmul_rn(w0, bj.y);
                float az = __fmul_rn(w0, bj.z), aw = __fmul_rn(w0, bj.w), ws = w0;
                for (int t = s_first[kidx - Kprev]; t >= 0; t = s_next[t]) {
                    const int m = (int)(s_mlist[t] & 1023u);
                    const float4 bm = s_box[m];
                    const float wm = s_conf[m];
                    ax = __fadd_rn(ax, __fmul_rn(wm, bm.x));
                    ay = __fadd_rn(ay, __fmul_rn(wm, bm.y));
                    az = __fadd_rn(az, __fmul_rn(wm, bm.z));
                    aw = __fadd_rn(aw, __fmul_rn(wm, bm.w));
                    ws = __fadd_rn(ws, wm);
                }
                finish(j, kidx, ax, ay, az, aw, ws);
            }
            if (any_cross) {
                // Rows of this chunk owned by keepers of earlier chunks add to those keepers' running sums (in row order,
                // per owner).  The rows are compacted, sorted by (owner, row) with a bitonic network in the — now free —
                // mask triangle, and the first row of every owner run walks its run.  (Every earlier keeper scanning all
                // rows of the chunk was 27 % of the kernel on the dense-crowd configuration.)
                __syncthreads();                                            // in-chunk merge is done with s_member / s_mpre
                unsigned long long* s_x = s_L;                              // keys: owner << 32 | row
                unsigned* xbits = reinterpret_cast<unsigned*>(s_member);
                for (int g = tid >> 5; g < CT / 32; g += NT / 32) {
                    const int j = (g << 5) + lane;
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, j < nc && s_pre[j] >= 0);
                    if (lane == 0) xbits[g] = bal;
                }
                __syncthreads();
                if (tid < 32) {
                    const int c = lane < CT / 32 ? __popc(xbits[lane]) : 0;
                    int inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                        if (lane >= o) inc += t;
                    }
                    if (lane <= CT / 32) s_mpre[lane] = inc - c;
                }
                __syncthreads();
                const int Mx = s_mpre[CT / 32];
                int P = 32;
                while (P < Mx) P <<= 1;                                       // <= 512
                for (int i = tid; i < P; i += NT) s_x[i] = ~0ull;
                __syncthreads();
                for (int j = tid; j < nc; j += NT) {
                    if (s_pre[j] >= 0) {
                        const unsigned bw = xbits[j >> 5];
                        s_x[s_mpre[j >> 5] + __popc(bw & ((1u << (j & 31)) - 1u))] = ((unsigned long long)(unsigned)s_pre[j] << 32) | (unsigned)j;
                    }
                }
                __syncthreads();
                for (int k = 2; k <= P; k <<= 1) {
                    for (int jj = k >> 1; jj > 0; jj >>= 1) {
                        for (int i = tid; i < P; i += NT) {
                            const int l = i ^ jj;
                            if (l > i) {
                                const unsigned long long a = s_x[i], bq = s_x[l];
                                const bool up = (i & k) == 0;
                                if ((a > bq) == up) { s_x[i] = bq; s_x[l] = a; }
                            }
                        }
                        __syncthreads();
                    }
                }
                for (int t = tid; t < Mx; t += NT) {
                    const unsigned kidx = (unsigned)(s_x[t] >> 32);
                    if (t > 0 && (unsigned)(s_x[t - 1] >> 32) == kidx) continue;             // not the first row of its owner
                    float* acc = p.kacc + (img + s + kidx) * 5;
                    float ax = acc[0], ay = acc[1], az = acc[2], aw = acc[3], ws = acc[4];
                    for (int u = t; u < Mx && (unsigned)(s_x[u] >> 32) == kidx; ++u) {
                        const int m = (int)(unsigned)s_x[u];
                        const float4 bm = s_box[m];
                        const float wm = s_conf[m];
                        ax = __fadd_rn(ax, __fmul_rn(wm, bm.x));
                        ay = __fadd_rn(ay, __fmul_rn(wm, bm.y));
                        az = __fadd_rn(az, __fmul_rn(wm, bm.z));
                        aw = __fadd_rn(aw, __fmul_rn(wm, bm.w));
                        ws = __fadd_rn(ws, wm);
                    }
                    acc[0] = ax; acc[1] = ay; acc[2] = az; acc[3] = aw; acc[4] = ws;
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
                for (int j = tid; j < nc; j += NT) {
                    const bool is_keeper = s_pre[j] < 0 && ((s_keptw[wq][j >> 6] >> (j & 63)) & 1ull);
                    if (!is_keeper && s_own[j] == last_k) ++cnt;
                }
                if (cnt) atomicAdd(&s_last_members, cnt);
            }
        }
        Kprev += Kc;
        __syncthreads();   // kacc / kbox writes of this chunk are visible to the next one (same CTA)
        if (c0 == s) nms_stamp(tr, 5);
    }

    if (VARIANT == 0) {
        if (!single) {
            for (int kidx = tid; kidx < Kprev; kidx += NT) {
                const float* acc = p.kacc + (img + s + kidx) * 5;
                const uint32_t pos = p.kpos[img + s + kidx];
                const uint32_t r = p.srank[img + pos];
                const float ws = acc[4];
                p.kpay[img + r] = p.spay[img + pos];
                p.mbox[img + r] = make_float4(__fdiv_rn(acc[0], ws), __fdiv_rn(acc[1], ws), __fdiv_rn(acc[2], ws),
                                              __fdiv_rn(acc[3], ws));
                atomicAdd(&p.chunk_cnt[(size_t)b * p.n_chunks + (r >> kEmitShift)], 1u);
            }
        }
    } else {
        int K = Kprev;
        if (p.compat && K > 0 && s_last_members == 0) K -= 1;          // SSD.py:277-278
        if (tid == 0) p.out_count[b] = K;
        const uint32_t* tp = p.tile_prefix + (size_t)b * p.n_tiles;
        for (int kidx = tid; kidx < K; kidx += NT) {
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
                // ... i.e. row k of the output carries prior number f = (filtered index of the kept candidate): decode that
                // prior here (<= topk rows per image) instead of keeping dense box / label arrays of all B x P priors around
                const int f = (int)(tp[slot >> kTileShift] + (slot & (kTile - 1)));
                bx = prior_decode_box(*reinterpret_cast<const float4*>(p.loc + ((size_t)b * p.P + f) * 4),
                                      *reinterpret_cast<const float4*>(p.priors + (size_t)f * 4));
                label = prior_row_argmax(p.pcls + ((size_t)b * p.P + f) * p.C, p.C);
                src = f;
            }
            float* o = p.out_rows + ((size_t)b * p.topk + kidx) * 7;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = 0.0f; o[5] = score; o[6] = (float)label;
            if (p.out_index) p.out_index[(size_t)b * p.topk + kidx] = src;   // prior whose box/label the row carries
        }
    }
}

// K3b — ordered emit: compact the kept rows of every image in ascending score rank.  One CTA per 1024 ranks;
// its output base is the sum of the kept-row counts of the preceding chunks (accumulated by the NMS kernel).
struct EmitParams {
    const uint32_t* count;
    const uint32_t* kpay;
    const float4* mbox;
    const float2* cc2;
    const uint32_t* orig;
    const uint32_t* chunk_cnt;
    float* out_rows;       // [B][n_pad][7]
    int32_t* out_index;    // [B][n_pad] or null
    int32_t* out_count;    // [B]
    int n_pad, n_chunks;
    const int32_t* out_base;   // packed output: first output row of every image ([B+1], from yolo_emit_prefix_kernel), or null
};

// Packed output only: per-image kept-row totals (sum of the chunk counters the NMS kernel accumulated) and their exclusive
// prefix = the first output row of every image.  One CTA; a thread per image.
__global__ void __launch_bounds__(1024) yolo_emit_prefix_kernel(const uint32_t* __restrict__ count,
                                                                const uint32_t* __restrict__ chunk_cnt, int B, int n_chunks,
                                                                int32_t* __restrict__ out_base,
                                                                int32_t* __restrict__ early /* [2B+1] or null */) {
    __shared__ int s_scan[33];
    int carry = 0;
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int b = b0 + (int)threadIdx.x;
        int tot = 0;
        if (b < B) {
            const int used = min(n_chunks, ((int)count[b] + kEmitChunk - 1) >> kEmitShift);
            const uint32_t* cc = chunk_cnt + (size_t)b * n_chunks;
            for (int c = 0; c < used; ++c) tot += (int)cc[c];
        }
        int total;
        const int ex = block_exclusive_scan(tot, s_scan, &total);
        if (b < B) {
            if (out_base) out_base[b] = carry + ex;
            if (early) { early[b] = tot; early[B + b] = carry + ex; }
        }
        carry += total;
    }
    if (threadIdx.x == 0) {
        if (out_base) out_base[B] = carry;
        if (early) early[2 * B] = carry;
    }
}

constexpr int kEmitThreads = kEmitChunk / 4;   // 4 consecutive ranks per thread
constexpr int kEmitPerCta = 1;                 // chunks per CTA (4 per CTA measured slower: 49 vs 39 us - the kernel is
                                               // latency-bound per CTA, so more, smaller CTAs win)

__global__ void __launch_bounds__(kEmitThreads) yolo_emit_kernel(const EmitParams p) {
    __shared__ int s_scan[33];
    __shared__ __align__(16) float s_rows[kEmitChunk * 7 + 4];    // kept rows of one chunk, packed (+ alignment shift)
    __shared__ int s_idx[kEmitChunk];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int c_first = blockIdx.x * kEmitPerCta;
    const size_t img = (size_t)b * p.n_pad;
    const size_t oimg = p.out_base ? (size_t)p.out_base[b] : img;          // first output row of the image
    const int n = (int)p.count[b];
    if (c_first > 0 && (c_first << kEmitShift) >= n) return;
    const uint32_t* cc = p.chunk_cnt + (size_t)b * p.n_chunks;
    // output base = kept rows of the preceding chunks; every warp sums them on its own (no block barrier)
    const int upto = c_first == 0 ? p.n_chunks : c_first;          // the first CTA also totals the image
    int base = 0;
    for (int t = lane; t < upto; t += 32) base += (int)cc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xFFFFFFFFu, base, o);
    if (c_first == 0) {
        if (tid == 0) p.out_count[b] = base;
        base = 0;
    }
    for (int c = c_first; c < c_first + kEmitPerCta && (c << kEmitShift) < n; ++c) {
        const int r = (c << kEmitShift) + tid * 4;
        uint32_t pay[4] = {kNone, kNone, kNone, kNone};
        if (r + 3 < n) {
            const uint4 v = *reinterpret_cast<const uint4*>(p.kpay + img + r);
            pay[0] = v.x; pay[1] = v.y; pay[2] = v.z; pay[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (r + i < n) pay[i] = p.kpay[img + r + i];
        }
        // gathers issued before the scan so that their latency overlaps it
        float4 mb[4];
        float2 cf[4];
        int og[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (pay[i] != kNone) {
                const uint32_t slot = pay[i] & kSlotMask;
                mb[i] = p.mbox[img + r + i];
                cf[i] = p.cc2[img + slot];
                og[i] = p.out_index ? (int)p.orig[img + slot] : 0;
            }
        }
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) cnt += pay[i] != kNone ? 1 : 0;
        int total;
        int ex = block_exclusive_scan(cnt, s_scan, &total);     // trailing barrier also protects s_rows reuse
        // rows are staged with the same alignment (mod 4 floats) as their place in the output, so the copy below
        // moves whole aligned float4 words
        const size_t gfloat = (oimg + (size_t)base) * 7;
        const int sh = (int)(gfloat & 3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (pay[i] != kNone) {
                float* o = s_rows + sh + ex * 7;
                o[0] = mb[i].x; o[1] = mb[i].y; o[2] = mb[i].z; o[3] = mb[i].w;
                o[4] = cf[i].x; o[5] = cf[i].y; o[6] = (float)(pay[i] >> kSlotBits);     // YOLOV3.py:318-319
                if (p.out_index) s_idx[ex] = og[i];
                ++ex;
            }
        }
        __syncthreads();
        {
            const int nfl = total * 7;                            // floats to write
            float4* dst4 = reinterpret_cast<float4*>(p.out_rows + (gfloat - sh));
            const float4* src4 = reinterpret_cast<const float4*>(s_rows);
            const int nvec = (sh + nfl + 3) >> 2;
            for (int i = tid; i < nvec; i += kEmitThreads) {
                const float4 v = src4[i];
                const int f0 = i * 4 - sh;                        // index of v.x among the chunk's floats
                if (f0 >= 0 && f0 + 3 < nfl) {
                    dst4[i] = v;
                } else {
                    float* d = reinterpret_cast<float*>(dst4 + i);
                    if (f0 >= 0 && f0 < nfl) d[0] = v.x;
                    if (f0 + 1 >= 0 && f0 + 1 < nfl) d[1] = v.y;
                    if (f0 + 2 >= 0 && f0 + 2 < nfl) d[2] = v.z;
                    if (f0 + 3 >= 0 && f0 + 3 < nfl) d[3] = v.w;
                }
            }
        }
        if (p.out_index) {
            int32_t* di = p.out_index + oimg + base;
            for (int i = tid; i < total; i += kEmitThreads) di[i] = s_idx[i];
        }
        base += total;
    }
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
    p.kbox = w.kbox; p.kacc = w.kacc; p.kpos = w.kpos; p.chunk_cnt = w.chunk_cnt; p.n_chunks = w.n_chunks;
    p.n_pad = w.n_pad; p.C = w.C; p.thr = d->nms_thres;
    dim3 grid(d->num_classes, d->batch);
    // candidates per (image, class) on average, if everything survived: longer than a chunk -> the wide variant,
    // at most 160 -> the small one
    const int avg = w.N / d->num_classes;
    const char* pf = getenv("B200DET_NMS");           // "tab": table pre-filter + queued exact tests; "half2": pair-wise half2 test
    const bool tab = pf && strcmp(pf, "tab") == 0;
    if (d->nms_thres >= 0.0f && tab) {
        if (avg > kNmsT) nms_segment_kernel<0, true, 512, kNmsT, true><<<grid, 512, 0, st>>>(p);
        else if (avg <= 160) nms_segment_kernel<0, true, 128, 192, true><<<grid, 128, 0, st>>>(p);
        else nms_segment_kernel<0, true, kNmsThreads, kNmsT, true><<<grid, kNmsThreads, 0, st>>>(p);
    } else if (d->nms_thres >= 0.0f) {
        // wide variant (segments longer than a chunk: few classes, dense crowds, single images): 512-row chunks; 1024 threads
        // when there are not even enough segments for one CTA per SM.  Measured (profiles/README.md, round 2): crowd shard
        // 313 us (512 threads x 384 rows) -> 299 us (512 x 512), 360 us (1024 x 512); one 20-class image 99 -> 90 -> 80 us.
        if (avg > kNmsT && (long long)d->num_classes * d->batch <= 148) nms_segment_kernel<0, true, 1024, 512><<<grid, 1024, 0, st>>>(p);
        else if (avg > kNmsT) nms_segment_kernel<0, true, 512, 512><<<grid, 512, 0, st>>>(p);
        else if (avg <= 160) nms_segment_kernel<0, true, 128, 192><<<grid, 128, 0, st>>>(p);
        else nms_segment_kernel<0, true, kNmsThreads, kNmsT><<<grid, kNmsThreads, 0, st>>>(p);
    } else {
        nms_segment_kernel<0, false, kNmsThreads, kNmsT><<<grid, kNmsThreads, 0, st>>>(p);
    }
    B2_LAUNCH_CHECK("nms_segment_kernel<0>");
    return 0;
}

int yolo_stage_emit(const b200det_yolo_desc* d, void* ws, size_t ws_bytes, float* out_rows, int32_t* out_index,
                    int32_t* out_count, int32_t* out_offsets, int32_t* counts_early, cudaEvent_t counts_ready, cudaStream_t st) {
    int rc = yolo_validate(d, ws, ws_bytes);
    if (rc) return rc;
    B2_CHECK_ARG(out_rows && out_count, "out_rows / out_count is null");
    B2_CHECK_ARG(((uintptr_t)out_rows & 15) == 0, "out_rows must be 16-byte aligned");
    YoloWs w;
    yolo_ws_layout(d, ws, &w);
    if (out_offsets || counts_early) {
        yolo_emit_prefix_kernel<<<1, 1024, 0, st>>>(w.count, w.chunk_cnt, d->batch, w.n_chunks, out_offsets, counts_early);
        B2_LAUNCH_CHECK("yolo_emit_prefix_kernel");
    }
    if (counts_ready) B2_CUDA(cudaEventRecord(counts_ready, st));
    EmitParams p;
    p.out_base = out_offsets;
    p.count = w.count; p.kpay = w.kpay; p.mbox = w.mbox; p.cc2 = w.cc2; p.orig = w.orig; p.chunk_cnt = w.chunk_cnt;
    p.out_rows = out_rows; p.out_index = out_index; p.out_count = out_count; p.n_pad = w.n_pad; p.n_chunks = w.n_chunks;
    dim3 grid(ceil_div(w.n_chunks, kEmitPerCta), d->batch);
    yolo_emit_kernel<<<grid, kEmitThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("yolo_emit_kernel");
    return 0;
}

int prior_nms_launch_raw(const uint32_t* count, const uint32_t* spay, const float4* box4, const float2* cc2,
                         float4* kbox, uint32_t* kpos, int n_pad, float thr, int topk, int compat,
                         const uint32_t* tile_prefix, int n_tiles, const uint32_t* orig, const float* loc,
                         const float* cls, const float* priors, int P, int C, float* out_rows, int32_t* out_index, int32_t* out_count,
                         int batch, int mode_min, cudaStream_t st) {
    NmsParams p;
    memset(&p, 0, sizeof(p));
    p.count = count; p.spay = spay; p.box4 = box4; p.cc2 = cc2; p.kbox = kbox; p.kpos = kpos;
    p.n_pad = n_pad; p.thr = thr; p.topk = topk; p.compat = compat; p.tile_prefix = tile_prefix; p.n_tiles = n_tiles;
    p.orig = orig; p.loc = loc; p.pcls = cls; p.priors = priors; p.P = P; p.C = C;
    p.out_rows = out_rows; p.out_index = out_index; p.out_count = out_count;
    dim3 grid(1, batch);
    if (mode_min) nms_segment_kernel<2, false, kNmsThreads, kNmsT><<<grid, kNmsThreads, 0, st>>>(p);
    else nms_segment_kernel<1, false, kNmsThreads, kNmsT><<<grid, kNmsThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("nms_segment_kernel<prior>");
    return 0;
}

}  // namespace b200det

extern "C" int b200det_debug_set_nms_trace(void* dev_ptr) {
    unsigned long long* p = (unsigned long long*)dev_ptr;
    return (int)cudaMemcpyToSymbol(b200det::g_nms_trace, &p, sizeof(p));
}

#ifdef B200DET_NMS_DEBUG
extern "C" int b200det_debug_nms_counters(unsigned long long* out8_host, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out8_host, b200det::g_nms_dbg, 64);
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0};
        e = cudaMemcpyToSymbol(b200det::g_nms_dbg, z, 64);
    }
    return (int)e;
}
#endif
