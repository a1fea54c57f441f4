// M1 / M2 — detection metrics on the device: true-positive matching of the NMS output against
// the labels (`get_batch_statistics`, LightningFunc/accuracy.py:116-154) and per-class average precision
// (`ap_per_class` / `compute_ap`, accuracy.py:207-287).
//
// M1.  The reference walks the detections of an image in order and lets a detection claim the target with the highest
// +1-IoU when its label occurs among the image's labels, the IoU reaches the threshold and the target is unclaimed.
// Which target a detection aims at does not depend on the walk, so the walk reduces to "the first detection (in row
// order) aiming at a target with a passing IoU gets it": one pass computes the aim of every detection and an
// atomicMin of the row index per target, a second pass reads the winners.  (The reference's early exit once every
// target is claimed changes nothing — later rows could only aim at claimed targets.)
//
// M2.  Detections are ordered by descending confidence (ties by position) with the score sort of the NMS pipeline
// (one "image" = the whole test set, payload = position); their (class, true-positive) pairs are gathered into that
// order once, and one CTA per evaluated class streams the array forward (totals) and backward (precision envelope from
// the right, sum of recall steps) in fp64.
#include <limits.h>

#include "boxmath.cuh"
#include "yolo_ws.cuh"

namespace b200det {

int seg_scan_launch(const uint32_t* cls_hist, uint32_t* seg_off, int C, int batch, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------------------
// M1
// ------------------------------------------------------------------------------------------------------------------
struct BsWs {
    int* list;     // [B][nt]  target rows of image b, in the order of `targets`
    int* cnt;      // [B]
    int* first;    // [B][nt]  lowest detection row aiming at list entry m with a passing IoU
    size_t bytes;
};
static BsWs bs_layout(void* ws, int B, int nt) {
    BsWs w;
    const size_t n = (size_t)B * (size_t)(nt > 0 ? nt : 1);
    char* p = (char*)ws;
    size_t off = 0;
    w.list = (int*)(p + off); off = align_up(off + n * 4, 256);
    w.cnt = (int*)(p + off); off = align_up(off + (size_t)B * 4, 256);
    w.first = (int*)(p + off); off = align_up(off + n * 4, 256);
    w.bytes = off;
    return w;
}
size_t batch_statistics_ws_bytes(int B, int nt) { return bs_layout(nullptr, B, nt).bytes; }

// targets[:, 0] == b (accuracy.py:132), order preserved; also arms the claim slots
__global__ void __launch_bounds__(1024) bs_group_kernel(const float* __restrict__ targets, int nt, int* __restrict__ list,
                                                        int* __restrict__ cnt, int* __restrict__ first) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    int base = 0;
    for (int t0 = 0; t0 < nt; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        const int f = (t < nt && targets[(size_t)t * 6] == (float)b) ? 1 : 0;
        if (t < nt) first[(size_t)b * nt + t] = INT_MAX;
        int total;
        const int ex = block_exclusive_scan(f, s_scan, &total);
        if (f) list[(size_t)b * nt + base + ex] = t;
        base += total;
    }
    if (threadIdx.x == 0) cnt[b] = base;
}

// pass 1: the target a detection aims at (first maximal +1-IoU over ALL targets of the image, accuracy.py:149) or -1,
// written into the tp slot as an integer; atomicMin of the row index on the aimed-at target
__global__ void __launch_bounds__(256) bs_aim_kernel(const float* __restrict__ rows, const long long* __restrict__ row_start,
                                                     const int* __restrict__ count, const float* __restrict__ targets, int nt,
                                                     const int* __restrict__ list, const int* __restrict__ cnt,
                                                     int* __restrict__ first, float thr, float* __restrict__ tp) {
    __shared__ float4 s_box[256];
    __shared__ float s_lab[256];
    __shared__ float s_area[256];
    const int b = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    const int K = count[b];
    if (blockIdx.x * 256 >= K) return;
    const int M = cnt[b];
    const bool live = k < K;
    const long long r0 = row_start[b];
    float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
    float pl = 0.f;
    if (live) {
        const float* r = rows + (r0 + k) * 7;
        pb = make_float4(r[0], r[1], r[2], r[3]);
        pl = r[6];                                                          // output[:, -1]  (accuracy.py:128)
    }
    const float pa = box_area_plus1(pb);
    bool label_ok = false, have = false;
    float best = 0.f;
    int bi = -1;
    for (int m0 = 0; m0 < M; m0 += 256) {
        const int mm = min(256, M - m0);
        __syncthreads();
        if ((int)threadIdx.x < mm) {
            const float* t = targets + (size_t)list[(size_t)b * nt + m0 + threadIdx.x] * 6;
            s_lab[threadIdx.x] = t[1];
            s_box[threadIdx.x] = make_float4(t[2], t[3], t[4], t[5]);      // corners, as bbox_iou's default reads them
            s_area[threadIdx.x] = box_area_plus1(s_box[threadIdx.x]);
        }
        __syncthreads();
        for (int m = 0; m < mm; ++m) {
            label_ok |= s_lab[m] == pl;                                     // `pred_label not in target_labels` (:146)
            const float v = iou_plus1_eps_pre(pb, s_box[m], pa, s_area[m]);  // most pairs do not touch: no division
            if (!have || (!(v <= best) && (best == best))) { best = v; bi = m0 + m; have = true; }   // torch.max(0)
        }
    }
    if (!live) return;
    const bool aim = label_ok && have && best >= thr;                       // :150 (NaN fails the comparison)
    if (aim) atomicMin(&first[(size_t)b * nt + bi], k);
    tp[r0 + k] = __int_as_float(aim ? bi : -1);
}

// pass 2: a detection is a true positive iff it is the first one aiming at its target (:150-152)
__global__ void __launch_bounds__(256) bs_claim_kernel(const long long* __restrict__ row_start, const int* __restrict__ count,
                                                       int nt, const int* __restrict__ first, float* __restrict__ tp) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= count[b]) return;
    float* slot = tp + row_start[b] + k;
    const int aim = __float_as_int(*slot);
    *slot = (aim >= 0 && first[(size_t)b * nt + aim] == k) ? 1.0f : 0.0f;
}

int batch_statistics_launch(const float* rows, const long long* row_start, const int* count, int B, int max_count,
                            const float* targets, int nt, float thr, void* ws, float* tp, cudaStream_t st) {
    BsWs w = bs_layout(ws, B, nt);
    bs_group_kernel<<<B, 1024, 0, st>>>(targets, nt, w.list, w.cnt, w.first);
    B2_LAUNCH_CHECK("bs_group_kernel");
    if (max_count <= 0) return 0;
    dim3 grid(ceil_div(max_count, 256), B);
    bs_aim_kernel<<<grid, 256, 0, st>>>(rows, row_start, count, targets, nt, w.list, w.cnt, w.first, thr, tp);
    B2_LAUNCH_CHECK("bs_aim_kernel");
    bs_claim_kernel<<<grid, 256, 0, st>>>(row_start, count, nt, w.first, tp);
    B2_LAUNCH_CHECK("bs_claim_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// M2
// ------------------------------------------------------------------------------------------------------------------
int score_sort_launch(const uint32_t* tile_count, const uint32_t* count, uint32_t* digit_hist, uint32_t* ticket,
                      uint32_t* status, uint32_t* key[2], uint32_t* pay[2], int n_pad, int n_tiles, int batch,
                      cudaStream_t st);
bool score_sort_is_lookback(int n_pad);
int partition_pass_launch(const uint32_t* count, uint32_t* digit_hist, uint32_t* ticket, uint32_t* status,
                          const uint32_t* pay_in, uint32_t* pay_out, int n_pad, int n_tiles, int pass, int shift,
                          cudaStream_t st);

struct ApWs {
    uint32_t* count;        // [1]
    uint32_t* digit_hist;   // [kMaxPasses][256]   (look-back sort only)
    uint32_t* ticket;       // [kMaxPasses]
    size_t zero_bytes;
    uint32_t* status;       // [kMaxPasses][sort_tiles][256]
    uint32_t* tile_count;   // [n_tiles]
    uint32_t* key[2];
    uint32_t* pay[2];       // payload = position of the detection in the caller's arrays
    int32_t* packed;        // [n_pad] by confidence rank: class << 1 | true positive, or -1 (class not evaluable)
    int n_pad, n_tiles, sort_tiles;
    size_t bytes;
};
static ApWs ap_layout(void* base, int n) {
    ApWs w;
    const size_t P = align_up((size_t)(n > 0 ? n : 1), kTile);
    w.n_pad = (int)P; w.n_tiles = (int)(P / kTile); w.sort_tiles = (int)((P + kSortTile - 1) / kSortTile);
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes, 256); return r; };
    w.count = (uint32_t*)take(4);
    w.digit_hist = (uint32_t*)take((size_t)kMaxPasses * 256 * 4);
    w.ticket = (uint32_t*)take((size_t)kMaxPasses * 4);
    w.zero_bytes = off;
    w.status = (uint32_t*)take((size_t)kMaxPasses * w.sort_tiles * 256 * 4);
    w.tile_count = (uint32_t*)take((size_t)w.n_tiles * 4);
    for (int i = 0; i < 2; ++i) w.key[i] = (uint32_t*)take(P * 4);
    for (int i = 0; i < 2; ++i) w.pay[i] = (uint32_t*)take(P * 4);
    w.packed = (int32_t*)take(P * 4);
    w.bytes = off;
    return w;
}
size_t ap_per_class_ws_bytes(int n) { return ap_layout(nullptr, n).bytes; }

// sort input: key = descending-confidence key (np.argsort(-conf), accuracy.py:221), payload = position; dense tiles
__global__ void __launch_bounds__(256) ap_prepare_kernel(const float* __restrict__ conf, int n, uint32_t* __restrict__ key,
                                                         uint32_t* __restrict__ pay, uint32_t* __restrict__ tile_count,
                                                         uint32_t* __restrict__ count, int n_tiles) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_tiles) tile_count[i] = (uint32_t)max(0, min(kTile, n - i * kTile));
    if (i == 0) count[0] = (uint32_t)n;
    if (i >= n) return;
    key[i] = score_sort_key(conf[i]);
    pay[i] = (uint32_t)i;
}

// by confidence rank: class and true-positive flag of the detection, one word (streamed by every class CTA)
__global__ void __launch_bounds__(256) ap_gather_kernel(const uint32_t* __restrict__ spay, const float* __restrict__ tp,
                                                        const float* __restrict__ pred_cls, int n, int32_t* __restrict__ packed) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = spay[s];
    const float c = pred_cls[i];
    int32_t v = -1;
    if (c >= 0.0f && c < 1073741824.0f && c == floorf(c)) v = ((int32_t)c << 1) | (tp[i] != 0.0f ? 1 : 0);
    packed[s] = v;
}

// The same gather for the partitioned route (<= 255 evaluated classes): the word is (u << 8 | true positive) with u the
// FIRST index of the detection's class in `classes` (255: not evaluated), and the bin totals of u are counted on the way —
// one more stable radix pass on u then leaves every evaluated class as one contiguous, confidence-ordered run.
__global__ void __launch_bounds__(256) ap_gather_bins_kernel(const uint32_t* __restrict__ spay, const float* __restrict__ tp,
                                                             const float* __restrict__ pred_cls, int n,
                                                             const int* __restrict__ classes, int num_classes,
                                                             uint32_t* __restrict__ binned, uint32_t* __restrict__ bin_hist) {
    __shared__ int s_cls[256];
    __shared__ int s_hist[256];
    s_hist[threadIdx.x] = 0;
    if ((int)threadIdx.x < num_classes) s_cls[threadIdx.x] = classes[threadIdx.x];
    __syncthreads();
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s < n) {
        const uint32_t i = spay[s];
        const float c = pred_cls[i];
        int u = 255;
        if (c >= 0.0f && c < 1073741824.0f && c == floorf(c)) {
            const int ci = (int)c;
            for (int j = 0; j < num_classes; ++j)
                if (s_cls[j] == ci) { u = j; break; }
        }
        binned[s] = ((uint32_t)u << 8) | (tp[i] != 0.0f ? 1u : 0u);
        atomicAdd(&s_hist[u], 1);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&bin_hist[threadIdx.x], (uint32_t)s_hist[threadIdx.x]);
}

constexpr int kApThreads = 1024;
constexpr int kApRows = 8;               // rows per thread and chunk in the backward pass

// inclusive block scan of two counters packed in one 64-bit word (hi = rows of the class, lo = true positives)
__device__ __forceinline__ unsigned long long block_inclusive_add64(unsigned long long v, unsigned long long* s_warp /*[32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    unsigned long long add = 0ull;
    for (int w = 0; w < warp; ++w) add += s_warp[w];
    return v + add;
}
// maximum over the threads AFTER me (0 when there is none; the values are >= 0)
__device__ __forceinline__ double block_suffix_max_excl(double v, double* s_warp /*[32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_down_sync(0xFFFFFFFFu, v, o);
        if (lane + o < 32) v = fmax(v, t);
    }
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    double ex = __shfl_down_sync(0xFFFFFFFFu, v, 1);
    if (lane == 31) ex = 0.0;
    for (int w = warp + 1; w < kApThreads / 32; ++w) ex = fmax(ex, s_warp[w]);
    return ex;
}

// One CTA per evaluated class u (classes[u] from np.unique(target_cls), n_gt[u] its label count).  The CTA streams the
// whole rank-ordered array twice: forward for the totals (rows of the class n_p, true positives), backward for the
// precision envelope from the right (accuracy.py:277-278) and the sum of recall steps times it (:282-285).  In the
// backward pass the running counts follow from the totals, so nothing is stored per row.
// SEG: the array is partitioned by class (ap_gather_bins_kernel + one radix pass): the CTA walks only its own run, found
// from the bin totals; rows are (u << 8 | tp) and all belong to the class.
template <bool SEG>
__global__ void __launch_bounds__(kApThreads) ap_class_kernel(const int32_t* __restrict__ packed, int n,
                                                              const int* __restrict__ classes, const int* __restrict__ n_gt,
                                                              const uint32_t* __restrict__ bin_hist,
                                                              double* __restrict__ out_p, double* __restrict__ out_r,
                                                              double* __restrict__ out_ap, double* __restrict__ out_f1) {
    __shared__ unsigned long long s_w[33];
    __shared__ double s_d[32];
    __shared__ double s_red[32];
    const int u = blockIdx.x, tid = threadIdx.x;
    int c = classes[u];
    if (SEG) {
        int u0 = u;                                       // a repeated class shares the run of its first occurrence
        for (int j = u - 1; j >= 0; --j)
            if (classes[j] == c) u0 = j;
        int start = 0;
        for (int j = 0; j < u0; ++j) start += (int)bin_hist[j];
        packed += start;
        n = (int)bin_hist[u0];
        c = u0 << 7;                                      // (row >> 1) of (u0 << 8 | tp)
    }
    const double denom = (double)n_gt[u] + 1e-16;                                  // accuracy.py:247
    // forward: totals
    unsigned long long tot = 0ull;
    for (int k = tid; k < n; k += kApThreads) {
        const int32_t v = packed[k];
        if (v >= 0 && (v >> 1) == c) tot += (1ull << 32) | (unsigned long long)(v & 1);
    }
    tot = block_inclusive_add64(tot, s_w);
    __syncthreads();
    if (tid == kApThreads - 1) s_w[32] = tot;
    __syncthreads();
    tot = s_w[32];
    const long long n_p = (long long)(tot >> 32), tp_total = (long long)(tot & 0xFFFFFFFFull);
    if (n_p == 0 || n_gt[u] == 0) {                                                // :238-241
        if (tid == 0) { out_p[u] = 0.0; out_r[u] = 0.0; out_ap[u] = 0.0; out_f1[u] = 0.0; }
        return;
    }
    // backward over chunks of kApThreads x kApRows rows; a thread owns kApRows consecutive rows
    double env = 0.0, acc = 0.0;
    long long after_rows = 0, after_tp = 0;          // rows / true positives of the class in the chunks already done
    const int chunk_rows = kApThreads * kApRows;
    const int nchunks = (n + chunk_rows - 1) / chunk_rows;
    for (int ch = nchunks - 1; ch >= 0; --ch) {
        const int k0 = ch * chunk_rows + tid * kApRows;
        int32_t v[kApRows];
        unsigned long long mine_cnt = 0ull;
#pragma unroll
        for (int j = 0; j < kApRows; ++j) {
            v[j] = (k0 + j) < n ? packed[k0 + j] : -1;
            if (v[j] >= 0 && (v[j] >> 1) == c) mine_cnt += (1ull << 32) | (unsigned long long)(v[j] & 1);
            else v[j] = -1;
        }
        const unsigned long long inc = block_inclusive_add64(mine_cnt, s_w);
        __syncthreads();
        if (tid == kApThreads - 1) s_w[32] = inc;
        __syncthreads();
        const unsigned long long chunk = s_w[32];
        // counts up to and including my last row = totals - (everything after it)
        const long long rows_end = n_p - after_rows - (long long)(chunk >> 32) + (long long)(inc >> 32);
        const long long tp_end = tp_total - after_tp - (long long)(chunk & 0xFFFFFFFFull) + (long long)(inc & 0xFFFFFFFFull);
        double prec[kApRows];
        double tmax = 0.0;
        {
            long long r = rows_end, t = tp_end;
#pragma unroll
            for (int j = kApRows - 1; j >= 0; --j) {
                prec[j] = 0.0;
                if (v[j] >= 0) {
                    prec[j] = (double)t / (double)r;                                  // tpc / (tpc + fpc)  (:251)
                    tmax = fmax(tmax, prec[j]);
                    r -= 1; t -= (v[j] & 1);
                }
            }
        }
        double e = fmax(block_suffix_max_excl(tmax, s_d), env);                      // envelope right after my rows
        {
            long long t = tp_end;
#pragma unroll
            for (int j = kApRows - 1; j >= 0; --j) {
                if (v[j] >= 0) {
                    e = fmax(e, prec[j]);
                    if (v[j] & 1) {                                                   // recall moves at true positives
                        acc += ((double)t / denom - (double)(t - 1) / denom) * e;
                        t -= 1;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_d[0] = e;                                                    // envelope at the chunk's first row
        __syncthreads();
        env = s_d[0];
        after_rows += (long long)(chunk >> 32);
        after_tp += (long long)(chunk & 0xFFFFFFFFull);
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double ap = 0.0;
        for (int w = 0; w < kApThreads / 32; ++w) ap += s_red[w];
        const double r = (double)tp_total / denom;                                 // recall_curve[-1]  (:248)
        const double p = (double)tp_total / (double)n_p;                           // precision_curve[-1] (:252)
        out_p[u] = p; out_r[u] = r; out_ap[u] = ap;
        out_f1[u] = 2.0 * p * r / (p + r + 1e-16);                                 // :259
    }
}

int ap_per_class_launch(const float* tp, const float* conf, const float* pred_cls, int n, const int* classes,
                        const int* n_gt, int num_classes, void* ws, double* out_p, double* out_r, double* out_ap,
                        double* out_f1, cudaStream_t st) {
    ApWs w = ap_layout(ws, n);
    int rc = zero_fill_launch(w.count, w.zero_bytes, st);
    if (rc) return rc;
    const int items = n > w.n_tiles ? n : w.n_tiles;
    ap_prepare_kernel<<<ceil_div(items, 256), 256, 0, st>>>(conf, n, w.key[0], w.pay[0], w.tile_count, w.count, w.n_tiles);
    B2_LAUNCH_CHECK("ap_prepare_kernel");
    rc = score_sort_launch(w.tile_count, w.count, w.digit_hist, w.ticket, w.status, w.key, w.pay, w.n_pad, w.n_tiles, 1, st);
    if (rc) return rc;
    if (n > 0 && num_classes > 0 && num_classes <= 255 && score_sort_is_lookback(w.n_pad)) {
        // large inputs: partition the confidence-ordered rows by class so that a class CTA walks ~n / classes rows, not n
        uint32_t* bin_hist = w.digit_hist + (size_t)kScorePasses * 256;            // zeroed above, unused by the score passes
        ap_gather_bins_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w.pay[0], tp, pred_cls, n, classes, num_classes, w.pay[1], bin_hist);
        B2_LAUNCH_CHECK("ap_gather_bins_kernel");
        rc = partition_pass_launch(w.count, w.digit_hist, w.ticket, w.status, w.pay[1], (uint32_t*)w.packed, w.n_pad, w.n_tiles,
                                   kScorePasses, 8, st);
        if (rc) return rc;
        ap_class_kernel<true><<<num_classes, kApThreads, 0, st>>>(w.packed, n, classes, n_gt, bin_hist, out_p, out_r, out_ap, out_f1);
        B2_LAUNCH_CHECK("ap_class_kernel<seg>");
        return 0;
    }
    if (n > 0) {
        ap_gather_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w.pay[0], tp, pred_cls, n, w.packed);
        B2_LAUNCH_CHECK("ap_gather_kernel");
    }
    if (num_classes > 0) {
        ap_class_kernel<false><<<num_classes, kApThreads, 0, st>>>(w.packed, n, classes, n_gt, nullptr, out_p, out_r, out_ap, out_f1);
        B2_LAUNCH_CHECK("ap_class_kernel");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// M3 — get_yolo_statistics for one level (LightningFunc/accuracy.py:382-470, the YOLOv2..v4 test step, step.py:99):
// D1 decode of the planar head, build_targets on the decoded map, six scalar metrics, and the decoded map itself.
// ------------------------------------------------------------------------------------------------------------------
int decode_box_launch(const float*, int, int, int, int, int, const float*, float, float*, cudaStream_t);
int decode_box_aux_launch(const float*, int, int, int, int, const float*, float, float*, float*, cudaStream_t);
bool decode_box_tileable(int, int, const void*, const void*);
size_t build_targets_ws_bytes(int, int, int, int);
int build_targets_launch_ex(const float*, int, const float*, int, const float*, const float*, int, int, int, int, int, float,
                            void*, float*, float*, uint8_t*, uint8_t*, float*, float*, float*, float*, float*, int32_t*,
                            cudaStream_t);

struct YsWs {
    double* acc;          // [8] n_obj, n_noobj, sum class_mask[obj], sum conf[obj], sum conf[noobj], sum conf50, sum iou50*det, sum iou75*det
    int32_t* status;      // build_targets guard bits
    float* iou_scores;    // [cells]
    float* class_mask;    // [cells]
    uint8_t* obj;         // [cells]
    uint8_t* noobj;       // [cells]
    float* aux;           // [cells][5] grid-unit box + objectness (tileable planes only)
    void* bt_ws;
    size_t bytes;
};
static YsWs ys_layout(void* base, int B, int A, int G, int nt) {
    YsWs w;
    const size_t cells = (size_t)B * A * G * G;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes, 256); return r; };
    w.acc = (double*)take(8 * 8);
    w.status = (int32_t*)take(4);
    w.iou_scores = (float*)take(cells * 4);
    w.class_mask = (float*)take(cells * 4);
    w.obj = (uint8_t*)take(cells);
    w.noobj = (uint8_t*)take(cells);
    w.aux = (float*)take(cells * 5 * 4);
    w.bt_ws = take(build_targets_ws_bytes(B, A, G, nt));
    w.bytes = off;
    return w;
}
size_t yolo_statistics_ws_bytes(int B, int A, int G, int nt) { return ys_layout(nullptr, B, A, G, nt).bytes; }

// One pass over the cells: the eight sums of accuracy.py:447-457, and the boxes of the decoded rows scaled from grid
// units to pixels (`pred_boxes.view(...) * self.stride`, :461) now that build_targets has read them.
__global__ void __launch_bounds__(256) ys_reduce_kernel(float* __restrict__ rows, int F, long long cells, float stride,
                                                        const float* __restrict__ iou_scores, const float* __restrict__ class_mask,
                                                        const uint8_t* __restrict__ obj, const uint8_t* __restrict__ noobj,
                                                        double* __restrict__ acc) {
    __shared__ double s_red[8][8];
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long c = (long long)blockIdx.x * 256 + threadIdx.x; c < cells; c += (long long)gridDim.x * 256) {
        float* r = rows + c * F;
        const float conf = r[4];
        r[0] = __fmul_rn(r[0], stride); r[1] = __fmul_rn(r[1], stride);
        r[2] = __fmul_rn(r[2], stride); r[3] = __fmul_rn(r[3], stride);
        const bool o = obj[c] != 0, no = noobj[c] != 0;
        const float cm = class_mask[c], iou = iou_scores[c];
        const float conf50 = conf > 0.5f ? 1.0f : 0.0f;                       // :450
        const float det = conf50 * cm * (o ? 1.0f : 0.0f);                    // :453 (tconf = obj_mask.float())
        if (o) { a[0] += 1.0; a[2] += cm; a[3] += conf; }
        if (no) { a[1] += 1.0; a[4] += conf; }
        a[5] += conf50;
        a[6] += (iou > 0.5f ? 1.0f : 0.0f) * det;                             // :451, :454-455
        a[7] += (iou > 0.75f ? 1.0f : 0.0f) * det;                            // :452, :456
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double v = a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_red[threadIdx.x][w];
        atomicAdd(&acc[threadIdx.x], v);
    }
}

// The same sums when the decode left a compact side table (aux[c] = grid box + objectness): nothing wide is touched.
__global__ void __launch_bounds__(256) ys_reduce_aux_kernel(const float* __restrict__ aux, long long cells,
                                                            const float* __restrict__ iou_scores, const float* __restrict__ class_mask,
                                                            const uint8_t* __restrict__ obj, const uint8_t* __restrict__ noobj,
                                                            double* __restrict__ acc) {
    __shared__ double s_red[8][8];
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long c = (long long)blockIdx.x * 256 + threadIdx.x; c < cells; c += (long long)gridDim.x * 256) {
        const float conf = aux[c * 5 + 4];
        const bool o = obj[c] != 0, no = noobj[c] != 0;
        const float cm = class_mask[c], iou = iou_scores[c];
        const float conf50 = conf > 0.5f ? 1.0f : 0.0f;
        const float det = conf50 * cm * (o ? 1.0f : 0.0f);
        if (o) { a[0] += 1.0; a[2] += cm; a[3] += conf; }
        if (no) { a[1] += 1.0; a[4] += conf; }
        a[5] += conf50;
        a[6] += (iou > 0.5f ? 1.0f : 0.0f) * det;
        a[7] += (iou > 0.75f ? 1.0f : 0.0f) * det;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double v = a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_red[threadIdx.x][w];
        atomicAdd(&acc[threadIdx.x], v);
    }
}

// metrics[6] = cls_acc, recall50, recall75, precision, conf_obj, conf_noobj  (the order of batch_metrics, :468)
__global__ void ys_final_kernel(const double* __restrict__ acc, float* __restrict__ metrics) {
    const float n_obj = (float)acc[0];
    metrics[0] = 100.0f * (float)(acc[2] / acc[0]);                           // mean over an empty selection is NaN, as in torch
    metrics[1] = (float)acc[6] / (n_obj + 1e-16f);
    metrics[2] = (float)acc[7] / (n_obj + 1e-16f);
    metrics[3] = (float)acc[6] / ((float)acc[5] + 1e-16f);
    metrics[4] = (float)(acc[3] / acc[0]);
    metrics[5] = (float)(acc[4] / acc[1]);
}

int yolo_statistics_launch(const float* head, int B, int A, int C, int G, const float* scaled_anchors, float stride,
                           const float* target, int nt, float ignore_thres, void* ws, float* out_rows, float* metrics,
                           cudaStream_t st) {
    YsWs w = ys_layout(ws, B, A, G, nt);
    const int F = 5 + C;
    const long long cells = (long long)B * A * G * G;
    const int grid = (int)((cells + 255) / 256 < 148 * 8 ? (cells + 255) / 256 : 148 * 8);
    if (decode_box_tileable(G, F, head, out_rows)) {
        // one pass writes the final rows (pixels) and the grid-unit boxes + objectness build_targets and the sums read
        int rc = decode_box_aux_launch(head, B, A, C, G, scaled_anchors, stride, out_rows, w.aux, st);
        if (rc) return rc;
        rc = build_targets_launch_ex(w.aux, 5, out_rows + 5, F, target, scaled_anchors, B, A, G, C, nt, ignore_thres, w.bt_ws,
                                     w.iou_scores, w.class_mask, w.obj, w.noobj, nullptr, nullptr, nullptr, nullptr, nullptr,
                                     w.status, st);
        if (rc) return rc;
        B2_CUDA(cudaMemsetAsync(w.acc, 0, 64, st));
        ys_reduce_aux_kernel<<<grid, 256, 0, st>>>(w.aux, cells, w.iou_scores, w.class_mask, w.obj, w.noobj, w.acc);
        B2_LAUNCH_CHECK("ys_reduce_aux_kernel");
        ys_final_kernel<<<1, 1, 0, st>>>(w.acc, metrics);
        B2_LAUNCH_CHECK("ys_final_kernel");
        return 0;
    }
    // odd planes (13 x 13): decoded map in GRID units first (stride 1): x = sigma + gx, w = exp * anchor  (:412-435)
    int rc = decode_box_launch(head, B, A, C, G, B200DET_DECODE_YOLO_EXP, scaled_anchors, 1.0f, out_rows, st);
    if (rc) return rc;
    rc = build_targets_launch_ex(out_rows, F, out_rows + 5, F, target, scaled_anchors, B, A, G, C, nt, ignore_thres, w.bt_ws,
                                 w.iou_scores, w.class_mask, w.obj, w.noobj, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 w.status, st);
    if (rc) return rc;
    B2_CUDA(cudaMemsetAsync(w.acc, 0, 64, st));
    ys_reduce_kernel<<<grid, 256, 0, st>>>(out_rows, F, cells, stride, w.iou_scores, w.class_mask, w.obj, w.noobj, w.acc);
    B2_LAUNCH_CHECK("ys_reduce_kernel");
    ys_final_kernel<<<1, 1, 0, st>>>(w.acc, metrics);
    B2_LAUNCH_CHECK("ys_final_kernel");
    return 0;
}

}  // namespace b200det
