// K2 — per-image (segmented) LSD radix sort of the surviving candidates.
//
// Order produced: (class ascending, score descending, candidate index ascending) — i.e. the score
// order of `(-score).argsort()` (model/YOLOV3.py:317, ties by ascending candidate index = stable),
// then a stable partition by class so that every (image, class) is one contiguous segment for the
// class-aware NMS (label_match, YOLOV3.py:324).  The position after the four score passes IS the
// global score rank; it is carried through the class pass(es) so kept rows can be emitted in the
// reference's descending-score order without a second sort.
//
// Layout: one CTA sorts one 4096-key tile of one image per pass.  The cross-tile prefix of a digit is
// obtained by decoupled look-back over the preceding tiles OF THE SAME IMAGE (one 32-bit status word per
// (tile, digit): 2 flag bits + 30-bit count).  Tiles are handed out by a per-(pass, image) ticket counter, so
// every tile a CTA waits for belongs to a CTA that has already started — no forward-progress assumption on
// the block scheduler.  The result is independent of timing (counts are integers), i.e. bit-reproducible.
// Per-image digit totals of all passes come from one up-front histogram kernel (totals are
// permutation-invariant).  In-tile ranking is the stable warp-match ranking (8 ballots per digit + per-warp
// digit counters), 8 bits per pass.
//
// This multi-launch sort is the path for images with more slots than one thread-block cluster holds (RetinaNet-sized
// inputs are handled by the radix select of topk.cu instead) and the A/B reference of the one-launch cluster sort in
// clustersort.cu, which is what normally runs (B200DET_SORT=global forces this one).
#include <stdlib.h>

#include "yolo_ws.cuh"

namespace b200det {

int seg_scan_launch(const uint32_t* cls_hist, uint32_t* seg_off, int C, int batch, cudaStream_t st);
int cluster_sort_capacity();
int cluster_sort_launch(const uint32_t* tile_count, uint32_t* count, bool count_from_tiles, uint32_t* chunk_cnt,
                        int n_chunks, uint32_t* seg_off, uint32_t* key[2], uint32_t* pay[2], uint32_t* rank[2], int n_pad,
                        int n_tiles, int C, int n_cls_passes, int batch, cudaStream_t st);

// The one-launch cluster sort (clustersort.cu) is used whenever an image's slots fit one cluster; B200DET_SORT=global
// forces the multi-launch sort below (kept for larger images, and as the A/B reference of the tests).
static bool use_cluster_sort(int n_pad) {
    const char* e = getenv("B200DET_SORT");
    if (e && strcmp(e, "global") == 0) return false;
    return n_pad <= cluster_sort_capacity();
}

int cluster_sort_dense_capacity();
int cluster_sort_dense_launch(const uint32_t* tile_count, uint32_t* count, uint32_t* seg_off, const uint32_t* dense_key,
                              const uint32_t* dense_pay, const uint32_t* overflow, uint32_t* key[2], uint32_t* pay[2],
                              uint32_t* rank[2], int n_pad, int n_tiles, int C, int n_cls_passes, int batch, cudaStream_t st);

// Images with more slots than one cluster sorts (1280-pixel heads: 101 376 slots) usually keep a small fraction of them
// (conf_thres 0.001: ~10 k).  Those take the DENSE route: a compaction kernel moves the survivors to [0, count) (stable, via the
// prefix of the tile counts), a cluster of 4 sorts them in one launch, and the multi-launch sort below only runs — gated by a
// per-image flag, its kernels return at once otherwise — for images with more survivors than the cluster holds.
// B200DET_SORT=global forces the plain multi-launch sort (the A/B reference).
static bool use_dense_sort(int n_pad, int n_cls_passes) {
    const char* e = getenv("B200DET_SORT");
    if (e && strcmp(e, "global") == 0) return false;
    return n_pad > cluster_sort_capacity() && n_cls_passes == 1;
}

// True when the sort stage derives count / class offsets / zeroed counters itself, so that the reset stage launches nothing and
// the decode kernel skips its global counter atomics.
bool yolo_fast_path(const YoloWs& w) {
    return (use_cluster_sort(w.n_pad) && w.n_cls_passes == 1) || use_dense_sort(w.n_pad, w.n_cls_passes);
}

struct SortParams {
    const uint32_t* tile_count;  // [B][n_tiles] (first pass: tile-sparse input), else unused
    const uint32_t* count;       // [B]
    uint32_t* digit_hist;        // [B][kMaxPasses][256]
    uint32_t* ticket;            // [kMaxPasses][B]
    uint32_t* status;            // [kMaxPasses][B][sort_tiles][256]
    int sort_tiles, B;
    const uint32_t* key_in;
    const uint32_t* pay_in;
    const uint32_t* rank_in;
    uint32_t* key_out;
    uint32_t* pay_out;
    uint32_t* rank_out;
    int n_pad, n_tiles;
    int pass;                    // index into digit_hist
    int shift;                   // bit offset of the digit
    int n_cls_passes;
    const uint32_t* only_flagged; // [B] or null: when set, images whose word is 0 are skipped (dense route, see use_dense_sort)
    int dense_hist;              // the histogram kernel reads dense [0, count) input instead of the tile-sparse layout
};

__device__ __forceinline__ bool sparse_valid(const uint32_t* tile_count_img, int e) {
    return (uint32_t)(e & (kTile - 1)) < tile_count_img[e >> kTileShift];
}

// Lanes of the warp holding the same 8-bit digit (all 32 lanes must call).  Built from 8 ballots: the
// hardware match.any instruction measured ~64 issue cycles per warp on B200, these are plain votes.
__device__ __forceinline__ unsigned match_digit8(uint32_t digit, bool valid) {
    unsigned peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool one = (digit >> bit) & 1u;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, one);
        peers &= one ? bal : ~bal;
    }
    return peers;
}

// Up-front per-image digit totals for every pass (score bytes 0..3 from the key, class digits from
// the payload).  Input is the tile-sparse K1 output.
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const SortParams p) {
    __shared__ int h[kMaxPasses][256];
    const int b = blockIdx.y;
    const int e0 = blockIdx.x * kSortTile;
    if (e0 >= p.n_pad) return;
    if (p.only_flagged && p.only_flagged[b] == 0u) return;
    const int dense_n = p.dense_hist ? (int)p.count[b] : 0;
    const uint32_t* tc = p.tile_count + (size_t)b * p.n_tiles;
    for (int i = threadIdx.x; i < kMaxPasses * 256; i += kSortThreads) (&h[0][0])[i] = 0;
    for (int ps = 0; ps < kMaxPasses; ++ps)      // look-back words of this (image, tile), all passes
        p.status[(((size_t)ps * p.B + b) * p.sort_tiles + blockIdx.x) * 256 + threadIdx.x] = 0u;
    __syncthreads();
    const size_t img = (size_t)b * p.n_pad;
    const int npass = kScorePasses + p.n_cls_passes;
#pragma unroll 4
    for (int k = 0; k < kSortItems; ++k) {
        const int e = e0 + k * kSortThreads + threadIdx.x;
        const bool valid = p.dense_hist ? e < dense_n : (e < p.n_pad && sparse_valid(tc, e));
        uint32_t key = 0, pay = 0;
        if (valid) { key = p.key_in[img + e]; pay = p.pay_in[img + e]; }
        // plain shared-memory atomics: measured 14 us vs 51 us for ballot-aggregated increments on B200
        if (valid) {
#pragma unroll
            for (int s = 0; s < kScorePasses; ++s) atomicAdd(&h[s][(key >> (8 * s)) & 0xFFu], 1);
            atomicAdd(&h[kScorePasses][(pay >> kSlotBits) & 0xFFu], 1);
            if (npass > kScorePasses + 1) atomicAdd(&h[kScorePasses + 1][(pay >> (kSlotBits + 8)) & 0xFFu], 1);
        }

    }
    __syncthreads();
    uint32_t* g = p.digit_hist + (size_t)b * kMaxPasses * 256;
    for (int i = threadIdx.x; i < npass * 256; i += kSortThreads) {
        int v = (&h[0][0])[i];
        if (v) atomicAdd(&g[i], (uint32_t)v);
    }
}

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr uint32_t kFlagAgg = 1u << 30, kFlagPrefix = 2u << 30, kValMask = (1u << 30) - 1u;

// One radix pass.  FIRST: input is tile-sparse (validity from tile_count), else dense [0, count).
// SRC: 0 = digit from key, 1 = digit from payload.  RANK: 0 none, 1 = write input position as rank,
// 2 = carry rank_in -> rank_out.  MOVE_KEY: keys are only moved while score passes remain.
template <bool FIRST, int SRC, int RANK, bool MOVE_KEY>
__global__ void __launch_bounds__(kSortThreads) sort_pass_kernel(const SortParams p) {
    constexpr int NW = kSortThreads / 32;
    __shared__ int s_warp[NW][256];       // per-warp running digit counters -> destination bases
    __shared__ int s_scan[33];
    __shared__ int s_tile;

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (p.only_flagged && p.only_flagged[b] == 0u) return;
    if (tid == 0) s_tile = (int)atomicAdd(&p.ticket[(size_t)p.pass * p.B + b], 1u);
#pragma unroll
    for (int w = 0; w < NW; ++w) s_warp[w][tid] = 0;
    __syncthreads();
    const int tile = s_tile;
    const int e0 = tile * kSortTile;
    const int limit = FIRST ? p.n_pad : (int)p.count[b];
    if (e0 >= limit) return;
    const uint32_t* tc = p.tile_count + (size_t)b * p.n_tiles;
    const size_t img = (size_t)b * p.n_pad;

    // ---- load own items (warp-striped: item k of lane l = warp_base + 32k + l) and rank them ---------
    uint32_t key[kSortItems], pay[kSortItems], rnk[kSortItems];
    int lrank[kSortItems];               // rank inside the warp's digit run, or -1
    uint32_t dig[kSortItems];
    const int wbase = e0 + warp * (32 * kSortItems);
    bool valid[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int e = wbase + k * 32 + lane;
        valid[k] = e < limit && (FIRST ? sparse_valid(tc, e) : true);
        key[k] = 0; pay[k] = 0; rnk[k] = 0;
        if (valid[k]) {
            if (MOVE_KEY || SRC == 0) key[k] = p.key_in[img + e];
            pay[k] = p.pay_in[img + e];
            if (RANK == 1) rnk[k] = (uint32_t)e;
            if (RANK == 2) rnk[k] = p.rank_in[img + e];
        }
    }
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t src = SRC == 0 ? key[k] : pay[k];
        dig[k] = (src >> p.shift) & 0xFFu;
        const unsigned peers = match_digit8(dig[k], valid[k]);
        int base = 0;
        if (valid[k]) base = s_warp[warp][dig[k]];
        __syncwarp();
        if (valid[k] && (peers & lanemask_lt()) == 0) s_warp[warp][dig[k]] = base + __popc(peers);
        __syncwarp();
        lrank[k] = valid[k] ? base + __popc(peers & lanemask_lt()) : -1;
    }
    __syncthreads();

    // ---- per-digit: tile count -> publish -> look back over the preceding tiles of this image ------
    {
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) cnt += s_warp[w][tid];
        uint32_t* stat = p.status + (((size_t)p.pass * p.B + b) * p.sort_tiles) * 256 + tid;
        int excl = 0;
        if (tile == 0) {
            st_relaxed(stat, kFlagPrefix | (uint32_t)cnt);
        } else {
            st_relaxed(stat + (size_t)tile * 256, kFlagAgg | (uint32_t)cnt);
            for (int tt = tile - 1; tt >= 0; --tt) {
                uint32_t v;
                while (((v = ld_relaxed(stat + (size_t)tt * 256)) >> 30) == 0u) __nanosleep(40);
                excl += (int)(v & kValMask);
                if (v & kFlagPrefix) break;
            }
            st_relaxed(stat + (size_t)tile * 256, kFlagPrefix | (uint32_t)(excl + cnt));
        }
        // destination bases: digit_start (image totals) + preceding tiles + preceding warps
        const uint32_t* tot = p.digit_hist + ((size_t)b * kMaxPasses + p.pass) * 256;
        int total_unused;
        const int dstart = block_exclusive_scan((int)tot[tid], s_scan, &total_unused);
        int run = dstart + excl;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = s_warp[w][tid];
            s_warp[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        if (lrank[k] >= 0) {
            const size_t o = img + (size_t)(s_warp[warp][dig[k]] + lrank[k]);
            if (MOVE_KEY) p.key_out[o] = key[k];
            p.pay_out[o] = pay[k];
            if (RANK != 0) p.rank_out[o] = rnk[k];
        }
    }
}


// Dense route, step 1: move the survivors of every image from the tile-sparse K1 layout to positions [0, count) of
// (dkey, dpay), in slot order (= candidate order: the later passes are stable).  One CTA per 8 candidate tiles; the first CTA of
// an image also writes the image's count, zeroes what the later stages expect zeroed (emit chunk counters; digit histogram and
// tickets of the gated multi-launch sort) and raises the overflow flag when the survivors do not fit the cluster.
constexpr int kCompactTiles = 8;
__global__ void __launch_bounds__(256) yolo_compact_kernel(const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ key,
                                                           const uint32_t* __restrict__ pay, uint32_t* __restrict__ dkey,
                                                           uint32_t* __restrict__ dpay, uint32_t* __restrict__ count,
                                                           uint32_t* __restrict__ chunk_cnt, int n_chunks,
                                                           uint32_t* __restrict__ digit_hist, uint32_t* __restrict__ ticket,
                                                           uint32_t* __restrict__ overflow, int capacity, int n_pad, int n_tiles, int B) {
    __shared__ int s_scan[33];
    __shared__ int s_base[kCompactTiles + 1];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int t0 = blockIdx.x * kCompactTiles;
    const uint32_t* tc = tile_count + (size_t)b * n_tiles;
    // survivors in the tiles before this CTA's (and, for the first CTA, in the whole image)
    const int upto = blockIdx.x == 0 ? n_tiles : t0;
    int part = 0;
    for (int t = tid; t < upto; t += 256) part += (int)tc[t];
    int before;
    block_exclusive_scan(part, s_scan, &before);
    if (blockIdx.x == 0) {
        const int total = before;
        before = 0;
        if (tid == 0) {
            count[b] = (uint32_t)total;
            overflow[b] = total > capacity ? 1u : 0u;
        }
        for (int i = tid; i < n_chunks; i += 256) chunk_cnt[(size_t)b * n_chunks + i] = 0u;
        for (int i = tid; i < kMaxPasses * 256; i += 256) digit_hist[(size_t)b * kMaxPasses * 256 + i] = 0u;
        if (tid < kMaxPasses) ticket[(size_t)tid * B + b] = 0u;
    }
    if (tid == 0) {
        int run = before;
        for (int k = 0; k < kCompactTiles; ++k) {
            s_base[k] = run;
            run += (t0 + k < n_tiles) ? (int)tc[t0 + k] : 0;
        }
        s_base[kCompactTiles] = run;
    }
    __syncthreads();
    const size_t img = (size_t)b * n_pad;
    static_assert(kCompactTiles == 256 / 32, "one warp per tile");
    {
        const int k = tid >> 5, lane = tid & 31;
        const int cnt = s_base[k + 1] - s_base[k];
        const size_t src = img + (size_t)(t0 + k) * kTile, dst = img + (size_t)s_base[k];
        for (int i = lane; i < cnt; i += 32) {
            dkey[dst + i] = key[src + i];
            dpay[dst + i] = pay[src + i];
        }
    }
}

// class segment offsets of the flagged images from the class digit totals of the multi-launch sort's histogram
__global__ void __launch_bounds__(256) seg_from_hist_kernel(const uint32_t* __restrict__ digit_hist, const uint32_t* __restrict__ flagged,
                                                            const uint32_t* __restrict__ count, uint32_t* __restrict__ seg_off, int C) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    if (flagged[b] == 0u) return;
    const uint32_t* h = digit_hist + ((size_t)b * kMaxPasses + kScorePasses) * 256;
    const int c = threadIdx.x;
    int total;
    const int ex = block_exclusive_scan(c < C ? (int)h[c] : 0, s_scan, &total);
    if (c < C) seg_off[(size_t)b * (C + 1) + c] = (uint32_t)ex;
    if (c == 0) seg_off[(size_t)b * (C + 1) + C] = count[b];
}

// Score-only sort (4 passes) for the prior pipeline; the sorted payload ends in pay[0].
int score_sort_launch(const uint32_t* tile_count, const uint32_t* count, uint32_t* digit_hist, uint32_t* ticket,
                      uint32_t* status, uint32_t* key[2], uint32_t* pay[2], int n_pad, int n_tiles, int batch,
                      cudaStream_t st) {
    if (use_cluster_sort(n_pad))
        return cluster_sort_launch(tile_count, const_cast<uint32_t*>(count), false, nullptr, 0, nullptr, key, pay, nullptr, n_pad,
                                   n_tiles, 0, 0, batch, st);
    SortParams p;
    memset(&p, 0, sizeof(p));
    p.tile_count = tile_count; p.count = count; p.digit_hist = digit_hist;
    p.ticket = ticket; p.status = status; p.sort_tiles = ceil_div(n_pad, kSortTile); p.B = batch;
    p.n_pad = n_pad; p.n_tiles = n_tiles; p.n_cls_passes = 0;
    dim3 grid(ceil_div(n_pad, kSortTile), batch);
    p.key_in = key[0]; p.pay_in = pay[0];
    sort_hist_kernel<<<grid, kSortThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("sort_hist_kernel");
    for (int pass = 0; pass < kScorePasses; ++pass) {
        const int src = pass & 1, dst = src ^ 1;
        p.key_in = key[src]; p.pay_in = pay[src];
        p.key_out = key[dst]; p.pay_out = pay[dst];
        p.pass = pass; p.shift = 8 * pass;
        if (pass == 0) sort_pass_kernel<true, 0, 0, true><<<grid, kSortThreads, 0, st>>>(p);
        else if (pass < kScorePasses - 1) sort_pass_kernel<false, 0, 0, true><<<grid, kSortThreads, 0, st>>>(p);
        else sort_pass_kernel<false, 0, 0, false><<<grid, kSortThreads, 0, st>>>(p);
        B2_LAUNCH_CHECK("sort_pass_kernel(score)");
    }
    return 0;
}

// True when score_sort_launch takes the multi-launch route for this size (its histogram kernel then leaves zeroed look-back
// words and tickets for all kMaxPasses passes, which partition_pass_launch below relies on).
bool score_sort_is_lookback(int n_pad) { return !use_cluster_sort(n_pad); }

// One more stable 8-bit pass over the payloads of a SINGLE sequence sorted by score_sort_launch (look-back route): a stable
// partition into <= 256 bins by (pay >> shift) & 255.  `digit_hist[pass]` must hold the bin totals (the caller counts them).
int partition_pass_launch(const uint32_t* count, uint32_t* digit_hist, uint32_t* ticket, uint32_t* status,
                          const uint32_t* pay_in, uint32_t* pay_out, int n_pad, int n_tiles, int pass, int shift,
                          cudaStream_t st) {
    SortParams p;
    memset(&p, 0, sizeof(p));
    p.count = count; p.digit_hist = digit_hist; p.ticket = ticket; p.status = status;
    p.sort_tiles = ceil_div(n_pad, kSortTile); p.B = 1; p.n_pad = n_pad; p.n_tiles = n_tiles;
    p.pay_in = pay_in; p.pay_out = pay_out; p.pass = pass; p.shift = shift;
    dim3 grid(ceil_div(n_pad, kSortTile), 1);
    sort_pass_kernel<false, 1, 0, false><<<grid, kSortThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("sort_pass_kernel(partition)");
    return 0;
}

int yolo_stage_sort(const b200det_yolo_desc* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = yolo_validate(d, ws, ws_bytes);
    if (rc) return rc;
    YoloWs w;
    yolo_ws_layout(d, ws, &w);
    return class_score_sort(w, yolo_fast_path(w), st);
}

// (class asc, score desc, index asc) sort of the tile-sparse (key, payload) pairs described by `w` (also used by the
// AP kernels of metrics.cu with a hand-filled YoloWs).  `fast`: count / chunk_cnt are derived here (yolo_fast_path).
int class_score_sort(const YoloWs& w, bool fast, cudaStream_t st) {
    int rc = 0;
    struct { int num_classes, batch; } dd = {w.C, w.B};
    const auto* d = &dd;
    if (use_cluster_sort(w.n_pad)) {
        uint32_t* seg_off = w.n_cls_passes == 1 ? w.seg_off : nullptr;
        if (!seg_off) {
            rc = seg_scan_launch(w.cls_hist, w.seg_off, d->num_classes, d->batch, st);
            if (rc) return rc;
        }
        uint32_t* key[2] = {w.key[0], w.key[1]};
        uint32_t* pay[2] = {w.pay[0], w.pay[1]};
        uint32_t* rank[2] = {w.rank[0], w.rank[1]};
        return cluster_sort_launch(w.tile_count, w.count, fast, fast ? w.chunk_cnt : nullptr, w.n_chunks, seg_off, key,
                                   pay, rank, w.n_pad, w.n_tiles, d->num_classes, w.n_cls_passes, d->batch, st);
    }
    SortParams p;
    memset(&p, 0, sizeof(p));
    p.tile_count = w.tile_count; p.count = w.count; p.digit_hist = w.digit_hist;
    p.ticket = w.ticket; p.status = w.status; p.sort_tiles = w.sort_tiles; p.B = w.B;
    p.n_pad = w.n_pad; p.n_tiles = w.n_tiles; p.n_cls_passes = w.n_cls_passes;
    dim3 grid(ceil_div(w.n_pad, kSortTile), d->batch);

    const bool dense = fast && use_dense_sort(w.n_pad, w.n_cls_passes);
    uint32_t* const overflow = w.cls_hist;         // [B] words of the (otherwise unused on this route) class histogram
    if (dense) {
        // survivors -> (rank[0], rank[1]) (free until the class pass), cluster of 4 per image, then the gated multi-launch sort
        dim3 cgrid(ceil_div(w.n_tiles, kCompactTiles), d->batch);
        yolo_compact_kernel<<<cgrid, 256, 0, st>>>(w.tile_count, w.key[0], w.pay[0], w.rank[0], w.rank[1], w.count, w.chunk_cnt,
                                                   w.n_chunks, w.digit_hist, w.ticket, overflow, cluster_sort_dense_capacity(),
                                                   w.n_pad, w.n_tiles, w.B);
        B2_LAUNCH_CHECK("yolo_compact_kernel");
        uint32_t* key[2] = {w.key[0], w.key[1]};
        uint32_t* pay[2] = {w.pay[0], w.pay[1]};
        uint32_t* rank[2] = {w.rank[0], w.rank[1]};
        rc = cluster_sort_dense_launch(w.tile_count, w.count, w.seg_off, w.rank[0], w.rank[1], overflow, key, pay, rank, w.n_pad,
                                       w.n_tiles, d->num_classes, w.n_cls_passes, d->batch, st);
        if (rc) return rc;
        p.only_flagged = overflow;
        p.dense_hist = 1;
    } else {
        rc = seg_scan_launch(w.cls_hist, w.seg_off, d->num_classes, d->batch, st);
        if (rc) return rc;
    }
    p.key_in = dense ? w.rank[0] : w.key[0]; p.pay_in = dense ? w.rank[1] : w.pay[0];
    sort_hist_kernel<<<grid, kSortThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("sort_hist_kernel");
    if (dense) {
        seg_from_hist_kernel<<<d->batch, 256, 0, st>>>(w.digit_hist, overflow, w.count, w.seg_off, d->num_classes);
        B2_LAUNCH_CHECK("seg_from_hist_kernel");
    }

    for (int pass = 0; pass < kScorePasses; ++pass) {
        const int src = pass & 1, dst = src ^ 1;
        p.key_in = w.key[src]; p.pay_in = w.pay[src];
        p.key_out = w.key[dst]; p.pay_out = w.pay[dst];
        if (dense && pass == 0) { p.key_in = w.rank[0]; p.pay_in = w.rank[1]; }
        p.pass = pass; p.shift = 8 * pass;
        const bool last_score = pass == kScorePasses - 1;
        if (pass == 0 && dense) sort_pass_kernel<false, 0, 0, true><<<grid, kSortThreads, 0, st>>>(p);
        else if (pass == 0) sort_pass_kernel<true, 0, 0, true><<<grid, kSortThreads, 0, st>>>(p);
        else if (!last_score) sort_pass_kernel<false, 0, 0, true><<<grid, kSortThreads, 0, st>>>(p);
        else sort_pass_kernel<false, 0, 0, false><<<grid, kSortThreads, 0, st>>>(p);
        B2_LAUNCH_CHECK("sort_pass_kernel(score)");
    }
    // after 4 score passes the data sits in buffer 0 (keys are no longer needed)
    p.key_in = nullptr; p.key_out = nullptr;
    p.pay_in = w.pay[0]; p.pay_out = w.pay[1]; p.rank_in = nullptr; p.rank_out = w.rank[0];
    p.pass = kScorePasses; p.shift = kSlotBits;
    sort_pass_kernel<false, 1, 1, false><<<grid, kSortThreads, 0, st>>>(p);
    B2_LAUNCH_CHECK("sort_pass_kernel(class lo)");
    if (w.n_cls_passes == 2) {
        p.pay_in = w.pay[1]; p.pay_out = w.pay[0]; p.rank_in = w.rank[0]; p.rank_out = w.rank[1];
        p.pass = kScorePasses + 1; p.shift = kSlotBits + 8;
        sort_pass_kernel<false, 1, 2, false><<<grid, kSortThreads, 0, st>>>(p);
        B2_LAUNCH_CHECK("sort_pass_kernel(class hi)");
    }
    return 0;
}

}  // namespace b200det
