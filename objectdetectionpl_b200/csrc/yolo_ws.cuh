// Workspace layout of the YOLO post-processing pipeline (shared by the stage files).
#pragma once
#include "common.cuh"

namespace b200det {

constexpr int kScorePasses = 4;   // 32-bit score key, 8 bits per pass
constexpr int kMaxPasses = 6;     // + up to two class passes (12-bit class id)
constexpr int kEmitShift = 10;    // emit chunk = 1024 score ranks
constexpr int kEmitChunk = 1 << kEmitShift;

struct YoloWs {
    // --- zeroed at the start of every call (one contiguous memset) ---
    uint32_t* count;       // [B]            surviving candidates per image
    uint32_t* cls_hist;    // [B][C]         survivors per (image, class)
    uint32_t* digit_hist;  // [B][kMaxPasses][256] per-image digit totals of every radix pass
    uint32_t* ticket;      // [kMaxPasses][B]     dynamic tile tickets of the radix passes
    uint32_t* chunk_cnt;   // [B][n_chunks]       kept rows per 1024 score ranks
    size_t zero_bytes;
    uint32_t* status;      // [kMaxPasses][B][sort_tiles][256] decoupled look-back words (zeroed by the histogram kernel)
    // --- plain scratch ---
    uint32_t* seg_off;     // [B][C+1]       class segment offsets in the (class, score) order
    uint32_t* tile_count;  // [B][n_tiles]   survivors per candidate tile (tile-sparse layout)
    float4* box4;          // [B][n_pad]     corner box of the candidate in slot s
    float2* cc2;           // [B][n_pad]     (obj conf, class conf)
    uint32_t* orig;        // [B][n_pad]     original candidate index of slot s
    uint32_t* key[2];      // [B][n_pad]     radix keys (ping-pong); key[0] is written tile-sparse by K1
    uint32_t* pay[2];      // [B][n_pad]     payload = class << 20 | slot
    uint32_t* rank[2];     // [B][n_pad]     global score rank of the element at a class-sorted position
    uint32_t* kpay;        // [B][n_pad]     by score rank: payload of a kept row, kNone otherwise
    float4* mbox;          // [B][n_pad]     by score rank: merged box of a kept row
    float4* kbox;          // [B][n_pad]     per segment: original boxes of the keepers found so far
    float* kacc;           // [B][n_pad][5]  per segment: running merge sums (multi-chunk segments)
    uint32_t* kpos;        // [B][n_pad]     per segment: sorted position of keeper k
    size_t total_bytes;
    int B, C, N, n_pad, n_tiles, n_cls_passes, n_chunks, sort_tiles;
};

// Candidates per image and the slot count of an image's region.  Every level starts on a tile boundary (so that the
// cells a decode thread owns are aligned inside their plane whatever the order of the levels — YOLOv3/v4 list the odd
// 13x13 level first), i.e. n_pad = sum over levels of roundup(A * G^2, tile).
inline int yolo_counts(const b200det_yolo_desc* d, int* n_out, int* n_pad_out) {
    long long n = 0, pad = 0;
    for (int l = 0; l < d->num_levels; ++l) {
        const long long nl = (long long)d->num_anchors * d->grid[l] * d->grid[l];
        n += nl;
        pad += (long long)align_up((size_t)nl, kTile);
    }
    if (n <= 0 || pad > B200DET_MAX_CANDIDATES) return B200DET_ELIMIT;
    *n_out = (int)n;
    *n_pad_out = (int)pad;
    return 0;
}

inline void yolo_ws_layout(const b200det_yolo_desc* d, void* base, YoloWs* w) {
    int N = 0, n_pad = 0;
    yolo_counts(d, &N, &n_pad);
    const size_t B = (size_t)d->batch, C = (size_t)d->num_classes, P = (size_t)n_pad;
    w->B = d->batch; w->C = d->num_classes; w->N = N; w->n_pad = n_pad; w->n_tiles = n_pad / kTile;
    w->n_cls_passes = d->num_classes <= 256 ? 1 : 2;
    w->n_chunks = (n_pad + kEmitChunk - 1) / kEmitChunk;
    w->sort_tiles = (n_pad + kSortTile - 1) / kSortTile;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes, 256); return r; };
    w->count = (uint32_t*)take(B * 4);
    w->cls_hist = (uint32_t*)take(B * C * 4);
    w->digit_hist = (uint32_t*)take(B * kMaxPasses * 256 * 4);
    w->ticket = (uint32_t*)take((size_t)kMaxPasses * B * 4);
    w->chunk_cnt = (uint32_t*)take(B * (size_t)w->n_chunks * 4);
    w->zero_bytes = off;
    w->status = (uint32_t*)take((size_t)kMaxPasses * B * w->sort_tiles * 256 * 4);
    w->seg_off = (uint32_t*)take(B * (C + 1) * 4);
    w->tile_count = (uint32_t*)take(B * (size_t)w->n_tiles * 4);
    w->box4 = (float4*)take(B * P * 16);
    w->cc2 = (float2*)take(B * P * 8);
    w->orig = (uint32_t*)take(B * P * 4);
    for (int i = 0; i < 2; ++i) w->key[i] = (uint32_t*)take(B * P * 4);
    for (int i = 0; i < 2; ++i) w->pay[i] = (uint32_t*)take(B * P * 4);
    for (int i = 0; i < 2; ++i) w->rank[i] = (uint32_t*)take(B * P * 4);
    w->kpay = (uint32_t*)take(B * P * 4);
    w->mbox = (float4*)take(B * P * 16);
    w->kbox = (float4*)take(B * P * 16);
    w->kacc = (float*)take(B * P * 5 * 4);
    w->kpos = (uint32_t*)take(B * P * 4);
    w->total_bytes = off;
}

// Buffers holding the final (class, score)-sorted payload / rank after all passes.
inline const uint32_t* yolo_sorted_pay(const YoloWs& w) { return w.n_cls_passes == 1 ? w.pay[1] : w.pay[0]; }
inline const uint32_t* yolo_sorted_rank(const YoloWs& w) { return w.n_cls_passes == 1 ? w.rank[0] : w.rank[1]; }

int yolo_validate(const b200det_yolo_desc* d, const void* ws, size_t ws_bytes);

// True when the sort stage is the one-launch cluster sort AND it derives everything the later stages need (count,
// class offsets, zeroed chunk counters) itself: then the reset stage launches nothing and the decode kernel skips its
// global counter atomics.  (segsort.cu)
bool yolo_fast_path(const YoloWs& w);
int class_score_sort(const YoloWs& w, bool fast, cudaStream_t st);
int zero_fill_launch(void* p, size_t bytes, cudaStream_t st);

}  // namespace b200det
