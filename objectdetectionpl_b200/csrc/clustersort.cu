// K2c — per-image LSD radix sort as ONE launch: a thread-block cluster per image, all passes inside the kernel.
//
// Same order and same output buffers as the multi-launch sort in segsort.cu ((class asc, score desc, candidate index
// asc); sorted payload / score rank in yolo_sorted_pay / yolo_sorted_rank; score-only variant ends in pay[0]), so the two
// are interchangeable and tested bit-identical.  What changes is the plumbing around the radix passes:
//   * a cluster of CL CTAs (CL = 1, 2, 4, 8; 512 threads x 13 keys = 6 656 keys per CTA) owns one image; the CTAs
//     hold the image's keys in position order, so "preceding tiles of the same image" are simply the lower cluster ranks;
//   * per pass each CTA ranks its keys (stable warp-match ranking, 8 bits), publishes its 256 digit counts in shared
//     memory, `barrier.cluster`, and every CTA reads its peers' counts over DSMEM: digit totals (-> digit starts, no
//     up-front histogram kernel) and the counts of the lower ranks (-> no decoupled look-back, no status words);
//   * keys / payloads are scattered through L2 (`st.global.cg`) and re-read by the next pass after the second
//     `barrier.cluster` of the pass (release/acquire at cluster scope orders the global writes); only the 8-bit digit
//     and the in-warp rank stay in registers between ranking and scatter, the moved words are re-read (coalesced L2 hits);
//   * the class pass's digit starts ARE the class segment offsets, so seg_off is written here as well (C <= 256).
// Five launches + histogram + segment scan (~140 us at the headline) become one (~see profiles/).
#include <cooperative_groups.h>

#include "yolo_ws.cuh"

namespace cg = cooperative_groups;

namespace b200det {

constexpr int kCsThreads = 512;
constexpr int kCsItems = 13;
constexpr int kCsCap = kCsThreads * kCsItems;          // 6 656 keys per CTA = 13 candidate tiles
constexpr int kCsWarps = kCsThreads / 32;
constexpr int kCsMaxCluster = 8;
static_assert(kCsCap % kTile == 0, "CTA capacity must cover whole candidate tiles");

struct ClusterSortParams {
    const uint32_t* tile_count;   // [B][n_tiles]
    uint32_t* count;              // [B]  read, or (count_from_tiles) written: sum of the image's tile counts
    uint32_t* chunk_cnt;          // [B][n_chunks] zeroed here for the NMS stage when count_from_tiles
    int n_chunks;
    int count_from_tiles;
    uint32_t* seg_off;            // [B][C+1] or null (score-only sort / C > 256)
    uint32_t* key[2];
    uint32_t* pay[2];
    uint32_t* rank[2];
    int n_pad, n_tiles, C;
    int n_cls_passes;             // 0 (score only), 1 or 2
    // dense-first mode (large, sparsely surviving images — see yolo_compact_kernel in segsort.cu): the survivors were
    // compacted to positions [0, count) of (dense_key, dense_pay); pass 0 reads those instead of the tile-sparse key[0] / pay[0].
    // An image with more survivors than the cluster holds is left to the multi-launch sort (its `overflow` word is set).
    const uint32_t* dense_key;
    const uint32_t* dense_pay;
    const uint32_t* overflow;     // [B] or null
};

__device__ __forceinline__ bool sparse_valid_cs(const uint32_t* tile_count_img, int e) {
    return (uint32_t)(e & (kTile - 1)) < tile_count_img[e >> kTileShift];
}

template <int CL>
__device__ __forceinline__ void cs_cluster_sync() {
    if (CL == 1) {
        __syncthreads();
    } else {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

// Development aid: when a trace buffer is installed (b200det_debug_set_trace), thread 0 of every CTA stamps
// %globaltimer at the phase boundaries of every pass: trace[((image * CL + rank) * 8 + pass) * 8 + point].
__device__ unsigned long long* g_cs_trace = nullptr;
__device__ __forceinline__ void cs_stamp(unsigned long long* tr, int point) {
    if (tr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tr[point] = t;
    }
}

struct CsShared {
    uint2 tab[kCsWarps][256];        // per (warp, digit): x = lanes of the current step holding the digit, y = running count
    uint16_t pref[kCsWarps][256];    // exclusive prefix of the digit count over the CTA's warps
    int hist[256];                   // this CTA's digit counts (read by the cluster peers over DSMEM)
    int lstart[256];                 // start of a digit inside this CTA's locally sorted staging buffer
    int gbase[256];                  // global position of staged element i with digit d = i + gbase[d]
    int scan[33];
    int cnt;                         // keys held by this CTA in this pass
    uint32_t stage_a[kCsCap];        // key (score passes) / score rank (class passes), locally sorted by digit
    uint32_t stage_b[kCsCap];        // payload
};

// One radix pass of one CTA.  FIRST: tile-sparse input (validity from tile_count), else dense [0, n).
// SCORE: digit from the key (else from the payload).  MOVE_KEY: keys only move while score passes remain.
// RANK: 0 none, 1 = write the input position as score rank, 2 = carry rank_in -> rank_out.
template <int CL, bool FIRST, bool SCORE, bool MOVE_KEY, int RANK>
__device__ __forceinline__ void cs_pass(CsShared& sm, const ClusterSortParams& p, const uint32_t* __restrict__ tc,
                                        const size_t img, const int n, const int crank, const int shift,
                                        const uint32_t* in_key, const uint32_t* in_pay, const uint32_t* in_rank,
                                        uint32_t* out_key, uint32_t* out_pay, uint32_t* out_rank, uint32_t* seg_off_img,
                                        unsigned long long* tr, const int per_warp = 32 * kCsItems) {
    // per_warp: positions owned by a warp (multiple of 32, <= 32 * kCsItems).  The dense route spreads a small image evenly over
    // the cluster's CTAs (fewer serial ranking steps per warp); everywhere else a warp owns its full 13 x 32 positions.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wbase = crank * (per_warp * kCsWarps) + warp * per_warp;
    const int limit = FIRST ? p.n_pad : n;
    const uint32_t* in_digit = SCORE ? in_key : in_pay;
    const unsigned lt = lanemask_lt();
    cs_stamp(tr, 0);

    // ---- rank own keys: item k of lane l = wbase + 32k + l (position order = stability order) ----
    uint32_t packed[kCsItems];          // digit | in-warp rank << 8, or kNone
    uint32_t srcw[kCsItems];
#pragma unroll
    for (int k = 0; k < kCsItems; ++k) {
        const int e = wbase + k * 32 + lane;
        bool valid = e < limit && k * 32 < per_warp;
        if (FIRST) valid = valid && (uint32_t)(e & (kTile - 1)) < tc[e >> kTileShift];
        srcw[k] = valid ? __ldcg(in_digit + img + e) : 0u;
        packed[k] = valid ? 0u : kNone;
    }
    cs_stamp(tr, 1);
    // Lanes holding the same digit find each other through a shared-memory bitmask (one atomicOr + one 64-bit load
    // per key instead of 8 ballots + mask arithmetic); the lowest lane of a group advances the group's counter and
    // clears the mask again, so the table is self-cleaning.
#pragma unroll
    for (int k = 0; k < kCsItems; ++k) {
        if (wbase + k * 32 >= limit || k * 32 >= per_warp) continue;               // warp-uniform
        const bool valid = packed[k] != kNone;
        const uint32_t dig = (srcw[k] >> shift) & 0xFFu;
        uint2* slot = &sm.tab[warp][dig];
        if (valid) atomicOr(&slot->x, 1u << lane);
        __syncwarp();
        uint2 v = make_uint2(0u, 0u);
        if (valid) v = *slot;
        __syncwarp();
        if (valid) {
            if ((v.x & lt) == 0u) *slot = make_uint2(0u, v.y + (uint32_t)__popc(v.x));
            packed[k] = dig | ((v.y + (uint32_t)__popc(v.x & lt)) << 8);
        }
        __syncwarp();
    }
    __syncthreads();
    cs_stamp(tr, 2);

    // ---- CTA digit counts; per-warp counts become exclusive prefixes over the warps; counters reset ----
    if (tid < 256) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < kCsWarps; ++w) {
            const int c = (int)sm.tab[w][tid].y;
            sm.tab[w][tid].y = 0u;
            sm.pref[w][tid] = (uint16_t)run;
            run += c;
        }
        sm.hist[tid] = run;
    }
    cs_cluster_sync<CL>();
    cs_stamp(tr, 3);

    // ---- digit totals over the cluster, counts of the lower ranks, digit starts ----
    {
        int tot = 0, before = 0;
        if (tid < 256) {
            if (CL == 1) {
                tot = sm.hist[tid];
            } else {
                cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
                for (int c = 0; c < CL; ++c) {
                    const int h = *cluster.map_shared_rank(&sm.hist[tid], c);
                    tot += h;
                    if (c < crank) before += h;
                }
            }
        }
        // one scan for both prefixes: digit starts over the whole image (< 2^16) and inside this CTA (< 2^13)
        const int own = tid < 256 ? sm.hist[tid] : 0;
        int total_unused;
        const int ex = block_exclusive_scan((tot << 13) | own, sm.scan, &total_unused);
        const int dstart = ex >> 13, lstart = ex & 8191;
        if (tid < 256) {
            sm.lstart[tid] = lstart;
            sm.gbase[tid] = dstart + before - lstart;
            if (tid == 255) sm.cnt = lstart + own;
            if (RANK == 1 && seg_off_img && crank == 0) {
                // class digit starts = class segment offsets (model/YOLOV3.py:324 label_match segments)
                if (tid < p.C) seg_off_img[tid] = (uint32_t)dstart;
                if (tid == 0) seg_off_img[p.C] = (uint32_t)n;
            }
        }
    }
    __syncthreads();
    cs_stamp(tr, 4);

    // ---- local reorder: stage the CTA's keys sorted by digit in shared memory, then write every digit run to its
    //      global destination with consecutive threads on consecutive words.  (Scattering 4-byte words straight from
    //      registers costs one 32-byte L2 sector write per word: measured 11-16 us per pass at the headline.) ----
    {
        uint32_t second[kCsItems];
#pragma unroll
        for (int k = 0; k < kCsItems; ++k) {
            const int e = wbase + k * 32 + lane;
            second[k] = 0u;
            if (packed[k] != kNone) {
                if (SCORE) second[k] = __ldcg(in_pay + img + e);
                else if (RANK == 1) second[k] = (uint32_t)e;
                else second[k] = __ldcg(in_rank + img + e);
            }
        }
#pragma unroll
        for (int k = 0; k < kCsItems; ++k) {
            if (packed[k] != kNone) {
                const uint32_t dig = packed[k] & 0xFFu;
                const int lp = sm.lstart[dig] + (int)sm.pref[warp][dig] + (int)(packed[k] >> 8);
                sm.stage_a[lp] = SCORE ? srcw[k] : second[k];
                sm.stage_b[lp] = SCORE ? second[k] : srcw[k];
            }
        }
    }
    __syncthreads();
    {
        const int cnt = sm.cnt;
        for (int i = tid; i < cnt; i += kCsThreads) {
            const uint32_t a = sm.stage_a[i], pb = sm.stage_b[i];
            const uint32_t dig = ((SCORE ? a : pb) >> shift) & 0xFFu;
            const size_t o = img + (size_t)(i + sm.gbase[dig]);
            if (MOVE_KEY) __stcg(out_key + o, a);
            __stcg(out_pay + o, pb);
            if (RANK != 0) __stcg(out_rank + o, a);
        }
    }
    cs_stamp(tr, 5);
    cs_cluster_sync<CL>();      // scattered words visible to the peers; hist may be overwritten
    cs_stamp(tr, 6);
}

template <int CL, bool DENSE = false>
__global__ void __launch_bounds__(kCsThreads, 2) cluster_sort_kernel(const ClusterSortParams p) {
    extern __shared__ __align__(16) unsigned char cs_smem_raw[];
    CsShared& sm = *reinterpret_cast<CsShared*>(cs_smem_raw);
    const int b = blockIdx.y;
    const int crank = CL == 1 ? 0 : (int)blockIdx.x;     // cluster dims are (CL, 1, 1) and gridDim.x == CL
    const int tid = threadIdx.x;
    const size_t img = (size_t)b * p.n_pad;
    const uint32_t* tc = p.tile_count + (size_t)b * p.n_tiles;
    int n;
    if (p.count_from_tiles) {
        int part = 0;
        for (int t = tid; t < p.n_tiles; t += kCsThreads) part += (int)tc[t];
        block_exclusive_scan(part, sm.scan, &n);
        if (crank == 0) {
            if (tid == 0) p.count[b] = (uint32_t)n;
            for (int i = tid; i < p.n_chunks; i += kCsThreads) p.chunk_cnt[(size_t)b * p.n_chunks + i] = 0u;
        }
    } else {
        n = (int)p.count[b];
    }
    unsigned long long* tr = g_cs_trace ? g_cs_trace + ((size_t)(b * CL + crank) * 8) * 8 : nullptr;
    if (DENSE && p.overflow[b]) return;                 // cluster-uniform: too many survivors, the multi-launch sort takes it

    for (int i = tid; i < kCsWarps * 256; i += kCsThreads) (&sm.tab[0][0])[i] = make_uint2(0u, 0u);
    __syncthreads();

    // positions per warp: the full 13 x 32 except on the dense route, where the image's n keys are dealt evenly to the CL CTAs
    int pw = 32 * kCsItems;
    if (DENSE) {
        const int per_cta = ((n + CL - 1) / CL + kCsThreads - 1) / kCsThreads * kCsThreads;      // multiple of 512, <= kCsCap
        pw = min(32 * kCsItems, max(32, per_cta / kCsWarps));
    }
    // score passes: key/pay ping-pong 0 -> 1 -> 0 -> 1 -> 0 (the last one moves the payload only)
    if (DENSE)
        cs_pass<CL, false, true, true, 0>(sm, p, tc, img, n, crank, 0, p.dense_key, p.dense_pay, nullptr, p.key[1], p.pay[1], nullptr, nullptr, tr, pw);
    else
        cs_pass<CL, true, true, true, 0>(sm, p, tc, img, n, crank, 0, p.key[0], p.pay[0], nullptr, p.key[1], p.pay[1], nullptr, nullptr, tr, pw);
    cs_pass<CL, false, true, true, 0>(sm, p, tc, img, n, crank, 8, p.key[1], p.pay[1], nullptr, p.key[0], p.pay[0], nullptr, nullptr, tr ? tr + 8 : tr, pw);
    cs_pass<CL, false, true, true, 0>(sm, p, tc, img, n, crank, 16, p.key[0], p.pay[0], nullptr, p.key[1], p.pay[1], nullptr, nullptr, tr ? tr + 16 : tr, pw);
    cs_pass<CL, false, true, false, 0>(sm, p, tc, img, n, crank, 24, p.key[1], p.pay[1], nullptr, nullptr, p.pay[0], nullptr, nullptr, tr ? tr + 24 : tr, pw);
    if (p.n_cls_passes >= 1) {
        uint32_t* so = p.seg_off ? p.seg_off + (size_t)b * (p.C + 1) : nullptr;
        cs_pass<CL, false, false, false, 1>(sm, p, tc, img, n, crank, (int)kSlotBits, nullptr, p.pay[0], nullptr, nullptr, p.pay[1], p.rank[0], so, tr ? tr + 32 : tr, pw);
    }
    if (p.n_cls_passes >= 2)
        cs_pass<CL, false, false, false, 2>(sm, p, tc, img, n, crank, (int)kSlotBits + 8, nullptr, p.pay[1], p.rank[0], nullptr, p.pay[0], p.rank[1], nullptr, tr ? tr + 40 : tr, pw);
}

// Largest candidate-slot count per image the cluster sort handles (8 CTAs x 6 656 keys).
int cluster_sort_capacity() { return kCsCap * kCsMaxCluster; }

template <int CL, bool DENSE = false>
static int cluster_sort_launch_cl(const ClusterSortParams& p, int batch, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, batch, 1);
    cfg.blockDim = dim3(kCsThreads, 1, 1);
    cfg.dynamicSmemBytes = sizeof(CsShared);
    // The opt-in above the 48 KB default is a per-DEVICE function attribute: cache it per device ordinal (one process may
    // drive several GPUs).  The attribute is idempotent, so a race between threads only repeats the call.
    static bool attr_set[64] = {};
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        B2_CUDA(cudaFuncSetAttribute(cluster_sort_kernel<CL, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CsShared)));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2_CUDA(cudaLaunchKernelEx(&cfg, cluster_sort_kernel<CL, DENSE>, p));
    return 0;
}

// One-launch sort of every image's candidates.  n_cls_passes = 0: score order only, result in pay[0];
// otherwise (class, score) order in yolo_sorted_pay / yolo_sorted_rank and, when seg_off != null, the class offsets.
// Dense-first variant for images with more slots than a cluster holds: survivors already compacted into (dense_key, dense_pay),
// count[b] final, overflow[b] set for images with more than cluster_sort_dense_capacity() survivors (those are skipped here).
int cluster_sort_dense_capacity() { return kCsCap * 4; }          // cluster of 4: 64 images = 256 CTAs = one wave on 148 SMs

int cluster_sort_dense_launch(const uint32_t* tile_count, uint32_t* count, uint32_t* seg_off, const uint32_t* dense_key,
                              const uint32_t* dense_pay, const uint32_t* overflow, uint32_t* key[2], uint32_t* pay[2],
                              uint32_t* rank[2], int n_pad, int n_tiles, int C, int n_cls_passes, int batch, cudaStream_t st) {
    ClusterSortParams p;
    memset(&p, 0, sizeof(p));
    p.tile_count = tile_count; p.count = count; p.seg_off = seg_off;
    for (int i = 0; i < 2; ++i) { p.key[i] = key[i]; p.pay[i] = pay[i]; p.rank[i] = rank[i]; }
    p.n_pad = n_pad; p.n_tiles = n_tiles; p.C = C; p.n_cls_passes = n_cls_passes;
    p.dense_key = dense_key; p.dense_pay = dense_pay; p.overflow = overflow;
    return cluster_sort_launch_cl<4, true>(p, batch, st);
}

int cluster_sort_launch(const uint32_t* tile_count, uint32_t* count, bool count_from_tiles, uint32_t* chunk_cnt,
                        int n_chunks, uint32_t* seg_off, uint32_t* key[2], uint32_t* pay[2], uint32_t* rank[2], int n_pad,
                        int n_tiles, int C, int n_cls_passes, int batch, cudaStream_t st) {
    B2_CHECK_LIMIT(n_pad <= cluster_sort_capacity(), "cluster sort: %d slots per image > %d", n_pad, cluster_sort_capacity());
    ClusterSortParams p;
    memset(&p, 0, sizeof(p));
    p.tile_count = tile_count; p.count = count; p.seg_off = seg_off;
    p.count_from_tiles = count_from_tiles ? 1 : 0; p.chunk_cnt = chunk_cnt; p.n_chunks = n_chunks;
    for (int i = 0; i < 2; ++i) { p.key[i] = key[i]; p.pay[i] = pay[i]; p.rank[i] = rank ? rank[i] : nullptr; }
    p.n_pad = n_pad; p.n_tiles = n_tiles; p.C = C; p.n_cls_passes = n_cls_passes;
    if (n_pad <= kCsCap) return cluster_sort_launch_cl<1>(p, batch, st);
    if (n_pad <= 2 * kCsCap) return cluster_sort_launch_cl<2>(p, batch, st);
    if (n_pad <= 4 * kCsCap) return cluster_sort_launch_cl<4>(p, batch, st);
    return cluster_sort_launch_cl<8>(p, batch, st);
}

}  // namespace b200det

// Development aid (not part of include/b200det.h): install / remove the phase-trace buffer of the cluster sort.
extern "C" int b200det_debug_set_trace(void* dev_ptr) {
    unsigned long long* p = (unsigned long long*)dev_ptr;
    return (int)cudaMemcpyToSymbol(b200det::g_cs_trace, &p, sizeof(p));
}
