// K2s — per-image top-k by radix SELECT (the "per-image top-k / radix-select kernel" of the north-star), for the SSD /
// RetinaNet path: model/SSD.py:265-273 sorts all scores and keeps `order[:topk]`; only those rows reach the greedy NMS.
//
// Sorting every candidate to keep 100 of up to 120 087 is 42 % of the RetinaNet-800 step (207 us of look-back radix passes);
// the k smallest keys (= k highest scores; ties by ascending slot, the order of the stable sort it replaces) are found
// instead by three most-significant-digit-first histogram passes (11 + 11 + 10 bits) over the image's keys, which narrow down
// the key T of the k-th element; the elements below T plus the first ties at T are collected (<= k of them) and sorted in
// shared memory.  One 1024-thread CTA per image; the keys are read four times from L2 (480 KB per read at RetinaNet-800).
// Output: pay_out[r], r < min(n, k), in the order the full sort would have produced — same consumer (nms_segment_kernel<1|2>).
#include <cooperative_groups.h>

#include "yolo_ws.cuh"

namespace cg = cooperative_groups;

namespace b200det {

constexpr int kSelThreads = 1024;
constexpr int kSelMaxK = 1024;
constexpr int kSelCluster = 4;          // CTAs per image: the passes are instruction-bound inside one SM (ncu: 54 us at 59 % issue)
constexpr int kSelCap = 2048;            // elements the shared-memory list holds

struct TopkParams {
    const uint32_t* tile_count;   // [B][n_tiles]
    const uint32_t* key;          // [B][n_pad] tile-sparse
    const uint32_t* pay;          // [B][n_pad] tile-sparse  (class << 20 | slot)
    uint32_t* pay_out;            // [B][n_pad] rows 0 .. min(n, k) - 1  (may alias `pay`: written after every read)
    uint32_t* count_out;          // [B] or null: candidates of the image (= sum of its tile counts)
    uint32_t* tile_prefix_out;    // [B][n_tiles] or null: exclusive prefix of the tile counts
    int n_pad, n_tiles, k;
};

template <int CL>
__device__ __forceinline__ void sel_cluster_sync() {
    if (CL == 1) __syncthreads();
    else asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// A cluster of CL CTAs per image (4 for large images, 1 below 16 k slots where the cluster barriers cost more than they save): every CTA histograms / collects its quarter of the image's slots, the histograms
// are merged over DSMEM (every CTA ends up with the same merged copy and takes the same decisions), rank 0 gathers the
// collected elements, sorts and writes them.
template <int CL>
__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const TopkParams p) {
    __shared__ int s_hist[2048];                        // this CTA's counts (read by the peers)
    __shared__ int s_hist_m[2048];                      // merged over the cluster
    __shared__ uint16_t s_tc[B200DET_MAX_CANDIDATES / kTile <= 2048 ? B200DET_MAX_CANDIDATES / kTile : 2048];
    __shared__ unsigned long long s_sel[kSelCap];      // key << 32 | payload  (payload's low 20 bits = slot: the tie order)
    __shared__ int s_cum_le;                            // elements with the known key prefix <= the K-th element's, after a pass
    __shared__ int s_scan[33];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_want, s_nsel, s_tie_take;

    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.y, tid = threadIdx.x;
    const int crank = CL == 1 ? 0 : (int)blockIdx.x;    // cluster dims (CL, 1, 1), gridDim.x == CL
    const size_t img = (size_t)b * p.n_pad;
    const uint32_t* tc = p.tile_count + (size_t)b * p.n_tiles;
    const bool tc_cached = p.n_tiles <= 2048;
    // candidates of the image and, when asked for, the exclusive prefix of its tile counts (consumed by the NMS stage for
    // the reference's filtered-index quirk) — folded in here so that the step needs no counter reset and no prefix launch
    int n = 0;
    for (int t0 = 0; t0 < p.n_tiles; t0 += kSelThreads) {
        const int t = t0 + tid;
        const int c = t < p.n_tiles ? (int)tc[t] : 0;
        if (t < p.n_tiles && tc_cached) s_tc[t] = (uint16_t)c;
        int tot;
        const int ex = block_exclusive_scan(c, s_scan, &tot);
        if (p.tile_prefix_out && crank == 0 && t < p.n_tiles) p.tile_prefix_out[(size_t)b * p.n_tiles + t] = (uint32_t)(n + ex);
        n += tot;
    }
    if (p.count_out && crank == 0 && tid == 0) p.count_out[b] = (uint32_t)n;
    const int K = min(n, p.k);
    auto valid = [&](int e) -> bool {
        const uint32_t c = tc_cached ? s_tc[e >> kTileShift] : tc[e >> kTileShift];
        return (uint32_t)(e & (kTile - 1)) < c;
    };
    // the image's keys are walked as aligned groups of 4 (one 128-bit load; a group never straddles a 512-slot tile), two
    // groups in flight per thread: with one scalar load per iteration the passes were pure L2 latency (~150 us at 120 k slots)
    const uint4* key4 = reinterpret_cast<const uint4*>(p.key + img);
    const int nvec = p.n_pad >> 2;
    const int vchunk = (nvec + CL - 1) / CL;
    const int v_lo = crank * vchunk, v_hi = min(nvec, v_lo + vchunk);      // this CTA's share of the image
    auto group_count = [&](int v) -> int {                  // valid keys at the front of group v
        const int e = v << 2;
        const int c = (int)(tc_cached ? s_tc[e >> kTileShift] : tc[e >> kTileShift]) - (e & (kTile - 1));
        return c < 0 ? 0 : (c > 4 ? 4 : c);
    };
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_want = K; s_nsel = 0; s_tie_take = 0; }
    __syncthreads();
    if (K == 0) return;                                 // cluster-uniform (every CTA computed the same n)

    // ---- three MSD passes: find the key T of the K-th smallest element and how many ties at T are taken ----
    int early_shift = -1;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int ps = 0; ps < 3; ++ps) {
        const int shift = shifts[ps], nb = 1 << widths[ps];
        for (int i = tid; i < nb; i += kSelThreads) s_hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        const int want = s_want;                             // read before this pass's owner thread updates it
        for (int v = v_lo + tid; v < v_hi; v += 2 * kSelThreads) {
            const int v1 = v + kSelThreads;
            const int c0 = group_count(v), c1 = v1 < v_hi ? group_count(v1) : 0;
            uint4 k0 = make_uint4(0u, 0u, 0u, 0u), k1 = k0;
            if (c0) k0 = key4[v];
            if (c1) k1 = key4[v1];
            const uint32_t ks[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool ok = (q & 3) < (q < 4 ? c0 : c1);
                if (ok && (ks[q] & mask) == prefix) atomicAdd(&s_hist[(ks[q] >> shift) & (nb - 1)], 1);
            }
        }
        sel_cluster_sync<CL>();                              // every CTA's counts are complete
        for (int i = tid; i < nb; i += kSelThreads) {
            int m = 0;
#pragma unroll
            for (int c = 0; c < CL; ++c) m += CL == 1 ? s_hist[i] : *cluster.map_shared_rank(&s_hist[i], c);
            s_hist_m[i] = m;
        }
        sel_cluster_sync<CL>();                              // the peers are done reading s_hist (it is zeroed again next pass)
        // smallest digit d with cum(d) >= want: two-level scan (each thread owns nb / 1024 <= 2 bins)
        const int per = nb / kSelThreads > 0 ? nb / kSelThreads : 1;
        int mine = 0;
        if (tid * per < nb)
            for (int q = 0; q < per; ++q) mine += s_hist_m[tid * per + q];
        int total;
        const int ex = block_exclusive_scan(mine, s_scan, &total);
        if (tid * per < nb && ex < want && ex + mine >= want) {
            int run = ex;
            for (int q = 0; q < per; ++q) {
                const int h = s_hist_m[tid * per + q];
                if (run + h >= want) {
                    s_prefix = prefix | ((uint32_t)(tid * per + q) << shift);
                    s_mask = mask | ((uint32_t)(nb - 1) << shift);
                    s_want = want - run;                 // rank of the K-th element inside this bin
                    if (ps == 2) s_tie_take = want - run;   // ties at T that belong to the top K (T's bin is one key now)
                    s_cum_le = (K - (want - run)) + h;
                    break;
                }
                run += h;
            }
        }
        __syncthreads();
        if (s_cum_le <= kSelCap) { early_shift = shift; break; }     // cluster-uniform
    }

    // ---- collect the candidates of this CTA's share into its shared list ----
    //   early finish (the usual case after one or two passes): few enough elements have a key prefix up to the K-th
    //     element's — take them all, the sort below keeps the first K;
    //   after the third pass: everything below T, and the ties at T (all of them when they all belong to the top K;
    //     otherwise the first `tie_take` in slot order — rank 0 walks the image alone for that rare case).
    const uint32_t T = s_prefix;
    const int tie_take = s_tie_take;
    const bool early = early_shift >= 0;
    const bool ordered_ties = !early && tie_take != s_hist_m[T & 1023u];
    if (!ordered_ties) {
        const int cshift = early ? early_shift : 0;
        const uint32_t top = T >> cshift;
        for (int v = v_lo + tid; v < v_hi; v += 2 * kSelThreads) {
            const int v1 = v + kSelThreads;
            const int c0 = group_count(v), c1 = v1 < v_hi ? group_count(v1) : 0;
            uint4 k0 = make_uint4(0u, 0u, 0u, 0u), k1 = k0;
            if (c0) k0 = key4[v];
            if (c1) k1 = key4[v1];
            const uint32_t ks[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool ok = (q & 3) < (q < 4 ? c0 : c1);
                if (ok && (ks[q] >> cshift) <= top) {
                    const int e = ((q < 4 ? v : v1) << 2) + (q & 3);
                    s_sel[atomicAdd(&s_nsel, 1)] = ((unsigned long long)ks[q] << 32) | p.pay[img + e];
                }
            }
        }
    } else if (crank == 0) {
        int tie_base = 0;                                  // ties seen in earlier chunks (position order)
        for (int e0 = 0; e0 < p.n_pad; e0 += kSelThreads) {
            const int e = e0 + tid;
            uint32_t key = 0xFFFFFFFFu;
            bool ok = false;
            if (e < p.n_pad && valid(e)) { key = p.key[img + e]; ok = true; }
            const bool is_tie = ok && key == T;
            int tot;
            const int rank = block_exclusive_scan(is_tie ? 1 : 0, s_scan, &tot) + tie_base;
            if (ok && (key < T || (is_tie && rank < tie_take)))
                s_sel[atomicAdd(&s_nsel, 1)] = ((unsigned long long)key << 32) | p.pay[img + e];
            tie_base += tot;
        }
    }
    sel_cluster_sync<CL>();                                    // every CTA's list is complete
    if (crank == 0 && !ordered_ties) {                     // rank 0 appends its peers' lists to its own
        int base = s_nsel;
        for (int c = 1; c < CL; ++c) {
            const int nc = *cluster.map_shared_rank(&s_nsel, c);
            const unsigned long long* src = cluster.map_shared_rank(&s_sel[0], c);
            for (int i = tid; i < nc; i += kSelThreads) s_sel[base + i] = src[i];
            base += nc;
        }
        __syncthreads();
        if (tid == 0) s_nsel = base;
    }
    sel_cluster_sync<CL>();                                    // the peers' shared memory is no longer needed
    if (crank != 0) return;
    __syncthreads();
    // ---- rank 0: sort the collected elements by (key, slot), write the first K ----
    const int M = s_nsel;
    int P = 32;
    while (P < M) P <<= 1;
    for (int i = M + tid; i < P; i += kSelThreads) s_sel[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < P; i += kSelThreads) {
                const int l = i ^ jj;
                if (l > i) {
                    const unsigned long long a = s_sel[i], c = s_sel[l];
                    // order by key, then by slot (low 20 bits of the payload; the class bits above it must not decide)
                    const unsigned long long ka = (a & 0xFFFFFFFF00000000ull) | (a & kSlotMask);
                    const unsigned long long kc = (c & 0xFFFFFFFF00000000ull) | (c & kSlotMask);
                    const bool up = (i & k) == 0;
                    if ((ka > kc) == up) { s_sel[i] = c; s_sel[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += kSelThreads) p.pay_out[img + i] = (uint32_t)s_sel[i];
}

bool topk_select_supported(int k) { return k >= 1 && k <= kSelMaxK; }

int topk_select_launch(const uint32_t* tile_count, const uint32_t* key, const uint32_t* pay, uint32_t* pay_out, int n_pad,
                       int n_tiles, int k, int batch, cudaStream_t st, uint32_t* count_out, uint32_t* tile_prefix_out) {
    TopkParams p;
    p.tile_count = tile_count; p.key = key; p.pay = pay; p.pay_out = pay_out; p.n_pad = n_pad; p.n_tiles = n_tiles; p.k = k;
    p.count_out = count_out; p.tile_prefix_out = tile_prefix_out;
    if (n_pad <= 16384) {
        topk_select_kernel<1><<<dim3(1, batch, 1), kSelThreads, 0, st>>>(p);
        B2_LAUNCH_CHECK("topk_select_kernel<1>");
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kSelCluster, batch, 1);
    cfg.blockDim = dim3(kSelThreads, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kSelCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2_CUDA(cudaLaunchKernelEx(&cfg, topk_select_kernel<kSelCluster>, p));
    return 0;
}

}  // namespace b200det
