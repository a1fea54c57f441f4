// Detection packing for the one exchange step after the hot path: the padded per-image output of the NMS
// pipeline, rows [B, row_pitch, 7] + count [B], becomes ONE dense [K, 8] array — the 7 detection columns plus the GLOBAL
// image id — that an all-gather can move as a single buffer (LightningFunc/step.py:95,102-130 consume the detections of
// all images together).  Image b's rows land at [offsets[b], offsets[b+1]), in their score order.
#include "common.cuh"

namespace b200det {

constexpr int kPackThreads = 256;

__global__ void __launch_bounds__(kPackThreads)
pack_detections_kernel(const float* __restrict__ rows, const int32_t* __restrict__ count, int B, long long row_pitch,
                       int image_offset, float4* __restrict__ out, long long cap, int32_t* __restrict__ offsets) {
    __shared__ int s_part[kPackThreads / 32];
    __shared__ int s_base;
    const int b = blockIdx.y, tid = threadIdx.x;
    const int n = max(count[b], 0);
    const int r = blockIdx.x * kPackThreads + tid;
    if (blockIdx.x * kPackThreads >= n && blockIdx.x != 0) return;         // CTA-uniform
    // offset of the image = sum of the counts before it (B is a batch size: a few strided loads per thread)
    int part = 0;
    for (int i = tid; i < b; i += kPackThreads) part += max(count[i], 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        int base = 0;
        for (int w = 0; w < kPackThreads / 32; ++w) base += s_part[w];
        s_base = base;
        if (blockIdx.x == 0 && offsets) {
            offsets[b] = base;
            if (b == B - 1) offsets[B] = base + n;
        }
    }
    __syncthreads();
    if (r >= n) return;
    const long long dst = (long long)s_base + r;
    if (dst >= cap) return;
    const float* src = rows + ((long long)b * row_pitch + r) * 7;
    out[2 * dst] = make_float4(src[0], src[1], src[2], src[3]);
    out[2 * dst + 1] = make_float4(src[4], src[5], src[6], (float)(image_offset + b));
}

int pack_detections_launch(const float* rows, const int32_t* count, int B, long long row_pitch, int max_rows, int image_offset,
                           float* out, long long cap, int32_t* offsets, cudaStream_t st) {
    dim3 grid(max(1, ceil_div(max_rows, kPackThreads)), B);
    pack_detections_kernel<<<grid, kPackThreads, 0, st>>>(rows, count, B, row_pitch, image_offset, (float4*)out, cap, offsets);
    B2_LAUNCH_CHECK("pack_detections_kernel");
    return 0;
}

}  // namespace b200det
