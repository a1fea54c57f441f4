// N3 / N4 / T3 / T5 — elementwise box maths with the reference's exact operation order.
//   xywh2xyxy      LightningFunc/accuracy.py:289-295
//   bbox_iou       LightningFunc/accuracy.py:39-69   (IoU with the +1 pixel convention, +1e-16)
//   iou            LightningFunc/accuracy.py:6-37    (corner boxes, no +1, no eps)
//   bbox_iou_v5    LightningFunc/accuracy.py:71-114  (IoU/GIoU/DIoU/CIoU on transposed [4,n] boxes) + backward
// All arithmetic is explicit round-to-nearest (no FMA contraction) in the order the reference's eager
// ops execute, so results agree with the fp32 CPU path to the last bit wherever no transcendental
// (atan) is involved.
#include "common.cuh"
#include "boxmath.cuh"

namespace b200det {

__global__ void xywh2xyxy_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = x[i];
    const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
    y[i] = make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
}

__global__ void bbox_iou_plus1_kernel(const float4* __restrict__ b1, long long n1, const float4* __restrict__ b2,
                                      long long n, int corner, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = b1[n1 == 1 ? 0 : i], b = b2[i];
    if (!corner) { a = cxcywh_to_corners(a); b = cxcywh_to_corners(b); }
    out[i] = iou_plus1_eps(a, b);
}

__global__ void pair_iou_kernel(const float4* __restrict__ t1, const float4* __restrict__ t2, long long n,
                                float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = iou_plain(t1[i], t2[i]);
}

struct Strided4 {
    const float* p;
    long long ld, inc;
    __device__ __forceinline__ float4 get(long long i) const {
        const float* q = p + i * inc;
        return make_float4(q[0], q[ld], q[2 * ld], q[3 * ld]);
    }
};

__global__ void bbox_iou_v5_fwd_kernel(Strided4 b1, Strided4 b2, long long n, int corner, int kind, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = iou_v5_forward(b1.get(i), b2.get(i), corner != 0, kind);
}

__global__ void bbox_iou_v5_bwd_kernel(Strided4 b1, Strided4 b2, long long n, int corner, int kind,
                                       const float* __restrict__ gout, float* __restrict__ gin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g[4];
    iou_v5_backward(b1.get(i), b2.get(i), corner != 0, kind, gout[i], g);
    gin[i] = g[0]; gin[n + i] = g[1]; gin[2 * n + i] = g[2]; gin[3 * n + i] = g[3];
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

int xywh2xyxy_launch(const float* x, float* y, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    xywh2xyxy_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n);
    B2_LAUNCH_CHECK("xywh2xyxy_kernel");
    return 0;
}
int bbox_iou_plus1_launch(const float* b1, long long n1, const float* b2, long long n, int corner, float* out, cudaStream_t st) {
    if (n == 0) return 0;
    bbox_iou_plus1_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const float4*)b1, n1, (const float4*)b2, n, corner, out);
    B2_LAUNCH_CHECK("bbox_iou_plus1_kernel");
    return 0;
}
int pair_iou_launch(const float* a, const float* b, long long n, float* out, cudaStream_t st) {
    if (n == 0) return 0;
    pair_iou_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const float4*)a, (const float4*)b, n, out);
    B2_LAUNCH_CHECK("pair_iou_kernel");
    return 0;
}
int bbox_iou_v5_fwd_launch(const float* b1, long long ld1, long long inc1, const float* b2, long long ld2, long long inc2,
                           long long n, int corner, int kind, float* out, cudaStream_t st) {
    if (n == 0) return 0;
    bbox_iou_v5_fwd_kernel<<<blocks_for(n, 256), 256, 0, st>>>(Strided4{b1, ld1, inc1}, Strided4{b2, ld2, inc2}, n, corner, kind, out);
    B2_LAUNCH_CHECK("bbox_iou_v5_fwd_kernel");
    return 0;
}
int bbox_iou_v5_bwd_launch(const float* b1, long long ld1, long long inc1, const float* b2, long long ld2, long long inc2,
                           long long n, int corner, int kind, const float* gout, float* gin, cudaStream_t st) {
    if (n == 0) return 0;
    bbox_iou_v5_bwd_kernel<<<blocks_for(n, 256), 256, 0, st>>>(Strided4{b1, ld1, inc1}, Strided4{b2, ld2, inc2}, n, corner, kind, gout, gin);
    B2_LAUNCH_CHECK("bbox_iou_v5_bwd_kernel");
    return 0;
}

}  // namespace b200det
