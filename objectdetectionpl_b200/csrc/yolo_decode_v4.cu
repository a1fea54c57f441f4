// D3 — `yolo_forward_dynamic` (LightningFunc/utils/YoloV4Utils.py:36-176): one YOLOv4 head level, planar
// [B, A*(5+C), H, W], decoded into normalised corner boxes [B, A*H*W, 1, 4] and per-class confidences
// [B, A*H*W, C] = sigmoid(cls) * sigmoid(obj).
//
//   bxy = sigmoid(t_xy) * scale_x_y - 0.5 * (scale_x_y - 1)          (:84)
//   bwh = exp(t_wh)                                                   (:85)
//   bx  = (bxy.x + grid_x) / W,  bw = bwh.x * anchor_w / W            (:117-121, 147-148; anchors in grid units)
//   x1  = bx - bw * 0.5,  x2 = x1 + bw   (y likewise with H)          (:156-159)
//   confs = sigmoid(t_cls) * sigmoid(t_obj)                           (:86-87, 171)
//
// HBM-bound transpose, same scheme as decode_box_tile_kernel: one CTA owns `tc` consecutive cells of one (image, anchor)
// slab with all their 5+C planes; the planes are read along the cells (128-bit loads when the plane size allows), the
// activation is applied on the way into shared memory ([cell][field], odd row pitch), and the two results — for
// consecutive cells ONE contiguous block each — are written out linearly.
#include "common.cuh"

namespace b200det {

struct V4Params {
    const float* head;
    const float* anchors;     // [A,2] device, grid units
    float* boxes;  long long ld_boxes;    // row pitch in floats (4 for the reference's [B,N,1,4])
    float* confs;  long long ld_confs;    // row pitch in floats (C for [B,N,C])
    float* det;    long long ld_det;      // optional: sigmoid(obj) per candidate
    int A, C, H, W;
    float sxy, soff;          // scale_x_y and 0.5 * (scale_x_y - 1), both rounded to fp32 like the reference's Python scalars
    int tc, tc_shift;         // cells per CTA (power of two)
};

template <bool VEC>
__global__ void __launch_bounds__(256) yolo_forward_dynamic_kernel(const V4Params p) {
    extern __shared__ __align__(16) float s_v4[];            // [tc][ldF]
    const int GG = p.H * p.W, F = 5 + p.C;
    const int ldF = F | 1;
    const int ba = blockIdx.y;
    const int a = ba % p.A;
    const int cell0 = blockIdx.x << p.tc_shift;
    const int ncell = min(p.tc, GG - cell0);
    const float* src = p.head + (size_t)ba * F * GG + cell0;
    const int tid = threadIdx.x;

    if (VEC) {
        // thread -> (cell group g of 4 cells, planes f0, f0 + step, ...), three independent 128-bit loads in flight
        const int groups = p.tc >> 2;                        // <= 16
        const int g = tid & (groups - 1);
        const int step = 256 / groups;
        if ((g << 2) < ncell) {
            for (int f0 = tid / groups; f0 < F; f0 += 3 * step) {
                float4 ld[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + u * step;
                    if (f < F) ld[u] = ldg_stream4(src + (size_t)f * GG + (g << 2));
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + u * step;
                    if (f >= F) continue;
                    const float v[4] = {ld[u].x, ld[u].y, ld[u].z, ld[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float x = v[k];
                        s_v4[((g << 2) + k) * ldF + f] = (f == 2 || f == 3) ? expf(x) : sigmoidf_acc(x);
                    }
                }
            }
        }
    } else {
        const int c = tid & (p.tc - 1);
        const int step = 256 >> p.tc_shift;
        if (c < ncell) {
            for (int f0 = tid >> p.tc_shift; f0 < F; f0 += 3 * step) {
                float ld[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + u * step;
                    if (f < F) ld[u] = ldg_stream1(src + (size_t)f * GG + c);
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int f = f0 + u * step;
                    if (f >= F) continue;
                    s_v4[c * ldF + f] = (f == 2 || f == 3) ? expf(ld[u]) : sigmoidf_acc(ld[u]);
                }
            }
        }
    }
    __syncthreads();

    const size_t row0 = (size_t)ba * GG + cell0;             // candidate index of the CTA's first cell: (b*A + a)*H*W + cell
    // boxes: one thread per (cell, corner pair)
    for (int i = tid; i < ncell * 2; i += 256) {
        const int c = i >> 1, ax = i & 1;                    // ax 0: x, 1: y
        const int cell = cell0 + c;
        const int gy = cell / p.W;
        const float g = ax ? (float)gy : (float)(cell - gy * p.W);
        const float dim = ax ? (float)p.H : (float)p.W;
        const float an = p.anchors[a * 2 + ax];
        const float bxy = __fsub_rn(__fmul_rn(s_v4[c * ldF + ax], p.sxy), p.soff);
        const float ctr = __fdiv_rn(__fadd_rn(bxy, g), dim);
        const float ext = __fdiv_rn(__fmul_rn(s_v4[c * ldF + 2 + ax], an), dim);
        const float lo = __fsub_rn(ctr, __fmul_rn(ext, 0.5f));
        float* o = p.boxes + (row0 + c) * p.ld_boxes + ax;
        o[0] = lo;
        o[2] = __fadd_rn(lo, ext);
    }
    if (p.det) {
        for (int c = tid; c < ncell; c += 256) p.det[(row0 + c) * p.ld_det] = s_v4[c * ldF + 4];
    }
    // confs: linear over [cell][class]
    const int C = p.C;
    for (int i = tid; i < ncell * C; i += 256) {
        const int c = i / C, k = i - c * C;
        p.confs[(row0 + c) * p.ld_confs + k] = __fmul_rn(s_v4[c * ldF + 5 + k], s_v4[c * ldF + 4]);
    }
}

int yolo_forward_dynamic_launch(const float* head, int B, int A, int C, int H, int W, const float* anchors_dev, float scale_x_y,
                                float* boxes, long long ld_boxes, float* confs, long long ld_confs, float* det, long long ld_det,
                                cudaStream_t st) {
    V4Params p;
    p.head = head; p.anchors = anchors_dev;
    p.boxes = boxes; p.ld_boxes = ld_boxes; p.confs = confs; p.ld_confs = ld_confs; p.det = det; p.ld_det = ld_det;
    p.A = A; p.C = C; p.H = H; p.W = W;
    p.sxy = scale_x_y;
    p.soff = (float)(0.5 * ((double)scale_x_y - 1.0));
    const int F = 5 + C, ldF = F | 1, GG = H * W;
    int tc = 64;
    while (tc > 4 && (size_t)tc * ldF * sizeof(float) > 96 * 1024) tc >>= 1;   // 4 cells x 4101 fields = 64 KB at the class limit
    p.tc = tc;
    p.tc_shift = 0;
    while ((1 << p.tc_shift) < tc) ++p.tc_shift;
    const size_t smem = (size_t)tc * ldF * sizeof(float);
    const bool vec = (GG & 3) == 0 && ((uintptr_t)head & 15) == 0;
    dim3 grid(ceil_div(GG, tc), B * A);
    if (vec) {
        B2_CUDA(cudaFuncSetAttribute(yolo_forward_dynamic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        yolo_forward_dynamic_kernel<true><<<grid, 256, smem, st>>>(p);
    } else {
        B2_CUDA(cudaFuncSetAttribute(yolo_forward_dynamic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        yolo_forward_dynamic_kernel<false><<<grid, 256, smem, st>>>(p);
    }
    B2_LAUNCH_CHECK("yolo_forward_dynamic_kernel");
    return 0;
}

}  // namespace b200det
