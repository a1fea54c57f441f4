// Device-side box arithmetic shared by the elementwise kernels and the target-assignment kernels.
#pragma once
#include "common.cuh"

namespace b200det {

__device__ __forceinline__ float4 cxcywh_to_corners(const float4 v) {   // accuracy.py:45-48 / :80-83 / :289-295
    const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
    return make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
}

// accuracy.py:54-68 on corner boxes
__device__ __forceinline__ float iou_plus1_eps(const float4 a, const float4 b) {
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
    const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float iw = fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.0f), 0.0f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.0f), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float a1 = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.0f), __fadd_rn(__fsub_rn(a.w, a.y), 1.0f));
    const float a2 = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-16f));
}

// the same value with the two +1-areas supplied by the caller and without the IEEE division when the boxes do not touch
// (0 / den is +-0 for any non-zero, non-NaN den; the sign of a zero is invisible to the comparisons that consume IoUs)
__device__ __forceinline__ float box_area_plus1(const float4 a) {
    return __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.0f), __fadd_rn(__fsub_rn(a.w, a.y), 1.0f));
}
__device__ __forceinline__ float iou_plus1_eps_pre(const float4 a, const float4 b, const float a1, const float a2) {
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
    const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float iw = fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.0f), 0.0f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.0f), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-16f);
    if (inter == 0.0f && (den > 0.0f || den < 0.0f)) return 0.0f;
    return __fdiv_rn(inter, den);
}

// accuracy.py:19-32 on corner boxes
__device__ __forceinline__ float iou_plain(const float4 a, const float4 b) {
    const float dx = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float dy = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(dx, dy);
    const float a1 = __fmul_rn(fmaxf(__fsub_rn(a.z, a.x), 0.0f), fmaxf(__fsub_rn(a.w, a.y), 0.0f));
    const float a2 = __fmul_rn(fmaxf(__fsub_rn(b.z, b.x), 0.0f), fmaxf(__fsub_rn(b.w, b.y), 0.0f));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

// accuracy.py:71-114.  `p`,`t` as given by the caller (corner or centre-size).
__device__ __forceinline__ float iou_v5_forward(float4 p, float4 t, bool corner, int kind) {
    if (!corner) { p = cxcywh_to_corners(p); t = cxcywh_to_corners(t); }
    const float iw = fmaxf(__fsub_rn(fminf(p.z, t.z), fmaxf(p.x, t.x)), 0.0f);
    const float ih = fmaxf(__fsub_rn(fminf(p.w, t.w), fmaxf(p.y, t.y)), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float w1 = __fsub_rn(p.z, p.x), h1 = __fsub_rn(p.w, p.y);
    const float w2 = __fsub_rn(t.z, t.x), h2 = __fsub_rn(t.w, t.y);
    const float uni = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, h1), 1e-16f), __fmul_rn(w2, h2)), inter);
    const float iou = __fdiv_rn(inter, uni);
    if (kind == B200DET_IOU) return iou;
    const float cw = __fsub_rn(fmaxf(p.z, t.z), fminf(p.x, t.x));
    const float ch = __fsub_rn(fmaxf(p.w, t.w), fminf(p.y, t.y));
    if (kind == B200DET_GIOU) {
        const float ca = __fadd_rn(__fmul_rn(cw, ch), 1e-16f);
        return __fsub_rn(iou, __fdiv_rn(__fsub_rn(ca, uni), ca));
    }
    const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)), 1e-16f);
    const float sx = __fsub_rn(__fadd_rn(t.x, t.z), __fadd_rn(p.x, p.z));
    const float sy = __fsub_rn(__fadd_rn(t.y, t.w), __fadd_rn(p.y, p.w));
    const float rho2 = __fadd_rn(__fmul_rn(__fmul_rn(sx, sx), 0.25f), __fmul_rn(__fmul_rn(sy, sy), 0.25f));
    if (kind == B200DET_DIOU) return __fsub_rn(iou, __fdiv_rn(rho2, c2));
    const float dt = __fsub_rn(atanf(__fdiv_rn(w2, h2)), atanf(__fdiv_rn(w1, h1)));
    const float v = __fmul_rn(0.40528473456935109f, __fmul_rn(dt, dt));
    const float alpha = __fdiv_rn(v, __fadd_rn(__fsub_rn(1.0f, iou), v));
    return __fsub_rn(iou, __fadd_rn(__fdiv_rn(rho2, c2), __fmul_rn(v, alpha)));
}

// d(out)/d(p) * go.  Sub-gradient conventions follow torch autograd: clamp(0) passes gradient when
// x >= 0, elementwise max/min split the gradient evenly on ties.
__device__ __forceinline__ void iou_v5_backward(float4 pin, float4 tin, bool corner, int kind, float go, float g[4]) {
    float4 p = pin, t = tin;
    if (!corner) { p = cxcywh_to_corners(p); t = cxcywh_to_corners(t); }
    auto gt = [](float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); };   // d max(a,b)/da
    auto lt = [](float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); };   // d min(a,b)/da
    const float dx = fminf(p.z, t.z) - fmaxf(p.x, t.x);
    const float dy = fminf(p.w, t.w) - fmaxf(p.y, t.y);
    const float iw = fmaxf(dx, 0.0f), ih = fmaxf(dy, 0.0f);
    const float mx = dx >= 0.0f ? 1.0f : 0.0f, my = dy >= 0.0f ? 1.0f : 0.0f;
    const float inter = iw * ih;
    const float w1 = p.z - p.x, h1 = p.w - p.y, w2 = t.z - t.x, h2 = t.w - t.y;
    const float uni = (w1 * h1 + 1e-16f) + w2 * h2 - inter;
    const float iou = inter / uni;
    // derivatives w.r.t. the corners (x1, y1, x2, y2) of p
    float di[4] = {-ih * mx * gt(p.x, t.x), -iw * my * gt(p.y, t.y), ih * mx * lt(p.z, t.z), iw * my * lt(p.w, t.w)};
    float du[4] = {-h1 - di[0], -w1 - di[1], h1 - di[2], w1 - di[3]};
    float gc[4];
    const float inv_u2 = 1.0f / (uni * uni);
    for (int k = 0; k < 4; ++k) gc[k] = (di[k] * uni - inter * du[k]) * inv_u2;          // d iou
    if (kind != B200DET_IOU) {
        const float cw = fmaxf(p.z, t.z) - fminf(p.x, t.x);
        const float ch = fmaxf(p.w, t.w) - fminf(p.y, t.y);
        const float dcw[4] = {-lt(p.x, t.x), 0.0f, gt(p.z, t.z), 0.0f};
        const float dch[4] = {0.0f, -lt(p.y, t.y), 0.0f, gt(p.w, t.w)};
        if (kind == B200DET_GIOU) {
            const float ca = cw * ch + 1e-16f;
            const float inv_c2 = 1.0f / (ca * ca);
            for (int k = 0; k < 4; ++k) {
                const float dca = dcw[k] * ch + cw * dch[k];
                gc[k] += (du[k] * ca - uni * dca) * inv_c2;                                 // d(union / c_area)
            }
        } else {
            const float c2 = cw * cw + ch * ch + 1e-16f;
            const float sx = (t.x + t.z) - (p.x + p.z), sy = (t.y + t.w) - (p.y + p.w);
            const float rho2 = sx * sx * 0.25f + sy * sy * 0.25f;
            const float drho[4] = {-0.5f * sx, -0.5f * sy, -0.5f * sx, -0.5f * sy};
            const float inv_c4 = 1.0f / (c2 * c2);
            float dv[4] = {0.f, 0.f, 0.f, 0.f};
            float alpha = 0.f;
            if (kind == B200DET_CIOU) {
                const float dt = atanf(w2 / h2) - atanf(w1 / h1);
                const float kk = 0.40528473456935109f;
                const float v = kk * dt * dt;
                alpha = v / (1.0f - iou + v);
                const float den = w1 * w1 + h1 * h1;
                const float dv_dw1 = -2.0f * kk * dt * h1 / den;
                const float dv_dh1 = 2.0f * kk * dt * w1 / den;
                dv[0] = -dv_dw1; dv[2] = dv_dw1; dv[1] = -dv_dh1; dv[3] = dv_dh1;
            }
            for (int k = 0; k < 4; ++k) {
                const float dc2 = 2.0f * cw * dcw[k] + 2.0f * ch * dch[k];
                gc[k] -= (drho[k] * c2 - rho2 * dc2) * inv_c4 + alpha * dv[k];
            }
        }
    }
    if (corner) {
        for (int k = 0; k < 4; ++k) g[k] = go * gc[k];
    } else {
        g[0] = go * (gc[0] + gc[2]);
        g[1] = go * (gc[1] + gc[3]);
        g[2] = go * 0.5f * (gc[2] - gc[0]);
        g[3] = go * 0.5f * (gc[3] - gc[1]);
    }
}

}  // namespace b200det
