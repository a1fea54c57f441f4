/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's YOLO test-time merge-NMS
 * (model/YOLOV3.py:304-335 == YOLOV5.py:187-216, with xywh2xyxy accuracy.py:289-295 and bbox_iou
 * accuracy.py:39-69).  It exists so that the CUDA path can be checked at the FULL benchmark size
 * (25 200 candidates x 64 images), where the torch port in oracle/ref_port.py needs ~15 s per image.
 * It is validated against that port and against the golden vectors of the unmodified reference in
 * tests/test_oracle_golden.py.  Never linked or called by the product package.
 *
 * Arithmetic: fp32 throughout, one rounding per operation in the reference's operation order
 * (compile with -ffp-contract=off, no -ffast-math), IEEE division, thresholds as fp32.
 * Ties in the score are ordered by ascending candidate index (the build's published rule).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t idx; } keyed_t;

static int cmp_desc(const void* a, const void* b) {
    const keyed_t* x = (const keyed_t*)a;
    const keyed_t* y = (const keyed_t*)b;
    const int xn = x->score != x->score, yn = y->score != y->score;   /* NaN last */
    if (xn != yn) return xn - yn;
    if (!xn) {
        if (x->score > y->score) return -1;
        if (x->score < y->score) return 1;
    }
    return (x->idx > y->idx) - (x->idx < y->idx);                     /* ties: ascending candidate index */
}

/* One image.  rows: [N][F] = (cx, cy, w, h, conf, cls[0..C-1]).  out: [N][7], out_idx: [N].
 * Returns the number of kept rows (descending score order). */
int yolo_nms_oracle_image(const float* rows, int N, int F, float conf_thres, float nms_thres, float* out,
                          int32_t* out_idx) {
    const int C = F - 5;
    keyed_t* order = (keyed_t*)malloc(sizeof(keyed_t) * (size_t)(N > 0 ? N : 1));
    float* det = (float*)malloc(sizeof(float) * 7 * (size_t)(N > 0 ? N : 1));
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    int n = 0;
    for (int i = 0; i < N; ++i) {
        const float* r = rows + (size_t)i * F;
        if (!(r[4] >= conf_thres)) continue;                          /* YOLOV3.py:310 */
        float best = r[5];
        for (int c = 1; c < C; ++c)                                   /* torch.max(1)[0] */
            if (!(r[5 + c] <= best) && best == best) best = r[5 + c];
        order[n].score = r[4] * best;                                 /* YOLOV3.py:315 */
        order[n].idx = i;
        ++n;
    }
    qsort(order, (size_t)n, sizeof(keyed_t), cmp_desc);               /* YOLOV3.py:317 */
    for (int k = 0; k < n; ++k) {                                     /* YOLOV3.py:305, 318-319 */
        const int i = order[k].idx;
        const float* r = rows + (size_t)i * F;
        float* d = det + (size_t)k * 7;
        float best = r[5];
        int bi = 0;
        for (int c = 1; c < C; ++c)
            if (!(r[5 + c] <= best) && best == best) { best = r[5 + c]; bi = c; }
        const float hw = r[2] / 2, hh = r[3] / 2;                     /* accuracy.py:291-294 */
        d[0] = r[0] - hw; d[1] = r[1] - hh; d[2] = r[0] + hw; d[3] = r[1] + hh;
        d[4] = r[4]; d[5] = best; d[6] = (float)bi;
        cand[k] = i;
    }
    int K = 0;
    while (n > 0) {                                                   /* YOLOV3.py:322-331 */
        const float tx1 = det[0], ty1 = det[1], tx2 = det[2], ty2 = det[3];
        const float tconf = det[4], tcc = det[5], tl = det[6];
        const int32_t tidx = cand[0];
        const float ta = (tx2 - tx1 + 1) * (ty2 - ty1 + 1);
        float sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, sw = 0;
        int m = 0, first = 1;
        for (int k = 0; k < n; ++k) {
            float* d = det + (size_t)k * 7;
            const float ix1 = tx1 > d[0] ? tx1 : d[0], iy1 = ty1 > d[1] ? ty1 : d[1];
            const float ix2 = tx2 < d[2] ? tx2 : d[2], iy2 = ty2 < d[3] ? ty2 : d[3];
            float iw = ix2 - ix1 + 1, ih = iy2 - iy1 + 1;             /* accuracy.py:60-62 */
            iw = iw > 0 ? iw : 0; ih = ih > 0 ? ih : 0;
            const float inter = iw * ih;
            const float a2 = (d[2] - d[0] + 1) * (d[3] - d[1] + 1);
            const float iou = inter / (ta + a2 - inter + 1e-16f);     /* accuracy.py:66-68 */
            if ((iou > nms_thres) && (d[6] == tl)) {                  /* YOLOV3.py:323-326 */
                const float w = d[4];                                 /* YOLOV3.py:327-329 */
                if (first) { sx1 = w * d[0]; sy1 = w * d[1]; sx2 = w * d[2]; sy2 = w * d[3]; sw = w; first = 0; }
                else { sx1 += w * d[0]; sy1 += w * d[1]; sx2 += w * d[2]; sy2 += w * d[3]; sw += w; }
            } else {
                if (m != k) { memcpy(det + (size_t)m * 7, d, sizeof(float) * 7); cand[m] = cand[k]; }
                ++m;
            }
        }
        if (m == n) break;   /* top row does not hit itself (w <= -1 or h <= -1): the reference never terminates */
        float* o = out + (size_t)K * 7;
        o[0] = sx1 / sw; o[1] = sy1 / sw; o[2] = sx2 / sw; o[3] = sy2 / sw;
        o[4] = tconf; o[5] = tcc; o[6] = tl;
        out_idx[K] = tidx;
        ++K;
        n = m;
    }
    free(order); free(det); free(cand);
    return K;
}

/* Batch entry: rows [B][N][F]; out [B][N][7]; out_idx [B][N]; out_count [B].  Images are independent. */
void yolo_nms_oracle_batch(const float* rows, int B, int N, int F, float conf_thres, float nms_thres, float* out,
                           int32_t* out_idx, int32_t* out_count) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        out_count[b] = yolo_nms_oracle_image(rows + (size_t)b * N * F, N, F, conf_thres, nms_thres,
                                             out + (size_t)b * N * 7, out_idx + (size_t)b * N);
}
