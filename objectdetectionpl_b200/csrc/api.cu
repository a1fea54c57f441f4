// C ABI of libb200det.so (include/b200det.h): argument checks + dispatch to the stage launchers.
#include "yolo_ws.cuh"
#include "targets.cuh"

namespace b200det {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

// launchers defined in the stage files
int yolo_stage_reset(const b200det_yolo_desc*, void*, size_t, cudaStream_t);
int yolo_stage_decode(const b200det_yolo_desc*, void*, size_t, cudaStream_t);
int yolo_stage_sort(const b200det_yolo_desc*, void*, size_t, cudaStream_t);
int yolo_stage_nms(const b200det_yolo_desc*, void*, size_t, cudaStream_t);
int yolo_stage_emit(const b200det_yolo_desc*, void*, size_t, float*, int32_t*, int32_t*, int32_t*, int32_t*, cudaEvent_t,
                    cudaStream_t);
int decode_box_launch(const float*, int, int, int, int, int, const float*, float, float*, cudaStream_t);
int yolo_forward_dynamic_launch(const float*, int, int, int, int, int, const float*, float, float*, long long, float*, long long,
                                float*, long long, cudaStream_t);
size_t prior_workspace_bytes(const b200det_prior_desc*);
int prior_nms_pipeline(const b200det_prior_desc*, void*, size_t, float*, int32_t*, int32_t*, int32_t*, cudaStream_t);
int prior_stage_decode(const b200det_prior_desc*, void*, size_t, cudaStream_t);
int prior_stage_select_nms(const b200det_prior_desc*, void*, size_t, float*, int32_t*, int32_t*, int32_t*, cudaStream_t);
int xywh2xyxy_launch(const float*, float*, long long, cudaStream_t);
int bbox_iou_plus1_launch(const float*, long long, const float*, long long, int, float*, cudaStream_t);
int pair_iou_launch(const float*, const float*, long long, float*, cudaStream_t);
int bbox_iou_v5_fwd_launch(const float*, long long, long long, const float*, long long, long long, long long, int, int,
                           float*, cudaStream_t);
int bbox_iou_v5_bwd_launch(const float*, long long, long long, const float*, long long, long long, long long, int, int,
                           const float*, float*, cudaStream_t);
int build_targets_v5_launch(const float*, int, const float*, int, int, int, int32_t*, int32_t*, int32_t*, int32_t*,
                            int32_t*, float*, float*, int32_t*, cudaStream_t);
int build_targets_v5_multi_launch(const float*, int, int, const float*, int, const int32_t*, const int32_t*, int32_t* const*,
                                  int32_t* const*, int32_t* const*, int32_t* const*, int32_t* const*, float* const*,
                                  float* const*, int32_t*, cudaStream_t);
int v5_match_fwd_launch(const float*, int, int, int, int, int, const int32_t*, const int32_t*, const int32_t*,
                        const int32_t*, const float*, const float*, int, float*, float*, cudaStream_t);
int v5_match_bwd_launch(const float*, int, int, int, int, int, const int32_t*, const int32_t*, const int32_t*,
                        const int32_t*, const float*, const float*, int, const float*, float*, cudaStream_t);
int v5_loss_fwd_launch(const float*, int, int, int, int, int, const int32_t*, const int32_t*, const int32_t*, const int32_t*,
                       const int32_t*, const float*, const float*, int, float, float, float, float, int, float*, float*,
                       double*, const int32_t*, float*, cudaStream_t);
int v5_loss_bwd_launch(const float*, int, int, int, int, int, const int32_t*, const int32_t*, const int32_t*, const int32_t*,
                       const int32_t*, const float*, const float*, int, float, float, float, float, int, const float*,
                       const float*, float, float, float, float*, int, const int32_t*, const float*, cudaStream_t);
int v5_loss_combine_launch(const double*, int, float, float, float, float*, cudaStream_t);
int v5_loss_combine_bwd_launch(const float*, const float*, const float*, const float*, float, float, float, float*, cudaStream_t);
size_t build_targets_ws_bytes(int, int, int, int);
int build_targets_launch(const float*, const float*, const float*, const float*, int, int, int, int, int, float, void*,
                         float*, float*, uint8_t*, uint8_t*, float*, float*, float*, float*, float*, int32_t*, cudaStream_t);
size_t ssd_match_ws_bytes(int, int);
int ssd_match_launch(const float*, int, const float*, int, float, void*, int32_t*, uint8_t*, cudaStream_t);
size_t retina_assign_ws_bytes(int, int);
int retina_assign_launch(const float*, int, const float*, int, int, float, void*, float*, int32_t*, cudaStream_t);

int pack_detections_launch(const float*, const int32_t*, int, long long, int, int, float*, long long, int32_t*, cudaStream_t);
size_t batch_statistics_ws_bytes(int, int);
int batch_statistics_launch(const float*, const long long*, const int*, int, int, const float*, int, float, void*, float*,
                            cudaStream_t);
size_t yolo_statistics_ws_bytes(int, int, int, int);
int yolo_statistics_launch(const float*, int, int, int, int, const float*, float, const float*, int, float, void*, float*,
                           float*, cudaStream_t);
size_t ap_per_class_ws_bytes(int);
int ap_per_class_launch(const float*, const float*, const float*, int, const int*, const int*, int, void*, double*, double*,
                        double*, double*, cudaStream_t);

}  // namespace b200det

using namespace b200det;

extern "C" {

int b200det_version(void) { return B200DET_VERSION; }
const char* b200det_last_error(void) { return g_err; }

int b200det_yolo_num_candidates(const b200det_yolo_desc* d, int32_t* n, int32_t* n_pad) {
    B2_CHECK_ARG(d && n && n_pad, "null argument");
    int a = 0, b = 0;
    B2_CHECK_LIMIT(yolo_counts(d, &a, &b) == 0, "candidates per image out of (0, %d]", B200DET_MAX_CANDIDATES);
    *n = a; *n_pad = b;
    return 0;
}

size_t b200det_yolo_workspace_bytes(const b200det_yolo_desc* d) {
    if (!d) return 0;
    int a = 0, b = 0;
    if (yolo_counts(d, &a, &b) != 0 || d->batch <= 0 || d->num_classes <= 0) return 0;
    YoloWs w;
    yolo_ws_layout(d, nullptr, &w);
    return w.total_bytes;
}

int b200det_yolo_stage_reset(const b200det_yolo_desc* d, void* ws, size_t n, void* st) {
    return yolo_stage_reset(d, ws, n, (cudaStream_t)st);
}
int b200det_yolo_stage_decode(const b200det_yolo_desc* d, void* ws, size_t n, void* st) {
    return yolo_stage_decode(d, ws, n, (cudaStream_t)st);
}
int b200det_yolo_stage_sort(const b200det_yolo_desc* d, void* ws, size_t n, void* st) {
    return yolo_stage_sort(d, ws, n, (cudaStream_t)st);
}
int b200det_yolo_stage_nms(const b200det_yolo_desc* d, void* ws, size_t n, void* st) {
    return yolo_stage_nms(d, ws, n, (cudaStream_t)st);
}
int b200det_yolo_stage_emit(const b200det_yolo_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                            int32_t* out_count, void* st) {
    return yolo_stage_emit(d, ws, n, out_rows, out_index, out_count, nullptr, nullptr, nullptr, (cudaStream_t)st);
}
int b200det_yolo_stage_emit_packed(const b200det_yolo_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                                   int32_t* out_count, int32_t* out_offsets, void* st) {
    B2_CHECK_ARG(out_offsets != nullptr, "out_offsets is null");
    return yolo_stage_emit(d, ws, n, out_rows, out_index, out_count, out_offsets, nullptr, nullptr, (cudaStream_t)st);
}

static int yolo_pipeline(const b200det_yolo_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index, int32_t* out_count,
                         int32_t* out_offsets, int32_t* counts_early, void* counts_ready, void* st) {
    int rc = yolo_stage_reset(d, ws, n, (cudaStream_t)st);
    if (rc) return rc;
    rc = yolo_stage_decode(d, ws, n, (cudaStream_t)st);
    if (rc) return rc;
    rc = yolo_stage_sort(d, ws, n, (cudaStream_t)st);
    if (rc) return rc;
    rc = yolo_stage_nms(d, ws, n, (cudaStream_t)st);
    if (rc) return rc;
    return yolo_stage_emit(d, ws, n, out_rows, out_index, out_count, out_offsets, counts_early, (cudaEvent_t)counts_ready,
                           (cudaStream_t)st);
}

int b200det_yolo_nms_packed(const b200det_yolo_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                            int32_t* out_count, int32_t* out_offsets, int32_t* counts_early, void* counts_ready_event, void* st) {
    B2_CHECK_ARG(out_offsets != nullptr, "out_offsets is null");
    return yolo_pipeline(d, ws, n, out_rows, out_index, out_count, out_offsets, counts_early, counts_ready_event, st);
}

int b200det_yolo_nms_early(const b200det_yolo_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                           int32_t* out_count, int32_t* counts_early, void* counts_ready_event, void* st) {
    return yolo_pipeline(d, ws, n, out_rows, out_index, out_count, nullptr, counts_early, counts_ready_event, st);
}

int b200det_yolo_nms(const b200det_yolo_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                     int32_t* out_count, void* st) {
    return yolo_pipeline(d, ws, n, out_rows, out_index, out_count, nullptr, nullptr, nullptr, st);
}

int b200det_yolo_workspace_field(const b200det_yolo_desc* d, const char* name, size_t* offset, size_t* bytes) {
    B2_CHECK_ARG(d && name && offset && bytes, "null argument");
    int a = 0, b = 0;
    B2_CHECK_LIMIT(yolo_counts(d, &a, &b) == 0, "bad descriptor");
    YoloWs w;
    yolo_ws_layout(d, nullptr, &w);
    const size_t B = (size_t)w.B, P = (size_t)w.n_pad, C = (size_t)w.C;
    struct F { const char* n; const void* p; size_t bytes; };
    const F fields[] = {
        {"count", w.count, B * 4}, {"cls_hist", w.cls_hist, B * C * 4}, {"seg_off", w.seg_off, B * (C + 1) * 4},
        {"tile_count", w.tile_count, B * (size_t)w.n_tiles * 4}, {"box4", w.box4, B * P * 16}, {"cc2", w.cc2, B * P * 8},
        {"orig", w.orig, B * P * 4}, {"key", w.key[0], B * P * 4}, {"pay", w.pay[0], B * P * 4},
        {"sorted_pay", yolo_sorted_pay(w), B * P * 4}, {"sorted_rank", yolo_sorted_rank(w), B * P * 4},
        {"kpay", w.kpay, B * P * 4}, {"mbox", w.mbox, B * P * 16},
    };
    for (const F& f : fields) {
        if (strcmp(f.n, name) == 0) {
            *offset = (size_t)(uintptr_t)f.p;
            *bytes = f.bytes;
            return 0;
        }
    }
    set_error("unknown workspace field '%s'", name);
    return B200DET_EINVAL;
}

int b200det_decode_box(const float* head, int32_t B, int32_t A, int32_t C, int32_t G, int32_t mode,
                       const float* anchors_dev, float stride, float* out, void* st) {
    B2_CHECK_ARG(head && out, "head / out is null");
    B2_CHECK_ARG(B > 0 && A > 0 && C > 0 && G > 0, "B, A, C, G must be > 0");
    B2_CHECK_ARG(mode >= B200DET_DECODE_NONE && mode <= B200DET_DECODE_YOLOV5, "bad decode mode %d", mode);
    B2_CHECK_ARG(mode == B200DET_DECODE_NONE || anchors_dev, "anchors required for decode modes");
    B2_CHECK_LIMIT((long long)B * A <= 65535, "B*A %lld > 65535", (long long)B * A);
    return decode_box_launch(head, B, A, C, G, mode, anchors_dev, stride, out, (cudaStream_t)st);
}

int b200det_yolo_forward_dynamic(const float* head, int32_t B, int32_t A, int32_t C, int32_t H, int32_t W, const float* anchors_dev,
                                 float scale_x_y, float* boxes, int64_t ld_boxes, float* confs, int64_t ld_confs, float* det,
                                 int64_t ld_det, void* st) {
    B2_CHECK_ARG(head && anchors_dev && boxes && confs, "head / anchors / boxes / confs is null");
    B2_CHECK_ARG(B > 0 && A > 0 && C > 0 && H > 0 && W > 0, "B, A, C, H, W must be > 0");
    B2_CHECK_ARG(ld_boxes >= 4 && ld_confs >= C && (!det || ld_det >= 1), "row pitches too small");
    B2_CHECK_LIMIT((long long)B * A <= 65535, "B*A %lld > 65535", (long long)B * A);
    B2_CHECK_LIMIT(C <= B200DET_MAX_CLASSES, "num_classes %d > %d", C, B200DET_MAX_CLASSES);
    B2_CHECK_LIMIT((long long)H * W < (1ll << 30), "plane too large");
    return yolo_forward_dynamic_launch(head, B, A, C, H, W, anchors_dev, scale_x_y, boxes, (long long)ld_boxes, confs,
                                       (long long)ld_confs, det, (long long)ld_det, (cudaStream_t)st);
}

size_t b200det_prior_workspace_bytes(const b200det_prior_desc* d) {
    if (!d || d->batch <= 0 || d->num_priors <= 0) return 0;
    return prior_workspace_bytes(d);
}
int b200det_prior_nms(const b200det_prior_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                      int32_t* out_count, int32_t* cand_count, void* st) {
    return prior_nms_pipeline(d, ws, n, out_rows, out_index, out_count, cand_count, (cudaStream_t)st);
}

int b200det_prior_stage_decode(const b200det_prior_desc* d, void* ws, size_t n, void* st) {
    return prior_stage_decode(d, ws, n, (cudaStream_t)st);
}
int b200det_prior_stage_select_nms(const b200det_prior_desc* d, void* ws, size_t n, float* out_rows, int32_t* out_index,
                                   int32_t* out_count, int32_t* cand_count, void* st) {
    return prior_stage_select_nms(d, ws, n, out_rows, out_index, out_count, cand_count, (cudaStream_t)st);
}

int b200det_xywh2xyxy(const float* x, float* y, int64_t n, void* st) {
    B2_CHECK_ARG(n >= 0 && (n == 0 || (x && y)), "null argument");
    B2_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "x and y must be 16-byte aligned");
    return xywh2xyxy_launch(x, y, n, (cudaStream_t)st);
}
int b200det_bbox_iou_plus1(const float* box1, int64_t n1, const float* box2, int64_t n, int32_t corner, float* out, void* st) {
    B2_CHECK_ARG(n >= 0 && (n == 0 || (box1 && box2 && out)), "null argument");
    B2_CHECK_ARG(n1 == 1 || n1 == n, "box1 must hold 1 or n rows (got %lld vs %lld)", (long long)n1, (long long)n);
    B2_CHECK_ARG((((uintptr_t)box1 | (uintptr_t)box2) & 15) == 0, "boxes must be 16-byte aligned");
    return bbox_iou_plus1_launch(box1, n1, box2, n, corner, out, (cudaStream_t)st);
}
int b200det_pair_iou(const float* a, const float* b, int64_t n, float* out, void* st) {
    B2_CHECK_ARG(n >= 0 && (n == 0 || (a && b && out)), "null argument");
    B2_CHECK_ARG((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "boxes must be 16-byte aligned");
    return pair_iou_launch(a, b, n, out, (cudaStream_t)st);
}
int b200det_bbox_iou_v5_fwd(const float* box1, int64_t ld1, int64_t inc1, const float* box2, int64_t ld2, int64_t inc2,
                            int64_t n, int32_t corner, int32_t kind, float* out, void* st) {
    B2_CHECK_ARG(n >= 0 && (n == 0 || (box1 && box2 && out)), "null argument");
    B2_CHECK_ARG(kind >= B200DET_IOU && kind <= B200DET_CIOU, "bad IoU kind %d", kind);
    return bbox_iou_v5_fwd_launch(box1, ld1, inc1, box2, ld2, inc2, n, corner, kind, out, (cudaStream_t)st);
}
int b200det_bbox_iou_v5_bwd(const float* box1, int64_t ld1, int64_t inc1, const float* box2, int64_t ld2, int64_t inc2,
                            int64_t n, int32_t corner, int32_t kind, const float* grad_out, float* grad_box1, void* st) {
    B2_CHECK_ARG(n >= 0 && (n == 0 || (box1 && box2 && grad_out && grad_box1)), "null argument");
    B2_CHECK_ARG(kind >= B200DET_IOU && kind <= B200DET_CIOU, "bad IoU kind %d", kind);
    return bbox_iou_v5_bwd_launch(box1, ld1, inc1, box2, ld2, inc2, n, corner, kind, grad_out, grad_box1, (cudaStream_t)st);
}

int b200det_build_targets_v5_level(const float* targets, int32_t nt, const float* anchors_host, int32_t na, int32_t nx,
                                   int32_t ny, int32_t* ob, int32_t* oa, int32_t* ogj, int32_t* ogi, int32_t* ocls,
                                   float* otbox, float* oanch, int32_t* ocount, void* st) {
    B2_CHECK_ARG(nt >= 0 && na > 0 && nx > 0 && ny > 0, "bad sizes");
    B2_CHECK_LIMIT(na <= B200DET_MAX_ANCHORS, "num_anchors %d > %d", na, B200DET_MAX_ANCHORS);
    B2_CHECK_ARG(anchors_host && ocount, "null argument");
    B2_CHECK_ARG(nt == 0 || (targets && ob && oa && ogj && ogi && ocls && otbox && oanch), "null argument");
    B2_CHECK_LIMIT((long long)nt * na < (1ll << 30), "too many targets");
    return build_targets_v5_launch(targets, nt, anchors_host, na, nx, ny, ob, oa, ogj, ogi, ocls, otbox, oanch, ocount,
                                   (cudaStream_t)st);
}
int b200det_build_targets_v5(const float* targets, int32_t nt, int32_t nl, const float* anchors_host, int32_t na,
                             const int32_t* nx_host, const int32_t* ny_host, int32_t* const* ob, int32_t* const* oa,
                             int32_t* const* ogj, int32_t* const* ogi, int32_t* const* ocls, float* const* otbox,
                             float* const* oanch, int32_t* ocount, void* st) {
    B2_CHECK_ARG(nt >= 0 && na > 0 && nl > 0, "bad sizes");
    B2_CHECK_LIMIT(na <= B200DET_MAX_ANCHORS && nl <= B200DET_MAX_LEVELS, "num_anchors %d / levels %d beyond the limits", na, nl);
    B2_CHECK_ARG(anchors_host && nx_host && ny_host && ob && oa && ogj && ogi && ocls && otbox && oanch && ocount && (nt == 0 || targets),
                 "null argument");
    B2_CHECK_LIMIT((long long)nt * na < (1ll << 30), "too many targets");
    for (int l = 0; l < nl; ++l) B2_CHECK_ARG(nx_host[l] > 0 && ny_host[l] > 0, "bad grid of level %d", l);
    return build_targets_v5_multi_launch(targets, nt, nl, anchors_host, na, nx_host, ny_host, ob, oa, ogj, ogi, ocls, otbox,
                                         oanch, ocount, (cudaStream_t)st);
}
int b200det_v5_match_fwd(const float* pi, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t F, const int32_t* b,
                         const int32_t* a, const int32_t* gj, const int32_t* gi, const float* tbox, const float* anch,
                         int32_t m, float* giou, float* tobj, void* st) {
    B2_CHECK_ARG(m >= 0 && F >= 5, "bad sizes");
    B2_CHECK_ARG(m == 0 || (pi && b && a && gj && gi && tbox && anch && giou && tobj), "null argument");
    B2_CHECK_ARG(((uintptr_t)tbox & 15) == 0, "tbox must be 16-byte aligned");
    return v5_match_fwd_launch(pi, B, na, ny, nx, F, b, a, gj, gi, tbox, anch, m, giou, tobj, (cudaStream_t)st);
}
int b200det_v5_match_bwd(const float* pi, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t F, const int32_t* b,
                         const int32_t* a, const int32_t* gj, const int32_t* gi, const float* tbox, const float* anch,
                         int32_t m, const float* ggiou, float* gpi, void* st) {
    B2_CHECK_ARG(m >= 0 && F >= 5, "bad sizes");
    B2_CHECK_ARG(m == 0 || (pi && b && a && gj && gi && tbox && anch && ggiou && gpi), "null argument");
    B2_CHECK_ARG(((uintptr_t)tbox & 15) == 0, "tbox must be 16-byte aligned");
    return v5_match_bwd_launch(pi, B, na, ny, nx, F, b, a, gj, gi, tbox, anch, m, ggiou, gpi, (cudaStream_t)st);
}

int b200det_v5_loss_fwd(const float* pi, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t F, const int32_t* b,
                        const int32_t* a, const int32_t* gj, const int32_t* gi, const int32_t* tcls, const float* tbox,
                        const float* anch, int32_t m, float cp, float cn, float gamma, float alpha, int32_t with_cls,
                        float* giou, float* tobj, double* sums, void* st) {
    B2_CHECK_ARG(B > 0 && na > 0 && ny > 0 && nx > 0 && m >= 0 && F >= 5, "bad sizes");
    B2_CHECK_ARG(pi && tobj && sums && (m == 0 || (b && a && gj && gi && tcls && tbox && anch && giou)), "null argument");
    B2_CHECK_ARG(((uintptr_t)tbox & 15) == 0 && ((uintptr_t)sums & 7) == 0, "tbox must be 16-byte, sums 8-byte aligned");
    return v5_loss_fwd_launch(pi, B, na, ny, nx, F, b, a, gj, gi, tcls, tbox, anch, m, cp, cn, gamma, alpha, with_cls, giou,
                              tobj, sums, nullptr, nullptr, (cudaStream_t)st);
}
int b200det_v5_loss_bwd(const float* pi, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t F, const int32_t* b,
                        const int32_t* a, const int32_t* gj, const int32_t* gi, const int32_t* tcls, const float* tbox,
                        const float* anch, int32_t m, float cp, float cn, float gamma, float alpha, int32_t with_cls,
                        const float* tobj, const float* g3, float inv_nbox, float inv_cells, float inv_ncls, float* gpi,
                        void* st) {
    B2_CHECK_ARG(B > 0 && na > 0 && ny > 0 && nx > 0 && m >= 0 && F >= 5, "bad sizes");
    B2_CHECK_ARG(pi && tobj && gpi && g3 && (m == 0 || (b && a && gj && gi && tcls && tbox && anch)), "null argument");
    B2_CHECK_ARG(((uintptr_t)tbox & 15) == 0, "tbox must be 16-byte aligned");
    return v5_loss_bwd_launch(pi, B, na, ny, nx, F, b, a, gj, gi, tcls, tbox, anch, m, cp, cn, gamma, alpha, with_cls, tobj,
                              g3, inv_nbox, inv_cells, inv_ncls, gpi, 0, nullptr, nullptr, (cudaStream_t)st);
}
int b200det_v5_loss_bwd_full(const float* pi, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t F, const int32_t* b,
                             const int32_t* a, const int32_t* gj, const int32_t* gi, const int32_t* tcls, const float* tbox,
                             const float* anch, int32_t m, float cp, float cn, float gamma, float alpha, int32_t with_cls,
                             const float* tobj, const float* g3, float inv_nbox, float inv_cells, float inv_ncls, float* gpi,
                             void* st) {
    B2_CHECK_ARG(B > 0 && na > 0 && ny > 0 && nx > 0 && m >= 0 && F >= 5, "bad sizes");
    B2_CHECK_ARG(pi && tobj && gpi && g3 && (m == 0 || (b && a && gj && gi && tcls && tbox && anch)), "null argument");
    B2_CHECK_ARG(((uintptr_t)tbox & 15) == 0, "tbox must be 16-byte aligned");
    return v5_loss_bwd_launch(pi, B, na, ny, nx, F, b, a, gj, gi, tcls, tbox, anch, m, cp, cn, gamma, alpha, with_cls, tobj,
                              g3, inv_nbox, inv_cells, inv_ncls, gpi, 1, nullptr, nullptr, (cudaStream_t)st);
}
// ---- all levels of the criterion per call (targets.cu: V5Multi) ----
static int v5_fill_levels(V5Multi* mp, const b200det_v5_level* levels, int32_t nl, int32_t cap, bool backward) {
    B2_CHECK_ARG(levels && nl > 0 && nl <= kV5MaxLevels, "need 1..5 levels");
    B2_CHECK_ARG(cap > 0, "bad row capacity");
    memset(mp, 0, sizeof(*mp));
    mp->nl = nl;
    for (int i = 0; i < nl; ++i) {
        const b200det_v5_level& l = levels[i];
        B2_CHECK_ARG(l.batch > 0 && l.na > 0 && l.ny > 0 && l.nx > 0 && l.fields >= 5, "bad level sizes");
        B2_CHECK_ARG(l.pi && l.b && l.a && l.gj && l.gi && l.tcls && l.tbox && l.anch && l.m_dev, "null level argument");
        B2_CHECK_ARG(((uintptr_t)l.tbox & 15) == 0, "tbox must be 16-byte aligned");
        B2_CHECK_ARG(!backward || (l.gpi && ((uintptr_t)l.gpi & 15) == 0), "gpi must be non-null and 16-byte aligned");
        V5Level& L = mp->lv[i];
        L.p = MatchParams{l.pi, l.batch, l.na, l.ny, l.nx, l.fields, l.b, l.a, l.gj, l.gi, l.tbox, l.anch, cap, l.m_dev};
        L.tcls = l.tcls;
        L.gpi = l.gpi;
        L.cells = (long long)l.batch * l.na * l.ny * l.nx;
    }
    return 0;
}
int b200det_v5_loss_fwd_all(const b200det_v5_level* levels, int32_t nl, int32_t cap, float cp, float cn, float gamma, float alpha,
                            int32_t with_cls, float wbox, float wobj, float wcls, float* giou, float* tobj, float* obj_grad,
                            double* sums, float* out4, void* st) {
    V5Multi mp;
    int rc = v5_fill_levels(&mp, levels, nl, cap, false);
    if (rc) return rc;
    B2_CHECK_ARG(giou && tobj && sums && out4, "null argument");
    B2_CHECK_ARG(((uintptr_t)sums & 7) == 0, "sums must be 8-byte aligned");
    mp.cp = cp; mp.cn = cn; mp.gamma = gamma; mp.alpha = alpha; mp.with_cls = with_cls;
    mp.wbox = wbox; mp.wobj = wobj; mp.wcls = wcls;
    return v5_loss_fwd_all_launch(mp, cap, giou, tobj, obj_grad, sums, out4, (cudaStream_t)st);
}
int b200det_v5_loss_bwd_all(const b200det_v5_level* levels, int32_t nl, int32_t cap, float cp, float cn, float gamma, float alpha,
                            int32_t with_cls, float wbox, float wobj, float wcls, const float* obj_grad, const float* g_loss,
                            const float* g_box, const float* g_cls, const float* g_obj, float* g3, void* st) {
    V5Multi mp;
    int rc = v5_fill_levels(&mp, levels, nl, cap, true);
    if (rc) return rc;
    B2_CHECK_ARG(obj_grad && g3, "null argument");
    mp.cp = cp; mp.cn = cn; mp.gamma = gamma; mp.alpha = alpha; mp.with_cls = with_cls;
    mp.wbox = wbox; mp.wobj = wobj; mp.wcls = wcls;
    return v5_loss_bwd_all_launch(mp, cap, obj_grad, g_loss, g_box, g_cls, g_obj, g3, (cudaStream_t)st);
}

int b200det_v5_loss_combine(const double* means, int32_t nl, float wbox, float wobj, float wcls, float* out4, void* st) {
    B2_CHECK_ARG(means && out4 && nl > 0, "null argument / bad level count");
    return v5_loss_combine_launch(means, nl, wbox, wobj, wcls, out4, (cudaStream_t)st);
}
int b200det_v5_loss_combine_bwd(const float* g_loss, const float* g_box, const float* g_cls, const float* g_obj, float wbox,
                                float wobj, float wcls, float* g3, void* st) {
    B2_CHECK_ARG(g3 != nullptr, "null argument");
    return v5_loss_combine_bwd_launch(g_loss, g_box, g_cls, g_obj, wbox, wobj, wcls, g3, (cudaStream_t)st);
}

size_t b200det_build_targets_workspace_bytes(int32_t B, int32_t A, int32_t G, int32_t nt) {
    if (B <= 0 || A <= 0 || G <= 0 || nt < 0) return 0;
    return build_targets_ws_bytes(B, A, G, nt);
}
int b200det_build_targets(const float* pred_boxes, const float* pred_cls, const float* target, const float* anchors,
                          int32_t B, int32_t A, int32_t G, int32_t C, int32_t nt, float ignore_thres, void* ws,
                          size_t ws_bytes, float* iou_scores, float* class_mask, uint8_t* obj_mask, uint8_t* noobj_mask,
                          float* tx, float* ty, float* tw, float* th, float* tcls, int32_t* status, void* st) {
    B2_CHECK_ARG(B > 0 && A > 0 && G > 0 && C > 0 && nt >= 0, "bad sizes");
    B2_CHECK_ARG(pred_boxes && pred_cls && anchors && ws && iou_scores && class_mask && obj_mask && noobj_mask && tx &&
                     ty && tw && th && tcls && status && (nt == 0 || target), "null argument");
    B2_CHECK_ARG(((uintptr_t)pred_boxes & 15) == 0, "pred_boxes must be 16-byte aligned");
    if (ws_bytes < build_targets_ws_bytes(B, A, G, nt)) {
        set_error("workspace too small");
        return B200DET_EWORKSPACE;
    }
    return build_targets_launch(pred_boxes, pred_cls, target, anchors, B, A, G, C, nt, ignore_thres, ws, iou_scores,
                                class_mask, obj_mask, noobj_mask, tx, ty, tw, th, tcls, status, (cudaStream_t)st);
}

size_t b200det_ssd_match_workspace_bytes(int32_t P, int32_t M) {
    if (P <= 0 || M < 0) return 0;
    return ssd_match_ws_bytes(P, M);
}
int b200det_ssd_match(const float* priors, int32_t P, const float* gt, int32_t M, float thresh, void* ws, size_t ws_bytes,
                      int32_t* idx, uint8_t* matched, void* st) {
    B2_CHECK_ARG(P > 0 && M >= 0 && priors && ws && idx && matched && (M == 0 || gt), "bad argument");
    B2_CHECK_ARG((((uintptr_t)priors | (uintptr_t)gt) & 15) == 0, "priors / gt must be 16-byte aligned");
    if (ws_bytes < ssd_match_ws_bytes(P, M)) {
        set_error("workspace too small");
        return B200DET_EWORKSPACE;
    }
    return ssd_match_launch(priors, P, gt, M, thresh, ws, idx, matched, (cudaStream_t)st);
}

size_t b200det_retina_assign_workspace_bytes(int32_t B, int32_t nt) {
    if (B <= 0 || nt < 0) return 0;
    return retina_assign_ws_bytes(B, nt);
}
int b200det_retina_assign(const float* anchors, int32_t A, const float* targets, int32_t nt, int32_t B, float img_size,
                          void* ws, size_t ws_bytes, float* loc, int32_t* cls, void* st) {
    B2_CHECK_ARG(A > 0 && B > 0 && nt >= 0 && anchors && ws && loc && cls && (nt == 0 || targets), "bad argument");
    B2_CHECK_LIMIT(B <= 65535, "batch %d > 65535", B);
    B2_CHECK_ARG((((uintptr_t)anchors | (uintptr_t)loc) & 15) == 0, "anchors / loc_targets must be 16-byte aligned");
    if (ws_bytes < retina_assign_ws_bytes(B, nt)) {
        set_error("workspace too small");
        return B200DET_EWORKSPACE;
    }
    return retina_assign_launch(anchors, A, targets, nt, B, img_size, ws, loc, cls, (cudaStream_t)st);
}

int b200det_pack_detections(const float* rows, const int32_t* count, int32_t B, int64_t row_pitch, int32_t max_rows,
                            int32_t image_offset, float* out, int64_t cap, int32_t* offsets, void* st) {
    B2_CHECK_ARG(B > 0 && row_pitch >= 0 && max_rows >= 0 && cap >= 0, "bad sizes");
    B2_CHECK_LIMIT(B <= 65535, "batch %d > 65535", B);
    B2_CHECK_ARG(count && (max_rows == 0 || (rows && out)), "null argument");
    B2_CHECK_ARG(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    return pack_detections_launch(rows, count, B, (long long)row_pitch, max_rows, image_offset, out, (long long)cap, offsets,
                                  (cudaStream_t)st);
}

size_t b200det_batch_statistics_workspace_bytes(int32_t B, int32_t nt) {
    if (B <= 0 || nt < 0) return 0;
    return batch_statistics_ws_bytes(B, nt);
}
int b200det_batch_statistics(const float* rows, const int64_t* row_start, const int32_t* count, int32_t B, int32_t max_count,
                             const float* targets, int32_t nt, float thr, void* ws, size_t ws_bytes, float* tp, void* st) {
    B2_CHECK_ARG(B > 0 && nt >= 0 && max_count >= 0, "bad sizes");
    B2_CHECK_LIMIT(B <= 65535, "batch %d > 65535", B);
    B2_CHECK_ARG(row_start && count && ws && (nt == 0 || targets) && (max_count == 0 || (rows && tp)), "null argument");
    if (ws_bytes < batch_statistics_ws_bytes(B, nt)) {
        set_error("workspace too small");
        return B200DET_EWORKSPACE;
    }
    return batch_statistics_launch(rows, (const long long*)row_start, count, B, max_count, targets, nt, thr, ws, tp,
                                   (cudaStream_t)st);
}

size_t b200det_yolo_statistics_workspace_bytes(int32_t B, int32_t A, int32_t G, int32_t nt) {
    if (B <= 0 || A <= 0 || G <= 0 || nt < 0) return 0;
    return yolo_statistics_ws_bytes(B, A, G, nt);
}
int b200det_yolo_statistics_level(const float* head, int32_t B, int32_t A, int32_t C, int32_t G, const float* scaled_anchors,
                                  float stride, const float* target, int32_t nt, float ignore_thres, void* ws,
                                  size_t ws_bytes, float* out_rows, float* metrics, void* st) {
    B2_CHECK_ARG(B > 0 && A > 0 && C > 0 && G > 0 && nt >= 0, "bad sizes");
    B2_CHECK_ARG(head && scaled_anchors && ws && out_rows && metrics && (nt == 0 || target), "null argument");
    B2_CHECK_LIMIT((long long)B * A <= 65535, "B*A %lld > 65535", (long long)B * A);
    if (ws_bytes < yolo_statistics_ws_bytes(B, A, G, nt)) {
        set_error("workspace too small");
        return B200DET_EWORKSPACE;
    }
    return yolo_statistics_launch(head, B, A, C, G, scaled_anchors, stride, target, nt, ignore_thres, ws, out_rows, metrics,
                                  (cudaStream_t)st);
}

size_t b200det_ap_per_class_workspace_bytes(int32_t n) {
    if (n < 0 || n > (1 << 30)) return 0;
    return ap_per_class_ws_bytes(n);
}
int b200det_ap_per_class(const float* tp, const float* conf, const float* pred_cls, int32_t n, const int32_t* classes,
                         const int32_t* n_gt, int32_t num_classes, void* ws, size_t ws_bytes, double* p, double* r, double* ap,
                         double* f1, void* st) {
    B2_CHECK_ARG(n >= 0 && num_classes >= 0, "bad sizes");
    B2_CHECK_LIMIT(n <= (1 << 30), "num_pred %d > 2^30", n);
    B2_CHECK_LIMIT(num_classes <= 65535, "num_classes %d > 65535", num_classes);
    B2_CHECK_ARG(ws && (n == 0 || (tp && conf && pred_cls)) && (num_classes == 0 || (classes && n_gt && p && r && ap && f1)),
                 "null argument");
    if (ws_bytes < ap_per_class_ws_bytes(n)) {
        set_error("workspace too small");
        return B200DET_EWORKSPACE;
    }
    return ap_per_class_launch(tp, conf, pred_cls, n, classes, n_gt, num_classes, ws, p, r, ap, f1, (cudaStream_t)st);
}

}  // extern "C"
