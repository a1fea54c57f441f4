/*
 * b200det — C ABI of the B200-native detection post-processing / target-assignment library.
 *
 * This is the drop-in boundary for the hot path of Leyan529/ObjectDetectionPL.  The reference has no
 * FFI layer (it is pure Python); its seams are Python functions monkey-patched onto the Lightning
 * modules (model/YOLOV5.py:134-150) and module globals of LightningFunc/losses.py:5-6.  The Python
 * package `objectdetectionpl_b200` re-creates those functions with the reference signatures and calls
 * ONLY the entry points below (ctypes).  Each entry point names the reference
 * code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller owns all memory (inputs, outputs, workspace); the library never allocates, frees or
 *     keeps pointers.  Workspace sizes come from the *_workspace_bytes() queries;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it, no
 *     device-wide synchronisation happens inside the library;
 *   - return value: 0 on success, a negative B200DET_E* code or a positive cudaError_t otherwise;
 *     b200det_last_error() returns a thread-local message;
 *   - fp32 everywhere; indices int32 on the ABI (the Python layer widens to int64 where the reference
 *     returns int64);
 *   - there is no CPU fallback.
 */
#ifndef B200DET_H_
#define B200DET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DET_VERSION 100

#define B200DET_OK 0
#define B200DET_EINVAL (-1)      /* bad argument (shape, range, alignment, null pointer)   */
#define B200DET_ELIMIT (-2)      /* size beyond a compiled limit (see the limits below)     */
#define B200DET_EWORKSPACE (-3)  /* workspace too small                                     */

/* compiled limits */
#define B200DET_MAX_LEVELS 8     /* detection levels per call                               */
#define B200DET_MAX_ANCHORS 16   /* anchors per level                                       */
#define B200DET_MAX_CLASSES 4095 /* class id is packed in 12 bits of the sort payload       */
#define B200DET_MAX_CANDIDATES (1 << 20) /* candidates per image (20-bit slot)              */
#define B200DET_TILE 512         /* candidate tile: per-image regions are padded to this    */

/* decode modes of the YOLO head kernel */
#define B200DET_DECODE_NONE 0      /* values used as-is: what every reference NMS does (model/YOLOV3.py:289-305) */
#define B200DET_DECODE_YOLO_EXP 1  /* D1: sigmoid xy + grid, exp wh * anchor, * stride (accuracy.py:412-435,461) */
#define B200DET_DECODE_YOLOV5 2    /* D2: (2s-0.5+grid)*stride, (2s)^2*anchor (utils/YoloV5Utils.py:244-248)    */
#define B200DET_DECODE_YOLOV4_NORM 3 /* D3: ((s*sxy - 0.5(sxy-1)) + grid)/G, exp*anchor/G -> normalised corners
                                        x1 = bx - bw/2, x2 = x1 + bw (utils/YoloV4Utils.py:84-159); anchors in grid units */

/* memory layout of the head levels */
#define B200DET_LAYOUT_PLANAR 0         /* [B, A, 5+C, G, G]: what every reference NMS reads (model/YOLOV3.py:294-300)  */
#define B200DET_LAYOUT_CHANNELS_LAST 1  /* [B, A, G, G, 5+C]: what YOLOv5's Yolo_Layers really writes (model/YOLOV5.py:96);
                                           same candidate order, no permute copy                    */

int b200det_version(void);
const char* b200det_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * N1 — YOLOv2..v5 test-time post-processing: planar head -> filter -> score sort -> class-aware
 * merge-NMS.  Replaces `non_max_suppression(self, predictions, conf_thres, nms_thres)` of
 * model/YOLOV5.py:157-218, YOLOV3.py:273-335, YOLOV4.py:221-283, YOLOV2.py:159-222 together with the
 * helpers it calls (xywh2xyxy accuracy.py:289, bbox_iou accuracy.py:39).
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200det_yolo_desc {
    int32_t batch;                                   /* B                                              */
    int32_t num_anchors;                             /* A (3; YOLOv2: 5)                               */
    int32_t num_classes;                             /* C; every level holds A*(5+C) planes            */
    int32_t num_levels;                              /* <= B200DET_MAX_LEVELS                          */
    const float* head[B200DET_MAX_LEVELS];           /* level storage read as planar [B,A,5+C,G,G]     */
    int32_t grid[B200DET_MAX_LEVELS];                /* G per level, in the caller's list order        */
    int32_t decode_mode;                             /* B200DET_DECODE_*                               */
    float stride[B200DET_MAX_LEVELS];                /* decode modes only                              */
    float anchors[B200DET_MAX_LEVELS][B200DET_MAX_ANCHORS][2]; /* decode modes only: D1 scaled anchors
                                                        (grid units), D2 pixel anchors                 */
    float conf_thres;                                /* keep rows with conf >= conf_thres              */
    float nms_thres;                                 /* suppress when IoU_+1 > nms_thres               */
    int32_t layout;                                  /* B200DET_LAYOUT_* (0 = planar)                  */
    float scale_x_y;                                 /* B200DET_DECODE_YOLOV4_NORM only (YoloV4Utils.py:84); 0 = 1.0 */
} b200det_yolo_desc;

/* candidates per image N = sum_l A*G_l^2.  Every level starts on a tile boundary, so the per-image region holds
 * n_pad = sum_l roundup(A*G_l^2, TILE) slots (>= roundup(N, TILE): YOLOv3-416 has N = 10647, n_pad = 11264).
 * out_rows / out_index of the calls below MUST be sized from the n_pad this query returns. */
int b200det_yolo_num_candidates(const b200det_yolo_desc* d, int32_t* n, int32_t* n_pad);
size_t b200det_yolo_workspace_bytes(const b200det_yolo_desc* d);

/*
 * Full pipeline (decode+filter, per-image sort, merge-NMS, ordered emit), enqueued on `stream`.
 *   out_rows  [B, n_pad, 7] fp32: x1,y1,x2,y2 (cluster-merged), obj conf, class conf, class id;
 *                                 image b holds out_count[b] rows in descending score order
 *   out_index [B, n_pad] int32 (may be NULL): original candidate index of every kept row
 *   out_count [B] int32
 */
int b200det_yolo_nms(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes,
                     float* out_rows, int32_t* out_index, int32_t* out_count, void* stream);

/* The same pipeline with a PACKED result: the rows of all images back to back, image b at rows
 * [out_offsets[b], out_offsets[b+1]) of out_rows [>= sum of counts, 7] (capacity B*n_pad always suffices) and likewise
 * out_index (may be NULL); out_offsets [B+1] int32 (device).  One tiny extra launch (the prefix over the images); the
 * host splits the result with ONE slicing call and copies only the kept rows when it needs them on the host.
 * The counts are final once the NMS stage is done, one stage before the rows are: `counts_early` (may be NULL) receives
 * count [B] | offsets [B+1] as 2B+1 int32 at that point — it may be MAPPED PINNED HOST memory, the kernel writes it directly —
 * and `counts_ready_event` (a cudaEvent_t, may be NULL) is recorded right after, before the emit stage is enqueued.  A host
 * that waits on that event can size and slice the result while the emit kernel is still writing the rows. */
int b200det_yolo_nms_packed(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, float* out_rows,
                            int32_t* out_index, int32_t* out_count, int32_t* out_offsets, int32_t* counts_early,
                            void* counts_ready_event, void* stream);

/* The PADDED pipeline (b200det_yolo_nms) with the early counts of the packed form: `counts_early` (may be NULL; device or
 * mapped pinned host memory) receives count [B] | exclusive offsets [B+1] as soon as the NMS stage is done, and
 * `counts_ready_event` (cudaEvent_t, may be NULL) is recorded right after, before the emit stage is enqueued. */
int b200det_yolo_nms_early(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, float* out_rows,
                           int32_t* out_index, int32_t* out_count, int32_t* counts_early, void* counts_ready_event,
                           void* stream);

/* Stage entry points (the pipeline above is exactly these five calls in order); used by the tests
 * and by profiling.  All operate on the workspace laid out by b200det_yolo_workspace_bytes().
 * reset = one cudaMemsetAsync of the counter header; decode = the fused decode+filter kernel alone. */
int b200det_yolo_stage_reset(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, void* stream);
int b200det_yolo_stage_decode(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, void* stream);
int b200det_yolo_stage_sort(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, void* stream);
int b200det_yolo_stage_nms(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, void* stream);
int b200det_yolo_stage_emit(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes,
                            float* out_rows, int32_t* out_index, int32_t* out_count, void* stream);
int b200det_yolo_stage_emit_packed(const b200det_yolo_desc* d, void* workspace, size_t workspace_bytes, float* out_rows,
                                   int32_t* out_index, int32_t* out_count, int32_t* out_offsets, void* stream);

/* Workspace introspection for tests: byte offset and element count of a named internal array
 * ("box4","cc2","orig","key","pay","rank","count","tile_count","cls_hist","seg_off","sorted_pay",
 *  "sorted_rank","kpay","mbox").  Returns B200DET_EINVAL for an unknown name. */
int b200det_yolo_workspace_field(const b200det_yolo_desc* d, const char* name, size_t* offset, size_t* bytes);

/* ------------------------------------------------------------------------------------------------
 * decode_box — full decoded map [B, A*G*G, 5+C] of one level (north-star API; restates D1/D2).
 * Replaces the inline decode of accuracy.py:402-435,459-466 / losses.py:679-703 (mode YOLO_EXP) and
 * utils/YoloV5Utils.py:241-248 (mode YOLOV5).  mode NONE is the planar->rows permute of
 * model/YOLOV3.py:294-300 with xywh left untouched.
 * ---------------------------------------------------------------------------------------------- */
int b200det_decode_box(const float* head, int32_t batch, int32_t num_anchors, int32_t num_classes,
                       int32_t grid, int32_t decode_mode, const float* anchors /*[A,2] device; NULL for DECODE_NONE*/, float stride,
                       float* out /*[B, A*G*G, 5+C]*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * D3 — `yolo_forward_dynamic(output, conf_thresh, num_classes, anchors, num_anchors, scale_x_y, ...)` of
 * LightningFunc/utils/YoloV4Utils.py:36-176 for one head level (planar [B, A*(5+C), H, W]):
 *   boxes  normalised corner boxes, row r = (b, a*H*W + cell) at boxes[r*ld_boxes + 0..3]   (reference: [B,N,1,4], ld 4)
 *   confs  sigmoid(cls) * sigmoid(obj) at confs[r*ld_confs + 0..C-1]                         (reference: [B,N,C],   ld C)
 *   det    optional (may be NULL): sigmoid(obj) at det[r*ld_det]
 * anchors [A,2] on the device, in grid units (what the reference's callers pass, YoloV4Utils.py:99-102).
 * ---------------------------------------------------------------------------------------------- */
int b200det_yolo_forward_dynamic(const float* head, int32_t batch, int32_t num_anchors, int32_t num_classes, int32_t height,
                                 int32_t width, const float* anchors, float scale_x_y, float* boxes, int64_t ld_boxes,
                                 float* confs, int64_t ld_confs, float* det, int64_t ld_det, void* stream);

/* ------------------------------------------------------------------------------------------------
 * D4 + N2 — SSD / RetinaNet prior decode, sigmoid-argmax, score filter, top-k, class-agnostic greedy
 * NMS.  Replaces `non_max_suppression(self, predictions, topk, nms_thresh, class_thresh, mode)` of
 * model/SSD.py:249-310 == model/RetinaNet.py:117-178.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200det_prior_desc {
    int32_t batch;            /* B */
    int32_t num_priors;       /* P */
    int32_t num_classes;      /* C */
    const float* loc;         /* [B,P,4] */
    const float* cls;         /* [B,P,C] logits */
    const float* priors;      /* [P,4] cx,cy,w,h */
    int32_t topk;             /* 100 */
    float nms_thresh;         /* 0.5: a box survives a keeper when ovr <= nms_thresh */
    float class_thresh;       /* 0.45: candidate when sigmoid(max logit) > class_thresh */
    int32_t mode_min;         /* 0: 'union', 1: 'min' (SSD.py:291-296) */
    int32_t compat;           /* 1: reproduce quirks (i) drop-last and (ii) filtered-index gather */
} b200det_prior_desc;

size_t b200det_prior_workspace_bytes(const b200det_prior_desc* d);
/*   out_rows  [B, topk, 7]: x1,y1,x2,y2, 0, score, label ; out_index [B, topk] (may be NULL): index
 *   of the prior whose box/label the row carries ; out_count [B] ; cand_count [B] (may be NULL):
 *   number of candidates that passed class_thresh (the Python layer raises IndexError on 1, SSD.py:266) */
int b200det_prior_nms(const b200det_prior_desc* d, void* workspace, size_t workspace_bytes,
                      float* out_rows, int32_t* out_index, int32_t* out_count, int32_t* cand_count, void* stream);
/* Stage entry points (b200det_prior_nms is exactly these two calls in order); used by profiling and by bench.py to put
 * CUDA events around the streaming kernel alone.  decode = prior decode + sigmoid-argmax + score filter (reads loc/cls
 * once: the HBM-bound part); select_nms = top-k selection + greedy NMS + output. */
int b200det_prior_stage_decode(const b200det_prior_desc* d, void* workspace, size_t workspace_bytes, void* stream);
int b200det_prior_stage_select_nms(const b200det_prior_desc* d, void* workspace, size_t workspace_bytes, float* out_rows,
                                   int32_t* out_index, int32_t* out_count, int32_t* cand_count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N3 / N4 / T3 / T5 — elementwise box maths.
 * ---------------------------------------------------------------------------------------------- */
/* xywh2xyxy, accuracy.py:289-295.  x,y: [n,4] */
int b200det_xywh2xyxy(const float* x, float* y, int64_t n, void* stream);
/* bbox_iou (IoU_+1, +1e-16), accuracy.py:39-69.  box1 [n1,4] with n1 in {1,n}, box2 [n,4] -> out[n] */
int b200det_bbox_iou_plus1(const float* box1, int64_t n1, const float* box2, int64_t n, int32_t x1y1x2y2,
                           float* out, void* stream);
/* iou, accuracy.py:6-37 (corner boxes, no +1, no eps).  a,b: [n,4] -> out[n] */
int b200det_pair_iou(const float* a, const float* b, int64_t n, float* out, void* stream);

#define B200DET_IOU 0
#define B200DET_GIOU 1
#define B200DET_DIOU 2
#define B200DET_CIOU 3
/* bbox_iou_v5, accuracy.py:71-114.  box1, box2 are the reference's TRANSPOSED [4,n] tensors
 * (row stride `ld1` / `ld2` elements, so .t() views need no copy: element (k,i) at p[k*ld + i*inc]).
 * backward: grad wrt box1 only (box2 is the target; CIoU's alpha is a constant, accuracy.py:110). */
int b200det_bbox_iou_v5_fwd(const float* box1, int64_t ld1, int64_t inc1, const float* box2, int64_t ld2,
                            int64_t inc2, int64_t n, int32_t x1y1x2y2, int32_t kind, float* out, void* stream);
int b200det_bbox_iou_v5_bwd(const float* box1, int64_t ld1, int64_t inc1, const float* box2, int64_t ld2,
                            int64_t inc2, int64_t n, int32_t x1y1x2y2, int32_t kind, const float* grad_out,
                            float* grad_box1 /*[4,n] contiguous*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * T2 / T4 — YOLOv5 target assignment.  Replaces build_targets_v5 (accuracy.py:472-521) and the
 * matched-row gather + decode + GIoU + objectness scatter of losses.py:105-123.
 * ---------------------------------------------------------------------------------------------- */
/* One level.  targets [nt,6] (img,cls,cx,cy,w,h normalised); anchors_host [na,2] (grid units).
 * Outputs have capacity 5*na*nt rows, written in the reference's row order:
 *   out_b,out_a,out_gj,out_gi,out_cls int32 [cap]; out_tbox [cap,4]; out_anch [cap,2]; out_count[1] */
int b200det_build_targets_v5_level(const float* targets, int32_t num_targets, const float* anchors_host,
                                   int32_t num_anchors, int32_t nx, int32_t ny, int32_t* out_b, int32_t* out_a,
                                   int32_t* out_gj, int32_t* out_gi, int32_t* out_cls, float* out_tbox,
                                   float* out_anch, int32_t* out_count, void* stream);
/* Matched-row forward: pi [B,na,ny,nx,5+C] channels-last (losses.py:112), rows from the call above.
 *   giou[m]; tobj[B,na,ny,nx] must be zero-filled by the caller, receives clamp(giou,0) with
 *   "last row wins" on duplicate cells (losses.py:123). */
/* all levels in one launch (one CTA per level): per-level host arrays nx/ny [nl], anchors [nl][na][2], and host
 * arrays of the nl device output pointers; count_out [nl] (device).  Same rows as nl calls of the _level entry. */
int b200det_build_targets_v5(const float* targets, int32_t num_targets, int32_t num_levels, const float* anchors_host,
                             int32_t num_anchors, const int32_t* nx_host, const int32_t* ny_host, int32_t* const* out_b,
                             int32_t* const* out_a, int32_t* const* out_gj, int32_t* const* out_gi, int32_t* const* out_cls,
                             float* const* out_tbox, float* const* out_anch, int32_t* count_out, void* stream);
int b200det_v5_match_fwd(const float* pi, int32_t batch, int32_t num_anchors, int32_t ny, int32_t nx,
                         int32_t fields, const int32_t* b, const int32_t* a, const int32_t* gj,
                         const int32_t* gi, const float* tbox, const float* anch, int32_t m, float* giou,
                         float* tobj, void* stream);
/* Backward of giou wrt pi (only the 4 box fields of the matched rows receive gradient; accumulated
 * with atomics into grad_pi, which the caller zero-fills). */
int b200det_v5_match_bwd(const float* pi, int32_t batch, int32_t num_anchors, int32_t ny, int32_t nx,
                         int32_t fields, const int32_t* b, const int32_t* a, const int32_t* gj,
                         const int32_t* gi, const float* tbox, const float* anch, int32_t m,
                         const float* grad_giou, float* grad_pi, void* stream);

/* ------------------------------------------------------------------------------------------------
 * T1 — build_targets (YOLOv2..v4 grid scatter builder), accuracy.py:305-380.
 * pred_boxes [B,A,G,G,4] (grid units), pred_cls [B,A,G,G,C], target [nt,6], anchors [A,2] device.
 * Outputs (caller-allocated, any content): iou_scores, class_mask, tx,ty,tw,th fp32 [B,A,G,G];
 * obj_mask, noobj_mask uint8 [B,A,G,G]; tcls fp32 [B,A,G,G,C].  Duplicate cells: highest target row
 * wins (the reference's CPU index_put_ order); tcls is multi-hot.  status[0] (int32, device) is set
 * to a bit mask: bit0 = index guard of accuracy.py:340-344 tripped, bit1 = label guard of :361-367,
 * bit2 = a negative image / cell / label index below -size (the reference's indexing raises IndexError
 * there; here every scatter is skipped and the Python layer raises from the bit). */
size_t b200det_build_targets_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t grid, int32_t num_targets);
int b200det_build_targets(const float* pred_boxes, const float* pred_cls, const float* target,
                          const float* anchors, int32_t batch, int32_t num_anchors, int32_t grid,
                          int32_t num_classes, int32_t num_targets, float ignore_thres, void* workspace,
                          size_t workspace_bytes, float* iou_scores, float* class_mask, uint8_t* obj_mask,
                          uint8_t* noobj_mask, float* tx, float* ty, float* tw, float* th, float* tcls,
                          int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * T5 — SSD prior matching (SSDLoss.match, losses.py:199-218) and
 * T6 — RetinaNet anchor assignment + target encoding (losses.py:375-403, 423-443).
 * ---------------------------------------------------------------------------------------------- */
size_t b200det_ssd_match_workspace_bytes(int32_t num_priors, int32_t num_gt);
/* priors [P,4], gt [M,4] (cx,cy,w,h in [0,1]) -> box_with_annotation int32 [P], matched uint8 [P] */
int b200det_ssd_match(const float* priors, int32_t num_priors, const float* gt, int32_t num_gt,
                      float match_thresh, void* workspace, size_t workspace_bytes,
                      int32_t* box_with_annotation, uint8_t* matched, void* stream);
/* anchors [A,4] pixel cxcywh; targets [nt,6] sorted or not; per image b the rows with targets[:,0]==b
 * in their original order.  loc_targets [B,A,4], cls_targets int32 [B,A] (1+label, 0 background,
 * -1 ignore).  An image without targets gets all-zero rows (the reference raises there). */
size_t b200det_retina_assign_workspace_bytes(int32_t batch, int32_t num_targets);
int b200det_retina_assign(const float* anchors, int32_t num_anchors, const float* targets,
                          int32_t num_targets, int32_t batch, float img_size, void* workspace,
                          size_t workspace_bytes, float* loc_targets, int32_t* cls_targets, void* stream);

/* ------------------------------------------------------------------------------------------------
 * T4' — the per-level loss terms of `MultiScaleRegionLoss_v5.forward` (LightningFunc/losses.py:98-152) fused
 * (SURVEY §8f row 2): matched-row gather + D2 decode + GIoU (:105-119), objectness target scatter (:123), class
 * focal-BCE over the matched rows (:127-133) and objectness focal-BCE over every cell (:137; FocalLoss :37-64 around
 * BCEWithLogitsLoss with pos_weight 1).
 *   fwd: giou[m], tobj[B,na,ny,nx] (zero-filled here), sums[3] fp64 (device) = the three MEANS on return:
 *        sum(1-giou) / max(m,1), sum FL_obj / cells, sum FL_cls / max(m*C,1) (accumulated in place, divided by the last
 *        kernel).  with_cls = 0 skips the class term (nc == 1, :127).
 *   bwd: gpi (zero-filled by the caller) += d/dpi of  g3[0]*inv_nbox*sum(1-giou) + g3[1]*inv_cells*sum FL_obj +
 *        g3[2]*inv_ncls*sum FL_cls, g3 = the three upstream gradients on the DEVICE (no host sync in backward);
 *        tobj is treated as a constant (`giou.detach()`, :123).
 * ---------------------------------------------------------------------------------------------- */
int b200det_v5_loss_fwd(const float* pi, int32_t batch, int32_t na, int32_t ny, int32_t nx, int32_t fields,
                        const int32_t* b, const int32_t* a, const int32_t* gj, const int32_t* gi, const int32_t* tcls,
                        const float* tbox, const float* anch, int32_t m, float cp, float cn, float gamma, float alpha,
                        int32_t with_cls, float* giou, float* tobj, double* sums, void* stream);
int b200det_v5_loss_bwd(const float* pi, int32_t batch, int32_t na, int32_t ny, int32_t nx, int32_t fields,
                        const int32_t* b, const int32_t* a, const int32_t* gj, const int32_t* gi, const int32_t* tcls,
                        const float* tbox, const float* anch, int32_t m, float cp, float cn, float gamma, float alpha,
                        int32_t with_cls, const float* tobj, const float* g3, float inv_nbox, float inv_cells,
                        float inv_ncls, float* gpi, void* stream);
/* Same as b200det_v5_loss_bwd, but gpi need NOT be initialised: the call defines every element (zeros outside the objectness
 * column and the matched rows), written as one linear stream instead of a zero-fill pass followed by strided updates. */
int b200det_v5_loss_bwd_full(const float* pi, int32_t batch, int32_t na, int32_t ny, int32_t nx, int32_t fields,
                             const int32_t* b, const int32_t* a, const int32_t* gj, const int32_t* gi, const int32_t* tcls,
                             const float* tbox, const float* anch, int32_t m, float cp, float cn, float gamma, float alpha,
                             int32_t with_cls, const float* tobj, const float* g3, float inv_nbox, float inv_cells,
                             float inv_ncls, float* gpi, void* stream);
/* All levels of the criterion per call, sync-free (what `v5_loss` of the Python layer uses): one launch per STAGE for all levels
 * (the per-level kernels are 5-45 us each; ~30 launches per step level by level, 11 this way), the matched-row count of every
 * level stays on the DEVICE (m_dev = the count_out word b200det_build_targets_v5 wrote for it; the row arrays are passed at their
 * capacity `cap` = 5 * na * nt and every kernel bounds itself by min(cap, *m_dev), the means' divisors max(m, 1), max(m * C, 1)
 * are formed on the device from the same word), and the gain-weighted combination (losses.py:139-152) is part of the call:
 * out4 = (loss, Localization, Classification, Conf_obj).  Results are identical to the per-level host-count forms + combine.
 *   giou [nl, cap], tobj [sum of cells] (levels back to back; scratch), sums [nl, 3] fp64 (scratch; holds the three means of
 *   every level afterwards), obj_grad [sum of cells]: d FL / d logit of every cell's objectness term, which the backward reads
 *   (4 contiguous bytes per cell) instead of re-reading column 4 of pi with a (5+C)-float stride next to its write stream.
 * _bwd_all: every level's gpi (same shape as pi, 16-byte aligned) is DEFINED by the call (zeros outside column 4 and the
 *   matched rows; no zero-fill needed); g_loss / g_box / g_cls / g_obj are the upstream gradients of out4 (device scalars, NULL =
 *   none), g3 [3] is scratch. */
typedef struct {
    const float* pi;                                       /* [batch, na, ny, nx, fields] fp32                         */
    int32_t batch, na, ny, nx, fields;
    const int32_t *b, *a, *gj, *gi, *tcls;                 /* [cap] rows of b200det_build_targets_v5 for this level     */
    const float* tbox;                                     /* [cap, 4], 16-byte aligned                                */
    const float* anch;                                     /* [cap, 2]                                                 */
    const int32_t* m_dev;                                  /* device word: rows in use                                  */
    float* gpi;                                            /* backward only                                            */
} b200det_v5_level;
int b200det_v5_loss_fwd_all(const b200det_v5_level* levels, int32_t nl, int32_t cap, float cp, float cn, float gamma, float alpha,
                            int32_t with_cls, float wbox, float wobj, float wcls, float* giou, float* tobj, float* obj_grad,
                            double* sums, float* out4, void* stream);
int b200det_v5_loss_bwd_all(const b200det_v5_level* levels, int32_t nl, int32_t cap, float cp, float cn, float gamma, float alpha,
                            int32_t with_cls, float wbox, float wobj, float wcls, const float* obj_grad, const float* g_loss,
                            const float* g_box, const float* g_cls, const float* g_obj, float* g3, void* stream);
/* The tail of the same forward (losses.py:139-152): means[nl][3] (fp64, what b200det_v5_loss_fwd left per level) are added
 * in level order in fp32, scaled by the three gains and summed: out4 = (loss, Localization, Classification, Conf_obj).
 * _bwd: g3[3] for b200det_v5_loss_bwd (the same for every level) from the upstream gradients of those four outputs
 * (device scalars, NULL = no gradient). */
int b200det_v5_loss_combine(const double* means, int32_t nl, float wbox, float wobj, float wcls, float* out4, void* stream);
int b200det_v5_loss_combine_bwd(const float* g_loss, const float* g_box, const float* g_cls, const float* g_obj, float wbox,
                                float wobj, float wcls, float* g3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * M1 — true-positive matching of detections against labels.  Replaces `get_batch_statistics(outputs,
 * targets, iou_threshold)` (LightningFunc/accuracy.py:116-154, called from LightningFunc/step.py:95).
 *   rows       detection rows of 7 floats (x1,y1,x2,y2,conf,cls_conf,label); image b owns rows
 *              [row_start[b], row_start[b] + count[b])  (row units; row_start int64 on the device) — covers both the
 *              padded [B,n_pad,7] output of b200det_yolo_nms (row_start[b] = b*n_pad) and a packed concatenation;
 *   targets    [num_targets,6] = (image, label, x1, y1, x2, y2) exactly as the reference passes them;
 *   tp         out, indexed like rows: 1.0 for a true positive, else 0.0;  max_count >= max(count) sizes the grid.
 * ---------------------------------------------------------------------------------------------- */
size_t b200det_batch_statistics_workspace_bytes(int32_t batch, int32_t num_targets);
int b200det_batch_statistics(const float* rows, const int64_t* row_start, const int32_t* count, int32_t batch,
                             int32_t max_count, const float* targets, int32_t num_targets, float iou_threshold,
                             void* ws, size_t ws_bytes, float* tp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Detection packing for the multi-GPU exchange step (the detections of all images are consumed together by the test
 * epoch, LightningFunc/step.py:95,102-130): rows [B, row_pitch, 7] + count [B] (the padded output of b200det_yolo_nms,
 * row_pitch = n_pad; or b200det_prior_nms, row_pitch = topk) -> out [cap, 8] dense rows = 7 detection columns + the
 * GLOBAL image id (image_offset + b), image b at [offsets[b], offsets[b+1]) in its score order.  offsets [B+1] int32
 * (device, may be NULL).  max_rows >= max(count) sizes the grid (row_pitch is always enough); rows beyond `cap` are
 * dropped (offsets[B] still reports the full total, so the caller can detect it).
 * ---------------------------------------------------------------------------------------------- */
int b200det_pack_detections(const float* rows, const int32_t* count, int32_t batch, int64_t row_pitch, int32_t max_rows,
                            int32_t image_offset, float* out, int64_t cap, int32_t* offsets, void* stream);

/* ------------------------------------------------------------------------------------------------
 * M2 — per-class precision / recall / AP / F1.  Replaces `ap_per_class(tp, conf, pred_cls, target_cls)`
 * with `compute_ap` (LightningFunc/accuracy.py:207-287, called from LightningFunc/step.py:115).
 *   tp, conf, pred_cls   [num_pred] fp32 (tp != 0 counts as a true positive; labels as floats, like the rows);
 *   classes, n_gt        [num_classes] int32: np.unique(target_cls) and the label count of each (device);
 *   p, r, ap, f1         [num_classes] fp64 out.  Ties in conf are ordered by position (numpy's argsort is unstable).
 * Limits: num_pred <= 2^30; class ids that are not integers in [0, 2^30) match no evaluated class.
 * ---------------------------------------------------------------------------------------------- */
size_t b200det_ap_per_class_workspace_bytes(int32_t num_pred);
int b200det_ap_per_class(const float* tp, const float* conf, const float* pred_cls, int32_t num_pred,
                         const int32_t* classes, const int32_t* n_gt, int32_t num_classes, void* ws, size_t ws_bytes,
                         double* p, double* r, double* ap, double* f1, void* stream);

/* ------------------------------------------------------------------------------------------------
 * M3 — test-time statistics of one YOLOv2..v4 level.  Replaces the body of `get_yolo_statistics(self, output,
 * target)` (LightningFunc/accuracy.py:382-470, called from LightningFunc/step.py:99) for one head tensor:
 * D1 decode (:412-435), build_targets on the decoded map (:437-443), the six metrics (:447-457) and the decoded map
 * `output` (:459-466).
 *   head            planar [B, A, 5+C, G, G] raw logits;  scaled_anchors [A,2] device = anchors / stride (:427);
 *   out_rows        [B, A*G*G, 5+C] out: (x, y, w, h) * stride, sigmoid(conf), sigmoid(cls)   (the reference's `output`)
 *   metrics         [6] fp32 out (device): cls_acc, recall50, recall75, precision, conf_obj, conf_noobj.
 * ---------------------------------------------------------------------------------------------- */
size_t b200det_yolo_statistics_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t grid, int32_t num_targets);
int b200det_yolo_statistics_level(const float* head, int32_t batch, int32_t num_anchors, int32_t num_classes, int32_t grid,
                                  const float* scaled_anchors, float stride, const float* target, int32_t num_targets,
                                  float ignore_thres, void* ws, size_t ws_bytes, float* out_rows, float* metrics,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DET_H_ */
