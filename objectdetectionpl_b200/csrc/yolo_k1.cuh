// Shared pieces of the two K1 (fused YOLO decode + filter) kernels: parameters, argmax step, per-candidate decode.
#pragma once
#include "yolo_ws.cuh"

namespace b200det {

struct K1Params {
    const float* head[B200DET_MAX_LEVELS];
    int G[B200DET_MAX_LEVELS];
    int GG[B200DET_MAX_LEVELS];
    int off[B200DET_MAX_LEVELS + 1];        // first candidate index of a level
    int tile_off[B200DET_MAX_LEVELS + 1];   // first tile of a level (levels start on tile boundaries)
    float stride[B200DET_MAX_LEVELS];
    float anc[B200DET_MAX_LEVELS][B200DET_MAX_ANCHORS][2];
    int nlevels, A, C, N, n_pad, n_tiles;
    int slow_first;                         // > 0: grid is (batch, tiles) and CTA row y takes tile (y + slow_first - 1) % n_tiles
    float conf_thres;
    float sxy, soff;                        // DECODE_YOLOV4_NORM: scale_x_y and 0.5 * (scale_x_y - 1)
    // outputs
    float4* box4;
    float2* cc2;
    uint32_t* orig;
    uint32_t* key;
    uint32_t* pay;
    uint32_t* tile_count;
    uint32_t* count;
    uint32_t* cls_hist;
};

// torch.max(dim) semantics (aten TensorCompareKernel): update when !(v <= best); stop at the first NaN.
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& besti) {
    if (!(v <= best) && (best == best)) { best = v; besti = c; }
}

// Decode one candidate (fields t[0..4] = raw x, y, w, h, obj; `best` = max class value) into corner box, obj conf and
// class conf.  MODE NONE: values used as-is (what every reference NMS does, model/YOLOV3.py:289-305).
template <int MODE>
__device__ __forceinline__ void k1_finish(const K1Params& p, int lvl, int a, int cell, const float (&t)[5], float best,
                                          float (&box)[4], float& conf, float& ccf) {
    float cx, cy, w, h, cf, cc;
    if (MODE == B200DET_DECODE_NONE) {
        cx = t[0]; cy = t[1]; w = t[2]; h = t[3]; cf = t[4]; cc = best;
    } else {
        const int G = p.G[lvl];
        const int gy = cell / G;
        const float gx = (float)(cell - gy * G);
        const float st = p.stride[lvl];
        const float aw = p.anc[lvl][a][0], ah = p.anc[lvl][a][1];
        if (MODE == B200DET_DECODE_YOLOV4_NORM) {
            // utils/YoloV4Utils.py:84-85,117-121,147-148,156-159: normalised corners, x2 = x1 + bw
            const float Gf = (float)G;
            const float bx = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(sigmoidf_acc(t[0]), p.sxy), p.soff), gx), Gf);
            const float by = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(sigmoidf_acc(t[1]), p.sxy), p.soff), (float)gy), Gf);
            const float bw = __fdiv_rn(__fmul_rn(expf(t[2]), aw), Gf);
            const float bh = __fdiv_rn(__fmul_rn(expf(t[3]), ah), Gf);
            box[0] = __fsub_rn(bx, __fmul_rn(bw, 0.5f));
            box[1] = __fsub_rn(by, __fmul_rn(bh, 0.5f));
            box[2] = __fadd_rn(box[0], bw);
            box[3] = __fadd_rn(box[1], bh);
            conf = sigmoidf_acc(t[4]);
            ccf = sigmoidf_acc(best);
            return;
        }
        if (MODE == B200DET_DECODE_YOLO_EXP) {
            // accuracy.py:432-435 then *stride (:461)
            cx = __fmul_rn(__fadd_rn(sigmoidf_acc(t[0]), gx), st);
            cy = __fmul_rn(__fadd_rn(sigmoidf_acc(t[1]), (float)gy), st);
            w = __fmul_rn(__fmul_rn(expf(t[2]), aw), st);
            h = __fmul_rn(__fmul_rn(expf(t[3]), ah), st);
        } else {
            // utils/YoloV5Utils.py:246-247
            cx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sigmoidf_acc(t[0]), 2.0f), 0.5f), gx), st);
            cy = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sigmoidf_acc(t[1]), 2.0f), 0.5f), (float)gy), st);
            const float sw = __fmul_rn(sigmoidf_acc(t[2]), 2.0f);
            const float sh = __fmul_rn(sigmoidf_acc(t[3]), 2.0f);
            w = __fmul_rn(__fmul_rn(sw, sw), aw);
            h = __fmul_rn(__fmul_rn(sh, sh), ah);
        }
        cf = sigmoidf_acc(t[4]);
        cc = sigmoidf_acc(best);   // sigmoid is monotone: argmax taken on the logits
    }
    // xywh2xyxy, accuracy.py:289-295 (x/2 == x*0.5 exactly)
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    box[0] = __fsub_rn(cx, hw);
    box[1] = __fsub_rn(cy, hh);
    box[2] = __fadd_rn(cx, hw);
    box[3] = __fadd_rn(cy, hh);
    conf = cf; ccf = cc;
}


// Shared staging of one tile's surviving candidates.  The survivors are written to shared memory at their compacted slot
// and then copied out with stores whose lanes cover consecutive addresses, so every global store fills whole 32-byte
// sectors.  (Measured on B200: writing the five candidate arrays straight from registers — four 16/8/4-byte pieces per
// thread at a 64/32/16-byte lane stride — cost 20 us per launch in partial-sector L2 traffic; staged it costs ~7 us.)
struct K1Stage {
    float4 box[kTile];
    float2 cc[kTile];
    uint32_t orig[kTile];
    uint32_t key[kTile];
    uint32_t cls[kTile];
};

__device__ __forceinline__ void k1_stage_put(K1Stage& st, int ofs, const float (&box)[4], float conf, float ccf, uint32_t orig,
                                             int cls) {
    st.box[ofs] = make_float4(box[0], box[1], box[2], box[3]);
    st.cc[ofs] = make_float2(conf, ccf);
    st.orig[ofs] = orig;
    st.key[ofs] = score_sort_key(__fmul_rn(conf, ccf));     // model/YOLOV3.py:315
    st.cls[ofs] = (uint32_t)cls;
}

// copy-out by NT threads (call after a barrier that orders the k1_stage_put calls)
template <int NT>
__device__ __forceinline__ void k1_stage_flush(const K1Stage& st, const K1Params& p, size_t img, int tile, int total, int tid) {
    const size_t base = img + (size_t)tile * kTile;
    for (int i = tid; i < total; i += NT) {
        p.box4[base + i] = st.box[i];
        p.cc2[base + i] = st.cc[i];
        p.orig[base + i] = st.orig[i];
        p.key[base + i] = st.key[i];
        p.pay[base + i] = (st.cls[i] << kSlotBits) | (uint32_t)(tile * kTile + i);
    }
}

}  // namespace b200det
