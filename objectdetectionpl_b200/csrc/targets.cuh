// Parameter blocks of the YOLOv5 loss kernels (targets.cu) that the C-ABI layer (api.cu) fills in.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200det {

struct MatchParams {
    const float* pi;
    int B, na, ny, nx, F;
    const int32_t *b, *a, *gj, *gi;
    const float* tbox;
    const float* anch;
    int m;                      // matched rows; with m_dev set: the CAPACITY of the row arrays (sizes the grids)
    const int32_t* m_dev;       // optional: the row count lives on the device (no host sync between build_targets_v5 and the loss)
};

constexpr int kV5MaxLevels = 5;
struct V5Level {
    MatchParams p;              // m = capacity of the row arrays, m_dev = the device-side row count
    const int32_t* tcls;
    float* giou;                // [cap]
    float* tobj;                // [cells]
    float* obj_grad;            // [cells]   forward: written, backward: read
    double* sums;               // [3]       sum(1 - giou), sum FL_obj, sum FL_cls -> the three means
    float* gpi;                 // backward: the level's gradient tensor (fully written)
    long long cells;
    float inv_cells;
    int obj_blocks;             // CTAs of the per-cell kernels
};
struct V5Multi {
    V5Level lv[kV5MaxLevels];
    int nl;
    float cp, cn, gamma, alpha;
    int with_cls;
    float wbox, wobj, wcls;
    const float* g3;
};

int v5_loss_fwd_all_launch(V5Multi& mp, int cap, float* giou, float* tobj, float* obj_grad, double* sums, float* out4,
                           cudaStream_t st);
int v5_loss_bwd_all_launch(V5Multi& mp, int cap, const float* obj_grad, const float* g_loss, const float* g_box,
                           const float* g_cls, const float* g_obj, float* g3, cudaStream_t st);

}  // namespace b200det
