// Micro-benchmark: what bounds the "whole gradient tensor" writer of the YOLOv5 objectness backward (targets.cu:
// v5_loss_obj_bwd_full_kernel)?  cells x F floats are written (zeros except column 4); variants isolate the write stream, the
// strided column-4 reads and the column-4 patch stores.      nvcc -O3 -arch=sm_100a -o gradfill gradfill.cu && ./gradfill
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) fill_only(float* __restrict__ gpi, int F, long long cells) {
    const int tid = threadIdx.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long c0 = (long long)blockIdx.x * 256; c0 < cells; c0 += (long long)gridDim.x * 256) {
        const int ncell = (int)min((long long)256, cells - c0);
        float4* o = reinterpret_cast<float4*>(gpi + c0 * F);
        const int n4 = (ncell * F) >> 2;
        for (int v = tid; v < n4; v += 256) o[v] = z;
    }
}
// + strided reads of column 4 and tobj, folded into ONE extra 4-byte store per CTA step (keeps the loads alive)
__global__ void __launch_bounds__(256) fill_read(const float* __restrict__ pi, const float* __restrict__ tobj,
                                                 float* __restrict__ gpi, float* __restrict__ sink, int F, long long cells) {
    const int tid = threadIdx.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc = 0.f;
    for (long long c0 = (long long)blockIdx.x * 256; c0 < cells; c0 += (long long)gridDim.x * 256) {
        const long long c = c0 + tid;
        if (c < cells) acc += pi[c * F + 4] + tobj[c];
        const int ncell = (int)min((long long)256, cells - c0);
        float4* o = reinterpret_cast<float4*>(gpi + c0 * F);
        const int n4 = (ncell * F) >> 2;
        for (int v = tid; v < n4; v += 256) o[v] = z;
    }
    if (acc == 12345.678f) sink[0] = acc;
}
// + the column-4 patch after a CTA barrier (the product kernel's structure, 64-cell sub-blocks)
__global__ void __launch_bounds__(256) fill_read_patch(const float* __restrict__ pi, const float* __restrict__ tobj,
                                                       float* __restrict__ gpi, int F, long long cells, int sub) {
    const int tid = threadIdx.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long c0 = (long long)blockIdx.x * 256; c0 < cells; c0 += (long long)gridDim.x * 256) {
        const long long c = c0 + tid;
        float g = 0.f;
        if (c < cells) g = pi[c * F + 4] + tobj[c];
        const int ncell = (int)min((long long)256, cells - c0);
        float* out = gpi + c0 * F;
        for (int cb = 0; cb < ncell; cb += sub) {
            const int nsub = min(sub, ncell - cb);
            float4* o = reinterpret_cast<float4*>(out + cb * F);
            const int n4 = (nsub * F) >> 2;
            for (int v = tid; v < n4; v += 256) o[v] = z;
            __syncthreads();
            if (tid >= cb && tid < cb + nsub) out[tid * F + 4] = g;
        }
    }
}
// single pass: every float4 is composed (zeros, or the cell's value in the lane of column 4) — no partial writes.
// The values come from shared memory (computed by the CTA first), the (cell, field) of a float4 from an incremental walk.
__global__ void __launch_bounds__(256) compose(const float* __restrict__ pi, const float* __restrict__ tobj,
                                               float* __restrict__ gpi, int F, long long cells) {
    __shared__ float s_g[256];
    const int tid = threadIdx.x;
    const int step_q = 1024 / F, step_r = 1024 - step_q * F;
    for (long long c0 = (long long)blockIdx.x * 256; c0 < cells; c0 += (long long)gridDim.x * 256) {
        const long long c = c0 + tid;
        __syncthreads();
        s_g[tid] = c < cells ? pi[c * F + 4] + tobj[c] : 0.f;
        __syncthreads();
        const int ncell = (int)min((long long)256, cells - c0);
        float4* o = reinterpret_cast<float4*>(gpi + c0 * F);
        const int n4 = (ncell * F) >> 2;
        int cell = (tid << 2) / F, f = (tid << 2) - cell * F;
        for (int v = tid; v < n4; v += 256) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            // column 4 of `cell` sits in this float4 iff 1 <= f <= 4 (lane 4 - f); of cell + 1 iff f + 3 >= F + 4, i.e. never for F > 7
            if (f >= 1 && f <= 4) {
                const float g = s_g[cell];
                if (f == 4) r.x = g; else if (f == 3) r.y = g; else if (f == 2) r.z = g; else r.w = g;
            } else if (f + 3 >= F + 4) {
                const float g = s_g[cell + 1];
                const int l = F + 4 - f;
                if (l == 0) r.x = g; else if (l == 1) r.y = g; else if (l == 2) r.z = g; else r.w = g;
            }
            o[v] = r;
            cell += step_q; f += step_r;
            if (f >= F) { f -= F; ++cell; }
        }
    }
}

int main() {
    const int F = 85;
    const long long cells = 64LL * 3 * 80 * 80;       // level 0 of the headline: 1 228 800 cells, 417.8 MB
    float *pi, *gpi, *tobj, *sink;
    cudaMalloc(&pi, cells * F * 4); cudaMalloc(&gpi, cells * F * 4); cudaMalloc(&tobj, cells * 4); cudaMalloc(&sink, 4);
    cudaMemset(pi, 0, cells * F * 4); cudaMemset(tobj, 0, cells * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double mb = cells * F * 4 / 1e6;
    auto timeit = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %7.1f us  %6.0f GB/s written  (%s)\n", name, ms / 20 * 1e3, mb / (ms / 20 * 1e3) * 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    for (int grid : {1184, 2368, 4800}) {
        printf("grid %d\n", grid);
        timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(gpi, 0, cells * F * 4); });
        timeit("fill only", [&] { fill_only<<<grid, 256>>>(gpi, F, cells); });
        timeit("fill + strided column reads", [&] { fill_read<<<grid, 256>>>(pi, tobj, gpi, sink, F, cells); });
        timeit("fill + reads + patch (sub-block 256)", [&] { fill_read_patch<<<grid, 256>>>(pi, tobj, gpi, F, cells, 256); });
        timeit("fill + reads + patch (sub-block 64)", [&] { fill_read_patch<<<grid, 256>>>(pi, tobj, gpi, F, cells, 64); });
        timeit("single-pass compose", [&] { compose<<<grid, 256>>>(pi, tobj, gpi, F, cells); });
    }
    return 0;
}
