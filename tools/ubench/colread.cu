// Micro-benchmark: cost of reading ONE 4-byte column of a [cells, F] fp32 row array (F = 85: a 340-byte stride), the access the
// YOLOv5 objectness loss makes (targets.cu: v5_loss_obj_fwd_body).  Variants: loads in flight per thread, ld.global.nc vs plain,
// with / without the per-cell store.        nvcc -O3 -arch=sm_100a -o colread colread.cu && ./colread
#include <cstdio>
#include <cuda_runtime.h>

template <int U, bool NC, bool STORE>
__global__ void __launch_bounds__(256) colread(const float* pi, const float* __restrict__ tobj, float* __restrict__ out,
                                               float* __restrict__ sink, int F, long long cells) {
    const long long stride = (long long)gridDim.x * 256;
    float acc = 0.f;
    for (long long c = (long long)blockIdx.x * 256 + threadIdx.x; c < cells; c += U * stride) {
        float x[U], t[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long long ck = c + k * stride;
            x[k] = 0.f; t[k] = 0.f;
            if (ck < cells) {
                x[k] = NC ? __ldg(pi + ck * F + 4) : *(volatile const float*)(pi + ck * F + 4);
                t[k] = tobj[ck];
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long long ck = c + k * stride;
            if (ck < cells) {
                acc += x[k] * t[k];
                if (STORE) out[ck] = x[k] + t[k];
            }
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

int main() {
    const int F = 85;
    const long long cells = 64LL * 3 * 80 * 80;
    float *pi, *tobj, *out, *sink, *flush;
    cudaMalloc(&pi, cells * F * 4); cudaMalloc(&tobj, cells * 4); cudaMalloc(&out, cells * 4); cudaMalloc(&sink, 4);
    cudaMalloc(&flush, 256 << 20);
    cudaMemset(pi, 0, cells * F * 4); cudaMemset(tobj, 0, cells * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto launch) {
        float tot = 0.f;
        for (int i = 0; i < 12; ++i) {
            cudaMemsetAsync(flush, i, 256 << 20);                   // L2 holds nothing of pi
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 2) tot += ms;
        }
        printf("%-52s %7.1f us  (%s)\n", name, tot / 10 * 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    const int grid = 1184;
    for (int gran : {0, 32, 64, 128}) {
    if (gran) {
        cudaError_t rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("-- cudaLimitMaxL2FetchGranularity = %d (%s; now %zu)\n", gran, cudaGetErrorString(rc), got);
    } else {
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("-- default cudaLimitMaxL2FetchGranularity = %zu\n", got);
    }
    timeit("1 load in flight, ld.global.nc, no store", [&] { colread<1, true, false><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    timeit("4 loads in flight, ld.global.nc, no store", [&] { colread<4, true, false><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    timeit("1 load in flight, plain ld, no store", [&] { colread<1, false, false><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    timeit("4 loads in flight, plain ld, no store", [&] { colread<4, false, false><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    timeit("1 load in flight, ld.global.nc, store per cell", [&] { colread<1, true, true><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    timeit("4 loads in flight, ld.global.nc, store per cell", [&] { colread<4, true, true><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    timeit("4 loads in flight, plain ld, store per cell", [&] { colread<4, false, true><<<grid, 256>>>(pi, tobj, out, sink, F, cells); });
    }
    return 0;
}
