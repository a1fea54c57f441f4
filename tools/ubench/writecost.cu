// Micro-benchmark: what do the candidate writes cost next to the 418 MB strided read stream?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// WMODE 0: no writes; 1: coalesced float4 writes (36 B/cell as 2.25 float4 per cell -> we write 2 float4 + 1 float per cell,
// lanes contiguous); 2: same but to a small (1 MB) recycled buffer; 3: coalesced with st.global.cs; 4: only 16 B/cell
template <int WMODE>
__global__ void __launch_bounds__(128, 6) rd_wr(const float* __restrict__ p, int GG, int F, float4* w4a, float4* w4b, float* w1,
                                                size_t wmask, float* out) {
    const int slab = blockIdx.y, tile = blockIdx.x;
    const int cell = tile * 512 + threadIdx.x * 4;
    float acc = 0.f;
    if (cell < GG) {
        const float* base = p + (size_t)slab * F * GG + cell;
        int f = 0;
        for (; f + 8 <= F; f += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(base + (size_t)(f + u) * GG));
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
        }
        for (; f < F; ++f) acc += __ldcs(reinterpret_cast<const float4*>(base + (size_t)f * GG)).y;
    }
    if (WMODE != 0) {
        const size_t tbase = ((size_t)slab * GG + (size_t)tile * 512);
        for (int k = 0; k < 4; ++k) {                       // 512 cells: lanes write consecutive elements
            const size_t i = (tbase + k * 128 + threadIdx.x) & wmask;
            const float4 val = make_float4(acc, acc, acc, acc);
            if (WMODE == 3) { __stcs(w4a + i, val); if (true) __stcs(w4b + i, val); __stcs(w1 + i, acc); }
            else if (WMODE == 4) { w4a[i] = val; }
            else { w4a[i] = val; w4b[i] = val; w1[i] = acc; }
        }
    }
    if (acc == 123.456f) *out = acc;
}

template <int WMODE>
static float run(const float* d, int GG, int F, int B, int A, float4* a, float4* b, float* c, size_t wmask, float* o) {
    dim3 g((GG + 511) / 512, B * A);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) rd_wr<WMODE><<<g, 128>>>(d, GG, F, a, b, c, wmask, o);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) rd_wr<WMODE><<<g, 128>>>(d, GG, F, a, b, c, wmask, o);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 100;   // us per launch
}

int main() {
    const int B = 64, A = 3, F = 85, GG = 6400;
    const size_t n = (size_t)B * A * F * GG, cells = (size_t)B * A * GG;
    float *d, *o, *w1; float4 *wa, *wb;
    CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&o, 4)); CK(cudaMemset(d, 0, n * 4));
    size_t cap = 1; while (cap < cells) cap <<= 1;
    CK(cudaMalloc(&wa, cap * 16)); CK(cudaMalloc(&wb, cap * 16)); CK(cudaMalloc(&w1, cap * 4));
    printf("reads only                      : %.1f us\n", run<0>(d, GG, F, B, A, wa, wb, w1, cap - 1, o));
    printf("+ 36 B/cell coalesced           : %.1f us\n", run<1>(d, GG, F, B, A, wa, wb, w1, cap - 1, o));
    printf("+ 36 B/cell into a 1 MB window  : %.1f us\n", run<2>(d, GG, F, B, A, wa, wb, w1, (1u << 15) - 1, o));
    printf("+ 36 B/cell coalesced, st.cs    : %.1f us\n", run<3>(d, GG, F, B, A, wa, wb, w1, cap - 1, o));
    printf("+ 16 B/cell coalesced           : %.1f us\n", run<4>(d, GG, F, B, A, wa, wb, w1, cap - 1, o));
    return 0;
}
