// FFMA vs FFMA2 (fma.rn.f32x2, sm_100+) throughput: same number of fp32 FMAs issued as scalar or packed instructions.
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void __launch_bounds__(1024, 1) k(float* out, float w, int iters, long long* cyc) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 ww = make_float2(w, w * 0.5f), cc = make_float2(0.25f, -0.25f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) a[i] = __ffma2_rn(a[i], ww, cc);
            else { a[i].x = fmaf(a[i].x, ww.x, cc.x); a[i].y = fmaf(a[i].y, ww.y, cc.y); }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* o; long long* c; cudaMalloc(&o, 1024 * 4); cudaMalloc(&c, 8);
    long long h;
    for (int warps : {4, 8, 16, 32}) {
        for (int packed = 0; packed < 2; ++packed) {
            if (packed) k<true><<<1, warps * 32>>>(o, 0.999f, 4096, c); else k<false><<<1, warps * 32>>>(o, 0.999f, 4096, c);
            cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%2d warps, %s: %.2f fp32 FMA / cycle / SM\n", warps, packed ? "FFMA2" : "FFMA ", warps * 32.0 * 16 * 4096 / h);
        }
    }
    return 0;
}
