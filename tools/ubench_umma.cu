// Micro-benchmarks behind the fused-kernel design decisions (one CTA on one SM):
//   * tcgen05.mma kind::f16 issue-to-complete cycles per instruction for the shapes the front end uses (M=128, K=16, small N),
//     with the A operand 128-byte aligned or shifted by one 16-byte row (the "Dx = -1" filter taps), same / rotating accumulators;
//   * tcgen05.ld (32x32b.x16) throughput with 4 / 8 / 16 warps.
// Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I chess_vision_b200/csrc tools/ubench_umma.cu -o /tmp/ubench && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace umma;

struct MmaCase { int N, a_shift_bytes, lbo, sbo, n_acc, n_mma, a_stride; int lsu_warps = 0; int commit_every = 0; int use_elect = 0; int alt = 0; int fence_every = 0; int wait_every = 0; };

template <int COMMIT_EVERY, bool ELECT>
__global__ void __launch_bounds__(544, 1) mma_bench(MmaCase c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32 && (ELECT ? elect_one() : threadIdx.x == 0)) {
        const uint32_t a0 = smem_u32(smem) + 1024 + c.a_shift_bytes, b0 = smem_u32(smem) + 160 * 1024;
        const uint32_t idesc = make_idesc_bf16(128, c.N), idesc_h = c.alt ? make_idesc_f16(128, c.N) : idesc;
        // descriptors and accumulator addresses precomputed: the issue loop is one UTCHMMA + nothing else per instruction
        uint64_t ad[8], bd = make_smem_desc(b0, c.N * 16, 128);
        uint32_t dc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ad[j] = make_smem_desc(a0 + j * c.a_stride, c.lbo, c.sbo); dc[j] = tm + (j % c.n_acc) * c.N; }
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < c.n_mma; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c.fence_every && (j & 3) == 0) tc_fence_after();
                    if (c.wait_every && (j & 3) == 0) { (void)mbar_try_wait(&bar2, 1u); }      // one (satisfied) mbarrier poll per tile, like a real consumer wait
                    mma_bf16_ss(dc[j], ad[j], bd, (j & 4) ? idesc_h : idesc, (COMMIT_EVERY && (j & 3) == 0) ? 0u : 1u);     // alt: kind flips every 4 MMAs
                    if (COMMIT_EVERY && (j & 3) == 3) mma_commit(&bar2);       // a barrier nobody waits on: cost of the commit itself
                }
            }
            const long long t1 = clock64();
            mma_commit(&bar);
            mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + 32 * c.lsu_warps) {
        // interference: conflict-free 16-byte loads + stores on a private region for the whole duration of the MMA loop
        uint4* q = reinterpret_cast<uint4*>(smem + 100 * 1024) + (threadIdx.x - 32);
        uint4 v = make_uint4(0, 0, 0, 0);
        for (int i = 0; i < 3000; ++i) {
            const uint4 a = q[(i & 3) * 512];
            v.x += a.x; v.y ^= a.y;
            q[((i + 1) & 3) * 512] = v;
        }
        if (v.x == 0x12345u) out[7] = v.y;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// Several warps issue MMAs concurrently (one thread each, own accumulators, own barrier): is the 48-cycle floor the tensor pipe or
// the single issuing thread?
__global__ void __launch_bounds__(128, 1) mma_multi_bench(MmaCase c, int n_issuers, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < n_issuers) {
        const uint32_t a0 = smem_u32(smem) + 1024 + c.a_shift_bytes + w * 32768, b0 = smem_u32(smem) + 160 * 1024;
        const uint32_t idesc = make_idesc_bf16(128, c.N);
        uint64_t ad[8], bd = make_smem_desc(b0, c.N * 16, 128);
        uint32_t dc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ad[j] = make_smem_desc(a0 + j * c.a_stride, c.lbo, c.sbo); dc[j] = tm + w * 128 + (j % c.n_acc) * c.N; }
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < c.n_mma; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) mma_bf16_ss(dc[j], ad[j], bd, idesc, 1u);
            }
            const long long t1 = clock64();
            mma_commit(&bar[w]);
            mbar_wait(&bar[w], rep & 1);
            const long long t2 = clock64();
            out[2 * w] = t1 - t0;
            out[2 * w + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int X32>
__global__ void __launch_bounds__(512, 1) ld_bench(int iters, long long* out) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r0[16], r1[16];
        tmem_ld16(tm + lane_base + ((i * 32) & 255) + (warp >> 2) * 64, r0);
        tmem_ld16(tm + lane_base + ((i * 32 + 16) & 255) + (warp >> 2) * 64, r1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += r0[k] ^ r1[k];
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    if (acc == 0x12345678u) out[63] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 64 * 8);
    long long h[64];
    cudaFuncSetAttribute(mma_bench<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(mma_bench<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(mma_bench<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const MmaCase cases[] = {
        // N, shift, lbo, sbo, n_acc, n_mma, a_stride, lsu_warps, commit_every, use_elect, alt, fence_every, wait_every
        {32, 0, 2048, 128, 4, 64, 2048, 0, 4, 1, 0, 0, 0}, {32, 0, 2048, 128, 4, 64, 2048, 0, 4, 1, 0, 0, 0}, {32, 0, 2048, 128, 4, 64, 2048, 0, 4, 1, 0, 1, 0},
        {32, 0, 2048, 128, 4, 64, 2048, 0, 4, 1, 0, 0, 1}, {32, 0, 2048, 128, 4, 64, 2048, 0, 4, 1, 0, 1, 1},
    };
    for (const MmaCase& c : cases) {
        if (!c.use_elect) mma_bench<0, false><<<1, 32 + 32 * c.lsu_warps, 200 * 1024>>>(c, d);
        else if (!c.commit_every) mma_bench<0, true><<<1, 32 + 32 * c.lsu_warps, 200 * 1024>>>(c, d);
        else mma_bench<4, true><<<1, 32 + 32 * c.lsu_warps, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma N=%d: %s\n", c.N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mma M=128 N=%3d K=16 a_shift=%2d lbo=%5d sbo=%3d n_acc=%d a_stride=%4d lsu_warps=%2d commit_every=%d elect=%d alt=%d fence=%d wait=%d: issue %5.1f cyc/mma, complete %6.1f cyc/mma\n", c.N, c.a_shift_bytes, c.lbo, c.sbo, c.n_acc,
               c.a_stride, c.lsu_warps, c.commit_every, c.use_elect, c.alt, c.fence_every, c.wait_every, (double)h[0] / c.n_mma, (double)h[1] / c.n_mma);
    }
    cudaFuncSetAttribute(mma_multi_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {32, 64}) for (int ni : {1, 2, 4}) {
        const MmaCase c = {N, 0, 2048, 128, 2, 64, 2048};
        mma_multi_bench<<<1, 128, 200 * 1024>>>(c, ni, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("multi: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < ni; ++w) mx = mx > h[2 * w + 1] ? mx : h[2 * w + 1];
        printf("multi-issuer N=%d: %d warps x 64 MMAs -> %6.1f cyc per MMA overall (%.0f total)\n", N, ni, (double)mx / (64.0 * ni), (double)mx);
    }
    for (int nw : {1, 4, 8, 16}) {
        ld_bench<0><<<1, nw * 32>>>(256, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ld: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < nw; ++w) mx = mx > h[w] ? mx : h[w];
        printf("tcgen05.ld 32x32b.x16 x2 per iter, %2d warps: %6.1f cyc/iter/warp -> %6.1f B/cyc/SM\n", nw, (double)mx / 256, 256.0 * nw * 4096 / mx);
    }
    return 0;
}
