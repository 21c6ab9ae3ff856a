// Micro-benchmark (one CTA): what an epilogue warp pays per accumulator tile while the tensor pipe is busy.
//   * latency of {tcgen05.fence::after, 2 x tcgen05.ld 32x32b.x16, tcgen05.wait::ld} with 0 / N MMAs queued ahead of it;
//   * latency from "4 MMAs + tcgen05.commit issued" to an mbarrier waiter waking up (try_wait and test_wait spin);
//   * latency of fence.proxy.async.shared::cta after 4 STS.128.
// Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I chess_vision_b200/csrc tools/ubench_ldlat.cu -o /tmp/ubench_ldlat && /tmp/ubench_ldlat
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace umma;

// mode bit 0: MMA stream on; bit 1: waiter spins (test_wait); burst = MMAs per commit
__global__ void __launch_bounds__(160, 1) lat_bench(int mode, int burst, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_tile, bar_go, bar_done;
    __shared__ uint32_t slot;
    __shared__ volatile long long t_commit[64];
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar_tile, 1); mbar_init(&bar_go, 1); mbar_init(&bar_done, 4); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t a0 = smem_u32(smem) + 1024, b0 = smem_u32(smem) + 160 * 1024;
            const uint32_t idesc = make_idesc_bf16(128, 32);
            const uint64_t bd = make_smem_desc(b0, 32 * 16, 128);
            for (int r = 0; r < reps; ++r) {
                // tile: 4 MMAs into columns [0,32) + commit -> the waiters' barrier
                for (int j = 0; j < 4; ++j) mma_bf16_ss(tm, make_smem_desc(a0 + j * 2048, 2048, 128), bd, idesc, j ? 1u : 0u);
                mma_commit(&bar_tile);
                t_commit[r] = clock64();
                // followed by `burst` more MMAs into other columns (the work the pipe has queued while the epilogue runs)
                if (mode & 1)
                    for (int j = 0; j < burst; ++j) mma_bf16_ss(tm + 64 + (j & 3) * 32, make_smem_desc(a0 + (j & 7) * 2048, 2048, 128), bd, idesc, 1u);
                mma_commit(&bar_go);
                mbar_wait_spin(&bar_go, r & 1);
                mbar_wait_spin(&bar_done, r & 1);          // the four waiter warps are done with this round
            }
        }
        __syncwarp();
    } else {
        long long acc_wake = 0, acc_ld = 0, acc_sts = 0;
        const uint32_t trow = tm + ((uint32_t)((warp - 1) * 32) << 16);
        uint32_t sink = 0;
        for (int r = 0; r < reps; ++r) {
            if (mode & 2) { if (lane == 0) mbar_wait_spin(&bar_tile, r & 1); __syncwarp(); } else mbar_wait(&bar_tile, r & 1);
            const long long t0 = clock64();
            tc_fence_after();
            uint32_t r0[16], r1[16];
            tmem_ld16(trow, r0);
            tmem_ld16(trow + 16, r1);
            tmem_ld_wait();
            const long long t1 = clock64();
            tc_fence_before();
#pragma unroll
            for (int k = 0; k < 16; ++k) sink += r0[k] ^ r1[k];
            uint4* dst = reinterpret_cast<uint4*>(smem + 64 * 1024) + threadIdx.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c * 256] = make_uint4(sink, r0[c], r1[c], c);
            fence_proxy_async_smem();
            __syncwarp();
            const long long t2 = clock64();
            if (r > 0) { acc_wake += t0 - t_commit[r]; acc_ld += t1 - t0; acc_sts += t2 - t1; }
            if (lane == 0) mbar_arrive(&bar_done);
        }
        if (lane == 0) { out[(warp - 1) * 4] = acc_wake / (reps - 1); out[(warp - 1) * 4 + 1] = acc_ld / (reps - 1); out[(warp - 1) * 4 + 2] = acc_sts / (reps - 1); }
        if (sink == 0x12345u) out[63] = sink;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// Round trip of an op-synchronous GEMM step as the fused stages B-D run it: CTA barrier, one elected thread issues n_mma MMAs
// (M=128, N, K=16) + commit, ALL warps wait on the mbarrier.  Reports cycles from the barrier to "awake" for thread 0 (issuer's warp)
// and for the last warp.  mode 1: only warp 0 polls the mbarrier, the others wait at a second CTA barrier.
__global__ void __launch_bounds__(512, 1) roundtrip_bench(int n_mma, int N, int mode, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t a0 = smem_u32(smem) + 1024, b0 = smem_u32(smem) + 160 * 1024;
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t bd = make_smem_desc(b0, N * 16, 128);
    long long acc = 0, acc_issue = 0;
    for (int r = 0; r < reps; ++r) {
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        const long long t0 = clock64();
        long long t_issue = t0;
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            for (int j = 0; j < n_mma; ++j) mma_bf16_ss(tm + (j & 1) * N, make_smem_desc(a0 + (j & 7) * 2048, 2048, 128), bd, idesc, j > 1 ? 1u : 0u);
            mma_commit(&bar);
            t_issue = clock64();
        }
        if (mode == 0 || warp == 0) mbar_wait(&bar, r & 1);
        if (mode == 1) __syncthreads();
        tc_fence_after();
        const long long t1 = clock64();
        if (r > 0) { acc += t1 - t0; acc_issue += t_issue - t0; }
    }
    if (threadIdx.x == 0) { out[0] = acc / (reps - 1); }
    if (threadIdx.x == 511) out[1] = acc / (reps - 1);
    if (warp == 0 && acc_issue) out[2] = acc_issue / (reps - 1);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 64 * 8);
    long long h[64];
    cudaFuncSetAttribute(lat_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode : {0, 1, 2, 3}) for (int burst : {0, 8, 18, 36}) {
        if (!(mode & 1) && burst) continue;
        lat_bench<<<1, 160, 200 * 1024>>>(mode, burst, 32, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lat: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
        printf("%s waiter, %2d MMAs queued behind the tile: commit->wake %4lld cyc, fence+2xld16+wait %4lld cyc, 4xSTS.128+fence.proxy.async %4lld cyc (warp 1; warp 4: %lld / %lld / %lld)\n",
               (mode & 2) ? "spin    " : "try_wait", (mode & 1) ? burst : 0, h[0], h[1], h[2], h[12], h[13], h[14]);
    }
    cudaFuncSetAttribute(roundtrip_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode : {0, 1}) for (int N : {48, 96}) for (int n : {0, 4, 8, 12, 24}) {
        cudaMemset(d, 0, 64 * 8);
        roundtrip_bench<<<1, 512, 200 * 1024>>>(n, N, mode, 32, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("roundtrip: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 3 * 8, cudaMemcpyDeviceToHost);
        printf("round trip, %s, %2d MMAs N=%2d: barrier -> awake %5lld cyc (thread 0), %5lld (last warp); issue + commit %4lld cyc\n",
               mode ? "warp 0 polls + CTA barrier" : "all 16 warps poll         ", n, N, h[0], h[1], h[2]);
    }
    return 0;
}
