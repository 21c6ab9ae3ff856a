// global_head of ChessSquareCNN (models/square.py:34-41,107-113) on the tensor cores:
//
//     hidden = ReLU(features (B, 30720) @ W_g^T (30720, 64) + b_g);  turn = hidden @ w_t + b_t;  castling = hidden @ W_c^T + b_c
//
// The 30720 -> 64 projection is 99.98 % of the head's work: a skinny GEMM with M = boards, N = 64, K = 30720 whose A operand
// (the pooled trunk features, 120 KB of fp32 per board) is the only large stream of the whole path after the uint8 boards.
// It runs as a split-K tcgen05 GEMM in kind::tf32 with SPLIT OPERANDS (fp32 accumulation in TMEM): a tf32 operand keeps 10
// explicit significand bits (the tensor core ignores the low 13 bits of the fp32 word), which alone put 3-5e-3 of relative error
// on the turn / castling logits (these dot products cancel heavily) -- so features and weights are held as x = hi + lo with
// hi = x & 0xffffe000 and lo = x - hi (exact), and every k-step issues  A_hi W_hi + A_lo W_hi + A_hi W_lo  (the dropped
// A_lo W_lo term is 2^-22 relative): fp32-grade products at three MMAs per k-step of a GEMM that is memory-bound anyway.
// The fused tail (kernels_backend.cu, stage D) writes the two feature planes directly in the operand layout
//
//     FT[m_tile][plane hi | lo][k/4][128 rows = boards][4 floats]     (K-major / no-swizzle UMMA tiles per 128 boards)
//
// so a K-block of 32 features x 128 boards of one plane is one contiguous 16 KB block = ONE TMA bulk copy; the weights are
// pre-tiled the same way ([plane][k/4][64][4]).  Grid = (board tiles) x (K splits); partial sums go to a small fp32 buffer and a
// finishing kernel adds them in a fixed order (deterministic), applies bias + ReLU and the 64 -> 1 + 4 output layers.
#include "internal.h"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int KCH = 7680;                  // 30720 / 4: 16-byte K chunks
constexpr int KB_CHUNKS = 8;               // chunks per pipeline stage (32 features)
constexpr int A_PLANE = KB_CHUNKS * 2048;  // 16384 per plane
constexpr int B_PLANE = KB_CHUNKS * 1024;  // 8192 per plane
constexpr int A_STAGE = 2 * A_PLANE;       // hi | lo
constexpr int B_STAGE = 2 * B_PLANE;
constexpr int STAGES = 4;
constexpr int OFF_B = STAGES * A_STAGE;
constexpr int OFF_BAR = OFF_B + STAGES * B_STAGE;    // 196608
constexpr int SMEM = OFF_BAR + 128;

// kind::tf32: D = f32, A = B = tf32 (format 2), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct HeadParams {
    const float* ft;         // FT features: [m_tile][2 planes][KCH][128][4]
    const float* wt;         // tiled global_head weight [2 planes][KCH][64][4]
    float* partial;          // [ksplit][m_tiles*128][64]
    int m_tiles, ksplit;
};

// 6 warps: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue
__global__ void __launch_bounds__(192, 1) global_head_umma_kernel(const __grid_constant__ HeadParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t *full = bars, *empty = bars + STAGES, *done = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % p.m_tiles, ks = blockIdx.x / p.m_tiles;
    const int blocks = (KCH / KB_CHUNKS) / p.ksplit;          // K-blocks of this CTA
    const int chunk0 = ks * blocks * KB_CHUNKS;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        if (lane == 0) {
            const uint8_t* a = reinterpret_cast<const uint8_t*>(p.ft) + ((size_t)mt * 2 * KCH + chunk0) * 2048;
            const uint8_t* b = reinterpret_cast<const uint8_t*>(p.wt) + (size_t)chunk0 * 1024;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < blocks; ++kb) {
                mbar_wait(empty + stage, phase ^ 1u);
                mbar_arrive_expect_tx(full + stage, A_STAGE + B_STAGE);
                for (int pl = 0; pl < 2; ++pl) {
                    bulk_g2s(smem + stage * A_STAGE + pl * A_PLANE, a + ((size_t)pl * KCH * 2048) + (size_t)kb * A_PLANE, A_PLANE, full + stage);
                    bulk_g2s(smem + OFF_B + stage * B_STAGE + pl * B_PLANE, b + ((size_t)pl * KCH * 1024) + (size_t)kb * B_PLANE, B_PLANE, full + stage);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32(128, 64);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < blocks; ++kb) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(smem + stage * A_STAGE), b0 = smem_u32(smem + OFF_B + stage * B_STAGE);
#pragma unroll
                for (int j = 0; j < KB_CHUNKS / 2; ++j) {      // one MMA = K 8 tf32 = two 16-byte chunks; three terms per k-step
                    const uint64_t ah = make_smem_desc(a0 + j * 4096, 2048, 128), al = make_smem_desc(a0 + A_PLANE + j * 4096, 2048, 128);
                    const uint64_t bh = make_smem_desc(b0 + j * 2048, 1024, 128), bl = make_smem_desc(b0 + B_PLANE + j * 2048, 1024, 128);
                    mma_tf32_ss(tmem, al, bh, idesc, (kb | j) ? 1u : 0u);       // small terms first
                    mma_tf32_ss(tmem, ah, bl, idesc, 1u);
                    mma_tf32_ss(tmem, ah, bh, idesc, 1u);
                }
                mma_commit(empty + stage);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    } else {
        const int q = warp & 3, row = q * 32 + lane;              // TMEM lane quadrant a warp may read = warp id % 4
        mbar_wait(done, 0);
        tc_fence_after();
        float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)ks * p.m_tiles * 128 + (size_t)mt * 128 + row) * 64);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[c * 4 + i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                             __uint_as_float(r[4 * i + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

// One warp per board: fixed-order sum of the K-split partials + bias, ReLU, then the 64 -> 1 (turn) and 64 -> 4 (castling) layers.
__global__ void __launch_bounds__(256) global_head_finish_kernel(const float* __restrict__ partial, int ksplit, int m_rows,
                                                                 const float* __restrict__ gb, const float* __restrict__ tc_w,
                                                                 const float* __restrict__ tc_b, int B, float* __restrict__ turn,
                                                                 float* __restrict__ castling) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    float h0 = 0.f, h1 = 0.f;
    for (int s = 0; s < ksplit; ++s) {
        const float* q = partial + ((size_t)s * m_rows + b) * 64;
        h0 += q[lane];
        h1 += q[lane + 32];
    }
    h0 = fmaxf(h0 + gb[lane], 0.f);
    h1 = fmaxf(h1 + gb[lane + 32], 0.f);
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        float v = fmaf(h0, tc_w[r * 64 + lane], h1 * tc_w[r * 64 + lane + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            v += tc_b[r];
            if (r == 0) turn[b] = v; else castling[(size_t)b * 4 + (r - 1)] = v;
        }
    }
}

__global__ void tile_glob_w_kernel(const float* __restrict__ w /*[64][30720]*/, float* __restrict__ wt /*[2][KCH][64][4]*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 30720) return;
    const int kk = i & 3, n = (i >> 2) & 63, chunk = i >> 8;
    const float v = w[(size_t)n * 30720 + chunk * 4 + kk];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    wt[i] = hi;
    wt[(size_t)64 * 30720 + i] = v - hi;
}

// FT -> row-major [board*64 + square][480] (only when the caller asked for the features)
__global__ void untile_features_kernel(const float4* __restrict__ ft, float4* __restrict__ out, int B) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // over B * 7680 float4
    if (i >= (int64_t)B * KCH) return;
    const int b = (int)(i / KCH), chunk = (int)(i - (int64_t)b * KCH);
    const float4 h = ft[((size_t)(b >> 7) * 2 * KCH + chunk) * 128 + (b & 127)], l = ft[((size_t)((b >> 7) * 2 + 1) * KCH + chunk) * 128 + (b & 127)];
    out[i] = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);            // hi + lo is exact
}

}  // namespace

int launch_tile_glob_w(const float* glob_w, float* wt, cudaStream_t s) {
    tile_glob_w_kernel<<<(64 * 30720 + 255) / 256, 256, 0, s>>>(glob_w, wt);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

// K is always cut into the same KSPLIT ranges, whatever the batch size: a board's logits do not depend on how many other
// boards are in the call (and the fixed-order finishing sum keeps them bit-reproducible).
// 32 ranges of 30 K-blocks: a single board (one M tile) still spreads the 15.7 MB weight stream (hi + lo planes) over 32 SMs (the
// largest kernel of a one-board call), and a 4096-board chunk gets 1024 CTAs on 148 SMs.
constexpr int KSPLIT = 32;
static_assert((KCH / KB_CHUNKS) % KSPLIT == 0, "K blocks must divide evenly over the splits");
size_t global_head_partial_floats(int B, int num_sms) {
    (void)num_sms;
    return (size_t)KSPLIT * ((B + 127) / 128) * 128 * 64;
}

int launch_global_head_umma(const float* ft, const float* wt, float* partial, const float* glob_b, const float* tc_w, const float* tc_b,
                            int B, int num_sms, float* turn, float* castling, cudaStream_t s) {
    if (B == 0) return CV_OK;
    const int m_tiles = (B + 127) / 128;
    const int ksplit = KSPLIT;
    (void)num_sms;
    HeadParams p{ft, wt, partial, m_tiles, ksplit};
    CV_CUDA(cudaFuncSetAttribute(global_head_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    global_head_umma_kernel<<<m_tiles * ksplit, 192, SMEM, s>>>(p);
    CV_CHECK_LAUNCH();
    global_head_finish_kernel<<<(B + 7) / 8, 256, 0, s>>>(partial, ksplit, m_tiles * 128, glob_b, tc_w, tc_b, B, turn, castling);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_untile_features(const float* ft, float* out, int B, cudaStream_t s) {
    const int64_t n = (int64_t)B * KCH;
    if (n == 0) return CV_OK;
    untile_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(ft), reinterpret_cast<float4*>(out), B);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
