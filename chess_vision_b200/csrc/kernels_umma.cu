// bf16 tensor-core kernels of the ChessSquareCNN trunk for sm_100a: tcgen05.mma with TMEM accumulators,
// operands staged in shared memory by TMA bulk copies, warp-specialised persistent CTAs (one per SM).
//
//   pointwise_umma_kernel   the 27 pointwise 1x1 convs (+folded BN, ReLU, residual)        timm trunk, square.py:86
//   dense_umma_kernel       the 3 dense 3x3 stride-2 convs as implicit GEMM (software im2col gather)
//   depthwise_t8_kernel     the 15 depthwise convs, 16-byte vectorised over channels (CUDA cores, HBM/L2-bound)
//
// Data layout ("T8", internal.h): activations live in HBM as [M/128][C/8][128 rows][8 ch] bf16, so a 128-row
// tile of a layer input is one contiguous block that already IS the K-major / no-swizzle UMMA operand image:
// the producer warp moves it with a single cp.async.bulk, and the epilogue writes 16-byte channel chunks with
// consecutive rows at consecutive addresses (fully coalesced, directly consumable by the next layer).
//
// GEMM view: D[128 rows, N=Cout] = A[128 rows, K] * W[N, K]^T, fp32 accumulate in TMEM (lane = row, column = n).
// K is tiny (16..288) so the whole K extent of a tile is one pipeline stage; the pipeline runs over M tiles:
//   producer warp  : TMA bulk load of A tile t+s           (full/empty mbarriers, kStages deep)
//   MMA warp       : K/16 tcgen05.mma per N sub-tile, tcgen05.commit -> frees the stage, publishes the accumulator
//   epilogue warps : tcgen05.ld -> +bias, ReLU, +residual -> bf16 -> coalesced 16 B stores (double-buffered TMEM)
#include "internal.h"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int TILE_M = 128;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_BUDGET = 200 * 1024;

struct GemmParams {
    const bf16* x;        // A source: T8 [M][K] (pointwise) | T8 [Min][Cin] or row-major [.,3] crops (dense)
    const bf16* wimg;     // B image [K/8][N][8]
    const float* bias;    // [N]
    const bf16* skip;     // T8 [M][N] or nullptr
    bf16* y;              // T8 [M][N]
    int m_tiles, K, N, relu, stages, n_split, n_tile, num_acc;
    int w_parts;          // 1: bf16 weights; 2: W = W_hi + W_lo (two images back to back), two MMAs per k-step
    int hin, hout, cin;   // dense only
};

struct SmemPlan {
    uint32_t b_bytes, a_bytes, off_a, off_bias, off_bar, total;
};
__host__ __device__ inline SmemPlan plan_smem(int K, int N, int stages, int w_parts) {
    SmemPlan s;
    s.b_bytes = (uint32_t)N * K * 2 * w_parts;
    s.a_bytes = (uint32_t)TILE_M * K * 2;
    s.off_a = (s.b_bytes + 127u) & ~127u;
    s.off_bias = s.off_a + stages * s.a_bytes;
    s.off_bar = (s.off_bias + N * 4 + 15u) & ~15u;
    s.total = s.off_bar + (2 * stages + 5) * 8 + 16;
    return s;
}

struct Pipe {            // smem pointers shared by all roles
    uint8_t* b;
    uint8_t* a;
    float* bias;
    uint64_t *full, *empty, *tfull, *tempty, *wbar;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ Pipe carve(uint8_t* smem, const GemmParams& p) {
    SmemPlan s = plan_smem(p.K, p.N, p.stages, p.w_parts);
    Pipe q;
    q.b = smem;
    q.a = smem + s.off_a;
    q.bias = reinterpret_cast<float*>(smem + s.off_bias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + s.off_bar);
    q.full = bars;
    q.empty = bars + p.stages;
    q.tfull = bars + 2 * p.stages;
    q.tempty = q.tfull + 2;
    q.wbar = q.tempty + 2;
    q.tmem_slot = reinterpret_cast<uint32_t*>(q.wbar + 1);
    return q;
}

// ---- MMA issuer: one warp, one elected lane issues --------------------------------------------------------
__device__ __forceinline__ void mma_role(const GemmParams& p, const Pipe& q, uint32_t tmem_base, int lane) {
    const uint32_t idesc = make_idesc_bf16(TILE_M, p.n_tile);
    const uint32_t a_lbo = TILE_M * 16, b_lbo = (uint32_t)p.N * 16;   // byte distance between K-adjacent core matrices
    const SmemPlan s = plan_smem(p.K, p.N, p.stages, p.w_parts);
    mbar_wait(q.wbar, 0);                                             // weights landed
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(q.full + stage, phase);
        mbar_wait(q.tempty + acc, acc_phase ^ 1u);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t a_base = smem_u32(q.a + (size_t)stage * s.a_bytes);
            const uint32_t b_base = smem_u32(q.b);
            for (int nt = 0; nt < p.n_split; ++nt) {
                const uint32_t d = tmem_base + (uint32_t)(acc * 256 + nt * p.n_tile);
                const uint32_t part_bytes = (uint32_t)p.N * p.K * 2;
                for (int k = 0; k < p.K / 16; ++k) {
                    const uint64_t ad = make_smem_desc(a_base + k * 2 * a_lbo, a_lbo, 128);
                    for (int part = 0; part < p.w_parts; ++part) {
                        const uint64_t bd = make_smem_desc(b_base + part * part_bytes + k * 2 * b_lbo + nt * p.n_tile * 16, b_lbo, 128);
                        mma_bf16_ss(d, ad, bd, idesc, (k | part) ? 1u : 0u);
                    }
                }
            }
            mma_commit(q.empty + stage);     // smem stage reusable once these MMAs have read it
            mma_commit(q.tfull + acc);       // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        if (p.num_acc == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1u; } else { acc_phase ^= 1u; }
    }
}

// ---- epilogue: 4 warps, warp w owns TMEM lanes [32w, 32w+32) = tile rows ------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}

__device__ __forceinline__ void epilogue_role(const GemmParams& p, const Pipe& q, uint32_t tmem_base, int warp, int lane) {
    const int row = warp * 32 + lane;
    const int n8 = p.N >> 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(q.tfull + acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 256);
        uint4* yt = reinterpret_cast<uint4*>(p.y) + ((size_t)tile * n8) * TILE_M + row;
        const uint4* st = p.skip ? reinterpret_cast<const uint4*>(p.skip) + ((size_t)tile * n8) * TILE_M + row : nullptr;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int chunk = (c0 >> 3) + j;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i] = __uint_as_float(r[8 * j + i]) + q.bias[c0 + 8 * j + i];
                    if (p.relu) v[i] = fmaxf(v[i], 0.f);
                }
                if (st) {
                    float sk[8];
                    unpack_bf16x8(__ldg(st + (size_t)chunk * TILE_M), sk);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += sk[i];
                }
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                yt[(size_t)chunk * TILE_M] = o;
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(q.tempty + acc);
        if (p.num_acc == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1u; } else { acc_phase ^= 1u; }
    }
}

// ---- common prologue / epilogue of both GEMM kernels ----------------------------------------------------------
__device__ __forceinline__ uint32_t gemm_setup(const GemmParams& p, const Pipe& q, int warp, int mma_warp, int full_count) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(q.full + i, full_count); mbar_init(q.empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(q.tfull + i, 1); mbar_init(q.tempty + i, 4); }
        mbar_init(q.wbar, 1);
        fence_barrier_init();
    }
    if (warp == mma_warp) tmem_alloc(q.tmem_slot, TMEM_COLS);
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) q.bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *q.tmem_slot;
}

// =================================== pointwise 1x1 ==================================================================
// 6 warps: 0-3 epilogue, 4 TMA producer, 5 MMA issuer (+ TMEM owner)
__global__ void __launch_bounds__(192, 1) pointwise_umma_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Pipe q = carve(smem, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_base = gemm_setup(p, q, warp, 5, 1);
    const SmemPlan s = plan_smem(p.K, p.N, p.stages, p.w_parts);
    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q.wbar, s.b_bytes);
            bulk_g2s(q.b, p.wimg, s.b_bytes, q.wbar);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                mbar_wait(q.empty + stage, phase ^ 1u);
                mbar_arrive_expect_tx(q.full + stage, s.a_bytes);
                bulk_g2s(q.a + (size_t)stage * s.a_bytes, reinterpret_cast<const uint8_t*>(p.x) + (size_t)tile * s.a_bytes,
                         s.a_bytes, q.full + stage);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 5) {
        mma_role(p, q, tmem_base, lane);
    } else {
        epilogue_role(p, q, tmem_base, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =================================== dense 3x3 stride 2 (implicit GEMM) ===========================================
// 9 warps: 0-3 epilogue, 4-7 im2col gather (thread = tile row), 8 MMA issuer (+ TMEM owner)
// K index = (ky*3+kx)*Cin + ci, matching the packed weight rows.  CIN8 = Cin/8 (0: the 3-channel stem, K 27->32).
template <int CIN8>
__global__ void __launch_bounds__(288, 1) dense_umma_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Pipe q = carve(smem, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_base = gemm_setup(p, q, warp, 8, 128);
    const SmemPlan s = plan_smem(p.K, p.N, p.stages, p.w_parts);
    if (warp >= 4 && warp < 8) {
        const int r = threadIdx.x - 128;
        if (r == 0) {
            mbar_arrive_expect_tx(q.wbar, s.b_bytes);
            bulk_g2s(q.b, p.wimg, s.b_bytes, q.wbar);
        }
        const int hw = p.hout * p.hout;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
            mbar_wait(q.empty + stage, phase ^ 1u);
            const int64_t m = (int64_t)tile * TILE_M + r;
            const int64_t n = m / hw;
            const int rem = (int)(m - n * hw);
            const int oy = rem / p.hout, ox = rem - oy * p.hout;
            uint4* dst = reinterpret_cast<uint4*>(q.a + (size_t)stage * s.a_bytes) + r;
            if (CIN8 > 0) {
                const uint4* src = reinterpret_cast<const uint4*>(p.x);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int iy = 2 * oy - 1 + t / 3, ix = 2 * ox - 1 + t % 3;
                    const bool ok = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.hin;
                    const int64_t pin = (n * p.hin + iy) * p.hin + ix;
                    const size_t base = ((size_t)(pin >> 7) * CIN8) * TILE_M + (pin & 127);      // T8 chunk index of channel-chunk 0
#pragma unroll
                    for (int c = 0; c < CIN8; ++c) {
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (ok) v = __ldg(src + base + (size_t)c * TILE_M);
                        dst[(size_t)(t * CIN8 + c) * TILE_M] = v;
                    }
                }
            } else {
                // stem: 3-channel row-major crops [N,64,64,3]; k = tap*3 + c, 27 values zero-padded to 32
                const unsigned short* src = reinterpret_cast<const unsigned short*>(p.x);
                unsigned short vals[32];
#pragma unroll
                for (int i = 27; i < 32; ++i) vals[i] = 0;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int iy = 2 * oy - 1 + t / 3, ix = 2 * ox - 1 + t % 3;
                    const bool ok = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.hin;
                    const int64_t pin = ((n * p.hin + iy) * p.hin + ix) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) vals[t * 3 + c] = ok ? __ldg(src + pin + c) : (unsigned short)0;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 v;
                    v.x = vals[8 * c + 0] | ((uint32_t)vals[8 * c + 1] << 16);
                    v.y = vals[8 * c + 2] | ((uint32_t)vals[8 * c + 3] << 16);
                    v.z = vals[8 * c + 4] | ((uint32_t)vals[8 * c + 5] << 16);
                    v.w = vals[8 * c + 6] | ((uint32_t)vals[8 * c + 7] << 16);
                    dst[(size_t)c * TILE_M] = v;
                }
            }
            fence_proxy_async_smem();             // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(q.full + stage);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 8) {
        mma_role(p, q, tmem_base, lane);
    } else {
        epilogue_role(p, q, tmem_base, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =================================== depthwise (CUDA cores, vectorised) ==============================================
// Thread = one 16-byte output chunk (row r of a tile, 8 channels); thread index == output chunk index, so stores
// are perfectly coalesced and loads are coalesced along rows.  Weights [tap][C] fp32 (BN folded) via the
// read-only path; fp32 accumulate.
template <int K, int S>
__global__ void __launch_bounds__(256)
depthwise_t8_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    bf16* __restrict__ y, int64_t total_chunks, int hin, int hout, int C, int relu) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= total_chunks) return;
    const int c8n = C >> 3;
    const int r = (int)(idx & 127);
    const int64_t tc = idx >> 7;
    const int c = (int)(tc % c8n);
    const int64_t m = (tc / c8n) * TILE_M + r;
    const int hw = hout * hout;
    const int64_t n = m / hw;
    const int rem = (int)(m - n * hw);
    const int oy = rem / hout, ox = rem - oy * hout;
    constexpr int PAD = ((S - 1) + (K - 1)) / 2;
    float acc[8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c * 8) + 1);
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
        acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
    const uint4* src = reinterpret_cast<const uint4*>(x);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S - PAD + ky;
        if (iy < 0 || iy >= hin) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int ix = ox * S - PAD + kx;
            if (ix < 0 || ix >= hin) continue;
            const int64_t pin = (n * hin + iy) * hin + ix;
            float v[8];
            unpack_bf16x8(__ldg(src + ((size_t)(pin >> 7) * c8n + c) * TILE_M + (pin & 127)), v);
            const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(ky * K + kx) * C + c * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
            acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
            acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
            acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
        }
    }
    if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    reinterpret_cast<uint4*>(y)[idx] = o;
}

// Row-tiled version for the shapes of this trunk (square maps of side 8, 4 or 2): thread = one OUTPUT ROW of one crop and one
// 8-channel chunk.  It loads each contributing input row once (HIN consecutive 16-byte chunks: rows of a crop are consecutive T8
// rows) and each filter tap once, and feeds all HOUT outputs of the row from registers: HIN*K + 2*K*K loads for HOUT outputs
// instead of HOUT * 3*K*K.  Same accumulation order per output as the kernel above (bias, then taps in (ky, kx) order): results are
// bit-identical.  Consecutive threads own consecutive output rows of one chunk, so loads and stores stay coalesced.
template <int K, int S, int HIN>
__global__ void __launch_bounds__(128)
depthwise_t8_rows_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         bf16* __restrict__ y, int64_t total_rows, int C, int relu) {
    constexpr int HOUT = HIN / S, PAD = ((S - 1) + (K - 1)) / 2, GROUPS = TILE_M / HOUT;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total_rows) return;
    const int c8n = C >> 3;
    const int g = (int)(t % GROUPS);
    const int64_t tc = t / GROUPS;
    const int c = (int)(tc % c8n);
    const int64_t m_out = (tc / c8n) * TILE_M + (int64_t)g * HOUT;       // first output row-of-tile (pixel) this thread writes
    const int64_t crop = m_out / (HOUT * HOUT);
    const int oy = (int)(m_out - crop * (HOUT * HOUT)) / HOUT;
    float acc[HOUT][8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c * 8) + 1);
#pragma unroll
        for (int ox = 0; ox < HOUT; ++ox) {
            acc[ox][0] = b0.x; acc[ox][1] = b0.y; acc[ox][2] = b0.z; acc[ox][3] = b0.w;
            acc[ox][4] = b1.x; acc[ox][5] = b1.y; acc[ox][6] = b1.z; acc[ox][7] = b1.w;
        }
    }
    const uint4* src = reinterpret_cast<const uint4*>(x);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S - PAD + ky;
        if (iy < 0 || iy >= HIN) continue;
        const int64_t m_in = (crop * HIN + iy) * HIN;                      // HIN consecutive rows inside one tile
        const uint4* row = src + ((size_t)(m_in >> 7) * c8n + c) * TILE_M + (m_in & 127);
        uint4 px[HIN];
#pragma unroll
        for (int i = 0; i < HIN; ++i) px[i] = __ldg(row + i);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(ky * K + kx) * C + c * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
#pragma unroll
            for (int ox = 0; ox < HOUT; ++ox) {
                const int ix = ox * S - PAD + kx;
                if (ix < 0 || ix >= HIN) continue;
                float v[8];
                unpack_bf16x8(px[ix], v);
                acc[ox][0] = fmaf(v[0], w0.x, acc[ox][0]); acc[ox][1] = fmaf(v[1], w0.y, acc[ox][1]);
                acc[ox][2] = fmaf(v[2], w0.z, acc[ox][2]); acc[ox][3] = fmaf(v[3], w0.w, acc[ox][3]);
                acc[ox][4] = fmaf(v[4], w1.x, acc[ox][4]); acc[ox][5] = fmaf(v[5], w1.y, acc[ox][5]);
                acc[ox][6] = fmaf(v[6], w1.z, acc[ox][6]); acc[ox][7] = fmaf(v[7], w1.w, acc[ox][7]);
            }
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(y) + ((size_t)(m_out >> 7) * c8n + c) * TILE_M + (m_out & 127);
#pragma unroll
    for (int ox = 0; ox < HOUT; ++ox) {
        if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[ox][i] = fmaxf(acc[ox][i], 0.f);
        }
        uint4 o;
        o.x = pack_bf16x2(acc[ox][0], acc[ox][1]); o.y = pack_bf16x2(acc[ox][2], acc[ox][3]);
        o.z = pack_bf16x2(acc[ox][4], acc[ox][5]); o.w = pack_bf16x2(acc[ox][6], acc[ox][7]);
        dst[ox] = o;
    }
}

// =================================== weight images =============================================================
// image element (n, k) at ((k/8)*N + n)*8 + k%8; hi image then lo image (w - bf16(w), itself rounded to bf16)
__global__ void prep_weight_kernel(const float* __restrict__ w, bf16* __restrict__ img, int K, int Kpad, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * N) return;
    const int kk = i & 7, n = (i >> 3) % N, chunk = (i >> 3) / N;
    const int k = chunk * 8 + kk;
    const float v = k < K ? w[(size_t)k * N + n] : 0.f;
    const bf16 hi = __float2bfloat16_rn(v);
    img[i] = hi;
    img[(size_t)Kpad * N + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

inline int gemm_k(const cv_layer_info& L) { return L.k * L.k * L.cin; }
inline int gemm_kpad(const cv_layer_info& L) { return (gemm_k(L) + 15) / 16 * 16; }

int fill_params(const cv_layer_info& L, GemmParams* p, int64_t n_crops, bool split_weights) {
    p->K = gemm_kpad(L);
    p->w_parts = split_weights ? 2 : 1;
    p->N = L.cout;
    p->relu = L.relu;
    const int64_t rows = n_crops * L.hout * L.hout;
    if (rows % TILE_M != 0) { cv_set_error("umma: row count %lld is not a multiple of 128", (long long)rows); return CV_ERR_ARG; }
    p->m_tiles = (int)(rows / TILE_M);
    p->n_split = p->N > 256 ? 2 : 1;
    p->n_tile = p->N / p->n_split;
    if (p->n_tile % 16 != 0 || p->N % 16 != 0 || p->N > 512) { cv_set_error("umma: unsupported N=%d", p->N); return CV_ERR_ARG; }
    p->num_acc = p->N <= 256 ? 2 : 1;
    const int a_bytes = TILE_M * p->K * 2, b_bytes = p->N * p->K * 2 * p->w_parts;
    int stages = (SMEM_BUDGET - b_bytes - 4096) / a_bytes;
    p->stages = stages < 2 ? 2 : (stages > 6 ? 6 : stages);
    p->hin = L.hin; p->hout = L.hout; p->cin = L.cin;
    return CV_OK;
}

}  // namespace

size_t umma_weight_image_elems() {
    size_t n = 0;
    const cv_layer_info* L = cv_layers();
    for (int i = 0; i < cv_num_layers(); ++i)
        if (L[i].kind != CV_KIND_DEPTHWISE) n += (size_t)gemm_kpad(L[i]) * L[i].cout;
    return n;
}

int64_t umma_weight_image_offset(int layer) {
    const cv_layer_info* L = cv_layers();
    if (layer < 0 || layer >= cv_num_layers() || L[layer].kind == CV_KIND_DEPTHWISE) return -1;
    int64_t n = 0;
    for (int i = 0; i < layer; ++i)
        if (L[i].kind != CV_KIND_DEPTHWISE) n += 2 * (int64_t)gemm_kpad(L[i]) * L[i].cout;    // hi + lo
    return n;          // every image size is a multiple of 8 elements -> 16-byte aligned
}

int launch_umma_prep_weights(const float* blob, bf16* wimg, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    for (int i = 0; i < cv_num_layers(); ++i) {
        if (L[i].kind == CV_KIND_DEPTHWISE) continue;
        const int K = gemm_k(L[i]), Kp = gemm_kpad(L[i]), N = L[i].cout;
        prep_weight_kernel<<<(Kp * N + 255) / 256, 256, 0, s>>>(blob + L[i].w_offset, wimg + umma_weight_image_offset(i), K, Kp, N);
        CV_CHECK_LAUNCH();
    }
    return CV_OK;
}

int launch_pointwise_umma(const cv_layer_info& L, const bf16* x, const bf16* wimg, const float* bias, const bf16* skip,
                          bf16* y, int64_t n_crops, int num_sms, bool split_weights, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    GemmParams p{};
    int rc = fill_params(L, &p, n_crops, split_weights);
    if (rc) return rc;
    p.x = x; p.wimg = wimg; p.bias = bias; p.skip = skip; p.y = y;
    const SmemPlan sp = plan_smem(p.K, p.N, p.stages, p.w_parts);
    CV_CUDA(cudaFuncSetAttribute(pointwise_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
    pointwise_umma_kernel<<<grid, 192, sp.total, s>>>(p);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_dense_umma(const cv_layer_info& L, const bf16* x, bool in_rowmajor3, const bf16* wimg, const float* bias,
                      bf16* y, int64_t n_crops, int num_sms, bool split_weights, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if (L.k != 3 || L.stride != 2) { cv_set_error("dense_umma: only 3x3 stride 2"); return CV_ERR_ARG; }
    GemmParams p{};
    int rc = fill_params(L, &p, n_crops, split_weights);
    if (rc) return rc;
    p.x = x; p.wimg = wimg; p.bias = bias; p.skip = nullptr; p.y = y;
    const SmemPlan sp = plan_smem(p.K, p.N, p.stages, p.w_parts);
    const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
#define DENSE_LAUNCH(C8)                                                                                                   \
    {                                                                                                                      \
        CV_CUDA(cudaFuncSetAttribute(dense_umma_kernel<C8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));    \
        dense_umma_kernel<C8><<<grid, 288, sp.total, s>>>(p);                                                              \
    }
    if (in_rowmajor3 && L.cin == 3) DENSE_LAUNCH(0)
    else if (!in_rowmajor3 && L.cin == 16) DENSE_LAUNCH(2)
    else if (!in_rowmajor3 && L.cin == 32) DENSE_LAUNCH(4)
    else { cv_set_error("dense_umma: unsupported Cin=%d", L.cin); return CV_ERR_ARG; }
#undef DENSE_LAUNCH
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_depthwise_t8(const cv_layer_info& L, const bf16* x, const float* w, const float* bias, bf16* y, int64_t n_crops,
                        cudaStream_t s) {
    const int64_t total = n_crops * L.hout * L.hout * (L.cout / 8);
    if (total == 0) return CV_OK;
    // the trunk's own shapes: row-tiled kernel (a thread per output row); whole 128-row tiles only (n_crops * hout^2 % 128 == 0)
    if ((n_crops * L.hout * L.hout) % TILE_M == 0 && L.hin == L.hout * L.stride) {
        const int64_t rows = total / L.hout;
        const unsigned rgrid = (unsigned)((rows + 127) / 128);
#define DW_ROWS(KK, SS, HH)                                                                                             \
        if (L.k == KK && L.stride == SS && L.hin == HH) {                                                               \
            depthwise_t8_rows_kernel<KK, SS, HH><<<rgrid, 128, 0, s>>>(x, w, bias, y, rows, L.cout, L.relu);            \
            CV_CHECK_LAUNCH();                                                                                          \
            return CV_OK;                                                                                               \
        }
        DW_ROWS(5, 1, 8) DW_ROWS(5, 2, 8) DW_ROWS(3, 1, 4) DW_ROWS(3, 2, 4) DW_ROWS(5, 1, 2) DW_ROWS(3, 1, 2)
#undef DW_ROWS
    }
    const unsigned grid = (unsigned)((total + 255) / 256);
#define DW_CASE(KK, SS)                                                                                                 \
    if (L.k == KK && L.stride == SS) {                                                                                  \
        depthwise_t8_kernel<KK, SS><<<grid, 256, 0, s>>>(x, w, bias, y, total, L.hin, L.hout, L.cout, L.relu);         \
        CV_CHECK_LAUNCH();                                                                                              \
        return CV_OK;                                                                                                   \
    }
    DW_CASE(3, 1) DW_CASE(3, 2) DW_CASE(5, 1) DW_CASE(5, 2)
#undef DW_CASE
    cv_set_error("depthwise_t8: unsupported k=%d stride=%d", L.k, L.stride);
    return CV_ERR_ARG;
}
