// fp32-grade arithmetic on the tensor cores: the layer-granular trunk with SPLIT fp16 operands (CV_PRECISION_FP32_SPLIT).
//
// The exact mode of the path (logits within 1e-5 of the reference, identical FEN strings: predict.py:24-42 runs the model in fp32) used to
// mean CUDA-core kernels.  Here every activation and every weight is carried as x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22
// significant bits; weights are pre-scaled by a per-layer power of two so that their lo part stays a normal fp16), and every GEMM
// k-step issues three tcgen05 MMAs into fp32 TMEM accumulators:
//
//      D0 += A_hi W_hi        D0 += A_lo W_hi        D1 += A_hi W_lo              (the dropped A_lo W_lo term is 2^-22 relative)
//
// Layout "X2": a tensor with C channels lives in HBM as the T8 layout of 2C channels -- per 128-row tile, C/8 chunks of hi values
// followed by C/8 chunks of lo values -- so a tile is still ONE contiguous block = one TMA bulk copy = the K-major / no-swizzle
// UMMA A operand of both halves.  Same byte count as fp32 activations.
//   pointwise_x2_kernel       the 27 pointwise convs (+bias, ReLU, residual): TMA producer | MMA issuer | two epilogue groups of 4 warps
//   dense_x2_kernel           the stem (crop gather fused in) and blocks.1.0: 3x3 stride-2 convs by software im2col of both halves
//   dense_b00_x2_kernel       blocks.0.0: 3x3 stride-2 conv as an implicit GEMM over a parity-split slab image (no im2col copy)
//   depthwise_x2_smem_kernel  the 15 depthwise convs on the CUDA cores in fp32 (hi + lo summed on load, split on store), tile staged in smem
//   pool_heads_x2_kernel      2x2 mean, type / color heads, type + color combine (square.py:87-104, common.py:24)
// fp16 overflows above 65504: every store of a hi value checks for non-finite lanes and raises the handle's overflow flag
// (cv_square_fp16_status); the caller then re-runs with CV_PRECISION_FP32 (the CUDA-core kernels have fp32 range).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "internal.h"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int TILE_M = 128;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 227 * 1024 - 2048;

struct X2Params {
    const uint16_t* x;        // A source: X2 tiles (pointwise / dense) ...
    const float* x_f32;       // ... or the fp32 NHWC crops [N,64,64,3] (stem)
    const uint16_t* wimg;     // B image [K/8][2N][8] fp16: columns [0,N) = hi(W * 2^s), [N,2N) = lo
    const float* bias;        // [N]
    const uint16_t* skip;     // X2 tiles [M][N] or nullptr
    uint16_t* y;              // X2 tiles [M][N]
    int* ovf;
    float unscale;            // 2^-s
    int m_tiles, K, N, relu, stages, n_split, n_tile, num_acc;
    int slices;               // dense kernels: the K extent of a tile passes through the smem ring in `slices` pieces (one filter row each)
    const uint8_t* boards;    // stem with the crop gather fused in: uint8 HWC boards, normalisation table, board side
    const float* lut;
    int H;
    int raw_ok;               // the board pointer is 4-byte aligned: the patch rows can be staged with word loads
    int groups;               // the large term A_hi W_hi is accumulated in `groups` separate TMEM ranges (k-steps dealt round robin), summed in the epilogue
    int hin, hout, cin;       // dense only
};

// fused crop gather (stem only): normalisation table + tap tables + one crop patch per gather group behind the barriers
constexpr int PATCH_ROWS = 9, PATCH_COLS = 65, PATCH_BYTES = (PATCH_ROWS * PATCH_COLS * 3 * 4 + 15) & ~15, MAX_GG = 4;
constexpr int RAW_ROWS = 16, RAW_WORDS = 80;          // staged board rows under a patch, per gather group (fits 512-pixel boards: 14 rows x 75 words)
constexpr int CROP_SMEM = 768 * 4 + (((int)sizeof(CropTaps) + 15) & ~15) + MAX_GG * PATCH_BYTES + MAX_GG * RAW_ROWS * RAW_WORDS * 4;
struct Plan { uint32_t b_bytes, a_bytes, off_a, off_bias, off_bar, off_crop, total; };
__host__ __device__ inline Plan plan_smem(int K, int N, int stages, int slices = 1, bool crop = false) {
    Plan s;
    s.b_bytes = (uint32_t)K * 2u * N * 2u;
    s.a_bytes = (uint32_t)TILE_M * (K / slices) * 4u;       // one ring stage = one K slice of a tile, hi chunks then lo chunks
    s.off_a = (s.b_bytes + 127u) & ~127u;
    s.off_bias = s.off_a + stages * s.a_bytes;
    s.off_bar = (s.off_bias + N * 4 + 15u) & ~15u;
    s.off_crop = (s.off_bar + (2 * stages + 5) * 8 + 16 + 15u) & ~15u;
    s.total = s.off_crop + (crop ? CROP_SMEM : 0);
    return s;
}
struct Pipe {
    uint8_t *b, *a;
    float* bias;
    uint64_t *full, *empty, *tfull, *tempty, *wbar;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ Pipe carve(uint8_t* smem, const X2Params& p) {
    const Plan s = plan_smem(p.K, p.N, p.stages, p.slices);
    Pipe q;
    q.b = smem;
    q.a = smem + s.off_a;
    q.bias = reinterpret_cast<float*>(smem + s.off_bias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + s.off_bar);
    q.full = bars; q.empty = bars + p.stages; q.tfull = bars + 2 * p.stages; q.tempty = q.tfull + 2; q.wbar = q.tempty + 2;
    q.tmem_slot = reinterpret_cast<uint32_t*>(q.wbar + 1);
    return q;
}

// x -> (hi, lo) fp16 pair of two values each: hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pk2<true>(a, b);
    const float2 h = up2<true>(hi);
    lo = pk2<true>(a - h.x, b - h.y);
}
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&f)[8]) {
    const float2 a = up2<true>(hi.x), b = up2<true>(hi.y), c = up2<true>(hi.z), d = up2<true>(hi.w);
    const float2 e = up2<true>(lo.x), g = up2<true>(lo.y), h = up2<true>(lo.z), i = up2<true>(lo.w);
    f[0] = a.x + e.x; f[1] = a.y + e.y; f[2] = b.x + g.x; f[3] = b.y + g.y;
    f[4] = c.x + h.x; f[5] = c.y + h.y; f[6] = d.x + i.x; f[7] = d.y + i.y;
}
__device__ __forceinline__ uint32_t split8(const float (&v)[8], uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y); split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
    return f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y) | f16x2_nonfinite(hi.z) | f16x2_nonfinite(hi.w);
}

// ---- MMA issuer: work items = (tile, N split); the A stage of a tile serves all of its splits --------------------------------------
__device__ __forceinline__ void mma_role(const X2Params& p, const Pipe& q, uint32_t tmem_base) {
    const uint32_t idesc = make_idesc_f16(TILE_M, p.n_tile);
    const uint32_t a_lbo = TILE_M * 16, b_lbo = (uint32_t)(2 * p.N) * 16;
    const Plan s = plan_smem(p.K, p.N, p.stages);
    mbar_wait(q.wbar, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const int k8 = p.K >> 3;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(q.full + stage, phase);
        for (int nt = 0; nt < p.n_split; ++nt) {
            mbar_wait(q.tempty + acc, acc_phase ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_base = smem_u32(q.a + (size_t)stage * s.a_bytes), b_base = smem_u32(q.b);
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 256), d1 = d0 + (uint32_t)(p.groups * p.n_tile);
                for (int k = 0, g = 0; k < p.K / 16; ++k, g = (g + 1 == p.groups ? 0 : g + 1)) {
                    const uint64_t ah = make_smem_desc(a_base + (2 * k) * a_lbo, a_lbo, 128);
                    const uint64_t al = make_smem_desc(a_base + (k8 + 2 * k) * a_lbo, a_lbo, 128);
                    const uint64_t bh = make_smem_desc(b_base + (2 * k) * b_lbo + (uint32_t)(nt * p.n_tile) * 16, b_lbo, 128);
                    const uint64_t bl = make_smem_desc(b_base + (2 * k) * b_lbo + (uint32_t)(p.N + nt * p.n_tile) * 16, b_lbo, 128);
                    // The tensor core rounds every accumulation toward zero, so the number of MMAs that touch a LARGE accumulator sets the
                    // error of the sum (measured): the large term A_hi W_hi goes to its own accumulators -- `groups` of
                    // them, k-steps dealt round robin, added in the epilogue with round-to-nearest -- and both small terms share D1.
                    mma_bf16_ss(d0 + (uint32_t)(g * p.n_tile), ah, bh, idesc, k >= p.groups ? 1u : 0u);       // (kind::f16: the idesc selects fp16 operands)
                    mma_bf16_ss(d1, al, bh, idesc, k ? 1u : 0u);
                    mma_bf16_ss(d1, ah, bl, idesc, 1u);
                }
                if (nt == p.n_split - 1) mma_commit(q.empty + stage);
                mma_commit(q.tfull + acc);
            }
            __syncwarp();
            if (p.num_acc == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1u; } else { acc_phase ^= 1u; }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
}

// ---- epilogue: groups of 4 warps, warp w of a group owns TMEM lanes [32w, 32w+32) = tile rows.  With two groups (pointwise kernel) group g
//      drains accumulator buffer g, i.e. every other work item: an epilogue is a chain of TMEM loads, global skip loads and stores whose
//      latency one group of four warps cannot hide (pw_proj layers sat at 40 % of the HBM peak), two groups work on two tiles at a time.
//      The skip tile of a work item (N <= 64: the pw_proj layers) is fetched into registers BEFORE the accumulator wait, under the MMAs.
template <bool PW>   // PW: pointwise kernel (two groups, skip prefetch); the dense kernels have one group and no residual
__device__ __forceinline__ void epilogue_role(const X2Params& p, const Pipe& q, uint32_t tmem_base, int warp, int lane, int group, int n_groups) {
    const int row = warp * 32 + lane, n8 = p.N >> 3;
    uint32_t bad = 0, it = 0;
    const bool pre = PW && p.skip != nullptr && p.n_split == 1 && n8 <= 8;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        uint4* yt = reinterpret_cast<uint4*>(p.y) + ((size_t)tile * 2 * n8) * TILE_M + row;
        const uint4* st = PW && p.skip ? reinterpret_cast<const uint4*>(p.skip) + ((size_t)tile * 2 * n8) * TILE_M + row : nullptr;
        for (int nt = 0; nt < p.n_split; ++nt, ++it) {
            const int acc = (int)(it & 1u);
            if (n_groups == 2 && acc != group) continue;
            const uint32_t acc_phase = (it >> 1) & 1u;
            uint4 skh[8], skl[8];
            if (PW && pre) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < n8) { skh[c] = __ldg(st + (size_t)c * TILE_M); skl[c] = __ldg(st + (size_t)(n8 + c) * TILE_M); }
            }
            mbar_wait(q.tfull + acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 256);
            auto block16 = [&](int c0, const uint4* sh, const uint4* sl) {       // 16 accumulator columns from c0; sh / sl: prefetched skip chunks or null
                uint32_t r0[16], r1[16];
                tmem_ld16(taddr + c0, r0);
                tmem_ld16(taddr + p.groups * p.n_tile + c0, r1);
                tmem_ld_wait();
                for (int g = 1; g < p.groups; ++g) {              // large-term partial sums, fixed order
                    uint32_t rg[16];
                    tmem_ld16(taddr + g * p.n_tile + c0, rg);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r0[i] = __float_as_uint(__uint_as_float(r0[i]) + __uint_as_float(rg[i]));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = nt * p.n_tile + c0 + 8 * j, chunk = col >> 3;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] = fmaf(__uint_as_float(r0[8 * j + i]) + __uint_as_float(r1[8 * j + i]), p.unscale, q.bias[col + i]);
                        if (p.relu) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (st) {
                        float sk[8];
                        if (sh) join8(sh[j], sl[j], sk);
                        else join8(__ldg(st + (size_t)chunk * TILE_M), __ldg(st + (size_t)(n8 + chunk) * TILE_M), sk);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += sk[i];
                    }
                    uint4 hi, lo;
                    bad |= split8(v, hi, lo);
                    yt[(size_t)chunk * TILE_M] = hi;
                    yt[(size_t)(n8 + chunk) * TILE_M] = lo;
                }
            };
            if (PW && pre) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (b * 16 < p.n_tile) block16(b * 16, skh + 2 * b, skl + 2 * b);
            } else {
                for (int c0 = 0; c0 < p.n_tile; c0 += 16) block16(c0, nullptr, nullptr);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q.tempty + acc);
        }
    }
    if (bad) atomicOr(p.ovf, 1);
}

__device__ __forceinline__ uint32_t gemm_setup(const X2Params& p, const Pipe& q, int warp, int mma_warp, int full_count) {
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

// 10 warps: 0-3 epilogue group 0, 4 TMA producer, 5 MMA issuer (+ TMEM owner), 6-9 epilogue group 1
__global__ void __launch_bounds__(320, 1) pointwise_x2_kernel(const __grid_constant__ X2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Pipe q = carve(smem, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_base = gemm_setup(p, q, warp, 5, 1);
    const Plan s = plan_smem(p.K, p.N, p.stages);
    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q.wbar, s.b_bytes);
            bulk_g2s(q.b, p.wimg, s.b_bytes, q.wbar);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                mbar_wait(q.empty + stage, phase ^ 1u);
                mbar_arrive_expect_tx(q.full + stage, s.a_bytes);
                bulk_g2s(q.a + (size_t)stage * s.a_bytes, reinterpret_cast<const uint8_t*>(p.x) + (size_t)tile * s.a_bytes, s.a_bytes, q.full + stage);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 5) {
        mma_role(p, q, tmem_base);
    } else {
        epilogue_role<true>(p, q, tmem_base, warp & 3 /* TMEM lane quadrant = warp % 4: warps 6-9 own quadrants 2, 3, 0, 1 */, lane, warp >= 6 ? 1 : 0, 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// MMA issuer of the dense kernels: the ring stages are K SLICES (one filter row = 3 taps each); a tile's accumulators collect its slices.
__device__ __forceinline__ void mma_role_dense(const X2Params& p, const Pipe& q, uint32_t tmem_base) {
    const uint32_t idesc = make_idesc_f16(TILE_M, p.N);
    const uint32_t a_lbo = TILE_M * 16, b_lbo = (uint32_t)(2 * p.N) * 16;
    const Plan s = plan_smem(p.K, p.N, p.stages, p.slices);
    mbar_wait(q.wbar, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const int ksl = (p.K / 16) / p.slices, kc = (p.K >> 3) / p.slices;       // k-steps and hi chunks per slice
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int sl = 0; sl < p.slices; ++sl) {
            mbar_wait(q.full + stage, phase);
            if (sl == 0) mbar_wait(q.tempty + acc, acc_phase ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_base = smem_u32(q.a + (size_t)stage * s.a_bytes), b_base = smem_u32(q.b);
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 256), d1 = d0 + (uint32_t)(p.groups * p.N);
                for (int kk = 0; kk < ksl; ++kk) {
                    const int k = sl * ksl + kk, g = k % p.groups;
                    const uint64_t ah = make_smem_desc(a_base + (2 * kk) * a_lbo, a_lbo, 128);
                    const uint64_t al = make_smem_desc(a_base + (kc + 2 * kk) * a_lbo, a_lbo, 128);
                    const uint64_t bh = make_smem_desc(b_base + (2 * k) * b_lbo, b_lbo, 128);
                    const uint64_t bl = make_smem_desc(b_base + (2 * k) * b_lbo + (uint32_t)p.N * 16, b_lbo, 128);
                    mma_bf16_ss(d0 + (uint32_t)(g * p.N), ah, bh, idesc, k >= p.groups ? 1u : 0u);
                    mma_bf16_ss(d1, al, bh, idesc, k ? 1u : 0u);
                    mma_bf16_ss(d1, ah, bl, idesc, 1u);
                }
                mma_commit(q.empty + stage);
                if (sl == p.slices - 1) mma_commit(q.tfull + acc);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
    }
}

// 13 warps: 0-3 epilogue, 4-7 and 8-11 two im2col gather groups (thread = tile row; the groups take alternate ring stages, so two
// gathers are in flight: the kernel is bound by the latency of the gather, not by bandwidth), 12 MMA issuer (+ TMEM owner).
// K index = (ky*3+kx)*Cin + ci.  CIN8 = Cin/8: one ring stage = filter row ky of one tile = 3 taps x CIN8 hi chunks, then as many lo
// chunks.  CIN8 == 0: the 3-channel stem (K 27 -> 32, one stage per tile), SRC 0 = from the fp32 NHWC crops, SRC 1 = with the crop gather
// fused in: every tap value is the bilinear blend of four uint8 board pixels (ChessSquareCNN._crop_squares, square.py:43-74; the same
// crop_blend() in the same order as crop_kernel, so the values are the reference's bit for bit) -- the 48 KB per crop of fp32 crops never exist.
template <int CIN8, int SRC, int GG /* gather groups of 4 warps */>
__global__ void __launch_bounds__((5 + 4 * GG) * 32, 1) dense_x2_kernel(const __grid_constant__ X2Params p, const __grid_constant__ CropTaps tp_param) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Pipe q = carve(smem, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Plan s = plan_smem(p.K, p.N, p.stages, p.slices, CIN8 == 0 && SRC == 1);
    float* lut_s = reinterpret_cast<float*>(smem + s.off_crop);
    const CropTaps& tp = *reinterpret_cast<const CropTaps*>(smem + s.off_crop + 768 * 4);
    float* patches = reinterpret_cast<float*>(smem + s.off_crop + 768 * 4 + (((int)sizeof(CropTaps) + 15) & ~15));
    if (CIN8 == 0 && SRC == 1) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x) lut_s[i] = p.lut[i];
        for (int i = threadIdx.x; i < (int)(sizeof(CropTaps) / 4); i += blockDim.x)
            reinterpret_cast<uint32_t*>(smem + s.off_crop + 768 * 4)[i] = reinterpret_cast<const uint32_t*>(&tp_param)[i];
        for (int i = threadIdx.x; i < MAX_GG * (PATCH_BYTES / 4); i += blockDim.x) patches[i] = 0.f;      // patch column ix = -1 (conv padding) stays zero
    }
    constexpr int MMA_WARP = 4 + 4 * GG;
    const uint32_t tmem_base = gemm_setup(p, q, warp, MMA_WARP, 128);
    if (warp >= 4 && warp < MMA_WARP) {
        const int grp = (warp - 4) >> 2, r = (threadIdx.x - 128) & 127;
        if (threadIdx.x == 128) {
            mbar_arrive_expect_tx(q.wbar, s.b_bytes);
            bulk_g2s(q.b, p.wimg, s.b_bytes, q.wbar);
        }
        const int kc = CIN8 > 0 ? 3 * CIN8 : 4;
        const int lh = 31 - __clz(p.hout);                   // hout is 32, 16 or 8: shifts, not the 64-bit divisions this loop used to spend 12 % of its instructions on
        // ring position of the current item = (tile iteration, slice); a group fills every GG-th item.  Stem (one item per tile): the group
        // walks ITS tiles only and advances the ring by GG items at a time; the other layers count all items and skip the foreign ones.
        constexpr bool OWN = CIN8 == 0;
        int stage = OWN ? grp % p.stages : 0, turn = 0;
        uint32_t phase = OWN ? (uint32_t)(grp / p.stages) & 1u : 0u;
        auto next_item = [&]() {
            if (OWN) {
                stage += GG;
                while (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
            } else {
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                if (++turn == GG) turn = 0;
            }
        };
        for (int tile = blockIdx.x + (OWN ? grp * (int)gridDim.x : 0); tile < p.m_tiles; tile += (OWN ? GG : 1) * (int)gridDim.x) {
            const int64_t m = (int64_t)tile * TILE_M + r;
            const int64_t n = m >> (2 * lh);
            const int rem = (int)m & ((1 << (2 * lh)) - 1);
            const int oy = rem >> lh, ox = rem & (p.hout - 1);
            for (int sl = 0; sl < p.slices; ++sl, next_item()) {
                if (!OWN && turn != grp) continue;
                mbar_wait(q.empty + stage, phase ^ 1u);
                uint4* dst = reinterpret_cast<uint4*>(q.a + (size_t)stage * s.a_bytes) + r;
                if (CIN8 > 0) {
                    const uint4* src = reinterpret_cast<const uint4*>(p.x);
                    const int iy = 2 * oy - 1 + sl;          // slice = filter row ky
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = 2 * ox - 1 + kx;
                        const bool ok = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.hin;
                        const int64_t pin = (n * p.hin + iy) * p.hin + ix;
                        const size_t base = ((size_t)(pin >> 7) * 2 * CIN8) * TILE_M + (pin & 127);      // X2 tile: hi chunks [0,CIN8), lo [CIN8,2 CIN8)
#pragma unroll
                        for (int c = 0; c < CIN8; ++c) {
                            uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
                            if (ok) { vh = __ldg(src + base + (size_t)c * TILE_M); vl = __ldg(src + base + (size_t)(CIN8 + c) * TILE_M); }
                            dst[(size_t)(kx * CIN8 + c) * TILE_M] = vh;
                            dst[(size_t)(kc + kx * CIN8 + c) * TILE_M] = vl;
                        }
                    }
                } else {
                    float vals[32];
#pragma unroll
                    for (int i = 27; i < 32; ++i) vals[i] = 0.f;
                    if (SRC == 0) {
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const int iy = 2 * oy - 1 + t / 3, ix = 2 * ox - 1 + t % 3;
                            const bool ok = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.hin;
                            const int64_t pin = ((n * p.hin + iy) * p.hin + ix) * 3;
#pragma unroll
                            for (int c = 0; c < 3; ++c) vals[t * 3 + c] = ok ? __ldg(p.x_f32 + pin + c) : 0.f;
                        }
                    } else {
                        // The 128 rows of a tile are 4 output rows x 32 columns of ONE crop: their 3x3 stride-2 windows cover crop rows
                        // [2 oy0 - 1, 2 oy0 + 7] and columns [-1, 63].  The group computes that patch once (every crop pixel = the bilinear
                        // blend of four board pixels, crop_blend() as in crop_kernel: the reference's values bit for bit) and each thread then
                        // picks its 27 taps from shared memory -- 2.6x fewer instructions than blending per tap.
                        float* P = patches + grp * (PATCH_BYTES / 4);
                        const int col = (int)(n & 7), row = (int)((n >> 3) & 7);
                        const uint8_t* board = p.boards + (n >> 6) * (int64_t)p.H * p.H * 3;
                        const int oy0 = oy - (r >> 5);                       // first output row of the tile (r >> 5 = this thread's row inside it)
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");          // the previous tile's readers are done with P and the staged rows
                        // The board rows under the patch (<= RAW_ROWS rows of <= RAW_WORDS aligned words: 9 x 39 for 256-pixel boards) are staged
                        // first -- about three coalesced word loads per thread and ONE exposed global latency per tile, where blending straight
                        // from global memory cost every thread five dependent rounds of twelve byte loads (26 % of the kernel's stall samples).
                        const int iyA = max(2 * oy0 - 1, 0), iyB = min(2 * oy0 + 7, 63);
                        const int yb0 = tp.p0[row][iyA], nrows = tp.p1[row][iyB] - yb0 + 1;
                        const int xw0 = (tp.p0[col][0] * 3) >> 2, nwords = ((tp.p1[col][63] * 3 + 3 + 3) >> 2) - xw0;
                        const bool staged = p.raw_ok && nrows <= RAW_ROWS && nwords <= RAW_WORDS;
                        uint32_t* RW = reinterpret_cast<uint32_t*>(patches + MAX_GG * (PATCH_BYTES / 4)) + grp * (RAW_ROWS * RAW_WORDS);
                        if (staged) {
                            const uint32_t* bw = reinterpret_cast<const uint32_t*>(board);
                            const int hw4 = p.H * 3 / 4;                          // words per board row (H % 32 == 0)
                            for (int i = r; i < nrows * RAW_WORDS; i += 128) {
                                const int rr = i / RAW_WORDS, ww = i - rr * RAW_WORDS;
                                if (ww < nwords) RW[i] = __ldg(bw + (size_t)(yb0 + rr) * hw4 + xw0 + ww);
                            }
                            asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                        }
                        const uint8_t* RB = reinterpret_cast<const uint8_t*>(RW);
                        if (staged) {
                            // thread = crop column ix = r & 63 and every other patch row: the column's taps, weights and byte offsets are
                            // loop-invariant (the zero column ix = -1 of the patch is written once at kernel start)
                            const int ix = r & 63;
                            const int x0 = tp.p0[col][ix] * 3 - xw0 * 4, x1 = tp.p1[col][ix] * 3 - xw0 * 4;
                            const float lx = tp.lam[ix];
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                const int pr = (r >> 6) + 2 * k;
                                if (pr >= PATCH_ROWS) break;
                                const int iy = 2 * oy0 - 1 + pr;
                                float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                                if (iy >= 0 && iy < 64) {
                                    const uint8_t *r0 = RB + (tp.p0[row][iy] - yb0) * (RAW_WORDS * 4), *r1 = RB + (tp.p1[row][iy] - yb0) * (RAW_WORDS * 4);
                                    const float ly = tp.lam[iy];
                                    const uint8_t *p00 = r0 + x0, *p01 = r0 + x1, *p10 = r1 + x0, *p11 = r1 + x1;
                                    v0 = crop_blend(lut_s[p00[0]], lut_s[p01[0]], lut_s[p10[0]], lut_s[p11[0]], lx, ly);
                                    v1 = crop_blend(lut_s[256 + p00[1]], lut_s[256 + p01[1]], lut_s[256 + p10[1]], lut_s[256 + p11[1]], lx, ly);
                                    v2 = crop_blend(lut_s[512 + p00[2]], lut_s[512 + p01[2]], lut_s[512 + p10[2]], lut_s[512 + p11[2]], lx, ly);
                                }
                                float* o = P + (pr * PATCH_COLS + ix + 1) * 3;
                                o[0] = v0; o[1] = v1; o[2] = v2;
                            }
                        } else
                        for (int e = r; e < PATCH_ROWS * PATCH_COLS; e += 128) {
                            const int pr = e / PATCH_COLS, pc = e - pr * PATCH_COLS;
                            const int iy = 2 * oy0 - 1 + pr, ix = pc - 1;
                            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                            if (iy >= 0 && iy < 64 && ix >= 0) {
                                const int y0 = tp.p0[row][iy], y1 = tp.p1[row][iy], x0 = tp.p0[col][ix], x1 = tp.p1[col][ix];
                                const float ly = tp.lam[iy], lx = tp.lam[ix];
                                const uint8_t *p00 = board + (y0 * p.H + x0) * 3, *p01 = board + (y0 * p.H + x1) * 3, *p10 = board + (y1 * p.H + x0) * 3,
                                              *p11 = board + (y1 * p.H + x1) * 3;
                                v0 = crop_blend(lut_s[__ldg(p00)], lut_s[__ldg(p01)], lut_s[__ldg(p10)], lut_s[__ldg(p11)], lx, ly);
                                v1 = crop_blend(lut_s[256 + __ldg(p00 + 1)], lut_s[256 + __ldg(p01 + 1)], lut_s[256 + __ldg(p10 + 1)], lut_s[256 + __ldg(p11 + 1)], lx, ly);
                                v2 = crop_blend(lut_s[512 + __ldg(p00 + 2)], lut_s[512 + __ldg(p01 + 2)], lut_s[512 + __ldg(p10 + 2)], lut_s[512 + __ldg(p11 + 2)], lx, ly);
                            }
                            P[e * 3] = v0; P[e * 3 + 1] = v1; P[e * 3 + 2] = v2;
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");          // patch complete
                        const int oyl = r >> 5;
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const float* src = P + ((2 * oyl + t / 3) * PATCH_COLS + 2 * ox + t % 3) * 3;
                            vals[t * 3] = src[0]; vals[t * 3 + 1] = src[1]; vals[t * 3 + 2] = src[2];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = vals[8 * c + i];
                        uint4 hi, lo;
                        split8(v, hi, lo);                         // normalised pixels: |v| < 3, no overflow possible
                        dst[(size_t)c * TILE_M] = hi;
                        dst[(size_t)(kc + c) * TILE_M] = lo;
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(q.full + stage);
            }
        }
    } else if (warp == MMA_WARP) {
        mma_role_dense(p, q, tmem_base);
    } else {
        epilogue_role<false>(p, q, tmem_base, warp, lane, 0, 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// blocks.0.0 (3x3 stride 2, 32 -> 16 channels, 32x32 -> 16x16) as an IMPLICIT GEMM: no im2col copy.  The gather kernel above rebuilds
// every tile's K = 288 operand image from global memory -- each stem pixel is fetched 2.25 times through L1 / L2 and stored again to
// shared memory -- and sits at half of the HBM roofline, bound by the latency of its gathers.  Here a work item is a COLUMN SLAB of one
// crop's output (16 rows x 8 columns = 128 GEMM rows, as in the 16-bit front end): its input region -- 32 rows x 17 columns x 8 chunk
// planes (4 hi, 4 lo), 70 KB -- is copied ONCE (LDG.128 + STS.128 by two loader groups) into a parity-split image: per plane 33 rows (row 0 = the zero
// padding above the crop) of 9 odd-column units then 8 even-column units.  With that order every filter tap's A operand is a plain
// K-major descriptor into the image: the 8 rows of a core matrix (output columns 8s .. 8s+7, input columns 2 ox - 1 + kx) are 8
// consecutive units of one parity, successive core matrices (output rows) are two image rows apart (SBO), the two K core matrices of a
// k-step are one plane apart (LBO).  Two slab images in flight; the MMA issue order, the split scheme (A_hi W_hi in `groups` accumulators
// dealt round robin, A_lo W_hi + A_hi W_lo in another) and the weight image are those of dense_x2_kernel, so the results are bit-identical.
// 21 warps: 0-3 epilogue, 4-19 loaders (two groups of 8, warp = plane), 20 MMA issuer (+ TMEM owner).
namespace b00 {
constexpr int RPU = 17, ROWS = 33, PLU = RPU * ROWS, PLANES = 8, STAGE_BYTES = PLANES * PLU * 16, NSTAGE = 2;   // units of 16 bytes
constexpr int W_BYTES = 288 * 2 * 16 * 2;
constexpr int OFF_STAGE = (W_BYTES + 127) & ~127, OFF_BIAS = OFF_STAGE + NSTAGE * STAGE_BYTES, OFF_BAR = OFF_BIAS + 64, OFF_PK = OFF_BAR + 96, SMEM = OFF_PK + 17 * 32 * 4;
constexpr int LOADERS = 256;
}
__global__ void __launch_bounds__(672, 1) dense_b00_x2_kernel(const __grid_constant__ X2Params p) {
    using namespace b00;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wimg_s = smem;
    uint8_t* stage_s = smem + OFF_STAGE;
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6, *wbar = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int N = 16, GROUPS = 6, KSTEPS = 18;
    const int n_items = p.m_tiles;                           // 2 slabs per crop = the layer's 128-row tile count
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(full + i, LOADERS); mbar_init(empty + i, 1); mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 20) tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x < N) bias_s[threadIdx.x] = p.bias[threadIdx.x];
    uint32_t* pk_s = reinterpret_cast<uint32_t*>(smem + OFF_PK);
    for (int u = threadIdx.x; u < 17 * 32; u += blockDim.x) {              // loader table, entry u = it * 32 + lane (see the loaders)
        const int iy = u / 17, ixl = u - iy * 17;
        pk_s[u] = (uint32_t)((iy >> 2) * 1024 + (iy & 3) * 32 + ixl) | ((uint32_t)(iy * RPU + ((ixl & 1) ? 9 + (ixl >> 1) : (ixl >> 1))) << 16);
    }
    auto zmask_of = [](int lane) {                                         // bit it: unit it * 32 + lane is column -1 of slab 0 (zero)
        uint32_t z = 0;
#pragma unroll
        for (int it = 0; it < 17; ++it) z |= ((it * 32 + lane) % 17 == 0 ? 1u : 0u) << it;
        return z;
    };
    for (int i = threadIdx.x; i < NSTAGE * STAGE_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(stage_s)[i] = make_uint4(0u, 0u, 0u, 0u);   // row 0 of every plane stays zero
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 20) {
        // ---- loaders: two groups of 8 warps (group g fills image g with every other item, so two items' loads are in flight), warp = plane
        //      (part = plane >> 2, chunk = plane & 3); 17 rounds of 32 units = the plane's 32 rows x 17 columns, all loads of a thread issued
        //      before its first store.  (16-byte cp.async instead of LDG + STS: 3.6 ms per 1024 boards against 3.0 for the gather kernel --
        //      the issuing warps stall on every cp.async.)
        const int grp = (warp - 4) >> 3, plane = (warp - 4) & 7;            // plane = part * 4 + chunk: hi planes 0-3, lo 4-7
        if (threadIdx.x == 128) {
            mbar_arrive_expect_tx(wbar, W_BYTES);
            bulk_g2s(wimg_s, p.wimg, W_BYTES, wbar);
        }
        const uint4* src = reinterpret_cast<const uint4*>(p.x);
        uint4* dst0 = reinterpret_cast<uint4*>(stage_s + (size_t)grp * STAGE_BYTES) + plane * PLU + RPU;      // image row 1 = crop row 0
        // Unit u = it * 32 + lane of the plane: image row iy = u / 17, column ixl = u % 17 (crop column 16 slab - 1 + ixl).  Its global unit
        // offset inside the crop's 8 input tiles and its place in the image do not depend on the item: packed once (16 bits each), so a
        // copy costs an add and a load -- computing them per item took 9 k warp instructions and spread a thread's 17 loads over ~2.5 k cycles.
        //      (the table lives in shared memory: 17 registers more per thread do not fit beside the 68 of the loads in flight at 21 warps)
        const uint32_t* pk = pk_s + lane;
        const uint32_t zmask = zmask_of(lane);
        int i = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
            if ((i & 1) != grp) continue;
            const int sl = item & 1;
            const uint4* base = src + (size_t)(item >> 1) * 8192 + plane * 128 + 16 * sl - 1;
            const uint32_t zm = sl ? 0u : zmask;
            uint4 v[17];
#pragma unroll
            for (int it = 0; it < 17; ++it) {
                v[it] = make_uint4(0u, 0u, 0u, 0u);
                if (!((zm >> it) & 1u)) v[it] = __ldg(base + (pk[it * 32] & 0xFFFFu));
            }
            mbar_wait(empty + grp, ((uint32_t)(i >> 1) & 1u) ^ 1u);                 // the MMAs of item i - 2 have read this image
#pragma unroll
            for (int it = 0; it < 17; ++it) dst0[pk[it * 32] >> 16] = v[it];
            fence_proxy_async_smem();
            mbar_arrive(full + grp);
        }
    } else if (warp == 20) {
        // ---- MMA issuer
        const uint32_t idesc = make_idesc_f16(TILE_M, N);
        constexpr uint32_t b_lbo = 2 * N * 16;
        mbar_wait(wbar, 0);
        int i = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
            const int stage = i & 1, acc = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(full + stage, ph);
            mbar_wait(tempty + acc, ph ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t img = smem_u32(stage_s + (size_t)stage * STAGE_BYTES), b_base = smem_u32(wimg_s);
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 256), d1 = d0 + GROUPS * N;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                    const int tap = k >> 1, kk = k & 1, ky = tap / 3, kx = tap % 3, g = k % GROUPS;
                    const int xoff = kx == 0 ? 0 : kx == 1 ? 9 : 1;
                    const uint32_t a_hi = img + (uint32_t)((2 * kk) * PLU + ky * RPU + xoff) * 16, a_lo = a_hi + 4 * PLU * 16;
                    const uint64_t ah = make_smem_desc(a_hi, PLU * 16, 2 * RPU * 16), al = make_smem_desc(a_lo, PLU * 16, 2 * RPU * 16);
                    const uint64_t bh = make_smem_desc(b_base + (2 * k) * b_lbo, b_lbo, 128);
                    const uint64_t bl = make_smem_desc(b_base + (2 * k) * b_lbo + N * 16, b_lbo, 128);
                    mma_bf16_ss(d0 + (uint32_t)(g * N), ah, bh, idesc, k >= GROUPS ? 1u : 0u);
                    mma_bf16_ss(d1, al, bh, idesc, k ? 1u : 0u);
                    mma_bf16_ss(d1, ah, bl, idesc, 1u);
                }
                mma_commit(empty + stage);
                mma_commit(tfull + acc);
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue: TMEM lane = tile row = output row (lane >> 3) + 4 warp, output column 8 slab + (lane & 7)
        uint32_t bad = 0;
        int i = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
            const int acc = i & 1;
            const int64_t n = item >> 1;
            const int oy = 4 * warp + (lane >> 3), ox = 8 * (item & 1) + (lane & 7);
            const int64_t m = n * 256 + oy * 16 + ox;
            uint4* yt = reinterpret_cast<uint4*>(p.y) + ((size_t)(m >> 7) * 4) * TILE_M + (m & 127);
            mbar_wait(tfull + acc, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 256);
            uint32_t r0[16], r1[16];
            tmem_ld16(taddr, r0);
            tmem_ld16(taddr + GROUPS * N, r1);
            tmem_ld_wait();
#pragma unroll
            for (int g = 1; g < GROUPS; ++g) {                   // large-term partial sums, fixed order
                uint32_t rg[16];
                tmem_ld16(taddr + g * N, rg);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) r0[e] = __float_as_uint(__uint_as_float(r0[e]) + __uint_as_float(rg[e]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[e] = fmaf(__uint_as_float(r0[8 * j + e]) + __uint_as_float(r1[8 * j + e]), p.unscale, bias_s[8 * j + e]);
                    if (p.relu) v[e] = fmaxf(v[e], 0.f);
                }
                uint4 hi, lo;
                bad |= split8(v, hi, lo);
                yt[(size_t)j * TILE_M] = hi;
                yt[(size_t)(2 + j) * TILE_M] = lo;
            }
        }
        if (bad) atomicOr(p.ovf, 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 20) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Depthwise KxK on square maps of side HIN (8, 4 or 2) on X2 tensors: hi + lo are summed on load (fp32), the taps accumulate in fp32 in
// the reference's order (bias, then taps in (ky, kx) order), the result is split on store.  The input tile goes through shared memory (the
// first version, a thread per output row reading its rows straight from global memory -- 128-byte pieces at a 128-byte lane stride, 150-180
// registers, 8 warps per SM -- was latency-bound at 20-45 % of the HBM peak).  A CTA owns one INPUT tile (128 rows = 128 / HIN^2 crops) and a group of `cg` channel chunks: all threads copy the
// 2 x cg chunk planes (hi, lo) with 16-byte cp.async (coalesced: consecutive threads, consecutive rows; here, with few copies per thread and many
// small CTAs per SM, cp.async beats LDG + STS: 0.50 against 0.64 ms for blocks.2.0.dw_mid) into an image whose rows are
// padded by one 16-byte unit (row pitch HIN + 1 units, plane pitch odd: the lanes of a warp -- consecutive image rows -- spread over all
// bank groups, every LDS.128 at its 4-wavefront minimum), then thread = (chunk, crop, output row) runs the same fp32 arithmetic in the
// same order (bias, taps in (ky, kx) order) from shared memory.  Several CTAs per SM overlap each other's copy and compute phases.
template <int K, int S, int HIN>
__global__ void __launch_bounds__(128)
depthwise_x2_smem_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, uint16_t* __restrict__ y,
                         int C, int cg, int relu, int* __restrict__ ovf) {
    constexpr int HOUT = HIN / S, PAD = ((S - 1) + (K - 1)) / 2, ROWS = TILE_M / HIN, RP = HIN + 1, PU = (ROWS * RP) | 1;
    constexpr int TPC = ROWS / S;                          // tasks per chunk: (crop, output row) pairs of the tile
    extern __shared__ __align__(16) uint8_t dsm[];
    uint4* img = reinterpret_cast<uint4*>(dsm);            // [hi | lo][cg][PU] 16-byte units
    const int c8n = C >> 3, n_grp = c8n / cg;
    const int tile = blockIdx.x / n_grp, c0 = (blockIdx.x - tile * n_grp) * cg;
    {
        const uint4* src = reinterpret_cast<const uint4*>(x) + (size_t)tile * 2 * c8n * TILE_M;
        const uint32_t base = smem_u32(img);
        for (int i = threadIdx.x; i < 2 * cg * TILE_M; i += blockDim.x) {
            const int m = i & (TILE_M - 1), pc = i >> 7, part = pc >= cg ? 1 : 0, cl = pc - part * cg;
            const uint4* g = src + ((size_t)(part * c8n + c0 + cl)) * TILE_M + m;
            const uint32_t d = base + (uint32_t)((pc * PU + (m / HIN) * RP + (m % HIN)) * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    uint32_t bad = 0;
    for (int task = threadIdx.x; task < cg * TPC; task += blockDim.x) {
        const int cl = task / TPC, rt = task - cl * TPC, crop_l = rt / HOUT, oy = rt - crop_l * HOUT, c = c0 + cl;
        float acc[HOUT][8];
        {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c * 8) + 1);
#pragma unroll
            for (int ox = 0; ox < HOUT; ++ox) {
                acc[ox][0] = b0.x; acc[ox][1] = b0.y; acc[ox][2] = b0.z; acc[ox][3] = b0.w;
                acc[ox][4] = b1.x; acc[ox][5] = b1.y; acc[ox][6] = b1.z; acc[ox][7] = b1.w;
            }
        }
        const uint4* hi = img + cl * PU + crop_l * HIN * RP;
        const uint4* lo = hi + cg * PU;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int iy = oy * S - PAD + ky;
            if (iy < 0 || iy >= HIN) continue;
            float wk[K][8];                                  // the K taps of this filter row
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(ky * K + kx) * C + c * 8);
                const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
                wk[kx][0] = w0.x; wk[kx][1] = w0.y; wk[kx][2] = w0.z; wk[kx][3] = w0.w;
                wk[kx][4] = w1.x; wk[kx][5] = w1.y; wk[kx][6] = w1.z; wk[kx][7] = w1.w;
            }
            // input pixels left to right: pixel ix feeds output ox through tap kx = ix - ox S + PAD, so every accumulator still
            // receives its taps in increasing kx order
#pragma unroll
            for (int ix = 0; ix < HIN; ++ix) {
                float px[8];
                join8(hi[iy * RP + ix], lo[iy * RP + ix], px);
#pragma unroll
                for (int ox = 0; ox < HOUT; ++ox) {
                    const int kx = ix - ox * S + PAD;
                    if (kx < 0 || kx >= K) continue;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[ox][e] = fmaf(px[e], wk[kx][e], acc[ox][e]);
                }
            }
        }
        const int64_t m_out = ((int64_t)tile * (TILE_M / (HIN * HIN)) + crop_l) * (HOUT * HOUT) + oy * HOUT;
        uint4* dst = reinterpret_cast<uint4*>(y) + ((size_t)(m_out >> 7) * 2 * c8n + c) * TILE_M + (m_out & 127);
#pragma unroll
        for (int ox = 0; ox < HOUT; ++ox) {
            if (relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[ox][e] = fmaxf(acc[ox][e], 0.f);
            }
            uint4 h4, l4;
            bad |= split8(acc[ox], h4, l4);
            dst[ox] = h4;
            dst[(size_t)c8n * TILE_M + ox] = l4;
        }
    }
    if (bad) atomicOr(ovf, 1);
}

// Pool + heads.  Thread = one ROW of the final 2x2 map (lane = crop_local * 4 + pixel: a warp covers 8 crops and its LDG.128 of a chunk plane
// is 512 contiguous bytes; the first version walked single halves with one warp per crop).  Per 8-channel chunk: hi + lo joined, the 2x2
// mean by two xor-shuffles -- (a0 + a1) + (a2 + a3), the reference's order -- after which all four lanes of a crop hold it; lane p stores
// channels 2p, 2p + 1 of the pooled features and accumulates its share of the 7 + 3 head dot products (p = 0: type 0-2, 1: type 3-5,
// 2: type 6 + color 0, 3: color 1-2); type + color -> 13 joint logits at the end (square.py:87-104, common.py:24).
__global__ void __launch_bounds__(256)
pool_heads_x2_kernel(const uint16_t* __restrict__ fmap /* X2 [n_crops*4 rows][480] */, const float* __restrict__ head_w, const float* __restrict__ head_b,
                     int64_t n_crops, float* __restrict__ features, float* __restrict__ squares) {
    const int lane = threadIdx.x & 31, px = lane & 3;
    const int64_t m = (blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + lane;      // row of the X2 tensor
    const int64_t n = m >> 2;
    if ((m & ~(int64_t)31) >= n_crops * 4) return;                    // whole warps only (n_crops * 4 is a multiple of 128)
    const uint4* src = reinterpret_cast<const uint4*>(fmap) + ((size_t)(m >> 7) * 120) * TILE_M + (m & 127);
    const int h0 = px == 0 ? 0 : px == 1 ? 3 : px == 2 ? 6 : 8, nh = px < 2 ? 3 : 2;      // this lane's heads [h0, h0 + nh)
    float part[3] = {0.f, 0.f, 0.f};
#pragma unroll 4
    for (int ch = 0; ch < 60; ++ch) {
        float v[8];
        join8(__ldg(src + (size_t)ch * TILE_M), __ldg(src + (size_t)(60 + ch) * TILE_M), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            v[e] += __shfl_xor_sync(0xffffffffu, v[e], 1);
            v[e] += __shfl_xor_sync(0xffffffffu, v[e], 2);
            v[e] *= 0.25f;
        }
        float2 mine = px == 0 ? make_float2(v[0], v[1]) : px == 1 ? make_float2(v[2], v[3]) : px == 2 ? make_float2(v[4], v[5]) : make_float2(v[6], v[7]);
        *reinterpret_cast<float2*>(features + n * 480 + ch * 8 + px * 2) = mine;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r < nh) {
                const float4* wp = reinterpret_cast<const float4*>(head_w + (h0 + r) * 480 + ch * 8);
                const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
                part[r] = fmaf(v[0], w0.x, part[r]); part[r] = fmaf(v[1], w0.y, part[r]); part[r] = fmaf(v[2], w0.z, part[r]); part[r] = fmaf(v[3], w0.w, part[r]);
                part[r] = fmaf(v[4], w1.x, part[r]); part[r] = fmaf(v[5], w1.y, part[r]); part[r] = fmaf(v[6], w1.z, part[r]); part[r] = fmaf(v[7], w1.w, part[r]);
            }
        }
    }
    // all ten head outputs on every lane of the crop: head h lives on lane base + (h < 3 ? 0 : h < 6 ? 1 : h < 8 ? 2 : 3), slot h - h0
    float head[10];
    const int base = lane & ~3;
#pragma unroll
    for (int h = 0; h < 10; ++h) {
        const int owner = h < 3 ? 0 : h < 6 ? 1 : h < 8 ? 2 : 3, slot = h - (owner == 0 ? 0 : owner == 1 ? 3 : owner == 2 ? 6 : 8);
        head[h] = __shfl_sync(0xffffffffu, part[slot], base + owner) + __ldg(head_b + h);
    }
    const int kT[13] = {0, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6}, kC[13] = {0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2};      // dataset.py:31-32
#pragma unroll
    for (int o = 0; o < 13; ++o)
        if ((o & 3) == px) squares[n * 13 + o] = head[kT[o]] + head[7 + kC[o]];
}

// X2 -> row-major fp32 [rows][C] (debug taps)
__global__ void x2_to_f32_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, size_t n, int C) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t m = i / C;
    const int ch = (int)(i - m * C);
    const __half* f = reinterpret_cast<const __half*>(src);
    const size_t base = (((m >> 7) * (size_t)(2 * (C >> 3)) + (ch >> 3)) * TILE_M + (m & 127)) * 8 + (ch & 7);
    dst[i] = __half2float(f[base]) + __half2float(f[base + (size_t)(C >> 3) * TILE_M * 8]);
}

// ---- weight images: [K/8][2N][8] fp16, columns [0,N) hi(w * 2^s), [N,2N) lo ----------------------------------------------------------
__global__ void prep_x2_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int K, int Kpad, int N, float scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * N) return;
    const int kk = i & 7, n = (i >> 3) % N, chunk = (i >> 3) / N;
    const int k = chunk * 8 + kk;
    const float v = (k < K ? w[(size_t)k * N + n] : 0.f) * scale;
    const __half hi = __float2half_rn(v);
    const size_t row = (size_t)chunk * 2 * N;
    img[(row + n) * 8 + kk] = __half_as_ushort(hi);
    img[(row + N + n) * 8 + kk] = __half_as_ushort(__float2half_rn(v - __half2float(hi)));
}
__global__ void absmax_kernel(const float* __restrict__ w, int n, float* __restrict__ out) {
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
    __shared__ float red[256];
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

inline int gemm_k(const cv_layer_info& L) { return L.k * L.k * L.cin; }
inline int gemm_kpad(const cv_layer_info& L) { return (gemm_k(L) + 15) / 16 * 16; }

int fill_params(const cv_layer_info& L, X2Params* p, int64_t n_crops) {
    p->K = gemm_kpad(L);
    p->N = L.cout;
    p->relu = L.relu;
    const int64_t rows = n_crops * L.hout * L.hout;
    if (rows % TILE_M != 0) { cv_set_error("x2: row count %lld is not a multiple of 128", (long long)rows); return CV_ERR_ARG; }
    p->m_tiles = (int)(rows / TILE_M);
    // two accumulator ranges (D0, D1) of n_tile columns each per work item, 256 TMEM columns per accumulator buffer
    p->n_tile = p->N;
    p->n_split = 1;
    if (p->N > 128) {
        p->n_tile = p->N % 96 == 0 ? 96 : 128;
        if (p->N % p->n_tile != 0) { cv_set_error("x2: unsupported N=%d", p->N); return CV_ERR_ARG; }
        p->n_split = p->N / p->n_tile;
    }
    if (p->n_tile % 16 != 0) { cv_set_error("x2: unsupported N=%d", p->N); return CV_ERR_ARG; }
    p->num_acc = 2;
    p->slices = L.kind == CV_KIND_DENSE && L.cin >= 8 ? 3 : 1;
    {   // (groups + 1) accumulator ranges of n_tile columns inside one 256-column buffer
        int g = 256 / p->n_tile - 1;
        const int steps = p->K / 16;
        g = g > steps ? steps : g;
        p->groups = g < 1 ? 1 : (g > 6 ? 6 : g);
        if (steps <= 2) p->groups = 1;               // stem (K = 32): two MMAs per large-term accumulator is fewer than any other layer gets; its epilogue is issue-bound
    }
    const Plan one = plan_smem(p->K, p->N, 1, p->slices);
    int stages = 1 + (int)((SMEM_LIMIT - (int)one.total) / ((int)one.a_bytes + 16));
    if ((int)one.total > SMEM_LIMIT) { cv_set_error("x2: layer K=%d N=%d does not fit shared memory", p->K, p->N); return CV_ERR_ARG; }
    p->stages = stages < 1 ? 1 : (stages > 6 ? 6 : stages);
    // blocks.0.0: the im2col gather re-reads every input pixel 2.25 times; with three 48 KB stages instead of four the L1 that is left
    // (the shared-memory carve-out follows the request) holds a tile's input region and L2 traffic drops (3.8 -> 3.0 ms per 1024 boards)
    if (p->slices == 3 && one.a_bytes > 40000 && p->stages > 3) p->stages = 3;
    p->hin = L.hin; p->hout = L.hout; p->cin = L.cin;
    return CV_OK;
}

}  // namespace

size_t x2_weight_image_elems() {
    size_t n = 0;
    const cv_layer_info* L = cv_layers();
    for (int i = 0; i < cv_num_layers(); ++i)
        if (L[i].kind != CV_KIND_DEPTHWISE) n += 2 * (size_t)gemm_kpad(L[i]) * L[i].cout;
    return n;
}
int64_t x2_weight_image_offset(int layer) {
    const cv_layer_info* L = cv_layers();
    int64_t n = 0;
    for (int i = 0; i < layer; ++i)
        if (L[i].kind != CV_KIND_DEPTHWISE) n += 2 * (int64_t)gemm_kpad(L[i]) * L[i].cout;
    return n;
}

// Builds every GEMM layer's image; unscale[layer] (host array of cv_num_layers()) receives 2^-s.  Synchronises `s` (reads the maxima).
int launch_x2_prep_weights(const float* blob, uint16_t* wimg, float* unscale_host, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    const int nl = cv_num_layers();
    float* d_max = nullptr;
    CV_CUDA(cudaMalloc(&d_max, nl * sizeof(float)));
    CV_CUDA(cudaMemsetAsync(d_max, 0, nl * sizeof(float), s));
    for (int i = 0; i < nl; ++i)
        if (L[i].kind != CV_KIND_DEPTHWISE) absmax_kernel<<<1, 256, 0, s>>>(blob + L[i].w_offset, gemm_k(L[i]) * L[i].cout, d_max + i);
    std::vector<float> mx(nl, 0.f);
    CV_CUDA(cudaMemcpyAsync(mx.data(), d_max, nl * sizeof(float), cudaMemcpyDeviceToHost, s));
    CV_CUDA(cudaStreamSynchronize(s));
    CV_CUDA(cudaFree(d_max));
    for (int i = 0; i < nl; ++i) {
        unscale_host[i] = 1.f;
        if (L[i].kind == CV_KIND_DEPTHWISE) continue;
        // largest power of two that keeps |w| * 2^s <= 16384: hi stays far from the fp16 limit, lo = w - hi (2^-11 of it) stays normal
        int e = 0;
        if (mx[i] > 0.f && std::isfinite(mx[i])) {
            e = (int)std::floor(std::log2(16384.0 / (double)mx[i]));
            e = e < -14 ? -14 : (e > 24 ? 24 : e);
        }
        const float scale = std::ldexp(1.f, e);
        unscale_host[i] = std::ldexp(1.f, -e);
        const int K = gemm_k(L[i]), Kp = gemm_kpad(L[i]), N = L[i].cout;
        prep_x2_weight_kernel<<<(Kp * N + 255) / 256, 256, 0, s>>>(blob + L[i].w_offset, wimg + x2_weight_image_offset(i), K, Kp, N, scale);
        CV_CHECK_LAUNCH();
    }
    return CV_OK;
}

int launch_pointwise_x2(const cv_layer_info& L, const uint16_t* x, const uint16_t* wimg, const float* bias, float unscale, const uint16_t* skip,
                        uint16_t* y, int64_t n_crops, int num_sms, int* ovf, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    X2Params p{};
    int rc = fill_params(L, &p, n_crops);
    if (rc) return rc;
    p.x = x; p.wimg = wimg; p.bias = bias; p.skip = skip; p.y = y; p.ovf = ovf; p.unscale = unscale;
    const Plan sp = plan_smem(p.K, p.N, p.stages);
    CV_CUDA(cudaFuncSetAttribute(pointwise_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
    pointwise_x2_kernel<<<grid, 320, sp.total, s>>>(p);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

// x_f32_crops non-null: stem from the fp32 crops; boards non-null: stem with the crop gather fused in (uint8 HWC boards, `taps` of the board size)
int launch_dense_x2(const cv_layer_info& L, const uint16_t* x, const float* x_f32_crops, const uint8_t* boards, int H, const float* lut, const CropTaps* taps,
                    const uint16_t* wimg, const float* bias, float unscale, uint16_t* y, int64_t n_crops, int num_sms, int* ovf, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if (L.k != 3 || L.stride != 2) { cv_set_error("dense_x2: only 3x3 stride 2"); return CV_ERR_ARG; }
    X2Params p{};
    int rc = fill_params(L, &p, n_crops);
    if (rc) return rc;
    p.x = x; p.x_f32 = x_f32_crops; p.boards = boards; p.H = H; p.lut = lut;
    p.raw_ok = boards != nullptr && (reinterpret_cast<uintptr_t>(boards) & 3) == 0 && H % 4 == 0;
    p.wimg = wimg; p.bias = bias; p.skip = nullptr; p.y = y; p.ovf = ovf; p.unscale = unscale;
    const Plan sp = plan_smem(p.K, p.N, p.stages, p.slices, L.cin == 3 && boards && taps);
    const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
    static const CropTaps no_taps{};
    const CropTaps& tp = taps ? *taps : no_taps;
#define DENSE_LAUNCH(C8, SRC, GGR)                                                                                              \
    {                                                                                                                           \
        CV_CUDA(cudaFuncSetAttribute(dense_x2_kernel<C8, SRC, GGR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        dense_x2_kernel<C8, SRC, GGR><<<grid, (5 + 4 * GGR) * 32, sp.total, s>>>(p, tp);                                        \
    }
    const bool stem = L.cin == 3;
    if (!stem && x && L.cin == 32 && L.cout == 16 && L.hin == 32 && L.hout == 16 && p.groups == 6) {      // blocks.0.0: implicit GEMM
        CV_CUDA(cudaFuncSetAttribute(dense_b00_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b00::SMEM));
        dense_b00_x2_kernel<<<grid, 672, b00::SMEM, s>>>(p);
        CV_CHECK_LAUNCH();
        return CV_OK;
    }
    if (stem && boards && taps) DENSE_LAUNCH(0, 1, 4)        // measured: 4 gather groups 5.8 ms per 1024 boards, 2 groups 6.7 (blend per tap)
    else if (stem && x_f32_crops) DENSE_LAUNCH(0, 0, 2)
    else if (!stem && x && L.cin == 16) DENSE_LAUNCH(2, 0, 2)
    else { cv_set_error("dense_x2: unsupported Cin=%d / source", L.cin); return CV_ERR_ARG; }
#undef DENSE_LAUNCH
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_depthwise_x2(const cv_layer_info& L, const uint16_t* x, const float* w, const float* bias, uint16_t* y, int64_t n_crops, int* ovf,
                        cudaStream_t s) {
    const int64_t total = n_crops * L.hout * L.hout * (L.cout / 8);
    if (total == 0) return CV_OK;
    if ((n_crops * L.hout * L.hout) % TILE_M != 0 || L.hin != L.hout * L.stride) { cv_set_error("depthwise_x2: crop count / shape not tiled"); return CV_ERR_ARG; }
    const int64_t in_tiles = n_crops * L.hin * L.hin / TILE_M;
    // chunk group: the largest divisor of C / 8 that keeps a CTA at <= 64 tasks (tasks per chunk = 128 / HIN / stride): two-warp CTAs, many
    // per SM, overlap copy and compute best (128 tasks: 0.55 ms for blocks.2.0.dw_mid, 64: 0.50, 32: 0.51)
    const int c8n = L.cout / 8, tpc = TILE_M / L.hin / L.stride;
    int cg = 1;
    for (int d = 1; d <= c8n; ++d)
        if (c8n % d == 0 && d * tpc <= 64) cg = d;
    if (in_tiles * (c8n / cg) >= (int64_t)1 << 31) { cv_set_error("depthwise_x2: wave too large"); return CV_ERR_ARG; }
    const int threads = std::min(128, (cg * tpc + 31) / 32 * 32);
    const unsigned grid = (unsigned)(in_tiles * (c8n / cg));
    const size_t smem = (size_t)2 * cg * (((TILE_M / L.hin) * (L.hin + 1)) | 1) * 16;
#define DW_SMEM(KK, SS, HH)                                                                                                              \
    if (L.k == KK && L.stride == SS && L.hin == HH) {                                                                                    \
        if (smem > 48 * 1024) CV_CUDA(cudaFuncSetAttribute(depthwise_x2_smem_kernel<KK, SS, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        depthwise_x2_smem_kernel<KK, SS, HH><<<grid, threads, smem, s>>>(x, w, bias, y, L.cout, cg, L.relu, ovf);                       \
        CV_CHECK_LAUNCH();                                                                                                               \
        return CV_OK;                                                                                                                    \
    }
    DW_SMEM(5, 1, 8) DW_SMEM(5, 2, 8) DW_SMEM(3, 1, 4) DW_SMEM(3, 2, 4) DW_SMEM(5, 1, 2) DW_SMEM(3, 1, 2)
#undef DW_SMEM
    cv_set_error("depthwise_x2: unsupported k=%d stride=%d hin=%d", L.k, L.stride, L.hin);
    return CV_ERR_ARG;
}

int launch_pool_heads_x2(const uint16_t* fmap, const float* head_w, const float* head_b, int64_t n_crops, float* features, float* squares,
                         cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if ((n_crops * 4) % TILE_M != 0) { cv_set_error("pool_heads_x2: crop count not tiled"); return CV_ERR_ARG; }
    pool_heads_x2_kernel<<<(unsigned)((n_crops * 4 + 255) / 256), 256, 0, s>>>(fmap, head_w, head_b, n_crops, features, squares);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_x2_to_f32(const uint16_t* src, float* dst, size_t n, int C, cudaStream_t s) {
    if (n == 0) return CV_OK;
    x2_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, n, C);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
