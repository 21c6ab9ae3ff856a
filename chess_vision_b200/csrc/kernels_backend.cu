// Fused back-end stages of the ChessSquareCNN trunk for sm_100a (timm mobilenetv4_conv_small_050 forward_features
// as called at models/square.py:86, then global_pool + type/color heads + combine, square.py:87-104).
//
//   stage D ("tail")  blocks.3.0 .. blocks.3.5, blocks.4.0, 2x2 average pool, type/color heads, type+color combine:
//                     21 conv layers + pooling + heads in ONE persistent kernel.  A CTA owns a tile of 32 crops; every
//                     intermediate activation lives in shared memory (bf16 UMMA operand tiles) or tensor memory, only
//                     the 1.5 KB/crop stage input is read from HBM and 1.9 KB/crop of pooled features + 13 logits are
//                     written.  Pointwise convs are tcgen05.mma GEMMs (M = 128 rows = 32 crops x 2x2 pixels, weights
//                     streamed L2 -> smem by TMA bulk copies one op ahead); depthwise convs run on the CUDA cores
//                     in place on the shared-memory tile; the RESIDUAL STREAM of the inverted-bottleneck blocks never
//                     leaves tensor memory: every pw_proj GEMM accumulates (fp32) onto the TMEM columns that hold the
//                     block input, so skip connections cost nothing and are never rounded to bf16.
//
// Execution inside a CTA is op-synchronous (all 16 warps work on the same layer, __syncthreads between ops): the
// per-layer work is dominated by CUDA-core epilogue / depthwise instructions, the MMAs are short.
//
// Row orders.  4x4-resolution tiles (stage input, blocks.3.0 head) are "P8": row = pixel*8 + crop_local, which makes
// every depthwise neighbour access a contiguous, bank-conflict-free 16-byte-per-lane shared-memory read.  2x2-resolution
// tiles are pixel-major as well (row = pixel*32 + crop): the in-place 2x2 depthwise then reads/writes contiguous rows per pixel.
#include <cstdio>
#include <cstdlib>
#include "internal.h"
#include "umma.cuh"
#include "arch_table.inc"     // CV_OFF_* blob offsets (kLayers itself is reached through cv_layers())

namespace {

using namespace umma;

constexpr int NT = 512;                 // 16 warps: warp w -> TMEM lane quadrant w&3, column slice w>>2

// Two fp32 FMAs in one instruction (FFMA2, sm_100+): same FMA-pipe throughput as two FFMAs but ONE issue slot -- the depthwise
// loops are issue-bound (FMAs share the slots with loads, conversions and packs), so this is where their time goes down.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
// Every device function below is templated on the 16-bit operand format of the stage (umma.cuh: F16 = fp16, else bf16).
template <bool F16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const float2 a = up2<F16>(v.x), b = up2<F16>(v.y), c = up2<F16>(v.z), d = up2<F16>(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <bool F16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pk2<F16>(f[0], f[1]), pk2<F16>(f[2], f[3]), pk2<F16>(f[4], f[5]), pk2<F16>(f[6], f[7]));
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// One 128-row GEMM on the tensor core, issued by ONE thread:  D[128 x n] (+)= A[128 x K] * B[n x K]^T.
// A: shared-memory operand tile [K/8][128 rows][8] 16-bit (K-major, no swizzle; LBO 2048 B, SBO 128 B).
// B: weight image [K/8][n_total][8] 16-bit, columns [n0, n0+n).  D: fp32 TMEM columns starting at d_tmem.
// w_parts = 2 (bf16 stages B / C): the image is W_hi followed by W_lo (W = W_hi + W_lo, both bf16): two MMAs per k-step on the
// same A.  fp16 weights carry 11 significant bits and need no second image.
template <bool F16>
__device__ __forceinline__ void issue_gemm(uint32_t a_addr, int K, uint32_t b_addr, int n_total, int n0, int n, uint32_t d_tmem,
                                           bool accumulate, int w_parts = 1) {
    const uint32_t idesc = make_idesc_16<F16>(128, n);
    const uint32_t b_lbo = (uint32_t)n_total * 16u;
    const uint32_t part_rows = ((uint32_t)K * (uint32_t)n_total * 2u) >> 4;           // descriptor words count 16-byte units
    uint32_t a_lo = desc_lo(a_addr, 2048u), b_lo = desc_lo(b_addr + (uint32_t)n0 * 16u, b_lbo);
    constexpr uint32_t hi = desc_hi(128u);
    for (int k = 0; k < K / 16; ++k, a_lo += 4096u >> 4, b_lo += (2u * b_lbo) >> 4)
        for (int part = 0; part < w_parts; ++part)
            mma_f16_ss2(d_tmem, a_lo, hi, b_lo + part * part_rows, hi, idesc, (accumulate || k > 0 || part > 0) ? 1u : 0u);
}

// TMEM accumulator columns [col0, col0+ncols) of this thread's row -> (+bias, ReLU) -> 16-bit -> operand tile
// dst[(chunk0 + col/8)][row][8].  The 16-column groups are dealt round-robin to the 4 column-slice warps.
// fp16 without ReLU (the copies of the fp32 residual stream): `bad` collects non-finite results -- the residual stream lives in
// TMEM in fp32 for the whole stage, so whatever overflowed anywhere in a block ends up in one of these copies.
template <bool F16, bool RELU>
__device__ __forceinline__ void epi_to_tile(uint32_t trow, int col0, int ncols, const float* bias, uint8_t* dst, int chunk0, int row,
                                            int cs, int n_slices, uint32_t& bad) {
    for (int g = cs; g < (ncols >> 4); g += n_slices) {
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)(col0 + g * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float v[8];
            load8(bias + g * 16 + j * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __uint_as_float(r[8 * j + i]);
            const uint4 o = RELU ? make_uint4(pk2r<F16>(v[0], v[1]), pk2r<F16>(v[2], v[3]), pk2r<F16>(v[4], v[5]), pk2r<F16>(v[6], v[7])) : pack8<F16>(v);
            if (F16 && !RELU) bad |= f16x2_nonfinite(o.x) | f16x2_nonfinite(o.y) | f16x2_nonfinite(o.z) | f16x2_nonfinite(o.w);
            *reinterpret_cast<uint4*>(dst + (((size_t)(chunk0 + g * 2 + j) * 128 + row) << 4)) = o;
        }
    }
}

// Depthwise 3x3 stride 1 on the 4x4 maps of ONE 128-row P8 tile, src -> dst (not in place), a task per OUTPUT ROW: (chunk, output row,
// crop, channel half) = 64 tasks per chunk, four times the parallelism of dw3x3_p8_rt for the price of reading every input row up to
// three times.  For the small layers (blocks.3.0.dw_start: 48 channels = 96 whole-map tasks, three warps of sixteen) the latency of
// the load -> convert -> FMA chain is what counts, not the instruction total.  The output row is warp-uniform (no divergence).
template <bool F16, bool RELU>
__device__ __forceinline__ void dw3x3_p8_rows(const uint8_t* src, uint8_t* dst, int C8, const float* w, const float* bias, int tid) {
    const int C = C8 * 8;
    for (int task = tid; task < C8 * 64; task += NT) {
        const int l = task & 15, crop = l >> 1, half = l & 1, oy = (task >> 5) & 3, c = ((task >> 7) << 1) | ((task >> 4) & 1);
        if (c >= C8) continue;
        const uint8_t* sbase = src + (size_t)c * 2048 + crop * 16 + half * 8;
        const float* wp = w + c * 8 + half * 4;
        const float4 b = *reinterpret_cast<const float4*>(bias + c * 8 + half * 4);
        float2 acc[4][2];
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) { acc[ox][0] = lo2(b); acc[ox][1] = hi2(b); }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy - 1 + ky;
            if (iy < 0 || iy > 3) continue;
            uint2 in[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) in[x] = *reinterpret_cast<const uint2*>(sbase + (iy * 4 + x) * 128);
            float4 wt[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) wt[kx] = *reinterpret_cast<const float4*>(wp + (ky * 3 + kx) * C);
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = ox - 1 + kx;
                    if (ix < 0 || ix > 3) continue;
                    acc[ox][0] = fma2(up2<F16>(in[ix].x), lo2(wt[kx]), acc[ox][0]);
                    acc[ox][1] = fma2(up2<F16>(in[ix].y), hi2(wt[kx]), acc[ox][1]);
                }
            }
        }
        uint8_t* dp = dst + (size_t)c * 2048 + crop * 16 + half * 8 + oy * 4 * 128;
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
            uint2 o;
            if (RELU) { o.x = pk2r<F16>(acc[ox][0].x, acc[ox][0].y); o.y = pk2r<F16>(acc[ox][1].x, acc[ox][1].y); }
            else { o.x = pk2<F16>(acc[ox][0].x, acc[ox][0].y); o.y = pk2<F16>(acc[ox][1].x, acc[ox][1].y); }
            *reinterpret_cast<uint2*>(dp + ox * 128) = o;
        }
    }
}

// Depthwise 3x3 stride 1 on 4x4 maps, P8 rows, IN PLACE over consecutive 128-row tiles of C8 chunks, register tiled.
// Task = (tile*C8 + chunk, crop, channel half): the thread loads the whole 4x4 map of 4 channels of one crop (16 x 8 bytes,
// a half warp reads 128 contiguous bytes per pixel), converts it once, computes all 16 outputs from registers (static border
// handling: no branches, every load issued up front) and writes them back over its own inputs -- no other thread touches those
// bytes, so no synchronisation is needed.  Against one task per output: 1.9x fewer instructions, 3x fewer LSU wavefronts.
template <bool F16, bool RELU>
__device__ __forceinline__ void dw3x3_p8_rt(const uint8_t* src, uint8_t* buf, int n_tasks, int C8, const float* w, const float* bias, int tid,
                                            int nt = NT) {
    const int C = C8 * 8;
    for (int task = tid; task < n_tasks; task += nt) {
        const int l = task & 15, crop = l >> 1, half = l & 1, tc = task >> 4, c = tc % C8;
        const uint8_t* sbase = src + (size_t)tc * 2048 + crop * 16 + half * 8;
        uint8_t* base = buf + (size_t)tc * 2048 + crop * 16 + half * 8;
        uint2 in[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) in[p] = *reinterpret_cast<const uint2*>(sbase + p * 128);
        const float* wp = w + c * 8 + half * 4;
        float4 wt[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const float4*>(wp + t * C);
        const float4 b = *reinterpret_cast<const float4*>(bias + c * 8 + half * 4);
        float2 x[16][2];
#pragma unroll
        for (int p = 0; p < 16; ++p) { x[p][0] = up2<F16>(in[p].x); x[p][1] = up2<F16>(in[p].y); }
#pragma unroll
        for (int oy = 0; oy < 4; ++oy) {
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) {
                float2 a01 = lo2(b), a23 = hi2(b);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = oy - 1 + ky;
                    if (iy < 0 || iy > 3) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = ox - 1 + kx;
                        if (ix < 0 || ix > 3) continue;
                        a01 = fma2(x[iy * 4 + ix][0], lo2(wt[ky * 3 + kx]), a01);
                        a23 = fma2(x[iy * 4 + ix][1], hi2(wt[ky * 3 + kx]), a23);
                    }
                }
                uint2 o;
                if (RELU) { o.x = pk2r<F16>(a01.x, a01.y); o.y = pk2r<F16>(a23.x, a23.y); }
                else { o.x = pk2<F16>(a01.x, a01.y); o.y = pk2<F16>(a23.x, a23.y); }
                *reinterpret_cast<uint2*>(base + (oy * 4 + ox) * 128) = o;
            }
        }
    }
}

// Depthwise 3x3 stride 2 (+ReLU) 4x4 -> 2x2: src [C8][128 P8 rows][8] (8 crops) -> dst rows p*32 + crop0 + crop (pixel-major
// 2x2 tile of 32 crops), chunks chunk0 + c of a [.][128][8] tile.
// Register tiled: task = (chunk, crop, channel half) loads the crop's whole 4x4 map of 4 channels once (a half
// warp reads 128 contiguous bytes per pixel), converts it once and computes the four stride-2 outputs with static border handling.
template <bool F16>
__device__ __forceinline__ void dw3x3s2_p8_rt(const uint8_t* src, uint8_t* dst, int C8, int C_total, int chunk0, int crop0, const float* w,
                                              const float* bias, int tid) {
    for (int task = tid; task < 16 * C8; task += NT) {
        const int l = task & 15, crop = l >> 1, half = l & 1, c = task >> 4, cg = chunk0 + c;
        const uint8_t* sbase = src + (size_t)c * 2048 + crop * 16 + half * 8;
        uint2 in[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) in[p] = *reinterpret_cast<const uint2*>(sbase + p * 128);
        const float* wp = w + cg * 8 + half * 4;
        float4 wt[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const float4*>(wp + t * C_total);
        const float4 b = *reinterpret_cast<const float4*>(bias + cg * 8 + half * 4);
        float2 x[16][2];
#pragma unroll
        for (int p = 0; p < 16; ++p) { x[p][0] = up2<F16>(in[p].x); x[p][1] = up2<F16>(in[p].y); }
        uint8_t* dbase = dst + (((size_t)cg * 128 + crop0 + crop) << 4) + half * 8;
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
                float2 a01 = lo2(b), a23 = hi2(b);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = 2 * oy - 1 + ky;
                    if (iy < 0 || iy > 3) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = 2 * ox - 1 + kx;
                        if (ix < 0 || ix > 3) continue;
                        a01 = fma2(x[iy * 4 + ix][0], lo2(wt[ky * 3 + kx]), a01);
                        a23 = fma2(x[iy * 4 + ix][1], hi2(wt[ky * 3 + kx]), a23);
                    }
                }
                *reinterpret_cast<uint2*>(dbase + (oy * 2 + ox) * 512) = make_uint2(pk2r<F16>(a01.x, a01.y), pk2r<F16>(a23.x, a23.y));
            }
        }
    }
}

// Depthwise KxK stride 1 on 2x2 maps, IN PLACE on a pixel-major tile [C8][128][8] (row = pixel*32 + crop, 32 crops).
// Thread = (crop, 8-channel chunk): it reads the crop's 4 pixels, computes all 4 outputs (each output sees all 4 inputs
// on a 2x2 map) and writes them back -- no other thread touches these 64 bytes, so no synchronisation is needed; a warp
// is 32 crops of one chunk, so activation accesses are contiguous and the weight reads are warp-uniform broadcasts.
// Weights: wq[chunk][p][q][8] fp32 = tap (q - p) of the KxK filter for output pixel p / input pixel q (prep_dw2x2_kernel).
template <bool F16>
__device__ __forceinline__ void dw2x2_pm(uint8_t* buf, int C8, const float* wq, const float* bias, bool relu, int tid) {
    for (int task = tid; task < 32 * C8; task += NT) {
        const int c = task >> 5, crop = task & 31;
        uint8_t* base = buf + (((size_t)c * 128 + crop) << 4);
        float2 x[4][4];                          // packed fp32 pairs: the 128 multiply-adds of a task are 64 FFMA2
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + q * 512);
            x[q][0] = up2<F16>(v.x); x[q][1] = up2<F16>(v.y); x[q][2] = up2<F16>(v.z); x[q][3] = up2<F16>(v.w);
        }
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c * 8), b1 = *reinterpret_cast<const float4*>(bias + c * 8 + 4);
        const float* wc = wq + c * 128;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            float2 acc[4] = {lo2(b0), hi2(b0), lo2(b1), hi2(b1)};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w0 = *reinterpret_cast<const float4*>(wc + (pp * 4 + q) * 8), w1 = *reinterpret_cast<const float4*>(wc + (pp * 4 + q) * 8 + 4);
                acc[0] = fma2(x[q][0], lo2(w0), acc[0]);
                acc[1] = fma2(x[q][1], hi2(w0), acc[1]);
                acc[2] = fma2(x[q][2], lo2(w1), acc[2]);
                acc[3] = fma2(x[q][3], hi2(w1), acc[3]);
            }
            uint4 o;                             // ReLU rides on the conversion (cvt.rn.relu), same values as max(acc, 0) then round
            if (relu) o = make_uint4(pk2r<F16>(acc[0].x, acc[0].y), pk2r<F16>(acc[1].x, acc[1].y), pk2r<F16>(acc[2].x, acc[2].y), pk2r<F16>(acc[3].x, acc[3].y));
            else o = make_uint4(pk2<F16>(acc[0].x, acc[0].y), pk2<F16>(acc[1].x, acc[1].y), pk2<F16>(acc[2].x, acc[2].y), pk2<F16>(acc[3].x, acc[3].y));
            *reinterpret_cast<uint4*>(base + pp * 512) = o;
        }
    }
}

// =====================================================================================================================
// stage D
// =====================================================================================================================
namespace sd {
constexpr int NOPS = 24;
// op:      0 dw24  1 pw25  2 dw26 | 3 pw27 | 4 dw28 5 pw29 6 dw30 7 pw31 | 8 pw32 9 dw33 10 pw34 | 11 pw35 12 dw36 13 pw37 |
//          14 pw38 15 dw39 16 pw40 | 17 pw41 18 dw42 19 pw43 | 20,21,22 pw44 (three 160-column parts) | 23 heads
constexpr int OFF_IN = 0;                       // 2 x 12288: stage-input sub-tile ring (128 P8 rows x 48 ch)
constexpr int OFF_A24 = 24576;                  // 12288: blocks.3.0.dw_start output (A operand of pw_exp)
constexpr int OFF_EH = 36864;                   // 36864: half (144 ch) of the blocks.3.0.pw_exp output | body: X16 (128 x 64) + head partials
constexpr int OFF_BIG = 73728;                  // 73728: blocks.3.0.dw_mid output (128 x 288) | body: expanded activation (128 x <=256)
constexpr int OFF_W = 147456;                   // weight arena: slot 0 at +0 (also the 3 resident head-phase blobs), slot 1 at +W_SLOT1
constexpr int W_SLOT1 = 42496;
constexpr int W_ARENA = W_SLOT1 + 37376;        // 79872
constexpr int OFF_BAR = OFF_W + W_ARENA;        // 227328
constexpr int SMEM = OFF_BAR + 64;
static_assert(SMEM <= 227 * 1024, "stage D shared memory budget");
constexpr int H_OFF1 = 1920, H_OFF2 = 30720;    // head-phase blobs inside slot 0: dw24 @0, pw25 @1920, dw26 @30720 (ends 42240)
constexpr int S_COL = 448;                      // TMEM columns [448,512): fp32 residual stream (64 ch); [0,288): pw_exp / blocks.4.0 accumulators
}  // namespace sd

struct StageDParams {
    const bf16* x;            // stage input: T8 tiles of 128 P8 rows x 48 ch (one tile = 8 crops)
    const uint8_t* wimg;      // stage weight image (build_stageD_image), CV_W_REPLICAS copies w_stride bytes apart
    int w_rep;                // CTA b streams its weights from copy b % w_rep (see CV_W_REPLICAS)
    uint32_t w_stride;
    float* features;          // pooled trunk features (square.py:90): [n_crops][480], or the FT layout of kernels_head.cu (tiled)
    float* squares;           // [n_crops][13]  combined type+color logits (common.py:24)
    int n_tiles;              // n_crops / 32
    int tiled;
    long long crop_base;      // tiled: index of this launch's first crop inside the chunk that `features` (FT base) covers
    int perm_boards;          // > 0: the input tiles are in the permuted crop order of a launch of this many boards (perm_pos)
    uint32_t off[sd::NOPS], bytes[sd::NOPS];
    int debug;                // CV_SD_DEBUG ablation bits (timing experiments only: results are wrong)
    const int* gate;          // non-null: the kernel runs only when (*gate != 0) == gate_want (fp16 pass: want 0; bf16 fall-back pass: want 1)
    int gate_want;
    int* ovf;                 // fp16 pass: set to 1 when a residual-stream value left the fp16 range (the bf16 pass then redoes the wave)
};

__constant__ int kTypeOf[13] = {0, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6};     // dataset.py:31
__constant__ int kColorOf[13] = {0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2};    // dataset.py:32

// Timing experiments (-DCV_SC_PROFILE, CV_SD_DEBUG & 256): cycles thread 0 of CTA 0 spends per section of a stage D tile (same caveats as SC_MARK).
#ifdef CV_SC_PROFILE
__device__ unsigned long long g_sd_prof[40];
#define SD_MARK(k) do { if (prof_on) { const long long _n = clock64(); pacc[k] += _n - plast; plast = _n; } } while (0)
#else
#define SD_MARK(k)
#endif

template <bool F16>
__global__ void __launch_bounds__(NT, 1) stageD_kernel(const __grid_constant__ StageDParams p) {
    using namespace sd;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_want != 0)) return;      // uniform over the grid up to races that only skip redundant work
    uint32_t bad = 0;
    const uint8_t* wsrc = p.wimg + (size_t)(blockIdx.x % p.w_rep) * p.w_stride;
    uint8_t* IN = smem + OFF_IN;
    uint8_t* A24 = smem + OFF_A24;
    uint8_t* EH = smem + OFF_EH;
    uint8_t* BIG = smem + OFF_BIG;
    uint8_t* WA = smem + OFF_W;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* wbar = bars;           // [2] weight slots
    uint64_t* inbar = bars + 2;      // [2] input ring
    uint64_t* mbar = bars + 4;       // MMA completion
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, cs = warp >> 2, row = quad * 32 + lane;

    if (tid == 0) {
        mbar_init(wbar, 1); mbar_init(wbar + 1, 1); mbar_init(inbar, 1); mbar_init(inbar + 1, 1); mbar_init(mbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t wph0 = 0, wph1 = 0, inph0 = 0, inph1 = 0, mph = 0;
#ifdef CV_SC_PROFILE
    const bool prof_on = (p.debug & 256) && blockIdx.x == 0 && tid == 0;
    long long pacc[40];
    for (int i = 0; i < 40; ++i) pacc[i] = 0;
    long long plast = clock64();
#endif

    auto load_head_weights = [&]() {            // thread 0: the three blobs that stay resident during the 4x4 phase
        mbar_arrive_expect_tx(wbar, p.bytes[0] + p.bytes[1] + p.bytes[2]);
        bulk_g2s(WA, wsrc + p.off[0], p.bytes[0], wbar);
        bulk_g2s(WA + H_OFF1, wsrc + p.off[1], p.bytes[1], wbar);
        bulk_g2s(WA + H_OFF2, wsrc + p.off[2], p.bytes[2], wbar);
    };
    auto load_in = [&](int tile, int t) {       // thread 0: sub-tile t (8 crops) of `tile` into ring slot t&1
        uint64_t* b = inbar + (t & 1);
        mbar_arrive_expect_tx(b, 12288);
        bulk_g2s(IN + (t & 1) * 12288, reinterpret_cast<const uint8_t*>(p.x) + ((size_t)tile * 4 + t) * 12288, 12288, b);
    };
    auto wait_in = [&](int s) {
        if (s) { mbar_wait(inbar + 1, inph1); inph1 ^= 1u; } else { mbar_wait(inbar, inph0); inph0 ^= 1u; }
    };
    // body op `op` (3..23) lives in slot op&1; entering it prefetches op+1 into the other slot (free: op-1 is complete)
    auto prefetch = [&](int nx) {                // any one thread: op nx -> slot nx & 1
        uint64_t* b = wbar + (nx & 1);
        mbar_arrive_expect_tx(b, p.bytes[nx]);
        bulk_g2s(WA + ((nx & 1) ? W_SLOT1 : 0), wsrc + p.off[nx], p.bytes[nx], b);
    };
    // pf false: a GEMM op issues the prefetch after the CTA barrier that precedes its MMA issue, from a warp that is NOT the MMA issuer
    // (the bulk copy's issue latency, a few hundred cycles, then runs while that warp would wait for the tensor pipe anyway; on the
    // issuing warp it delayed the op's epilogue, and every closing barrier waits for the slowest warp)
    constexpr int PF_WARP = 15;
    auto begin_op = [&](int op, bool pf = true) -> uint8_t* {
        if (pf && warp == PF_WARP && op + 1 < NOPS && elect_one()) prefetch(op + 1);
#ifdef CV_SC_PROFILE
        if (prof_on) {                           // how often, and for how long, the weights of an op are NOT there when it begins
            const long long t = clock64();
            if (!mbar_try_wait(wbar + (op & 1), (op & 1) ? wph1 : wph0)) { ++pacc[33]; mbar_wait(wbar + (op & 1), (op & 1) ? wph1 : wph0); }
            pacc[32] += clock64() - t; ++pacc[34];
            if (op & 1) wph1 ^= 1u; else wph0 ^= 1u;
            return WA + ((op & 1) ? W_SLOT1 : 0);
        }
#endif
        if (op & 1) { mbar_wait(wbar + 1, wph1); wph1 ^= 1u; } else { mbar_wait(wbar, wph0); wph0 ^= 1u; }
        return WA + ((op & 1) ? W_SLOT1 : 0);
    };
    // generic-proxy smem writes + TMEM reads of all threads ordered before the MMAs the elected thread issues next
    auto sync_before_mma = [&]() {
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    auto wait_mma = [&]() {
        mbar_wait(mbar, mph);
        mph ^= 1u;
        tc_fence_after();
    };

    if (tid == 0 && blockIdx.x < p.n_tiles) {
        load_head_weights();
        load_in(blockIdx.x, 0);
        load_in(blockIdx.x, 1);
    }

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // ------------------------------ blocks.3.0 at 4x4: four sub-tiles of 8 crops -----------------------------------
        SD_MARK(39);
        mbar_wait(wbar, wph0); wph0 ^= 1u;
        const float* b24 = reinterpret_cast<const float*>(WA);
        const float* w24 = b24 + 48;
        const float* b25 = reinterpret_cast<const float*>(WA + H_OFF1);
        const uint32_t w25 = smem_u32(WA + H_OFF1 + 288 * 4);
        const float* b26 = reinterpret_cast<const float*>(WA + H_OFF2);
        const float* w26 = b26 + 288;
        const bool loader = warp == PF_WARP && elect_one();      // the depthwise of this phase has 96 tasks: the last warp is idle in it
        if (loader) {                            // blocks.3.0.pw_proj weights -> slot 1 while the 4x4 phase runs
            mbar_arrive_expect_tx(wbar + 1, p.bytes[3]);
            bulk_g2s(WA + W_SLOT1, wsrc + p.off[3], p.bytes[3], wbar + 1);
        }
        for (int t = 0; t < 4; ++t) {
            wait_in(t & 1);
            SD_MARK(0);
            const uint8_t* in = IN + (t & 1) * 12288;
            if (!(p.debug & 1)) dw3x3_p8_rows<F16, false>(in, A24, 6, w24, b24, tid);          // L24 dw_start (no act)
            SD_MARK(1);
            sync_before_mma();
            if (loader) {                        // ring slot t & 1 is consumed: sub-tile t + 2 (of this tile, or the next tile's first two) -> it
                const int tn = t < 2 ? tile : tile + (int)gridDim.x;
                if (tn < p.n_tiles) load_in(tn, (t + 2) & 3);
            }
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                issue_gemm<F16>(smem_u32(A24), 48, w25, 288, 0, 144, tmem, false);             // L25 pw_exp 48 -> 288
                issue_gemm<F16>(smem_u32(A24), 48, w25, 288, 144, 144, tmem + 144, false);
                mma_commit(mbar);
            }
            wait_mma();
            SD_MARK(2);
            for (int h = 0; h < 2; ++h) {
                epi_to_tile<F16, true>(trow, 144 * h, 144, b25 + 144 * h, EH, 0, row, cs, 4, bad);
                SD_MARK(3);
                __syncthreads();
                if (!(p.debug & 2)) dw3x3s2_p8_rt<F16>(EH, BIG, 18, 288, 18 * h, 8 * t, w26, b26, tid);   // L26 dw_mid s2 (+ReLU) -> rows of the 2x2 tile
                SD_MARK(4);
                __syncthreads();
                SD_MARK(5);
            }
        }
        // ------------------------------ 2x2 phase: 128 rows = 32 crops ------------------------------------------------------
        uint8_t* X16 = EH;                       // block input as bf16 operand tile [8][128][8]
        int op = 3;
        {   // L27 blocks.3.0.pw_proj 288 -> 64: starts the residual stream in TMEM
            uint8_t* wb = begin_op(op, false);
            sync_before_mma();
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                issue_gemm<F16>(smem_u32(BIG), 288, smem_u32(wb + 256), 64, 0, 64, tmem + S_COL, false);
                mma_commit(mbar);
            }
            if (warp == PF_WARP && op + 1 < NOPS && elect_one()) prefetch(op + 1);
            wait_mma();
            epi_to_tile<F16, false>(trow, S_COL, 64, reinterpret_cast<const float*>(wb), X16, 0, row, cs, 4, bad);
            __syncthreads();
            SD_MARK(6 + op); ++op;
        }
#pragma unroll 1
        for (int blk = 1; blk <= 5; ++blk) {
            const int cexp = blk == 3 ? 192 : 256;
            if (blk == 1) {                      // L28 blocks.3.1.dw_start 5x5 on the block input (no act)
                uint8_t* wb = begin_op(op);
                const float* b = reinterpret_cast<const float*>(wb);
                if (!(p.debug & 4)) dw2x2_pm<F16>(X16, 8, b + 64, b, false, tid);
                SD_MARK(6 + op); ++op;           // no closing barrier: pw_exp opens with sync_before_mma()
            }
            {   // pw_exp 64 -> cexp (+ReLU)
                uint8_t* wb = begin_op(op, false);
                sync_before_mma();
                if (warp == 0 && elect_one()) {
                    tc_fence_after();
                    issue_gemm<F16>(smem_u32(X16), 64, smem_u32(wb + cexp * 4), cexp, 0, cexp, tmem, false);
                    mma_commit(mbar);
                }
                if (warp == PF_WARP && op + 1 < NOPS && elect_one()) prefetch(op + 1);
                wait_mma();
                epi_to_tile<F16, true>(trow, 0, cexp, reinterpret_cast<const float*>(wb), BIG, 0, row, cs, 4, bad);
                __syncthreads();
                SD_MARK(6 + op); ++op;
            }
            {   // dw_mid 5x5 / 3x3 (+ReLU), in place
                uint8_t* wb = begin_op(op);
                const float* b = reinterpret_cast<const float*>(wb);
                if (!(p.debug & 4)) dw2x2_pm<F16>(BIG, cexp >> 3, b + cexp, b, true, tid);
                SD_MARK(6 + op); ++op;           // no closing barrier: pw_proj opens with sync_before_mma(), and prefetches behind it
            }
            {   // pw_proj cexp -> 64, accumulated onto the residual stream in TMEM (skip connection)
                uint8_t* wb = begin_op(op, false);
                sync_before_mma();
                if (warp == 0 && elect_one()) {
                    tc_fence_after();
                    issue_gemm<F16>(smem_u32(BIG), cexp, smem_u32(wb + 256), 64, 0, 64, tmem + S_COL, true);
                    mma_commit(mbar);
                }
                if (warp == PF_WARP && op + 1 < NOPS && elect_one()) prefetch(op + 1);
                wait_mma();
                epi_to_tile<F16, false>(trow, S_COL, 64, reinterpret_cast<const float*>(wb), X16, 0, row, cs, 4, bad);   // + cumulative bias
                SD_MARK(6 + op); ++op;           // no closing barrier: the next op (pw_exp / blocks.4.0) opens with sync_before_mma()
            }
        }
        // ------------------------------ blocks.4.0 (64 -> 480, ReLU) + average pool + heads -------------------------------------
        for (int part = 0; part < 3; ++part, ++op) {          // three 160-column weight parts (ops 20..22)
            uint8_t* wb = begin_op(op, false);
            sync_before_mma();
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                issue_gemm<F16>(smem_u32(X16), 64, smem_u32(wb), 160, 0, 160, tmem + 160 * part, false);
                mma_commit(mbar);
            }
            if (warp == PF_WARP && op + 1 < NOPS && elect_one()) prefetch(op + 1);
            wait_mma();
            SD_MARK(6 + op);
        }
        // Weight slot 0 (last used by op 22, whose MMAs have completed) is free from here on: the next tile's resident weights arrive
        // under the pool + heads epilogue instead of after it (its first two input sub-tiles were requested in the 4x4 phase).
        if (warp == PF_WARP && tile + (int)gridDim.x < p.n_tiles && elect_one()) load_head_weights();
        {
            uint8_t* wb = begin_op(op);          // op 23 (slot 1): [head_b 16][head_w 10x480][blocks.4.0 bias 480] fp32
            const float* hb = reinterpret_cast<const float*>(wb);
            const float* hw = hb + 16;
            const float* bias44 = hw + 4800;
            // TMEM lane = pixel*32 + crop: warp (quad = pixel, cs) holds one pixel of 32 crops.  The 2x2 mean needs the 4
            // pixel warps of a column slice to exchange through shared memory (named barrier per slice, 128 threads).
            float part_acc[10];
#pragma unroll
            for (int r = 0; r < 10; ++r) part_acc[r] = 0.f;
            float* xbuf = reinterpret_cast<float*>(BIG) + (size_t)cs * 4 * 16 * 32;      // [pixel][16 ch][32 crops] of this slice
            int64_t crop = (int64_t)tile * 32 + lane;                                    // position in the launch -> crop index b * 64 + square
            if (p.perm_boards > 0) {
                int64_t pb;
                int psq;
                perm_inv(crop, p.perm_boards, pb, psq);
                crop = pb * 64 + psq;
            }
            for (int g = cs; g < ((p.debug & 8) ? 0 : 30); g += 4) {
                uint32_t rr[16];
                tmem_ld16(trow + (uint32_t)(g * 16), rr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    xbuf[(quad * 16 + i) * 32 + lane] = fmaxf(__uint_as_float(rr[i]) + bias44[g * 16 + i], 0.f);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + cs) : "memory");
                float f4[4];                                                             // channels g*16 + 4*quad + i of crop `lane`
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ch = 4 * quad + i;
                    f4[i] = ((xbuf[(0 * 16 + ch) * 32 + lane] + xbuf[(1 * 16 + ch) * 32 + lane]) +
                             (xbuf[(2 * 16 + ch) * 32 + lane] + xbuf[(3 * 16 + ch) * 32 + lane])) * 0.25f;   // global_pool: mean over 2x2
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + cs) : "memory");
                const int ch = g * 16 + 4 * quad;
                if (p.tiled) {                                                           // operand layout of the global-head GEMM
                    const long long cg = p.crop_base + crop;
                    const long long b = cg >> 6;
                    const int sq = (int)(cg & 63);
                    // two planes for the split-tf32 GEMM of the global head: hi = what a tf32 operand keeps, lo = the exact remainder
                    float hi4[4], lo4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { hi4[i] = __uint_as_float(__float_as_uint(f4[i]) & 0xffffe000u); lo4[i] = f4[i] - hi4[i]; }
                    float4* ftp = reinterpret_cast<float4*>(p.features) + (((b >> 7) * 2) * 7680 + sq * 120 + (ch >> 2)) * 128 + (b & 127);
                    ftp[0] = make_float4(hi4[0], hi4[1], hi4[2], hi4[3]);
                    ftp[7680 * 128] = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
                } else {
                    *reinterpret_cast<float4*>(p.features + crop * 480 + ch) = make_float4(f4[0], f4[1], f4[2], f4[3]);
                }
#pragma unroll
                for (int r = 0; r < 10; ++r) {                                           // type_head rows 0..6, color_head rows 7..9
                    const float4 w4 = *reinterpret_cast<const float4*>(hw + r * 480 + ch);
                    part_acc[r] = fmaf(f4[0], w4.x, fmaf(f4[1], w4.y, fmaf(f4[2], w4.z, fmaf(f4[3], w4.w, part_acc[r]))));
                }
            }
            SD_MARK(30);
            float* red = reinterpret_cast<float*>(EH + 16384);               // [16 warps][32 crops][10]
#pragma unroll
            for (int r = 0; r < 10; ++r) red[(warp * 32 + lane) * 10 + r] = part_acc[r];
            tc_fence_before();
            __syncthreads();
            if (tid < 32 * 13) {                                             // combine_type_color (common.py:24)
                const int c = tid / 13, cls = tid - c * 13;
                const int ti = kTypeOf[cls], ci = 7 + kColorOf[cls];
                float t = hb[ti], cl = hb[ci];
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                    t += red[(w * 32 + c) * 10 + ti];
                    cl += red[(w * 32 + c) * 10 + ci];
                }
                int64_t oc = (int64_t)tile * 32 + c;
                if (p.perm_boards > 0) {
                    int64_t pb;
                    int psq;
                    perm_inv(oc, p.perm_boards, pb, psq);
                    oc = pb * 64 + psq;
                }
                p.squares[oc * 13 + cls] = t + cl;
            }
            SD_MARK(31);
            __syncthreads();                     // heads blob and partials fully consumed
        }
    }
#ifdef CV_SC_PROFILE
    if (prof_on) for (int i = 0; i < 40; ++i) g_sd_prof[i] = (unsigned long long)pacc[i];
#endif
    if (F16 && bad != 0) atomicOr(p.ovf, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// =====================================================================================================================
// stage C: blocks.2.0 .. blocks.2.5 (19 conv layers), 8x8 -> 4x4 maps.  Tile = 16 crops.
// =====================================================================================================================
// TMEM accumulator columns -> (+bias) -> 16-bit -> global T8-chunked tile dst[chunk][128 rows][8] (coalesced 16-byte stores).
// `dst` already points at this thread's row slot (tile base + row, or the permuted position of its crop); chunks are 128 slots apart.
template <bool F16>
__device__ __forceinline__ void epi_to_global(uint32_t trow, int col0, int ncols, const float* bias, uint4* dst, int cs, int n_slices,
                                              uint32_t& bad) {
    for (int g = cs; g < (ncols >> 4); g += n_slices) {
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)(col0 + g * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float v[8];
            load8(bias + g * 16 + j * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __uint_as_float(r[8 * j + i]);
            const uint4 o = pack8<F16>(v);
            if (F16) bad |= f16x2_nonfinite(o.x) | f16x2_nonfinite(o.y) | f16x2_nonfinite(o.z) | f16x2_nonfinite(o.w);
            dst[(size_t)(g * 2 + j) * 128] = o;
        }
    }
}

// ---- row-streaming 5x5 depthwise kernels of the 8x8 phase --------------------------------------------------------------------
// Both take a task = (one output row, 4 channels): the thread streams over the (up to 5) input rows; each row is loaded once
// (8 pixels x 8 bytes, every load issued before the first use), converted once, and feeds every output of the row from
// registers with static column-border handling -- long runs of independent FMAs instead of load -> convert -> FMA chains with
// a border branch per tap.  The source layouts are pixel-major with the (chunk, crop, channel-half) index fastest, so a half
// warp reads 128 contiguous bytes per pixel: no bank conflicts.

// blocks.2.0.dw_start: 5x5 stride 1, 32 channels, no activation.  src = stage input of the whole tile: 8 sub-tiles (2 crops each)
// x [64 pixels][4 chunks][2 crops][8 ch]  ("P2X").  dst = 8 operand images [4 chunks][128 rows = pixel*2 + crop][8 ch].
template <bool F16>
__device__ __forceinline__ void dw5x5_rows(const uint8_t* src, uint8_t* dst, const float* w, const float* bias, int tid) {
#pragma unroll 1
    for (int task = tid; task < 1024; task += NT) {
        const int l = task & 15, t = ((task >> 8) << 1) | ((task >> 4) & 1), y = (task >> 5) & 7;     // y is warp-uniform
        const uint8_t* sp = src + t * 8192 + l * 8;
        const int coff = (l >> 2) * 8 + (l & 1) * 4;             // first of this task's 4 channels
        const float4 b = *reinterpret_cast<const float4*>(bias + coff);
        float2 acc[8][2];
#pragma unroll
        for (int ox = 0; ox < 8; ++ox) { acc[ox][0] = lo2(b); acc[ox][1] = hi2(b); }
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int iy = y + ky - 2;
            if (iy < 0 || iy > 7) continue;
            uint2 in[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) in[x] = *reinterpret_cast<const uint2*>(sp + (iy * 8 + x) * 128);
            float4 wt[5];
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) wt[kx] = *reinterpret_cast<const float4*>(w + (ky * 5 + kx) * 32 + coff);
            float2 xv[8][2];
#pragma unroll
            for (int x = 0; x < 8; ++x) { xv[x][0] = up2<F16>(in[x].x); xv[x][1] = up2<F16>(in[x].y); }
#pragma unroll
            for (int ox = 0; ox < 8; ++ox) {
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const int ix = ox + kx - 2;
                    if (ix < 0 || ix > 7) continue;
                    acc[ox][0] = fma2(xv[ix][0], lo2(wt[kx]), acc[ox][0]);
                    acc[ox][1] = fma2(xv[ix][1], hi2(wt[kx]), acc[ox][1]);
                }
            }
        }
        uint8_t* dp = dst + t * 8192 + (l >> 2) * 2048 + ((l >> 1) & 1) * 16 + (l & 1) * 8 + y * 8 * 32;
#pragma unroll
        for (int ox = 0; ox < 8; ++ox) *reinterpret_cast<uint2*>(dp + ox * 32) = make_uint2(pk2<F16>(acc[ox][0].x, acc[ox][0].y), pk2<F16>(acc[ox][1].x, acc[ox][1].y));
    }
}

// E6: the expanded 8x8 activation (96 ch) of one sub-tile (2 crops), pixel-major: byte = pixel*384 + ((chunk ^ (pixel & 3))*2 + crop)*16.
// The XOR spreads the four pixels an epilogue quarter-warp writes over the four 32-byte bank groups; a group of four chunks
// stays one contiguous 128-byte block per pixel for the depthwise reader.
__device__ __forceinline__ int e6_off(int pix, int chunk, int crop) { return pix * 384 + (((chunk ^ (pix & 3)) << 1) | crop) * 16; }

// TMEM accumulator columns [col0, col0+96) of this thread's row (= pixel*2 + crop) -> (+bias, ReLU) -> bf16 -> E6.
template <bool F16>
__device__ __forceinline__ void epi_to_e6(uint32_t trow, int col0, const float* bias, uint8_t* dst, int row, int cs, int n_slices) {
    const int pix = row >> 1, crop = row & 1;
    for (int g = cs; g < 6; g += n_slices) {
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)(col0 + g * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float v[8];
            load8(bias + g * 16 + j * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __uint_as_float(r[8 * j + i]);
            *reinterpret_cast<uint4*>(dst + e6_off(pix, g * 2 + j, crop)) =
                make_uint4(pk2r<F16>(v[0], v[1]), pk2r<F16>(v[2], v[3]), pk2r<F16>(v[4], v[5]), pk2r<F16>(v[6], v[7]));
        }
    }
}

// blocks.2.0.dw_mid: 5x5 stride 2 (+ReLU) 8x8 -> 4x4, 96 channels, over two sub-tiles (4 crops): src0 / src1 in the E6 layout ->
// dst P8 tile rows opix*8 + crop0 + 2*sub + crop of [12 chunks][128][8 ch].  384 tasks = (sub, output row, 4 channels).
template <bool F16>
__device__ __forceinline__ void dw5x5s2_rows(const uint8_t* src0, const uint8_t* src1, uint8_t* dst, int crop0, const float* w, const float* bias,
                                             int tid) {
    if (tid >= 384) return;
    const int l = tid & 15, sub = (tid >> 4) & 1, oy = (tid >> 5) & 3, g = tid >> 7;          // oy and g are warp-uniform
    const int chunk = g * 4 + (l >> 2), crop = (l >> 1) & 1, half = l & 1;
    const uint8_t* sp = (sub ? src1 : src0) + half * 8;
    const int coff = chunk * 8 + half * 4;
    const float4 b = *reinterpret_cast<const float4*>(bias + coff);
    float2 acc[4][2];
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) { acc[ox][0] = lo2(b); acc[ox][1] = hi2(b); }
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
        const int iy = 2 * oy + ky - 2;
        if (iy < 0 || iy > 7) continue;
        uint2 in[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) in[x] = *reinterpret_cast<const uint2*>(sp + e6_off(iy * 8 + x, chunk, crop));
        float4 wt[5];
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) wt[kx] = *reinterpret_cast<const float4*>(w + (ky * 5 + kx) * 96 + coff);
        float2 xv[8][2];
#pragma unroll
        for (int x = 0; x < 8; ++x) { xv[x][0] = up2<F16>(in[x].x); xv[x][1] = up2<F16>(in[x].y); }
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                const int ix = 2 * ox + kx - 2;
                if (ix < 0 || ix > 7) continue;
                acc[ox][0] = fma2(xv[ix][0], lo2(wt[kx]), acc[ox][0]);
                    acc[ox][1] = fma2(xv[ix][1], hi2(wt[kx]), acc[ox][1]);
            }
        }
    }
    uint8_t* dp = dst + chunk * 2048 + (oy * 4 * 8 + crop0 + 2 * sub + crop) * 16 + half * 8;
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) *reinterpret_cast<uint2*>(dp + ox * 128) = make_uint2(pk2r<F16>(acc[ox][0].x, acc[ox][0].y), pk2r<F16>(acc[ox][1].x, acc[ox][1].y));
}

namespace sc {
constexpr int NOPS = 20;
// op: 0 dw5  1 pw6  2 dw7 | 3 pw8 | 4+3j pw_exp  5+3j dw_mid  6+3j pw_proj (blocks.2.1..2.4, j = 0..3) |
//     16 dw21 + pw22 (output columns 0..95)  17 pw22 (columns 96..191)  18 pw23 (K rows 0..95)  19 pw23 (K rows 96..191)
constexpr int OFF_A5 = 0;                 // 8 x 8192: blocks.2.0.dw_start output of the whole tile (one operand image per 2-crop sub-tile)
constexpr int OFF_R = 65536;              // 73728: E6a (24576) | A7 (2 x 24576)   ||  body: E (2 x 24576)  ||  E22a (49152) | E22b runs into X16
constexpr int OFF_X16 = 139264;           // 2 x 12288: block input operand tiles (2 M-tiles of 8 crops, P8 rows, 48 ch); E6b during the 8x8 phase
constexpr int OFF_IN = OFF_R + 24576;     // 65536: the tile's stage input (8 sub-tiles, P2X) lies over A7 + X16, both dead until dw_start has consumed it
constexpr int OFF_W = 163840;             // weight arena, a ring of three slots: 0 (26112; also the resident 8x8-phase blobs) | 1 | 2 (20992 each)
constexpr int W_SLOT1 = 26112, W_SLOT2 = W_SLOT1 + 20992;
constexpr int W_ARENA = W_SLOT2 + 20992;  // 68096
constexpr int OFF_BAR = OFF_W + W_ARENA;  // 169984
constexpr int SMEM = OFF_BAR + 64;
constexpr int H_OFF1 = 3328, H_OFF2 = 16000;    // slot 0 during the 8x8 phase: dw5 @0 (3328) | pw6 @3328 (12672) | dw7 @16000 (9984)
constexpr int ACC = 0;                    // TMEM: accumulators [0,384); residual stream (48 ch) of M-tile m at S_COL + 48 m
constexpr int S_COL = 400;
}  // namespace sc

struct StageCParams {
    const bf16* x;            // stage input: P2 tiles (128 rows = 2 crops, row = pix*2 + crop) x 32 ch, T8 chunking
    const uint8_t* wimg;      // CV_W_REPLICAS copies w_stride bytes apart; CTA b streams from copy b % w_rep
    int w_rep;
    uint32_t w_stride;
    bf16* y;                  // stage output: P8 tiles (128 rows = 8 crops, row = pix*8 + crop) x 48 ch  == stage D input
    int n_tiles;              // n_crops / 16
    int perm_boards;          // > 0: write the output tiles in the permuted crop order of a launch of this many boards (perm_pos)
    uint32_t off[sc::NOPS], bytes[sc::NOPS];
    int debug;                // CV_SC_DEBUG ablation bits (timing experiments only: results are wrong): 1 no dw5x5, 2 no dw5x5s2, 4 no dw3x3,
                              // 32 no MMA / epilogue in the 4x4 phase (what is left is the per-op barrier + weight-stream cost), 256 section timers
    const int* gate;          // non-null: the kernel runs only when (*gate != 0) == gate_want (fp16 pass: want 0; bf16 fall-back pass: want 1)
    int gate_want;
    int* ovf;                 // fp16 pass: set to 1 when a residual-stream value left the fp16 range (the bf16 pass then redoes the wave)
};

// Timing experiments (-DCV_SC_PROFILE, CV_SC_DEBUG & 256): cycles thread 0 of CTA 0 spends per section of a tile
// (0 wait weights+input | 1 dw_start | 2 pw6 MMA | 3 pw6 epilogue | 4 dw_mid 5x5 s2 | 8+op: 4x4-phase op).  clock64 is not ordered
// against bar.sync: a section can contain the barrier wait of its neighbour -- trust sums, not single sections.
#ifdef CV_SC_PROFILE
__device__ unsigned long long g_sc_prof[40];
#define SC_MARK(k) do { if (prof_on) { const long long _n = clock64(); pacc[k] += _n - plast; plast = _n; } } while (0)
#else
#define SC_MARK(k)
#endif

template <bool F16>
__global__ void __launch_bounds__(NT, 1) stageC_kernel(const __grid_constant__ StageCParams p) {
    using namespace sc;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_want != 0)) return;
    uint32_t bad = 0;
    const uint8_t* wsrc = p.wimg + (size_t)(blockIdx.x % p.w_rep) * p.w_stride;
    constexpr int WP = F16 ? 1 : 2;                 // weight images per pointwise blob: fp16 W | bf16 W_hi, W_lo
    uint8_t* IN = smem + OFF_IN;
    uint8_t* A5 = smem + OFF_A5;
    uint8_t* R = smem + OFF_R;
    uint8_t* E6 = R;
    uint8_t* A7 = R + 24576;
    uint8_t* X16 = smem + OFF_X16;
    uint8_t* WA = smem + OFF_W;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* wbar = bars;                          // one per weight slot
    uint64_t* inbar = bars + 3;
    uint64_t* mbar = bars + 4;
    uint64_t* mbar2 = bars + 6;                     // 4x4 phase: second M-tile of a GEMM op (its own commit: see wait_mma2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, cs = warp >> 2, row = quad * 32 + lane;
    const int mt = cs & 1, half = cs >> 1;          // body epilogues: M-tile and column half handled by this warp

    if (tid == 0) {
        mbar_init(wbar, 1); mbar_init(wbar + 1, 1); mbar_init(wbar + 2, 1); mbar_init(inbar, 1); mbar_init(mbar, 1); mbar_init(mbar2, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t wph = 0, inph = 0, mph = 0, mph2 = 0;    // wph: phase bit per weight slot
#ifdef CV_SC_PROFILE
    const bool prof_on = (p.debug & 256) && blockIdx.x == 0 && tid == 0;
    long long pacc[40];
    for (int i = 0; i < 40; ++i) pacc[i] = 0;
    long long plast = clock64();
#endif

    auto load_head_weights = [&]() {
        mbar_arrive_expect_tx(wbar, p.bytes[0] + p.bytes[1] + p.bytes[2]);
        bulk_g2s(WA, wsrc + p.off[0], p.bytes[0], wbar);
        bulk_g2s(WA + H_OFF1, wsrc + p.off[1], p.bytes[1], wbar);
        bulk_g2s(WA + H_OFF2, wsrc + p.off[2], p.bytes[2], wbar);
    };
    auto load_in = [&](int tile) {             // the whole tile's stage input: 8 sub-tiles = one 64 KB bulk copy
        mbar_arrive_expect_tx(inbar, 65536);
        bulk_g2s(IN, reinterpret_cast<const uint8_t*>(p.x) + (size_t)tile * 65536, 65536, inbar);
    };
    // Weights stream L2 -> smem TWO ops ahead through the three-slot ring (a blob is 4-21 KB and its fetch takes about as long
    // as two of the short 4x4-phase ops: with one op of look-ahead every op waited for its weights).  Op `op` >= 3 lives in
    // slot (op - 2) % 3; the slot being refilled at begin_op(op) held op - 1, whose last reader passed a CTA barrier.
    auto slot_of = [&](int op) { return (op - 2) % 3; };
    auto slot_ptr = [&](int slot) { return WA + (slot == 0 ? 0 : slot == 1 ? W_SLOT1 : W_SLOT2); };
    auto prefetch = [&](int op) {
        const int sl = slot_of(op);
        mbar_arrive_expect_tx(wbar + sl, p.bytes[op]);
        bulk_g2s(slot_ptr(sl), wsrc + p.off[op], p.bytes[op], wbar + sl);
    };
    auto wait_slot = [&](int sl) {
#ifdef CV_SC_PROFILE
        if (prof_on) {                           // how often, and for how long, the weights of an op are NOT there when it begins
            const long long t = clock64();
            if (!mbar_try_wait(wbar + sl, (wph >> sl) & 1u)) { ++pacc[6]; mbar_wait(wbar + sl, (wph >> sl) & 1u); }
            pacc[5] += clock64() - t; ++pacc[7];
            wph ^= 1u << sl;
            return;
        }
#endif
        mbar_wait(wbar + sl, (wph >> sl) & 1u);
        wph ^= 1u << sl;
    };
    // Issuing a bulk copy costs its thread a few hundred cycles.  GEMM ops (`pf` false here): a warp that is NOT the MMA issuer does it
    // right after the barrier that precedes the MMA issue, while it would otherwise wait for the tensor pipe (the issuing warp is the
    // one every op's closing barrier waits for); depthwise ops: the last warp, which has no depthwise task (384 tasks on 512 threads).
    constexpr int PF_WARP = 15;
    auto begin_op = [&](int op, bool pf) -> uint8_t* {
        if (pf && warp == PF_WARP && op + 2 < NOPS && elect_one()) prefetch(op + 2);
        const int sl = slot_of(op);
        wait_slot(sl);
        return slot_ptr(sl);
    };
    auto sync_before_mma = [&]() {
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    auto wait_mma = [&]() {
        mbar_wait(mbar, mph);
        mph ^= 1u;
        tc_fence_after();
    };
    // 4x4 phase: the two M-tiles of a GEMM op are committed separately and ALL 16 warps run the epilogue of M-tile 0 (four column
    // slices) while the tensor pipe still works on M-tile 1, then that of M-tile 1 (slices rotated by two, so that every warp gets the
    // same number of 16-column groups over both): the epilogue starts one M-tile's worth of MMAs (~350 cycles) earlier per op.
    auto wait_mma2 = [&]() {
        mbar_wait(mbar2, mph2);
        mph2 ^= 1u;
        tc_fence_after();
    };
    const int cs2 = (cs + 2) & 3;

    if (tid == 0 && blockIdx.x < p.n_tiles) {
        load_head_weights();
        load_in(blockIdx.x);
    }

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // ------------------------------ blocks.2.0 at 8x8: dw_start over the whole tile, then four pairs of 2-crop sub-tiles ----------
        SC_MARK(39);
        wait_slot(0);
        const float* b5 = reinterpret_cast<const float*>(WA);
        const float* w5 = b5 + 32;
        const float* b6 = reinterpret_cast<const float*>(WA + H_OFF1);
        const uint32_t w6 = smem_u32(WA + H_OFF1 + 96 * 4);
        const float* b7 = reinterpret_cast<const float*>(WA + H_OFF2);
        const float* w7 = b7 + 96;
        mbar_wait(inbar, inph); inph ^= 1u;
        SC_MARK(0);
        if (!(p.debug & 1)) dw5x5_rows<F16>(IN, A5, w5, b5, tid);                                // L5 dw_start 5x5 (no act), all 16 crops
        sync_before_mma();
        SC_MARK(1);
        // L6 pw_exp 32 -> 96 on sub-tiles 2j, 2j+1 into accumulator set j & 1: the GEMMs of pair j + 1 are issued BEFORE the depthwise
        // of pair j (they only read A5), so their round trip through the tensor pipe runs under it instead of in front of every epilogue.
        auto issue_pw6 = [&](int j) {
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                for (int m = 0; m < 2; ++m)
                    issue_gemm<F16>(smem_u32(A5 + (2 * j + m) * 8192), 32, w6, 96, 0, 96, tmem + ACC + 192 * (j & 1) + 96 * m, false, WP);
                mma_commit(mbar);
            }
        };
        issue_pw6(0);
        if (warp == PF_WARP && elect_one()) { prefetch(3); prefetch(4); }   // first two ops of the 4x4 phase -> slots 1, 2 while the 8x8 phase runs
        for (int j = 0; j < 4; ++j) {
            wait_mma();
            SC_MARK(2);
            epi_to_e6<F16>(trow, ACC + 192 * (j & 1) + 96 * mt, b6, mt ? X16 : E6, row, half, 2);
            tc_fence_before();
            __syncthreads();
            SC_MARK(3);
            if (j < 3) issue_pw6(j + 1);                // set (j + 1) & 1 was drained by the epilogue of pair j - 1, two barriers ago
            if (!(p.debug & 2)) dw5x5s2_rows<F16>(E6, X16, A7 + (j >> 1) * 24576, (4 * j) & 7, w7, b7, tid);   // L7 dw_mid 5x5 s2 (+ReLU) -> 4x4 P8 tile
            __syncthreads();
            SC_MARK(4);
        }
        // ------------------------------ 4x4 phase: 2 M-tiles of 8 crops -----------------------------------------------------------
        int op = 3;
        {   // L8 blocks.2.0.pw_proj 96 -> 48: starts the residual stream
            uint8_t* wb = begin_op(op, false);
            sync_before_mma();
            if (!(p.debug & 32) && warp == 0 && elect_one()) {
                tc_fence_after();
                for (int m = 0; m < 2; ++m) {
                    issue_gemm<F16>(smem_u32(A7 + m * 24576), 96, smem_u32(wb + 192), 48, 0, 48, tmem + S_COL + 48 * m, false, WP);
                    mma_commit(m ? mbar2 : mbar);
                }
            }
            if (warp == PF_WARP && op + 2 < NOPS && elect_one()) prefetch(op + 2);
            if (!(p.debug & 32)) {
                wait_mma();
                epi_to_tile<F16, false>(trow, S_COL, 48, reinterpret_cast<const float*>(wb), X16, 0, row, cs, 4, bad);
                wait_mma2();
                epi_to_tile<F16, false>(trow, S_COL + 48, 48, reinterpret_cast<const float*>(wb), X16 + 12288, 0, row, cs2, 4, bad);
            }
            SC_MARK(8 + op); ++op;                       // no closing barrier: the next op opens with sync_before_mma()
        }
#pragma unroll 1
        for (int blk = 1; blk <= 4; ++blk) {
            {   // pw_exp 48 -> 96 (+ReLU)
                uint8_t* wb = begin_op(op, false);
                SC_MARK(28);
                sync_before_mma();
                SC_MARK(29);
                if (!(p.debug & 32) && warp == 0 && elect_one()) {
                    tc_fence_after();
                    for (int m = 0; m < 2; ++m) {
                        issue_gemm<F16>(smem_u32(X16 + m * 12288), 48, smem_u32(wb + 384), 96, 0, 96, tmem + ACC + 96 * m, false, WP);
                        mma_commit(m ? mbar2 : mbar);
                    }
                }
                if (warp == PF_WARP && op + 2 < NOPS && elect_one()) prefetch(op + 2);
                SC_MARK(30);
                if (!(p.debug & 32)) {
                    wait_mma();
                    SC_MARK(31);
                    epi_to_tile<F16, true>(trow, ACC, 96, reinterpret_cast<const float*>(wb), R, 0, row, cs, 4, bad);
                    SC_MARK(32);
                    wait_mma2();
                    SC_MARK(33);
                    epi_to_tile<F16, true>(trow, ACC + 96, 96, reinterpret_cast<const float*>(wb), R + 24576, 0, row, cs2, 4, bad);
                    SC_MARK(34);
                }
                __syncthreads();
                SC_MARK(8 + op); ++op;
            }
            {   // dw_mid 3x3 (+ReLU), in place on both M-tiles
                uint8_t* wb = begin_op(op, true);
                SC_MARK(35);
                const float* b = reinterpret_cast<const float*>(wb);
                if (!(p.debug & 4)) dw3x3_p8_rt<F16, true>(R, R, 2 * 12 * 16, 12, b + 96, b, tid);
                SC_MARK(36);
                SC_MARK(8 + op); ++op;                   // no closing barrier: pw_proj opens with sync_before_mma()
            }
            {   // pw_proj 96 -> 48 accumulated onto the residual stream
                uint8_t* wb = begin_op(op, false);
                sync_before_mma();
                if (!(p.debug & 32) && warp == 0 && elect_one()) {
                    tc_fence_after();
                    for (int m = 0; m < 2; ++m) {
                        issue_gemm<F16>(smem_u32(R + m * 24576), 96, smem_u32(wb + 192), 48, 0, 48, tmem + S_COL + 48 * m, true, WP);
                        mma_commit(m ? mbar2 : mbar);
                    }
                }
                if (warp == PF_WARP && op + 2 < NOPS && elect_one()) prefetch(op + 2);
                if (!(p.debug & 32)) {
                    wait_mma();
                    epi_to_tile<F16, false>(trow, S_COL, 48, reinterpret_cast<const float*>(wb), X16, 0, row, cs, 4, bad);
                    wait_mma2();
                    epi_to_tile<F16, false>(trow, S_COL + 48, 48, reinterpret_cast<const float*>(wb), X16 + 12288, 0, row, cs2, 4, bad);
                }
                if (blk == 4) __syncthreads();           // blocks.2.5.dw_start reads X16 next; the other blocks go on with sync_before_mma()
                SC_MARK(8 + op); ++op;
            }
        }
        // ------------------------------ blocks.2.5: dw_start 3x3, pw_exp 48 -> 192 (two column halves), pw_proj 192 -> 48 ------------
        uint8_t* E22a = R;
        uint8_t* E22b = R + 49152;               // second half runs into the X16 region (dead once the L22 MMAs have read it)
        {   // op 16: [dw21 blob 1920 B][bias22[0:96] | W22 columns 0..95]
            uint8_t* wb = begin_op(op, false);
            const float* b21 = reinterpret_cast<const float*>(wb);
            if (!(p.debug & 4)) dw3x3_p8_rt<F16, false>(X16, X16, 2 * 6 * 16, 6, b21 + 48, b21, tid);
            sync_before_mma();
            if (!(p.debug & 32) && warp == 0 && elect_one()) {
                tc_fence_after();
                for (int m = 0; m < 2; ++m) {
                    issue_gemm<F16>(smem_u32(X16 + m * 12288), 48, smem_u32(wb + 1920 + 384), 96, 0, 96, tmem + ACC + 96 * m, false, WP);
                    mma_commit(m ? mbar2 : mbar);
                }
            }
            if (warp == PF_WARP && op + 2 < NOPS && elect_one()) prefetch(op + 2);
            if (!(p.debug & 32)) {
                wait_mma();
                epi_to_tile<F16, true>(trow, ACC, 96, reinterpret_cast<const float*>(wb + 1920), E22a, 0, row, cs, 4, bad);
                wait_mma2();
                epi_to_tile<F16, true>(trow, ACC + 96, 96, reinterpret_cast<const float*>(wb + 1920), E22a + 24576, 0, row, cs2, 4, bad);
            }
            SC_MARK(8 + op); ++op;
        }
        {   // op 17: W22 columns 96..191
            uint8_t* wb = begin_op(op, false);
            sync_before_mma();
            if (!(p.debug & 32) && warp == 0 && elect_one()) {
                tc_fence_after();
                for (int m = 0; m < 2; ++m) {
                    issue_gemm<F16>(smem_u32(X16 + m * 12288), 48, smem_u32(wb + 384), 96, 0, 96, tmem + ACC + 192 + 96 * m, false, WP);
                    mma_commit(m ? mbar2 : mbar);
                }
            }
            if (warp == PF_WARP && op + 2 < NOPS && elect_one()) prefetch(op + 2);
            if (!(p.debug & 32)) {
                wait_mma();
                epi_to_tile<F16, true>(trow, ACC + 192, 96, reinterpret_cast<const float*>(wb), E22b, 0, row, cs, 4, bad);
                wait_mma2();
                epi_to_tile<F16, true>(trow, ACC + 192 + 96, 96, reinterpret_cast<const float*>(wb), E22b + 24576, 0, row, cs2, 4, bad);
            }
            SC_MARK(8 + op); ++op;
        }
        const float* cum23;
        {   // op 18: W23 K rows 0..95 (+ cumulative bias)
            uint8_t* wb = begin_op(op, false);
            cum23 = reinterpret_cast<const float*>(wb);
            sync_before_mma();
            if (!(p.debug & 32) && warp == 0 && elect_one()) {
                tc_fence_after();
                for (int m = 0; m < 2; ++m)
                    issue_gemm<F16>(smem_u32(E22a + m * 24576), 96, smem_u32(wb + 192), 48, 0, 48, tmem + S_COL + 48 * m, true, WP);
            }
            SC_MARK(8 + op); ++op;
        }
        {   // op 19: W23 K rows 96..191, then the stage output
            uint8_t* wb = begin_op(op, false);
            if (!(p.debug & 32) && warp == 0 && elect_one()) {
                for (int m = 0; m < 2; ++m)
                    issue_gemm<F16>(smem_u32(E22b + m * 24576), 96, smem_u32(wb), 48, 0, 48, tmem + S_COL + 48 * m, true, WP);
                mma_commit(mbar);
            }
            if (!(p.debug & 32)) wait_mma();
            // row = pixel * 8 + crop_local of M-tile mt; its P8 slot in the hand-off: the same position, or the crop's permuted position
            int64_t pos = ((int64_t)tile * 2 + mt) * 8 + (row & 7);
            if (p.perm_boards > 0) pos = perm_pos(pos, p.perm_boards);
            uint4* dst = reinterpret_cast<uint4*>(p.y) + (pos >> 3) * 6 * 128 + (row >> 3) * 8 + (pos & 7);
            if (!(p.debug & 32)) epi_to_global<F16>(trow, S_COL + 48 * mt, 48, cum23, dst, half, 2, bad);
            SC_MARK(8 + op); ++op;
        }
        const int next = tile + gridDim.x;
        tc_fence_before();
        __syncthreads();
        if (tid == 0 && next < p.n_tiles) {
            load_head_weights();
            load_in(next);
        }
    }
#ifdef CV_SC_PROFILE
    if (prof_on) for (int i = 0; i < 40; ++i) g_sc_prof[i] = (unsigned long long)pacc[i];
#endif
    if (F16 && bad != 0) atomicOr(p.ovf, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// =====================================================================================================================
// stage B: blocks.0.1 (1x1 16->16 @16x16), blocks.1.0 (3x3 s2 16->48 -> 8x8), blocks.1.1 (1x1 48->32).  Tile = 2 crops,
// 256 threads, ~103 KB smem and 256 TMEM columns per CTA so that TWO CTAs share an SM and overlap each other's
// MMA / epilogue phases.  All weights (hi + lo) stay resident in shared memory for the whole kernel.
// =====================================================================================================================
namespace sb {
constexpr int NTB = 256;
constexpr int W_BYTES_MAX = 384 + 1024 + 27648 + 6144;     // [b2 16 | b3 48 | b4 32] fp32, then W2, W3, W4: bf16 hi|lo pairs (fp16: single images, half of it)
constexpr int OFF_W = 0;
constexpr int OFF_A3 = 35328;             // 36864: im2col image of the blocks.0.1 output = A operand of blocks.1.0 [18][128][8]
constexpr int OFF_IN = OFF_A3 + 36864;    // 2 x 16384: input ring (2 crops = 4 T8 tiles of 128 rows x 16 ch); the consumed slot is reused as A4
constexpr int OFF_BAR = OFF_IN + 32768;   // 104960
constexpr int SMEM = OFF_BAR + 64;
constexpr int w_bytes(bool f16) { return 384 + (512 + 13824 + 3072) * (f16 ? 1 : 2); }
}  // namespace sb

struct StageBParams {
    const bf16* x;            // front-end output: T8 tiles, rows = crop*256 + pixel (16x16), 16 ch
    const uint8_t* wimg;      // sb::w_bytes(F16)
    bf16* y;                  // P2 tiles (128 rows = 2 crops at 8x8, row = pix*2 + crop) x 32 ch == stage C input
    int n_tiles;              // n_crops / 2
    const int* gate;          // non-null: the kernel runs only when (*gate != 0) == gate_want (fp16 pass: want 0; bf16 fall-back pass: want 1)
    int gate_want;
    int* ovf;                 // fp16 pass: set to 1 when a residual-stream value left the fp16 range (the bf16 pass then redoes the wave)
};

template <bool F16>
__global__ void __launch_bounds__(sb::NTB, 2) stageB_kernel(const __grid_constant__ StageBParams p) {
    using namespace sb;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_want != 0)) return;
    uint32_t bad = 0;                 // every output of this stage passes a ReLU conversion: nothing to check, +inf travels on to stage C
    constexpr int WP = F16 ? 1 : 2;
    constexpr int W2_OFF = 384, W3_OFF = W2_OFF + 512 * WP, W4_OFF = W3_OFF + 13824 * WP, W_BYTES = W4_OFF + 3072 * WP;
    uint8_t* W = smem + OFF_W;
    uint8_t* A3 = smem + OFF_A3;
    uint8_t* IN = smem + OFF_IN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* wbar = bars;
    uint64_t* inbar = bars + 1;       // [2]
    uint64_t* mbar = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, hi2 = warp >> 2, row = quad * 32 + lane;

    if (tid == 0) {
        mbar_init(wbar, 1); mbar_init(inbar, 1); mbar_init(inbar + 1, 1); mbar_init(mbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    for (int i = tid; i < 36864 / 16; i += NTB) reinterpret_cast<uint4*>(A3)[i] = make_uint4(0, 0, 0, 0);   // padding taps stay zero
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t inph0 = 0, inph1 = 0, mph = 0;
    const float* b2 = reinterpret_cast<const float*>(W);
    const float* b3 = b2 + 16;
    const float* b4 = b2 + 64;

    auto load_in = [&](int tile, int s) {
        mbar_arrive_expect_tx(inbar + s, 16384);
        bulk_g2s(IN + s * 16384, reinterpret_cast<const uint8_t*>(p.x) + (size_t)tile * 16384, 16384, inbar + s);
    };
    auto wait_mma = [&]() {
        mbar_wait(mbar, mph);
        mph ^= 1u;
        tc_fence_after();
    };
    auto sync_before_mma = [&]() {
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    if (tid == 0) {
        mbar_arrive_expect_tx(wbar, W_BYTES);
        bulk_g2s(W, p.wimg, W_BYTES, wbar);
        if (blockIdx.x < p.n_tiles) load_in(blockIdx.x, 0);
    }
    mbar_wait(wbar, 0);

    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        uint8_t* in = IN + s * 16384;
        if (tid == 0 && tile + (int)gridDim.x < p.n_tiles) load_in(tile + gridDim.x, s ^ 1);
        if (s) { mbar_wait(inbar + 1, inph1); inph1 ^= 1u; } else { mbar_wait(inbar, inph0); inph0 ^= 1u; }
        // ---- blocks.0.1: 1x1 16 -> 16 (+ReLU) on 512 rows = 4 M-tiles
        //      (issuing this GEMM one stage early, behind the blocks.1.0 GEMM of the previous tile, was measured: no change -- with two
        //      CTAs per SM its round trip already runs under the other CTA's work)
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            for (int j = 0; j < 4; ++j) issue_gemm<F16>(smem_u32(in + j * 4096), 16, smem_u32(W + W2_OFF), 16, 0, 16, tmem + 16 * j, false, WP);
            mma_commit(mbar);
        }
        wait_mma();
        // epilogue: scatter every output pixel into the im2col image of the 3x3 stride-2 conv that follows
        // (A3 chunk = tap*2 + channel-half, row = (oy*8+ox)*2 + crop): a pixel feeds 1, 2 or 4 (tap, output) pairs.
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int j = hi2 * 2 + jj;
            uint32_t r[16];
            tmem_ld16(trow + (uint32_t)(16 * j), r);
            tmem_ld_wait();
            uint4 o[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(r[8 * h + i]) + b2[8 * h + i], 0.f);
                o[h] = pack8<F16>(v);
            }
            const int rg = j * 128 + row, crop = rg >> 8, iy = (rg >> 4) & 15, ix = rg & 15;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int ty = iy + 1 - ky;
                if ((ty & 1) || ty < 0 || ty > 14) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int tx = ix + 1 - kx;
                    if ((tx & 1) || tx < 0 || tx > 14) continue;
                    const int ra = ((ty >> 1) * 8 + (tx >> 1)) * 2 + crop, ch = (ky * 3 + kx) * 2;
                    *reinterpret_cast<uint4*>(A3 + (((size_t)ch * 128 + ra) << 4)) = o[0];
                    *reinterpret_cast<uint4*>(A3 + (((size_t)(ch + 1) * 128 + ra) << 4)) = o[1];
                }
            }
        }
        sync_before_mma();
        // ---- blocks.1.0: 3x3 s2 16 -> 48 (+ReLU) as one K = 144 GEMM on 128 rows (2 crops x 8x8, P2 order)
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            issue_gemm<F16>(smem_u32(A3), 144, smem_u32(W + W3_OFF), 48, 0, 48, tmem + 64, false, WP);
            mma_commit(mbar);
        }
        wait_mma();
        uint8_t* A4 = in;                       // the input slot is dead (its MMAs completed): reuse it for the next operand
        epi_to_tile<F16, true>(trow, 64, 48, b3, A4, 0, row, hi2, 2, bad);
        sync_before_mma();
        // ---- blocks.1.1: 1x1 48 -> 32 (+ReLU) -> global P2 tile
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            issue_gemm<F16>(smem_u32(A4), 48, smem_u32(W + W4_OFF), 32, 0, 32, tmem + 128, false, WP);
            mma_commit(mbar);
        }
        wait_mma();
        {
            uint32_t r[16];
            tmem_ld16(trow + (uint32_t)(128 + 16 * hi2), r);
            tmem_ld_wait();
            uint4* dst = reinterpret_cast<uint4*>(p.y) + (size_t)tile * 4 * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(r[8 * h + i]) + b4[16 * hi2 + 8 * h + i], 0.f);
                dst[(size_t)(row >> 1) * 8 + (hi2 * 2 + h) * 2 + (row & 1)] = pack8<F16>(v);      // P2X: [pixel][chunk][crop]
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}


// crop-major T8 tiles (row = crop*16 + pix) -> P8 tiles (row = pix*8 + crop_local); one thread per 16-byte chunk.
__global__ void permute_p8_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n_chunks) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    const int r = (int)(i & 127);
    out[(i & ~(int64_t)127) + ((r & 15) * 8 + (r >> 4))] = in[i];
}

// crop-major T8 tiles of 8x8 maps, 32 ch (128 rows = 2 crops, row = crop*64 + pix; 4 chunks) -> P2X sub-tiles
// [64 pixels][4 chunks][2 crops] of 16-byte chunks (the stage C input layout).
__global__ void permute_p2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n_chunks) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    const int r = (int)(i & 127), chunk = (int)((i >> 7) & 3);
    out[(i & ~(int64_t)511) + (r & 63) * 8 + chunk * 2 + (r >> 6)] = in[i];
}

// ---- stage weight image construction --------------------------------------------------------------------------------------
// rows [k0, k0+K) x columns [n0, n0+n) of w[.][n_total] -> UMMA B image [K/8][n][8].  bf16: parts == 2 appends the lo image
// (w - bf16(w), itself rounded to bf16) right after the hi image.  fp16 (f16_flag non-null): one image; *f16_flag is raised when a
// weight does not fit the fp16 range (the handle then never takes the fp16 kernels).
__global__ void prep_pw_part_kernel(const float* __restrict__ w, int k0, int K, int n_total, int n0, int n, int parts, uint16_t* __restrict__ dst,
                                    int* __restrict__ f16_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * n) return;
    const int kk = i & 7, j = (i >> 3) % n, kc = (i >> 3) / n;
    const float v = w[(size_t)(k0 + kc * 8 + kk) * n_total + n0 + j];
    if (f16_flag != nullptr) {
        if (!(fabsf(v) <= 65504.f)) atomicOr(f16_flag, 1);
        dst[i] = __half_as_ushort(__float2half_rn(v));
        return;
    }
    const bf16 hi = __float2bfloat16_rn(v);
    dst[i] = __bfloat16_as_ushort(hi);
    if (parts == 2) dst[(size_t)K * n + i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hi)));
}
// 2x2-map depthwise weights: dst[((c8*4 + p)*4 + q)*8 + i] = w[tap(p,q)][c8*8 + i], tap = (qy-py+pad)*K + (qx-px+pad).
__global__ void prep_dw2x2_kernel(const float* __restrict__ w, int K, int C, float* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 16 * C) return;
    const int i = idx & 7, q = (idx >> 3) & 3, p = (idx >> 5) & 3, c8 = idx >> 7, pad = (K - 1) / 2;
    const int tap = ((q >> 1) - (p >> 1) + pad) * K + ((q & 1) - (p & 1) + pad);
    dst[idx] = w[(size_t)tap * C + c8 * 8 + i];
}
__global__ void add_f32_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (a ? a[i] : 0.f) + b[i];
}

struct DOp { int layer, kind; };   // kind 0 depthwise | 1 pointwise (own bias) | 2 pw_proj (cumulative bias of its residual chain) | 3 blocks.4.0 part | 4 heads
const DOp kDOps[sd::NOPS] = {{24, 0}, {25, 1}, {26, 0}, {27, 2}, {28, 0}, {29, 1}, {30, 0}, {31, 2}, {32, 1}, {33, 0}, {34, 2}, {35, 1},
                             {36, 0}, {37, 2}, {38, 1}, {39, 0}, {40, 2}, {41, 1}, {42, 0}, {43, 2}, {44, 3}, {44, 3}, {44, 3}, {-1, 4}};

uint32_t dop_bytes(int op) {
    const cv_layer_info* L = cv_layers();
    const DOp& d = kDOps[op];
    if (d.kind == 4) return (16 + 4800 + 480) * 4;
    const cv_layer_info& l = L[d.layer];
    if (d.kind == 0) return (uint32_t)(l.cout + (l.hin == 2 ? 16 : l.k * l.k) * l.cout) * 4;      // 2x2 maps: [chunk][p][q][8] layout
    if (d.kind == 3) return 64 * 160 * 2;
    return (uint32_t)l.cout * 4 + (uint32_t)l.cin * l.cout * 2;
}

}  // namespace

size_t stageD_image_bytes() {
    size_t n = 0;
    for (int op = 0; op < sd::NOPS; ++op) n += (dop_bytes(op) + 127) / 128 * 128;
    return n;
}

// Builds the stage-D weight image from the packed fp32 blob (device) and fills off[]/bytes[] (host arrays of 24).
// f16_flag non-null: fp16 weight images (and the range check of prep_pw_part_kernel); null: bf16.
int build_stageD_image(const float* blob, uint8_t* img, uint32_t* off, uint32_t* bytes, int* f16_flag, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    size_t o = 0;
    const float* prev_cum = nullptr;
    int part = 0;
    for (int op = 0; op < sd::NOPS; ++op) {
        const DOp& d = kDOps[op];
        off[op] = (uint32_t)o;
        bytes[op] = dop_bytes(op);
        if (bytes[op] % 16 != 0) { cv_set_error("stage D: blob %d size %u not a multiple of 16", op, bytes[op]); return CV_ERR_STATE; }
        uint8_t* dst = img + o;
        if (d.kind == 4) {
            float* f = reinterpret_cast<float*>(dst);
            CV_CUDA(cudaMemsetAsync(f, 0, 64, s));
            CV_CUDA(cudaMemcpyAsync(f, blob + CV_OFF_HEAD_B, 10 * sizeof(float), cudaMemcpyDeviceToDevice, s));
            CV_CUDA(cudaMemcpyAsync(f + 16, blob + CV_OFF_HEAD_W, 4800 * sizeof(float), cudaMemcpyDeviceToDevice, s));
            CV_CUDA(cudaMemcpyAsync(f + 16 + 4800, blob + L[44].b_offset, 480 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        } else {
            const cv_layer_info& l = L[d.layer];
            if (d.kind == 0) {
                float* f = reinterpret_cast<float*>(dst);
                CV_CUDA(cudaMemcpyAsync(f, blob + l.b_offset, l.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
                if (l.hin == 2) {
                    prep_dw2x2_kernel<<<(16 * l.cout + 255) / 256, 256, 0, s>>>(blob + l.w_offset, l.k, l.cout, f + l.cout);
                    CV_CHECK_LAUNCH();
                } else {
                    CV_CUDA(cudaMemcpyAsync(f + l.cout, blob + l.w_offset, (size_t)l.k * l.k * l.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
                }
            } else if (d.kind == 3) {
                prep_pw_part_kernel<<<(64 * 160 + 255) / 256, 256, 0, s>>>(blob + l.w_offset, 0, 64, 480, 160 * part, 160, 1, reinterpret_cast<uint16_t*>(dst), f16_flag);
                CV_CHECK_LAUNCH();
                ++part;
            } else {
                float* f = reinterpret_cast<float*>(dst);
                if (d.kind == 2) {
                    // residual chain: blocks.3.0.pw_proj (no skip) starts it, every later pw_proj adds its bias to the running sum
                    add_f32_kernel<<<1, 64, 0, s>>>(f, l.skip >= 0 ? prev_cum : nullptr, blob + l.b_offset, l.cout);
                    CV_CHECK_LAUNCH();
                    prev_cum = f;
                } else {
                    CV_CUDA(cudaMemcpyAsync(f, blob + l.b_offset, l.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
                }
                prep_pw_part_kernel<<<(l.cin * l.cout + 255) / 256, 256, 0, s>>>(blob + l.w_offset, 0, l.cin, l.cout, 0, l.cout, 1,
                                                                                reinterpret_cast<uint16_t*>(dst + (size_t)l.cout * 4), f16_flag);
                CV_CHECK_LAUNCH();
            }
        }
        o += (bytes[op] + 127) / 128 * 128;
    }
    return CV_OK;
}

int launch_permute_p8(const bf16* in, bf16* out, int64_t n_crops, int C, cudaStream_t s) {
    const int64_t n_chunks = n_crops * 16 * (C / 8);
    if (n_chunks == 0) return CV_OK;
    permute_p8_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n_chunks);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_stageD(const bf16* x_p8, int64_t n_crops, const uint8_t* wimg, const uint32_t* off, const uint32_t* bytes, float* features,
                  int tiled, int64_t crop_base, float* squares, int num_sms, const StageGate& gate, int perm_boards, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if (n_crops % 32 != 0) { cv_set_error("stage D: crop count %lld is not a multiple of 32", (long long)n_crops); return CV_ERR_ARG; }
    StageDParams p{};
    p.w_rep = weight_replicas(); p.w_stride = (uint32_t)stageD_image_bytes();
    p.x = x_p8; p.wimg = wimg; p.features = features; p.squares = squares; p.n_tiles = (int)(n_crops / 32);
    p.tiled = tiled; p.crop_base = crop_base; p.perm_boards = perm_boards;
    p.gate = gate.flag; p.gate_want = gate.want; p.ovf = gate.ovf;
#ifdef CV_EXPERIMENTS                 // ablation / timing switches change the results: compiled in only with -DCV_EXPERIMENTS (CV_NVCC_EXTRA)
    { const char* d = getenv("CV_SD_DEBUG"); p.debug = d ? atoi(d) : 0; }
#endif
    for (int i = 0; i < sd::NOPS; ++i) { p.off[i] = off[i]; p.bytes[i] = bytes[i]; }
    auto kern = gate.f16 ? stageD_kernel<true> : stageD_kernel<false>;
    CV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sd::SMEM));
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    kern<<<grid, NT, sd::SMEM, s>>>(p);
    CV_CHECK_LAUNCH();
#ifdef CV_SC_PROFILE
    if (p.debug & 256) {
        unsigned long long h[40];
        CV_CUDA(cudaStreamSynchronize(s));
        CV_CUDA(cudaMemcpyFromSymbol(h, g_sd_prof, sizeof(h)));
        const double tiles = (double)((p.n_tiles + grid - 1) / grid);
        fprintf(stderr, "stageD cycles per 32-crop tile: wait w %.0f | 4x4 phase (4 sub-tiles): wait in %.0f, dw24 %.0f, bar+pw25 %.0f, epi %.0f, dw26 %.0f, its closing barrier %.0f\n   2x2 ops 3..22:",
                h[39] / tiles, h[0] / tiles, h[1] / tiles, h[2] / tiles, h[3] / tiles, h[4] / tiles, h[5] / tiles);
        double sum = 0;
        for (int op = 3; op < 23; ++op) { fprintf(stderr, " %.0f", h[6 + op] / tiles); sum += h[6 + op] / tiles; }
        fprintf(stderr, "  (sum %.0f) | pool + heads %.0f | combine %.0f\n   weight waits per tile: %.1f of %.1f not ready at first poll, %.0f cycles in all\n", sum, h[30] / tiles, h[31] / tiles,
                h[33] / tiles, h[34] / tiles, h[32] / tiles);
    }
#endif
    return CV_OK;
}

// ---- stage C image: every pointwise blob carries W_hi | W_lo (bf16) or one fp16 image ------------------------------------------
namespace {
struct COp { int layer, kind; };   // 0 dw | 1 pw own bias | 2 pw_proj cumulative bias | 5 dw21 + pw22 cols 0..95 | 6 pw22 cols 96..191 | 7 pw23 K 0..95 | 8 pw23 K 96..191
const COp kCOps[sc::NOPS] = {{5, 0}, {6, 1}, {7, 0}, {8, 2}, {9, 1}, {10, 0}, {11, 2}, {12, 1}, {13, 0}, {14, 2}, {15, 1}, {16, 0}, {17, 2},
                             {18, 1}, {19, 0}, {20, 2}, {22, 5}, {22, 6}, {23, 7}, {23, 8}};
uint32_t cop_bytes(int op, bool f16) {
    const cv_layer_info* L = cv_layers();
    const COp& d = kCOps[op];
    const cv_layer_info& l = L[d.layer];
    const uint32_t wb = f16 ? 2 : 4;            // bytes per weight: one fp16 image | bf16 hi + lo
    switch (d.kind) {
        case 0: return (uint32_t)(l.cout + l.k * l.k * l.cout) * 4;
        case 1: case 2: return (uint32_t)l.cout * 4 + (uint32_t)l.cin * l.cout * wb;
        case 5: return 1920 + 96 * 4 + 48 * 96 * wb;
        case 6: return 96 * 4 + 48 * 96 * wb;
        case 7: return 48 * 4 + 96 * 48 * wb;
        default: return 96 * 48 * wb;
    }
}
}  // namespace

int weight_replicas() {
#ifdef CV_EXPERIMENTS
    if (const char* e = getenv("CV_W_REP")) { const int r = atoi(e); if (r >= 1 && r <= CV_W_REPLICAS) return r; }
#endif
    return CV_W_REPLICAS;
}

size_t stageC_image_bytes() {               // the larger (bf16) variant
    size_t n = 0;
    for (int op = 0; op < sc::NOPS; ++op) n += (cop_bytes(op, false) + 127) / 128 * 128;
    return n;
}

int build_stageC_image(const float* blob, uint8_t* img, uint32_t* off, uint32_t* bytes, int* f16_flag, cudaStream_t s) {
    const bool f16 = f16_flag != nullptr;
    const cv_layer_info* L = cv_layers();
    size_t o = 0;
    const float* prev_cum = nullptr;
    auto copy_f = [&](float* dst, const float* src, size_t n) { return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s); };
    auto pw = [&](const cv_layer_info& l, int k0, int K, int n0, int n, uint8_t* dst) {
        prep_pw_part_kernel<<<(K * n + 255) / 256, 256, 0, s>>>(blob + l.w_offset, k0, K, l.cout, n0, n, 2, reinterpret_cast<uint16_t*>(dst), f16_flag);
    };
    for (int op = 0; op < sc::NOPS; ++op) {
        const COp& d = kCOps[op];
        const cv_layer_info& l = L[d.layer];
        off[op] = (uint32_t)o;
        bytes[op] = cop_bytes(op, f16);
        uint8_t* dst = img + o;
        float* f = reinterpret_cast<float*>(dst);
        switch (d.kind) {
            case 0:
                CV_CUDA(copy_f(f, blob + l.b_offset, l.cout));
                CV_CUDA(copy_f(f + l.cout, blob + l.w_offset, (size_t)l.k * l.k * l.cout));
                break;
            case 1:
                CV_CUDA(copy_f(f, blob + l.b_offset, l.cout));
                pw(l, 0, l.cin, 0, l.cout, dst + (size_t)l.cout * 4);
                break;
            case 2:
                add_f32_kernel<<<1, 64, 0, s>>>(f, l.skip >= 0 ? prev_cum : nullptr, blob + l.b_offset, l.cout);
                prev_cum = f;
                pw(l, 0, l.cin, 0, l.cout, dst + (size_t)l.cout * 4);
                break;
            case 5: {
                const cv_layer_info& l21 = L[21];
                CV_CUDA(copy_f(f, blob + l21.b_offset, 48));
                CV_CUDA(copy_f(f + 48, blob + l21.w_offset, 9 * 48));
                CV_CUDA(copy_f(f + 480, blob + l.b_offset, 96));
                pw(l, 0, 48, 0, 96, dst + 1920 + 384);
                break;
            }
            case 6:
                CV_CUDA(copy_f(f, blob + l.b_offset + 96, 96));
                pw(l, 0, 48, 96, 96, dst + 384);
                break;
            case 7:
                add_f32_kernel<<<1, 64, 0, s>>>(f, prev_cum, blob + l.b_offset, 48);
                pw(l, 0, 96, 0, 48, dst + 192);
                break;
            default:
                pw(l, 96, 96, 0, 48, dst);
                break;
        }
        CV_CHECK_LAUNCH();
        o += (bytes[op] + 127) / 128 * 128;
    }
    return CV_OK;
}

int launch_permute_p2(const bf16* in, bf16* out, int64_t n_crops, int C, cudaStream_t s) {
    const int64_t n_chunks = n_crops * 64 * (C / 8);
    if (n_chunks == 0) return CV_OK;
    permute_p2_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n_chunks);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_stageC(const bf16* x_p2, int64_t n_crops, const uint8_t* wimg, const uint32_t* off, const uint32_t* bytes, bf16* y_p8, int num_sms,
                  const StageGate& gate, int perm_boards, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if (n_crops % 16 != 0) { cv_set_error("stage C: crop count %lld is not a multiple of 16", (long long)n_crops); return CV_ERR_ARG; }
    StageCParams p{};
    p.w_rep = weight_replicas(); p.w_stride = (uint32_t)stageC_image_bytes();
    p.x = x_p2; p.wimg = wimg; p.y = y_p8; p.n_tiles = (int)(n_crops / 16); p.perm_boards = perm_boards;
    p.gate = gate.flag; p.gate_want = gate.want; p.ovf = gate.ovf;
#ifdef CV_EXPERIMENTS                 // ablation / timing switches change the results: compiled in only with -DCV_EXPERIMENTS (CV_NVCC_EXTRA)
    { const char* d = getenv("CV_SC_DEBUG"); p.debug = d ? atoi(d) : 0; }
#endif
    for (int i = 0; i < sc::NOPS; ++i) { p.off[i] = off[i]; p.bytes[i] = bytes[i]; }
    auto kern = gate.f16 ? stageC_kernel<true> : stageC_kernel<false>;
    CV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sc::SMEM));
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    kern<<<grid, NT, sc::SMEM, s>>>(p);
    CV_CHECK_LAUNCH();
#ifdef CV_SC_PROFILE
    if (p.debug & 256) {
        unsigned long long h[40];
        CV_CUDA(cudaStreamSynchronize(s));
        CV_CUDA(cudaMemcpyFromSymbol(h, g_sc_prof, sizeof(h)));
        const double tiles = (double)((p.n_tiles + grid - 1) / grid);
        fprintf(stderr, "stageC cycles per 16-crop tile: wait in/w %.0f | dw_start %.0f | pw6 mma %.0f | pw6 epi %.0f | dw_mid s2 %.0f | tile turnaround %.0f\n   4x4 ops 3..19:",
                h[0] / tiles, h[1] / tiles, h[2] / tiles, h[3] / tiles, h[4] / tiles, h[39] / tiles);
        double sum = 0;
        for (int op = 3; op < 20; ++op) { fprintf(stderr, " %.0f", h[8 + op] / tiles); sum += h[8 + op] / tiles; }
        fprintf(stderr, "  (sum %.0f)\n", sum);
        fprintf(stderr, "   inside the four pw_exp ops (sums; the op figures above hold only their trailing barrier): weights %.0f | fence+bar %.0f | issue %.0f | wait m0 %.0f | epi m0 %.0f | wait m1 %.0f | epi m1 %.0f"
                        "   dw3x3: weights %.0f | work %.0f\n   weight waits per tile: %.1f of %.1f not ready at first poll, %.0f cycles in all\n", h[28] / tiles, h[29] / tiles, h[30] / tiles, h[31] / tiles, h[32] / tiles, h[33] / tiles, h[34] / tiles, h[35] / tiles, h[36] / tiles, h[6] / tiles, h[7] / tiles, h[5] / tiles);
    }
#endif
    return CV_OK;
}

// ---- stage B ----------------------------------------------------------------------------------------------------------------------
size_t stageB_image_bytes() { return sb::w_bytes(false); }          // the larger (bf16 hi|lo) variant

int build_stageB_image(const float* blob, uint8_t* img, int* f16_flag, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    const int wp = f16_flag ? 1 : 2;
    const int w2 = 384, w3 = w2 + 512 * wp, w4 = w3 + 13824 * wp;     // == the offsets stageB_kernel<F16> derives
    float* f = reinterpret_cast<float*>(img);
    CV_CUDA(cudaMemcpyAsync(f, blob + L[2].b_offset, 16 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CV_CUDA(cudaMemcpyAsync(f + 16, blob + L[3].b_offset, 48 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CV_CUDA(cudaMemcpyAsync(f + 64, blob + L[4].b_offset, 32 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    prep_pw_part_kernel<<<1, 256, 0, s>>>(blob + L[2].w_offset, 0, 16, 16, 0, 16, 2, reinterpret_cast<uint16_t*>(img + w2), f16_flag);
    prep_pw_part_kernel<<<(144 * 48 + 255) / 256, 256, 0, s>>>(blob + L[3].w_offset, 0, 144, 48, 0, 48, 2, reinterpret_cast<uint16_t*>(img + w3), f16_flag);
    prep_pw_part_kernel<<<(48 * 32 + 255) / 256, 256, 0, s>>>(blob + L[4].w_offset, 0, 48, 32, 0, 32, 2, reinterpret_cast<uint16_t*>(img + w4), f16_flag);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_stageB(const bf16* x, int64_t n_crops, const uint8_t* wimg, bf16* y_p2, int num_sms, const StageGate& gate, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if (n_crops % 2 != 0) { cv_set_error("stage B: odd crop count %lld", (long long)n_crops); return CV_ERR_ARG; }
    StageBParams p{x, wimg, y_p2, (int)(n_crops / 2), gate.flag, gate.want, gate.ovf};
    auto kern = gate.f16 ? stageB_kernel<true> : stageB_kernel<false>;
    CV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sb::SMEM));
    const int grid = p.n_tiles < 2 * num_sms ? p.n_tiles : 2 * num_sms;
    kern<<<grid, sb::NTB, sb::SMEM, s>>>(p);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
