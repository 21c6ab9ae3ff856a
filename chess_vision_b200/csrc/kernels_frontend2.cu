// Fused front end, second generation (uint8 HWC boards): crop gather + normalise + bilinear resize + conv_stem + blocks.0.0
// in one persistent kernel (ChessSquareCNN._crop_squares, models/square.py:43-74, + the first two convs of the trunk, :86).
//
// What changed against kernels_frontend.cu (which stays as the path for float / CHW sources):
//   * the board window of a crop (crop x crop pixels, clamped to the board) is staged in shared memory by TMA bulk copies
//     (one per window row, issued by a dedicated warp, double buffered) -- each board byte is fetched from HBM/L2 exactly
//     once per crop and no thread ever waits on a global load;
//   * the bilinear resize is separable: a horizontal pass (window rows -> fp16 rows of 64 normalised pixels) and a vertical
//     pass that writes the space-to-depth bf16 operand image of the stem conv directly;
//   * 8 epilogue warps (two groups that alternate over the accumulator tiles); W_hi | W_lo stay concatenated along N (one
//     MMA and ONE shared-memory read of the A operand per tap -- small-N MMAs are bound by the A fetch, not by the math),
//     the epilogue adds the two column halves; biases live in registers;
//   * activation images use the minimal halo pitch (33 / 17 positions per row instead of 40 / 24): 9 instead of 10 stem
//     tiles per crop and 35 % less shared memory, which pays for a double-buffered stem operand;
//   * blocks.0.0 of crop n is issued after the first stem tiles of crop n+1, so the tensor core never waits for the
//     epilogue to finish a crop.
//
// Warp roles (17 warps): 0-7 epilogue (group = warp>>2, TMEM lane quadrant = warp&3), 8-15 resize producers (warp 8 also
// issues the TMA row copies of the next crop), 16 MMA issuer (+ TMEM owner).
#include "internal.h"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int XP = 33, XLEAD = 8, XPOS = XLEAD + 33 * XP;      // 1097 positions per channel-chunk plane
constexpr int X_CHUNK = XPOS * 16;                             // 17552
constexpr int X_BYTES = 2 * X_CHUNK;                           // 35104
constexpr int YP = 17, YLEAD = 8, YPOS = YLEAD + 17 * YP;      // 297
constexpr int Y_CHUNK = YPOS * 16;                             // 4752
constexpr int Y_PLANE = 4 * Y_CHUNK;                           // 19008 (= 64 mod 128: the two x-parities hit different banks)
constexpr int Y_BYTES = 4 * Y_PLANE;                           // 76032
constexpr int W_STEM_BYTES = 4 * 2048;                         // taps x [2 chunks][64 n = hi 32 | lo 32][8]
constexpr int W_B00_BYTES = 18 * 1024;                         // (tap, k-step) x [2 chunks][32 n = hi 16 | lo 16][8]
constexpr int W_BYTES = W_STEM_BYTES + W_B00_BYTES;            // 26624
constexpr int HB_PITCH = 64 * 3 * 2;                           // fp16 row of 64 normalised pixels
constexpr int ND = 4;                                          // stem accumulator ring depth (64 columns each: hi | lo)
constexpr int NTHREADS = 17 * 32;
constexpr int NPROD = 256;
constexpr int STEM_TILES = 9, B00_TILES = 3;

struct FrontTables {                  // per launch, copied to shared memory
    int16_t yo0[8][64], yo1[8][64];   // window-relative source rows of output row d for square row r
    int16_t xo0[8][64], xo1[8][64];   // byte offsets inside a staged window row of output column d for square column c
    float lam[64];
    int32_t row0[8], nrows[8];        // first board row and number of rows staged for square row r
    int32_t byte0[8], nbytes[8];      // first byte (16-aligned) and byte count (multiple of 16) of a staged row for square column c
};

struct Front2Params {
    const uint8_t* boards;            // (B, H, H, 3) uint8
    const bf16* wimg;
    const float* bias_stem;
    const float* bias_b00;
    bf16* y;                          // T8 [crops*256 rows][16 ch]
    int n_crops, H;
    int raw_pitch, raw_bytes, hb_bytes, n_xbuf, n_rawbuf;
    int off_raw, off_hb, off_y, off_w, off_tab, off_bar, smem_total;
    float na[3], nb[3];               // normalisation v = na[c] * u8 + nb[c]
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(NTHREADS, 1)
frontend2_kernel(const __grid_constant__ Front2Params p, const __grid_constant__ FrontTables tab_param) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* X = smem;                                            // n_xbuf x X_BYTES
    uint8_t* RAW = smem + p.off_raw;                              // n_rawbuf x raw_bytes
    uint8_t* HB = smem + p.off_hb;
    uint8_t* Y = smem + p.off_y;
    uint8_t* W = smem + p.off_w;
    const FrontTables& tab = *reinterpret_cast<const FrontTables*>(smem + p.off_tab);
    float* bias = reinterpret_cast<float*>(smem + p.off_tab + sizeof(FrontTables));      // [0,32) stem, [32,48) blocks.0.0
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t *raw_full = bars, *raw_empty = bars + 2, *x_full = bars + 4, *x_empty = bars + 6, *y_full = bars + 8, *y_empty = bars + 9,
             *e_full = bars + 10, *e_empty = bars + 11, *wbar = bars + 12, *d_full = bars + 13, *d_empty = bars + 13 + ND;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13 + 2 * ND);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: zero X / Y (halos stay zero for the whole kernel), tables, barriers, TMEM
    for (int i = threadIdx.x; i < (p.n_xbuf * X_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4*>(X)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (Y_BYTES + 2048) / 16; i += NTHREADS) reinterpret_cast<uint4*>(Y)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (int)(sizeof(FrontTables) / 4); i += NTHREADS)
        reinterpret_cast<uint32_t*>(smem + p.off_tab)[i] = reinterpret_cast<const uint32_t*>(&tab_param)[i];
    for (int i = threadIdx.x; i < 48; i += NTHREADS) bias[i] = i < 32 ? p.bias_stem[i] : p.bias_b00[i - 32];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 1); mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
        mbar_init(y_full, 256); mbar_init(y_empty, 1);
        mbar_init(e_full, 1); mbar_init(e_empty, 8);
        mbar_init(wbar, 1);
        for (int i = 0; i < ND; ++i) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, 8); }
        fence_barrier_init();
    }
    if (warp == 16) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_e = tmem_base + ND * 64;                  // blocks.0.0 accumulators: 3 tiles x 32 columns (hi | lo)
    const int rawmask = p.n_rawbuf - 1, xmask = p.n_xbuf - 1;

    if (warp >= 8 && warp < 16) {
        // =========================== resize producers (256 threads) ==========================================================
        const int t = threadIdx.x - 256;
        const float a0 = p.na[0], a1 = p.na[1], a2 = p.na[2], b0 = p.nb[0], b1 = p.nb[1], b2 = p.nb[2];
        // stage the window rows of crop `nn` (iteration index iti) into RAW slot iti & rawmask: executed by warp 8 only.  The slot
        // is free: its previous user is the horizontal pass of an earlier crop, which ended at a producer barrier.
        auto stage_window = [&](int nn, uint32_t iti) {
            const int slot = iti & rawmask;
            const int64_t b = nn >> 6;
            const int rr = (nn >> 3) & 7, cc = nn & 7;
            const int nr = tab.nrows[rr], nbytes = tab.nbytes[cc];
            if (lane == 0) mbar_arrive_expect_tx(raw_full + slot, (uint32_t)(nr * nbytes));
            __syncwarp();
            const uint8_t* src = p.boards + (b * p.H + tab.row0[rr]) * (int64_t)p.H * 3 + tab.byte0[cc];
            uint8_t* dst = RAW + slot * p.raw_bytes;
            for (int i = lane; i < nr; i += 32) bulk_g2s(dst + i * p.raw_pitch, src + (int64_t)i * p.H * 3, (uint32_t)nbytes, raw_full + slot);
        };
        if (warp == 8 && blockIdx.x < p.n_crops) stage_window(blockIdx.x, 0);
        uint32_t it = 0;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            const int r = (n >> 3) & 7, c = n & 7;
            const int rslot = it & rawmask, xslot = it & xmask;
            const uint8_t* raw = RAW + rslot * p.raw_bytes;
            if (warp == 8 && p.n_rawbuf == 2 && n + (int)gridDim.x < p.n_crops) stage_window(n + gridDim.x, it + 1);   // prefetch
            mbar_wait(raw_full + rslot, (it / p.n_rawbuf) & 1u);
            // ---- horizontal pass: window row i, output columns 2xp, 2xp+1 -> 6 fp16 (normalised) at HB[i][xp]
            const int nrows = tab.nrows[r];
            const int xp = t & 31;                               // fixed per thread: its 4 column taps are looked up once per crop
            const int xoff[4] = {tab.xo0[c][2 * xp], tab.xo1[c][2 * xp], tab.xo0[c][2 * xp + 1], tab.xo1[c][2 * xp + 1]};
            const float lxs[2] = {tab.lam[2 * xp], tab.lam[2 * xp + 1]};
            for (int i = t >> 5; i < nrows; i += NPROD / 32) {
                const uint8_t* rowp = raw + i * p.raw_pitch;
                float v[6];
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const uint8_t* q0 = rowp + xoff[2 * dx];
                    const uint8_t* q1 = rowp + xoff[2 * dx + 1];
                    const float lx = lxs[dx];
                    const float u00 = (float)q0[0], u01 = (float)q0[1], u02 = (float)q0[2];
                    const float u10 = (float)q1[0], u11 = (float)q1[1], u12 = (float)q1[2];
                    v[dx * 3 + 0] = fmaf(a0, fmaf(lx, u10 - u00, u00), b0);
                    v[dx * 3 + 1] = fmaf(a1, fmaf(lx, u11 - u01, u01), b1);
                    v[dx * 3 + 2] = fmaf(a2, fmaf(lx, u12 - u02, u02), b2);
                }
                uint32_t* o = reinterpret_cast<uint32_t*>(HB + i * HB_PITCH + xp * 12);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                    o[k] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");                       // HB complete, RAW slot consumed
            if (warp == 8 && p.n_rawbuf == 1 && n + (int)gridDim.x < p.n_crops) stage_window(n + gridDim.x, it + 1);   // single buffer: refill now
            mbar_wait(x_empty + xslot, ((it / p.n_xbuf) & 1u) ^ 1u);
            // ---- vertical pass: s2d position (py, px) -> 12 bf16 channels (dy*6 + dx*3 + c) into the stem operand image
            uint8_t* Xb = X + xslot * X_BYTES;
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int sp = t + NPROD * k, py = sp >> 5, px = sp & 31;
                float o[12];
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const int y = 2 * py + dy;
                    const uint32_t* h0 = reinterpret_cast<const uint32_t*>(HB + tab.yo0[r][y] * HB_PITCH + px * 12);
                    const uint32_t* h1 = reinterpret_cast<const uint32_t*>(HB + tab.yo1[r][y] * HB_PITCH + px * 12);
                    const float ly = tab.lam[y];
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        uint32_t w0 = h0[w], w1 = h1[w];
                        const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&w0));
                        const float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&w1));
                        o[dy * 6 + 2 * w] = fmaf(ly, f1.x - f0.x, f0.x);
                        o[dy * 6 + 2 * w + 1] = fmaf(ly, f1.y - f0.y, f0.y);
                    }
                }
                const int pos = XLEAD + (py + 1) * XP + px;
                *reinterpret_cast<uint4*>(Xb + pos * 16) = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
                *reinterpret_cast<uint4*>(Xb + X_CHUNK + pos * 16) = make_uint4(pack2(o[8], o[9]), pack2(o[10], o[11]), 0u, 0u);
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");                       // operand image complete; HB free again
            if (t == 0) mbar_arrive(x_full + xslot);
        }
    } else if (warp == 16) {
        // =========================== MMA issuer ===========================================================================
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, W_BYTES);
            bulk_g2s(W, p.wimg, W_BYTES, wbar);
        }
        mbar_wait(wbar, 0);
        const uint32_t idesc_s = make_idesc_bf16(128, 64), idesc_1 = make_idesc_bf16(128, 32);
        const uint32_t yb = smem_u32(Y), ws = smem_u32(W), w1 = ws + W_STEM_BYTES;
        auto issue_b00 = [&](uint32_t itb) {                     // blocks.0.0 of the crop with iteration index itb
            mbar_wait(y_full, itb & 1u);
            mbar_wait(e_empty, (itb & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                for (int j = 0; j < B00_TILES; ++j) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        const int plane = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
                        const int Dy = ky == 0 ? -1 : 0, Dx = kx == 0 ? -1 : 0;
                        const uint32_t a0 = yb + (uint32_t)plane * Y_PLANE + (uint32_t)(YLEAD + YP + 128 * j + Dy * YP + Dx) * 16u;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            mma_bf16_ss(tmem_e + j * 32, make_smem_desc(a0 + ks * 2 * Y_CHUNK, Y_CHUNK, 128),
                                        make_smem_desc(w1 + (tap * 2 + ks) * 1024, 32 * 16, 128), idesc_1, (tap | ks) ? 1u : 0u);
                    }
                }
                mma_commit(e_full);
                mma_commit(y_empty);
            }
            __syncwarp();
        };
        uint32_t it = 0, tcount = 0;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            const int xslot = it & xmask;
            mbar_wait(x_full + xslot, (it / p.n_xbuf) & 1u);
            tc_fence_after();
            const uint32_t xb = smem_u32(X + xslot * X_BYTES);
            for (int j = 0; j < STEM_TILES; ++j, ++tcount) {
                if (j == ND && it > 0) issue_b00(it - 1);        // previous crop's second conv, once this crop's first tiles are in flight
                const uint32_t buf = tcount & (ND - 1), ph = (tcount / ND) & 1u;
                mbar_wait(d_empty + buf, ph ^ 1u);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 4; ++tap) {
                        const int Dy = (tap >> 1) - 1, Dx = (tap & 1) - 1;
                        const uint32_t a = xb + (uint32_t)(XLEAD + XP + 128 * j + Dy * XP + Dx) * 16u;
                        mma_bf16_ss(tmem_base + buf * 64, make_smem_desc(a, X_CHUNK, 128), make_smem_desc(ws + tap * 2048, 64 * 16, 128),
                                    idesc_s, tap > 0 ? 1u : 0u);
                    }
                    mma_commit(d_full + buf);
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(x_empty + xslot);          // operand image free once the stem MMAs have read it
            __syncwarp();
        }
        if (it > 0) issue_b00(it - 1);
    } else {
        // =========================== epilogue warps 0-7: group g = warp>>2 takes output channels [16g, 16g+16) of every stem tile ====
        const int g = warp >> 2, qd = warp & 3, i = qd * 32 + lane;
        const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
        float bs[16];                                            // this group's stem biases in registers (shared-memory bandwidth is the bound here)
#pragma unroll
        for (int k = 0; k < 16; ++k) bs[k] = bias[16 * g + k];
        const float* bb = bias + 32;
        auto epilogue_b00 = [&](int n, uint32_t itb) {           // blocks.0.0 accumulators of crop n -> global T8 tile rows
            mbar_wait(e_full, itb & 1u);
            tc_fence_after();
            const int j0 = g, j1 = g + 2;                        // group 0: tiles 0, 2; group 1: tile 1
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = k == 0 ? j0 : j1;
                uint32_t eh[16], el[16];                         // hi | lo halves of the 16 output channels
                if (j < B00_TILES) {
                    tmem_ld16(tmem_e + lane_base + j * 32, eh);
                    tmem_ld16(tmem_e + lane_base + j * 32 + 16, el);
                    tmem_ld_wait();
                }
                if (k == 1) {                                    // every accumulator this warp reads is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(e_empty);
                }
                if (j >= B00_TILES) continue;
                const int Q = 128 * j + i, Oy = Q / YP, Ox = Q - Oy * YP;
                if (Ox < 16 && Oy < 16) {
                    const int64_t m = (int64_t)n * 256 + Oy * 16 + Ox;
                    uint4* dst = reinterpret_cast<uint4*>(p.y) + ((m >> 7) * 2) * 128 + (m & 127);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            v[q] = fmaxf(__uint_as_float(eh[8 * c + q]) + __uint_as_float(el[8 * c + q]) + bb[8 * c + q], 0.f);
                        dst[c * 128] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                    }
                }
            }
        };
        uint32_t it = 0, tcount = 0;
        int prev_n = -1;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            for (int j = 0; j < STEM_TILES; ++j, ++tcount) {
                if (j == ND && it > 0) epilogue_b00(prev_n, it - 1);
                const uint32_t buf = tcount & (ND - 1), ph = (tcount / ND) & 1u;
                mbar_wait(d_full + buf, ph);
                tc_fence_after();
                uint32_t rh[16], rl[16];                         // hi and lo halves of this group's 16 channels
                tmem_ld16(tmem_base + lane_base + buf * 64 + g * 16, rh);
                tmem_ld16(tmem_base + lane_base + buf * 64 + 32 + g * 16, rl);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty + buf);       // accumulator drained into registers
                if (j == 0 && it > 0) mbar_wait(y_empty, (it - 1) & 1u);   // first Y write of this crop: previous blocks.0.0 MMAs done
                const int q = 128 * j + i, oy = q / XP, ox = q - oy * XP;
                if (ox < 32 && oy < 32) {
                    const int plane = (oy & 1) * 2 + (ox & 1);
                    uint8_t* dst = Y + plane * Y_PLANE + (YLEAD + ((oy >> 1) + 1) * YP + (ox >> 1)) * 16 + 2 * g * Y_CHUNK;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            v[e] = fmaxf(__uint_as_float(rh[cc * 8 + e]) + __uint_as_float(rl[cc * 8 + e]) + bs[cc * 8 + e], 0.f);
                        *reinterpret_cast<uint4*>(dst + cc * Y_CHUNK) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(y_full);
            prev_n = n;
        }
        if (it > 0) epilogue_b00(prev_n, it - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// Weight images: conv_stem as a 2x2 conv on the space-to-depth crop (4 taps x K 16), blocks.0.0 as 9 taps x 2 k-steps;
// W_hi | W_lo concatenated along N.
__global__ void prep_frontend2_weights_kernel(const float* __restrict__ w_stem /*[27][32]*/, const float* __restrict__ w_b00 /*[288][16]*/,
                                              bf16* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W_BYTES / 2) return;
    float v = 0.f;
    bool lo;
    if (i < W_STEM_BYTES / 2) {
        const int kk = i & 7, n = (i >> 3) & 63, chunk = (i >> 9) & 1, tap = i >> 10;
        lo = n >= 32;
        const int k = chunk * 8 + kk;                                  // s2d channel = dy*6 + dx*3 + c
        if (k < 12) {
            const int dy = k / 6, dx = (k % 6) / 3, c = k % 3;
            const int ky = 2 * ((tap >> 1) - 1) + dy + 1, kx = 2 * ((tap & 1) - 1) + dx + 1;
            if (ky >= 0 && ky < 3 && kx >= 0 && kx < 3) v = w_stem[((ky * 3 + kx) * 3 + c) * 32 + (n & 31)];
        }
    } else {
        const int j = i - W_STEM_BYTES / 2;
        const int kk = j & 7, n = (j >> 3) & 31, chunk = (j >> 8) & 1, ks = (j >> 9) & 1, tap = j >> 10;
        lo = n >= 16;
        const int ci = ks * 16 + chunk * 8 + kk;
        v = w_b00[(tap * 32 + ci) * 16 + (n & 15)];
    }
    const bf16 hi = __float2bfloat16_rn(v);
    img[i] = lo ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
}

}  // namespace

size_t frontend2_weight_image_elems() { return W_BYTES / 2; }

int launch_frontend2_prep_weights(const float* blob, bf16* img, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    prep_frontend2_weights_kernel<<<(W_BYTES / 2 + 255) / 256, 256, 0, s>>>(blob + L[0].w_offset, blob + L[1].w_offset, img);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

// Returns CV_OK and sets *supported = 0 when this kernel cannot take the configuration (the caller then uses the
// first-generation kernel): non-affine normalisation table, window too large for shared memory.
int launch_frontend2(const uint8_t* boards_hwc, int nb, int H, const CropGeom& g, const float* lut_host, const bf16* wimg,
                     const float* bias_stem, const float* bias_b00, bf16* y, int num_sms, int* supported, cudaStream_t s) {
    *supported = 0;
    if (nb == 0) { *supported = 1; return CV_OK; }
    Front2Params p{};
    // normalisation must be affine in the byte value: v = na*u + nb (true for ToTensor + Normalize)
    for (int c = 0; c < 3; ++c) {
        const float* l = lut_host + c * 256;
        p.nb[c] = l[0];
        p.na[c] = (l[255] - l[0]) / 255.0f;
        for (int u = 0; u < 256; ++u)
            if (fabsf(p.na[c] * u + p.nb[c] - l[u]) > 1e-5f * (1.0f + fabsf(l[u]))) return CV_OK;
    }
    FrontTables tab{};
    const CropTaps tp = make_taps(g);
    int max_rows = 0, max_bytes = 0;
    for (int r = 0; r < 8; ++r) {
        int lo = g.H, hi = -1;
        for (int d = 0; d < 64; ++d) {
            lo = lo < tp.p0[r][d] ? lo : tp.p0[r][d];
            hi = hi > tp.p1[r][d] ? hi : tp.p1[r][d];
        }
        tab.row0[r] = lo;
        tab.nrows[r] = hi - lo + 1;
        const int b0 = (lo * 3) & ~15, b1 = ((hi + 1) * 3 + 15) & ~15;     // row pitch H*3 is a multiple of 16 (H % 32 == 0): stays inside the row
        tab.byte0[r] = b0;
        tab.nbytes[r] = b1 - b0;
        for (int d = 0; d < 64; ++d) {
            tab.yo0[r][d] = (int16_t)(tp.p0[r][d] - lo);
            tab.yo1[r][d] = (int16_t)(tp.p1[r][d] - lo);
            tab.xo0[r][d] = (int16_t)(tp.p0[r][d] * 3 - b0);
            tab.xo1[r][d] = (int16_t)(tp.p1[r][d] * 3 - b0);
        }
        max_rows = max_rows > tab.nrows[r] ? max_rows : tab.nrows[r];
        max_bytes = max_bytes > tab.nbytes[r] ? max_bytes : tab.nbytes[r];
    }
    for (int d = 0; d < 64; ++d) tab.lam[d] = g.lam[d];
    p.raw_pitch = max_bytes;
    p.raw_bytes = (max_rows * max_bytes + 127) & ~127;
    p.hb_bytes = (max_rows * HB_PITCH + 127) & ~127;
    const int fixed = Y_BYTES + 2048 + W_BYTES + (int)sizeof(FrontTables) + 48 * 4 + 256;
    p.n_xbuf = 2; p.n_rawbuf = 2;
    auto total = [&]() { return p.n_xbuf * X_BYTES + p.n_rawbuf * p.raw_bytes + p.hb_bytes + fixed + 2048; };
    if (total() > 227 * 1024) p.n_rawbuf = 1;
    if (total() > 227 * 1024) p.n_xbuf = 1;
    if (total() > 227 * 1024) return CV_OK;                       // window does not fit: not supported
    int off = p.n_xbuf * X_BYTES;
    off = (off + 127) & ~127; p.off_raw = off; off += p.n_rawbuf * p.raw_bytes;
    p.off_hb = off; off += p.hb_bytes;
    off = (off + 127) & ~127; p.off_y = off; off += Y_BYTES + 2048;            // + slack: the last M-tile of a plane reads past its end
    p.off_w = off; off += W_BYTES;
    p.off_tab = off; off += (int)sizeof(FrontTables) + 48 * 4;
    off = (off + 15) & ~15; p.off_bar = off; off += 256;
    p.smem_total = off;
    if (p.smem_total > 227 * 1024) return CV_OK;
    p.boards = boards_hwc; p.wimg = wimg; p.bias_stem = bias_stem; p.bias_b00 = bias_b00; p.y = y;
    p.n_crops = nb * 64; p.H = H;
    *supported = 1;
    CV_CUDA(cudaFuncSetAttribute(frontend2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_total));
    const int grid = p.n_crops < num_sms ? p.n_crops : num_sms;
    frontend2_kernel<<<grid, NTHREADS, p.smem_total, s>>>(p, tab);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
