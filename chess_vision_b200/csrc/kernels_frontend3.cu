// Fused front end, third generation (uint8 HWC boards): crop gather + normalise + bilinear resize + conv_stem + blocks.0.0
// in one persistent kernel (ChessSquareCNN._crop_squares, models/square.py:43-74, + the first two convs of the trunk, :86).
//
// Built on what the instrumented second generation and tools/ubench_umma.cu showed on B200:
//   * a tcgen05.mma with M=128, K=16 costs ~48 cycles whatever N <= 64 is (the shared-memory fetch of the 4 KB A slab), so
//     the tensor-pipe time of this kernel is (number of MMAs) x 48: the M tiles are cut as COLUMN SLABS of the activation
//     image (8 pixels wide, 16 rows high: sixteen 8-row core matrices one image row apart, SBO = row pitch), which tiles the
//     32x32 stem output in exactly 8 and the 16x16 blocks.0.0 output in exactly 2 tiles (9 and 3 with linear tiles);
//   * the stem conv consumes the resized pixels as fp16 (they are bounded: |v| < 2.7) against fp16 weights: 2^-12 weight
//     rounding instead of the bf16 W_hi + W_lo pair, so N = 32 instead of 64 -- half the TMEM traffic, no hi+lo adds in the
//     epilogue -- and the vertical resize pass is pure half2 arithmetic writing the operand image directly; its folded-BN
//     bias rides on a constant-one input channel; ReLU is folded into the bf16x2 conversion (cvt.rn.relu);
//   * second generation serialised "epilogue of crop n -> blocks.0.0 MMAs of crop n -> epilogue of crop n+1" on the single
//     stem-output image.  Here that image is split into LEFT / RIGHT halves (stem columns 0..15 / 16..31, each the input of
//     one blocks.0.0 tile; column 15 is duplicated as the halo of the right half) rotating through three buffers, and the
//     MMA warp interleaves  stem(n) left | b00(n-1) right | stem(n) right | b00(n) left.  The tensor pipe executes in issue
//     order, so a stem tile's completion implies every earlier blocks.0.0 tile has finished reading its buffer: no
//     "buffer empty" barriers are needed at all;
//   * the issuing thread adds a constant to a precomputed descriptor word per MMA (its dependent-instruction latency is exposed).
//
// Warp roles (22 warps): 0-7 epilogue (group g = warp>>2 takes the stem tiles of image rows [16g, 16g+16) and blocks.0.0
// tile g; TMEM lane quadrant = warp&3), 8-19 resize producers, 20 MMA issuer (+ TMEM owner), 21 issues the TMA tile copies of the
// board windows (on a producer warp that issue cost the whole producer group ~1 k cycles per crop at its barrier).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include "internal.h"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int XP = 33, XLEAD = 8, XPOS = XLEAD + 33 * XP;      // stem operand image: 1097 positions per channel-chunk plane
constexpr int X_CHUNK = XPOS * 16;                             // 17552
constexpr int X_BYTES = 2 * X_CHUNK;                           // 35104
constexpr int X_ALLOC = (X_BYTES + 127) & ~127;
constexpr int YHP = 9, YLEAD = 8, YHPOS = YLEAD + 17 * YHP;    // half image of one parity plane: 17 rows x (halo + 8) positions
constexpr int YH_CHUNK = YHPOS * 16;                           // 2576
constexpr int YH_PLANE = 4 * YH_CHUNK;                         // 10304 (= 64 mod 128: the two x-parities hit different banks)
constexpr int YH_BYTES = 4 * YH_PLANE;                         // 41216 per half
constexpr int NYBUF = 2;                                       // left / right half images of the stem output
constexpr int NXBUF = 2;                                       // stem operand image, double buffered
constexpr int W_STEM_BYTES = 4 * 1024;                         // taps x [2 chunks][32 n][8] fp16
constexpr int W_B00_BYTES = 18 * 1024;                         // (tap, k-step) x [2 chunks][32 n = hi 16 | lo 16][8] bf16
constexpr int W_B00_F16_BYTES = 18 * 512;                      // (tap, k-step) x [2 chunks][16 n][8] fp16 (fp16 mode: one image, N = 16)
constexpr int W_BYTES = W_STEM_BYTES + W_B00_BYTES;            // 22528: what a CTA holds in shared memory (stem + one blocks.0.0 variant)
constexpr int W_IMG_BYTES = W_BYTES + W_B00_F16_BYTES;         // global image: stem | blocks.0.0 bf16 hi|lo | blocks.0.0 fp16
constexpr int NPROD = 384;                                      // 12 resize producer warps (the resize is the longest per-crop job)
constexpr int MMA_WARP = 8 + NPROD / 32;
constexpr int TMA_WARP = MMA_WARP + 1;                         // stages the board windows: the tile copy's issue latency stays off the producers' chain
constexpr int NTHREADS = (TMA_WARP + 1) * 32;
constexpr int TM_B00 = 256;                                    // TMEM: stem tile k at column 32 k, blocks.0.0 tile t at 256 + 32 t

struct Front3Tables {                 // per launch, copied to shared memory
    uint32_t vy[8][64];               // vertical taps of output row d for square row r: row0 | row1 << 8 | fp16(lambda) << 16
    uint8_t xh0[8][64], xh1[8][64];   // horizontal taps of output column d for square column c: pixel index in the staged window
    uint16_t lamh[64];                // fp16(lambda) per output column
    int32_t row0[8], nrows[8];        // first board row and number of rows staged for square row r
    int32_t byte0[8], boff[8];        // square column c: first staged byte of a board row (16-byte aligned, as the TMA tile needs) and the offset of the
                                      // first pixel GROUP inside the staged row (pixel index a multiple of 4 -> byte offset a multiple of 4: word-aligned groups)
};

struct Front3Params {
    int perm_boards;                                              // > 0: output position n holds the crop perm_inv(n, perm_boards) names (umma.cuh)
    const uint8_t* boards;            // (B, H, H, 3) uint8
    const uint8_t* wimg;
    const float* bias_b00;
    bf16* y;                          // T8 [crops*256 rows][16 ch]
    int n_crops, H;
    int raw_pitch, raw_bytes, v_pitch, v_bytes, n_groups, g_magic, n_rawbuf, n_xbuf, box_bytes;
    int off_raw, off_v, off_y, off_w, off_tab, off_bar, smem_total;
    float na[3], nb[3];               // normalisation v = na[c] * u8 + nb[c]
    const int* skip_flag;             // non-null: the kernel exits at once when *skip_flag != 0 (float source that is not a uint8 image, see below)
    const int* gate;                  // non-null: run only when (*gate != 0) == gate_want (fp16 pass / bf16 fall-back pass, internal.h StageGate)
    int gate_want;
    int debug;                        // experiment builds (CV_FE3_DEBUG): 1 / 2 skip the resize passes, 4 / 8 skip epilogue TMEM loads / stores, 16 / 128 / 1024 /
                                      // 2048 skip MMA-thread waits, 32 skip the x_full wait, 64 spin in the epilogue, 256 wait-cycle report, 512 event trace
};

// Timing experiments (-DCV_FE_PROFILE, CV_FE3_DEBUG & 256): cycles the lead lane of each role spends in each barrier wait.
#if defined(CV_EXPERIMENTS) || defined(CV_FE_PROFILE)
#define FE_DBG(bits) ((p.debug & (bits)) != 0)        // ablation switches exist in experiment builds only
#else
#define FE_DBG(bits) (0)
#endif
#ifdef CV_FE_PROFILE
__device__ uint2 g_fe3_trace[4 * 1024];            // CV_FE3_DEBUG & 512: event trace (tag, clock) of CTA 0, crops 8..15: one lane per role
#define TRACE(tag) do { if (trace_on && tr_n < 1024) { g_fe3_trace[tr_role * 1024 + tr_n] = make_uint2((uint32_t)(tag), (uint32_t)clock64()); ++tr_n; } } while (0)
__device__ unsigned long long g_fe3_prof[4 * 16];
#define TWAIT(k, call)                                                   \
    do {                                                                 \
        const long long _t = clock64();                                  \
        call;                                                            \
        if (prof_on) pacc[k] += clock64() - _t;                          \
    } while (0)
#else
#define TWAIT(k, call) call
#define TRACE(tag)
#endif

template <bool F16>                   // F16: stem output image, blocks.0.0 weights and the T8 output are fp16; otherwise bf16 (W_hi | W_lo along N)
__global__ void __launch_bounds__(NTHREADS, 1)
frontend3_kernel(const __grid_constant__ Front3Params p, const __grid_constant__ Front3Tables tab_param, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* X = smem;
    uint8_t* RAW = smem + p.off_raw;                              // n_rawbuf x raw_bytes
    uint8_t* V = smem + p.off_v;                                  // 64 rows x window pixels x [R G B -] fp16, vertically resized + normalised
    uint8_t* Y = smem + p.off_y;                                  // NYBUF half images
    uint8_t* W = smem + p.off_w;
    const Front3Tables& tab = *reinterpret_cast<const Front3Tables*>(smem + p.off_tab);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t *raw_full = bars + 32 /*3*/, *x_full = bars + 2, *x_empty = bars + 4, *wbar = bars + 6, *yh_full = bars + 7 /*2*/, *e_full = bars + 9 /*2*/,
             *e_empty = bars + 11 /*2*/, *d_full = bars + 13 /*8*/, *d_empty = bars + 21 /*8*/;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
    uint64_t* raw_empty = bars + 35 /*3*/;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.skip_flag != nullptr && *p.skip_flag != 0) return;     // uniform over the grid: decided by an earlier kernel of the stream
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_want != 0)) return;

    // ---- one-time setup: zero X / Y (halos stay zero for the whole kernel), tables, barriers, TMEM
    for (int i = threadIdx.x; i < p.n_xbuf * X_ALLOC / 16; i += NTHREADS) reinterpret_cast<uint4*>(X)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (NYBUF * YH_BYTES + 256) / 16; i += NTHREADS) reinterpret_cast<uint4*>(Y)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (int)(sizeof(Front3Tables) / 4); i += NTHREADS)
        reinterpret_cast<uint32_t*>(smem + p.off_tab)[i] = reinterpret_cast<const uint32_t*>(&tab_param)[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 1); }
        for (int i = 0; i < NXBUF; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
        mbar_init(wbar, 1);
        for (int i = 0; i < NYBUF; ++i) mbar_init(yh_full + i, 8);
        for (int i = 0; i < 2; ++i) { mbar_init(e_full + i, 1); mbar_init(e_empty + i, 4); }
        for (int i = 0; i < 8; ++i) mbar_init(d_full + i, 1);
        mbar_init(d_empty, 8);                                   // one arrival per epilogue warp and crop: all its stem accumulators are drained
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int xmask = p.n_xbuf - 1, xshift = p.n_xbuf - 1;   // 1 or 2 operand images: slot = it & mask, use = it >> shift
#ifdef CV_FE_PROFILE
    const int pw = FE_DBG(512) ? 12 : 8;                     // which producer warp is sampled
    const bool prof_on = FE_DBG(256) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == pw || warp == MMA_WARP);
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_role0 = clock64();
    const int tr_role = warp == 0 ? 0 : warp == 4 ? 1 : warp == pw ? 2 : 3;
    bool trace_on = false;
    int tr_n = 0;
#define TRACE_WINDOW(it_) trace_on = FE_DBG(512) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == pw || warp == MMA_WARP) && (it_) >= 8 && (it_) < 16
#else
#define TRACE_WINDOW(it_)
#endif

    if (warp >= 8 && warp < MMA_WARP) {
        // =========================== resize producers (256 threads) ==========================================================
        const int t = threadIdx.x - 256;
        // normalisation v = na[c] * u8 + nb[c] as half2 constants for the channel pairs a pixel group runs through: (R,G) (B,R) (G,B)
        const __half2 na2[3] = {__floats2half2_rn(p.na[0], p.na[1]), __floats2half2_rn(p.na[2], p.na[0]), __floats2half2_rn(p.na[1], p.na[2])};
        const __half2 nb2[3] = {__floats2half2_rn(p.nb[0], p.nb[1]), __floats2half2_rn(p.nb[2], p.nb[0]), __floats2half2_rn(p.nb[1], p.nb[2])};
        const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
        uint32_t it = 0, rphase = 0;
        int rslot = 0;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            int sq = n & 63;                                             // square of the crop that output position n holds
            if (p.perm_boards > 0) { int64_t pb; perm_inv(n, p.perm_boards, pb, sq); }
            const int r = sq >> 3, c = sq & 7;
            const uint8_t* raw = RAW + rslot * p.raw_bytes;
            TRACE_WINDOW(it);
            TWAIT(0, mbar_wait(raw_full + rslot, rphase));
            TRACE(0xB00);
            // ---- vertical pass first, on the raw bytes: item = (output row y, group of 4 window pixels = 12 bytes = 3 aligned words).
            //      Bytes become fp16 pairs by PRMT (0x6400 | b = 1024 + b exactly), the lerp u0 + ly (u1 - u0) and the normalisation
            //      a u + b run in half2 arithmetic; the 4 pixels are stored as [R G B -] (8 bytes each): V[y][pixel].
#ifdef CV_FE_PROFILE
            const long long t_h0 = clock64();
#endif
            {
                const int n_items = 64 * p.n_groups, boff = tab.boff[c];
                for (int item = t; item < (FE_DBG(1) ? 0 : n_items); item += NPROD) {
                    const int y = __umulhi((unsigned)item, (unsigned)p.g_magic), gq = item - y * p.n_groups;
                    const uint32_t vy = tab.vy[r][y];
                    const uint32_t* r0 = reinterpret_cast<const uint32_t*>(raw + (vy & 255u) * p.raw_pitch + boff + gq * 12);
                    const uint32_t* r1 = reinterpret_cast<const uint32_t*>(raw + ((vy >> 8) & 255u) * p.raw_pitch + boff + gq * 12);
                    const uint32_t l2 = (vy >> 16) * 0x10001u;
                    const __half2 ly = *reinterpret_cast<const __half2*>(&l2);
                    uint32_t n[6];
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const uint32_t a = r0[w], b = r1[w];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const uint32_t ma = __byte_perm(a, 0x64646464u, hh ? 0x4342u : 0x4140u), mb = __byte_perm(b, 0x64646464u, hh ? 0x4342u : 0x4140u);
                            const __half2 m0 = *reinterpret_cast<const __half2*>(&ma), m1 = *reinterpret_cast<const __half2*>(&mb);
                            const __half2 u = __hfma2(ly, __hsub2(m1, m0), __hsub2(m0, k1024));     // both differences are exact
                            const int k = (2 * w + hh) % 3;                                        // channel pair (R,G) / (B,R) / (G,B)
                            const __half2 v = __hfma2(na2[k], u, nb2[k]);
                            n[2 * w + hh] = *reinterpret_cast<const uint32_t*>(&v);
                        }
                    }
                    // V is dense (offset = item * 32), so consecutive lanes own consecutive 32-byte slots: lanes 0-3 of a quarter warp store
                    // their first 16 bytes while lanes 4-7 store their second (and vice versa) -> every STS.128 covers all 32 banks once
                    uint4* dst = reinterpret_cast<uint4*>(V + y * p.v_pitch + gq * 32);
                    const uint4 h0 = make_uint4(n[0], n[1] & 0xffffu, __byte_perm(n[1], n[2], 0x5432u), n[2] >> 16);
                    const uint4 h1 = make_uint4(n[3], n[4] & 0xffffu, __byte_perm(n[4], n[5], 0x5432u), n[5] >> 16);
                    const bool sw = (t & 4) != 0;
                    dst[sw ? 1 : 0] = sw ? h1 : h0;
                    dst[sw ? 0 : 1] = sw ? h0 : h1;
                }
            }
#ifdef CV_FE_PROFILE
            if (prof_on) pacc[4] += clock64() - t_h0;
#endif
            TRACE(0xB10);
            TWAIT(2, asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory"));                       // V complete, RAW slot consumed
            if (t == 0) mbar_arrive(raw_empty + rslot);                                             // the TMA warp may refill it
            if (++rslot == p.n_rawbuf) { rslot = 0; rphase ^= 1u; }                                 // 1 to 3 window buffers, used round robin
            const int xslot = it & xmask;
            uint8_t* Xb = X + xslot * X_ALLOC;
            TWAIT(1, mbar_wait(x_empty + xslot, ((it >> xshift) & 1u) ^ 1u));                   // stem MMAs of the crop before last have read this image
            TRACE(0xB20);
#ifdef CV_FE_PROFILE
            const long long t_v0 = clock64();
#endif
            // ---- horizontal pass (half2): s2d position (py, px) -> 12 fp16 channels (dy*6 + dx*3 + c) + the constant-one channel
            {
                const int px = t & 31;                               // fixed per thread: its column taps are looked up once per crop
                // the four taps of output columns 2 px, 2 px + 1 lie in three consecutive window pixels (checked at launch): three
                // 8-byte loads per row instead of four, the taps are selected in registers
                __half2 lx[2];
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const uint32_t l2 = tab.lamh[2 * px + dx] * 0x10001u;
                    lx[dx] = *reinterpret_cast<const __half2*>(&l2);
                }
                const int xbase = tab.xh0[c][2 * px], xlast = p.n_groups * 4 - 1;
                const int x0 = xbase * 8, x1 = min(xbase + 1, xlast) * 8, x2 = min(xbase + 2, xlast) * 8;
                const bool sb0 = tab.xh1[c][2 * px] != xbase, sa1 = tab.xh0[c][2 * px + 1] != xbase;
                const int sb1 = tab.xh1[c][2 * px + 1] - xbase;
#pragma unroll
                for (int k = 0; k < (1024 + NPROD - 1) / NPROD; ++k) {
                    const int py = (t >> 5) + (NPROD / 32) * k;
                    if (py >= 32 || FE_DBG(2)) break;
                    uint32_t o[6];
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy) {
                        const uint8_t* rowb = V + (2 * py + dy) * p.v_pitch;
                        uint32_t rg[2], bx[2];
                        const uint2 q0 = *reinterpret_cast<const uint2*>(rowb + x0), q1 = *reinterpret_cast<const uint2*>(rowb + x1),
                                    q2 = *reinterpret_cast<const uint2*>(rowb + x2);
#pragma unroll
                        for (int dx = 0; dx < 2; ++dx) {
                            const uint2 pa = dx == 0 ? q0 : (sa1 ? q1 : q0);
                            const uint2 pb = dx == 0 ? (sb0 ? q1 : q0) : (sb1 == 2 ? q2 : sb1 == 1 ? q1 : q0);
                            const __half2 a0 = *reinterpret_cast<const __half2*>(&pa.x), a1 = *reinterpret_cast<const __half2*>(&pa.y);
                            const __half2 b0 = *reinterpret_cast<const __half2*>(&pb.x), b1 = *reinterpret_cast<const __half2*>(&pb.y);
                            const __half2 v0 = __hfma2(lx[dx], __hsub2(b0, a0), a0), v1 = __hfma2(lx[dx], __hsub2(b1, a1), a1);
                            rg[dx] = *reinterpret_cast<const uint32_t*>(&v0);
                            bx[dx] = *reinterpret_cast<const uint32_t*>(&v1);
                        }
                        o[dy * 3 + 0] = rg[0];                                       // (R0, G0)
                        o[dy * 3 + 1] = __byte_perm(bx[0], rg[1], 0x5410u);          // (B0, R1)
                        o[dy * 3 + 2] = __byte_perm(rg[1], bx[1], 0x5432u);          // (G1, B1)
                    }
                    const int pos = XLEAD + (py + 1) * XP + px;
                    *reinterpret_cast<uint4*>(Xb + pos * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                    // channel 12 = 1.0 (fp16 0x3C00): the centre tap's weight row 12 holds the folded-BN bias, so the GEMM adds it
                    *reinterpret_cast<uint4*>(Xb + X_CHUNK + pos * 16) = make_uint4(o[4], o[5], 0x3C00u, 0u);
                }
            }
            fence_proxy_async_smem();
#ifdef CV_FE_PROFILE
            if (prof_on) pacc[5] += clock64() - t_v0;
#endif
            TWAIT(3, asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory"));                       // operand image complete; HB free again
            if (t == 0) mbar_arrive(x_full + xslot);
            TRACE(0xB30);
        }
    } else if (warp == TMA_WARP) {
        // =========================== board-window stager ==================================================================
        // ONE 2D TMA tile copy per crop (rows x bytes box of the (B*H rows, H*3 bytes) board tensor at the clamped window origin; what
        // lies beyond the window is never read) into the RAW slots round robin, as soon as the vertical pass of the slot's previous
        // crop has ended: up to two crops ahead of the producers (a window is 48 short rows from as many DRAM pages: ~2-4 k cycles).
        if (elect_one()) {
            uint32_t it = 0, phase = 0;
            int slot = 0;
            for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
                if ((int)it >= p.n_rawbuf) mbar_wait(raw_empty + slot, phase ^ 1u);      // the slot's previous use (one round ago) has been consumed
                int64_t b = n >> 6;
                int sq = n & 63;
                if (p.perm_boards > 0) perm_inv(n, p.perm_boards, b, sq);
                const int rr = sq >> 3, cc = sq & 7;
                uint64_t* bar = raw_full + slot;
                mbar_arrive_expect_tx(bar, (uint32_t)p.box_bytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(RAW + slot * p.raw_bytes)), "l"(&tmap), "r"(tab.byte0[cc] >> 2), "r"((int)b * p.H + tab.row0[rr]), "r"(smem_u32(bar))
                             : "memory");
                if (++slot == p.n_rawbuf) { slot = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ===========================================================================
        // ONE thread runs the whole role (elected once): its barrier waits, descriptor adds and UTCHMMAs form one instruction stream
        // without a warp-wide wait + election + reconvergence per tile.  With 4 MMAs (~160 cycles of tensor-pipe work) per stem tile,
        // that per-block overhead (~150 cycles) had kept the pipe half idle: the pipe queue does not run far ahead of the issuer.
        if (elect_one()) {
            constexpr int N1 = F16 ? 16 : 32;                      // blocks.0.0 GEMM width: fp16 W | bf16 W_hi, W_lo
            if (F16) {
                mbar_arrive_expect_tx(wbar, W_STEM_BYTES + W_B00_F16_BYTES);
                bulk_g2s(W, p.wimg, W_STEM_BYTES, wbar);
                bulk_g2s(W + W_STEM_BYTES, p.wimg + W_BYTES, W_B00_F16_BYTES, wbar);
            } else {
                mbar_arrive_expect_tx(wbar, W_BYTES);
                bulk_g2s(W, p.wimg, W_BYTES, wbar);
            }
            mbar_wait_spin(wbar, 0);
            constexpr uint32_t idesc_s = make_idesc_f16(128, 32), idesc_1 = make_idesc_16<F16>(128, N1);
            const uint32_t x_lo = desc_lo(smem_u32(X), X_CHUNK), y_lo = desc_lo(smem_u32(Y), YH_CHUNK);
            const uint32_t ws_lo = desc_lo(smem_u32(W), 32 * 16), w1_lo = desc_lo(smem_u32(W + W_STEM_BYTES), N1 * 16);
            constexpr uint32_t x_hi = desc_hi(XP * 16), y_hi = desc_hi(YHP * 16), w_hi = desc_hi(128);
            // stem tiles k = 2 s + h, k in [k0, k1), of the crop in X slot it & xmask: output columns [8s, 8s+8), rows [16h, 16h+16)
            auto issue_stem = [&](int k0, int k1, uint32_t it) {
                const uint32_t xb_lo = x_lo + (it & xmask) * (X_ALLOC >> 4);
#pragma unroll
                for (int k = k0; k < k1; ++k) {
                    const int s = k >> 1, h = k & 1;
#pragma unroll
                    for (int tap = 0; tap < 4; ++tap) {
                        const int Dy = (tap >> 1) - 1, Dx = (tap & 1) - 1;
                        mma_f16_ss2(tmem_base + k * 32, xb_lo + (uint32_t)(XLEAD + (16 * h + 1 + Dy) * XP + 8 * s + Dx), x_hi, ws_lo + tap * 64, w_hi,
                                    idesc_s, tap > 0 ? 1u : 0u);
                    }
                    mma_commit(d_full + k);
                    TRACE(0x200 + k);
                }
            };
            // blocks.0.0 tile t (output columns [8t, 8t+8), all 16 rows) of the crop with iteration index itb: reads half image t
            auto issue_b00 = [&](int t, uint32_t itb) {
                const uint32_t buf = t;
                if (!FE_DBG(16 | 1024)) TWAIT(0, mbar_wait_spin(yh_full + buf, itb & 1u));
                TRACE(0x300 + t);
                tc_fence_after();
                const uint32_t yb_lo = y_lo + buf * (YH_BYTES >> 4);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    const int plane = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
                    const int Dy = ky == 0 ? -1 : 0, Dx = kx == 0 ? -1 : 0;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        mma_f16_ss2(tmem_base + TM_B00 + t * 32,
                                    yb_lo + (uint32_t)(plane * (YH_PLANE >> 4) + ks * 2 * YHPOS + YLEAD + (1 + Dy) * YHP + 1 + Dx), y_hi,
                                    w1_lo + (tap * 2 + ks) * (N1 * 2), w_hi, idesc_1, (tap | ks) ? 1u : 0u);
                }
                mma_commit(e_full + t);
                TRACE(0x400 + t);
            };
            uint32_t it = 0;
            for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
                TRACE_WINDOW(it);
                if (!FE_DBG(32)) TWAIT(2, mbar_wait_spin(x_full + (it & xmask), (it >> xshift) & 1u));
                TRACE(0x500);
                // one wait per crop for the eight stem accumulators (each wait costs the issuing thread ~100+ cycles of exposed latency
                // even when it is satisfied: thirteen of them per crop had kept the tensor pipe at half rate)
                if (!FE_DBG(16 | 128)) TWAIT(3, mbar_wait_spin(d_empty, (it & 1u) ^ 1u));
                tc_fence_after();
                // stem left half | previous crop's right tile | stem right half | this crop's left tile: each blocks.0.0 tile has a whole
                // other tile + half a stem (~1.4 k cycles of tensor-pipe work) queued between the stem tiles it depends on and itself, which
                // covers the epilogue's round trip.  The right half image (and its halo column, which the epilogue of slab 1 holds back in
                // registers until its first right-half tile has completed) is only written after the previous crop's right tile has read it.
                issue_stem(0, 4, it);                                // slabs 0, 1: left half image
                if (it > 0) issue_b00(1, it - 1);                    // previous crop, right tile
                issue_stem(4, 8, it);                                // slabs 2, 3: right half image
                mma_commit(x_empty + (it & xmask));                  // operand image free once the stem MMAs have read it
                issue_b00(0, it);                                    // this crop, left tile
            }
            if (it > 0) issue_b00(1, it - 1);
        }
        __syncwarp();
    } else {
        // =========================== epilogue warps 0-7 ====================================================================
        const int g = warp >> 2, qd = warp & 3, i = qd * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
        float bb[16];                                            // blocks.0.0 bias in registers (the stem bias is folded into its GEMM)
#pragma unroll
        for (int k = 0; k < 16; ++k) bb[k] = p.bias_b00[k];
        // stem tile (s, h = g): this lane holds pixel y = 16 g + (i >> 3), x = 8 s + (i & 7)
        const int y = 16 * g + (i >> 3), xl = i & 7;
        const int yoff = ((y & 1) * 2 + (xl & 1)) * YH_PLANE + (YLEAD + ((y >> 1) + 1) * YHP + (xl >> 1) + 1) * 16;   // + 64 for odd s
        const int halo_off = ((y & 1) * 2 + 1) * YH_PLANE + (YLEAD + ((y >> 1) + 1) * YHP) * 16;                        // column 0 of the odd-x plane
        auto epilogue_b00 = [&](int n, uint32_t itb) {           // blocks.0.0 tile g of crop n -> global T8 tile rows
            TWAIT(0, mbar_wait(e_full + g, itb & 1u));
            TRACE(0xA00 + g);
            tc_fence_after();
            uint32_t eh[16], el[16];                             // bf16: hi | lo halves of the 16 output channels
            tmem_ld16(trow + TM_B00 + g * 32, eh);
            if (!F16) tmem_ld16(trow + TM_B00 + g * 32 + 16, el);
            tmem_ld_wait();

            tc_fence_before();
            __syncwarp();
            const int64_t m = (int64_t)n * 256 + (i >> 3) * 16 + 8 * g + (i & 7);
            uint4* dst = reinterpret_cast<uint4*>(p.y) + ((m >> 7) * 2) * 128 + (m & 127);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = 8 * c + 2 * q;
                    const float v0 = F16 ? __uint_as_float(eh[e]) : __uint_as_float(eh[e]) + __uint_as_float(el[e]);
                    const float v1 = F16 ? __uint_as_float(eh[e + 1]) : __uint_as_float(eh[e + 1]) + __uint_as_float(el[e + 1]);
                    w[q] = pk2r<F16>(v0 + bb[e], v1 + bb[e + 1]);
                }
                dst[c * 128] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        };
        uint32_t it = 0;
        int prev_n = -1;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            uint4 halo[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int k = 2 * s + g;
                uint8_t* yb = Y + (s >> 1) * YH_BYTES;
                if (s == 0) TRACE_WINDOW(it);
                if (FE_DBG(64)) { if (lane == 0) mbar_wait_spin(d_full + k, it & 1u); __syncwarp(); } else TWAIT(1, mbar_wait(d_full + k, it & 1u));
                TRACE(0x600 + k);
                tc_fence_after();
                uint32_t r0[16], r1[16];
                if (!FE_DBG(4)) {
                    tmem_ld16(trow + k * 32, r0);
                    tmem_ld16(trow + k * 32 + 16, r1);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) r0[q] = r1[q] = 0x3f800000u;
                }
                tc_fence_before();
                __syncwarp();
                if (s == 3 && lane == 0) mbar_arrive(d_empty);   // this warp's four stem accumulators of the crop are drained
                TRACE(0x700 + k);
                uint4 o[4];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    o[c] = make_uint4(pk2r<F16>(__uint_as_float(r0[8 * c]), __uint_as_float(r0[8 * c + 1])),
                                      pk2r<F16>(__uint_as_float(r0[8 * c + 2]), __uint_as_float(r0[8 * c + 3])),
                                      pk2r<F16>(__uint_as_float(r0[8 * c + 4]), __uint_as_float(r0[8 * c + 5])),
                                      pk2r<F16>(__uint_as_float(r0[8 * c + 6]), __uint_as_float(r0[8 * c + 7])));
                    o[2 + c] = make_uint4(pk2r<F16>(__uint_as_float(r1[8 * c]), __uint_as_float(r1[8 * c + 1])),
                                          pk2r<F16>(__uint_as_float(r1[8 * c + 2]), __uint_as_float(r1[8 * c + 3])),
                                          pk2r<F16>(__uint_as_float(r1[8 * c + 4]), __uint_as_float(r1[8 * c + 5])),
                                          pk2r<F16>(__uint_as_float(r1[8 * c + 6]), __uint_as_float(r1[8 * c + 7])));
                }
                uint8_t* dst = yb + yoff + (s & 1) * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (!FE_DBG(8) || c == 0) *reinterpret_cast<uint4*>(dst + c * YH_CHUNK) = o[c];
                if (s == 1) {                                    // x = 15 is also the halo column of the right half: held back until the
#pragma unroll                                                   // previous crop's right tile has read that image (see the MMA role)
                    for (int c = 0; c < 4; ++c) halo[c] = o[c];
                }
                if (s == 2 && xl == 7) {                         // tile 4 + g has completed, so has every MMA issued before it
                    uint8_t* yr = Y + YH_BYTES + halo_off;
#pragma unroll
                    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(yr + c * YH_CHUNK) = halo[c];
                }
                if (s & 1) {                                     // half image complete (this warp's share)
                    // First drain this group's blocks.0.0 tile of the PREVIOUS crop (tile g was issued before the stem tiles just
                    // processed, so it has completed: the tensor pipe executes in order).  The half-image barrier below then also
                    // tells the MMA thread that accumulator g is free -- the next writer of accumulator g is exactly the
                    // blocks.0.0 tile that waits for half image g of this crop -- so no separate "accumulator empty" wait exists.
                    if (it > 0 && (s >> 1) == g) epilogue_b00(prev_n, it - 1);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(yh_full + (s >> 1));
                    TRACE(0x900 + (s >> 1));
                }
            }
            prev_n = n;
        }
        if (it > 0) epilogue_b00(prev_n, it - 1);
    }
#ifdef CV_FE_PROFILE
    if (prof_on) {
        const int role = warp == 0 ? 0 : warp == 4 ? 1 : warp == pw ? 2 : 3;
        for (int k = 0; k < 8; ++k) g_fe3_prof[role * 16 + k] = (unsigned long long)pacc[k];
        g_fe3_prof[role * 16 + 8] = (unsigned long long)(clock64() - t_role0);
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

// Weight images: conv_stem as a 2x2 conv on the space-to-depth crop (4 taps x K 16, fp16, bias in row 12 of the centre tap),
// blocks.0.0 as 9 taps x 2 k-steps, bf16 W_hi | W_lo concatenated along N.  *flag is set when a stem weight does not fit fp16.
__global__ void prep_frontend3_weights_kernel(const float* __restrict__ w_stem /*[27][32]*/, const float* __restrict__ b_stem /*[32]*/,
                                              const float* __restrict__ w_b00 /*[288][16]*/, uint16_t* __restrict__ img, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W_IMG_BYTES / 2) return;
    float v = 0.f;
    if (i < W_STEM_BYTES / 2) {
        const int kk = i & 7, n = (i >> 3) & 31, chunk = (i >> 8) & 1, tap = i >> 9;
        const int k = chunk * 8 + kk;                                  // s2d channel = dy*6 + dx*3 + c
        if (k < 12) {
            const int dy = k / 6, dx = (k % 6) / 3, c = k % 3;
            const int ky = 2 * ((tap >> 1) - 1) + dy + 1, kx = 2 * ((tap & 1) - 1) + dx + 1;
            if (ky >= 0 && ky < 3 && kx >= 0 && kx < 3) v = w_stem[((ky * 3 + kx) * 3 + c) * 32 + n];
        } else if (k == 12 && tap == 3) {
            v = b_stem[n];                                             // bias row: multiplied by the constant-one channel of the centre tap
        }
        if (!(fabsf(v) <= 65504.f)) atomicOr(flag, 1);
        img[i] = __half_as_ushort(__float2half_rn(v));
    } else if (i < W_BYTES / 2) {
        const int j = i - W_STEM_BYTES / 2;
        const int kk = j & 7, n = (j >> 3) & 31, chunk = (j >> 8) & 1, ks = (j >> 9) & 1, tap = j >> 10;
        const int ci = ks * 16 + chunk * 8 + kk;
        v = w_b00[(tap * 32 + ci) * 16 + (n & 15)];
        const bf16 hi = __float2bfloat16_rn(v);
        img[i] = __bfloat16_as_ushort(n >= 16 ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi);
    } else {                                                           // fp16 blocks.0.0 image: (tap, k-step) x [2 chunks][16 n][8]
        const int j = i - W_BYTES / 2;
        const int kk = j & 7, n = (j >> 3) & 15, chunk = (j >> 7) & 1, ks = (j >> 8) & 1, tap = j >> 9;
        const int ci = ks * 16 + chunk * 8 + kk;
        v = w_b00[(tap * 32 + ci) * 16 + n];
        if (!(fabsf(v) <= 65504.f)) atomicOr(flag, 1);
        img[i] = __half_as_ushort(__float2half_rn(v));
    }
}

}  // namespace

size_t frontend3_weight_image_bytes() { return W_IMG_BYTES; }

int launch_frontend3_prep_weights(const float* blob, uint8_t* img, int* flag_dev, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    CV_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int), s));
    prep_frontend3_weights_kernel<<<(W_IMG_BYTES / 2 + 255) / 256, 256, 0, s>>>(blob + L[0].w_offset, blob + L[0].b_offset, blob + L[1].w_offset,
                                                                            reinterpret_cast<uint16_t*>(img), flag_dev);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

// Returns CV_OK and sets *supported = 0 when this kernel cannot take the configuration (the caller then uses an earlier
// generation): non-affine normalisation table, window too large for shared memory (512x512 boards), > 255 window rows.
int launch_frontend3(const uint8_t* boards_hwc, int nb, int H, const CropGeom& g, const float* lut_host, const uint8_t* wimg,
                     const float* bias_b00, bf16* y, int num_sms, int* supported, cudaStream_t s, const int* skip_flag, const StageGate& gate,
                     int perm_boards) {
    *supported = 0;
    if (nb == 0) { *supported = 1; return CV_OK; }
    Front3Params p{};
    p.perm_boards = perm_boards;
    // normalisation must be affine in the byte value: v = na*u + nb (true for ToTensor + Normalize)
    for (int c = 0; c < 3; ++c) {
        const float* l = lut_host + c * 256;
        p.nb[c] = l[0];
        p.na[c] = (l[255] - l[0]) / 255.0f;
        for (int u = 0; u < 256; ++u)
            if (fabsf(p.na[c] * u + p.nb[c] - l[u]) > 1e-5f * (1.0f + fabsf(l[u]))) return CV_OK;
        if (fabsf(l[0]) > 1000.f || fabsf(l[255]) > 1000.f) return CV_OK;      // the resized pixels are held in fp16
    }
    Front3Tables tab{};
    const CropTaps tp = make_taps(g);
    int max_rows = 0, max_px = 0, max_bytes = 0;
    for (int r = 0; r < 8; ++r) {                                   // the same taps serve square rows (vertical) and square columns (horizontal)
        int lo = g.H, hi = -1;
        for (int d = 0; d < 64; ++d) {
            lo = lo < tp.p0[r][d] ? lo : tp.p0[r][d];
            hi = hi > tp.p1[r][d] ? hi : tp.p1[r][d];
        }
        if (hi - lo > 250) return CV_OK;
        tab.row0[r] = lo;
        tab.nrows[r] = hi - lo + 1;
        const int lo4 = lo & ~3;                                    // staged rows start on a 4-pixel (12-byte) group: word-aligned pixel groups
        tab.byte0[r] = (lo4 * 3) & ~15;
        tab.boff[r] = lo4 * 3 - tab.byte0[r];                       // 0, 4, 8 or 12
        const int npx = (hi - lo4 + 1 + 3) & ~3;
        max_bytes = max_bytes > tab.boff[r] + npx * 3 ? max_bytes : tab.boff[r] + npx * 3;
        for (int d = 0; d < 64; ++d) {
            const __half lh = __float2half_rn(g.lam[d]);
            if (__half2float(lh) != g.lam[d]) return CV_OK;                 // the interpolation weights must be exact in fp16 (they are k/128)
            tab.vy[r][d] = (uint32_t)(tp.p0[r][d] - lo) | ((uint32_t)(tp.p1[r][d] - lo) << 8) | ((uint32_t)__half_as_ushort(lh) << 16);
            tab.xh0[r][d] = (uint8_t)(tp.p0[r][d] - lo4);
            tab.xh1[r][d] = (uint8_t)(tp.p1[r][d] - lo4);
            tab.lamh[d] = __half_as_ushort(lh);
        }
        max_rows = max_rows > tab.nrows[r] ? max_rows : tab.nrows[r];
        max_px = max_px > npx ? max_px : npx;
    }
    for (int r = 0; r < 8; ++r)                                     // horizontal pass: the taps of an output column pair within three pixels
        for (int px = 0; px < 32; ++px) {
            const int a0 = tab.xh0[r][2 * px], b0 = tab.xh1[r][2 * px], a1 = tab.xh0[r][2 * px + 1], b1 = tab.xh1[r][2 * px + 1];
            if (b0 < a0 || b0 > a0 + 1 || a1 < a0 || a1 > a0 + 1 || b1 < a0 || b1 > a0 + 2) return CV_OK;
        }
    max_bytes = (max_bytes + 15) & ~15;                             // TMA box: inner extent a multiple of 16 bytes
    p.raw_pitch = max_bytes;
    p.raw_bytes = (max_rows * max_bytes + 127) & ~127;
    p.n_groups = max_px / 4;
    p.g_magic = (int)((0x100000000ull + p.n_groups - 1) / p.n_groups);      // item / n_groups = umulhi(item, g_magic) for item < 2^16
    p.v_pitch = max_px * 8;
    p.v_bytes = (64 * p.v_pitch + 127) & ~127;
    const int fixed = NYBUF * YH_BYTES + 256 + W_BYTES + (int)sizeof(Front3Tables) + 384 + 1024;
    p.n_rawbuf = 3; p.n_xbuf = NXBUF;
    auto total = [&]() { return p.n_xbuf * X_ALLOC + p.n_rawbuf * p.raw_bytes + p.v_bytes + fixed; };
    if (total() > 227 * 1024) p.n_rawbuf = 2;
    if (total() > 227 * 1024) p.n_rawbuf = 1;                     // larger windows (512x512 boards): single window buffer,
    if (total() > 227 * 1024) p.n_xbuf = 1;                       // then a single stem operand image
    if (total() > 227 * 1024) return CV_OK;                       // window does not fit: not supported
    int off = p.n_xbuf * X_ALLOC;
    p.off_raw = off; off += p.n_rawbuf * p.raw_bytes;
    off = (off + 127) & ~127; p.off_v = off; off += p.v_bytes;
    off = (off + 127) & ~127; p.off_y = off; off += NYBUF * YH_BYTES + 256;
    off = (off + 127) & ~127; p.off_w = off; off += W_BYTES;
    p.off_tab = off; off += ((int)sizeof(Front3Tables) + 15) & ~15;
    off = (off + 15) & ~15; p.off_bar = off; off += 384;
    p.smem_total = off;
    if (p.smem_total > 227 * 1024) return CV_OK;
    p.boards = boards_hwc; p.wimg = wimg; p.bias_b00 = bias_b00; p.y = y;
    p.n_crops = nb * 64; p.H = H;
    p.skip_flag = skip_flag;
    p.gate = gate.flag; p.gate_want = gate.want;
#if defined(CV_EXPERIMENTS) || defined(CV_FE_PROFILE)   // ablation / timing / trace switches (some change the results): experiment builds only
    { const char* d = getenv("CV_FE3_DEBUG"); p.debug = d ? atoi(d) : 0; }
#endif
    // 2D tensor map over the boards of this launch: (nb*H rows) x (H*3 bytes as uint32 elements); box = the largest window
    if ((reinterpret_cast<uintptr_t>(boards_hwc) & 15) || (H * 3) % 16 || max_bytes / 4 > 256 || max_rows > 256) return CV_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return CV_OK;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    alignas(64) CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)(H * 3 / 4), (cuuint64_t)nb * H};
    const cuuint64_t gstride[1] = {(cuuint64_t)H * 3};
    const cuuint32_t box[2] = {(cuuint32_t)(max_bytes / 4), (cuuint32_t)max_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(boards_hwc), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { cv_set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return CV_ERR_CUDA; }
    p.box_bytes = max_rows * max_bytes;
    *supported = 1;
    auto kern = gate.f16 ? frontend3_kernel<true> : frontend3_kernel<false>;
    CV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_total));
    const int grid = p.n_crops < num_sms ? p.n_crops : num_sms;
    kern<<<grid, NTHREADS, p.smem_total, s>>>(p, tab, tmap);
    CV_CHECK_LAUNCH();
#ifdef CV_FE_PROFILE
    if (p.debug & 256) {                                          // timing experiment: print block 0's wait-cycle counters
        unsigned long long h[64];
        CV_CUDA(cudaStreamSynchronize(s));
        CV_CUDA(cudaMemcpyFromSymbol(h, g_fe3_prof, sizeof(h)));
        const int per_cta = (p.n_crops + grid - 1) / grid;
        const char* names[4] = {"epilogue g0", "epilogue g1", "producer   ", "mma        "};
        const char* what[4][4] = {{"e_full", "d_full", "-", "-"}, {"e_full", "d_full", "-", "-"},
                                  {"raw_full", "x_empty", "bar A", "bar B"}, {"yh_full", "e_empty", "x_full", "d_empty"}};
        if (p.debug & 512) {
            static uint2 tr[4 * 1024];
            CV_CUDA(cudaMemcpyFromSymbol(tr, g_fe3_trace, sizeof(tr)));
            FILE* tf = fopen("gpurun_out/fe3_trace.txt", "w");
            if (tf) {
                for (int r = 0; r < 4; ++r)
                    for (int i = 0; i < 1024 && tr[r * 1024 + i].x; ++i) fprintf(tf, "%d %x %u\n", r, tr[r * 1024 + i].x, tr[r * 1024 + i].y);
                fclose(tf);
            }
        }
        fprintf(stderr, "fe3 config: n_xbuf %d n_rawbuf %d smem %d raw_bytes %d v_bytes %d groups %d\n", p.n_xbuf, p.n_rawbuf, p.smem_total, p.raw_bytes, p.v_bytes, p.n_groups);
        for (int r = 0; r < 4; ++r) {
            fprintf(stderr, "fe3 %s total %7.0f cyc/crop | waits:", names[r], (double)h[r * 16 + 8] / per_cta);
            for (int k = 0; k < 4; ++k) fprintf(stderr, " %s %6.0f", what[r][k], (double)h[r * 16 + k] / per_cta);
            if (r == 2) fprintf(stderr, " | horizontal %6.0f vertical %6.0f", (double)h[r * 16 + 4] / per_cta, (double)h[r * 16 + 5] / per_cta);
            fprintf(stderr, "\n");
        }
    }
#endif
    return CV_OK;
}


// ---- float source that is a uint8 image in disguise -------------------------------------------------------------------------------
// The reference's own call is model(images) with images = Normalize(ToTensor(uint8 image)) (dataset.py:177-181): every value is
// na[c] * u + nb[c] for an integer u in 0..255.  This kernel inverts that (4 pixels per thread: three float4 loads, three 32-bit stores,
// NCHW float -> HWC uint8) and raises *flag when a value is NOT within 1e-5 of such a grid point (the grid step is ~0.017); the third-
// generation front end then runs on the bytes (and exits at once if the flag is up), the first-generation one on the floats (and exits
// at once if it is not): the float entry point gets the fast front end for what the reference actually feeds it, without a host sync.
namespace {
__global__ void __launch_bounds__(256) f32_nchw_to_u8_hwc_kernel(const float* __restrict__ x, int64_t n_quads, int HH, float3 na, float3 nb, float3 inv,
                                                                 uint8_t* __restrict__ out, int* __restrict__ flag) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool bad = false;
    if (q < n_quads) {
        const int64_t px = q * 4, b = px / HH;
        const int p0 = (int)(px - b * HH);
        const float* base = x + b * 3 * (int64_t)HH + p0;
        const float4 r = *reinterpret_cast<const float4*>(base), g = *reinterpret_cast<const float4*>(base + HH),
                     bl = *reinterpret_cast<const float4*>(base + 2 * (int64_t)HH);
        const float v[4][3] = {{r.x, g.x, bl.x}, {r.y, g.y, bl.y}, {r.z, g.z, bl.z}, {r.w, g.w, bl.w}};
        const float a3[3] = {na.x, na.y, na.z}, b3[3] = {nb.x, nb.y, nb.z}, i3[3] = {inv.x, inv.y, inv.z};
        uint32_t bytes[12];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float u = fminf(fmaxf(rintf((v[i][c] - b3[c]) * i3[c]), 0.f), 255.f);
                bad |= !(fabsf(fmaf(a3[c], u, b3[c]) - v[i][c]) <= 1e-5f);
                bytes[i * 3 + c] = (uint32_t)u;
            }
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + px * 3);
#pragma unroll
        for (int w = 0; w < 3; ++w) dst[w] = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | (bytes[4 * w + 3] << 24);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
}  // namespace

int launch_f32_to_u8_boards(const float* x_nchw, int nb, int H, const float* lut_host, uint8_t* out_hwc, int* flag_dev, cudaStream_t s) {
    if (nb == 0) return CV_OK;
    float a[3], b[3], inv[3];
    for (int c = 0; c < 3; ++c) {
        b[c] = lut_host[c * 256];
        a[c] = (lut_host[c * 256 + 255] - lut_host[c * 256]) / 255.0f;
        inv[c] = a[c] != 0.f ? 1.0f / a[c] : 0.f;
    }
    CV_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int), s));
    const int64_t n_quads = (int64_t)nb * H * H / 4;
    f32_nchw_to_u8_hwc_kernel<<<(unsigned)((n_quads + 255) / 256), 256, 0, s>>>(x_nchw, n_quads, H * H, make_float3(a[0], a[1], a[2]),
                                                                               make_float3(b[0], b[1], b[2]), make_float3(inv[0], inv[1], inv[2]),
                                                                               out_hwc, flag_dev);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
