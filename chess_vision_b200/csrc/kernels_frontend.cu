// Fused front end of the ChessSquareCNN hot path for sm_100a:
//
//     uint8 board --> [crop gather + normalise + bilinear 48->64] --> conv_stem 3x3 s2 (+BN+ReLU)
//                 --> blocks.0.0 3x3 s2 (+BN+ReLU) --> T8 activations [crops*256 rows][16 ch] in HBM
//
// (ChessSquareCNN._crop_squares, models/square.py:43-74, and the first two convs of the timm trunk,
//  models/square.py:86.)  One persistent CTA per SM walks over crops; nothing but the 8 KB/crop result is
// written to HBM -- the 24 KB crop and the 64 KB stem activation (40 % of the layer-granular traffic of the
// whole network) live only in shared memory.  Both convolutions run on tcgen05 tensor cores as implicit GEMMs
// WITHOUT an im2col copy: the activations are laid out in shared memory so that every filter tap's A operand is
// a plain K-major / no-swizzle UMMA descriptor pointing into the activation itself.
//
//   stem   : the crop is stored space-to-depth (2x2 pixels x 3 ch -> 12(+4 zero) channels per s2d pixel, one
//            16-byte chunk pair), 33 x 40 positions per channel chunk (zero row on top, 8 zero columns on the
//            right).  A 3x3 stride-2 conv on the crop == a 2x2 stride-1 conv on the s2d image, so tap (Dy,Dx) in
//            {-1,0}^2 for the 128 consecutive positions of M-tile j is the descriptor at
//            X + (lead + 40 + 128 j + 40 Dy + Dx) * 16 B, SBO 128 B, LBO = chunk-plane size: 4 MMAs (K=16) per tile.
//            Position -1 of a row is the previous row's zero halo => zero padding comes for free; the 8 junk
//            columns per row cost 20 % extra MMA rows and are dropped in the epilogue.
//   b0.0   : the stem epilogue scatters its bf16 output into four parity planes (y%2, x%2) of 17 x 24 positions
//            x 4 channel chunks; tap (ky,kx) of the stride-2 conv reads plane ((ky!=1),(kx!=1)) shifted by
//            (Dy,Dx) = (-(ky==0), -(kx==0)): again one descriptor per tap and k-step, 18 MMAs (K=16) per M-tile,
//            3 M-tiles per crop.
//   weights: W = W_hi + W_lo (bf16 each) concatenated along N (stem N = 32+32, b0.0 N = 16+16); the epilogue adds
//            the two column halves, so the bf16 rounding of the weights costs no accuracy and no extra A reads.
//
// Warp roles (13 warps): 0-3 epilogue (TMEM lane quadrants), 4 MMA issuer + TMEM owner, 5-12 crop producers.
// mbarrier protocol per crop: x_full/x_empty (producers <-> MMA), d_full/d_empty[4] (stem accumulators, 4-deep),
// y_full/y_empty (stem epilogue <-> b0.0 MMAs), e_full/e_empty (b0.0 accumulators <-> epilogue).
#include "internal.h"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int XP = 40, XROWS = 33, XLEAD = 8;
constexpr int XPOS = XLEAD + XROWS * XP;              // 1328 positions per channel-chunk plane
constexpr int X_CHUNK_BYTES = XPOS * 16;              // 21248
constexpr int X_BYTES = 2 * X_CHUNK_BYTES;            // 42496
constexpr int YP = 24, YROWS = 17, YLEAD = 8;
constexpr int YPOS = YLEAD + YROWS * YP;              // 416
constexpr int Y_CHUNK_BYTES = YPOS * 16;              // 6656
constexpr int Y_PLANE_STRIDE = 4 * Y_CHUNK_BYTES + 64;   // +64 B skew: the two x-parities hit different banks
constexpr int Y_BYTES = 4 * Y_PLANE_STRIDE;           // 106752
constexpr int WS_ELEMS = 4 * 2 * 64 * 8;              // stem B: 4 taps x [2 chunks][64 n][8]      = 4096 bf16
constexpr int W1_ELEMS = 18 * 2 * 32 * 8;             // b0.0 B: 18 k-steps x [2 chunks][32 n][8]  = 9216 bf16
constexpr int W_BYTES = (WS_ELEMS + W1_ELEMS) * 2;    // 26624
constexpr int ND = 4;                                 // stem accumulator ring depth
constexpr int NPROD = 256;                            // producer threads
constexpr int NTHREADS = 160 + NPROD;                 // 416

constexpr int OFF_X = 0;
constexpr int OFF_Y = OFF_X + X_BYTES;
constexpr int OFF_W = OFF_Y + Y_BYTES;
constexpr int OFF_LUT = OFF_W + W_BYTES;
constexpr int OFF_BIAS = OFF_LUT + 768 * 4;
constexpr int OFF_TAPS = OFF_BIAS + 48 * 4;             // CropTaps copy: divergent indexing is slow from the constant bank
constexpr int OFF_BAR = OFF_TAPS + (int)sizeof(CropTaps);
constexpr int NBAR = 7 + 2 * ND;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_TOTAL = OFF_TMEM + 16;
static_assert(SMEM_TOTAL <= 227 * 1024, "front-end shared memory budget");
static_assert(OFF_Y % 128 == 0 && OFF_W % 128 == 0 && OFF_BAR % 8 == 0 && OFF_TAPS % 4 == 0 && sizeof(CropTaps) % 4 == 0, "alignment");

struct FrontParams {
    const void* src;
    const float* lut;
    const bf16* wimg;
    const float* bias_stem;
    const float* bias_b00;
    bf16* y;
    int n_crops, H;
    const int* run_flag;      // non-null: the kernel exits at once when *run_flag == 0 (the third-generation front end took the call)
    int out_f16;              // the T8 output is fp16 (feeds the fp16 stage B) instead of bf16; the convolutions themselves stay bf16 hi|lo
    const int* gate;          // internal.h StageGate: run only when (*gate != 0) == gate_want
    int gate_want;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int SRC>
__device__ __forceinline__ void load_taps(const FrontParams& p, const float* lut, int64_t b, int y0, int y1, int x0, int x1,
                                          float (&v)[3][4]) {
    const int H = p.H;
    if (SRC == CV_SRC_F32_NCHW) {
        const float* base = reinterpret_cast<const float*>(p.src) + b * 3 * (int64_t)H * H;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* pc = base + (int64_t)c * H * H;
            v[c][0] = __ldg(pc + y0 * H + x0); v[c][1] = __ldg(pc + y0 * H + x1);
            v[c][2] = __ldg(pc + y1 * H + x0); v[c][3] = __ldg(pc + y1 * H + x1);
        }
    } else {
        const uint8_t* base = reinterpret_cast<const uint8_t*>(p.src) + b * 3 * (int64_t)H * H;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* l = lut + c * 256;
            if (SRC == CV_SRC_U8_CHW) {
                const uint8_t* pc = base + (int64_t)c * H * H;
                v[c][0] = l[__ldg(pc + y0 * H + x0)]; v[c][1] = l[__ldg(pc + y0 * H + x1)];
                v[c][2] = l[__ldg(pc + y1 * H + x0)]; v[c][3] = l[__ldg(pc + y1 * H + x1)];
            } else {
                v[c][0] = l[__ldg(base + (y0 * H + x0) * 3 + c)]; v[c][1] = l[__ldg(base + (y0 * H + x1) * 3 + c)];
                v[c][2] = l[__ldg(base + (y1 * H + x0) * 3 + c)]; v[c][3] = l[__ldg(base + (y1 * H + x1) * 3 + c)];
            }
        }
    }
}

template <int SRC>
__global__ void __launch_bounds__(NTHREADS, 1)
frontend_kernel(const __grid_constant__ FrontParams p, const __grid_constant__ CropTaps tp_param) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (p.run_flag != nullptr && *p.run_flag == 0) return;       // uniform over the grid
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_want != 0)) return;
    uint8_t* X = smem + OFF_X;
    uint8_t* Y = smem + OFF_Y;
    uint8_t* W = smem + OFF_W;
    float* lut = reinterpret_cast<float*>(smem + OFF_LUT);
    float* bias = reinterpret_cast<float*>(smem + OFF_BIAS);      // [0,32) stem, [32,48) blocks.0.0
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t *x_full = bars, *x_empty = bars + 1, *y_full = bars + 2, *y_empty = bars + 3, *e_full = bars + 4,
             *e_empty = bars + 5, *wbar = bars + 6, *d_full = bars + 7, *d_empty = bars + 7 + ND;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: zero the activation buffers (halos stay zero for the whole kernel), tables, barriers
    for (int i = threadIdx.x; i < (X_BYTES + Y_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 768; i += NTHREADS) lut[i] = p.lut ? p.lut[i] : 0.f;
    for (int i = threadIdx.x; i < 48; i += NTHREADS) bias[i] = i < 32 ? p.bias_stem[i] : p.bias_b00[i - 32];
    for (int i = threadIdx.x; i < (int)(sizeof(CropTaps) / 4); i += NTHREADS)
        reinterpret_cast<uint32_t*>(smem + OFF_TAPS)[i] = reinterpret_cast<const uint32_t*>(&tp_param)[i];
    const CropTaps& tp = *reinterpret_cast<const CropTaps*>(smem + OFF_TAPS);
    if (threadIdx.x == 0) {
        mbar_init(x_full, NPROD); mbar_init(x_empty, 1);
        mbar_init(y_full, 128); mbar_init(y_empty, 1);
        mbar_init(e_full, 1); mbar_init(e_empty, 4);
        mbar_init(wbar, 1);
        for (int i = 0; i < ND; ++i) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, 4); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_e = tmem_base + ND * 64;          // blocks.0.0 accumulators: 3 tiles x 32 columns

    if (warp >= 5) {
        // =========================== crop producers: 256 threads, 4 s2d pixels each ===========================
        const int t = threadIdx.x - 160;
        uint32_t it = 0;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            const int64_t b = n >> 6;
            const int row = (n >> 3) & 7, col = n & 7;
            mbar_wait(x_empty, (it & 1u) ^ 1u);
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int sp = t + NPROD * k, py = sp >> 5, px = sp & 31;
                uint32_t w[6];                                     // 12 bf16: channel = (dy*2+dx)*3 + c
                float vals[12];
#pragma unroll
                for (int sub = 0; sub < 4; ++sub) {
                    const int y = 2 * py + (sub >> 1), x = 2 * px + (sub & 1);
                    float v[3][4];
                    load_taps<SRC>(p, lut, b, tp.p0[row][y], tp.p1[row][y], tp.p0[col][x], tp.p1[col][x], v);
                    const float ly = tp.lam[y], lx = tp.lam[x];
#pragma unroll
                    for (int c = 0; c < 3; ++c) vals[sub * 3 + c] = crop_blend(v[c][0], v[c][1], v[c][2], v[c][3], lx, ly);
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) w[i] = pack2(vals[2 * i], vals[2 * i + 1]);
                const int pos = XLEAD + (py + 1) * XP + px;
                *reinterpret_cast<uint4*>(X + pos * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(X + X_CHUNK_BYTES + pos * 16) = make_uint4(w[4], w[5], 0u, 0u);
            }
            fence_proxy_async_smem();
            mbar_arrive(x_full);
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ================================================================
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, W_BYTES);
            bulk_g2s(W, p.wimg, W_BYTES, wbar);
        }
        mbar_wait(wbar, 0);
        const uint32_t idesc_s = make_idesc_bf16(128, 64), idesc_1 = make_idesc_bf16(128, 32);
        const uint32_t xb = smem_u32(X), yb = smem_u32(Y), ws = smem_u32(W), w1 = ws + WS_ELEMS * 2;
        uint32_t it = 0, tcount = 0;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            mbar_wait(x_full, it & 1u);
            tc_fence_after();
            for (int j = 0; j < 10; ++j, ++tcount) {
                const uint32_t buf = tcount & (ND - 1), ph = (tcount / ND) & 1u;
                mbar_wait(d_empty + buf, ph ^ 1u);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 4; ++tap) {
                        const int Dy = (tap >> 1) - 1, Dx = (tap & 1) - 1;
                        const uint32_t a = xb + (uint32_t)(XLEAD + XP + 128 * j + Dy * XP + Dx) * 16u;
                        mma_bf16_ss(tmem_base + buf * 64, make_smem_desc(a, X_CHUNK_BYTES, 128),
                                    make_smem_desc(ws + tap * 2048, 64 * 16, 128), idesc_s, tap > 0 ? 1u : 0u);
                    }
                    mma_commit(d_full + buf);
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(x_empty);            // crop buffer free once the stem MMAs have read it
            __syncwarp();
            mbar_wait(y_full, it & 1u);
            mbar_wait(e_empty, (it & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        const int plane = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
                        const int Dy = ky == 0 ? -1 : 0, Dx = kx == 0 ? -1 : 0;
                        const uint32_t a0 = yb + (uint32_t)plane * Y_PLANE_STRIDE + (uint32_t)(YLEAD + YP + 128 * j + Dy * YP + Dx) * 16u;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            mma_bf16_ss(tmem_e + j * 32, make_smem_desc(a0 + ks * 2 * Y_CHUNK_BYTES, Y_CHUNK_BYTES, 128),
                                        make_smem_desc(w1 + (tap * 2 + ks) * 1024, 32 * 16, 128), idesc_1,
                                        (tap | ks) ? 1u : 0u);
                    }
                }
                mma_commit(e_full);
                mma_commit(y_empty);
            }
            __syncwarp();
        }
    } else {
        // =========================== epilogue warps 0-3 ========================================================
        const int i = warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        uint32_t it = 0, tcount = 0;
        for (int n = blockIdx.x; n < p.n_crops; n += gridDim.x, ++it) {
            for (int j = 0; j < 10; ++j, ++tcount) {
                const uint32_t buf = tcount & (ND - 1), ph = (tcount / ND) & 1u;
                mbar_wait(d_full + buf, ph);
                tc_fence_after();
                uint32_t r[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld16(tmem_base + lane_base + buf * 64 + c * 16, r[c]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty + buf);           // accumulator drained into registers
                if (j == 0) mbar_wait(y_empty, (it & 1u) ^ 1u);      // previous crop's b0.0 MMAs are done with Y
                const int q = 128 * j + i, oy = q / XP, ox = q - oy * XP;
                if (ox < 32) {
                    const int plane = (oy & 1) * 2 + (ox & 1);
                    uint8_t* dst = Y + plane * Y_PLANE_STRIDE + (YLEAD + ((oy >> 1) + 1) * YP + (ox >> 1)) * 16;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {                     // output channels 8c .. 8c+7
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int ch = 8 * c + e;                // hi half: columns [0,32), lo half: [32,64)
                            v[e] = fmaxf(__uint_as_float(r[ch >> 4][ch & 15]) + __uint_as_float(r[2 + (ch >> 4)][ch & 15]) + bias[ch], 0.f);
                        }
                        *reinterpret_cast<uint4*>(dst + c * Y_CHUNK_BYTES) =
                            make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(y_full);
            mbar_wait(e_full, it & 1u);
            tc_fence_after();
            uint32_t e[3][2][16];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                tmem_ld16(tmem_e + lane_base + j * 32, e[j][0]);
                tmem_ld16(tmem_e + lane_base + j * 32 + 16, e[j][1]);
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(e_empty);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int Q = 128 * j + i, Oy = Q / YP, Ox = Q - Oy * YP;
                if (Ox < 16) {
                    const int64_t m = (int64_t)n * 256 + Oy * 16 + Ox;
                    uint4* dst = reinterpret_cast<uint4*>(p.y) + ((m >> 7) * 2) * 128 + (m & 127);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            v[k] = fmaxf(__uint_as_float(e[j][0][8 * c + k]) + __uint_as_float(e[j][1][8 * c + k]) + bias[32 + 8 * c + k], 0.f);
                        dst[c * 128] = p.out_f16 ? make_uint4(pk2<true>(v[0], v[1]), pk2<true>(v[2], v[3]), pk2<true>(v[4], v[5]), pk2<true>(v[6], v[7]))
                                                 : make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// Weight images of the two fused convs: hi|lo halves concatenated along N (see file header).
__global__ void prep_frontend_weights_kernel(const float* __restrict__ w_stem /*[27][32]*/, const float* __restrict__ w_b00 /*[288][16]*/,
                                             bf16* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= WS_ELEMS + W1_ELEMS) return;
    float v = 0.f;
    bool lo;
    if (i < WS_ELEMS) {
        const int kk = i & 7, n = (i >> 3) & 63, chunk = (i >> 9) & 1, tap = i >> 10;
        const int k = chunk * 8 + kk;                                  // s2d channel = (dy*2+dx)*3 + c
        lo = n >= 32;
        if (k < 12) {
            const int sub = k / 3, c = k % 3, dy = sub >> 1, dx = sub & 1;
            const int ky = 2 * ((tap >> 1) - 1) + dy + 1, kx = 2 * ((tap & 1) - 1) + dx + 1;
            if (ky >= 0 && ky < 3 && kx >= 0 && kx < 3) v = w_stem[((ky * 3 + kx) * 3 + c) * 32 + (n & 31)];
        }
    } else {
        const int j = i - WS_ELEMS;
        const int kk = j & 7, n = (j >> 3) & 31, chunk = (j >> 8) & 1, ks = (j >> 9) & 1, tap = j >> 10;
        const int ci = ks * 16 + chunk * 8 + kk;
        lo = n >= 16;
        v = w_b00[(tap * 32 + ci) * 16 + (n & 15)];
    }
    const bf16 hi = __float2bfloat16_rn(v);
    img[i] = lo ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
}

}  // namespace

size_t frontend_weight_image_elems() { return WS_ELEMS + W1_ELEMS; }

int launch_frontend_prep_weights(const float* blob, bf16* img, cudaStream_t s) {
    const cv_layer_info* L = cv_layers();
    prep_frontend_weights_kernel<<<(WS_ELEMS + W1_ELEMS + 255) / 256, 256, 0, s>>>(blob + L[0].w_offset, blob + L[1].w_offset, img);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_frontend(const void* src, int src_kind, int nb, int H, const CropGeom& g, const float* lut_dev, const bf16* wimg,
                    const float* bias_stem, const float* bias_b00, bf16* y, int num_sms, cudaStream_t s, const int* run_flag, const StageGate& gate) {
    if (nb == 0) return CV_OK;
    FrontParams p{src, lut_dev, wimg, bias_stem, bias_b00, y, nb * 64, H, run_flag, gate.f16 ? 1 : 0, gate.flag, gate.want};
    const CropTaps tp = make_taps(g);
    const int grid = p.n_crops < num_sms ? p.n_crops : num_sms;
#define FE_LAUNCH(KIND)                                                                                              \
    {                                                                                                                \
        CV_CUDA(cudaFuncSetAttribute(frontend_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL)); \
        frontend_kernel<KIND><<<grid, NTHREADS, SMEM_TOTAL, s>>>(p, tp);                                           \
    }
    if (src_kind == CV_SRC_U8_HWC) FE_LAUNCH(CV_SRC_U8_HWC)
    else if (src_kind == CV_SRC_U8_CHW) FE_LAUNCH(CV_SRC_U8_CHW)
    else if (src_kind == CV_SRC_F32_NCHW) FE_LAUNCH(CV_SRC_F32_NCHW)
    else { cv_set_error("frontend: bad source kind %d", src_kind); return CV_ERR_ARG; }
#undef FE_LAUNCH
    CV_CHECK_LAUNCH();
    return CV_OK;
}
