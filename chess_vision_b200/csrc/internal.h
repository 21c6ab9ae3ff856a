// Internal declarations shared by the translation units of libchessvision_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/chessvision_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing ---------------------------------------------------------------------------
void cv_set_error(const char* fmt, ...);
#define CV_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            cv_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return CV_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)
#define CV_CHECK_LAUNCH() CV_CUDA(cudaGetLastError())
#define CV_ARG(cond, msg)                                     \
    do {                                                      \
        if (!(cond)) {                                        \
            cv_set_error("%s: %s", __func__, msg);            \
            return CV_ERR_ARG;                                \
        }                                                     \
    } while (0)

extern const cv_layer_info* cv_layers();   // the compiled-in table (arch_table.inc)

// ---- activation layouts ------------------------------------------------------------------------------
// RowMajorL: NHWC, element (pixel-row m, channel ch) at m*C + ch.                       (fp32 path, crops)
// T8L: "tiled-K8" -- rows grouped in tiles of 128, channels in chunks of 8 (16 bytes of bf16):
//      offset = (((m/128)*(C/8) + ch/8)*128 + m%128)*8 + ch%8.
//      One 128-row tile of a C-channel tensor is a contiguous 128*C*2-byte block that IS the shared-memory
//      image of a K-major, no-swizzle UMMA A operand ([C/8] core-matrix columns x [128 rows] x 16 B), so a
//      GEMM layer loads it with ONE TMA bulk copy, and epilogues store 16-byte chunks fully coalesced.
struct RowMajorL {
    __host__ __device__ static inline int64_t off(int64_t m, int ch, int C) { return m * C + ch; }
};
struct T8L {
    __host__ __device__ static inline int64_t off(int64_t m, int ch, int C) {
        return ((((m >> 7) * (C >> 3) + (ch >> 3)) << 7) + (m & 127)) * 8 + (ch & 7);
    }
};

// ---- crop geometry (models/square.py:53-55) + bilinear taps (ATen upsample_bilinear2d,
//      align_corners=False), computed once on the host and passed by value to the kernels ----------
struct CropGeom {
    int H, sq, crop, pad;
    int16_t i0[64], i1[64];   // crop-space taps per output pixel
    float lam[64];
};
int cv_make_crop_geom(int H, CropGeom* g);

// Per-axis bilinear taps in BOARD space (after replicate-pad clamping) for the 8 squares of a rank/file:
// passed to the crop kernels by value as a __grid_constant__ parameter (2.3 KB).
struct CropTaps {
    int16_t p0[8][64], p1[8][64];
    float lam[64];
};
inline CropTaps make_taps(const CropGeom& g) {
    CropTaps t;
    for (int r = 0; r < 8; ++r)
        for (int d = 0; d < 64; ++d) {
            int a = r * g.sq + g.i0[d] - g.pad, b = r * g.sq + g.i1[d] - g.pad;
            a = a < 0 ? 0 : (a > g.H - 1 ? g.H - 1 : a);
            b = b < 0 ? 0 : (b > g.H - 1 ? g.H - 1 : b);
            t.p0[r][d] = (int16_t)a; t.p1[r][d] = (int16_t)b;
        }
    for (int d = 0; d < 64; ++d) t.lam[d] = g.lam[d];
    return t;
}
#ifdef __CUDACC__
// Blend order pinned to the oracle: (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11), no FMA
// contraction, so fp32 crops are bit-identical to the CPU restatement.
__device__ __forceinline__ float crop_blend(float v00, float v01, float v10, float v11, float lx, float ly) {
    float wx0 = __fsub_rn(1.0f, lx), wy0 = __fsub_rn(1.0f, ly);
    float top = __fadd_rn(__fmul_rn(wx0, v00), __fmul_rn(lx, v01));
    float bot = __fadd_rn(__fmul_rn(wx0, v10), __fmul_rn(lx, v11));
    return __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly, bot));
}
#endif

// ---- generic (precision-templated) kernels: kernels_generic.cu ----------------------------------
// All activations NHWC; T = float (CV_PRECISION_FP32) or bf16; accumulation always fp32.
template <typename T>
int launch_crop_f32(const float* x_nchw, int B, int H, const CropGeom& g, T* out_nhwc, float* out_nchw,
                    cudaStream_t s);
template <typename T>
int launch_crop_u8(const uint8_t* boards, int layout, int B, int H, const CropGeom& g, const float* lut_dev,
                   T* out_nhwc, float* out_nchw, cudaStream_t s);
template <typename T>
int launch_conv_generic(const cv_layer_info& L, const T* in, const float* w, const float* bias, const T* skip,
                        T* out, int64_t n_crops, bool in_t8, bool out_t8, cudaStream_t s);
template <typename T>
int launch_depthwise_generic(const cv_layer_info& L, const T* in, const float* w, const float* bias, T* out,
                             int64_t n_crops, bool t8, cudaStream_t s);
template <typename T>
int launch_pool_heads(const T* feat_map /*[N,2,2,480]*/, const float* head_w, const float* head_b,
                      int64_t n_crops, float* features /*[N,480]*/, float* squares, bool t8, cudaStream_t s);
int launch_global_head(const float* features /*[B,30720]*/, const float* glob_wt /*[30720,64]*/,
                       const float* glob_b, const float* tc_w, const float* tc_b, int B, float* turn,
                       float* castling, bool exact_fp64_accumulate, cudaStream_t s);
size_t global_head_f64_partial_bytes(int B);            // the same fp64 head with the K reduction spread over the grid (deterministic partial sums)
int launch_global_head_f64_split(const float* features, const float* glob_wt, const float* glob_b, const float* tc_w, const float* tc_b, int B,
                                 double* partial, float* turn, float* castling, cudaStream_t s);
template <typename T>
int launch_to_f32(const T* src, float* dst, size_t n, int C, bool t8, cudaStream_t s);   // -> row-major fp32
int launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t s);   // dst[c][r]=src[r][c]

// ---- fen.cu / synth.cu ---------------------------------------------------------------------------
int launch_fen(const float* squares, const float* turn, const float* castling, const uint8_t* flipped, int B,
               char* fen, uint8_t* fen_len, cudaStream_t s);
int launch_synth(uint8_t* boards, int layout, int64_t first_board, int B, int H, uint32_t seed, int dist,
                 uint8_t* flipped, cudaStream_t s);
int launch_combine(const float* t, const float* c, int64_t n, float* joint, cudaStream_t s);

// ---- kernels_umma.cu: bf16 tensor-core path (tcgen05 / TMEM / TMA bulk copies), T8 activations ------------
size_t umma_weight_image_elems();                      // bf16 elements of all GEMM-layer weight images
int64_t umma_weight_image_offset(int layer);           // element offset of a layer's image (-1: not a GEMM layer)
int launch_umma_prep_weights(const float* blob, bf16* wimg, cudaStream_t s);
int launch_pointwise_umma(const cv_layer_info& L, const bf16* x, const bf16* wimg, const float* bias, const bf16* skip,
                          bf16* y, int64_t n_crops, int num_sms, bool split_weights, cudaStream_t s);
int launch_dense_umma(const cv_layer_info& L, const bf16* x, bool in_rowmajor3, const bf16* wimg, const float* bias,
                      bf16* y, int64_t n_crops, int num_sms, bool split_weights, cudaStream_t s);
int launch_depthwise_t8(const cv_layer_info& L, const bf16* x, const float* w, const float* bias, bf16* y,
                        int64_t n_crops, cudaStream_t s);

// ---- operand format and gating of the fused kernels -------------------------------------------------------------------------
// The 16-bit modes run the same fused kernels in one of two operand formats: fp16 (CV_PRECISION_FP16: 11-bit significands, one
// weight image, half the MMAs) or bf16 (W = W_hi + W_lo in stages B / C).  fp16 can overflow (|v| > 65504): the fp16 kernels raise
// `ovf` when a non-finite value reaches a residual-stream copy, and a forward in fp16 mode enqueues BOTH chains per wave -- the
// fp16 kernels gated on *flag == 0, then the bf16 kernels gated on *flag != 0 -- so the fall-back needs no host synchronisation
// (a gated-off persistent kernel exits in its first instruction).
struct StageGate {
    bool f16 = false;             // operand format of this launch
    const int* flag = nullptr;    // null: run unconditionally
    int want = 0;                 // run iff (*flag != 0) == want
    int* ovf = nullptr;           // fp16 launches: the overflow flag to raise (== flag)
};

// ---- kernels_frontend.cu: fused crop gather + conv_stem + blocks.0.0 (tcgen05), output T8 [crops*256][16] ----
enum { CV_SRC_U8_HWC = 0, CV_SRC_U8_CHW = 1, CV_SRC_F32_NCHW = 2 };
size_t frontend_weight_image_elems();
int launch_frontend_prep_weights(const float* blob, bf16* img, cudaStream_t s);
int launch_frontend(const void* src, int src_kind, int nb, int H, const CropGeom& g, const float* lut_dev,
                    const bf16* wimg, const float* bias_stem, const float* bias_b00, bf16* y, int num_sms,
                    cudaStream_t s, const int* run_flag = nullptr, const StageGate& gate = StageGate());

// ---- kernels_backend.cu: fused back-end stages (persistent tcgen05 kernels, activations in smem / TMEM) -----------------
// stage D = blocks.3.* + blocks.4.0 + average pool + type/color heads + combine.  Input: "P8" tiles (128 rows = 8 crops,
// row = pixel*8 + crop_local, 48 channels, T8 chunking); output: features [crops][480] fp32, squares [crops][13] fp32.
enum { CV_STAGE_D_OPS = 24 };
// Stages C and D stream their weights L2 -> shared memory once per tile, every CTA at about the same time: their images are kept
// in CV_W_REPLICAS copies (stageX_image_bytes() apart) and CTA b reads copy b % replicas, which spreads the requests over the L2.
enum { CV_W_REPLICAS = 8 };
int weight_replicas();                      // how many of the copies the kernels use (CV_W_REPLICAS; experiment builds: CV_W_REP)
size_t stageD_image_bytes();
int build_stageD_image(const float* blob, uint8_t* img, uint32_t* off, uint32_t* bytes, int* f16_flag /* non-null: fp16 images */, cudaStream_t s);
int launch_permute_p8(const bf16* in_t8, bf16* out_p8, int64_t n_crops, int C, cudaStream_t s);
// features: row-major [crops][480] (tiled == 0, indexed from this launch's first crop) or the FT operand layout of the
// tensor-core global head (tiled == 1: `features` is the chunk's FT base and crop_base the launch's first crop in the chunk).
int launch_stageD(const bf16* x_p8, int64_t n_crops, const uint8_t* wimg, const uint32_t* off, const uint32_t* bytes,
                  float* features, int tiled, int64_t crop_base, float* squares, int num_sms, const StageGate& gate,
                  int perm_boards /* > 0: input in the permuted crop order stage C wrote for a launch of this many boards */, cudaStream_t s);

// stage C = blocks.2.* (19 conv layers).  Input: "P2" tiles (128 rows = 2 crops at 8x8, row = pixel*2 + crop_local, 32 ch);
// output: the P8 tiles stage D consumes.
enum { CV_STAGE_C_OPS = 20 };
size_t stageC_image_bytes();
int build_stageC_image(const float* blob, uint8_t* img, uint32_t* off, uint32_t* bytes, int* f16_flag, cudaStream_t s);
int launch_permute_p2(const bf16* in_t8, bf16* out_p2, int64_t n_crops, int C, cudaStream_t s);
int launch_stageC(const bf16* x_p2, int64_t n_crops, const uint8_t* wimg, const uint32_t* off, const uint32_t* bytes, bf16* y_p8,
                  int num_sms, const StageGate& gate, int perm_boards, cudaStream_t s);

// stage B = blocks.0.1 + blocks.1.0 + blocks.1.1.  Input: the front end's T8 output (rows = crop*256 + pixel, 16 ch); output: P2 tiles.
size_t stageB_image_bytes();
int build_stageB_image(const float* blob, uint8_t* img, int* f16_flag, cudaStream_t s);
int launch_stageB(const bf16* x_t8, int64_t n_crops, const uint8_t* wimg, bf16* y_p2, int num_sms, const StageGate& gate, cudaStream_t s);

// ---- kernels_head.cu: global_head as a split-K tcgen05 (kind::tf32) GEMM over FT-tiled features -----------------------------
//   FT[m_tile][plane][k/4][128 boards][4]: float index (((b/128 * 2 + plane) * 7680 + k/4) * 128 + b%128) * 4 + k%4,  k = square*480 + channel;
//   plane 0 = x & 0xffffe000 (what the tensor core reads of a tf32 operand), plane 1 = the exact remainder
inline size_t ft_floats(int boards) { return (size_t)((boards + 127) / 128) * 128 * 30720 * 2; }
int launch_tile_glob_w(const float* glob_w, float* wt, cudaStream_t s);
size_t global_head_partial_floats(int B, int num_sms);
int launch_global_head_umma(const float* ft, const float* wt, float* partial, const float* glob_b, const float* tc_w, const float* tc_b,
                            int B, int num_sms, float* turn, float* castling, cudaStream_t s);
int launch_untile_features(const float* ft, float* out_rowmajor, int B, cudaStream_t s);

// ---- kernels_frontend3.cu: third generation (column-slab M tiles, fp16 stem operands, pipelined half images) ---------------
// out_f16: the stem output image, the blocks.0.0 weights and the T8 output are fp16 (bf16 otherwise); the weight image holds both
// blocks.0.0 variants.  skip_flag / gate: the kernel exits at once when *skip_flag != 0, or when gate says so (StageGate).
size_t frontend3_weight_image_bytes();
int launch_frontend3_prep_weights(const float* blob, uint8_t* img, int* flag_dev, cudaStream_t s);
int launch_frontend3(const uint8_t* boards_hwc, int nb, int H, const CropGeom& g, const float* lut_host, const uint8_t* wimg,
                     const float* bias_b00, bf16* y, int num_sms, int* supported, cudaStream_t s, const int* skip_flag = nullptr,
                     const StageGate& gate = StageGate(),
                     int perm_boards = 0 /* > 0: write the crops in the permuted order of a launch of this many boards (umma.cuh perm_pos) */);
// float NCHW boards that are Normalize(ToTensor(uint8)) -> the uint8 HWC image + a device flag (1 = some value is not on the uint8 grid)
int launch_f32_to_u8_boards(const float* x_nchw, int nb, int H, const float* lut_host, uint8_t* out_hwc, int* flag_dev, cudaStream_t s);

// ---- kernels_exact.cu: fp32-grade layer-granular trunk on the tensor cores, split fp16 operands (CV_PRECISION_FP32_SPLIT) --------------
//   X2 layout of a C-channel tensor = T8 layout of 2C channels: per 128-row tile C/8 chunks of hi = fp16(x), then C/8 chunks of lo = fp16(x - hi)
size_t x2_weight_image_elems();
int64_t x2_weight_image_offset(int layer);
int launch_x2_prep_weights(const float* blob, uint16_t* wimg, float* unscale_host /*[cv_num_layers()]*/, cudaStream_t s);
int launch_pointwise_x2(const cv_layer_info& L, const uint16_t* x, const uint16_t* wimg, const float* bias, float unscale, const uint16_t* skip,
                        uint16_t* y, int64_t n_crops, int num_sms, int* ovf, cudaStream_t s);
int launch_dense_x2(const cv_layer_info& L, const uint16_t* x, const float* x_f32_crops, const uint8_t* boards_hwc, int H, const float* lut_dev,
                    const CropTaps* taps, const uint16_t* wimg, const float* bias, float unscale, uint16_t* y, int64_t n_crops, int num_sms, int* ovf,
                    cudaStream_t s);
int launch_depthwise_x2(const cv_layer_info& L, const uint16_t* x, const float* w, const float* bias, uint16_t* y, int64_t n_crops, int* ovf,
                        cudaStream_t s);
int launch_pool_heads_x2(const uint16_t* fmap, const float* head_w, const float* head_b, int64_t n_crops, float* features, float* squares,
                         cudaStream_t s);
int launch_x2_to_f32(const uint16_t* src, float* dst, size_t n, int C, cudaStream_t s);
