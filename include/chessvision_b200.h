/*
 * chessvision_b200 -- C-ABI of the B200-native ChessSquareCNN inference hot path.
 *
 * The reference (cloudui/chess-vision) is pure Python/PyTorch and has no FFI; the hot path is one
 * Python call, `model(images)` -> ChessSquareCNN.forward (models/square.py:92-114), followed by the
 * FEN assembly of predict.py:27-42.  This header is the boundary a maintainer binds instead of those
 * Python bodies (ctypes stub in INTEGRATION.md).  Every entry point below cites the reference code
 * it replaces.
 *
 * Conventions
 *   - plain C types only; all `const void* / void*` data pointers are DEVICE pointers owned by the
 *     caller unless the name says `host`;
 *   - every call returns CV_OK (0) or a negative cv_status; cv_last_error() gives the text of the last
 *     failure on the calling thread; nothing throws across the boundary;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue work
 *     on it -- no device synchronisation unless documented;
 *   - one handle per device; a handle is NOT re-entrant and serves ONE stream at a time (its workspace-independent scratch -- overflow and
 *     recovery flags, staging slots of the host path -- is per handle); distinct handles are independent;
 *   - device pointers of float inputs should be 16-byte aligned (cudaMalloc / torch allocations are); an unaligned x_nchw is accepted
 *     and takes a slower front end;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CV_ERR_CUDA.
 */
#ifndef CHESSVISION_B200_H
#define CHESSVISION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CV_ABI_VERSION 1

typedef enum cv_status {
    CV_OK = 0,
    CV_ERR_ARG = -1,        /* bad argument (null pointer, bad size, H not a multiple of 32, ...) */
    CV_ERR_CUDA = -2,       /* CUDA runtime / launch failure; text in cv_last_error() */
    CV_ERR_STATE = -3,      /* call order (e.g. forward before weights were loaded) */
    CV_ERR_WORKSPACE = -4   /* workspace too small */
} cv_status;

/* Arithmetic of a forward:
 *   CV_PRECISION_FP32  fp32 activations and weights, fp32 accumulation (CUDA-core kernels): the 1e-5 parity mode;
 *   CV_PRECISION_FP16  (the default of the Python surface) tensor-core path with IEEE fp16 operands, fp32 accumulation in tensor
 *                      memory, fp32 residual stream, fp32 pooled features.  Needs the fused kernels (default cv_square_set_impl mask).
 *                      fp16 overflows above 65504: weights are range-checked when loaded, activations on the device -- a wave in which
 *                      a value leaves the range is recomputed by the bf16 kernels inside the same call (no host synchronisation;
 *                      cv_square_fp16_status reports it).  Weights that do not fit make this mode identical to CV_PRECISION_BF16;
 *   CV_PRECISION_BF16  the same kernels with bf16 operands (8-bit significands, fp32 range; W = W_hi + W_lo in the early stages);
 *   CV_PRECISION_FP32_SPLIT  fp32-grade results on the tensor cores: every activation and weight is carried as fp16 hi + lo (22 significant
 *                      bits) and every GEMM k-step issues three MMAs into fp32 accumulators (A_hi W_hi + A_lo W_hi + A_hi W_lo); depthwise
 *                      convolutions and the crop gather stay fp32.  Logits within 1e-5 of the reference and identical FEN strings like
 *                      CV_PRECISION_FP32, at 77 k boards/s against 7 k.  Activations must stay below 65504 in magnitude: cv_square_fp16_status
 *                      reports an overflow of the last call (re-run it with CV_PRECISION_FP32, whose kernels have fp32 range). */
enum { CV_PRECISION_FP32 = 0, CV_PRECISION_BF16 = 1, CV_PRECISION_FP16 = 2, CV_PRECISION_FP32_SPLIT = 3 };
enum { CV_LAYOUT_HWC = 0, CV_LAYOUT_CHW = 1 };           /* uint8 board layouts: (B,H,H,3) / (B,3,H,H) */
enum { CV_DIST_UNIFORM = 0, CV_DIST_STRUCTURED = 1 };    /* synthetic board distributions */
enum { CV_KIND_DENSE = 0, CV_KIND_POINTWISE = 1, CV_KIND_DEPTHWISE = 2 };
enum { CV_FEN_STRIDE = 80 };                             /* bytes per FEN record, NUL padded (max 78) */
enum { CV_NUM_SQUARES = 64, CV_NUM_CLASSES = 13, CV_FEATURE_DIM = 480 };

typedef struct cv_square cv_square;                      /* opaque per-device model handle */

/* One conv(+folded BN)(+ReLU)(+residual) layer of the trunk (timm mobilenetv4_conv_small_050
 * forward_features as called at models/square.py:86). */
typedef struct cv_layer_info {
    int32_t kind, cin, cout, k, stride, relu, hin, hout, skip;
    int64_t w_offset, b_offset;                          /* float offsets into the weight blob */
} cv_layer_info;

/* ---- library / table queries (no GPU needed) ------------------------------------------------- */
const char* cv_last_error(void);
int         cv_abi_version(void);
int         cv_num_layers(void);                         /* 45 */
int         cv_layer_info_get(int index, cv_layer_info* out);
size_t      cv_weight_blob_floats(void);                 /* size of the packed fp32 weight blob */

/* ---- handle life cycle: replaces build_square()/build_model() module construction
 *      (models/square.py:117-138, models/__init__.py:8-30) on the device side ------------------ */
int cv_square_create(int device, cv_square** out);
int cv_square_destroy(cv_square* h);

/* Weights: `blob` is a DEVICE pointer to cv_weight_blob_floats() fp32 values laid out as documented in
 * chess_vision_b200/arch.py (BatchNorm already folded, eval mode eps=1e-5: models/square.py:83-84).
 * Replaces model.load_state_dict(ckpt["model"]) (predict.py:57) for the device copy.  Derives the bf16
 * tensor-core operand images.  Synchronises `stream` before returning. */
int cv_square_load_weights(cv_square* h, const float* blob, size_t n_floats, void* stream);

/* Host-side packer: the tensors of a reference state_dict (HOST fp32 pointers, named exactly as in the checkpoint's ckpt["model"]:
 * backbone.conv_stem.weight, backbone.bn1.{weight,bias,running_mean,running_var}, backbone.blocks.S.B. ... , type_head.1.weight, ...;
 * tensors the path never executes -- conv_head, norm_head, num_batches_tracked, class_to_* -- may be omitted) -> the packed blob in
 * blob_host (cv_weight_blob_floats() floats, HOST).  Folds eval-mode BatchNorm (eps 1e-5) in double precision exactly like
 * chess_vision_b200/weights.py: both packers emit the same bits.  Replaces build_model() + load_state_dict() (predict.py:56-57) for a
 * host that has no Python; copy blob_host to the device and hand it to cv_square_load_weights.  No GPU needed. */
typedef struct cv_named_tensor { const char* name; const float* data; int64_t numel; } cv_named_tensor;
int cv_square_pack_weights(const cv_named_tensor* tensors, int n_tensors, float* blob_host, size_t blob_floats);

/* Normalisation table lut[c*256+u] = (u/255 - mean[c]) / std[c] for the uint8 entry points; the default is
 * the timm IMAGENET mean/std the reference reads from pretrained_cfg (dataset.py:157-160).  HOST pointer,
 * 768 floats. */
int cv_square_set_norm_lut(cv_square* h, const float* lut_host);

/* Which kernels the 16-bit modes run (bit mask; default = all fused).  Clearing a bit selects a less fused kernel for that part of the
 * path -- used by the parity tests to cross-check every fused stage against layer-granular kernels.  The layer-granular kernels are
 * bf16 only, so CV_PRECISION_FP16 needs FRONTEND | EARLY | MID | TAIL.
 * CV_IMPL_POINTWISE_UMMA / CV_IMPL_DENSE_UMMA / CV_IMPL_DEPTHWISE_VEC: layer-granular tcgen05 GEMMs for the pointwise / dense 3x3
 *   convolutions and the 16-byte-vectorised depthwise kernel (cleared: plain CUDA-core kernels).
 * CV_IMPL_SPLIT_WEIGHTS: layer-granular GEMM weights held as W = W_hi + W_lo (two bf16 images, two MMAs per k-step).
 * CV_IMPL_FRONTEND: crop gather + conv_stem + blocks.0.0 as ONE persistent tcgen05 kernel (crops and stem activations never leave
 *   shared memory); CV_IMPL_FRONTEND3 (with it, uint8 HWC boards whose crop window fits shared memory: 256x256 and 512x512 do): the
 *   third-generation kernel (kernels_frontend3.cu: TMA-staged board windows, separable half2 resize, column-slab M tiles, fp16 stem
 *   operands, pipelined half images); other sources take the first generation (kernels_frontend.cu).
 * CV_IMPL_TAIL: blocks.3.* + blocks.4.0 + pool + type/color heads + combine as ONE persistent kernel (21 conv layers; activations in
 *   shared memory, residual stream in tensor memory).  CV_IMPL_MID (needs TAIL): blocks.2.* (19 conv layers) likewise.
 *   CV_IMPL_EARLY (needs MID): blocks.0.1 + blocks.1.0 + blocks.1.1 likewise (2 CTAs per SM).
 * Bits 256 and 1024 selected kernels that were retired in round 2 (second-generation front end, warp-group stage C): accepted, ignored. */
enum { CV_IMPL_POINTWISE_UMMA = 1, CV_IMPL_DENSE_UMMA = 2, CV_IMPL_DEPTHWISE_VEC = 4, CV_IMPL_SPLIT_WEIGHTS = 8,
       CV_IMPL_FRONTEND = 16, CV_IMPL_TAIL = 32, CV_IMPL_MID = 64, CV_IMPL_EARLY = 128, CV_IMPL_RETIRED_256 = 256, CV_IMPL_FRONTEND3 = 512,
       CV_IMPL_RETIRED_1024 = 1024, CV_IMPL_DEFAULT = 1023, CV_IMPL_ALL = 2047 };
int cv_square_set_impl(cv_square* h, int mask);

/* Boards per internal wave (the stage hand-offs of one wave share the workspace). 0 = library default (4096 with the fused kernels). */
int cv_square_set_wave(cv_square* h, int boards);

size_t cv_square_workspace_bytes(const cv_square* h, int max_boards, int H, int precision);

/* ChessSquareCNN.forward (models/square.py:92-114) on an already-normalised fp32 NCHW batch
 * x (B,3,H,H), H % 32 == 0 -- the tensor get_transform()'s eval branch produces (dataset.py:177-181).
 * Outputs fp32: squares (B,832) index = square*13+class; turn (B,1); castling (B,4);
 * features (B*64,480) optional (NULL to skip) = the pooled trunk output of square.py:90.
 * bf16 mode: when every value of x lies on the grid Normalize(ToTensor(u)) of a uint8 image (checked on the device, 1e-5), the call
 * computes exactly what cv_square_forward_u8 computes on those bytes (the fast front end); otherwise the floats are used as they are.
 * The handle keeps a scratch copy of one wave's bytes (B*H*H*3, allocated on first use: that first call synchronises the device). */
int cv_square_forward_f32(cv_square* h, const float* x_nchw, int B, int H, int precision,
                          float* squares, float* turn, float* castling, float* features,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Same forward from raw uint8 boards; ToTensor+Normalize (dataset.py:177-181) is fused into the crop
 * gather: value = (u8/255 - mean[c]) / std[c] in fp32, mean/std = timm IMAGENET defaults. */
int cv_square_forward_u8(cv_square* h, const uint8_t* boards, int layout, int B, int H, int precision,
                         float* squares, float* turn, float* castling, float* features,
                         void* workspace, size_t workspace_bytes, void* stream);

/* FEN assembly (predict.py:27-42 + dataset.py:52-70 labels_to_fen), batched on the device:
 * argmax over 13 (first max wins), run-length placement, turn>0 -> 'b', castling>0 -> subset of "KQkq"
 * or "-".  `flipped` (B bytes or NULL): non-zero re-indexes the 64 labels by 63-i (board rendered from
 * Black's side; datagen/render-worker.js:14-24).  fen: B records of CV_FEN_STRIDE bytes, NUL padded;
 * fen_len: B bytes. */
int cv_square_fen(const float* squares, const float* turn, const float* castling,
                  const uint8_t* flipped, int B, char* fen, uint8_t* fen_len, void* stream);

/* forward_u8 + fen in one call; squares/turn/castling scratch lives in the workspace. */
int cv_square_predict_u8(cv_square* h, const uint8_t* boards, int layout, const uint8_t* flipped,
                         int B, int H, int precision, char* fen, uint8_t* fen_len,
                         void* workspace, size_t workspace_bytes, void* stream);

/* End-to-end with HOST buffers (the call predict.py's user makes): boards_host -> chunked H2D on an
 * internal copy stream overlapped with compute -> FEN records -> D2H into fen_host/fen_len_host.
 * Blocks until the results are in host memory.  Host buffers should be pinned for full PCIe speed. */
int cv_square_predict_host_u8(cv_square* h, const uint8_t* boards_host, int layout,
                              const uint8_t* flipped_host, int B, int H, int precision,
                              char* fen_host, uint8_t* fen_len_host);

/* combine_type_color (models/common.py:10-24): joint[n][c] = type[n][T[c]] + color[n][C[c]]. */
int cv_combine_type_color(const float* type_logits, const float* color_logits, int64_t n,
                          float* joint, void* stream);

/* Crop gather alone (ChessSquareCNN._crop_squares, models/square.py:43-74), NCHW fp32 output
 * (B*64,3,64,64) like the reference, for parity tests; plus the integer source-index tables the kernel
 * uses: y0,y1 are (8,64) int32 board rows/cols after replicate-pad clamping, lam (64) fp32 blend weights. */
int cv_crop_squares_f32(const float* x_nchw, int B, int H, float* crops_nchw, void* stream);
int cv_crop_squares_u8(const uint8_t* boards, int layout, int B, int H, float* crops_nchw, void* stream);
int cv_crop_index_table(int H, int32_t* y0_host, int32_t* y1_host, float* lam_host);   /* HOST outputs, no GPU */

/* Debug tap: after layer `layer` (0..44) of the NEXT forward, its output activation of the first wave is
 * converted to fp32 NHWC and copied to `dst` (capacity n_floats).  layer < 0 clears the tap. */
int cv_square_set_tap(cv_square* h, int layer, float* dst, size_t n_floats);

/* Counter-based synthetic boards keyed by (seed, first_board + i), bit-identical to
 * chess_vision_b200/synthetic.py.  flipped (B bytes) may be NULL. */
int cv_synth_boards(uint8_t* boards, int layout, int64_t first_board, int B, int H, uint32_t seed,
                    int dist, uint8_t* flipped, void* stream);

/* Host mirrors used by the no-GPU tests (they run the same inline code as the kernels):
 * cv_synth_boards_host fills HOST memory; cv_fen_from_classes_host encodes 64 class indices + turn/castling
 * logits into one NUL-padded CV_FEN_STRIDE record and returns its length (or a negative cv_status). */
int cv_synth_boards_host(uint8_t* boards_host, int layout, int64_t first_board, int B, int H, uint32_t seed,
                         int dist, uint8_t* flipped_host);
int cv_fen_from_classes_host(const int8_t* classes, float turn, const float* castling, char* rec80);

/* Per-kernel timing: while enabled, a CUDA event is recorded on the launch stream before every kernel of the
 * path.  cv_square_profile_read synchronises the device and returns, per slot, the summed milliseconds and the
 * number of launches since the last read: slot CV_PROF_CROP, CV_PROF_LAYER0 + layer (0..44),
 * CV_PROF_POOL_HEADS, CV_PROF_GLOBAL_HEAD, CV_PROF_FEN, CV_PROF_FRONTEND (fused crop+stem+blocks.0.0).  ms/counts: HOST arrays of CV_PROF_SLOTS. */
enum { CV_PROF_CROP = 0, CV_PROF_LAYER0 = 1, CV_PROF_POOL_HEADS = 46, CV_PROF_GLOBAL_HEAD = 47, CV_PROF_FEN = 48,
       CV_PROF_FRONTEND = 49, CV_PROF_TAIL = 50, CV_PROF_MID = 51, CV_PROF_EARLY = 52,
       CV_PROF_FALLBACK = 53 /* the gated bf16 chain of a CV_PRECISION_FP16 wave: ~4 kernels that exit at once unless fp16 overflowed */,
       CV_PROF_SLOTS = 54 };
int cv_square_profile(cv_square* h, int enable);
int cv_square_profile_read(cv_square* h, double* ms, int64_t* counts);

/* Number of kernels this library has launched on behalf of `h` since creation (bench bookkeeping). */
int64_t cv_square_launch_count(const cv_square* h);

/* CV_PRECISION_FP16 bookkeeping: *weights_fit = 1 when every GEMM weight loaded into `h` fits fp16 (otherwise the mode runs the bf16
 * kernels), *overflowed = 1 when the LAST CV_PRECISION_FP16 forward on `h` had to recompute a wave with the bf16 kernels.
 * Synchronises the device (reads a device flag).  Either pointer may be NULL. */
int cv_square_fp16_status(cv_square* h, int* weights_fit, int* overflowed);

/* ---- on-device evaluation bookkeeping (replaces the per-batch host loop of evaluate.py:74-155) --------------------------
 * One batch of logits (device, fp32: squares (B,832), turn (B,1), castling (B,4)) against labels (device, uint8: class per
 * square (B,64), turn (B), castling bits (B,4), legal flag (B)) is ADDED to `counters` (device, int64[CV_EVAL_COUNTERS], the
 * caller zeroes it once): exact integer counts, laid out as below (confusion: rows = true class, evaluate.py:132-133).
 * per_sample (device, uint8 (B,4)): squares wrong, board correct, turn correct, castling-all correct (the last two are 255 for
 * positions that are not legal: the reference stores None, evaluate.py:142-143).  board_loss (device, float (B)): the sum of the
 * 64 per-square cross-entropies of each board (evaluate.py:95): loss = sum / (64 * boards). */
enum { CV_EVAL_TOTAL_BOARDS = 0, CV_EVAL_TOTAL_SQUARES = 1, CV_EVAL_CORRECT_SQUARES = 2, CV_EVAL_CORRECT_BOARDS = 3,
       CV_EVAL_TOTAL_LEGAL = 4, CV_EVAL_CORRECT_TURN = 5, CV_EVAL_CORRECT_CASTLING_RIGHT = 6 /* 4 */, CV_EVAL_CORRECT_CASTLING_ALL = 10,
       CV_EVAL_CORRECT_FULL_FEN = 11, CV_EVAL_PIECE_CORRECT = 12 /* 13 */, CV_EVAL_PIECE_TOTAL = 25 /* 13 */,
       CV_EVAL_CONFUSION = 38 /* 13 x 13 */, CV_EVAL_TURN_CONFUSION = 207 /* 2 x 2 */, CV_EVAL_COUNTERS = 211 };
int cv_eval_accumulate(const float* squares, const float* turn, const float* castling, const uint8_t* sq_labels,
                       const uint8_t* turn_labels, const uint8_t* castling_labels, const uint8_t* legal, int B,
                       int64_t* counters, uint8_t* per_sample, float* board_loss, void* stream);

/* ---- board resize of the input transform (replaces transforms.Resize((S, S)) on a PIL image, dataset.py:177-181 fed at
 * predict.py:19-20, i.e. Pillow's Image.resize((S, S), BILINEAR): src/libImaging/Resample.c) --------------------------------
 * src (device, uint8 (B, in_h, in_w, 3), decoded RGB images of one size) -> dst (device, uint8 (B, out_h, out_w, 3)), BIT-EXACT
 * with Pillow 12.2: antialiased triangle filter, 22-bit fixed-point weights, uint8 rounding between the horizontal and the vertical
 * pass; equal sizes copy.  dst is what cv_square_forward_u8 / cv_square_predict_u8 take (CV_LAYOUT_HWC): ToTensor + Normalize are
 * fused there.  cv_resize_coeffs_host (no GPU needed) returns the per-axis tables the kernel uses: *ksize taps per output index,
 * bounds_host (out_size, 2) = (first tap, tap count), coeffs_host (out_size, *ksize) int32 weights; either buffer may be NULL. */
int cv_resize_bilinear_u8(const uint8_t* src, int B, int in_h, int in_w, uint8_t* dst, int out_h, int out_w, void* stream);
int cv_resize_coeffs_host(int in_size, int out_size, int* ksize, int32_t* bounds_host, int32_t* coeffs_host, int coeffs_capacity);

/* ---- JPEG decode of the board files (replaces PIL's Image.open(path).convert("RGB"), predict.py:19 / dataset.py ChessDataset, for the
 * files the reference's datagen writes: datagen/generate.js:26-27) -----------------------------------------------------------------------
 * BIT-EXACT with Pillow 12.2 on libjpeg-turbo at libjpeg's decompression defaults: baseline / extended-sequential Huffman entropy decoding
 * with restart intervals (jdhuff.c), JDCT_ISLOW integer IDCT (jidctint.c), "fancy" triangle-filter chroma upsampling with replicated
 * edges (jdsample.c, jdmainct.c), fixed-point YCbCr -> RGB (jdcolor.c); grayscale files give R = G = B.  Handles 8-bit files with 1 or 3
 * components and 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 sampling; anything else (progressive, arithmetic, CMYK, RGB-coded) is refused with
 * CV_ERR_ARG and a reason in cv_last_error() -- the caller then decodes that file the way the reference does.
 * cv_jpeg_info: size and component count from the header (HOST pointer, no GPU).
 * cv_jpeg_decode_batch: n files of ONE size (HOST pointers) -> rgb (DEVICE, uint8 (n, height, width, 3)): the layout
 *   cv_resize_bilinear_u8 / cv_square_predict_u8 take.  By default the COMPRESSED bytes cross PCIe and the Huffman streams are walked on
 *   the device, 256-byte chunks of every file in parallel: each chunk is decoded speculatively from its first byte, re-decoded from its
 *   predecessor's exit state until the chain of states closes (a Huffman stream resynchronises by itself), then decoded once more into
 *   the coefficient buffer; a file whose chain does not close within the rounds is walked by one thread instead (same result either
 *   way).  entropy_on_host != 0 walks the streams on the host and ships coefficients (same result; the quicker way for a few files).
 *   Blocks until the pixels are in `rgb` (staging buffers are per call).
 * cv_jpeg_decode_coefficients_host: the entropy decoder alone, on the host (the same routine the device kernel runs), for no-GPU tests:
 *   quantised coefficients in natural order, component after component, [block rows][block columns][64]; block_grid (3 x (rows, cols)).
 * cv_jpeg_decode_coefficients_host_chunked: the same coefficients through the chunked scheme of the device path, run on the host round by
 *   round (no-GPU tests of the scheme itself); stats (nullable, int32[4]) = chunks, chunk decodes over all rounds, intervals that fell back
 *   to the serial walk, last round that changed a state. */
int cv_jpeg_info(const uint8_t* file_host, size_t size, int* width, int* height, int* components);
int cv_jpeg_decode_batch(const uint8_t* const* files_host, const size_t* sizes, int n, int width, int height, uint8_t* rgb,
                         int entropy_on_host, void* stream);
int cv_jpeg_decode_coefficients_host(const uint8_t* file_host, size_t size, int16_t* coef_host, size_t capacity, int32_t* block_grid);
int cv_jpeg_decode_coefficients_host_chunked(const uint8_t* file_host, size_t size, int16_t* coef_host, size_t capacity, int chunk_bytes,
                                             int rounds, int32_t* stats);

#ifdef __cplusplus
}
#endif
#endif /* CHESSVISION_B200_H */
