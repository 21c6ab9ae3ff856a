// C-ABI of libchessvision_b200.so: handle life cycle, weight loading, wave scheduler, forward / predict
// entry points (include/chessvision_b200.h).  Host-side orchestration only; kernels live in the other
// translation units.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

#include "internal.h"
#include "arch_table.inc"

static thread_local char g_err[1024] = "";

void cv_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const cv_layer_info* cv_layers() { return kLayers; }

namespace {

constexpr int NBUF_SMALL = 4;
constexpr int64_t EL_CROPS = 64 * 64 * 3, EL_STEM = 32 * 32 * 32, EL_SMALL = 6144;
constexpr int DEFAULT_WAVE_FP32 = 64, DEFAULT_WAVE_BF16 = 128, DEFAULT_WAVE_FUSED = 512, DEFAULT_WAVE_SPLIT = 1024;   // fused stages: persistent kernels want many tiles per SM
// all four fused stages: only the 8 / 4 / 1.5 KB per crop hand-offs live in the workspace, and every kernel boundary costs ~13 us of
// drained SMs, so a wave is a whole chunk (measured per 4096 boards: 13.81 ms at 512, 13.64 at 1024, 13.55 at 2048, 13.45 at 4096)
constexpr int DEFAULT_WAVE_ALL_FUSED = 4096;
constexpr int MAX_CHUNK = 4096;      // boards whose pooled features are kept for one global-head launch

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

}  // namespace

struct cv_square {
    int device = 0;
    bool loaded = false;
    float* blob = nullptr;        // fp32 packed weights (device, owned)
    float* glob_wt = nullptr;     // global_head weight transposed to [30720][64] (device, owned)
    float* glob_wtile = nullptr;  // global_head weight in the tf32 UMMA operand layout [hi | lo][30720/4][64][4] (kernels_head.cu)
    float* head_w = nullptr;      // aligned copies of the small heads: head_w[10*480], head_b[10], glob_b[64], tc_w[320], tc_b[5]
    float* lut = nullptr;         // normalisation LUT [3][256] (device, owned)
    bf16* wimg = nullptr;         // bf16 UMMA weight images of the 30 GEMM layers (device, owned)
    bf16* fe_wimg = nullptr;      // hi|lo weight images of the fused front end (conv_stem + blocks.0.0)
    uint8_t* fe3_wimg = nullptr;  // third generation: fp16 stem image + blocks.0.0 images (bf16 hi|lo, fp16), then an int "weights exceed fp16" flag
    bool fe3_ok = false;          // the stem / blocks.0.0 weights fit fp16 (checked when the weights are packed)
    float lut_host[768];          // host copy of the normalisation table (affinity check of the third-generation front end)
    // weight images of the fused stages, index 0 = bf16 (stages B / C: W_hi | W_lo), 1 = fp16
    uint8_t* sd_img[2] = {};      // fused tail (stage D)
    uint32_t sd_off[2][CV_STAGE_D_OPS], sd_bytes[2][CV_STAGE_D_OPS];
    uint8_t* sb_img[2] = {};      // fused blocks.0.1 + blocks.1 (stage B)
    uint8_t* sc_img[2] = {};      // fused blocks.2 (stage C)
    uint32_t sc_off[2][CV_STAGE_C_OPS], sc_bytes[2][CV_STAGE_C_OPS];
    int* flags = nullptr;         // device ints: [0] a weight of stages B-D does not fit fp16, [1] fp16 overflow of the current call (StageGate),
                                  // [2] float entry point: "the input is not a uint8 image"
    bool f16_ok = false;          // every GEMM weight of the fused stages fits fp16 (checked when the weights are packed)
    uint16_t* x2_wimg = nullptr;  // split-fp16 weight images of the 30 GEMM layers (kernels_exact.cu): [K/8][hi N | lo N][8], scaled by 2^s
    float x2_unscale[CV_NUM_LAYERS];   // 2^-s per layer
    int num_sms = 148;
    int impl = CV_IMPL_DEFAULT;   // which bf16 kernels run (cv_square_set_impl)
    int wave = 0;                 // boards per wave, 0 = default per precision
    int tap_layer = -1;
    float* tap_dst = nullptr;
    size_t tap_n = 0;
    int64_t launches = 0;
    bool profiling = false;                 // CUDA-event marks before every launch (cv_square_profile)
    bool prof_suspended = false;            // inside the gated bf16 fall-back pass: its launches are timed as ONE slot (CV_PROF_FALLBACK)
    std::vector<cudaEvent_t> prof_pool;
    std::vector<int> prof_slot;             // slot of the launch that FOLLOWS mark i (-1 = end of a call)
    int out_buf[CV_NUM_LAYERS];   // small-buffer index each layer writes (-1: dedicated stem buffer)
    // host-pointer pipeline resources (lazily created)
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    static constexpr int kStages = 3;     // staging slots of the host path: the copy stream may run two chunks ahead of the compute stream
    cudaEvent_t ev_h2d[kStages] = {}, ev_done[kStages] = {};
    static constexpr int kPieces = 4;     // a chunk is copied in up to kPieces pieces: the front end starts on a piece as soon as it has landed
    cudaEvent_t ev_piece[kStages][kPieces] = {};
    const cudaEvent_t* piece_ev = nullptr;  // set by the host path around one forward: piece i (piece_boards boards) is ready after piece_ev[i]
    int piece_boards = 0, n_pieces = 0;
    void* stage[kStages] = {};
    uint8_t* stage_flip[kStages] = {};
    size_t stage_bytes = 0;
    int stage_boards = 0;                 // boards a staging slot (and its flip buffer) holds
    void* own_ws = nullptr;
    size_t own_ws_bytes = 0;
    char* dev_fen = nullptr;
    uint8_t* dev_fen_len = nullptr;
    int dev_fen_cap = 0;
};

namespace {

// Greedy assignment of the rotating small activation buffers: a layer's output stays live until its last
// reader (the next layer, or the pw_proj that adds it back as a residual).
void plan_buffers(cv_square* h) {
    int last_use[CV_NUM_LAYERS];
    for (int i = 0; i < CV_NUM_LAYERS; ++i) last_use[i] = i + 1;
    for (int i = 0; i < CV_NUM_LAYERS; ++i)
        if (kLayers[i].skip >= 0) last_use[kLayers[i].skip] = std::max(last_use[kLayers[i].skip], i);
    int owner[NBUF_SMALL];
    for (int b = 0; b < NBUF_SMALL; ++b) owner[b] = -1;
    h->out_buf[0] = -1;
    for (int i = 1; i < CV_NUM_LAYERS; ++i) {
        int pick = -1;
        for (int b = 0; b < NBUF_SMALL && pick < 0; ++b)
            if (owner[b] < 0 || last_use[owner[b]] < i) pick = b;     // free once every reader has run
        h->out_buf[i] = pick;                                           // never -1: 4 buffers suffice (checked at create)
        if (pick >= 0) owner[pick] = i;
    }
}

// Profiling mark: an event recorded on the launch stream right before the kernel(s) of `slot`.
inline int prof_mark(cv_square* h, int slot, cudaStream_t s) {
    if (!h->profiling || h->prof_suspended) return CV_OK;
    size_t i = h->prof_slot.size();
    if (i >= h->prof_pool.size()) {
        cudaEvent_t e;
        CV_CUDA(cudaEventCreate(&e));
        h->prof_pool.push_back(e);
    }
    CV_CUDA(cudaEventRecord(h->prof_pool[i], s));
    h->prof_slot.push_back(slot);
    return CV_OK;
}

struct WavePlan {
    int wave;
    size_t es;                        // activation element size
    size_t off_crops, off_stem, off_small[NBUF_SMALL], off_feat, off_partial, off_sq, off_turn, off_cast, off_f2u, f2u_bytes, total;
};

// H > 0 adds the scratch of the float entry point in the 16-bit modes (the uint8 image recovered from Normalize(ToTensor(uint8))
// inputs, one wave of H x H x 3 bytes per board): it lives in the caller's workspace, so a forward never allocates.
WavePlan make_plan(const cv_square* h, int max_boards, int precision, int H) {
    WavePlan p;
    const bool bf = precision == CV_PRECISION_BF16 || precision == CV_PRECISION_FP16;
    const int all_fused = CV_IMPL_FRONTEND | CV_IMPL_EARLY | CV_IMPL_MID | CV_IMPL_TAIL;
    int def = precision == CV_PRECISION_FP32_SPLIT ? DEFAULT_WAVE_SPLIT : !bf ? DEFAULT_WAVE_FP32 : (h->impl & all_fused) == all_fused ? DEFAULT_WAVE_ALL_FUSED : (h->impl & CV_IMPL_TAIL) ? DEFAULT_WAVE_FUSED : DEFAULT_WAVE_BF16;
    p.wave = h->wave > 0 ? h->wave : def;
    if (p.wave > max_boards) p.wave = std::max(max_boards, 1);
    p.es = bf ? 2 : 4;
    const size_t n = (size_t)p.wave * 64;
    const bool fused_front = bf && (h->impl & CV_IMPL_FRONTEND);    // crops and stem output never reach HBM
    size_t off = 0;
    p.off_crops = off; off = align_up(off + (fused_front ? 0 : n * EL_CROPS * p.es));
    p.off_stem = off; off = align_up(off + (fused_front ? 0 : n * EL_STEM * p.es));
    for (int b = 0; b < NBUF_SMALL; ++b) { p.off_small[b] = off; off = align_up(off + n * EL_SMALL * p.es); }
    const size_t chunk = (size_t)std::min(std::max(max_boards, 1), MAX_CHUNK);
    // pooled features of one chunk: row-major, or FT-tiled (whole 128-board tiles) for the tensor-core global head
    const bool tiled = bf && (h->impl & CV_IMPL_TAIL);
    p.off_feat = off; off = align_up(off + (tiled ? ft_floats((int)chunk) : chunk * 64 * 480) * sizeof(float));
    p.off_partial = off; off = align_up(off + (tiled ? global_head_partial_floats((int)chunk, h->num_sms) * sizeof(float)
                                                     : !bf ? global_head_f64_partial_bytes((int)chunk) : 0));
    // scratch logits for the predict entry points (whole batch)
    p.off_sq = off; off = align_up(off + (size_t)max_boards * 832 * sizeof(float));
    p.off_turn = off; off = align_up(off + (size_t)max_boards * sizeof(float));
    p.off_cast = off; off = align_up(off + (size_t)max_boards * 4 * sizeof(float));
    p.f2u_bytes = (bf && (h->impl & CV_IMPL_FRONTEND) && (h->impl & CV_IMPL_FRONTEND3) && H > 0) ? (size_t)p.wave * H * H * 3 : 0;
    p.off_f2u = off; off = align_up(off + p.f2u_bytes);
    p.total = off;
    return p;
}

template <typename T>
int run_layer(cv_square* h, int i, const T* in, const T* skip, T* out, int64_t n, bool t8, cudaStream_t s);

template <>
int run_layer<float>(cv_square* h, int i, const float* in, const float* skip, float* out, int64_t n, bool, cudaStream_t s) {
    const cv_layer_info& L = kLayers[i];
    const float* w = h->blob + L.w_offset;
    const float* b = h->blob + L.b_offset;
    if (L.kind == CV_KIND_DEPTHWISE) return launch_depthwise_generic<float>(L, in, w, b, out, n, false, s);
    return launch_conv_generic<float>(L, in, w, b, skip, out, n, false, false, s);
}

// bf16 mode: T8 activations everywhere except the 3-channel crops feeding the stem.
template <>
int run_layer<bf16>(cv_square* h, int i, const bf16* in, const bf16* skip, bf16* out, int64_t n, bool, cudaStream_t s) {
    const cv_layer_info& L = kLayers[i];
    const float* w = h->blob + L.w_offset;
    const float* b = h->blob + L.b_offset;
    const bool in_t8 = i > 0;
    if (L.kind == CV_KIND_DEPTHWISE) {
        if (h->impl & CV_IMPL_DEPTHWISE_VEC) return launch_depthwise_t8(L, in, w, b, out, n, s);
        return launch_depthwise_generic<bf16>(L, in, w, b, out, n, true, s);
    }
    const bf16* wi = h->wimg + umma_weight_image_offset(i);
    if (L.kind == CV_KIND_POINTWISE && (h->impl & CV_IMPL_POINTWISE_UMMA))
        return launch_pointwise_umma(L, in, wi, b, skip, out, n, h->num_sms, (h->impl & CV_IMPL_SPLIT_WEIGHTS) != 0, s);
    if (L.kind == CV_KIND_DENSE && (h->impl & CV_IMPL_DENSE_UMMA))
        return launch_dense_umma(L, in, !in_t8, wi, b, out, n, h->num_sms, (h->impl & CV_IMPL_SPLIT_WEIGHTS) != 0, s);
    return launch_conv_generic<bf16>(L, in, w, b, skip, out, n, in_t8, true, s);
}

template <typename T>
int run_wave(cv_square* h, const WavePlan& p, char* ws, int nb, float* squares, float* feat, float* feat_chunk, int64_t crop_base,
             bool first_wave, int first_layer, const StageGate& gate, bool fe_permuted, cudaStream_t s) {
    const int fi = gate.f16 ? 1 : 0;                   // which weight images the fused stages read
    const bool t8 = sizeof(T) == 2;
    const int64_t n = (int64_t)nb * 64;
    T* crops = reinterpret_cast<T*>(ws + p.off_crops);
    T* stem = reinterpret_cast<T*>(ws + p.off_stem);
    T* small[NBUF_SMALL];
    for (int b = 0; b < NBUF_SMALL; ++b) small[b] = reinterpret_cast<T*>(ws + p.off_small[b]);
    auto buf_of = [&](int layer) -> T* { return layer < 0 ? crops : (h->out_buf[layer] < 0 ? stem : small[h->out_buf[layer]]); };
    // bf16: blocks.3.* + blocks.4.0 + pool + heads fused into one persistent kernel (stage D) fed by layer 23's output
    const bool fused_tail = sizeof(T) == 2 && (h->impl & CV_IMPL_TAIL);
    const bool fused_mid = fused_tail && (h->impl & CV_IMPL_MID);          // blocks.2.* as one kernel too (stage C)
    const bool fused_early = fused_mid && (h->impl & CV_IMPL_EARLY) && first_layer == 2;   // blocks.0.1 + blocks.1.* (stage B) after the fused front end
    const int end_layer = fused_early ? 2 : fused_mid ? 5 : fused_tail ? 24 : CV_NUM_LAYERS;
    for (int i = 0; i < end_layer; ++i) {
        const cv_layer_info& L = kLayers[i];
        T* out = buf_of(i);
        if (i >= first_layer) {
            int rc = prof_mark(h, CV_PROF_LAYER0 + i, s);
            if (rc) return rc;
            rc = run_layer<T>(h, i, buf_of(i - 1), L.skip >= 0 ? buf_of(L.skip) : nullptr, out, n, t8, s);
            if (rc) return rc;
            ++h->launches;
        }
        int rc;
        if (i >= first_layer - 1 && first_wave && h->tap_layer == i && h->tap_dst) {
            size_t cnt = std::min(h->tap_n, (size_t)n * L.hout * L.hout * L.cout);
            rc = launch_to_f32<T>(out, h->tap_dst, cnt, L.cout, t8, s);
            if (rc) return rc;
        }
    }
    if (fused_tail) {
        int rc;
        T* p8;
        if (fused_mid) {
            T* p2;
            if (fused_early) {
                rc = prof_mark(h, CV_PROF_EARLY, s);
                if (rc) return rc;
                p2 = small[(h->out_buf[1] + 1) % NBUF_SMALL];               // any two buffers but layer 1's own
                p8 = small[(h->out_buf[1] + 2) % NBUF_SMALL];
                rc = launch_stageB(reinterpret_cast<const bf16*>(buf_of(1)), n, h->sb_img[fi], reinterpret_cast<bf16*>(p2), h->num_sms, gate, s);
                if (rc) return rc;
                ++h->launches;
                rc = prof_mark(h, CV_PROF_MID, s);
                if (rc) return rc;
            } else {
                rc = prof_mark(h, CV_PROF_MID, s);
                if (rc) return rc;
                p2 = small[(h->out_buf[4] + 1) % NBUF_SMALL];               // any two buffers but layer 4's own
                p8 = small[(h->out_buf[4] + 2) % NBUF_SMALL];
                rc = launch_permute_p2(reinterpret_cast<const bf16*>(buf_of(4)), reinterpret_cast<bf16*>(p2), n, 32, s);
                if (rc) return rc;
                ++h->launches;
            }
            rc = launch_stageC(reinterpret_cast<const bf16*>(p2), n, h->sc_img[fi], h->sc_off[fi], h->sc_bytes[fi], reinterpret_cast<bf16*>(p8),
                               h->num_sms, gate, fe_permuted ? 0 : nb /* the crops arrive in board order: stage C permutes its hand-off */, s);
            if (rc) return rc;
            ++h->launches;
            rc = prof_mark(h, CV_PROF_TAIL, s);
            if (rc) return rc;
        } else {
            rc = prof_mark(h, CV_PROF_TAIL, s);
            if (rc) return rc;
            p8 = small[(h->out_buf[23] + 1) % NBUF_SMALL];                 // any buffer but layer 23's own
            rc = launch_permute_p8(reinterpret_cast<const bf16*>(buf_of(23)), reinterpret_cast<bf16*>(p8), n, 48, s);
            if (rc) return rc;
            ++h->launches;
        }
        rc = launch_stageD(reinterpret_cast<const bf16*>(p8), n, h->sd_img[fi], h->sd_off[fi], h->sd_bytes[fi], feat_chunk, 1, crop_base, squares,
                           h->num_sms, gate, fused_mid ? nb : 0, s);
        if (rc) return rc;
        ++h->launches;
        return CV_OK;
    }
    int rc = prof_mark(h, CV_PROF_POOL_HEADS, s);
    if (rc) return rc;
    rc = launch_pool_heads<T>(buf_of(CV_NUM_LAYERS - 1), h->head_w, h->head_w + 4800, n, feat, squares, t8, s);
    if (rc) return rc;
    ++h->launches;
    return CV_OK;
}

// One front-end launch sequence for `nb` boards of a wave in the 16-bit modes (crop gather + conv_stem + blocks.0.0 -> front_out):
// third generation for uint8 HWC boards (piece by piece behind the host path's copies when `by_pieces`), first generation for the
// other sources and for windows the third cannot stage; float sources that are uint8 images in disguise take the third generation
// on the recovered bytes (f2u: recovered by the caller once per wave; its device flag picks exactly one of the two kernels).
// permute: the third-generation kernel may write the crops in the pipeline's permuted order (umma.cuh perm_pos; *permuted says whether
// it did -- only uint8 HWC sources it supports, and pieces that are whole groups of 32 boards).
int launch_front(cv_square* h, const void* src, int kind, int nb, int H, const CropGeom& g, bf16* front_out, bool by_pieces,
                 const uint8_t* f2u, const int* f2u_flag, const StageGate& gate, bool permute, bool* permuted, cudaStream_t s) {
    int rc = CV_OK, done = 0;
    *permuted = false;
    if (by_pieces && h->piece_boards % 32 != 0 && h->piece_boards < nb) permute = false;
    const bool v3 = (h->impl & CV_IMPL_FRONTEND3) && h->fe3_ok;
    const float* bias_b00 = h->blob + kLayers[1].b_offset;
    if (kind == CV_SRC_U8_HWC && v3) {
        if (by_pieces) {
            for (int i = 0, q0 = 0; i < h->n_pieces && q0 < nb; ++i, q0 += h->piece_boards) {
                const int qn = std::min(h->piece_boards, nb - q0);
                CV_CUDA(cudaStreamWaitEvent(s, h->piece_ev[i], 0));
                rc = launch_frontend3(static_cast<const uint8_t*>(src) + (size_t)q0 * H * H * 3, qn, H, g, h->lut_host, h->fe3_wimg, bias_b00,
                                      front_out + (size_t)q0 * 64 * 256 * 16, h->num_sms, &done, s, nullptr, gate, permute ? qn : 0);
                if (rc) return rc;
                if (!done) break;                              // configuration not supported: nothing was launched
                ++h->launches;
            }
            if (!done)                                         // not supported after all: the first generation reads the whole chunk
                for (int i = 0; i < h->n_pieces; ++i) CV_CUDA(cudaStreamWaitEvent(s, h->piece_ev[i], 0));
        } else {
            rc = launch_frontend3(static_cast<const uint8_t*>(src), nb, H, g, h->lut_host, h->fe3_wimg, bias_b00, front_out, h->num_sms, &done, s,
                                  nullptr, gate, permute ? nb : 0);
            if (rc) return rc;
            h->launches += done;
        }
        *permuted = permute && done;
    }
    const int* v1_run_flag = nullptr;
    if (kind == CV_SRC_F32_NCHW && f2u != nullptr) {
        int took = 0;
        rc = launch_frontend3(f2u, nb, H, g, h->lut_host, h->fe3_wimg, bias_b00, front_out, h->num_sms, &took, s, f2u_flag, gate);
        if (rc) return rc;
        h->launches += took;
        if (took) v1_run_flag = f2u_flag;                      // otherwise (configuration not supported) the floats go the old way
    }
    if (!done) {
        rc = launch_frontend(src, kind, nb, H, g, h->lut, h->fe_wimg, h->blob + kLayers[0].b_offset, bias_b00, front_out, h->num_sms, s,
                             v1_run_flag, gate);
        if (rc) return rc;
        ++h->launches;
    }
    return CV_OK;
}

template <typename T>
int forward_impl(cv_square* h, const float* x_f32, const uint8_t* x_u8, int layout, int B, int H, int precision,
                 float* squares, float* turn, float* castling, float* features, void* workspace, size_t ws_bytes,
                 cudaStream_t s) {
    CropGeom g;
    int rc = cv_make_crop_geom(H, &g);
    if (rc) return rc;
    WavePlan p = make_plan(h, B, precision, x_f32 ? H : 0);
    if (ws_bytes < p.total) {
        cv_set_error("workspace too small: need %zu bytes, got %zu", p.total, ws_bytes);
        return CV_ERR_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    T* crops = reinterpret_cast<T*>(ws + p.off_crops);
    float* feat = reinterpret_cast<float*>(ws + p.off_feat);
    // 16-bit modes: crop gather + conv_stem + blocks.0.0 fused in one tensor-core kernel (activations stay in smem)
    const bool fused_front = sizeof(T) == 2 && (h->impl & CV_IMPL_FRONTEND);
    bf16* front_out = fused_front ? reinterpret_cast<bf16*>(ws + p.off_small[h->out_buf[1]]) : nullptr;
    // fp16 operands exist in the fused kernels only; weights that do not fit fp16 (checked at load time) make the mode run its bf16 kernels
    const int all_fused = CV_IMPL_FRONTEND | CV_IMPL_EARLY | CV_IMPL_MID | CV_IMPL_TAIL;
    const bool want_f16 = precision == CV_PRECISION_FP16;
    if (want_f16 && (h->impl & all_fused) != all_fused) {
        cv_set_error("CV_PRECISION_FP16 needs the fused kernels (CV_IMPL_FRONTEND | EARLY | MID | TAIL); the layer-granular kernels are bf16 / fp32");
        return CV_ERR_STATE;
    }
    const bool f16 = want_f16 && h->f16_ok;
    int* ovf = h->flags + 1;
    StageGate passes[2];
    int n_passes = 1;
    if (f16) {
        // both chains are enqueued: fp16 kernels run while *ovf == 0, the bf16 kernels after them only if it was raised
        CV_CUDA(cudaMemsetAsync(ovf, 0, sizeof(int), s));
        passes[0].f16 = true; passes[0].flag = ovf; passes[0].want = 0; passes[0].ovf = ovf;
        passes[1].f16 = false; passes[1].flag = ovf; passes[1].want = 1;
        n_passes = 2;
    }
    // host path: the boards of this call arrive in pieces (one event each, cv_square_predict_host_u8).  Only the third-generation front
    // end of a single-wave call follows them piece by piece; every other configuration waits for the whole chunk first.
    const bool by_pieces = h->n_pieces > 1 && fused_front && x_u8 && layout != CV_LAYOUT_CHW && (h->impl & CV_IMPL_FRONTEND3) && h->fe3_ok &&
                           B <= p.wave && B <= MAX_CHUNK;
    if (h->n_pieces > 0 && !by_pieces)
        for (int i = 0; i < h->n_pieces; ++i) CV_CUDA(cudaStreamWaitEvent(s, h->piece_ev[i], 0));
    // float source (the reference's own call, model(images)): recover the uint8 image the transform started from (one wave at a time,
    // into the workspace); needs 16-byte aligned rows for the float4 loads of the recovery kernel
    const bool recover = fused_front && x_f32 && (h->impl & CV_IMPL_FRONTEND3) && h->fe3_ok && p.f2u_bytes > 0 &&
                         (reinterpret_cast<uintptr_t>(x_f32) & 15) == 0 && (H * H) % 4 == 0;
    uint8_t* f2u = recover ? reinterpret_cast<uint8_t*>(ws + p.off_f2u) : nullptr;
    int* f2u_flag = h->flags + 2;
    for (int c0 = 0; c0 < B; c0 += MAX_CHUNK) {                    // chunk: one global-head launch
        const int cb = std::min(MAX_CHUNK, B - c0);
        for (int w0 = 0; w0 < cb; w0 += p.wave) {                  // wave: the stage hand-offs share the workspace
            const int b0 = c0 + w0;
            const int nb = std::min(p.wave, cb - w0);
            if (!fused_front) {
                rc = prof_mark(h, CV_PROF_CROP, s);
                if (rc) return rc;
                if (x_u8) rc = launch_crop_u8<T>(x_u8 + (size_t)b0 * H * H * 3, layout, nb, H, g, h->lut, crops, nullptr, s);
                else rc = launch_crop_f32<T>(x_f32 + (size_t)b0 * 3 * H * H, nb, H, g, crops, nullptr, s);
                if (rc) return rc;
                ++h->launches;
                rc = run_wave<T>(h, p, ws, nb, squares + (size_t)b0 * 832, feat + (size_t)w0 * 30720, feat, (int64_t)w0 * 64, b0 == 0, 0,
                                 StageGate(), false, s);
                if (rc) return rc;
                continue;
            }
            const void* src = x_u8 ? static_cast<const void*>(x_u8 + (size_t)b0 * H * H * 3)
                                   : static_cast<const void*>(x_f32 + (size_t)b0 * 3 * H * H);
            const int kind = x_u8 ? (layout == CV_LAYOUT_CHW ? CV_SRC_U8_CHW : CV_SRC_U8_HWC) : CV_SRC_F32_NCHW;
            if (recover) {
                rc = launch_f32_to_u8_boards(static_cast<const float*>(src), nb, H, h->lut_host, f2u, f2u_flag, s);
                if (rc) return rc;
                ++h->launches;
            }
            for (int pass = 0; pass < n_passes; ++pass) {
                rc = prof_mark(h, pass == 0 ? CV_PROF_FRONTEND : CV_PROF_FALLBACK, s);
                if (rc) return rc;
                h->prof_suspended = pass > 0;                  // the gated fall-back chain is one profiling slot
                // crop order: the front end gathers in the permuted order when the whole fused chain follows it and nobody taps a layer
                const bool permute = (h->impl & all_fused) == all_fused && !(h->tap_dst && h->tap_layer >= 0);
                bool permuted = false;
                rc = launch_front(h, src, kind, nb, H, g, front_out, by_pieces && pass == 0, f2u, f2u_flag, passes[pass], permute, &permuted, s);
                if (rc == CV_OK)
                    rc = run_wave<T>(h, p, ws, nb, squares + (size_t)b0 * 832, feat + (size_t)w0 * 30720, feat, (int64_t)w0 * 64, b0 == 0, 2,
                                     passes[pass], permuted, s);
                h->prof_suspended = false;
                if (rc) return rc;
            }
        }
        rc = prof_mark(h, CV_PROF_GLOBAL_HEAD, s);
        if (rc) return rc;
        if (sizeof(T) == 2 && (h->impl & CV_IMPL_TAIL)) {          // features are FT-tiled: tensor-core split-K GEMM + finishing kernel
            rc = launch_global_head_umma(feat, h->glob_wtile, reinterpret_cast<float*>(ws + p.off_partial), h->head_w + 4816,
                                         h->head_w + 4880, h->head_w + 5200, cb, h->num_sms, turn + c0, castling + (size_t)c0 * 4, s);
            if (rc) return rc;
            h->launches += 2;
            if (features) {
                rc = launch_untile_features(feat, features + (size_t)c0 * 30720, cb, s);
                if (rc) return rc;
            }
        } else {
            if (precision == CV_PRECISION_FP32) {
                rc = launch_global_head_f64_split(feat, h->glob_wt, h->head_w + 4816, h->head_w + 4880, h->head_w + 5200, cb,
                                                  reinterpret_cast<double*>(ws + p.off_partial), turn + c0, castling + (size_t)c0 * 4, s);
                ++h->launches;
            } else {
                rc = launch_global_head(feat, h->glob_wt, h->head_w + 4816, h->head_w + 4880, h->head_w + 5200, cb, turn + c0,
                                        castling + (size_t)c0 * 4, false, s);
            }
            if (rc) return rc;
            ++h->launches;
            if (features)
                CV_CUDA(cudaMemcpyAsync(features + (size_t)c0 * 30720, feat, (size_t)cb * 30720 * sizeof(float),
                                        cudaMemcpyDeviceToDevice, s));
        }
    }
    return prof_mark(h, -1, s);
}

// CV_PRECISION_FP32_SPLIT: crop gather in fp32 (bit-exact with the reference), then the 45 layers as layer-granular tensor-core kernels on
// split fp16 operands (kernels_exact.cu), pooling + heads, fp64-accumulating global head.  Activations are X2 tensors: 4 bytes per element,
// so the fp32 plan of the workspace is the right size.
int forward_split(cv_square* h, const float* x_f32, const uint8_t* x_u8, int layout, int B, int H, float* squares, float* turn, float* castling,
                  float* features, void* workspace, size_t ws_bytes, cudaStream_t s) {
    CropGeom g;
    int rc = cv_make_crop_geom(H, &g);
    if (rc) return rc;
    const WavePlan p = make_plan(h, B, CV_PRECISION_FP32_SPLIT, 0);
    if (ws_bytes < p.total) { cv_set_error("workspace too small: need %zu bytes, got %zu", p.total, ws_bytes); return CV_ERR_WORKSPACE; }
    const CropTaps taps = make_taps(g);
    char* ws = static_cast<char*>(workspace);
    float* crops = reinterpret_cast<float*>(ws + p.off_crops);
    float* feat = reinterpret_cast<float*>(ws + p.off_feat);
    int* ovf = h->flags + 1;
    CV_CUDA(cudaMemsetAsync(ovf, 0, sizeof(int), s));
    if (h->n_pieces > 0)
        for (int i = 0; i < h->n_pieces; ++i) CV_CUDA(cudaStreamWaitEvent(s, h->piece_ev[i], 0));
    auto buf_of = [&](int layer) -> uint16_t* {
        return reinterpret_cast<uint16_t*>(ws + (h->out_buf[layer] < 0 ? p.off_stem : p.off_small[h->out_buf[layer]]));
    };
    for (int c0 = 0; c0 < B; c0 += MAX_CHUNK) {
        const int cb = std::min(MAX_CHUNK, B - c0);
        for (int w0 = 0; w0 < cb; w0 += p.wave) {
            const int b0 = c0 + w0, nb = std::min(p.wave, cb - w0);
            const int64_t n = (int64_t)nb * 64;
            // uint8 HWC boards: the crop gather is fused into the stem's im2col gather (the fp32 crops never exist); other sources go
            // through the crop kernel first
            const bool fused_crop = x_u8 != nullptr && layout == CV_LAYOUT_HWC;
            const uint8_t* wave_boards = fused_crop ? x_u8 + (size_t)b0 * H * H * 3 : nullptr;
            if (!fused_crop) {
                rc = prof_mark(h, CV_PROF_CROP, s);
                if (rc) return rc;
                if (x_u8) rc = launch_crop_u8<float>(x_u8 + (size_t)b0 * H * H * 3, layout, nb, H, g, h->lut, crops, nullptr, s);
                else rc = launch_crop_f32<float>(x_f32 + (size_t)b0 * 3 * H * H, nb, H, g, crops, nullptr, s);
                if (rc) return rc;
                ++h->launches;
            }
            for (int i = 0; i < CV_NUM_LAYERS; ++i) {
                const cv_layer_info& L = kLayers[i];
                rc = prof_mark(h, CV_PROF_LAYER0 + i, s);
                if (rc) return rc;
                uint16_t* out = buf_of(i);
                const uint16_t* in = i > 0 ? buf_of(i - 1) : nullptr;
                const float* bias = h->blob + L.b_offset;
                if (L.kind == CV_KIND_DEPTHWISE) rc = launch_depthwise_x2(L, in, h->blob + L.w_offset, bias, out, n, ovf, s);
                else if (L.kind == CV_KIND_DENSE)
                    rc = launch_dense_x2(L, in, i == 0 && !fused_crop ? crops : nullptr, i == 0 ? wave_boards : nullptr, H, h->lut, &taps,
                                         h->x2_wimg + x2_weight_image_offset(i), bias, h->x2_unscale[i], out, n, h->num_sms, ovf, s);
                else
                    rc = launch_pointwise_x2(L, in, h->x2_wimg + x2_weight_image_offset(i), bias, h->x2_unscale[i], L.skip >= 0 ? buf_of(L.skip) : nullptr,
                                             out, n, h->num_sms, ovf, s);
                if (rc) return rc;
                ++h->launches;
                if (b0 == 0 && h->tap_layer == i && h->tap_dst) {
                    const size_t cnt = std::min(h->tap_n, (size_t)n * L.hout * L.hout * L.cout);
                    rc = launch_x2_to_f32(out, h->tap_dst, cnt, L.cout, s);
                    if (rc) return rc;
                }
            }
            rc = prof_mark(h, CV_PROF_POOL_HEADS, s);
            if (rc) return rc;
            rc = launch_pool_heads_x2(buf_of(CV_NUM_LAYERS - 1), h->head_w, h->head_w + 4800, n, feat + (size_t)w0 * 30720, squares + (size_t)b0 * 832, s);
            if (rc) return rc;
            ++h->launches;
        }
        rc = prof_mark(h, CV_PROF_GLOBAL_HEAD, s);
        if (rc) return rc;
        rc = launch_global_head_f64_split(feat, h->glob_wt, h->head_w + 4816, h->head_w + 4880, h->head_w + 5200, cb,
                                          reinterpret_cast<double*>(ws + p.off_partial), turn + c0, castling + (size_t)c0 * 4, s);
        if (rc) return rc;
        h->launches += 2;
        if (features) CV_CUDA(cudaMemcpyAsync(features + (size_t)c0 * 30720, feat, (size_t)cb * 30720 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return prof_mark(h, -1, s);
}

int check_forward_args(const cv_square* h, const void* x, int B, int H, int precision, const void* squares,
                       const void* turn, const void* castling, const void* ws) {
    if (!h) { cv_set_error("null handle"); return CV_ERR_ARG; }
    if (!h->loaded) { cv_set_error("weights not loaded: call cv_square_load_weights first"); return CV_ERR_STATE; }
    if (B < 0) { cv_set_error("negative batch"); return CV_ERR_ARG; }
    if (precision < CV_PRECISION_FP32 || precision > CV_PRECISION_FP32_SPLIT) { cv_set_error("bad precision %d", precision); return CV_ERR_ARG; }
    if (H < 32 || H % 32) { cv_set_error("board side H=%d must be a positive multiple of 32", H); return CV_ERR_ARG; }
    if (B > 0 && (!x || !squares || !turn || !castling || !ws)) { cv_set_error("null data pointer"); return CV_ERR_ARG; }
    return CV_OK;
}

void default_lut(float* lut) {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int c = 0; c < 3; ++c)
        for (int u = 0; u < 256; ++u) {
            volatile float t = (float)u / 255.0f;       // ToTensor: .div(255)
            volatile float d = t - mean[c];             // Normalize: .sub_(mean)
            lut[c * 256 + u] = d / stdv[c];             //            .div_(std)
        }
}

}  // namespace

extern "C" {

const char* cv_last_error(void) { return g_err; }
int cv_abi_version(void) { return CV_ABI_VERSION; }
int cv_num_layers(void) { return CV_NUM_LAYERS; }
size_t cv_weight_blob_floats(void) { return (size_t)CV_BLOB_FLOATS; }

int cv_layer_info_get(int index, cv_layer_info* out) {
    CV_ARG(out != nullptr, "null out");
    CV_ARG(index >= 0 && index < CV_NUM_LAYERS, "layer index out of range");
    *out = kLayers[index];
    return CV_OK;
}

int cv_square_create(int device, cv_square** out) {
    CV_ARG(out != nullptr, "null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cv_set_error("no CUDA device available (%s): chessvision_b200 has no CPU fallback", cudaGetErrorString(e));
        return CV_ERR_CUDA;
    }
    CV_ARG(device >= 0 && device < count, "device index out of range");
    CV_CUDA(cudaSetDevice(device));
    cv_square* h = new cv_square();
    h->device = device;
    plan_buffers(h);
    for (int i = 1; i < CV_NUM_LAYERS; ++i)
        if (h->out_buf[i] < 0 || (int64_t)kLayers[i].hout * kLayers[i].hout * kLayers[i].cout > EL_SMALL) {
            cv_set_error("internal: activation buffer plan failed at layer %d", i);
            delete h;
            return CV_ERR_STATE;
        }
    CV_CUDA(cudaMalloc(&h->blob, (size_t)CV_BLOB_FLOATS * sizeof(float)));
    CV_CUDA(cudaMalloc(&h->glob_wt, (size_t)30720 * 64 * sizeof(float)));
    CV_CUDA(cudaMalloc(&h->glob_wtile, (size_t)2 * 30720 * 64 * sizeof(float)));     // hi | lo planes
    CV_CUDA(cudaMalloc(&h->head_w, (4800 + 16 + 64 + 320 + 8) * sizeof(float)));
    CV_CUDA(cudaMalloc(&h->lut, 768 * sizeof(float)));
    CV_CUDA(cudaMalloc(&h->wimg, 2 * umma_weight_image_elems() * sizeof(bf16)));   // hi + lo images
    CV_CUDA(cudaMalloc(&h->fe_wimg, frontend_weight_image_elems() * sizeof(bf16)));
    CV_CUDA(cudaMalloc(&h->fe3_wimg, frontend3_weight_image_bytes() + 16));
    for (int f = 0; f < 2; ++f) {
        CV_CUDA(cudaMalloc(&h->sd_img[f], stageD_image_bytes() * CV_W_REPLICAS));
        CV_CUDA(cudaMalloc(&h->sc_img[f], stageC_image_bytes() * CV_W_REPLICAS));
        CV_CUDA(cudaMalloc(&h->sb_img[f], stageB_image_bytes()));
    }
    CV_CUDA(cudaMalloc(&h->x2_wimg, x2_weight_image_elems() * sizeof(uint16_t)));
    CV_CUDA(cudaMalloc(&h->flags, 4 * sizeof(int)));
    CV_CUDA(cudaMemset(h->flags, 0, 4 * sizeof(int)));
    CV_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    default_lut(h->lut_host);
    CV_CUDA(cudaMemcpy(h->lut, h->lut_host, sizeof(h->lut_host), cudaMemcpyHostToDevice));
    *out = h;
    return CV_OK;
}

int cv_square_destroy(cv_square* h) {
    if (!h) return CV_OK;
    cudaSetDevice(h->device);
    cudaFree(h->blob); cudaFree(h->glob_wt); cudaFree(h->glob_wtile); cudaFree(h->head_w); cudaFree(h->lut); cudaFree(h->wimg); cudaFree(h->fe_wimg); cudaFree(h->fe3_wimg); cudaFree(h->flags); cudaFree(h->x2_wimg);
    for (int f = 0; f < 2; ++f) { cudaFree(h->sd_img[f]); cudaFree(h->sc_img[f]); cudaFree(h->sb_img[f]); }
    for (int i = 0; i < cv_square::kStages; ++i) {
        if (h->stage[i]) cudaFree(h->stage[i]);
        if (h->stage_flip[i]) cudaFree(h->stage_flip[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        for (int j = 0; j < cv_square::kPieces; ++j)
            if (h->ev_piece[i][j]) cudaEventDestroy(h->ev_piece[i][j]);
    }
    for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
    if (h->own_ws) cudaFree(h->own_ws);
    if (h->dev_fen) cudaFree(h->dev_fen);
    if (h->dev_fen_len) cudaFree(h->dev_fen_len);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
    delete h;
    return CV_OK;
}

int cv_square_set_norm_lut(cv_square* h, const float* lut_host) {
    CV_ARG(h && lut_host, "null argument");
    CV_CUDA(cudaSetDevice(h->device));
    CV_CUDA(cudaMemcpy(h->lut, lut_host, 768 * sizeof(float), cudaMemcpyHostToDevice));
    memcpy(h->lut_host, lut_host, sizeof(h->lut_host));
    return CV_OK;
}

int cv_square_load_weights(cv_square* h, const float* blob, size_t n_floats, void* stream) {
    CV_ARG(h && blob, "null argument");
    CV_ARG(n_floats == (size_t)CV_BLOB_FLOATS, "weight blob has the wrong number of floats");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CV_CUDA(cudaSetDevice(h->device));
    CV_CUDA(cudaMemcpyAsync(h->blob, blob, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    int rc = launch_transpose_f32(h->blob + CV_OFF_GLOB_W, h->glob_wt, 64, 30720, s);
    if (rc) return rc;
    rc = launch_tile_glob_w(h->blob + CV_OFF_GLOB_W, h->glob_wtile, s);
    if (rc) return rc;
    rc = launch_umma_prep_weights(h->blob, h->wimg, s);
    if (rc) return rc;
    rc = launch_frontend_prep_weights(h->blob, h->fe_wimg, s);
    if (rc) return rc;
    int* fe3_flag = reinterpret_cast<int*>(h->fe3_wimg + frontend3_weight_image_bytes());
    rc = launch_frontend3_prep_weights(h->blob, h->fe3_wimg, fe3_flag, s);
    if (rc) return rc;
    rc = launch_x2_prep_weights(h->blob, h->x2_wimg, h->x2_unscale, s);
    if (rc) return rc;
    CV_CUDA(cudaMemsetAsync(h->flags, 0, 4 * sizeof(int), s));
    for (int f = 0; f < 2; ++f) {                 // bf16 images, then fp16 images (+ the fp16 range check of every GEMM weight)
        int* chk = f ? h->flags : nullptr;
        rc = build_stageD_image(h->blob, h->sd_img[f], h->sd_off[f], h->sd_bytes[f], chk, s);
        if (rc) return rc;
        rc = build_stageC_image(h->blob, h->sc_img[f], h->sc_off[f], h->sc_bytes[f], chk, s);
        if (rc) return rc;
        for (int r = 1; r < CV_W_REPLICAS; ++r) {
            CV_CUDA(cudaMemcpyAsync(h->sd_img[f] + r * stageD_image_bytes(), h->sd_img[f], stageD_image_bytes(), cudaMemcpyDeviceToDevice, s));
            CV_CUDA(cudaMemcpyAsync(h->sc_img[f] + r * stageC_image_bytes(), h->sc_img[f], stageC_image_bytes(), cudaMemcpyDeviceToDevice, s));
        }
        rc = build_stageB_image(h->blob, h->sb_img[f], chk, s);
        if (rc) return rc;
    }
    float* hw = h->head_w;
    CV_CUDA(cudaMemcpyAsync(hw, h->blob + CV_OFF_HEAD_W, 4800 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CV_CUDA(cudaMemcpyAsync(hw + 4800, h->blob + CV_OFF_HEAD_B, 10 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CV_CUDA(cudaMemcpyAsync(hw + 4816, h->blob + CV_OFF_GLOB_B, 64 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CV_CUDA(cudaMemcpyAsync(hw + 4880, h->blob + CV_OFF_TC_W, 320 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CV_CUDA(cudaMemcpyAsync(hw + 5200, h->blob + CV_OFF_TC_B, 5 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    int fe3_overflow = 1, w_overflow = 1;
    CV_CUDA(cudaMemcpyAsync(&fe3_overflow, fe3_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CV_CUDA(cudaMemcpyAsync(&w_overflow, h->flags, sizeof(int), cudaMemcpyDeviceToHost, s));
    CV_CUDA(cudaStreamSynchronize(s));
    h->fe3_ok = fe3_overflow == 0;
    h->f16_ok = w_overflow == 0 && h->fe3_ok;
    h->loaded = true;
    return CV_OK;
}

int cv_square_set_impl(cv_square* h, int mask) {
    CV_ARG(h != nullptr, "null handle");
    CV_ARG(mask >= 0 && mask <= CV_IMPL_ALL, "bad implementation mask");
    h->impl = mask;
    return CV_OK;
}

int cv_square_set_wave(cv_square* h, int boards) {
    CV_ARG(h != nullptr, "null handle");
    CV_ARG(boards >= 0 && boards <= 4096, "wave must be in [0,4096]");
    h->wave = boards;
    return CV_OK;
}

size_t cv_square_workspace_bytes(const cv_square* h, int max_boards, int H, int precision) {
    if (!h || max_boards < 0) return 0;
    return make_plan(h, std::max(max_boards, 1), precision, H).total;
}

int cv_square_forward_f32(cv_square* h, const float* x, int B, int H, int precision, float* squares, float* turn,
                          float* castling, float* features, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_forward_args(h, x, B, H, precision, squares, turn, castling, ws);
    if (rc) return rc;
    if (B == 0) return CV_OK;
    CV_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (precision == CV_PRECISION_FP32_SPLIT) return forward_split(h, x, nullptr, 0, B, H, squares, turn, castling, features, ws, ws_bytes, s);
    if (precision == CV_PRECISION_FP32)
        return forward_impl<float>(h, x, nullptr, 0, B, H, precision, squares, turn, castling, features, ws, ws_bytes, s);
    return forward_impl<bf16>(h, x, nullptr, 0, B, H, precision, squares, turn, castling, features, ws, ws_bytes, s);
}

int cv_square_forward_u8(cv_square* h, const uint8_t* boards, int layout, int B, int H, int precision, float* squares,
                         float* turn, float* castling, float* features, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_forward_args(h, boards, B, H, precision, squares, turn, castling, ws);
    if (rc) return rc;
    CV_ARG(layout == CV_LAYOUT_HWC || layout == CV_LAYOUT_CHW, "bad layout");
    if (B == 0) return CV_OK;
    CV_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (precision == CV_PRECISION_FP32_SPLIT) return forward_split(h, nullptr, boards, layout, B, H, squares, turn, castling, features, ws, ws_bytes, s);
    if (precision == CV_PRECISION_FP32)
        return forward_impl<float>(h, nullptr, boards, layout, B, H, precision, squares, turn, castling, features, ws, ws_bytes, s);
    return forward_impl<bf16>(h, nullptr, boards, layout, B, H, precision, squares, turn, castling, features, ws, ws_bytes, s);
}

int cv_square_fen(const float* squares, const float* turn, const float* castling, const uint8_t* flipped, int B,
                  char* fen, uint8_t* fen_len, void* stream) {
    CV_ARG(B >= 0, "negative batch");
    if (B == 0) return CV_OK;
    CV_ARG(squares && turn && castling && fen && fen_len, "null data pointer");
    return launch_fen(squares, turn, castling, flipped, B, fen, fen_len, static_cast<cudaStream_t>(stream));
}

int cv_square_predict_u8(cv_square* h, const uint8_t* boards, int layout, const uint8_t* flipped, int B, int H,
                         int precision, char* fen, uint8_t* fen_len, void* ws, size_t ws_bytes, void* stream) {
    CV_ARG(h != nullptr, "null handle");
    CV_ARG(B >= 0, "negative batch");
    if (B == 0) return CV_OK;
    CV_ARG(fen && fen_len && ws, "null data pointer");
    WavePlan p = make_plan(h, B, precision, 0);
    char* w = static_cast<char*>(ws);
    float* sq = reinterpret_cast<float*>(w + p.off_sq);
    float* tu = reinterpret_cast<float*>(w + p.off_turn);
    float* ca = reinterpret_cast<float*>(w + p.off_cast);
    int rc = cv_square_forward_u8(h, boards, layout, B, H, precision, sq, tu, ca, nullptr, ws, ws_bytes, stream);
    if (rc) return rc;
    rc = prof_mark(h, CV_PROF_FEN, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    rc = launch_fen(sq, tu, ca, flipped, B, fen, fen_len, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    ++h->launches;
    return prof_mark(h, -1, static_cast<cudaStream_t>(stream));
}

// Host-buffer end-to-end: chunks of `chunk` boards are copied H2D on a copy stream, piece by piece, into one of three staging
// buffers while the previous chunk computes; FEN records come back with one D2H at the end.
int cv_square_predict_host_u8(cv_square* h, const uint8_t* boards_host, int layout, const uint8_t* flipped_host, int B,
                              int H, int precision, char* fen_host, uint8_t* fen_len_host) {
    CV_ARG(h != nullptr, "null handle");
    CV_ARG(B >= 0, "negative batch");
    if (!h->loaded) { cv_set_error("weights not loaded"); return CV_ERR_STATE; }
    if (B == 0) return CV_OK;
    CV_ARG(boards_host && fen_host && fen_len_host, "null host pointer");
    CV_ARG(H >= 32 && H % 32 == 0, "H must be a positive multiple of 32");
    CV_CUDA(cudaSetDevice(h->device));
    if (!h->copy_stream) {
        CV_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CV_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
        for (int i = 0; i < cv_square::kStages; ++i) {
            CV_CUDA(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            CV_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            for (int j = 0; j < cv_square::kPieces; ++j) CV_CUDA(cudaEventCreateWithFlags(&h->ev_piece[i][j], cudaEventDisableTiming));
        }
    }
    const size_t per_board = (size_t)H * H * 3;
    int max_chunk = 512, piece = 128;
    bool ramp_head = false, shrink_tail = false, tail256 = false;      // measured (tools/gpu_host_trace.py): flat 512-board chunks are best
#ifdef CV_EXPERIMENTS                 // tuning knobs of the host pipeline: experiment builds only
    if (const char* e = getenv("CV_HOST_CHUNK")) max_chunk = std::min(4096, std::max(128, atoi(e) / 128 * 128));
    if (const char* e = getenv("CV_HOST_PIECE")) piece = std::max(32, atoi(e));
    if (const char* e = getenv("CV_HOST_SCHED")) { const int v = atoi(e); ramp_head = v & 1; shrink_tail = v & 2; tail256 = v & 4; }
#endif
    const int chunk = std::min(B, max_chunk);
    if (h->stage_bytes < chunk * per_board || h->stage_boards < chunk) {
        for (int i = 0; i < cv_square::kStages; ++i) {
            if (h->stage[i]) CV_CUDA(cudaFree(h->stage[i]));
            if (h->stage_flip[i]) CV_CUDA(cudaFree(h->stage_flip[i]));
            h->stage[i] = nullptr; h->stage_flip[i] = nullptr;
        }
        h->stage_bytes = 0; h->stage_boards = 0;
        const size_t bytes = std::max(h->stage_bytes, chunk * per_board);
        for (int i = 0; i < cv_square::kStages; ++i) {
            CV_CUDA(cudaMalloc(&h->stage[i], bytes));
            CV_CUDA(cudaMalloc(&h->stage_flip[i], (size_t)max_chunk));       // one flag per board of the largest chunk
        }
        h->stage_bytes = bytes;
        h->stage_boards = max_chunk;
    }
    size_t need = cv_square_workspace_bytes(h, chunk, H, precision);
    if (h->own_ws_bytes < need) {
        if (h->own_ws) CV_CUDA(cudaFree(h->own_ws));
        CV_CUDA(cudaMalloc(&h->own_ws, need));
        h->own_ws_bytes = need;
    }
    if (h->dev_fen_cap < B) {
        if (h->dev_fen) CV_CUDA(cudaFree(h->dev_fen));
        if (h->dev_fen_len) CV_CUDA(cudaFree(h->dev_fen_len));
        CV_CUDA(cudaMalloc(&h->dev_fen, (size_t)B * CV_FEN_STRIDE));
        CV_CUDA(cudaMalloc(&h->dev_fen_len, (size_t)B));
        h->dev_fen_cap = B;
    }
    // Chunks of 512 boards (one wave), each copied in four pieces with an event per piece: the front end of a chunk follows the copy
    // piece by piece, so what is exposed is the first piece's copy and, after the last piece has landed, one front-end piece plus
    // stages B-D of the last chunk (measured per 4096 boards: 16.0 ms; 16.6 ms with whole-chunk events and growing / shrinking
    // chunk sizes, 17.7 ms with equal chunks and whole-chunk events).  CV_HOST_SCHED: bit 0 growing head, bit 1 / 2 shrinking tails.
    // CV_HOST_TRACE=1: per-chunk timeline (copy end, compute end, ms since the first copy was enqueued) on stderr
#ifdef CV_EXPERIMENTS
    const bool trace = getenv("CV_HOST_TRACE") != nullptr;
#else
    const bool trace = false;
#endif
    std::vector<cudaEvent_t> tr_copy, tr_comp;
    std::vector<int> tr_nb;
    cudaEvent_t tr0 = nullptr;
    if (trace) { CV_CUDA(cudaEventCreate(&tr0)); CV_CUDA(cudaEventRecord(tr0, h->copy_stream)); }
    int slot = 0, it = 0;
    for (int b0 = 0, nb = 0; b0 < B; b0 += nb, slot = (slot + 1) % cv_square::kStages, ++it) {
        const int left = B - b0;
        nb = max_chunk;
        if (ramp_head) nb = it < 2 ? 128 : it == 2 ? 256 : max_chunk;
        if (shrink_tail) {
            if (left <= 256) nb = std::min(nb, 128);
            else if (left <= 512 + 256) nb = std::min(nb, 256);
        } else if (tail256 && left <= 512 && left > 256) nb = std::min(nb, 256);
        nb = std::min(std::min(nb, chunk), left);
        if (it >= cv_square::kStages) CV_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[slot], 0));   // staging slot free again
        if (flipped_host)
            CV_CUDA(cudaMemcpyAsync(h->stage_flip[slot], flipped_host + b0, nb, cudaMemcpyHostToDevice, h->copy_stream));
        // the chunk goes over in pieces of `piece` boards, one event each: the forward below waits for them one by one
        const int n_pieces = std::min((nb + piece - 1) / piece, (int)cv_square::kPieces);
        const int per_piece = n_pieces == cv_square::kPieces ? (nb + n_pieces - 1) / n_pieces : piece;
        for (int i = 0, q0 = 0; i < n_pieces; ++i, q0 += per_piece) {
            const int qn = std::min(per_piece, nb - q0);
            CV_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(h->stage[slot]) + (size_t)q0 * per_board, boards_host + (size_t)(b0 + q0) * per_board,
                                    qn * per_board, cudaMemcpyHostToDevice, h->copy_stream));
            CV_CUDA(cudaEventRecord(h->ev_piece[slot][i], h->copy_stream));
        }
        if (trace) { cudaEvent_t e; CV_CUDA(cudaEventCreate(&e)); CV_CUDA(cudaEventRecord(e, h->copy_stream)); tr_copy.push_back(e); tr_nb.push_back(nb); }
        h->piece_ev = h->ev_piece[slot];
        h->piece_boards = per_piece;
        h->n_pieces = n_pieces;
        int rc = cv_square_predict_u8(h, static_cast<const uint8_t*>(h->stage[slot]), layout,
                                      flipped_host ? h->stage_flip[slot] : nullptr, nb, H, precision,
                                      h->dev_fen + (size_t)b0 * CV_FEN_STRIDE, h->dev_fen_len + b0, h->own_ws,
                                      h->own_ws_bytes, h->compute_stream);
        h->n_pieces = 0;
        h->piece_ev = nullptr;
        if (rc) return rc;
        CV_CUDA(cudaEventRecord(h->ev_done[slot], h->compute_stream));
        if (trace) { cudaEvent_t e; CV_CUDA(cudaEventCreate(&e)); CV_CUDA(cudaEventRecord(e, h->compute_stream)); tr_comp.push_back(e); }
    }
    CV_CUDA(cudaMemcpyAsync(fen_host, h->dev_fen, (size_t)B * CV_FEN_STRIDE, cudaMemcpyDeviceToHost, h->compute_stream));
    CV_CUDA(cudaMemcpyAsync(fen_len_host, h->dev_fen_len, (size_t)B, cudaMemcpyDeviceToHost, h->compute_stream));
    CV_CUDA(cudaStreamSynchronize(h->compute_stream));
    if (trace) {
        for (size_t i = 0; i < tr_copy.size(); ++i) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tr0, tr_copy[i]);
            cudaEventElapsedTime(&b, tr0, tr_comp[i]);
            fprintf(stderr, "chunk %2zu (%4d boards): copy done %7.3f ms, compute done %7.3f ms\n", i, tr_nb[i], a, b);
            cudaEventDestroy(tr_copy[i]); cudaEventDestroy(tr_comp[i]);
        }
        cudaEventDestroy(tr0);
    }
    return CV_OK;
}

int cv_combine_type_color(const float* t, const float* c, int64_t n, float* joint, void* stream) {
    CV_ARG(n >= 0, "negative n");
    if (n == 0) return CV_OK;
    CV_ARG(t && c && joint, "null data pointer");
    return launch_combine(t, c, n, joint, static_cast<cudaStream_t>(stream));
}

int cv_crop_squares_f32(const float* x, int B, int H, float* crops_nchw, void* stream) {
    CV_ARG(B >= 0, "negative batch");
    if (B == 0) return CV_OK;
    CV_ARG(x && crops_nchw, "null data pointer");
    CropGeom g;
    int rc = cv_make_crop_geom(H, &g);
    if (rc) return rc;
    return launch_crop_f32<float>(x, B, H, g, nullptr, crops_nchw, static_cast<cudaStream_t>(stream));
}

int cv_crop_squares_u8(const uint8_t* boards, int layout, int B, int H, float* crops_nchw, void* stream) {
    CV_ARG(B >= 0, "negative batch");
    if (B == 0) return CV_OK;
    CV_ARG(boards && crops_nchw, "null data pointer");
    CV_ARG(layout == CV_LAYOUT_HWC || layout == CV_LAYOUT_CHW, "bad layout");
    CropGeom g;
    int rc = cv_make_crop_geom(H, &g);
    if (rc) return rc;
    float lut[768];
    default_lut(lut);
    float* dl = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CV_CUDA(cudaMallocAsync(&dl, sizeof(lut), s));
    CV_CUDA(cudaMemcpyAsync(dl, lut, sizeof(lut), cudaMemcpyHostToDevice, s));
    rc = launch_crop_u8<float>(boards, layout, B, H, g, dl, nullptr, crops_nchw, s);
    CV_CUDA(cudaStreamSynchronize(s));     // lut[] is a stack buffer
    CV_CUDA(cudaFreeAsync(dl, s));
    return rc;
}

// Host-side query of the crop source-index tables (pure integer/fp32 arithmetic, no GPU needed):
// y0,y1: (8,64) board rows/cols after replicate-pad clamping; lam: (64).  Written to HOST memory.
int cv_crop_index_table(int H, int32_t* y0, int32_t* y1, float* lam) {
    CV_ARG(y0 && y1 && lam, "null output pointer");
    CropGeom g;
    int rc = cv_make_crop_geom(H, &g);
    if (rc) return rc;
    for (int r = 0; r < 8; ++r)
        for (int d = 0; d < 64; ++d) {
            int a = r * g.sq + g.i0[d] - g.pad, b = r * g.sq + g.i1[d] - g.pad;
            y0[r * 64 + d] = std::min(std::max(a, 0), H - 1);
            y1[r * 64 + d] = std::min(std::max(b, 0), H - 1);
        }
    for (int d = 0; d < 64; ++d) lam[d] = g.lam[d];
    return CV_OK;
}

int cv_square_set_tap(cv_square* h, int layer, float* dst, size_t n_floats) {
    CV_ARG(h != nullptr, "null handle");
    CV_ARG(layer < CV_NUM_LAYERS, "layer index out of range");
    h->tap_layer = layer;
    h->tap_dst = layer < 0 ? nullptr : dst;
    h->tap_n = n_floats;
    return CV_OK;
}

int cv_synth_boards(uint8_t* boards, int layout, int64_t first_board, int B, int H, uint32_t seed, int dist,
                    uint8_t* flipped, void* stream) {
    CV_ARG(B >= 0, "negative batch");
    CV_ARG(H >= 32 && H % 32 == 0, "H must be a positive multiple of 32");
    CV_ARG(layout == CV_LAYOUT_HWC || layout == CV_LAYOUT_CHW, "bad layout");
    CV_ARG(dist == CV_DIST_UNIFORM || dist == CV_DIST_STRUCTURED, "bad distribution");
    return launch_synth(boards, layout, first_board, B, H, seed, dist, flipped, static_cast<cudaStream_t>(stream));
}

int cv_square_profile(cv_square* h, int enable) {
    CV_ARG(h != nullptr, "null handle");
    h->profiling = enable != 0;
    h->prof_slot.clear();
    return CV_OK;
}

int cv_square_profile_read(cv_square* h, double* ms, int64_t* counts) {
    CV_ARG(h && ms && counts, "null argument");
    CV_CUDA(cudaSetDevice(h->device));
    CV_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < CV_PROF_SLOTS; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (size_t i = 0; i + 1 < h->prof_slot.size(); ++i) {
        int slot = h->prof_slot[i];
        if (slot < 0) continue;
        float t = 0.f;
        CV_CUDA(cudaEventElapsedTime(&t, h->prof_pool[i], h->prof_pool[i + 1]));
        ms[slot] += t;
        ++counts[slot];
    }
    h->prof_slot.clear();
    return CV_OK;
}

int64_t cv_square_launch_count(const cv_square* h) { return h ? h->launches : 0; }

int cv_square_fp16_status(cv_square* h, int* weights_fit, int* overflowed) {
    CV_ARG(h != nullptr, "null handle");
    CV_CUDA(cudaSetDevice(h->device));
    int ovf = 0;
    CV_CUDA(cudaMemcpy(&ovf, h->flags + 1, sizeof(int), cudaMemcpyDeviceToHost));     // synchronises with the null stream's work
    if (weights_fit) *weights_fit = h->f16_ok ? 1 : 0;
    if (overflowed) *overflowed = ovf;
    return CV_OK;
}

}  // extern "C"
