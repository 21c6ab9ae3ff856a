// Warp-level FEN assembly (sm_100a): argmax over the 13 joint classes, run-length placement string,
// turn and castling fields -- predict.py:27-42 + dataset.py:52-70 (labels_to_fen), batched on the device.
// Also combine_type_color (models/common.py:10-24) as a stand-alone op.
//
// One warp per board; lane l owns squares 2l and 2l+1 (same rank).  The run-length encoding is done with
// warp ballots and bit arithmetic instead of a serial scan:
//   E  = 64-bit "square is empty" mask
//   a square emits a byte iff it holds a piece, or it is the LAST empty square of a run inside its rank;
//   its byte position = popcount(emit mask below it) + rank index (one '/' per completed rank).
#include "internal.h"

namespace {

__device__ __forceinline__ uint64_t spread_bits(uint32_t x) {      // bit i -> bit 2i
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

__device__ __forceinline__ int argmax13(const float* __restrict__ p) {
    float best = p[0];
    int bi = 0;
#pragma unroll
    for (int c = 1; c < 13; ++c) {
        float v = p[c];
        if (v > best) { best = v; bi = c; }       // strict '>' : first maximum wins, like torch.argmax
    }
    return bi;
}

// ---- bit-parallel run-length logic, shared by the kernel and the host test hook --------------------
__host__ __device__ __forceinline__ int popc64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
__host__ __device__ __forceinline__ int high_bit(uint32_t v) {     // index of the highest set bit, v != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz(v);
#endif
}
// squares that emit a byte: pieces, plus empties closing a run (file h, or next square holds a piece)
__host__ __device__ __forceinline__ uint64_t emit_mask(uint64_t E) {
    return ~E | (E & (0x8080808080808080ull | ~(E >> 1)));
}
// byte position of square s (also where its rank starts when s is on file a)
__host__ __device__ __forceinline__ int emit_pos(uint64_t emit, int s) {
    return popc64(emit & ((1ull << s) - 1ull)) + (s >> 3);
}
__host__ __device__ __forceinline__ char emit_char(uint64_t E, int s, int cls) {
    if (cls != 0) return ".PNBRQKpnbrqk"[cls];
    const int r = s >> 3, f = s & 7;
    const uint32_t rank_empty = (uint32_t)(E >> (8 * r)) & 0xFFu;
    const uint32_t pieces_below = ~rank_empty & ((1u << f) - 1u);
    return (char)('0' + (pieces_below ? f - high_bit(pieces_below) : f + 1));
}
// tail " w KQkq"; returns the total length
__host__ __device__ __forceinline__ int emit_tail(char* rec, int pos, float turn, const float* castling) {
    rec[pos++] = ' ';
    rec[pos++] = turn > 0.f ? 'b' : 'w';                      // predict.py:32
    rec[pos++] = ' ';
    const char kq[4] = {'K', 'Q', 'k', 'q'};
    int n = 0;
    for (int i = 0; i < 4; ++i)
        if (castling[i] > 0.f) { rec[pos++] = kq[i]; ++n; }   // predict.py:35-39
    if (n == 0) rec[pos++] = '-';                             // predict.py:40
    return pos;
}

constexpr int WARPS = 8;

__global__ void __launch_bounds__(WARPS * 32)
fen_kernel(const float* __restrict__ squares, const float* __restrict__ turn, const float* __restrict__ castling,
           const uint8_t* __restrict__ flipped, int B, char* __restrict__ fen, uint8_t* __restrict__ fen_len) {
    __shared__ __align__(16) char rec[WARPS][CV_FEN_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * WARPS + warp;
    if (b >= B) return;
    const bool flip = flipped != nullptr && flipped[b] != 0;
    const float* sq = squares + (int64_t)b * 832;
    const int s0 = 2 * lane, s1 = s0 + 1;
    const int c0 = argmax13(sq + (flip ? 63 - s0 : s0) * 13);
    const int c1 = argmax13(sq + (flip ? 63 - s1 : s1) * 13);
    const uint64_t E = spread_bits(__ballot_sync(0xffffffffu, c0 == 0)) |
                       (spread_bits(__ballot_sync(0xffffffffu, c1 == 0)) << 1);
    const uint64_t emit = emit_mask(E);
    for (int i = lane; i < CV_FEN_STRIDE / 4; i += 32) reinterpret_cast<uint32_t*>(rec[warp])[i] = 0u;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int s = k ? s1 : s0, c = k ? c1 : c0;
        const int pos = emit_pos(emit, s);
        if ((s & 7) == 0 && s > 0) rec[warp][pos - 1] = '/';      // rank separator precedes file a
        if ((emit >> s) & 1ull) rec[warp][pos] = emit_char(E, s, c);
    }
    if (lane == 0)
        fen_len[b] = (uint8_t)emit_tail(rec[warp], popc64(emit) + 7, turn[b], castling + (int64_t)b * 4);
    __syncwarp();
    uint32_t* dst = reinterpret_cast<uint32_t*>(fen + (int64_t)b * CV_FEN_STRIDE);
    for (int i = lane; i < CV_FEN_STRIDE / 4; i += 32) dst[i] = reinterpret_cast<const uint32_t*>(rec[warp])[i];
}

__global__ void combine_kernel(const float* __restrict__ t, const float* __restrict__ c, int64_t n,
                               float* __restrict__ joint) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * 13) return;
    int k = (int)(i % 13);
    int64_t r = i / 13;
    int ti = k == 0 ? 0 : 1 + (k - 1) % 6;            // CLASS_TO_TYPE  (dataset.py:31)
    int ci = k == 0 ? 0 : (k <= 6 ? 1 : 2);           // CLASS_TO_COLOR (dataset.py:32)
    joint[i] = t[r * 7 + ti] + c[r * 3 + ci];
}

}  // namespace

int launch_fen(const float* squares, const float* turn, const float* castling, const uint8_t* flipped, int B, char* fen,
               uint8_t* fen_len, cudaStream_t s) {
    if (B == 0) return CV_OK;
    fen_kernel<<<(B + WARPS - 1) / WARPS, WARPS * 32, 0, s>>>(squares, turn, castling, flipped, B, fen, fen_len);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

// Host test hook: the same bit-parallel encoder driven square by square (no GPU needed).
extern "C" int cv_fen_from_classes_host(const int8_t* classes, float turn, const float* castling, char* rec80) {
    if (!classes || !castling || !rec80) { cv_set_error("cv_fen_from_classes_host: null argument"); return CV_ERR_ARG; }
    uint64_t E = 0;
    for (int s = 0; s < 64; ++s) {
        if (classes[s] < 0 || classes[s] > 12) { cv_set_error("class index out of range"); return CV_ERR_ARG; }
        if (classes[s] == 0) E |= 1ull << s;
    }
    const uint64_t emit = emit_mask(E);
    for (int i = 0; i < CV_FEN_STRIDE; ++i) rec80[i] = 0;
    for (int s = 0; s < 64; ++s) {
        const int pos = emit_pos(emit, s);
        if ((s & 7) == 0 && s > 0) rec80[pos - 1] = '/';
        if ((emit >> s) & 1ull) rec80[pos] = emit_char(E, s, classes[s]);
    }
    return emit_tail(rec80, popc64(emit) + 7, turn, castling);
}

int launch_combine(const float* t, const float* c, int64_t n, float* joint, cudaStream_t s) {
    if (n == 0) return CV_OK;
    int64_t total = n * 13;
    combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(t, c, n, joint);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
