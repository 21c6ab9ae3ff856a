// Counter-based synthetic boards: bit-identical to chess_vision_b200/synthetic.py.
// Every byte is a pure function of (seed, global board index, y, x, channel), so any sharding of a board
// stream over ranks regenerates the same boards.  Also the host mirror used by the no-GPU tests.
#include "internal.h"

namespace {

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {   // lowbias32
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t board_key(uint32_t seed, int64_t board) {
    return hash32(seed ^ hash32((uint32_t)((uint64_t)board & 0xFFFFFFFFull) + 0x9E3779B9u));
}
__host__ __device__ __forceinline__ uint8_t board_byte(uint32_t key, int dist, int H, int y, int x, int c) {
    if (dist == CV_DIST_UNIFORM) return (uint8_t)(hash32(key + 5000u + (uint32_t)((y * H + x) * 3 + c)) & 0xFFu);
    const int sq = H / 8, cell = H / 32;
    const uint32_t base = hash32(key + 1u + (uint32_t)(((y / sq) * 8 + (x / sq)) * 3 + c)) & 0xFFu;
    const uint32_t pat = (hash32(key + 1000u + (uint32_t)(((y / cell) * 32 + (x / cell)) * 3 + c)) >> 8) & 0xFFu;
    return (uint8_t)((base + pat + 1u) >> 1);
}

// One thread per 4 consecutive output bytes (one 32-bit store).
__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* __restrict__ out, int layout, int64_t first_board, int64_t total_words, int H, uint32_t seed,
             int dist) {
    int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (wi >= total_words) return;
    const int64_t per_board = (int64_t)H * H * 3;
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int64_t i = wi * 4 + k;
        int64_t b = i / per_board;
        int r = (int)(i - b * per_board);
        int y, x, c;
        if (layout == CV_LAYOUT_CHW) { c = r / (H * H); r -= c * H * H; y = r / H; x = r - y * H; }
        else { c = r % 3; r /= 3; y = r / H; x = r - y * H; }
        word |= (uint32_t)board_byte(board_key(seed, first_board + b), dist, H, y, x, c) << (8 * k);
    }
    reinterpret_cast<uint32_t*>(out)[wi] = word;
}

__global__ void flipped_kernel(uint8_t* __restrict__ flipped, int64_t first_board, int B, uint32_t seed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) flipped[i] = (uint8_t)(hash32(board_key(seed, first_board + i) + 7777u) & 1u);
}

}  // namespace

int launch_synth(uint8_t* boards, int layout, int64_t first_board, int B, int H, uint32_t seed, int dist,
                 uint8_t* flipped, cudaStream_t s) {
    if (B == 0) return CV_OK;
    if (boards) {
        int64_t words = (int64_t)B * H * H * 3 / 4;       // H % 32 == 0 -> divisible
        synth_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(boards, layout, first_board, words, H, seed, dist);
        CV_CHECK_LAUNCH();
    }
    if (flipped) {
        flipped_kernel<<<(B + 255) / 256, 256, 0, s>>>(flipped, first_board, B, seed);
        CV_CHECK_LAUNCH();
    }
    return CV_OK;
}

// Host mirror (no GPU): fills HOST memory; used by bench.py to build pinned host inputs quickly and by the
// no-GPU tests to pin the device hash against chess_vision_b200/synthetic.py.
extern "C" int cv_synth_boards_host(uint8_t* boards_host, int layout, int64_t first_board, int B, int H, uint32_t seed,
                                    int dist, uint8_t* flipped_host) {
    if (H < 32 || H % 32 != 0) { cv_set_error("cv_synth_boards_host: H must be a multiple of 32"); return CV_ERR_ARG; }
    if (B < 0) { cv_set_error("cv_synth_boards_host: negative B"); return CV_ERR_ARG; }
    for (int b = 0; b < B; ++b) {
        const uint32_t key = board_key(seed, first_board + b);
        if (flipped_host) flipped_host[b] = (uint8_t)(hash32(key + 7777u) & 1u);
        if (!boards_host) continue;
        uint8_t* o = boards_host + (int64_t)b * H * H * 3;
        if (layout == CV_LAYOUT_CHW) {
            for (int c = 0; c < 3; ++c)
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < H; ++x) *o++ = board_byte(key, dist, H, y, x, c);
        } else {
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < H; ++x)
                    for (int c = 0; c < 3; ++c) *o++ = board_byte(key, dist, H, y, x, c);
        }
    }
    return CV_OK;
}
