// On-device evaluation bookkeeping: the step right after the hot path.  Replaces the reference's
// per-batch host loop (/root/reference/evaluate.py:74-155: ~20 device->host syncs and O(B*64) Python iterations per batch) by
// one kernel that turns the logits and labels of a batch into exact integer counters, per-sample flags and per-board loss sums.
//   warp per board, lane = 2 squares: argmax over 13 classes (first maximum wins, like torch.argmax), log-sum-exp for the
//   cross-entropy, ballots for the per-board results; 13x13 confusion + per-piece counts go through a shared-memory histogram,
//   one 64-bit global atomic per non-zero bin and block.
#include "internal.h"

namespace {

constexpr int N_COUNTERS = CV_EVAL_COUNTERS;

__global__ void __launch_bounds__(256) eval_kernel(const float* __restrict__ squares, const float* __restrict__ turn, const float* __restrict__ castling,
                                                   const uint8_t* __restrict__ sq_labels, const uint8_t* __restrict__ turn_labels,
                                                   const uint8_t* __restrict__ castling_labels, const uint8_t* __restrict__ legal, int B,
                                                   unsigned long long* __restrict__ counters, uint8_t* __restrict__ per_sample,
                                                   float* __restrict__ board_loss) {
    __shared__ unsigned int hist[N_COUNTERS];
    for (int i = threadIdx.x; i < N_COUNTERS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b < B) {
        int wrong = 0;
        float loss = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int sq = 2 * lane + k;
            const float* x = squares + ((size_t)b * 64 + sq) * 13;
            float best = x[0];
            int arg = 0;
#pragma unroll
            for (int c = 1; c < 13; ++c) {
                const float v = x[c];
                if (v > best) { best = v; arg = c; }                       // strict: the first maximum wins
            }
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < 13; ++c) se += expf(x[c] - best);
            const int label = min((int)sq_labels[(size_t)b * 64 + sq], 12);      // labels are 0..12 (dataset.py:14-19)
            loss += (logf(se) + best) - x[label];
            wrong += arg != label;
            atomicAdd(&hist[CV_EVAL_CONFUSION + label * 13 + arg], 1u);
            atomicAdd(&hist[CV_EVAL_PIECE_TOTAL + label], 1u);
            if (arg == label) atomicAdd(&hist[CV_EVAL_PIECE_CORRECT + label], 1u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                                 // fixed-order tree: deterministic per-board sums
            wrong += __shfl_xor_sync(0xffffffffu, wrong, o);
            loss += __shfl_xor_sync(0xffffffffu, loss, o);
        }
        if (lane == 0) {
            const bool board_ok = wrong == 0;
            const bool is_legal = legal[b] != 0;
            const bool turn_pred = turn[b] > 0.f, turn_true = turn_labels[b] != 0;
            const bool turn_ok = turn_pred == turn_true;
            bool cast_all = true;
            atomicAdd(&hist[CV_EVAL_TOTAL_BOARDS], 1u);
            atomicAdd(&hist[CV_EVAL_TOTAL_SQUARES], 64u);
            atomicAdd(&hist[CV_EVAL_CORRECT_SQUARES], (unsigned)(64 - wrong));
            if (board_ok) atomicAdd(&hist[CV_EVAL_CORRECT_BOARDS], 1u);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool ok = (castling[b * 4 + r] > 0.f) == (castling_labels[b * 4 + r] != 0);
                cast_all = cast_all && ok;
                if (ok && is_legal) atomicAdd(&hist[CV_EVAL_CORRECT_CASTLING_RIGHT + r], 1u);
            }
            if (is_legal) {
                atomicAdd(&hist[CV_EVAL_TOTAL_LEGAL], 1u);
                if (turn_ok) atomicAdd(&hist[CV_EVAL_CORRECT_TURN], 1u);
                atomicAdd(&hist[CV_EVAL_TURN_CONFUSION + 2 * (int)turn_true + (int)turn_pred], 1u);
                if (cast_all) atomicAdd(&hist[CV_EVAL_CORRECT_CASTLING_ALL], 1u);
                if (board_ok && turn_ok && cast_all) atomicAdd(&hist[CV_EVAL_CORRECT_FULL_FEN], 1u);
            }
            per_sample[b * 4 + 0] = (uint8_t)wrong;
            per_sample[b * 4 + 1] = board_ok;
            per_sample[b * 4 + 2] = is_legal ? (uint8_t)turn_ok : 255;
            per_sample[b * 4 + 3] = is_legal ? (uint8_t)cast_all : 255;
            board_loss[b] = loss;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N_COUNTERS; i += blockDim.x)
        if (hist[i]) atomicAdd(&counters[i], (unsigned long long)hist[i]);
}

}  // namespace

extern "C" int cv_eval_accumulate(const float* squares, const float* turn, const float* castling, const uint8_t* sq_labels,
                                  const uint8_t* turn_labels, const uint8_t* castling_labels, const uint8_t* legal, int B,
                                  int64_t* counters, uint8_t* per_sample, float* board_loss, void* stream) {
    CV_ARG(B >= 0, "negative batch");
    if (B == 0) return CV_OK;
    CV_ARG(squares && turn && castling && sq_labels && turn_labels && castling_labels && legal && counters && per_sample && board_loss,
           "null data pointer");
    eval_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(squares, turn, castling, sq_labels, turn_labels, castling_labels, legal, B,
                                                                           reinterpret_cast<unsigned long long*>(counters), per_sample, board_loss);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
