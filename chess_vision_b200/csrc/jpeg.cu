// Baseline JPEG decode for the step BEFORE the hot path: replaces `Image.open(path).convert("RGB")` (predict.py:19,
// dataset.py ChessDataset) for the files the reference's datagen writes (datagen/generate.js:26-27: JPEG, quality 90).
//
// The arithmetic of that call lives in Pillow on libjpeg-turbo with libjpeg's decompression defaults; it is integer work end to end, so
// the bar is BIT-EXACT (oracle/jpeg_oracle.py pins the restatement to Pillow's own decodes):
//   entropy decode   baseline sequential Huffman, restart intervals (jdhuff.c)            -> huff_decode_interval  (host AND device)
//   dequant + IDCT   JDCT_ISLOW, "accurate integer" (jidctint.c jpeg_idct_islow)          -> jpeg_idct_kernel
//   upsampling       do_fancy_upsampling triangle filters (jdsample.c), replicated edges  -> jpeg_color_kernel (on the fly)
//   colour           ycc_rgb_convert 16-bit fixed-point tables (jdcolor.c)                -> jpeg_color_kernel
//
// Data flow per batch: the COMPRESSED files go over PCIe (about a tenth of the decoded pixels), one device thread per restart interval
// (or per file) walks its bit stream and writes quantised coefficients [block][64] int16; the IDCT kernel turns 32 blocks per CTA into
// samples through shared memory (8 threads per block: column pass, row pass); the colour kernel reads the three planes, upsamples
// chroma on the fly and writes RGB bytes (B, H, W, 3) -- the layout cv_resize_bilinear_u8 / cv_square_predict_u8 take.
// The same Huffman routine compiled for the host backs cv_jpeg_decode_coefficients_host (no-GPU tests against the oracle).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"

namespace {

// ---- per-file description (built on the host by parse_jpeg) ------------------------------------------------------------------------
struct HuffTab {
    uint8_t look_nbits[256];      // 8-bit look-ahead: code length (0 = longer than 8 bits)
    uint8_t look_sym[256];
    int32_t maxcode[18];          // largest code of length l (-1 if none), maxcode[17] = sentinel
    int32_t valoffset[17];        // vals index of the first code of length l minus that code
    uint8_t vals[256];
};

struct JpegDesc {
    int32_t width, height, ncomp;
    int32_t hs, vs;               // luma sampling factors (chroma is 1x1)
    int32_t mcux, mcuy;           // MCU grid
    int32_t restart;              // MCUs per restart interval (0 = none)
    int32_t dc_sel[3], ac_sel[3], q_sel[3];
    int32_t tset;                 // index of this file's Huffman table set in the batch (files of one encoder share one set)
    uint16_t qt[4][64];           // quantisation tables, natural order
    int64_t data_off, data_len;   // entropy-coded segment inside the staged file bytes (offset from the batch buffer start)
    int64_t coef_off[3];          // int16 offset of each component's first block inside the batch coefficient buffer
    int32_t bw[3], bh[3];         // block grid of each component (whole MCUs)
    int64_t plane_off[3];         // byte offset of each component's sample plane inside the batch plane buffer
};

struct TableSet { HuffTab dc[4], ac[4]; };      // 7.3 KB: the DHT segments of one file

struct Interval {                 // one independently decodable run of MCUs
    int32_t image;
    int32_t mcu0, n_mcu;
    int64_t byte0, byte1;         // within the batch buffer
};

inline const uint8_t* zigzag_tab() {
    static const uint8_t z[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return z;
}
__constant__ uint8_t kZigzagDev[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                       35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ---- bit reader + Huffman decode (jdhuff.c semantics: 0xFF00 -> 0xFF, zeros after a marker), shared by host and device --------------
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc;
    int n;
    int fake;                     // of the n unread bits, how many are zeros fed after a marker / the end of the range (the youngest ones)
    bool hit_marker;
#ifdef __CUDA_ARCH__
    // device: the stream is read in aligned 8-byte words (one global load per 8 bytes instead of one dependent load per byte; the
    // staged buffer is padded, so a word may reach past `end` -- those bytes are never consumed)
    uint64_t word;
    int word_left;
    __device__ inline void init() {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const int skip = (int)(a & 7);
        word = *reinterpret_cast<const uint64_t*>(a - skip) >> (8 * skip);
        word_left = 8 - skip;
    }
    __device__ inline uint32_t next_raw() {                  // byte at p, p advances
        if (word_left == 0) { word = *reinterpret_cast<const uint64_t*>(p); word_left = 8; }
        const uint32_t b = (uint32_t)word & 0xffu;
        word >>= 8; --word_left; ++p;
        return b;
    }
    __device__ inline uint32_t peek_raw() {                  // byte at p without advancing
        if (word_left == 0) { word = *reinterpret_cast<const uint64_t*>(p); word_left = 8; }
        return (uint32_t)word & 0xffu;
    }
#else
    inline void init() {}
    inline uint32_t next_raw() { return *p++; }
    inline uint32_t peek_raw() { return *p; }
#endif
    __host__ __device__ inline void step_byte() {            // one stream byte into the accumulator, the careful way (stuffing, markers, the end)
        uint32_t b = 0;
        if (!hit_marker && p < end) {
            b = next_raw();
            if (b == 0xFF) {
                if (p < end && peek_raw() == 0) next_raw();              // stuffed zero
                else { hit_marker = true; b = 0; fake += 8; }             // a marker (or the end): zeros from here on, like jdhuff.c
            }
        } else {
            fake += 8;
        }
        acc = (acc << 8) | b;
        n += 8;
    }
#ifdef __CUDA_ARCH__
    // device: until more than 32 bits are unread (a symbol takes at most 16 + 16; callers refill below 32 / 16).  Four stream bytes at
    // once whenever none of them is 0xFF (no stuffing, no marker); single careful bytes otherwise.  (The byte-at-a-time loop of the
    // host version was 55 % of the decoder's instructions at 3.4 active lanes: ncu, profiles/r02_jpeg_*.)
    __device__ inline void fill() {
        while (n <= 32) {
            if (!hit_marker && p + 4 <= end) {
                if (word_left == 0) { word = *reinterpret_cast<const uint64_t*>(p); word_left = 8; }
                const uint32_t x = (uint32_t)word, y = ~x;
                if (word_left >= 4 && ((y - 0x01010101u) & x & 0x80808080u) == 0) {      // no byte of y is zero <=> no byte of x is 0xFF
                    acc = (acc << 32) | __byte_perm(x, 0, 0x0123);
                    n += 32;
                    word >>= 32; word_left -= 4; p += 4;
                    return;
                }
            }
            step_byte();
        }
    }
#else
    inline void fill() {
        while (n <= 56) step_byte();
    }
#endif
    __host__ __device__ inline uint32_t peek(int k) { return (uint32_t)(acc >> (n - k)) & ((1u << k) - 1u); }
    __host__ __device__ inline void skip(int k) { n -= k; }
};

__host__ __device__ inline int huff_symbol(BitReader& br, const HuffTab& t) {
    if (br.n < 32) br.fill();                               // the slow path below peeks up to 17 bits
    const uint32_t look = br.peek(8);
    int nb = t.look_nbits[look];
    if (nb) { br.skip(nb); return t.look_sym[look]; }
    int32_t code = (int32_t)look;
    nb = 8;
    do {
        ++nb;
        code = (int32_t)br.peek(nb);
    } while (nb < 17 && code > t.maxcode[nb]);
    if (nb > 16) { br.skip(16); return 0; }                 // corrupt code: libjpeg warns and returns 0
    br.skip(nb);
    return t.vals[(code + t.valoffset[nb]) & 255];
}

__host__ __device__ inline int huff_extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// Decodes the MCUs [mcu0, mcu0 + n_mcu) of one file from its own byte range.  Every coefficient block is zeroed with eight 16-byte
// stores and the non-zero coefficients are stored as they are decoded (stores do not stall the thread; no local array).
__host__ __device__ inline void huff_decode_interval(const JpegDesc& d, const TableSet& ts, const uint8_t* bytes, int64_t byte0, int64_t byte1, int mcu0,
                                                    int n_mcu, int16_t* coef, const uint8_t* zz, int16_t* dcbuf = nullptr /* DC of block k also to dcbuf[k] */) {
    BitReader br;
    br.p = bytes + byte0; br.end = bytes + byte1; br.acc = 0; br.n = 0; br.fake = 0; br.hit_marker = false;
    br.init();
    int pred[3] = {0, 0, 0};
    // the fields of the descriptor the loop needs, read once (the coefficient stores below could alias `d` as far as the compiler knows)
    const int mcux = d.mcux, ncomp = d.ncomp, hs = d.hs, vs = d.vs;
    const HuffTab* dctab[3] = {&ts.dc[d.dc_sel[0]], &ts.dc[d.dc_sel[1]], &ts.dc[d.dc_sel[2]]};
    const HuffTab* actab[3] = {&ts.ac[d.ac_sel[0]], &ts.ac[d.ac_sel[1]], &ts.ac[d.ac_sel[2]]};
    int16_t* cbase[3] = {coef + d.coef_off[0], coef + d.coef_off[1], coef + d.coef_off[2]};
    const int bws[3] = {d.bw[0], d.bw[1], d.bw[2]};
    int my = mcu0 / mcux, mx = mcu0 - my * mcux;
    for (int m = 0; m < n_mcu; ++m, ++mx) {
        if (mx == mcux) { mx = 0; ++my; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c >= ncomp) break;
            const int ch = c == 0 ? hs : 1, cv = c == 0 ? vs : 1;
            const HuffTab& dct = *dctab[c];
            const HuffTab& act = *actab[c];
            for (int by = 0; by < cv; ++by)
                for (int bx = 0; bx < ch; ++bx) {
                    int16_t* blk = cbase[c] + ((int64_t)(my * cv + by) * bws[c] + (mx * ch + bx)) * 64;
                    uint4* blk4 = reinterpret_cast<uint4*>(blk);
#pragma unroll
                    for (int i = 0; i < 8; ++i) blk4[i] = make_uint4(0, 0, 0, 0);
                    int s = huff_symbol(br, dct);
                    if (s) {
                        if (br.n < 16) br.fill();
                        const int v = (int)br.peek(s);
                        br.skip(s);
                        pred[c] += huff_extend(v, s);
                    }
                    blk[0] = (int16_t)pred[c];
                    if (dcbuf != nullptr) dcbuf[(blk - coef) >> 6] = (int16_t)pred[c];
                    for (int k = 1; k < 64;) {
                        const int rs = huff_symbol(br, act);
                        const int r = rs >> 4;
                        s = rs & 15;
                        if (s == 0) {
                            if (r != 15) break;
                            k += 16;
                            continue;
                        }
                        k += r;
                        if (k > 63) break;
                        if (br.n < 16) br.fill();
                        const int v = (int)br.peek(s);
                        br.skip(s);
                        blk[zz[k]] = (int16_t)huff_extend(v, s);
                        ++k;
                    }
                }
        }
    }
}

// ---- intra-file parallelism: speculative decoding of fixed-size chunks of the entropy-coded bytes ------------------------------------------
// A baseline Huffman stream has no entry points (the reference's files carry no restart markers), so one thread per file leaves the
// device nearly empty (4096 files = 4096 threads, ~10 ms whatever the batch).  Chunk i of an interval = the symbols (a Huffman code
// with its extra bits) that START in raw bits [i*S*8, (i+1)*S*8).  The decoder state at a symbol boundary is (bit position, block index
// inside the MCU, next zig-zag index).  Round 0 decodes every chunk from a GUESS (chunk start, state 0); Huffman streams resynchronise
// after a few symbols (and the block / component phase after a few MCUs), so most exit states are already the true ones.  Round r
// re-decodes chunk i from the exit state of chunk i-1 wherever that changed.  After a fixed number of rounds a chain
// "entry(i) == exit(i-1) for all i, entry(0) = the true start" PROVES every chunk was decoded from its true state (induction), a prefix
// sum of the per-chunk block counts places each chunk in the coefficient buffer, and a final pass decodes again and writes.  Intervals
// whose chain is still broken fall back to the one-thread kernel.  DC values are written as differences and prefix-summed per
// component afterwards (the predictor is the only state that runs across chunks).  Same routine on host and device.
struct IvChunks { int32_t chunk0, n_chunks, chunk_bytes; };     // chunks of one interval inside the batch's chunk arrays; their size

__host__ __device__ inline uint64_t pack_state(int64_t bit, int c, int z) { return ((uint64_t)bit << 16) | ((uint64_t)c << 8) | (uint64_t)z; }

// raw bit position (relative to `base`, stuffed bytes counted) of the next unread bit
__host__ __device__ inline int64_t raw_bitpos(const BitReader& br, const uint8_t* base) {
    const int real = br.n - br.fake;
    if (real <= 0) return (int64_t)(br.p - base) * 8;       // everything read (or the range ran out): where the reader stands
    const uint8_t* q = br.p - (br.hit_marker && br.p[-1] == 0xFF ? 1 : 0);   // the 0xFF of the marker that ended the data was consumed, it is not data
    for (int j = (real + 7) >> 3; j > 0; --j) {              // walk back over the unread DATA bytes; a 0x00 behind a 0xFF is stuffing
        --q;
        if (*q == 0 && q > base && q[-1] == 0xFF) --q;
    }
    return (int64_t)(q - base) * 8 + ((8 - (real & 7)) & 7);
}

// Decodes chunk symbols from `entry` until a symbol boundary at or beyond end_bit (or, WRITE, until block blk_end of the interval).
// WRITE: coefficients go to their blocks (the buffer was zeroed), DC as the DIFFERENCE; blk0 = blocks of the interval completed before
// `entry`.  Returns the exit state; *n_done = blocks completed.
template <bool WRITE>
__host__ __device__ inline uint64_t huff_decode_chunk(const JpegDesc& d, const TableSet& ts, const uint8_t* bytes, int64_t byte0, int64_t byte1,
                                                     uint64_t entry, int64_t end_bit, int mcu0, int64_t blk0, int64_t blk_end, int16_t* coef,
                                                     const uint8_t* zz, int* n_done, int16_t* dcbuf = nullptr /* WRITE: DC differences go to dcbuf[block] */) {
    const uint8_t* base = bytes + byte0;
    const int64_t bit0 = (int64_t)(entry >> 16);
    int c = (int)(entry >> 8) & 0xff, z = (int)entry & 0xff;
    BitReader br;
    br.p = base + (bit0 >> 3); br.end = bytes + byte1; br.acc = 0; br.n = 0; br.fake = 0; br.hit_marker = false;
    br.init();
    br.fill();
    br.skip((int)(bit0 & 7));
    const int hs = d.hs, vs = d.vs, nlum = hs * vs, bpm = nlum + d.ncomp - 1, mcux = d.mcux;
    const HuffTab* dctab[3] = {&ts.dc[d.dc_sel[0]], &ts.dc[d.dc_sel[1]], &ts.dc[d.dc_sel[2]]};
    const HuffTab* actab[3] = {&ts.ac[d.ac_sel[0]], &ts.ac[d.ac_sel[1]], &ts.ac[d.ac_sel[2]]};
    int comp = c < nlum ? 0 : c - nlum + 1;
    int64_t blk_i = blk0;                                    // index (inside the interval) of the block in progress
    int mx = 0, my = 0;
    int16_t* blk = nullptr;
    auto locate = [&]() {                                    // WRITE: the block in progress -> its place in the coefficient buffer
        const int ch = comp == 0 ? hs : 1, cv = comp == 0 ? vs : 1;
        const int by = comp == 0 ? c / hs : 0, bx = comp == 0 ? c - by * hs : 0;
        blk = coef + d.coef_off[comp] + ((int64_t)(my * cv + by) * d.bw[comp] + (mx * ch + bx)) * 64;
    };
    if (WRITE) {
        const int mcu = mcu0 + (int)(blk0 / bpm);
        my = mcu / mcux; mx = mcu - my * mcux;
        locate();
    }
    int done = 0;
    int64_t pos = bit0;
    for (;;) {
        if (WRITE && blk_i >= blk_end) break;
        if ((int64_t)(br.p - base) * 8 - (br.n - br.fake) >= end_bit) {      // an upper bound of the position: the exact one only near the end
            pos = raw_bitpos(br, base);
            if (pos >= end_bit) break;
        }
        // one symbol per iteration, the same instruction path for DC and AC symbols (a warp's lanes are in different states)
        const bool dc = z == 0;
        const int sym = huff_symbol(br, dc ? *dctab[comp] : *actab[comp]);
        const int r = dc ? 0 : sym >> 4, s = dc ? sym : sym & 15;
        const int zi = z + r;                                // zig-zag index of the coefficient this symbol carries (DC: 0)
        const bool has_value = s != 0 && zi <= 63;           // run past the block's end (corrupt / speculative garbage): no value bits, block over
        if (has_value) {
            if (br.n < 16) br.fill();
            const int v = (int)br.peek(s);
            br.skip(s);
            if (WRITE) {                                                  // DC: the DIFFERENCE (a zero difference is what the buffers hold)
                if (dc && dcbuf != nullptr) dcbuf[(blk - coef) >> 6] = (int16_t)huff_extend(v, s);
                else blk[zz[zi]] = (int16_t)huff_extend(v, s);
            }
        }
        const bool eob = !dc && s == 0 && r != 15;
        z = (!dc && s == 0) ? z + 16 : zi + 1;               // ZRL skips 16; everything else moves behind the coefficient
        const bool block_done = eob || z > 63;
        if (block_done) {
            z = 0; ++done; ++blk_i;
            if (++c == bpm) { c = 0; if (++mx == mcux) { mx = 0; ++my; } }
            comp = c < nlum ? 0 : c - nlum + 1;
            if (WRITE) locate();
        }
    }
    if (WRITE && blk_i >= blk_end) pos = raw_bitpos(br, base);
    *n_done = done;
    return pack_state(pos, c, z);
}

// the guess a chunk starts from in round 0: its first byte (past a stuffed zero), state 0
__host__ __device__ inline uint64_t chunk_guess(const uint8_t* base, int64_t byte_off) {
    if (byte_off > 0 && base[byte_off] == 0 && base[byte_off - 1] == 0xFF) ++byte_off;
    return pack_state(byte_off * 8, 0, 0);
}

// DC differences -> DC values: blocks of component `comp` of an interval in decode order (MCU by MCU, rows of the MCU, columns)
__host__ __device__ inline int16_t* dc_block(const JpegDesc& d, int16_t* coef, int comp, int mcu0, int64_t j, int stride = 64 /* 1: compact DC buffer */) {
    const int ch = comp == 0 ? d.hs : 1, cv = comp == 0 ? d.vs : 1, nb = ch * cv;
    const int mcu = mcu0 + (int)(j / nb), sub = (int)(j % nb);
    const int my = mcu / d.mcux, mx = mcu - my * d.mcux, by = sub / ch, bx = sub - by * ch;
    return coef + (d.coef_off[comp] / 64 + ((int64_t)(my * cv + by) * d.bw[comp] + (mx * ch + bx))) * stride;
}

// One thread per interval.  The Huffman table sets of the batch (files written by one encoder share a set) are copied into shared
// memory when at most `smem_sets` of them exist: a symbol is then two shared-memory look-ups instead of two L2 round trips.
constexpr int kHuffThreads = 64;
__global__ void __launch_bounds__(kHuffThreads) jpeg_huffman_kernel(const JpegDesc* __restrict__ descs, const TableSet* __restrict__ tsets, int n_sets,
                                                                    int smem_sets, const Interval* __restrict__ iv, int n_iv,
                                                                    const uint8_t* __restrict__ bytes, int16_t* __restrict__ coef,
                                                                    const int* __restrict__ only_failed /* non-null: intervals with a 0 here are skipped */,
                                                                    int16_t* __restrict__ dcbuf /* non-null: DC values also go there */) {
    extern __shared__ __align__(16) uint8_t huff_smem[];
    __shared__ uint8_t zz[64];
    if (smem_sets > 0) {
        const uint4* src = reinterpret_cast<const uint4*>(tsets);
        uint4* dst = reinterpret_cast<uint4*>(huff_smem);
        for (int i = threadIdx.x; i < (int)(n_sets * sizeof(TableSet) / 16); i += kHuffThreads) dst[i] = src[i];
    }
    if (threadIdx.x < 64) zz[threadIdx.x] = kZigzagDev[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_iv) return;
    if (only_failed != nullptr && only_failed[i] == 0) return;
    const Interval v = iv[i];
    const JpegDesc& d = descs[v.image];
    const TableSet& ts = smem_sets > 0 ? reinterpret_cast<const TableSet*>(huff_smem)[d.tset] : tsets[d.tset];
    huff_decode_interval(d, ts, bytes, v.byte0, v.byte1, v.mcu0, v.n_mcu, coef, zz, dcbuf);
}

// Chunked decoding (see huff_decode_chunk): one thread per chunk.  Per-chunk state lives in four arrays of the batch.
struct ChunkArrays {
    uint64_t* used_entry;         // the entry state a chunk was last decoded from
    uint64_t* exit_state;         // where that decode ended = the entry of the next chunk
    int32_t* count;               // blocks it completed
    int32_t* blk0;                // blocks of the interval completed before the chunk (prefix sum, jpeg_chunk_scan_kernel)
    int32_t* failed;              // per interval: the chain did not close in the rounds given
    int16_t* dcbuf;               // DC of block k of the batch (differences, then values): the IDCT kernel reads it instead of coefficient 0
};
constexpr int kChunkThreads = 128;

__device__ __forceinline__ void load_table_sets(const TableSet* tsets, int n_sets, int smem_sets, uint8_t* huff_smem, uint8_t* zz) {
    if (smem_sets > 0) {
        const uint4* src = reinterpret_cast<const uint4*>(tsets);
        uint4* dst = reinterpret_cast<uint4*>(huff_smem);
        for (int i = threadIdx.x; i < (int)(n_sets * sizeof(TableSet) / 16); i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x < 64) zz[threadIdx.x] = kZigzagDev[threadIdx.x];
    __syncthreads();
}

// round 0: every chunk but the last of its interval from its guess; later rounds: only chunks whose predecessor's exit state changed
__global__ void __launch_bounds__(kChunkThreads) jpeg_spec_kernel(const JpegDesc* __restrict__ descs, const TableSet* __restrict__ tsets, int n_sets,
                                                                  int smem_sets, const Interval* __restrict__ iv, const IvChunks* __restrict__ ivc,
                                                                  const int32_t* __restrict__ chunk_iv, int n_chunks, const uint8_t* __restrict__ bytes,
                                                                  int round, ChunkArrays a) {
    extern __shared__ __align__(16) uint8_t huff_smem[];
    __shared__ uint8_t zz[64];
    load_table_sets(tsets, n_sets, smem_sets, huff_smem, zz);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks) return;
    const int vi = chunk_iv[g];
    const IvChunks ic = ivc[vi];
    const int i = g - ic.chunk0;
    if (i == ic.n_chunks - 1) return;                        // nobody needs the exit state of the last chunk
    const Interval v = iv[vi];
    uint64_t entry;
    if (i == 0) entry = pack_state(0, 0, 0);
    else if (round == 0) entry = chunk_guess(bytes + v.byte0, (int64_t)i * ic.chunk_bytes);
    else entry = a.exit_state[g - 1];                        // a neighbour may be rewriting it right now: either value is a legal entry
    if (round > 0 && entry == a.used_entry[g]) return;
    const JpegDesc& d = descs[v.image];
    const TableSet& ts = smem_sets > 0 ? reinterpret_cast<const TableSet*>(huff_smem)[d.tset] : tsets[d.tset];
    int done = 0;
    const uint64_t ex = huff_decode_chunk<false>(d, ts, bytes, v.byte0, v.byte1, entry, (int64_t)(i + 1) * ic.chunk_bytes * 8, v.mcu0, 0, 0, nullptr, zz, &done);
    a.used_entry[g] = entry;
    a.exit_state[g] = ex;
    a.count[g] = done;
}

// per interval: does the chain close (entry(i) == exit(i-1))?  Block offsets of the chunks.
__global__ void jpeg_chunk_scan_kernel(const IvChunks* __restrict__ ivc, int n_iv, ChunkArrays a) {
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= n_iv) return;
    const IvChunks ic = ivc[vi];
    int ok = 1, run = 0;
    for (int i = 0; i < ic.n_chunks; ++i) {
        const int g = ic.chunk0 + i;
        a.blk0[g] = run;
        if (i < ic.n_chunks - 1) {
            if (i > 0 && a.used_entry[g] != a.exit_state[g - 1]) ok = 0;
            run += a.count[g];
        }
    }
    a.failed[vi] = !ok;
}

__global__ void __launch_bounds__(kChunkThreads) jpeg_write_kernel(const JpegDesc* __restrict__ descs, const TableSet* __restrict__ tsets, int n_sets,
                                                                   int smem_sets, const Interval* __restrict__ iv, const IvChunks* __restrict__ ivc,
                                                                   const int32_t* __restrict__ chunk_iv, int n_chunks, const uint8_t* __restrict__ bytes,
                                                                   ChunkArrays a, int16_t* __restrict__ coef) {
    extern __shared__ __align__(16) uint8_t huff_smem[];
    __shared__ uint8_t zz[64];
    load_table_sets(tsets, n_sets, smem_sets, huff_smem, zz);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks) return;
    const int vi = chunk_iv[g];
    if (a.failed[vi]) return;                                // the one-thread kernel decodes this interval
    const IvChunks ic = ivc[vi];
    const int i = g - ic.chunk0;
    const Interval v = iv[vi];
    const JpegDesc& d = descs[v.image];
    const TableSet& ts = smem_sets > 0 ? reinterpret_cast<const TableSet*>(huff_smem)[d.tset] : tsets[d.tset];
    const uint64_t entry = i == 0 ? pack_state(0, 0, 0) : a.exit_state[g - 1];
    const int64_t end_bit = i == ic.n_chunks - 1 ? ((int64_t)1 << 46) : (int64_t)(i + 1) * ic.chunk_bytes * 8;
    const int64_t blk_end = (int64_t)v.n_mcu * (d.hs * d.vs + d.ncomp - 1);
    int done = 0;
    huff_decode_chunk<true>(d, ts, bytes, v.byte0, v.byte1, entry, end_bit, v.mcu0, a.blk0[g], blk_end, coef, zz, &done, a.dcbuf);
}

// DC differences -> DC values: one warp per (interval, component); a lane sums a run of consecutive blocks, the warp scans the run sums
__global__ void __launch_bounds__(128) jpeg_dc_scan_kernel(const JpegDesc* __restrict__ descs, const Interval* __restrict__ iv, int n_iv,
                                                           const int32_t* __restrict__ failed, int16_t* __restrict__ dcbuf) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int vi = w / 3, comp = w - vi * 3;
    if (vi >= n_iv || failed[vi]) return;                    // the one-thread kernel wrote absolute values
    const Interval v = iv[vi];
    const JpegDesc& d = descs[v.image];
    if (comp >= d.ncomp) return;
    const int64_t total = (int64_t)v.n_mcu * (comp == 0 ? d.hs * d.vs : 1);
    const int64_t per = (total + 31) / 32, j0 = lane * per, j1 = j0 + per < total ? j0 + per : total;
    int sum = 0;
    for (int64_t j = j0; j < j1; ++j) sum += dc_block(d, dcbuf, comp, v.mcu0, j, 1)[0];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int run = incl - sum;
    for (int64_t j = j0; j < j1; ++j) {
        int16_t* b = dc_block(d, dcbuf, comp, v.mcu0, j, 1);
        run += b[0];
        b[0] = (int16_t)run;
    }
}

// ---- jidctint.c jpeg_idct_islow --------------------------------------------------------------------------------------------------------
#define CONST_BITS 13
#define PASS1_BITS 2
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
// the 8-point pass: d[0..7] -> out[0..7], already descaled by `shift` (32-bit arithmetic: |dequantised coefficient| < 2^15 by construction of
// JPEG's 11-bit DCT range x 8-bit tables holds for baseline files; the SIMD code Pillow actually runs works in 16/32 bits as well)
__device__ __forceinline__ void idct8(const int (&d)[8], int (&o)[8], int shift) {
    int z2 = d[2], z3 = d[6];
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137);
    int tmp3 = z1 + z2 * 6270;
    z2 = d[0]; z3 = d[4];
    int tmp0 = (z2 + z3) << CONST_BITS;
    int tmp1 = (z2 - z3) << CONST_BITS;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = d[7]; tmp1 = d[5]; tmp2 = d[3]; tmp3 = d[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
    z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    o[0] = descale(tmp10 + tmp3, shift); o[7] = descale(tmp10 - tmp3, shift);
    o[1] = descale(tmp11 + tmp2, shift); o[6] = descale(tmp11 - tmp2, shift);
    o[2] = descale(tmp12 + tmp1, shift); o[5] = descale(tmp12 - tmp1, shift);
    o[3] = descale(tmp13 + tmp0, shift); o[4] = descale(tmp13 - tmp0, shift);
}
// sample_range_limit + CENTERJSAMPLE indexed with (v & RANGE_MASK) (jdmaster.c prepare_range_limit_table)
__device__ __forceinline__ uint32_t range_limit(int v) {
    const int i = v & 1023;
    return i < 128 ? i + 128 : i < 512 ? 255 : i < 896 ? 0 : i - 896;
}

struct IdctJob { int32_t image, comp; int64_t first_block, n_blocks; };      // blocks of one component of one image, consecutive in `coef`

// 256 threads = 32 blocks x 8 threads.  Blocks are addressed through a flat (image, component) job list: block g of the batch.
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegDesc* __restrict__ descs, const int64_t* __restrict__ job_start /*[n_jobs+1]*/,
                                                        const IdctJob* __restrict__ jobs, int n_jobs, const int16_t* __restrict__ coef,
                                                        const int16_t* __restrict__ dcbuf /* non-null: DC of block k of the batch = dcbuf[k] */,
                                                        uint8_t* __restrict__ planes) {
    __shared__ int ws[32][8][9];
    const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
    const int64_t g = (int64_t)blockIdx.x * 32 + lb;
    const bool active = g < job_start[n_jobs];
    int lo = 0;
    if (active) {                                            // binary search of the job that owns block g
        int hi = n_jobs;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (job_start[mid] <= g) lo = mid; else hi = mid;
        }
    }
    const IdctJob job = jobs[active ? lo : 0];
    const JpegDesc& d = descs[job.image];
    const int64_t b = g - job_start[active ? lo : 0];        // block index inside the component
    if (active) {                                            // row t of the block: one 16-byte load (a block is 128 contiguous bytes), dequantised
        const int16_t* cb = coef + (job.first_block + b) * 64;
        const uint4 raw = *reinterpret_cast<const uint4*>(cb + t * 8);
        const uint2* qrow = reinterpret_cast<const uint2*>(d.qt[d.q_sel[job.comp]] + t * 8);       // the tables are 8-byte aligned inside JpegDesc
        const uint2 q0 = qrow[0], q1 = qrow[1];
        const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w}, qq[4] = {q0.x, q0.y, q1.x, q1.y};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            ws[lb][t][2 * c] = (int)(int16_t)(rw[c] & 0xffffu) * (int)(qq[c] & 0xffffu);
            ws[lb][t][2 * c + 1] = (int)(int16_t)(rw[c] >> 16) * (int)(qq[c] >> 16);
        }
        if (dcbuf != nullptr && t == 0) ws[lb][0][0] = (int)dcbuf[job.first_block + b] * (int)(qq[0] & 0xffffu);
    }
    __syncthreads();
    if (active) {                                            // pass 1: column t (read and written back by the same thread)
        int in[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = ws[lb][r][t];
        idct8(in, o, CONST_BITS - PASS1_BITS);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[lb][r][t] = o[r];
    }
    __syncthreads();
    if (active) {                                            // pass 2: row t -> 8 samples
        int in[8], o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) in[c] = ws[lb][t][c];
        idct8(in, o, CONST_BITS + PASS1_BITS + 3);
        const int bw = d.bw[job.comp];
        const int by = (int)(b / bw), bx = (int)(b - (int64_t)by * bw);
        uint8_t* dst = planes + d.plane_off[job.comp] + ((int64_t)(by * 8 + t) * bw + bx) * 8;
        const uint32_t w0 = range_limit(o[0]) | (range_limit(o[1]) << 8) | (range_limit(o[2]) << 16) | (range_limit(o[3]) << 24);
        const uint32_t w1 = range_limit(o[4]) | (range_limit(o[5]) << 8) | (range_limit(o[6]) << 16) | (range_limit(o[7]) << 24);
        *reinterpret_cast<uint2*>(dst) = make_uint2(w0, w1);
    }
}

// ---- jdsample.c fancy upsampling (on the fly) + jdcolor.c ycc_rgb_convert ---------------------------------------------------------------
// One thread per output pixel.  Chroma planes have pitch bw*8; their REAL extent is (dw, dh) = ceil(width / hs), ceil(height / vs): the
// triangle filters replicate the first / last real row and column (jdmainct.c context rows, the edge cases of h2v*_fancy_upsample).
__device__ __forceinline__ int chroma_at(const uint8_t* pl, int pitch, int dw, int dh, int hs, int vs, int x, int y) {
    const int cx = hs == 2 ? x >> 1 : x, cy = vs == 2 ? y >> 1 : y;
    const bool fancy_h = hs == 2 && dw > 2;                  // jinit_upsampler: fancy only when downsampled_width > 2
    if (hs == 2 && vs == 2) {
        if (!fancy_h) return pl[cy * pitch + cx];
        const int oy = (y & 1) ? min(cy + 1, dh - 1) : max(cy - 1, 0);
        auto colsum = [&](int c) { return 3 * (int)pl[cy * pitch + c] + (int)pl[oy * pitch + c]; };
        const int cur = colsum(cx);
        if (x & 1) return cx == dw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + colsum(cx + 1) + 7) >> 4;
        return cx == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + colsum(cx - 1) + 8) >> 4;
    }
    if (hs == 2) {                                           // h2v1
        const int cur = pl[cy * pitch + cx];
        if (!fancy_h) return cur;
        if (x & 1) return cx == dw - 1 ? cur : (3 * cur + (int)pl[cy * pitch + cx + 1] + 2) >> 2;
        return cx == 0 ? cur : (3 * cur + (int)pl[cy * pitch + cx - 1] + 1) >> 2;
    }
    if (vs == 2) {                                           // h1v2 (libjpeg-turbo h1v2_fancy_upsample)
        const int cur = pl[cy * pitch + cx];
        if (y & 1) return (3 * cur + (int)pl[min(cy + 1, dh - 1) * pitch + cx] + 2) >> 2;
        return (3 * cur + (int)pl[max(cy - 1, 0) * pitch + cx] + 1) >> 2;
    }
    return pl[cy * pitch + cx];
}
#define FIX16(x) ((int)((x) * 65536.0 + 0.5))
__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegDesc* __restrict__ descs, const uint8_t* __restrict__ planes, int W, int H,
                                                         uint8_t* __restrict__ rgb) {
    const int img = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const JpegDesc& d = descs[img];
    const int y = i / W, x = i - y * W;
    const int Y = planes[d.plane_off[0] + (int64_t)y * d.bw[0] * 8 + x];
    int r = Y, g = Y, b = Y;
    if (d.ncomp == 3) {
        const int dw = (d.width + d.hs - 1) / d.hs, dh = (d.height + d.vs - 1) / d.vs;
        const int cb = chroma_at(planes + d.plane_off[1], d.bw[1] * 8, dw, dh, d.hs, d.vs, x, y) - 128;
        const int cr = chroma_at(planes + d.plane_off[2], d.bw[2] * 8, dw, dh, d.hs, d.vs, x, y) - 128;
        r = Y + ((FIX16(1.40200) * cr + 32768) >> 16);
        g = Y + ((-FIX16(0.34414) * cb + 32768 - FIX16(0.71414) * cr) >> 16);
        b = Y + ((FIX16(1.77200) * cb + 32768) >> 16);
        r = min(max(r, 0), 255); g = min(max(g, 0), 255); b = min(max(b, 0), 255);
    }
    uint8_t* o = rgb + ((int64_t)img * H * W + i) * 3;
    o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b;
}

// Four horizontally consecutive pixels per thread (image width a multiple of 4): the luma bytes are one word, the h2v2 column sums
// 3 * near + far are computed once for the four chroma columns the pixels touch (17 loads per 4 pixels instead of 36), and the twelve
// output bytes leave as three words.  Other samplings take chroma_at() per pixel.  Same arithmetic as jpeg_color_kernel.
__device__ __forceinline__ void ycc_to_rgb(int Y, int cb, int cr, int& r, int& g, int& b) {
    r = Y + ((FIX16(1.40200) * cr + 32768) >> 16);
    g = Y + ((-FIX16(0.34414) * cb + 32768 - FIX16(0.71414) * cr) >> 16);
    b = Y + ((FIX16(1.77200) * cb + 32768) >> 16);
    r = min(max(r, 0), 255); g = min(max(g, 0), 255); b = min(max(b, 0), 255);
}
__device__ __forceinline__ void h2v2_fancy4(const uint8_t* pl, int pitch, int dw, int dh, int x0, int y, int (&out)[4]) {
    const int cy = y >> 1, oy = (y & 1) ? min(cy + 1, dh - 1) : max(cy - 1, 0), cx0 = x0 >> 1;
    const uint8_t* near = pl + cy * pitch;
    const uint8_t* far = pl + oy * pitch;
    int cs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = min(max(cx0 - 1 + k, 0), dw - 1);
        cs[k] = 3 * (int)near[c] + (int)far[c];
    }
    out[0] = cx0 == 0 ? (cs[1] * 4 + 8) >> 4 : (cs[1] * 3 + cs[0] + 8) >> 4;
    out[1] = cx0 == dw - 1 ? (cs[1] * 4 + 7) >> 4 : (cs[1] * 3 + cs[2] + 7) >> 4;
    out[2] = (cs[2] * 3 + cs[1] + 8) >> 4;
    out[3] = cx0 + 1 == dw - 1 ? (cs[2] * 4 + 7) >> 4 : (cs[2] * 3 + cs[3] + 7) >> 4;
}
__global__ void __launch_bounds__(256) jpeg_color4_kernel(const JpegDesc* __restrict__ descs, const uint8_t* __restrict__ planes, int W, int H,
                                                          uint8_t* __restrict__ rgb) {
    const int img = blockIdx.y;
    const int g4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (g4 >= (W * H) >> 2) return;
    const JpegDesc& d = descs[img];
    const int i0 = g4 << 2, y = i0 / W, x0 = i0 - y * W;
    const uint32_t yw = *reinterpret_cast<const uint32_t*>(planes + d.plane_off[0] + (int64_t)y * d.bw[0] * 8 + x0);
    int r[4], g[4], b[4];
    if (d.ncomp == 3) {
        const int dw = (d.width + d.hs - 1) / d.hs, dh = (d.height + d.vs - 1) / d.vs;
        int cb[4], cr[4];
        if (d.hs == 2 && d.vs == 2 && dw > 2) {
            h2v2_fancy4(planes + d.plane_off[1], d.bw[1] * 8, dw, dh, x0, y, cb);
            h2v2_fancy4(planes + d.plane_off[2], d.bw[2] * 8, dw, dh, x0, y, cr);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cb[k] = chroma_at(planes + d.plane_off[1], d.bw[1] * 8, dw, dh, d.hs, d.vs, x0 + k, y);
                cr[k] = chroma_at(planes + d.plane_off[2], d.bw[2] * 8, dw, dh, d.hs, d.vs, x0 + k, y);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) ycc_to_rgb((int)((yw >> (8 * k)) & 0xffu), cb[k] - 128, cr[k] - 128, r[k], g[k], b[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = g[k] = b[k] = (int)((yw >> (8 * k)) & 0xffu);
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(rgb + ((int64_t)img * H * W + i0) * 3);
    o[0] = (uint32_t)r[0] | ((uint32_t)g[0] << 8) | ((uint32_t)b[0] << 16) | ((uint32_t)r[1] << 24);
    o[1] = (uint32_t)g[1] | ((uint32_t)b[1] << 8) | ((uint32_t)r[2] << 16) | ((uint32_t)g[2] << 24);
    o[2] = (uint32_t)b[2] | ((uint32_t)r[3] << 8) | ((uint32_t)g[3] << 16) | ((uint32_t)b[3] << 24);
}

// ---- host: header parsing (jdmarker.c) ----------------------------------------------------------------------------------------------------
void build_hufftab(const uint8_t* counts, const uint8_t* vals, int nvals, HuffTab* t) {
    memset(t, 0, sizeof(*t));
    memcpy(t->vals, vals, (size_t)std::min(nvals, 256));
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        t->valoffset[l] = k - code;
        if (counts[l - 1]) {
            for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code)
                if (l <= 8) {                                // every 8-bit pattern starting with this code
                    const int first = code << (8 - l);
                    for (int f = 0; f < (1 << (8 - l)); ++f) { t->look_nbits[first + f] = (uint8_t)l; t->look_sym[first + f] = vals[k]; }
                }
            t->maxcode[l] = code - 1;
        } else {
            t->maxcode[l] = -1;
        }
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
}

// Returns nullptr on success, otherwise the reason the file is not handled here (the caller may then decode it another way).
const char* parse_jpeg(const uint8_t* f, size_t size, JpegDesc* d, TableSet* ts) {
    memset(d, 0, sizeof(*d));
    memset(ts, 0, sizeof(*ts));
    if (size < 4 || f[0] != 0xFF || f[1] != 0xD8) return "not a JPEG (no SOI)";
    size_t pos = 2;
    bool have_frame = false, have_q[4] = {false, false, false, false}, have_dc[4] = {}, have_ac[4] = {};
    int comp_id[3] = {0, 0, 0}, comp_h[3] = {1, 1, 1}, comp_v[3] = {1, 1, 1};
    int adobe_transform = -1;
    while (true) {
        if (pos + 4 > size) return "truncated before SOS";
        if (f[pos] != 0xFF) return "marker expected";
        while (pos + 1 < size && f[pos + 1] == 0xFF) ++pos;
        const int m = f[pos + 1];
        pos += 2;
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (pos + 2 > size) return "truncated segment";
        const size_t L = ((size_t)f[pos] << 8) | f[pos + 1];
        if (L < 2 || pos + L > size) return "truncated segment";
        const uint8_t* seg = f + pos + 2;
        const size_t n = L - 2;
        if (m == 0xDB) {
            for (size_t i = 0; i < n;) {
                const int pq = seg[i] >> 4, tq = seg[i] & 15;
                ++i;
                if (tq > 3 || i + (pq ? 128 : 64) > n) return "bad DQT";
                for (int k = 0; k < 64; ++k) {
                    const int v = pq ? ((seg[i + 2 * k] << 8) | seg[i + 2 * k + 1]) : seg[i + k];
                    d->qt[tq][zigzag_tab()[k]] = (uint16_t)v;
                }
                i += pq ? 128 : 64;
                have_q[tq] = true;
            }
        } else if (m == 0xC4) {
            for (size_t i = 0; i < n;) {
                if (i + 17 > n) return "bad DHT";
                const int tc = seg[i] >> 4, th = seg[i] & 15;
                int total = 0;
                for (int k = 0; k < 16; ++k) total += seg[i + 1 + k];
                if (th > 3 || tc > 1 || total > 256 || i + 17 + total > n) return "bad DHT";
                build_hufftab(seg + i + 1, seg + i + 17, total, tc == 0 ? &ts->dc[th] : &ts->ac[th]);
                (tc == 0 ? have_dc : have_ac)[th] = true;
                i += 17 + total;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (n < 6 || seg[0] != 8) return "only 8-bit samples";
            d->height = (seg[1] << 8) | seg[2];
            d->width = (seg[3] << 8) | seg[4];
            d->ncomp = seg[5];
            if (d->ncomp != 1 && d->ncomp != 3) return "1 or 3 components only (CMYK?)";
            if (n < (size_t)(6 + 3 * d->ncomp) || d->width <= 0 || d->height <= 0) return "bad SOF";
            for (int c = 0; c < d->ncomp; ++c) {
                comp_id[c] = seg[6 + 3 * c];
                comp_h[c] = seg[7 + 3 * c] >> 4;
                comp_v[c] = seg[7 + 3 * c] & 15;
                d->q_sel[c] = seg[8 + 3 * c];
                if (d->q_sel[c] > 3) return "bad quantisation table selector";
            }
            have_frame = true;
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
            return "progressive / lossless / arithmetic JPEG";
        } else if (m == 0xDD) {
            if (n < 2) return "bad DRI";
            d->restart = (seg[0] << 8) | seg[1];
        } else if (m == 0xEE && n >= 12 && memcmp(seg, "Adobe", 5) == 0) {
            adobe_transform = seg[11];
        } else if (m == 0xDA) {
            if (!have_frame) return "SOS before SOF";
            if (n < 1 || seg[0] != d->ncomp || n < (size_t)(1 + 2 * d->ncomp)) return "non-interleaved multi-scan file";
            for (int k = 0; k < d->ncomp; ++k) {
                int c = -1;
                for (int j = 0; j < d->ncomp; ++j) if (comp_id[j] == seg[1 + 2 * k]) c = j;
                if (c != k) return "scan components out of frame order";
                d->dc_sel[c] = seg[2 + 2 * k] >> 4;
                d->ac_sel[c] = seg[2 + 2 * k] & 15;
                if (d->dc_sel[c] > 3 || d->ac_sel[c] > 3 || !have_dc[d->dc_sel[c]] || !have_ac[d->ac_sel[c]] || !have_q[d->q_sel[c]]) return "missing table";
            }
            if (d->ncomp == 3) {
                if (adobe_transform == 0 || (adobe_transform < 0 && comp_id[0] == 82 && comp_id[1] == 71 && comp_id[2] == 66)) return "RGB-coded JPEG";
                if ((comp_h[0] != 1 && comp_h[0] != 2) || (comp_v[0] != 1 && comp_v[0] != 2) || comp_h[1] != 1 || comp_v[1] != 1 || comp_h[2] != 1 || comp_v[2] != 1)
                    return "sampling factors other than full-resolution luma with 1x1 chroma";
                d->hs = comp_h[0];
                d->vs = comp_v[0];
            } else {
                d->hs = d->vs = 1;                            // a single-component scan is never interleaved: 1 block per MCU
            }
            d->mcux = (d->width + 8 * d->hs - 1) / (8 * d->hs);
            d->mcuy = (d->height + 8 * d->vs - 1) / (8 * d->vs);
            for (int c = 0; c < d->ncomp; ++c) {
                d->bw[c] = d->mcux * (c == 0 ? d->hs : 1);
                d->bh[c] = d->mcuy * (c == 0 ? d->vs : 1);
            }
            d->data_off = (int64_t)(pos + L);
            d->data_len = (int64_t)size - d->data_off;
            return nullptr;
        }
        pos += L;
    }
}

// restart intervals of one file: byte ranges between RSTn markers (a stuffed 0xFF00 is data, any other marker ends the scan)
void split_intervals(const uint8_t* f, const JpegDesc& d, int image, int64_t base, std::vector<Interval>* out) {
    const int total = d.mcux * d.mcuy;
    const int64_t end = d.data_off + d.data_len;
    if (d.restart <= 0) {
        out->push_back(Interval{image, 0, total, base + d.data_off, base + end});
        return;
    }
    int64_t p = d.data_off, start = d.data_off;
    int mcu = 0;
    while (mcu < total) {
        while (p + 1 < end && !(f[p] == 0xFF && f[p + 1] != 0)) ++p;          // next marker (or the end)
        const bool rst = p + 1 < end && f[p + 1] >= 0xD0 && f[p + 1] <= 0xD7;
        const int n = std::min(d.restart, total - mcu);
        out->push_back(Interval{image, mcu, n, base + start, base + (p + 1 < end ? p : end)});
        mcu += n;
        if (!rst) {                                           // EOI / garbage: remaining MCUs (if any) decode from an empty range (zeros)
            if (mcu < total) out->push_back(Interval{image, mcu, total - mcu, base + end, base + end});
            break;
        }
        p += 2;
        start = p;
    }
}

struct Batch {                     // host-side layout of one decode call
    std::vector<JpegDesc> descs;
    std::vector<TableSet> tsets;                             // distinct Huffman table sets of the batch
    std::vector<Interval> intervals;
    std::vector<IdctJob> jobs;
    std::vector<int64_t> job_start, file_base;               // file_base[i]: offset of file i inside the staged byte buffer
    std::vector<IvChunks> ivc;                               // chunked entropy decoding: chunks of every interval ...
    std::vector<int32_t> chunk_iv;                           // ... and the interval of every chunk
    int64_t bytes = 0, coef_elems = 0, plane_bytes = 0;
};

// Chunk size of an interval: what resynchronises is the bit stream (a few symbols) AND the block / component phase (a few MCUs), so the
// size follows the interval's bytes per MCU: chunk_mcus MCUs' worth, at least min_bytes (chunk_mcus = 0: min_bytes for everyone).
// Measured on the host (tests/test_jpeg_cpu.py): 8 MCUs close the chain in 2-4 rounds from board files (37 B / MCU) to noise (230 B / MCU).
void plan_chunks(Batch* b, int min_bytes, int chunk_mcus) {
    b->ivc.clear(); b->chunk_iv.clear();
    for (size_t i = 0; i < b->intervals.size(); ++i) {
        const Interval& v = b->intervals[i];
        const int64_t len = v.byte1 - v.byte0;
        int64_t cb = min_bytes;
        if (chunk_mcus > 0 && v.n_mcu > 0) cb = std::max<int64_t>(cb, ((int64_t)chunk_mcus * len / v.n_mcu + 63) / 64 * 64);
        cb = std::min<int64_t>(cb, 1 << 20);
        const int n = (int)std::max<int64_t>(1, (len + cb - 1) / cb);
        b->ivc.push_back(IvChunks{(int32_t)b->chunk_iv.size(), n, (int32_t)cb});
        b->chunk_iv.insert(b->chunk_iv.end(), n, (int32_t)i);
    }
}

// defaults of the chunked decoder (experiment builds: CV_JPEG_CHUNK bytes, CV_JPEG_MCUS, CV_JPEG_ROUNDS; CV_JPEG_CHUNK=0 = the one-thread
// kernel).  A round in which nothing changed costs a chunk two loads, so spare rounds are nearly free.
constexpr int kChunkBytes = 256, kChunkMcus = 8, kSpecRounds = 6;

// fn(first, last) over [0, n) on up to 8 host threads (parsing and staging a few thousand files is milliseconds of memory-bound work)
template <typename F>
void parallel_ranges(int n, int min_per_thread, F fn) {
    int t = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    t = std::min(t, std::max(1, n / std::max(1, min_per_thread)));
    if (t <= 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    const int per = (n + t - 1) / t;
    for (int k = 1; k < t; ++k) pool.emplace_back(fn, std::min(n, k * per), std::min(n, (k + 1) * per));
    fn(0, std::min(n, per));
    for (auto& th : pool) th.join();
}

int plan_batch(const uint8_t* const* files, const size_t* sizes, int n, int W, int H, Batch* b) {
    b->descs.resize(n);
    b->job_start.push_back(0);
    // headers in parallel; a file whose Huffman tables equal those of the file before it (the usual case: one encoder) is only marked so,
    // the others leave their set in ts_new for the sequential pass below (7.3 KB per set: not one per file)
    std::vector<const char*> why_of(n, nullptr);
    std::vector<int32_t> ts_ref(n, -1);                      // index into ts_new, or -1 = same tables as file i - 1
    std::vector<std::vector<TableSet>> ts_new_of;
    ts_new_of.reserve(16);
    for (int k = 0; k < 16; ++k) ts_new_of.emplace_back();
    std::atomic<int> next_slot{0};
    std::vector<int32_t> slot_of(n, 0);
    parallel_ranges(n, 128, [&](int i0, int i1) {
        const int slot = next_slot.fetch_add(1);
        std::vector<TableSet>& mine = ts_new_of[slot];
        TableSet ts;
        for (int i = i0; i < i1; ++i) {
            why_of[i] = parse_jpeg(files[i], sizes[i], &b->descs[i], &ts);
            slot_of[i] = slot;
            if (why_of[i]) continue;
            if (i > i0 && !mine.empty() && memcmp(&mine.back(), &ts, sizeof(ts)) == 0) { ts_ref[i] = -1; continue; }
            mine.push_back(ts);
            ts_ref[i] = (int32_t)mine.size() - 1;
        }
    });
    for (int i = 0; i < n; ++i) {
        JpegDesc& d = b->descs[i];
        if (why_of[i]) { cv_set_error("cv_jpeg_decode: file %d: %s", i, why_of[i]); return CV_ERR_ARG; }
        if (ts_ref[i] < 0) {
            d.tset = b->descs[i - 1].tset;                    // same tables as the file before (same parsing thread)
        } else {
            const TableSet& ts = ts_new_of[slot_of[i]][ts_ref[i]];
            d.tset = -1;
            for (int k = (int)b->tsets.size() - 1; k >= 0 && k >= (int)b->tsets.size() - 8 && d.tset < 0; --k)      // look at the most recent sets
                if (memcmp(&b->tsets[k], &ts, sizeof(ts)) == 0) d.tset = k;
            if (d.tset < 0) { b->tsets.push_back(ts); d.tset = (int)b->tsets.size() - 1; }
        }
        if (d.width != W || d.height != H) { cv_set_error("cv_jpeg_decode: file %d is %dx%d, the batch is %dx%d", i, d.width, d.height, W, H); return CV_ERR_ARG; }
        b->file_base.push_back(b->bytes);
        split_intervals(files[i], d, i, b->bytes, &b->intervals);
        d.data_off += b->bytes;                               // offsets become batch-relative
        b->bytes += (int64_t)((sizes[i] + 15) & ~(size_t)15);
        for (int c = 0; c < d.ncomp; ++c) {
            const int64_t blocks = (int64_t)d.bw[c] * d.bh[c];
            d.coef_off[c] = b->coef_elems;
            d.plane_off[c] = b->plane_bytes;
            b->jobs.push_back(IdctJob{i, c, b->coef_elems / 64, blocks});
            b->job_start.push_back(b->job_start.back() + blocks);
            b->coef_elems += blocks * 64;
            b->plane_bytes += blocks * 64;
        }
    }
    return CV_OK;
}

// Staging memory of cv_jpeg_decode_batch, per device, grow-only (cudaMallocHost / cudaMalloc cost milliseconds: never per call).
struct Scratch {
    // two staging slots: the host fills slot k & 1 for sub-batch k while the device still works on sub-batch k - 1 out of the other
    uint8_t *h_stage2[2] = {nullptr, nullptr}, *d_stage2[2] = {nullptr, nullptr};
    size_t stage_cap2[2] = {0, 0};
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};     // the copy out of h_stage2[i] has completed (the host may overwrite it)
    bool h2d_pending[2] = {false, false};
    uint8_t *h_stage = nullptr, *d_stage = nullptr;   // the slot in use (set by use_slot)
    uint8_t *d_planes = nullptr, *d_chunks = nullptr;
    int16_t* d_coef = nullptr;
    size_t stage_cap = 0, coef_cap = 0, plane_cap = 0, chunk_cap = 0;
    int slot = 0;
    int use_slot(int i) {                             // waits until the host may write the slot's pinned block again
        if (h_stage) { h_stage2[slot] = h_stage; d_stage2[slot] = d_stage; stage_cap2[slot] = stage_cap; }
        slot = i;
        if (!h2d_done[i]) CV_CUDA(cudaEventCreateWithFlags(&h2d_done[i], cudaEventDisableTiming));
        if (h2d_pending[i]) { CV_CUDA(cudaEventSynchronize(h2d_done[i])); h2d_pending[i] = false; }
        h_stage = h_stage2[i]; d_stage = d_stage2[i]; stage_cap = stage_cap2[i];
        return CV_OK;
    }
    int copied(cudaStream_t s) {                      // call right after enqueuing the H2D copy of the slot in use
        CV_CUDA(cudaEventRecord(h2d_done[slot], s));
        h2d_pending[slot] = true;
        return CV_OK;
    }
    int reserve_chunks(size_t bytes) {
        if (bytes > chunk_cap) {
            CV_CUDA(cudaDeviceSynchronize());
            if (d_chunks) cudaFree(d_chunks);
            d_chunks = nullptr; chunk_cap = 0;
            CV_CUDA(cudaMalloc(&d_chunks, bytes + bytes / 4));
            chunk_cap = bytes + bytes / 4;
        }
        return CV_OK;
    }
    int reserve(size_t stage, size_t coef, size_t planes) {
        if (stage > stage_cap || coef > coef_cap || planes > plane_cap) CV_CUDA(cudaDeviceSynchronize());     // nothing in flight uses what is freed below
        if (stage > stage_cap) {
            if (h_stage) cudaFreeHost(h_stage);
            if (d_stage) cudaFree(d_stage);
            h_stage = d_stage = nullptr; stage_cap = 0;
            const size_t cap = stage + stage / 4;
            CV_CUDA(cudaMallocHost(&h_stage, cap));
            CV_CUDA(cudaMalloc(&d_stage, cap));
            stage_cap = cap;
        }
        if (coef > coef_cap) {
            if (d_coef) cudaFree(d_coef);
            d_coef = nullptr; coef_cap = 0;
            CV_CUDA(cudaMalloc(&d_coef, coef + coef / 4));
            coef_cap = coef + coef / 4;
        }
        if (planes > plane_cap) {
            if (d_planes) cudaFree(d_planes);
            d_planes = nullptr; plane_cap = 0;
            CV_CUDA(cudaMalloc(&d_planes, planes + planes / 4));
            plane_cap = planes + planes / 4;
        }
        return CV_OK;
    }
};
std::mutex g_scratch_mutex;
std::map<int, Scratch> g_scratch;

}  // namespace

extern "C" {

int cv_jpeg_info(const uint8_t* file_host, size_t size, int* width, int* height, int* components) {
    CV_ARG(file_host != nullptr, "null file");
    JpegDesc d;
    TableSet ts;
    const char* why = parse_jpeg(file_host, size, &d, &ts);
    if (why) { cv_set_error("cv_jpeg_info: %s", why); return CV_ERR_ARG; }
    if (width) *width = d.width;
    if (height) *height = d.height;
    if (components) *components = d.ncomp;
    return CV_OK;
}

// Host mirror of the entropy decoder (the same inline routine the device kernel runs): quantised coefficients of one file, natural order,
// component after component, each as [block rows][block columns][64] over whole MCUs.  For the no-GPU tests.
int cv_jpeg_decode_coefficients_host(const uint8_t* file_host, size_t size, int16_t* coef_host, size_t capacity, int32_t* block_grid /*[3][2] = (bh, bw)*/) {
    CV_ARG(file_host != nullptr, "null file");
    Batch b;
    JpegDesc probe;
    TableSet probe_ts;
    const char* why = parse_jpeg(file_host, size, &probe, &probe_ts);
    if (why) { cv_set_error("cv_jpeg_decode_coefficients_host: %s", why); return CV_ERR_ARG; }
    const uint8_t* files[1] = {file_host};
    int rc = plan_batch(files, &size, 1, probe.width, probe.height, &b);
    if (rc) return rc;
    if (block_grid)
        for (int c = 0; c < 3; ++c) { block_grid[2 * c] = c < probe.ncomp ? b.descs[0].bh[c] : 0; block_grid[2 * c + 1] = c < probe.ncomp ? b.descs[0].bw[c] : 0; }
    if (!coef_host) return CV_OK;
    CV_ARG(capacity >= (size_t)b.coef_elems, "coefficient buffer too small");
    std::vector<int16_t> aligned((size_t)b.coef_elems + 8);
    int16_t* dst = reinterpret_cast<int16_t*>((reinterpret_cast<uintptr_t>(aligned.data()) + 15) & ~(uintptr_t)15);
    for (const Interval& v : b.intervals) huff_decode_interval(b.descs[0], b.tsets[0], file_host, v.byte0, v.byte1, v.mcu0, v.n_mcu, dst, zigzag_tab());
    memcpy(coef_host, dst, (size_t)b.coef_elems * sizeof(int16_t));
    return CV_OK;
}

// The chunked decoder (huff_decode_chunk: what the device runs by default) executed on the host, round by round as the kernels do it
// (every round reads the exit states of the round before).  Output = cv_jpeg_decode_coefficients_host's.  stats (nullable, int32[4]):
// chunks, chunk decodes over all rounds, intervals whose chain did not close (decoded serially instead), last round that changed a state.
int cv_jpeg_decode_coefficients_host_chunked(const uint8_t* file_host, size_t size, int16_t* coef_host, size_t capacity, int chunk_bytes, int rounds,
                                             int32_t* stats) {
    CV_ARG(file_host != nullptr && coef_host != nullptr, "null argument");
    CV_ARG((chunk_bytes >= 16 || chunk_bytes == -1) && rounds >= 1, "chunk_bytes >= 16 (or -1: the device path's own rule) and rounds >= 1");
    Batch b;
    JpegDesc probe;
    TableSet probe_ts;
    const char* why = parse_jpeg(file_host, size, &probe, &probe_ts);
    if (why) { cv_set_error("cv_jpeg_decode_coefficients_host_chunked: %s", why); return CV_ERR_ARG; }
    const uint8_t* files[1] = {file_host};
    int rc = plan_batch(files, &size, 1, probe.width, probe.height, &b);
    if (rc) return rc;
    CV_ARG(capacity >= (size_t)b.coef_elems, "coefficient buffer too small");
    if (chunk_bytes == -1) plan_chunks(&b, kChunkBytes, kChunkMcus);
    else plan_chunks(&b, chunk_bytes, 0);
    const JpegDesc& d = b.descs[0];
    const TableSet& ts = b.tsets[0];
    const size_t nc = b.chunk_iv.size();
    std::vector<uint64_t> used(nc, ~0ull), ex(nc, 0), ex_prev;
    std::vector<int32_t> count(nc, 0), blk0(nc, 0);
    std::vector<int16_t> aligned((size_t)b.coef_elems + 8, 0);
    int16_t* dst = reinterpret_cast<int16_t*>((reinterpret_cast<uintptr_t>(aligned.data()) + 15) & ~(uintptr_t)15);
    int decodes = 0, last_change = -1, failed = 0;
    for (int r = 0; r < rounds; ++r) {
        ex_prev = ex;
        for (size_t g = 0; g < nc; ++g) {
            const int vi = b.chunk_iv[g];
            const IvChunks ic = b.ivc[vi];
            const Interval& v = b.intervals[vi];
            const int i = (int)g - ic.chunk0;
            if (i == ic.n_chunks - 1) continue;
            const uint64_t entry = i == 0 ? pack_state(0, 0, 0) : r == 0 ? chunk_guess(file_host + v.byte0, (int64_t)i * ic.chunk_bytes) : ex_prev[g - 1];
            if (r > 0 && entry == used[g]) continue;
            int done = 0;
            ex[g] = huff_decode_chunk<false>(d, ts, file_host, v.byte0, v.byte1, entry, (int64_t)(i + 1) * ic.chunk_bytes * 8, v.mcu0, 0, 0, nullptr,
                                             zigzag_tab(), &done);
            used[g] = entry; count[g] = done;
            ++decodes; last_change = r;
        }
    }
    for (size_t vi = 0; vi < b.intervals.size(); ++vi) {
        const IvChunks ic = b.ivc[vi];
        const Interval& v = b.intervals[vi];
        bool ok = true;
        int run = 0;
        for (int i = 0; i < ic.n_chunks; ++i) {
            const int g = ic.chunk0 + i;
            blk0[g] = run;
            if (i < ic.n_chunks - 1) {
                if (i > 0 && used[g] != ex[g - 1]) ok = false;
                run += count[g];
            }
        }
        if (!ok) {
            ++failed;
            huff_decode_interval(d, ts, file_host, v.byte0, v.byte1, v.mcu0, v.n_mcu, dst, zigzag_tab());
            continue;
        }
        const int64_t blk_end = (int64_t)v.n_mcu * (d.hs * d.vs + d.ncomp - 1);
        for (int i = 0; i < ic.n_chunks; ++i) {
            const int g = ic.chunk0 + i;
            int done = 0;
            huff_decode_chunk<true>(d, ts, file_host, v.byte0, v.byte1, i == 0 ? pack_state(0, 0, 0) : ex[g - 1],
                                    i == ic.n_chunks - 1 ? ((int64_t)1 << 46) : (int64_t)(i + 1) * ic.chunk_bytes * 8, v.mcu0, blk0[g], blk_end, dst,
                                    zigzag_tab(), &done);
        }
        for (int c = 0; c < d.ncomp; ++c) {
            const int64_t total = (int64_t)v.n_mcu * (c == 0 ? d.hs * d.vs : 1);
            int run_dc = 0;
            for (int64_t j = 0; j < total; ++j) {
                int16_t* blk = dc_block(d, dst, c, v.mcu0, j);
                run_dc += blk[0];
                blk[0] = (int16_t)run_dc;
            }
        }
    }
    memcpy(coef_host, dst, (size_t)b.coef_elems * sizeof(int16_t));
    if (stats) { stats[0] = (int32_t)nc; stats[1] = decodes; stats[2] = failed; stats[3] = last_change; }
    return CV_OK;
}

}  // extern "C"

namespace {
int decode_sub(Scratch& sc, int slot, const uint8_t* const* files_host, const size_t* sizes, int n, int W, int H, uint8_t* rgb, int entropy_on_host,
               cudaStream_t s) {
#ifdef CV_EXPERIMENTS                 // CV_JPEG_TRACE=1: wall-clock split of one call (stream synchronised after every stage)
    const bool trace = getenv("CV_JPEG_TRACE") != nullptr && atoi(getenv("CV_JPEG_TRACE")) == 1;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "  jpeg %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
#else
    auto mark = [](const char*) {};
#endif
    Batch b;
    int rc = plan_batch(files_host, sizes, n, W, H, &b);
    if (rc) return rc;
    int chunk_bytes = kChunkBytes, chunk_mcus = kChunkMcus, spec_rounds = kSpecRounds;
#ifdef CV_EXPERIMENTS
    if (const char* e = getenv("CV_JPEG_CHUNK")) chunk_bytes = atoi(e);
    if (const char* e = getenv("CV_JPEG_MCUS")) chunk_mcus = atoi(e);
    if (const char* e = getenv("CV_JPEG_ROUNDS")) spec_rounds = std::max(1, atoi(e));
#endif
    const bool chunked = !entropy_on_host && chunk_bytes >= 16;
    if (chunked) plan_chunks(&b, chunk_bytes, chunk_mcus);
    mark("parse headers + plan");
    rc = sc.use_slot(slot);
    if (rc) return rc;
    const size_t desc_b = b.descs.size() * sizeof(JpegDesc), ts_b = b.tsets.size() * sizeof(TableSet), iv_b = b.intervals.size() * sizeof(Interval),
                 job_b = b.jobs.size() * sizeof(IdctJob), js_b = b.job_start.size() * sizeof(int64_t);
    // one pinned staging block: [descs | table sets | intervals | jobs | job starts | payload], mirrored by one device block
    size_t off = 0;
    auto place = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_desc = place(desc_b), o_ts = place(ts_b), o_iv = place(iv_b), o_job = place(job_b), o_js = place(js_b);
    const size_t ivc_b = b.ivc.size() * sizeof(IvChunks), civ_b = b.chunk_iv.size() * sizeof(int32_t);
    const size_t o_ivc = place(ivc_b), o_civ = place(civ_b);
    const size_t payload = entropy_on_host ? (size_t)b.coef_elems * sizeof(int16_t) : (size_t)b.bytes + 16;
    const size_t o_pay = place(payload);
    rc = sc.reserve(off, (size_t)b.coef_elems * sizeof(int16_t), (size_t)b.plane_bytes);
    if (rc) return rc;
    memcpy(sc.h_stage + o_desc, b.descs.data(), desc_b);
    memcpy(sc.h_stage + o_ts, b.tsets.data(), ts_b);
    memcpy(sc.h_stage + o_iv, b.intervals.data(), iv_b);
    memcpy(sc.h_stage + o_job, b.jobs.data(), job_b);
    memcpy(sc.h_stage + o_js, b.job_start.data(), js_b);
    if (ivc_b) memcpy(sc.h_stage + o_ivc, b.ivc.data(), ivc_b);
    if (civ_b) memcpy(sc.h_stage + o_civ, b.chunk_iv.data(), civ_b);
    const JpegDesc* d_desc = reinterpret_cast<const JpegDesc*>(sc.d_stage + o_desc);
    const TableSet* d_ts = reinterpret_cast<const TableSet*>(sc.d_stage + o_ts);
    const Interval* d_iv = reinterpret_cast<const Interval*>(sc.d_stage + o_iv);
    const IdctJob* d_jobs = reinterpret_cast<const IdctJob*>(sc.d_stage + o_job);
    const int64_t* d_jstart = reinterpret_cast<const int64_t*>(sc.d_stage + o_js);
    const int16_t* d_coef = sc.d_coef;
    int16_t* d_dcbuf = nullptr;                              // chunked device decoding: DC values live in a compact array
    if (entropy_on_host) {
        int16_t* h_coef = reinterpret_cast<int16_t*>(sc.h_stage + o_pay);
        for (const Interval& v : b.intervals) {
            const int64_t base = b.file_base[v.image];        // interval byte ranges are batch-relative: rebase them onto this file
            huff_decode_interval(b.descs[v.image], b.tsets[b.descs[v.image].tset], files_host[v.image], v.byte0 - base, v.byte1 - base, v.mcu0, v.n_mcu,
                                 h_coef, zigzag_tab());
        }
        mark("entropy decode (host)");
        CV_CUDA(cudaMemcpyAsync(sc.d_stage, sc.h_stage, off, cudaMemcpyHostToDevice, s));       // tables + coefficients in one copy
        rc = sc.copied(s);
        if (rc) return rc;
        d_coef = reinterpret_cast<const int16_t*>(sc.d_stage + o_pay);
        mark("H2D tables + coefficients");
    } else {
        parallel_ranges(n, 256, [&](int i0, int i1) {
            for (int i = i0; i < i1; ++i) memcpy(sc.h_stage + o_pay + b.file_base[i], files_host[i], sizes[i]);
        });
        CV_CUDA(cudaMemcpyAsync(sc.d_stage, sc.h_stage, off, cudaMemcpyHostToDevice, s));       // tables + compressed bytes in one copy
        rc = sc.copied(s);
        if (rc) return rc;
        mark("stage + H2D compressed bytes");
        const int n_iv = (int)b.intervals.size(), n_sets = (int)b.tsets.size();
        const int smem_sets = n_sets <= 6 ? n_sets : 0;          // up to 44 KB of shared memory for the table sets
        const size_t smem = (size_t)smem_sets * sizeof(TableSet);
        const uint8_t* d_bytes = sc.d_stage + o_pay;
        if (smem > 48 * 1024 - 1024) {
            CV_CUDA(cudaFuncSetAttribute(jpeg_huffman_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CV_CUDA(cudaFuncSetAttribute(jpeg_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CV_CUDA(cudaFuncSetAttribute(jpeg_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        const int* only_failed = nullptr;
        if (chunked) {
            // speculative rounds -> chain check + block offsets -> writing pass -> DC prefix sums; the one-thread kernel behind them
            // takes the intervals whose chain did not close (none, normally: it exits at once)
            const int n_ch = (int)b.chunk_iv.size();
            const size_t per_chunk = 2 * sizeof(uint64_t) + 2 * sizeof(int32_t);
            const size_t n_blocks = (size_t)(b.coef_elems / 64);
            rc = sc.reserve_chunks((size_t)n_ch * per_chunk + (size_t)n_iv * sizeof(int32_t) + n_blocks * sizeof(int16_t) + 64);
            if (rc) return rc;
            ChunkArrays a;
            a.used_entry = reinterpret_cast<uint64_t*>(sc.d_chunks);
            a.exit_state = a.used_entry + n_ch;
            a.count = reinterpret_cast<int32_t*>(a.exit_state + n_ch);
            a.blk0 = a.count + n_ch;
            a.failed = a.blk0 + n_ch;
            a.dcbuf = reinterpret_cast<int16_t*>(a.failed + n_iv);
            d_dcbuf = a.dcbuf;
            CV_CUDA(cudaMemsetAsync(a.dcbuf, 0, n_blocks * sizeof(int16_t), s));
            const IvChunks* d_ivc = reinterpret_cast<const IvChunks*>(sc.d_stage + o_ivc);
            const int32_t* d_civ = reinterpret_cast<const int32_t*>(sc.d_stage + o_civ);
            CV_CUDA(cudaMemsetAsync(sc.d_coef, 0, (size_t)b.coef_elems * sizeof(int16_t), s));
            mark("  chunked: zero coefficients");
            const int grid = (n_ch + kChunkThreads - 1) / kChunkThreads;
            for (int r = 0; r < spec_rounds; ++r) {
                jpeg_spec_kernel<<<grid, kChunkThreads, smem, s>>>(d_desc, d_ts, n_sets, smem_sets, d_iv, d_ivc, d_civ, n_ch, d_bytes, r, a);
                CV_CHECK_LAUNCH();
                mark("  chunked: speculative round");
            }
            jpeg_chunk_scan_kernel<<<(n_iv + 127) / 128, 128, 0, s>>>(d_ivc, n_iv, a);
            CV_CHECK_LAUNCH();
            mark("  chunked: chain check + offsets");
            jpeg_write_kernel<<<grid, kChunkThreads, smem, s>>>(d_desc, d_ts, n_sets, smem_sets, d_iv, d_ivc, d_civ, n_ch, d_bytes, a, sc.d_coef);
            CV_CHECK_LAUNCH();
            mark("  chunked: writing pass");
            jpeg_dc_scan_kernel<<<(n_iv * 3 * 32 + 127) / 128, 128, 0, s>>>(d_desc, d_iv, n_iv, a.failed, a.dcbuf);
            CV_CHECK_LAUNCH();
            mark("  chunked: DC prefix sums");
            only_failed = a.failed;
        }
        jpeg_huffman_kernel<<<(n_iv + kHuffThreads - 1) / kHuffThreads, kHuffThreads, smem, s>>>(d_desc, d_ts, n_sets, smem_sets, d_iv, n_iv, d_bytes,
                                                                                              sc.d_coef, only_failed, d_dcbuf);
        CV_CHECK_LAUNCH();
#ifdef CV_EXPERIMENTS
        if (trace && chunked) {
            std::vector<int32_t> f(n_iv);
            cudaStreamSynchronize(s);
            cudaMemcpy(f.data(), only_failed, n_iv * sizeof(int32_t), cudaMemcpyDeviceToHost);
            int nf = 0;
            for (int v : f) nf += v != 0;
            fprintf(stderr, "  jpeg chunked: %zu chunks (>= %d bytes, %d MCUs), %d rounds, %d of %d intervals fell back\n", b.chunk_iv.size(), chunk_bytes, chunk_mcus, spec_rounds, nf, n_iv);
        }
#endif
        mark("entropy decode (device)");
    }
    const int64_t total_blocks = b.job_start.back();
    jpeg_idct_kernel<<<(unsigned)((total_blocks + 31) / 32), 256, 0, s>>>(d_desc, d_jstart, d_jobs, (int)b.jobs.size(), d_coef, d_dcbuf, sc.d_planes);
    CV_CHECK_LAUNCH();
    mark("idct");
    if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(rgb) & 3) == 0)
        jpeg_color4_kernel<<<dim3((unsigned)((W * H / 4 + 255) / 256), (unsigned)n), 256, 0, s>>>(d_desc, sc.d_planes, W, H, rgb);
    else
        jpeg_color_kernel<<<dim3((unsigned)((W * H + 255) / 256), (unsigned)n), 256, 0, s>>>(d_desc, sc.d_planes, W, H, rgb);
    CV_CHECK_LAUNCH();
    mark("upsampling + colour");
    return CV_OK;
}
}  // namespace

extern "C" {
// n baseline JPEG files of ONE size (host pointers) -> rgb (DEVICE, uint8 (n, H, W, 3)), bit-exact with PIL.Image.open(f).convert("RGB").
// entropy_on_host != 0 decodes the Huffman streams on the host and ships coefficients (3 bytes per pixel at 4:2:0) instead of the
// compressed bytes; results are identical.  Synchronises `stream` before returning.  Staging memory (pinned host + device) is kept
// per device for the life of the process and only grows: a steady stream of equal batches allocates once.
int cv_jpeg_decode_batch(const uint8_t* const* files_host, const size_t* sizes, int n, int W, int H, uint8_t* rgb, int entropy_on_host, void* stream) {
    CV_ARG(n >= 0, "negative batch");
    if (n == 0) return CV_OK;
    CV_ARG(files_host && sizes && rgb, "null argument");
    CV_ARG(W > 0 && H > 0 && W <= 16384 && H <= 16384, "bad image size");
    CV_ARG(n <= 65535, "at most 65535 files per call");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    CV_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    Scratch& sc = g_scratch[dev];
    // Sub-batches through two staging slots: the host parses and stages sub-batch k + 1 while the device decodes sub-batch k (the
    // coefficient / plane / chunk-state buffers are shared: the kernels of all sub-batches run in stream order).
    int sub = n;
    if (n > 4096) sub = 4096;                                // bounds the scratch memory (0.3 MB per 256x256 file) and overlaps host and device work
#ifdef CV_EXPERIMENTS
    if (const char* e = getenv("CV_JPEG_SUB")) sub = std::max(1, atoi(e));
#endif
    int rc = CV_OK;
#ifdef CV_EXPERIMENTS                 // CV_JPEG_TRACE=2: host-side timeline of the sub-batches (no extra synchronisation)
    const char* tr = getenv("CV_JPEG_TRACE");
    const bool timeline = tr && atoi(tr) == 2;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
#endif
    for (int k = 0, i0 = 0; i0 < n && rc == CV_OK; ++k, i0 += sub) {
        rc = decode_sub(sc, k & 1, files_host + i0, sizes + i0, std::min(sub, n - i0), W, H, rgb + (size_t)i0 * W * H * 3, entropy_on_host, s);
#ifdef CV_EXPERIMENTS
        if (timeline) fprintf(stderr, "  jpeg sub-batch %d enqueued at %.3f ms\n", k, since());
#endif
    }
    const cudaError_t e = cudaStreamSynchronize(s);        // the pixels are in `rgb`; the staging slots are free for the next call
#ifdef CV_EXPERIMENTS
    if (timeline) fprintf(stderr, "  jpeg all done at %.3f ms\n", since());
#endif
    sc.h2d_pending[0] = sc.h2d_pending[1] = false;
    if (rc) return rc;
    CV_CUDA(e);
    return CV_OK;
}

}  // extern "C"
