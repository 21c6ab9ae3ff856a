// Precision-templated CUDA-core kernels of the ChessSquareCNN hot path (sm_100a).
//
// These are the fp32 "exact" mode of the path (1e-5 parity against the oracle) and the bring-up /
// cross-check implementation for the bf16 tensor-core kernels (umma_*.cu).  Activations are NHWC;
// accumulation is always fp32; BatchNorm is pre-folded into (w, bias) by the packer.
//
//   crop_*        ChessSquareCNN._crop_squares           models/square.py:43-74
//   conv_generic  dense 3x3 s2 / pointwise 1x1 convs     timm trunk, models/square.py:86
//   depthwise     depthwise k3/k5 s1/s2 convs            timm trunk, models/square.py:86
//   pool_heads    global_pool + type/color heads + combine  square.py:87-104, common.py:24
//   global_head   global_head -> turn/castling heads     square.py:107-113
#include "internal.h"

namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// One thread per (crop, oy, ox); 3 channels each.
template <typename T, bool U8, bool CHW>
__global__ void __launch_bounds__(256)
crop_kernel(const void* __restrict__ src, int64_t total, int H, const __grid_constant__ CropTaps tp,
            const float* __restrict__ lut, T* __restrict__ out_nhwc, float* __restrict__ out_nchw) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int ox = idx & 63, oy = (idx >> 6) & 63;
    int64_t n = idx >> 12;
    int col = n & 7, row = (n >> 3) & 7;
    int64_t b = n >> 6;
    int y0 = tp.p0[row][oy], y1 = tp.p1[row][oy], x0 = tp.p0[col][ox], x1 = tp.p1[col][ox];
    float ly = tp.lam[oy], lx = tp.lam[ox];
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v00, v01, v10, v11;
        if (U8) {
            const uint8_t* p = (const uint8_t*)src + b * (int64_t)H * H * 3;
            const float* l = lut + c * 256;
            if (CHW) {
                const uint8_t* pc = p + (int64_t)c * H * H;
                v00 = l[pc[y0 * H + x0]]; v01 = l[pc[y0 * H + x1]];
                v10 = l[pc[y1 * H + x0]]; v11 = l[pc[y1 * H + x1]];
            } else {
                v00 = l[p[(y0 * H + x0) * 3 + c]]; v01 = l[p[(y0 * H + x1) * 3 + c]];
                v10 = l[p[(y1 * H + x0) * 3 + c]]; v11 = l[p[(y1 * H + x1) * 3 + c]];
            }
        } else {
            const float* pc = (const float*)src + (b * 3 + c) * (int64_t)H * H;
            v00 = pc[y0 * H + x0]; v01 = pc[y0 * H + x1];
            v10 = pc[y1 * H + x0]; v11 = pc[y1 * H + x1];
        }
        r[c] = crop_blend(v00, v01, v10, v11, lx, ly);
    }
    if (out_nhwc) {
        T* o = out_nhwc + idx * 3;
        stf<T>(o, r[0]); stf<T>(o + 1, r[1]); stf<T>(o + 2, r[2]);
    }
    if (out_nchw) {
        float* o = out_nchw + n * 3 * 4096 + oy * 64 + ox;
        o[0] = r[0]; o[4096] = r[1]; o[8192] = r[2];
    }
}

// Dense KxK / 1x1 convolution.  Thread = (output pixel, group of CO_T output channels); lanes of a warp
// share the pixel (input loads broadcast) and read contiguous weight vectors.
template <typename T, int CO_T, typename LIn, typename LOut>
__global__ void __launch_bounds__(256)
conv_generic_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                    const T* __restrict__ skip, T* __restrict__ out, int64_t total, int Hin, int Hout, int Cin,
                    int Cout, int K, int S, int pad, int relu) {
    const int groups = Cout / CO_T;
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int g = (int)(idx % groups);
    int64_t p = idx / groups;
    int ox = (int)(p % Hout), oy = (int)((p / Hout) % Hout);
    int64_t n = p / ((int64_t)Hout * Hout);
    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < K; ++ky) {
        int iy = oy * S - pad + ky;
        if (iy < 0 || iy >= Hin) continue;
        for (int kx = 0; kx < K; ++kx) {
            int ix = ox * S - pad + kx;
            if (ix < 0 || ix >= Hin) continue;
            const int64_t pin = (n * Hin + iy) * Hin + ix;
            const float* wp = w + (int64_t)((ky * K + kx) * Cin) * Cout + g * CO_T;
            for (int ci = 0; ci < Cin; ++ci) {
                float v = ldf<T>(in + LIn::off(pin, ci, Cin));
                const float4* w4 = reinterpret_cast<const float4*>(wp + (int64_t)ci * Cout);
#pragma unroll
                for (int j = 0; j < CO_T / 4; ++j) {
                    float4 q = __ldg(w4 + j);
                    acc[4 * j + 0] = fmaf(v, q.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(v, q.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(v, q.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(v, q.w, acc[4 * j + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
        const int64_t o = LOut::off(p, g * CO_T + j, Cout);
        float v = acc[j] + __ldg(bias + g * CO_T + j);
        if (relu) v = fmaxf(v, 0.f);
        if (skip) v += ldf<T>(skip + o);
        stf<T>(out + o, v);
    }
}

// Depthwise KxK.  Thread = (output pixel, channel), channel fastest -> coalesced.
template <typename T, typename L>
__global__ void __launch_bounds__(256)
depthwise_generic_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                         T* __restrict__ out, int64_t total, int Hin, int Hout, int C, int K, int S, int pad,
                         int relu) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c = (int)(idx % C);
    int64_t p = idx / C;
    int ox = (int)(p % Hout), oy = (int)((p / Hout) % Hout);
    int64_t n = p / ((int64_t)Hout * Hout);
    float acc = 0.f;
    for (int ky = 0; ky < K; ++ky) {
        int iy = oy * S - pad + ky;
        if (iy < 0 || iy >= Hin) continue;
        for (int kx = 0; kx < K; ++kx) {
            int ix = ox * S - pad + kx;
            if (ix < 0 || ix >= Hin) continue;
            acc = fmaf(ldf<T>(in + L::off((n * Hin + iy) * Hin + ix, c, C)), __ldg(w + (ky * K + kx) * C + c), acc);
        }
    }
    acc += __ldg(bias + c);
    if (relu) acc = fmaxf(acc, 0.f);
    stf<T>(out + L::off(p, c, C), acc);
}

__constant__ int kClassToType[13] = {0, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6};    // dataset.py:31
__constant__ int kClassToColor[13] = {0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2};   // dataset.py:32

// One warp per crop: mean over the 2x2 map, the 7+3 head dot products, type+color -> 13 joint logits.
template <typename T, typename L>
__global__ void __launch_bounds__(256)
pool_heads_kernel(const T* __restrict__ fmap, const float* __restrict__ head_w, const float* __restrict__ head_b,
                  int64_t n_crops, float* __restrict__ features, float* __restrict__ squares) {
    int64_t n = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (n >= n_crops) return;
    float part[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) part[r] = 0.f;
    for (int c = lane; c < 480; c += 32) {
        float m = ((ldf<T>(fmap + L::off(n * 4, c, 480)) + ldf<T>(fmap + L::off(n * 4 + 1, c, 480))) +
                   (ldf<T>(fmap + L::off(n * 4 + 2, c, 480)) + ldf<T>(fmap + L::off(n * 4 + 3, c, 480)))) * 0.25f;
        features[n * 480 + c] = m;
#pragma unroll
        for (int r = 0; r < 10; ++r) part[r] = fmaf(m, __ldg(head_w + r * 480 + c), part[r]);
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
        part[r] += __ldg(head_b + r);
    }
    if (lane < 13) {
        float t = 0.f, cl = 0.f;
        int ti = kClassToType[lane], ci = kClassToColor[lane];
#pragma unroll
        for (int r = 0; r < 7; ++r) if (r == ti) t = part[r];
#pragma unroll
        for (int r = 0; r < 3; ++r) if (r == ci) cl = part[7 + r];
        squares[n * 13 + lane] = t + cl;
    }
}

// global_head Linear(30720,64)+ReLU -> turn(1)/castling(4).  Block = GB boards; 256 threads = 64 hidden
// units x 4 K-slices; the board tile's features are staged through shared memory in K chunks.  AccT = double
// in the fp32 "exact" mode: the 30720-term dot products cancel heavily (bias = -W.mean), so fp32 summation order
// alone moves the result by ~1e-5; fp64 accumulation makes this side exact to fp32 rounding.
constexpr int GB = 8, GK = 512;
template <typename AccT>
__global__ void __launch_bounds__(256)
global_head_kernel(const float* __restrict__ feat, const float* __restrict__ wt, const float* __restrict__ gb,
                   const float* __restrict__ tc_w, const float* __restrict__ tc_b, int B, float* __restrict__ turn,
                   float* __restrict__ castling) {
    __shared__ float sf[GB][GK];
    __shared__ AccT red[4][GB][64];
    __shared__ float hid[GB][64];
    const int j = threadIdx.x & 63, slice = threadIdx.x >> 6;
    const int b0 = blockIdx.x * GB;
    AccT acc[GB];
#pragma unroll
    for (int i = 0; i < GB; ++i) acc[i] = 0;
    for (int k0 = 0; k0 < 30720; k0 += GK) {
        for (int t = threadIdx.x; t < GB * GK / 4; t += 256) {
            int bi = t / (GK / 4), kk = t % (GK / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + bi < B) v = __ldg(reinterpret_cast<const float4*>(feat + (int64_t)(b0 + bi) * 30720 + k0) + kk);
            reinterpret_cast<float4*>(&sf[bi][0])[kk] = v;
        }
        __syncthreads();
        float part[GB];                      // two-level summation
#pragma unroll
        for (int i = 0; i < GB; ++i) part[i] = 0.f;
        for (int kk = slice; kk < GK; kk += 4) {
            float wv = __ldg(wt + (int64_t)(k0 + kk) * 64 + j);
            if (sizeof(AccT) == 8) {
#pragma unroll
                for (int i = 0; i < GB; ++i) acc[i] += (AccT)sf[i][kk] * (AccT)wv;
            } else {
#pragma unroll
                for (int i = 0; i < GB; ++i) part[i] = fmaf(sf[i][kk], wv, part[i]);
            }
        }
        if (sizeof(AccT) != 8) {
#pragma unroll
            for (int i = 0; i < GB; ++i) acc[i] += part[i];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < GB; ++i) red[slice][i][j] = acc[i];
    __syncthreads();
    for (int t = threadIdx.x; t < GB * 64; t += 256) {
        int bi = t >> 6, jj = t & 63;
        AccT v = ((red[0][bi][jj] + red[1][bi][jj]) + (red[2][bi][jj] + red[3][bi][jj])) + (AccT)gb[jj];
        hid[bi][jj] = fmaxf((float)v, 0.f);
    }
    __syncthreads();
    if (threadIdx.x < GB * 5) {
        int bi = threadIdx.x / 5, r = threadIdx.x % 5;
        if (b0 + bi < B) {
            AccT v = 0;
            for (int jj = 0; jj < 64; ++jj) v += (AccT)hid[bi][jj] * (AccT)tc_w[r * 64 + jj];
            v += (AccT)tc_b[r];
            if (r == 0) turn[b0 + bi] = (float)v; else castling[(int64_t)(b0 + bi) * 4 + (r - 1)] = (float)v;
        }
    }
}

// The fp64 head as a register-tiled GEMM: partial[ks][b][j] = sum over K range ks of feat[b][k] * wt[k][j], products and sums in double.
// CTA = 64 boards x 64 outputs x one of HEAD_KS K ranges (1024 boards: 16 x 30 CTAs, three per SM); thread = 4 boards x 4 outputs (j, j + 16,
// j + 32, j + 48: a warp's weight loads are 128 contiguous bytes).  Features and weights are converted to double ONCE when a 32-deep K slab
// is staged in shared memory (the feature slab transposed, rows padded to 65 doubles: conflict-free both ways), so the inner loop is
// 8 LDS.64 per 16 DFMA -- the first version converted both operands in front of every DFMA and ran at 2.4 TFLOP/s (0.8 ms per 1024 boards).
// The finishing kernel adds the partial sums in a fixed order (deterministic), then bias, ReLU, 64 -> 1 + 4.
constexpr int HEAD_KS = 30, HEAD_BT = 64, HEAD_KC = 32;
__global__ void __launch_bounds__(256)
global_head_f64_partial_kernel(const float* __restrict__ feat, const float* __restrict__ wt, int B, double* __restrict__ partial) {
    __shared__ double fs[HEAD_KC][HEAD_BT + 1];
    __shared__ double ws[HEAD_KC][64];
    const int tj = threadIdx.x & 15, tb = threadIdx.x >> 4;
    const int b0 = blockIdx.x * HEAD_BT, ks = blockIdx.y;
    const int k_lo = ks * (30720 / HEAD_KS), k_hi = k_lo + 30720 / HEAD_KS;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = k_lo; k0 < k_hi; k0 += HEAD_KC) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {                      // 64 boards x 8 float4 of features, 32 x 16 float4 of weights
            const int idx = threadIdx.x + 256 * t;
            const int bi = idx >> 3, kq = idx & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + bi < B) v = __ldg(reinterpret_cast<const float4*>(feat + (int64_t)(b0 + bi) * 30720 + k0) + kq);
            fs[kq * 4 + 0][bi] = (double)v.x; fs[kq * 4 + 1][bi] = (double)v.y; fs[kq * 4 + 2][bi] = (double)v.z; fs[kq * 4 + 3][bi] = (double)v.w;
            const int kk = idx >> 4, jq = idx & 15;
            const float4 u = __ldg(reinterpret_cast<const float4*>(wt + (int64_t)(k0 + kk) * 64) + jq);
            ws[kk][jq * 4 + 0] = (double)u.x; ws[kk][jq * 4 + 1] = (double)u.y; ws[kk][jq * 4 + 2] = (double)u.z; ws[kk][jq * 4 + 3] = (double)u.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < HEAD_KC; ++kk) {
            double f[4], w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { f[i] = fs[kk][tb * 4 + i]; w4[i] = ws[kk][tj + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(f[i], w4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = b0 + tb * 4 + i;
        if (b < B) {
#pragma unroll
            for (int j = 0; j < 4; ++j) partial[((size_t)ks * B + b) * 64 + tj + 16 * j] = acc[i][j];
        }
    }
}
__global__ void __launch_bounds__(64)
global_head_f64_finish_kernel(const double* __restrict__ partial, const float* __restrict__ gb, const float* __restrict__ tc_w, const float* __restrict__ tc_b,
                              int B, float* __restrict__ turn, float* __restrict__ castling) {
    __shared__ float hid[64];
    const int b = blockIdx.x, j = threadIdx.x;
    double v = 0;
    for (int ks = 0; ks < HEAD_KS; ++ks) v += partial[((size_t)ks * B + b) * 64 + j];
    hid[j] = fmaxf((float)(v + (double)gb[j]), 0.f);
    __syncthreads();
    if (j < 5) {
        double o = 0;
        for (int jj = 0; jj < 64; ++jj) o += (double)hid[jj] * (double)tc_w[j * 64 + jj];
        o += (double)tc_b[j];
        if (j == 0) turn[b] = (float)o; else castling[(int64_t)b * 4 + (j - 1)] = (float)o;
    }
}

template <typename T, typename L>
__global__ void to_f32_kernel(const T* __restrict__ s, float* __restrict__ d, size_t n, int C) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) d[i] = ldf<T>(s + L::off((int64_t)(i / C), (int)(i % C), C));
}

__global__ void transpose_kernel(const float* __restrict__ s, float* __restrict__ d, int rows, int cols) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = s[(int64_t)(r0 + i) * cols + c];
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < cols && r < rows) d[(int64_t)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

int cv_make_crop_geom(int H, CropGeom* g) {
    if (H < 32 || H % 32 != 0 || H > 4096) { cv_set_error("board side H=%d must be a multiple of 32 in [32,4096]", H); return CV_ERR_ARG; }
    g->H = H;
    g->sq = H / 8;
    g->crop = (int)(g->sq * 1.5);                    // int(sq_size * square_overlap), square.py:54
    g->pad = (g->crop - g->sq) / 2;                  // square.py:55
    const float scale = (float)g->crop / 64.0f;      // ATen area_pixel_compute_scale (align_corners=False)
    for (int d = 0; d < 64; ++d) {
        if (g->crop == 64) { g->i0[d] = g->i1[d] = (int16_t)d; g->lam[d] = 0.f; continue; }
        volatile float t = scale * ((float)d + 0.5f);    // volatile: no host-side FMA contraction
        float src = t - 0.5f;
        if (src < 0.f) src = 0.f;
        int a = (int)src;
        g->i0[d] = (int16_t)a;
        g->i1[d] = (int16_t)(a + 1 < g->crop ? a + 1 : g->crop - 1);
        g->lam[d] = src - (float)a;
    }
    return CV_OK;
}

template <typename T>
int launch_crop_f32(const float* x, int B, int H, const CropGeom& g, T* out_nhwc, float* out_nchw, cudaStream_t s) {
    int64_t total = (int64_t)B * 64 * 4096;
    if (total == 0) return CV_OK;
    CropTaps tp = make_taps(g);
    crop_kernel<T, false, false><<<blocks_for(total, 256), 256, 0, s>>>(x, total, H, tp, nullptr, out_nhwc, out_nchw);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

template <typename T>
int launch_crop_u8(const uint8_t* boards, int layout, int B, int H, const CropGeom& g, const float* lut, T* out_nhwc,
                   float* out_nchw, cudaStream_t s) {
    int64_t total = (int64_t)B * 64 * 4096;
    if (total == 0) return CV_OK;
    CropTaps tp = make_taps(g);
    if (layout == CV_LAYOUT_CHW)
        crop_kernel<T, true, true><<<blocks_for(total, 256), 256, 0, s>>>(boards, total, H, tp, lut, out_nhwc, out_nchw);
    else
        crop_kernel<T, true, false><<<blocks_for(total, 256), 256, 0, s>>>(boards, total, H, tp, lut, out_nhwc, out_nchw);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

template <typename T, typename LIn, typename LOut>
static int launch_conv_l(const cv_layer_info& L, const T* in, const float* w, const float* bias, const T* skip, T* out,
                         int64_t n_crops, cudaStream_t s) {
    int pad = ((L.stride - 1) + (L.k - 1)) / 2;
    int64_t px = n_crops * L.hout * L.hout;
    if (px == 0) return CV_OK;
    if (L.cout % 16 == 0 && L.cout >= 64) {
        int64_t total = px * (L.cout / 16);
        conv_generic_kernel<T, 16, LIn, LOut><<<blocks_for(total, 256), 256, 0, s>>>(
            in, w, bias, skip, out, total, L.hin, L.hout, L.cin, L.cout, L.k, L.stride, pad, L.relu);
    } else {
        int64_t total = px * (L.cout / 8);
        conv_generic_kernel<T, 8, LIn, LOut><<<blocks_for(total, 256), 256, 0, s>>>(
            in, w, bias, skip, out, total, L.hin, L.hout, L.cin, L.cout, L.k, L.stride, pad, L.relu);
    }
    CV_CHECK_LAUNCH();
    return CV_OK;
}

template <typename T>
int launch_conv_generic(const cv_layer_info& L, const T* in, const float* w, const float* bias, const T* skip, T* out,
                        int64_t n_crops, bool in_t8, bool out_t8, cudaStream_t s) {
    if (in_t8 && out_t8) return launch_conv_l<T, T8L, T8L>(L, in, w, bias, skip, out, n_crops, s);
    if (!in_t8 && out_t8) return launch_conv_l<T, RowMajorL, T8L>(L, in, w, bias, skip, out, n_crops, s);
    if (!in_t8 && !out_t8) return launch_conv_l<T, RowMajorL, RowMajorL>(L, in, w, bias, skip, out, n_crops, s);
    cv_set_error("conv_generic: T8 -> row-major is not instantiated");
    return CV_ERR_ARG;
}

template <typename T>
int launch_depthwise_generic(const cv_layer_info& L, const T* in, const float* w, const float* bias, T* out,
                             int64_t n_crops, bool t8, cudaStream_t s) {
    int pad = ((L.stride - 1) + (L.k - 1)) / 2;
    int64_t total = n_crops * L.hout * L.hout * L.cout;
    if (total == 0) return CV_OK;
    if (t8)
        depthwise_generic_kernel<T, T8L><<<blocks_for(total, 256), 256, 0, s>>>(in, w, bias, out, total, L.hin, L.hout,
                                                                               L.cout, L.k, L.stride, pad, L.relu);
    else
        depthwise_generic_kernel<T, RowMajorL><<<blocks_for(total, 256), 256, 0, s>>>(in, w, bias, out, total, L.hin, L.hout,
                                                                                     L.cout, L.k, L.stride, pad, L.relu);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

template <typename T>
int launch_pool_heads(const T* fmap, const float* head_w, const float* head_b, int64_t n_crops, float* features,
                      float* squares, bool t8, cudaStream_t s) {
    if (n_crops == 0) return CV_OK;
    if (t8)
        pool_heads_kernel<T, T8L><<<blocks_for(n_crops, 8), 256, 0, s>>>(fmap, head_w, head_b, n_crops, features, squares);
    else
        pool_heads_kernel<T, RowMajorL><<<blocks_for(n_crops, 8), 256, 0, s>>>(fmap, head_w, head_b, n_crops, features,
                                                                              squares);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_global_head(const float* features, const float* glob_wt, const float* glob_b, const float* tc_w,
                       const float* tc_b, int B, float* turn, float* castling, bool exact, cudaStream_t s) {
    if (B == 0) return CV_OK;
    if (exact)
        global_head_kernel<double><<<blocks_for(B, GB), 256, 0, s>>>(features, glob_wt, glob_b, tc_w, tc_b, B, turn, castling);
    else
        global_head_kernel<float><<<blocks_for(B, GB), 256, 0, s>>>(features, glob_wt, glob_b, tc_w, tc_b, B, turn, castling);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

size_t global_head_f64_partial_bytes(int B) { return (size_t)HEAD_KS * B * 64 * sizeof(double); }
int launch_global_head_f64_split(const float* features, const float* glob_wt, const float* glob_b, const float* tc_w, const float* tc_b, int B,
                                 double* partial, float* turn, float* castling, cudaStream_t s) {
    if (B == 0) return CV_OK;
    static_assert(30720 % HEAD_KS == 0 && (30720 / HEAD_KS) % HEAD_KC == 0, "K ranges must be whole slabs");
    global_head_f64_partial_kernel<<<dim3((unsigned)blocks_for(B, HEAD_BT), HEAD_KS), 256, 0, s>>>(features, glob_wt, B, partial);
    CV_CHECK_LAUNCH();
    global_head_f64_finish_kernel<<<B, 64, 0, s>>>(partial, glob_b, tc_w, tc_b, B, turn, castling);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

template <typename T>
int launch_to_f32(const T* src, float* dst, size_t n, int C, bool t8, cudaStream_t s) {
    if (n == 0) return CV_OK;
    if (t8) to_f32_kernel<T, T8L><<<blocks_for((int64_t)n, 256), 256, 0, s>>>(src, dst, n, C);
    else to_f32_kernel<T, RowMajorL><<<blocks_for((int64_t)n, 256), 256, 0, s>>>(src, dst, n, C);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

int launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(src, dst, rows, cols);
    CV_CHECK_LAUNCH();
    return CV_OK;
}

#define INSTANTIATE(T)                                                                                              \
    template int launch_crop_f32<T>(const float*, int, int, const CropGeom&, T*, float*, cudaStream_t);             \
    template int launch_crop_u8<T>(const uint8_t*, int, int, int, const CropGeom&, const float*, T*, float*,       \
                                   cudaStream_t);                                                                   \
    template int launch_conv_generic<T>(const cv_layer_info&, const T*, const float*, const float*, const T*, T*,  \
                                        int64_t, bool, bool, cudaStream_t);                                         \
    template int launch_depthwise_generic<T>(const cv_layer_info&, const T*, const float*, const float*, T*,       \
                                             int64_t, bool, cudaStream_t);                                          \
    template int launch_pool_heads<T>(const T*, const float*, const float*, int64_t, float*, float*, bool,         \
                                      cudaStream_t);                                                                \
    template int launch_to_f32<T>(const T*, float*, size_t, int, bool, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(bf16)
