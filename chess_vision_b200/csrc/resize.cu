// Board resize of the reference's input transform on the GPU (the step right before the hot path).
// Replaces `transforms.Resize((S, S))` on a PIL image (/root/reference/dataset.py:177-181, fed at predict.py:19-20), i.e. Pillow's
// `Image.resize((S, S), BILINEAR)` (src/libImaging/Resample.c, Pillow 12.2.0): a separable triangle filter whose support grows with
// the shrink factor (antialiasing), int32 fixed-point weights with 22 fractional bits, the horizontal result rounded to uint8
// before the vertical pass.  Integer work: the output is BIT-EXACT with Pillow (tests/test_resize_gpu.py, golden vectors made by
// Pillow itself).  ToTensor + Normalize stay fused into the crop gather of the hot path (cv_square_forward_u8).
//
// One kernel, HBM-bound: a CTA produces TY output rows of one image.  It (1) runs the horizontal pass over the input rows those
// output rows need -- source bytes read straight from global memory as aligned 32-bit words (neighbouring threads share their
// taps through L1), weights in registers -- into a shared-memory image of uint8 rows, (2) runs the vertical pass from shared
// memory, eight output bytes per thread, one 64-bit coalesced store each.  Algorithmic bytes per image: in_h*in_w*3 read + out_h*out_w*3 written.
#include <cmath>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "internal.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;      // Resample.c: PRECISION_BITS
constexpr int RS_THREADS = 256;

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter, in the same double arithmetic.
void make_coeffs(int in_size, int out_size, int* ksize_out, std::vector<int32_t>& bounds, std::vector<int32_t>& kk) {
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> w((size_t)ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < ksize; ++x) w[x] = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w[x] = a < 1.0 ? 1.0 - a : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) w[x] /= ww;
        for (int x = 0; x < ksize; ++x)
            kk[(size_t)xx * ksize + x] = w[x] < 0 ? (int32_t)(-0.5 + w[x] * (double)(1 << PRECISION_BITS)) : (int32_t)(0.5 + w[x] * (double)(1 << PRECISION_BITS));
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    *ksize_out = ksize;
}

// What the kernels rely on to drop the clamp of clip8: weights >= 0 and row sums within 2^22 +- 2^12.
bool coeffs_need_no_clamp(int out_size, int ksize, const std::vector<int32_t>& kk) {
    for (int xx = 0; xx < out_size; ++xx) {
        int64_t sum = 0;
        for (int x = 0; x < ksize; ++x) {
            if (kk[(size_t)xx * ksize + x] < 0) return false;
            sum += kk[(size_t)xx * ksize + x];
        }
        if (sum > (1 << PRECISION_BITS) + 4096 || sum < (1 << PRECISION_BITS) - 4096) return false;
    }
    return true;
}

struct ResizeParams {
    const uint8_t* src;
    uint8_t* dst;
    const int32_t *xb, *xk, *yb, *yk;      // device tables: bounds (out, 2), weights (out, ksize)
    int in_h, in_w, out_h, out_w, kx, ky, ty, pitch;
    size_t src_bytes;                       // whole source batch: word loads never start beyond its last byte
};

// Pillow's clip8().  The bilinear weights are non-negative and sum to 2^22 (+- rounding of <= ksize/2, checked when the table is
// built), so 0 < acc < 256 * 2^22 always holds and the clamp of clip8 never acts: the shift alone is exact.
__device__ __forceinline__ uint32_t clip8(int v) { return (uint32_t)v >> PRECISION_BITS; }
__device__ __forceinline__ int byte_of(uint32_t w, int k) { return (int)__byte_perm(w, 0u, 0x4440u + (uint32_t)k); }   // one PRMT

// KMAX > 0: at most KMAX taps per output column (checked on the host): a thread keeps the weights of its column in registers and
// reads the 3*KMAX source bytes of a row as aligned 32-bit words (realigned with funnel shifts), i.e. (3*KMAX+6)/4 loads instead of
// 3*KMAX byte loads -- the horizontal pass is bound by load/store instructions.  KMAX = 0: any tap count, byte loads.
template <int KMAX>
__global__ void __launch_bounds__(RS_THREADS) resize_bilinear_kernel(const ResizeParams p) {
    extern __shared__ __align__(16) uint8_t hbuf[];              // horizontal-pass rows [r0, r1) of this tile: pitch bytes each
    const int b = blockIdx.y;
    const int y0 = blockIdx.x * p.ty, y1 = min(y0 + p.ty, p.out_h);
    const int r0 = p.yb[2 * y0], r1 = p.yb[2 * (y1 - 1)] + p.yb[2 * (y1 - 1) + 1];
    const int nrows = r1 - r0;
    const uint8_t* img = p.src + (size_t)b * p.in_h * p.in_w * 3;
    // ---- horizontal pass (ImagingResampleHorizontal_8bpc)
    if (KMAX > 0) {
        constexpr int NW = ((3 * (KMAX > 0 ? KMAX : 1) - 1) >> 2) + 2;       // aligned words covering 3*KMAX bytes that start at any byte of the first
        const uintptr_t last_word = (reinterpret_cast<uintptr_t>(p.src) + p.src_bytes - 1) & ~(uintptr_t)3;
        // lanes run over output columns, the rest of the block over rows (narrow outputs keep every warp busy)
        const int xg = min(RS_THREADS, (p.out_w + 31) & ~31), rgroups = RS_THREADS / xg;
        const int rg = threadIdx.x / xg;
        for (int x = threadIdx.x - rg * xg; x < p.out_w && rg < rgroups; x += xg) {
            const int xmin = __ldg(p.xb + 2 * x), n = __ldg(p.xb + 2 * x + 1);
            int c[KMAX > 0 ? KMAX : 1];
#pragma unroll
            for (int t = 0; t < KMAX; ++t) c[t] = t < n ? __ldg(p.xk + x * p.kx + t) : 0;
            const uint8_t* s = img + ((size_t)(r0 + rg) * p.in_w + xmin) * 3;
            uint8_t* d = hbuf + rg * p.pitch + x * 3;
            for (int row = rg; row < nrows; row += rgroups, s += (size_t)rgroups * p.in_w * 3, d += rgroups * p.pitch) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(s);
                const uintptr_t wa = a & ~(uintptr_t)3;
                const uint32_t sh = (uint32_t)(a & 3) * 8;
                uint32_t w[NW];
#pragma unroll
                for (int i = 0; i < NW; ++i) w[i] = wa + 4 * i <= last_word ? __ldg(reinterpret_cast<const uint32_t*>(wa) + i) : 0u;
#pragma unroll
                for (int i = 0; i + 1 < NW; ++i) w[i] = __funnelshift_r(w[i], w[i + 1], sh);   // byte j of the taps = byte j of w[]
                int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll
                for (int t = 0; t < KMAX; ++t) {
                    a0 += byte_of(w[(3 * t) >> 2], (3 * t) & 3) * c[t];
                    a1 += byte_of(w[(3 * t + 1) >> 2], (3 * t + 1) & 3) * c[t];
                    a2 += byte_of(w[(3 * t + 2) >> 2], (3 * t + 2) & 3) * c[t];
                }
                d[0] = (uint8_t)clip8(a0);
                d[1] = (uint8_t)clip8(a1);
                d[2] = (uint8_t)clip8(a2);
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < nrows * p.out_w; idx += RS_THREADS) {      // one output pixel (3 channels) per thread and step
            const int row = idx / p.out_w, x = idx - row * p.out_w;
            const int xmin = __ldg(p.xb + 2 * x), n = __ldg(p.xb + 2 * x + 1);
            const uint8_t* s = img + ((size_t)(r0 + row) * p.in_w + xmin) * 3;
            const int32_t* k = p.xk + x * p.kx;
            int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
            for (int t = 0; t < n; ++t) {
                const int c = __ldg(k + t);
                a0 += (int)__ldg(s + 3 * t) * c;
                a1 += (int)__ldg(s + 3 * t + 1) * c;
                a2 += (int)__ldg(s + 3 * t + 2) * c;
            }
            uint8_t* d = hbuf + row * p.pitch + x * 3;
            d[0] = (uint8_t)clip8(a0);
            d[1] = (uint8_t)clip8(a1);
            d[2] = (uint8_t)clip8(a2);
        }
    }
    __syncthreads();
    // ---- vertical pass (ImagingResampleVertical_8bpc): eight (aligned rows) or four consecutive output bytes per thread and step
    const int rowbytes = p.out_w * 3;
    uint8_t* out = p.dst + (size_t)b * p.out_h * rowbytes;
    if (((reinterpret_cast<uintptr_t>(out) | (uintptr_t)rowbytes) & 7) == 0) {
        const int groups = rowbytes >> 3;
        for (int idx = threadIdx.x; idx < (y1 - y0) * groups; idx += RS_THREADS) {
            const int yy = idx / groups, g = idx - yy * groups, y = y0 + yy;
            const int ymin = __ldg(p.yb + 2 * y) - r0, n = __ldg(p.yb + 2 * y + 1);
            const int32_t* k = p.yk + y * p.ky;
            int a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = 1 << (PRECISION_BITS - 1);
            for (int t = 0; t < n; ++t) {
                const int c = __ldg(k + t);
                const uint2 w = *reinterpret_cast<const uint2*>(hbuf + (ymin + t) * p.pitch + 8 * g);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    a[j] += byte_of(w.x, j) * c;
                    a[4 + j] += byte_of(w.y, j) * c;
                }
            }
            *reinterpret_cast<uint2*>(out + (size_t)y * rowbytes + 8 * g) =
                make_uint2(clip8(a[0]) | (clip8(a[1]) << 8) | (clip8(a[2]) << 16) | (clip8(a[3]) << 24),
                           clip8(a[4]) | (clip8(a[5]) << 8) | (clip8(a[6]) << 16) | (clip8(a[7]) << 24));
        }
        return;
    }
    const int groups = (rowbytes + 3) >> 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(out) | (uintptr_t)rowbytes) & 3) == 0;
    for (int idx = threadIdx.x; idx < (y1 - y0) * groups; idx += RS_THREADS) {
        const int yy = idx / groups, g = idx - yy * groups, y = y0 + yy;
        const int ymin = __ldg(p.yb + 2 * y) - r0, n = __ldg(p.yb + 2 * y + 1);
        const int32_t* k = p.yk + y * p.ky;
        int a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = 1 << (PRECISION_BITS - 1);
        for (int t = 0; t < n; ++t) {
            const int c = __ldg(k + t);
            const uint32_t w = *reinterpret_cast<const uint32_t*>(hbuf + (ymin + t) * p.pitch + 4 * g);
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] += (int)((w >> (8 * j)) & 255u) * c;
        }
        uint8_t* d = out + (size_t)y * rowbytes + 4 * g;
        if (aligned) {
            *reinterpret_cast<uint32_t*>(d) = clip8(a[0]) | (clip8(a[1]) << 8) | (clip8(a[2]) << 16) | (clip8(a[3]) << 24);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * g + j < rowbytes) d[j] = (uint8_t)clip8(a[j]);
        }
    }
}

// Device copies of the coefficient tables, one entry per (device, in, out) pair, built on first use.
struct Table {
    int ksize = 0, max_span = 0;            // max_span: most input rows any run of `ty` outputs touches is derived from bounds on the host
    std::vector<int32_t> bounds, kk;
    int32_t *d_bounds = nullptr, *d_kk = nullptr;
};
std::mutex g_mu;
std::map<std::tuple<int, int, int>, Table*> g_tables;

int get_table(int in_size, int out_size, Table** out) {
    int dev = 0;
    CV_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_mu);
    auto key = std::make_tuple(dev, in_size, out_size);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { *out = it->second; return CV_OK; }
    Table* t = new Table();
    make_coeffs(in_size, out_size, &t->ksize, t->bounds, t->kk);
    if (!coeffs_need_no_clamp(out_size, t->ksize, t->kk)) {
        delete t;
        cv_set_error("cv_resize_bilinear_u8: coefficient table %d -> %d outside the range the kernel assumes", in_size, out_size);
        return CV_ERR_ARG;
    }
    CV_CUDA(cudaMalloc(&t->d_bounds, t->bounds.size() * sizeof(int32_t)));
    CV_CUDA(cudaMalloc(&t->d_kk, t->kk.size() * sizeof(int32_t)));
    CV_CUDA(cudaMemcpy(t->d_bounds, t->bounds.data(), t->bounds.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CV_CUDA(cudaMemcpy(t->d_kk, t->kk.data(), t->kk.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    g_tables[key] = t;
    *out = t;
    return CV_OK;
}

}  // namespace

extern "C" int cv_resize_coeffs_host(int in_size, int out_size, int* ksize, int32_t* bounds_host, int32_t* coeffs_host, int coeffs_capacity) {
    CV_ARG(in_size > 0 && out_size > 0 && ksize, "sizes must be positive");
    std::vector<int32_t> b, k;
    make_coeffs(in_size, out_size, ksize, b, k);
    if (bounds_host)
        for (size_t i = 0; i < b.size(); ++i) bounds_host[i] = b[i];
    if (coeffs_host) {
        CV_ARG((size_t)coeffs_capacity >= k.size(), "coefficient buffer too small (need out_size * ksize)");
        for (size_t i = 0; i < k.size(); ++i) coeffs_host[i] = k[i];
    }
    return CV_OK;
}

extern "C" int cv_resize_bilinear_u8(const uint8_t* src, int B, int in_h, int in_w, uint8_t* dst, int out_h, int out_w, void* stream) {
    CV_ARG(B >= 0, "negative batch");
    CV_ARG(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "sizes must be positive");
    CV_ARG(in_h <= 16384 && in_w <= 16384 && out_h <= 16384 && out_w <= 16384, "sizes above 16384 are not supported");
    if (B == 0) return CV_OK;
    CV_ARG(src && dst, "null data pointer");
    CV_ARG(B <= 65535, "at most 65535 images per call");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (in_h == out_h && in_w == out_w) {                       // Image.resize returns a copy when the size does not change
        CV_CUDA(cudaMemcpyAsync(dst, src, (size_t)B * in_h * in_w * 3, cudaMemcpyDeviceToDevice, s));
        return CV_OK;
    }
    Table *tx = nullptr, *ty = nullptr;
    int rc = get_table(in_w, out_w, &tx);
    if (rc) return rc;
    rc = get_table(in_h, out_h, &ty);
    if (rc) return rc;
    ResizeParams p{};
    p.src = src; p.dst = dst;
    p.xb = tx->d_bounds; p.xk = tx->d_kk; p.yb = ty->d_bounds; p.yk = ty->d_kk;
    p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.kx = tx->ksize; p.ky = ty->ksize;
    p.pitch = (out_w * 3 + 7) & ~7;
    p.src_bytes = (size_t)B * in_h * in_w * 3;
    int max_taps = 0;
    for (int x = 0; x < out_w; ++x) max_taps = std::max(max_taps, tx->bounds[2 * x + 1]);
    // rows per tile: as many as keep the horizontal-pass image of the tile within 64 KB of shared memory (>= 3 CTAs per SM)
    auto span = [&](int tyrows) {
        int m = 0;
        for (int y0 = 0; y0 < out_h; y0 += tyrows) {
            const int y1 = std::min(y0 + tyrows, out_h);
            m = std::max(m, ty->bounds[2 * (y1 - 1)] + ty->bounds[2 * (y1 - 1) + 1] - ty->bounds[2 * y0]);
        }
        return m;
    };
    int rows = 32;
    while (rows > 1 && (size_t)span(rows) * p.pitch > 64 * 1024) rows >>= 1;
    const size_t smem = (size_t)span(rows) * p.pitch;
    if (smem > 200 * 1024) { cv_set_error("cv_resize_bilinear_u8: one output row needs %zu bytes of shared memory", smem); return CV_ERR_ARG; }
    p.ty = rows;
    const dim3 grid((out_h + rows - 1) / rows, B);
#define CV_RESIZE_LAUNCH(K)                                                                                              \
    do {                                                                                                                 \
        CV_CUDA(cudaFuncSetAttribute(resize_bilinear_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        resize_bilinear_kernel<K><<<grid, RS_THREADS, smem, s>>>(p);                                                     \
    } while (0)
    if (max_taps <= 3) CV_RESIZE_LAUNCH(3);             // enlarging (two taps, three at a clipped border)
    else if (max_taps <= 4) CV_RESIZE_LAUNCH(4);        // shrinking by up to 1.5x or so (400 -> 256)
    else if (max_taps <= 5) CV_RESIZE_LAUNCH(5);        // shrinking by up to 2x (512 -> 256)
    else if (max_taps <= 9) CV_RESIZE_LAUNCH(9);        // up to 4x
    else CV_RESIZE_LAUNCH(0);
#undef CV_RESIZE_LAUNCH
    CV_CHECK_LAUNCH();
    return CV_OK;
}
