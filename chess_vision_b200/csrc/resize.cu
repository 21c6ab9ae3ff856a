// Board resize of the reference's input transform on the GPU (the step right before the hot path).
// Replaces `transforms.Resize((S, S))` on a PIL image (/root/reference/dataset.py:177-181, fed at predict.py:19-20), i.e. Pillow's
// `Image.resize((S, S), BILINEAR)` (src/libImaging/Resample.c, Pillow 12.2.0): a separable triangle filter whose support grows with
// the shrink factor (antialiasing), int32 fixed-point weights with 22 fractional bits, the horizontal result rounded to uint8
// before the vertical pass.  Integer work: the output is BIT-EXACT with Pillow (tests/test_resize_gpu.py, golden vectors made by
// Pillow itself).  ToTensor + Normalize stay fused into the crop gather of the hot path (cv_square_forward_u8).
//
// One kernel, HBM-bound: a CTA produces TY output rows of one image.  It (1) runs the horizontal pass over the input rows those
// output rows need -- source bytes read straight from global memory (neighbouring threads share their taps through L1) -- into a
// shared-memory image of uint8 rows, (2) runs the vertical pass from shared memory, four output bytes per thread, one 32-bit
// coalesced store each.  Algorithmic bytes per image: in_h*in_w*3 read + out_h*out_w*3 written.
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

struct ResizeParams {
    const uint8_t* src;
    uint8_t* dst;
    const int32_t *xb, *xk, *yb, *yk;      // device tables: bounds (out, 2), weights (out, ksize)
    int in_h, in_w, out_h, out_w, kx, ky, ty, pitch;
};

__device__ __forceinline__ uint32_t clip8(int v) { return (uint32_t)min(max(v >> PRECISION_BITS, 0), 255); }

__global__ void __launch_bounds__(RS_THREADS) resize_bilinear_kernel(const ResizeParams p) {
    extern __shared__ __align__(16) uint8_t hbuf[];              // horizontal-pass rows [r0, r1) of this tile: pitch bytes each
    const int b = blockIdx.y;
    const int y0 = blockIdx.x * p.ty, y1 = min(y0 + p.ty, p.out_h);
    const int r0 = p.yb[2 * y0], r1 = p.yb[2 * (y1 - 1)] + p.yb[2 * (y1 - 1) + 1];
    const int nrows = r1 - r0;
    const uint8_t* img = p.src + (size_t)b * p.in_h * p.in_w * 3;
    // ---- horizontal pass (ImagingResampleHorizontal_8bpc): one output pixel (3 channels) per thread and step
    for (int idx = threadIdx.x; idx < nrows * p.out_w; idx += RS_THREADS) {
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
    __syncthreads();
    // ---- vertical pass (ImagingResampleVertical_8bpc): four consecutive output bytes per thread and step
    const int rowbytes = p.out_w * 3, groups = (rowbytes + 3) >> 2;
    uint8_t* out = p.dst + (size_t)b * p.out_h * rowbytes;
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
    p.pitch = (out_w * 3 + 3) & ~3;
    // rows per tile: as many as keep the horizontal-pass image of the tile within 64 KB of shared memory (>= 3 CTAs per SM)
    auto span = [&](int tyrows) {
        int m = 0;
        for (int y0 = 0; y0 < out_h; y0 += tyrows) {
            const int y1 = std::min(y0 + tyrows, out_h);
            m = std::max(m, ty->bounds[2 * (y1 - 1)] + ty->bounds[2 * (y1 - 1) + 1] - ty->bounds[2 * y0]);
        }
        return m;
    };
    int rows = 16;
    while (rows > 1 && (size_t)span(rows) * p.pitch > 64 * 1024) rows >>= 1;
    const size_t smem = (size_t)span(rows) * p.pitch;
    if (smem > 200 * 1024) { cv_set_error("cv_resize_bilinear_u8: one output row needs %zu bytes of shared memory", smem); return CV_ERR_ARG; }
    p.ty = rows;
    CV_CUDA(cudaFuncSetAttribute(resize_bilinear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resize_bilinear_kernel<<<dim3((out_h + rows - 1) / rows, B), RS_THREADS, smem, s>>>(p);
    CV_CHECK_LAUNCH();
    return CV_OK;
}
