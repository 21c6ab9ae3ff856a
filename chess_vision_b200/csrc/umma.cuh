// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA bulk copies (cp.async.bulk),
// tcgen05 (TMEM allocation, UMMA issue/commit, TMEM loads) and the shared-memory / instruction descriptors.
// Hand-written PTX; field layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (cross-checked against the CUTLASS sm100 headers vendored in this image).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace umma {

// Waits cannot hang the GPU: after kWaitTimeoutNs of polling a barrier the thread traps (-> CUDA error).
constexpr uint64_t kWaitTimeoutNs = 2000000000ull;
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (uniform predicate: ptxas emits straight-line UTCHMMA sequences for the elected thread, where
// `if (lane == 0)` makes it wrap every tcgen05 instruction in an ELECT/branch loop that costs ~40 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // Hot path: back-to-back try_wait (each may suspend the thread for a short, hardware-chosen time).  The global timer is
    // slow to read, so the hang guard only looks at it once every 4096 failed polls.
    uint32_t polls = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
#ifdef CV_WAIT_SLEEP_NS
        __nanosleep(CV_WAIT_SLEEP_NS);
#endif
        if ((++polls & 4095u) == 0) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kWaitTimeoutNs) __trap();
        }
    }
}

// Busy-polling variant (mbarrier.test_wait never suspends the thread): for waits on the critical path of an op-synchronous
// kernel, where every other warp of the CTA is waiting for the same event anyway.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, polls = 0;
    uint64_t t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && (++polls & 0xFFFFu) == 0) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kWaitTimeoutNs) __trap();
        }
    } while (!ok);
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ---------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05: TMEM allocation -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_NONE ("interleaved" core matrices):
//   a core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
//   SBO = byte distance between core matrices adjacent in M/N (next 8 rows);
//   LBO = byte distance between core matrices adjacent in K (next 16 bytes of K).
//   bits [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// (A and B must have the same 16-bit format: bf16 activations x fp16 weights raises an illegal-instruction fault on B200.)
// Same with A = B = fp16 (format code 0).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Descriptor halves: lo = addr>>4 | (LBO>>4) << 16, hi = SBO>>4 | version.  An operand window that moves by whole 16-byte
// rows is `lo + rows`: one integer add per MMA in the single issuing thread (whose dependent-instruction latency is exposed).
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

// ---- tcgen05.mma / commit / ld -----------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA (SASS: UTCHMMA)
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same instruction, descriptors passed as 32-bit halves (kind::f16 covers fp16 x fp16 and bf16 x bf16; the idesc says which)
__device__ __forceinline__ void mma_f16_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// two fp32 -> packed bf16x2 with ReLU in the conversion (lo -> bits [0,16), hi -> bits [16,32))
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// ---- 16-bit operand formats of the fused stages: F16 = true -> IEEE fp16 (11-bit significand; the default mode), false -> bf16.
// pk2 / pk2r: two fp32 -> one packed pair (first argument in bits [0,16)), pk2r with ReLU folded into the conversion;
// up2: packed pair -> float2.  fp16 overflows to +-inf above 65504: the fused kernels raise a device flag (see cv_square) and the
// bf16 kernels redo the wave.
template <bool F16>
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
    uint32_t d;
    if (F16) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t pk2r(float lo, float hi) {
    uint32_t d;
    if (F16) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <bool F16>
__device__ __forceinline__ float2 up2(uint32_t p) {
    if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&p));
    return make_float2(__uint_as_float(p << 16), __uint_as_float(p & 0xffff0000u));
}
// any fp16 lane of the packed pair is +-inf or NaN (exponent field all ones)
__device__ __forceinline__ uint32_t f16x2_nonfinite(uint32_t p) { return ((p & 0x7C007C00u) + 0x04000400u) & 0x80008000u; }
template <bool F16>
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N) {
    return (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// arrives on `bar` once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane base + t) (SASS: LDTM)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Crop order of the fused pipeline ("square-major inside groups of 32 boards").  Stage D's tile is 32 consecutive crops and
// its pooled features go to the operand layout of the global head, which interleaves 128 BOARDS at 16-byte granularity: with crops in
// board order a tile is half of ONE board and every lane's 16-byte store lands 245 KB from its neighbour's.  In the permuted order a tile
// is one square of 32 consecutive boards, and a warp's store is 512 contiguous bytes.  The front end applies the order when it picks
// the board window of output position n' (every stage between is indifferent to the crop order); for the sources it does not serve,
// stage C's last epilogue permutes its 1.5 KB per crop instead.  Position n' of crop (b, sq) in a launch of nb boards:
//   full groups g = b / 32 < nb / 32:  n' = g * 2048 + sq * 32 + b % 32;      last partial group of r = nb % 32 boards:  n' = full * 2048 + sq * r + (b - 32 full)
__device__ __forceinline__ int64_t perm_pos(int64_t n, int nb) {
    const int64_t b = n >> 6;
    const int sq = (int)(n & 63), full = nb >> 5;
    const int64_t g = b >> 5;
    if (g < full) return g * 2048 + sq * 32 + (b & 31);
    return (int64_t)full * 2048 + (int64_t)sq * (nb & 31) + (b - (int64_t)full * 32);
}
__device__ __forceinline__ void perm_inv(int64_t np, int nb, int64_t& b, int& sq) {
    const int full = nb >> 5;
    const int64_t g = np >> 11;
    if (g < full) {
        const int rem = (int)(np & 2047);
        sq = rem >> 5;
        b = g * 32 + (rem & 31);
    } else {
        const int rem = (int)(np - (int64_t)full * 2048), r = nb & 31;
        sq = rem / r;
        b = (int64_t)full * 32 + (rem - sq * r);
    }
}

}  // namespace umma
