// common.cuh - error handling, sm_100a PTX wrappers (mbarrier / TMA / tcgen05), math helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace ptts {

// ---------------------------------------------------------------- errors
extern thread_local char g_last_error[512];
void set_error(const char *fmt, ...);

#define PTTS_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::ptts::set_error("CUDA error %s at %s:%d: %s%s", cudaGetErrorName(_e), __FILE__,   \
                              __LINE__, cudaGetErrorString(_e),                                 \
                              _e == cudaErrorLaunchFailure ? " (a device-side trap: every mbarrier / grid-barrier wait in the kernels is bounded and traps on time-out)" : ""); \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

#define PTTS_CHECK(cond, ...)                                                                   \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            ::ptts::set_error(__VA_ARGS__);                                                     \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

// ---------------------------------------------------------------- small device math
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the engine is launched with cudaLaunchAttributeProgrammaticStreamSerialization
// (engine.cuh launch_k): the next kernel's CTAs may become resident while this one is still
// running. pdl_wait() blocks until the preceding grid has completed and its writes are visible;
// it must precede the first access to any buffer another kernel writes. pdl_trigger() lets the
// following grid start being scheduled. Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2, ACT_SILU = 3, ACT_ELU = 4 };

__device__ __forceinline__ float act_apply(int act, float x) {
    switch (act) {
        case ACT_GELU_ERF: return 0.5f * x * (1.0f + erff(x * 0.7071067811865475f));       // ptts_flowlm.c:207-210
        case ACT_GELU_TANH: return 0.5f * x * (1.0f + tanhf(0.7978845608f * (x + 0.044715f * x * x * x)));  // ptts_mimi.c:208-211
        case ACT_SILU: return x / (1.0f + expf(-x));                                        // ptts_flowlm.c:199-201
        case ACT_ELU: return x >= 0.0f ? x : (expf(x) - 1.0f);                               // ptts_kernels.c:215-218
        default: return x;
    }
}

// 16-bit storage helpers: `bf` selects bf16 (true) or fp16 (false) encoding of a 16-bit lane.
__device__ __forceinline__ uint16_t f2h16(float v, bool bf) {
    return bf ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
}
__device__ __forceinline__ float h162f(uint16_t u, bool bf) {
    return bf ? __bfloat162float(__ushort_as_bfloat16(u)) : __half2float(__ushort_as_half(u));
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait (about one second): a pipeline bug must trap and surface as a CUDA error (cudaErrorLaunchFailure; the host message
// says so) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    // a failed try_wait comes back after ~0.1 us (measured: 11 iterations per ~1 us wait); 2^26 of them are several seconds, and at
    // least 0.7 s even if every attempt returned at once - longer than any legitimate wait (the longest: a kernel launched early by
    // programmatic dependent launch behind a predecessor of tens of milliseconds). Counting keeps the loop at five instructions and
    // holds no 64-bit clock value in registers (these loops are 20-30 % of the instructions the GEMM kernels execute)
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == (1u << 26)) __trap();   // no device printf here: the call alone costs k_flow_mega 8 us per frame in register allocation
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 3D tiled load: coordinates (c0 = innermost element index, c1 = row, c2 = batch). Out-of-bounds
// elements (including negative coordinates) are zero-filled; the full box byte count is always
// credited to the mbarrier.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 3D tiled STORE shared -> global (bulk async group): the box is read from shared memory in the tensor map's
// swizzle and written as whole lines; elements outside the tensor are clipped.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(smem_src)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_s(const CUtensorMap *map, uint32_t smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_src), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }              // stores complete

// 1D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 or bf16 inputs, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Wait for outstanding tcgen05.ld and tie the destination registers to the wait, so the compiler
// cannot hoist their uses above it (the loads are asynchronous until the wait retires).
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Activations for 16-bit outputs (Mimi): hardware exp / tanh approximations (relative error
// ~2^-11, the same as the fp16 rounding that follows).
__device__ __forceinline__ float act_apply_fast(int act, float x) {
    switch (act) {
        case ACT_GELU_TANH: return 0.5f * x * (1.0f + tanh_fast(0.7978845608f * (x + 0.044715f * x * x * x)));
        case ACT_ELU: return x >= 0.0f ? x : (__expf(x) - 1.0f);
        case ACT_SILU: return __fdividef(x, 1.0f + __expf(-x));
        case ACT_GELU_ERF: return 0.5f * x * (1.0f + erff(x * 0.7071067811865475f));
        default: return x;
    }
}

// Compile-time activation (the epilogues dispatch on `act` once per 16-column chunk, not per element;
// everything inside is straight-line code: selects instead of branches).
template <int ACT, bool FAST>
__device__ __forceinline__ float act_t(float x) {
    if constexpr (ACT == ACT_GELU_ERF) return 0.5f * x * (1.0f + erff(x * 0.7071067811865475f));
    else if constexpr (ACT == ACT_GELU_TANH)
        return 0.5f * x * (1.0f + (FAST ? tanh_fast(0.7978845608f * (x + 0.044715f * x * x * x)) : tanhf(0.7978845608f * (x + 0.044715f * x * x * x))));
    else if constexpr (ACT == ACT_SILU) return FAST ? __fdividef(x, 1.0f + __expf(-x)) : x / (1.0f + expf(-x));
    else if constexpr (ACT == ACT_ELU) {
        const float e = (FAST ? __expf(fminf(x, 0.0f)) : expf(fminf(x, 0.0f))) - 1.0f;
        return x >= 0.0f ? x : e;
    } else return x;
}
// 16 fp32 values -> activation -> 8 packed 16-bit pairs (hi) and optionally the rounding residues (lo)
template <int ACT, bool FAST, bool BF, bool LO>
__device__ __forceinline__ void act_pack16(const float (&v)[16], uint32_t (&hw)[8], uint32_t (&lw)[8]) {
    if constexpr (ACT == ACT_ELU && FAST && !BF && !LO) {
        // fp16 output: ELU evaluated on packed halves, ELU(x) = max(x,0) + (2^(min(x,0)*log2 e) - 1). h2exp2 goes through fp32
        // (two MUFU.EX2 and conversions per pair); ex2.approx.f16x2 would save 4 instructions per pair but costs 1-1.7 dB of PCM
        // parity for -1 % of the decode time (profiles/r02_step_kernel_experiments.txt, experiment 14)
        const __half2 zero = __float2half2_rn(0.0f), one = __float2half2_rn(1.0f), l2e = __float2half2_rn(1.4426950408889634f);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            const __half2 e = h2exp2(__hmul2(__hmin2(h, zero), l2e));
            const __half2 r = __hadd2(__hmax2(h, zero), __hsub2(e, one));
            hw[j] = *(const uint32_t *)&r;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float a0 = act_t<ACT, FAST>(v[2 * j]), a1 = act_t<ACT, FAST>(v[2 * j + 1]);
        if constexpr (BF) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
            hw[j] = *(const uint32_t *)&h;
            if constexpr (LO) {
                const __nv_bfloat162 l = __floats2bfloat162_rn(a0 - __low2float(h), a1 - __high2float(h));
                lw[j] = *(const uint32_t *)&l;
            }
        } else {
            const __half2 h = __floats2half2_rn(a0, a1);
            hw[j] = *(const uint32_t *)&h;
            if constexpr (LO) {
                const __half2 l = __floats2half2_rn(a0 - __low2float(h), a1 - __high2float(h));
                lw[j] = *(const uint32_t *)&l;
            }
        }
    }
}
// Run-time dispatch to the compile-time variants actually used by the engine (one uniform branch per chunk).
__device__ __forceinline__ void act_pack16_dispatch(int act, bool fast, bool bf, bool lo, const float (&v)[16], uint32_t (&hw)[8],
                                                    uint32_t (&lw)[8]) {
    if (bf) {   // FlowLM: bf16, precise math
        if (lo) {
            if (act == ACT_GELU_ERF) act_pack16<ACT_GELU_ERF, false, true, true>(v, hw, lw);
            else if (act == ACT_SILU) act_pack16<ACT_SILU, false, true, true>(v, hw, lw);
            else act_pack16<ACT_NONE, false, true, true>(v, hw, lw);
        } else {
            if (act == ACT_GELU_ERF) act_pack16<ACT_GELU_ERF, false, true, false>(v, hw, lw);
            else if (act == ACT_SILU) act_pack16<ACT_SILU, false, true, false>(v, hw, lw);
            else act_pack16<ACT_NONE, false, true, false>(v, hw, lw);
        }
    } else if (fast) {   // Mimi: fp16, hardware exp / tanh
        if (act == ACT_ELU) act_pack16<ACT_ELU, true, false, false>(v, hw, lw);
        else if (act == ACT_GELU_TANH) act_pack16<ACT_GELU_TANH, true, false, false>(v, hw, lw);
        else act_pack16<ACT_NONE, true, false, false>(v, hw, lw);
    } else {
        if (act == ACT_ELU) act_pack16<ACT_ELU, false, false, false>(v, hw, lw);
        else if (act == ACT_GELU_TANH) act_pack16<ACT_GELU_TANH, false, false, false>(v, hw, lw);
        else act_pack16<ACT_NONE, false, false, false>(v, hw, lw);
    }
}

// Shared-memory matrix descriptor for a K-major tile whose rows are exactly 128 bytes
// (64 16-bit elements) laid out by TMA with CU_TENSOR_MAP_SWIZZLE_128B: 8-row core groups
// are 1024 bytes apart (SBO), LBO is unused for swizzled K-major, descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  (Field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor.)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // LBO (ignored), bits [16,30)
    d |= (uint64_t)(1024u >> 4) << 32;               // SBO = 1024 B, bits [32,46)
    d |= (uint64_t)1 << 46;                          // version = 1
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16, fp32 accumulate, A and B K-major.
// (Field layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor.)
__host__ __device__ inline uint32_t umma_idesc_f16(int M, int N, bool a_bf16, bool b_bf16) {
    uint32_t d = 0;
    d |= 1u << 4;                        // c_format = F32
    d |= (a_bf16 ? 1u : 0u) << 7;        // a_format: 0 = F16, 1 = BF16
    d |= (b_bf16 ? 1u : 0u) << 10;       // b_format
    d |= (uint32_t)(N >> 3) << 17;       // n_dim
    d |= (uint32_t)(M >> 4) << 24;       // m_dim
    return d;
}

}  // namespace ptts
