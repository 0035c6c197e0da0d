// Shared device helpers for the sm_100a kernels of the ViT-detector hot path.
//
// Everything here is a thin wrapper around one PTX instruction (mbarrier, TMA bulk-tensor copy,
// tcgen05 MMA / TMEM load / TMEM allocation) or a piece of activation arithmetic that several
// kernels share.  No CUTLASS/CuTe types: the kernels in this directory are written against the
// PTX ISA directly.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vitdet {

// ---------------------------------------------------------------------------------------------
// Activation ids shared by host and device (mirrors `use_mish` of the reference,
// vision_transformer_detector.py:396-402 and :478-483).
// ---------------------------------------------------------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_MISH = 1, ACT_GELU = 2 };

// Mish(x) = x * tanh(softplus(x))  (tfa.activations.mish, reference det.py:129).
// With n = e^x:  tanh(log(1+n)) = ((1+n)^2 - 1) / ((1+n)^2 + 1) = 1 - 2 / (n^2 + 2n + 2), so
//   mish(x) = x - 2x / (n^2 + 2n + 2):  one ex2 and one rcp on the SFU.
// The fast form needs no clamp: n -> inf gives rcp(inf) = 0 and mish = x; n -> 0 gives x - x = 0.
// It is 7 instructions (FMUL, MUFU.EX2, FADD, FFMA, MUFU.RCP, FMUL, FFMA); the epilogues of the wide
// layers are issue/SFU-bound, so the count matters (profiles/r01b_ncu_mlp1.md: 24.5 -> 9 per element).
template <bool PRECISE>
__device__ __forceinline__ float mish(float x) {
    if (PRECISE) {
        float n = expf(fminf(x, 30.f));
        float a = n * (n + 2.f);
        return x * (a / (a + 2.f));
    } else {
        float n, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(n) : "f"(x * 1.4426950408889634f));
        const float q = fmaf(n, n + 2.f, 2.f);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
        return fmaf(x * -2.f, r, x);
    }
}

// tfa.layers.GELU() default approximate=True (reference det.py:402):
// 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) )).
template <bool PRECISE>
__device__ __forceinline__ float gelu_tanh(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (x + k1 * x * x * x);
    float t;
    if (PRECISE) {
        t = tanhf(u);
    } else {
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    }
    return 0.5f * x * (1.f + t);
}

template <int ACT, bool PRECISE>
__device__ __forceinline__ float apply_act(float x) {
    if (ACT == ACT_MISH) return mish<PRECISE>(x);
    if (ACT == ACT_GELU) return gelu_tanh<PRECISE>(x);
    return x;
}

// Activation of the fp32-accumulate mode's tensor-core epilogues.  Mish keeps the one-ex2 / one-rcp form: ex2.approx and
// rcp.approx are accurate to 2 and 1 float32 ulp, the form's absolute error is 4e-7 |x| (tests: test_mish_fast_form_accuracy),
// three orders of magnitude inside the mode's 1e-3 budget, and expf + an IEEE division made the 28 -> 3584 layer's epilogue
// three times slower.  tanh.approx (2^-11) is NOT good enough for GELU there: tanhf.
template <int ACT>
__device__ __forceinline__ float apply_act_f32acc(float x) {
    if (ACT == ACT_MISH) return mish<false>(x);
    if (ACT == ACT_GELU) return gelu_tanh<true>(x);
    return x;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------
// Packed float32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 work on an aligned register pair, one issue slot for two
// lanes of arithmetic).  The epilogues are issue-bound (profiles/r01c_ncu_gemm_mlp1.md), so the bias add and the
// five FMA-pipe steps of Mish are done on pairs: 4.5 instead of 7 instructions per element, same roundings.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// mish<false> on a pair (same operations and roundings as the scalar form above).
// Measured and dropped (profiles/r02_experiments.md): a one-SFU form — reciprocal of n^2 + 2n + 2 from an integer seed and
// three packed Newton steps on the FMA pipe, float32-accurate — made the 28 -> 3584 layer SLOWER (1.39 -> 1.48 ms per
// step): that epilogue is bound by instruction issue, not by the SFU, and the form costs 8 instead of 6.5 slots per element.
// Also measured and dropped (r03, gpurun_out/r03a_ab.log): ONE reciprocal per pair, 1 / (q0 q1), and two multiplications
// (1.5 SFU operations per element): 1.381 -> 1.341 ms per step on that layer although ncu has its XU pipe at 85 % — not worth
// two more roundings in the activation.  And (r03k_mish_poly_ab.log) one of four / one of two of the exponentials on the FMA pipe
// (Cody-Waite + degree-5 polynomial, 2.1e-7, 11 issue slots): 1.382 -> 1.416 / 1.500 ms.  Every variant that adds instructions
// to this epilogue loses, whatever it takes off the SFU: the four epilogue warps of a sub-partition are bound by their own
// in-order instruction streams (ncu: 0.87 eligible warps per scheduler, stalls "wait" and "mio_throttle"), and 576 threads at
// 96 registers are all the register file allows.
__device__ __forceinline__ uint64_t mish2_fast(uint64_t x) {
    float t0, t1, n0, n1, q0, q1, r0, r1;
    f2_unpack(f2_mul(x, f2_pack(1.4426950408889634f, 1.4426950408889634f)), t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(n0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(n1) : "f"(t1));
    const uint64_t two = f2_pack(2.f, 2.f), n = f2_pack(n0, n1);
    f2_unpack(f2_fma(n, f2_add(n, two), two), q0, q1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(q0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(q1));
    return f2_fma(f2_mul(x, f2_pack(-2.f, -2.f)), f2_pack(r0, r1), x);
}

// act(a + b) for two neighbouring columns, rounded to a packed bf16 pair (the epilogues' inner step).
template <int ACT>
__device__ __forceinline__ uint32_t bias_act_bf16x2(float a0, float a1, float b0, float b1) {
    if (ACT == ACT_MISH) {
        float y0, y1;
        f2_unpack(mish2_fast(f2_add(f2_pack(a0, a1), f2_pack(b0, b1))), y0, y1);
        return pack_bf16x2(y0, y1);
    }
    if (ACT == ACT_NONE) {
        float y0, y1;
        f2_unpack(f2_add(f2_pack(a0, a1), f2_pack(b0, b1)), y0, y1);
        return pack_bf16x2(y0, y1);
    }
    return pack_bf16x2(apply_act<ACT, false>(a0 + b0), apply_act<ACT, false>(a1 + b1));
}

// ---------------------------------------------------------------------------------------------
// Coalesced float32 epilogue store for wide outputs.  In the epilogues a thread owns one ROW of the accumulator tile
// (its TMEM lane), so storing its values directly makes every warp-wide 16-byte access touch 32 different rows — fine for
// the default 28-wide residual stream (the whole row is one 112-byte segment) but 8x sector over-fetch on the residual
// read and scattered partial-line writes when the row is hundreds of floats wide (embedding_dim 768).  Here the lane's
// 16 finished values of a 32-row x 16-column half chunk go through the warp's 2 KB staging tile (64 B per row, 16-byte
// chunk c of row r at c ^ ((r >> 1) & 3): conflict-free both ways) and come back as 8 rows x 64 contiguous bytes per
// instruction; the residual is added on that side from values fetched with the same coalesced pattern (prefetch_resid_chunk).
// ---------------------------------------------------------------------------------------------
// The residual values of a 32 x 32 chunk in that coalesced mapping ([half][i]: row 8 i + lane / 4, columns 16 half + 4 (lane & 3)
// .. + 3).  Issued right after the chunk's tcgen05.ld, BEFORE its wait, so that the global-load latency hides behind the
// TMEM load and the activation arithmetic instead of sitting between the staging-tile read and the store.
__device__ __forceinline__ void prefetch_resid_chunk(float4 (&rr)[2][4], const float* resid, int ldr, int lane, int row0, int M, int col0) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 8 * i + (lane >> 2);
            rr[hh][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (resid != nullptr && row < M)
                rr[hh][i] = *reinterpret_cast<const float4*>(resid + static_cast<size_t>(row) * ldr + col0 + 16 * hh + 4 * (lane & 3));
        }
}

__device__ __forceinline__ void store_f32_half_chunk_coalesced(uint32_t s_tile, const float (&y)[16], int lane, int row0, int M,
                                                               float* out, int ldc, const float4 (&rr4)[4], int col0) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t dst = s_tile + static_cast<uint32_t>(lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(y[4 * g]), "f"(y[4 * g + 1]), "f"(y[4 * g + 2]),
                     "f"(y[4 * g + 3]) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2), ch = lane & 3;
        const uint32_t src = s_tile + static_cast<uint32_t>(r * 64 + ((ch ^ ((r >> 1) & 3)) << 4));
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src) : "memory");
        const int row = row0 + r;
        if (row < M) {
            const int col = col0 + 4 * ch;
            v.x += rr4[i].x; v.y += rr4[i].y; v.z += rr4[i].z; v.w += rr4[i].w;
            *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * ldc + col) = v;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (see launch.h)
// ---------------------------------------------------------------------------------------------
// Lets the next kernel of the stream start its prologue; call as early as possible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Returns when the predecessor grid has completed and its global writes are visible.  No-op without PDL.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Shared-memory addresses, mbarriers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Makes mbarrier.init visible to the async proxy (TMA / tcgen05.commit arrive on these barriers).
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Generic-proxy writes to shared memory -> visible to async-proxy readers (TMA store, UMMA).
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure reported
// through the C ABI), never as a hung GPU.  try_wait suspends the thread in hardware for a
// time slice, so the poll count is small on the normal path; the clock is only read after many
// failed polls.  ~4 s at 2 GHz is far beyond any legitimate wait in these kernels — but clock64 keeps counting
// while a context is preempted or time-sliced (MPS, a debugger, an oversubscribed GPU), so deployments that share
// the GPU build with -DVITDET_NO_MBAR_WATCHDOG and get plain unbounded waits.
#ifndef VITDET_NO_MBAR_WATCHDOG
#define VITDET_WATCHDOG_CHECK(polls, mask, t0)                        \
    if ((++polls & (mask)) == 0) {                                    \
        long long now = clock64();                                    \
        if (t0 == 0) t0 = now;                                        \
        else if (now - t0 > 8000000000ll) __trap();                   \
    }
#else
#define VITDET_WATCHDOG_CHECK(polls, mask, t0) (void)polls; (void)t0;
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        VITDET_WATCHDOG_CHECK(polls, 0x3fffu, t0)
    }
}

// Wait used where latency does not matter (a producer running stages ahead): back off with nanosleep
// so that the polling lane does not steal issue slots from the compute warps of its sub-partition.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(200);
        VITDET_WATCHDOG_CHECK(polls, 0x3ffu, t0)
    }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — tiled 2-D loads into 128B-swizzled shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// L2 eviction-priority policies for TMA loads (createpolicy encodings used as immediates).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d_hint(uint32_t smem_dst, const CUtensorMap* map,
                                                 uint32_t bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

// TMA store of a 2-D box from shared memory (bulk async-group completion).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Waits until the bulk groups of this thread have finished READING their shared-memory source.
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM -> register loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_result_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// One lane of a converged warp (the compiler knows the result is warp-uniform, so code under `if (elect_one())`
// can use the uniform datapath — unlike `if (lane == 0)`, which made the MMA issue loop 98 instructions per k-block).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, f32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: A (bf16, 128 rows = lanes, two K elements per 32-bit column) read from
// tensor memory, B through a shared-memory descriptor.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// All previously issued tcgen05.mma of this thread arrive (once) on `bar` when they complete.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i),
// columns [col, col+32).  A warp may only touch the TMEM lane quadrant 32*(warp_id % 4).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// Register -> TMEM stores, same lane/column addressing as the loads.
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): cluster helpers, 2-SM TMA / MMA / commit / TMEM allocation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair.
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_result_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T over the CTA pair: 256 x N x 16, issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Completion of all prior MMAs of this thread arrives on the barrier at the same offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle (what a TMA box of 64 bf16 x R rows
// with CU_TENSOR_MAP_SWIZZLE_128B produces): rows are 128 B, 8 rows form a 1024 B swizzle atom.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused: 1)
//   bits [32,46) stride byte offset >> 4 (1024 B between 8-row groups)
//   bits [46,48) descriptor version = 1 (sm_100)     bits [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// MN-major operand tile (the N/M index is contiguous): rows are K steps of 128 B (64 bf16 along MN),
// 8 rows form a 1024 B swizzle atom; what a TMA box of 64 bf16 x R rows produces for a matrix stored
// [K][MN].  SBO = 1024 B between 8-row K groups; LBO (stride between 64-element MN chunks) unused for
// MN <= 64.  Used for V in the attention PV product.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
    return umma_desc_sw128_kmajor(smem_addr);     // same fields; the major-ness lives in the instruction descriptor
}

// Instruction descriptor for kind::f16 with bf16 A/B (K-major both), f32 accumulator, M x N tile:
//   [4,6) c_format = 1 (f32)   [7,10) a_format = 1 (bf16)   [10,13) b_format = 1 (bf16)
//   bit 15 / 16 a_major / b_major = 0 (K-major)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_f32(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}
// Same with B MN-major (bit 16).
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_f32_bmn(int m, int n) {
    return umma_idesc_bf16_f32(m, n) | (1u << 16);
}

}  // namespace vitdet
