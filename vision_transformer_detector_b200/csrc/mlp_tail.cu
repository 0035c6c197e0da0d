// Fused tail of the encoder MLP pyramid: the last three Dense + activation layers of a block, the residual
// add and the LayerNorm that follows (reference det.py:388-412, and det.py:353-357 of the next block), in ONE
// kernel.  In the default model these are 224 -> 112 -> 56 -> 28: as separate GEMM launches they are pure
// latency (one k-block of work per tile), 0.6 ms per forward; here the intermediate activations never leave
// the SM — the epilogue of layer l writes them to shared memory in the 128B-swizzled K-major layout that the
// tcgen05.mma of layer l+1 reads as its A operand.
//
// One persistent CTA per SM, 288 threads, a tile = 128 token rows:
//   warp 8      control: loads the three weight matrices once (TMA), then per tile: A tile by TMA, MMA of
//               layer 0, waits for act0, prefetches the NEXT tile's A (its buffer is free once MMA 0 is done),
//               MMA 1, waits for act1, MMA 2.  Accumulators D0 | D1 | D2 live in TMEM columns 0 | 128 | 192.
//   warps 0..7  epilogues (thread = row = TMEM lane; warps w and w+4 split the column chunks of layers 0, 1): tcgen05.ld, bias, Mish/GELU, bf16, st.shared (swizzled),
//               fence.proxy.async, arrive; the last one adds the residual, writes the f32 residual stream and
//               the LayerNorm'd bf16 row for the next block.
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace vitdet {

namespace {

constexpr int kTM = 128;
constexpr int kTK = 64;
constexpr int kTailEpiWarps = 8;        // two per TMEM lane quadrant: they split the 32-column chunks of layers 0 and 1
constexpr int kTailCtlWarp = kTailEpiWarps;
constexpr int kTailThreads = 32 * (kTailEpiWarps + 1);
constexpr int kATileBytes = kTM * kTK * 2;     // 16 KiB per k-block of an A operand
constexpr int kTailTmemCols = 256;

struct TailArgs {
    int M;
    int N[3], K[3], npad[3], kb[3];
    uint32_t off_w[3], off_act[2];      // byte offsets in dynamic smem (A tile at 0)
    const float* bias[3];
    float* x; int ldx;                  // residual stream (read as residual, written in place)
    const float* ln_gamma; const float* ln_beta; float ln_eps;
    __nv_bfloat16* ln_out; int ln_ld;   // may be null (last block)
    int num_tiles;
};

struct TailMaps {
    CUtensorMap a;
    CUtensorMap w[3];
};

// 8 consecutive bf16 of row r, columns [col, col+8) of a K-major 128B-swizzled operand made of 64-column k-blocks
__device__ __forceinline__ uint32_t act_addr(uint32_t base, int r, int col) {
    const int kbk = col >> 6, chunk = (col & 63) >> 3;
    return base + static_cast<uint32_t>(kbk * kATileBytes + r * 128 + ((chunk ^ (r & 7)) << 4));
}

template <int ACT>
__global__ void __launch_bounds__(kTailThreads, 1)
mlp_tail_kernel(const __grid_constant__ TailMaps maps, const TailArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();

    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0u) __trap();
    const uint32_t sA = base;
    const uint32_t bar_w = smem_u32(&bars[0]);          // weights landed (once)
    const uint32_t bar_a = smem_u32(&bars[1]);          // A tile landed (per tile)
    const uint32_t bar_d = smem_u32(&bars[2]);          // [3] MMA of layer l complete (per tile)
    const uint32_t bar_act = smem_u32(&bars[5]);        // [2] activations of layer l are in smem (per tile, 4 warps)

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_a, 1);
        for (int l = 0; l < 3; ++l) mbar_init(bar_d + 8 * l, 1);
        for (int l = 0; l < 2; ++l) mbar_init(bar_act + 8 * l, kTailEpiWarps);
        fence_mbar_init();
    }
    if (warp == kTailCtlWarp) {
        tmem_alloc(smem_u32(&tmem_base_s), kTailTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t col_d[3] = {0u, 128u, 192u};
    pdl_wait();

    if (warp == kTailCtlWarp) {
        // The whole warp runs the control loop (warp-uniform); TMA / MMA / commit are issued by one elected lane.
        const uint32_t a_bytes = static_cast<uint32_t>(p.kb[0]) * kATileBytes;
        int tile = blockIdx.x;
        if (elect_one()) {
            // ---- weights, once; first A tile ----
            uint32_t wbytes = 0;
            for (int l = 0; l < 3; ++l) wbytes += static_cast<uint32_t>(p.kb[l] * p.npad[l] * 128);
            mbar_arrive_expect_tx(bar_w, wbytes);
            for (int l = 0; l < 3; ++l)
                for (int kb = 0; kb < p.kb[l]; ++kb)
                    tma_load_2d(base + p.off_w[l] + kb * p.npad[l] * 128, &maps.w[l], bar_w, kb * kTK, 0);
            if (tile < p.num_tiles) {
                mbar_arrive_expect_tx(bar_a, a_bytes);
                for (int kb = 0; kb < p.kb[0]; ++kb) tma_load_2d(sA + kb * kATileBytes, &maps.a, bar_a, kb * kTK, tile * kTM);
            }
        }
        __syncwarp();
        mbar_wait(bar_w, 0);
        uint32_t ph = 0;
        for (; tile < p.num_tiles; tile += gridDim.x, ph ^= 1u) {
            for (int l = 0; l < 3; ++l) {
                uint32_t a_base;
                if (l == 0) { mbar_wait(bar_a, ph); a_base = sA; }
                else { mbar_wait(bar_act + 8 * (l - 1), ph); a_base = base + p.off_act[l - 1]; }
                tc_fence_after();
                if (elect_one()) {
                    if (l == 1) {
                        // MMA 0 is complete (its epilogue has run): the A buffer is free for the next tile
                        const int nt = tile + gridDim.x;
                        if (nt < p.num_tiles) {
                            mbar_arrive_expect_tx(bar_a, a_bytes);
                            for (int kb = 0; kb < p.kb[0]; ++kb) tma_load_2d(sA + kb * kATileBytes, &maps.a, bar_a, kb * kTK, nt * kTM);
                        }
                    }
                    const uint32_t idesc = umma_idesc_bf16_f32(kTM, p.npad[l]);
                    const uint32_t d_tmem = tmem_base + col_d[l];
                    for (int kb = 0; kb < p.kb[l]; ++kb) {
                        const uint64_t da = umma_desc_sw128_kmajor(a_base + kb * kATileBytes);
                        const uint64_t db = umma_desc_sw128_kmajor(base + p.off_w[l] + kb * p.npad[l] * 128);
                        int ksteps = kTK / 16;
                        if (kb == p.kb[l] - 1) ksteps = (p.K[l] - kb * kTK + 15) / 16;
                        for (int k = 0; k < ksteps; ++k) umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(bar_d + 8 * l);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------ epilogues ----------------------------------
        const int quad = warp & 3, half = warp >> 2;    // TMEM lane quadrant; which chunks of layers 0 / 1
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int r = quad * 32 + lane;                 // row within the tile
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ph ^= 1u) {
            const int row = tile * kTM + r;
            // the residual row of the last layer is fetched NOW (the warps that will need it), so that its global-load latency
            // hides behind the three MMA -> epilogue stages of the tile instead of sitting in the last one
            float4 xpre[8];
            if (half == 0 && row < p.M) {
                const float* xrow_pre = p.x + static_cast<size_t>(row) * p.ldx;
#pragma unroll
                for (int g = 0; g < 8; ++g) xpre[g] = (4 * g < p.ldx) ? *reinterpret_cast<const float4*>(xrow_pre + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // ---- layers 0 and 1: activations -> shared memory (A operand of the next layer) ----
            for (int l = 0; l < 2; ++l) {
                mbar_wait(bar_d + 8 * l, ph);
                tc_fence_after();
                const uint32_t act_base = base + p.off_act[l];
                const int N = p.N[l], npad = p.npad[l];
                const float* bias = p.bias[l];
                for (int c0 = 32 * half; c0 < npad; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + lane_off + col_d[l] + c0, v);
                    tmem_ld_wait();
                    if (c0 + 32 <= N) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * g));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * g + 4));
                            const uint32_t o0 = bias_act_bf16x2<ACT>(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]), b0.x, b0.y);
                            const uint32_t o1 = bias_act_bf16x2<ACT>(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]), b0.z, b0.w);
                            const uint32_t o2 = bias_act_bf16x2<ACT>(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]), b1.x, b1.y);
                            const uint32_t o3 = bias_act_bf16x2<ACT>(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]), b1.z, b1.w);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(act_addr(act_base, r, c0 + 8 * g)), "r"(o0),
                                         "r"(o1), "r"(o2), "r"(o3) : "memory");
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int c = c0 + 8 * g;
                            if (c < npad) {
                                uint32_t o[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int n = c + 2 * j;
                                    const float y0 = n < N ? apply_act<ACT, false>(__uint_as_float(v[8 * g + 2 * j]) + __ldg(bias + n)) : 0.f;
                                    const float y1 = n + 1 < N ? apply_act<ACT, false>(__uint_as_float(v[8 * g + 2 * j + 1]) + __ldg(bias + n + 1)) : 0.f;
                                    o[j] = pack_bf16x2(y0, y1);
                                }
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(act_addr(act_base, r, c)), "r"(o[0]),
                                             "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                            }
                        }
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_act + 8 * l);
            }
            if (half != 0) continue;        // the 32-wide last layer needs one warp per quadrant only
            // ---- layer 2: + residual -> x (f32), LayerNorm -> y (bf16) ----
            mbar_wait(bar_d + 16, ph);
            tc_fence_after();
            {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_off + col_d[2], v);
                tmem_ld_wait();
                if (row < p.M) {
                    float* xrow = p.x + static_cast<size_t>(row) * p.ldx;
                    float xr[32];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int n = 4 * g;
                        float h4[4] = {0.f, 0.f, 0.f, 0.f};
                        if (n < p.ldx) {
                            const float4 r4 = xpre[g];
                            const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                h4[j] = (n + j < p.N[2]) ? apply_act<ACT, false>(__uint_as_float(v[n + j]) + __ldg(p.bias[2] + n + j)) + rr[j] : 0.f;
                            *reinterpret_cast<float4*>(xrow + n) = make_float4(h4[0], h4[1], h4[2], h4[3]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) xr[n + j] = h4[j];
                    }
                    if (p.ln_out != nullptr) {
                        const float inv_n = 1.f / static_cast<float>(p.N[2]);
                        float sum = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum += xr[i];
                        const float mean = sum * inv_n;
                        float sq = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) { const float dv = (i < p.N[2]) ? xr[i] - mean : 0.f; sq = fmaf(dv, dv, sq); }
                        const float rstd = rsqrtf(sq * inv_n + p.ln_eps);
                        __nv_bfloat16* yrow = p.ln_out + static_cast<size_t>(row) * p.ln_ld;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (8 * g < p.ln_ld) {
                                float o8[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int i = 8 * g + j;
                                    o8[j] = (i < p.N[2]) ? (xr[i] - mean) * rstd * __ldg(p.ln_gamma + i) + __ldg(p.ln_beta + i) : 0.f;
                                }
                                uint4 w;
                                w.x = pack_bf16x2(o8[0], o8[1]); w.y = pack_bf16x2(o8[2], o8[3]);
                                w.z = pack_bf16x2(o8[4], o8[5]); w.w = pack_bf16x2(o8[6], o8[7]);
                                *reinterpret_cast<uint4*>(yrow + 8 * g) = w;
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTailCtlWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTailTmemCols);
    }
}

}  // namespace

bool mlp_tail_supported(const int N[3], const int K[3]) {
    // widths that fit the fixed TMEM / shared-memory layout of the kernel
    if (K[0] > 256 || N[0] > 128 || N[1] > 64 || N[2] > 32) return false;
    if (K[1] != N[0] || K[2] != N[1]) return false;
    for (int l = 0; l < 3; ++l) if (N[l] < 8 || K[l] < 8 || (K[l] % 8)) return false;
    return true;
}

int mlp_tail_make_plan(MlpTailPlan* plan, const MlpTailDesc& d, int num_sms) {
    if (!mlp_tail_supported(d.N, d.K)) return -30;
    if ((d.lda % 8) || (reinterpret_cast<uintptr_t>(d.A) & 15)) return -31;
    if ((d.ldx % 4) || d.ldx < (d.N[2] + 3) / 4 * 4 || d.ldx > 32) return -32;
    if (d.ln_out && ((d.ln_ld % 8) || d.ln_ld > 32 || d.ln_ld < (d.N[2] + 7) / 8 * 8)) return -33;
    plan->desc = d;
    uint32_t off = 0;
    for (int l = 0; l < 3; ++l) {
        plan->npad[l] = (d.N[l] + 15) / 16 * 16;
        plan->kb[l] = (d.K[l] + kTK - 1) / kTK;
    }
    off += static_cast<uint32_t>(plan->kb[0]) * kATileBytes;                 // A tile
    for (int l = 0; l < 3; ++l) {
        plan->off_w[l] = off;
        off += static_cast<uint32_t>(plan->kb[l] * plan->npad[l] * 128);
        off = (off + 1023u) & ~1023u;
    }
    for (int l = 0; l < 2; ++l) {
        plan->off_act[l] = off;
        off += static_cast<uint32_t>(plan->kb[l + 1]) * kATileBytes;
    }
    plan->smem_bytes = off;
    if (off > 224u * 1024u) return -34;
    plan->num_tiles = (d.M + kTM - 1) / kTM;
    plan->grid = plan->num_tiles < num_sms ? plan->num_tiles : num_sms;
    int r = make_tmap_bf16_2d(&plan->tmA, d.A, d.M, d.K[0], d.lda, kTM);
    if (r) return r;
    for (int l = 0; l < 3; ++l) {
        r = make_tmap_bf16_2d(&plan->tmW[l], d.W[l], d.N[l], d.K[l], d.ldw[l], plan->npad[l]);
        if (r) return r;
    }
    return 0;
}

template <int ACT>
static cudaError_t tail_launch_variant(const MlpTailPlan& plan, const TailMaps& maps, const TailArgs& a, cudaStream_t stream) {
    auto kern = mlp_tail_kernel<ACT>;
    {
        cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(kern), 224 * 1024);
        if (e != cudaSuccess) return e;
    }
    return launch_kernel(kern, dim3(plan.grid), dim3(kTailThreads), plan.smem_bytes, stream, 1, maps, a);
}

cudaError_t mlp_tail_launch(const MlpTailPlan& plan, cudaStream_t stream) {
    const MlpTailDesc& d = plan.desc;
    TailMaps maps;
    maps.a = plan.tmA;
    for (int l = 0; l < 3; ++l) maps.w[l] = plan.tmW[l];
    TailArgs a;
    a.M = d.M;
    for (int l = 0; l < 3; ++l) {
        a.N[l] = d.N[l]; a.K[l] = d.K[l]; a.npad[l] = plan.npad[l]; a.kb[l] = plan.kb[l];
        a.off_w[l] = plan.off_w[l]; a.bias[l] = d.bias[l];
    }
    a.off_act[0] = plan.off_act[0]; a.off_act[1] = plan.off_act[1];
    a.x = d.x; a.ldx = d.ldx;
    a.ln_gamma = d.ln_gamma; a.ln_beta = d.ln_beta; a.ln_eps = d.ln_eps;
    a.ln_out = static_cast<__nv_bfloat16*>(d.ln_out); a.ln_ld = d.ln_ld;
    a.num_tiles = plan.num_tiles;
    switch (d.act) {
        case ACT_NONE: return tail_launch_variant<ACT_NONE>(plan, maps, a, stream);
        case ACT_MISH: return tail_launch_variant<ACT_MISH>(plan, maps, a, stream);
        case ACT_GELU: return tail_launch_variant<ACT_GELU>(plan, maps, a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vitdet
