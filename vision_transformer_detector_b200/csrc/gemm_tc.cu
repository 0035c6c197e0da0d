// Dense layer on the 5th-gen tensor cores:  out = epilogue( A[M,K] * W[N,K]^T )
//
// Replaces every keras.layers.Dense / EinsumDense of the reference's forward pass in bf16 mode:
// linear_projection (det.py:297-301), the q/k/v and attention_output projections inside
// MultiHeadAttention (det.py:364-369), MLP_i_j (det.py:388-398) and the mlp_head Dense chain
// (det.py:468-480), together with the element-wise ops that follow them in the reference graph
// (bias add, MishActivation det.py:119-129 / tfa GELU det.py:402, keras.layers.add det.py:305,
// 371, 408-412), which run here as the epilogue of the same kernel.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes of A (128 x 64 bf16) and W (BN x 64
//               bf16) into a ring of 128B-swizzled shared-memory stages, completion on mbarriers.
//   warp 1      allocates 512 TMEM columns; one elected lane issues tcgen05.mma (M=128, N=BN, K=16,
//               bf16 x bf16 -> f32 in TMEM) and tcgen05.commit to free stages / publish accumulators.
//   warps 2..5  epilogue: tcgen05.ld the f32 accumulator (thread = one row of the tile), bias /
//               position scalar / activation / residual, 16-byte global stores.
// Two TMEM accumulator stages (2 x 256 columns) let the MMA of tile i+1 overlap the epilogue of
// tile i.  BN and the ring depth are run-time values (N lives in the instruction descriptor), so one
// binary serves every layer width of the model.
#include "common.cuh"
#include "kernels.h"

namespace vitdet {

namespace {

constexpr int kBM = 128;              // UMMA M (cta_group::1)
constexpr int kBK = 64;               // one 128-byte swizzle row of bf16
constexpr int kUK = 16;               // UMMA K for 16-bit inputs
constexpr int kMaxBN = 256;
constexpr int kMaxStages = 8;
constexpr int kStageBytesA = kBM * kBK * 2;   // 16 KiB
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kRingBytes = 200 * 1024;            // budget of the A/B stage ring
constexpr int kMaxDynSmem = kRingBytes + 1024;    // + alignment slack; static smem (~2.2 KB) must also fit in 227 KB

struct TcGemmArgs {
    int M, N, K;
    int block_n;
    int num_stages;
    int n_tiles;
    int num_tiles;
    int num_kb;
    const float* bias;
    const float* pos;
    int pos_period;
    const float* resid;
    int ldr;
    void* out;
    int ldc;
    int n_store;   // round_up(N, store vector): columns [N, n_store) are written as zero
};

template <int ACT, bool OUT_F32>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcGemmArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[2][kMaxBN];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes_b = static_cast<uint32_t>(p.block_n) * (kBK * 2);
    const uint32_t sA0 = smem_base;
    const uint32_t sB0 = smem_base + static_cast<uint32_t>(p.num_stages) * kStageBytesA;

    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
    const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
    const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, kEpiThreads);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&tmem_base_s), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = kStageBytesA + stage_bytes_b;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * kBM;
                const int n0 = (tile % p.n_tiles) * p.block_n;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    mbar_arrive_expect_tx(bar_full + 8 * stage, tx_bytes);
                    tma_load_2d(sA0 + stage * kStageBytesA, &tmA, bar_full + 8 * stage, kb * kBK, m0);
                    tma_load_2d(sB0 + stage * stage_bytes_b, &tmB, bar_full + 8 * stage, kb * kBK, n0);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_f32(kBM, p.block_n);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kMaxBN);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128_kmajor(sA0 + stage * kStageBytesA);
                    const uint64_t db = umma_desc_sw128_kmajor(sB0 + stage * stage_bytes_b);
                    int ksteps = kBK / kUK;
                    if (kb == p.num_kb - 1) ksteps = (p.K - kb * kBK + kUK - 1) / kUK;
                    for (int k = 0; k < ksteps; ++k) {
                        // +32 B per K step inside the 128 B swizzle row: +2 in the (addr >> 4) field
                        umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
    } else {
        // ------------------------------ epilogue ----------------------------------
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int et = threadIdx.x - 64;              // 0..127 within the epilogue group
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (tile / p.n_tiles) * kBM;
            const int n0 = (tile % p.n_tiles) * p.block_n;

            // Stage this tile's bias slice in shared memory (double-buffered by tile parity; the
            // single named barrier per tile also orders reuse of the other buffer).
            float* bs = bias_s[it & 1];
            for (int c = et; c < p.block_n; c += kEpiThreads) {
                const int n = n0 + c;
                bs[c] = (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");

            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();

            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < p.M;
            float pos_v = 0.f;
            if (p.pos != nullptr && row_ok) pos_v = __ldg(p.pos + (row % p.pos_period));
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                   static_cast<uint32_t>(acc * kMaxBN);

            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c0, v);
                tmem_ld_wait();
                if (row_ok) {
                    if (OUT_F32) {
                        float* orow = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldc;
                        const float* rrow = p.resid ? p.resid + static_cast<size_t>(row) * p.ldr : nullptr;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const int c = c0 + 4 * g;
                            const int n = n0 + c;
                            if (c < p.block_n && n < p.n_store) {
                                const float4 b4 = *reinterpret_cast<const float4*>(bs + c);
                                float x[4] = {__uint_as_float(v[4 * g + 0]) + b4.x,
                                              __uint_as_float(v[4 * g + 1]) + b4.y,
                                              __uint_as_float(v[4 * g + 2]) + b4.z,
                                              __uint_as_float(v[4 * g + 3]) + b4.w};
                                float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (rrow) r4 = *reinterpret_cast<const float4*>(rrow + n);
                                const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float y = apply_act<ACT, false>(x[j] + pos_v) + r[j];
                                    x[j] = (n + j < p.N) ? y : 0.f;
                                }
                                *reinterpret_cast<float4*>(orow + n) = make_float4(x[0], x[1], x[2], x[3]);
                            }
                        }
                    } else {
                        __nv_bfloat16* orow =
                            reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldc;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int c = c0 + 8 * g;
                            const int n = n0 + c;
                            if (c < p.block_n && n < p.n_store) {
                                const float4 b0 = *reinterpret_cast<const float4*>(bs + c);
                                const float4 b1 = *reinterpret_cast<const float4*>(bs + c + 4);
                                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                                float x[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float y = apply_act<ACT, false>(__uint_as_float(v[8 * g + j]) + b[j] + pos_v);
                                    x[j] = (n + j < p.N) ? y : 0.f;
                                }
                                uint4 o;
                                o.x = pack_bf16x2(x[0], x[1]);
                                o.y = pack_bf16x2(x[2], x[3]);
                                o.z = pack_bf16x2(x[4], x[5]);
                                o.w = pack_bf16x2(x[6], x[7]);
                                *reinterpret_cast<uint4*>(orow + n) = o;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * acc);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

template <int ACT, bool OUT_F32>
cudaError_t launch_variant(const TcGemmPlan& plan, const TcGemmArgs& a, cudaStream_t stream) {
    auto kern = gemm_tc_kernel<ACT, OUT_F32>;
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    kern<<<plan.grid, kThreads, plan.smem_bytes, stream>>>(plan.tmA, plan.tmB, a);
    return cudaGetLastError();
}

}  // namespace

int choose_block_n(int N) {
    if (N <= kMaxBN) return (N + 15) / 16 * 16;
    // Largest multiple of 16 whose padded width wastes <= 3 %; otherwise the least wasteful >= 64.
    int best = kMaxBN;
    double best_waste = 1e9;
    for (int bn = kMaxBN; bn >= 64; bn -= 16) {
        const int tiles = (N + bn - 1) / bn;
        const double waste = static_cast<double>(tiles) * bn / N - 1.0;
        if (waste <= 0.03) return bn;
        if (waste < best_waste - 1e-9) { best_waste = waste; best = bn; }
    }
    return best;
}

// 2-D bf16 tensor map, row-major [rows, cols] with `ld` elements between rows, box = 64 x box_rows,
// 128B swizzle, out-of-bounds elements read as zero (this is what pads K, M and N tails).
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

int tc_gemm_make_plan(TcGemmPlan* plan, const GemmDesc& d, int num_sms) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) return -2;
    if ((d.lda % 8) || (d.ldw % 8)) return -3;                        // 16-byte row pitch for TMA
    if ((reinterpret_cast<uintptr_t>(d.A) & 15) || (reinterpret_cast<uintptr_t>(d.W) & 15) ||
        (reinterpret_cast<uintptr_t>(d.out) & 15))
        return -4;
    const int vec = d.out_f32 ? 4 : 8;
    if ((d.ldc % vec) || d.ldc < (d.N + vec - 1) / vec * vec) return -5;
    if (d.resid && (!d.out_f32 || (d.ldr % 4) || d.ldr < (d.N + 3) / 4 * 4)) return -6;
    int bn = d.block_n > 0 ? d.block_n : choose_block_n(d.N);
    if (bn % 16 || bn < 16 || bn > kMaxBN) return -7;

    plan->desc = d;
    plan->block_n = bn;
    const int stage = kStageBytesA + bn * kBK * 2;
    int stages = kRingBytes / stage;
    if (stages > kMaxStages) stages = kMaxStages;
    plan->num_stages = stages;
    plan->smem_bytes = static_cast<size_t>(stages) * stage + 1024;
    const int m_tiles = (d.M + kBM - 1) / kBM;
    const int n_tiles = (d.N + bn - 1) / bn;
    plan->n_tiles = n_tiles;
    plan->num_tiles = m_tiles * n_tiles;
    plan->grid = plan->num_tiles < num_sms ? plan->num_tiles : num_sms;
    int r = make_tmap_bf16_2d(&plan->tmA, d.A, d.M, d.K, d.lda, kBM);
    if (r) return r;
    r = make_tmap_bf16_2d(&plan->tmB, d.W, d.N, d.K, d.ldw, bn);
    return r;
}

cudaError_t tc_gemm_launch(const TcGemmPlan& plan, cudaStream_t stream) {
    const GemmDesc& d = plan.desc;
    TcGemmArgs a;
    a.M = d.M; a.N = d.N; a.K = d.K;
    a.block_n = plan.block_n;
    a.num_stages = plan.num_stages;
    a.n_tiles = plan.n_tiles;
    a.num_tiles = plan.num_tiles;
    a.num_kb = (d.K + kBK - 1) / kBK;
    a.bias = d.bias;
    a.pos = d.pos;
    a.pos_period = d.pos_period > 0 ? d.pos_period : 1;
    a.resid = d.resid;
    a.ldr = d.ldr;
    a.out = d.out;
    a.ldc = d.ldc;
    a.n_store = d.out_f32 ? (d.N + 3) / 4 * 4 : (d.N + 7) / 8 * 8;
    if (d.out_f32) {
        switch (d.act) {
            case ACT_NONE: return launch_variant<ACT_NONE, true>(plan, a, stream);
            case ACT_MISH: return launch_variant<ACT_MISH, true>(plan, a, stream);
            case ACT_GELU: return launch_variant<ACT_GELU, true>(plan, a, stream);
        }
    } else {
        switch (d.act) {
            case ACT_NONE: return launch_variant<ACT_NONE, false>(plan, a, stream);
            case ACT_MISH: return launch_variant<ACT_MISH, false>(plan, a, stream);
            case ACT_GELU: return launch_variant<ACT_GELU, false>(plan, a, stream);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace vitdet
