// f32 CUDA-core Dense layer for the fp32 parity mode: same contract as the tensor-core kernel
// (gemm_tc.cu) but every product and sum is IEEE f32, as in the reference (which never sets a
// mixed-precision policy; all Keras layers of det.py:239-495 run in float32).
//
// 128 x 128 x 8 register-tiled kernel, 256 threads, 8 x 8 outputs per thread split as 2 x 2
// quadrants of 4 x 4 so that shared-memory reads are conflict-free float4 broadcasts.
#include "common.cuh"
#include "kernels.h"

namespace vitdet {

namespace {

constexpr int SBM = 128, SBN = 128, SBK = 8, SPAD = 4;

struct SimtArgs {
    const float* A; int lda;
    const float* W; int ldw;
    int M, N, K4;               // K4 = round_up(K, 4); pads are zero in both operands
    const float* bias;
    const float* pos; int pos_period;
    const float* resid; int ldr;
    float* out; int ldc;
    int n_store;
};

template <int ACT>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtArgs p) {
    __shared__ __align__(16) float As[2][SBK][SBM + SPAD];
    __shared__ __align__(16) float Bs[2][SBK][SBN + SPAD];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;

    // global -> register staging: one float4 of A and one of W per thread per k-block
    const int lrow = t >> 1, lk = (t & 1) * 4;
    const bool a_ok = (m0 + lrow) < p.M;
    const bool b_ok = (n0 + lrow) < p.N;
    const float* ap = p.A + static_cast<size_t>(a_ok ? m0 + lrow : 0) * p.lda + lk;
    const float* bp = p.W + static_cast<size_t>(b_ok ? n0 + lrow : 0) * p.ldw + lk;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nkb = (p.K4 + SBK - 1) / SBK;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
    auto gload = [&](int kb) {
        const int k = kb * SBK + lk;
        ra = (a_ok && k < p.K4) ? *reinterpret_cast<const float4*>(ap + kb * SBK) : make_float4(0.f, 0.f, 0.f, 0.f);
        rb = (b_ok && k < p.K4) ? __ldg(reinterpret_cast<const float4*>(bp + kb * SBK)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sstore = [&](int buf) {
        As[buf][lk + 0][lrow] = ra.x; As[buf][lk + 1][lrow] = ra.y;
        As[buf][lk + 2][lrow] = ra.z; As[buf][lk + 3][lrow] = ra.w;
        Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y;
        Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
    };

    gload(0);
    sstore(0);
    __syncthreads();
    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nkb) gload(kb + 1);
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kb + 1 < nkb) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: act(acc + bias + pos) + resid, float4 stores; columns [N, n_store) written as zero
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= p.M) continue;
        const float pos_v = p.pos ? __ldg(p.pos + (row % p.pos_period)) : 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * 64 + tx * 4;
            if (n >= p.n_store) continue;
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.resid) r4 = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(row) * p.ldr + n);
            const float r[4] = {r4.x, r4.y, r4.z, r4.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float bv = (p.bias && n + j < p.N) ? __ldg(p.bias + n + j) : 0.f;
                const float y = apply_act<ACT, true>(acc[i][h * 4 + j] + bv + pos_v) + r[j];
                o[j] = (n + j < p.N) ? y : 0.f;
            }
            *reinterpret_cast<float4*>(p.out + static_cast<size_t>(row) * p.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace

cudaError_t simt_gemm_launch(const GemmDesc& d, cudaStream_t stream) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0 || !d.out_f32) return cudaErrorInvalidValue;
    const int K4 = (d.K + 3) / 4 * 4;
    if ((d.lda % 4) || (d.ldw % 4) || d.lda < K4 || d.ldw < K4) return cudaErrorInvalidValue;
    const int n_store = (d.N + 3) / 4 * 4;
    if ((d.ldc % 4) || d.ldc < n_store) return cudaErrorInvalidValue;
    if (d.resid && ((d.ldr % 4) || d.ldr < n_store)) return cudaErrorInvalidValue;
    SimtArgs a;
    a.A = static_cast<const float*>(d.A); a.lda = d.lda;
    a.W = static_cast<const float*>(d.W); a.ldw = d.ldw;
    a.M = d.M; a.N = d.N; a.K4 = K4;
    a.bias = d.bias; a.pos = d.pos; a.pos_period = d.pos_period > 0 ? d.pos_period : 1;
    a.resid = d.resid; a.ldr = d.ldr;
    a.out = static_cast<float*>(d.out); a.ldc = d.ldc; a.n_store = n_store;
    dim3 grid((d.N + SBN - 1) / SBN, (d.M + SBM - 1) / SBM);
    switch (d.act) {
        case ACT_NONE: gemm_simt_kernel<ACT_NONE><<<grid, 256, 0, stream>>>(a); break;
        case ACT_MISH: gemm_simt_kernel<ACT_MISH><<<grid, 256, 0, stream>>>(a); break;
        case ACT_GELU: gemm_simt_kernel<ACT_GELU><<<grid, 256, 0, stream>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace vitdet
