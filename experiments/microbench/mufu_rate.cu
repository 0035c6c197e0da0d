// Micro-benchmark: how fast can ONE warp (and two warps of one sub-partition) run the attention kernel's exponential
// loop (FFMA2 scale-and-shift, MUFU.EX2, FADD2 row sum, F2FP pack) on sm_100a, as a function of the distance between a
// MUFU and the instructions that consume its result.  Prints clocks per MUFU per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

// MODE 0: product form (consumers right behind the MUFU pair in the source)
// MODE 1: consumers lag the MUFUs by LAG groups of 4 elements in the source
// MODE 2: MUFU + plain add only (no pack)
// MODE 3: no row sum (ones-column form): MUFU + F2FP
template <int MODE, int LAG>
__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ in, uint32_t* out, long long* clk, int iters, float scale, float negm) {
    __shared__ float4 s[1024 + 512];     // 32 float4 per thread would be 128 KB; share one row set per warp instead
    const int tid = threadIdx.x;
    for (int i = tid; i < 1536; i += blockDim.x) s[i] = reinterpret_cast<const float4*>(in)[i % 1024];
    __syncthreads();
    uint32_t acc = 0;
    float l = 0.f;
    const uint64_t sc2 = f2_pack(scale, scale), nm2 = f2_pack(negm, negm);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float v[128];
        const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(&s[0])) + (((tid & 31) + (it & 7) * 32) << 4);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float4 q;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(sbase + c * 512));
            v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
        }
        uint32_t pk[64];
        uint64_t sum2[2] = {0ull, 0ull};
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                float a0, a1, a2, a3;
                f2_unpack(f2_fma(f2_pack(v[i], v[i + 1]), sc2, nm2), a0, a1);
                f2_unpack(f2_fma(f2_pack(v[i + 2], v[i + 3]), sc2, nm2), a2, a3);
                const float e0 = ex2f(a0), e1 = ex2f(a1), e2 = ex2f(a2), e3 = ex2f(a3);
                sum2[0] = f2_add(sum2[0], f2_pack(e0, e1));
                sum2[1] = f2_add(sum2[1], f2_pack(e2, e3));
                pk[i / 2] = pack_bf16x2(e0, e1);
                pk[i / 2 + 1] = pack_bf16x2(e2, e3);
            }
        } else if (MODE == 1) {
            float e[32 + 1][4];
#pragma unroll
            for (int g = 0; g < 32 + LAG; ++g) {
                if (g < 32) {
                    float a0, a1, a2, a3;
                    f2_unpack(f2_fma(f2_pack(v[4 * g], v[4 * g + 1]), sc2, nm2), a0, a1);
                    f2_unpack(f2_fma(f2_pack(v[4 * g + 2], v[4 * g + 3]), sc2, nm2), a2, a3);
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[g][0]) : "f"(a0));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[g][1]) : "f"(a1));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[g][2]) : "f"(a2));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[g][3]) : "f"(a3));
                }
                if (g >= LAG) {
                    const int h = g - LAG;
                    uint64_t r0, r1;
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r0) : "l"(sum2[0]), "l"(f2_pack(e[h][0], e[h][1])));
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r1) : "l"(sum2[1]), "l"(f2_pack(e[h][2], e[h][3])));
                    sum2[0] = r0; sum2[1] = r1;
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[2 * h]) : "f"(e[h][1]), "f"(e[h][0]));
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[2 * h + 1]) : "f"(e[h][3]), "f"(e[h][2]));
                }
            }
        } else if (MODE == 2) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                float a0, a1, a2, a3;
                f2_unpack(f2_fma(f2_pack(v[i], v[i + 1]), sc2, nm2), a0, a1);
                f2_unpack(f2_fma(f2_pack(v[i + 2], v[i + 3]), sc2, nm2), a2, a3);
                s4[0] += ex2f(a0); s4[1] += ex2f(a1); s4[2] += ex2f(a2); s4[3] += ex2f(a3);
            }
            sum2[0] = f2_pack(s4[0], s4[1]); sum2[1] = f2_pack(s4[2], s4[3]);
#pragma unroll
            for (int i = 0; i < 64; ++i) pk[i] = 0;
        } else {
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                float a0, a1, a2, a3;
                f2_unpack(f2_fma(f2_pack(v[i], v[i + 1]), sc2, nm2), a0, a1);
                f2_unpack(f2_fma(f2_pack(v[i + 2], v[i + 3]), sc2, nm2), a2, a3);
                const float e0 = ex2f(a0), e1 = ex2f(a1), e2 = ex2f(a2), e3 = ex2f(a3);
                pk[i / 2] = pack_bf16x2(e0, e1);
                pk[i / 2 + 1] = pack_bf16x2(e2, e3);
            }
        }
        float s0, s1, s2, s3;
        f2_unpack(sum2[0], s0, s1); f2_unpack(sum2[1], s2, s3);
        l += (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int i = 0; i < 64; ++i) acc ^= pk[i];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc ^ __float_as_uint(l);
    if ((tid & 31) == 0) clk[blockIdx.x * (blockDim.x / 32) + tid / 32] = t1 - t0;
}

template <int MODE, int LAG>
void run(const char* name, int warps_per_smsp, const float* in, uint32_t* out, long long* clk) {
    const int threads = 128 * warps_per_smsp, blocks = 148, iters = 2000;
    k<MODE, LAG><<<blocks, threads>>>(in, out, clk, 10, 0.2f, -1.f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, LAG><<<blocks, threads>>>(in, out, clk, iters, 0.2f, -1.f);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const int nw = blocks * threads / 32;
    std::vector<long long> h(nw);
    cudaMemcpy(h.data(), clk, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    printf("%-34s warps/SMSP %d: %6.2f clk per MUFU per warp (median; min %.2f max %.2f)  %.3f ms  %s\n", name, warps_per_smsp,
           double(h[nw / 2]) / (iters * 128.0), double(h[0]) / (iters * 128.0), double(h[nw - 1]) / (iters * 128.0), ms, cudaGetErrorString(err));
}

int main() {
    float* in; uint32_t* out; long long* clk;
    cudaMalloc(&in, 1024 * 16); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&clk, 148 * 8 * 8);
    std::vector<float> h(4096);
    for (int i = 0; i < 4096; ++i) h[i] = -float((i * 37) % 101) * 0.05f;
    cudaMemcpy(in, h.data(), 4096 * 4, cudaMemcpyHostToDevice);
    for (int w = 1; w <= 2; ++w) {
        run<0, 0>("product form", w, in, out, clk);
        run<1, 1>("lag 1 group (4 MUFU)", w, in, out, clk);
        run<1, 2>("lag 2 groups (8 MUFU)", w, in, out, clk);
        run<1, 4>("lag 4 groups (16 MUFU)", w, in, out, clk);
        run<2, 0>("MUFU + FADD only", w, in, out, clk);
        run<3, 0>("no row sum (MUFU + F2FP)", w, in, out, clk);
    }
    return 0;
}
