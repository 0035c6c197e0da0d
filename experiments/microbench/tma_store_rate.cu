// Micro-benchmark: write bandwidth of the GEMM epilogues' store path as a function of the TMA box shape.
// 148 CTAs x 16 warps; every warp fills a staging tile in shared memory (64B- or 128B-swizzled rows, contents irrelevant)
// and TMA-stores 32-row boxes of a [M, N] bf16 matrix, tile order as in gemm_tc.cu (CTA = 128 rows x 256 columns, warp =
// (row quadrant, column chunk)).  Variants: 32 x 32 boxes (64-byte rows, the r02 product), 32 x 64 boxes (128-byte rows),
// 32 x 128 boxes (256-byte rows = two 128B-swizzle atoms are not expressible; skipped), and the 32 x 32 box with two staging
// tiles per warp (wait_group.read 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_rate tma_store_rate.cu -lcuda && ./tma_store_rate
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, CUtensorMapSwizzle sw) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, 32}, estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); }
    return m;
}

template <int BOXC, int NBUF>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ CUtensorMap tm, int M, int N, int work) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quad = warp & 3, sub = warp >> 2;
    constexpr int TILE = 32 * BOXC * 2;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem) + warp * NBUF * TILE;
    const int n_tiles = N / 256, num_tiles = (M / 128) * n_tiles;
    int buf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * 256;
        for (int c0 = sub * BOXC; c0 < 256; c0 += 4 * BOXC) {
            // stand-in for the epilogue arithmetic
            uint32_t o[BOXC / 2];
#pragma unroll
            for (int i = 0; i < BOXC / 2; ++i) o[i] = tile * 131 + c0 + i + lane;
            for (int w = 0; w < work; ++w)
#pragma unroll
                for (int i = 0; i < BOXC / 2; ++i) o[i] = o[i] * 1664525u + 1013904223u;
            if (lane == 0) {
                if (NBUF == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            __syncwarp();
            const uint32_t st = s0 + buf * TILE;
#pragma unroll
            for (int g = 0; g < BOXC / 8; ++g) {
                const uint32_t dst = st + lane * (BOXC * 2) + (((BOXC == 32) ? (g ^ ((lane >> 1) & 3)) : (g ^ (lane & 7))) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[4 * g]), "r"(o[4 * g + 1]), "r"(o[4 * g + 2]), "r"(o[4 * g + 3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(st), "r"(n0 + c0), "r"(m0 + quad * 32) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (NBUF == 2) buf ^= 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int BOXC, int NBUF>
void run(const char* name, EncodeFn fn, void* buf, int M, int N, int work) {
    CUtensorMap tm = make_map(fn, buf, M, N, BOXC, BOXC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    const int smem = 16 * NBUF * 32 * BOXC * 2 + 1024;
    cudaFuncSetAttribute(k<BOXC, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<BOXC, NBUF><<<148, 512, smem>>>(tm, M, N, work);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<BOXC, NBUF><<<148, 512, smem>>>(tm, M, N, work);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%-44s work %3d: %.3f ms  %.2f TB/s\n", name, work, best, double(M) * N * 2 / (best * 1e-3) / 1e12);
}

int main() {
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeFn fn = (EncodeFn)sym;
    const int M = 82944, N = 3584;
    void* buf; cudaMalloc(&buf, size_t(M) * N * 2);
    for (int work : {0, 8, 24}) {
        run<32, 1>("32 x 32 box (64 B rows), 1 staging tile", fn, buf, M, N, work);
        run<32, 2>("32 x 32 box (64 B rows), 2 staging tiles", fn, buf, M, N, work);
        run<64, 1>("32 x 64 box (128 B rows), 1 staging tile", fn, buf, M, N, work);
        run<64, 2>("32 x 64 box (128 B rows), 2 staging tiles", fn, buf, M, N, work);
    }
    return 0;
}
