// mma_probe.cu — development probe: what paces a chain of small tcgen05.mma (kind::f16, SS / TS) on sm_100a?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tetris_mcts_b200/csrc -o scripts/probe/mma_probe scripts/probe/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include "valuenet_tc.cuh"
using namespace b200;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int M, int N, int ALTN, int NACC, int TS, int SHIFT, int STEP>
__global__ void __launch_bounds__(128, 1) k_probe(int outer, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
        const uint64_t a0 = umma_desc(a_addr, 152 * 16, 128), b0 = umma_desc(b_addr, 256 * 16, 128);
        long long t0 = clock64();
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int r = 0; r < 72; ++r) {
                const uint32_t d = tm + (uint32_t)(r % NACC) * 64;
                const uint32_t idesc = (r & 1) ? umma_idesc_f16(M, ALTN ? ALTN : N) : umma_idesc_f16(M, N);
                const uint64_t ad = a0 + (uint64_t)(SHIFT + (r % 9) * STEP);
                if (TS) umma_f16_ts(d, tm + 256, b0 + (uint64_t)(r % 18) * 128, idesc, 1u);
                else umma_f16(d, ad, b0 + (uint64_t)(r % 18) * 128, idesc, 1u);
            }
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int M, int N, int ALTN, int NACC, int TS, int SHIFT, int STEP>
void run(int grid, long long *d) {
    auto kf = k_probe<M, N, ALTN, NACC, TS, SHIFT, STEP>;
    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    long long h[2] = {0, 0};
    const int outer = 10;
    for (int it = 0; it < 2; ++it) {
        kf<<<grid, 128, 160 * 1024>>>(outer, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    double nn = ALTN ? 0.5 * (N + ALTN) : N;
    printf("%4d %4d %5d %4d %2d %4d %4d | %13.1f | %13.1f | %6.1f\n", M, N, SHIFT, NACC, TS, ALTN, STEP, (double)h[0] / (72 * outer),
           (double)h[1] / (72 * outer), (M > 128 ? M : 128) * nn / 256.0);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    for (int grid : {1, 148}) {
        printf("grid %d\n   M    N shift nacc ts altN step | issue clk/mma | total clk/mma | guide floor\n", grid);
        run<128, 64, 0, 1, 0, 0, 0>(grid, d);  run<128, 64, 0, 1, 0, 1, 0>(grid, d);  run<128, 64, 0, 2, 0, 0, 0>(grid, d);  run<128, 64, 0, 4, 0, 0, 0>(grid, d);
        run<128, 32, 0, 1, 0, 0, 0>(grid, d);  run<128, 32, 0, 1, 0, 1, 0>(grid, d);  run<128, 32, 0, 4, 0, 0, 0>(grid, d);
        run<128, 64, 32, 1, 0, 0, 0>(grid, d); run<128, 64, 32, 2, 0, 0, 0>(grid, d); run<128, 64, 32, 4, 0, 0, 0>(grid, d);
        run<128, 64, 32, 1, 0, 0, 1>(grid, d); run<128, 64, 32, 1, 0, 0, 8>(grid, d); run<128, 64, 32, 4, 0, 0, 1>(grid, d);
        run<128, 128, 0, 1, 0, 0, 0>(grid, d); run<128, 128, 0, 1, 0, 1, 0>(grid, d); run<128, 256, 0, 1, 0, 0, 0>(grid, d); run<128, 256, 0, 1, 0, 1, 0>(grid, d);
        run<128, 96, 0, 1, 0, 0, 0>(grid, d);  run<128, 96, 0, 2, 0, 0, 1>(grid, d);
        run<64, 64, 0, 1, 0, 0, 0>(grid, d);   run<64, 128, 0, 1, 0, 0, 0>(grid, d);  run<64, 256, 0, 1, 0, 0, 0>(grid, d);
        run<128, 64, 0, 1, 1, 0, 0>(grid, d);  run<128, 64, 0, 4, 1, 0, 0>(grid, d);  run<128, 32, 0, 1, 1, 0, 0>(grid, d);  run<128, 128, 0, 1, 1, 0, 0>(grid, d);
        run<128, 256, 0, 1, 1, 0, 0>(grid, d); run<128, 64, 32, 1, 1, 0, 0>(grid, d);
    }
    return 0;
}
