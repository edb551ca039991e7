// mma_contend.cu — development probe: does worker traffic (TMEM loads, shared-memory stores, plain ALU) slow a queue of
// N=96 tcgen05.mma?  One issuer lane + 16 worker warps in the same CTA.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include "valuenet_tc.cuh"
using namespace b200;

template <int MODE, int COMMIT_EVERY>
__global__ void __launch_bounds__(544, 1) k_contend(int outer, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tptr;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 180 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); stop = 0; }
    if (threadIdx.x >= 512) tmem_alloc<512>(&tptr);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 16) {
        if (lane == 0) {
            const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
            const uint64_t a0 = umma_desc(a_addr, 144 * 16, 128), b0 = umma_desc(b_addr, 96 * 16, 128);
            long long t0 = clock64();
            for (int o = 0; o < outer; ++o) {
#pragma unroll
                for (int r = 0; r < 72; ++r) {
                    umma_f16(tm + (r / 18) * 128, a0 + (uint64_t)((r % 6) * 288 + (r % 3) * 8), b0 + (uint64_t)(r % 12) * 192, umma_idesc_f16(128, 96), 1u);
                    if (COMMIT_EVERY && (r % COMMIT_EVERY) == COMMIT_EVERY - 1) umma_commit(&bar[(r / COMMIT_EVERY) & 3]);
                }
            }
            umma_commit(&bar[7]);
            mbar_wait(&bar[7], 0);
            long long t1 = clock64();
            stop = 1;
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
    } else {
        const int q = warp & 3, cq = warp >> 2;
        float acc = 0.f;
        uint32_t it = 0;
        while (!stop) {
            if (MODE == 1) {                 // TMEM loads
                float v[8];
                tmem_ld8(tm + ((uint32_t)(q * 32) << 16) + cq * 8 + (it & 3) * 128, v);
                acc += v[0] + v[7];
            } else if (MODE == 2) {          // 16-byte shared stores, conflict free
                *reinterpret_cast<uint4 *>(smem + 128 * 1024 + ((it & 3) * 512 + threadIdx.x) * 16) = make_uint4(it, it, it, it);
            } else if (MODE == 3) {          // ALU only
#pragma unroll
                for (int k = 0; k < 16; ++k) acc = fmaf(acc, 1.0001f, 0.5f);
            } else if (MODE == 4) {          // shared loads
                acc += reinterpret_cast<volatile float *>(smem + 128 * 1024)[(it * 32 + threadIdx.x) & 8191];
            } else {
                __nanosleep(200);
            }
            ++it;
        }
        if (acc == 123.456f) out[1] = it;
    }
    __syncthreads();
    if (threadIdx.x >= 512) tmem_dealloc<512>(tm);
}

template <int MODE, int CE>
void run(const char *name, long long *d) {
    auto kf = k_contend<MODE, CE>;
    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
    long long h[2] = {0, 0};
    const int outer = 10;
    for (int it = 0; it < 2; ++it) {
        kf<<<148, 544, 180 * 1024>>>(outer, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s commit every %2d : %6.1f clk/mma\n", name, CE, (double)h[0] / (72 * outer));
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    run<0, 0>("workers asleep", d);
    run<0, 18>("workers asleep", d);
    run<1, 0>("workers: tcgen05.ld x8 in a loop", d);
    run<2, 0>("workers: st.shared.v4 in a loop", d);
    run<4, 0>("workers: ld.shared in a loop", d);
    run<3, 0>("workers: FMA chain", d);
    run<1, 18>("workers: tcgen05.ld x8 in a loop", d);
    return 0;
}
