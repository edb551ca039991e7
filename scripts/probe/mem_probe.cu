// mem_probe.cu — development probe: what does a RANDOM access into a large HBM footprint cost on B200, and how many random
// 64-byte bursts per second does the memory system deliver?  (The tree walk of k_select_expand is two dependent random accesses per
// level into 82 GB of arenas; profiles/ncu_summary_r2.md has it at ~27 G bursts/s and 1.6-1.9 k clk per access.)
//   part 1  pointer chase, ONE thread: latency per hop against the footprint (L2 hit / DRAM / TLB reach)
//   part 2  independent random 16-byte loads from every SM: bursts per second against footprint, occupancy and loads in flight
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/mem_probe scripts/probe/mem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr unsigned long long LCG_A = 6364136223846793005ull, LCG_C = 1442695040888963407ull;

// line i (128 B) holds the index of the next line: a full-period LCG over the n_lines (a power of two) = one cycle through all lines
__global__ void k_init_chain(uint32_t *p, unsigned long long n_lines) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_lines; i += (unsigned long long)gridDim.x * blockDim.x)
        p[i * 32] = (uint32_t)((LCG_A * i + LCG_C) & (n_lines - 1));
}

__global__ void k_chase(const uint32_t *p, int warm, int hops, long long *out) {
    uint32_t i = 12345u;
    for (int k = 0; k < warm; ++k) i = p[(size_t)i * 32];
    long long t0 = clock64();
    for (int k = 0; k < hops; ++k) i = p[(size_t)i * 32];
    long long t1 = clock64();
    out[0] = t1 - t0; out[1] = i;
}

__device__ __forceinline__ uint32_t xs32(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// every thread: `iters` rounds of MLP independent random 16-byte loads (64-byte aligned addresses over n_bursts bursts)
template <int MLP>
__global__ void k_gather(const int4 *p, unsigned long long n_bursts, int iters, unsigned long long *sink) {
    uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1u);
    int acc = 0;
    for (int k = 0; k < iters; ++k) {
        int4 v[MLP];
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            const unsigned long long r = ((unsigned long long)xs32(s) << 20) ^ xs32(s);
            v[j] = p[(r & (n_bursts - 1)) * 4];
        }
#pragma unroll
        for (int j = 0; j < MLP; ++j) acc += v[j].x ^ v[j].w;
    }
    if (acc == 0x7fffffff) sink[0] = acc;
}

// the same with a DEPENDENT chain per thread (address of load k+1 comes from the value of load k, as in a tree walk)
__global__ void k_gather_dep(const int4 *p, unsigned long long n_bursts, int iters, unsigned long long *sink, int lane_stride) {
    uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1u);
    unsigned long long r = s;
    int acc = 0;
    if ((threadIdx.x & 31) % lane_stride) return;           // lane_stride 8: four loading lanes per warp (one per 8-lane group, as the walk's row load)
    for (int k = 0; k < iters; ++k) {
        const int4 v = p[(r & (n_bursts - 1)) * 4];
        acc += v.x;
        r = ((unsigned long long)xs32(s) << 20) ^ xs32(s) ^ (unsigned)v.y;    // v.y is 0: the dependence is real, the address still random
    }
    if (acc == 0x7fffffff) sink[0] = acc;
}

int main(int argc, char **argv) {
    int dev = 0, n_sm = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    printf("SMs %d, clock attr %.0f MHz, free %.1f GB\n", n_sm, khz / 1e3, free_b / 1e9);
    const size_t max_bytes = (size_t)128 << 30;
    size_t bytes = max_bytes;
    while (bytes + ((size_t)4 << 30) > free_b) bytes >>= 1;
    uint8_t *buf = nullptr;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    long long *d_out; unsigned long long *d_sink;
    CK(cudaMalloc(&d_out, 16)); CK(cudaMalloc(&d_sink, 8));
    printf("buffer %.1f GB\n", bytes / 1e9);

    printf("\n== part 1: pointer chase, one thread, 128-byte lines (clk per hop)\n");
    for (size_t fp = (size_t)16 << 20; fp <= bytes; fp <<= 2) {
        const unsigned long long n_lines = fp / 128;
        if (n_lines > 0xffffffffull) break;
        k_init_chain<<<n_sm * 8, 256>>>(reinterpret_cast<uint32_t *>(buf), n_lines);
        CK(cudaDeviceSynchronize());
        k_chase<<<1, 1>>>(reinterpret_cast<const uint32_t *>(buf), 256, 4096, d_out);
        CK(cudaDeviceSynchronize());
        long long h[2];
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("footprint %8.0f MB   %7.1f clk/hop\n", fp / 1048576.0, h[0] / 4096.0);
    }
    CK(cudaMemset(buf, 0, bytes));

    printf("\n== part 2: random 16-byte loads at 64-byte granularity (G bursts/s); threads = SMs x ctas x 256\n");
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char *name, int mlp, int ctas_per_sm, size_t fp, int lane_stride = 1) {
        const unsigned long long n_bursts = fp / 64;
        const int iters = 512 / (mlp > 0 ? mlp : 1);
        const int grid = n_sm * ctas_per_sm;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            if (mlp == 0) k_gather_dep<<<grid, 256>>>(reinterpret_cast<const int4 *>(buf), n_bursts, 512, d_sink, lane_stride);
            else if (mlp == 1) k_gather<1><<<grid, 256>>>(reinterpret_cast<const int4 *>(buf), n_bursts, iters, d_sink);
            else if (mlp == 2) k_gather<2><<<grid, 256>>>(reinterpret_cast<const int4 *>(buf), n_bursts, iters, d_sink);
            else if (mlp == 4) k_gather<4><<<grid, 256>>>(reinterpret_cast<const int4 *>(buf), n_bursts, iters, d_sink);
            else k_gather<8><<<grid, 256>>>(reinterpret_cast<const int4 *>(buf), n_bursts, iters, d_sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double loads = (double)grid * 256 * 512 / lane_stride;
        printf("%-10s footprint %8.0f MB  ctas/SM %d (%2d warps/SM)  %7.3f ms  %6.1f G loads/s  %6.2f TB/s of 64-byte bursts\n", name, fp / 1048576.0, ctas_per_sm,
               ctas_per_sm * 8, best, loads / best / 1e6, loads * 64 / best / 1e9);
    };
    const size_t fps[] = {(size_t)64 << 20, (size_t)1 << 30, (size_t)8 << 30, bytes};
    for (size_t fp : fps) {
        if (fp > bytes) continue;
        run("dependent", 0, 4, fp);
        run("dependent", 0, 8, fp);
        run("dep 4/warp", 0, 4, fp, 8);
        run("dep 4/warp", 0, 8, fp, 8);
        run("mlp1", 1, 8, fp);
        run("mlp2", 2, 8, fp);
        run("mlp4", 4, 4, fp);
        run("mlp4", 4, 8, fp);
        run("mlp8", 8, 8, fp);
    }
    return 0;
}
