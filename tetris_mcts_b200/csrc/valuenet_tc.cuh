// valuenet_tc.cuh — the reference value network (model/model_vv.py:13-52) on Blackwell tensor cores (sm_100a).
//
// Precision: north_star asks value outputs within 1e-5 of the reference's fp32.  Plain bf16/tf32/fp16 MMAs cannot
// reach that, so every fp32 operand x is split into two fp16 terms x = x1 + x2 (11 + 11 = 22 mantissa bits, the
// precision class of 3xTF32) and each product a*b is accumulated in fp32 (TMEM) as a1*b2 + a2*b1 + a1*b1 (the dropped
// a2*b2 is < 2^-22 relative).  Operands are pre-scaled by exact powers of two (activations x16, weights x64) so that the
// low terms stay in fp16's normal range; the epilogues undo the 2^10.  Three fp16 MMAs replace one fp32 product.
//
//   k_tc_conv  one persistent CTA per SM, two boards in flight (ping-pong):
//              decode obs key -> conv1 on CUDA cores (K = 9 is too small for an MMA) -> split -> smem
//              conv2 / conv3 as shift-GEMMs: activations live in shared memory channel-chunk-major
//              ([8-channel chunk][pixel row][16 B]), so the A operand of filter tap (dy,dx) is the SAME array
//              started (dy*W+dx) rows later — a canonical no-swizzle K-major UMMA layout with SBO = 128 B,
//              LBO = rows*16 B.  tcgen05.mma (M=128 pixels, N=32 couts, K=16) issued by one thread, accumulators
//              in TMEM, completion through tcgen05.commit -> mbarrier; epilogues read TMEM with tcgen05.ld, apply
//              bias+ReLU, re-split (fp16 x2) and write the next layer's operand (or act3 to HBM in the FC kernel's tile layout).
//   k_tc_fc    [R,1792] x [1792,256]: 128-row tiles, operands streamed by cp.async.bulk (1-D TMA) into a 5-stage
//              mbarrier ring — both operands are stored in HBM already in the canonical UMMA layout, so one bulk copy
//              per operand block needs no tensor map; warp-specialised (producer / MMA issuer / 4 epilogue warps);
//              epilogue fuses bias+ReLU+fc_out+sigmoid+affine and scatters (v, var) to the requesting tree slot.
#pragma once
#include <cuda_fp16.h>
#include "search_dev.cuh"
#include "valuenet_simt.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {   // 1-D TMA (UBLKCP)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {   // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// rows of a core matrix are 16 B apart, 8-row groups SBO apart, the two 16-byte K chunks LBO apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=fp16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr float TC_SCALE_A = 16.f, TC_SCALE_W = 64.f, TC_UNSCALE = 1.f / 1024.f;
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {   // lane i of the warp <- TMEM lane (quadrant*32 + i), 16 columns
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8_sum2(uint32_t taddr, float (&v)[8]) {
    float w1[8], w2[8];
    tmem_ld8(taddr + 32, w2);
    tmem_ld8(taddr, w1);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (w2[i] + w1[i]) * TC_UNSCALE;
}

// accumulator blocks a*W1 | a*W2 sit 32 columns apart: add the small one first, undo the operand scaling
__device__ __forceinline__ void tmem_ld16_sum2(uint32_t taddr, float (&v)[16]) {
    float w1[16], w2[16];
    tmem_ld16(taddr + 32, w2);
    tmem_ld16(taddr, w1);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (w2[i] + w1[i]) * TC_UNSCALE;
}

// x = x1 + x2 with fp16 terms (round-to-nearest each step); eight fp32 values -> two 16-byte chunks (one per split)
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &c1, uint4 &c2) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        __half h1 = __float2half_rn(x[i]);
        __half h2 = __float2half_rn(x[i] - __half2float(h1));
        a[i] = __half_as_ushort(h1); b[i] = __half_as_ushort(h2);
    }
    c1 = make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), a[4] | (a[5] << 16), a[6] | (a[7] << 16));
    c2 = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4] | (b[5] << 16), b[6] | (b[7] << 16));
}

// ---------------------------------------------------------------------------------------------------- conv kernel
constexpr int TCC_WORKERS = 512;            // warps 0-15: decode, conv1, epilogues
constexpr int TCC_ISSUER = TCC_WORKERS / 32; // warp 16: MMA issuer (one elected lane)
constexpr int TCC_THREADS = TCC_WORKERS + 32;
constexpr int TCC_R1 = 152;                 // act1 rows per board: 18x8 grid = 144 (+8: tap shifts read up to row 145)
constexpr int TCC_R2 = 144;                 // act2 rows per board: 16x6 grid = 96 (+48: M=128 tile + shifts read up to row 141)
constexpr int TCC_WBLOCK = 2 * 64 * 16;      // one (tap, half) block: [chunk 2][n = split*32 + cout][16 B]
constexpr int TCC_WBYTES = 18 * TCC_WBLOCK;  // one conv layer = 36864 B
constexpr int TCC_SLOTS = 3;                // boards in flight
constexpr int TCC_ASLOT = 2 * 4 * TCC_R1 * 16;   // operand buffer of one slot: act1 [split][chunk 4][152 rows][16 B], later overwritten
                                                 // in place by act2 [split][chunk 4][144 rows][16 B] (conv2 has finished reading by then)
constexpr int TCC_OFF_W2 = 0;
constexpr int TCC_OFF_W3 = TCC_OFF_W2 + TCC_WBYTES;
constexpr int TCC_OFF_A1 = TCC_OFF_W3 + TCC_WBYTES;
constexpr int TCC_OFF_IN = TCC_OFF_A1 + TCC_SLOTS * TCC_ASLOT;   // 2 x 200 floats
constexpr int TCC_OFF_W1 = TCC_OFF_IN + 2 * 200 * 4;       // 288 floats + 96 floats of biases
constexpr int TCC_OFF_KEY = TCC_OFF_W1 + (288 + 96) * 4;   // 2 x 12 key words of the current pair
constexpr int TCC_OFF_BAR = TCC_OFF_KEY + 2 * 12 * 4;      // 4 x TCC_SLOTS mbarriers + tmem pointer
constexpr int TCC_SMEM = TCC_OFF_BAR + 4 * TCC_SLOTS * 8 + 16;
constexpr int TCC_TMEM_COLS = 512;          // 3 slots x 2 layers x 64 columns (two partial sums of 32 couts) = 384 -> 512
constexpr int ACT3_KCHUNKS = 224;           // 1792 / 8

struct TcWeights {
    const uint8_t *wc2, *wc3;   // TCC_WBYTES each, already in the shared-memory layout
    const uint8_t *wfc;         // [split 3][k16 block 112][chunk 2][n 256][16 B]
};

// act3 in HBM, FC-tile layout: [split][tile of 128 rows][k chunk 224][row 128][8 bf16], k' = (y*4 + x)*32 + c
__device__ __forceinline__ size_t act3_off(int split, int n_tiles, int ridx, int kchunk) {
    return ((((size_t)split * n_tiles + (ridx >> 7)) * ACT3_KCHUNKS + kchunk) * 128 + (ridx & 127)) * 16;
}

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// One 3x3 convolution layer as 36 tcgen05.mma (9 taps x 2 halves of the 32 input channels x 2 activation splits).
// The operand fetch of an SS-mode MMA is shared-memory bound (~64 B/clk measured), so the two weight splits are stacked
// along N: activation split 1 meets [W1;W2] (N=64), split 2 meets W1 (N=32).  The three products land in two 32-column
// accumulator blocks (a*W1 | a*W2) that the epilogue adds.
//   R = rows per operand chunk, WGRID = width of the pixel grid the operand is stored on (tap shift = dy*WGRID + dx)
template <int R, int WGRID>
__device__ __forceinline__ void issue_conv_layer(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr) {
    const uint64_t a0 = umma_desc(a_addr, R * 16, 128), b0 = umma_desc(w_addr, 64 * 16, 128);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int sa = 0; sa < 2; ++sa) {
                const uint32_t a_off = sa * 4 * R + 2 * h * R + (tap / 3) * WGRID + (tap % 3);   // 16-byte units
                const uint32_t b_off = (tap * 2 + h) * (TCC_WBLOCK / 16);
                // the first MMA (sa = 0, N = 64) initialises both accumulator blocks; everything after accumulates
                umma_f16(tmem_d, a0 + a_off, b0 + b_off, umma_idesc_f16(128, 64 - 32 * sa), (tap | h | sa) ? 1u : 0u);
            }
        }
    }
}

__global__ void __launch_bounds__(TCC_THREADS, 1)
k_tc_conv(NetWeights W, TcWeights TW, const uint2 *req, const int32_t *n_req_ptr, const uint32_t *keys, int M, uint8_t *act3, int n_tiles,
          unsigned long long *prof) {
#define PROF_T(i) do { if (prof && do_prof) { long long _n = clock64(); pacc[i] += _n - ptick; ptick = _n; } } while (0)
    extern __shared__ __align__(128) uint8_t smem[];
    float *sIn = reinterpret_cast<float *>(smem + TCC_OFF_IN);
    float *sW1 = reinterpret_cast<float *>(smem + TCC_OFF_W1);
    float *sB = sW1 + 288;
    uint32_t *sKey = reinterpret_cast<uint32_t *>(smem + TCC_OFF_KEY);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TCC_OFF_BAR);
    uint64_t *bar_c2 = bars, *bar_c3 = bars + TCC_SLOTS;                       // tensor core -> workers: conv2 / conv3 of slot done
    uint64_t *bar_a1 = bars + 2 * TCC_SLOTS, *bar_a2 = bars + 3 * TCC_SLOTS;   // workers -> issuer: act1 / act2 of slot written
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + TCC_OFF_BAR + 4 * TCC_SLOTS * 8);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // ---- one-time setup: weights into smem, zeroed activations, barriers, TMEM
    for (int i = t; i < TCC_WBYTES / 16; i += TCC_THREADS) {
        reinterpret_cast<uint4 *>(smem + TCC_OFF_W2)[i] = reinterpret_cast<const uint4 *>(TW.wc2)[i];
        reinterpret_cast<uint4 *>(smem + TCC_OFF_W3)[i] = reinterpret_cast<const uint4 *>(TW.wc3)[i];
    }
    for (int i = t; i < TCC_SLOTS * TCC_ASLOT / 16; i += TCC_THREADS) reinterpret_cast<uint4 *>(smem + TCC_OFF_A1)[i] = make_uint4(0, 0, 0, 0);
    for (int i = t; i < 288; i += TCC_THREADS) sW1[i] = W.w1[i];
    if (t < 32) { sB[t] = W.b1[t]; sB[32 + t] = W.b2[t]; sB[64 + t] = W.b3[t]; }
    if (t == 0) {
        for (int i = 0; i < 2 * TCC_SLOTS; ++i) mbar_init(&bars[i], 1);
        for (int i = 2 * TCC_SLOTS; i < 4 * TCC_SLOTS; ++i) mbar_init(&bars[i], TCC_WORKERS);
        fence_barrier_init();
    }
    if (warp == TCC_ISSUER) tmem_alloc<TCC_TMEM_COLS>(tmem_ptr);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_req = *n_req_ptr;
    // Boards are handed out in runs of 8 consecutive requests (neighbouring act3 rows get written close in time); board i of
    // this CTA's sequence lives in slot i % 3.  Three boards are in flight at different stages (software pipeline):
    //   workers, iteration i :  S1(i) decode + conv1 | S2(i-1) conv2 epilogue | S3(i-2) conv3 epilogue
    //   issuer               :  conv2(i) then conv3(i-1)
    // Every MMA batch is issued about one full iteration before its result is consumed, so the tensor pipe works on one
    // board while the CUDA cores prepare / drain two others.
    const int n_runs = (n_req + 7) >> 3;
    int n_local = 0;
    for (int run = blockIdx.x; run < n_runs; run += gridDim.x) n_local += min(8, n_req - run * 8);
    auto board_of = [&](int i) -> int { return ((i >> 3) * (int)gridDim.x + (int)blockIdx.x) * 8 + (i & 7); };
    constexpr int NS = TCC_SLOTS;
    if (warp == TCC_ISSUER) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t s_w2 = smem_u32(smem + TCC_OFF_W2), s_w3 = smem_u32(smem + TCC_OFF_W3);
            const uint32_t s_act = smem_u32(smem + TCC_OFF_A1);
            const bool do_prof = blockIdx.x == 0;
            long long pacc[12] = {0}, ptick = clock64();
            for (int i = 0; i <= n_local; ++i) {
                if (i < n_local) {                               // conv2 (model_vv.py:34): act1 on the 18x8 grid
                    const int slot = i % NS;
                    mbar_wait(&bar_a1[slot], (uint32_t)(i / NS) & 1u);
                    PROF_T(6);
                    tc_fence_after();
                    issue_conv_layer<TCC_R1, 8>(tmem_base + slot * 128, s_act + slot * TCC_ASLOT, s_w2);
                    umma_commit(&bar_c2[slot]);
                    PROF_T(7);
                }
                if (i >= 1) {                                    // conv3 (model_vv.py:36): act2 on the compact 16x6 grid
                    const int j = i - 1, slot = j % NS;
                    mbar_wait(&bar_a2[slot], (uint32_t)(j / NS) & 1u);
                    PROF_T(8);
                    tc_fence_after();
                    issue_conv_layer<TCC_R2, 6>(tmem_base + slot * 128 + 64, s_act + slot * TCC_ASLOT, s_w3);
                    umma_commit(&bar_c3[slot]);
                    PROF_T(9);
                }
            }
            if (prof && do_prof) for (int i = 6; i < 10; ++i) atomicAdd(&prof[i], (unsigned long long)pacc[i]);
        }
    } else {
        // ===================================================== workers (512 threads)
        const bool do_prof = blockIdx.x == 0 && t == 0;
        long long pacc[12] = {0}, ptick = clock64();
        auto fetch_key = [&](int i) -> uint32_t {               // key word t of this CTA's i-th board (threads 0..11)
            if (t < 12 && i < n_local) {
                uint2 rq = req[board_of(i)];
                return keys[((size_t)rq.x * M + (rq.y & 0x0fffffffu)) * KEY_WORDS + t];
            }
            return 0u;
        };
        uint32_t kpre = fetch_key(0);
        const int q = warp & 3, cq = warp >> 2, m = q * 32 + lane;
        for (int i = 0; i < n_local + 2; ++i) {
            // ---- S1(i): decode + conv1 -> act1 of the slot
            if (i < n_local) {
                const int slot = i % NS;
                if (t < 12) sKey[t] = kpre;
                worker_barrier();
                kpre = fetch_key(i + 1);                         // latency hidden behind this iteration's work
                if (t < 200) {                                   // {-1,0,1} (model_vv.py:212)
                    const int r = t / 10, c = t - r * 10;
                    float v = (float)((sKey[r >> 1] >> ((r & 1) * 16 + c)) & 1u);
                    uint32_t pc = sKey[10], ci = (uint32_t)t;
                    if ((pc & 0xffu) == ci || ((pc >> 8) & 0xffu) == ci || ((pc >> 16) & 0xffu) == ci || (pc >> 24) == ci) v = -1.f;
                    sIn[t] = v;
                }
                worker_barrier();
                PROF_T(0);
                // conv1 (model_vv.py:32) on CUDA cores: task = (8-cout chunk, pixel of the 18x8 grid); consecutive lanes take
                // consecutive pixels, so the 16-byte operand stores of a warp are contiguous (no bank conflicts)
                for (int task = t; task < 576; task += TCC_WORKERS) {
                    const int c4 = task / 144, pix = task - c4 * 144, y = pix >> 3, x = pix & 7;
                    float acc[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = sB[c4 * 8 + e];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            float in = sIn[(y + dy) * 10 + x + dx];
                            const float4 *w = reinterpret_cast<const float4 *>(sW1 + (dy * 3 + dx) * 32 + c4 * 8);
                            float4 wa = w[0], wb = w[1];
                            acc[0] = fmaf(in, wa.x, acc[0]); acc[1] = fmaf(in, wa.y, acc[1]); acc[2] = fmaf(in, wa.z, acc[2]); acc[3] = fmaf(in, wa.w, acc[3]);
                            acc[4] = fmaf(in, wb.x, acc[4]); acc[5] = fmaf(in, wb.y, acc[5]); acc[6] = fmaf(in, wb.z, acc[6]); acc[7] = fmaf(in, wb.w, acc[7]);
                        }
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.f) * TC_SCALE_A;
                    uint4 c1, c2;
                    split8(acc, c1, c2);
                    uint8_t *base = smem + TCC_OFF_A1 + slot * TCC_ASLOT + (c4 * TCC_R1 + pix) * 16;
                    *reinterpret_cast<uint4 *>(base) = c1;
                    *reinterpret_cast<uint4 *>(base + 4 * TCC_R1 * 16) = c2;
                }
                fence_async_smem();
                mbar_arrive(&bar_a1[slot]);
                PROF_T(1);
            }
            // ---- S2(i-1): conv2 epilogue: bias + ReLU + split -> act2 (16x6 compact grid), in place of the slot's act1
            if (i >= 1 && i - 1 < n_local) {
                const int j = i - 1, slot = j % NS;
                mbar_wait(&bar_c2[slot], (uint32_t)(j / NS) & 1u);
                PROF_T(2);
                tc_fence_after();
                float v[8];
                tmem_ld8_sum2(tmem_base + ((uint32_t)(q * 32) << 16) + slot * 128 + cq * 8, v);
                const int y = m >> 3, x = m & 7;
                if (x < 6) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[e] + sB[32 + cq * 8 + e], 0.f) * TC_SCALE_A;
                    uint4 c1, c2;
                    split8(o, c1, c2);
                    uint8_t *base = smem + TCC_OFF_A1 + slot * TCC_ASLOT + (cq * TCC_R2 + y * 6 + x) * 16;
                    *reinterpret_cast<uint4 *>(base) = c1;
                    *reinterpret_cast<uint4 *>(base + 4 * TCC_R2 * 16) = c2;
                }
                tc_fence_before();
                fence_async_smem();
                mbar_arrive(&bar_a2[slot]);
                PROF_T(3);
            }
            // ---- S3(i-2): conv3 epilogue: bias + ReLU + split -> act3 in HBM (FC tile layout)
            if (i >= 2 && i - 2 < n_local) {
                const int j = i - 2, slot = j % NS, ridx = board_of(j);
                mbar_wait(&bar_c3[slot], (uint32_t)(j / NS) & 1u);
                PROF_T(4);
                tc_fence_after();
                const int y = m / 6, x = m - y * 6;
                float v[8];
                tmem_ld8_sum2(tmem_base + ((uint32_t)(q * 32) << 16) + slot * 128 + 64 + cq * 8, v);
                if (m < 84 && x < 4) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[e] + sB[64 + cq * 8 + e], 0.f) * TC_SCALE_A;
                    uint4 c1, c2;
                    split8(o, c1, c2);
                    const int kc = (y * 4 + x) * 4 + cq;
                    *reinterpret_cast<uint4 *>(act3 + act3_off(0, n_tiles, ridx, kc)) = c1;
                    *reinterpret_cast<uint4 *>(act3 + act3_off(1, n_tiles, ridx, kc)) = c2;
                }
                tc_fence_before();
                PROF_T(5);
            }
        }
        if (prof && do_prof) for (int i = 0; i < 6; ++i) atomicAdd(&prof[i], (unsigned long long)pacc[i]);
    }
#undef PROF_T
    tc_fence_before();
    __syncthreads();
    if (warp == TCC_ISSUER) tmem_dealloc<TCC_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------- fc kernel
constexpr int TCF_THREADS = 192;            // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int TCF_STAGES = 8;
constexpr int TCF_A_BYTES = 2 * 128 * 16;   // one split of one k16 block of the A tile
constexpr int TCF_B_BYTES = 2 * 256 * 16;
constexpr int TCF_STAGE = 2 * TCF_A_BYTES + 2 * TCF_B_BYTES;   // 24576
constexpr int TCF_KBLOCKS = 112;            // 1792 / 16
constexpr int TCF_OFF_BAR = TCF_STAGES * TCF_STAGE;
constexpr int TCF_OFF_EPI = TCF_OFF_BAR + 256;                 // (2*stages + 2 mbarriers + tmem ptr fit in 256 B) bias[256] | wout[2][256] | bout/ub/lb
static_assert((2 * TCF_STAGES + 2) * 8 + 4 <= 256, "barrier block overflows into the epilogue constants");
constexpr int TCF_SMEM = TCF_OFF_EPI + (256 * 3 + 8) * 4;
constexpr int TCF_TMEM_COLS = 256;

__global__ void __launch_bounds__(TCF_THREADS, 1)
k_tc_fc(NetWeights W, TcWeights TW, const uint8_t *act3, int n_tiles_alloc, const uint2 *req, const int32_t *n_req_ptr, float2 *eval_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + TCF_OFF_BAR);     // [stage] operands landed
    uint64_t *empty = full + TCF_STAGES;                                    // [stage] operands consumed
    uint64_t *acc_full = empty + TCF_STAGES;                                // accumulator complete
    uint64_t *acc_empty = acc_full + 1;                                     // accumulator drained by the epilogue
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 1);
    float *sBias = reinterpret_cast<float *>(smem + TCF_OFF_EPI), *sWo = sBias + 256, *sTail = sWo + 512;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < 256; i += TCF_THREADS) { sBias[i] = W.bfc1[i]; sWo[i] = W.wout[i]; sWo[256 + i] = W.wout[256 + i]; }
    if (t < 2) { sTail[t] = W.bout[t]; sTail[2 + t] = W.ub[t]; sTail[4 + t] = W.lb[t]; }
    if (t == 0) {
        for (int i = 0; i < TCF_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TCF_TMEM_COLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_req = *n_req_ptr;
    const int n_tiles = (n_req + 127) >> 7;
    if (warp == 0) {
        if (lane == 0) {   // ===== producer: bulk copies of the pre-laid-out operand blocks
            int stage = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int j = 0; j < TCF_KBLOCKS; ++j) {
                    mbar_wait(&empty[stage], ph ^ 1);
                    mbar_expect_tx(&full[stage], TCF_STAGE);
                    uint8_t *dst = smem + stage * TCF_STAGE;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        bulk_g2s(dst + s * TCF_A_BYTES, act3 + (((size_t)s * n_tiles_alloc + tile) * ACT3_KCHUNKS + 2 * j) * 2048, TCF_A_BYTES, &full[stage]);
                        bulk_g2s(dst + 2 * TCF_A_BYTES + s * TCF_B_BYTES, TW.wfc + ((size_t)s * TCF_KBLOCKS + j) * TCF_B_BYTES, TCF_B_BYTES, &full[stage]);
                    }
                    if (++stage == TCF_STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer: D[128 x 256] += A[128 x 16] * B[256 x 16]^T, six split terms per k block
            const uint32_t idesc = umma_idesc_f16(128, 256);
            int stage = 0; uint32_t ph = 0, aph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(acc_empty, aph ^ 1);
                tc_fence_after();
                uint32_t acc = 0;
                for (int j = 0; j < TCF_KBLOCKS; ++j) {
                    mbar_wait(&full[stage], ph);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + stage * TCF_STAGE);
#pragma unroll
                    for (int term = 0; term < 3; ++term) {   // a1*b2, a2*b1, a1*b1 (small terms first)
                        const int sa = term == 1 ? 1 : 0, sb = term == 0 ? 1 : 0;
                        uint64_t ad = umma_desc(sbase + sa * TCF_A_BYTES, 128 * 16, 128);
                        uint64_t bd = umma_desc(sbase + 2 * TCF_A_BYTES + sb * TCF_B_BYTES, 256 * 16, 128);
                        umma_f16(tmem_base, ad, bd, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == TCF_STAGES) { stage = 0; ph ^= 1; }
                }
                umma_commit(acc_full);
                aph ^= 1;
            }
        }
    } else {   // ===== epilogue warps 2..5: TMEM quadrant = warp % 4, one row per thread
        const int q = warp & 3, row = q * 32 + lane;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(acc_full, aph);
            tc_fence_after();
            float p0 = 0.f, p1 = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float h = fmaxf(v[j] * TC_UNSCALE + sBias[c0 + j], 0.f);   // model_vv.py:39-40
                    p0 = fmaf(h, sWo[c0 + j], p0); p1 = fmaf(h, sWo[256 + c0 + j], p1);   // :41
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
            const int ridx = tile * 128 + row;
            if (ridx < n_req) {
                float x0 = p0 + sTail[0], x1 = p1 + sTail[1];
                float s0 = 1.f / (1.f + expf(-x0)), s1 = 1.f / (1.f + expf(-x1));      // :42
                uint2 rq = req[ridx];
                eval_out[(size_t)rq.x * 8 + (rq.y >> 28)] =
                    make_float2(__fadd_rn(__fmul_rn(s0, sTail[2]), sTail[4]), __fadd_rn(__fmul_rn(s1, sTail[3]), sTail[5]));   // :51
            }
            aph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TCF_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------- host side
struct TcState {
    uint8_t *d_w = nullptr;      // wc2 | wc3 | wfc
    TcWeights TW{};
    uint8_t *d_act3 = nullptr; size_t tiles = 0;
};

static inline void host_split2(float x, uint16_t *o) {   // x*scale = h1 + h2 in fp16
    __half h1 = __float2half_rn(x);
    __half h2 = __float2half_rn(x - __half2float(h1));
    memcpy(&o[0], &h1, 2); memcpy(&o[1], &h2, 2);
}
static inline float host_half_f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }

// w = the state_dict-order weight vector of include/b200_tetris_mcts.h.  Pure re-layout + bf16 splitting.
static int tc_prepare(void **state, const float *w, cudaStream_t stream) {
    TcState *st = (TcState *)*state;
    if (!st) { st = new TcState(); *state = st; }
    const float *c2w = w + 288 + 32, *c3w = c2w + 9216 + 32, *f1w = c3w + 9216 + 32;
    const size_t fc_bytes = (size_t)2 * TCF_KBLOCKS * TCF_B_BYTES;
    std::vector<uint8_t> h(2 * (size_t)TCC_WBYTES + fc_bytes);
    uint16_t *p2 = reinterpret_cast<uint16_t *>(h.data()), *p3 = reinterpret_cast<uint16_t *>(h.data() + TCC_WBYTES);
    uint16_t *pf = reinterpret_cast<uint16_t *>(h.data() + 2 * (size_t)TCC_WBYTES);
    for (int layer = 0; layer < 2; ++layer) {
        const float *cw = layer ? c3w : c2w;
        uint16_t *dst = layer ? p3 : p2;
        for (int tap = 0; tap < 9; ++tap)
            for (int hh = 0; hh < 2; ++hh)
                for (int c2 = 0; c2 < 2; ++c2)
                    for (int n = 0; n < 32; ++n)
                        for (int e = 0; e < 8; ++e) {
                            int ci = 16 * hh + 8 * c2 + e;
                            uint16_t s2[2];
                            host_split2(cw[(n * 32 + ci) * 9 + tap] * TC_SCALE_W, s2);
                            for (int s = 0; s < 2; ++s) dst[((((size_t)(tap * 2 + hh)) * 2 + c2) * 64 + s * 32 + n) * 8 + e] = s2[s];
                        }
    }
    for (int j = 0; j < TCF_KBLOCKS; ++j)
        for (int c2 = 0; c2 < 2; ++c2)
            for (int n = 0; n < 256; ++n)
                for (int e = 0; e < 8; ++e) {
                    int kp = j * 16 + c2 * 8 + e, p = kp >> 5, c = kp & 31;      // k' = pixel*32 + channel, pixel = y*4 + x
                    uint16_t s2[2];
                    host_split2(f1w[(size_t)n * 1792 + c * 56 + p] * TC_SCALE_W, s2);
                    for (int s = 0; s < 2; ++s) pf[((((size_t)s * TCF_KBLOCKS + j) * 2 + c2) * 256 + n) * 8 + e] = s2[s];
                }
    if (!st->d_w && cudaMalloc(&st->d_w, h.size()) != cudaSuccess) return 1;
    if (cudaMemcpyAsync(st->d_w, h.data(), h.size(), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;
    st->TW.wc2 = st->d_w; st->TW.wc3 = st->d_w + TCC_WBYTES; st->TW.wfc = st->d_w + 2 * (size_t)TCC_WBYTES;
    if (cudaFuncSetAttribute(k_tc_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, TCC_SMEM) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tc_fc, cudaFuncAttributeMaxDynamicSharedMemorySize, TCF_SMEM) != cudaSuccess) return 1;
    return 0;
}

static int tc_ensure_act3(TcState *st, size_t max_rows, cudaStream_t stream) {
    size_t tiles = (max_rows + 127) / 128;
    if (st->tiles >= tiles) return 0;
    if (st->d_act3) { cudaStreamSynchronize(stream); cudaFree(st->d_act3); st->d_act3 = nullptr; }
    size_t bytes = (size_t)2 * tiles * ACT3_KCHUNKS * 2048;
    if (cudaMalloc(&st->d_act3, bytes) != cudaSuccess) return 1;
    cudaMemsetAsync(st->d_act3, 0, bytes, stream);
    st->tiles = tiles;
    return 0;
}

static void tc_destroy(void *state) {
    TcState *st = (TcState *)state;
    if (!st) return;
    cudaFree(st->d_w); cudaFree(st->d_act3);
    delete st;
}

}  // namespace b200
