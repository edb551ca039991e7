// TEST AID — a HOST build of the device board-step header (tetris_mcts_b200/csrc/tetris_dev.cuh), so that the exact code the
// kernels run (bitboard step, closed-form hard drop, pack/unpack, observation key) is checked against the oracle in the CPU test
// suite as well, before any GPU time is spent.  Compiled by tests/test_cpu_device_env_header.py with g++; the CUDA intrinsics the
// header uses are given their documented meaning below.  Not part of the product: the library has no CPU path.
#include <stdint.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define __constant__ static const
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {   // offset-th set bit at or above base (offset >= 1)
    int n = 0;
    for (unsigned i = base; i < 32; ++i) if ((mask >> i) & 1u) { if (++n == offset) return i; }
    return 0xffffffffu;
}
#include "tetris_dev.cuh"
extern "C" void host_obskeys(const uint32_t *recs, uint32_t *keys, int n) {
    for (int i = 0; i < n; ++i) {
        uint32_t w[b200::REC_WORDS], k[b200::KEY_WORDS];
        memcpy(w, recs + (size_t)i * b200::REC_WORDS, sizeof(w));
        b200::Game g;
        b200::unpack(g, w);
        b200::obskey(g, k);
        memcpy(keys + (size_t)i * b200::KEY_WORDS, k, sizeof(k));
    }
}
extern "C" void host_play_records(uint32_t *recs, const int32_t *actions, int n) {
    for (int i = 0; i < n; ++i) {
        uint32_t w[b200::REC_WORDS];
        memcpy(w, recs + (size_t)i * b200::REC_WORDS, sizeof(w));
        b200::Game g;
        b200::unpack(g, w);
        b200::play(g, actions[i]);
        b200::pack(g, w);
        memcpy(recs + (size_t)i * b200::REC_WORDS, w, sizeof(w));
    }
}
