/*
 * rand_shim.c — TEST INFRASTRUCTURE ONLY (golden generation; never shipped, never on the product path).
 *
 * The reference draws its search randomness from two global generators: libc rand() in check_low
 * (agents/cppmodule/core.h:62,76) and Python's random.randint in the Vanilla rollout (agents/Vanilla.py:52).  Neither can be
 * reproduced on a device, so oracle and engine use one xorshift32 stream per game.  To pin the oracle's ValueSim /
 * Vanilla loops against the reference's OWN code, tests/golden/gen_golden.py runs the reference with this library LD_PRELOADed:
 * the reference's compiled core.cpp then calls THIS rand(), and gen_golden.py points Vanilla.randint at shim_next() — the
 * reference's code is unmodified, only its random source is the oracle's stream.
 *
 * rand() returns stream % 420: check_low only ever uses rand() % n with 1 <= n <= 7, and every such n divides 420, so
 * rand() % n == stream % n — the exact value the oracle computes from the 32-bit stream (a plain (int) cast could be negative).
 */
#include <stdint.h>

static uint32_t shim_state = 0x2545F491u;

void shim_seed(uint32_t s) { shim_state = s ? s : 0x2545F491u; }

uint32_t shim_next(void) {   /* xorshift32, same as agent_rand in mcts_oracle.c and rng_next in tetris_dev.cuh */
    uint32_t s = shim_state;
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    shim_state = s;
    return s;
}

int rand(void) { return (int)(shim_next() % 420u); }
