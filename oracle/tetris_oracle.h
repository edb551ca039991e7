/*
 * tetris_oracle.h — TEST INFRASTRUCTURE ONLY (CPU oracle; never shipped, never on the product path).
 *
 * CPU restatement of the board step contract in /SPEC_PYTETRIS.md.  The reference's environment is the
 * third-party module hrpan/pyTetris (pinned pyTetris==1.0.0, reference requirements.txt:15, README.md:30),
 * whose source is absent from /root/reference.  PARITY UNPINNED: nothing in the reference tree (no test, no
 * golden vector) fixes these rules; the call sites this restatement serves are agents/agent.py:103,114,143-144,
 * agents/cppmodule/agent.cpp:205,233,243 and play.py:150.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 */
#ifndef TETRIS_ORACLE_H
#define TETRIS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TO_ROWS 20
#define TO_COLS 10
#define TO_RECORD_WORDS 20
#define TO_OBSKEY_WORDS 12

typedef struct {
    uint16_t rows[TO_ROWS];   /* SPEC §2: bit x of rows[r] = locked cell (r, x) */
    int piece, rot, px, py;   /* falling piece (SPEC §2) */
    uint32_t bag;             /* SPEC §4: pieces left in the 7-bag */
    int dropcnt;              /* actions since the last gravity tick (SPEC §3.2) */
    int end;
    int app, scoring, randomizer;
    int combo;
    uint32_t rng;
    int32_t score, line_clears, line_stats[4];
} to_game;

void to_init(to_game *g, int app, int scoring, int randomizer);  /* SPEC §4 construction */
void to_seed(to_game *g, uint32_t seed);
void to_reset(to_game *g);
void to_play(to_game *g, int action);                            /* SPEC §3 */
void to_state(const to_game *g, int8_t *out200);                 /* SPEC §1 observation */
void to_pack(const to_game *g, uint32_t *rec20);                 /* SPEC §6 */
void to_unpack(to_game *g, const uint32_t *rec20);
void to_obskey(const to_game *g, uint32_t *key12);               /* SPEC §6 */
int to_equal(const to_game *a, const to_game *b);                /* SPEC §5 */
uint64_t to_hash(const to_game *g);
/* batched helper used by the tests: play actions[i] on recs[i] (80-byte records), in place */
void to_play_records(uint32_t *recs, const int32_t *actions, int n);

#ifdef __cplusplus
}
#endif
#endif
