/*
 * tetris_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of /SPEC_PYTETRIS.md (see tetris_oracle.h for
 * why the reference cannot pin it: the env is the absent third-party hrpan/pyTetris 1.0.0).  Written cell by
 * cell on purpose, so that it shares no arithmetic with the bitboard device code it checks.
 */
#include "tetris_oracle.h"
#include <string.h>

/* SPEC §2 cell tables: [piece][rot][cell] = {row, col} inside the 4x4 box */
static const int8_t CELLS[7][4][4][2] = {
    /* I */ {{{1,0},{1,1},{1,2},{1,3}}, {{0,2},{1,2},{2,2},{3,2}}, {{2,0},{2,1},{2,2},{2,3}}, {{0,1},{1,1},{2,1},{3,1}}},
    /* O */ {{{0,1},{0,2},{1,1},{1,2}}, {{0,1},{0,2},{1,1},{1,2}}, {{0,1},{0,2},{1,1},{1,2}}, {{0,1},{0,2},{1,1},{1,2}}},
    /* T */ {{{0,1},{1,0},{1,1},{1,2}}, {{0,1},{1,1},{1,2},{2,1}}, {{1,0},{1,1},{1,2},{2,1}}, {{0,1},{1,0},{1,1},{2,1}}},
    /* S */ {{{0,1},{0,2},{1,0},{1,1}}, {{0,1},{1,1},{1,2},{2,2}}, {{1,1},{1,2},{2,0},{2,1}}, {{0,0},{1,0},{1,1},{2,1}}},
    /* Z */ {{{0,0},{0,1},{1,1},{1,2}}, {{0,2},{1,1},{1,2},{2,1}}, {{1,0},{1,1},{2,1},{2,2}}, {{0,1},{1,0},{1,1},{2,0}}},
    /* J */ {{{0,0},{1,0},{1,1},{1,2}}, {{0,1},{0,2},{1,1},{2,1}}, {{1,0},{1,1},{1,2},{2,2}}, {{0,1},{1,1},{2,0},{2,1}}},
    /* L */ {{{0,2},{1,0},{1,1},{1,2}}, {{0,1},{1,1},{2,1},{2,2}}, {{1,0},{1,1},{1,2},{2,0}}, {{0,0},{0,1},{1,1},{2,1}}},
};

static uint32_t rng_next(to_game *g) { /* SPEC §4 xorshift32 */
    uint32_t s = g->rng;
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    g->rng = s;
    return s;
}

static int collides(const to_game *g, int piece, int rot, int px, int py) {
    for (int i = 0; i < 4; ++i) {
        int r = py + CELLS[piece][rot][i][0], c = px + CELLS[piece][rot][i][1];
        if (r < 0 || r >= TO_ROWS || c < 0 || c >= TO_COLS) return 1;
        if (g->rows[r] >> c & 1) return 1;
    }
    return 0;
}

static void spawn(to_game *g) { /* SPEC §3.4 + §4 */
    int piece;
    if (g->randomizer == 0) {
        if (g->bag == 0) g->bag = 0x7f;
        int n = __builtin_popcount(g->bag);
        int k = (int)(rng_next(g) % (uint32_t)n);
        piece = 0;
        for (int b = 0; b < 7; ++b) {
            if (g->bag >> b & 1) {
                if (k == 0) { piece = b; break; }
                --k;
            }
        }
        g->bag &= ~(1u << piece);
    } else {
        piece = (int)(rng_next(g) % 7u);
    }
    g->piece = piece; g->rot = 0; g->px = 3; g->py = (piece == 0) ? -1 : 0;
    if (collides(g, g->piece, g->rot, g->px, g->py)) g->end = 1;
}

static void lock_piece(to_game *g) { /* SPEC §3.3 */
    for (int i = 0; i < 4; ++i) {
        int r = g->py + CELLS[g->piece][g->rot][i][0], c = g->px + CELLS[g->piece][g->rot][i][1];
        g->rows[r] |= (uint16_t)(1u << c);
    }
    int n = 0;
    for (int r = TO_ROWS - 1; r >= 0;) {
        if (g->rows[r] == 0x3ff) {
            for (int k = r; k > 0; --k) g->rows[k] = g->rows[k - 1];
            g->rows[0] = 0;
            ++n;
        } else {
            --r;
        }
    }
    if (n > 0) {
        static const int base[4] = {100, 300, 500, 800};
        g->combo += 1;
        g->line_clears += n;
        g->line_stats[n - 1] += 1;
        if (g->scoring == 0) g->score += base[n - 1] + 50 * (g->combo - 1);
        else g->score += n;
    } else {
        g->combo = 0;
    }
    spawn(g);
}

void to_init(to_game *g, int app, int scoring, int randomizer) {
    memset(g, 0, sizeof(*g));
    g->app = app < 1 ? 1 : (app > 255 ? 255 : app);
    g->scoring = scoring ? 1 : 0;
    g->randomizer = randomizer ? 1 : 0;
    g->rng = 0x9E3779B9u;
    g->bag = 0x7f;
    spawn(g);
}

void to_seed(to_game *g, uint32_t seed) {
    int app = g->app, sc = g->scoring, rz = g->randomizer;
    memset(g, 0, sizeof(*g));
    g->app = app; g->scoring = sc; g->randomizer = rz;
    g->rng = seed ? seed : 0x9E3779B9u;
    g->bag = 0x7f;
    spawn(g);
}

void to_reset(to_game *g) {
    uint32_t rng = g->rng;
    int app = g->app, sc = g->scoring, rz = g->randomizer;
    memset(g, 0, sizeof(*g));
    g->app = app; g->scoring = sc; g->randomizer = rz;
    g->rng = rng;
    g->bag = 0x7f;
    spawn(g);
}

void to_play(to_game *g, int action) {
    if (g->end) return;
    switch (action) {
    case 1: if (!collides(g, g->piece, g->rot, g->px - 1, g->py)) g->px -= 1; break;
    case 2: if (!collides(g, g->piece, g->rot, g->px + 1, g->py)) g->px += 1; break;
    case 3: { int r = (g->rot + 1) & 3; if (!collides(g, g->piece, r, g->px, g->py)) g->rot = r; } break;
    case 4: { int r = (g->rot + 3) & 3; if (!collides(g, g->piece, r, g->px, g->py)) g->rot = r; } break;
    case 5: {
        int d = 0;
        while (!collides(g, g->piece, g->rot, g->px, g->py + 1)) { g->py += 1; ++d; }
        if (g->scoring == 0) g->score += 2 * d;
        g->dropcnt = 0;
        lock_piece(g);
        return;
    }
    case 6:
        if (!collides(g, g->piece, g->rot, g->px, g->py + 1)) {
            g->py += 1;
            if (g->scoring == 0) g->score += 1;
        }
        break;
    default: break;
    }
    g->dropcnt += 1;
    if (g->dropcnt >= g->app) {
        g->dropcnt = 0;
        if (!collides(g, g->piece, g->rot, g->px, g->py + 1)) g->py += 1;
        else lock_piece(g);
    }
}

void to_state(const to_game *g, int8_t *out) {
    for (int r = 0; r < TO_ROWS; ++r)
        for (int c = 0; c < TO_COLS; ++c) out[r * TO_COLS + c] = (int8_t)(g->rows[r] >> c & 1);
    for (int i = 0; i < 4; ++i) {
        int r = g->py + CELLS[g->piece][g->rot][i][0], c = g->px + CELLS[g->piece][g->rot][i][1];
        if (r >= 0 && r < TO_ROWS && c >= 0 && c < TO_COLS) out[r * TO_COLS + c] = -1;
    }
}

void to_pack(const to_game *g, uint32_t *w) {
    for (int i = 0; i < 10; ++i) w[i] = (uint32_t)g->rows[2 * i] | ((uint32_t)g->rows[2 * i + 1] << 16);
    w[10] = (uint32_t)g->piece | ((uint32_t)g->rot << 3) | ((uint32_t)(g->px + 2) << 5) | ((uint32_t)(g->py + 2) << 9) |
            ((g->bag & 0x7fu) << 14) | ((uint32_t)(g->end & 1) << 21) | ((uint32_t)g->scoring << 22) |
            ((uint32_t)g->randomizer << 23) | ((uint32_t)(g->dropcnt & 0xff) << 24);
    w[11] = (uint32_t)(g->app & 0xff) | ((uint32_t)(g->combo & 0xffffff) << 8);
    w[12] = g->rng;
    w[13] = (uint32_t)g->score;
    w[14] = (uint32_t)g->line_clears;
    for (int i = 0; i < 4; ++i) w[15 + i] = (uint32_t)g->line_stats[i];
    w[19] = 0;
}

void to_unpack(to_game *g, const uint32_t *w) {
    for (int i = 0; i < 10; ++i) { g->rows[2 * i] = (uint16_t)(w[i] & 0xffff); g->rows[2 * i + 1] = (uint16_t)(w[i] >> 16); }
    g->piece = (int)(w[10] & 7); g->rot = (int)(w[10] >> 3 & 3);
    g->px = (int)(w[10] >> 5 & 15) - 2; g->py = (int)(w[10] >> 9 & 31) - 2;
    g->bag = w[10] >> 14 & 0x7f; g->end = (int)(w[10] >> 21 & 1);
    g->scoring = (int)(w[10] >> 22 & 1); g->randomizer = (int)(w[10] >> 23 & 1);
    g->dropcnt = (int)(w[10] >> 24 & 0xff);
    g->app = (int)(w[11] & 0xff); g->combo = (int)(w[11] >> 8);
    g->rng = w[12]; g->score = (int32_t)w[13]; g->line_clears = (int32_t)w[14];
    for (int i = 0; i < 4; ++i) g->line_stats[i] = (int32_t)w[15 + i];
}

void to_obskey(const to_game *g, uint32_t *k) {
    uint16_t rows[TO_ROWS];
    memcpy(rows, g->rows, sizeof(rows));
    int cell[4];
    for (int i = 0; i < 4; ++i) {
        int r = g->py + CELLS[g->piece][g->rot][i][0], c = g->px + CELLS[g->piece][g->rot][i][1];
        cell[i] = r * 10 + c;
        rows[r] &= (uint16_t)~(1u << c);   /* -1 overrides a locked cell in getState(), so the key must not see it */
    }
    for (int i = 0; i < 10; ++i) k[i] = (uint32_t)rows[2 * i] | ((uint32_t)rows[2 * i + 1] << 16);
    for (int i = 1; i < 4; ++i) { /* ascending */
        int v = cell[i], j = i - 1;
        while (j >= 0 && cell[j] > v) { cell[j + 1] = cell[j]; --j; }
        cell[j + 1] = v;
    }
    k[10] = (uint32_t)cell[0] | ((uint32_t)cell[1] << 8) | ((uint32_t)cell[2] << 16) | ((uint32_t)cell[3] << 24);
    k[11] = 0;
}

int to_equal(const to_game *a, const to_game *b) {
    uint32_t x[TO_RECORD_WORDS], y[TO_RECORD_WORDS];
    to_pack(a, x); to_pack(b, y);
    return memcmp(x, y, sizeof(x)) == 0;
}

uint64_t to_hash(const to_game *g) {
    uint32_t x[TO_RECORD_WORDS];
    to_pack(g, x);
    uint64_t h = 1469598103934665603ull; /* FNV-1a over the record words */
    for (int i = 0; i < TO_RECORD_WORDS; ++i) { h ^= x[i]; h *= 1099511628211ull; }
    return h;
}

void to_play_records(uint32_t *recs, const int32_t *actions, int n) {
    for (int i = 0; i < n; ++i) {
        to_game g;
        to_unpack(&g, recs + (size_t)i * TO_RECORD_WORDS);
        to_play(&g, actions[i]);
        to_pack(&g, recs + (size_t)i * TO_RECORD_WORDS);
    }
}
