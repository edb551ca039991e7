/*
 * pyTetris.h — TEST INFRASTRUCTURE ONLY.  Stand-in for the header the reference includes as <pyTetris.h>
 * (agents/cppmodule/core.h:12, agent.cpp:17; expected at <python include>/pyTetris/, core.cpp:4).  The real
 * header belongs to the absent third-party hrpan/pyTetris 1.0.0; this one wraps oracle/tetris_oracle.c (the CPU
 * restatement of /SPEC_PYTETRIS.md) behind the member names the reference's agent.cpp uses
 * (agent.cpp:32-38,94,99,123,205,233,236,243,275-281,420,440,458), so that the reference's own core.cpp and
 * agent.cpp compile UNCHANGED from /root/reference into oracle/_ref/.  Never on the product path.
 */
#ifndef PYTETRIS_ORACLE_SHIM_H
#define PYTETRIS_ORACLE_SHIM_H
#include <vector>
#include <array>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include "../tetris_oracle.h"

class Tetris {
  public:
    /* g MUST stay the first member: agent.cpp:275-276 reinterprets the Python object's buffer pointer as Tetris* */
    to_game g;
    /* mirrors of the attributes agent.cpp reads as plain members (g.end, g.score: agent.cpp:236,251,280,420) */
    bool end;
    int score;
    int combo, line_clears;
    std::array<int, 4> line_stats;

    Tetris() { to_init(&g, 1, 0, 0); sync(); }
    Tetris(std::pair<int, int> boardsize, int actions_per_drop, int scoring, int randomizer) {
        if (boardsize.first != 20 || boardsize.second != 10) throw std::invalid_argument("only 20x10 boards (SPEC §1)");
        to_init(&g, actions_per_drop, scoring, randomizer);
        sync();
    }
    void sync() {
        end = g.end != 0; score = g.score; combo = g.combo; line_clears = g.line_clears;
        for (int i = 0; i < 4; ++i) line_stats[i] = g.line_stats[i];
    }
    void play(int a) { to_play(&g, a); sync(); }
    void reset() { to_reset(&g); sync(); }
    void seed(uint32_t s) { to_seed(&g, s); sync(); }
    void copy_from(const Tetris &o) { g = o.g; sync(); }
    Tetris clone() const { Tetris t; t.copy_from(*this); return t; }
    bool equiv(const Tetris &o) const { return to_equal(&g, &o.g) != 0; }
    bool operator==(const Tetris &o) const { return to_equal(&g, &o.g) != 0; }
    size_t hash() const { return (size_t)to_hash(&g); }
    int getScore() const { return g.score; }
    std::vector<char> _getState() const {
        std::vector<char> v(200);
        to_state(&g, reinterpret_cast<int8_t *>(v.data()));
        return v;
    }
    pybind11::array_t<int8_t> getState() const {
        pybind11::array_t<int8_t> a({20, 10});
        to_state(&g, a.mutable_data());
        return a;
    }
    void printState() const {
        int8_t s[200];
        to_state(&g, s);
        for (int r = 0; r < 20; ++r) {
            for (int c = 0; c < 10; ++c) std::fputc(s[r * 10 + c] == 0 ? '.' : (s[r * 10 + c] > 0 ? '#' : 'o'), stdout);
            std::fputc('\n', stdout);
        }
        std::fflush(stdout);
    }
    std::vector<uint32_t> get_record() const { std::vector<uint32_t> w(TO_RECORD_WORDS); to_pack(&g, w.data()); return w; }
    void set_record(const std::vector<uint32_t> &w) {
        if (w.size() != TO_RECORD_WORDS) throw std::invalid_argument("record must be 20 words");
        to_unpack(&g, w.data());
        sync();
    }
};
#endif
