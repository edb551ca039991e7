"""pyTetris.Tetris — drop-in for the (absent, third-party) hrpan/pyTetris module consumed by the reference:
play.py:1,75-76,150,169, agents/agent.py:70,103,114,143-144.  The object is a handle around the 80-byte packed
record of SPEC_PYTETRIS.md §6.  Construction, reset(), play() and getState() all run on the GPU through the C-ABI
(b200_tetris_new / b200_tetris_step / b200_tetris_state); no rule of the game is implemented on the host."""
import numpy as np

from . import _lib as L


class Tetris:
    __slots__ = ("_rec", "_args")

    def __init__(self, boardsize=(20, 10), actions_per_drop=1, scoring=0, randomizer=0, _record=None):
        if tuple(boardsize) != (20, 10):
            raise ValueError("only 20x10 boards are supported (SPEC_PYTETRIS.md §1)")
        self._args = (int(actions_per_drop), int(scoring), int(randomizer))
        if _record is not None:
            self._rec = np.ascontiguousarray(_record, np.uint32).copy()
        else:
            self._rec = new_games(1, self._args, None)[0]

    # ---- state-changing calls: all on the device
    def play(self, action):
        a = np.array([int(action)], np.int32)
        L.check(L.lib().b200_tetris_step(L.ptr(self._rec), L.ptr(a), 1))

    def reset(self):
        L.check(L.lib().b200_tetris_new(L.ptr(self._rec), 1, *self._args, None, 1))

    def seed(self, s):
        self._rec = new_games(1, self._args, np.array([int(s) & 0xffffffff], np.uint32))[0]

    def copy_from(self, other):
        self._rec[:] = other._rec
        self._args = other._args

    def clone(self):
        return Tetris((20, 10), *self._args, _record=self._rec)

    def equiv(self, other):
        return bool(np.array_equal(self._rec, other._rec))

    def __eq__(self, other):
        return isinstance(other, Tetris) and bool(np.array_equal(self._rec, other._rec))

    def __hash__(self):
        return hash(self._rec.tobytes())

    def hash(self):
        return hash(self._rec.tobytes()) & 0xffffffffffffffff

    # ---- observation (agents/agent.py:116)
    def getState(self):
        out = np.zeros((20, 10), np.int8)
        L.check(L.lib().b200_tetris_state(L.ptr(self._rec), L.ptr(out), 1))
        return out

    def _getState(self):
        return self.getState().ravel().tolist()

    def printState(self):
        for row in self.getState():
            print("".join("." if v == 0 else ("#" if v > 0 else "o") for v in row))

    def getScore(self):
        return self.score

    def get_record(self):
        return self._rec.copy()

    def set_record(self, rec):
        self._rec[:] = np.asarray(rec, np.uint32)

    # ---- attributes read by play.py:144-148, util/Data.py:74-77 (decoded from the record, SPEC §6)
    @property
    def end(self):
        return bool((int(self._rec[10]) >> 21) & 1)

    @property
    def score(self):
        return int(np.int32(self._rec[13]))

    @property
    def combo(self):
        return int(self._rec[11]) >> 8

    @property
    def line_clears(self):
        return int(np.int32(self._rec[14]))

    @property
    def line_stats(self):
        return self._rec[15:19].astype(np.int32)


def new_games(n, env_args=(1, 0, 0), seeds=None):
    """n fresh packed games built on the device; seeds: uint32[n] or None (default seed)."""
    if len(env_args) == 4:
        env_args = env_args[1:]
    recs = np.zeros((n, L.REC_WORDS), np.uint32)
    s = None if seeds is None else np.ascontiguousarray(seeds, np.uint32)
    L.check(L.lib().b200_tetris_new(L.ptr(recs), int(n), int(env_args[0]), int(env_args[1]), int(env_args[2]), L.ptr(s), 0))
    return recs


def step_games(recs, actions):
    """Tetris.play over a batch of packed games (in place)."""
    recs = np.ascontiguousarray(recs, np.uint32)
    a = np.ascontiguousarray(actions, np.int32)
    L.check(L.lib().b200_tetris_step(L.ptr(recs), L.ptr(a), len(a)))
    return recs


def states_of(recs):
    recs = np.ascontiguousarray(recs, np.uint32).reshape(-1, L.REC_WORDS)
    out = np.zeros((len(recs), 20, 10), np.int8)
    L.check(L.lib().b200_tetris_state(L.ptr(recs), L.ptr(out), len(recs)))
    return out
