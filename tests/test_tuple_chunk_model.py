"""Model of the per-warp chunked reservation of tuple-list records in `k_expand_grid` (expand_grid.cuh: the `cm` branch
of the drain and the epilogue): a warp reserves records 8, 16, 32, 64, 64, ... at a time with one atomic, a round of
k <= 32 finished tuples that does not fit the rest of the chunk fills it and continues in a fresh one, and the unused
rest of the LAST chunk is marked as holes.  Properties the kernel relies on: every tuple gets its own record, the
records of all warps tile [0, tuple_slots) exactly (tuples + holes), and the holes of a warp are at most one chunk."""
import random

K_TUPLE_CHUNK = 64


class Warp:
    def __init__(self, cursor):
        self.cursor = cursor            # the shared counter (a one-element list)
        self.t_next, self.t_left, self.t_chunk, self.n_tup = 0, 0, 8, 0
        self.tuples, self.holes = [], []

    def round(self, k):                 # k lanes finished a tuple in this round
        t_next, t_left = self.t_next, self.t_left
        fresh = 0
        want = max(self.t_chunk, k - t_left)
        if k > t_left:
            fresh = self.cursor[0]
            self.cursor[0] += want      # atomicAdd(&tuple_slots, want)
        for r in range(k):
            self.tuples.append(t_next + r if r < t_left else fresh + (r - t_left))
        if k > t_left:
            self.t_next, self.t_left = fresh + (k - t_left), want - (k - t_left)
            self.t_chunk = min(K_TUPLE_CHUNK, 2 * self.t_chunk)
        else:
            self.t_next, self.t_left = t_next + k, t_left - k
        self.n_tup += k

    def finish(self):
        self.holes = [self.t_next + i for i in range(self.t_left)]


def test_records_tile_the_tuple_list_exactly():
    rng = random.Random(5)
    for trial in range(200):
        cursor = [0]
        warps = [Warp(cursor) for _ in range(rng.randint(1, 12))]
        for _ in range(rng.randint(0, 400)):
            w = rng.choice(warps)
            w.round(rng.choice([1, 1, 2, 3, 7, 8, 9, 31, 32]) if rng.random() < 0.9 else rng.randint(1, 32))
        for w in warps:
            w.finish()
        tuples = [s for w in warps for s in w.tuples]
        holes = [s for w in warps for s in w.holes]
        assert len(set(tuples)) == len(tuples) == sum(w.n_tup for w in warps)          # one record per tuple
        assert sorted(tuples + holes) == list(range(cursor[0]))                         # tuples + holes = every reserved record
        for w in warps:
            assert len(w.holes) < K_TUPLE_CHUNK                                         # at most the rest of ONE chunk
            if w.n_tup and w.n_tup <= 8:
                assert len(w.holes) < 8                                                 # a handful of tuples, a handful of holes
