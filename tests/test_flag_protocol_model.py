"""CPU: a discrete-event model of the fused exchange's flag protocol (plan.cu fuse_writer / fuse_reader).

Every rank runs two chains of launches, writers and readers, one launch per tile and chain.  A launch may START once the
launch before it in its chain has started (dependent-launch chains: a launch is let in as soon as its predecessor's CTAs
are past their flag wait) and its own flag condition holds; started launches FINISH in any order.  Tile number s
(1, 2, ...) lives in ring slot (s - 1) % depth.

  writer s on rank r:  waits until released[slot][j] >= s - depth for all members j (the slot's previous tenant has been
                       read on every member), stores its block into every member's slot, then its last CTA sets
                       arrived[slot][r] = s in every member's flag block.
  reader s on rank r:  waits until arrived[slot][j] >= s for all j, reads its own slot, then sets
                       released[slot][r] = s in every member's flag block.

The model checks, over random schedules: no deadlock, every reader sees exactly tile s from every member in its slot
(nothing stale, nothing overwritten early), and flag words only ever grow.  With ONE word per (phase, member) instead of
one per ring slot - the layout that was enough while launches ran one after the other, before the dependent-launch
chains - launches finishing out of order make a word go backwards, which the last test reproduces."""
import random

import pytest


class World:
    def __init__(self, ranks, tiles, depth, per_slot_flags=True, seed=0):
        self.P, self.nb, self.D = ranks, tiles, depth
        self.per_slot = per_slot_flags
        self.rng = random.Random(seed)
        nslots = depth if per_slot_flags else 1
        # flags[r][kind][slot][j]: word in rank r's flag block written by member j
        self.flags = [{k: [[0] * ranks for _ in range(nslots)] for k in ("arrived", "released")} for _ in range(ranks)]
        self.slots = [[[0] * ranks for _ in range(depth)] for _ in range(ranks)]   # slots[r][slot][j]: tile number of j's block
        self.started = {(r, k): 0 for r in range(ranks) for k in "wr"}            # launches started per chain
        self.running = []                                                          # (rank, kind, s)
        self.finished = set()
        self.went_backwards = False

    def word(self, r, kind, slot):
        return self.flags[r][kind][slot if self.per_slot else 0]

    def set_flag(self, r, kind, slot, j, value):
        w = self.word(r, kind, slot)
        if value < w[j]:
            self.went_backwards = True
        w[j] = value

    def can_start(self, r, kind):
        s = self.started[(r, kind)] + 1
        if s > self.nb:
            return None
        slot = (s - 1) % self.D
        if kind == "w":
            ok = s <= self.D or all(v >= s - self.D for v in self.word(r, "released", slot))
        else:
            ok = all(v >= s for v in self.word(r, "arrived", slot))
        return s if ok else None

    def start(self, r, kind, s):
        self.started[(r, kind)] = s
        self.running.append((r, kind, s))
        if kind == "w":                                         # adversarial: the stores land the moment the wait is over
            for m in range(self.P):
                self.slots[m][(s - 1) % self.D][r] = s

    def finish(self, r, kind, s):
        slot = (s - 1) % self.D
        if kind == "w":
            for m in range(self.P):
                self.set_flag(m, "arrived", slot, r, s)
        else:
            got = list(self.slots[r][slot])
            assert got == [s] * self.P, f"reader {s} on rank {r} found tiles {got} in slot {slot}"
            for m in range(self.P):
                self.set_flag(m, "released", slot, r, s)
        self.finished.add((r, kind, s))

    def run(self):
        total = 2 * self.P * self.nb
        while len(self.finished) < total:
            moves = [("finish",) + x for x in self.running]
            for r in range(self.P):
                for kind in "wr":
                    s = self.can_start(r, kind)
                    if s:
                        moves.append(("start", r, kind, s))
            assert moves, f"deadlock with {len(self.finished)} of {total} launches finished"
            what, r, kind, s = self.rng.choice(moves)
            if what == "start":
                self.start(r, kind, s)
            else:
                self.running.remove((r, kind, s))
                self.finish(r, kind, s)


@pytest.mark.parametrize("ranks,tiles,depth", [(2, 2, 1), (2, 7, 3), (4, 16, 3), (8, 16, 4), (3, 5, 11), (8, 33, 2)])
def test_per_slot_flags_never_deadlock_never_go_backwards_and_readers_see_their_tile(ranks, tiles, depth):
    for seed in range(40):
        w = World(ranks, tiles, depth, per_slot_flags=True, seed=seed)
        w.run()
        assert not w.went_backwards


def test_one_word_per_member_goes_backwards_when_launches_finish_out_of_order():
    """the layout from before the chains: found on real GPUs as a flag time-out of a 2-tile plan once launches overlapped"""
    bad = 0
    for seed in range(200):
        w = World(2, 4, 3, per_slot_flags=False, seed=seed)
        try:
            w.run()
        except AssertionError:
            bad += 1
            continue
        bad += w.went_backwards
    assert bad > 0
