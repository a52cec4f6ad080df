"""Model check of the peer-memory reduction protocol of dd_field_update_p2p_k
(csrc/dd_kernels.cu: p2p_reduce / p2p_finish), with threads as ranks and Python lists as the
IPC-mapped buffers.  Each "rank" runs what its stream runs per iteration:

    particle kernel : add this iteration's deposits to MY accumulators
    field kernel    : ready[me] := seq in every rank's buffer; wait until my ready[*] >= seq;
                      sum every rank's accumulators (rank order); done[me] := seq in every rank's
                      buffer; (field phase); wait until my done[*] >= seq; zero MY accumulators

with random delays everywhere, and with sequence numbers that are skipped by all ranks alike (the
no-op launches of the enqueue-ahead loop).  Checked: every rank obtains the exact sum of every
iteration -- i.e. nobody reads accumulators that are incomplete, already zeroed or already being
refilled -- and nobody deadlocks.  The CUDA kernel needs the same argument plus memory ordering
(release/acquire at system scope), which this model does not cover."""
import random
import threading
import time

import pytest


class Rank(threading.Thread):
    def __init__(self, rank, world, bufs, flags, seqs, out, seed, jitter):
        super().__init__(daemon=True)
        self.rank, self.world, self.bufs, self.flags, self.seqs, self.out = rank, world, bufs, flags, seqs, out
        self.rng = random.Random(seed)
        self.jitter = jitter
        self.error = None

    def nap(self):
        if self.jitter and self.rng.random() < 0.5:
            time.sleep(self.rng.random() * self.jitter)

    def wait(self, which, seq):
        mine = self.flags[self.rank][which]
        t0 = time.time()
        for t in range(self.world):
            while mine[t] - seq < 0:
                if time.time() - t0 > 20:
                    raise TimeoutError("rank %d: %s[%d] stuck at %d < %d" % (self.rank, which, t, mine[t], seq))
                time.sleep(0)

    def run(self):
        try:
            me, W = self.rank, self.world
            for it, seq in enumerate(self.seqs):
                self.nap()
                acc = self.bufs[me]                                   # particle kernel
                for i in range(len(acc)):
                    acc[i] += (me + 1) * 1000 + it * 10 + i
                self.nap()
                for t in range(W):                                    # field kernel: ready
                    self.flags[t]["ready"][me] = seq
                self.wait("ready", seq)
                self.nap()
                total = [sum(self.bufs[r][i] for r in range(W)) for i in range(len(acc))]
                for t in range(W):
                    self.flags[t]["done"][me] = seq
                self.nap()                                            # (field phase)
                self.wait("done", seq)
                for i in range(len(acc)):
                    acc[i] = 0
                self.out[me].append(total)
        except Exception as e:                                        # noqa: BLE001
            self.error = e


@pytest.mark.parametrize("world,jitter", [(2, 0.0), (2, 2e-4), (4, 2e-4), (8, 1e-4)])
def test_every_rank_gets_every_sum_and_nobody_deadlocks(world, jitter):
    n, iters = 5, 150
    rng = random.Random(world)
    seqs, s = [], 0
    for _ in range(iters):
        s += 1 + (rng.random() < 0.2) * rng.randint(1, 3)      # gaps: launches that were no-ops on every rank
        seqs.append(s)
    bufs = [[0] * n for _ in range(world)]
    flags = [dict(ready=[0] * world, done=[0] * world) for _ in range(world)]
    out = [[] for _ in range(world)]
    ranks = [Rank(r, world, bufs, flags, seqs, out, 100 + r, jitter) for r in range(world)]
    for r in ranks:
        r.start()
    for r in ranks:
        r.join(timeout=60)
        assert not r.is_alive(), "deadlock"
        assert r.error is None, r.error
    for it in range(iters):
        expect = [sum((r + 1) * 1000 + it * 10 + i for r in range(world)) for i in range(n)]
        for r in range(world):
            assert out[r][it] == expect, (r, it)
    assert all(v == 0 for b in bufs for v in b)
