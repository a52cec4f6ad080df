"""Model check of the peer-memory reduction protocol of dd_field_update_p2p_k
(csrc/dd_kernels.cu: p2p_push / p2p_collect), with threads as ranks and Python lists as the
IPC-mapped buffers.  Each "rank" runs what its stream runs per Picard iteration:

    particle kernel : add this iteration's deposits to MY accumulators
    field kernel    : n = reductions done so far + 1, b = n & 1;
                      copy my accumulators into inbox[b][me] of EVERY rank and zero them;
                      ready[b][me] := n in every rank's buffer;
                      wait until MY ready[b][*] >= n; sum MY inbox[b][*] (rank order); (field phase)

with random delays everywhere and NO acknowledgement of the reads: the double buffering alone must
keep a fast rank from overwriting a slot a slow rank is still summing (a rank can push reduction n+2
only after its own reduction n+1 completed, which needed every peer's push of n+1, which every peer
issued after it had finished reading reduction n).  Checked: every rank obtains the exact sum of
every iteration and nobody deadlocks.  The CUDA kernel needs the same argument plus memory ordering
(release/acquire at system scope), which this model does not cover."""
import random
import threading
import time

import pytest


class Rank(threading.Thread):
    def __init__(self, rank, world, acc, inbox, ready, iters, out, seed, jitter):
        super().__init__(daemon=True)
        self.rank, self.world, self.acc, self.inbox, self.ready = rank, world, acc, inbox, ready
        self.iters, self.out = iters, out
        self.rng = random.Random(seed)
        self.jitter = jitter
        self.error = None

    def nap(self, scale=1.0):
        if self.jitter and self.rng.random() < 0.5:
            time.sleep(self.rng.random() * self.jitter * scale)

    def run(self):
        try:
            me, W = self.rank, self.world
            acc = self.acc[me]
            n = 0
            for it in range(self.iters):
                self.nap()
                for i in range(len(acc)):                              # particle kernel
                    acc[i] += (me + 1) * 1000 + it * 10 + i
                self.nap()
                n += 1                                                 # field kernel: push
                b = n & 1
                vals = list(acc)
                for i in range(len(acc)):
                    acc[i] = 0
                for t in range(W):
                    slot = self.inbox[t][b][me]
                    for i, v in enumerate(vals):
                        slot[i] = v
                        if i == 0:
                            self.nap(0.2)                              # a slot is never written atomically
                for t in range(W):
                    self.ready[t][b][me] = n
                t0 = time.time()
                for t in range(W):                                     # collect
                    while self.ready[me][b][t] - n < 0:
                        if time.time() - t0 > 20:
                            raise TimeoutError("rank %d: ready[%d][%d] stuck at %d < %d" % (me, b, t, self.ready[me][b][t], n))
                        time.sleep(0)
                self.nap(3.0)                                          # a slow reader
                total = [sum(self.inbox[me][b][r][i] for r in range(W)) for i in range(len(acc))]
                self.nap()                                             # (field phase)
                self.out[me].append(total)
        except Exception as e:                                        # noqa: BLE001
            self.error = e


@pytest.mark.parametrize("world,jitter", [(2, 0.0), (2, 2e-4), (4, 2e-4), (8, 1e-4)])
def test_every_rank_gets_every_sum_and_nobody_deadlocks(world, jitter):
    n, iters = 5, 150
    acc = [[0] * n for _ in range(world)]
    inbox = [[[[0] * n for _ in range(world)] for _ in range(2)] for _ in range(world)]
    ready = [[[0] * world for _ in range(2)] for _ in range(world)]
    out = [[] for _ in range(world)]
    ranks = [Rank(r, world, acc, inbox, ready, iters, out, 100 + r, jitter) for r in range(world)]
    for r in ranks:
        r.start()
    for r in ranks:
        r.join(timeout=90)
        assert not r.is_alive(), "deadlock"
        assert r.error is None, r.error
    for it in range(iters):
        expect = [sum((r + 1) * 1000 + it * 10 + i for r in range(world)) for i in range(n)]
        for r in range(world):
            assert out[r][it] == expect, (r, it)
    assert all(v == 0 for b in acc for v in b)
