"""CPU test of the host logic of SheathSim.picard (enqueue-ahead Picard loop): the C ABI is replaced
by a fake device that implements only the CONTROL contract of pic_dev_dd_picard_iter5 /
pic_dev_dd_field_update2 (include/pic_b200.h: launches are no-ops once the flag is up; the field
kernel counts iterations, records residuals and raises the flag when `r > tol and k < maxiter`
fails) with scripted residuals.  Checked: iteration counts, which launches ran as light / full
iterations, when the repair pass runs, and that the commit makes the buffer written by the LAST
EXECUTED iteration the new x0."""
import types

import numpy as np
import pytest
import torch

from pypic_b200 import sheath as S


class FakeDevice:
    def __init__(self, sim, scripts):
        self.sim, self.scripts, self.step = sim, scripts, -1
        self.log = []

    def tensor_at(self, p):
        for name, t in vars(self.sim).items():
            if isinstance(t, torch.Tensor) and t.data_ptr() <= p < t.data_ptr() + max(t.numel() * t.element_size(), 1):
                return name, t, (p - t.data_ptr()) // t.element_size()
        raise KeyError(p)

    def call(self, name, *a):
        sim = self.sim
        if name == "pic_dev_dd_picard_iter5":
            assert a[15] == len([e for e in self.log if e[0] in ("iter", "noop")])      # 0-based iteration number of the launch
            ctl = self.tensor_at(a[11])[1]
            if int(ctl[0]):
                self.log.append(("noop",))
                return
            xin, xout = self.tensor_at(a[3])[0], self.tensor_at(a[4])
            it = int(sim.stats[3]) + 1
            xout[1][0] = 100 * self.step + it                 # marks the buffer this iteration wrote
            self.log.append(("iter", it, xin, xout[0], a[5] is not None, a[9]))
        elif name == "pic_dev_dd_field_update2":
            ctl = self.tensor_at(a[10])[1]
            if int(ctl[0]):
                return
            _, st, off = self.tensor_at(a[9])
            assert off == 8
            tol, maxiter = a[11], a[12]
            it = int(sim.stats[3]) + 1
            r = self.scripts[self.step][it - 1]
            sim.stats[0] = r; sim.stats[3] = it; st[off + it - 1] = r
            if not (r > tol) or it >= maxiter:
                ctl[0] = 1
        elif name == "pic_dev_dd_step_begin":
            sim.Es.copy_(sim.E0); sim.wall_cum.zero_(); sim.stats.zero_(); sim.ctl.zero_()
        elif name in ("pic_dev_dd_commit_u2", "pic_dev_dd_j1_finish"):
            if name == "pic_dev_dd_commit_u2":
                self.log.append(("repair", self.tensor_at(a[3])[0], self.tensor_at(a[4])[0], a[8]))
        else:
            raise AssertionError(name)


def make_sim(maxiter=6, tol=1e-5, enqueue_ahead=True):
    sim = object.__new__(S.SheathSim)
    f = lambda n: torch.zeros(n, dtype=torch.float64)
    sim.maxiter, sim.tol, sim.elide_u, sim.enqueue_ahead, sim.det = maxiter, tol, True, enqueue_ahead, False
    sim.dev = torch.device("cpu")
    sim.Ng = 4
    sim.fused_moments = False
    sim._begun = False
    sim.p2p = None
    sim.params = S._lib.DDParams()
    for nm in ("x0", "u0", "x1", "x1b", "u1", "E0", "Es", "E1", "Es_prev", "j0", "acc", "wall_cum"):
        setattr(sim, nm, f(4))
    sim.active = torch.ones(4, dtype=torch.int8)
    sim.stats = f(8 + maxiter + 4)
    sim.ctl = torch.zeros(1, dtype=torch.int32)
    sim.dead_buf = torch.zeros(4 + 4 * 16, dtype=torch.int32); sim.dead_cap = 16; sim.oid = None
    sim.range_err = torch.zeros(1, dtype=torch.int32)
    sim.comm = types.SimpleNamespace(world=1, allreduce_sum=lambda t: t)
    sim._ratio = sim._r1 = sim._prev_hist = None
    sim.u_repairs = sim.kernel_launches = 0
    sim.iter_events = sim.resid_trace = None
    return sim


@pytest.fixture
def fake(monkeypatch):
    def install(sim, scripts):
        dev = FakeDevice(sim, scripts)
        monkeypatch.setattr(S._lib, "call", dev.call)
        monkeypatch.setattr(S.D, "stream", lambda: 0)
        monkeypatch.setattr(S.D, "read_f64", lambda t, n=None: t[:n].numpy().copy())
        return dev
    return install


def run_step(sim, dev):
    dev.step += 1
    dev.log.clear()
    names = {id(getattr(sim, n)): n for n in ("x0", "x1", "x1b")}
    k, r = sim.picard()
    return k, r, list(dev.log), names


SCRIPTS = [
    [1e3, 2.0, 4e-3, 8e-6],                # no history: synchronous loop, 4 iterations
    [1e3, 2.0, 4e-3, 8e-6],                # queued ahead: 4, all run
    [1e3, 1e-2, 1e-7],                     # converges after 3: the 4th queued launch is a no-op
    [1e3, 1.0, 1e-2, 1e-4, 1e-6],          # needs 5: 3 queued, 2 more one at a time
    [1e3, 9e2, 8e2, 7e2, 6e2, 5e2],        # never converges: stops at maxiter = 6
    [1e3, 2.0, 4e-3, 8e-6],                # back to normal after a step that hit maxiter
]


def test_enqueue_ahead_host_logic(fake):
    sim = make_sim()
    dev = fake(sim, SCRIPTS)
    for step, script in enumerate(SCRIPTS):
        k, r, log, names = run_step(sim, dev)
        iters = [e for e in log if e[0] == "iter"]
        assert k == len(script) == len(iters) and r == script[-1], (step, k, log)
        assert [e[1] for e in iters] == list(range(1, k + 1))
        assert [e[5] for e in iters] == [1] + [0] * (k - 1)                       # `first` only for iteration 1
        # ping-pong: every iteration reads what the previous one wrote
        for a, b in zip(iters, iters[1:]):
            assert b[2] == a[3] and b[3] == a[2]
        # the commit makes the last executed iteration's output the new x0
        assert int(sim.x0[0]) == 100 * step + k
        assert len({sim.x0.data_ptr(), sim.x1.data_ptr(), sim.x1b.data_ptr()}) == 3
        # repair pass iff the last executed iteration ran light, on that iteration's buffers
        repairs = [e for e in log if e[0] == "repair"]
        last_full = iters[-1][4]
        assert len(repairs) == (0 if last_full else 1)
        if repairs:
            assert repairs[0][1:3] == (iters[-1][2], iters[-1][3]) and repairs[0][3] == (1 if k == 1 else 0)
        assert sim._prev_hist == script and int(sim.ctl[0]) == 1
        if step == 2:
            assert [e[0] for e in log].count("noop") == 1
        if step == 1:
            assert [e[4] for e in iters] == [False, False, False, True]           # l l l F, no repair
        if step == 5:
            # the slow contraction seen in the step that hit maxiter makes the predictor call no
            # iteration "the last": all light, then the repair pass -- slower, never wrong
            assert [e[4] for e in iters] == [False] * 4 and len(repairs) == 1
    assert sim.u_repairs >= 1          # step 2 ended one iteration early on a light one


def test_synchronous_loop_gives_the_same_counts(fake):
    sim = make_sim(enqueue_ahead=False)
    dev = fake(sim, SCRIPTS)
    for step, script in enumerate(SCRIPTS):
        k, r, log, _ = run_step(sim, dev)
        assert k == len(script) and r == script[-1]
        assert not [e for e in log if e[0] == "noop"]
        assert int(sim.x0[0]) == 100 * step + k


def test_raising_maxiter_after_construction_grows_the_residual_history(fake):
    sim = make_sim(maxiter=3)
    dev = fake(sim, [[1e3, 9e2, 8e2], [1e3, 5e2, 1e2, 50.0, 20.0, 10.0, 5.0, 1.0]])
    k, r, _, _ = run_step(sim, dev)
    assert k == 3
    sim.maxiter = 8
    k, r, log, _ = run_step(sim, dev)
    assert k == 8 and r == 1.0 and sim.stats.numel() >= 16 and sim._prev_hist[-1] == 1.0
