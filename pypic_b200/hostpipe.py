"""End-to-end stepping with HOST buffers for a COUPLED multi-GPU run.

`pic_host_dd_step_batches` (C ABI) advances independent batches on one GPU.  A particle-decomposed
run over several GPUs is one coupled system -- every Picard iteration sums the ranks' grid
accumulators -- so its host-buffer path goes through the resident driver (SheathSim), whose Picard
loop holds the reduction: per step this rank's shard of x0,u0 (+E0) is copied in from pinned host
memory, the whole coupled step runs, and x1,u1,flags (+E1,j1) are copied back out.  Three streams
and two staging slots pipeline the copies around the compute exactly like the C function does:
upload of step b+1 and download of step b-1 overlap the Picard loop of step b.
"""
import torch

from . import _lib, device as D


class HostPipelinedSheath:
    def __init__(self, sim):
        self.sim = sim
        dev = sim.dev
        n = max(sim.N, 1)
        self.h2d, self.d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        f = lambda k, dt=torch.float64: torch.empty(k, dtype=dt, device=dev)
        self.stage_in = [dict(x=f(n), u=f(n), E=f(sim.Ng)) for _ in range(2)]
        self.stage_out = [dict(x=f(n), u=f(n), a=f(n, torch.int8), E=f(sim.Ng), j=f(sim.Ng)) for _ in range(2)]
        self.ev_up = [torch.cuda.Event() for _ in range(2)]
        self.ev_in_free = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.ev_out_free = [torch.cuda.Event() for _ in range(2)]

    def run(self, batches, outs):
        """batches: list of dicts of PINNED host tensors x0,u0,E0 (this rank's shard); outs: list of
        dicts of pinned x1,u1,act,E1,j1 (len >= min(len(batches), 2), reused round-robin).  Returns the
        Picard iteration count of every step."""
        sim, n = self.sim, self.sim.N
        cmp_stream = torch.cuda.current_stream(sim.dev)
        iters = []

        def upload(b):
            s = b & 1
            with torch.cuda.stream(self.h2d):
                if b >= 2:
                    self.h2d.wait_event(self.ev_in_free[s])
                st = self.stage_in[s]
                st["x"][:n].copy_(batches[b]["x0"], non_blocking=True)
                st["u"][:n].copy_(batches[b]["u0"], non_blocking=True)
                st["E"].copy_(batches[b]["E0"], non_blocking=True)
                self.ev_up[s].record(self.h2d)
        upload(0)
        for b in range(len(batches)):
            s = b & 1
            if b + 1 < len(batches):
                upload(b + 1)
            cmp_stream.wait_event(self.ev_up[s])
            st = self.stage_in[s]
            _lib.call("pic_dev_copy", D.ptr(sim.x0), D.ptr(st["x"]), n * 8, D.stream())
            _lib.call("pic_dev_copy", D.ptr(sim.u0), D.ptr(st["u"]), n * 8, D.stream())
            _lib.call("pic_dev_copy", D.ptr(sim.E0), D.ptr(st["E"]), sim.Ng * 8, D.stream())
            self.ev_in_free[s].record(cmp_stream)
            sim.active.fill_(1)
            sim._log_valid = False
            k, r = sim.picard()                       # the coupled step: one reduction over the ranks per iteration
            iters.append(k)
            so = self.stage_out[s]
            if b >= 2:
                cmp_stream.wait_event(self.ev_out_free[s])
            _lib.call("pic_dev_copy", D.ptr(so["x"]), D.ptr(sim.x0), n * 8, D.stream())
            _lib.call("pic_dev_copy", D.ptr(so["u"]), D.ptr(sim.u0), n * 8, D.stream())
            _lib.call("pic_dev_copy", D.ptr(so["a"]), D.ptr(sim.active), n, D.stream())
            _lib.call("pic_dev_copy", D.ptr(so["E"]), D.ptr(sim.E0), sim.Ng * 8, D.stream())
            _lib.call("pic_dev_copy", D.ptr(so["j"]), D.ptr(sim.j0), sim.Ng * 8, D.stream())
            self.ev_done[s].record(cmp_stream)
            o = outs[b % len(outs)]
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self.ev_done[s])
                o["x1"].copy_(so["x"][:n], non_blocking=True)
                o["u1"].copy_(so["u"][:n], non_blocking=True)
                o["act"].copy_(so["a"][:n], non_blocking=True)
                o["E1"].copy_(so["E"], non_blocking=True)
                o["j1"].copy_(so["j"], non_blocking=True)
                self.ev_out_free[s].record(self.d2h)
        self.d2h.synchronize()
        return iters
