"""Peer-memory accumulators for the particle decomposition (one process per GPU): every rank owns a
buffer (its grid accumulators + an inbox with one slot per rank) that all ranks map through CUDA IPC;
the field kernel pushes the accumulators into every rank's inbox over NVLink and sums its own inbox
(pic_dev_dd_field_update_p2p) instead of calling a library all-reduce.  torch.distributed is only used
to exchange the 64-byte IPC handles at start-up."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class PeerAccumulators:
    def __init__(self, comm, nacc, device):
        self.comm, self.nacc = comm, int(nacc)
        self.rank, self.world = comm.rank, comm.world
        torch.cuda.set_device(device)        # pic_p2p_alloc allocates on the current device
        mine, handle = C.c_void_p(), (C.c_char * 64)()
        _lib.call("pic_p2p_alloc", self.nacc, self.world, C.byref(mine), handle)
        self.mine = mine.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=comm.group)
        self.ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.mine)
                continue
            q = C.c_void_p()
            _lib.call("pic_p2p_open", (C.c_char * 64).from_buffer_copy(h), C.byref(q))
            self.ptrs.append(q.value)
            self._opened.append(q.value)
        self.peers_dev = torch.tensor(self.ptrs, dtype=torch.int64, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.seq = 0                         # one per reduction launched; identical on every rank
        torch.cuda.synchronize(device)
        comm.barrier()                       # every rank has mapped every buffer before the first kernel uses them

    def next_seq(self):
        self.seq = (self.seq + 1) & 0xffffffff
        return self.seq

    def check(self):
        if int(self.err.item()):
            self.err.zero_()
            raise _lib.PicError(_lib.PIC_ERR_CUDA, "peer-memory reduction timed out waiting for another rank")

    def close(self):
        """Collective: no rank may free its buffer while another still has it mapped."""
        if self.mine is None:
            return
        torch.cuda.synchronize()
        for q in self._opened:
            _lib.call("pic_p2p_close", q)
        self._opened = []
        self.comm.barrier()
        _lib.call("pic_p2p_free", self.mine)
        self.mine = None
