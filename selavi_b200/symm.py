"""Symmetric peer-mapped device buffers (CUDA IPC) for in-kernel NVSwitch P2P exchange.

torch.distributed is only the plumbing that moves the 64-byte IPC handles between the one-process-per-GPU
ranks; the data path afterwards is plain loads/stores on mapped peer pointers inside our kernels.
"""
import ctypes

import torch

from . import _lib


class SymmetricBuffer:
    def __init__(self, nbytes, group=None):
        import torch.distributed as dist
        lib = _lib.lib()
        self.nbytes = int(nbytes)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._local = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(lib.selavi_symm_alloc(self.nbytes, ctypes.byref(self._local), handle), "selavi_symm_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_ptrs = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peer_ptrs.append(self._local.value)
                continue
            p = ctypes.c_void_p()
            hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            _lib.check(lib.selavi_symm_open(hb, ctypes.byref(p)), "selavi_symm_open")
            self.peer_ptrs.append(p.value)
            self._opened.append(p.value)
        dist.barrier(group)

    def zero_(self):
        _lib.check(_lib.lib().selavi_symm_memset(self._local, 0, self.nbytes, _lib.stream_ptr()), "selavi_symm_memset")

    def close(self):
        lib = _lib.lib()
        for p in self._opened:
            lib.selavi_symm_close(ctypes.c_void_p(p))
        self._opened = []
        if self._local:
            torch.cuda.synchronize()
            lib.selavi_symm_free(self._local)
            self._local = ctypes.c_void_p()


class P2PAllReduce:
    """Sum all-reduce of small float64 vectors through peer-mapped rings (one tiny kernel, no NCCL): the SyncBatchNorm
    statistic exchange.  Every rank must issue the same sequence of calls (it does: same program, same layer order)."""
    RING_DOUBLES = 4 << 20       # 32 MB receive ring per rank
    FLAGS = 8192

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise _lib.SelaviError("P2PAllReduce supports one NVSwitch domain (world <= 8)")
        self.recv = SymmetricBuffer(self.RING_DOUBLES * 8, group)
        self.flag = SymmetricBuffer(self.FLAGS * self.world * 8, group)
        self._recv_arr = (ctypes.c_void_p * self.world)(*self.recv.peer_ptrs)
        self._flag_arr = (ctypes.c_void_p * self.world)(*self.flag.peer_ptrs)
        self.seq = 0
        self.off = 0

    def allreduce_(self, t):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError("P2PAllReduce needs a contiguous float64 CUDA tensor")
        n = t.numel()
        need = self.world * n
        if need > self.RING_DOUBLES // 4:
            raise ValueError("vector too large for the peer ring")
        if self.off + need > self.RING_DOUBLES:
            self.off = 0
        self.seq += 1
        _lib.check(_lib.lib().selavi_p2p_allreduce_f64(_lib.ptr(t), n, self.world, self.rank, self._recv_arr, self._flag_arr,
                                                       self.off, self.seq % self.FLAGS, self.seq, _lib.stream_ptr()),
                   "selavi_p2p_allreduce_f64")
        self.off += need
        return t
