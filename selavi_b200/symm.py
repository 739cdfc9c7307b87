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
