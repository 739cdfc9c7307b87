"""Symmetric peer-mapped device buffers (CUDA IPC) for in-kernel NVSwitch P2P exchange.

torch.distributed is only the plumbing that moves the 64-byte IPC handles between the one-process-per-GPU
ranks; the data path afterwards is plain loads/stores on mapped peer pointers inside our kernels.
"""
import ctypes

import torch

from . import _lib


def _host_id():
    import socket
    try:
        return open("/proc/sys/kernel/random/boot_id").read().strip()
    except OSError:
        return socket.gethostname()


def _agree(ok, group, device):
    """True only if EVERY rank of the group reports ok (one tiny all-reduce): a rank that failed to allocate or map a
    peer buffer must not leave the others waiting in a later barrier or spinning in a P2P kernel."""
    import torch.distributed as dist
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class SymmetricBuffer:
    """One cudaMalloc'd buffer per rank, mapped into every peer through CUDA IPC.  Single node, one NVSwitch domain only:
    construction raises SelaviError ON EVERY RANK when any rank cannot allocate / export / map (ranks on different hosts,
    IPC disabled in the container, ...), so callers can fall back to NCCL collectively."""

    def __init__(self, nbytes, group=None):
        import torch.distributed as dist
        lib = _lib.lib()
        self.nbytes = int(nbytes)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._local = ctypes.c_void_p()
        self._opened = []
        self.peer_ptrs = []
        dev = torch.device("cuda", torch.cuda.current_device())
        handle = (ctypes.c_ubyte * 64)()
        err = None
        code = lib.selavi_symm_alloc(self.nbytes, ctypes.byref(self._local), handle)
        if code != 0:
            err = f"selavi_symm_alloc failed with code {code}: {lib.selavi_last_error().decode()}"
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle), err is None, _host_id()), group=group)
        if err is None and len({h[2] for h in handles}) > 1:
            err = "ranks live on different hosts: CUDA IPC peer mapping needs one node"
        if err is None and not all(h[1] for h in handles):
            err = "a peer rank could not allocate its symmetric buffer"
        if err is None:
            for r, (h, _, _) in enumerate(handles):
                if r == self.rank:
                    self.peer_ptrs.append(self._local.value)
                    continue
                p = ctypes.c_void_p()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                code = lib.selavi_symm_open(hb, ctypes.byref(p))
                if code != 0:
                    err = f"selavi_symm_open(rank {r}) failed with code {code}: {lib.selavi_last_error().decode()}"
                    break
                self.peer_ptrs.append(p.value)
                self._opened.append(p.value)
        if not _agree(err is None, group, dev):
            self.close()
            raise _lib.SelaviError(err or "a peer rank could not map the symmetric buffers (P2P unavailable)")

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
    statistic exchange.  Every rank must issue the same sequence of calls (it does: same program, same layer order).

    Flow control: every call is itself a barrier between the ranks (a rank leaves call k only after ALL ranks have pushed
    call k, and a rank pushes call k+1 only after its own call k has finished reading), so no rank runs more than one
    call ahead of the slowest one per stream; a ring slot / flag index is reused only after RING_DOUBLES / (world * n) >
    250 resp. FLAGS = 8192 further calls, far beyond that window."""
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
        try:
            self.flag = SymmetricBuffer(self.FLAGS * self.world * 8, group)
        except _lib.SelaviError:      # raised on every rank (SymmetricBuffer agrees collectively)
            self.recv.close()
            raise
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
