"""Row-sharding of the tall matrix across GPUs (one process per GPU, torch.distributed / NCCL).

The reference has no distributed code.  Every O(m n) kernel on the hot path is a
row-wise map followed by a sum of a small result (the d x (n+1) sketch, the (n+1)-vector
[A^T u~ | |u~|^2] per LSQR iteration, n x k / k x k blocks on the low-rank path), so the only
collective needed is an all-reduce(sum) of those small buffers over NVLink/NVSwitch.  Everything
n-sized (R, v, w, x, the LSQR scalars) is replicated and evolves identically on every rank.
"""
import ctypes
import os
import socket

import torch
import torch.distributed as dist


class RowSharded:
    """This rank's block of rows ``[row_offset, row_offset + local.shape[0])`` of an m x n matrix
    (or of an m-vector).  ``shape`` reports the GLOBAL shape so driver-level checks read as in the
    reference."""

    def __init__(self, local, row_offset, m_global, group=None):
        self.local = local
        self.row_offset = int(row_offset)
        self.m_global = int(m_global)
        self.group = group
        self.shape = (self.m_global,) + tuple(local.shape[1:])
        self.ndim = local.dim()
        self.device = local.device
        self.dtype = local.dtype

    @classmethod
    def from_rank(cls, local, group=None):
        """Build from equally sized shards laid out in rank order."""
        ws, rk = dist.get_world_size(group), dist.get_rank(group)
        return cls(local, rk * local.shape[0], ws * local.shape[0], group)


def unwrap(A):
    """-> (local tensor, row_offset, group or None if not sharded)."""
    if isinstance(A, RowSharded):
        return A.local, A.row_offset, (A.group if A.group is not None else dist.group.WORLD)
    return A, 0, None


def allreduce_(t, group):
    """In-place sum over the ranks sharing the rows (no-op for a single GPU)."""
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class PeerComm:
    """Exchange buffers of the fused reduce + cross-GPU sum of the streaming pass (``pla_stream_pass_peer_f64``,
    csrc/stream_pass.cu): one cudaMalloc block per rank, mapped into every other rank of the node with CUDA IPC, so
    the (n + 1)-vector [A^T u~ | |u~|^2] of an LSQR iteration is summed over the GPUs by plain NVLink stores inside
    the reduce kernel instead of a separate NCCL all-reduce.  ``epoch`` counts the calls (identical on all ranks)."""

    MAX_WORLD = 8

    def __init__(self, group, device, slot_lines):
        from . import _lib
        lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        self.slot_lines = int(slot_lines)
        self.device = torch.device(device)
        self.epoch = 0
        self._lib = lib
        self._imported = []
        nbytes = lib.pla_peer_exchange_bytes(self.world, self.slot_lines)
        local = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.pla_peer_alloc(nbytes, ctypes.byref(local)), "pla_peer_alloc")
            self.local = local.value
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.pla_peer_export(self.local, handle), "pla_peer_export")
            mine = (bytes(handle), socket.gethostname())
            everyone = [mine]
            if self.world > 1:
                everyone = [None] * self.world
                dist.all_gather_object(everyone, mine, group=group)
            self.ptrs = (ctypes.c_void_p * self.world)()
            ok = len({h for _, h in everyone}) == 1               # CUDA IPC maps memory of the same node only
            for r, (hb, _) in enumerate(everyone):
                if r == self.rank:
                    self.ptrs[r] = self.local
                elif ok:
                    q = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
                    if lib.pla_peer_import(buf, ctypes.byref(q)) != 0:
                        ok = False
                    else:
                        self.ptrs[r] = q.value
                        self._imported.append(q.value)
            if self.world > 1:                                    # all ranks take the same path
                flags = [None] * self.world
                dist.all_gather_object(flags, bool(ok), group=group)
                ok = all(flags)
        self.ok = ok
        if not ok:
            self.close()

    def next_epoch(self):
        self.epoch = self.epoch % 0xFFFFFFFF + 1        # 1 .. 2^32 - 1: the flag is 32 bits wide and 0 means "never written"
        return self.epoch

    def close(self):
        lib = self._lib
        with torch.cuda.device(self.device):
            for q in self._imported:
                lib.pla_peer_close(q)
            self._imported = []
            if self.local is not None:
                lib.pla_peer_free(self.local)
                self.local = None


_PEER_COMMS = {}


def peer_comm(group, device, slot_lines=8193):
    """The PeerComm of (group, device), created on first use (a collective call: every rank of the group must make
    it); None when the fused exchange does not apply -- not sharded, more than 8 ranks, ranks on different nodes,
    an IPC mapping that failed, or PLA_PEER_ALLREDUCE=0 -- and the caller uses an NCCL all-reduce instead."""
    if group is None or not dist.is_initialized() or os.environ.get("PLA_PEER_ALLREDUCE", "1") == "0":
        return None
    world = dist.get_world_size(group)
    if world < 2 or world > PeerComm.MAX_WORLD or torch.device(device).type != "cuda":
        return None
    key = (id(group), str(device))
    if key not in _PEER_COMMS:
        _PEER_COMMS[key] = PeerComm(group, device, slot_lines)
    comm = _PEER_COMMS[key]
    return comm if comm.ok else None
