"""Row-sharding of the tall matrix across GPUs (one process per GPU, torch.distributed / NCCL).

The reference has no distributed code.  Every O(m n) kernel on the hot path is a
row-wise map followed by a sum of a small result (the d x (n+1) sketch, the (n+1)-vector
[A^T u~ | |u~|^2] per LSQR iteration, n x k / k x k blocks on the low-rank path), so the only
collective needed is an all-reduce(sum) of those small buffers over NVLink/NVSwitch.  Everything
n-sized (R, v, w, x, the LSQR scalars) is replicated and evolves identically on every rank.
"""
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
