"""The cosine contraction of one large group sharded over the GPUs of a box (SURVEY.md section 8e;
BASELINE configs[4]: 10^6 x 3000 profiles; reference call site Feature_select_cosine_ami.py:145-149).

Every rank holds a block of rows.  It normalises them and splits them into the two bf16 planes of
the tensor-core kernel (``ips_cosine_split_rows``), the planes are all-gathered once
(``ips_allgather_blocks``: NCCL inside libips.so, rows as blocks -- 6 GB for 10^6 x 3000), every rank
runs the tcgen05 pass over its share of the upper-triangular tile schedule
(``ips_cosine_triu_part``: equal tile counts = equal triangle areas), and the ``world`` partial sums
are exchanged with one 8-byte-per-rank all-gather and added in rank order, so every rank returns the
same float64 sum.
"""
import ctypes as C

import torch

from . import capi
from .plate import _make_comm, _ptr, _stream


class ShardedCosine:
    """Reusable communicator + planes for ``mean_triu`` calls of one shape."""

    def __init__(self, n_local, D, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_local, self.D = int(n_local), int(D)
        self.N = self.n_local * self.world                     # padded row count (zero rows add nothing)
        if self.N < 2:
            raise ValueError("at least two rows are needed")
        self._comm = _make_comm(dist, group, self.rank, self.world) if self.world > 1 else None
        dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = int(capi.call("ips_cosine_planes_bytes", self.N, self.D))
        self.row_bytes = int(capi.call("ips_cosine_plane_row_bytes", self.D))
        self.stride = int(capi.call("ips_cosine_plane_stride_bytes", self.N, self.D))
        buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        off = (-buf.data_ptr()) % 1024
        self.planes = buf[off:off + nbytes]
        self.sums = torch.zeros((self.world,), dtype=torch.float64, device=dev)

    def sum_triu(self, x_local):
        """x_local [n_local][D] float32 (this rank's rows; NaN already replaced by 0) -> float64 sum over
        all i < j of cos(i, j) across every rank's rows, identical on every rank."""
        if x_local.dtype != torch.float32 or tuple(x_local.shape) != (self.n_local, self.D) or not x_local.is_cuda \
                or not x_local.is_contiguous():
            raise ValueError("x_local must be a contiguous float32 CUDA tensor [%d][%d]" % (self.n_local, self.D))
        dev = x_local.device
        base = self.planes.data_ptr()
        with torch.cuda.device(dev):
            st = _stream(dev)
            capi.call("ips_cosine_split_rows", _ptr(x_local), self.n_local, self.rank * self.n_local, self.N, self.D,
                      C.c_void_p(base), st)
            if self.world > 1:
                for p in range(2):                                # one all-gather per plane: rows are the blocks
                    capi.call("ips_allgather_blocks", self._comm, C.c_void_p(base + p * self.stride), self.n_local,
                              self.row_bytes, st)
            capi.call("ips_cosine_triu_part", C.c_void_p(base), C.c_void_p(self.sums.data_ptr() + 8 * self.rank),
                      self.N, self.D, self.rank, self.world, st)
            if self.world > 1:
                capi.call("ips_allgather_blocks", self._comm, _ptr(self.sums), 1, 8, st)
        total = 0.0
        for v in self.sums.cpu().tolist():                        # rank order: every rank adds the same way
            total += v
        return total

    def close(self):
        if self._comm is not None:
            capi.call("ips_comm_destroy", self._comm)
            self._comm = None


def mean_triu(x_local, n_rows_total=None, group=None):
    """Mean of the strict upper triangle of the cosine-similarity matrix of all ranks' rows
    (``cosine_similarity(X)[np.triu_indices(n, 1)].mean()``, Feature_select_cosine_ami.py:145-149).
    ``n_rows_total``: the true row count when the last rank's block is padded with zero rows."""
    sc = ShardedCosine(x_local.shape[0], x_local.shape[1], group)
    try:
        s = sc.sum_triu(x_local)
        n = int(n_rows_total) if n_rows_total is not None else sc.N
        return s / (n * (n - 1) / 2.0)
    finally:
        sc.close()
