"""Data-parallel plumbing of the per-ray path: rays are the sharded unit, gradients the one exchange.

The reference wraps the model in nerfstudio's DDP (gfnerf/gf_pipeline.py:136-138) with one process per GPU, each
drawing its own ray batch from seed + rank (scripts/train.py:84); DDP then averages the gradients of the
*registered* parameters -- which silently excludes the hash table (feat_pool is not an nn.Parameter,
gfnerf/hash_3d_anchored.py:26,73-78) and the octree's occupancy votes.  Here every per-step exchange is explicit:

  * gradient buffers: all-reduce(SUM), the mean is taken by gf_adam_step's grad_div;
  * octree votes (int64 adders / marks / visit counts, atomicMax'ed in the reference): all-reduce(MAX), so every
    replica applies the votes one process seeing all rays would have produced and the replicated octrees stay
    bit-identical.

Backend-agnostic on purpose: NCCL over NVLink on the GPU box, gloo in the CPU tests (tests/test_ddp_gloo.py).
On CUDA the gradient reduce runs on a side stream so that it overlaps whatever the compute stream does next that
needs neither the table nor the MLP (the small bucket's reduce runs under the hash scatter; without sampling ahead,
the table's reduce runs under the next batch's ray sampling).
"""
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def rank_seed(base_seed: int, rank: int) -> int:
    """scripts/train.py:84 -- every rank samples its own rays."""
    return int(base_seed) + int(rank)


def shard_slice(n_items: int, rank: int, world: int) -> slice:
    """Contiguous, balanced split of n_items independent units (rays / blocks) over the ranks."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def table_level_rows(local_size: int, level: int, n_levels: int = 16):
    """Rows [lo, hi) of the hash table that level `level` reads and scatters into.  The reference adds the per-level
    offset level * local_size to a pointer to SCALARS (Hash3DAnchored_cuda.cu:38,105), so the window starts at row
    level * local_size / 2 and consecutive levels overlap by half (csrc/hash_common.cuh level_base_row)."""
    lo = (int(level) * int(local_size)) // 2
    return lo, lo + int(local_size)


def table_reduce_ranges(local_size: int, level_group: int, n_levels: int = 16):
    """Row ranges to all-reduce after each group of `level_group` levels has been scattered, levels ascending:
    [(l0, l1, row_lo, row_hi)].  A row is handed over once no later level can still add to it: after levels < l1
    that is every row below level l1's window; the last group takes what is left of the reachable rows.  The ranges
    are disjoint and cover exactly [0, (n_levels - 1) * local_size / 2 + local_size) -- the rows beyond are never
    touched on any rank and need no exchange."""
    out = []
    for l0 in range(0, n_levels, level_group):
        l1 = min(l0 + level_group, n_levels)
        lo = table_level_rows(local_size, l0)[0]
        hi = table_level_rows(local_size, l1)[0] if l1 < n_levels else table_level_rows(local_size, n_levels - 1)[1]
        out.append((l0, l1, lo, hi))
    return out


class FlatBucket:
    """Several small tensors carved out of ONE flat buffer, so that one collective covers all of them
    (MLP + appearance-embedding gradients: ~100 KB, latency-bound as separate messages)."""

    ALIGN = 64   # elements; every view starts 256-byte aligned (the kernels use 16-byte vector accesses)

    def __init__(self, shapes: Sequence[Sequence[int]], dtype=torch.float32, device="cpu"):
        self.shapes = [tuple(int(x) for x in s) for s in shapes]
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        self.padded = [(n + self.ALIGN - 1) // self.ALIGN * self.ALIGN for n in sizes]
        self.offsets = [sum(self.padded[:i]) for i in range(len(sizes))]     # element offset of each view
        self.rebind(torch.zeros(sum(self.padded), dtype=dtype, device=device))

    def rebind(self, flat: torch.Tensor) -> None:
        """Carve the views out of another flat buffer of the same size (peer-mapped memory, peer.PeerExchange)."""
        assert flat.numel() == sum(self.padded)
        self.flat = flat
        self.views: List[torch.Tensor] = []
        for s, o in zip(self.shapes, self.offsets):
            n = int(torch.Size(s).numel())
            self.views.append(self.flat[o:o + n].view(*s))


class GradSync:
    """all-reduce of gradient buffers (SUM, asynchronous on a side stream for CUDA tensors) and of octree votes
    (MAX, in stream order)."""

    def __init__(self, group=None, device: Optional[torch.device] = None):
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=self.device) if (self.cuda and self.world > 1) else None
        self._pending = False

    # -- gradients -----------------------------------------------------------------------------
    def start_sum(self, buffers: Sequence[torch.Tensor]) -> None:
        """Launch all-reduce(SUM) of `buffers` (small ones first).  CUDA: on the comm stream, after everything
        already queued on the current stream; call wait() before consuming.  CPU (gloo): synchronous."""
        if self.world == 1:
            return
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                for b in buffers:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
                    b.record_stream(self.comm_stream)
            self._pending = True
        else:
            for b in buffers:
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)

    def wait(self) -> None:
        """Make the current stream wait for the reduce launched by start_sum()."""
        if self._pending:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
            self._pending = False

    # -- octree votes -----------------------------------------------------------------------------
    def max_(self, buffers: Sequence[torch.Tensor]) -> None:
        if self.world == 1:
            return
        for b in buffers:
            dist.all_reduce(b, op=dist.ReduceOp.MAX, group=self.group)

    # -- parameters -----------------------------------------------------------------------------
    def broadcast_(self, tensors: Sequence[torch.Tensor], src: int = 0) -> None:
        """Identical initial parameters on every rank (what DDP's constructor does)."""
        if self.world == 1:
            return
        for t in tensors:
            dist.broadcast(t, src, group=self.group)
