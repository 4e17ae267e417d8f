"""Peer-memory plumbing of the data-parallel gradient exchange (csrc/peer.cu): buffers every rank of the node can
address directly over NVLink / NVSwitch.

One process per GPU (torchrun).  Each rank `cudaMalloc`s its exchange buffers through the C-ABI (`gf_peer_alloc`; the
torch caching allocator hands out sub-blocks, which CUDA IPC cannot export), publishes their IPC handles through
`torch.distributed` (all_gather_object: plumbing), and maps every peer's buffers (`gf_peer_import`,
cudaIpcMemLazyEnablePeerAccess).  What comes out is, per named buffer, a table of `world` device pointers -- index r
is rank r's copy, mine included -- that the exchange kernels take as `void* const*`.

There is no fallback inside this module: if a handle cannot be exported or imported the constructor raises and the
caller (ddp.GradSync) keeps the NCCL all-reduce path, saying so once on stderr.
"""
import ctypes as C
from typing import Dict, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class _DevMem:
    """A raw device allocation as something `torch.as_tensor` can wrap without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = int(ptr), int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                         "version": 3, "strides": None}


class PeerExchange:
    """Named exchange buffers of equal size on every rank + the two flag arrays of the cross-GPU barrier."""

    def __init__(self, group, device: torch.device, sizes: Dict[str, int]):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        if self.world > 8:
            raise RuntimeError("PeerExchange: at most 8 ranks (one NVSwitch domain)")
        L = _lib.lib()
        self._local: Dict[str, _DevMem] = {}
        self._imported = []
        self._tables: Dict[str, Sequence[int]] = {}
        sizes = dict(sizes)
        sizes["__flags_in"] = 256
        sizes["__flags_out"] = 256
        sizes["__flags_vote"] = 256    # the octree-vote barrier runs on another stream: its own flags and epochs
        with torch.cuda.device(self.device):
            handles = {}
            for name, nbytes in sizes.items():
                nbytes = (int(nbytes) + 255) // 256 * 256
                p = C.c_void_p()
                _lib.check(L.gf_peer_alloc(nbytes, C.byref(p)), "gf_peer_alloc")
                self._local[name] = _DevMem(p.value, nbytes)
                h = (C.c_ubyte * 64)()
                _lib.check(L.gf_peer_export(p, h), "gf_peer_export")
                handles[name] = bytes(h)
            torch.cuda.synchronize(self.device)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handles, group=group)
            for name in sizes:
                table = []
                for r in range(self.world):
                    if r == self.rank:
                        table.append(self._local[name].ptr)
                        continue
                    h = (C.c_ubyte * 64).from_buffer_copy(everyone[r][name])
                    p = C.c_void_p()
                    _lib.check(L.gf_peer_import(h, C.byref(p)), "gf_peer_import")
                    self._imported.append(p.value)
                    table.append(p.value)
                self._tables[name] = table
        self.epoch = 0
        self.vote_epoch = 0
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.any_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group=group)    # every rank has mapped every buffer before anyone uses them

    # ---- views / pointer tables --------------------------------------------------------------------------------
    def tensor(self, name: str, dtype: torch.dtype, shape: Tuple[int, ...], byte_offset: int = 0) -> torch.Tensor:
        """This rank's copy of buffer `name` (or a slice of it) as a torch tensor -- no copy, not owned by torch."""
        mem = self._local[name]
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        assert byte_offset + n <= mem.nbytes, (name, byte_offset, n, mem.nbytes)
        whole = torch.as_tensor(mem, device=self.device)           # uint8 [nbytes]
        return whole[byte_offset:byte_offset + n].view(dtype).view(*shape)

    def ptrs(self, name: str, byte_offset: int = 0):
        """`void* const*` (host array, index = rank) of every rank's copy of `name`, each advanced by byte_offset."""
        return (C.c_void_p * self.world)(*[p + int(byte_offset) for p in self._tables[name]])

    # ---- cross-GPU barrier ---------------------------------------------------------------------------------------
    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def barrier(self, which: str, epoch: int, local_flag: torch.Tensor = None, want_any: bool = False):
        """Enqueue the barrier kernel on the current stream.  which = "in" / "out" (separate flag arrays).  With
        want_any, self.any_flag receives the OR of every rank's local_flag."""
        _lib.check(_lib.lib().gf_peer_barrier(
            self.world, self.rank, int(epoch), self.ptrs("__flags_" + which), _lib.ptr(local_flag),
            _lib.ptr(self.any_flag) if want_any else None, _lib.ptr(self.error), _lib.cur_stream()), "gf_peer_barrier")

    def check(self):
        """Raises if a barrier ever timed out (host sync: call outside the hot loop)."""
        if int(self.error.item()) != 0:
            raise RuntimeError("gfnerf_b200 peer exchange: a cross-GPU barrier timed out (a rank died or fell behind "
                               "by more than 2 s)")

    def close(self):
        L = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._imported:
                L.gf_peer_close(C.c_void_p(p))
            self._imported = []
            for m in self._local.values():
                L.gf_peer_free(C.c_void_p(m.ptr))
            self._local = {}
