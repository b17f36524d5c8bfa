"""One-process-per-GPU communicator: torch.distributed (NCCL) for bootstrap and the baseline collectives, plus a
symmetric peer-memory arena that the fused kernels write into directly over NVLink.

Reference component replaced: NnNetwork::serve/connect + NnNetworkNodeSynchronizer (src/nn/nn-network.cpp:295-632):
root/worker TCP mesh bootstrap and the per-segment sync steps. Here rank 0 is the root, ranks 1..n-1 are the workers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import torch
import torch.distributed as dist

from ..ops import cuda_lib as cl

MAX_CTAS = 256


@dataclass
class ArenaLayout:
    slots_off: int
    flags_off: int
    cand_val_off: int
    cand_idx_off: int
    cand_flag_off: int
    gather_off: int
    prefill_slots_off: int
    prefill_slot_stride: int
    total: int


def arena_layout(n_ranks: int, max_batch: int, dim: int, vocab_full: int, max_prefill: int = 0) -> ArenaLayout:
    def align(x):
        return (x + 255) // 256 * 256
    off = 0
    slots = off; off = align(off + 2 * n_ranks * max_batch * dim * 8)   # LL words: (f32 payload, flag)
    flags = off; off = align(off + 2 * n_ranks * MAX_CTAS * 4)
    cv = off; off = align(off + 8 * 8)                                   # arg-max candidates, one LL word per rank
    ci = off; off = align(off + 64)
    cf = off; off = align(off + 64)
    gather = off; off = align(off + max_batch * vocab_full * 4)
    pslots = off; off = align(off + 2 * n_ranks * max_prefill * dim * 8)
    return ArenaLayout(slots, flags, cv, ci, cf, gather, pslots, max_prefill * dim, off)


class Communicator:
    """Wraps the default process group. `alloc_arena` creates the symmetric buffer and maps every peer's copy."""

    def __init__(self):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one rank per GPU, backend nccl)")
        self.rank = dist.get_rank()
        self.world_size = dist.get_world_size()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = cl.lib()
        self.arena_ptrs: List[int] = []
        self.layout: ArenaLayout | None = None
        self._local = None
        # Peer-memory (CUDA IPC over NVLink) collectives need every rank on one host; otherwise the engine falls back to NCCL.
        import socket
        hosts = [None] * self.world_size
        dist.all_gather_object(hosts, socket.gethostname())
        self.single_node = len(set(hosts)) == 1

    @property
    def is_root(self) -> bool:
        return self.rank == 0

    def alloc_arena(self, layout: ArenaLayout) -> None:
        self.layout = layout
        ptr = C.c_void_p()
        cl.check(self._lib.dl_comm_alloc(layout.total, C.byref(ptr)), "comm_alloc")
        self._local = ptr.value
        handle = (C.c_ubyte * 64)()
        cl.check(self._lib.dl_comm_ipc_handle(ptr, handle), "comm_ipc_handle")
        mine = (bytes(handle), torch.cuda.current_device())
        gathered = [None] * self.world_size
        dist.all_gather_object(gathered, mine)
        self.arena_ptrs = []
        for r, (h, _dev) in enumerate(gathered):
            if r == self.rank:
                self.arena_ptrs.append(self._local)
            else:
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                out = C.c_void_p()
                cl.check(self._lib.dl_comm_ipc_open(buf, C.byref(out)), "comm_ipc_open")
                self.arena_ptrs.append(out.value)
        dist.barrier()

    def comm_ptrs(self, slot_stride: int) -> cl.CommPtrs:
        L = self.layout
        arr = (C.c_void_p * 8)(*([C.c_void_p(p) for p in self.arena_ptrs] + [None] * (8 - len(self.arena_ptrs))))
        return cl.CommPtrs(nRanks=self.world_size, rank=self.rank, maxCtas=MAX_CTAS, slotStride=slot_stride, arena=arr,
                           slotsOff=L.slots_off, flagsOff=L.flags_off, candValOff=L.cand_val_off, candIdxOff=L.cand_idx_off,
                           candFlagOff=L.cand_flag_off, gatherOff=L.gather_off, prefillSlotsOff=L.prefill_slots_off,
                           prefillSlotStride=L.prefill_slot_stride)

    # ---- baseline collectives (NCCL) ----
    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t)
        return t

    def all_gather_cat(self, t: torch.Tensor, dim: int = -1) -> torch.Tensor:
        parts = [torch.empty_like(t) for _ in range(self.world_size)]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts, dim=dim)

    def broadcast_int(self, value: int, src: int = 0) -> int:
        t = torch.tensor([value], dtype=torch.int64, device=self.device)
        dist.broadcast(t, src)
        return int(t.item())

    def barrier(self):
        dist.barrier()
