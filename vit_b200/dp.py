"""Data parallelism for the ViT step: one process per GPU, parameters replicated, the batch sharded by
rank, ONE exchange step -- the gradient all-reduce (mean) over the flat gradient arena
(reference: Lightning DDP via strategy='ddp', src/hardware_utils.py:86-95, src/basemodule.py:226-241).

The arena is laid out so each layer's gradients are contiguous; backward produces buckets in the
order head -> layer L-1 -> ... -> layer 0 -> embeddings, and each bucket's all-reduce is issued as soon
as its kernels are enqueued, so NCCL (NVLink/NVSwitch) overlaps the remaining backward kernels.
`vit.pooler.dense.*` lies outside every bucket: it never has a gradient, which is
exactly what makes stock DDP raise on the reference model; here it is simply not exchanged.

These helpers are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def backward_bucket_order(buckets: Sequence[Tuple[str, int, int]]) -> List[Tuple[str, int, int]]:
    """ParamLayout.buckets is in forward order (embeddings, layer0.., head); backward finishes them reversed."""
    return list(reversed(list(buckets)))


def allreduce_bucket(flat_grad: torch.Tensor, start: int, end: int, group=None, async_op: bool = True):
    """SUM all-reduce of flat_grad[start:end] in place; the 1/world mean factor is folded into the
    optimizer kernel's grad_scale, so no extra pass over the gradients is needed."""
    if end <= start:
        return None
    return dist.all_reduce(flat_grad[start:end], op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def allreduce_all(flat_grad: torch.Tensor, buckets: Sequence[Tuple[str, int, int]], group=None) -> None:
    works = [allreduce_bucket(flat_grad, s, e, group=group, async_op=True) for _, s, e in backward_bucket_order(buckets)]
    for w in works:
        if w is not None:
            w.wait()


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the default group if needed."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def broadcast_parameters(flat_data: torch.Tensor, group=None, src: int = 0) -> None:
    """Replicate rank `src`'s parameter arena (DDP's constructor does the same)."""
    dist.broadcast(flat_data, src=src, group=group)


class PeerUnavailable(RuntimeError):
    """The ranks cannot map each other's memory (different nodes, P2P disabled): use the NCCL all-reduce."""


class PeerExchange:
    """Per-rank exchange buffers for the in-kernel gradient all-reduce (vitb200_clip_adamw_fused_dp).

    Every rank cudaMalloc's one buffer, the 64-byte CUDA IPC handles are all-gathered (host side, once), and each rank
    maps its peers' buffers (NVLink / NVSwitch peer access inside one node, one process per GPU).  `table` is the device
    array of the `world` buffer pointers in rank order that the kernel receives.  There is no NCCL call per step."""

    def __init__(self, n_floats: int, device: torch.device, group=None):
        import ctypes

        from . import _lib

        self.lib = _lib.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:    # the exchange buffers hold 8 source ranks (one NVLink / NVSwitch node)
            raise PeerUnavailable("the in-kernel exchange covers up to 8 ranks of one node")
        nbytes = int(self.lib.vitb200_peer_buffer_bytes(n_floats))
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(self.lib.vitb200_peer_alloc(nbytes, ctypes.byref(own), handle), "peer_alloc")
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.opened = []
        ptrs = []
        ok = 1
        for q, h in enumerate(handles):
            if q == self.rank:
                ptrs.append(self.own)
                continue
            p = ctypes.c_void_p()
            if self.lib.vitb200_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)) != 0:
                ok = 0      # no CUDA IPC / peer access to that rank (another node, or P2P disabled)
                break
            self.opened.append(p.value)
            ptrs.append(p.value)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag[0]) == 0:   # every rank takes the same decision: the caller falls back to the NCCL all-reduce
            for p in self.opened:
                self.lib.vitb200_peer_close(p)
            self.lib.vitb200_peer_free(self.own)
            self.own, self.opened = None, []
            raise PeerUnavailable("CUDA IPC peer mapping failed on at least one rank")
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        # the optimizer kernel's barrier / launch-count words live and die with the exchange buffers: the flag protocol
        # compares peer flags with the local launch count, so a re-zeroed workspace next to surviving buffers (or the
        # reverse) would let a wait pass on stale flags
        self.tail_ws = torch.zeros(int(self.lib.vitb200_clip_adamw_fused_ws_bytes()), dtype=torch.uint8, device=device)
        self.group = group
        torch.cuda.synchronize(device)
        dist.barrier(group=group)   # every rank has mapped every buffer before the first launch touches them

    @staticmethod
    def check(state: torch.Tensor) -> None:
        """Raise if an optimizer launch gave up waiting for a peer (state[7] is set by the kernel; one host read --
        call it at epoch / checkpoint boundaries, not per step)."""
        if float(state[7]) != 0.0:
            raise RuntimeError("vit_b200: a data-parallel optimizer launch timed out waiting for a peer rank's gradients; "
                               "the replicas are no longer consistent")

    def close(self) -> None:
        if self.own is None:
            return
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)   # nobody unmaps while a peer's last kernel may still read
        for p in self.opened:
            self.lib.vitb200_peer_close(p)
        self.lib.vitb200_peer_free(self.own)
        self.own, self.opened = None, []
