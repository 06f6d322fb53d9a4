"""Peer-memory transport for the expert-parallel exchange (csrc/peer.cu): one process per GPU on one NVSwitch box, every
rank maps the staging buffers of all ranks (CUDA IPC through torch's storage sharing) and the equal-split all-to-all /
all-gather become [copy into my staging buffer] -> [device-side barrier] -> [pull my segments from every peer over
NVLink].  No NCCL call and no host state on the data path, so the whole expert-parallel train step records into ONE CUDA
graph per half like the data-parallel step.  torch.distributed is used once per buffer, at allocation time, to exchange
the IPC handles.

Buffer reuse: every call site owns its staging buffer (key), and the layer issues at least one more barrier on the same
stream before a buffer is written again (dispatch -> combine -> ... -> next step), which a rank only passes after its
peers have finished the pulls they issued before signalling; no second barrier per exchange is needed."""
import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib as L


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PeerGroup:
    """Symmetric buffers + device-side barrier over the ranks of `group` (all on one node, peer access available)."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._keep: List = []                    # opened peer storages must stay alive as long as their pointers are used
        self._bufs: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.flags = torch.zeros(64, dtype=torch.int32, device=self.device)
        self.epoch = torch.zeros(2, dtype=torch.int32, device=self.device)     # [barriers passed, sticky failure flag]
        torch.cuda.synchronize()
        self.flag_ptrs = self._share(self.flags)
        dist.barrier(group=group)                # every rank's flags are zero and mapped before anyone signals
        torch.cuda.synchronize()

    # -------------------------------------------------------------------------------------------- mapping
    def _share(self, t: torch.Tensor) -> torch.Tensor:
        """Map `t` (this rank's buffer) into every peer; returns the device int64 table of all ranks' base addresses."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("hdmoe_b200.peer: staging buffers must be created before stream capture (run a warm-up step)")
        st = t.untyped_storage()
        handle = st._share_cuda_()
        off = t.storage_offset() * t.element_size()
        mine = (handle, off)
        everyone: List = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        ptrs = []
        for r, (h, o) in enumerate(everyone):
            if r == self.rank:
                ptrs.append(t.data_ptr())
                continue
            # opened under THIS rank's device (not the sender's index): cudaIpcOpenMemHandle maps the peer allocation into
            # the current context with lazy peer access, so kernels running here dereference it directly over NVLink
            peer = torch.UntypedStorage._new_shared_cuda(self.device.index, *h[1:])
            self._keep.append(peer)
            ptrs.append(peer.data_ptr() + o)
        return torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def staging(self, key, shape, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
        """(this rank's staging buffer, device table of every rank's buffer address) for a call site; created on first
        use (collective: every rank reaches the same call sites in the same order)."""
        k = (key, tuple(shape), dtype)
        hit = self._bufs.get(k)
        if hit is None:
            buf = torch.empty(shape, dtype=dtype, device=self.device)
            hit = (buf, self._share(buf))
            self._bufs[k] = hit
        return hit

    # -------------------------------------------------------------------------------------------- device ops
    def barrier(self) -> None:
        L.check(L.lib().hdmoe_peer_barrier(C.c_void_p(self.flags.data_ptr()), C.c_void_p(self.flag_ptrs.data_ptr()),
                                           C.c_void_p(self.epoch.data_ptr()), self.rank, self.world, _st()), "peer_barrier")

    def check(self) -> None:
        """Raise if a barrier timed out since the group was created (synchronises)."""
        if int(self.epoch[1]):
            raise RuntimeError("hdmoe_b200.peer: a device-side barrier timed out (a peer rank did not arrive within 8 s); "
                               "the exchanged data of this group is invalid")

    def exchange(self, x: torch.Tensor, key, gather: bool = False) -> torch.Tensor:
        """gather = False: equal-split all-to-all of x [world * C, ...] (segment g of x goes to rank g; segment s of the
        result came from rank s).  gather = True: all-gather of x -> [world, *x.shape]."""
        x = x.contiguous()
        buf, table = self.staging(key, x.shape, x.dtype)
        buf.copy_(x)
        self.barrier()
        nbytes = x.numel() * x.element_size()
        if gather:
            seg, out = nbytes, torch.empty((self.world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        else:
            if x.shape[0] % self.world:
                raise ValueError("peer all-to-all: the leading dimension must be a multiple of the world size")
            seg, out = nbytes // self.world, torch.empty_like(x)
        if seg % 16:
            raise ValueError("peer exchange: segments must be multiples of 16 bytes")
        L.check(L.lib().hdmoe_peer_pull(C.c_void_p(out.data_ptr()), C.c_void_p(table.data_ptr()), seg,
                                        -1 if gather else self.rank, self.world, _st()), "peer_pull")
        return out


_GROUPS: Dict[int, PeerGroup] = {}


def peer_group(group=None) -> PeerGroup:
    k = id(group) if group is not None else 0
    pg = _GROUPS.get(k)
    if pg is not None and (pg.world != dist.get_world_size(group) or pg.rank != dist.get_rank(group)):
        pg = None                                # a new process group re-used the key: map afresh
    if pg is None:
        pg = _GROUPS[k] = PeerGroup(group)
    return pg


class _PeerAllToAll(torch.autograd.Function):
    """Equal-split all-to-all over peer memory; its backward is the same exchange of the gradient (own staging buffer)."""

    @staticmethod
    def forward(ctx, x, pg, key):
        ctx.pg, ctx.key = pg, key
        return pg.exchange(x, ("f",) + tuple(key))

    @staticmethod
    def backward(ctx, g):
        return ctx.pg.exchange(g, ("b",) + tuple(ctx.key)), None, None


def all_to_all_equal(x: torch.Tensor, key, group=None) -> torch.Tensor:
    return _PeerAllToAll.apply(x, peer_group(group), key)


def all_gather(x: torch.Tensor, key, group=None) -> torch.Tensor:
    """[world, *x.shape]; no gradient (used for the per-expert counts)."""
    return peer_group(group).exchange(x.detach(), ("g",) + tuple(key), gather=True)


def any_failed() -> bool:
    """True if a barrier of any group of this process timed out (synchronises)."""
    return any(bool(int(pg.epoch[1])) for pg in _GROUPS.values())


def check_all() -> None:
    """PeerGroup.check() for every group of this process."""
    for pg in _GROUPS.values():
        pg.check()
