"""Device memory that every GPU of one NVSwitch box can address (the reference is single-device).

`PeerBuffer(numel)` gives every rank of a torch.distributed group `numel` floats of its own (`.local`, a torch tensor
over memory from eb_peer_alloc) plus the device addresses under which THIS process sees every rank's block
(`.ptrs[r]`; `.ptrs[rank]` is the local block).  The kernels of csrc/peer.cu and the PEER mode of csrc/bpr_train.cu take
those addresses and load / atomically add rows wherever they live, so the cross-shard row traffic of the training
step rides NVLink inside the kernel instead of going through an all-to-all.

Mapping: CUDA IPC handles (eb_peer_export / eb_peer_open) exchanged with `all_gather_object` — plumbing only.  If that
fails (e.g. a driver without IPC in the container) torch's symmetric memory is tried; if both fail PeerUnavailable is
raised and the caller falls back to the NCCL formulation (parallel.ShardedTable / ReplicatedTableSync).
"""
import ctypes

import torch
import torch.distributed as dist

from ._lib import EbError, check, lib


class PeerUnavailable(RuntimeError):
    pass


class _Raw:
    """Minimal __cuda_array_interface__ carrier so torch can wrap memory it did not allocate."""

    def __init__(self, ptr, numel, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}


def _agree(ok, group, device):
    """True only if every rank succeeded (keeps the ranks on the same code path)."""
    t = torch.tensor([0 if ok else 1], dtype=torch.int32, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item()) == 0


class PeerBuffer:
    def __init__(self, numel, group=None, device=None, method="auto"):
        self.numel = int(numel)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self._own = None           # pointer from eb_peer_alloc (freed in close)
        self._opened = []          # pointers from eb_peer_open
        self._keep = None          # symmetric-memory handle / tensors kept alive
        self.kind = None
        if self.world > 8:
            raise PeerUnavailable("peer-addressed tables span at most 8 GPUs (one NVSwitch box)")
        errors = []
        for m in (("ipc", "symm") if method == "auto" else (method,)):
            try:
                getattr(self, "_init_" + m)()
                self.kind = m if self.world > 1 else "single"
                break
            except PeerUnavailable as e:
                errors.append(f"{m}: {e}")
        if self.kind is None:
            raise PeerUnavailable("; ".join(errors))

    # ---- CUDA IPC through the C ABI
    def _init_ipc(self):
        L = lib()
        nbytes = self.numel * 4
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            ok, err = True, ""
            try:
                check(L.eb_peer_alloc(nbytes, ctypes.byref(p)))
                self._own = p.value
            except EbError as e:
                ok, err = False, str(e)
            handle = ctypes.create_string_buffer(64)
            if ok and self.world > 1:
                try:
                    check(L.eb_peer_export(self._own, handle))
                except EbError as e:
                    ok, err = False, str(e)
            if self.world == 1:
                if not ok:
                    raise PeerUnavailable(err)
                self.ptrs = [self._own]
            else:
                got = [None] * self.world
                dist.all_gather_object(got, (ok, bytes(handle.raw)), group=self.group)
                ptrs = [0] * self.world
                if all(g[0] for g in got):
                    for r, (_, h) in enumerate(got):
                        if r == self.rank:
                            ptrs[r] = self._own
                            continue
                        q = ctypes.c_void_p()
                        try:
                            check(L.eb_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q)))
                            ptrs[r] = q.value; self._opened.append(q.value)
                        except EbError as e:
                            ok, err = False, str(e)
                            break
                else:
                    ok = False; err = err or "a peer could not allocate/export"
                if not _agree(ok, self.group, self.device):
                    self._release_ipc()
                    raise PeerUnavailable(err or "a peer could not map the buffers")
                self.ptrs = ptrs
            self.local = torch.as_tensor(_Raw(self._own, self.numel), device=self.device)

    def _release_ipc(self):
        L = lib()
        for q in self._opened:
            L.eb_peer_close(q)
        self._opened = []
        if self._own:
            L.eb_peer_free(self._own)
            self._own = None

    # ---- torch symmetric memory (CUDA VMM handles exchanged by torch)
    def _init_symm(self):
        if self.world == 1:
            self.local = torch.zeros(self.numel, device=self.device)
            self.ptrs = [self.local.data_ptr()]
            return
        ok, err = True, ""
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(self.numel, dtype=torch.float32, device=self.device)
            h = symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            t.zero_()
            self.local, self.ptrs, self._keep = t, [int(p) for p in h.buffer_ptrs], (t, h)
        except Exception as e:                                          # noqa: BLE001
            ok, err = False, repr(e)
        if not _agree(ok, self.group, self.device):
            raise PeerUnavailable(err or "a peer could not rendezvous")

    def ptr_array(self, offset_elems=0):
        """ctypes array of the ranks' base addresses (+ offset in floats) — the `*_shards` argument of the C ABI."""
        return (ctypes.c_void_p * self.world)(*[p + 4 * int(offset_elems) for p in self.ptrs])

    def barrier(self):
        """All ranks' previously enqueued work on the current stream is finished and visible (before peers read, or
        before the memory goes away)."""
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self):
        """Collective: nobody may touch the peers' memory afterwards."""
        self.barrier()
        self.local = None
        if self.kind in ("ipc", "single") and (self._own or self._opened):
            self._release_ipc()
        self._keep = None
        if self.world > 1:
            dist.barrier(group=self.group)
