"""Final gather of a sharded batch over NVLink peer memory (collective C1).

One process per GPU.  The gathering rank owns the destination buffer in its HBM; every other rank maps it
(`lsr_peer_export` / `lsr_peer_open`, a CUDA IPC handle carried by whatever transport the job already has) and hands
`slice_ptr` to the commitment entry points as their output pointer.  The fused commitment kernel then stores its
container rows straight into the gathering rank's memory through NVLink / NVSwitch: the gather is the kernel's own
epilogue, not a collective that follows it, and the containers are written exactly once.

What bounds it: the gathering GPU's NVLink ingress (900 GB/s nominal per direction on B200).  A container is 64 KiB, so
one sink absorbs at most ~13.7 M commitments/s whatever the number of producers; a job that needs more than that has to
gather digests (32 bytes per commitment, `digest_all_gather`) and leave the containers with their owners.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

from . import capi


class PeerGather:
    """rank 0 holds `world * slice_bytes` bytes; rank r writes its slice at offset r * slice_bytes.

    `bcast(obj_or_None) -> obj` broadcasts a small Python object from rank 0 (e.g. a wrapper around
    torch.distributed.broadcast_object_list); it is the only communication this class does itself.
    """

    def __init__(self, rank: int, world: int, slice_bytes: int, bcast: Callable[[object], object]):
        lib = capi.load()
        self._lib = lib
        self.rank, self.world, self.slice_bytes = rank, world, int(slice_bytes)
        self._own = None
        self._mapped = None
        if rank == 0:
            self._own = lib.lsr_device_alloc(self.slice_bytes * world)
            if not self._own:
                raise MemoryError("lsr_device_alloc failed: " + lib.lsr_last_error().decode())
            h = C.create_string_buffer(64)
            if lib.lsr_peer_export(self._own, h) != 0:
                raise RuntimeError("lsr_peer_export failed: " + lib.lsr_last_error().decode())
            bcast(bytes(h.raw))
            self.base = self._own
        else:
            raw = bcast(None)
            self._mapped = lib.lsr_peer_open(C.create_string_buffer(raw, 64))
            if not self._mapped:
                raise RuntimeError("lsr_peer_open failed (no peer path between the devices?): " + lib.lsr_last_error().decode())
            self.base = self._mapped

    @property
    def slice_ptr(self) -> int:
        """Device address (valid in THIS process) of this rank's slice of the gathered buffer."""
        return int(self.base) + self.rank * self.slice_bytes

    def close(self) -> None:
        if self._mapped:
            self._lib.lsr_peer_close(self._mapped)
            self._mapped = None
        if self._own:
            self._lib.lsr_device_free(self._own)
            self._own = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def torch_bcast(group=None) -> Callable[[object], object]:
    """bcast callable over torch.distributed (any backend)."""
    import torch.distributed as dist

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0, group=group)
        return box[0]
    return bcast


class _RawDeviceWords:
    def __init__(self, ptr: int, words: int):
        self.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def device_view(ptr: int, words: int, device=None):
    """torch int64 view (no copy) of `words` 64-bit words of plain device memory at `ptr` (e.g. the gathered buffer)."""
    import torch
    return torch.as_tensor(_RawDeviceWords(ptr, words), device=device or torch.device("cuda", torch.cuda.current_device()))
