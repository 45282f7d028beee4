"""Peer-memory gradient exchange of the ray-sharded step (kernel: csrc/p2p_reduce.cu).

One process per GPU on one node.  Each rank allocates an arena for its flat gradients and a small flag block through
the library (exportable allocations), sends their 64-byte handles to the other ranks over the existing
``torch.distributed`` group (the only use of the process group: plumbing), and maps the peers' allocations.  From then
on ``all_reduce()`` is one kernel launch per step with no host-side argument that changes, so it is recorded inside
the step's CUDA graph.  ``tensor`` is the arena as a torch tensor: the trainer points the parameters' ``.grad`` into
it, the backward kernels accumulate there, and after the exchange every rank reads the summed gradients in place.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class _DeviceMemory:
    """Exposes a raw device allocation to torch (zero copy) through the CUDA array interface."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class P2PExchange:
    def __init__(self, n_floats, device, group=None, n_ctas=0):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("P2PExchange needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.SNERF_P2P_MAX_RANKS:
            raise RuntimeError(f"P2PExchange: at most {_lib.SNERF_P2P_MAX_RANKS} ranks (one node)")
        self.n_floats = (int(n_floats) + 3) // 4 * 4
        self.n_ctas = int(n_ctas)
        self._opened = []
        self._arena = ctypes.c_void_p()
        self._flags = ctypes.c_void_p()
        # Every rank runs the same collectives whatever happens locally (a rank that raised between them would leave
        # the others waiting): local failures are recorded, agreed on with one MIN all-reduce, and raised everywhere.
        err = None
        mine = [b"", b""]
        with torch.cuda.device(self.device):
            try:
                check(self.lib.snerf_p2p_alloc(self.n_floats * 4, ctypes.byref(self._arena)), "p2p arena alloc")
                check(self.lib.snerf_p2p_alloc(self.lib.snerf_p2p_flag_bytes(), ctypes.byref(self._flags)), "p2p flag alloc")
                torch.cuda.synchronize()
                for k, p in enumerate((self._arena, self._flags)):
                    h = ctypes.create_string_buffer(_lib.SNERF_P2P_HANDLE_BYTES)
                    check(self.lib.snerf_p2p_export(p, h), "p2p export")
                    mine[k] = bytes(h.raw)
            except Exception as e:  # noqa: BLE001 -- reported below, on every rank
                err = e
            everyone = [None] * self.world
            dist.all_gather_object(everyone, (self.rank, mine[0], mine[1]), group=group)
            self.peers = _lib.P2PPeers()
            if err is None:
                try:
                    for r, h_arena, h_flags in everyone:
                        if r == self.rank:
                            self.peers.buf[r], self.peers.flags[r] = self._arena.value, self._flags.value
                            continue
                        if len(h_arena) != _lib.SNERF_P2P_HANDLE_BYTES:
                            raise RuntimeError(f"rank {r} exported no arena")
                        mapped = []
                        for h in (h_arena, h_flags):
                            q = ctypes.c_void_p()
                            check(self.lib.snerf_p2p_open(ctypes.create_string_buffer(h, _lib.SNERF_P2P_HANDLE_BYTES),
                                                          ctypes.byref(q)), f"p2p open (rank {r})")
                            self._opened.append(q)
                            mapped.append(q.value)
                        self.peers.buf[r], self.peers.flags[r] = mapped
                    self.tensor = torch.as_tensor(_DeviceMemory(self._arena.value, self.n_floats), device=self.device)
                except Exception as e:  # noqa: BLE001
                    err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also: everybody has mapped everybody
            if int(ok.item()) == 0:
                self._release()
                raise RuntimeError(f"peer-memory exchange unavailable on rank {self.rank}: "
                                   f"{err if err is not None else 'another rank failed to map the arenas'}")

    def _release(self):
        for q in self._opened:
            self.lib.snerf_p2p_close(q)
        self._opened = []
        self.tensor = None
        for a in (self._arena, self._flags):
            if a.value:
                self.lib.snerf_p2p_free(a)
        self._arena, self._flags = ctypes.c_void_p(), ctypes.c_void_p()

    def all_reduce(self, lo=0, hi=None, channel=0):
        """Sum floats [lo, hi) of the arenas of all ranks in place (current stream; capturable in a CUDA graph).
        Calls that may overlap in time (different streams) must use different channels."""
        hi = self.n_floats if hi is None else int(hi)
        if lo % 4 or hi % 4:
            raise ValueError("p2p all_reduce: range bounds must be multiples of 4 floats")
        check(self.lib.snerf_p2p_allreduce(ctypes.byref(self.peers), self.rank, self.world, int(lo), hi - int(lo), int(channel),
                                           self.n_ctas, _lib.stream()), "p2p all-reduce")

    def status(self):
        """(completed calls, bounded waits that ran out), summed over the channels -- synchronises the device."""
        calls = waits = 0
        for c in range(_lib.SNERF_P2P_CHANNELS):
            epoch, timeouts = ctypes.c_uint32(), ctypes.c_uint32()
            check(self.lib.snerf_p2p_status(self._flags, c, ctypes.byref(epoch), ctypes.byref(timeouts)), "p2p status")
            calls, waits = calls + int(epoch.value), waits + int(timeouts.value)
        return calls, waits

    def close(self):
        """Unmap the peers and free the arena.  The caller must have dropped every view of ``tensor``."""
        if not self._arena.value:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody is still reading this rank's arena
        self._release()
