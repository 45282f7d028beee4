"""Gradient exchange of the ray-sharded step over NVLink / NVSwitch (kernels: csrc/p2p_reduce.cu).

One process per GPU on one node.  Each rank keeps its flat gradients in an arena and a small flag block allocated
through the library; ``all_reduce()`` is then ONE kernel launch per step with no host-side argument that changes, so
it is recorded inside the step's CUDA graph.  ``tensor`` is the arena as a torch tensor: the trainer points the
parameters' ``.grad`` into it, the backward kernels accumulate there, and after the exchange every rank reads the summed
gradients in place.  Two ways of moving the data, chosen at set-up (``algo``):

* ``"nvls"``: the arenas are physical allocations bound to one NVSwitch multicast object; the kernel reduces inside the
  switch (``multimem.ld_reduce`` / ``multimem.st``).  The multicast object's file descriptor travels from rank 0 to the
  other processes over an abstract unix socket (``SCM_RIGHTS``); the existing ``torch.distributed`` group only carries
  the socket's name and the agreement on every step of the set-up.
* ``"peer"``: the arenas are mapped into every process through CUDA IPC and the kernel adds the copies in rank order
  with peer loads / stores (bit-reproducible; the fallback where multicast is unavailable).

``"auto"`` picks NVLS from 4 ranks up and the peer path at 2-3 ranks, where each GPU moves the same bytes either way and
the peer loads are faster (measured on 2 B200: 49.2 MB in 102 us peer, 159 us NVLS, 132 us NCCL; scripts/exchange_probe.py).  A rank that does not reach an exchange within ``timeout_ms`` makes the exchange fail for
good: no data moves, a sticky error is left on the device and in a pinned host word, and ``raise_on_error()`` (called
by ``TrainStep`` every step, without synchronising) raises.
"""
import ctypes
import os
import socket
import uuid

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
    def __init__(self, n_floats, device, group=None, n_ctas=0, algo="auto", timeout_ms=0):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("P2PExchange needs an initialised torch.distributed process group")
        if algo not in ("auto", "nvls", "peer"):
            raise ValueError(f"algo must be 'auto', 'nvls' or 'peer', got {algo!r}")
        self.lib = _lib.load()
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.SNERF_P2P_MAX_RANKS:
            raise RuntimeError(f"P2PExchange: at most {_lib.SNERF_P2P_MAX_RANKS} ranks (one node)")
        self.n_floats = (int(n_floats) + 3) // 4 * 4
        self.n_ctas = int(n_ctas)
        self.algo, self.nvls_error = None, None
        self._opened = []
        self._arena = ctypes.c_void_p()
        self._flags = ctypes.c_void_p()
        self._mc = dict(ptr=ctypes.c_void_p(), arena=ctypes.c_void_p(), mc=ctypes.c_uint64(), mem=ctypes.c_uint64(), bytes=0)
        self.tensor = None
        self.peers = _lib.P2PPeers()
        self.error_host = torch.zeros(1, dtype=torch.int32).pin_memory()  # written by the kernel when a wait runs out
        self.peers.host_error = self.error_host.data_ptr()
        self.peers.timeout_ms = int(timeout_ms)
        with torch.cuda.device(self.device):
            self._setup_flags()
            if algo == "nvls" or (algo == "auto" and self.world >= 4):
                self.nvls_error = self._setup_nvls()
                if self.nvls_error is None:
                    self.algo = "nvls"
                    if self.n_ctas == 0:
                        # 32 CTAs saturate the switch's reduction path; more only add contention (8 B200, 49.2 MB:
                        # 150 us at 32 CTAs, 157 / 173 / 179 us at 64 / 128 / 256 -- scripts/exchange_probe.py)
                        self.n_ctas = 32
                elif algo == "nvls":
                    self._release()
                    raise RuntimeError(f"NVLS exchange unavailable on rank {self.rank}: {self.nvls_error}")
            if self.algo is None:
                self._setup_peer()
                self.algo = "peer"

    # ---- set-up.  Every rank runs the same collectives whatever happens locally (a rank that raised between them would
    # leave the others waiting): local failures are recorded, agreed on with one MIN all-reduce, and handled everywhere.
    def _agree(self, err):
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        return int(ok.item()) == 1

    def _share_ipc(self, ptr, what):
        """export `ptr` (a snerf_p2p_alloc allocation), return the list of all ranks' mappings (own entry: ptr itself)"""
        err, mine = None, b""
        try:
            h = ctypes.create_string_buffer(_lib.SNERF_P2P_HANDLE_BYTES)
            check(self.lib.snerf_p2p_export(ptr, h), f"p2p export ({what})")
            mine = bytes(h.raw)
        except Exception as e:  # noqa: BLE001 -- reported below, on every rank
            err = e
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (self.rank, mine), group=self.group)
        mapped = [None] * self.world
        if err is None:
            try:
                for r, h in everyone:
                    if r == self.rank:
                        mapped[r] = ptr.value
                        continue
                    if len(h) != _lib.SNERF_P2P_HANDLE_BYTES:
                        raise RuntimeError(f"rank {r} exported no {what}")
                    q = ctypes.c_void_p()
                    check(self.lib.snerf_p2p_open(ctypes.create_string_buffer(h, _lib.SNERF_P2P_HANDLE_BYTES), ctypes.byref(q)),
                          f"p2p open ({what}, rank {r})")
                    self._opened.append(q)
                    mapped[r] = q.value
            except Exception as e:  # noqa: BLE001
                err = e
        if not self._agree(err):
            self._release()
            raise RuntimeError(f"peer-memory exchange unavailable on rank {self.rank}: "
                               f"{err if err is not None else 'another rank failed to map the ' + what}")
        return mapped

    def _setup_flags(self):
        err = None
        try:
            check(self.lib.snerf_p2p_alloc(self.lib.snerf_p2p_flag_bytes(), ctypes.byref(self._flags)), "p2p flag alloc")
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            err = e
        if not self._agree(err):
            self._release()
            raise RuntimeError(f"peer-memory exchange unavailable on rank {self.rank}: {err or 'another rank failed'}")
        for r, q in enumerate(self._share_ipc(self._flags, "flag block")):
            self.peers.flags[r] = q

    def _setup_peer(self):
        err = None
        try:
            check(self.lib.snerf_p2p_alloc(self.n_floats * 4, ctypes.byref(self._arena)), "p2p arena alloc")
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            err = e
        if not self._agree(err):
            self._release()
            raise RuntimeError(f"peer-memory exchange unavailable on rank {self.rank}: {err or 'another rank failed'}")
        for r, q in enumerate(self._share_ipc(self._arena, "arena")):
            self.peers.buf[r] = q
        self.peers.mc_buf = None
        self.tensor = torch.as_tensor(_DeviceMemory(self._arena.value, self.n_floats), device=self.device)

    def _setup_nvls(self):
        """Returns None on success, else why NVLS is not used (the same decision on every rank)."""
        lib, m = self.lib, self._mc
        err = None
        gran = 0
        try:
            if not lib.snerf_mc_supported():
                raise RuntimeError("device reports no multicast support")
            gran = int(lib.snerf_mc_granularity(self.world, self.n_floats * 4))
            if gran <= 0:
                raise RuntimeError("multicast granularity query failed")
            m["bytes"] = (self.n_floats * 4 + gran - 1) // gran * gran
            check(lib.snerf_mc_arena_create(m["bytes"], gran, ctypes.byref(m["arena"]), ctypes.byref(m["mem"])), "mc arena")
        except Exception as e:  # noqa: BLE001
            err = e
        if not self._agree(err):
            self._release_mc()
            return str(err) if err is not None else "another rank has no multicast support"
        # rank 0 creates the multicast object; its file descriptor goes to the other processes over a unix socket
        fd, name, server = ctypes.c_int(-1), None, None
        if self.rank == 0:
            try:
                check(lib.snerf_mc_create(self.world, m["bytes"], ctypes.byref(m["mc"]), ctypes.byref(fd)), "mc create")
                name = f"snerf_mc_{os.getpid()}_{uuid.uuid4().hex}"
                server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                server.bind("\0" + name)
                server.listen(self.world)
            except Exception as e:  # noqa: BLE001
                err, name = e, None
        box = [name]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        name = box[0]
        if name is None:
            if server is not None:
                server.close()
            self._release_mc()
            return f"rank 0 could not create the multicast object ({err})" if self.rank == 0 else "rank 0 could not create the multicast object"
        try:
            if self.rank == 0:
                server.settimeout(60)
                for _ in range(self.world - 1):
                    conn, _addr = server.accept()
                    socket.send_fds(conn, [b"m"], [fd.value])
                    conn.close()
            else:
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                c.settimeout(60)
                c.connect("\0" + name)
                _msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
                c.close()
                if not fds:
                    raise RuntimeError("no file descriptor received from rank 0")
                fd = ctypes.c_int(fds[0])
                check(lib.snerf_mc_import(fd.value, ctypes.byref(m["mc"])), "mc import")
            check(lib.snerf_mc_add_device(m["mc"]), "mc add device")
        except Exception as e:  # noqa: BLE001
            err = e
        finally:
            if server is not None:
                server.close()
            if fd.value >= 0:
                os.close(fd.value)  # the driver holds its own reference once exported / imported
        if not self._agree(err):  # also the barrier: every device is part of the team before anybody binds memory
            self._release_mc()
            return str(err) if err is not None else "another rank failed to join the multicast team"
        try:
            check(lib.snerf_mc_bind_and_map(m["mc"], m["mem"], m["bytes"], gran, ctypes.byref(m["ptr"])), "mc bind + map")
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            err = e
        if not self._agree(err):
            self._release_mc()
            return str(err) if err is not None else "another rank failed to bind its arena"
        self.peers.mc_buf = m["ptr"].value
        for r in range(self.world):
            self.peers.buf[r] = None
        self.tensor = torch.as_tensor(_DeviceMemory(m["arena"].value, self.n_floats), device=self.device)
        return None

    def _release_mc(self):
        m = self._mc
        if m["arena"].value or m["mc"].value or m["mem"].value or m["ptr"].value:
            self.lib.snerf_mc_release(m["ptr"], m["arena"], m["mc"], m["mem"], m["bytes"])
        self._mc = dict(ptr=ctypes.c_void_p(), arena=ctypes.c_void_p(), mc=ctypes.c_uint64(), mem=ctypes.c_uint64(), bytes=0)
        self.peers.mc_buf = None

    def _release(self):
        for q in self._opened:
            self.lib.snerf_p2p_close(q)
        self._opened = []
        self.tensor = None
        self._release_mc()
        for a in (self._arena, self._flags):
            if a.value:
                self.lib.snerf_p2p_free(a)
        self._arena, self._flags = ctypes.c_void_p(), ctypes.c_void_p()

    # ---- use
    def all_reduce(self, lo=0, hi=None, channel=0, n_ctas=None):
        """Sum floats [lo, hi) of the arenas of all ranks in place (current stream; capturable in a CUDA graph).
        Calls that may overlap in time (different streams) must use different channels.  ``n_ctas`` overrides the
        exchange's CTA count for this call (an exchange that runs beside another kernel wants fewer)."""
        hi = self.n_floats if hi is None else int(hi)
        if lo % 4 or hi % 4:
            raise ValueError("p2p all_reduce: range bounds must be multiples of 4 floats")
        n_ctas = self.n_ctas if not n_ctas else int(n_ctas)
        check(self.lib.snerf_p2p_allreduce(ctypes.byref(self.peers), self.rank, self.world, int(lo), hi - int(lo), int(channel),
                                           n_ctas, _lib.stream()), "p2p all-reduce")

    def raise_on_error(self):
        """Raises when an exchange of this rank gave up waiting for another rank (the kernel left the gradients untouched
        and wrote 1 + rank into the pinned word).  A plain host read: no synchronisation, safe to call every step."""
        v = int(self.error_host[0])
        if v:
            raise RuntimeError(f"gradient exchange failed: rank {v - 1} waited longer than "
                               f"{self.peers.timeout_ms or 30000} ms for another rank; the gradients of this step were not "
                               "summed and every later exchange is refused -- restart the job")

    def status(self):
        """(completed calls, waits that ran out), summed over the channels -- synchronises the device."""
        calls = waits = 0
        for c in range(_lib.SNERF_P2P_CHANNELS):
            epoch, timeouts = ctypes.c_uint32(), ctypes.c_uint32()
            check(self.lib.snerf_p2p_status(self._flags, c, ctypes.byref(epoch), ctypes.byref(timeouts)), "p2p status")
            calls, waits = calls + int(epoch.value), waits + int(timeouts.value)
        return calls, waits

    def close(self):
        """Unmap the peers and free the arena.  The caller must have dropped every view of ``tensor``."""
        if not (self._flags.value or self._arena.value or self._mc["arena"].value):
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody is still reading this rank's arena
        self._release()
