"""Peer-memory gradient exchange (csrc/p2p_reduce.cu) on ONE device: the W "ranks" are W arenas and
flag blocks in the same memory, played by ONE cooperative launch (SNERF_P2P_EMULATE_RANKS: rank = blockIdx.y) -- kernels
that wait for each other must not be separate launches on one GPU, nothing guarantees that they run at the same time.
The CTAs wait for each other through the flag protocol exactly as across GPUs; what this cannot cover is the IPC mapping
and the NVLS path (tests/test_multi_gpu.py, 2 GPUs).  The oracle is the sum in rank order, which the kernel promises bit
for bit on every rank."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def make_ranks(world, n, cuda, seed):
    from stable_nerf_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    arenas = [torch.randn(n, generator=g).to(cuda) for _ in range(world)]
    flags = [torch.zeros(_lib.load().snerf_p2p_flag_bytes() // 4, dtype=torch.int32, device=cuda) for _ in range(world)]
    peers = _lib.P2PPeers()
    for r in range(world):
        peers.buf[r], peers.flags[r] = arenas[r].data_ptr(), flags[r].data_ptr()
    peers.flags_word = _lib.SNERF_P2P_EMULATE_RANKS
    return arenas, flags, peers


def launch_all(lib, peers, world, n, streams, n_ctas=8, lo=0, channel=0):
    """all `world` ranks in one cooperative launch on streams[0]"""
    from stable_nerf_b200._lib import check
    check(lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, world, lo, n, channel, n_ctas, ctypes.c_void_p(streams[0].cuda_stream)),
          "p2p all-reduce (all ranks)")


def status(lib, flags, channel=0):
    e, t = ctypes.c_uint32(), ctypes.c_uint32()
    assert lib.snerf_p2p_status(ctypes.c_void_p(flags.data_ptr()), channel, ctypes.byref(e), ctypes.byref(t)) == 0
    return e.value, t.value


@pytest.mark.parametrize("world,n", [(2, 4 * 4096), (3, 4 * 1001), (8, 4 * 50000), (4, 4 * 3), (5, 4 * 7777)])
def test_allreduce_is_the_rank_ordered_sum_on_every_rank(world, n, built_lib, cuda):
    arenas, flags, peers = make_ranks(world, n, cuda, seed=world * 7 + n)
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for call in range(3):  # the epoch advances inside the kernel: identical launches, step after step
        want = torch.zeros(n, device=cuda)
        for a in arenas:
            want += a
        launch_all(built_lib, peers, world, n, streams)
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(arenas[r], want), f"call {call}, rank {r}"
            assert status(built_lib, flags[r]) == (call + 1, 0)
        for r, a in enumerate(arenas):  # new "gradients" for the next step, different on every rank
            a.mul_(0.5).add_(float(r + call))


def test_two_ranges_on_two_channels_overlap(built_lib, cuda):
    """Two slices of the arena exchanged by concurrent calls (different streams, different channels), as the step does
    with the fine / coarse halves of the table gradient; the untouched middle stays as it was."""
    world, n = 3, 4 * 6000
    arenas, flags, peers = make_ranks(world, n, cuda, seed=11)
    before = [a.clone() for a in arenas]
    s0 = [torch.cuda.Stream() for _ in range(world)]
    s1 = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    launch_all(built_lib, peers, world, 4 * 2500, s0, lo=4 * 3500, channel=0)
    launch_all(built_lib, peers, world, 4 * 3000, s1, lo=0, channel=1)
    torch.cuda.synchronize()
    want = torch.zeros(n, device=cuda)
    for a in before:
        want += a
    for r in range(world):
        assert torch.equal(arenas[r][:4 * 3000], want[:4 * 3000]) and torch.equal(arenas[r][4 * 3500:], want[4 * 3500:])
        assert torch.equal(arenas[r][4 * 3000:4 * 3500], before[r][4 * 3000:4 * 3500])
        assert status(built_lib, flags[r], 0) == (1, 0) and status(built_lib, flags[r], 1) == (1, 0)


def test_absent_rank_is_fatal_and_visible_not_a_partial_sum(built_lib, cuda):
    """Rank 1 never launches: rank 0's wait runs out (budget 300 ms here, 30 s by default), NO data moves, the error is
    sticky (the next call is refused too) and lands in the registered pinned host word without a synchronising read."""
    arenas, flags, peers = make_ranks(2, 4 * 256, cuda, seed=3)
    launch_all(built_lib, peers, 1, 4 * 256, [torch.cuda.current_stream()])  # world == 1 is a no-op
    torch.cuda.synchronize()
    assert status(built_lib, flags[0]) == (0, 0)
    from stable_nerf_b200._lib import check
    host_error = torch.zeros(1, dtype=torch.int32).pin_memory()
    peers.flags_word = 0  # a real rank: rank 0 alone
    peers.timeout_ms = 300
    peers.host_error = host_error.data_ptr()
    before = [a.clone() for a in arenas]
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for call in range(2):
        check(built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, 2, 0, 4 * 256, 0, 2, s), "p2p")
        torch.cuda.synchronize()
        epoch, timeouts = status(built_lib, flags[0])
        assert timeouts > 0 and int(host_error[0]) == 1
        assert torch.equal(arenas[0], before[0]) and torch.equal(arenas[1], before[1]), "no partial sums"


def test_bad_arguments(built_lib, cuda):
    from stable_nerf_b200 import _lib
    arenas, flags, peers = make_ranks(2, 16, cuda, seed=1)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 2, 2, 0, 16, 0, 0, s) != 0      # rank out of range
    assert built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, 2, 0, 18, 0, 0, s) != 0      # not a multiple of 4
    assert built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, 2, 2, 16, 0, 0, s) != 0      # misaligned offset
    assert built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, 17, 0, 16, 0, 0, s) != 0     # more ranks than a node has
    assert built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, 2, 0, 16, 4, 0, s) != 0      # no such channel
    peers.buf[1] = None
    assert built_lib.snerf_p2p_allreduce(ctypes.byref(peers), 0, 2, 0, 16, 0, 0, s) != 0      # unmapped peer
    assert _lib.SNERF_P2P_MAX_RANKS == 16
