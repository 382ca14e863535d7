"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the static aircraft split, the max-over-ranks
timing and the end-of-run statistics gather that bench.py uses with NCCL on the GPU box (aircraft are
independent, so there is no collective on the data path)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import importlib.util

    import torch.distributed as dist
    # the sharding module alone: importing the package would load libf16_b200.so, which is fine, but is not needed here
    spec = importlib.util.spec_from_file_location("shard", os.path.join(REPO, "f16_mpc_oop_py_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_total = 1001
        r = np.random.default_rng(5)
        x = r.normal(size=(18, n_total))
        st = (r.uniform(size=n_total) < 0.1).astype(np.int32) * 4
        lo, hi = shard.shard_range(n_total, rank, world)
        xs, ss = shard.shard_soa(x, rank, world), st[lo:hi]
        assert xs.flags["C_CONTIGUOUS"] and xs.shape == (18, hi - lo)
        t = shard.max_over_ranks(dist, [10.0 + rank, 5.0 - rank])
        merged = shard.gather_summaries(dist, shard.summarise(xs, ss))
        dist.barrier()
        q.put((rank, lo, hi, t, merged, shard.rank_seed(0xF16, rank)))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_batch_exactly():
    sys.path.insert(0, REPO)
    from f16_mpc_oop_py_b200.shard import shard_range
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gloo_split_and_statistics():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, m0, s0), (r1, lo1, hi1, t1, m1, s1) = res
    assert (lo0, hi1) == (0, 1001) and hi0 == lo1 and s0 != s1
    assert t0 == t1 == [11.0, 5.0]                      # slowest rank's time, identical on every rank
    # merged statistics equal the single-process statistics of the whole batch
    r = np.random.default_rng(5)
    x = r.normal(size=(18, 1001))
    st = (r.uniform(size=1001) < 0.1).astype(np.int32) * 4
    alive = st == 0
    for m in (m0, m1):
        assert m["n"] == 1001 and m["alive"] == int(alive.sum())
        np.testing.assert_allclose(m["mean"], x[:, alive].mean(axis=1), rtol=0, atol=1e-13)
        np.testing.assert_allclose(m["var"], x[:, alive].var(axis=1), rtol=1e-12)
        np.testing.assert_array_equal(m["min"], x[:, alive].min(axis=1))
        np.testing.assert_array_equal(m["max"], x[:, alive].max(axis=1))


def test_library_slices_and_chunks_cover_the_batch():
    """the in-library multi-GPU split (f16_init_devices) and the chunk pipeline, host logic only: slices are contiguous, start
    on multiples of 32 aircraft (a warp-task never straddles two devices), cover [0, N) and are balanced to within one warp-task;
    a batch too small to feed every device uses fewer; chunks are multiples of 32 and at most three slots are in use"""
    import ctypes
    import f16_mpc_oop_py_b200 as f16
    L = f16.lib
    for n, mpd, ctx in ((1 << 20, 16384, 8), (300_001, 16384, 8), (70_001, 16384, 8), (5, 16384, 8), (64 << 20, 16384, 8),
                        (4096 * 3 + 17, 4096, 4), (1, 1, 64), (2_000_000_011, 16384, 3)):
        out = (ctypes.c_longlong * (2 * ctx))()
        k = L.f16_plan_slices(n, mpd, ctx, out)
        assert 1 <= k <= ctx and k == max(1, min(ctx, n // mpd))
        lo = [out[2 * i] for i in range(k)]
        cnt = [out[2 * i + 1] for i in range(k)]
        assert lo[0] == 0 and all(l % 32 == 0 for l in lo) and all(c > 0 for c in cnt)
        assert all(lo[i] + cnt[i] == lo[i + 1] for i in range(k - 1)) and lo[-1] + cnt[-1] == n
        assert max(cnt) - min(cnt) <= 32 + 31
    assert L.f16_plan_slices(0, 16384, 8, (ctypes.c_longlong * 16)()) == 0
    assert L.f16_plan_slices(10, 1, 0, (ctypes.c_longlong * 2)()) == -3
    for n, mc, mx in ((1 << 20, 1 << 18, 4), (1 << 20, 1 << 16, 8), (300_001, 1 << 16, 8), (100, 1 << 16, 8), (1 << 27, 1 << 18, 4)):
        chunk, slots = ctypes.c_longlong(0), ctypes.c_int(0)
        c = L.f16_plan_chunks(n, mc, mx, ctypes.byref(chunk), ctypes.byref(slots))
        assert 1 <= c <= mx and chunk.value % 32 == 0 and slots.value == min(c, 3)
        assert (c - 1) * chunk.value < n <= c * chunk.value and (c == 1 or chunk.value >= mc)
