"""Host logic of the sharded build on the CPU: the splitter / exchange plan the ranks derive from their
per-bin suffix counts (esa_shard_plan, the function the CUDA path calls), alone and across two real processes that
talk over gloo and play the exchange with numpy arrays instead of GPU buffers."""
import ctypes
import os
import sys

import numpy as np
import pytest

import vstree_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan(lib, hist, me):
    """hist: (nranks, 256) uint64.  Returns owner[256], send_cnt, recv_cnt, start[256]."""
    nranks = hist.shape[0]
    h = np.ascontiguousarray(hist, np.uint64)
    owner = np.zeros(256, np.uint8)
    send, recv = np.zeros(nranks, np.uint64), np.zeros(nranks, np.uint64)
    start = np.zeros(256, np.uint32)
    rc = lib.esa_shard_plan(h.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(256), nranks, me,
                            owner.ctypes.data_as(ctypes.c_void_p), send.ctypes.data_as(ctypes.c_void_p),
                            recv.ctypes.data_as(ctypes.c_void_p), start.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return owner, send, recv, start


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(vstree_b200.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    raw = ctypes.CDLL(vstree_b200.library_path())
    raw.esa_shard_plan.restype = ctypes.c_int
    return raw


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
@pytest.mark.parametrize("shape", ["uniform", "skewed", "one_bin", "empty", "sparse"])
def test_plan_properties(lib, nranks, shape):
    rng = np.random.default_rng(nranks * 17 + len(shape))
    if shape == "uniform":
        hist = rng.integers(900, 1100, size=(nranks, 256)).astype(np.uint64)
    elif shape == "skewed":
        hist = (rng.pareto(0.7, size=(nranks, 256)) * 100).astype(np.uint64)
    elif shape == "one_bin":
        hist = np.zeros((nranks, 256), np.uint64)
        hist[:, 77] = 5000
    elif shape == "empty":
        hist = np.zeros((nranks, 256), np.uint64)
    else:
        hist = np.zeros((nranks, 256), np.uint64)
        hist[:, rng.choice(256, 9, replace=False)] = rng.integers(1, 50, size=(nranks, 9)).astype(np.uint64)
    plans = [plan(lib, hist, me) for me in range(nranks)]
    owner = plans[0][0]
    total = hist.sum(axis=0)
    assert all(np.array_equal(p[0], owner) for p in plans), "ranks disagree on the owners"
    assert np.all(np.diff(owner.astype(int)) >= 0) and owner.max() < nranks, "owners must be non-decreasing"
    for me, (_, send, recv, start) in enumerate(plans):
        for r in range(nranks):
            assert send[r] == hist[me, owner == r].sum()
            assert recv[r] == plans[r][1][me], "what I expect from r is what r sends me"
    # the runs (sender, bin) tile every owner's buffer exactly: bin-major, then sender
    for o in range(nranks):
        cursor = 0
        for b in np.flatnonzero(owner == o):
            for sender in range(nranks):
                assert plans[sender][3][b] == cursor
                cursor += int(hist[sender, b])
        assert cursor == int(total[owner == o].sum()) == int(plans[o][2].sum())
    # balance: no rank gets more than its quota plus one bin
    if total.sum() > 0:
        per_rank = np.array([total[owner == r].sum() for r in range(nranks)])
        assert per_rank.max() <= total.sum() / nranks + total.max() + 1


# ---------------------------------------------------------------------------------------------------------
def _worker(rank, world, port, n, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raw = ctypes.CDLL(vstree_b200.library_path())
        raw.esa_shard_plan.restype = ctypes.c_int
        rng = np.random.default_rng(99)          # every rank holds the same (replicated) sequence
        tis = rng.integers(0, 4, size=n + 16, dtype=np.uint64)
        tis[: n // 3] %= 2                         # skew: a third of the text uses two letters only
        h0 = 15
        code = np.zeros(n, np.uint64)
        for i in range(h0):
            code = code * 4 + tis[i:i + n]
        key = (code << 1).astype(np.uint32)        # full h0-mers only (no specials in this text)
        sl = -(-n // world)
        lo, hi = rank * sl, min(n, (rank + 1) * sl)
        my_key, my_pos = key[lo:hi], np.arange(lo, hi, dtype=np.uint32)
        bins = (my_key >> 23).astype(np.int64)
        local = np.bincount(bins, minlength=256).astype(np.uint64)
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        hist = np.stack(gathered)
        owner, send, recv, start = plan(raw, hist, rank)
        # play the fused partition + all-to-all: a stable partition by bin, every run stored at start[bin] of its owner
        order = np.argsort(bins, kind="stable")
        outgoing = [dict() for _ in range(world)]
        cursor = start.astype(np.int64).copy()
        for j in order:
            b = bins[j]
            outgoing[owner[b]][int(cursor[b])] = (int(my_key[j]), int(my_pos[j]))
            cursor[b] += 1
        inbox = [None] * world
        dist.all_gather_object(inbox, outgoing)
        mine = {}
        for sender in range(world):
            part = inbox[sender][rank]
            assert len(part) == recv[sender], "receive count"
            assert not (mine.keys() & part.keys()), "two runs overlap in my buffer"
            mine.update(part)
        size = int(recv.sum())
        assert sorted(mine) == list(range(size)), "my buffer has holes"
        buf = np.array([mine[i] for i in range(size)], np.int64).reshape(size, 2)
        # my range: sort it (stable, as the LSD radix sort does) and hand it back for the global comparison
        s = np.argsort(buf[:, 0], kind="stable") if size else np.zeros(0, np.int64)
        result = [None] * world
        dist.all_gather_object(result, buf[s])
        if rank == 0:
            whole = np.concatenate([r.reshape(-1, 2) for r in result])
            want = np.argsort(key, kind="stable")
            assert np.array_equal(whole[:, 1], want), "ranges do not concatenate to the globally sorted order"
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_between_real_processes_over_gloo(world):
    """world_size > 1 on the CPU: every process derives the plan from the gathered counts, sends its runs to the
    owners, and the owners' sorted ranges concatenate to the global (stable) order of the initial sort."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29800 + world + (os.getpid() % 100)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 40000, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    assert out.get(timeout=5) == "ok"


def test_plan_counts_the_special_tail_in_the_last_ranks_quota(lib):
    """Rows of 258 words: [257] = special-initial suffixes, which all land in the last range."""
    nranks = 8
    rng = np.random.default_rng(3)
    hist = np.zeros((nranks, 258), np.uint64)
    hist[:, :256] = rng.integers(900, 1100, size=(nranks, 256))
    hist[:, 257] = 256 * 1000 // 20  # 5 % of the suffixes start with a special
    h = np.ascontiguousarray(hist)
    owner = np.zeros(256, np.uint8)
    send, recv = np.zeros(nranks, np.uint64), np.zeros(nranks, np.uint64)
    start = np.zeros(256, np.uint32)
    loads = []
    for me in range(nranks):
        rc = lib.esa_shard_plan(h.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(258), nranks, me,
                                owner.ctypes.data_as(ctypes.c_void_p), send.ctypes.data_as(ctypes.c_void_p),
                                recv.ctypes.data_as(ctypes.c_void_p), start.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        loads.append(int(recv.sum()) + (int(hist[:, 257].sum()) if me == nranks - 1 else 0))
    fair = sum(loads) / nranks
    assert max(loads) < 1.05 * fair and min(loads) > 0.95 * fair, loads
