"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: row-range sharding + the final
reduction of scalar aggregates. Each rank runs the step on its shard with the CPU oracle, the
partials are exchanged with torch.distributed (gloo) exactly as bench.py does over NCCL, and the
result must equal the single-process result over the whole table (bit-exact for integers and
float min/max; float sum within the documented tolerance)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_table(n):
    rng = np.random.default_rng(123)
    vals = rng.integers(-2**62, 2**62, n, dtype=np.int64)
    fvals = rng.random(n) * 2e6 - 1e6
    fvals[rng.integers(0, n, 16)] = np.nan
    fvals[rng.integers(0, n, 16)] = -np.inf
    valid = rng.random(n) >= 0.05
    pred = rng.random(n) < 0.1
    return vals, fvals, valid, pred


def _shard_step(orc, acu, abi, vals, fvals, valid, pred, lo, hi):
    from acu import HostArray
    col = HostArray.from_numpy(abi.I64, vals[lo:hi], valid[lo:hi])
    fcol = HostArray.from_numpy(abi.F64, fvals[lo:hi], valid[lo:hi])
    p = HostArray.bool_from_numpy(pred[lo:hi])
    f = orc.filter(col, p)
    ff = orc.filter(fcol, p)
    idx = HostArray.from_numpy(abi.U32, np.arange(0, f.length, 2, dtype=np.uint32))  # monotone half-sample of the filtered rows
    t, tf = orc.take(f, idx), orc.take(ff, idx)
    nvalid = int(t.valid_mask().sum())
    return {"rows": f.length, "sum": (orc.sum(t), nvalid), "min": (orc.min(tf), nvalid), "max": (orc.max(tf), nvalid), "fsum": (orc.sum(tf), nvalid)}


def _worker(rank, world, port, n, out_q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
    sys.path.insert(0, HERE)
    import acu
    from acu import _abi as abi
    from acu import shard
    from oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    vals, fvals, valid, pred = _make_table(n)
    lo, hi = shard.shard_ranges(n, world)[rank]
    part = _shard_step(orc, acu, abi, vals, fvals, valid, pred, lo, hi)
    # --- the exchange: row counts + integer sum as int64 all-reduce(sum), min/max on totalOrder keys ---
    t = torch.tensor([part["rows"], part["sum"][1], part["sum"][0] or 0], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # int64 sum wraps like add_wrapping
    kmin = torch.tensor([shard.total_order_key(part["min"][0], abi.F64) if part["min"][0] is not None else 2**63 - 1], dtype=torch.int64)
    kmax = torch.tensor([shard.total_order_key(part["max"][0], abi.F64) if part["max"][0] is not None else -2**63], dtype=torch.int64)
    dist.all_reduce(kmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    fs = torch.tensor([part["fsum"][0] if part["fsum"][0] is not None and not np.isnan(part["fsum"][0]) else 0.0], dtype=torch.float64)
    dist.all_reduce(fs, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    if rank == 0:
        out_q.put({"rows": int(t[0]), "valid": int(t[1]), "sum": int(t[2]), "min": shard.from_total_order_key(int(kmin[0]), abi.F64),
                   "max": shard.from_total_order_key(int(kmax[0]), abi.F64), "fsum": float(fs[0]), "parts": gathered})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [100_003, 4096])
def test_two_rank_sharded_step_matches_single_process(n):
    import torch.multiprocessing as mp  # imported lazily: collecting the GPU tests must not pay for `import torch`
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
    import acu
    from acu import _abi as abi
    from acu import shard
    from oracle import Oracle
    orc = Oracle()
    vals, fvals, valid, pred = _make_table(n)
    ranges = shard.shard_ranges(n, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(r[0] % 64 == 0 for r in ranges)
    # the same shard-local pipeline, single process: fold with combine_aggregates
    parts = [_shard_step(orc, acu, abi, vals, fvals, valid, pred, lo, hi) for lo, hi in ranges]
    assert got["rows"] == sum(p["rows"] for p in parts)
    exp_sum, exp_valid = shard.combine_aggregates(abi.SUM, abi.I64, [p["sum"] for p in parts])
    assert got["valid"] == exp_valid
    assert np.int64(np.uint64(got["sum"] & 0xFFFFFFFFFFFFFFFF)) == np.int64(exp_sum)
    exp_min, _ = shard.combine_aggregates(abi.MIN, abi.F64, [p["min"] for p in parts])
    exp_max, _ = shard.combine_aggregates(abi.MAX, abi.F64, [p["max"] for p in parts])
    for g, e in ((got["min"], exp_min), (got["max"], exp_max)):
        assert (np.isnan(g) and np.isnan(e) and np.signbit(g) == np.signbit(e)) or g == e
    # and the min/max over shards equals the oracle's min/max over the concatenated taken column (totalOrder is associative)
    from acu import HostArray
    allvals, allmask = [], []
    for (lo, hi) in ranges:
        fcol = HostArray.from_numpy(abi.F64, fvals[lo:hi], valid[lo:hi])
        ff = orc.filter(fcol, HostArray.bool_from_numpy(pred[lo:hi]))
        tf = orc.take(ff, HostArray.from_numpy(abi.U32, np.arange(0, ff.length, 2, dtype=np.uint32)))
        allvals.append(tf.value_array())
        allmask.append(tf.valid_mask())
    whole = HostArray.from_numpy(abi.F64, np.concatenate(allvals), np.concatenate(allmask))
    wmin, wmax = orc.min(whole), orc.max(whole)
    for g, e in ((got["min"], wmin), (got["max"], wmax)):
        assert (np.isnan(g) and np.isnan(e) and np.signbit(g) == np.signbit(e)) or g == e
    assert [p["rows"] for p in got["parts"]] == [p["rows"] for p in parts]


def test_shard_ranges_properties():
    sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
    from acu import shard
    for n in [0, 1, 63, 64, 65, 1000, 10**9]:
        for w in [1, 2, 3, 8]:
            r = shard.shard_ranges(n, w)
            assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo % 64 == 0 for lo, _ in r if lo < n)


def test_total_order_key_round_trip_and_order():
    sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
    from acu import _abi as abi
    from acu import shard
    neg_nan = float(np.array([0xFFF8000000000000], dtype=np.uint64).view(np.float64)[0])
    xs = [neg_nan, -np.inf, -1.5, -0.0, 0.0, 5e-324, 2.0, np.inf, float("nan")]  # ascending totalOrder
    keys = [shard.total_order_key(x, abi.F64) for x in xs]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    for x, k in zip(xs, keys):
        y = shard.from_total_order_key(k, abi.F64)
        assert np.array([x]).view(np.uint64)[0] == np.array([y]).view(np.uint64)[0]


def _rendezvous_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    from acu.rendezvous import Group
    g = object.__new__(Group)  # the transport alone: no ctx / NCCL on a CPU box
    g.rank, g.world = rank, world
    payload = bytes(range(128)) if rank == 0 else b""
    q.put((rank, g._socket_broadcast(payload)))


def test_socket_rendezvous_world_3():
    """The torch-free NCCL-unique-id exchange of the benches (acu/rendezvous.py): rank 0 -> every rank."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 500)
    procs = [ctx.Process(target=_rendezvous_worker, args=(r, 3, port, q)) for r in (1, 2, 0)]  # rank 0 last: the others must retry
    for p in procs:
        p.start()
    got = dict(q.get(timeout=60) for _ in range(3))
    for p in procs:
        p.join(timeout=30)
    assert all(got[r] == bytes(range(128)) for r in range(3))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_outputs_concatenate_to_the_global_result(world):
    """with contiguous 64-row-aligned row ranges, the filter output of shard g is the g-th contiguous
    piece of the global filter output (values, validity, Utf8 strings), global offsets = exclusive scan of the counts."""
    sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
    sys.path.insert(0, HERE)
    import acu
    from acu import _abi as abi
    from acu import HostArray, shard
    from golden_util import strings_of
    from oracle import Oracle
    orc = Oracle()
    rng = np.random.default_rng(world)
    n = 10_000 + world
    vals = rng.integers(-1000, 1000, n).astype(np.int64)
    valid = rng.random(n) >= 0.1
    pred = rng.random(n) < 0.3
    lens = rng.integers(0, 9, n)
    offs = np.zeros(n + 1, dtype=np.int32)
    offs[1:] = np.cumsum(lens)
    data = rng.integers(97, 123, int(offs[-1]) + 16).astype(np.uint8)

    def strings(lo, hi):
        o = (offs[lo:hi + 1] - offs[lo]).astype(np.int32)
        d = data[offs[lo]: offs[hi] + 16].copy()
        nl = HostArray(abi.U8, np.zeros(0, np.uint8), hi - lo, acu.pack_bits(valid[lo:hi]), 0, 0, int((~valid[lo:hi]).sum()))
        return o, d, nl

    whole = orc.filter(HostArray.from_numpy(abi.I64, vals, valid), HostArray.bool_from_numpy(pred))
    whole_s = strings_of(*orc.filter_bytes(*strings(0, n), HostArray.bool_from_numpy(pred)))
    pieces, pieces_s, counts = [], [], []
    for lo, hi in shard.shard_ranges(n, world):
        assert lo % 64 == 0
        p = HostArray.bool_from_numpy(pred[lo:hi])
        f = orc.filter(HostArray.from_numpy(abi.I64, vals[lo:hi], valid[lo:hi]), p)
        pieces += f.to_list()
        pieces_s += strings_of(*orc.filter_bytes(*strings(lo, hi), p))
        counts.append(f.length)
    assert pieces == whole.to_list() and pieces_s == whole_s
    assert sum(counts) == whole.length and list(np.cumsum([0] + counts[:-1])) == [sum(counts[:g]) for g in range(world)]


def test_socket_rendezvous_ignores_strays_and_serves_each_rank_once(monkeypatch):
    """acu/rendezvous.py: the NCCL unique id exchange over TCP (no torch): a stray connection must not consume a slot, a
    peer that closes early must raise instead of spinning, every rank gets the payload."""
    import threading
    import time
    sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
    from acu.rendezvous import Group
    port = _free_port()
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MASTER_PORT", str(port - 17))
    world, payload, got = 3, bytes(range(128)), {}

    def run(rank):
        g = Group.__new__(Group)
        g.rank, g.world = rank, world
        got[rank] = g._socket_broadcast(payload if rank == 0 else b"", timeout=20.0)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    ts[0].start()
    time.sleep(0.3)
    with socket.create_connection(("127.0.0.1", port), timeout=5) as stray:  # connects, sends garbage, leaves
        stray.sendall(b"hello")
    for t in ts[1:]:
        t.start()
    for t in ts:
        t.join(timeout=30)
    assert got == {0: payload, 1: payload, 2: payload}
    with pytest.raises(ConnectionError):
        a, b = socket.socketpair()
        b.close()
        Group._recv_exact(a, 4)
