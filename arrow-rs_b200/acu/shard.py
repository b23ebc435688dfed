"""Row-range sharding of a table across ranks.

filter / take-within-shard / arithmetic / cmp / cast are row-local, so a RecordBatch is split
into contiguous row ranges aligned to 64 rows (validity / predicate bitmaps then split on u64
words) with no data-path collective. The only exchange is the final reduction of scalar
aggregates: sum -> wrapping integer / IEEE float sum, min/max -> reduced on IEEE-754 totalOrder
integer keys (arrow-array/src/arithmetic.rs:400-437), because a plain float min/max is not the
reference's totalOrder. The same key transform is implemented in csrc/comm.cu for NCCL.
"""
import numpy as np

from . import _abi as abi


def shard_ranges(n_rows, world, align=64):
    """Contiguous [lo, hi) per rank; every boundary except the last is a multiple of `align`."""
    per = -(-n_rows // world)
    per = -(-per // align) * align
    out = []
    for r in range(world):
        lo = min(r * per, n_rows)
        out.append((lo, min(lo + per, n_rows)))
    return out


def total_order_key(value, dtype):
    """Native scalar -> python int whose ordering is the reference's totalOrder."""
    if dtype == abi.F64:
        b = int(np.array([value], dtype=np.float64).view(np.int64)[0])
        return b ^ (((b >> 63) & 0xFFFFFFFFFFFFFFFF) >> 1)
    if dtype == abi.F32:
        b = int(np.array([value], dtype=np.float32).view(np.int32)[0])
        return b ^ (((b >> 31) & 0xFFFFFFFF) >> 1)
    return int(value)


def from_total_order_key(key, dtype):
    if dtype == abi.F64:
        b = key ^ (((key >> 63) & 0xFFFFFFFFFFFFFFFF) >> 1)
        return float(np.array([b], dtype=np.int64).view(np.float64)[0])
    if dtype == abi.F32:
        b = key ^ (((key >> 31) & 0xFFFFFFFF) >> 1)
        return float(np.array([b], dtype=np.int32).view(np.float32)[0])
    return key


def combine_aggregates(op, dtype, partials):
    """Fold per-shard `(value_or_None, valid_count)` pairs exactly as the all-reduce does.
    Returns (value_or_None, total_valid_count)."""
    vals = [(v, c) for v, c in partials if c > 0 and v is not None]
    total = sum(c for _, c in partials)
    if not vals:
        return None, total
    if op == abi.SUM:
        if dtype in (abi.F32, abi.F64):
            npdt = np.float32 if dtype == abi.F32 else np.float64
            acc = npdt(0)
            for v, _ in vals:
                acc = npdt(acc + npdt(v))
            return float(acc), total
        bits = abi.DTYPE_SIZE[dtype] * 8
        s = sum(int(v) for v, _ in vals) & ((1 << bits) - 1)  # add_wrapping
        if dtype in (abi.I8, abi.I16, abi.I32, abi.I64) and s >= 1 << (bits - 1):
            s -= 1 << bits
        return s, total
    keys = [total_order_key(v, dtype) for v, _ in vals]
    k = min(keys) if op == abi.MIN else max(keys)
    return from_total_order_key(k, dtype), total
