"""Helpers for the Utf8View / BinaryView coalescing tests (host side)."""
import numpy as np

from acu import ViewColumn


def view_batch(n, items, block=8192):
    """stringview_batch_repeated (arrow-select/src/coalesce.rs tests): n rows cycling through `items`, built like
    StringViewBuilder::with_fixed_block_size(block): long values fill blocks of `block` bytes (capacity = block)."""
    vals = [items[i % len(items)] for i in range(n)]
    c = ViewColumn.from_values(vals, block)
    c.buffer_capacities = [max(block, int(b.nbytes)) for b in c.buffers]
    return c


def view_slice(col, off, n):
    c = ViewColumn(col.views[off:off + n], col.buffers, col.nulls.slice(off, n))
    c.buffer_capacities = getattr(col, "buffer_capacities", None)
    return c


def view_values(col):
    """Logical values (bytes or None) of a ViewColumn."""
    out = []
    valid = col.nulls.valid_mask() if col.nulls.validity is not None else np.ones(col.length, dtype=bool)
    for i in range(col.length):
        if not valid[i]:
            out.append(None)
            continue
        v = col.views[i]
        ln = int(np.frombuffer(v[:4].tobytes(), dtype=np.uint32)[0])
        if ln <= 12:
            out.append(v[4:4 + ln].tobytes())
        else:
            bi, off = (int(x) for x in np.frombuffer(v[8:16].tobytes(), dtype=np.uint32))
            out.append(col.buffers[bi][off:off + ln].tobytes())
    return out


def as_bytes(items):
    return [None if x is None else (x.encode() if isinstance(x, str) else bytes(x)) for x in items]
