#!/usr/bin/env python
"""tools/opbench.py — per-op device-time table for every config of BASELINE.json.

For each op: algorithmic bytes (each input read once, each output written once, bitmaps
ceil(rows/8)) / kernel-only CUDA-event time (acu_kernel_stats) vs the measured HBM peak.
Writes one JSON object per line and a markdown table to stdout; used to fill profiles/.
"""
import argparse
import ctypes as C
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "arrow-rs_b200"))
import numpy as np  # noqa: E402

import acu  # noqa: E402
from acu import _abi as abi  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


class Bench:
    def __init__(self, ctx, reps):
        self.ctx, self.lib, self.h, self.reps, self.rows = ctx, ctx.lib, ctx.h, reps, []
        self.peak = peak()
        self.only = None
        self.quiet = False

    def arr(self, values, validity, n, nc, voff=0, scalar=0):
        a = abi.Array()
        a.values, a.values_offset, a.validity, a.validity_offset, a.len, a.null_count, a.is_scalar = values, voff, validity, 0, n, nc, scalar
        return a

    def out(self, vbytes, rows):
        o = abi.ArrayOut()
        o.values, o.validity = self.ctx.malloc(vbytes + 64), self.ctx.malloc(abi.bitmap_bytes(rows) + 64)
        return o

    def gen(self, kind, seed, n, width, param=0):
        d = self.ctx.malloc(n * width + 64)
        self.ctx.check(self.lib.acu_generate_values(self.h, kind, seed, 0, param, d, n))
        return d

    def bits(self, seed, p, n):
        d = self.ctx.malloc(abi.bitmap_bytes(n) + 64)
        self.ctx.check(self.lib.acu_generate_bits(self.h, seed, 0, p, d, n))
        c = C.c_int64(0)
        self.ctx.check(self.lib.acu_bitmap_count(self.h, d, 0, None, 0, n, C.byref(c)))
        return d, c.value

    def timed(self, name, classes, alg_bytes, rows, fn, note=""):
        if self.only and not any(t in name for t in self.only.split("|")):
            return
        for _ in range(2):
            fn()
        self.ctx.check(self.lib.acu_kernel_stats_reset(self.h))
        ms = C.c_float(0)
        self.ctx.check(self.lib.acu_timer_start_slot(self.h, 3))
        for _ in range(self.reps):
            fn()
        self.ctx.check(self.lib.acu_timer_stop_slot(self.h, 3, C.byref(ms)))
        k_ms = 0.0
        for cls in classes:
            tot, cnt = C.c_double(0), C.c_int64(0)
            self.ctx.check(self.lib.acu_kernel_stats(self.h, cls, C.byref(tot), C.byref(cnt)))
            k_ms += tot.value
        k_ms /= self.reps
        gbs = alg_bytes / (k_ms * 1e-3) / 1e9
        row = {"op": name, "rows": rows, "kernel_ms": round(k_ms, 4), "call_ms": round(ms.value / self.reps, 4),
               "algorithmic_bytes": alg_bytes, "achieved_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / self.peak, 4),
               "mrows_s": round(rows / (k_ms * 1e-3) / 1e6, 1), "note": note}
        self.rows.append(row)
        if not self.quiet:
            print(json.dumps(row), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--small-rows", type=int, default=100_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None, help="substring of the op names to run (alternatives separated by |)")
    args = ap.parse_args()
    n, ns = args.rows, args.small_rows
    with acu.Context(0) as ctx:
        b = Bench(ctx, args.reps)
        b.only = args.only
        lib, h = ctx.lib, ctx.h
        bb = abi.bitmap_bytes(n)
        # ---------------- config 3: binary add/mul + cmp Float64 1e9, 5 % nulls ----------------
        da, dbv = b.gen(2, 42, n, 8), b.gen(2, 43, n, 8)
        va, nva = b.bits(44, 0.95, n)
        vb, nvb = b.bits(45, 0.95, n)
        A, Bv = b.arr(da, va, n, n - nva), b.arr(dbv, vb, n, n - nvb)
        o = b.out(n * 8, n)
        for name, op in [("add f64", abi.ADD), ("mul f64", abi.MUL), ("div f64", abi.DIV)]:
            b.timed(name, [abi.K_ARITH], 24 * n + 3 * n / 8, n, lambda op=op: ctx.check(lib.acu_arith(h, abi.F64, op, C.byref(A), C.byref(Bv), C.byref(o))))
        for name, op in [("lt f64", abi.LT), ("eq f64", abi.EQ)]:
            b.timed(name, [abi.K_CMP], 16 * n + 4 * n / 8, n, lambda op=op: ctx.check(lib.acu_cmp(h, abi.F64, op, C.byref(A), C.byref(Bv), C.byref(o))))
        sc = b.arr(dbv, None, 1, 0, scalar=1)
        b.timed("add f64 array+scalar", [abi.K_ARITH], 16 * n + 2 * n / 8, n, lambda: ctx.check(lib.acu_arith(h, abi.F64, abi.ADD, C.byref(A), C.byref(sc), C.byref(o))))
        # Int64 checked add on the same buffers reinterpreted (values in [-2^61, 2^61) -> no overflow)
        di, dj = b.gen(1, 42, n, 8), None
        ctx.free(dbv)
        dj = b.gen(1, 43, n, 8)
        I, J = b.arr(di, va, n, n - nva), b.arr(dj, vb, n, n - nvb)
        b.timed("add i64 checked", [abi.K_ARITH], 24 * n + 3 * n / 8, n, lambda: ctx.check(lib.acu_arith(h, abi.I64, abi.ADD, C.byref(I), C.byref(J), C.byref(o))))
        b.timed("add_wrapping i64", [abi.K_ARITH], 24 * n + 3 * n / 8, n, lambda: ctx.check(lib.acu_arith(h, abi.I64, abi.ADD_WRAPPING, C.byref(I), C.byref(J), C.byref(o))))
        # predicate construction (arrow-arith/src/boolean.rs): bitmaps only, 1e9 rows
        bl, nbl = b.bits(50, 0.5, n)
        br, nbr = b.bits(51, 0.5, n)
        BL, BR = b.arr(bl, va, n, n - nva), b.arr(br, vb, n, n - nvb)
        b.timed("and_kleene bool (nulls both sides)", [abi.K_CMP], 6 * n / 8, n,
                lambda: ctx.check(lib.acu_boolean(h, abi.BOOL_AND_KLEENE, C.byref(BL), C.byref(BR), C.byref(o))))
        b.timed("is_not_null", [abi.K_CMP], 2 * n / 8, n, lambda: ctx.check(lib.acu_boolean(h, abi.BOOL_IS_NOT_NULL, C.byref(A), None, C.byref(o))))
        ctx.free(bl)
        ctx.free(br)
        # aggregates over a full column
        bits_, cnt_ = C.c_uint64(0), C.c_int64(0)
        for name, dt, arr_, op in [("sum i64", abi.I64, I, abi.SUM), ("min f64", abi.F64, A, abi.MIN), ("sum f64", abi.F64, A, abi.SUM)]:
            b.timed(name, [abi.K_REDUCE], 8 * n + n / 8, n, lambda dt=dt, arr_=arr_, op=op: ctx.check(lib.acu_aggregate(h, dt, op, C.byref(arr_), C.byref(bits_), C.byref(cnt_))))
        ctx.free(dj)
        # ---------------- config 2: filter + take Int64 1e9 ----------------
        for sel in (0.01, 0.1, 0.5, 0.9):
            dp, m = b.bits(46, sel, n)
            pred = b.arr(dp, None, n, 0)
            of = b.out(m * 8, m)
            plan = C.c_void_p()

            def run_filter():
                p = C.c_void_p()
                ctx.check(lib.acu_filter_plan_create(h, C.byref(pred), C.byref(p)))
                ctx.check(lib.acu_filter_primitive(h, p, 8, C.byref(I), C.byref(of)))
                lib.acu_filter_plan_destroy(h, p)

            b.timed(f"filter i64 s={sel}", [abi.K_FILTER, abi.K_FILTER_PLAN], 8 * n + 2 * n / 8 + 8 * m + m / 8, n, run_filter, note=f"selected {m}; plan + values + validity kernels")
            if sel in (0.1, 0.5):
                didx = ctx.malloc(m * 4 + 64)
                ctx.check(lib.acu_filter_plan_create(h, C.byref(pred), C.byref(plan)))
                ctx.check(lib.acu_filter_plan_indices(h, plan, abi.U32, didx))
                lib.acu_filter_plan_destroy(h, plan)
                ix = b.arr(didx, None, m, 0)
                b.timed(f"take i64 monotone M/N={sel}", [abi.K_TAKE], 20.25 * m, m, lambda: ctx.check(lib.acu_take_primitive(h, 8, C.byref(I), C.byref(ix), abi.U32, 0, C.byref(of))),
                        note="index distribution A (selected rows)")
                ctx.free(didx)
            ctx.free(dp)
            ctx._free_out(of)
        m = ns
        drand = b.gen(3, 47, m, 4, param=n)
        ix = b.arr(drand, None, m, 0)
        ot = b.out(m * 8, m)
        b.timed("take i64 uniform random M=1e8", [abi.K_TAKE], 20.25 * m, m, lambda: ctx.check(lib.acu_take_primitive(h, 8, C.byref(I), C.byref(ix), abi.U32, 0, C.byref(ot))),
                note="index distribution B")
        ctx.free(drand)
        ctx._free_out(ot)
        # ---------------- config 4: cast Int64 -> Float64 (1e8) and Dictionary<Int32,Utf8> -> Utf8 (1e8) ----------------
        Is = b.arr(di, va, ns, -1)
        oc = b.out(ns * 8, ns)
        b.timed("cast i64->f64", [abi.K_CAST], 16 * ns + 2 * ns / 8, ns, lambda: ctx.check(lib.acu_cast_numeric(h, abi.I64, abi.F64, 1, C.byref(Is), C.byref(oc))))
        D = 4096
        rng = np.random.default_rng(1)
        lens = rng.integers(4, 13, D)
        offs = np.zeros(D + 1, dtype=np.int32)
        offs[1:] = np.cumsum(lens)
        data = rng.integers(97, 123, int(offs[-1]) + 16).astype(np.uint8)
        d_off, d_data = ctx.malloc(offs.nbytes + 64), ctx.malloc(data.nbytes + 64)
        ctx.h2d(d_off, offs)
        ctx.h2d(d_data, data)
        dkeys = b.gen(4, 48, ns, 4, param=D)
        kv, nkv = b.bits(49, 0.95, ns)
        keys = b.arr(dkeys, kv, ns, ns - nkv)
        dict_nulls = b.arr(None, None, D, 0)
        d_out_off = ctx.malloc((ns + 1) * 4 + 64)
        d_out_data = ctx.malloc(ns * 13 + 64)
        on = abi.ArrayOut()
        on.validity = ctx.malloc(abi.bitmap_bytes(ns) + 64)
        total = C.c_int64(0)
        b.timed("cast dict<i32,utf8>->utf8", [abi.K_TAKE, abi.K_BYTES], 4 * ns + ns / 8 + 4 * (ns + 1) + 0.95 * ns * 8 + ns / 8, ns,
                lambda: ctx.check(lib.acu_take_bytes(h, 4, d_off, d_data, C.byref(dict_nulls), C.byref(keys), abi.I32, 0, d_out_off, d_out_data, ns * 13, C.byref(total), C.byref(on))),
                note="kernel_ms = validity + table + lengths + copy kernels (the three tiny scan kernels only show in call_ms)")
    print("\n| op | rows | kernel ms | GB/s (algorithmic) | % of measured HBM peak | Mrows/s |")
    print("|---|---|---|---|---|---|")
    for r in b.rows:
        print(f"| {r['op']} | {r['rows']:.3g} | {r['kernel_ms']} | {r['achieved_gbs']} | {100 * r['frac_of_measured_peak']:.1f} | {r['mrows_s']} |")


if __name__ == "__main__":
    main()
