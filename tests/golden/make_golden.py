#!/usr/bin/env python
"""Transcribes literal known-answer tests of the reference (apache/arrow-rs @ cd7c6b83) into
tests/golden/vectors.json.

Nothing here runs the reference (there is no Rust toolchain in this environment): every case
is a literal input -> literal expected output copied from the cited test, with the iterator
expressions of the Rust test (`(1..=65)`, `i % 3 == 0`, ...) re-evaluated in Python. Run
`python tests/golden/make_golden.py` to regenerate the committed JSON.

Value encoding: JSON numbers, null = Arrow null, and the float tokens "nan" (any NaN),
"+nan", "-nan", "inf", "-inf", "-0.0"; "*" in an expectation means "not asserted by the
reference test".
"""
import json
import os

I32_MIN, I32_MAX = -2**31, 2**31 - 1
I64_MIN, I64_MAX = -2**63, 2**63 - 1
I16_MIN, I16_MAX = -2**15, 2**15 - 1
I8_MIN, I8_MAX = -128, 127
F32_MAX = 3.4028234663852886e38
F64_MAX = 1.7976931348623157e308

cases = []


def arr(dtype, data, slice=None, scalar=False, force_validity=False):
    d = {"dtype": dtype, "data": data}
    if slice is not None:
        d["slice"] = list(slice)
    if scalar:
        d["scalar"] = True
    if force_validity:
        d["force_validity"] = True
    return d


def case(id, ref, op, **kw):
    c = {"id": id, "ref": ref, "op": op}
    c.update(kw)
    cases.append(c)


# =========================================================================================
# filter — arrow-select/src/filter.rs
# =========================================================================================
F = "arrow-select/src/filter.rs"
case("filter_doc_example", F + ":191-200", "filter",
     values=arr("int32", [5, 6, 7, 8, 9]), predicate=arr("bool", [True, False, False, True, False]),
     expect={"data": [5, 8]})
case("filter_array_slice", F + ":1177", "filter",
     values=arr("int32", [5, 6, 7, 8, 9], slice=(1, 4)), predicate=arr("bool", [True, False, False, True]),
     expect={"data": [6, 9]})
data = list(range(1, 66)) + [66, 67]
pred = [(i % 65) == 0 for i in range(1, 66)] + [False, True]
case("filter_array_low_density", F + ":1191", "filter",
     values=arr("int32", data), predicate=arr("bool", pred), expect={"data": [65, 67]})
data = list(range(1, 66))
data[1] = None
data += [66, None, 67, None]
pred = [(i % 65) != 0 for i in range(1, 66)] + [False, True, True, True]
case("filter_array_high_density", F + ":1208", "filter",
     values=arr("int32", data), predicate=arr("bool", pred),
     expect={"len": 67, "null_count": 3, "at": {"0": 1, "1": None, "63": 64, "64": None, "65": 67}})
case("filter_primitive_array_with_null", F + ":1244", "filter",
     values=arr("int32", [5, None]), predicate=arr("bool", [False, True]), expect={"data": [None]})
case("filter_array_slice_with_null", F + ":1414", "filter",
     values=arr("int32", [5, None, 7, 8, 9], slice=(1, 4)), predicate=arr("bool", [True, False, False, True]),
     expect={"data": [None, 9]})
case("filter_null_mask", F + ":1718", "filter",
     values=arr("int64", [1, 2, None]), predicate=arr("bool", [True, True, None]), expect={"data": [1, 2]})
case("filter_fast_path_all_true", F + ":1738", "filter",
     values=arr("int64", [1, 2, None]), predicate=arr("bool", [True, True, True]), expect={"data": [1, 2, None]})
case("filter_fast_path_all_false", F + ":1750", "filter",
     values=arr("int64", [1, 2, None]), predicate=arr("bool", [False, False, False]), expect={"data": []})
case("filter_predicate_longer_than_values", F + ":536-542", "filter",
     values=arr("int32", [1, 2]), predicate=arr("bool", [True, False, True]),
     expect_error="Invalid argument error: Filter predicate of length 3 is larger than target array of length 2")
case("filter_boolean_values", F + ":723-729", "filter",
     values=arr("bool", [True, None, False, True]), predicate=arr("bool", [True, True, False, True]),
     expect={"data": [True, None, True]})
# SlicesIterator / IndexIterator ground truth: the selected positions (filter.rs:1642-1678, :1758)
case("slices_iterator_bits", F + ":1642", "selected",
     predicate=arr("bool", [i == 1 for i in range(64)]), expect={"positions": [1], "count": 1})
case("slices_iterator_bits1", F + ":1655", "selected",
     predicate=arr("bool", [i != 1 for i in range(64)]),
     expect={"positions": [0] + list(range(2, 64)), "count": 63})
case("slices_iterator_chunk_and_bits", F + ":1668", "selected",
     predicate=arr("bool", [i % 62 != 0 for i in range(130)]),
     expect={"positions": list(range(1, 62)) + list(range(63, 124)) + list(range(125, 130)), "count": 127})
bools = [True] * 10 + [False] * 30 + [True] * 20 + [False] * 17 + [True] * 4
case("slices_full", F + ":1758", "selected", predicate=arr("bool", bools),
     expect={"positions": list(range(0, 10)) + list(range(40, 60)) + list(range(77, 81)), "count": 34})
case("slices_sliced_offset_truncated", F + ":1774-1783", "selected",
     predicate=arr("bool", bools, slice=(7, len(bools) - 10)),
     expect={"positions": list(range(0, 3)) + list(range(33, 53)) + list(range(70, 71)), "count": 24})
# strings (filter_bytes)
case("filter_string_array_simple", F + ":1233", "filter_utf8",
     values={"strings": ["hello", " ", "world", "!"]}, predicate=arr("bool", [True, False, True, False]),
     expect={"strings": ["hello", "world"]})
case("filter_string_array_with_null", F + ":1254", "filter_utf8",
     values={"strings": ["hello", None, "world", None]}, predicate=arr("bool", [True, False, False, True]),
     expect={"strings": ["hello", None]})

# =========================================================================================
# take — arrow-select/src/take.rs
# =========================================================================================
T = "arrow-select/src/take.rs"
case("take_primitive_non_null_indices", T + ":1295", "take",
     values=arr("int8", [None, 3, 5, 2, 3, None]), indices=arr("uint32", [0, 5, 3, 1, 4, 2]),
     expect={"data": [None, None, 2, 3, 3, 5]})
case("take_primitive_non_null_values", T + ":1307", "take",
     values=arr("int8", [0, 1, 2, 3, 4]), indices=arr("uint32", [3, None, 1, 3, 2]),
     expect={"data": [3, None, 1, 3, 2]})
case("take_primitive_non_null", T + ":1319", "take",
     values=arr("int8", [0, 3, 5, 2, 3, 1]), indices=arr("uint32", [0, 5, 3, 1, 4, 2]),
     expect={"data": [0, 1, 2, 3, 3, 5]})
case("take_nullable_indices_non_null_values_with_offset", T + ":1331", "take",
     values=arr("int64", [0, 10, 20, 30, 40, 50]), indices=arr("uint32", [0, 1, 2, 3, None, None], slice=(2, 4)),
     expect={"data": [20, 30, None, None]})
case("take_nullable_indices_nullable_values_with_offset", T + ":1351", "take",
     values=arr("int64", [None, None, 20, 30, 40, 50]), indices=arr("uint32", [0, 1, 2, 3, None, None], slice=(2, 4)),
     expect={"data": [20, 30, None, None]})
for dt in ["int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64"]:
    case("take_primitive_" + dt, T + ":1371-1440", "take",
         values=arr(dt, [0, None, 2, 3, None]), indices=arr("uint32", [3, None, 1, 3, 2]),
         expect={"data": [3, None, None, 3, 2]})
case("take_primitive_int64_negative", T + ":1434", "take",
     values=arr("int64", [0, None, 2, -15, None]), indices=arr("uint32", [3, None, 1, 3, 2]),
     expect={"data": [-15, None, None, -15, 2]})
case("take_int64_indices_int16", T + ":1553", "take",
     values=arr("int16", [0, None, 2, 3, None]), indices=arr("int64", [3, None, 1, 3, 2]),
     expect={"data": [3, None, None, 3, 2]})
case("take_int64_indices_int64", T + ":1565", "take",
     values=arr("int64", [0, None, 2, -15, None]), indices=arr("int64", [3, None, 1, 3, 2]),
     expect={"data": [-15, None, None, -15, 2]})
case("take_int64_indices_float32", T + ":1589", "take",
     values=arr("float32", [0.0, None, 2.21, -3.1, None]), indices=arr("int64", [3, None, 1, 3, 2]),
     expect={"data": [-3.1, None, None, -3.1, 2.21], "f32": True})
case("take_uint8_indices_int16", T + ":1598", "take",
     values=arr("int16", [0, None, 2, 3, None]), indices=arr("uint8", [3, None, 1, 3, 2]),
     expect={"data": [3, None, None, 3, 2]})
case("take_bool", T + ":1627", "take",
     values=arr("bool", [False, None, True, False, None]), indices=arr("uint32", [3, None, 1, 3, 2]),
     expect={"data": [False, None, None, False, True]})
# indices whose masked (null) slots would be out of bounds (take.rs:1640-1680)
oob_idx = {"dtype": "uint32", "raw_values": [99, 0, 999, 1, 9999, 2], "raw_validity": [False, True, False, True, False, True]}
case("take_bool_nullable_index", T + ":1640", "take",
     values=arr("bool", [True, None, False]), indices=oob_idx, expect={"data": [None, True, None, None, None, False]})
case("take_bool_nullable_index_nonnull_values", T + ":1663", "take",
     values=arr("bool", [True, True, False]), indices=oob_idx, expect={"data": [None, True, None, True, None, False]})
case("take_bool_with_offset", T + ":1686", "take",
     values=arr("bool", [False, None, True, False, None]), indices=arr("uint32", [3, None, 1, 3, 2, None], slice=(2, 4)),
     expect={"data": [None, False, True, None]})
case("take_out_of_bounds_checked", T + ":2408", "take", check_bounds=True,
     values=arr("int64", [0, None, 2, 3, None]), indices=arr("uint32", [3, None, 1, 3, 6]),
     expect_error="Compute error: Array index out of bounds, cannot get item at index 6 from 5 entries")
case("take_out_of_bounds_panic", T + ":2423", "take",
     values=arr("int64", [0, 1, 2, 3]), indices=arr("uint32", [1000]), expect_panic=True)
case("take_check_bounds_message", T + ":2455-2465", "take", check_bounds=True,
     values=arr("int32", [0, 0, 0, 0, 0]), indices=arr("uint32", [0, None, 15]),
     expect_error="Compute error: Array index out of bounds, cannot get item at index 15 from 5 entries")
case("take_null_indices_oob_masked", T + ":2686", "take",
     values=arr("int32", [1, 23, 4, 5]),
     indices={"dtype": "int32", "raw_values": [1, 2, 400, 400], "raw_validity": [True, True, False, False]},
     expect={"data": [23, 4, None, None]})
case("take_doc_example_strings", T + ":76-88", "take_utf8",
     values={"strings": ["zero", "one", "two"]}, indices=arr("uint32", [2, 1]), expect={"strings": ["two", "one"]})
case("take_bytes_null_indices", T + ":2719", "take_utf8",
     values={"strings": ["foo", None]},
     indices={"dtype": "int32", "raw_values": [0, 1, 400, 400], "raw_validity": [True, True, False, False]},
     expect={"strings": ["foo", None, None, None]})

# =========================================================================================
# numeric — arrow-arith/src/numeric.rs, arity.rs
# =========================================================================================
N = "arrow-arith/src/numeric.rs"
a, b = arr("int32", [4, 3, 5, -6, 100]), arr("int32", [6, 2, 5, -7, 3])
for op, exp in [("add", [10, 5, 10, -13, 103]), ("sub", [-2, 1, 0, 1, 97]), ("div", [0, 1, 1, 0, 33]),
                ("mul", [24, 6, 25, 42, 300]), ("rem", [4, 1, 0, -6, 1])]:
    case("integer_" + op, N + ":1296-1311", op, a=a, b=b, expect={"data": exp})
case("integer_add_nulls", N + ":1313-1316", "add",
     a=arr("int8", [2, None, 45]), b=arr("int8", [5, 3, None]), expect={"data": [7, None, None]})
u8a, u8b = arr("uint8", [56, 5, 3]), arr("uint8", [200, 2, 5])
case("u8_add_overflow", N + ":1318-1321", "add", a=u8a, b=u8b,
     expect_error="Arithmetic overflow: Overflow happened on: 56 + 200")
case("u8_add_wrapping", N + ":1322-1323", "add_wrapping", a=u8a, b=u8b, expect={"data": [0, 7, 8]})
u8a = arr("uint8", [34, 5, 3])
case("u8_sub_overflow", N + ":1325-1328", "sub", a=u8a, b=u8b,
     expect_error="Arithmetic overflow: Overflow happened on: 34 - 200")
case("u8_sub_wrapping", N + ":1329-1330", "sub_wrapping", a=u8a, b=u8b, expect={"data": [90, 3, 254]})
case("u8_mul_overflow", N + ":1332-1335", "mul", a=u8a, b=u8b,
     expect_error="Arithmetic overflow: Overflow happened on: 34 * 200")
case("u8_mul_wrapping", N + ":1336-1337", "mul_wrapping", a=u8a, b=u8b, expect={"data": [144, 10, 15]})
case("i16_div_overflow", N + ":1339-1346", "div", a=arr("int16", [I16_MIN]), b=arr("int16", [-1]),
     expect_error="Arithmetic overflow: Overflow happened on: -32768 / -1")
case("i16_rem_min_minus_one", N + ":1348-1351", "rem", a=arr("int16", [I16_MIN]), b=arr("int16", [-1]), expect={"data": [0]})
case("i16_div_zero", N + ":1353-1356", "div", a=arr("int16", [21]), b=arr("int16", [0]), expect_error="Divide by zero error")
case("i16_rem_zero", N + ":1358-1361", "rem", a=arr("int16", [21]), b=arr("int16", [0]), expect_error="Divide by zero error")
fa = arr("float32", [1.0, F32_MAX, 6.0, -4.0, -1.0, 0.0])
fb = arr("float32", [1.0, F32_MAX, F32_MAX, -3.0, 45.0, 0.0])
case("float_add", N + ":1366-1372", "add", a=fa, b=fb, expect={"data": [2.0, "inf", F32_MAX, -7.0, 44.0, 0.0], "f32": True})
case("float_sub", N + ":1374-1378", "sub", a=fa, b=fb, expect={"data": [0.0, 0.0, -F32_MAX, -1.0, -46.0, 0.0], "f32": True})
case("float_mul", N + ":1380-1384", "mul", a=fa, b=fb, expect={"data": [1.0, "inf", "inf", 12.0, -45.0, 0.0], "f32": True})
case("float_div", N + ":1386-1392", "div", a=fa, b=fb,
     expect={"data": [1.0, 1.0, "*", 1.3333333730697632, "*", "nan"], "f32": True})
case("float_rem", N + ":1394-1397", "rem", a=fa, b=fb, expect={"data": [0.0, 0.0, 6.0, -1.0, -1.0, "nan"], "f32": True})
case("neg_i32", N + ":1152-1154", "neg", a=arr("int32", [1, -5, 2, 693, 3929]), expect={"data": [-1, 5, -2, -693, -3929]})
case("neg_i64", N + ":1156-1158", "neg", a=arr("int64", [1, -5, 2, 693, 3929]), expect={"data": [-1, 5, -2, -693, -3929]})
case("neg_f32", N + ":1164-1166", "neg", a=arr("float32", [F32_MAX, -F32_MAX, "inf", 1.3, 0.5]),
     expect={"data": [-F32_MAX, F32_MAX, "-inf", -1.3, -0.5], "f32": True})
case("neg_i32_overflow", N + ":1168-1171", "neg", a=arr("int32", [I32_MIN]),
     expect_error="Arithmetic overflow: Overflow happened on: - -2147483648")
case("neg_i64_overflow", N + ":1172-1175", "neg", a=arr("int64", [I64_MIN]),
     expect_error="Arithmetic overflow: Overflow happened on: - -9223372036854775808")
case("neg_wrapping_i32_min", N + ":1181-1182", "neg_wrapping", a=arr("int32", [I32_MIN]), expect={"data": [I32_MIN]})
case("neg_wrapping_i64_min", N + ":1184-1185", "neg_wrapping", a=arr("int64", [I64_MIN]), expect={"data": [I64_MIN]})
case("neg_unsigned_unsupported", N + ":174-176", "neg", a=arr("uint32", [1]),
     expect_error="Invalid argument error: Invalid arithmetic operation: !UInt32")
A = "arrow-arith/src/arity.rs"
case("binary_length_mismatch", A + ":115-119", "add_wrapping", a=arr("int32", [1, 2]), b=arr("int32", [1]),
     expect_error="Compute error: Cannot perform binary operation on arrays of different length")
case("try_binary_length_mismatch", A + ":263-267", "add", a=arr("int32", [1, 2]), b=arr("int32", [1]),
     expect_error="Compute error: Cannot perform a binary operation on arrays of different length")
case("binary_doc_example", A + ":93-103", "add_wrapping",
     a=arr("int32", [5, None, 7, 8]), b=arr("int32", [1, 2, 3, 8]), expect={"data": [6, None, 10, 16]})
# scalar broadcast (numeric.rs:278-317)
case("add_scalar_rhs", N + ":278-292", "add", a=arr("int64", [1, None, 3]), b=arr("int64", [10], scalar=True),
     expect={"data": [11, None, 13]})
case("sub_scalar_lhs", N + ":278-292", "sub", a=arr("int64", [10], scalar=True), b=arr("int64", [1, None, 3]),
     expect={"data": [9, None, 7]})
case("add_null_scalar", N + ":282-289", "add", a=arr("int64", [1, 2, 3]), b=arr("int64", [None], scalar=True),
     expect={"data": [None, None, None]})
case("checked_overflow_under_null_is_ignored", A + ":285-294", "add",
     a=arr("int64", [I64_MAX, 1]), b=arr("int64", [None, 1]), expect={"data": [None, 2]})
case("checked_overflow_reports_lowest_index", A + ":383-400", "mul",
     a=arr("int64", [1, I64_MAX, I64_MIN]), b=arr("int64", [1, 2, 2]),
     expect_error="Arithmetic overflow: Overflow happened on: 9223372036854775807 * 2")

# =========================================================================================
# cmp — arrow-ord/src/cmp.rs, comparison.rs
# =========================================================================================
K = "arrow-ord/src/comparison.rs"
for dt in ["float32", "float64"]:
    a1 = arr(dt, ["nan", 7.0, 8.0, 8.0, 10.0])
    a2 = arr(dt, ["nan", "nan", 8.0, 8.0, 10.0])
    case("eq_nan_" + dt, K + ":2475-2503", "eq", a=a1, b=a2, expect={"data": [True, False, True, True, True]})
    case("neq_nan_" + dt, K + ":2475-2503", "neq", a=a1, b=a2, expect={"data": [False, True, False, False, False]})
    a1 = arr(dt, ["nan", 7.0, 8.0, 8.0, 11.0, "nan"])
    a2 = arr(dt, ["nan", "nan", 8.0, 9.0, 10.0, 1.0])
    case("lt_nan_" + dt, K + ":2507-2537", "lt", a=a1, b=a2, expect={"data": [False, True, False, True, False, False]})
    case("lt_eq_nan_" + dt, K + ":2507-2537", "lt_eq", a=a1, b=a2, expect={"data": [True, True, True, True, False, False]})
    case("gt_nan_" + dt, K + ":2541-2571", "gt", a=a1, b=a2, expect={"data": [False, False, False, False, True, True]})
    case("gt_eq_nan_" + dt, K + ":2541-2571", "gt_eq", a=a1, b=a2, expect={"data": [True, False, True, False, True, True]})
    a1 = arr(dt, ["nan", 7.0, 8.0, 8.0, 10.0])
    s = arr(dt, ["nan"], scalar=True)
    case("eq_scalar_nan_" + dt, K + ":2575-2600", "eq", a=a1, b=s, expect={"data": [True, False, False, False, False]})
    case("neq_scalar_nan_" + dt, K + ":2575-2600", "neq", a=a1, b=s, expect={"data": [False, True, True, True, True]})
    case("lt_scalar_nan_" + dt, K + ":2604-2629", "lt", a=a1, b=s, expect={"data": [False, True, True, True, True]})
    case("lt_eq_scalar_nan_" + dt, K + ":2604-2629", "lt_eq", a=a1, b=s, expect={"data": [True, True, True, True, True]})
    case("gt_scalar_nan_" + dt, K + ":2633-2664", "gt", a=a1, b=s, expect={"data": [False, False, False, False, False]})
    case("gt_eq_scalar_nan_" + dt, K + ":2633-2664", "gt_eq", a=a1, b=s, expect={"data": [True, False, False, False, False]})
za, zb = arr("float32", [0.0, "-0.0"]), arr("float32", ["-0.0", 0.0])
case("floating_zeros_eq", K + ":3558-3564", "eq", a=za, b=zb, expect={"data": [False, False]})
case("floating_zeros_eq_scalar_pos", K + ":3566-3569", "eq", a=za, b=arr("float32", [0.0], scalar=True), expect={"data": [True, False]})
case("floating_zeros_eq_scalar_neg", K + ":3571-3574", "eq", a=za, b=arr("float32", ["-0.0"], scalar=True), expect={"data": [False, True]})
C = "arrow-ord/src/cmp.rs"
l, r = arr("int32", [0, 1, 2, 3, 4]), arr("int32", [4, 3, 2, 1, 0])
case("distinct_non_nulls", C + ":1028-1041", "distinct", a=l, b=r, expect={"data": [True, True, False, True, True]})
case("not_distinct_non_nulls", C + ":1028-1041", "not_distinct", a=l, b=r, expect={"data": [False, False, True, False, False]})
l = {"dtype": "int32", "raw_values": [0, 0, 1, 3, 0, 0], "raw_validity": [True, True, False, True, True, True]}
r = {"dtype": "int32", "raw_values": [0] * 6, "raw_validity": [True, False, False, False, True, False]}
case("distinct_nulls", C + ":1044-1067", "distinct", a=l, b=r, expect={"data": [False, True, False, True, False, True]})
case("not_distinct_nulls", C + ":1044-1067", "not_distinct", a=l, b=r, expect={"data": [True, False, True, False, True, False]})
s12, null1 = arr("int32", [12], scalar=True), arr("int32", [None])
case("distinct_scalar_scalar", C + ":1071-1074", "distinct", a=s12, b=s12, expect={"data": [False]})
case("not_distinct_scalar_scalar", C + ":1071-1074", "not_distinct", a=s12, b=s12, expect={"data": [True]})
case("distinct_scalar_vs_null_array", C + ":1076-1081", "distinct", a=s12, b=null1, expect={"data": [True]})
case("not_distinct_scalar_vs_null_array", C + ":1076-1081", "not_distinct", a=s12, b=null1, expect={"data": [False]})
case("distinct_null_array_vs_scalar", C + ":1080", "distinct", a=null1, b=s12, expect={"data": [True]})
nulls = arr("int32", [None], scalar=True)
case("distinct_scalar_vs_null_scalar", C + ":1083-1085", "distinct", a=s12, b=nulls, expect={"data": [True]})
case("not_distinct_null_scalar_self", C + ":1087-1088", "not_distinct", a=nulls, b=nulls, expect={"data": [True]})
case("distinct_null_scalar_self", C + ":1087", "distinct", a=nulls, b=nulls, expect={"data": [False]})
av = {"dtype": "int32", "raw_values": [0, 1, 2, 3], "raw_validity": [False, False, True, True]}
case("distinct_array_vs_null_scalar", C + ":1090-1096", "distinct", a=av, b=nulls, expect={"data": [False, False, True, True]})
case("distinct_null_scalar_vs_array", C + ":1096", "distinct", a=nulls, b=av, expect={"data": [False, False, True, True]})
case("not_distinct_array_vs_null_scalar", C + ":1098-1100", "not_distinct", a=av, b=nulls, expect={"data": [True, True, False, False]})
s1 = arr("int32", [1], scalar=True)
case("distinct_array_vs_scalar_1", C + ":1102-1105", "distinct", a=av, b=s1, expect={"data": [True, True, True, True]})
case("not_distinct_scalar_1_vs_array", C + ":1106-1108", "not_distinct", a=s1, b=av, expect={"data": [False, False, False, False]})
s3 = arr("int32", [3], scalar=True)
case("distinct_array_vs_scalar_3", C + ":1110-1113", "distinct", a=av, b=s3, expect={"data": [True, True, True, False]})
case("not_distinct_array_vs_scalar_3", C + ":1114-1116", "not_distinct", a=s3, b=av, expect={"data": [False, False, False, True]})
s54 = arr("int32", [54], scalar=True)
case("scalar_negation_eq", C + ":1151-1159", "eq", a=s54, b=s54, expect={"data": [True]})
case("scalar_negation_neq", C + ":1151-1159", "neq", a=s54, b=s54, expect={"data": [False]})
case("scalar_empty", C + ":1162-1169", "eq", a=arr("int32", [], force_validity=True), b=arr("int32", [23], scalar=True), expect={"data": []})
case("cmp_length_mismatch", C + ":228-232", "eq", a=arr("int32", [1, 2, 3]), b=arr("int32", [1, 2]),
     expect_error="Invalid argument error: Cannot compare arrays of different lengths, got 3 vs 2")
case("eq_null_scalar_is_all_null", C + ":353,364", "eq", a=arr("int32", [1, 2]), b=nulls, expect={"data": [None, None]})
case("lt_nullable_both", C + ":345", "lt", a=arr("int64", [1, None, 3, 4]), b=arr("int64", [2, 2, None, 4]),
     expect={"data": [True, None, None, False]})

# =========================================================================================
# cast — arrow-cast/src/cast/mod.rs
# =========================================================================================
M = "arrow-cast/src/cast/mod.rs"
i64v = [I64_MIN, I32_MIN, I16_MIN, I8_MIN, 0, I8_MAX, I16_MAX, I32_MAX, I64_MAX]
src = arr("int64", i64v)
case("cast_i64_f64", M + ":8449-8481", "cast", a=src, to="float64",
     expect={"data": [-9223372036854775808.0, -2147483648.0, -32768.0, -128.0, 0.0, 127.0, 32767.0, 2147483647.0, 9223372036854775808.0]})
case("cast_i64_f32", M + ":8483-8501", "cast", a=src, to="float32",
     expect={"data": [-9223372036854775808.0, -2147483648.0, -32768.0, -128.0, 0.0, 127.0, 32767.0, 2147483648.0, 9223372036854775808.0], "f32": True})
case("cast_i64_i64", M + ":8524-8539", "cast", a=src, to="int64", expect={"data": i64v})
case("cast_i64_i32", M + ":8541-8556", "cast", a=src, to="int32",
     expect={"data": [None, I32_MIN, I16_MIN, I8_MIN, 0, I8_MAX, I16_MAX, I32_MAX, None]})
case("cast_i64_i16", M + ":8563-8569", "cast", a=src, to="int16",
     expect={"data": [None, None, I16_MIN, I8_MIN, 0, I8_MAX, I16_MAX, None, None]})
f64v = [float(I64_MIN), float(I32_MIN), float(I16_MIN), float(I8_MIN), 0.0, 255.0, 65535.0, 4294967295.0, 18446744073709551616.0]
fsrc = arr("float64", f64v)
case("cast_f64_i64", M + ":7890-7904", "cast", a=fsrc, to="int64",
     expect={"data": [I64_MIN, I32_MIN, I16_MIN, I8_MIN, 0, 255, 65535, 4294967295, None]})
case("cast_f64_i32", M + ":7906-7920", "cast", a=fsrc, to="int32",
     expect={"data": [None, I32_MIN, I16_MIN, I8_MIN, 0, 255, 65535, None, None]})
case("cast_f64_i16", M + ":7922-7928", "cast", a=fsrc, to="int16",
     expect={"data": [None, None, I16_MIN, I8_MIN, 0, 255, None, None, None]})
case("cast_f64_i8", M + ":7930-7936", "cast", a=fsrc, to="int8",
     expect={"data": [None, None, None, I8_MIN, 0, None, None, None, None]})
case("cast_f64_u64", M + ":7938-7953", "cast", a=fsrc, to="uint64",
     expect={"data": [None, None, None, None, 0, 255, 65535, 4294967295, None]})
case("cast_f64_u32", M + ":7955-7970", "cast", a=fsrc, to="uint32",
     expect={"data": [None, None, None, None, 0, 255, 65535, 4294967295, None]})
case("cast_f64_u16", M + ":7972-7978", "cast", a=fsrc, to="uint16",
     expect={"data": [None, None, None, None, 0, 255, 65535, None, None]})
case("cast_f64_u8", M + ":7980-7986", "cast", a=fsrc, to="uint8",
     expect={"data": [None, None, None, None, 0, 255, None, None, None]})
u64v = [0, 255, 65535, 4294967295, 2**64 - 1]
usrc = arr("uint64", u64v)
case("cast_u64_i64", M + ":8181-8185", "cast", a=usrc, to="int64", expect={"data": [0, 255, 65535, 4294967295, None]})
case("cast_u64_i32", M + ":8187-8191", "cast", a=usrc, to="int32", expect={"data": [0, 255, 65535, None, None]})
case("cast_u64_i8", M + ":8199-8203", "cast", a=usrc, to="int8", expect={"data": [0, None, None, None, None]})
case("cast_u64_u32", M + ":8211-8215", "cast", a=usrc, to="uint32", expect={"data": [0, 255, 65535, 4294967295, None]})
case("cast_u64_u8", M + ":8223-8227", "cast", a=usrc, to="uint8", expect={"data": [0, 255, None, None, None]})
case("cast_u64_f32", M + ":8154-8161", "cast", a=usrc, to="float32",
     expect={"data": [0.0, 255.0, 65535.0, 4294967296.0, 18446744073709551616.0], "f32": True})
case("cast_i64_f64_nulls_zeroed", "arrow-array/src/array/primitive_array.rs:1065-1103", "cast",
     a=arr("int64", [1, None, 3]), to="float64", expect={"data": [1.0, None, 3.0], "always_validity": True})
case("cast_unsafe_error", M + ":2575-2591", "cast", a=arr("int64", [1, I64_MAX]), to="int32", safe=False,
     expect_error="Cast error: Can't cast value 9223372036854775807 to type Int32")
case("cast_dict_utf8", "arrow-cast/src/cast/dictionary.rs:310-317", "take_utf8",
     values={"strings": ["one", None, "three"]}, indices=arr("int32", [0, 1, 2, None, 0, 2]),
     expect={"strings": ["one", None, "three", None, "one", "three"]})

# =========================================================================================
# aggregate — arrow-arith/src/aggregate.rs
# =========================================================================================
G = "arrow-arith/src/aggregate.rs"
case("sum_i32", G + ":1039", "sum", a=arr("int32", [1, 2, 3, 4, 5]), expect={"scalar": 15})
case("sum_f64", G + ":1045", "sum", a=arr("float64", [1.1, 2.2, 3.3, 4.4, 5.5]), expect={"scalar": 16.5})
case("sum_all_nulls", G + ":1118", "sum", a=arr("int32", [None, None, None]), expect={"scalar": None})
case("sum_large_f64", G + ":1124-1127", "sum", a=arr("float64", [float(x) for x in range(1, 101)]), expect={"scalar": 5050.0})
case("sum_large_f64_nullable", G + ":1129-1137", "sum",
     a={"dtype": "float64", "raw_values": [float(x) for x in range(1, 101)], "raw_validity": [x % 3 == 0 for x in range(1, 101)]},
     expect={"scalar": float(sum(i for i in range(1, 101) if i % 3 == 0))})
case("sum_overflow_wraps", G + ":1985-1991", "sum", a=arr("int32", [I32_MAX, 1]), expect={"scalar": I32_MIN})
case("min_i32", G + ":1299", "min", a=arr("int32", [5, 6, 7, 8, 9]), expect={"scalar": 5})
case("max_i32", G + ":1299", "max", a=arr("int32", [5, 6, 7, 8, 9]), expect={"scalar": 9})
case("min_nulls", G + ":1306", "min", a=arr("int32", [5, None, None, 8, 9]), expect={"scalar": 5})
case("max_nulls", G + ":1306", "max", a=arr("int32", [5, None, None, 8, 9]), expect={"scalar": 9})
case("min_1", G + ":1313", "min", a=arr("int32", [None, None, 5, 2]), expect={"scalar": 2})
case("max_1", G + ":1313", "max", a=arr("int32", [None, None, 5, 2]), expect={"scalar": 5})
for name, v in [("neg_inf", "-inf"), ("f64_min", -F64_MAX), ("f64_max", F64_MAX), ("inf", "inf")]:
    case("min_edge_" + name, G + ":1380-1396", "min", a=arr("float64", [v] * 100), expect={"scalar": v})
    case("max_edge_" + name, G + ":1380-1396", "max", a=arr("float64", [v] * 100), expect={"scalar": v})
case("min_all_nans", G + ":1399", "min", a=arr("float64", ["nan"] * 100), expect={"scalar": "nan"})
case("max_all_nans", G + ":1399", "max", a=arr("float64", ["nan"] * 100), expect={"scalar": "nan"})
nn = arr("float64", ["-inf", "+nan", "inf", "-nan"])
case("max_negative_nan", G + ":1406-1416", "max", a=nn, expect={"scalar": "+nan"})
case("min_negative_nan", G + ":1406-1416", "min", a=nn, expect={"scalar": "-nan"})
first_nan = ["nan"] + [float(i) for i in range(1, 100)]
case("min_first_nan_nonnull", G + ":1419", "min", a=arr("float64", first_nan), expect={"scalar": 1.0})
case("max_first_nan_nonnull", G + ":1419", "max", a=arr("float64", first_nan), expect={"scalar": "nan"})
last_nan = [float(i + 1) for i in range(99)] + ["nan"]
case("min_last_nan_nonnull", G + ":1434", "min", a=arr("float64", last_nan), expect={"scalar": 1.0})
case("max_last_nan_nonnull", G + ":1434", "max", a=arr("float64", last_nan), expect={"scalar": "nan"})
fnn = ["nan" if i == 0 else (None if i % 2 == 0 else float(i)) for i in range(100)]
case("min_first_nan_nullable", G + ":1449", "min", a=arr("float64", fnn), expect={"scalar": 1.0})
case("max_first_nan_nullable", G + ":1449", "max", a=arr("float64", fnn), expect={"scalar": "nan"})
lnn = ["nan" if i == 99 else (None if i % 2 == 0 else float(i)) for i in range(100)]
case("min_last_nan_nullable", G + ":1466", "min", a=arr("float64", lnn), expect={"scalar": 1.0})
case("max_last_nan_nullable", G + ":1466", "max", a=arr("float64", lnn), expect={"scalar": "nan"})
mix = [{0: "-inf", 1: -F64_MAX, 2: F64_MAX, 4: "inf", 5: "nan"}.get(i % 10, float(i)) for i in range(100)]
case("min_inf_and_nans", G + ":1483", "min", a=arr("float64", mix), expect={"scalar": "-inf"})
case("max_inf_and_nans", G + ":1483", "max", a=arr("float64", mix), expect={"scalar": "nan"})

# =========================================================================================
# boolean — arrow-arith/src/boolean.rs (predicate construction)
# =========================================================================================
B = "arrow-arith/src/boolean.rs"
T, Fa = True, False
case("bool_and", B + ":364", "and_", a=arr("bool", [Fa, Fa, T, T]), b=arr("bool", [Fa, T, Fa, T]), expect={"data": [Fa, Fa, Fa, T], "no_validity": True})
case("bool_or", B + ":375", "or_", a=arr("bool", [Fa, Fa, T, T]), b=arr("bool", [Fa, T, Fa, T]), expect={"data": [Fa, T, T, T], "no_validity": True})
case("bool_and_not", B + ":386", "and_not", a=arr("bool", [Fa, Fa, T, T]), b=arr("bool", [Fa, T, Fa, T]), expect={"data": [Fa, Fa, T, Fa]})
case("bool_and_not_sliced", B + ":398", "and_not", a=arr("bool", [T, Fa, T, Fa, T, Fa, T], slice=(2, 3)),
     b=arr("bool", [Fa, T, Fa, T, Fa, T, Fa], slice=(2, 3)), expect={"data": [T, Fa, T]})  # == and(a, not(b)) evaluated by hand
case("bool_and_not_sliced_different_offsets", B + ":410", "and_not", a=arr("bool", [Fa, T, T, Fa, T, Fa, T], slice=(1, 4)),
     b=arr("bool", [T, Fa, Fa, T, Fa, T, Fa], slice=(2, 4)), expect={"data": [T, Fa, Fa, Fa]})
NINE_A = [None, None, None, Fa, Fa, Fa, T, T, T]
NINE_B = [None, Fa, T, None, Fa, T, None, Fa, T]
case("bool_or_nulls", B + ":422", "or_", a=arr("bool", NINE_A), b=arr("bool", NINE_B), expect={"data": [None, None, None, None, Fa, T, None, T, T]})
case("bool_and_nulls", B + ":643", "and_", a=arr("bool", NINE_A), b=arr("bool", NINE_B), expect={"data": [None, None, None, None, Fa, Fa, None, Fa, T]})
case("bool_and_kleene_nulls", B + ":473", "and_kleene", a=arr("bool", NINE_A), b=arr("bool", NINE_B),
     expect={"data": [None, Fa, None, Fa, Fa, Fa, None, Fa, T]})
case("bool_or_kleene_nulls", B + ":514", "or_kleene", a=arr("bool", NINE_A), b=arr("bool", NINE_B),
     expect={"data": [None, None, T, None, Fa, T, T, T, T]})
case("bool_or_kleene_right_sided_nulls", B + ":555", "or_kleene", a=arr("bool", [Fa, Fa, Fa, T, T, T]), b=arr("bool", [T, Fa, None, T, Fa, None]),
     expect={"data": [T, Fa, None, T, T, T]})
case("bool_or_kleene_left_sided_nulls", B + ":588", "or_kleene", a=arr("bool", [T, Fa, None, T, Fa, None]), b=arr("bool", [Fa, Fa, Fa, T, T, T]),
     expect={"data": [T, Fa, None, T, T, T]})
case("bool_kleene_no_remainder", B + ":463", "or_kleene", a=arr("bool", [T] * 1024), b=arr("bool", [None] * 1024),
     expect={"data": [T] * 1024})
case("bool_and_kleene_doc", B + ":50-57", "and_kleene", a=arr("bool", [T, Fa, None]), b=arr("bool", [None, None, None]), expect={"data": [None, Fa, None]})
case("bool_or_kleene_doc", B + ":146-153", "or_kleene", a=arr("bool", [T, Fa, None]), b=arr("bool", [None, None, None]), expect={"data": [T, None, None]})
case("bool_not", B + ":621", "not_", a=arr("bool", [Fa, T]), expect={"data": [T, Fa], "no_validity": True})
case("bool_not_sliced", B + ":631", "not_", a=arr("bool", [None, T, Fa, None, T], slice=(1, 4)), expect={"data": [Fa, T, None, Fa]})
TW_A = [Fa] * 10 + [T, T]
TW_B = [Fa] * 9 + [T, Fa, T]
case("bool_and_sliced_same_offset", B + ":684", "and_", a=arr("bool", TW_A, slice=(8, 4)), b=arr("bool", TW_B, slice=(8, 4)), expect={"data": [Fa, Fa, Fa, T]})
case("bool_and_sliced_same_offset_mod8", B + ":705", "and_", a=arr("bool", [Fa, Fa, T, T] + [Fa] * 8, slice=(0, 4)), b=arr("bool", TW_B, slice=(8, 4)),
     expect={"data": [Fa, Fa, Fa, T]})
case("bool_and_sliced_offset1", B + ":726", "and_", a=arr("bool", TW_A, slice=(8, 4)), b=arr("bool", [Fa, T, Fa, T]), expect={"data": [Fa, Fa, Fa, T]})
case("bool_and_sliced_offset2", B + ":743", "and_", a=arr("bool", [Fa, Fa, T, T]), b=arr("bool", TW_B, slice=(8, 4)), expect={"data": [Fa, Fa, Fa, T]})
case("bool_and_nulls_offset", B + ":760", "and_", a=arr("bool", [None, Fa, T, None, T], slice=(1, 4)), b=arr("bool", [None, None, T, Fa, T, T], slice=(2, 4)),
     expect={"data": [Fa, Fa, None, T]})
case("bool_and_length_mismatch", B + ":232-236", "and_", a=arr("bool", [T, Fa]), b=arr("bool", [T]),
     expect_error="Compute error: Cannot perform bitwise operation on arrays of different length")
case("bool_and_kleene_length_mismatch", B + ":61-65", "and_kleene", a=arr("bool", [T, Fa]), b=arr("bool", [T]),
     expect_error="Compute error: Cannot perform bitwise operation on arrays of different length")
case("is_null_nonnull", B + ":785", "is_null", a=arr("int32", [1, 2, 3, 4]), expect={"data": [Fa] * 4, "no_validity": True})
case("is_null_nonnull_offset", B + ":797", "is_null", a=arr("int32", [1, 2, 3, 4, 5, 6, 7, 8, 7, 6, 5, 4, 3, 2, 1], slice=(8, 4)), expect={"data": [Fa] * 4, "no_validity": True})
case("is_not_null_nonnull", B + ":810", "is_not_null", a=arr("int32", [1, 2, 3, 4]), expect={"data": [T] * 4, "no_validity": True})
case("is_not_null_nonnull_offset", B + ":822", "is_not_null", a=arr("int32", [1, 2, 3, 4, 5, 6, 7, 8, 7, 6, 5, 4, 3, 2, 1], slice=(8, 4)),
     expect={"data": [T] * 4, "no_validity": True})
case("is_null_nullable", B + ":835", "is_null", a=arr("int32", [1, None, 3, None]), expect={"data": [Fa, T, Fa, T], "no_validity": True})
SIXTEEN = [None] * 8 + [1, None, 2, None, 3, 4, None, None]
case("is_null_nullable_offset", B + ":847", "is_null", a=arr("int32", SIXTEEN, slice=(8, 4)), expect={"data": [Fa, T, Fa, T], "no_validity": True})
case("is_not_null_nullable", B + ":878", "is_not_null", a=arr("int32", [1, None, 3, None]), expect={"data": [T, Fa, T, Fa], "no_validity": True})
case("is_not_null_nullable_offset", B + ":890", "is_not_null", a=arr("int32", SIXTEEN, slice=(8, 4)), expect={"data": [T, Fa, T, Fa], "no_validity": True})

T2 = "arrow-select/src/take.rs"
FIVE = ["one", None, "three", "four", "five"]
case("take_string", T2 + ":1702-1726", "take_utf8", values={"strings": FIVE}, indices=arr("uint32", [3, None, 1, 3, 4]),
     expect={"strings": ["four", None, None, "four", "five"]})
case("take_large_string", T2 + ":1728-1731", "take_utf8", values={"strings": FIVE, "large": True}, indices=arr("uint32", [3, None, 1, 3, 4]),
     expect={"strings": ["four", None, None, "four", "five"]})
case("take_slice_string", T2 + ":1733-1744", "take_utf8", values={"strings": ["hello", None, "world", None, "hi"]},
     indices=arr("int32", [0, 1, None, 0, 2], slice=(1, 4)), expect={"strings": [None, None, "hello", "world"]})
SEVEN = ["aaa", "bbb", None, "ccccc", "dd", None, "eeee"]
case("take_bytes_sliced_values_fast_path", T2 + ":1750-1775", "take_utf8", values={"strings": SEVEN, "slice": [2, 5]},
     indices=arr("int32", [1, 2, 4, 1]), expect={"strings": ["ccccc", "dd", "eeee", "ccccc"]})
case("take_bytes_sliced_values_nullable_path", T2 + ":1777-1783", "take_utf8", values={"strings": SEVEN, "slice": [2, 5]},
     indices=arr("int32", [1, None, 0, 4, 3]), expect={"strings": ["ccccc", None, None, "eeee", None]})
F2 = "arrow-select/src/filter.rs"
case("filter_string_array_sliced_values", F2 + ":893-928", "filter_utf8", values={"strings": SEVEN, "slice": [2, 5]},
     predicate=arr("bool", [True, True, False, True, True]), expect={"strings": [None, "ccccc", None, "eeee"]})

A = "arrow-arith/src/aggregate.rs"
case("sum_checked_overflow", A + ":1993", "sum_checked", a=arr("int32", [I32_MAX, 1]),
     expect_error="Arithmetic overflow: Overflow happened on: 2147483647 + 1")
case("sum_checked_prefix_overflow_even_if_total_fits", A + ":897-937", "sum_checked", a=arr("int32", [I32_MAX, 1, -5]),
     expect_error="Arithmetic overflow: Overflow happened on: 2147483647 + 1")  # try_fold stops at the first failing add
case("sum_checked_nulls_skipped", A + ":918-934", "sum_checked", a=arr("int64", [I64_MAX, None, -1, 1]), expect={"scalar": I64_MAX})
case("sum_checked_all_null", A + ":902-904", "sum_checked", a=arr("int32", [None, None]), expect={"scalar": None})
case("sum_checked_ok", A + ":897", "sum_checked", a=arr("int32", [1, 2, 3, 4, 5]), expect={"scalar": 15})

# =========================================================================================
# nullif — arrow-select/src/nullif.rs
# =========================================================================================
NI = "arrow-select/src/nullif.rs"
case("nullif_doc_example", NI + ":31-43", "nullif", left=arr("int32", [None, 8, 1, 9]), right=arr("bool", [False, True, False, None]),
     expect={"data": [None, None, 1, 9]})
case("nullif_int_array", NI + ":127", "nullif", left=arr("int32", [15, None, 8, 1, 9]), right=arr("bool", [False, None, True, False, None]),
     expect={"data": [15, None, None, 1, 9]})
COMP7 = [False, False, False, None, True, False, None]
case("nullif_int_array_offset", NI + ":166", "nullif", left=arr("int32", [None, 15, 8, 1, 9], slice=(1, 3)),
     right=arr("bool", COMP7, slice=(2, 3)), expect={"data": [15, 8, None]})
case("nullif_int_large_left_offset", NI + ":228", "nullif", left=arr("int32", [-1] * 16 + [None, 15, 8, 1, 9], slice=(17, 3)),
     right=arr("bool", COMP7, slice=(2, 3)), expect={"data": [15, 8, None]})
case("nullif_int_large_right_offset", NI + ":278", "nullif", left=arr("int32", [None, 15, 8, 1, 9], slice=(1, 3)),
     right=arr("bool", [False] * 19 + [None, True, False, None], slice=(18, 3)), expect={"data": [15, 8, None]})
case("nullif_boolean_offset", NI + ":327", "nullif", left=arr("bool", [None, True, False, True, True], slice=(1, 3)),
     right=arr("bool", COMP7, slice=(2, 3)), expect={"data": [True, False, None]})
case("nullif_no_nulls", NI + ":466", "nullif", left=arr("int32", [15, 7, 8, 1, 9]), right=arr("bool", [False, None, True, False, None]),
     expect={"data": [15, 7, None, 1, 9]})
case("nullif_nothing_nulled_has_no_null_buffer", NI + ":105-112", "nullif", left=arr("int32", [15, 7, 8]), right=arr("bool", [False, None, False]),
     expect={"data": [15, 7, 8], "no_validity": True})
case("nullif_empty", NI + ":477", "nullif", left=arr("int32", []), right=arr("bool", []), expect={"data": []})
case("nullif_length_mismatch", NI + ":47-51", "nullif", left=arr("int32", [1, 2]), right=arr("bool", [True]),
     expect_error="Compute error: Cannot perform comparison operation on arrays of different length")

# =========================================================================================
# zip — arrow-select/src/zip.rs
# =========================================================================================
Z = "arrow-select/src/zip.rs"
ZA, ZB = [5, None, 7, None, 1], [None, 3, 6, 7, 3]
M1, M2 = [True, True, False, False, True], [False, False, True, True, False]
case("zip_doc_example", Z + ":50-69", "zip", mask=arr("bool", [True, True, False, None, True]), truthy=arr("int32", [1, None, 3, 4, 5]),
     falsy=arr("int32", [10, 20, 30, 40, 50]), expect={"data": [1, None, 30, 40, 5]})
case("zip_doc_example_scalar", Z + ":79-97", "zip", mask=arr("bool", [True, True, False, None, True]), truthy=arr("int32", [1, None, 3, 4, 5]),
     falsy=arr("int32", [42], scalar=True), expect={"data": [1, None, 42, 42, 5]})
case("zip_kernel_one", Z + ":870", "zip", mask=arr("bool", M1), truthy=arr("int32", ZA), falsy=arr("int32", ZB), expect={"data": [5, None, 6, 7, 1]})
case("zip_kernel_two", Z + ":881", "zip", mask=arr("bool", M2), truthy=arr("int32", ZA), falsy=arr("int32", ZB), expect={"data": [None, 3, 7, None, 3]})
case("zip_kernel_scalar_falsy_1", Z + ":892", "zip", mask=arr("bool", M1), truthy=arr("int32", ZA), falsy=arr("int32", [42], scalar=True),
     expect={"data": [5, None, 42, 42, 1]})
case("zip_kernel_scalar_falsy_2", Z + ":905", "zip", mask=arr("bool", M2), truthy=arr("int32", ZA), falsy=arr("int32", [42], scalar=True),
     expect={"data": [42, 42, 7, None, 42]})
case("zip_kernel_scalar_truthy_1", Z + ":918", "zip", mask=arr("bool", M1), truthy=arr("int32", [42], scalar=True), falsy=arr("int32", ZA),
     expect={"data": [42, 42, 7, None, 42]})
case("zip_kernel_scalar_truthy_2", Z + ":931", "zip", mask=arr("bool", M2), truthy=arr("int32", [42], scalar=True), falsy=arr("int32", ZA),
     expect={"data": [5, None, 42, 42, 1]})
case("zip_kernel_scalar_both_mask_ends_with_true", Z + ":944", "zip", mask=arr("bool", M1), truthy=arr("int32", [42], scalar=True),
     falsy=arr("int32", [123], scalar=True), expect={"data": [42, 42, 123, 123, 42], "no_validity": True})
case("zip_kernel_scalar_both_mask_ends_with_false", Z + ":956", "zip", mask=arr("bool", [True, True, False, True, False, False]),
     truthy=arr("int32", [42], scalar=True), falsy=arr("int32", [123], scalar=True), expect={"data": [42, 42, 123, 42, 123, 123]})
case("zip_kernel_primitive_scalar_none_1", Z + ":975", "zip", mask=arr("bool", M1), truthy=arr("int32", [42], scalar=True),
     falsy=arr("int32", [None], scalar=True), expect={"data": [42, 42, None, None, 42]})
case("zip_kernel_primitive_scalar_none_2", Z + ":987", "zip", mask=arr("bool", M2), truthy=arr("int32", [42], scalar=True),
     falsy=arr("int32", [None], scalar=True), expect={"data": [None, None, 42, 42, None]})
case("zip_kernel_primitive_scalar_both_null", Z + ":999", "zip", mask=arr("bool", M2), truthy=arr("int32", [None], scalar=True),
     falsy=arr("int32", [None], scalar=True), expect={"data": [None] * 5})
MN = {"dtype": "bool", "raw_values": [True, True, False, True, False, False], "raw_validity": [True, True, True, False, True, True]}
case("zip_primitive_array_mask_nulls_treated_as_false", Z + ":1011", "zip", mask=MN, truthy=arr("int32", [1, 2, 3, 4, 5, 6]),
     falsy=arr("int32", [7, 8, 9, 10, 11, 12]), expect={"data": [1, 2, 9, 10, 11, 12], "no_validity": True})
case("zip_primitive_scalar_mask_nulls_treated_as_false", Z + ":1038", "zip", mask=MN, truthy=arr("int32", [42], scalar=True),
     falsy=arr("int32", [123], scalar=True), expect={"data": [42, 42, 123, 123, 123, 123]})
case("zip_length_mismatch", Z + ":128-132", "zip", mask=arr("bool", [True, False]), truthy=arr("int32", [1, 2, 3]), falsy=arr("int32", [1, 2]),
     expect_error="Invalid argument error: all arrays should have the same length")

# =========================================================================================
# cmp on Binary / Utf8 / Utf8View — arrow-ord/src/comparison.rs (test_binary!, test_utf8!, test_utf8_view! macros)
# byte strings are lists of ints when not ASCII
# =========================================================================================
CP = "arrow-ord/src/comparison.rs"
FF8, FF9 = [0xff, 0xf8], [0xff, 0xf9]
BL = ["arrow", "datafusion", "flight", "parquet"]
for name, line, op, left, right, exp in [
    ("binary_array_eq", 968, "eq", ["arrow", "arrow", "arrow", "arrow", FF8], ["arrow", "parquet", "datafusion", "flight", FF8], [True, False, False, False, True]),
    ("binary_array_neq", 984, "neq", ["arrow", "arrow", "arrow", "arrow", FF8], ["arrow", "parquet", "datafusion", "flight", FF9], [False, True, True, True, True]),
    ("binary_array_lt", 999, "lt", BL + [FF8], ["flight"] * 4 + [FF9], [True, True, False, False, True]),
    ("binary_array_lt_eq", 1014, "lt_eq", BL + [FF8], ["flight"] * 4 + [[0xff, 0xf8, 0xf9]], [True, True, True, False, True]),
    ("binary_array_gt", 1029, "gt", BL + [FF9], ["flight"] * 4 + [FF8], [False, False, False, True, True]),
    ("binary_array_gt_eq", 1044, "gt_eq", BL + [FF8], ["flight"] * 4 + [FF8], [False, False, True, True, True]),
]:
    for large in (False, True):
        case(name + ("_large" if large else ""), f"{CP}:{line}", "cmp_bytes", cmp=op, large=large, left={"bytes": left}, right={"bytes": right}, expect={"data": exp})
for name, line, op, left, right, exp in [
    ("binary_array_eq_scalar", 976, "eq", ["arrow", "parquet", "datafusion", "flight", FF8], "arrow", [True, False, False, False, False]),
    ("binary_array_neq_scalar", 991, "neq", ["arrow", "parquet", "datafusion", "flight", FF8], "arrow", [False, True, True, True, True]),
    ("binary_array_lt_scalar", 1006, "lt", BL + [FF8], "flight", [True, True, False, False, False]),
    ("binary_array_lt_eq_scalar", 1021, "lt_eq", BL + [FF8], "flight", [True, True, True, False, False]),
    ("binary_array_gt_scalar", 1036, "gt", BL + [FF8], "flight", [False, False, False, True, True]),
    ("binary_array_gt_eq_scalar", 1051, "gt_eq", BL + [FF8], "flight", [False, False, True, True, True]),
]:
    case(name, f"{CP}:{line}", "cmp_bytes", cmp=op, large=False, left={"bytes": left}, right={"bytes": [right], "scalar": True}, expect={"data": exp})
case("binary_eq_scalar_on_slice", CP + ":937", "cmp_bytes", cmp="eq", large=False, left={"bytes": ["hi", None, "hello", "world"], "slice": [1, 3]},
     right={"bytes": ["hello"], "scalar": True}, expect={"data": [None, True, False]})
case("utf8_eq_scalar_on_slice", CP + ":1147", "cmp_bytes", cmp="eq", large=False, left={"bytes": ["hi", None, "hello", "world", ""], "slice": [1, 4]},
     right={"bytes": ["hello"], "scalar": True}, expect={"data": [None, True, False, False]})
case("utf8_eq_empty_scalar_on_slice", CP + ":1157", "cmp_bytes", cmp="eq", large=False, left={"bytes": ["hi", None, "hello", "world", ""], "slice": [1, 4]},
     right={"bytes": [""], "scalar": True}, expect={"data": [None, False, False, True]})
UA = ["arrow", "arrow", "arrow", "arrow"]
UB = ["arrow", "parquet", "datafusion", "flight"]
for name, line, op, left, right, exp in [
    ("utf8_array_eq", 1246, "eq", UA, UB, [True, False, False, False]), ("utf8_array_neq", 1282, "neq", UA, UB, [False, True, True, True]),
    ("utf8_array_lt", 1318, "lt", BL, ["flight"] * 4, [True, True, False, False]), ("utf8_array_lt_eq", 1354, "lt_eq", BL, ["flight"] * 4, [True, True, True, False]),
    ("utf8_array_gt", 1383, "gt", BL, ["flight"] * 4, [False, False, False, True]), ("utf8_array_gt_eq", 1419, "gt_eq", BL, ["flight"] * 4, [False, False, True, True]),
]:
    case(name, f"{CP}:{line}", "cmp_bytes", cmp=op, large=False, left={"bytes": left}, right={"bytes": right}, expect={"data": exp})
    case(name + "_large", f"{CP}:{line}", "cmp_bytes", cmp=op, large=True, left={"bytes": left}, right={"bytes": right}, expect={"data": exp})
LARGE_1, LARGE_2, SMALL_1, SMALL_2 = "prefix-larger than 12 bytes string", "prefix-larger but different string", "pref1", "pref2"
TA1, TA2 = [LARGE_1, LARGE_1, SMALL_1, SMALL_1, LARGE_1], [LARGE_1, LARGE_2, SMALL_1, SMALL_2, SMALL_1]
for name, line, op, exp in [("utf8_view_array_eq", 1253, "eq", [True, False, True, False, False]), ("utf8_view_array_neq", 1289, "neq", [False, True, False, True, True]),
                            ("utf8_view_array_lt", 1325, "lt", [False, False, False, True, False]), ("utf8_view_array_lt_eq", 1361, "lt_eq", [True, False, True, True, False]),
                            ("utf8_view_array_gt", 1390, "gt", [False, True, False, False, True]), ("utf8_view_array_gt_eq", 1426, "gt_eq", [True, True, True, False, True])]:
    case(name, f"{CP}:{line}", "cmp_view", cmp=op, left={"bytes": TA1}, right={"bytes": TA2}, expect={"data": exp})
for name, line, op, sc, exp in [
    ("utf8_view_array_eq_large_scalar", 1267, "eq", LARGE_1, [True, False, False, False, False]), ("utf8_view_array_eq_small_scalar", 1274, "eq", SMALL_1, [False, False, True, False, True]),
    ("utf8_view_array_neq_scalar", 1303, "neq", LARGE_1, [False, True, True, True, True]), ("utf8_view_array_lt_scalar", 1339, "lt", LARGE_1, [False, True, True, True, True]),
    ("utf8_view_array_lt_scalar_small", 1346, "lt", SMALL_1, [False, False, False, False, False]), ("utf8_view_array_lt_eq_scalar", 1375, "lt_eq", LARGE_1, [True, True, True, True, True]),
    ("utf8_view_array_gt_scalar", 1404, "gt", LARGE_1, [False, False, False, False, False]), ("utf8_view_array_gt_scalar_small", 1411, "gt", SMALL_1, [True, True, False, True, False]),
    ("utf8_view_array_gt_eq_scalar", 1440, "gt_eq", LARGE_1, [True, False, False, False, False]), ("utf8_view_array_gt_eq_scalar_small", 1447, "gt_eq", SMALL_1, [True, True, True, True, True])]:
    case(name, f"{CP}:{line}", "cmp_view", cmp=op, left={"bytes": TA2}, right={"bytes": [sc], "scalar": True}, expect={"data": exp})

# =========================================================================================
# concat — arrow-select/src/concat.rs
# =========================================================================================
CC = "arrow-select/src/concat.rs"
case("concat_empty_vec", CC + ":698", "concat", arrays=[], expect_error="Compute error: concat requires input of at least one array")
case("concat_one_element_vec", CC + ":718", "concat", arrays=[arr("int64", [-1, 2, None])], expect={"data": [-1, 2, None]})
case("concat_primitive_arrays", CC + ":880", "concat",
     arrays=[arr("int64", [-1, -1, 2, None, None]), arr("int64", [101, 102, 103, None]), arr("int64", [256, 512, 1024])],
     expect={"data": [-1, -1, 2, None, None, 101, 102, 103, None, 256, 512, 1024]})
case("concat_primitive_array_slices", CC + ":907", "concat",
     arrays=[arr("int64", [-1, -1, 2, None, None], slice=(1, 3)), arr("int64", [101, 102, 103, None], slice=(1, 3))],
     expect={"data": [-1, 2, None, 102, 103, None]})
case("concat_boolean_primitive_arrays", CC + ":930", "concat",
     arrays=[arr("bool", [True, True, False, None, None, False]), arr("bool", [None, False, True, False])],
     expect={"data": [True, True, False, None, None, False, None, False, True, False]})
case("concat_no_nulls_has_no_null_buffer", CC + ":334-343", "concat", arrays=[arr("int32", [1, 2]), arr("int32", [3])],
     expect={"data": [1, 2, 3], "no_validity": True})
case("concat_string_arrays", CC + ":832", "concat_utf8",
     arrays=[{"strings": ["hello", "world"]}, {"strings": ["2", "3", "4"]}, {"strings": ["foo", "bar", None, "baz"]}],
     expect={"strings": ["hello", "world", "2", "3", "4", "foo", "bar", None, "baz"]})
case("concat_string_array_slices", CC + ":355-368", "concat_utf8",
     arrays=[{"strings": ["hello", "world", "x"], "slice": [1, 2]}, {"strings": ["a", None, "bcd", "e"], "slice": [1, 3]}],
     expect={"strings": ["world", "x", None, "bcd", "e"]})

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
with open(out, "w") as f:
    json.dump({"reference": "apache/arrow-rs 59.2.0 @ cd7c6b83", "cases": cases}, f, indent=0)
print(f"wrote {len(cases)} cases to {out}")
