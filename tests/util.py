"""Comparison helpers shared by the parity tests."""
from __future__ import annotations

import math

import pyarrow as pa
import pyarrow.compute as pc


def canon(table: pa.Table, sort: bool = True) -> pa.Table:
    """Canonical form: combined chunks, rows sorted by all columns (row order is engine-defined unless the
    plan has an ORDER BY -- SURVEY.md 7.3-7)."""
    if isinstance(table, pa.RecordBatch):
        table = pa.Table.from_batches([table])
    table = table.combine_chunks()
    if sort and table.num_rows > 1:
        keys = [(n, "ascending") for n in table.column_names]
        idx = pc.sort_indices(table, sort_keys=keys, null_placement="at_start")
        table = table.take(idx)
    return table


def assert_tables_equal(a, b, sort: bool = True, f64_rtol: float = 0.0, check_names: bool = True):
    """Bit-exact for everything but Float64 columns, which are compared with `f64_rtol`."""
    a, b = canon(a, sort), canon(b, sort)
    assert a.num_columns == b.num_columns, (a.schema, b.schema)
    assert a.num_rows == b.num_rows, f"row count {a.num_rows} != {b.num_rows}"
    for i in range(a.num_columns):
        ca, cb = a.column(i), b.column(i)
        fa, fb = a.schema.field(i), b.schema.field(i)
        if check_names:
            assert fa.name == fb.name, (fa.name, fb.name)
        assert fa.type == fb.type, f"column {fa.name}: {fa.type} != {fb.type}"
        if pa.types.is_floating(fa.type) and f64_rtol > 0:
            la, lb = ca.to_pylist(), cb.to_pylist()
            for x, y in zip(la, lb):
                if x is None or y is None:
                    assert x is None and y is None, (fa.name, x, y)
                elif math.isnan(x) or math.isnan(y):
                    assert math.isnan(x) and math.isnan(y)
                else:
                    assert abs(x - y) <= f64_rtol * max(abs(x), abs(y), 1e-300), (fa.name, x, y)
        elif pa.types.is_floating(fa.type):
            # bit-exact except that any NaN equals any NaN (pyarrow's equals() treats NaN != NaN)
            la, lb = ca.to_pylist(), cb.to_pylist()
            for x, y in zip(la, lb):
                if x is None or y is None:
                    assert x is None and y is None, (fa.name, x, y)
                elif math.isnan(x) or math.isnan(y):
                    assert math.isnan(x) and math.isnan(y), (fa.name, x, y)
                else:
                    assert x == y and math.copysign(1.0, x) == math.copysign(1.0, y), (fa.name, x, y)
        else:
            assert ca.equals(cb), f"column {fa.name} differs:\n{ca.to_pylist()[:20]}\nvs\n{cb.to_pylist()[:20]}"


def stats_tuples(stats):
    return sorted(s.as_tuple() for s in stats)
