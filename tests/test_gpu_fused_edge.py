"""GPU parity of the fused scan->filter->project->aggregate kernel (csrc/device/fused.cuh) on inputs
that leave its fast paths: the q1/q6 stage plans over hand-made lineitem batches with negative,
wide (> 31 bit, > 40 bit, products beyond 64 bit) decimals, long / empty strings, more groups than the register
directory holds, ragged and tiny row counts, unaligned slices.  CUDA engine (through the C-ABI) vs the
CPU oracle, bit-exact."""
import datetime
import decimal

import numpy as np
import pyarrow as pa
import pytest

import ballista_b200 as bb
from ballista_b200 import driver, tpch
from util import assert_tables_equal

pytestmark = pytest.mark.gpu

D = decimal.Decimal
DEC = pa.decimal128(15, 2)
EPOCH = datetime.date(1970, 1, 1)


def _lineitem(n, seed, qty=None, price=None, disc=None, tax=None, flags=None, status=None, dates=None):
    rng = np.random.default_rng(seed)

    def dec(vals):
        return pa.array([D(int(v)).scaleb(-2) for v in vals], type=DEC)

    qty = qty if qty is not None else rng.integers(100, 5001, n)
    price = price if price is not None else rng.integers(90000, 10495001, n)
    disc = disc if disc is not None else rng.integers(0, 11, n)
    tax = tax if tax is not None else rng.integers(0, 9, n)
    flags = flags if flags is not None else rng.choice(["A", "N", "R"], n)
    status = status if status is not None else rng.choice(["F", "O"], n)
    dates = dates if dates is not None else rng.integers(8036, 10592, n)  # 1992-01-02 .. 1998-12-31
    return pa.record_batch(
        [dec(qty), dec(price), dec(disc), dec(tax), pa.array(list(flags), type=pa.utf8()), pa.array(list(status), type=pa.utf8()),
         pa.array(np.asarray(dates, dtype=np.int32), type=pa.date32())],
        names=tpch.Q1_COLUMNS)


def _run_both(gpu, oracle, batches, stages, job):
    for e in (gpu, oracle):
        e.drop_table("lineitem")
        for p, b in enumerate(batches):
            e.register_batch("lineitem", p, b)
    got = driver.run_stages(gpu, stages, job)
    want = driver.run_stages(oracle, stages, job)
    return got, want


@pytest.mark.parametrize("n", [0, 1, 31, 63, 64, 65, 127, 4097, 100003])
def test_q1_row_counts(gpu, oracle, n):
    """ragged last warp tile, fewer rows than one tile, empty input"""
    f0, s0 = gpu.counter("fused"), gpu.counter("fused_static")
    got, want = _run_both(gpu, oracle, [_lineitem(n, 1 + n)], tpch.q1(3), f"edge-n{n}")
    assert_tables_equal(got, want, sort=False)
    if n > 0:  # the ahead-of-time q1 shape of the fused kernel is what ran, not the tile VM
        assert gpu.counter("fused") > f0 and gpu.counter("fused_static") > s0


def test_q1_negative_and_wide_values(gpu, oracle):
    """values outside [0, 2^31): negative quantities/prices (returns), prices beyond 32 and 64 bits --
    the speculative narrow path must hand these warp tiles to the exact general path / slow rows"""
    n = 20000
    rng = np.random.default_rng(7)
    price = rng.integers(90000, 10495001, n).astype(object)
    qty = rng.integers(100, 5001, n).astype(object)
    for i in range(0, n, 97):
        price[i] = -int(price[i])                      # negative
    for i in range(5, n, 211):
        price[i] = 3_000_000_000 + i                   # > 2^31
    for i in range(11, n, 1013):
        price[i] = 9_000_000_000_000 + i               # product*scale leaves 2^40 (large addend path)
    for i in range(3, n, 499):
        qty[i] = -int(qty[i])
    b = _lineitem(n, 8, qty=qty, price=price)
    f0 = gpu.counter("fused")
    got, want = _run_both(gpu, oracle, [b], tpch.q1(2), "edge-wide")
    assert_tables_equal(got, want, sort=False)
    assert gpu.counter("fused") > f0


def test_q1_many_groups_and_long_strings(gpu, oracle):
    """more groups than the 4-entry register directory (falls back to the hash-table sink), strings longer
    than the packed key image, empty strings"""
    n = 50000
    rng = np.random.default_rng(11)
    flags = rng.choice(["A", "N", "R", "", "returned-long-flag", "Z9", "abc", "abcd"], n)
    status = rng.choice(["F", "O", "", "a-status-longer-than-seven"], n)
    b = _lineitem(n, 12, flags=flags, status=status)
    got, want = _run_both(gpu, oracle, [b], tpch.q1(4), "edge-groups")
    assert_tables_equal(got, want, sort=False)


def test_q1_late_fifth_group(gpu, oracle):
    """four groups for almost the whole table, a fifth one only at the very end: the fused kernel gives
    up late (overflow flag) and the engine re-runs the aggregate on the hash-table sink"""
    n = 300000
    rng = np.random.default_rng(13)
    flags = list(rng.choice(["A", "N"], n))
    status = list(rng.choice(["F", "O"], n))
    flags[-1] = "R"
    b = _lineitem(n, 14, flags=flags, status=status)
    got, want = _run_both(gpu, oracle, [b], tpch.q1(2), "edge-late5")
    assert_tables_equal(got, want, sort=False)


def test_q1_unaligned_slices(gpu, oracle):
    """zero-copy slices of a larger batch: column pointers lose their 16-byte alignment (no TMA path)"""
    base = _lineitem(5000, 15)
    batches = [base.slice(1, 1777), base.slice(1778, 3001)]
    got, want = _run_both(gpu, oracle, batches, tpch.q1(3), "edge-slices")
    assert_tables_equal(got, want, sort=False)


def test_q1_extreme_in_range_values(gpu, oracle):
    """largest Decimal128(15,2) prices: disc_price ~2^57, charge beyond 2^64 -- every addend takes the
    large-addend path (merged straight into the global table), mixed with ordinary rows"""
    n = 5000
    rng = np.random.default_rng(16)
    price = rng.integers(90000, 10495001, n).astype(object)
    for i in range(0, n, 3):
        price[i] = 10 ** 15 - 1 - i
    b = _lineitem(n, 16, price=price)
    got, want = _run_both(gpu, oracle, [b], tpch.q1(1), "edge-max")
    assert_tables_equal(got, want, sort=False)


@pytest.mark.parametrize("n", [0, 1, 129, 70001])
def test_q6_row_counts_and_negatives(gpu, oracle, n):
    rng = np.random.default_rng(17 + n)
    price = rng.integers(90000, 10495001, n).astype(object)
    for i in range(0, n, 53):
        price[i] = -int(price[i])
    for i in range(7, n, 301):
        price[i] = 5_000_000_000 + i
    disc = rng.integers(4, 9, n)
    dates = rng.integers(8766, 9131, n)  # 1994
    b = _lineitem(n, 18, price=price, disc=disc, dates=dates)
    b6 = pa.record_batch([b.column(0), b.column(1), b.column(2), b.column(6)], names=tpch.Q6_COLUMNS)
    for e in (gpu, oracle):
        e.drop_table("lineitem")
        e.register_batch("lineitem", 0, b6)
    got = driver.run_stages(gpu, tpch.q6(2), f"edge6-{n}")
    want = driver.run_stages(oracle, tpch.q6(2), f"edge6-{n}")
    assert_tables_equal(got, want)
