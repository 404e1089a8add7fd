"""GPU parity of the fused kernel's RUN-TIME-described variant (fused.cuh, shape (0,0)): aggregate pipelines that
match the fused pattern but are not one of the ahead-of-time shapes (q1, q6) -- other filter lists, an integer
group key, a two-character string key, plain column sums, a*b products.  CUDA engine vs the CPU oracle, bit-exact;
the kernel-family counters prove which kernel produced the result."""
import decimal

import numpy as np
import pyarrow as pa
import pytest

from ballista_b200 import driver
from ballista_b200 import plan as P
from ballista_b200.plan import Stage
from util import assert_tables_equal

pytestmark = pytest.mark.gpu
D = decimal.Decimal
D122 = P.dec(12, 2)
SCHEMA = [P.field("k", "i32"), P.field("d", D122), P.field("e", D122), P.field("dt", "date32"), P.field("s", "utf8"), P.field("q", "i64")]


def _table(n, seed):
    rng = np.random.default_rng(seed)
    dec = lambda v: pa.array([D(int(x)).scaleb(-2) for x in v], type=pa.decimal128(12, 2))
    return pa.record_batch([
        pa.array(rng.integers(0, 4, n), type=pa.int32()),
        dec(rng.integers(0, 5_000_000, n)),
        dec(rng.integers(0, 100, n)),
        pa.array(rng.integers(9000, 9400, n).astype(np.int32), type=pa.date32()),
        pa.array(list(rng.choice(["ab", "cd", "e", ""], n)), type=pa.utf8()),
        pa.array(rng.integers(-50, 1000, n), type=pa.int64()),
    ], names=[f["name"] for f in SCHEMA])


def _two_phase(child, group_by, aggs, partial_schema, faggs, n_parts=2):
    c = P.col
    s1 = P.aggregate("Partial", group_by, aggs, child)
    keys = [c(i) for i in range(len(group_by))]
    st1 = Stage(1, P.shuffle_writer(s1, 1, keys if keys else None, n_parts if keys else 0))
    mode = "FinalPartitioned" if keys else "Final"
    src = P.shuffle_reader(1, partial_schema)
    s2 = P.aggregate(mode, [(c(i), g[1]) for i, g in enumerate(group_by)], faggs, src if keys else P.coalesce_partitions(src))
    return [st1, Stage(2, P.shuffle_writer(s2, 2), n_tasks=None if keys else 1)]


def _run(gpu, oracle, stages, job, n=150_000):
    b = _table(n, 21)
    for e in (gpu, oracle):
        e.drop_table("ft")
        e.register_batch("ft", 0, b.slice(0, n // 2))
        e.register_batch("ft", 1, b.slice(n // 2))
    f0, s0, v0 = gpu.counter("fused"), gpu.counter("fused_static"), gpu.counter("vm")
    got = driver.run_stages(gpu, stages, job)
    dyn = (gpu.counter("fused") - f0) - (gpu.counter("fused_static") - s0)
    want = driver.run_stages(oracle, stages, job)
    assert_tables_equal(got, want)
    return dyn


def test_scalar_product_sum_other_filters(gpu, oracle):
    """select sum(d*e), count(*) from t where dt >= X and q < 900  (q6-like, different filter list)"""
    c = P.col
    child = P.filter_(P.and_(P.binop(">=", c("dt"), P.lit_date("1994-10-01")), P.binop("<", c("q"), P.lit_i64(900))), P.scan("ft", SCHEMA))
    aggs = [P.agg("sum", P.binop("*", c("d"), c("e")), "rev"), P.agg("count", None, "n")]
    part = [P.field("rev[sum]", P.dec(35, 4), True), P.field("n[count]", "i64")]
    faggs = [P.agg("sum", None, "rev"), P.agg("count", None, "n")]
    dyn = _run(gpu, oracle, _two_phase(child, [], aggs, part, faggs), "fs-a")
    assert dyn >= 2  # both map tasks ran on the run-time-described fused variant


def test_int_key_column_sums(gpu, oracle):
    """select k, sum(d), sum(e), count(*) from t where dt < X group by k  (integer key, no products)"""
    c = P.col
    child = P.filter_(P.binop("<", c("dt"), P.lit_date("1995-06-01")), P.scan("ft", SCHEMA))
    aggs = [P.agg("sum", c("d"), "sd"), P.agg("sum", c("e"), "se"), P.agg("count", None, "n")]
    part = [P.field("k", "i32", True), P.field("sd[sum]", P.dec(22, 2), True), P.field("se[sum]", P.dec(22, 2), True), P.field("n[count]", "i64")]
    faggs = [P.agg("sum", None, "sd"), P.agg("sum", None, "se"), P.agg("count", None, "n")]
    dyn = _run(gpu, oracle, _two_phase(child, [(c("k"), "k")], aggs, part, faggs), "fs-b")
    assert dyn >= 2


def test_short_string_key_with_discounted_product(gpu, oracle):
    """select s, sum(d*(1-e)), count(*) from t group by s  (one short-string key, a*(lit-b) with negative results possible)"""
    c = P.col
    one_minus = P.binop("-", P.lit_dec(1, 20, 0), c("e"))
    aggs = [P.agg("sum", P.binop("*", c("d"), one_minus), "v"), P.agg("count", None, "n")]
    part = [P.field("s", "utf8", True), P.field("v[sum]", P.dec(38, 4), True), P.field("n[count]", "i64")]
    faggs = [P.agg("sum", None, "v"), P.agg("count", None, "n")]
    dyn = _run(gpu, oracle, _two_phase(P.scan("ft", SCHEMA), [(c("s"), "s")], aggs, part, faggs), "fs-c")
    assert dyn >= 2
