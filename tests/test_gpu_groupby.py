"""High-cardinality GROUP BY on the dedicated hash-aggregate kernel (csrc/device/groupby.cu) vs the CPU oracle:
q17 / q18 / q15 / q20-shaped partial aggregates and their FinalPartitioned merges, plus the rows that must leave the
kernel's pattern (key images that do not fit, wide decimals) and still produce the oracle's answer on the general sink."""
import decimal

import pyarrow as pa
import pytest

from ballista_b200 import driver, plan as P, tpch
from test_tpch_queries import load_tables
from util import assert_tables_equal

pytestmark = pytest.mark.gpu
D152 = P.dec(15, 2)
c = P.col


def _two_level(keys, aggs, src, state_fields, n_part=4, filt=None):
    """Partial -> hash shuffle -> FinalPartitioned -> un-partitioned writer."""
    s = src if filt is None else P.filter_(filt, src)
    gb = [(k, nm) for k, nm in keys]
    s1 = P.aggregate("Partial", gb, [P.agg(fn, arg, nm) for fn, arg, nm, _ in aggs], s)
    nk = len(keys)
    st1 = P.Stage(1, P.shuffle_writer(s1, 1, [c(i) for i in range(nk)], n_part))
    fin = P.aggregate("FinalPartitioned", [(c(i), nm) for i, (_, nm) in enumerate(keys)], [P.agg(fn, None, nm, it) for fn, _, nm, it in aggs],
                      P.shuffle_reader(1, state_fields))
    return [st1, P.Stage(2, P.shuffle_writer(fin, 2))]


CASES = {
    # q17 stage 1: GROUP BY l_partkey, AVG(l_quantity)
    "q17": (["l_partkey", "l_quantity"], [(c("l_partkey"), "l_partkey")], [("avg", c("l_quantity"), "a", D152)],
            [P.field("l_partkey", "i64", True), P.field("a[count]", "u64", True), P.field("a[sum]", P.dec(25, 2), True)], None),
    # q18 stage 1: GROUP BY l_orderkey, SUM(l_quantity)
    "q18": (["l_orderkey", "l_quantity"], [(c("l_orderkey"), "l_orderkey")], [("sum", c("l_quantity"), "q", None), ("count", None, "n", None)],
            [P.field("l_orderkey", "i64", True), P.field("q[sum]", P.dec(25, 2), True), P.field("n[count]", "i64", True)], None),
    # q20 stage 4: two integer keys + a date filter
    "q20": (["l_partkey", "l_suppkey", "l_quantity", "l_shipdate"], [(c("l_partkey"), "l_partkey"), (c("l_suppkey"), "l_suppkey")],
            [("sum", c("l_quantity"), "q", None)],
            [P.field("l_partkey", "i64", True), P.field("l_suppkey", "i64", True), P.field("q[sum]", P.dec(25, 2), True)],
            P.binop(">=", c("l_shipdate"), P.lit_date("1994-01-01"))),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("msf,parts", [(30, 2), (300, 1)])
def test_groupby_kernel_against_oracle(gpu, oracle, oracle_lib, name, msf, parts):
    cols, keys, aggs, states, filt = CASES[name]
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, msf, {"lineitem": cols}, parts)
    st = _two_level(keys, aggs, tpch.table_scan("lineitem", cols), states, filt=filt)
    g0 = gpu.counter("groupby")
    got = driver.run_stages(gpu, st, f"gb-{name}-{msf}")
    want = driver.run_stages(oracle, st, f"gb-{name}-{msf}")
    assert gpu.counter("groupby") > g0, "the aggregate did not run on the group-by kernel"
    assert want.num_rows > 1000
    assert_tables_equal(got, want, sort=True)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("bucket_slots", [64, 4096])
def test_groupby_partition_first(gpu, oracle, oracle_lib, name, bucket_slots):
    """The same aggregates with the rows radix-partitioned first and one table region per bucket (what tables larger than
    the L2 get), forced here onto small inputs: many tiny buckets (up to the 4096-bucket limit) and a few larger ones; run
    twice so that the second pass sizes the table from the learnt group count."""
    cols, keys, aggs, states, filt = CASES[name]
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 300, {"lineitem": cols}, 2)
    gpu.set_config("b200.agg.partition_first.bucket_slots", str(bucket_slots))
    gpu.set_config("b200.agg.partition_first.min_rows", "1000")
    try:
        st = _two_level(keys, aggs, tpch.table_scan("lineitem", cols), states, filt=filt)
        want = driver.run_stages(oracle, st, f"gbpf-{name}")
        p0 = gpu.counter("groupby_partition_first")
        for rep in range(2):
            got = driver.run_stages(gpu, st, f"gbpf-{name}-{rep}")
            assert_tables_equal(got, want, sort=True)
        assert gpu.counter("groupby_partition_first") >= p0 + 2
    finally:
        gpu.set_config("b200.agg.partition_first.bucket_slots", str(1 << 19))
        gpu.set_config("b200.agg.partition_first.min_rows", str(1 << 22))


def test_q15_shape_product_sum(gpu, oracle, oracle_lib):
    cols = tpch.Q15_TABLES["lineitem"]
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 100, {"lineitem": cols}, 2)
    st = tpch.q15(4)[:2]   # partial GROUP BY l_suppkey SUM(ext * (1 - disc)) with the shipdate filter + its final merge
    g0 = gpu.counter("groupby")
    got = driver.run_stages(gpu, st, "gb-q15")
    want = driver.run_stages(oracle, st, "gb-q15")
    assert gpu.counter("groupby") > g0
    assert want.num_rows > 100
    assert_tables_equal(got, want, sort=True)


def test_rows_outside_the_pattern_fall_back(gpu, oracle):
    """Two keys where one exceeds 2^32 and a SUM over decimals wider than 64 bits: the kernel bails, the general sink answers."""
    n = 5000
    k0 = pa.array([(i * 7919) % 1000 + (1 << 40) * (i % 3 == 0) for i in range(n)], pa.int64())
    k1 = pa.array([i % 17 for i in range(n)], pa.int64())
    big = decimal.Decimal(10) ** 30
    v = pa.array([big + i for i in range(n)], pa.decimal128(38, 0))
    batch = pa.RecordBatch.from_arrays([k0, k1, v], names=["k0", "k1", "v"])
    for e in (gpu, oracle):
        e.drop_table("t")
        e.register_batch("t", 0, batch)
    sch = [P.field("k0", "i64"), P.field("k1", "i64"), P.field("v", P.dec(38, 0))]
    states = [P.field("k0", "i64", True), P.field("k1", "i64", True), P.field("s[sum]", P.dec(38, 0), True), P.field("n[count]", "i64", True)]
    st = _two_level([(c(0), "k0"), (c(1), "k1")], [("sum", c(2), "s", None), ("count", None, "n", None)], P.scan("t", sch), states)
    got = driver.run_stages(gpu, st, "gb-fallback")
    want = driver.run_stages(oracle, st, "gb-fallback")
    assert want.num_rows > 100
    assert_tables_equal(got, want, sort=True)
