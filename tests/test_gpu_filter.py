"""FilterExec on the dedicated compaction kernel (csrc/device/filter.cu) vs the CPU oracle: comparisons of integer / date /
decimal / string columns with literals and with each other under AND / OR / NOT / IN, rows kept in input order; shapes the
kernel does not take (LIKE, arithmetic in the predicate, computed outputs, nullable operands) must still match on the VM."""
import pyarrow as pa
import pytest

from ballista_b200 import driver, plan as P, tpch
from test_tpch_queries import load_tables
from util import assert_tables_equal

pytestmark = pytest.mark.gpu
c = P.col
COLS = ["l_orderkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_shipdate", "l_commitdate", "l_receiptdate", "l_shipmode", "l_returnflag"]

PREDS = {
    "date_range": P.and_(P.binop(">=", c("l_shipdate"), P.lit_date("1994-01-01")), P.binop("<", c("l_shipdate"), P.lit_date("1995-01-01"))),
    "col_col": P.and_(P.binop("<", c("l_commitdate"), c("l_receiptdate")), P.binop("<", c("l_shipdate"), c("l_commitdate"))),
    "decimal_and_int": P.and_(P.binop("<", c("l_quantity"), P.lit_dec(2400, 15, 2)), P.binop(">=", c("l_discount"), P.lit_dec(5, 15, 2)),
                              P.binop("<>", c("l_suppkey"), P.lit_i64(7))),
    "in_list_str": P.and_(P.in_list(c("l_shipmode"), [P.lit_utf8("MAIL"), P.lit_utf8("SHIP"), P.lit_utf8("AIR")]), P.binop("=", c("l_returnflag"), P.lit_utf8("R"))),
    "or_not": P.or_(P.not_(P.binop("=", c("l_returnflag"), P.lit_utf8("N"))), P.and_(P.binop(">", c("l_quantity"), P.lit_dec(4900, 15, 2)),
                                                                                    P.binop("<=", c("l_orderkey"), P.lit_i64(1000)))),
    "not_in": P.in_list(c("l_shipmode"), [P.lit_utf8("TRUCK"), P.lit_utf8("RAIL")], negated=True),
    "nothing_passes": P.binop("<", c("l_shipdate"), P.lit_date("1970-01-01")),
}


@pytest.mark.parametrize("name", sorted(PREDS))
def test_fast_filter_against_oracle(gpu, oracle, oracle_lib, name):
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 60, {"lineitem": COLS}, 2)
    scan = tpch.table_scan("lineitem", COLS)
    f = P.filter_(PREDS[name], scan, projection=[0, 3, 8, 5])
    st = [P.Stage(1, P.shuffle_writer(f, 1))]
    n0 = gpu.counter("fastfilter")
    got = driver.run_stages(gpu, st, f"ff-{name}")
    want = driver.run_stages(oracle, st, f"ff-{name}")
    assert gpu.counter("fastfilter") > n0, "the filter did not run on the compaction kernel"
    if want is None or want.num_rows == 0:
        assert got is None or got.num_rows == 0
        return
    assert_tables_equal(got, want, sort=False)   # FilterExec keeps the input order


def test_filtered_hash_shuffle_keeps_partition_contract(gpu, oracle, oracle_lib):
    """filter -> hash repartition: the keys are compacted with the payload and hashed inside the partition kernels."""
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 60, {"lineitem": COLS}, 2)
    f = P.filter_(PREDS["date_range"], tpch.table_scan("lineitem", COLS), projection=[0, 1, 3])
    st = [P.Stage(1, P.shuffle_writer(f, 1, [c(0)], 7)), P.Stage(2, P.shuffle_writer(P.shuffle_reader(1, [P.field("l_orderkey", "i64", True), P.field("l_suppkey", "i64", True),
                                                                                                       P.field("l_extendedprice", P.dec(15, 2), True)]), 2))]
    n0 = gpu.counter("fastfilter")
    got = driver.run_stages(gpu, st, "ff-shuffle")
    want = driver.run_stages(oracle, st, "ff-shuffle")
    assert gpu.counter("fastfilter") > n0
    assert_tables_equal(got, want, sort=False)   # partition by partition, input order inside each


def test_shapes_outside_the_kernel_still_match(gpu, oracle, oracle_lib):
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 30, {"lineitem": COLS}, 1)
    scan = tpch.table_scan("lineitem", COLS)
    for pred in (P.like(c("l_shipmode"), "%AI%"), P.binop(">", P.binop("*", c("l_quantity"), c("l_discount")), P.lit_dec(20000, 31, 4))):
        st = [P.Stage(1, P.shuffle_writer(P.filter_(pred, scan, projection=[0, 2]), 1))]
        got = driver.run_stages(gpu, st, "ff-vm")
        want = driver.run_stages(oracle, st, "ff-vm")
        assert want.num_rows > 0
        assert_tables_equal(got, want, sort=False)


def test_filter_with_fetch(gpu, oracle, oracle_lib):
    """FilterExec { fetch }: the first rows that pass, in input order."""
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 30, {"lineitem": COLS}, 1)
    scan = tpch.table_scan("lineitem", COLS)
    for fetch in (0, 7, 5000, 10**9):
        f = P.filter_(PREDS["date_range"], scan, projection=[0, 3, 8])
        f["fetch"] = fetch
        st = [P.Stage(1, P.shuffle_writer(f, 1))]
        got = driver.run_stages(gpu, st, f"ff-fetch{fetch}")
        want = driver.run_stages(oracle, st, f"ff-fetch{fetch}")
        if want is None or want.num_rows == 0:
            assert got is None or got.num_rows == 0
        else:
            assert_tables_equal(got, want, sort=False)
