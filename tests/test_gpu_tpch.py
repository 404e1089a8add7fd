"""GPU parity: TPC-H-shaped queries, CUDA engine (through the C-ABI) vs the CPU oracle on the same
synthetic inputs.  Decimals / integers / strings / counts bit-exact."""
import pytest

import ballista_b200 as bb
from ballista_b200 import driver, tpch
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


def _load(engine, oracle_lib, table, msf, columns, n_parts):
    n = oracle_lib.lib().oracle_tpch_table_rows(table.encode(), msf)
    step = (n + n_parts - 1) // n_parts
    for p in range(n_parts):
        engine.tpch_generate(table, msf, p, min(n, p * step), min(n, (p + 1) * step), columns)
    return n


def test_generator_identical(gpu, oracle, oracle_lib):
    cols = tpch.Q1_COLUMNS + ["l_orderkey", "l_partkey", "l_suppkey", "l_linenumber", "l_shipmode", "l_comment"]
    for e in (gpu, oracle):
        e.drop_table("lineitem")
        _load(e, oracle_lib, "lineitem", 10, cols, 3)
    for p in range(3):
        assert_tables_equal(gpu.export_table("lineitem", p), oracle.export_table("lineitem", p), sort=False)


@pytest.mark.parametrize("msf,parts,P", [(10, 2, 4), (50, 3, 16), (1, 1, 1)])
def test_q1(gpu, oracle, oracle_lib, msf, parts, P):
    for e in (gpu, oracle):
        e.drop_table("lineitem")
        _load(e, oracle_lib, "lineitem", msf, tpch.Q1_COLUMNS, parts)
    job = f"q1-{msf}-{parts}-{P}"
    s0 = gpu.counter("fused_static")
    got = driver.run_stages(gpu, tpch.q1(P), job)
    assert gpu.counter("fused_static") >= s0 + parts  # stage 1 ran on the shape-specialised fused kernel
    want = driver.run_stages(oracle, tpch.q1(P), job)
    assert got.num_rows == 4
    assert_tables_equal(got, want, sort=False)  # ORDER BY l_returnflag, l_linestatus


def test_q1_without_prepacked_keys(gpu, oracle, oracle_lib, monkeypatch):
    """The same query with the key columns read as Arrow offsets + characters (the path a table takes whose short-string
    companion images do not exist, e.g. a key column with a string longer than 3 bytes)."""
    monkeypatch.setenv("B200_NO_PREPACK", "1")
    for e in (gpu, oracle):
        e.drop_table("lineitem")
        _load(e, oracle_lib, "lineitem", 20, tpch.Q1_COLUMNS, 2)
    s0 = gpu.counter("fused_static")
    got = driver.run_stages(gpu, tpch.q1(4), "q1-noprepack")
    assert gpu.counter("fused_static") >= s0 + 2
    want = driver.run_stages(oracle, tpch.q1(4), "q1-noprepack")
    assert_tables_equal(got, want, sort=False)


@pytest.mark.parametrize("msf,parts", [(10, 2), (50, 5)])
def test_q6(gpu, oracle, oracle_lib, msf, parts):
    for e in (gpu, oracle):
        e.drop_table("lineitem")
        _load(e, oracle_lib, "lineitem", msf, tpch.Q6_COLUMNS, parts)
    job = f"q6-{msf}-{parts}"
    s0 = gpu.counter("fused_static")
    got = driver.run_stages(gpu, tpch.q6(4), job)
    assert gpu.counter("fused_static") >= s0 + parts
    want = driver.run_stages(oracle, tpch.q6(4), job)
    assert_tables_equal(got, want)
