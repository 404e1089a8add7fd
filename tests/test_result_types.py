"""Result TYPES of the TPC-H q1 / q6 / q17 pipelines, written down from DataFusion 53's documented coercion rules and NOT
derived through csrc/common/plan.hpp (which both the engine and the oracle include for typing):
  Decimal128(p1,s1) * Decimal128(p2,s2) -> Decimal128(min(38, p1+p2+1), s1+s2)      [arrow-rs decimal multiply]
  Int64 literal next to a decimal        -> Decimal128(20,0); 1 - l_discount -> Decimal128(23,2)
  SUM(Decimal128(p,s))                   -> Decimal128(min(38, p+10), s)
  AVG(Decimal128(p,s))                   -> Decimal128(min(38, p+4), s+4)
  COUNT(*)                               -> Int64
so q1 = (15,2)*(23,2) -> (38,4) [39 capped], *(23,2) -> (38,6); SUM(l_quantity) -> (25,2); AVG(l_quantity) -> (19,6)."""
import pyarrow as pa
import pytest

from ballista_b200 import driver, tpch

Q1_TYPES = {
    "l_returnflag": pa.string(), "l_linestatus": pa.string(),
    "sum_qty": pa.decimal128(25, 2), "sum_base_price": pa.decimal128(25, 2),
    "sum_disc_price": pa.decimal128(38, 4), "sum_charge": pa.decimal128(38, 6),
    "avg_qty": pa.decimal128(19, 6), "avg_price": pa.decimal128(19, 6), "avg_disc": pa.decimal128(19, 6),
    "count_order": pa.int64(),
}


def _load(e, lib, cols, msf=5):
    n = lib.lib().oracle_tpch_table_rows(b"lineitem", msf)
    e.drop_table("lineitem")
    e.tpch_generate("lineitem", msf, 0, 0, n, cols)


def _check_q1(t):
    assert [f.name for f in t.schema] == list(Q1_TYPES)
    for f in t.schema:
        assert f.type == Q1_TYPES[f.name], (f.name, f.type)


def test_q1_types_oracle(oracle, oracle_lib):
    _load(oracle, oracle_lib, tpch.Q1_COLUMNS)
    _check_q1(driver.run_stages(oracle, tpch.q1(2), "types"))


def test_q6_type_oracle(oracle, oracle_lib):
    _load(oracle, oracle_lib, tpch.Q6_COLUMNS)
    t = driver.run_stages(oracle, tpch.q6(2), "types6")
    # SUM(l_extendedprice * l_discount): (15,2)*(15,2) -> (31,4); SUM -> (38,4)
    assert t.schema.field(0).type == pa.decimal128(38, 4)


@pytest.mark.gpu
def test_q1_types_gpu(gpu, oracle_lib):
    _load(gpu, oracle_lib, tpch.Q1_COLUMNS)
    _check_q1(driver.run_stages(gpu, tpch.q1(2), "types"))


@pytest.mark.gpu
def test_q6_type_gpu(gpu, oracle_lib):
    _load(gpu, oracle_lib, tpch.Q6_COLUMNS)
    t = driver.run_stages(gpu, tpch.q6(2), "types6")
    assert t.schema.field(0).type == pa.decimal128(38, 4)
