"""PhysicalExpr semantics (ballista/core/proto/datafusion.proto:851-901) on both engines.
`oracle` cases check the CPU oracle against hand-computed answers; `gpu` cases check the CUDA engine
against the oracle on the same inputs (bit-exact, incl. NULL handling and error behaviour)."""
import decimal

import pyarrow as pa
import pytest

import golden_data as G
import queries as Q
from ballista_b200 import driver
from ballista_b200 import plan as P
from util import assert_tables_equal

D = decimal.Decimal
c = P.col


def _table():
    return pa.table({
        "i": pa.array([1, -2, None, 4, 2147483647, 0], type=pa.int32()),
        "l": pa.array([10, None, 30, -40, 9223372036854775807, 7], type=pa.int64()),
        "f": pa.array([1.5, -0.0, None, float("nan"), 1e300, 2.25], type=pa.float64()),
        "d": pa.array([D("1.50"), D("-2.25"), None, D("0.00"), D("9999999999999.99"), D("3.10")], type=pa.decimal128(15, 2)),
        "e": pa.array([D("2.000"), None, D("0.001"), D("-7.125"), D("1.000"), D("3.100")], type=pa.decimal128(12, 3)),
        "s": pa.array(["MED BOX", "Brand#23", None, "", "promo pack", "ASIA"]),
        "b": pa.array([True, False, None, True, False, None]),
        "dt": pa.array([8035, 9298, None, 10591, 0, 10471], type=pa.int32()).cast(pa.date32()),
    })


SCH = [P.field("i", "i32", True), P.field("l", "i64", True), P.field("f", "f64", True), P.field("d", P.dec(15, 2), True),
       P.field("e", P.dec(12, 3), True), P.field("s", "utf8", True), P.field("b", "bool", True), P.field("dt", "date32", True)]

EXPRS = [
    ("i_plus_i", P.binop("+", c("i"), c("i"))),                      # Int32 wrapping
    ("i_times_l", P.binop("*", P.cast(c("i"), "i64"), c("l"))),     # Int64 wrapping
    ("l_minus_7", P.binop("-", c("l"), P.lit_i64(7))),
    ("f_times_2", P.binop("*", c("f"), P.lit_f64(2.0))),
    ("d_plus_e", P.binop("+", c("d"), c("e"))),                      # (15,2)+(12,3) -> (17,3)
    ("d_minus_e", P.binop("-", c("d"), c("e"))),
    ("d_times_e", P.binop("*", c("d"), c("e"))),                     # -> (28,5)
    ("one_minus_d", P.binop("-", P.lit_dec(1, 20, 0), c("d"))),      # (20,0)-(15,2) -> (23,2)
    ("e_times_1md", P.binop("*", c("e"), P.binop("-", P.lit_dec(1, 20, 0), c("d")))),
    ("d_as_f64", P.cast(c("d"), "f64")),
    ("i_as_dec", P.cast(c("i"), P.dec(12, 2))),
    ("d_rescaled", P.cast(c("d"), P.dec(15, 1))),                   # round half away from zero
    ("l_as_i32", P.cast(c("l"), "i32")),                            # overflow -> NULL (safe cast)
    ("cmp_lt", P.binop("<", c("d"), c("e"))),
    ("cmp_ge_lit", P.binop(">=", c("d"), P.lit_dec(150, 15, 2))),
    ("cmp_f", P.binop("<=", c("f"), P.lit_f64(1.5))),
    ("cmp_s", P.binop("=", c("s"), P.lit_utf8("ASIA"))),
    ("cmp_s_lt", P.binop("<", c("s"), P.lit_utf8("N"))),
    ("kleene_and", P.and_(c("b"), P.binop(">", c("i"), P.lit_i32(0)))),
    ("kleene_or", P.or_(c("b"), P.binop(">", c("i"), P.lit_i32(0)))),
    ("not_b", P.not_(c("b"))),
    ("is_null_s", P.is_null(c("s"))),
    ("is_not_null_d", P.is_not_null(c("d"))),
    ("neg_d", P.neg(c("d"))),
    ("case_when", P.case([[P.binop(">", c("i"), P.lit_i32(1)), c("d")], [P.binop("=", c("i"), P.lit_i32(1)), P.lit_dec(-100, 15, 2)]], None)),
    ("case_else", P.case([[c("b"), P.lit_i64(1)]], P.lit_i64(0))),
    ("in_list", P.in_list(c("s"), [P.lit_utf8("ASIA"), P.lit_utf8("MED BOX")])),
    ("not_in_list", P.in_list(c("i"), [P.lit_i32(1), P.lit_i32(4)], negated=True)),
    ("like_pct", P.like(c("s"), "%B%")),
    ("like_under", P.like(c("s"), "AS_A")),
    ("not_like", P.like(c("s"), "promo%", negated=True)),
    ("substr", P.fn("substr", c("s"), P.lit_i64(1), P.lit_i64(3))),
    ("year", P.fn("date_part_year", c("dt"))),
    ("dt_cmp", P.binop("<=", c("dt"), P.lit_date("1998-09-02"))),
]


def _plan(exprs):
    return [Q.Stage(1, P.shuffle_writer(P.project([(e, n) for n, e in exprs], P.scan("x", SCH)), 1))]


def test_oracle_expression_answers(oracle):
    G.register(oracle, "x", _table(), 1)
    out = driver.run_stages(oracle, _plan(EXPRS), "e1")
    g = {n: out.column(n).to_pylist() for n in out.column_names}
    t = {n: out.schema.field(n).type for n in out.column_names}
    assert g["i_plus_i"] == [2, -4, None, 8, -2, 0] and t["i_plus_i"] == pa.int32()
    assert g["i_times_l"][:4] == [10, None, None, -160] and g["i_times_l"][4] == (2147483647 * 9223372036854775807 + 2**63) % 2**64 - 2**63
    assert g["l_minus_7"] == [3, None, 23, -47, 9223372036854775800, 0]
    assert g["f_times_2"][:3] == [3.0, -0.0, None] and g["f_times_2"][5] == 4.5
    assert t["d_plus_e"] == pa.decimal128(17, 3) and g["d_plus_e"] == [D("3.500"), None, None, D("-7.125"), D("10000000000000.990"), D("6.200")]
    assert t["d_times_e"] == pa.decimal128(28, 5) and g["d_times_e"][0] == D("3.00000") and g["d_times_e"][3] == D("0.00000")
    assert t["one_minus_d"] == pa.decimal128(23, 2) and g["one_minus_d"][:2] == [D("-0.50"), D("3.25")]
    assert t["e_times_1md"] == pa.decimal128(36, 5) and g["e_times_1md"][0] == D("-1.00000")
    assert g["d_as_f64"][:2] == [1.5, -2.25]
    assert g["i_as_dec"][:2] == [D("1.00"), D("-2.00")]
    assert g["d_rescaled"] == [D("1.5"), D("-2.3"), None, D("0.0"), D("10000000000000.0"), D("3.1")]
    assert g["l_as_i32"] == [10, None, 30, -40, None, 7]
    assert g["cmp_lt"] == [True, None, None, False, False, False]
    assert g["cmp_ge_lit"] == [True, False, None, False, True, True]
    assert g["cmp_f"] == [True, True, None, False, False, False]   # NaN sorts above everything (total order)
    assert g["cmp_s"] == [False, False, None, False, False, True]
    assert g["kleene_and"] == [True, False, None, True, False, False]
    assert g["kleene_or"] == [True, False, None, True, True, None]
    assert g["not_b"] == [False, True, None, False, True, None]
    assert g["is_null_s"] == [False, False, True, False, False, False]
    assert g["neg_d"][:2] == [D("-1.50"), D("2.25")]
    assert g["case_when"] == [D("-1.00"), None, None, D("0.00"), D("9999999999999.99"), None]
    assert g["case_else"] == [1, 0, 0, 1, 0, 0]
    assert g["in_list"] == [True, False, None, False, False, True]
    assert g["not_in_list"] == [False, True, None, False, True, True]
    assert g["like_pct"] == [True, True, None, False, False, False]
    assert g["like_under"] == [False, False, None, False, False, True]
    assert g["not_like"] == [True, True, None, True, False, True]
    assert g["substr"] == ["MED", "Bra", None, "", "pro", "ASI"]
    assert g["year"] == [1992, 1995, None, 1998, 1970, 1998] and t["year"] == pa.int32()
    assert g["dt_cmp"] == [True, True, None, False, True, True]


@pytest.mark.gpu
def test_gpu_expressions_match_oracle(gpu, oracle):
    for e in (gpu, oracle):
        G.register(e, "x", _table(), 1)
    # one projection per expression keeps each pipeline small; plus everything at once in chunks
    for k in range(0, len(EXPRS), 6):
        chunk = EXPRS[k:k + 6]
        got = driver.run_stages(gpu, _plan(chunk), f"e2-{k}")
        want = driver.run_stages(oracle, _plan(chunk), f"e2-{k}")
        assert_tables_equal(got, want, sort=False)


@pytest.mark.gpu
def test_gpu_filter_with_nulls_and_strings(gpu, oracle):
    preds = [P.binop("=", c("s"), P.lit_utf8("ASIA")), P.and_(c("b"), P.binop(">", c("d"), P.lit_dec(0, 15, 2))),
             P.like(c("s"), "%a%"), P.or_(P.is_null(c("f")), P.binop("<", c("f"), P.lit_f64(2.0))),
             P.in_list(c("i"), [P.lit_i32(1), P.lit_i32(4), P.lit_i32(0)])]
    for e in (gpu, oracle):
        G.register(e, "x", _table(), 2)
    for k, pr in enumerate(preds):
        st = Q.q_filter("x", SCH, pr)
        assert_tables_equal(driver.run_stages(gpu, st, f"e3-{k}"), driver.run_stages(oracle, st, f"e3-{k}"))


ERR_CASES = [
    ("div_by_zero_int", P.binop("/", c("l"), P.lit_i64(0)), -3),
    ("div_by_zero_dec", P.binop("/", c("d"), P.lit_dec(0, 15, 2)), -3),
    ("decimal_overflow", P.binop("*", P.binop("*", P.lit_dec(10**37, 38, 0), c("d")), c("d")), -3),
]


@pytest.mark.parametrize("name,expr,code", ERR_CASES)
def test_oracle_error_behaviour(oracle, oracle_lib, name, expr, code):
    G.register(oracle, "x", _table(), 1)
    with pytest.raises(oracle_lib.OracleError) as ei:
        driver.run_stages(oracle, _plan([(name, expr)]), "err-" + name)
    assert ei.value.code == code


@pytest.mark.gpu
@pytest.mark.parametrize("name,expr,code", ERR_CASES)
def test_gpu_error_behaviour(gpu, name, expr, code):
    import ballista_b200 as bb
    G.register(gpu, "x", _table(), 1)
    with pytest.raises(bb.B200Error) as ei:
        driver.run_stages(gpu, _plan([(name, expr)]), "err-" + name)
    assert ei.value.code == code  # DataFusionError::Execution; the engine stays usable afterwards
    out = driver.run_stages(gpu, _plan(EXPRS[:2]), "after-err")
    assert out.num_rows == 6


def test_plan_errors(oracle, oracle_lib):
    G.register(oracle, "x", _table(), 1)
    bad_root = Q.Stage(1, P.project([(c("i"), "i")], P.scan("x", SCH)))
    with pytest.raises(oracle_lib.OracleError):
        oracle.create_query_stage_exec("j", 1, bad_root.json("j")).execute_query_stage(0)


@pytest.mark.gpu
def test_gpu_plan_errors(gpu):
    import ballista_b200 as bb
    G.register(gpu, "x", _table(), 1)
    bad_root = Q.Stage(1, P.project([(c("i"), "i")], P.scan("x", SCH)))
    with pytest.raises(bb.B200Error) as ei:   # execution_engine.rs:164-167
        gpu.create_query_stage_exec("j", 1, bad_root.json("j"))
    assert ei.value.code == -1 and "ShuffleWriterExec" in str(ei.value)
    with pytest.raises(bb.B200Error) as ei:
        gpu.create_query_stage_exec("j", 1, "{not json")
    assert ei.value.code == -1
    with pytest.raises(bb.B200Error) as ei:   # FetchFailed mapping
        gpu.partition_export("nojob", 1, 0)
    assert ei.value.code == -5
