"""Pins the CPU oracle against engines that are independent of it: pyarrow.compute / Acero, the stdlib
sqlite3, and exact Python big-int arithmetic for the Decimal128 rules."""
import decimal
import sqlite3

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import golden_data as G
import queries as Q
from ballista_b200 import driver, tpch
from ballista_b200 import plan as P
from util import assert_tables_equal


def _rand_table(n, seed=3):
    rng = np.random.default_rng(seed)
    return pa.table({
        "k": pa.array(rng.integers(0, 17, n), type=pa.int64()),
        "g": pa.array([["x", "yy", "zzz", None][i] for i in rng.integers(0, 4, n)]),
        "v": pa.array([None if r < 0.1 else int(x) for r, x in zip(rng.random(n), rng.integers(-1000, 1000, n))], type=pa.int64()),
        "f": pa.array(rng.normal(size=n), type=pa.float64()),
    })


SCH = [P.field("k", "i64"), P.field("g", "utf8", True), P.field("v", "i64", True), P.field("f", "f64")]


def _agg_plan(n_parts=3):
    c = P.col
    aggs = [P.agg("sum", c("v"), "s"), P.agg("count", c("v"), "cv"), P.agg("count", None, "c"), P.agg("min", c("v"), "mn"),
            P.agg("max", c("v"), "mx"), P.agg("avg", c("v"), "av"), P.agg("sum", c("f"), "sf")]
    s1 = P.filter_(P.binop("<", c("k"), P.lit_i64(12)), P.scan("t", SCH))
    s1 = P.aggregate("Partial", [(c("k"), "k"), (c("g"), "g")], aggs, s1)
    st1 = Q.Stage(1, P.shuffle_writer(s1, 1, [c(0), c(1)], n_parts))
    partial = [P.field("k", "i64", True), P.field("g", "utf8", True), P.field("s[sum]", "i64", True), P.field("cv[count]", "i64"),
               P.field("c[count]", "i64"), P.field("mn[min]", "i64", True), P.field("mx[max]", "i64", True),
               P.field("av[count]", "u64", True), P.field("av[sum]", "f64", True), P.field("sf[sum]", "f64", True)]
    f = [P.agg("sum", None, "s"), P.agg("count", None, "cv"), P.agg("count", None, "c"), P.agg("min", None, "mn"),
         P.agg("max", None, "mx"), P.agg("avg", None, "av", "i64"), P.agg("sum", None, "sf")]
    s2 = P.aggregate("FinalPartitioned", [(c(0), "k"), (c(1), "g")], f, P.shuffle_reader(1, partial))
    return [st1, Q.Stage(2, P.shuffle_writer(s2, 2))]


def test_groupby_against_sqlite_and_pyarrow(oracle):
    t = _rand_table(4000)
    G.register(oracle, "t", t, 4)
    got = driver.run_stages(oracle, _agg_plan(), "x1")
    # sqlite3
    con = sqlite3.connect(":memory:")
    con.execute("create table t(k integer, g text, v integer, f real)")
    con.executemany("insert into t values (?,?,?,?)", [tuple(r.values()) for r in t.to_pylist()])
    rows = con.execute("select k, g, sum(v), count(v), count(*), min(v), max(v), avg(v), sum(f) from t where k < 12 group by k, g").fetchall()
    want = {(r[0], r[1]): r[2:] for r in rows}
    assert got.num_rows == len(want)
    for r in got.to_pylist():
        w = want[(r["k"], r["g"])]
        assert (r["s"], r["cv"], r["c"], r["mn"], r["mx"]) == w[:5]
        assert (r["av"] is None and w[5] is None) or abs(r["av"] - w[5]) <= 1e-12 * max(1.0, abs(w[5]))
        assert abs(r["sf"] - w[6]) <= 1e-9 * max(1.0, abs(w[6]))
    # pyarrow (Acero hash aggregate)
    pa_out = t.filter(pc.less(t.column("k"), 12)).group_by(["k", "g"]).aggregate([("v", "sum"), ("v", "count"), ("v", "min"), ("v", "max")])
    pw = {(r["k"], r["g"]): r for r in pa_out.to_pylist()}
    for r in got.to_pylist():
        w = pw[(r["k"], r["g"])]
        assert (r["s"], r["cv"], r["mn"], r["mx"]) == (w["v_sum"], w["v_count"], w["v_min"], w["v_max"])


def test_join_against_pyarrow(oracle):
    rng = np.random.default_rng(11)
    l = pa.table({"id": pa.array(rng.integers(0, 50, 300), type=pa.int64()), "a": pa.array(rng.integers(0, 9, 300), type=pa.int64())})
    r = pa.table({"id": pa.array(rng.integers(0, 50, 200), type=pa.int64()), "b": pa.array([f"s{i}" for i in range(200)])})
    G.register(oracle, "l", l, 3)
    G.register(oracle, "r", r, 2)
    c = P.col
    ls, rs = [P.field("id", "i64"), P.field("a", "i64")], [P.field("id", "i64"), P.field("b", "utf8")]
    for jt, pa_jt in (("Inner", "inner"), ("Left", "left outer"), ("Right", "right outer"), ("Full", "full outer"),
                      ("LeftSemi", "left semi"), ("LeftAnti", "left anti"), ("RightSemi", "right semi"), ("RightAnti", "right anti")):
        st1 = Q.Stage(1, P.shuffle_writer(P.scan("l", ls), 1, [c(0)], 4))
        st2 = Q.Stage(2, P.shuffle_writer(P.scan("r", rs), 2, [c(0)], 4))
        j = P.hash_join(P.shuffle_reader(1, ls), P.shuffle_reader(2, rs), [[c(0), c(0)]], jt, "Partitioned")
        st3 = Q.Stage(3, P.shuffle_writer(j, 3))
        got = driver.run_stages(oracle, [st1, st2, st3], "j" + jt)
        want = l.join(r, keys="id", join_type=pa_jt, coalesce_keys=False, right_suffix="_r") if "semi" not in pa_jt and "anti" not in pa_jt \
            else l.join(r, keys="id", join_type=pa_jt)
        assert got.num_rows == want.num_rows, jt
        rows = lambda tb: sorted(tuple(str(v) for v in r) for r in zip(*[col.to_pylist() for col in tb.columns]))
        gl, wl = rows(got), rows(want)
        assert gl == wl, jt


def test_sort_merge_join_ordering(oracle):
    """SortMergeJoinExec: the same rows as the hash join, delivered ordered by the join keys (per task)"""
    rng = np.random.default_rng(12)
    l = pa.table({"id": pa.array(rng.integers(0, 40, 300), type=pa.int64()), "a": pa.array(rng.integers(0, 9, 300), type=pa.int64())})
    r = pa.table({"id": pa.array(rng.integers(0, 40, 200), type=pa.int64()), "b": pa.array([f"s{i}" for i in range(200)])})
    G.register(oracle, "l", l, 3)
    G.register(oracle, "r", r, 2)
    c = P.col
    ls, rs = [P.field("id", "i64"), P.field("a", "i64")], [P.field("id", "i64"), P.field("b", "utf8")]
    rows = lambda tb: sorted(tuple(str(v) for v in rw) for rw in zip(*[col.to_pylist() for col in tb.columns]))
    for jt in ("Inner", "Left", "Right", "LeftSemi", "LeftAnti", "RightSemi"):
        st1 = Q.Stage(1, P.shuffle_writer(P.scan("l", ls), 1, [c(0)], 1))
        st2 = Q.Stage(2, P.shuffle_writer(P.scan("r", rs), 2, [c(0)], 1))
        hj = P.hash_join(P.shuffle_reader(1, ls), P.shuffle_reader(2, rs), [[c(0), c(0)]], jt, "Partitioned")
        sm = P.sort_merge_join(P.shuffle_reader(1, ls), P.shuffle_reader(2, rs), [[c(0), c(0)]], jt,
                               sort_options=[{"asc": False, "nulls_first": True}])
        a = driver.run_stages(oracle, [st1, st2, Q.Stage(3, P.shuffle_writer(hj, 3))], "smj-h" + jt)
        b = driver.run_stages(oracle, [st1, st2, Q.Stage(3, P.shuffle_writer(sm, 3))], "smj-s" + jt)
        assert rows(a) == rows(b), jt
        key_col = 2 if jt == "Right" else 0
        keys = b.column(key_col).to_pylist()
        assert keys == sorted(keys, key=lambda v: (v is not None, -(v if v is not None else 0))), jt  # DESC NULLS FIRST


def test_sort_against_pyarrow(oracle):
    t = _rand_table(500, seed=5)
    G.register(oracle, "t", t, 1)
    c = P.col
    keys = [P.sort_key(c("g"), asc=True, nulls_first=False), P.sort_key(c("v"), asc=False, nulls_first=True), P.sort_key(c("f"))]
    st = [Q.Stage(1, P.shuffle_writer(P.sort(keys, P.scan("t", SCH)), 1))]
    got = driver.run_stages(oracle, st, "s1")
    # pyarrow has one null placement for all keys: emulate with explicit rank columns
    g = t.column("g").to_pylist()
    v = t.column("v").to_pylist()
    f = t.column("f").to_pylist()
    idx = sorted(range(t.num_rows), key=lambda i: ((g[i] is None, g[i] or ""), (v[i] is not None, -(v[i] or 0)), f[i]))
    assert got.column("f").to_pylist() == [f[i] for i in idx]


def test_q1_decimal_rules_against_python_integers(oracle, oracle_lib):
    """Decimal typing [EXT]: (15,2)*(23,2) -> (38,4); *(23,2) -> (38,6); SUM -> (38,s); AVG(15,2) -> (19,6) truncated."""
    msf = 2
    n = oracle_lib.lib().oracle_tpch_table_rows(b"lineitem", msf)
    oracle.tpch_generate("lineitem", msf, 0, 0, n, tpch.Q1_COLUMNS)
    t = pa.Table.from_batches([oracle.export_table("lineitem", 0)])
    got = driver.run_stages(oracle, tpch.q1(3), "d1")
    assert [str(f.type) for f in got.schema] == ["string", "string", "decimal128(25, 2)", "decimal128(25, 2)", "decimal128(38, 4)",
                                                 "decimal128(38, 6)", "decimal128(19, 6)", "decimal128(19, 6)", "decimal128(19, 6)", "int64"]
    un = lambda col: [int(x.as_py().scaleb(2)) for x in t.column(col)]
    qty, ext, disc, tax = un("l_quantity"), un("l_extendedprice"), un("l_discount"), un("l_tax")
    rf, ls = t.column("l_returnflag").to_pylist(), t.column("l_linestatus").to_pylist()
    ship = t.column("l_shipdate").cast(pa.int32()).to_pylist()
    cutoff = 10471  # 1998-09-02
    acc = {}
    for i in range(n):
        if ship[i] > cutoff:
            continue
        a = acc.setdefault((rf[i], ls[i]), [0, 0, 0, 0, 0, 0])
        dp = ext[i] * (100 - disc[i])
        a[0] += qty[i]; a[1] += ext[i]; a[2] += dp; a[3] += dp * (100 + tax[i]); a[4] += disc[i]; a[5] += 1
    D = decimal.Decimal
    for r in got.to_pylist():
        a = acc[(r["l_returnflag"], r["l_linestatus"])]
        assert r["sum_qty"] == D(a[0]).scaleb(-2) and r["sum_base_price"] == D(a[1]).scaleb(-2)
        assert r["sum_disc_price"] == D(a[2]).scaleb(-4) and r["sum_charge"] == D(a[3]).scaleb(-6)
        trunc = lambda s, c: D(int(s * 10**4 / c) if False else (abs(s * 10**4) // c) * (1 if s >= 0 else -1)).scaleb(-6)
        assert r["avg_qty"] == trunc(a[0], a[5]) and r["avg_price"] == trunc(a[1], a[5]) and r["avg_disc"] == trunc(a[4], a[5])
        assert r["count_order"] == a[5]
    assert got.column("l_returnflag").to_pylist() == sorted(got.column("l_returnflag").to_pylist())
