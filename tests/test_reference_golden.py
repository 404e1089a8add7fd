"""Golden / known-answer tests lifted from the reference's own test-suite (tests/golden/reference_tests.json
cites file:line), run through BOTH engines:
  * the CPU oracle (not gpu)  -- this is what pins the oracle against the reference's vectors;
  * the CUDA engine through the C-ABI (gpu) -- parity with the oracle and with the golden answers.
Independent cross-checks with pyarrow.compute and sqlite3 are in test_oracle_crosscheck.py.
"""
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import golden_data as G
import queries as Q
from ballista_b200 import driver
from ballista_b200 import plan as P

REF = G.reference_expectations()
ENGINES = [pytest.param("oracle", id="oracle"), pytest.param("gpu", marks=pytest.mark.gpu, id="gpu")]


@pytest.fixture(params=ENGINES)
def engine(request):
    return request.getfixturevalue(request.param)


def alltypes():
    t = G.load("alltypes_plain")
    return t, G.ir_schema("alltypes_plain")


def test_filter_id_gt_4(engine):
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_filter("test", sch, P.binop(">", P.col("id"), P.lit_i32(4))), "g1")
    assert out.num_rows == REF["filter_id_gt_4_rows"]
    assert sorted(out.column("id").to_pylist()) == [5, 6, 7]
    # every column of the surviving rows is carried through unchanged
    want = t.filter(pc.greater(t.column("id"), 4))
    got = out.sort_by("id")
    assert got.equals(want.sort_by("id"))


def test_python_binding_filter(engine):
    t = G.load("python_test")
    G.register(engine, "t", t, 1)
    out = driver.run_stages(engine, Q.q_filter("t", G.ir_schema("python_test"), P.binop(">", P.col("a"), P.lit_i64(2))), "g2")
    assert out.sort_by("a").to_pydict() == REF["python_filter_a_gt_2"]


@pytest.mark.parametrize("n_parts", [1, 2, 5])
def test_groupby_string_count(engine, n_parts):
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_groupby_count("test", sch, "string_col", P.binop(">", P.col("id"), P.lit_i32(4)), n_parts),
                            f"g3-{n_parts}")
    assert [[k, v] for k, v in zip(out.column(0).to_pylist(), out.column(1).to_pylist())] == REF["groupby_string_col_count_where_id_gt_4"]


def test_bool_col_counts(engine):
    t, sch = alltypes()
    G.register(engine, "test", t, 3)
    out = driver.run_stages(engine, Q.q_groupby_count("test", sch, "bool_col", None, 2), "g4")
    assert [[k, v] for k, v in zip(out.column(0).to_pylist(), out.column(1).to_pylist())] == REF["bool_col_counts"]


def test_scalar_aggregates(engine):
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_scalar_aggs("test", sch), "g5")
    row = {n: out.column(n)[0].as_py() for n in out.column_names}
    assert row == {"sum": REF["sum_id"], "avg": REF["avg_id"], "count": REF["count_star"], "min": REF["min_id"], "max": REF["max_id"]}
    assert out.schema.field("sum").type == pa.int64() and out.schema.field("avg").type == pa.float64()
    assert out.schema.field("min").type == pa.int32()


def test_scalar_aggregates_empty_input(engine):
    t, sch = alltypes()
    G.register(engine, "empty", t.slice(0, 0), 1)
    out = driver.run_stages(engine, Q.q_scalar_aggs("empty", sch), "g5e")
    row = {n: out.column(n)[0].as_py() for n in out.column_names}
    assert row == {"sum": None, "avg": None, "count": 0, "min": None, "max": None}


@pytest.mark.parametrize("asc,fetch,want", [(True, None, [0, 1, 2, 3, 4, 5, 6, 7]), (False, None, [7, 6, 5, 4, 3, 2, 1, 0]), (False, 3, [7, 6, 5])])
def test_order_by_limit(engine, asc, fetch, want):
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_order_limit("test", sch, "id", asc, fetch), f"g6-{asc}-{fetch}")
    assert out.column("id").to_pylist() == want
    # the other columns travel with their row
    ref = {r["id"]: r for r in t.to_pylist()}
    for r in out.to_pylist():
        assert r == ref[r["id"]]


def test_hash_join_when_opted_in(engine):
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_self_join("test", sch), "g7")
    assert out.column(0).to_pylist() == REF["hash_join_ids_desc"]


def test_default_sort_merge_join(engine):
    """same query with Ballista's default join strategy (SortMergeJoinExec): same golden result"""
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_self_join("test", sch, smj=True), "g7s")
    assert out.column(0).to_pylist() == REF["hash_join_ids_desc"]


@pytest.mark.parametrize("jt,want", [("LeftSemi", [7, 6, 5, 4, 3]), ("LeftAnti", []), ("RightSemi", [7, 6, 5, 4, 3]),
                                     ("RightAnti", [2, 1, 0]), ("Left", [7, 6, 5, 4, 3]), ("Right", [7, 6, 5, 4, 3, None, None, None]),
                                     ("Full", [7, 6, 5, 4, 3, None, None, None])])
def test_join_types(engine, jt, want):
    t, sch = alltypes()
    G.register(engine, "test", t, 2)
    out = driver.run_stages(engine, Q.q_self_join("test", sch, join_type=jt), "g8" + jt)
    got = out.column(0).to_pylist()
    assert sorted(got, key=lambda v: (v is None, -(v or 0))) == want


def test_config0_min_group_by(engine):
    t = G.load("aggregate_test_100")
    G.register(engine, "aggregate_test_100", t, 2)
    out = driver.run_stages(engine, Q.q_config0("aggregate_test_100", G.ir_schema("aggregate_test_100")), "g9")
    got = {str(k): v for k, v in zip(out.column(0).to_pylist(), out.column(1).to_pylist())}
    assert got == REF["config0_min_b_group_a"]


def test_config0_empty_result(engine):
    t = G.load("python_test")  # a=1..5, b=-2..-6: `a <= b` selects nothing
    G.register(engine, "example", t, 1)
    out = driver.run_stages(engine, Q.q_config0("example", G.ir_schema("python_test"), "a", "b", 2), "g10")
    assert out is None or out.num_rows == 0


def test_remote_sql_example(engine):
    t = G.load("aggregate_test_100")
    G.register(engine, "aggregate_test_100", t, 2)
    out = driver.run_stages(engine, Q.q_remote_sql("aggregate_test_100", G.ir_schema("aggregate_test_100")), "g11").sort_by("c1")
    assert out.column("c1").to_pylist() == REF["remote_sql_groups"]
    c11 = pc.cast(t.column("c11"), pa.float64())
    keep = pc.and_(pc.greater(c11, 0.1), pc.less(c11, 0.9))
    f = t.filter(keep)
    assert f.num_rows == REF["remote_sql_rows_passing"]
    want = f.group_by("c1").aggregate([("c12", "min"), ("c12", "max")]).sort_by("c1")
    assert out.column("min").to_pylist() == want.column("c12_min").to_pylist()  # MIN/MAX of f64 are exact
    assert out.column("max").to_pylist() == want.column("c12_max").to_pylist()


def test_shuffle_writer_unit(engine):
    """shuffle_writer.rs:614-710: 2 input partitions of (a:UInt32 in {1,3}, b:Utf8), P = 2."""
    b = pa.RecordBatch.from_arrays([pa.array([1, 1, 3, 3], type=pa.uint32()), pa.array(["hello", None, "world", "x"])], names=["a", "b"])
    engine.drop_table("mem")
    engine.register_batch("mem", 0, b)
    engine.register_batch("mem", 1, b)
    sch = [P.field("a", "u32"), P.field("b", "utf8", True)]
    st = Q.Stage(1, P.shuffle_writer(P.scan("mem", sch), 1, [P.col(0)], 2, sort_shuffle=False))
    qse = engine.create_query_stage_exec("jobOne", 1, st.json("jobOne"))
    total = 0
    for p in range(2):
        stats = qse.execute_query_stage(p)
        assert sum(s.num_rows for s in stats) == 4  # every input row lands in exactly one output partition
        for s in stats:
            assert s.file_id == p and s.is_sort_shuffle == 0 and s.num_batches == 1
        total += sum(s.num_rows for s in stats)
    assert total == REF["shuffle_writer_unit_total_rows"]
    # equal keys co-locate
    seen = {}
    for p in range(2):
        if engine.partition_rows("jobOne", 1, p) < 0:
            continue
        out = engine.partition_export("jobOne", 1, p)
        for k in set(out.column(0).to_pylist()):
            assert seen.setdefault(k, p) == p
        assert out.schema.names == ["a", "b"]
    assert sorted(seen) == [1, 3]
    engine.remove_job_data("jobOne")
    assert engine.partition_rows("jobOne", 1, 0) == -1


def test_sort_shuffle_all_rows_exactly_once(engine):
    """sort_shuffle/writer.rs:785-829 and :879-1011: every key 0..N-1 appears exactly once after write -> read,
    per-partition stats add up, empty partitions are not reported (:357-369)."""
    n, parts = 1000, 8
    b = pa.RecordBatch.from_arrays([pa.array(range(n), type=pa.int64()), pa.array([f"v{i}" for i in range(n)])], names=["k", "v"])
    engine.drop_table("keys")
    engine.register_batch("keys", 0, b)
    sch = [P.field("k", "i64"), P.field("v", "utf8")]
    st = Q.Stage(1, P.shuffle_writer(P.scan("keys", sch), 1, [P.col(0)], parts))
    qse = engine.create_query_stage_exec("ss", 1, st.json("ss"))
    stats = qse.execute_query_stage(0)
    assert all(s.is_sort_shuffle == 1 and s.num_rows > 0 and s.file_id == 0 for s in stats)
    assert sum(s.num_rows for s in stats) == n
    got = []
    for s in stats:
        out = engine.partition_export("ss", 1, s.partition_id)
        assert out.num_rows == s.num_rows
        ks = out.column("k").to_pylist()
        assert out.column("v").to_pylist() == [f"v{i}" for i in ks]
        # bytes = values + offsets + chars (validity not counted)
        assert s.num_bytes == 8 * len(ks) + 4 * (len(ks) + 1) + sum(len(f"v{i}") for i in ks)
        got += ks
    assert sorted(got) == list(range(n))


def test_sort_shuffle_identical_keys_one_partition(engine):
    """sort_shuffle/writer.rs:1060-1132: 256 identical keys, P = 8 => exactly one non-empty partition."""
    b = pa.RecordBatch.from_arrays([pa.array([42] * 256, type=pa.int64())], names=["k"])
    engine.drop_table("same")
    engine.register_batch("same", 0, b)
    st = Q.Stage(1, P.shuffle_writer(P.scan("same", [P.field("k", "i64")]), 1, [P.col(0)], 8))
    stats = engine.create_query_stage_exec("same", 1, st.json("same")).execute_query_stage(0)
    assert len(stats) == 1 and stats[0].num_rows == 256
