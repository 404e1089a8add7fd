"""Stage plans of the queries the reference's own tests run (hand-lowered the way Ballista's planner
cuts them), shared by the oracle tests (CPU) and the GPU parity tests."""
from ballista_b200 import plan as P
from ballista_b200.plan import Stage

c = P.col


def q_filter(table, schema, predicate):
    """select * from t where <predicate>      (context_checks.rs:58-75, test_context.py:66-75)"""
    s1 = P.filter_(predicate, P.scan(table, schema))
    return [Stage(1, P.shuffle_writer(s1, 1))]


def q_groupby_count(table, schema, key, predicate, n_parts=2):
    """select key, count(*) from t where pred group by key order by key      (context_checks.rs:813-827)"""
    ki = [f["name"] for f in schema].index(key)
    ktype = schema[ki]["type"]
    s1 = P.scan(table, schema)
    if predicate is not None:
        s1 = P.filter_(predicate, s1)
    s1 = P.aggregate("Partial", [(c(key), key)], [P.agg("count", None, "count(*)")], s1)
    st1 = Stage(1, P.shuffle_writer(s1, 1, [c(0)], n_parts))
    partial = [P.field(key, ktype, True), P.field("count(*)[count]", "i64")]
    s2 = P.aggregate("FinalPartitioned", [(c(0), key)], [P.agg("count", None, "count(*)")], P.shuffle_reader(1, partial))
    s2 = P.sort([P.sort_key(c(0))], s2, preserve_partitioning=True)
    st2 = Stage(2, P.shuffle_writer(s2, 2))
    final = [P.field(key, ktype, True), P.field("count(*)", "i64")]
    s3 = P.sort_preserving_merge([P.sort_key(c(0))], P.shuffle_reader(2, final))
    return [st1, st2, Stage(3, P.shuffle_writer(s3, 3), n_tasks=1)]


def q_scalar_aggs(table, schema, column="id", col_type="i32"):
    """select sum(id), avg(id), count(*), min(id), max(id) from t      (sort_shuffle.rs:212-299)"""
    aggs = [P.agg("sum", c(column), "sum"), P.agg("avg", c(column), "avg"), P.agg("count", None, "count"),
            P.agg("min", c(column), "min"), P.agg("max", c(column), "max")]
    s1 = P.aggregate("Partial", [], aggs, P.scan(table, schema))
    st1 = Stage(1, P.shuffle_writer(s1, 1))
    partial = [P.field("sum[sum]", "i64", True), P.field("avg[count]", "u64", True), P.field("avg[sum]", "f64", True),
               P.field("count[count]", "i64"), P.field("min[min]", col_type, True), P.field("max[max]", col_type, True)]
    faggs = [P.agg("sum", None, "sum"), P.agg("avg", None, "avg", col_type), P.agg("count", None, "count"),
             P.agg("min", None, "min"), P.agg("max", None, "max")]
    s2 = P.aggregate("Final", [], faggs, P.coalesce_partitions(P.shuffle_reader(1, partial)))
    return [st1, Stage(2, P.shuffle_writer(s2, 2), n_tasks=1)]


def q_order_limit(table, schema, key, asc, fetch):
    """select * from t order by key [desc] limit k      (sort_shuffle.rs:516-568)"""
    keys = [P.sort_key(c(key), asc=asc)]
    s1 = P.sort(keys, P.scan(table, schema), fetch=fetch)
    st1 = Stage(1, P.shuffle_writer(s1, 1))
    s2 = P.sort_preserving_merge(keys, P.shuffle_reader(1, [dict(f, nullable=True) for f in schema]), fetch=fetch)
    return [st1, Stage(2, P.shuffle_writer(s2, 2), n_tasks=1)]


def q_self_join(table, schema, key="id", gt=2, n_parts=2, join_type="Inner", smj=False):
    """select t1.id from t t1 join t t2 on t1.id = t2.id where t1.id > 2 order by id desc
    (context_checks.rs:1015-1066, prefer_hash_join=true, PartitionMode::Partitioned)"""
    ki = [f["name"] for f in schema].index(key)
    kt = schema[ki]["type"]
    s1 = P.filter_(P.binop(">", c(key), P.lit_i32(gt)), P.scan(table, schema), projection=[ki])
    st1 = Stage(1, P.shuffle_writer(s1, 1, [c(0)], n_parts))
    s2 = P.project([(c(key), key)], P.scan(table, schema))
    st2 = Stage(2, P.shuffle_writer(s2, 2, [c(0)], n_parts))
    side = [P.field(key, kt, True)]
    if smj:  # the default plan (prefer_hash_join=false, extension.rs:683): SortMergeJoinExec over sorted, co-partitioned inputs
        srt = lambda x: P.sort([P.sort_key(c(0))], x, preserve_partitioning=True)
        j = P.sort_merge_join(srt(P.shuffle_reader(1, side)), srt(P.shuffle_reader(2, side)), [[c(0), c(0)]], join_type)
    else:
        j = P.hash_join(P.shuffle_reader(1, side), P.shuffle_reader(2, side), [[c(0), c(0)]], join_type, "Partitioned")
    out_cols = 1 if join_type in ("LeftSemi", "LeftAnti", "RightSemi", "RightAnti") else 2
    s3 = P.sort([P.sort_key(c(0), asc=False)], P.project([(c(0), "id")], j), preserve_partitioning=True)
    st3 = Stage(3, P.shuffle_writer(s3, 3))
    s4 = P.sort_preserving_merge([P.sort_key(c(0), asc=False)], P.shuffle_reader(3, [P.field("id", kt, True)]))
    return [st1, st2, st3, Stage(4, P.shuffle_writer(s4, 4), n_tasks=1)]


def q_config0(table, schema, a="c2", b="c3", n_parts=2):
    """BASELINE.json configs[0]: SELECT a, MIN(b) FROM t WHERE a <= b GROUP BY a"""
    ai = [f["name"] for f in schema].index(a)
    at = schema[ai]["type"]
    bt = schema[[f["name"] for f in schema].index(b)]["type"]
    s1 = P.filter_(P.binop("<=", c(a), c(b)), P.scan(table, schema))
    s1 = P.aggregate("Partial", [(c(a), a)], [P.agg("min", c(b), "min_b")], s1)
    st1 = Stage(1, P.shuffle_writer(s1, 1, [c(0)], n_parts))
    partial = [P.field(a, at, True), P.field("min_b[min]", bt, True)]
    s2 = P.aggregate("FinalPartitioned", [(c(0), a)], [P.agg("min", None, "min_b")], P.shuffle_reader(1, partial))
    return [st1, Stage(2, P.shuffle_writer(s2, 2))]


def q_remote_sql(table, schema, n_parts=2):
    """examples/examples/remote-sql.rs:50-56:
    SELECT c1, MIN(c12), MAX(c12) FROM t WHERE c11 > 0.1 AND c11 < 0.9 GROUP BY c1"""
    pred = P.and_(P.binop(">", P.cast(c("c11"), "f64"), P.lit_f64(0.1)), P.binop("<", P.cast(c("c11"), "f64"), P.lit_f64(0.9)))
    s1 = P.filter_(pred, P.scan(table, schema))
    s1 = P.aggregate("Partial", [(c("c1"), "c1")], [P.agg("min", c("c12"), "min"), P.agg("max", c("c12"), "max")], s1)
    st1 = Stage(1, P.shuffle_writer(s1, 1, [c(0)], n_parts))
    partial = [P.field("c1", "utf8", True), P.field("min[min]", "f64", True), P.field("max[max]", "f64", True)]
    s2 = P.aggregate("FinalPartitioned", [(c(0), "c1")], [P.agg("min", None, "min"), P.agg("max", None, "max")],
                     P.shuffle_reader(1, partial))
    return [st1, Stage(2, P.shuffle_writer(s2, 2))]
