"""TPC-H schemas (benchmarks/src/bin/tpch.rs:960-1049) and hand-lowered distributed stage plans for
the benchmark queries (benchmarks/queries/q*.sql), in the shape Ballista's planner produces
(ballista/scheduler/src/planner.rs:126-263; example stage shapes :655-670): every stage is rooted
at a shuffle writer, cut at hash repartitions / SortPreservingMerge.

Joins are lowered as HashJoinExec (``datafusion.optimizer.prefer_hash_join=true``, the opt-in
setting the north_star names; Ballista's default is sort-merge join, extension.rs:683).
"""
from __future__ import annotations

from typing import Dict, List

from . import plan as P
from .plan import Stage

D152 = P.dec(15, 2)

SCHEMAS: Dict[str, List[dict]] = {
    "part": [P.field("p_partkey", "i64"), P.field("p_name", "utf8"), P.field("p_mfgr", "utf8"),
             P.field("p_brand", "utf8"), P.field("p_type", "utf8"), P.field("p_size", "i32"),
             P.field("p_container", "utf8"), P.field("p_retailprice", D152), P.field("p_comment", "utf8")],
    "supplier": [P.field("s_suppkey", "i64"), P.field("s_name", "utf8"), P.field("s_address", "utf8"),
                 P.field("s_nationkey", "i64"), P.field("s_phone", "utf8"), P.field("s_acctbal", D152),
                 P.field("s_comment", "utf8")],
    "partsupp": [P.field("ps_partkey", "i64"), P.field("ps_suppkey", "i64"), P.field("ps_availqty", "i32"),
                 P.field("ps_supplycost", D152), P.field("ps_comment", "utf8")],
    "customer": [P.field("c_custkey", "i64"), P.field("c_name", "utf8"), P.field("c_address", "utf8"),
                 P.field("c_nationkey", "i64"), P.field("c_phone", "utf8"), P.field("c_acctbal", D152),
                 P.field("c_mktsegment", "utf8"), P.field("c_comment", "utf8")],
    "orders": [P.field("o_orderkey", "i64"), P.field("o_custkey", "i64"), P.field("o_orderstatus", "utf8"),
               P.field("o_totalprice", D152), P.field("o_orderdate", "date32"), P.field("o_orderpriority", "utf8"),
               P.field("o_clerk", "utf8"), P.field("o_shippriority", "i32"), P.field("o_comment", "utf8")],
    "lineitem": [P.field("l_orderkey", "i64"), P.field("l_partkey", "i64"), P.field("l_suppkey", "i64"),
                 P.field("l_linenumber", "i32"), P.field("l_quantity", D152), P.field("l_extendedprice", D152),
                 P.field("l_discount", D152), P.field("l_tax", D152), P.field("l_returnflag", "utf8"),
                 P.field("l_linestatus", "utf8"), P.field("l_shipdate", "date32"), P.field("l_commitdate", "date32"),
                 P.field("l_receiptdate", "date32"), P.field("l_shipinstruct", "utf8"), P.field("l_shipmode", "utf8"),
                 P.field("l_comment", "utf8")],
    "nation": [P.field("n_nationkey", "i64"), P.field("n_name", "utf8"), P.field("n_regionkey", "i64"),
               P.field("n_comment", "utf8")],
    "region": [P.field("r_regionkey", "i64"), P.field("r_name", "utf8"), P.field("r_comment", "utf8")],
}


def col_index(table: str, name: str) -> int:
    for i, f in enumerate(SCHEMAS[table]):
        if f["name"] == name:
            return i
    raise KeyError(name)


def table_scan(table: str, columns: List[str]) -> dict:
    """DataSourceExec with projection push-down; the table must be registered with exactly `columns`
    (the harness generates only the referenced columns), so the projection is the identity."""
    sch = [f for name in columns for f in SCHEMAS[table] if f["name"] == name]
    return P.scan(table, sch)


def one_minus(x):  # `1 - x` with the Int64 literal already coerced to Decimal128(20,0) [EXT]
    return P.binop("-", P.lit_dec(1, 20, 0), x)


def one_plus(x):
    return P.binop("+", P.lit_dec(1, 20, 0), x)


Q1_COLUMNS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]


def q1(n_partitions: int = 16) -> List[Stage]:
    """benchmarks/queries/q1.sql -- scan + filter + projection + hash aggregate (low cardinality)."""
    c = P.col
    s1 = table_scan("lineitem", Q1_COLUMNS)
    s1 = P.filter_(P.binop("<=", c("l_shipdate"), P.lit_date("1998-09-02")), s1, projection=[0, 1, 2, 3, 4, 5])
    # DataFusion common-subexpression elimination: disc_price is computed once
    s1 = P.project([(P.binop("*", c("l_extendedprice"), one_minus(c("l_discount"))), "__common_expr_1"),
                    (c("l_quantity"), "l_quantity"), (c("l_extendedprice"), "l_extendedprice"),
                    (c("l_discount"), "l_discount"), (c("l_tax"), "l_tax"),
                    (c("l_returnflag"), "l_returnflag"), (c("l_linestatus"), "l_linestatus")], s1)
    aggs = [P.agg("sum", c("l_quantity"), "sum_qty"),
            P.agg("sum", c("l_extendedprice"), "sum_base_price"),
            P.agg("sum", c("__common_expr_1"), "sum_disc_price"),
            P.agg("sum", P.binop("*", c("__common_expr_1"), one_plus(c("l_tax"))), "sum_charge"),
            P.agg("avg", c("l_quantity"), "avg_qty"),
            P.agg("avg", c("l_extendedprice"), "avg_price"),
            P.agg("avg", c("l_discount"), "avg_disc"),
            P.agg("count", None, "count_order")]
    s1 = P.aggregate("Partial", [(c("l_returnflag"), "l_returnflag"), (c("l_linestatus"), "l_linestatus")], aggs, s1)
    st1 = Stage(1, P.shuffle_writer(s1, 1, [c(0), c(1)], n_partitions))

    partial_schema = [P.field("l_returnflag", "utf8"), P.field("l_linestatus", "utf8"),
                      P.field("sum_qty[sum]", P.dec(25, 2), True), P.field("sum_base_price[sum]", P.dec(25, 2), True),
                      P.field("sum_disc_price[sum]", P.dec(38, 4), True), P.field("sum_charge[sum]", P.dec(38, 6), True),
                      P.field("avg_qty[count]", "u64", True), P.field("avg_qty[sum]", P.dec(25, 2), True),
                      P.field("avg_price[count]", "u64", True), P.field("avg_price[sum]", P.dec(25, 2), True),
                      P.field("avg_disc[count]", "u64", True), P.field("avg_disc[sum]", P.dec(25, 2), True),
                      P.field("count_order[count]", "i64")]
    faggs = [P.agg("sum", None, "sum_qty"), P.agg("sum", None, "sum_base_price"),
             P.agg("sum", None, "sum_disc_price"), P.agg("sum", None, "sum_charge"),
             P.agg("avg", None, "avg_qty", D152), P.agg("avg", None, "avg_price", D152),
             P.agg("avg", None, "avg_disc", D152), P.agg("count", None, "count_order")]
    s2 = P.aggregate("FinalPartitioned", [(c(0), "l_returnflag"), (c(1), "l_linestatus")], faggs,
                     P.shuffle_reader(1, partial_schema))
    keys = [P.sort_key(c(0)), P.sort_key(c(1))]
    s2 = P.sort(keys, s2, preserve_partitioning=True)
    st2 = Stage(2, P.shuffle_writer(s2, 2))

    final_schema = [P.field("l_returnflag", "utf8"), P.field("l_linestatus", "utf8"),
                    P.field("sum_qty", P.dec(25, 2), True), P.field("sum_base_price", P.dec(25, 2), True),
                    P.field("sum_disc_price", P.dec(38, 4), True), P.field("sum_charge", P.dec(38, 6), True),
                    P.field("avg_qty", P.dec(19, 6), True), P.field("avg_price", P.dec(19, 6), True),
                    P.field("avg_disc", P.dec(19, 6), True), P.field("count_order", "i64")]
    s3 = P.sort_preserving_merge(keys, P.shuffle_reader(2, final_schema))
    st3 = Stage(3, P.shuffle_writer(s3, 3), n_tasks=1)
    return [st1, st2, st3]


Q6_COLUMNS = ["l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]


def q6(n_partitions: int = 16) -> List[Stage]:
    """benchmarks/queries/q6.sql -- multi-conjunct range filter + scalar aggregate."""
    c = P.col
    s1 = table_scan("lineitem", Q6_COLUMNS)
    pred = P.and_(P.binop(">=", c("l_shipdate"), P.lit_date("1994-01-01")),
                  P.binop("<", c("l_shipdate"), P.lit_date("1995-01-01")),
                  P.binop(">=", c("l_discount"), P.lit_dec(5, 15, 2)),
                  P.binop("<=", c("l_discount"), P.lit_dec(7, 15, 2)),
                  P.binop("<", c("l_quantity"), P.lit_dec(2400, 15, 2)))
    s1 = P.filter_(pred, s1, projection=[1, 2])
    s1 = P.aggregate("Partial", [], [P.agg("sum", P.binop("*", c("l_extendedprice"), c("l_discount")), "revenue")], s1)
    st1 = Stage(1, P.shuffle_writer(s1, 1))
    partial_schema = [P.field("revenue[sum]", P.dec(38, 4), True)]
    s2 = P.aggregate("Final", [], [P.agg("sum", None, "revenue")],
                     P.coalesce_partitions(P.shuffle_reader(1, partial_schema)))
    st2 = Stage(2, P.shuffle_writer(s2, 2), n_tasks=1)
    return [st1, st2]
