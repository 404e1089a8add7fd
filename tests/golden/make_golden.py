"""Generates the committed golden fixtures from the reference's own test data.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Outputs (committed; the GPU box never reads /root/reference):
  tests/golden/alltypes_plain.json      <- ballista/client/testdata/alltypes_plain.parquet
  tests/golden/aggregate_test_100.json  <- examples/testdata/aggregate_test_100.csv
  tests/golden/python_test.json         <- python/testdata/test.csv
  tests/golden/reference_tests.json     <- expected tables lifted from the reference's tests (file:line cited)
"""
import json
import os

import pyarrow as pa
import pyarrow.csv as pacsv
import pyarrow.parquet as pq

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def dump_table(t: pa.Table, path: str):
    cols = {}
    for name in t.column_names:
        col = t.column(name)
        typ = col.type
        if pa.types.is_binary(typ):
            vals = [None if v is None else v.decode("latin1") for v in col.to_pylist()]
            typ_s = "utf8"
        elif pa.types.is_timestamp(typ):
            vals = col.cast(pa.int64()).to_pylist()
            typ_s = "ts"
        elif pa.types.is_string(typ):
            vals, typ_s = col.to_pylist(), "utf8"
        elif pa.types.is_boolean(typ):
            vals, typ_s = col.to_pylist(), "bool"
        elif pa.types.is_float32(typ):
            vals, typ_s = col.to_pylist(), "f32"
        elif pa.types.is_float64(typ):
            vals, typ_s = col.to_pylist(), "f64"
        elif pa.types.is_int32(typ):
            vals, typ_s = col.to_pylist(), "i32"
        elif pa.types.is_int64(typ):
            vals, typ_s = col.to_pylist(), "i64"
        elif pa.types.is_uint64(typ):
            vals, typ_s = col.to_pylist(), "u64"
        else:
            vals, typ_s = col.cast(pa.int64()).to_pylist(), "i64"
        cols[name] = {"type": typ_s, "values": vals}
    with open(path, "w") as f:
        json.dump({"num_rows": t.num_rows, "columns": cols}, f, indent=0)


def main():
    dump_table(pq.read_table(f"{REF}/ballista/client/testdata/alltypes_plain.parquet"), f"{OUT}/alltypes_plain.json")
    conv = pacsv.ConvertOptions(column_types={"c1": pa.string(), "c2": pa.int64(), "c3": pa.int64(), "c4": pa.int64(),
                                              "c5": pa.int64(), "c6": pa.int64(), "c7": pa.int64(), "c8": pa.int64(),
                                              "c9": pa.int64(), "c10": pa.uint64(), "c11": pa.float32(),
                                              "c12": pa.float64(), "c13": pa.string()})
    dump_table(pacsv.read_csv(f"{REF}/examples/testdata/aggregate_test_100.csv", convert_options=conv),
               f"{OUT}/aggregate_test_100.json")
    dump_table(pacsv.read_csv(f"{REF}/python/testdata/test.csv"), f"{OUT}/python_test.json")
    ref = {
        # ballista/client/tests/context_checks.rs:58-75  `select ... where id > 4` -> 3 rows
        "filter_id_gt_4_rows": 3,
        # ballista/client/tests/context_checks.rs:813-827 (string_col shown as hex of the binary value)
        "groupby_string_col_count_where_id_gt_4": [["0", 1], ["1", 2]],
        # ballista/client/tests/sort_shuffle.rs:155-175  bool_col counts
        "bool_col_counts": [[False, 4], [True, 4]],
        # ballista/client/tests/sort_shuffle.rs:212-299
        "sum_id": 28, "avg_id": 3.5, "count_star": 8, "min_id": 0, "max_id": 7,
        # ballista/client/tests/context_checks.rs:1015-1066 (self-join t1.id = t2.id where t1.id > 2 -> ids 7,6,5,4,3 desc)
        "hash_join_ids_desc": [7, 6, 5, 4, 3],
        # ballista/core/src/execution_plans/shuffle_writer.rs:614-670: 2 input partitions x 4 rows, keys {1,3}, P = 2:
        # 8 rows in total over the output partitions; equal keys co-locate (absolute placement is hash-dependent: unpinned)
        "shuffle_writer_unit_total_rows": 8,
        # python/python/tests/test_context.py:66-75  filter a > 2
        "python_filter_a_gt_2": {"a": [3, 4, 5], "b": [-4, -5, -6]},
        # config[0] on aggregate_test_100.csv (a:=c2, b:=c3), pyarrow and sqlite3 agree
        "config0_min_b_group_a": {"1": 12, "2": 29, "3": 13, "4": 5, "5": 36},
        # examples/examples/remote-sql.rs:50-56: 5 groups a-e, 86 rows pass the c11 filter
        "remote_sql_groups": ["a", "b", "c", "d", "e"], "remote_sql_rows_passing": 86,
    }
    with open(f"{OUT}/reference_tests.json", "w") as f:
        json.dump(ref, f, indent=1)


if __name__ == "__main__":
    main()
