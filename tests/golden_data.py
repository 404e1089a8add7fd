"""Loads the committed golden fixtures (tests/golden/*.json) as pyarrow tables."""
import json
import os

import pyarrow as pa

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_TYPES = {"utf8": pa.string(), "ts": pa.timestamp("ns"), "bool": pa.bool_(), "f32": pa.float32(), "f64": pa.float64(),
          "i32": pa.int32(), "i64": pa.int64(), "u64": pa.uint64()}
IR_TYPES = {"utf8": "utf8", "ts": "ts", "bool": "bool", "f32": "f32", "f64": "f64", "i32": "i32", "i64": "i64", "u64": "u64"}


def load(name: str) -> pa.Table:
    with open(os.path.join(_DIR, name + ".json")) as f:
        d = json.load(f)
    arrays, names = [], []
    for n, c in d["columns"].items():
        arrays.append(pa.array(c["values"], type=pa.int64()).cast(_TYPES[c["type"]]) if c["type"] == "ts"
                      else pa.array(c["values"], type=_TYPES[c["type"]]))
        names.append(n)
    return pa.Table.from_arrays(arrays, names=names)


def ir_schema(name: str):
    with open(os.path.join(_DIR, name + ".json")) as f:
        d = json.load(f)
    return [{"name": n, "type": IR_TYPES[c["type"]], "nullable": False} for n, c in d["columns"].items()]


def reference_expectations():
    with open(os.path.join(_DIR, "reference_tests.json")) as f:
        return json.load(f)


def register(engine, table_name: str, table: pa.Table, n_partitions: int = 2):
    """Split `table` row-wise into partitions and register them on `engine`."""
    engine.drop_table(table_name)
    n = table.num_rows
    step = max(1, (n + n_partitions - 1) // n_partitions)
    for p in range(n_partitions):
        sl = table.slice(min(n, p * step), max(0, min(n, (p + 1) * step) - min(n, p * step)))
        engine.register_batch(table_name, p, sl.combine_chunks().to_batches()[0] if sl.num_rows else
                              pa.RecordBatch.from_arrays([pa.array([], type=f.type) for f in table.schema], schema=table.schema))
