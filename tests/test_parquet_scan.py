"""Device Parquet scan (DataSourceExec + ParquetSource, datafusion.proto:1058-1077): the host-side Thrift footer / page-header
reader against pyarrow's metadata (CPU), and the decoded columns against pyarrow's reader (GPU): PLAIN and RLE_DICTIONARY
pages, data page V1 and V2, nullable columns with and without NULLs, INT32 / INT64 / DOUBLE / BOOLEAN / strings, decimals as
FIXED_LEN_BYTE_ARRAY and as integers, dates, several row groups, projection; then TPC-H q1 / q6 on a Parquet-scanned lineitem."""
import datetime as dt
import decimal
import os
import random

import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import ballista_b200 as bb
from ballista_b200 import driver, tpch
from util import assert_tables_equal


def _table(n=20000, nulls=True, seed=7):
    rnd = random.Random(seed)

    def maybe(v):
        return None if (nulls and rnd.random() < 0.1) else v
    words = ["", "a", "MAIL", "DELIVER IN PERSON", "1-URGENT", "x" * 70, "Brand#23", "forest green", "é-utf8"]
    cols = {
        "i64": pa.array([maybe(rnd.randrange(-2**62, 2**62)) for _ in range(n)], pa.int64()),
        "i32": pa.array([maybe(rnd.randrange(-2**31, 2**31)) for _ in range(n)], pa.int32()),
        "small": pa.array([rnd.randrange(0, 7) for _ in range(n)], pa.int64()),                      # dictionary friendly, never NULL
        "f64": pa.array([maybe(rnd.random() * 1e6) for _ in range(n)], pa.float64()),
        "flag": pa.array([maybe(rnd.random() < 0.5) for _ in range(n)], pa.bool_()),
        "s": pa.array([maybe(rnd.choice(words)) for _ in range(n)], pa.string()),
        "u": pa.array([maybe("k%d" % rnd.randrange(0, 10**9)) for _ in range(n)], pa.string()),   # high cardinality: falls back to PLAIN
        "d152": pa.array([maybe(decimal.Decimal(rnd.randrange(-10**14, 10**14)).scaleb(-2)) for _ in range(n)], pa.decimal128(15, 2)),
        "d384": pa.array([maybe(decimal.Decimal(rnd.randrange(-10**37, 10**37)).scaleb(-4)) for _ in range(n)], pa.decimal128(38, 4)),
        "day": pa.array([maybe(dt.date(1992, 1, 1) + dt.timedelta(days=rnd.randrange(0, 2500))) for _ in range(n)], pa.date32()),
    }
    return pa.table(cols)


VARIANTS = [
    dict(use_dictionary=True, data_page_version="1.0"),
    dict(use_dictionary=False, data_page_version="1.0"),
    dict(use_dictionary=True, data_page_version="2.0", row_group_size=3000, data_page_size=4096),
    dict(use_dictionary=["small", "s"], data_page_version="2.0", store_decimal_as_integer=True),
]


def _write(tmp_path, t, **kw):
    path = os.path.join(tmp_path, "t.parquet")
    try:
        pq.write_table(t, path, compression="NONE", **kw)
    except TypeError:   # older pyarrow without store_decimal_as_integer
        kw.pop("store_decimal_as_integer", None)
        pq.write_table(t, path, compression="NONE", **kw)
    return path


@pytest.mark.parametrize("kw", VARIANTS)
def test_metadata_reader_matches_pyarrow(tmp_path, kw):
    t = _table(6000)
    path = _write(str(tmp_path), t, **kw)
    d = bb.engine.parquet_describe(path)
    md = pq.ParquetFile(path).metadata
    assert d["num_rows"] == md.num_rows == t.num_rows
    assert d["row_groups"] == md.num_row_groups
    assert [c["name"] for c in d["columns"]] == t.column_names
    for i, c in enumerate(d["columns"]):
        col = md.row_group(0).column(i)
        assert c["values"] == t.num_rows
        assert c["codec"] == 0
        assert c["optional"] is True
        assert (c["dict_pages"] > 0) == col.has_dictionary_page
        phys = {"BOOLEAN": 0, "INT32": 1, "INT64": 2, "DOUBLE": 5, "BYTE_ARRAY": 6, "FIXED_LEN_BYTE_ARRAY": 7}[col.physical_type]
        assert c["physical"] == phys
    by = {c["name"]: c for c in d["columns"]}
    assert (by["d152"]["precision"], by["d152"]["scale"]) == (15, 2)
    assert (by["d384"]["precision"], by["d384"]["scale"]) == (38, 4)


def test_describe_rejects_non_parquet(tmp_path):
    p = os.path.join(str(tmp_path), "x.parquet")
    open(p, "wb").write(b"not a parquet file at all")
    with pytest.raises(bb.B200Error):
        bb.engine.parquet_describe(p)


@pytest.mark.gpu
@pytest.mark.parametrize("nulls", [True, False])
@pytest.mark.parametrize("kw", VARIANTS)
def test_decoded_columns_match_pyarrow(gpu, tmp_path, kw, nulls):
    t = _table(20000, nulls=nulls)
    path = _write(str(tmp_path), t, **kw)
    gpu.drop_table("pqt")
    gpu.register_parquet("pqt", 0, path)
    got = pa.Table.from_batches([gpu.export_table("pqt", 0)])
    want = pq.read_table(path)
    assert_tables_equal(got, want, sort=False)
    # projection push-down: only the named chunks are read, in the order asked for
    gpu.register_parquet("pqt", 1, path, ["s", "d152", "i64"])
    got = pa.Table.from_batches([gpu.export_table("pqt", 1)])
    assert_tables_equal(got, want.select(["s", "d152", "i64"]), sort=False)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", VARIANTS)
def test_snappy_pages(gpu, tmp_path, kw):
    """pyarrow's default codec: the pages are decompressed on the device (one warp per page) before the decode."""
    t = _table(30000, nulls=True, seed=21)
    path = os.path.join(str(tmp_path), "s.parquet")
    kw = dict(kw)
    try:
        pq.write_table(t, path, compression="snappy", **kw)
    except TypeError:
        kw.pop("store_decimal_as_integer", None)
        pq.write_table(t, path, compression="snappy", **kw)
    gpu.drop_table("pqs")
    gpu.register_parquet("pqs", 0, path)
    got = pa.Table.from_batches([gpu.export_table("pqs", 0)])
    assert_tables_equal(got, pq.read_table(path), sort=False)


@pytest.mark.gpu
def test_other_codecs_are_refused(gpu, tmp_path):
    path = os.path.join(str(tmp_path), "z.parquet")
    pq.write_table(_table(100), path, compression="zstd")
    with pytest.raises(bb.B200Error) as ei:
        gpu.register_parquet("pqz", 0, path)
    assert ei.value.code == -2


@pytest.mark.gpu
def test_q1_q6_from_parquet(gpu, oracle, oracle_lib, tmp_path):
    """lineitem written as Parquet by pyarrow from the oracle's generated rows, scanned by the device decoder, then q1 / q6."""
    msf = 20
    cols = list(dict.fromkeys(tpch.Q1_COLUMNS + tpch.Q6_COLUMNS))
    n = oracle_lib.lib().oracle_tpch_table_rows(b"lineitem", msf)
    oracle.drop_table("lineitem")
    oracle.tpch_generate("lineitem", msf, 0, 0, n, cols)
    host = pa.Table.from_batches([oracle.export_table("lineitem", 0)])
    path = os.path.join(str(tmp_path), "lineitem.parquet")
    pq.write_table(host, path, compression="NONE", row_group_size=50000)
    tpch.TABLE_LAYOUT["lineitem"] = cols
    try:
        gpu.drop_table("lineitem")
        gpu.register_parquet("lineitem", 0, path, cols)
        for name, st in (("q1", tpch.q1(4)), ("q6", tpch.q6(4))):
            got = driver.run_stages(gpu, st, f"pq-{name}")
            want = driver.run_stages(oracle, st, f"pq-{name}")
            assert_tables_equal(got, want, sort=False)
    finally:
        tpch.TABLE_LAYOUT.clear()
