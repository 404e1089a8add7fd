"""The reference's shuffle file format -- Arrow IPC streams with LZ4_FRAME body compression (shuffle_writer.rs:317-328,
sort_shuffle/writer.rs:419-513, index.rs:18-33) -- restated in csrc/host/arrow_ipc.hpp, pinned against pyarrow's own reader and
writer (Arrow C++, the same format arrow-rs implements):
  * what b200_ipc_encode writes, pyarrow.ipc.open_stream reads back identically (compressed and not, chunked);
  * what pyarrow writes with compression='lz4' (real LZ4 frames with matches), b200_ipc_decode reads back identically;
  * concatenated streams with a leading schema-only stream decode as one batch (the sort-shuffle partition range);
  * GPU: shuffle_write_files produces the reference's hash and sort layouts (files + index) that pyarrow reads, and
    shuffle_read_file ingests files pyarrow wrote in those layouts."""
import datetime as dt
import decimal
import os
import random
import struct

import pyarrow as pa
import pyarrow.ipc
import pytest

import ballista_b200 as bb
from ballista_b200 import driver, plan as P, tpch
from ballista_b200.engine import ipc_decode, ipc_encode
from util import assert_tables_equal


def _batch(n=5000, nulls=True, seed=3):
    rnd = random.Random(seed)

    def maybe(v):
        return None if (nulls and rnd.random() < 0.15) else v
    words = ["", "a", "MAIL", "DELIVER IN PERSON", "x" * 300, "é-utf8", "Customer#000000001"]
    return pa.record_batch({
        "i64": pa.array([maybe(rnd.randrange(-2**62, 2**62)) for _ in range(n)], pa.int64()),
        "i32": pa.array([maybe(rnd.randrange(-2**31, 2**31)) for _ in range(n)], pa.int32()),
        "u8": pa.array([rnd.randrange(0, 256) for _ in range(n)], pa.uint8()),
        "rep": pa.array([7] * n, pa.int64()),                                     # compresses to almost nothing
        "f64": pa.array([maybe(rnd.random()) for _ in range(n)], pa.float64()),
        "flag": pa.array([maybe(rnd.random() < 0.5) for _ in range(n)], pa.bool_()),
        "s": pa.array([maybe(rnd.choice(words)) for _ in range(n)], pa.string()),
        "d": pa.array([maybe(decimal.Decimal(rnd.randrange(-10**20, 10**20)).scaleb(-4)) for _ in range(n)], pa.decimal128(38, 4)),
        "day": pa.array([maybe(dt.date(1992, 1, 1) + dt.timedelta(days=rnd.randrange(0, 2500))) for _ in range(n)], pa.date32()),
    })


def _as_table(b):
    return pa.Table.from_batches([b])


@pytest.mark.parametrize("compress", [False, True])
@pytest.mark.parametrize("n,chunk", [(5000, 0), (5000, 1024), (1, 0), (0, 0)])
def test_pyarrow_reads_what_we_write(compress, n, chunk):
    b = _batch(n) if n else _batch(5).slice(0, 0)
    data = ipc_encode(b, compress=compress, max_rows_per_message=chunk)
    rd = pa.ipc.open_stream(data)
    got = rd.read_all()
    assert rd.schema.names == b.schema.names
    assert [f.type for f in rd.schema] == [f.type for f in b.schema]
    assert_tables_equal(got, _as_table(b), sort=False)
    if chunk and n > chunk:
        assert len(pa.ipc.open_stream(data).read_all().to_batches()) == (n + chunk - 1) // chunk


@pytest.mark.parametrize("codec", [None, "lz4"])
@pytest.mark.parametrize("nulls", [True, False])
def test_we_read_what_pyarrow_writes(codec, nulls):
    b = _batch(8000, nulls=nulls, seed=11)
    sink = pa.BufferOutputStream()
    opts = pa.ipc.IpcWriteOptions(compression=codec)
    with pa.ipc.new_stream(sink, b.schema, options=opts) as w:
        for r0 in range(0, b.num_rows, 3000):
            w.write_batch(b.slice(r0, 3000))
    got = ipc_decode(sink.getvalue().to_pybytes())
    assert_tables_equal(_as_table(got), _as_table(b), sort=False)


def test_round_trip_through_both_codecs():
    b = _batch(3000, seed=5)
    assert_tables_equal(_as_table(ipc_decode(ipc_encode(b, compress=True, max_rows_per_message=700))), _as_table(b), sort=False)


def test_concatenated_streams_with_schema_header():
    """The byte range of one partition in a sort-shuffle data file: [schema-only stream] + several complete streams."""
    b = _batch(4000, seed=9)
    opts = pa.ipc.IpcWriteOptions(compression="lz4")
    parts = []
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, b.schema, options=opts):
        pass
    parts.append(sink.getvalue().to_pybytes())
    for r0 in (0, 1500, 2500):
        sink = pa.BufferOutputStream()
        with pa.ipc.new_stream(sink, b.schema, options=opts) as w:
            w.write_batch(b.slice(r0, 1500 if r0 == 0 else (1000 if r0 == 1500 else 1500)))
        parts.append(sink.getvalue().to_pybytes())
    got = ipc_decode(b"".join(parts))
    assert_tables_equal(_as_table(got), _as_table(b), sort=False)


def test_garbage_is_rejected():
    with pytest.raises(bb.B200Error):
        ipc_decode(b"\xff\xff\xff\xff\x10\x00\x00\x00" + b"\x01" * 16)


# ---- GPU: the engine's shuffle store <-> files in the reference's layouts -----------------------------------------------
def _read_sort_partition(data_path, p):
    idx = open(data_path + ".index", "rb").read()
    offs = struct.unpack("<%dq" % (len(idx) // 8), idx)
    raw = open(data_path, "rb").read()
    schema = pa.ipc.open_stream(raw[:offs[0]]).schema
    out, pos, seg = [], 0, raw[offs[p]:offs[p + 1]]
    while pos < len(seg):       # concatenated complete streams
        rd = pa.ipc.open_stream(seg[pos:])
        out += rd.read_all().to_batches()
        # a stream ends with the 8-byte end-of-stream marker; find it by re-encoding length is not possible: scan messages
        pos += _stream_length(seg[pos:])
    return schema, out


def _stream_length(buf):
    pos = 0
    while True:
        cont, msize = struct.unpack_from("<Ii", buf, pos)
        assert cont == 0xFFFFFFFF
        pos += 8
        if msize == 0:
            return pos
        meta = buf[pos:pos + msize]
        root = struct.unpack_from("<I", meta, 0)[0]
        so = struct.unpack_from("<i", meta, root)[0]
        vt = root - so
        vsize = struct.unpack_from("<H", meta, vt)[0]
        body = 0
        if 4 + 2 * 3 + 2 <= vsize:
            off = struct.unpack_from("<H", meta, vt + 4 + 2 * 3)[0]
            if off:
                body = struct.unpack_from("<q", meta, root + off)[0]
        pos += msize + body


@pytest.mark.gpu
@pytest.mark.parametrize("sort_layout", [False, True])
def test_shuffle_files_in_reference_layout(gpu, oracle, oracle_lib, tmp_path, sort_layout):
    from test_tpch_queries import load_tables
    cols = ["l_orderkey", "l_quantity", "l_shipmode", "l_shipdate"]
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 20, {"lineitem": cols}, 2)
    Pn = 5
    st = P.Stage(1, P.shuffle_writer(tpch.table_scan("lineitem", cols), 1, [P.col(0)], Pn, sort_shuffle=sort_layout))
    job = f"files-{int(sort_layout)}"
    for e in (gpu, oracle):
        q = e.create_query_stage_exec(job, 1, st.json(job))
        for p in range(2):
            q.execute_query_stage(p)
        q.release()
    wd = str(tmp_path)
    res = gpu.shuffle_write_files(job, 1, wd, Pn, sort_layout)
    assert res["files"] == (4 if sort_layout else 2 * Pn)
    for p in range(Pn):
        want = pa.Table.from_batches([oracle.partition_export(job, 1, p)])
        if sort_layout:
            got_batches = []
            for task in range(2):
                schema, bs = _read_sort_partition(os.path.join(wd, job, "1", str(task), "data.arrow"), p)
                got_batches += bs
            got = pa.Table.from_batches(got_batches, schema=schema)
        else:
            got = pa.concat_tables([pa.ipc.open_stream(open(os.path.join(wd, job, "1", str(p), f"data-{task}.arrow"), "rb").read()).read_all() for task in range(2)])
        assert_tables_equal(got, want, sort=False)
    # ... and back: a fresh job id reads the files (as a CPU executor's output would be read) and serves the same partitions
    job2 = job + "-reload"
    for p in range(Pn):
        for task in range(2):
            if sort_layout:
                gpu.shuffle_read_file(job2, 1, p, task, os.path.join(wd, job, "1", str(task), "data.arrow"), use_index=True)
            else:
                gpu.shuffle_read_file(job2, 1, p, task, os.path.join(wd, job, "1", str(p), f"data-{task}.arrow"))
        got = pa.Table.from_batches([gpu.partition_export(job2, 1, p)])
        want = pa.Table.from_batches([oracle.partition_export(job, 1, p)])
        assert_tables_equal(got, want, sort=False)
