"""GPU: the exchange layout of a shuffle partition (b200_partition_device_buffers) survives a pack
(b200_device_gather) -> unpack (b200_partition_import_device) round trip bit-exactly, for partitions that are
row slices of a hash-partitioned stage output with string, decimal, integer and nullable columns.  This is the
device side of exchange.exchange_stage without the NCCL hop (bench.py --gpus N covers that)."""
import json

import numpy as np
import pyarrow as pa
import pytest
import torch

import ballista_b200 as bb
from ballista_b200 import plan as P
from ballista_b200.plan import Stage
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


def test_pack_unpack_round_trip(gpu):
    n, parts = 20000, 5
    rng = np.random.default_rng(3)
    words = np.array(["", "a", "bb", "MAIL", "a-rather-longer-string-value", "çé"], dtype=object)
    ints = rng.integers(-1000, 1000, n)
    mask = rng.random(n) < 0.1
    batch = pa.record_batch([
        pa.array(rng.integers(0, 97, n), type=pa.int64()),
        pa.array(list(words[rng.integers(0, len(words), n)]), type=pa.utf8()),
        pa.array([None if m else int(v) for v, m in zip(ints, mask)], type=pa.int32()),
        pa.array([__import__("decimal").Decimal(int(v)).scaleb(-2) for v in rng.integers(-10**9, 10**9, n)], type=pa.decimal128(15, 2)),
        pa.array(list(words[rng.integers(0, len(words), n)]), type=pa.utf8()),
    ], names=["k", "s", "i", "d", "t"])
    schema = [P.field("k", "i64"), P.field("s", "utf8"), P.field("i", "i32", True), P.field("d", P.dec(15, 2)), P.field("t", "utf8")]
    gpu.drop_table("xt")
    gpu.register_batch("xt", 0, batch)
    job = "xbuf"
    st = Stage(1, P.shuffle_writer(P.scan("xt", schema), 1, [P.col(0)], parts))
    q = gpu.create_query_stage_exec(job, 1, st.json(job))
    q.execute_query_stage(0)
    q.release()
    want = {p: gpu.partition_export(job, 1, p) for p in range(parts) if gpu.partition_rows(job, 1, p) >= 0}
    assert sum(b.num_rows for b in want.values()) == n
    dev = torch.device("cuda", 0)
    for p, w in want.items():
        bufs, rows = gpu.partition_device_buffers(job, 1, p)
        assert rows == w.num_rows and len(bufs) == 3 * len(schema)
        total = sum(b for _, b in bufs)
        msg = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
        gpu.device_gather(bufs, msg.data_ptr(), total)
        gpu.synchronize()
        pos, moved = 0, []
        for _, nb in bufs:
            moved.append((msg.data_ptr() + pos if nb else 0, nb))
            pos += nb
        gpu.partition_import_device(job, 7, p, 0, json.dumps(schema), moved, rows)
        got = gpu.partition_export(job, 7, p)
        assert_tables_equal(pa.Table.from_batches([got]), pa.Table.from_batches([w]), sort=False)
    gpu.remove_job_data(job)
