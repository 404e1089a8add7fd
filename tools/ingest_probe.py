"""Host->HBM ingest rate of a Decimal128-heavy batch (TPC-H lineitem q1 columns) as a function of the
host-pool size (diagnostic for csrc/host/host_pool.hpp / import_batch)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import pyarrow as pa
import ballista_b200 as bb
from ballista_b200 import tpch

msf = int(os.environ.get("MSF", "10000"))
n = {1000: 5999995, 10000: 59986052}.get(msf) or int(6000000 * msf / 1000)
L = bb.engine.load_library()
src = bb.GpuExecutionEngine(0)
src.tpch_generate("lineitem", msf, 0, 0, n, tpch.Q1_COLUMNS)
host = src.export_table("lineitem", 0)
src.close()
# pinned copies of every buffer (what an executor's scan would hand over)
pinned, arrays = [], []
for col in host.columns:
    bufs = []
    for b in col.buffers():
        if b is None:
            bufs.append(None)
            continue
        p = L.b200_host_alloc_pinned(max(b.size, 64))
        C.memmove(p, b.address, b.size)
        pinned.append(p)
        bufs.append(pa.foreign_buffer(p, b.size))
    arrays.append(pa.Array.from_buffers(col.type, len(col), bufs, null_count=0))
batch = pa.RecordBatch.from_arrays(arrays, schema=host.schema)
total = sum(b.size for a in arrays for b in a.buffers() if b is not None)
chunks = [int(x) for x in os.environ.get("CHUNKS", "4194304").split(",")]
slots = [int(x) for x in os.environ.get("SLOTS", "3").split(",")]
for threads, chunk, nsl in [(t, c, s) for t in [int(x) for x in os.environ.get("THREADS", "0,8,16,32,64,128").split(",")] for c in chunks for s in slots]:
    eng = bb.GpuExecutionEngine(0)
    if threads == 0:
        eng.set_config("b200.ingest.narrow_decimals", "off")
    else:
        eng.set_config("b200.ingest.threads", threads)
        eng.set_config("b200.ingest.chunk_rows", chunk)
        eng.set_config("b200.ingest.slots", nsl)
    ts = []
    for rep in range(4):
        eng.drop_table("t")
        t0 = time.perf_counter()
        eng.register_batch("t", 0, batch)
        ts.append(time.perf_counter() - t0)
    saved = eng.counter("ingest_bytes_saved") // 4
    best = min(ts[1:])
    print(f"threads={threads:3d} chunk={chunk:9d} slots={nsl}  {best*1e3:7.2f} ms  {total/best/1e9:6.1f} GB/s of Arrow bytes  ({(total-saved)/1e9:.2f} GB over PCIe)", flush=True)
    eng.close()
