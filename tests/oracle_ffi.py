"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import pyarrow as pa

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "oracle", "liboracle.so")


class ShuffleWritePartition(C.Structure):
    _fields_ = [("partition_id", C.c_uint64), ("num_batches", C.c_uint64), ("num_rows", C.c_uint64),
                ("num_bytes", C.c_uint64), ("file_id", C.c_int64), ("is_sort_shuffle", C.c_int32),
                ("_pad", C.c_int32)]

    def as_tuple(self):
        return (self.partition_id, self.num_batches, self.num_rows, self.num_bytes, self.file_id, self.is_sort_shuffle)


class ArrowSchema(C.Structure):
    _fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                ("n_children", C.c_int64), ("children", C.c_void_p), ("dictionary", C.c_void_p),
                ("release", C.c_void_p), ("private_data", C.c_void_p)]


class ArrowArray(C.Structure):
    _fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                ("n_children", C.c_int64), ("buffers", C.c_void_p), ("children", C.c_void_p),
                ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_set_config.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.oracle_register_batch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_drop_table.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_tpch_table_rows.argtypes = [C.c_char_p, C.c_int64]
        L.oracle_tpch_table_rows.restype = C.c_int64
        L.oracle_tpch_generate.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_char_p]
        L.oracle_export_table.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_execute_stage.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int,
                                           C.POINTER(ShuffleWritePartition), C.c_int, C.POINTER(C.c_int)]
        L.oracle_partition_export.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p]
        L.oracle_partition_rows.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int]
        L.oracle_partition_rows.restype = C.c_int64
        L.oracle_remove_job_data.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_hash_partition_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int64,
                                                C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().oracle_last_error().decode())


def export_batch(batch: pa.RecordBatch):
    arr, sch = ArrowArray(), ArrowSchema()
    batch._export_to_c(C.addressof(arr), C.addressof(sch))
    return arr, sch


def import_batch(arr: ArrowArray, sch: ArrowSchema) -> pa.RecordBatch:
    return pa.RecordBatch._import_from_c(C.addressof(arr), C.addressof(sch))


class _StageExec:
    def __init__(self, eng, job_id, stage_id, plan_json):
        self.eng, self.job_id, self.stage_id, self.plan_json = eng, job_id, stage_id, plan_json

    def execute_query_stage(self, input_partition: int) -> List[ShuffleWritePartition]:
        cap = 65536
        out = (ShuffleWritePartition * cap)()
        n = C.c_int(0)
        _check(lib().oracle_execute_stage(self.eng.h, self.job_id.encode(), self.stage_id, self.plan_json.encode(),
                                          input_partition, out, cap, C.byref(n)))
        return [out[i] for i in range(n.value)]

    def collect_plan_metrics(self):
        return []

    def release(self):
        pass


class OracleEngine:
    """Same Python surface as ballista_b200.GpuExecutionEngine, backed by the CPU oracle."""

    def __init__(self):
        self.h = C.c_void_p(lib().oracle_create())
        self._parts = {}

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def set_config(self, key, value):
        lib().oracle_set_config(self.h, key.encode(), str(value).encode())

    def register_batch(self, table: str, partition: int, batch: pa.RecordBatch):
        arr, sch = export_batch(batch)
        _check(lib().oracle_register_batch(self.h, table.encode(), partition, C.addressof(arr), C.addressof(sch)))
        self._parts.setdefault(table, set()).add(partition)

    def drop_table(self, table):
        lib().oracle_drop_table(self.h, table.encode())
        self._parts.pop(table, None)

    def tpch_generate(self, table, msf, partition, row_begin, row_end, columns: Optional[List[str]] = None):
        csv = ",".join(columns).encode() if columns else None
        _check(lib().oracle_tpch_generate(self.h, table.encode(), msf, partition, row_begin, row_end, csv))
        self._parts.setdefault(table, set()).add(partition)

    def export_table(self, table, partition) -> pa.RecordBatch:
        arr, sch = ArrowArray(), ArrowSchema()
        _check(lib().oracle_export_table(self.h, table.encode(), partition, C.addressof(arr), C.addressof(sch)))
        return import_batch(arr, sch)

    def n_table_partitions(self, table):
        return max(self._parts[table]) + 1

    def create_query_stage_exec(self, job_id, stage_id, plan_json):
        return _StageExec(self, job_id, stage_id, plan_json)

    def partition_export(self, job_id, stage_id, out_partition) -> pa.RecordBatch:
        arr, sch = ArrowArray(), ArrowSchema()
        _check(lib().oracle_partition_export(self.h, job_id.encode(), stage_id, out_partition, None,
                                             C.addressof(arr), C.addressof(sch)))
        return import_batch(arr, sch)

    def partition_rows(self, job_id, stage_id, out_partition) -> int:
        return lib().oracle_partition_rows(self.h, job_id.encode(), stage_id, out_partition)

    def remove_job_data(self, job_id):
        lib().oracle_remove_job_data(self.h, job_id.encode())

    def hash_partition_ids(self, batch: pa.RecordBatch, key_cols: List[int], P: int):
        import numpy as np
        arr, sch = export_batch(batch)
        h = np.zeros(batch.num_rows, dtype=np.uint64)
        pid = np.zeros(batch.num_rows, dtype=np.int32)
        _check(lib().oracle_hash_partition_ids(self.h, C.addressof(arr), C.addressof(sch),
                                               ",".join(map(str, key_cols)).encode(), P,
                                               h.ctypes.data, pid.ctypes.data))
        return h, pid
