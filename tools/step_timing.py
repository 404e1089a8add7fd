"""Host-side timeline of one resident-input q1 step (B200_TIMING=1 makes the engine print per-operator and per-phase
host times to stderr).  Usage: B200_TIMING=1 python tools/step_timing.py [msf]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ballista_b200 as bb
from ballista_b200 import tpch

msf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
eng = bb.GpuExecutionEngine(0)
n = eng.tpch_table_rows("lineitem", msf)
eng.tpch_generate("lineitem", msf, 0, 0, n, tpch.Q1_COLUMNS)
stages = tpch.q1(1)
for it in range(4):
    job = f"t{it}"
    print(f"---- iteration {it}", file=sys.stderr)
    t0 = time.perf_counter()
    for st in stages:
        ta = time.perf_counter()
        q = eng.create_query_stage_exec(job, st.stage_id, st.json(job))
        tb = time.perf_counter()
        q.execute_query_stage(0)
        tc = time.perf_counter()
        q.release()
        print(f"[py] stage {st.stage_id}: prepare {1e3 * (tb - ta):.3f} ms, execute {1e3 * (tc - tb):.3f} ms", file=sys.stderr)
    ta = time.perf_counter()
    res = eng.partition_export(job, 3, 0)
    tb = time.perf_counter()
    eng.remove_job_data(job)
    print(f"[py] export {1e3 * (tb - ta):.3f} ms, step {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
eng.close()
