"""Print the device time of the fused stage-1 kernel for q1 / q6 (diagnostic)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ballista_b200 as bb
from ballista_b200 import tpch

msf = int(os.environ.get("MSF", "1000"))
reps = int(os.environ.get("REPS", "3"))
which = os.environ.get("Q", "q1,q6").split(",")
eng = bb.GpuExecutionEngine(0)
L = bb.engine.load_library()
import ctypes as C
sys.path.insert(0, os.path.join(ROOT, "tests"))
n = {1000: 5999995, 10000: 59986052}.get(msf) or int(6000000 * msf / 1000)
for q in which:
    cols = tpch.Q1_COLUMNS if q == "q1" else tpch.Q6_COLUMNS
    stages = tpch.q1(1) if q == "q1" else tpch.q6(1)
    eng.drop_table("lineitem")
    eng.tpch_generate("lineitem", msf, 0, 0, n, cols)
    for r in range(reps):
        job = f"{q}-{r}"
        s1 = eng.create_query_stage_exec(job, 1, stages[0].json(job))
        s1.execute_query_stage(0)
        m = [x for x in s1.collect_plan_metrics() if x["name"] == "AggregateExec"][0]
        ms = m["elapsed_compute_ns"] / 1e6
        by = m["bytes_read"]
        print(f"{q} rep{r} rows={n} kernel_ms={ms:.3f} GB/s={by/ms/1e6:.1f} launches={m['kernel_launches']}", flush=True)
        s1.release()
        eng.remove_job_data(job)
