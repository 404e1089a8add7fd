B200_TIMING=1 python - <<'PY' 2>&1 | tail -60
import os, sys, time
sys.path.insert(0, os.getcwd())
import ballista_b200 as bb
from ballista_b200 import tpch
eng = bb.GpuExecutionEngine(0)
n = 59986052
eng.tpch_generate("lineitem", 10000, 0, 0, n, tpch.Q1_COLUMNS)
stages = tpch.q1(1)
for rep in range(3):
    job = f"j{rep}"
    print(f"--- rep {rep}", file=sys.stderr)
    for sid, st in enumerate(stages, 1):
        t0 = time.perf_counter()
        s = eng.create_query_stage_exec(job, sid, st.json(job))
        t1 = time.perf_counter()
        s.execute_query_stage(0)
        t2 = time.perf_counter()
        s.release()
        print(f"[py] stage {sid}: prepare {1e3*(t1-t0):.3f} ms execute {1e3*(t2-t1):.3f} ms", file=sys.stderr)
    t0 = time.perf_counter(); out = eng.partition_export(job, 3, 0); eng.remove_job_data(job)
    print(f"[py] export+remove {1e3*(time.perf_counter()-t0):.3f} ms", file=sys.stderr)
PY
