"""Task-level behaviour of the engine behind the C-ABI: cancellation (Executor::cancel_task drops the task future
and its partial outputs, ballista/executor/src/executor.rs:217-237), K concurrent tasks on one engine
(`concurrent_tasks` workers of the DedicatedExecutor share the process' engine, cpu_bound_executor.rs:94-131,
executor_process.rs:202,263-268) and row order across tiles of FilterExec / ProjectionExec outputs."""
import ctypes as C
import threading
import time

import pyarrow as pa
import pytest

import ballista_b200 as bb
from ballista_b200 import driver, plan as P, tpch
from test_tpch_queries import load_tables
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


def _stored_partitions(gpu, job, stage_ids, n=64):
    return [(s, p) for s in stage_ids for p in range(n) if gpu.partition_rows(job, s, p) >= 0]


def test_cancel_before_start_leaves_nothing(gpu, oracle_lib):
    load_tables(gpu, oracle_lib, 20, tpch.Q3_TABLES, 2)
    st = tpch.q3(3, "BUILDING")[0]
    flag = C.c_int32(1)
    q = gpu.create_query_stage_exec("cancel0", st.stage_id, st.json("cancel0"))
    with pytest.raises(bb.B200Error) as ei:
        q.execute_query_stage(0, cancel_flag=flag)
    assert ei.value.code == -6
    q.release()
    assert _stored_partitions(gpu, "cancel0", [st.stage_id]) == []


def test_cancel_mid_join_drops_partial_outputs(gpu, oracle_lib):
    """The flag flips while the join stage of q3 runs on ~1.5 M lineitem rows: the call must return -6 and the stage must
    have stored nothing; a task that was not cancelled in time is simply retried with an earlier flip."""
    msf = 250
    load_tables(gpu, oracle_lib, msf, tpch.Q3_TABLES, 1)
    stages = tpch.q3(2, "BUILDING")
    job = "cancel1"
    join_stage = None
    for st in stages:     # run the scan stages, stop at the first stage that joins
        if '"HashJoinExec"' in st.json(job):
            join_stage = st
            break
        q = gpu.create_query_stage_exec(job, st.stage_id, st.json(job))
        for p in range(gpu.n_table_partitions(driver._probe_side_leaf(st.plan["input"])[1])):
            q.execute_query_stage(p)
        q.release()
    assert join_stage is not None
    cancelled = False
    for delay in (2e-3, 5e-4, 1e-4, 0.0):
        flag = C.c_int32(0)
        t = threading.Timer(delay, lambda: setattr(flag, "value", 1))
        q = gpu.create_query_stage_exec(job, join_stage.stage_id, join_stage.json(job))
        t.start()
        try:
            q.execute_query_stage(0, cancel_flag=flag)
        except bb.B200Error as ex:
            assert ex.code == -6
            cancelled = True
        t.join()
        q.release()
        if cancelled:
            break
        gpu.remove_stage_partitions(job, join_stage.stage_id)
    assert cancelled
    assert _stored_partitions(gpu, job, [join_stage.stage_id]) == []
    gpu.remove_job_data(job)


def test_four_concurrent_tasks_different_plans(gpu, oracle, oracle_lib):
    """4 host threads x 4 different plans on ONE engine, several rounds; every result must equal the oracle's."""
    msf = 30
    tables = {}
    for tb in (tpch.Q3_TABLES, tpch.Q12_TABLES, tpch.Q4_TABLES, {"lineitem": tpch.Q1_COLUMNS}, {"lineitem": tpch.Q6_COLUMNS}):
        for t, cols in tb.items():
            tables.setdefault(t, [])
            tables[t] += [c for c in cols if c not in tables[t]]
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, msf, tables, 2)
    # the tables hold the union of the four queries' columns: scans carry projections into that layout
    tpch.TABLE_LAYOUT.update(tables)
    try:
        plans = {"q1": tpch.q1(4), "q6": tpch.q6(4), "q12": tpch.q12(3), "q4": tpch.q4(3, "1993-01-01", "1996-01-01")}
    finally:
        tpch.TABLE_LAYOUT.clear()
    want = {k: driver.run_stages(oracle, v, f"cc-{k}") for k, v in plans.items()}
    errors, results = [], {}

    def work(name, rounds=3):
        try:
            for r in range(rounds):
                results[(name, r)] = driver.run_stages(gpu, plans[name], f"cc-{name}-{r}")
        except Exception as ex:  # pragma: no cover
            errors.append((name, repr(ex)))

    threads = [threading.Thread(target=work, args=(k,)) for k in plans]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (name, r), got in results.items():
        assert_tables_equal(got, want[name], sort=name in ("q6",))


def test_filter_keeps_input_order_across_tiles(gpu, oracle, oracle_lib):
    """FilterExec / ProjectionExec are order-preserving in DataFusion; 300 k rows = ~300 tiles of the materialising sink."""
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 50, {"lineitem": ["l_orderkey", "l_linenumber", "l_quantity", "l_shipdate"]}, 1)
    scan = tpch.table_scan("lineitem", ["l_orderkey", "l_linenumber", "l_quantity", "l_shipdate"])
    f = P.filter_(P.binop("<", P.col("l_quantity"), P.lit_dec(2500, 15, 2)), scan)
    st = [P.Stage(1, P.shuffle_writer(f, 1))]
    got = driver.run_stages(gpu, st, "order")
    want = driver.run_stages(oracle, st, "order")
    assert want.num_rows > 100000
    assert_tables_equal(got, want, sort=False)
