"""GPU: stages prepared from the protobuf plan bytes (b200_stage_prepare_proto, the form in which a Ballista scheduler ships a
task's plan) give the oracle's answer -- the fixtures of tests/golden/proto_plans.json for whole queries, every stage created
from bytes, none from IR text."""
import base64
import json
import os

import pytest

from ballista_b200 import driver, tpch
from test_tpch_queries import load_tables
from util import assert_tables_equal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "proto_plans.json")) as fh:
    PROTO = {c["name"]: base64.b64decode(c["proto_b64"]) for c in json.load(fh)["cases"]}


class _FromProto:
    """The engine with create_query_stage_exec swapped for the protobuf entry point (the IR text the driver passes is ignored)."""

    def __init__(self, eng, query):
        self._e, self._q = eng, query

    def __getattr__(self, name):
        return getattr(self._e, name)

    def create_query_stage_exec(self, job_id, stage_id, plan_json):
        return self._e.create_query_stage_exec_proto(job_id, stage_id, PROTO[f"{self._q}/stage{stage_id}"])


@pytest.mark.parametrize("q,tables,ordered", [("q1", {"lineitem": tpch.Q1_COLUMNS}, True), ("q6", {"lineitem": tpch.Q6_COLUMNS}, True),
                                               ("q5", tpch.Q5_TABLES, True), ("q12", tpch.Q12_TABLES, True), ("q4", None, True)])
def test_query_from_plan_bytes(gpu, oracle, oracle_lib, q, tables, ordered):
    if tables is None:
        tables = tpch.union_tables([q])
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 50, tables, 2)
    stages = tpch.QUERIES[q][1](4)          # the plans the fixtures were generated from (4 shuffle partitions)
    got = driver.run_stages(_FromProto(gpu, q), stages, f"{q}-proto")
    want = driver.run_stages(oracle, stages, f"{q}-proto")
    assert want is not None and want.num_rows > 0
    assert_tables_equal(got, want, sort=not ordered, f64_rtol=1e-12)


def test_stage_from_task_definition(gpu, oracle, oracle_lib):
    """b200_stage_prepare_task: the whole MultiTaskDefinition as received -- props applied, plan prepared, one execute per task id."""
    with open(os.path.join(HERE, "golden", "proto_plans.json")) as fh:
        t = json.load(fh)["tasks"]
    for e in (gpu, oracle):
        load_tables(e, oracle_lib, 30, {"lineitem": tpch.Q5_TABLES["lineitem"]}, 3)
    try:
        qse, info = gpu.create_query_stage_exec_task(base64.b64decode(t["multi_b64"]), multi=True)   # props: batch_size = 1024
        assert info == t["multi"]
        stage5 = [s for s in tpch.q5(4) if s.stage_id == 5][0]
        ref = oracle.create_query_stage_exec(info["job_id"], 5, stage5.json(info["job_id"]))
        for task in info["tasks"]:
            got = qse.execute_query_stage(task["partition_id"])
            want = ref.execute_query_stage(task["partition_id"])
            assert [(w.partition_id, w.num_rows) for w in got] == [(w.partition_id, w.num_rows) for w in want]
            assert sum(w.num_rows for w in got) > 0
            assert all(w.num_batches == -(-w.num_rows // 1024) for w in got)                          # the task's props were applied
        qse.release()
        ref.release()
        for p in range(4):
            assert_tables_equal(gpu.partition_export(info["job_id"], 5, p), oracle.partition_export(info["job_id"], 5, p), sort=False)
    finally:
        gpu.set_config("datafusion.execution.batch_size", "8192")
