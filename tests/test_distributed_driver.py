"""Host logic of the N-executor stage driver (ballista_b200/driver.py::run_stages_distributed) without a GPU: a recording fake
stands in for the engine and the walk is replayed for every rank of a 4-executor gang over the TPC-H stage plans.  What must
hold for the gang not to deadlock or lose data:
  * every rank issues the SAME sequence of collectives (exchange_stage / execute_query_stage_exchange), whatever tasks it runs;
  * the fused writer+exchange call is chosen only where it is defined: hash-repartitioned output without string columns, the
    same number of map tasks on every executor;
  * single-task stages and stages over replicated dimension tables run on rank 0 only; hash-placed reduce tasks run where
    their partition lives (p % world)."""
import pytest

from ballista_b200 import driver, tpch
from ballista_b200.engine import EXCHANGE_BROADCAST, EXCHANGE_GATHER, EXCHANGE_HASH


class _Stage:
    def __init__(self, eng, stage_id):
        self.eng, self.stage_id = eng, stage_id

    def execute_query_stage(self, p):
        self.eng.log.append(("task", self.stage_id, p))
        return []

    def execute_query_stage_exchange(self, p):
        self.eng.log.append(("fused", self.stage_id, p))
        return [], {"sent_bytes": 0, "recv_bytes": 0}

    def collect_plan_metrics(self):
        return []

    def release(self):
        pass


class RecordingEngine:
    def __init__(self, parts_per_table=2):
        self.log = []
        self.parts = parts_per_table

    def n_table_partitions(self, table):
        return 1 if table in driver.REPLICATED_TABLES else self.parts

    def create_query_stage_exec(self, job_id, stage_id, plan_json):
        return _Stage(self, stage_id)

    def exchange_stage(self, job_id, stage_id, n_out, schema, mode=EXCHANGE_HASH, root=0):
        self.log.append(("exchange", stage_id, mode, n_out))
        return {"sent_bytes": 0, "recv_bytes": 0}


def _walk(stages, world, fused, parts=2):
    logs = []
    for rank in range(world):
        e = RecordingEngine(parts)
        driver.run_stages_distributed(e, stages, "job", rank, world, collect=False, fused=fused)
        logs.append(e.log)
    return logs


def _collectives(log):
    out = []
    for ev in log:
        if ev[0] == "exchange":
            out.append(ev)
        elif ev[0] == "fused":
            out.append(("fused", ev[1]))
    return out


def _stage_schema(stages, stage_id):
    for later in stages:
        rs = driver._readers_of(later.plan["input"], stage_id)
        if rs:
            return rs[0][2]
    return None


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("qname", sorted(tpch.QUERIES, key=lambda q: int(q[1:])))
def test_every_rank_issues_the_same_collectives(qname, fused):
    world = 4
    stages = tpch.QUERIES[qname][1](2 * world)
    logs = _walk(stages, world, fused)
    want = _collectives(logs[0])
    for rank in range(1, world):
        assert _collectives(logs[rank]) == want, f"{qname}: rank {rank} would not meet rank 0 in the same collectives"
    # every stage but the last is exchanged exactly once per map task (fused) or once per stage (two-step)
    for st in stages[:-1]:
        ex = [c for c in want if c[1] == st.stage_id]
        assert ex, f"{qname}: stage {st.stage_id} is never exchanged"
        kinds = {c[0] for c in ex}
        assert len(kinds) == 1
        if kinds == {"exchange"}:
            assert len(ex) == 1
    if not fused:
        assert not [c for c in want if c[0] == "fused"]


@pytest.mark.parametrize("qname", ["q5", "q17", "q3", "q12", "q18"])
def test_fused_only_where_defined(qname):
    world = 4
    stages = tpch.QUERIES[qname][1](2 * world)
    logs = _walk(stages, world, True)
    for log in logs:
        fused_stages = {ev[1] for ev in log if ev[0] == "fused"}
        for sid in fused_stages:
            schema = _stage_schema(stages, sid)
            assert schema is not None and driver._fixed_width(schema), f"{qname}: stage {sid} has string columns but was fused"
            st = [s for s in stages if s.stage_id == sid][0]
            assert st.plan.get("partitioning"), f"{qname}: stage {sid} is not hash partitioned"
            assert st.n_tasks != 1
        # a fused stage never also goes through exchange_stage
        assert not fused_stages & {ev[1] for ev in log if ev[0] == "exchange"}
    # q5 / q17 do have fixed-width shuffles (lineitem, orders / lineitem): the fused path is exercised
    if qname in ("q5", "q17"):
        assert any(ev[0] == "fused" for ev in logs[0])
    # same number of fused calls per stage on every rank
    per_rank = [sorted((ev[1], ev[0]) for ev in log if ev[0] == "fused") for log in logs]
    assert all(p == per_rank[0] for p in per_rank)


def test_task_placement_q5():
    world = 4
    P = 2 * world
    stages = tpch.q5(P)
    logs = _walk(stages, world, False)
    tables = {}
    for st in stages:
        kind, what = driver._probe_side_leaf(st.plan["input"])
        tables[st.stage_id] = (kind, what, st.n_tasks)
    for rank, log in enumerate(logs):
        for ev in log:
            if ev[0] != "task":
                continue
            kind, what, n_tasks = tables[ev[1]]
            if n_tasks == 1 or (kind == "table" and what in driver.REPLICATED_TABLES):
                assert rank == 0, f"stage {ev[1]} is a single-task / replicated-table stage and must run on rank 0 only"
    # hash-placed reduce tasks: partition p runs on rank p % world, each exactly once over the gang
    modes = {ev[1]: ev[2] for ev in logs[0] if ev[0] == "exchange"}
    for st in stages:
        kind, what, n_tasks = tables[st.stage_id]
        if kind == "stage" and n_tasks != 1 and modes.get(what) == EXCHANGE_HASH:
            seen = {}
            for rank, log in enumerate(logs):
                for ev in log:
                    if ev[0] == "task" and ev[1] == st.stage_id:
                        assert ev[2] % world == rank
                        seen[ev[2]] = seen.get(ev[2], 0) + 1
            assert sorted(seen) == list(range(P)) and set(seen.values()) == {1}
    # the broadcast build side and the final gather are there
    assert EXCHANGE_BROADCAST in modes.values() and EXCHANGE_GATHER in modes.values()


def test_tables_the_loader_replicated_are_scanned_once():
    """A table the engine's loader put on every executor in full (tiny tables at small scale factors) must be scanned by one
    executor only, whatever its name -- otherwise every rank contributes the same rows."""
    world = 4
    stages = tpch.q5(2 * world)
    leaf = {st.stage_id: driver._probe_side_leaf(st.plan["input"]) for st in stages}
    supplier_stages = [sid for sid, (kind, what) in leaf.items() if kind == "table" and what == "supplier"]
    assert supplier_stages
    for rank in range(world):
        e = RecordingEngine(2)
        e.replicated_tables = {"supplier"}
        driver.run_stages_distributed(e, stages, "job", rank, world, collect=False)
        ran = {ev[1] for ev in e.log if ev[0] == "task"}
        assert (set(supplier_stages) <= ran) == (rank == 0)
