"""Minimal stand-in for the parts of Ballista that sit ABOVE the execution engine in a test or a
benchmark: it walks a list of stages in dependency order and, per stage, runs one task per input
partition -- what the scheduler's task binding (ballista/scheduler/src/cluster/mod.rs:354,438)
plus ``Executor::execute_query_stage`` (ballista/executor/src/executor.rs:186-212) amount to when
there is one executor.  It is deliberately engine-agnostic (the CPU oracle in tests/ exposes the
same Python surface) and contains no data-path logic.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import pyarrow as pa

from .plan import Stage


def _leaf_info(node: dict):
    """(scan tables, upstream stage ids) reachable from a plan node."""
    tables, readers = [], []
    op = node.get("op")
    if op in ("DataSourceExec", "MemoryScan", "Scan"):
        tables.append(node["table"])
    elif op in ("ShuffleReaderExec", "UnresolvedShuffleExec"):
        readers.append(node["stage_id"])
    for k in ("input", "left", "right"):
        if k in node:
            t, r = _leaf_info(node[k])
            tables += t
            readers += r
    return tables, readers


def _probe_side_leaf(node: dict):
    """The leaf that decides the task count: for a hash join it is the probe (right) side."""
    op = node.get("op")
    if op in ("DataSourceExec", "MemoryScan", "Scan"):
        return ("table", node["table"])
    if op in ("ShuffleReaderExec", "UnresolvedShuffleExec"):
        return ("stage", node["stage_id"])
    if op in ("HashJoinExec", "SortMergeJoinExec"):
        return _probe_side_leaf(node["right"])
    return _probe_side_leaf(node["input"])


def run_stages(engine, stages: List[Stage], job_id: str = "job", collect: bool = True,
               metrics_out: Optional[list] = None) -> Optional[pa.Table]:
    """Execute `stages` in order on `engine`; return the last stage's output as one Table."""
    out_parts: Dict[int, int] = {}
    last_stats = []
    for st in stages:
        root = st.plan
        if st.n_tasks is not None:
            n_tasks = st.n_tasks
        else:
            kind, what = _probe_side_leaf(root["input"])
            n_tasks = engine.n_table_partitions(what) if kind == "table" else out_parts[what]
        qse = engine.create_query_stage_exec(job_id, st.stage_id, st.json(job_id))
        stats = []
        for p in range(n_tasks):
            stats += qse.execute_query_stage(p)
        if metrics_out is not None:
            metrics_out.append((st.stage_id, qse.collect_plan_metrics()))
        qse.release()
        part = root.get("partitioning")
        out_parts[st.stage_id] = part["n"] if part else n_tasks
        last_stats = stats
    if not collect:
        return None
    last = stages[-1]
    batches = []
    for p in range(out_parts[last.stage_id]):
        if engine.partition_rows(job_id, last.stage_id, p) >= 0:
            batches.append(engine.partition_export(job_id, last.stage_id, p))
    if not batches:
        return None
    return pa.Table.from_batches(batches)
