"""N>1 host logic of the exchange step (ballista_b200/exchange.py) with world_size 2 on the gloo backend.
The engine is replaced by a host-memory fake that implements the four C-ABI calls the exchange uses
(b200_partition_rows / _device_buffers / _import_device / b200_remove_stage_data); what is checked is the
all-to-all bookkeeping: ownership p % world, sizes, per-sender pieces, nothing lost or duplicated."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHEMA = [{"name": "k", "type": "i64", "nullable": False}, {"name": "s", "type": "utf8", "nullable": True}]


class FakeEngine:
    def __init__(self):
        self.parts = {}   # (job, stage, p) -> list of (file_id, [np buffers], rows)
        self.keep = []

    def add(self, job, stage, p, file_id, rows_k, strs):
        valid = np.array([1 if s is not None else 0 for s in strs], dtype=np.uint8)
        offs = np.zeros(len(strs) + 1, dtype=np.int32)
        chars = bytearray()
        for i, s in enumerate(strs):
            chars += (s or "").encode()
            offs[i + 1] = len(chars)
        bufs = [np.zeros(0, np.uint8), np.array(rows_k, dtype=np.int64).view(np.uint8), np.zeros(0, np.uint8),
                valid, offs.view(np.uint8), np.frombuffer(bytes(chars), dtype=np.uint8).copy()]
        self.parts.setdefault((job, stage, p), []).append((file_id, bufs, len(rows_k)))

    def partition_rows(self, job, stage, p):
        v = self.parts.get((job, stage, p))
        return sum(x[2] for x in v) if v else -1

    def partition_device_buffers(self, job, stage, p):
        (fid, bufs, rows), = self.parts[(job, stage, p)]
        self.keep.append(bufs)
        return [(b.ctypes.data if b.size else 0, int(b.size)) for b in bufs], rows

    def remove_stage_partitions(self, job, stage):
        for k in [k for k in self.parts if k[0] == job and k[1] == stage]:
            del self.parts[k]

    def partition_import_device(self, job, stage, p, file_id, schema_json, bufs, rows):
        assert json.loads(schema_json) == SCHEMA
        import ctypes
        arrs = [np.frombuffer((ctypes.c_uint8 * nb).from_address(ptr), dtype=np.uint8).copy() if nb else np.zeros(0, np.uint8) for ptr, nb in bufs]
        self.parts.setdefault((job, stage, p), []).append((file_id, arrs, rows))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ballista_b200 import exchange
    eng = FakeEngine()
    P = 5
    # map output of this rank: partition p holds keys k with k % P == p (only some partitions non-empty)
    for p in range(P):
        if (p + rank) % 3 == 0:
            continue
        ks = [100 * rank + 10 * p + j for j in range(p + 1 + rank)]
        eng.add("job", 1, p, rank, ks, [None if k % 4 == 0 else f"r{rank}k{k}" for k in ks])
    before = {p: eng.parts[("job", 1, p)][0][2] for p in range(P) if ("job", 1, p) in eng.parts}
    res = exchange.exchange_stage(eng, "job", 1, P, SCHEMA, rank, world, torch.device("cpu"))
    owned = {}
    for (job, stage, p), pieces in eng.parts.items():
        assert p % world == rank, "a rank must only hold the partitions it owns"
        for fid, bufs, rows in pieces:
            ks = bufs[1].view(np.int64).tolist()
            assert len(ks) == rows
            offs = bufs[4].view(np.int32)
            strs = [bytes(bufs[5][offs[i]:offs[i + 1]]).decode() if bufs[3][i] else None for i in range(rows)]
            owned.setdefault(p, []).append((fid, ks, strs))
    q.put((rank, before, owned, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_exchange_all_to_all_world2():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    sent = {r: b for r, b, _, _ in out}
    for rank, _, owned, res in out:
        for p, pieces in owned.items():
            assert p % world == rank
            # one piece per sender that had rows for p, tagged with the sender's rank as file_id
            senders = sorted(f for f, _, _ in pieces)
            assert senders == sorted(r for r in range(world) if p in sent[r])
            for fid, ks, strs in pieces:
                assert ks == [100 * fid + 10 * p + j for j in range(p + 1 + fid)]
                assert strs == [None if k % 4 == 0 else f"r{fid}k{k}" for k in ks]
        assert res["sent_bytes"] >= 0 and res["recv_bytes"] >= 0
    # nothing lost
    total_before = sum(sum(b.values()) for b in sent.values())
    total_after = sum(len(ks) for _, _, owned, _ in out for pieces in owned.values() for _, ks, _ in pieces)
    assert total_before == total_after


def _worker_gather(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ballista_b200 import exchange
    eng = FakeEngine()
    # every rank wrote output partition `rank` of stage 2 (what bench.py's final merge gathers onto rank 0)
    ks = [1000 * rank + j for j in range(3 + rank)]
    eng.add("job", 2, rank, rank, ks, [f"g{k}" for k in ks])
    exchange.exchange_stage(eng, "job", 2, world, SCHEMA, rank, world, torch.device("cpu"), owner=lambda p: 0)
    held = {p: [(fid, bufs[1].view(np.int64).tolist()) for fid, bufs, rows in pieces] for (job, stage, p), pieces in eng.parts.items()}
    q.put((rank, held))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_exchange_gather_to_rank0_world2():
    """owner = rank 0 for every partition: the exchange degenerates into the gather the final merge stage needs"""
    world, port = 2, 29741
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_gather, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert out[1] == {}                                   # rank 1 keeps nothing
    assert sorted(out[0]) == [0, 1]                       # rank 0 holds both partitions ...
    for p in (0, 1):
        assert out[0][p] == [(p, [1000 * p + j for j in range(3 + p)])]   # ... each as one piece tagged with its sender
