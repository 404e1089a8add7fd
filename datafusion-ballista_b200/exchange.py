"""Exchange step between query stages when there is one executor per GPU.

Reference counterpart: the file-based hash shuffle -- map tasks write
``work_dir/job/stage/...`` (ballista/core/src/execution_plans/mod.rs:66-99) and reduce tasks pull
via Arrow Flight (shuffle_reader.rs:522-602, client.rs:143-220).  Semantically an all-to-all(v):
output partition p of every map task goes to the executor that runs reduce task p.

Here partitions stay in HBM; partition p is owned by rank ``p % world``.  The payload moves with
NCCL (torch.distributed.all_to_all_single over NVLink/NVSwitch) directly between the device
buffers the engine exposes (``b200_partition_device_buffers``) -- no host staging, no compression.
torch is plumbing only (communicator + stream); the engine never sees torch types.
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict, List

import torch
import torch.distributed as dist


_TRACE: Dict[str, float] = {}  # B200_BENCH_TRACE=1: accumulated ms per exchange phase (diagnostic)


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _as_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if torch.device(device).type == "cpu":  # host-memory "partitions" (gloo tests of the exchange logic)
        import ctypes
        return torch.frombuffer((ctypes.c_uint8 * nbytes).from_address(ptr), dtype=torch.uint8)
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()


def exchange_stage(engine, job_id: str, stage_id: int, n_out_partitions: int, schema: List[dict], rank: int, world: int,
                   device, owner=None) -> Dict[str, int]:
    """All ranks call this after finishing their map tasks of `stage_id`.

    For every output partition p (owner = owner(p), default p % world): each rank sends its local
    piece of p to the owner; the owner installs the received pieces under file_id = sender rank.
    Two collectives: the sizes (one small all-to-all) and the payload (one all-to-all-v of a single
    contiguous message per peer, packed by the engine with b200_device_gather).  Returns byte counts.
    """
    owner = owner or (lambda p: p % world)
    trace = _TRACE if os.environ.get("B200_BENCH_TRACE") else None
    t0 = time.perf_counter()

    def mark(name):
        nonlocal t0
        if trace is not None:
            _sync(device)
            t1 = time.perf_counter()
            trace[name] = trace.get(name, 0.0) + (t1 - t0) * 1e3
            t0 = t1

    ncols = len(schema)
    nbuf = 3 * ncols
    schema_json = json.dumps(schema)
    is_cuda = torch.device(device).type == "cuda"
    # 1. metadata: per (dest rank, partition, buffer) sizes + rows
    parts_of = {r: [p for p in range(n_out_partitions) if owner(p) == r] for r in range(world)}
    max_parts = max(1, max(len(v) for v in parts_of.values()))
    meta = torch.zeros((world, max_parts, nbuf + 1), dtype=torch.int64)
    order = []  # buffers in send order
    for r in range(world):
        for k, p in enumerate(parts_of[r]):
            if engine.partition_rows(job_id, stage_id, p) < 0:
                continue
            bufs, rows = engine.partition_device_buffers(job_id, stage_id, p)
            for b, (_, nb) in enumerate(bufs):
                meta[r, k, b] = nb
            meta[r, k, nbuf] = rows
            order.extend(bufs)
    mark("buffers")
    meta_dev = meta.to(device)
    recv_meta = torch.empty_like(meta_dev)
    dist.all_to_all_single(recv_meta.view(world, -1), meta_dev.view(world, -1))
    recv_meta_h = recv_meta.cpu()
    mark("meta")
    # 2. payload: one flat byte buffer per destination
    send_sizes = meta[:, :, :nbuf].sum(dim=(1, 2)).tolist()
    recv_sizes = recv_meta_h[:, :, :nbuf].sum(dim=(1, 2)).tolist()
    total_send = int(sum(send_sizes))
    send = torch.empty(max(total_send, 1), dtype=torch.uint8, device=device)
    if is_cuda:
        # the gather is enqueued on the ENGINE's stream, the collective below on torch's current stream: make the hand-over
        # explicit instead of relying on the caller having made them the same stream
        torch.cuda.current_stream(device).synchronize()   # `send` exists before the engine writes into it
        engine.device_gather(order, send.data_ptr(), total_send)
        engine.synchronize()                               # ... and is complete before NCCL reads it
    else:
        pos = 0
        for (ptr, nb) in order:
            if nb:
                send[pos:pos + nb].copy_(_as_tensor(ptr, nb, device))
                pos += nb
    total_recv = int(sum(recv_sizes))
    recv = torch.empty(max(total_recv, 1), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv[:total_recv], send[:total_send], output_split_sizes=[int(x) for x in recv_sizes],
                           input_split_sizes=[int(x) for x in send_sizes])
    _sync(device)
    mark("payload")
    # 3. install what this rank owns (drop its own un-exchanged local pieces first)
    mine = parts_of[rank]
    pos = 0
    base = recv.data_ptr()
    installs = []
    rm = recv_meta_h.tolist()
    for src in range(world):
        for k, p in enumerate(mine):
            sizes = rm[src][k][:nbuf]
            rows = rm[src][k][nbuf]
            bufs = []
            for nb in sizes:
                bufs.append((base + pos if nb else 0, nb))
                pos += nb
            if rows > 0:
                installs.append((p, src, bufs, rows))
    engine.remove_stage_partitions(job_id, stage_id)
    for p, src, bufs, rows in installs:
        engine.partition_import_device(job_id, stage_id, p, src, schema_json, bufs, rows)
    mark("install")
    return {"sent_bytes": int(total_send - send_sizes[rank]), "recv_bytes": int(total_recv - recv_sizes[rank])}
