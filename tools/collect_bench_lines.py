"""Collect the JSON lines of this round's bench runs (gpurun_out/ is scratch) into profiles/r02_bench_lines.json.

  python tools/collect_bench_lines.py
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [
    ("q1 SF10, 1 GPU (BASELINE.json configs[1])", "r02_bench5.json"),
    ("q1 SF10 per GPU, 4 GPUs (weak scaling)", "r02_q1_n4.json"),
    ("all 22 queries SF10, 1 GPU", "r02_allf.json"),
    ("all 22 queries SF10, 1 GPU (before partition-first aggregation)", "r02_alle.json"),
    ("q5 SF50, 2 GPUs, fused shuffle (warp-staged peer scatter)", "r02_q5_n2_wstaged.json"),
    ("q5 SF50, 2 GPUs, fused shuffle (direct peer scatter)", "r02_q5_n2_direct.json"),
    ("q5 SF50, 2 GPUs, two-step shuffle (writer, then NCCL exchange)", "r02_q5_n2_twostep.json"),
    ("all 22 queries SF50, 2 GPUs", "r02_all50_n2.json"),
    ("q5 SF100, 4 GPUs (BASELINE.json configs[2])", "r02_q5_n4.json"),
    ("q17 SF100, 8 GPUs (BASELINE.json configs[3])", "r02_q17_n8.json"),
    ("all 22 queries SF100, 8 GPUs (BASELINE.json configs[4])", "r02_all_n8.json"),
]
out = []
for what, f in SRC:
    p = os.path.join(ROOT, "gpurun_out", f)
    if not os.path.exists(p):
        continue
    lines = [x for x in open(p).read().splitlines() if x.startswith("{")]
    if not lines:
        continue
    line = json.loads(lines[-1])
    out.append({"what": what, "source": "gpurun_out/" + f, "line": line})
with open(os.path.join(ROOT, "profiles", "r02_bench_lines.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(len(out), "lines")
