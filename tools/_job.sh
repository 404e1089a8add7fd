python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_gpu8.log; tail -8 gpurun_out/r02_gpu8.log
THREADS=8,12,16,24 CHUNKS=1048576,4194304,16777216 SLOTS=3 python tools/ingest_probe.py 2>&1 | tail -14
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; tail -3 gpurun_out/r02_bench5.err; python - <<PY
import json
l=json.loads([x for x in open("gpurun_out/r02_bench5.json").read().splitlines() if x.startswith("{")][-1])
print("q1", l["ms_per_step"], l["value"], "frac", l["roofline"]["frac"], "e2e", l["e2e"]["ms_per_step"], l["e2e"]["value"], l["parity_checked"], l.get("cpu_baseline",{}).get("value"))
PY
python bench.py --workload all --sf 10 --steps 2 --warmup 1 > gpurun_out/r02_alle.json 2> gpurun_out/r02_alle.err; tail -3 gpurun_out/r02_alle.err
python - <<PY
import json
l=json.loads([x for x in open("gpurun_out/r02_alle.json").read().splitlines() if x.startswith("{")][-1])
print("all", round(l["ms_per_step"],3), "ms", l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"])
for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
PY
python tools/op_probe.py parquet 2> gpurun_out/r02_probe_parquet.err | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_probe_parquet.err
