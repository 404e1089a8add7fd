python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_gpu4.log; tail -6 gpurun_out/r02_gpu4.log
show() {
  python - <<PY
import json
try:
    l=json.load(open("gpurun_out/$1.json"))
    print("$1", round(l["ms_per_step"],3), "ms", l["value"], l["parity"].get("equal"), l["self_consistent_at_full_scale"])
    if "per_query_ms" in l and len(l["per_query_ms"])>1: print({k: round(v,2) for k,v in l["per_query_ms"].items()}, "qph", l["queries_per_hour"])
    for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
except Exception as ex: print("no line", ex)
PY
}
python bench.py --workload q5 --sf 10 --steps 2 --warmup 1 --partitions-per-gpu 8 > gpurun_out/r02_q5p8b.json 2> gpurun_out/r02_q5p8b.err; tail -3 gpurun_out/r02_q5p8b.err; show r02_q5p8b
python bench.py --workload all --sf 10 --steps 2 --warmup 1 > gpurun_out/r02_allb.json 2> gpurun_out/r02_allb.err; tail -5 gpurun_out/r02_allb.err; show r02_allb
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; tail -3 gpurun_out/r02_bench3.err; python - <<PY
import json
l=json.load(open("gpurun_out/r02_bench3.json"))
print(l["ms_per_step"], l["value"], l["roofline"]["frac"], l["e2e"]["ms_per_step"], l["e2e"]["value"], l["parity_checked"], l.get("cpu_baseline",{}).get("value"))
PY
