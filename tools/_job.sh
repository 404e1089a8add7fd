python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_gpu7.log; tail -6 gpurun_out/r02_gpu7.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; tail -3 gpurun_out/r02_bench4.err; python - <<PY
import json
l=json.load(open("gpurun_out/r02_bench4.json"))
print("q1", l["ms_per_step"], l["value"], "frac", l["roofline"]["frac"], "kernel_ms", l["roofline"]["kernel_ms"], "e2e", l["e2e"]["ms_per_step"], l["e2e"]["value"], l["parity_checked"], l.get("cpu_baseline",{}).get("value"))
PY
python bench.py --workload all --sf 10 --steps 2 --warmup 1 > gpurun_out/r02_alld.json 2> gpurun_out/r02_alld.err; tail -3 gpurun_out/r02_alld.err
python - <<PY
import json
l=json.load(open("gpurun_out/r02_alld.json"))
print("all", round(l["ms_per_step"],3), "ms", l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"])
print({k: round(v,2) for k,v in l["per_query_ms"].items()})
for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
PY
python tools/op_probe.py groupby > gpurun_out/r02_probe_groupby2.json 2>/dev/null; cat gpurun_out/r02_probe_groupby2.json | tr -d '\n '; echo
REPS=3 python tools/op_probe.py q1 > gpurun_out/r02_probe_q1b.json 2>/dev/null; cat gpurun_out/r02_probe_q1b.json | tr -d '\n '; echo
