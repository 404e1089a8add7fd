python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_gpu8.log; tail -8 gpurun_out/r02_gpu8.log
python bench.py --workload all --sf 10 --steps 2 --warmup 1 > gpurun_out/r02_alle.json 2> gpurun_out/r02_alle.err; tail -3 gpurun_out/r02_alle.err
python - <<PY
import json
l=json.load(open("gpurun_out/r02_alle.json"))
print("all", round(l["ms_per_step"],3), "ms", l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"])
for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
PY
python tools/op_probe.py parquet 2> gpurun_out/r02_probe_parquet.err | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_probe_parquet.err
