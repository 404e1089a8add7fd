timeout 80 python -m pytest tests/test_gpu_groupby.py -x -q -m gpu 2>&1 | tail -6
timeout 90 python bench.py --workload all --sf 10 --steps 1 --warmup 1 > gpurun_out/r02_allf.json 2> gpurun_out/r02_allf.err; tail -2 gpurun_out/r02_allf.err | cut -c1-300
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r02_allf.json").read().splitlines() if x.startswith("{")][-1])
    print("all", round(l["ms_per_step"],3), "ms", l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"])
    print({k: round(v,2) for k,v in l["per_query_ms"].items()})
    for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
except Exception as ex: print("no line", ex)
PY
timeout 100 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
