N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus $N --workload all --sf 50 --steps 2 --warmup 1 > gpurun_out/r02_all50_n2.json 2> gpurun_out/r02_all50_n2.err; grep -E "Error|error|Traceback" -A5 gpurun_out/r02_all50_n2.err | tail -12
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r02_all50_n2.json").read().splitlines() if x.startswith("{")][-1])
    print("all SF50 N=2", round(l["ms_per_step"],3), "ms", l["value"], l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"], l.get("fused_shuffle"))
    print({k: round(v,2) for k,v in l["per_query_ms"].items()})
    for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
except Exception as ex: print("no line", ex)
PY
