N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for W in q17 all; do
ST=3; [ $W = all ] && ST=2
timeout 900 $TR bench.py --gpus $N --workload $W --sf 100 --steps $ST --warmup 1 > gpurun_out/r02_${W}_n$N.json 2> gpurun_out/r02_${W}_n$N.err; grep -E "Error|error|Traceback" -A3 gpurun_out/r02_${W}_n$N.err | tail -8
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r02_${W}_n$N.json").read().splitlines() if x.startswith("{")][-1])
    print("$W N=$N", round(l["ms_per_step"],3), "ms", l["value"], l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"], l.get("fused_shuffle"))
    if len(l["per_query_ms"])>1: print({k: round(v,2) for k,v in l["per_query_ms"].items()})
    for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
except Exception as ex: print("no $W line", ex)
PY
done
