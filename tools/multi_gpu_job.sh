# N-GPU measurement job (run under gpurun --gpus N).  Usage:
#   N=4 SF=100 WORKLOADS="q5" bash tools/multi_gpu_job.sh          # BASELINE.json configs[2]
#   N=8 SF=100 WORKLOADS="q17 all" Q1=0 bash tools/multi_gpu_job.sh # configs[3], configs[4]
# Every bench.py call prints one JSON line into gpurun_out/r02_<workload>_n<N>.json; tools/collect_bench_lines.py copies
# the lines into profiles/.  Mind the budget: a call on N GPUs is charged N x its duration.
N=${N:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for W in ${WORKLOADS:-q5}; do
ST=3; [ $W = all ] && ST=2
timeout 900 $TR bench.py --gpus $N --workload $W --sf ${SF:-100} --steps $ST --warmup 1 > gpurun_out/r02_${W}_n$N.json 2> gpurun_out/r02_${W}_n$N.err; grep -E "Error|error|Traceback" -A3 gpurun_out/r02_${W}_n$N.err | tail -8
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r02_${W}_n$N.json").read().splitlines() if x.startswith("{")][-1])
    print("$W N=$N", round(l["ms_per_step"],3), "ms", l["value"], l["parity"].get("equal"), l["self_consistent_at_full_scale"], "qph", l["queries_per_hour"], l.get("fused_shuffle"))
    print("   exchange", l["exchange_rank0"])
    if len(l["per_query_ms"])>1: print({k: round(v,2) for k,v in l["per_query_ms"].items()})
    for k,v in l["kernels"].items(): print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
except Exception as ex: print("no $W line", ex)
PY
done
if [ "${Q1:-1}" = 1 ]; then
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_q1_n$N.json 2> gpurun_out/r02_q1_n$N.err; grep -E "Error|error|Traceback" gpurun_out/r02_q1_n$N.err | tail -6
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r02_q1_n$N.json").read().splitlines() if x.startswith("{")][-1])
    print("q1 N=$N", l["ms_per_step"], l["value"], l["roofline"]["frac"], "e2e", l["e2e"]["ms_per_step"], l["e2e"]["value"], "parity", l["parity_checked"])
except Exception as ex: print("no q1 line", ex)
PY
fi
