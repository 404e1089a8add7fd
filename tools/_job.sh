N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for MODE in wstaged direct; do
export B200_SCATTER_STAGED=-1; [ $MODE = direct ] && export B200_SCATTER_STAGED=0
timeout 600 $TR bench.py --gpus $N --workload q5 --sf 50 --steps 3 --warmup 1 --partitions-per-gpu ${PPG:-1} > gpurun_out/r02_q5_n2_$MODE.json 2> gpurun_out/r02_q5_n2_$MODE.err; grep -E "Error|error|Traceback" -A3 gpurun_out/r02_q5_n2_$MODE.err | tail -8
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r02_q5_n2_$MODE.json").read().splitlines() if x.startswith("{")][-1])
    print("q5 N=2 $MODE", round(l["ms_per_step"],3), "ms", l["value"], l["parity"].get("equal"), l["self_consistent_at_full_scale"], l.get("fused_shuffle"))
    for k,v in l["kernels"].items():
        if "scatter" in k: print("   ", k, round(v["ms_per_step"],3), round(v["launches_per_step"],1), round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3))
except Exception as ex: print("no line", ex)
PY
done
unset B200_SCATTER_STAGED
timeout 900 python -m pytest tests/test_gpu_exchange_nccl.py -x -q -m gpu 2>&1 | tail -5
