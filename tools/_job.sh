bash tools/ncu_r02.sh 2>&1 | tail -150
for op in join partition groupby groupby_small filter q1; do echo "== $op"; cat gpurun_out/r02_probe_$op.json | tr -d '\n' | cut -c1-900; echo; done
