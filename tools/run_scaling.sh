# strong-scaling point(s) of bench.py on one box: bash tools/run_scaling.sh <tag> <refs> <total queries> N [N ...]
# (BASELINE configs[3]: 100 000 full-length queries vs 500 000 references, a fixed total split over the GPUs)
tag=$1; refs=$2; total=$3; shift 3
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = 1 ]; then launch="python"; else launch="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N"; fi
  timeout 1500 $launch bench.py --gpus $N --steps 2 --warmup 3 --refs $refs --scaling strong --total-queries $total --no-shares --no-cpu-baseline > gpurun_out/${tag}_strong_n$N.json 2> gpurun_out/${tag}_strong_n$N.err; echo "n$N rc=$?"
  tail -2 gpurun_out/${tag}_strong_n$N.err
  python -c "
import json; j=json.loads(open('gpurun_out/${tag}_strong_n$N.json').read().strip().splitlines()[-1]); print('N', j['n_gpus'], 'value', round(j['value']), 'e2e', round(j['e2e']['value']), 'ms/step', round(j['ms_per_step'],2), j['scaling'], j['clocks'].get('sm_mhz'))"
done
