mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r01i_bench_n$N.json 2> gpurun_out/r01i_bench_n$N.err; echo "n$N rc=$?"
tail -2 gpurun_out/r01i_bench_n$N.err
python -c "
import json; j=json.loads(open('gpurun_out/r01i_bench_n$N.json').read().strip().splitlines()[-1]); print(j['n_gpus'], round(j['value']), round(j['e2e']['value']), j['ms_per_step'], j['clocks'])"
