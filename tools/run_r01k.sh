mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_cli.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/cli_bench.py --refs 5000 --queries 40000 > gpurun_out/r01k_cli_bench.log 2>&1; cat gpurun_out/r01k_cli_bench.log
python bench.py --no-cpu-baseline > gpurun_out/r01k_bench.json; python -c "
import json; j=json.load(open('gpurun_out/r01k_bench.json')); print(round(j['value']), round(j['e2e']['value']), j['ms_per_step'], j['stages_ms_per_step_isolated'])"
