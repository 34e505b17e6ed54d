mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/run_chunks.sh
timeout 600 python tools/cli_bench.py --refs 5000 --queries 20000 > gpurun_out/r01j_cli_bench.log 2>&1; cat gpurun_out/r01j_cli_bench.log
