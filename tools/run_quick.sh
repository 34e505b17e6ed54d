# GPU parity tests + DP probe + (optional) one default bench line. usage: bash tools/run_quick.sh <tag> [bench]
tag=${1:-quick}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python tools/dp_probe.py --refs 20000 --queries 1184 --reps 3 2>&1 | tee gpurun_out/${tag}_probe.log | tail -3
if [ "$2" = "bench" ]; then
  timeout 900 python bench.py --no-cpu-baseline --no-shares > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
  cat gpurun_out/${tag}_bench.json
fi
