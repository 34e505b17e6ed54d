# k-mer search A/B on the GPU box: find-related parity tests with the product library, then tools/find_ab.py
# usage: [REFS="50000 500000"] [FIND_AB_ENV="SG_TILE_WARPS=24"] bash tools/run_find_ab.sh <tag> <lib> [<lib> ...]
tag=${1:-findab}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-find or family or turn or search or pipeline_golden or production}" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for refs in ${REFS:-50000 500000}; do
  timeout 600 python tools/find_ab.py --refs $refs --queries 4096 --libs "$@" 2>&1 | tee gpurun_out/${tag}_${refs}.log | grep -v "^$" | tail -12
done
for e in $FIND_AB_ENV; do   # extra runs of the last library on the last reference size with one variable set
  echo "== $e"
  env $e timeout 600 python tools/find_ab.py --worker "${@: -1}" --refs $refs --queries 4096 2>&1 | tee gpurun_out/${tag}_${refs}_$e.log | grep -v "^$" | tail -4
done
