for cfg in "3 4096" "3 8192" "6 4096" "6 8192" "8 16384"; do
set -- $cfg
echo "== workers/GPU $1 batch $2"
SINA_B200_WORKERS=$1 python tools/cli_bench.py --refs 5000 --queries 160000 --batch-size $2 2>&1 | grep "Took\|busy"
done
