mkdir -p gpurun_out
bash tools/run_quick.sh r01h_b
timeout 1200 python tools/kmer_sweep.py --refs 50000,200000,500000 --out gpurun_out/r01h_kmer_sweep.jsonl > gpurun_out/r01h_kmer_sweep.log 2>&1; echo "sweep rc=$?"
tail -20 gpurun_out/r01h_kmer_sweep.log
timeout 1200 python bench.py --kind v4 --refs 500000 --queries 125000 --steps 2 --warmup 1 > gpurun_out/r01h_bench_v4_500k.json 2> gpurun_out/r01h_bench_v4_500k.err; echo "v4 bench rc=$?"
tail -3 gpurun_out/r01h_bench_v4_500k.err
cat gpurun_out/r01h_bench_v4_500k.json
