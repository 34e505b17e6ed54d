mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r01g_pytest.log
tail -5 gpurun_out/r01g_pytest.log
timeout 900 python bench.py > gpurun_out/r01g_bench.json 2> gpurun_out/r01g_bench.err; echo "bench rc=$?"
cat gpurun_out/r01g_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01g_launches_bench.log 2>&1
SG_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_v2_kernel|find_tile|graph_kernel|backtrack_kernel|find_merge|kmer_extract' --launch-skip 0 -c 12 -o gpurun_out/r01g_full python tools/dp_probe.py --refs 50000 --queries 1184 --reps 1 > gpurun_out/r01g_full.log 2>&1
tail -3 gpurun_out/r01g_full.log
ls -la gpurun_out
