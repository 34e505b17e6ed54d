mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r01i_pytest.log
tail -3 gpurun_out/r01i_pytest.log
timeout 900 python bench.py > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err; echo "bench rc=$?"
cat gpurun_out/r01i_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01i_ref_arm.json 2>/dev/null; cat gpurun_out/r01i_ref_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01i_launches_bench.log 2>&1
SG_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_v2_kernel|find_tile|graph_kernel|backtrack_kernel|find_merge|query_kmers' --launch-skip 0 -c 14 -o gpurun_out/r01i_full python tools/dp_probe.py --refs 50000 --queries 888 --reps 1 > gpurun_out/r01i_full.log 2>&1
tail -2 gpurun_out/r01i_full.log
timeout 900 python bench.py --refs 500000 --queries 12500 --steps 2 --warmup 1 > gpurun_out/r01i_bench_full_500k.json 2> gpurun_out/r01i_bench_full_500k.err; echo "full500k rc=$?"; cat gpurun_out/r01i_bench_full_500k.json
timeout 900 python bench.py --kind v4 --refs 500000 --queries 125000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01i_bench_v4_500k.json 2> gpurun_out/r01i_bench_v4_500k.err; echo "v4 rc=$?"; cat gpurun_out/r01i_bench_v4_500k.json
ls -la gpurun_out | tail -12
