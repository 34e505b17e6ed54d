mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r01l_pytest.log
tail -3 gpurun_out/r01l_pytest.log
timeout 900 python bench.py > gpurun_out/r01l_bench.json 2> gpurun_out/r01l_bench.err; echo "bench rc=$?"
cat gpurun_out/r01l_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01l_ref_arm.json 2>/dev/null; cat gpurun_out/r01l_ref_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01l_launches_bench.log 2>&1
SG_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_v2_kernel|find_tile|graph_kernel|backtrack_kernel|find_merge|query_kmers' --launch-skip 0 -c 14 -o gpurun_out/r01l_full python tools/dp_probe.py --refs 50000 --queries 888 --reps 1 > gpurun_out/r01l_full.log 2>&1
tail -2 gpurun_out/r01l_full.log
ls -la gpurun_out | tail -12
