mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "align_golden or wide_indegree or find_golden or insertion_forbid" > gpurun_out/r01k_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "Race reported\|hazard" gpurun_out/r01k_racecheck.log; grep -m8 "hazard\|Race reported\|at 0x\|in .*kernel" gpurun_out/r01k_racecheck.log | cut -c1-220; tail -4 gpurun_out/r01k_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or chunk_pipeline or index_and_find or pipeline_golden" > gpurun_out/r01k_memcheck2.log 2>&1; echo "memcheck2 rc=$?"; tail -3 gpurun_out/r01k_memcheck2.log
