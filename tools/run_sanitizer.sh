# compute-sanitizer over subsets of the GPU tests: bash tools/run_sanitizer.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_search.py -m gpu -x -q -k "align_golden or graph_matches or wide_indegree or weighted or oversized or penalties or identity_kernel or search_stage or family_identity or insertion_forbid" > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -m3 "ERROR SUMMARY" gpurun_out/${tag}_memcheck.log; tail -2 gpurun_out/${tag}_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_search.py -m gpu -x -q -k "align_golden or wide_indegree or identity_kernel or graph_matches" > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "Race reported\|hazard" gpurun_out/${tag}_racecheck.log; grep -m8 "hazard\|Race reported\|at 0x\|in .*kernel" gpurun_out/${tag}_racecheck.log | cut -c1-220; grep -m3 "RACECHECK SUMMARY" gpurun_out/${tag}_racecheck.log; tail -2 gpurun_out/${tag}_racecheck.log
