# one profiling round on the GPU box: tag given as $1 (e.g. r02i)
tag=${1:-r02x}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-shares > gpurun_out/${tag}_launches_bench.log 2>&1
SG_STREAMS=1 SG_BATCH=1184 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_kernel|find_tile|graph_kernel|backtrack_kernel|find_merge|query_kmers|family_kernel' --launch-skip 0 -c 14 -o gpurun_out/${tag}_full python tools/dp_probe.py --refs 50000 --queries 1184 --reps 1 > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
ls -la gpurun_out | tail -6
