mkdir -p gpurun_out
bash tools/run_quick.sh r01h_c
timeout 600 python tools/kmer_sweep.py --refs 500000 --k 10,12 --modes fast --out gpurun_out/r01h_c_sweep.jsonl > gpurun_out/r01h_c_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/r01h_c_sweep.jsonl
SG_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'graph_kernel' -c 1 -o gpurun_out/r01h_graph python tools/dp_probe.py --refs 50000 --queries 1184 --reps 1 > gpurun_out/r01h_graph.log 2>&1
tail -2 gpurun_out/r01h_graph.log
ls -la gpurun_out
