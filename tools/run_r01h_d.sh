mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'find_tile' -s 1 -c 1 -o gpurun_out/r01h_find50k python tools/find_probe.py --refs 50000 --queries 2048 --reps 1 > gpurun_out/r01h_find50k.log 2>&1
tail -1 gpurun_out/r01h_find50k.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'find_tile' -s 1 -c 1 -o gpurun_out/r01h_find500k python tools/find_probe.py --refs 500000 --queries 2048 --reps 1 > gpurun_out/r01h_find500k.log 2>&1
tail -1 gpurun_out/r01h_find500k.log
