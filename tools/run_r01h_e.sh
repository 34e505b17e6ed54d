mkdir -p gpurun_out
SG_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mesh_v2_kernel' -s 1 -c 1 -o gpurun_out/r01h_mesh python tools/dp_probe.py --refs 50000 --queries 1184 --reps 1 > gpurun_out/r01h_mesh.log 2>&1
tail -2 gpurun_out/r01h_mesh.log
