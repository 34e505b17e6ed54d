mkdir -p gpurun_out
: > gpurun_out/r01i_find_layouts.log
for N in 500000 50000; do
for lay in "4096 24" "8192 12" "16384 6" "4096 12" "2048 24" "8192 6"; do
set -- $lay
for kind in full v4; do
echo -n "N=$N sub=$1 warps=$2 kind=$kind: " >> gpurun_out/r01i_find_layouts.log
SG_SUBTILE=$1 SG_TILE_WARPS=$2 timeout 300 python tools/find_probe.py --refs $N --queries 4096 --reps 3 --kind $kind 2>&1 | tail -1 >> gpurun_out/r01i_find_layouts.log
done; done; done
cat gpurun_out/r01i_find_layouts.log
