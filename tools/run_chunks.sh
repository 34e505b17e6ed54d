mkdir -p gpurun_out
: > gpurun_out/r01i_chunks3.log
for cfg in "888 4" "888 5" "740 5" "592 6" "740 6" "444 8" "592 8"; do
set -- $cfg
echo -n "SG_BATCH=$1 SG_STREAMS=$2: " >> gpurun_out/r01i_chunks3.log
SG_BATCH=$1 SG_STREAMS=$2 timeout 300 python bench.py --no-cpu-baseline 2>>gpurun_out/r01i_chunks3.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['e2e']['value']), round(j['ms_per_step'],2))
except Exception as e: print('failed')" >> gpurun_out/r01i_chunks3.log
done
cat gpurun_out/r01i_chunks3.log; tail -2 gpurun_out/r01i_chunks3.err
python -c "import __graft_entry__ as g; g.smoke()"
