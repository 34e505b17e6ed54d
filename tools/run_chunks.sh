mkdir -p gpurun_out
: > gpurun_out/r01i_chunks2.log
for cfg in "888 3" "888 4" "740 3" "740 4" "592 4" "1184 3" "1184 4" "444 4" "1036 3"; do
set -- $cfg
echo -n "SG_BATCH=$1 SG_STREAMS=$2: " >> gpurun_out/r01i_chunks2.log
SG_BATCH=$1 SG_STREAMS=$2 timeout 300 python bench.py --no-cpu-baseline 2>>gpurun_out/r01i_chunks2.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['e2e']['value']), round(j['ms_per_step'],2))
except Exception as e: print('failed')" >> gpurun_out/r01i_chunks2.log
done
cat gpurun_out/r01i_chunks2.log; tail -3 gpurun_out/r01i_chunks2.err
