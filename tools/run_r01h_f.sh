mkdir -p gpurun_out
bash tools/run_quick.sh r01h_f
SG_BANKPLAN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r01h_f_bench_nobankplan.json 2>/dev/null; echo "nobankplan rc=$?"
python - <<'P'
import json
for f in ("gpurun_out/r01h_f_bench.json","gpurun_out/r01h_f_bench_nobankplan.json"):
    try:
        j=json.load(open(f)); print(f, round(j["value"]), round(j["e2e"]["value"]), j["roofline"]["gcups"], j["stages_ms_per_step_isolated"])
    except Exception as e: print(f, e)
P
