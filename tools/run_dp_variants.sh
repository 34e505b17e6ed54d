# DP kernel A/B runs on one GPU: isolated DP (one stream, whole-wave launches) and the chunk pipeline, per variant
for bp in 0 1 2 3; do
  echo "== SG_BANKPLAN=$bp"
  SG_BANKPLAN=$bp SG_STREAMS=1 SG_BATCH=1184 timeout 600 python tools/dp_probe.py --refs 20000 --queries 2368 --reps 3 2>&1 | tail -1
  SG_BANKPLAN=$bp timeout 600 python tools/dp_probe.py --refs 20000 --queries 3552 --reps 3 2>&1 | tail -1
done
