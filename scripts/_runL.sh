mkdir -p gpurun_out
for cfg in "32 2" "64 1" "128 2" "16 2"; do set -- $cfg
DJETS_STREAM_CTAS=$1 DJETS_EVAL_CTAS=$2 timeout 200 python scripts/bench_eval.py > gpurun_out/l_eval_s$1_e$2.jsonl 2>&1
echo "== stream $1 eval $2"; python - <<PY
import json
for l in open("gpurun_out/l_eval_s$1_e$2.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print(f"{d['op'][:44]:44s} {d['GBps']:8.0f}")
PY
done
DJETS_STREAM_CTAS=16 DJETS_EVAL_CTAS=2 timeout 200 python scripts/bench_eval.py 30 f32 > gpurun_out/l_eval_f32_s16_e2.jsonl 2>&1; echo "== f32 stream 16 eval 2"; cut -c1-120 gpurun_out/l_eval_f32_s16_e2.jsonl
