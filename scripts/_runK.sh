set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_dbarray.py -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -4 gpurun_out/k_pytest.log
timeout 200 python scripts/bench_eval.py > gpurun_out/k_eval_default.jsonl 2>&1
DJETS_STREAM_CTAS=16 timeout 200 python scripts/bench_eval.py > gpurun_out/k_eval_stream16.jsonl 2>&1
DJETS_STREAM_CTAS=4 DJETS_EVAL_CTAS=2 timeout 200 python scripts/bench_eval.py > gpurun_out/k_eval_s4_e2.jsonl 2>&1
DJETS_STREAM_CTAS=1 DJETS_EVAL_CTAS=2 DJETS_EVAL_RING_KB=64 timeout 200 python scripts/bench_eval.py > gpurun_out/k_eval_s1_e2_r64.jsonl 2>&1
DJETS_STREAM_CTAS=3 DJETS_EVAL_CTAS=3 DJETS_EVAL_RING_KB=48 timeout 200 python scripts/bench_eval.py > gpurun_out/k_eval_s3_e3_r48.jsonl 2>&1
for f in default stream16 s4_e2 s1_e2_r64 s3_e3_r48; do echo "== $f"; python - <<PY
import json
for l in open("gpurun_out/k_eval_$f.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print(f"{d['op'][:44]:44s} {d['GBps']:8.0f}")
PY
done
