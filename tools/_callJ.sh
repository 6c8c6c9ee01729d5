mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2zb_tests.log
timeout 900 bash tools/profile_step.sh r2zb > gpurun_out/r2zb_profile.log 2>&1
timeout 400 python bench.py > gpurun_out/r2zb_bench_config1.json 2> gpurun_out/r2zb_bench_config1.err
BADREAD_B200_SUBBATCHES=3 BADREAD_B200_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 2 --no_parity > gpurun_out/r2zb_trace_w3.json 2> gpurun_out/r2zb_trace_w3.err; echo "trace w3 rc=$?" >> gpurun_out/r2zb_tests.log
BADREAD_B200_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 2 --no_parity > /dev/null 2>&1; echo "trace w4 rc=$?" >> gpurun_out/r2zb_tests.log; cp gpurun_out/trace_config1_rank0.csv gpurun_out/r2zb_trace_c1.csv
timeout 500 python bench.py --config 2 --steps 2 --warmup 2 > gpurun_out/r2zb_bench_config2.json 2> gpurun_out/r2zb_bench_config2.err
timeout 500 python bench.py --config 3 --steps 2 --warmup 2 > gpurun_out/r2zb_bench_config3.json 2> gpurun_out/r2zb_bench_config3.err
cat gpurun_out/r2zb_tests.log; for c in 1 2 3; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2zb_bench_config$c.json').read().strip().splitlines()[-1]); print($c, d['value'], d['e2e']['value'], d['ms_per_step'], d['parity'], d['cpu_baseline']['value'])
except Exception as e: print($c, 'FAILED', e)
PY
done
