mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2t_tests.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no_parity > gpurun_out/r2t_$tag.json 2>/dev/null; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2t_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],4), round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms_rank0'].items() if k in ('error_loop','final_align')})
except Exception as e: print('$tag FAILED', e)
PY
}
run ring4 A=1 >> gpurun_out/r2t_sweep.log
run ring8 BADREAD_B200_RING_T=8 >> gpurun_out/r2t_sweep.log
run ring2 BADREAD_B200_RING_T=2 >> gpurun_out/r2t_sweep.log
run cb4 BADREAD_B200_CB_NARROW=4 >> gpurun_out/r2t_sweep.log
run ring4b A=1 >> gpurun_out/r2t_sweep.log
timeout 600 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err
cat gpurun_out/r2t_tests.log gpurun_out/r2t_sweep.log; tail -c 1500 gpurun_out/r2t_bench_default.json
