run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no_parity > gpurun_out/r2o_$tag.json 2>/dev/null; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2o_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],4), round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms_rank0'].items() if k in ('error_loop','final_align')})
except Exception as e: print('$tag FAILED', e)
PY
}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run default A=1
run lowmem BADREAD_B200_LOWMEM=1
run w3 BADREAD_B200_SUBBATCHES=3
run w5 BADREAD_B200_SUBBATCHES=5
run nohead BADREAD_B200_HEAD_WORKER=0
run lane1024 BADREAD_B200_LANE8_COLS=1024
run lane2048 BADREAD_B200_LANE8_COLS=2048
run lane8192 BADREAD_B200_LANE8_COLS=8192
run mut12 BADREAD_B200_GRID_MUTATE=12
run mut16 BADREAD_B200_GRID_MUTATE=16
run warp4x3 BADREAD_B200_GRID_WARP4=3
run warp4x4 BADREAD_B200_GRID_WARP4=4
run warp21x5 BADREAD_B200_GRID_WARP2=5 BADREAD_B200_GRID_WARP1=5
run default2 A=1
