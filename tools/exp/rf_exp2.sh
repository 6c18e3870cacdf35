mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_recurrent_bwd.py tests/test_gpu_train.py -x -q 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" $B > gpurun_out/s2_exp_$name.json 2> gpurun_out/s2_exp_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s2_exp_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'rb', d['kernels'].get('recurrent_bwd',{}).get('ms_per_step'), 'rf', d['kernels'].get('recurrent_fwd',{}).get('ms_per_step'))
except Exception as e: print('$name', 'FAILED', e)
PY
}
run g1 X=1
run g1f6 SSCVAE_RF_STAGES=6
run g1b5 SSCVAE_RB_STAGES=5
