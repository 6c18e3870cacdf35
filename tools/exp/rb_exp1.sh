mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_recurrent_bwd.py -x -q 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" $B > gpurun_out/s2_exp_$name.json 2> gpurun_out/s2_exp_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s2_exp_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'rb', d['kernels'].get('recurrent_bwd',{}).get('ms_per_step'), 'rf', d['kernels'].get('recurrent_fwd',{}).get('ms_per_step'))
except Exception as e: print('$name', 'FAILED', e)
PY
}
run default X=1
run att0 SSCVAE_RB_ATT_POLICY=0
run att2 SSCVAE_RB_ATT_POLICY=2
run w0 SSCVAE_RB_W_POLICY=0
run att2w0 SSCVAE_RB_ATT_POLICY=2 SSCVAE_RB_W_POLICY=0
run att0w0 SSCVAE_RB_ATT_POLICY=0 SSCVAE_RB_W_POLICY=0
run st3 SSCVAE_RB_STAGES=3
SSCVAE_RB_DBG=1 SSCVAE_NO_GRAPHS=1 python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-decode --no-gpu-eager --profile-steps 1 > gpurun_out/s2_dbg2.json 2> gpurun_out/s2_dbg2.err; grep rbdbg gpurun_out/s2_dbg2.err | head -5
