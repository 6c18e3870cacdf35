mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_gputests6.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s2_gputests6.log
B="python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" $B > gpurun_out/s2_exp_$name.json 2> gpurun_out/s2_exp_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s2_exp_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'rb', d['kernels'].get('recurrent_bwd',{}).get('ms_per_step'), 'rf', d['kernels'].get('recurrent_fwd',{}).get('ms_per_step'))
except Exception as e: print('$name', 'FAILED', e)
PY
}
run v10 X=1
SSCVAE_RB_DBG=1 SSCVAE_NO_GRAPHS=1 python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-decode --no-gpu-eager --profile-steps 1 > gpurun_out/s2_dbg11.json 2> gpurun_out/s2_dbg11.err; grep rbdbg gpurun_out/s2_dbg11.err | head -5
