mkdir -p gpurun_out
for pr in 0 1 2; do
SSCVAE_RB_PROBE=$pr SSCVAE_RB_DBG=1 SSCVAE_NO_GRAPHS=1 python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-decode --no-gpu-eager --profile-steps 1 > gpurun_out/s2_probe$pr.json 2> gpurun_out/s2_probe$pr.err; echo "probe $pr"; grep rbdbg gpurun_out/s2_probe$pr.err | sed -n '1p;3p' | cut -c1-700
done
