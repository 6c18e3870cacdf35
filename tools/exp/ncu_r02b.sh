mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-decode --no-gpu-eager --profile-steps 1"
$CMD > gpurun_out/r02b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 330 --csv --log-file gpurun_out/launches_r02b.csv $CMD > gpurun_out/r02b_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"recurrent_bwd_kernel|recurrent_fwd_kernel" -s 8 -c 2 -o gpurun_out/r02b_full $CMD > gpurun_out/r02b_ncu2.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/ | tail -5
