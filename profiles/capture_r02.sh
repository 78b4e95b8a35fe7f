set -x
P="python profiles/prof_kernels.py"
N="ncu --set full --clock-control none --import-source on -f"
$P all --iters 1 > gpurun_out/r02_plain.log 2>&1 || exit 1
$N -k regex:lip_frame_kernel -c 1 -o gpurun_out/r02_lip_f $P lip --iters 1 > gpurun_out/r02_ncu1.log 2>&1
$N -k regex:logfbank_kernel -c 1 -o gpurun_out/r02_logfbank_f $P logfbank --iters 1 > gpurun_out/r02_ncu2.log 2>&1
$N -k regex:noise_ -c 2 -o gpurun_out/r02_noise_f $P noise --iters 1 > gpurun_out/r02_ncu3.log 2>&1
$N -k regex:fuse_ln_tma -c 2 -o gpurun_out/r02_fuselntma_f $P fuse_ln_tma --iters 1 > gpurun_out/r02_ncu4.log 2>&1
$N -k regex:pep_ -c 3 -o gpurun_out/r02_proj_f $P proj --iters 1 > gpurun_out/r02_ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_f.csv $P all --iters 1 > gpurun_out/r02_ncu6.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_f.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu7.log 2>&1
