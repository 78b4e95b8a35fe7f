# per-launch device times (ncu gpu__time_duration.sum; cold-cache, serialised: shares, not absolutes)
set -x
K='regex:^(lip_|logmel|logfbank|noise_|fuse_|pep_|spec_|vfeats|gray_|tform|lm_fill|collate|pad_or_trim|peak_|similarity|warp_|cut_patch|video_feats)'
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r02_launches_bench_f.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu7.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r02_launches_f.csv python profiles/prof_kernels.py all --iters 1 > gpurun_out/r02_ncu6.log 2>&1
