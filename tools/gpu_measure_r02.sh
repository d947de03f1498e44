#!/bin/bash
# Round-2 measurement pass on one B200 (run through gpurun). ncu reports are reduced to CSV on the box and deleted:
# gpurun_out/ is limited to 64 MiB.
mkdir -p gpurun_out
P=gpurun_out
python tools/profile_small.py > $P/r02_small_profile.json 2> $P/small.err
python tools/profile_small.py --bilinear > $P/r02_small_profile_bilinear.json 2>> $P/small.err
python tools/bench_aux.py > $P/r02_aux_kernels.jsonl 2> $P/aux.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r02_launches_ncu.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $P/ncu_launch.log 2>&1
FI_AUX_ITERS=1 FI_AUX_WARM=0 ncu --set full --clock-control none --import-source on \
  -k regex:"pack_pair|head_post|upsample2x|ssim_psnr|stem_mma" -c 12 -f -o /tmp/prof_aux_r02 python tools/bench_aux.py > $P/ncu_aux.log 2>&1
python tools/summarize_ncu.py /tmp/prof_aux_r02.ncu-rep > $P/r02_aux_ncu_full.csv 2>> $P/ncu_aux.log
ncu --set full --clock-control none -k regex:"bn_|wgrad_kernel" -c 48 -f -o /tmp/prof_train_r02 \
  python tools/bench_train.py --steps 1 --warmup 0 --skip-torch > $P/ncu_train.log 2>&1
python tools/summarize_ncu.py /tmp/prof_train_r02.ncu-rep > $P/r02_train_ncu_full.csv 2>> $P/ncu_train.log
# one forward at 1 pair, every kernel, --set full (tensor-pipe activity and DRAM bytes per launch)
ncu --set full --clock-control none -k regex:"conv_|stem_mma" -c 22 -f -o /tmp/prof_fwd_r02 \
  python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline > $P/ncu_fwd.log 2>&1
python tools/summarize_ncu.py /tmp/prof_fwd_r02.ncu-rep > $P/r02_kernels_ncu_full.csv 2>> $P/ncu_fwd.log
python bench.py --workload 4k_eval --steps 10 > $P/r02_4k_eval.json 2> $P/r02_4k_eval.err
python bench.py --workload train > $P/r02_train1.json 2> $P/r02_train1.err
python bench.py --workload api256 --no-cpu-baseline > $P/r02_api256.json 2> $P/r02_api256.err
for pp in 4 6 8; do python bench.py --pairs $pp --steps 20 --no-cpu-baseline --profile-out $P/r02_launch_profile_p$pp.json > $P/r02_bench_p$pp.json 2> $P/bench_p$pp.err; done
python -m pytest tests/test_gpu_train_step.py tests/test_gpu_wgrad.py -q 2>&1 | tail -15 > $P/r02_pytest_train.log
du -sh $P; ls -la $P | tail -30
