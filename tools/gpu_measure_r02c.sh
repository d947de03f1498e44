#!/bin/bash
# Final pass of round 2 on one B200: what the driver runs (tests, smoke, bench, reference arm) + the files profiles/README.md is built from.
mkdir -p gpurun_out
P=gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $P/final_pytest.log 2>&1
python __graft_entry__.py smoke > $P/final_smoke.log 2>&1
python bench.py --steps 20 --warmup 3 --profile-out $P/r02_launch_profile.json > $P/r02_bench.json 2> $P/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $P/r02_ref.json 2> $P/final_ref.err
python bench.py --workload api256 --no-cpu-baseline > $P/r02_api256.json 2>/dev/null
python bench.py --workload 4k_eval --steps 10 > $P/r02_4k_eval.json 2>/dev/null
python bench.py --workload train > $P/r02_train1.json 2>/dev/null
python tools/profile_small.py > $P/r02_small_profile.json 2>/dev/null
python tools/profile_small.py --bilinear > $P/r02_small_profile_bilinear.json 2>/dev/null
python tools/bench_aux.py > $P/r02_aux_kernels.jsonl 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r02_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $P/ncu_launch.log 2>&1
ncu --set full --clock-control none -k regex:"conv_|stem_mma|inc_fused" -c 21 -f -o /tmp/prof_fwd_r02 \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline > $P/ncu_fwd.log 2>&1
python tools/summarize_ncu.py /tmp/prof_fwd_r02.ncu-rep > $P/r02_kernels_ncu_full.csv 2>> $P/ncu_fwd.log
tail -4 $P/final_pytest.log; tail -1 $P/final_smoke.log
python -c "
import json
for f in ['r02_bench.json','r02_ref.json','r02_api256.json','r02_4k_eval.json','r02_train1.json']:
    t=open('$P/'+f).read().strip().splitlines(); d=json.loads(t[-1]); print(f, len(t), round(d['value'],3), round(d['e2e']['value'],3), d.get('gpu_launches'))
"
du -sh $P
