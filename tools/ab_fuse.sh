#!/bin/bash
# A/B of the fused inc kernel against the two-launch schedule (FI_FUSE_INC=0), alternating to average out clock drift.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -x -k "fused_inc or other_channel or batch_sizes" 2>&1 | tail -2
for i in 1 2; do for v in 1 0; do
  FI_FUSE_INC=$v python bench.py --steps ${AB_STEPS:-60} --no-cpu-baseline --profile-out gpurun_out/prof_fuse$v.json 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fuse=$v', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done; done
python -c "
import json
p=json.load(open('gpurun_out/prof_fuse1.json')); print([(r['name'][-14:], round(r['ms_total']/r['calls'],3)) for r in p[:2]])"
