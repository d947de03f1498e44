#!/bin/bash
# Development loop of conv_rows.cu: layer parity (rows variant), network parity, then an A/B against the halo kernels.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_layers.py -q -x -k "rows" 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_unet.py -q -x -k "kernel_selection or golden or parity or bilinear" 2>&1 | tail -4
for v in 1 0; do
  FI_ROWS=$v timeout 200 python bench.py --steps ${AB_STEPS:-40} --no-cpu-baseline --profile-out gpurun_out/prof_rows$v.json 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rows=$v', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
python -c "
import json
for v in (1,0):
    p=json.load(open('gpurun_out/prof_rows%d.json'%v)); print(v, [(r['name'][-12:], round(r['ms_total']/r['calls'],3)) for r in p if 'up4.conv' in r['name']])"
