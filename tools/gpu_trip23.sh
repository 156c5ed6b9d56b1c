#!/bin/bash
# measured tile shapes (HRNB_AUTOTUNE=1) against the cycle model, inference
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4))
except Exception as e: print('FAILED', e)"; }
for r in 1 2; do
for a in 0 1; do
echo -n "config3 autotune=$a: "; HRNB_AUTOTUNE=$a timeout 400 python bench.py --config 3 --no-cpu-baseline 2>>$O/t23.err | brief
echo -n "infer256 autotune=$a: "; HRNB_AUTOTUNE=$a timeout 400 python bench.py --mode infer --batch 256 --no-cpu-baseline 2>>$O/t23.err | brief
echo -n "infer64 autotune=$a: "; HRNB_AUTOTUNE=$a timeout 400 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t23.err | brief
done; done
HRNB_AUTOTUNE=0 timeout 400 python bench.py --config 3 --no-cpu-baseline --detail $O/t23_config3_detail.json > /dev/null 2>>$O/t23.err
tail -3 $O/t23.err
