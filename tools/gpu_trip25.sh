#!/bin/bash
# two CTAs per SM (8 epilogue warps) for thin resident-weight layers, per launch: tests + in-trip A/B (HRNB_TWIN_MIN: tiles per CTA slot)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4))
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_train_kernels.py -m gpu -q -x > $O/t25_pytest.txt 2>&1; echo "tests rc=$?"; tail -n 3 $O/t25_pytest.txt
for r in 1 2; do for b in 256 64; do for m in 1000000 4 2 1; do
echo -n "infer$b twin_min=$m: "; HRNB_TWIN_MIN=$m timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t25.err | brief
done; done; done
for m in 1000000 4 1; do echo -n "config3 twin_min=$m: "; HRNB_TWIN_MIN=$m timeout 400 python bench.py --config 3 --no-cpu-baseline 2>>$O/t25.err | brief; done
for m in 1000000 4 1; do echo -n "train twin_min=$m: "; HRNB_TWIN_MIN=$m timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t25.err | brief; done
HRNB_TWIN_MIN=4 timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline --detail $O/t25_infer256_detail_twin.json >/dev/null 2>>$O/t25.err
tail -n 3 $O/t25.err
