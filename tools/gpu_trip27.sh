#!/bin/bash
# resident multi-chunk weight slabs (head conv with BN = 160 x 3 N tiles, 64-channel 3x3 layers with two chunks): tests + in-trip A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4))
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_train_kernels.py -m gpu -q -x > $O/t27_pytest.txt 2>&1; echo "tests rc=$?"; tail -n 3 $O/t27_pytest.txt
for r in 1 2; do for b in 256 64; do for m in 0 1; do
echo -n "infer$b no_slab=$m: "; HRNB_NO_SLAB=$m timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline --detail $O/t27_infer${b}_noslab$m.json 2>>$O/t27.err | brief
done; done; done
for m in 0 1; do echo -n "config3 no_slab=$m: "; HRNB_NO_SLAB=$m timeout 400 python bench.py --config 3 --no-cpu-baseline 2>>$O/t27.err | brief; done
for r in 1 2; do for m in 0 1; do echo -n "train no_slab=$m: "; HRNB_NO_SLAB=$m timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t27.err | brief; done; done
timeout 900 python -m pytest tests/test_gpu_train_network.py -m gpu -q -x > $O/t27_pytest_train.txt 2>&1; echo "train tests rc=$?"; tail -n 3 $O/t27_pytest_train.txt
tail -n 3 $O/t27.err
