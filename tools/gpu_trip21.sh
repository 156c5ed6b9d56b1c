#!/bin/bash
# in-trip A/B of library builds (inference): current / before the grouped launch; then training + kernel tests on the current build
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3))
except Exception as e: print('FAILED', e)"; }
for r in 1 2; do for b in 256 64; do for l in libhrnb.so libhrnb_pre.so; do
echo -n "infer$b $l: "; HRNB_LIB_LAX=1 HRNB_LIB=$l timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t21.err | brief
done; done; done
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -q -x > $O/t21_pytest.txt 2>&1; echo "tests rc=$?"; tail -2 $O/t21_pytest.txt
for r in 1 2; do echo -n "train: "; timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t21.err | brief; echo -n "train s2 split: "; HRNB_S2_SPLIT=1 timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t21.err | brief; done
tail -3 $O/t21.err
