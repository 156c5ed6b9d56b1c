#!/bin/bash
# dual MMA issuers: tests + in-trip A/B (HRNB_NO_DUAL=1 = one issuer)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4), 'launches', d['gpu_launches']//d['steps'])
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_train_kernels.py -m gpu -q -x > $O/t15_pytest.txt 2>&1; echo "tests rc=$?"; tail -3 $O/t15_pytest.txt
for b in 256 64; do
echo -n "infer$b dual: "; timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline --detail $O/t15_infer${b}_detail_dual.json 2>>$O/t15.err | brief
echo -n "infer$b single: "; HRNB_NO_DUAL=1 timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline --detail $O/t15_infer${b}_detail_single.json 2>>$O/t15.err | brief
echo -n "infer$b dual old-picker: "; HRNB_PICK_OLD=1 timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline --detail $O/t15_infer${b}_detail_dual_oldpick.json 2>>$O/t15.err | brief
echo -n "infer$b single old-picker: "; HRNB_PICK_OLD=1 HRNB_NO_DUAL=1 timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t15.err | brief
done
for r in 1 2; do
echo -n "train dual: "; timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t15.err | brief
echo -n "train single: "; HRNB_NO_DUAL=1 timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t15.err | brief
echo -n "train dual old-picker: "; HRNB_PICK_OLD=1 timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t15.err | brief
echo -n "train single old-picker: "; HRNB_PICK_OLD=1 HRNB_NO_DUAL=1 timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t15.err | brief
done

timeout 900 python -m pytest tests/test_gpu_train_network.py -m gpu -q -x > $O/t15_pytest_train.txt 2>&1; echo "train tests rc=$?"; tail -3 $O/t15_pytest_train.txt
tail -3 $O/t15.err
