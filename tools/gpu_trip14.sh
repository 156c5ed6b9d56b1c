#!/bin/bash
# in-trip A/B: library of commit fabf69c (libhrnb_base.so) against the current build
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4), 'launches', d['gpu_launches']//d['steps'])
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_train_kernels.py -m gpu -q > $O/t14_pytest.txt 2>&1; echo "tests rc=$?"; tail -3 $O/t14_pytest.txt
for r in 1 2; do
echo -n "train base: "; HRNB_LIB=libhrnb_base.so timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t14.err | brief
echo -n "train new: "; timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t14.err | brief
echo -n "train new no-wres: "; HRNB_NO_WRES=1 timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t14.err | brief
done
for b in 256 64; do
echo -n "infer$b base gather: "; HRNB_LIB=libhrnb_base.so HRNB_FUSE_HOST0=gather timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t14.err | brief
echo -n "infer$b new gather: "; HRNB_FUSE_HOST0=gather timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t14.err | brief
echo -n "infer$b new gather no-wres: "; HRNB_NO_WRES=1 HRNB_FUSE_HOST0=gather timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t14.err | brief
echo -n "infer$b new conv2: "; timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t14.err | brief
echo -n "infer$b new conv2 no-wres: "; HRNB_NO_WRES=1 timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t14.err | brief
done
tail -3 $O/t14.err
