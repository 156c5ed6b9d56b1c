#!/bin/bash
# resident weights + two CTAs per SM (8 epilogue warps build) A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4))
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_train_kernels.py -m gpu -q -x > $O/t12_pytest_default.txt 2>&1; echo "default lib tests rc=$?"; tail -3 $O/t12_pytest_default.txt
HRNB_LIB=libhrnb_epi8.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -q > $O/t12_pytest_epi8.txt 2>&1; echo "epi8 lib tests rc=$?"; tail -3 $O/t12_pytest_epi8.txt
for b in 256 64; do
echo -n "infer$b default: "; timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t12.err | brief
echo -n "infer$b epi8: "; HRNB_LIB=libhrnb_epi8.so timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t12.err | brief
echo -n "infer$b epi8 nopdl: "; HRNB_PDL=0 HRNB_LIB=libhrnb_epi8.so timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t12.err | brief
echo -n "infer$b epi8 MB=2: "; HRNB_MB=2 HRNB_LIB=libhrnb_epi8.so timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t12.err | brief
done
HRNB_LIB=libhrnb_epi8.so timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline --detail $O/t12_infer256_detail_epi8.json >/dev/null 2>>$O/t12.err
timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline --detail $O/t12_infer256_detail_default.json >/dev/null 2>>$O/t12.err
echo -n "train default: "; timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t12.err | brief
tail -5 $O/t12.err
