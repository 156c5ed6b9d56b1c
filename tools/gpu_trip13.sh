#!/bin/bash
# lean epilogue with phase outputs, fuse output 0 hosted by branch 0's last conv: tests + A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4), 'launches', d['gpu_launches']//d['steps'])
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_train_kernels.py tests/test_gpu_glue.py -m gpu -q > $O/t13_pytest.txt 2>&1; echo "tests rc=$?"; tail -8 $O/t13_pytest.txt
for b in 256 64; do
echo -n "infer$b host0=conv2: "; timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t13.err | brief
echo -n "infer$b host0=gather: "; HRNB_FUSE_HOST0=gather timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t13.err | brief
echo -n "infer$b fuse_sum kernels: "; HRNB_FUSE_EPILOGUE=0 timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t13.err | brief
done
timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline --detail $O/t13_infer256_detail.json >/dev/null 2>>$O/t13.err
echo -n "train: "; timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t13.err | brief
tail -5 $O/t13.err
