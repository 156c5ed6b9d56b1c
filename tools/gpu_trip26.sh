#!/bin/bash
# up-path fuse terms pre-summed on the low-resolution grids (HRNB_FUSE_TREE): tests + in-trip A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4))
except Exception as e: print('FAILED', e)"; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -q -x > $O/t26_pytest.txt 2>&1; echo "tests rc=$?"; tail -n 3 $O/t26_pytest.txt
for r in 1 2; do for b in 256 64; do for m in 1 0; do
echo -n "infer$b tree=$m: "; HRNB_FUSE_TREE=$m timeout 200 python bench.py --mode infer --batch $b --no-cpu-baseline 2>>$O/t26.err | brief
done; done; done
for m in 1 0; do echo -n "config3 tree=$m: "; HRNB_FUSE_TREE=$m timeout 400 python bench.py --config 3 --no-cpu-baseline 2>>$O/t26.err | brief; done
for m in 1 0; do echo -n "config1 tree=$m: "; HRNB_FUSE_TREE=$m timeout 400 python bench.py --config 1 --no-cpu-baseline 2>>$O/t26.err | brief; done
HRNB_FUSE_TREE=1 timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline --detail $O/t26_infer256_detail_tree.json >/dev/null 2>>$O/t26.err
tail -n 3 $O/t26.err
