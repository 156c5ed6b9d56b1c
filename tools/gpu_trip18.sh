#!/bin/bash
# training with the 8-epilogue-warp build (two CTAs per SM, optionally of different kernels)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
brief() { python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d.get('tensor_frac_of_burst_peak',0),4), 'launches', d['gpu_launches']//d['steps'])
except Exception as e: print('FAILED', e)"; }
echo -n "train default: "; timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t18.err | brief
echo -n "train epi8: "; HRNB_STATS_MAX=32 HRNB_LIB=libhrnb_epi8.so timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t18.err | brief
echo -n "train epi8 share: "; HRNB_TMEM_SHARE=1 HRNB_STATS_MAX=32 HRNB_LIB=libhrnb_epi8.so timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t18.err | brief
echo -n "train default stats<=32: "; HRNB_STATS_MAX=32 timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t18.err | brief
echo -n "infer64 epi8 share: "; HRNB_TMEM_SHARE=1 HRNB_LIB=libhrnb_epi8.so timeout 200 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t18.err | brief
echo -n "infer64 default: "; timeout 200 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t18.err | brief
tail -3 $O/t18.err
