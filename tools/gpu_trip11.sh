#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_network.py -m gpu -q -k trajectory > $O/t11_pytest.txt 2>&1; echo "traj rc=$?"; grep -E "Error|assert|rel\[" $O/t11_pytest.txt | head -20
grep trajectory $O/parity_report.jsonl | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/t11_launches_infer.csv python tools/run_once.py 256 2 > $O/t11_ncu_launches_infer.log 2>&1; echo "ncu infer list rc=$?"; tail -5 $O/t11_ncu_launches_infer.log; tail -3 $O/t11_launches_infer.csv | cut -c1-300
