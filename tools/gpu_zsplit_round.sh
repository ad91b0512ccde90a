#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r9
: > ${O}_time.log
for N in 160 144 128 96 64; do
  timeout 300 python tools/time_apply.py $N fcc chiral 16 >> ${O}_time.log 2>&1
done
timeout 300 python tools/time_apply.py 160 bcc_dg pseudochiral_crossdof 16 >> ${O}_time.log 2>&1
timeout 300 python tools/time_apply.py 120 fcc chiral 16 >> ${O}_time.log 2>&1
cat ${O}_time.log
timeout 900 python -m pytest tests/test_operator.py tests/test_gpu_parity_sizes.py -m gpu -x -q -k "z_split or baseline_sizes or five_sweep or vs_oracle_small or crossdof_pass" > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log
tail -4 ${O}_pytest.log
