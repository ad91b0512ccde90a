#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r8
timeout 900 python -m pytest tests/test_operator.py tests/test_gpu_parity_sizes.py -m gpu -x -q -k "z_split or baseline_sizes or five_sweep or vs_oracle_small" > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log
tail -6 ${O}_pytest.log
: > ${O}_time.log
for N in 160 144 128; do
  PCB200_DEBUG=1 timeout 300 python tools/time_apply.py $N bcc_dg pseudochiral_trivial 16 >> ${O}_time.log 2>&1
  PCB200_PLANE_COUPLED=0 timeout 300 python tools/time_apply.py $N bcc_dg pseudochiral_trivial 16 >> ${O}_time.log 2>&1
done
cat ${O}_time.log
