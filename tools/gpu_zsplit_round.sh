#!/bin/bash
# One GPU session: parity of the z-split plane mode, then its timings against the five-pass structure.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r4
timeout 900 python -m pytest tests/test_operator.py tests/test_gpu_parity_sizes.py -m gpu -x -q -k "z_split or baseline_sizes" > ${O}_pytest_zsplit.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_zsplit.log
tail -5 ${O}_pytest_zsplit.log
: > ${O}_time.log
for N in 160 144 128; do
  timeout 300 python tools/time_apply.py $N fcc chiral 16 >> ${O}_time.log 2>&1
  timeout 300 python tools/time_apply.py $N bcc_dg pseudochiral_crossdof 16 >> ${O}_time.log 2>&1
  PCB200_PLANE_CROSS=2 timeout 300 python tools/time_apply.py $N bcc_dg pseudochiral_crossdof 16 >> ${O}_time.log 2>&1
done
timeout 300 python tools/time_apply.py 160 bcc_dg pseudochiral_crossdof 32 >> ${O}_time.log 2>&1
cat ${O}_time.log
timeout 600 python tools/run_bandgap.py 160 bcc_dg pseudochiral_crossdof 20 2 > ${O}_band160.log 2>&1
tail -2 ${O}_band160.log
