#!/bin/bash
# One GPU session: timings after the occupancy fix of the inverse x pass, then the whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r5
: > ${O}_time.log
for N in 160 144 128; do
  timeout 300 python tools/time_apply.py $N fcc chiral 16 >> ${O}_time.log 2>&1
done
timeout 300 python tools/time_apply.py 160 bcc_dg pseudochiral_crossdof 16 >> ${O}_time.log 2>&1
timeout 300 python tools/time_apply.py 160 bcc_dg pseudochiral_trivial 16 >> ${O}_time.log 2>&1
timeout 300 python tools/time_apply.py 256 sc_curv chiral 8 >> ${O}_time.log 2>&1
timeout 300 python tools/time_apply.py 192 sc_curv chiral 8 >> ${O}_time.log 2>&1
timeout 300 python tools/time_apply.py 120 fcc chiral 16 >> ${O}_time.log 2>&1
cat ${O}_time.log
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_gpu.log
tail -4 ${O}_pytest_gpu.log
