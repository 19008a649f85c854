#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python scripts/diag_fused.py 120 21 > $O/${TAG}_diag_fused.log 2>&1; cat $O/${TAG}_diag_fused.log
timeout 120 python scripts/diag_fused.py 300 50 >> $O/${TAG}_diag_fused.log 2>&1; tail -12 $O/${TAG}_diag_fused.log
( timeout 600 python -m pytest tests/test_gpu_local_group.py tests/test_gpu_upload.py tests/test_gpu_list.py tests/test_gpu_parity.py -m gpu -q -p timeout --timeout 150 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -30 $O/${TAG}_pytest_gpu.log
