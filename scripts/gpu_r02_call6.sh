#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python scripts/diag_fused.py 300 50 3000 2 > $O/${TAG}_diag_fused.log 2>&1; cat $O/${TAG}_diag_fused.log
timeout 200 python scripts/diag_fused.py 300 32 3000 1 >> $O/${TAG}_diag_fused.log 2>&1; tail -14 $O/${TAG}_diag_fused.log
timeout 200 python scripts/diag_fused.py 300 40 3000 1 >> $O/${TAG}_diag_fused.log 2>&1; tail -14 $O/${TAG}_diag_fused.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python scripts/diag_fused.py 260 40 400 1 > $O/${TAG}_racecheck.log 2>&1; tail -30 $O/${TAG}_racecheck.log
timeout 400 compute-sanitizer --tool memcheck python scripts/diag_fused.py 260 40 400 1 > $O/${TAG}_memcheck.log 2>&1; tail -15 $O/${TAG}_memcheck.log
