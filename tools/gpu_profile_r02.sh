#!/bin/bash
# Round-2 profile captures on ONE GPU (run under gpurun).  Outputs land in gpurun_out/ and are summarised into profiles/ here.
#  1. launch list of one eager step of the bench command (times + DRAM bytes per launch; cold-cache, serialised: use SHARES)
#  2. `ncu --set full` of the dominant kernels at their C2 shapes (tools/gpu_one_kernel.py)
set -x
mkdir -p gpurun_out
[ -n "$SKIP_LAUNCHES" ] || timeout 900 ncu --nvtx --nvtx-include "vs_timed_eager/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-library-baseline --no-inversion > gpurun_out/r02_launches_bench.log 2>&1
[ -n "$SKIP_LAUNCHES" ] || tail -2 gpurun_out/r02_launches_bench.log
# -k: only OUR kernel (gpu_one_kernel.py also launches torch's randn / copy kernels while it builds the inputs -- the first
# round-2 capture of proj/qkv/conv/gn without the filter recorded those instead); -s: skip the first (cold) call
for k in ${KERNELS:-attn proj qkv conv gn}; do
  case $k in attn) filt="regex:.*attn_tcp_kernel.*"; skip=1; cnt=1;; gn) filt="regex:.*gn_(stats|apply)_kernel.*"; skip=2; cnt=2;; *) filt="regex:.*gemm_tc_kernel.*"; skip=1; cnt=1;; esac
  timeout 300 ncu --set full --clock-control none --import-source on -k "$filt" -c $cnt -s $skip -o gpurun_out/r02_ncu_$k -f python tools/gpu_one_kernel.py $k 3 > gpurun_out/r02_ncu_$k.log 2>&1
  tail -1 gpurun_out/r02_ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
