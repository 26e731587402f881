#!/bin/bash
# In-kernel timelines and size sweeps of the shipped code (run under gpurun; needs `make -C .../csrc trace`):
#   gpurun -- 'bash tools/make_timelines.sh'   ->  gpurun_out/r02_timelines.txt  (copied to profiles/ by hand)
cd "$(dirname "$0")/.."
O=gpurun_out/r02_timelines.txt
{
  echo "# In-kernel timelines (globaltimer), tracing build of the library, B200 (tools/make_timelines.sh)"
  echo; echo "## K1 (tools/k1_trace.py), 3840x2160"
  timeout 200 python tools/k1_trace.py 3840 2160 2>&1 | grep -v "^\s*$"
  echo; echo "## K1, 7680x4320"
  timeout 200 python tools/k1_trace.py 7680 4320 2>&1 | grep -v "^\s*$"
  echo; echo "## K1, 15360x8640 (steady state: phases of every warp's 4th tile)"
  timeout 200 python tools/k1_trace.py 15360 8640 2>&1 | sed -n '/4th tile/,$p'
  echo; echo "## K2 (tools/k2_trace.py), 3840x2160, per tile of 16 strips"
  timeout 200 python tools/k2_trace.py 3840 2160 2>&1 | grep -v "unused"
  echo; echo "## Event time per kernel vs image size (tools/round_scaling.py, one stream)"
  timeout 300 python tools/round_scaling.py 2>&1
  echo; echo "## Coefficients re-evaluated in reference order (tools/flag_rate.py), tensor-core transform"
  timeout 200 python tools/flag_rate.py 2>&1
} > $O
tail -40 $O
