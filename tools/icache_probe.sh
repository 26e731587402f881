#!/bin/bash
# Tuning aid: instruction-cache behaviour and stall mix of a kernel in steady state (8K image).
# usage (on the GPU box): bash tools/icache_probe.sh [out.csv] [kernel-name regex, default k_fused_blocks]
out=${1:-gpurun_out/icache_probe.csv}
kernel=${2:-k_fused_blocks}
python tools/k1_only.py > /dev/null 2>&1 || exit 1
ncu --clock-control none --kernel-name "regex:$kernel" --launch-skip 3 --launch-count 1 \
    --metrics smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum \
    --csv --log-file "$out" python tools/k1_only.py > /dev/null 2>&1
python - "$out" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
for r in rows[1:]:
    print(f"{r[h.index('Metric Name')]:90s} {r[h.index('Metric Value')]}")
PY
