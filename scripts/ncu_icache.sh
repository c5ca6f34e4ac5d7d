#!/bin/bash
# instruction-supply metrics of the solve kernel for several builds: scripts/ncu_icache.sh "<libA> <libB> ..."
D=$PWD/dyobav_mpcnwta_warehouse_b200/csrc
M=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,launch__registers_per_thread
for L in $1; do
  echo "== $L"
  MPCB_LIB_PATH=$D/$L ncu --metrics $M --clock-control none -k regex:solve_kernel -c 1 python scripts/profile_step.py 1184 8 200 4 2>/dev/null | grep -E "hit_rate|gcc__|no_instruction|stalled_wait|short_score|inst_executed|duration|registers"
done
