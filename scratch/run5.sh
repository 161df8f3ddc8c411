python -m pytest tests -m gpu -x -q -k "not pass_by_pass and not chained_plan_explained and not seeded_stress" 2>&1 | tail -5
python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA_SYNC=4 PIVB200_SOA_GROUP=4 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA64=1 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA64=1 PIVB200_SOA_SYNC=4 PIVB200_SOA_GROUP=4 python tools/_sweep.py 16 2>&1 | tail -1
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
for cfg in "20 0 4" "20 4 4"; do set -- $cfg
  echo "== nwarps=$1 sync=$2 group=$3"
  PIVB200_NWARPS=$1 PIVB200_SOA_SYNC=$2 PIVB200_SOA_GROUP=$3 ncu --metrics $M --clock-control none -k regex:"piv_soa" -c 1 python tools/_prof.py 8 2>&1 | grep -E "no_instruction|_wait_|math_pipe|short_score|issue_active|duration|inst_executed" | awk '{print "   ", $1, $NF}'
done
