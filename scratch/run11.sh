python -m pytest tests -m gpu -x -q -k "not pass_by_pass and not chained_plan_explained and not seeded_stress" 2>&1 | tail -3
python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA_SYNC=0 python tools/_sweep.py 16 2>&1 | tail -1
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:"piv_soa" -c 2 python tools/_prof.py 8 2>&1 | grep -E "piv_soa_kernel|no_instruction|issue_active|duration|inst_executed" | awk '{print "   ", $1, $2, $NF}'
