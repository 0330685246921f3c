M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_active.avg,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
for lib in "" rusty_marcher_b200/librm_b200_privq.so; do
  echo "== lib: $lib"
  RM_B200_LIB=$lib ncu --metrics $M --clock-control none --kernel-name regex:render_fast -c 3 --csv python tools/run_phases.py cornell_4k 4 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
H=rows[0]; i_n=H.index('Metric Name'); i_v=H.index('Metric Value'); i_id=H.index('ID')
last=max(int(r[i_id]) for r in rows[1:])
for r in rows[1:]:
    if int(r[i_id])==last: print('  %-90s %s' % (r[i_n], r[i_v]))
"
done
