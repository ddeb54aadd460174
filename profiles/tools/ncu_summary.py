import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.per_cycle_active','smsp__average_warp_latency_per_inst_issued.ratio','l1tex__data_bank_conflicts_pipe_lsu.sum','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
idx=[hdr.index(w) for w in want if w in hdr]
out=[[hdr[i] for i in idx],[units[i] for i in idx]]+[[r[i] for i in idx] for r in rows[2:]]
if len(sys.argv)>2:
    csv.writer(open(sys.argv[2],'w')).writerows(out)
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:60])
    for i in idx[1:]:
        print(f"   {hdr[i][:75]:75s} {r[i]:>16s} {units[i]}")
