import csv,sys,subprocess
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit','launch__waves_per_multiprocessor','smsp__warp_issue_stalled','sm__pipe_tensor','launch__grid_size','launch__block_size','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warp']
for vals in rows[2:]:
    print("----", vals[hdr.index('Kernel Name')][:70])
    for h,u,v in zip(hdr,units,vals):
        if any(p in h for p in pats): print(f"  {h} [{u}] = {v}")
