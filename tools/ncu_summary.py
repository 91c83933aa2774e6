"""Summarise an .ncu-rep (raw page CSV) into the handful of metrics the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [extra-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
pat = re.compile(r"gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|sm__pipe_tensor.*cycles_active.*(pct|avg$)|sm__inst_executed_pipe_(xu|fma|alu|tensor|uniform|lsu|fmaheavy)[a-z_]*\.sum$|smsp__inst_executed\.sum$|sm__warps_active\.avg\.pct_of_peak|launch__registers_per_thread|launch__occupancy_limit|sm__cycles_elapsed\.avg$|sm__cycles_active\.avg$|smsp__issue_active\.avg\.pct|smsp__average_warps?_issue_stalled_[a-z_]+_per_warp_active\.pct|gpu__dram_throughput|lts__t_bytes\.sum$|sm__throughput\.avg\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|smsp__inst_executed_op_local|local_(load|store)|sm__sass_inst_executed_op_local|lts__t_sector_hit_rate\.pct|sm__pipe_xu_cycles_active|sm__inst_executed_pipe_xu|Kernel Name|Grid Size|Block Size|sm__clock|gpc__cycles_elapsed\.max|dram__cycles_active")
if extra: pat = re.compile(pat.pattern + "|" + extra)
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            try:
                fv = float(v.replace(",", ""))
                if "stalled" in h and fv < 1.0: continue
            except ValueError:
                pass
            print(f"{h} [{u}] = {v}")
