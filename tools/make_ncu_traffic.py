"""profiles/ncu_traffic.json from the raw page of an `ncu --set full` capture of one window-kernel launch (what bench.py's roofline.traffic and
roofline.issue are read from):   python tools/make_ncu_traffic.py raw.csv windows "source text" [head commit] > profiles/ncu_traffic.json
   raw.csv = `ncu -i capture.ncu-rep --page raw --csv`; windows = windows the captured launch processed (tools/ncu_target.py prints it)"""
import csv, json, sys
raw, nwin, src = sys.argv[1], float(sys.argv[2]), sys.argv[3]
head = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name); v = float(val[i].replace(",", "")); u = unit[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0) if "byte" in u else v
dr, dw = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
inst = get("smsp__inst_executed.sum")
out = {"dram_bytes_per_window": int(round((dr + dw) / nwin)), "dram_read_bytes_per_window": int(round(dr / nwin)), "dram_write_bytes_per_window": int(round(dw / nwin)),
       "windows": int(nwin), "kernel": rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else None, "source": src, "head": head,
       "issue": {"warp_instructions_per_window": int(round(inst / nwin)), "lanes_per_instruction": round(get("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
                 "issue_active_pct": round(get("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2), "ipc_per_scheduler": round(get("smsp__issue_active.avg.per_cycle_active"), 3),
                 "warps_active_pct": round(get("sm__warps_active.avg.pct_of_peak_sustained_active"), 2), "l1_hit_pct": round(get("l1tex__t_sector_hit_rate.pct"), 1), "l2_hit_pct": round(get("lts__t_sector_hit_rate.pct"), 1),
                 "stall_long_scoreboard_per_issue": round(get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"), 2),
                 "stall_barrier_per_issue": round(get("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"), 2),
                 "note": "instruction-issue view of the same launch: the path is integer / latency bound (SURVEY 8d), peak issue = 1 warp instruction per scheduler per cycle"}}
print(json.dumps(out, indent=1))
