"""Summarise an .ncu-rep: headline metrics + per-source-line warp-instruction share and lane efficiency."""
import csv, io, subprocess, sys, json
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, vals = rows[0], rows[2]
keep = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
summary = {}
for h, u, v in zip(hdr, rows[1], vals):
    if h in keep:
        summary[h] = f"{v} {u}".strip()
        print(f"{h:90s} {v} {u}")
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 10 and r[0]:
        try:
            ln = int(r[0]); ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); sm = hdr.index('# Samples')
            agg.append((cur, ln, r[1].strip()[:80], int(r[ie]), int(r[te]), int(r[sm])))
        except Exception:
            pass
ti = sum(a[3] for a in agg); tt = sum(a[4] for a in agg); ts = sum(a[5] for a in agg)
print(f"total warp-inst {ti:.3e}  lane efficiency {tt/ti/32:.3f}")
lines = []
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    l = f"{a[0]}:{a[1]:4d} inst {a[3]/ti*100:5.2f}% lanes {a[4]/max(a[3],1):5.1f} stall-samples {a[5]/ts*100:5.2f}% | {a[2]}"
    print(l); lines.append(l)
if len(sys.argv) > 3:
    json.dump({"report": rep, "metrics": summary, "lane_efficiency": tt/ti/32, "warp_inst": ti, "top_lines": lines}, open(sys.argv[3], 'w'), indent=1)
