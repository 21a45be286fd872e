"""profiles/ncu_<WL>.json for bench.py's roofline.ncu: what bounds the dominant kernel, from an `ncu --set full` report.
python tools/ncu_summary.py <report.ncu-rep> <WL> <out.json> [kernel-regex]"""
import csv
import io
import json
import re
import subprocess
import sys

rep, wl, out = sys.argv[1], sys.argv[2], sys.argv[3]
pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name")
cand = [r for r in rows[2:] if len(r) == len(hdr) and (pat is None or pat.search(r[kn]))]
vals = max(cand, key=lambda r: float(r[hdr.index("gpu__time_duration.sum")].replace(",", "") or 0))


def get(name, default=None):
    if name not in hdr:
        return default
    i = hdr.index(name)
    try:
        v = float(vals[i].replace(",", ""))
    except ValueError:
        return default
    u = units[i].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3,
             "us": 1e-3, "msecond": 1, "ms": 1, "second": 1e3, "s": 1e3}.get(u, 1)
    return v * scale


rd, wr = get("dram__bytes_read.sum", 0.0), get("dram__bytes_write.sum", 0.0)
d = {
    "workload": wl, "kernel": vals[kn], "source": f"ncu {'--set full' if 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio' in hdr else '(short metric list)'} --clock-control none, {rep.split('/')[-1]}",
    "gpu_time_ms_under_ncu": get("gpu__time_duration.sum"),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "lsu_wavefronts_pct": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "lanes_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "alu_pipe_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
    "lts_bytes": get("lts__t_bytes.sum"), "l1_bytes": get("l1tex__t_bytes.sum"),
    "warp_inst": get("smsp__inst_executed.sum"), "registers": get("launch__registers_per_thread"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "local_load_requests": get("smsp__inst_executed_op_local_ld.sum"), "local_store_requests": get("smsp__inst_executed_op_local_st.sum"),
    "shared_mem_per_block": get("launch__shared_mem_per_block_dynamic"), "grid": get("launch__grid_size"),
}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
