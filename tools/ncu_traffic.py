"""profiles/ncu_<WL>.json for bench.py's roofline.traffic: DRAM bytes of the dominant kernel from an `ncu --set full` report."""
import csv, io, json, subprocess, sys
rep, wl, out = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(',', ''))
    u = units[i].lower()
    scale = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)
    return v * scale
rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
d = {"workload": wl, "kernel": vals[hdr.index('Kernel Name')], "dram_bytes_read": rd, "dram_bytes_write": wr,
     "dram_bytes_per_launch": rd + wr, "gpu_time_ms_under_ncu": get('gpu__time_duration.sum') / 1e6 if units[hdr.index('gpu__time_duration.sum')] in ('ns', 'nsecond') else None,
     "source": f"ncu --set full --clock-control none, {rep.split('/')[-1]}"}
json.dump(d, open(out, 'w'), indent=1)
print(d)
