O=gpurun_out/last; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; echo "bench rc=$?"
python bench.py --workload C5 --steps 2 > $O/bench_C5_n1.json 2> $O/bench_C5_n1.err; echo "C5 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/last/bench_c4_n1.json","gpurun_out/last/bench_C5_n1.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],2), round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "traffic", d["roofline"]["traffic"], d["frame_md5"], d.get("cpu_baseline",{}).get("value"))
PY
