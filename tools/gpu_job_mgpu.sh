# usage: bash tools/gpu_job_mgpu.sh N  — bench C4 and C5 on N GPUs (torchrun), JSON lines into gpurun_out/mgpu/
N=$1
mkdir -p gpurun_out/mgpu
for wl in C4 C5; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 3 --warmup 3 --workload $wl --no-secondary > gpurun_out/mgpu/bench_${wl}_n${N}.json 2> gpurun_out/mgpu/bench_${wl}_n${N}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --workload $wl > gpurun_out/mgpu/bench_${wl}_n${N}.json 2> gpurun_out/mgpu/bench_${wl}_n${N}.err
  fi
  echo "$wl N=$N rc=$?"; tail -c 300 gpurun_out/mgpu/bench_${wl}_n${N}.err
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/mgpu/bench_${wl}_n${N}.json") if x.startswith("{")][-1]
    d=json.loads(l); print("${wl}", d["n_gpus"], "ms", round(d["ms_per_step"],2), "Mrays/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],1))
except Exception as e: print("ERR", e)
PY
done
