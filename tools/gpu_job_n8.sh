# C4 on 8 GPUs (bench line into gpurun_out/n8/)
mkdir -p gpurun_out/n8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --workload ${WL:-C4} > gpurun_out/n8/final_${WL:-C4}.json 2> gpurun_out/n8/final_${WL:-C4}.err
python - <<PY
import json
try:
    d=json.loads([x for x in open("gpurun_out/n8/final_${WL:-C4}.json") if x.startswith("{")][-1]); print("${WL:-C4}", d["n_gpus"], "ms", round(d["ms_per_step"],3), [round(x,2) for x in d["step_ms"]], "Mrays/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["frame_md5"])
except Exception as e: print("ERR", e)
PY
