# C4 (and optionally C5) on 8 GPUs, with the chunk-size experiment
mkdir -p gpurun_out/n8
run() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --workload ${WL:-C4} > gpurun_out/n8/$name.json 2> gpurun_out/n8/$name.err
  python - <<PY
import json
try:
    d=json.loads([x for x in open("gpurun_out/n8/$name.json") if x.startswith("{")][-1]); print("$name", d["n_gpus"], "ms", round(d["ms_per_step"],3), [round(x,2) for x in d["step_ms"]], "Mrays/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("$name ERR", e)
PY
}
run c4_chunk8 RL_OW_CHUNK=8
run c4_chunk4 RL_OW_CHUNK=4
WL=C5 run c5_chunk8 RL_OW_CHUNK=8
