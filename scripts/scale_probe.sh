# node-shard scaling probe (run under gpurun --gpus 8): bench.py at the world sizes given as arguments (default 8)
export CCSIM_NO_REBUILD=1
for n in ${@:-8}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-objects > gpurun_out/r2_bench_c4_n$n.json 2> gpurun_out/r2_bench_c4_n$n.err; tail -2 gpurun_out/r2_bench_c4_n$n.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_c4_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","placements_per_sec","parity")}, (d.get("roofline") or {}).get("latency"))
    except Exception as e: print(f, "ERR", e)
PY
