export CCSIM_NO_REBUILD=1
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -6
for n in 8 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_n$n.json 2> gpurun_out/r2_bench_c4_n$n.err; tail -2 gpurun_out/r2_bench_c4_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --workload c5 --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/r2_bench_c5_n8.err; tail -2 gpurun_out/r2_bench_c5_n8.err
python - <<PY
import json
for f in ("r2_bench_c4_n8","r2_bench_c4_n2","r2_bench_c5_n8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","placements_per_sec","parity")}, (d.get("roofline") or {}).get("latency"))
    except Exception as e: print(f, "ERR", e)
PY
