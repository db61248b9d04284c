# 2-GPU evidence (run under gpurun --gpus 2): the process-per-GPU sharded parity tests (CUDA IPC) and the node-sharded bench lines
export CCSIM_NO_REBUILD=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_n2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-objects > gpurun_out/r2_bench_c4_n2.json 2> gpurun_out/r2_bench_c4_n2.err; echo "c4 n2 rc=$?"; tail -2 gpurun_out/r2_bench_c4_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --steps 3 --warmup 3 --no-objects > gpurun_out/r2_bench_c5_n2.json 2> gpurun_out/r2_bench_c5_n2.err; echo "c5 n2 rc=$?"; tail -2 gpurun_out/r2_bench_c5_n2.err
python - <<PY
import json
for f in ("r2_bench_c4_n2","r2_bench_c5_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","placements_per_sec")}, (d.get("parity") or {}).get("ok"), (d.get("roofline") or {}).get("latency"))
    except Exception as e: print(f, "ERR", e)
PY
