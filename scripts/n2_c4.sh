export CCSIM_NO_REBUILD=1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-objects > gpurun_out/r2_bench_c4_n2.json 2> gpurun_out/r2_bench_c4_n2.err; echo "c4 n2 rc=$?"; tail -2 gpurun_out/r2_bench_c4_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_c4_n2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","placements_per_sec")}, (d.get("parity") or {}).get("ok"), (d.get("roofline") or {}).get("latency"))
PY
