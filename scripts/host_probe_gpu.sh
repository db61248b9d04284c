# host-side (libcchost) evidence on the GPU box: object-level GPU tests, then the C4 bench line with the per-phase host timing
export CCSIM_NO_REBUILD=1
mkdir -p gpurun_out
nproc
timeout 300 python -m pytest tests/test_gpu_framework.py tests/test_pod_list.py tests/test_golden.py -m gpu -q 2>&1 | tail -3
CCHOST_TIMING=1 timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c4_host.json 2> gpurun_out/r2_bench_c4_host.err; echo "bench rc=$?"
grep "cchost" gpurun_out/r2_bench_c4_host.err | tail -45
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_c4_host.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, d["parity"]["ok"], json.dumps(d["e2e_objects"]))
PY
