# round-end evidence on ONE B200 (run under gpurun): the whole GPU test suite, the bench lines of both workloads and of the
# reference arm, and the ncu launch list of the default bench command. Outputs under gpurun_out/ (copied to profiles/ afterwards).
export CCSIM_NO_REBUILD=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -1 gpurun_out/r2_bench_c4.err
python bench.py --workload c5 --steps 3 --warmup 3 --no-objects > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c4_ref.json 2> gpurun_out/r2_bench_c4_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-objects --no-parity > gpurun_out/r2_bench_under_ncu.log 2>&1
python - <<PY
import json
for f in ("r2_bench_c4","r2_bench_c5","r2_bench_c4_ref"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","placements_per_sec")}, (d.get("parity") or {}).get("ok"), (d.get("e2e_objects") or {}).get("ms_per_step"), (d.get("clocks")))
    except Exception as e: print(f, "ERR", e)
PY
