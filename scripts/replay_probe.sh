# replay anatomy of the multi-commit kernel on C4: look-ahead waves (default) against strict waves (CCSIM_DEBUG_FLAGS & 16), same binary
export CCSIM_NO_REBUILD=1
echo "== look-ahead"; CCSIM_DEBUG_FLAGS=8 timeout 200 python scripts/perf_probe.py c4 2>&1 | tail -6
echo "== strict";     CCSIM_DEBUG_FLAGS=24 timeout 200 python scripts/perf_probe.py c4 2>&1 | tail -6
