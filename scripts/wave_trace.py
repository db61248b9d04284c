"""Per-wave trace of the multi-commit kernel on C4 (CCSIM_DEBUG_FLAGS=4: CTA 0 prints one line per wave): what a wave decided and why it ended.
Run on a GPU box:  CCSIM_DEBUG_FLAGS=4 python scripts/wave_trace.py > gpurun_out/wave_trace_c4.txt"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("cluster-capacity_b200.synth")
engine = importlib.import_module("cluster-capacity_b200.engine")
snap, tmpl, ctr = synth.c4()
with engine.Engine(device=0) as eng:
    eng.load_nodes(snap); eng.set_templates(tmpl, ctr)
    r = eng.run(0)
    print("placed", r.placed, "waves", r.waves, "ms", r.run_ms, flush=True)
