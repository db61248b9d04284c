"""Quick device-time probe of the wave kernel on the BASELINE configs (not the bench contract; see bench.py)."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
abi = importlib.import_module("cluster-capacity_b200._abi")
synth = importlib.import_module("cluster-capacity_b200.synth")
engine = importlib.import_module("cluster-capacity_b200.engine")

def probe(name, snap, tmpl, ctr, limit, bytes_per_eval):
    with engine.Engine(device=0) as eng:
        t0 = time.time(); eng.load_nodes(snap); eng.set_templates(tmpl, ctr); t1 = time.time()
        r = eng.run(limit)
        r = eng.run(limit)
        info = eng.device_info()
        st = eng.run_stats()
    us = r.run_ms * 1e3 / max(1, r.waves)
    print("%-28s n=%-8d grid=%-4d placed=%-8d waves=%-8d run=%9.3f ms  %6.2f us/wave  %.3g evals/s  %.0f GB/s algorithmic  (load %.1f ms)" % (
        name, snap.n, info["grid"], r.placed, r.waves, r.run_ms, us, r.evals / (r.run_ms * 1e-3),
        r.evals * bytes_per_eval / (r.run_ms * 1e-3) / 1e9, (t1 - t0) * 1e3), flush=True)
    if st["engine"] == "multi-commit":
        w = max(1, st["waves"])
        print("    engine=%s  placements/wave=%.2f  candidates/wave=%.1f  bar raised in %d waves  cycles/wave (CTA 0): scan=%d S1=%d merge+publish=%d gather=%d replay=%d tail=%d  smem=%d B" % (
            st["engine"], st["placed"] / w, st["candidates"] / w, st["bar_raised_waves"], *[c // w for c in st["phase_cycles"][:6]], st["smem_bytes"]), flush=True)
        print("    replay detail: rounds/wave=%.2f  setup=%d cycles/wave  waves ended by a binding minimum move: %d of %d" % (st["smem_bytes"] / w, st["phase_cycles"][6] // w, st["phase_cycles"][7], w), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    if "c2" in which: probe("C2 10k fit-only", *synth.c2(), 20000, 72)
    if "c2big" in which: probe("C2 100k fit-only", *synth.c2(n=100_000), 20000, 72)
    if "c3" in which: probe("C3 50k full filters", *synth.c3(), 20000, 88)
    if "c4" in which: probe("C4 100k PTS+IPA", *synth.c4(), 0, 96)
    if "c4spread" in which:      # spread constraints only (no hostname anti-affinity): nodes take several clones, winners can come back
        snap, tmpl, ctr = synth.c4()
        tmpl[0].n_anti = 0
        probe("C4 spread-only", snap, tmpl, ctr[:3], 20000, 92)
        os.environ["CCSIM_FORCE_SEQUENTIAL"] = "1"
        probe("C4 spread-only, sequential", snap, tmpl, ctr[:3], 20000, 92)
        del os.environ["CCSIM_FORCE_SEQUENTIAL"]
    if "c5" in which: probe("C5 1M x 64 templates", *synth.c5(), 6400, 72)
