"""Where does the end-to-end time of one C-ABI analysis go? (wall clock per call + device time of the run)"""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("cluster-capacity_b200.synth")
engine = importlib.import_module("cluster-capacity_b200.engine")
snap, tmpl, ctr = synth.c4()
with engine.Engine(device=0) as eng:
    for it in range(4):
        t0 = time.perf_counter(); eng.load_nodes(snap)
        t1 = time.perf_counter(); eng.set_templates(tmpl, ctr)
        t2 = time.perf_counter(); r = eng.run(0)
        t3 = time.perf_counter(); r2 = eng.run(0)
        t4 = time.perf_counter()
        print("iter %d: load_nodes %.2f ms  set_templates %.2f ms  run#1 wall %.2f ms (device %.2f)  run#2 wall %.2f ms (device %.2f)" % (
            it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, r.run_ms, (t4 - t3) * 1e3, r2.run_ms), flush=True)
