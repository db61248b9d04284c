"""Ingest + encode time of the C4 cluster (v1.Node / v1.Pod JSON, 91 MB) for the CCHOST_THREADS given on the command line: one fresh
process per setting (the thread count is read once per process). No GPU needed: cc_debug_encoded_snapshot stops after the encoder."""
import importlib, sys, os, time, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import ctypes as C
    sys.path.insert(0, ROOT)
    synth = importlib.import_module("cluster-capacity_b200.synth")
    fw = importlib.import_module("cluster-capacity_b200.framework")
    nodes, pods, tmpl = synth.c4_objects()
    nj = json.dumps({"items": nodes}).encode(); pj = json.dumps({"items": pods}).encode(); tj = json.dumps(tmpl).encode()
    L = fw.lib(); best = (1e9, 1e9)
    for it in range(4):
        h = C.c_void_p(); assert L.cc_new(None, tj, 0, None, 0, C.byref(h)) == 0
        t0 = time.perf_counter(); assert L.cc_sync_with_objects(h, nj, pj, None) == 0
        t1 = time.perf_counter()
        os.environ["CCHOST_TIMING"] = "1"
        L.cc_close(h)
        if it and t1 - t0 < best[0]: best = (t1 - t0, 0)
    print("CCHOST_THREADS=%s  cc_sync_with_objects best of 3: %.1f ms" % (os.environ.get("CCHOST_THREADS", "default"), best[0] * 1e3), flush=True)
else:
    for t in sys.argv[1:] or ["8", "16", "32", "64", "128"]:
        env = dict(os.environ, CCHOST_THREADS=t)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env)
