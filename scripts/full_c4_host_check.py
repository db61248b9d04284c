"""Full-size check of the host path without a GPU: the C4 cluster as v1.Node / v1.Pod objects (synth.c4_objects) through libcchost's ingest +
encoder must give the C oracle exactly the placement sequence the flat generator (synth.c4) gives it (31 071 placements). ~1-2 minutes.
    CCHOST_THREADS=64 TWICE=1 python scripts/full_c4_host_check.py     # second pass: recycled arrays, warm pool"""
import importlib, sys, json, numpy as np, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from oracle import binding as oracle
synth = importlib.import_module("cluster-capacity_b200.synth")
fw = importlib.import_module("cluster-capacity_b200.framework")
snap, tmpl, ctr = synth.c4()
nodes, pods, t = synth.c4_objects()
for it in range(2 if __import__('os').environ.get('TWICE') else 1):     # second pass: recycled arrays, warm pool
    cc = fw.New(None, None, t, 0, [])
    t0 = time.time(); cc.SyncWithClient(fw.ListClient(nodes, pods, ())); t1 = time.time()
    s2, T2, c2, _, _, names = helpers.from_encoded(cc.EncodedSnapshot())
    for f in ("alloc_cpu", "alloc_mem", "alloc_pods", "req_cpu", "req_mem", "npods", "nz_cpu", "nz_mem"):
        assert np.array_equal(getattr(snap, f), getattr(s2, f)), f
    assert len(c2) == len(ctr)
    ra = oracle.run(snap, tmpl, ctr, threads=8, memo=True); rb = oracle.run(s2, T2, c2, threads=8, memo=True)
    assert ra.placed == rb.placed == 31071 and np.array_equal(ra.pod_node, rb.pod_node) and np.array_equal(ra.reason_hist, rb.reason_hist)
    assert names[0] == "node-000000" and names[-1] == "node-%06d" % (snap.n - 1)
    cc.Close()
    print("pass", it, "ok, sync+json %.1fs" % (t1 - t0), flush=True)
