"""Small runs of every wave kernel for compute-sanitizer (memcheck / racecheck): multi-commit, lean, batched, generic soft scorers."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
abi = importlib.import_module("cluster-capacity_b200._abi")
synth = importlib.import_module("cluster-capacity_b200.synth")
engine = importlib.import_module("cluster-capacity_b200.engine")
from oracle import binding as oracle

def run(name, snap, tmpl, ctr, limit, kind=abi.ENGINE_AUTO):
    want = oracle.run(snap, tmpl, ctr, max_pods=limit, threads=4)
    with engine.Engine(device=0, engine=kind) as eng:
        eng.load_nodes(snap); eng.set_templates(tmpl, ctr)
        r = eng.run(limit)
    ok = r.placed == want.placed and np.array_equal(r.pod_node, want.pod_node)
    print("%-22s placed %6d waves %6d parity %s" % (name, r.placed, r.waves, ok), flush=True)
    assert ok

run("multi-commit (C4)", *synth.c4(n=20000, n_existing=30000, zones=16, racks=128, regions=4), 300)
run("lean sequential (C4)", *synth.c4(n=20000, n_existing=30000, zones=16, racks=128, regions=4), 100, abi.ENGINE_SEQUENTIAL)
run("batched (C2)", *synth.c2(n=5000), 2000)
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "test_gpu_parity.py")).read()
ns = {}
exec(src.replace("pytestmark = pytest.mark.gpu", ""), ns)
run("generic soft scorers", *ns["_soft_cluster"](4, n=1500), 150)
