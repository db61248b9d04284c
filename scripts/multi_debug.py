"""Debug probe: multi-commit engine vs the oracle on one synthetic C4 configuration; prints the first mismatch with context."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
abi = importlib.import_module("cluster-capacity_b200._abi")
synth = importlib.import_module("cluster-capacity_b200.synth")
engine = importlib.import_module("cluster-capacity_b200.engine")
from oracle import binding as oracle

def main():
    kw = dict(n=4000, n_existing=8000, zones=8, racks=64, regions=4)
    snap, tmpl, ctr = synth.c4(**kw)
    want = oracle.run(snap, tmpl, ctr, threads=4, memo=True)
    with engine.Engine(device=0) as eng:
        eng.load_nodes(snap); eng.set_templates(tmpl, ctr)
        got = eng.run(0)
        st = eng.run_stats()
    print("flags", os.environ.get("CCSIM_DEBUG_FLAGS"), "placed", got.placed, want.placed, "waves", got.waves, st["engine"])
    m = min(len(got.pod_node), len(want.pod_node))
    d = np.nonzero(got.pod_node[:m] != want.pod_node[:m])[0]
    if not len(d):
        print("sequences equal"); return
    k = int(d[0])
    zone, rack, region = snap.topo
    cz = np.array(ctr[0].init_np if hasattr(ctr[0], "init_np") else np.ctypeslib.as_array(ctr[0].init, (ctr[0].n_domains,))).copy()
    cr = np.ctypeslib.as_array(ctr[1].init, (ctr[1].n_domains,)).copy()
    cg = np.ctypeslib.as_array(ctr[2].init, (ctr[2].n_domains,)).copy()
    for w in want.pod_node[:k]:
        cz[zone[w]] += 1; cr[rack[w]] += 1; cg[region[w]] += 1
    print("first mismatch at pod", k, "got node", got.pod_node[k], "want node", want.pod_node[k])
    for name, nd in (("got", got.pod_node[k]), ("want", want.pod_node[k])):
        print(" ", name, "zone", zone[nd], cz[zone[nd]], "min", cz.min(), "| rack", rack[nd], cr[rack[nd]], "min", cr.min(), "| region", region[nd], cg[region[nd]], "min", cg.min(),
              "| already placed on it:", int((want.pod_node[:k] == nd).sum()))
    print("  got[k-3:k+3]", got.pod_node[max(0, k - 3):k + 3], "want", want.pod_node[max(0, k - 3):k + 3])
    # is the oracle's node later in our sequence / ours in the oracle's?
    print("  want node appears in got at", np.nonzero(got.pod_node == want.pod_node[k])[0][:3], " got node appears in want at", np.nonzero(want.pod_node == got.pod_node[k])[0][:3])

main()
