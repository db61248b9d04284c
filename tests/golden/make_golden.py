"""Generates tests/golden/*.json — committed known-answer fixtures for the hot path.

The reference holds no golden vectors for this path and cannot be run here (Go toolchain absent), so these are NOT outputs
of the reference binary: they are (a) the reference's own asserted outcomes and README numbers, written down by hand with
their source line, and (b) outputs of the object-level oracle (oracle/objref.py, a line-by-line restatement of the vendored
scheduler sources) for object-level scenarios. Regenerate with:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import objref  # noqa: E402


def main():
    fixed = {
        "source": "reference-asserted outcomes; see 'cite' of each entry",
        "cases": [
            {"name": "readme_demo", "cite": "README.md:44-66", "nodes": [helpers.make_node("kube-node-%d" % i, cpu="2", mem="4Gi", pods="110") for i in range(1, 5)],
             "pods": [], "template": helpers.make_pod("small-pod", cpu="150m", mem="100Mi"), "max_pods": 0,
             "expect": {"replicas": 52, "failType": "Unschedulable", "per_node": {"kube-node-1": 13, "kube-node-2": 13, "kube-node-3": 13, "kube-node-4": 13}}},
            # README.md:68-101: the same cluster after `examples/rc.yml` was scaled to 6 nginx replicas (150m / 100Mi each, two on
            # kube-node-1 and kube-node-2, one on kube-node-3 and kube-node-4): 46 instances, 11 / 12 / 11 / 12
            {"name": "readme_demo_with_rc", "cite": "README.md:68-101; examples/rc.yml",
             "nodes": [helpers.make_node("kube-node-%d" % i, cpu="2", mem="4Gi", pods="110") for i in range(1, 5)],
             "pods": [helpers.make_pod("nginx-%d" % j, cpu="150m", mem="100Mi", node="kube-node-%d" % nd, labels={"app": "nginx"})
                      for j, nd in enumerate([1, 1, 2, 2, 3, 4])],
             "template": helpers.make_pod("small-pod", cpu="150m", mem="100Mi"), "max_pods": 0,
             "expect": {"replicas": 46, "failType": "Unschedulable", "per_node": {"kube-node-1": 11, "kube-node-2": 11, "kube-node-3": 12, "kube-node-4": 12}}},
            {"name": "test_prediction_limit", "cite": "pkg/framework/simulator_test.go:162-173,250-252",
             "nodes": [helpers.make_node("test-node-1", cpu="300m", mem="1000000000", pods="3"), helpers.make_node("test-node-2", cpu="400m", mem="2000000000", pods="3"),
                       helpers.make_node("test-node-3", cpu="1200m", mem="1000000000", pods="3")],
             "pods": [], "template": helpers.make_pod("simulated-pod", cpu="100m", mem="5000000"), "max_pods": 6,
             "expect": {"replicas": 6, "failType": "LimitReached"}},
            {"name": "test_prediction_unlimited", "cite": "pkg/framework/simulator_test.go:162-173,250-252",
             "nodes": [helpers.make_node("test-node-1", cpu="300m", mem="1000000000", pods="3"), helpers.make_node("test-node-2", cpu="400m", mem="2000000000", pods="3"),
                       helpers.make_node("test-node-3", cpu="1200m", mem="1000000000", pods="3")],
             "pods": [], "template": helpers.make_pod("simulated-pod", cpu="100m", mem="5000000"), "max_pods": 0,
             "expect": {"replicas": 9, "failType": "Unschedulable"}},
            {"name": "e2e_limit_reached", "cite": "test/e2e/e2e_test.go:36-39,171-173", "nodes": [helpers.make_node("w%d" % i, cpu="4", mem="8Gi", pods="110") for i in range(2)],
             "pods": [], "template": helpers.make_pod("p", cpu="100m", mem="64Mi"), "max_pods": 5, "expect": {"replicas": 5, "failType": "LimitReached"}},
        ]}
    json.dump(fixed, open(os.path.join(HERE, "reference_asserted.json"), "w"), indent=1, sort_keys=True)
    gen = {"source": "oracle/objref.py (object-level restatement of the vendored kube-scheduler v1.34.1); NOT reference outputs", "cases": []}
    for seed in (21, 22):
        nodes, pods = helpers.random_cluster(seed, n_nodes=24, n_pods=40)
        for variant in helpers.TEMPLATE_VARIANTS:
            tmpl = helpers.template(variant, seed)
            sim = objref.Simulator(tmpl, 0, ())
            helpers.objref_sync(sim, nodes, pods, variant)
            sim.run()
            gen["cases"].append({"name": "%s_seed%d" % (variant, seed), "cluster_seed": seed, "variant": variant,
                                 "scheduled": sim.pods_status, "stop_reason": sim.stop_reason})
    json.dump(gen, open(os.path.join(HERE, "objref_sequences.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(fixed["cases"]), "+", len(gen["cases"]), "cases")


if __name__ == "__main__":
    main()
