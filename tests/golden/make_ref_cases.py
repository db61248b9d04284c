"""Writes tests/golden/ref_cases/*.json — the snapshots the Go harness (oracle/ref_harness/ccref_harness_test.go) feeds to the
UNMODIFIED reference on a Go-equipped box (`make -C oracle ref`). Each file: {"name", "nodes": [v1.Node], "pods": [v1.Pod],
"template": v1.Pod, "max_pods", "exclude_nodes"}. Regenerate with:  python tests/golden/make_ref_cases.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402

OUT = os.path.join(HERE, "ref_cases")
# variants whose outcome does not depend on the reference's random tie-breaking in a way that changes count / message
VARIANTS = ["plain", "selector", "tolerations", "affinity_terms", "hostports", "spread_zone", "anti_hostname", "extended", "best_effort",
            "init_overhead", "spread_everything"]


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = []
    asserted = json.load(open(os.path.join(HERE, "reference_asserted.json")))["cases"]
    for c in asserted:
        cases.append({"name": "asserted_" + c["name"], "nodes": c["nodes"], "pods": c["pods"], "template": c["template"], "max_pods": c["max_pods"],
                      "exclude_nodes": []})
    for seed in (21, 22):
        nodes, pods = helpers.random_cluster(seed, n_nodes=24, n_pods=40)
        # bound pods only: the reference would also schedule the source cluster's PENDING pods through its own scheduler and count
        # them as simulated instances (simulator.go:193-200,297-299) - a documented deviation of this repo, kept out of these cases
        pods = [p for p in pods if p["spec"].get("nodeName")]
        for v in VARIANTS:
            cases.append({"name": "%s_seed%d" % (v, seed), "nodes": nodes, "pods": pods, "template": helpers.template(v, seed), "max_pods": 0,
                          "exclude_nodes": []})
    for c in cases:
        json.dump(c, open(os.path.join(OUT, c["name"] + ".json"), "w"), sort_keys=True, separators=(",", ":"))
    print("wrote", len(cases), "cases to", OUT)


if __name__ == "__main__":
    main()
