"""`genpod` — namespace limits -> stub pod spec, the reference's second command (cmd/genpod, pkg/client/nspod.go:36-131):

    python -m cluster-capacity_b200.genpod --namespace NS --snapshot cluster.json [--output json|yaml]

RetrieveNamespacePod: a pause-container pod in the namespace whose requests == limits == the MINIMUM over all LimitRange
items of type Pod of `max` for cpu / memory / nvdia.com/gpu [sic], plus the `openshift.io/node-selector` namespace annotation
as nodeSelector. The objects come from `--snapshot` ({"namespaces": [...], "limitranges": [...]}) or, with `--kubeconfig`,
from kubectl. The resulting spec is what `cluster-capacity --podspec` (and, for several namespaces, the multi-template
run of BASELINE config C5) consumes.
"""
import argparse
import importlib
import json
import subprocess
import sys
from fractions import Fraction

import yaml

RESOURCES = ("memory", "cpu", "nvdia.com/gpu")
_SUF = {"n": Fraction(1, 10**9), "u": Fraction(1, 10**6), "m": Fraction(1, 1000), "": Fraction(1), "k": Fraction(10**3), "M": Fraction(10**6),
        "G": Fraction(10**9), "T": Fraction(10**12), "P": Fraction(10**15), "E": Fraction(10**18), "Ki": Fraction(2**10), "Mi": Fraction(2**20),
        "Gi": Fraction(2**30), "Ti": Fraction(2**40), "Pi": Fraction(2**50), "Ei": Fraction(2**60)}


def _qty(s):
    s = str(s)
    i = 0
    while i < len(s) and (s[i].isdigit() or s[i] in "+-."):
        i += 1
    suf = s[i:]
    if suf[:1] in ("e", "E") and len(suf) > 1 and (suf[1].isdigit() or suf[1] in "+-"):
        return Fraction(s[:i]) * Fraction(10) ** int(suf[1:])
    return Fraction(s[:i]) * _SUF[suf]


def convert_selector_to_labels_map(sel):
    """labels.ConvertSelectorToLabelsMap: "k1=v1,k2=v2" -> {k1: v1, k2: v2}; anything else is an error."""
    out = {}
    if not sel:
        return out
    for part in sel.split(","):
        kv = part.split("=")
        if len(kv) != 2:
            raise ValueError("invalid selector: %r" % part)
        out[kv[0].strip()] = kv[1].strip()
    return out


def retrieve_namespace_pod(namespaces, limitranges, namespace):
    """RetrieveNamespacePod (pkg/client/nspod.go:36-131) over listed objects."""
    ns = next((n for n in namespaces if n["metadata"]["name"] == namespace), None)
    if ns is None:
        raise LookupError("Namespace %s not found" % namespace)
    pod = {"apiVersion": "v1", "kind": "Pod",
           "metadata": {"name": "cluster-capacity-stub-container", "namespace": namespace},
           "spec": {"containers": [{"name": "cluster-capacity-stub-container", "image": "gcr.io/google_containers/pause:2.0",
                                    "imagePullPolicy": "Always"}],
                    "restartPolicy": "OnFailure", "dnsPolicy": "Default"}}
    best = {r: None for r in RESOURCES}
    for lr in limitranges:
        if (lr.get("metadata") or {}).get("namespace") != namespace:
            continue
        for item in (lr.get("spec") or {}).get("limits") or []:
            if item.get("type") != "Pod":
                continue
            for r in RESOURCES:
                amount = (item.get("max") or {}).get(r)
                if amount is None:
                    continue
                if best[r] is None or _qty(best[r]) > _qty(amount):     # Cmp(amount) == 1: keep the smaller
                    best[r] = amount
    if any(v is not None and _qty(v) != 0 for v in best.values()):
        rl = {r: str(v) for r, v in best.items() if v is not None}
        pod["spec"]["containers"][0]["resources"] = {"limits": dict(rl), "requests": dict(rl)}
    ann = (ns["metadata"].get("annotations") or {})
    if "openshift.io/node-selector" in ann:
        try:
            pod["spec"]["nodeSelector"] = convert_selector_to_labels_map(ann["openshift.io/node-selector"])
        except ValueError as e:
            raise ValueError("Unable to parse openshift.io/node-selector in %s namespace: %s" % (ann["openshift.io/node-selector"], e))
    return pod


def main(argv=None):
    ap = argparse.ArgumentParser(prog="genpod", description="Generate pod based on namespace resource limits and node selector annotations")
    ap.add_argument("--kubeconfig", default="")
    ap.add_argument("--namespace", default="", help="Cluster namespace (a comma-separated list with --output-dir: one podspec per namespace)")
    ap.add_argument("--output-dir", default="", help="write <namespace>.yaml|json per namespace into this directory (feeds cluster-capacity --podspec DIR)")
    ap.add_argument("--output", "-o", default="", help="Output format. One of: json|yaml")
    ap.add_argument("--snapshot", default="", help="JSON/YAML file with namespaces / limitranges lists")
    a = ap.parse_args(argv)
    if not a.namespace:
        print("Cluster namespace missing")
        ap.print_help()
        return 0
    if a.output and a.output not in ("json", "yaml"):
        print("Output format %s not recognized: only json and yaml are allowed" % a.output)
        return 0
    if a.snapshot:
        d = yaml.safe_load(open(a.snapshot))
        nss, lrs = d.get("namespaces") or [], d.get("limitranges") or []
    else:
        base = ["kubectl"] + (["--kubeconfig", a.kubeconfig] if a.kubeconfig else [])
        nss = json.loads(subprocess.check_output(base + ["get", "namespaces", "-o", "json"]))["items"]
        lrs = json.loads(subprocess.check_output(base + ["get", "limitranges", "-n", a.namespace, "-o", "json"]))["items"]
    if a.output_dir:       # several namespaces -> several podspecs (the multi-template run of BASELINE config C5)
        import os
        os.makedirs(a.output_dir, exist_ok=True)
        for ns in [x for x in a.namespace.split(",") if x]:
            try:
                pod = retrieve_namespace_pod(nss, lrs, ns)
            except (LookupError, ValueError) as e:
                print("Error: %s" % e)
                return 1
            pod["metadata"]["name"] = "cluster-capacity-stub-container-" + ns     # report rows are keyed by pod name
            with open(os.path.join(a.output_dir, ns + (".json" if a.output == "json" else ".yaml")), "w") as f:
                f.write(json.dumps(pod) if a.output == "json" else yaml.safe_dump(pod, default_flow_style=False, sort_keys=True))
        return 0
    try:
        pod = retrieve_namespace_pod(nss, lrs, a.namespace)
    except (LookupError, ValueError) as e:
        print("Error: %s" % e)
        return 1
    if a.output == "json":
        print(json.dumps(pod))
    else:                       # utils.PrintPod defaults to YAML
        print(yaml.safe_dump(pod, default_flow_style=False, sort_keys=True), end="")
    return 0


if __name__ == "__main__":
    sys.exit(main())
