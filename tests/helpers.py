"""Shared test helpers: encoded-snapshot -> ctypes structs, object-level scenario generators."""
import ctypes as C
import importlib
import random

import numpy as np

abi = importlib.import_module("cluster-capacity_b200._abi")


def from_encoded(enc):
    """cc_debug_encoded_snapshot JSON -> (Snapshot, [Template], [Counter], taint_dict, scalar_names, names)."""
    nd = enc["nodes"]
    n = nd["n"]
    u64 = lambda a: np.array([int(x) for x in a], dtype=np.uint64)
    scal = [(np.array(a, np.int64), np.array(r, np.int64)) for a, r in zip(nd["alloc_scalar"], nd["req_scalar"])]
    lists = [nd["taint_list"][nd["taint_off"][i]:nd["taint_off"][i + 1]] for i in range(n)]
    snap = abi.Snapshot(n, np.array(nd["alloc_cpu"], np.int64), np.array(nd["alloc_mem"], np.int64), np.array(nd["alloc_pods"], np.int32),
                        alloc_eph=np.array(nd["alloc_eph"], np.int64), req_cpu=np.array(nd["req_cpu"], np.int64),
                        req_mem=np.array(nd["req_mem"], np.int64), req_eph=np.array(nd["req_eph"], np.int64),
                        npods=np.array(nd["npods"], np.int32), nz_cpu=np.array(nd["nz_cpu"], np.int64), nz_mem=np.array(nd["nz_mem"], np.int64),
                        scalars=scal, taint_mask=u64(nd["taint_mask"]).reshape(nd["taint_words"], n) if n else None,
                        taint_nosched=[int(x) for x in nd["taint_nosched"]], taint_prefer=[int(x) for x in nd["taint_prefer"]],
                        static_mask=u64(nd["static_mask"]).reshape(nd["static_words"], n) if nd["static_words"] else None,
                        topo=[np.array(t, np.int32) for t in nd["topo"]], has_placed_mask=nd["has_placed_mask"],
                        taint_lists=lists, names=enc["names"])
    ts = []
    for k, hx in enumerate(enc.get("templates_hex") or [enc["template_hex"]]):
        t = abi.Template()
        raw = bytes.fromhex(hx)
        assert len(raw) == C.sizeof(abi.Template)
        C.memmove(C.byref(t), raw, len(raw))
        img = np.array((enc.get("image_scores") or [enc.get("image_score") or []])[k], np.uint8)   # the hex carries a pointer of the encoding process: replace it
        t._keep_img = img
        t.image_score = img.ctypes.data_as(C.POINTER(C.c_uint8)) if len(img) else None
        ts.append(t)
    ctr = [abi.make_counter(c["topo_col"], np.array(c["init"], np.int32), n_present=c["n_present"], inc=c["inc"], elig_bit=c.get("elig_bit", -1))
           for c in enc["counters"]]
    return snap, ts, ctr, nd["taint_dict"], nd["scalar_names"], enc["names"]


def reason_text(r, taint_dict, scalar_names):
    if r < abi.R_FIXED_COUNT:
        return abi.REASON_TEXT[r]
    if r < abi.R_TAINT0:
        return "Insufficient %s" % scalar_names[r - abi.R_SCALAR0]
    t = taint_dict[r - abi.R_TAINT0]
    return "node(s) had untolerated taint {%s: %s}" % (t["key"], t["value"])


def stop_reason_from_result(res, n, max_pods, taint_dict, scalar_names, preemption_never=False):
    if res.stop_code == abi.STOP_LIMIT_REACHED:
        return "LimitReached: Maximum number of pods simulated: %d" % max_pods
    hist = {i: int(c) for i, c in enumerate(res.reason_hist) if c}
    msg = abi.fit_error_message(n, hist, res.preempt_no_victims, res.preempt_not_helpful, lambda r: reason_text(r, taint_dict, scalar_names))
    if preemption_never:
        msg = msg.split(" preemption: ")[0] + " preemption: not eligible due to preemptionPolicy=Never."
    return "Unschedulable: " + msg


# ---- object-level scenarios ------------------------------------------------------------------------------------------
def make_node(name, cpu="4", mem="8Gi", pods="20", labels=None, taints=None, unschedulable=False, extra_alloc=None):
    alloc = {"cpu": cpu, "memory": mem, "pods": pods, "ephemeral-storage": "100Gi"}
    alloc.update(extra_alloc or {})
    lab = {"kubernetes.io/hostname": name}
    lab.update(labels or {})
    n = {"apiVersion": "v1", "kind": "Node", "metadata": {"name": name, "labels": lab}, "spec": {}, "status": {"allocatable": alloc}}
    if taints:
        n["spec"]["taints"] = taints
    if unschedulable:
        n["spec"]["unschedulable"] = True
    return n


def make_pod(name, cpu=None, mem=None, node=None, labels=None, ns="default", phase="Running", **spec_extra):
    req = {}
    if cpu:
        req["cpu"] = cpu
    if mem:
        req["memory"] = mem
    p = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": name, "namespace": ns, "labels": labels or {}},
         "spec": {"containers": [{"name": "c", "image": "img", "resources": {"requests": req}}]}, "status": {"phase": phase}}
    if node:
        p["spec"]["nodeName"] = node
    p["spec"].update(spec_extra)
    return p


def random_cluster(seed, n_nodes=40, n_pods=60, zones=3):
    rng = random.Random(seed)
    nodes, pods = [], []
    for i in range(n_nodes):
        labels = {}
        if rng.random() < 0.9:
            labels["topology.kubernetes.io/zone"] = "z%d" % rng.randrange(zones)
            labels["topology.kubernetes.io/region"] = "r%d" % rng.randrange(2)
        if rng.random() < 0.6:
            labels["disk"] = rng.choice(["ssd", "hdd"])
        if rng.random() < 0.5:
            labels["rank"] = str(rng.randrange(10))
        taints = []
        if rng.random() < 0.2:
            taints.append({"key": "dedicated", "value": rng.choice(["a", "b"]), "effect": "NoSchedule"})
        if rng.random() < 0.15:
            taints.append({"key": "flaky", "effect": "PreferNoSchedule"})
        if rng.random() < 0.1:
            taints.append({"key": "gpu", "value": "true", "effect": "NoExecute"})
        extra = {"example.com/foo": str(rng.randrange(0, 6))} if rng.random() < 0.5 else None
        nodes.append(make_node("node-%02d" % i, cpu=rng.choice(["2", "4", "8", "3500m"]), mem=rng.choice(["4Gi", "8Gi", "16Gi", "6000Mi"]),
                               pods=str(rng.choice([5, 8, 12, 110])), labels=labels, taints=taints,
                               unschedulable=rng.random() < 0.05, extra_alloc=extra))
        images = []
        if rng.random() < 0.3:
            images.append({"names": ["img:latest", "registry.local/img@sha256:0123"], "sizeBytes": rng.choice([120, 300, 700]) * 1024 * 1024})
        if rng.random() < 0.2:
            images.append({"names": ["y:latest"], "sizeBytes": 900 * 1024 * 1024})
        if images:
            nodes[-1]["status"]["images"] = images
    for j in range(n_pods):
        node = "node-%02d" % rng.randrange(n_nodes) if rng.random() < 0.92 else None
        labels = {"app": rng.choice(["web", "db", "sim"])}
        extra = {}
        if rng.random() < 0.15:
            extra["affinity"] = {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
                {"labelSelector": {"matchLabels": {"app": rng.choice(["sim", "db"])}}, "topologyKey": rng.choice(["kubernetes.io/hostname", "topology.kubernetes.io/zone"])}]}}
        r = rng.random()
        if r < 0.08:      # scored through hardPodAffinityWeight when it matches the incoming pod
            extra.setdefault("affinity", {})["podAffinity"] = {"requiredDuringSchedulingIgnoredDuringExecution": [
                {"labelSelector": {"matchLabels": {"app": rng.choice(["sim", "web"])}}, "topologyKey": "topology.kubernetes.io/zone"}]}
        elif r < 0.16:
            extra.setdefault("affinity", {})["podAffinity"] = {"preferredDuringSchedulingIgnoredDuringExecution": [
                {"weight": rng.choice([10, 35]), "podAffinityTerm": {"labelSelector": {"matchExpressions": [{"key": "app", "operator": "In", "values": ["sim", "web"]}]},
                                                                    "topologyKey": rng.choice(["topology.kubernetes.io/zone", "disk"])}}]}
        elif r < 0.24:
            aa = extra.setdefault("affinity", {}).setdefault("podAntiAffinity", {})
            aa["preferredDuringSchedulingIgnoredDuringExecution"] = [
                {"weight": rng.choice([5, 60]), "podAffinityTerm": {"labelSelector": {"matchLabels": {"app": "sim"}}, "topologyKey": "kubernetes.io/hostname"}}]
        p = make_pod("pod-%03d" % j, cpu=rng.choice([None, "100m", "250m", "1"]), mem=rng.choice([None, "64Mi", "256Mi", "1Gi"]),
                     node=node, labels=labels, phase=rng.choice(["Running"] * 8 + ["Succeeded", "Pending"]), **extra)
        if rng.random() < 0.2:
            p["spec"]["containers"][0]["ports"] = [{"containerPort": 80, "hostPort": rng.choice([8080, 9090]), "protocol": "TCP"}]
        if rng.random() < 0.15:
            p["spec"]["initContainers"] = [{"name": "init", "image": "img", "resources": {"requests": {"cpu": "500m", "memory": "32Mi"}}}]
        if rng.random() < 0.1:
            p["spec"]["overhead"] = {"cpu": "10m", "memory": "8Mi"}
        if rng.random() < 0.3:
            p["spec"]["containers"][0]["resources"]["requests"]["example.com/foo"] = "1"
        pods.append(p)
    return nodes, pods


TEMPLATE_VARIANTS = ["plain", "selector", "tolerations", "affinity_terms", "hostports", "spread_zone", "spread_two", "anti_hostname",
                     "anti_zone", "affinity_zone", "extended", "best_effort", "init_overhead", "never_preempt", "gt_lt", "name_in", "pref_affinity", "pref_and_required",
                     "soft_spread", "soft_and_hard", "pref_pod_affinity", "svc_default_spread", "owner_default_spread", "spread_everything"]


def workloads_for(variant):
    """Services / controllers synced next to the nodes and pods (only the *_default_spread variants need them)."""
    svc = lambda name, sel, ns="default": {"apiVersion": "v1", "kind": "Service", "metadata": {"name": name, "namespace": ns}, "spec": {"selector": sel}}
    if variant == "svc_default_spread":
        return {"services": [svc("sim", {"app": "sim"}), svc("other-ns", {"app": "sim", "x": "y"}, ns="kube-system"), svc("web", {"app": "web"}),
                             {"apiVersion": "v1", "kind": "Service", "metadata": {"name": "headless", "namespace": "default"}, "spec": {}}]}
    if variant == "owner_default_spread":
        return {"services": [svc("web", {"app": "web"})],
                "replica_sets": [{"apiVersion": "apps/v1", "kind": "ReplicaSet", "metadata": {"name": "sim-rs", "namespace": "default"},
                                  "spec": {"selector": {"matchExpressions": [{"key": "app", "operator": "In", "values": ["sim", "db"]}]}}}]}
    return {}


def list_client(fw, nodes, pods, variant=None, namespaces=()):
    return fw.ListClient(nodes, pods, namespaces, **workloads_for(variant))


def objref_sync(sim, nodes, pods, variant=None, namespaces=()):
    w = workloads_for(variant)
    sim.sync(nodes, pods, namespaces, services=w.get("services", ()), rcs=w.get("replication_controllers", ()),
             replicasets=w.get("replica_sets", ()), statefulsets=w.get("stateful_sets", ()))


def template(variant, seed=0):
    del seed   # variants are deterministic; the parameter keeps the call sites symmetrical with random_cluster
    p = make_pod("small-pod", cpu="150m", mem="100Mi", labels={"app": "sim"})
    s = p["spec"]
    if variant == "selector":
        s["nodeSelector"] = {"disk": "ssd"}
    elif variant == "tolerations":
        s["tolerations"] = [{"key": "dedicated", "operator": "Equal", "value": "a", "effect": "NoSchedule"}, {"key": "gpu", "operator": "Exists"},
                            {"key": "flaky", "operator": "Exists", "effect": "PreferNoSchedule"}]
    elif variant == "affinity_terms":
        s["affinity"] = {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
            {"matchExpressions": [{"key": "disk", "operator": "In", "values": ["ssd"]}, {"key": "rank", "operator": "Exists"}]},
            {"matchExpressions": [{"key": "topology.kubernetes.io/zone", "operator": "NotIn", "values": ["z0"]}, {"key": "disk", "operator": "DoesNotExist"}]}]}}}
    elif variant == "gt_lt":
        s["affinity"] = {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
            {"matchExpressions": [{"key": "rank", "operator": "Gt", "values": ["3"]}, {"key": "rank", "operator": "Lt", "values": ["8"]}]}]}}}
    elif variant == "name_in":
        s["affinity"] = {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
            {"matchFields": [{"key": "metadata.name", "operator": "In", "values": ["node-03"]}]},
            {"matchFields": [{"key": "metadata.name", "operator": "In", "values": ["node-07"]}]}]}}}
    elif variant == "pref_affinity":
        s["affinity"] = {"nodeAffinity": {"preferredDuringSchedulingIgnoredDuringExecution": [
            {"weight": 50, "preference": {"matchExpressions": [{"key": "disk", "operator": "In", "values": ["ssd"]}]}},
            {"weight": 20, "preference": {"matchExpressions": [{"key": "topology.kubernetes.io/zone", "operator": "In", "values": ["z1", "z2"]}]}},
            {"weight": 0, "preference": {"matchExpressions": [{"key": "rank", "operator": "Exists"}]}},
            {"weight": 7, "preference": {"matchExpressions": [{"key": "rank", "operator": "Gt", "values": ["4"]}]}}]}}
        s["tolerations"] = [{"key": "dedicated", "operator": "Exists"}]
    elif variant == "pref_and_required":
        s["affinity"] = {"nodeAffinity": {
            "requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [{"matchExpressions": [{"key": "disk", "operator": "Exists"}]}]},
            "preferredDuringSchedulingIgnoredDuringExecution": [
                {"weight": 100, "preference": {"matchExpressions": [{"key": "disk", "operator": "In", "values": ["hdd"]}]}},
                {"weight": 1, "preference": {"matchFields": [{"key": "metadata.name", "operator": "In", "values": ["node-05"]}]}}]}}
    elif variant == "soft_spread":
        s["topologySpreadConstraints"] = [
            {"maxSkew": 2, "topologyKey": "topology.kubernetes.io/zone", "whenUnsatisfiable": "ScheduleAnyway", "labelSelector": {"matchLabels": {"app": "sim"}}},
            {"maxSkew": 1, "topologyKey": "kubernetes.io/hostname", "whenUnsatisfiable": "ScheduleAnyway",
             "labelSelector": {"matchExpressions": [{"key": "app", "operator": "In", "values": ["sim", "web"]}]}}]
    elif variant == "soft_and_hard":
        s["topologySpreadConstraints"] = [
            {"maxSkew": 3, "topologyKey": "topology.kubernetes.io/zone", "whenUnsatisfiable": "DoNotSchedule", "labelSelector": {"matchLabels": {"app": "sim"}}},
            {"maxSkew": 1, "topologyKey": "topology.kubernetes.io/region", "whenUnsatisfiable": "ScheduleAnyway", "labelSelector": {"matchLabels": {"app": "sim"}},
             "nodeTaintsPolicy": "Honor"},
            {"maxSkew": 4, "topologyKey": "disk", "whenUnsatisfiable": "ScheduleAnyway", "labelSelector": {"matchLabels": {"app": "db"}}, "nodeAffinityPolicy": "Ignore"}]
    elif variant == "pref_pod_affinity":
        s["affinity"] = {"podAffinity": {"preferredDuringSchedulingIgnoredDuringExecution": [
            {"weight": 40, "podAffinityTerm": {"labelSelector": {"matchLabels": {"app": "db"}}, "topologyKey": "topology.kubernetes.io/zone"}},
            {"weight": 15, "podAffinityTerm": {"labelSelector": {"matchLabels": {"app": "sim"}}, "topologyKey": "disk"}}]},
            "podAntiAffinity": {"preferredDuringSchedulingIgnoredDuringExecution": [
                {"weight": 25, "podAffinityTerm": {"labelSelector": {"matchLabels": {"app": "sim"}}, "topologyKey": "kubernetes.io/hostname"}}]}}
    elif variant == "owner_default_spread":
        p["metadata"]["ownerReferences"] = [{"apiVersion": "apps/v1", "kind": "ReplicaSet", "name": "sim-rs", "controller": True, "uid": "u"}]
    elif variant == "hostports":
        s["containers"][0]["ports"] = [{"containerPort": 80, "hostPort": 8080}]
    elif variant == "spread_zone":
        s["topologySpreadConstraints"] = [{"maxSkew": 1, "topologyKey": "topology.kubernetes.io/zone", "whenUnsatisfiable": "DoNotSchedule",
                                           "labelSelector": {"matchLabels": {"app": "sim"}}}]
    elif variant == "spread_everything":
        # labelSelector {} = Everything: self-matches (filtering.go:341-344) but countPodsMatchSelector returns 0 for an empty
        # selector (common.go:144-147), so the counts never move and the constraint never blocks (skew = 1 - 0)
        s["topologySpreadConstraints"] = [{"maxSkew": 1, "topologyKey": "topology.kubernetes.io/zone", "whenUnsatisfiable": "DoNotSchedule",
                                           "labelSelector": {}}]
    elif variant == "spread_two":
        s["topologySpreadConstraints"] = [
            {"maxSkew": 2, "topologyKey": "topology.kubernetes.io/zone", "whenUnsatisfiable": "DoNotSchedule", "labelSelector": {"matchLabels": {"app": "sim"}}},
            {"maxSkew": 1, "topologyKey": "kubernetes.io/hostname", "whenUnsatisfiable": "DoNotSchedule", "labelSelector": {"matchLabels": {"app": "web"}},
             "minDomains": 2}]
        s["nodeSelector"] = {"disk": "ssd"}
    elif variant == "anti_hostname":
        s["affinity"] = {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
            {"labelSelector": {"matchLabels": {"app": "sim"}}, "topologyKey": "kubernetes.io/hostname"}]}}
    elif variant == "anti_zone":
        s["affinity"] = {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
            {"labelSelector": {"matchExpressions": [{"key": "app", "operator": "In", "values": ["db"]}]}, "topologyKey": "topology.kubernetes.io/zone"}]}}
    elif variant == "affinity_zone":
        s["affinity"] = {"podAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
            {"labelSelector": {"matchLabels": {"app": "sim"}}, "topologyKey": "topology.kubernetes.io/zone"}]}}
    elif variant == "extended":
        s["containers"][0]["resources"]["requests"]["example.com/foo"] = "2"
        s["containers"][0]["resources"]["requests"]["ephemeral-storage"] = "30Gi"
    elif variant == "best_effort":
        s["containers"][0]["resources"] = {}
    elif variant == "init_overhead":
        s["initContainers"] = [{"name": "i", "image": "x", "resources": {"requests": {"cpu": "1", "memory": "50Mi"}}},
                               {"name": "side", "image": "x", "restartPolicy": "Always", "resources": {"requests": {"cpu": "50m"}}}]
        s["overhead"] = {"cpu": "25m", "memory": "10Mi"}
        s["containers"].append({"name": "c2", "image": "y", "resources": {}})
    elif variant == "never_preempt":
        s["preemptionPolicy"] = "Never"
    return p
