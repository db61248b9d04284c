"""objref.py — OBJECT-LEVEL CPU ORACLE (test infrastructure, NOT the product path).

A direct, deliberately naive restatement of the reference's whole path on Kubernetes objects (dicts as parsed from
JSON/YAML): SyncWithClient -> scheduler cache (NodeInfo, nodeTree order) -> one scheduling cycle per simulated pod
(PreFilter recounts over ALL nodes and pods every cycle, exactly like the reference; Filter plugins in profile order with
string reasons; Score/Normalize; selectHost) -> assume/bind -> report. No dictionary encoding, no bitmasks, no
incremental counters: this is what checks the C++ encoder (cluster-capacity_b200/csrc/host) and, through it, the CUDA
path, on small clusters. Pure-Python loops: keep clusters to a few hundred nodes.

Deterministic contract: percentageOfNodesToScore=100 (every node filtered every cycle), ties ->
first max in nodeTree order. Every function cites the reference file:line it follows
(KS: vendor/k8s.io/kubernetes/pkg/scheduler/, PL: KS:framework/plugins/, CH: vendor/k8s.io/component-helpers/,
 R: the cluster-capacity repo root).
"""
import math
from fractions import Fraction

DEFAULT_MILLI_CPU = 100                 # KS:util/pod_resources.go:29
DEFAULT_MEMORY = 200 * 1024 * 1024      # KS:util/pod_resources.go:31

_SUFFIX = {"n": Fraction(1, 10**9), "u": Fraction(1, 10**6), "m": Fraction(1, 1000), "": Fraction(1), "k": Fraction(10**3),
           "M": Fraction(10**6), "G": Fraction(10**9), "T": Fraction(10**12), "P": Fraction(10**15), "E": Fraction(10**18),
           "Ki": Fraction(2**10), "Mi": Fraction(2**20), "Gi": Fraction(2**30), "Ti": Fraction(2**40), "Pi": Fraction(2**50),
           "Ei": Fraction(2**60)}


def parse_quantity(s):
    """resource.ParseQuantity: exact value as a Fraction (vendor/k8s.io/apimachinery/pkg/api/resource/quantity.go)."""
    s = str(s).strip()
    i = 0
    if s[i] in "+-":
        i += 1
    while i < len(s) and (s[i].isdigit() or s[i] == "."):
        i += 1
    num, suf = s[:i], s[i:]
    if suf[:1] in ("e", "E") and len(suf) > 1 and (suf[1].isdigit() or (suf[1] in "+-" and suf[2:3].isdigit())):
        return Fraction(num) * Fraction(10) ** int(suf[1:])
    return Fraction(num) * _SUFFIX[suf]


def qvalue(q):       # Quantity.Value(): rounds up (quantity.go:812-823)
    return math.ceil(q)


def qmilli(q):       # Quantity.MilliValue(): rounds up (quantity.go:825-834)
    return math.ceil(q * 1000)


def resource_list(d):
    return {k: parse_quantity(v) for k, v in (d or {}).items()}


def _add(dst, src):
    for k, v in src.items():
        dst[k] = dst.get(k, Fraction(0)) + v


def _max(dst, src):
    for k, v in src.items():
        if k not in dst or v > dst[k]:
            dst[k] = v


def pod_requests(pod, use_status=False, skip_pod_level=False, non_missing=None):
    """resourcehelper.PodRequests + AggregateContainerRequests (CH:resource/helpers.go:144-251)."""
    spec = pod.get("spec", {})
    statuses = {}
    if use_status:
        st = pod.get("status", {}) or {}
        for field in ("containerStatuses", "initContainerStatuses"):
            for cs in st.get(field) or []:
                statuses[cs["name"]] = cs
    infeasible = any(c.get("type") == "PodResizePending" and c.get("reason") == "Infeasible"
                     for c in (pod.get("status", {}) or {}).get("conditions") or [])

    def creqs(c, is_init):
        r = resource_list((c.get("resources") or {}).get("requests"))
        restartable = c.get("restartPolicy") == "Always"
        if use_status and (not is_init or restartable):
            cs = statuses.get(c.get("name"))
            if cs is not None and cs.get("resources") is not None:
                m = {}
                if not infeasible:
                    _max(m, r)
                _max(m, resource_list(cs["resources"].get("requests")))
                _max(m, resource_list(cs.get("allocatedResources")))
                r = m
        if non_missing:
            r = dict(r)
            for k, v in non_missing.items():
                if k not in r:
                    r[k] = v
        return r

    reqs = {}
    for c in spec.get("containers") or []:
        _add(reqs, creqs(c, False))
    restartable_sum, init_max = {}, {}
    for c in spec.get("initContainers") or []:
        cr = creqs(c, True)
        if c.get("restartPolicy") == "Always":
            _add(reqs, cr)
            _add(restartable_sum, cr)
            cr = dict(restartable_sum)
        else:
            tmp = {}
            _add(tmp, cr)
            _add(tmp, restartable_sum)
            cr = tmp
        _max(init_max, cr)
    _max(reqs, init_max)
    plr = resource_list((spec.get("resources") or {}).get("requests"))
    if not skip_pod_level and any(k in ("cpu", "memory") or k.startswith("hugepages-") for k in plr):
        for k, v in plr.items():
            if k in ("cpu", "memory") or k.startswith("hugepages-"):
                reqs[k] = v
    _add(reqs, resource_list(spec.get("overhead")))
    return reqs


def is_scalar_resource_name(n):          # KS:util/utils.go:139-143
    native = "/" not in n or "kubernetes.io/" in n
    extended = (not native) and not n.startswith("requests.")
    return extended or n.startswith("hugepages-") or "kubernetes.io/" in n or n.startswith("attachable-volumes-")


class Resource:
    """framework.Resource (KS:framework/types.go:940-1055)."""

    def __init__(self):
        self.cpu = self.mem = self.eph = 0
        self.pods = 0
        self.scalar = {}

    def add(self, rl):                   # types.go:979-1001
        for k, v in rl.items():
            if k == "cpu":
                self.cpu += qmilli(v)
            elif k == "memory":
                self.mem += qvalue(v)
            elif k == "pods":
                self.pods += qvalue(v)
            elif k == "ephemeral-storage":
                self.eph += qvalue(v)
            elif is_scalar_resource_name(k):
                self.scalar[k] = self.scalar.get(k, 0) + qvalue(v)


def calculate_resource(pod):
    """PodInfo.CalculateResource (KS:framework/types.go:700-734): (Resource, Non0CPU, Non0Mem)."""
    req = pod_requests(pod, use_status=True)
    plr = resource_list(((pod.get("spec") or {}).get("resources") or {}).get("requests"))
    pod_level = any(k in ("cpu", "memory") or k.startswith("hugepages-") for k in plr)
    defaults = {"cpu": Fraction(DEFAULT_MILLI_CPU, 1000), "memory": Fraction(DEFAULT_MEMORY)}
    nm = defaults if not pod_level else {k: v for k, v in defaults.items() if k not in req}
    non0 = pod_requests(pod, use_status=True, non_missing=nm) if nm else req
    r = Resource()
    r.add(req)
    return r, qmilli(non0.get("cpu", Fraction(0))), qvalue(non0.get("memory", Fraction(0)))


# ---- selectors --------------------------------------------------------------------------------------------------
def req_matches(r, labels):
    """labels.Requirement.Matches (vendor/k8s.io/apimachinery/pkg/labels/selector.go:246-293)."""
    key, op, values = r["key"], r["operator"], r.get("values") or []
    has = key in labels
    if op == "In":
        return has and labels[key] in values
    if op == "NotIn":
        return (not has) or labels[key] not in values
    if op == "Exists":
        return has
    if op == "DoesNotExist":
        return not has
    if op in ("Gt", "Lt"):
        if not has or len(values) != 1:
            return False
        try:
            a, b = int(labels[key]), int(values[0])
        except ValueError:
            return False
        return a > b if op == "Gt" else a < b
    return False


def label_selector_matches(sel, labels):
    """metav1.LabelSelectorAsSelector + Matches: None -> Nothing, {} -> Everything."""
    if sel is None:
        return False
    for k, v in (sel.get("matchLabels") or {}).items():
        if labels.get(k) != v:
            return False
    for e in sel.get("matchExpressions") or []:
        if not req_matches(e, labels):
            return False
    return True


def label_selector_empty(sel):
    return sel is not None and not (sel.get("matchLabels") or {}) and not (sel.get("matchExpressions") or [])


def tolerates(tol, taint):               # vendor/k8s.io/api/core/v1/toleration.go:38-57
    if tol.get("effect") and tol["effect"] != taint.get("effect", ""):
        return False
    if tol.get("key") and tol["key"] != taint.get("key", ""):
        return False
    op = tol.get("operator") or "Equal"
    if op == "Equal":
        return (tol.get("value") or "") == (taint.get("value") or "")
    return op == "Exists"


def tolerations_tolerate(tols, taint):
    return any(tolerates(t, taint) for t in tols or [])


def node_term_matches(term, node):       # CH:scheduling/corev1/nodeaffinity/nodeaffinity.go:190-202
    labels = (node.get("metadata") or {}).get("labels") or {}
    for e in term.get("matchExpressions") or []:
        if not req_matches(e, labels):
            return False
    name = node["metadata"].get("name", "")
    if (term.get("matchFields") or []) and name:
        for e in term["matchFields"]:
            if not req_matches(e, {"metadata.name": name}):
                return False
    return True


def required_node_affinity_match(pod, node):   # nodeaffinity.go:306-332
    labels = (node.get("metadata") or {}).get("labels") or {}
    sel = (pod.get("spec") or {}).get("nodeSelector")
    if sel:
        for k, v in sel.items():
            if labels.get(k) != v:
                return False
    req = (((pod["spec"].get("affinity") or {}).get("nodeAffinity") or {}).get("requiredDuringSchedulingIgnoredDuringExecution"))
    if req is not None:
        terms = [t for t in (req.get("nodeSelectorTerms") or []) if (t.get("matchExpressions") or t.get("matchFields"))]
        return any(node_term_matches(t, node) for t in terms)
    return True


def affinity_terms(pod, kind, required=True):
    aff = ((pod.get("spec") or {}).get("affinity") or {}).get(kind) or {}
    if required:
        terms = aff.get("requiredDuringSchedulingIgnoredDuringExecution") or []
    else:
        terms = [t["podAffinityTerm"] for t in aff.get("preferredDuringSchedulingIgnoredDuringExecution") or []]
    out = []
    ns = (pod.get("metadata") or {}).get("namespace") or "default"
    for t in terms:
        names = set(t.get("namespaces") or [])
        if not names and t.get("namespaceSelector") is None:   # KS:framework/types.go:927-935
            names = {ns}
        out.append({"namespaces": names, "selector": t.get("labelSelector"), "nsSelector": t.get("namespaceSelector"),
                    "topologyKey": t.get("topologyKey", "")})
    if not required:
        for o, wt in zip(out, aff.get("preferredDuringSchedulingIgnoredDuringExecution") or []):
            o["weight"] = wt.get("weight", 0)
    return out


def go_log(x):
    """Go's math.Log on amd64 (go/src/math/log.go:80-129, the FreeBSD e_log.c port; no FMA)."""
    Ln2Hi, Ln2Lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    L1, L2, L3, L4 = 6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01
    L5, L6, L7 = 1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01
    f1, ki = math.frexp(x)
    if f1 < math.sqrt(2) / 2:
        f1 *= 2
        ki -= 1
    f = f1 - 1
    k = float(ki)
    s = f / (2 + f)
    s2 = s * s
    s4 = s2 * s2
    t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)))
    t2 = s4 * (L2 + s4 * (L4 + s4 * L6))
    R = t1 + t2
    hfsq = 0.5 * f * f
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f)


def go_round(x):
    """math.Round: half away from zero."""
    t = math.trunc(x)
    if abs(x - t) >= 0.5:
        t += math.copysign(1.0, x)
    return int(t)


def normalized_image_name(name):          # PL:imagelocality/image_locality.go:124-131
    if name.rfind(":") <= name.rfind("/"):
        name += ":latest"
    return name


def term_matches(term, pod, ns_labels):  # AffinityTerm.Matches (KF:types.go:379-384)
    ns = (pod.get("metadata") or {}).get("namespace") or "default"
    if ns in term["namespaces"] or label_selector_matches(term["nsSelector"], ns_labels or {}):
        return label_selector_matches(term["selector"], (pod.get("metadata") or {}).get("labels") or {})
    return False


def host_ports(pod):                      # KS:util/utils.go:175-210
    out = []
    spec = pod.get("spec") or {}
    for c in spec.get("initContainers") or []:
        if c.get("restartPolicy") == "Always":
            out += [p for p in c.get("ports") or [] if (p.get("hostPort") or 0) > 0]
    for c in spec.get("containers") or []:
        out += [p for p in c.get("ports") or [] if (p.get("hostPort") or 0) > 0]
    return out


def ports_conflict(used, want):           # HostPortInfo.CheckConflict (KF:types.go:499-528)
    for w in want:
        wip, wproto, wport = w.get("hostIP") or "0.0.0.0", w.get("protocol") or "TCP", w["hostPort"]
        for (uip, uproto, uport) in used:
            if uproto != wproto or uport != wport:
                continue
            if wip == "0.0.0.0" or uip == "0.0.0.0" or uip == wip:
                return True
    return False


def zone_key(node):                        # CH:node/topology/helpers.go:31-58
    labels = (node.get("metadata") or {}).get("labels") or {}
    zone = labels.get("failure-domain.beta.kubernetes.io/zone", labels.get("topology.kubernetes.io/zone", ""))
    region = labels.get("failure-domain.beta.kubernetes.io/region", labels.get("topology.kubernetes.io/region", ""))
    if not region and not zone:
        return ""
    return region + ":\x00:" + zone


class NodeInfo:
    def __init__(self, node):
        self.node = node
        self.name = node["metadata"]["name"]
        self.labels = (node["metadata"].get("labels") or {})
        self.alloc = Resource()
        self.alloc.add(resource_list((node.get("status") or {}).get("allocatable")))
        self.requested = Resource()
        self.nz_cpu = self.nz_mem = 0
        self.pods = []
        self.used_ports = []

    def add_pod(self, pod):                # AddPodInfo/update(+1): KS:framework/types.go:333-343,409-427
        r, n0c, n0m = calculate_resource(pod)
        self.requested.cpu += r.cpu
        self.requested.mem += r.mem
        self.requested.eph += r.eph
        for k, v in r.scalar.items():
            self.requested.scalar[k] = self.requested.scalar.get(k, 0) + v
        self.nz_cpu += n0c
        self.nz_mem += n0m
        self.pods.append(pod)
        for p in host_ports(pod):
            self.used_ports.append((p.get("hostIP") or "0.0.0.0", p.get("protocol") or "TCP", p["hostPort"]))


class Simulator:
    """ClusterCapacity (R:pkg/framework/simulator.go) on plain objects, canonical deterministic mode."""

    def __init__(self, pod, max_pods=0, exclude_nodes=()):
        self.template = pod
        self.max_pods = max_pods
        self.exclude = set(exclude_nodes)
        self.pods_status = []      # node name per scheduled pod
        self.stop_reason = None

    def sync(self, nodes, pods, namespaces=(), services=(), rcs=(), replicasets=(), statefulsets=()):
        """SyncWithClient (simulator.go:176-295) + the informer-driven cache build (KS:backend/cache/node_tree.go:51-143)."""
        zones, tree, seen = [], {}, set()
        for n in nodes:
            name = n["metadata"]["name"]
            if name in self.exclude or name in seen:
                continue
            seen.add(name)
            z = zone_key(n)
            if z not in tree:
                zones.append(z)
                tree[z] = []
            tree[z].append(n)
        order, i = [], 0
        while len(order) < len(seen):
            for z in zones:
                if i < len(tree[z]):
                    order.append(tree[z][i])
            i += 1
        self.infos = [NodeInfo(n) for n in order]
        by_name = {ni.name: ni for ni in self.infos}
        for p in pods:
            if (p.get("status") or {}).get("phase") in ("Succeeded", "Failed"):     # simulator.go:196
                continue
            nn = (p.get("spec") or {}).get("nodeName")
            if nn and nn in by_name:
                by_name[nn].add_pod(p)
        self.ns_labels = {ns["metadata"]["name"]: (ns["metadata"].get("labels") or {}) for ns in namespaces}
        self.services, self.rcs, self.replicasets, self.statefulsets = list(services), list(rcs), list(replicasets), list(statefulsets)
        # image states (KS:backend/cache/cache.go:680-703): size as first registered, number of nodes holding the name
        self.image_states = {}
        for ni in self.infos:
            for im in (ni.node.get("status") or {}).get("images") or []:
                for name in im.get("names") or []:
                    stt = self.image_states.setdefault(name, [im.get("sizeBytes", 0), set()])
                    stt[1].add(ni.name)

    def _default_selector(self, pod):
        """helper.DefaultSelector (PL:helper/spread.go:40-113) as a list of requirements (None = selects nothing... never here)."""
        ns = pod["metadata"].get("namespace") or "default"
        labels = pod["metadata"].get("labels") or {}
        lset = {}
        for svc in self.services:
            if (svc["metadata"].get("namespace") or "default") != ns:
                continue
            sel = (svc.get("spec") or {}).get("selector")
            if sel is None:
                continue
            if all(labels.get(k) == v for k, v in sel.items()):
                lset.update(sel)
        reqs = {"matchLabels": dict(lset), "matchExpressions": []}
        owner = next((o for o in pod["metadata"].get("ownerReferences") or [] if o.get("controller")), None)
        if owner is None:
            return reqs
        def find(lst):
            return next((o for o in lst if (o["metadata"].get("namespace") or "default") == ns and o["metadata"]["name"] == owner.get("name")), None)
        if owner.get("kind") == "ReplicationController" and owner.get("apiVersion") == "v1":
            rc = find(self.rcs)
            if rc is not None:
                reqs["matchLabels"].update((rc.get("spec") or {}).get("selector") or {})
        elif owner.get("kind") in ("ReplicaSet", "StatefulSet") and owner.get("apiVersion") == "apps/v1":
            o = find(self.replicasets if owner["kind"] == "ReplicaSet" else self.statefulsets)
            if o is not None and (o.get("spec") or {}).get("selector") is not None:
                sel = o["spec"]["selector"]
                extra = [{"key": k, "operator": "In", "values": [v]} for k, v in (sel.get("matchLabels") or {}).items()]
                reqs["matchExpressions"] = extra + list(sel.get("matchExpressions") or [])
        return reqs

    def _soft_constraints(self, pod):
        """initPreScoreState's constraint list (PL:podtopologyspread/scoring.go:60-82; plugin.go:48-59; common.go:64-127)."""
        spec = pod["spec"]
        labels = pod["metadata"].get("labels") or {}
        out = []
        if spec.get("topologySpreadConstraints"):
            for c in spec["topologySpreadConstraints"]:
                if c.get("whenUnsatisfiable") != "ScheduleAnyway":
                    continue
                out.append({"key": c["topologyKey"], "maxSkew": c.get("maxSkew", 1), "selector": c.get("labelSelector"),
                            "mlk": {k: labels[k] for k in c.get("matchLabelKeys") or [] if k in labels},
                            "affPolicy": c.get("nodeAffinityPolicy") or "Honor", "taintPolicy": c.get("nodeTaintsPolicy") or "Ignore"})
            return out, True
        sel = self._default_selector(pod)
        if not sel["matchLabels"] and not sel["matchExpressions"]:
            return [], False
        for key, skew in (("kubernetes.io/hostname", 3), ("topology.kubernetes.io/zone", 5)):
            out.append({"key": key, "maxSkew": skew, "selector": sel, "mlk": {}, "affPolicy": "Honor", "taintPolicy": "Ignore"})
        return out, False

    @staticmethod
    def _csel_matches(c, lbls):
        if c["selector"] is None:
            return False
        for k, v in c["mlk"].items():
            if lbls.get(k) != v:
                return False
        return label_selector_matches(c["selector"], lbls)

    def _count_matching(self, c, ni, ns):          # countPodsMatchSelector (common.go:144-158)
        if label_selector_empty(c["selector"]) and not c["mlk"]:
            return 0
        return sum(1 for p in ni.pods if p["metadata"].get("deletionTimestamp") is None and
                   (p["metadata"].get("namespace") or "default") == ns and self._csel_matches(c, p["metadata"].get("labels") or {}))

    def _pts_scores(self, pod, feasible):
        """PodTopologySpread PreScore + Score + NormalizeScore (PL:podtopologyspread/scoring.go:60-265). None = Skip."""
        cons, require_all = self._soft_constraints(pod)
        if not cons:
            return None
        spec = pod["spec"]
        ns = pod["metadata"].get("namespace") or "default"
        ignored, pair_counts, topo_size = set(), [dict() for _ in cons], [0] * len(cons)
        for ni in feasible:
            if require_all and any(c["key"] not in ni.labels for c in cons):
                ignored.add(ni.name)
                continue
            for i, c in enumerate(cons):
                if c["key"] == "kubernetes.io/hostname":
                    continue
                v = ni.labels.get(c["key"], "")
                if v not in pair_counts[i]:
                    pair_counts[i][v] = 0
                    topo_size[i] += 1
        weights = []
        for i, c in enumerate(cons):
            sz = topo_size[i]
            if c["key"] == "kubernetes.io/hostname":
                sz = len(feasible) - len(ignored)
            weights.append(go_log(float(sz + 2)))
        for ni in self.infos:
            if require_all and any(c["key"] not in ni.labels for c in cons):
                continue
            for i, c in enumerate(cons):
                if c["affPolicy"] == "Honor" and not required_node_affinity_match(pod, ni.node):
                    continue
                if c["taintPolicy"] == "Honor" and any(t.get("effect") in ("NoSchedule", "NoExecute") and not tolerations_tolerate(spec.get("tolerations"), t)
                                                        for t in (ni.node.get("spec") or {}).get("taints") or []):
                    continue
                v = ni.labels.get(c["key"], "")
                if v not in pair_counts[i]:
                    continue
                pair_counts[i][v] += self._count_matching(c, ni, ns)
        raw = []
        for ni in feasible:
            if ni.name in ignored:
                raw.append(None)
                continue
            sc = 0.0
            for i, c in enumerate(cons):
                if c["key"] in ni.labels:
                    if c["key"] == "kubernetes.io/hostname":
                        cnt = self._count_matching(c, ni, ns)
                    else:
                        cnt = pair_counts[i][ni.labels[c["key"]]]
                    sc += float(cnt) * weights[i] + float(c["maxSkew"] - 1)
            raw.append(go_round(sc))
        vals = [r for r in raw if r is not None]
        mn = min(vals) if vals else (1 << 63) - 1
        mx = max(vals + [0])
        out = []
        for r in raw:
            if r is None:
                out.append(0)
            elif mx == 0:
                out.append(100)
            else:
                out.append(100 * (mx + mn - r) // mx)
        return out

    def _ipa_scores(self, pod, feasible, hard_weight=1):
        """InterPodAffinity PreScore + Score + NormalizeScore (PL:interpodaffinity/scoring.go:51-295). None = Skip."""
        ns = pod["metadata"].get("namespace") or "default"
        nsl = self.ns_labels.get(ns, {})
        paff = affinity_terms(pod, "podAffinity", required=False)
        panti = affinity_terms(pod, "podAntiAffinity", required=False)
        for t in paff + panti:       # mergeAffinityTermNamespacesIfNotEmpty
            if t["nsSelector"] is not None and not label_selector_empty(t["nsSelector"]):
                for name, nl in self.ns_labels.items():
                    if label_selector_matches(t["nsSelector"], nl):
                        t["namespaces"].add(name)
        topo = {}

        def process(term, weight, target, nslabels, node_labels, mult):
            if term_matches(term, target, nslabels) and term["topologyKey"] in node_labels:
                d = topo.setdefault(term["topologyKey"], {})
                v = node_labels[term["topologyKey"]]
                d[v] = d.get(v, 0) + weight * mult
        for ni in self.infos:
            if not ni.labels:
                continue
            for p in ni.pods:
                for t in paff:
                    process(t, t["weight"], p, None, ni.labels, 1)
                for t in panti:
                    process(t, t["weight"], p, None, ni.labels, -1)
                if hard_weight > 0:
                    for t in affinity_terms(p, "podAffinity"):
                        process(t, hard_weight, pod, nsl, ni.labels, 1)
                for t in affinity_terms(p, "podAffinity", required=False):
                    process(t, t["weight"], pod, nsl, ni.labels, 1)
                for t in affinity_terms(p, "podAntiAffinity", required=False):
                    process(t, t["weight"], pod, nsl, ni.labels, -1)
        if not topo:
            return None
        raw = [sum(vals.get(ni.labels[k], 0) for k, vals in topo.items() if k in ni.labels) for ni in feasible]
        mn, mx = min(raw), max(raw)
        return [int(100.0 * (float(r - mn) / float(mx - mn))) if mx > mn else 0 for r in raw]

    def _image_score(self, pod, ni):
        """ImageLocality.Score (PL:imagelocality/image_locality.go:54-122)."""
        spec = pod["spec"]
        conts = list(spec.get("initContainers") or []) + list(spec.get("containers") or [])
        here = set()
        for im in (ni.node.get("status") or {}).get("images") or []:
            here.update(im.get("names") or [])
        total = 0
        for c in conts:
            name = normalized_image_name(c.get("image", ""))
            if name in here:
                size, nodes = self.image_states[name]
                total += int(float(size) * (float(len(nodes)) / float(len(self.infos))))
        mb = 1024 * 1024
        lo, hi = 23 * mb, 1000 * mb * len(conts)
        total = min(max(total, lo), hi)
        return 100 * (total - lo) // (hi - lo)

    # ---- one scheduling cycle -----------------------------------------------------------------------------------
    def _prefilter(self, pod):
        spec = pod["spec"]
        ns = pod["metadata"].get("namespace") or "default"
        labels = pod["metadata"].get("labels") or {}
        st = {}
        # NodeResourcesFit (PL:noderesources/fit.go:224-233)
        fit = Resource()
        fit.add(pod_requests(pod))
        st["fit"] = fit
        # PodTopologySpread (PL:podtopologyspread/filtering.go:235-308)
        cons = []
        for c in spec.get("topologySpreadConstraints") or []:
            if (c.get("whenUnsatisfiable") or "DoNotSchedule") != "DoNotSchedule":
                continue
            sel = c.get("labelSelector")
            mlk = {k: labels[k] for k in c.get("matchLabelKeys") or [] if k in labels}
            cons.append({"key": c["topologyKey"], "maxSkew": c.get("maxSkew", 1), "selector": sel, "mlk": mlk,
                         "minDomains": c.get("minDomains") or 1, "affPolicy": c.get("nodeAffinityPolicy") or "Honor",
                         "taintPolicy": c.get("nodeTaintsPolicy") or "Ignore"})

        def csel_matches(c, lbls):
            if c["selector"] is None and not c["mlk"]:
                return False
            if c["selector"] is None:
                return False if not c["mlk"] else False
            for k, v in c["mlk"].items():
                if lbls.get(k) != v:
                    return False
            return label_selector_matches(c["selector"], lbls)

        def csel_empty(c):
            return label_selector_empty(c["selector"]) and not c["mlk"]

        counts = [dict() for _ in cons]
        for ni in self.infos:
            if any(c["key"] not in ni.labels for c in cons):
                continue
            for k, c in enumerate(cons):
                if c["affPolicy"] == "Honor" and not required_node_affinity_match(pod, ni.node):
                    continue
                if c["taintPolicy"] == "Honor" and any(t.get("effect") in ("NoSchedule", "NoExecute") and not tolerations_tolerate(spec.get("tolerations"), t)
                                                        for t in (ni.node.get("spec") or {}).get("taints") or []):
                    continue
                cnt = 0
                if not csel_empty(c):
                    for p in ni.pods:
                        if p["metadata"].get("deletionTimestamp") is None and (p["metadata"].get("namespace") or "default") == ns and \
                                csel_matches(c, p["metadata"].get("labels") or {}):
                            cnt += 1
                v = ni.labels[c["key"]]
                counts[k][v] = counts[k].get(v, 0) + cnt
        st["pts"] = (cons, counts, [csel_matches(c, labels) for c in cons])
        # InterPodAffinity (PL:interpodaffinity/filtering.go:204-309)
        aff = affinity_terms(pod, "podAffinity")
        anti = affinity_terms(pod, "podAntiAffinity")
        for t in aff + anti:       # mergeAffinityTermNamespacesIfNotEmpty
            if t["nsSelector"] is not None and not label_selector_empty(t["nsSelector"]):
                for name, nl in self.ns_labels.items():
                    if label_selector_matches(t["nsSelector"], nl):
                        t["namespaces"].add(name)
        nsl = self.ns_labels.get(ns, {})
        existing_anti, aff_counts, anti_counts = {}, {}, {}
        for ni in self.infos:
            for p in ni.pods:
                for t in affinity_terms(p, "podAntiAffinity"):
                    if term_matches(t, pod, nsl) and t["topologyKey"] in ni.labels:
                        k = (t["topologyKey"], ni.labels[t["topologyKey"]])
                        existing_anti[k] = existing_anti.get(k, 0) + 1
                if aff and all(term_matches(t, p, None) for t in aff):
                    for t in aff:
                        if t["topologyKey"] in ni.labels:
                            k = (t["topologyKey"], ni.labels[t["topologyKey"]])
                            aff_counts[k] = aff_counts.get(k, 0) + 1
                for t in anti:
                    if term_matches(t, p, None) and t["topologyKey"] in ni.labels:
                        k = (t["topologyKey"], ni.labels[t["topologyKey"]])
                        anti_counts[k] = anti_counts.get(k, 0) + 1
        st["ipa"] = (aff, anti, existing_anti, aff_counts, anti_counts,
                     bool(aff) and all(term_matches(t, pod, None) for t in aff))
        # NodeAffinity PreFilterResult (PL:nodeaffinity/node_affinity.go:147-194)
        st["node_names"] = None
        req = ((spec.get("affinity") or {}).get("nodeAffinity") or {}).get("requiredDuringSchedulingIgnoredDuringExecution")
        if req is not None and (req.get("nodeSelectorTerms") or []):
            names, all_pin = set(), True
            for t in req["nodeSelectorTerms"]:
                tn = None
                for r in t.get("matchFields") or []:
                    if r["key"] == "metadata.name" and r["operator"] == "In":
                        s = set(r.get("values") or [])
                        tn = s if tn is None else (tn & s)
                if tn is None:
                    all_pin = False
                    break
                names |= tn
            if all_pin:
                st["node_names"] = names
        return st

    def _filter(self, pod, ni, st):
        """RunFilterPlugins in default order (KS:apis/config/v1/default_plugins.go:33-53). Returns (code, reasons)."""
        spec = pod["spec"]
        node = ni.node
        tols = spec.get("tolerations") or []
        if st["node_names"] is not None and ni.name not in st["node_names"]:
            return "U2", ["node(s) didn't satisfy plugin(s) [NodeAffinity]"]
        if (node.get("spec") or {}).get("unschedulable") and not tolerations_tolerate(tols, {"key": "node.kubernetes.io/unschedulable", "effect": "NoSchedule"}):
            return "U2", ["node(s) were unschedulable"]
        if spec.get("nodeName") and spec["nodeName"] != ni.name:
            return "U2", ["node(s) didn't match the requested node name"]
        for t in (node.get("spec") or {}).get("taints") or []:
            if t.get("effect") in ("NoSchedule", "NoExecute") and not tolerations_tolerate(tols, t):
                return "U2", ["node(s) had untolerated taint {%s: %s}" % (t.get("key", ""), t.get("value") or "")]
        has_aff = ((spec.get("affinity") or {}).get("nodeAffinity") or {}).get("requiredDuringSchedulingIgnoredDuringExecution") is not None
        if (has_aff or spec.get("nodeSelector") is not None) and not required_node_affinity_match(pod, node):
            return "U2", ["node(s) didn't match Pod's node affinity/selector"]
        want = host_ports(pod)
        if want and ports_conflict(ni.used_ports, want):
            return "U1", ["node(s) didn't have free ports for the requested pod ports"]
        # NodeResourcesFit (PL:noderesources/fit.go:564-660)
        fit = st["fit"]
        reasons, unres = [], False
        if len(ni.pods) + 1 > ni.alloc.pods:
            reasons.append("Too many pods")
        if not (fit.cpu == 0 and fit.mem == 0 and fit.eph == 0 and not fit.scalar):
            if fit.cpu > 0 and fit.cpu > ni.alloc.cpu - ni.requested.cpu:
                reasons.append("Insufficient cpu"); unres |= fit.cpu > ni.alloc.cpu
            if fit.mem > 0 and fit.mem > ni.alloc.mem - ni.requested.mem:
                reasons.append("Insufficient memory"); unres |= fit.mem > ni.alloc.mem
            if fit.eph > 0 and fit.eph > ni.alloc.eph - ni.requested.eph:
                reasons.append("Insufficient ephemeral-storage"); unres |= fit.eph > ni.alloc.eph
            for k, q in fit.scalar.items():
                if q == 0:
                    continue
                if q > ni.alloc.scalar.get(k, 0) - ni.requested.scalar.get(k, 0):
                    reasons.append("Insufficient %s" % k); unres |= q > ni.alloc.scalar.get(k, 0)
        if reasons:
            return ("U2" if unres else "U1"), reasons
        cons, counts, selfm = st["pts"]
        for k, c in enumerate(cons):      # PL:podtopologyspread/filtering.go:311-356
            if c["key"] not in ni.labels:
                return "U2", ["node(s) didn't match pod topology spread constraints (missing required label)"]
            mn = min(counts[k].values()) if counts[k] else 2**31 - 1
            if len(counts[k]) < c["minDomains"]:
                mn = 0
            skew = counts[k].get(ni.labels[c["key"]], 0) + (1 if selfm[k] else 0) - mn
            if skew > c["maxSkew"]:
                return "U1", ["node(s) didn't match pod topology spread constraints"]
        aff, anti, existing_anti, aff_counts, anti_counts, self_all = st["ipa"]
        pods_exist = True
        for t in aff:                     # satisfyPodAffinity (filtering.go:382-408)
            if t["topologyKey"] not in ni.labels:
                return "U2", ["node(s) didn't match pod affinity rules"]
            if aff_counts.get((t["topologyKey"], ni.labels[t["topologyKey"]]), 0) <= 0:
                pods_exist = False
        if not pods_exist and not (len(aff_counts) == 0 and self_all):
            return "U2", ["node(s) didn't match pod affinity rules"]
        if anti_counts:
            for t in anti:
                if t["topologyKey"] in ni.labels and anti_counts.get((t["topologyKey"], ni.labels[t["topologyKey"]]), 0) > 0:
                    return "U1", ["node(s) didn't match pod anti-affinity rules"]
        if existing_anti:
            for k, v in ni.labels.items():
                if existing_anti.get((k, v), 0) > 0:
                    return "U1", ["node(s) didn't satisfy existing pods anti-affinity rules"]
        return "OK", []

    def _score(self, pod, ni):
        """LeastAllocated + BalancedAllocation (PL:noderesources/least_allocated.go:30-61, balanced_allocation.go:146-180)."""
        defaults = {"cpu": Fraction(DEFAULT_MILLI_CPU, 1000), "memory": Fraction(DEFAULT_MEMORY)}
        lr = pod_requests(pod, skip_pod_level=True, non_missing=defaults)
        lc, lm = qmilli(lr.get("cpu", Fraction(0))), qvalue(lr.get("memory", Fraction(0)))
        score, wsum = 0, 0
        for alloc, req in ((ni.alloc.cpu, ni.nz_cpu + lc), (ni.alloc.mem, ni.nz_mem + lm)):
            if alloc == 0:
                continue
            s = 0 if req > alloc else ((alloc - req) * 100) // alloc
            score += s
            wsum += 1
        least = score // wsum if wsum else 0
        br = pod_requests(pod, use_status=True)
        bc, bm = qmilli(br.get("cpu", Fraction(0))), qvalue(br.get("memory", Fraction(0)))
        bal = None
        if not (bc == 0 and bm == 0):
            fr = []
            for alloc, req in ((ni.alloc.cpu, ni.requested.cpu + bc), (ni.alloc.mem, ni.requested.mem + bm)):
                if alloc == 0:
                    continue
                f = float(req) / float(alloc)
                fr.append(min(f, 1.0))
            std = abs((fr[0] - fr[1]) / 2) if len(fr) == 2 else 0.0
            bal = int((1 - std) * 100.0)
        return least, bal

    def run(self):
        k = 0
        # one podspec, or a list simulated round-robin (the roadmap's "list of pods", README.md:305-306): pod k is a clone of
        # template k % T, the index parsePodsReview already uses (report.go:160)
        templates = self.template if isinstance(self.template, (list, tuple)) else [self.template]
        while True:
            pod = templates[k % len(templates)]
            tols_prefer = [t for t in (pod["spec"].get("tolerations") or []) if not t.get("effect") or t["effect"] == "PreferNoSchedule"]
            if not self.infos:
                self.stop_reason = "Unschedulable: no nodes available to schedule pods"
                return
            st = self._prefilter(pod)
            if st["node_names"] is not None and not st["node_names"]:
                n = len(self.infos)
                self.stop_reason = ("Unschedulable: 0/%d nodes are available: node(s) didn't match Pod's node affinity/selector. "
                                    "preemption: 0/%d nodes are available: %d Preemption is not helpful for scheduling." % (n, n, n))
                return
            feasible, hist, n_unsched = [], {}, 0
            for ni in self.infos:
                code, reasons = self._filter(pod, ni, st)
                if code == "OK":
                    feasible.append(ni)
                else:
                    for r in reasons:
                        hist[r] = hist.get(r, 0) + 1
                    if code == "U1":
                        n_unsched += 1
            if not feasible:
                n = len(self.infos)

                def body(h):
                    strs = sorted("%d %s" % (v, r) for r, v in h.items() if v)
                    return "0/%d nodes are available:" % n + ((" " + ", ".join(strs) + ".") if strs else "")
                if pod["spec"].get("preemptionPolicy") == "Never":
                    post = "not eligible due to preemptionPolicy=Never."
                else:
                    post = body({"No preemption victims found for incoming pod": n_unsched,
                                 "Preemption is not helpful for scheduling": n - n_unsched})
                self.stop_reason = "Unschedulable: " + body(hist) + " preemption: " + post
                return
            # prioritizeNodes + selectHost (KS:schedule_one.go:776-941): first max in node order
            pref = [(t.get("weight", 0), t.get("preference") or {}) for t in
                    (((pod["spec"].get("affinity") or {}).get("nodeAffinity") or {}).get("preferredDuringSchedulingIgnoredDuringExecution") or [])]
            pref = [(w, pr) for w, pr in pref if w != 0 and (pr.get("matchExpressions") or pr.get("matchFields"))]   # nodeaffinity.go:118-139
            na_raw = [sum(w for w, pr in pref if node_term_matches(pr, ni.node)) for ni in feasible]
            na_max = max(na_raw) if na_raw else 0
            raws = [sum(1 for t in (ni.node.get("spec") or {}).get("taints") or []
                        if t.get("effect") == "PreferNoSchedule" and not tolerations_tolerate(tols_prefer, t)) for ni in feasible]
            mx = max(raws)
            pts = self._pts_scores(pod, feasible)
            ipa = self._ipa_scores(pod, feasible)
            best, best_ni = None, None
            for idx, (ni, raw) in enumerate(zip(feasible, raws)):
                least, bal = self._score(pod, ni)
                tt = 100 if mx == 0 else 100 - (100 * raw // mx)
                total = 3 * tt + least + (bal if bal is not None else 0)
                if pref:      # NodeAffinity score, weight 2 (node_affinity.go:241-290; helper/normalize_score.go:28-56)
                    total += 2 * (na_raw[idx] if na_max == 0 else 100 * na_raw[idx] // na_max)
                if pts is not None:      # weight 2
                    total += 2 * pts[idx]
                if ipa is not None:      # weight 2
                    total += 2 * ipa[idx]
                total += self._image_score(pod, ni)     # weight 1
                if best is None or total > best:
                    best, best_ni = total, ni
            clone = {"metadata": dict(pod["metadata"]), "spec": dict(pod["spec"]), "status": {}}
            clone["spec"]["nodeName"] = best_ni.name
            best_ni.add_pod(clone)
            self.pods_status.append(best_ni.name)
            k += 1
            if self.max_pods > 0 and k >= self.max_pods:
                self.stop_reason = "LimitReached: Maximum number of pods simulated: %d" % self.max_pods
                return

    def report(self):
        """GetReport (R:pkg/framework/report.go:100-233), the fields parity is judged on."""
        order, counts = [], {}
        for n in self.pods_status:
            if n not in counts:
                order.append(n)
                counts[n] = 0
            counts[n] += 1
        colon = self.stop_reason.index(":")
        return {"replicas": len(self.pods_status), "failType": self.stop_reason[:colon],
                "failMessage": self.stop_reason[colon + 1:].strip(" "),
                "replicasOnNodes": [{"nodeName": n, "replicas": counts[n]} for n in order]}
