"""Seeded synthetic snapshots for BASELINE.json's configs C1..C5, generated directly in the flat
SoA form of include/ccsim.h (numpy PCG64). The object-level path (Node/Pod JSON -> cchost encoder) is exercised on
smaller clusters by tests/; these generators exist so that the 10k..1M-node workloads do not need 100 MB of JSON.
"""
import importlib

import numpy as np

abi = importlib.import_module("cluster-capacity_b200._abi")

GiB = 1 << 30
MiB = 1 << 20


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def c1():
    """README demo: 4 nodes x 2 CPU / 4 GiB / 110 pods, examples/pod.yaml (150m / 100Mi) -> 52 (README.md:44-66)."""
    snap = abi.Snapshot(4, np.full(4, 2000), np.full(4, 4 * GiB), np.full(4, 110),
                        names=["kube-node-%d" % (i + 1) for i in range(4)])
    return snap, [abi.default_template(150, 100 * MiB)], []


def _c2_nodes(n, rng):
    cores = rng.choice([4, 8, 16, 32, 64], size=n)
    alloc_cpu = cores.astype(np.int64) * 1000
    alloc_mem = cores.astype(np.int64) * rng.choice([2, 4, 8], size=n) * GiB
    alloc_pods = np.full(n, 110, np.int32)
    req_cpu = (rng.random(n) * 0.7 * alloc_cpu / 10).astype(np.int64) * 10
    req_mem = (rng.random(n) * 0.7 * alloc_mem / MiB).astype(np.int64) * MiB
    npods = rng.integers(0, 61, size=n).astype(np.int32)
    return alloc_cpu, alloc_mem, alloc_pods, req_cpu, req_mem, npods


def c2(n=10_000, seed=1, fit_only=True):
    """10k nodes, single podspec cpu=150m mem=100Mi, NodeResourcesFit only (+LeastAllocated, BalancedAllocation)."""
    rng = _rng(seed)
    a_cpu, a_mem, a_pods, r_cpu, r_mem, npods = _c2_nodes(n, rng)
    snap = abi.Snapshot(n, a_cpu, a_mem, a_pods, req_cpu=r_cpu, req_mem=r_mem, npods=npods)
    return snap, [abi.default_template(150, 100 * MiB, fit_only=fit_only)], []


def c3(n=50_000, seed=2, prefer_taints=False):
    """50k nodes, nodeSelector on 2 keys (~6% match) + 3 tolerations, 3 NoSchedule taints on 10% of nodes each,
    1% unschedulable, full default Filter set. With prefer_taints=True two PreferNoSchedule taints (ids 3,4) are added
    so that TaintToleration's normalized score is not constant."""
    rng = _rng(seed)
    a_cpu, a_mem, a_pods, r_cpu, r_mem, npods = _c2_nodes(n, rng)
    taint = np.zeros(n, np.uint64)
    for tid in range(3):
        taint |= (rng.random(n) < 0.10).astype(np.uint64) << np.uint64(tid)
    # one more NoSchedule taint (id 5) that the pod does NOT tolerate, on 2% of nodes
    taint |= (rng.random(n) < 0.02).astype(np.uint64) << np.uint64(5)
    nosched = 0b100111
    prefer = 0
    if prefer_taints:
        taint |= (rng.random(n) < 0.30).astype(np.uint64) << np.uint64(3)
        taint |= (rng.random(n) < 0.20).astype(np.uint64) << np.uint64(4)
        prefer = 0b011000
    taint |= (rng.random(n) < 0.01).astype(np.uint64) << np.uint64(abi.TAINT_UNSCHEDULABLE_BIT)
    # static bits 0,1: the two nodeSelector requirements (key==value), each true on 25% of nodes
    static = (rng.random(n) < 0.25).astype(np.uint64) | ((rng.random(n) < 0.25).astype(np.uint64) << np.uint64(1))
    lists = []
    for i in range(n):
        ids = [t for t in range(6) if (int(taint[i]) >> t) & 1]
        lists.append(ids)
    snap = abi.Snapshot(n, a_cpu, a_mem, a_pods, req_cpu=r_cpu, req_mem=r_mem, npods=npods,
                        taint_mask=taint.reshape(1, n), taint_nosched=[nosched], taint_prefer=[prefer],
                        static_mask=static.reshape(1, n), taint_lists=lists)
    t = abi.default_template(150, 100 * MiB)
    t.flags |= abi.TF_HAS_NODE_SELECTOR
    t.sel_mask[0] = 0b11
    t.tol_nosched[0] = 0b000111
    return snap, [t], []


def _c4_raw(n, seed, n_existing, zones, racks, regions, match_frac):
    """The random draws behind C4: node capacities and topology, and the existing pods (node, cpu, memory; the first n_match carry app=sim)."""
    rng = _rng(seed)
    a_cpu, a_mem, a_pods, _, _, _ = _c2_nodes(n, rng)
    rack = rng.integers(0, racks, size=n).astype(np.int32)
    zone = (rack % zones).astype(np.int32)
    region = (zone % regions).astype(np.int32)
    n_match = int(n_existing * match_frac)
    # matching pods: pod j -> rack j % racks, then a random node of that rack
    order = np.argsort(rack, kind="stable")
    rack_start = np.searchsorted(rack[order], np.arange(racks))
    rack_size = np.bincount(rack, minlength=racks)
    mrack = (np.arange(n_match) % racks)
    ok = rack_size[mrack] > 0
    mrack = mrack[ok]
    mnode = order[rack_start[mrack] + (rng.integers(0, 1 << 30, size=len(mrack)) % rack_size[mrack])]
    onode = rng.integers(0, n, size=n_existing - n_match)
    pn = np.concatenate([mnode, onode])
    pcpu = rng.integers(10, 51, size=len(pn)).astype(np.int64) * 10
    pmem = rng.integers(128, 1025, size=len(pn)).astype(np.int64) * MiB
    return a_cpu, a_mem, a_pods, zone, rack, region, pn, pcpu, pmem, len(mnode)


def c4(n=100_000, seed=3, n_existing=200_000, zones=64, racks=1024, regions=8, match_frac=0.30):
    """100k nodes, 3 DoNotSchedule topology-spread constraints (zone/rack/region; maxSkew 1/2/4) + required
    anti-affinity on kubernetes.io/hostname against app=sim, 200k pre-existing pods of which 30% carry app=sim.
    The template itself carries app=sim (it spreads / repels itself).

    Topology is hierarchical (rack -> zone = rack % zones -> region = zone % regions). The app=sim pods that already
    exist are spread evenly over racks (as if they had been scheduled under the same constraints): with SURVEY.md
    §8(d)'s literal "assigned uniformly at random" the per-zone counts differ by ~30 and maxSkew=1 stops the run
    after <10 placements, which measures nothing. The other 70% of the existing pods are assigned uniformly."""
    a_cpu, a_mem, a_pods, zone, rack, region, pn, pcpu, pmem, n_m = _c4_raw(n, seed, n_existing, zones, racks, regions, match_frac)
    mnode = pn[:n_m]
    req_cpu = np.bincount(pn, weights=pcpu, minlength=n).astype(np.int64)
    req_mem = np.bincount(pn, weights=pmem, minlength=n).astype(np.int64)
    npods = np.bincount(pn, minlength=n).astype(np.int32)
    mcount = np.bincount(mnode, minlength=n).astype(np.int32)   # app=sim pods per node
    snap = abi.Snapshot(n, a_cpu, a_mem, a_pods, req_cpu=req_cpu, req_mem=req_mem, npods=npods,
                        topo=[zone, rack, region])
    counters = [
        abi.make_counter(0, np.bincount(zone, weights=mcount, minlength=zones).astype(np.int32), inc=1),
        abi.make_counter(1, np.bincount(rack, weights=mcount, minlength=racks).astype(np.int32), inc=1),
        abi.make_counter(2, np.bincount(region, weights=mcount, minlength=regions).astype(np.int32), inc=1),
        abi.make_counter(-1, mcount, inc=1),   # hostname: every node its own domain
    ]
    t = abi.default_template(150, 100 * MiB)
    t.n_pts = 3
    for c, skew in enumerate((1, 2, 4)):
        t.pts[c].counter = c
        t.pts[c].max_skew = skew
        t.pts[c].self_match = 1
        t.pts[c].min_zero = 0
    t.n_anti = 1
    t.anti_counter[0] = 3
    return snap, [t], counters


def c4_objects(n=100_000, seed=3, n_existing=200_000, zones=64, racks=1024, regions=8, match_frac=0.30):
    """The same cluster as c4() as API objects (v1.Node / v1.Pod dicts + the podspec) for the reference-facing path
    (framework.New / SyncWithClient): node i of c4() is the i-th Node. The topology labels use example.com/* keys — the well-known
    zone / region labels would make nodeTree reorder the nodes (node_tree.go:119-143) and the two paths would no longer index
    nodes alike. Requests are whole milli-cpu / MiB values, so the encoded columns equal c4()'s exactly."""
    a_cpu, a_mem, a_pods, zone, rack, region, pn, pcpu, pmem, n_m = _c4_raw(n, seed, n_existing, zones, racks, regions, match_frac)
    nodes = [{"apiVersion": "v1", "kind": "Node",
              "metadata": {"name": "node-%06d" % i, "labels": {"kubernetes.io/hostname": "node-%06d" % i, "example.com/zone": "z%d" % zone[i],
                                                              "example.com/rack": "k%d" % rack[i], "example.com/region": "g%d" % region[i]}},
              "spec": {}, "status": {"allocatable": {"cpu": "%dm" % a_cpu[i], "memory": "%dMi" % (a_mem[i] // MiB), "pods": str(int(a_pods[i]))}}}
             for i in range(n)]
    pods = [{"apiVersion": "v1", "kind": "Pod",
             "metadata": {"name": "pod-%06d" % j, "namespace": "default", "labels": {"app": "sim"} if j < n_m else {"app": "other"}},
             "spec": {"nodeName": "node-%06d" % pn[j],
                      "containers": [{"name": "c", "image": "img", "resources": {"requests": {"cpu": "%dm" % pcpu[j], "memory": "%dMi" % (pmem[j] // MiB)}}}]},
             "status": {"phase": "Running"}} for j in range(len(pn))]
    sel = {"matchLabels": {"app": "sim"}}
    template = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "sim-pod", "namespace": "default", "labels": {"app": "sim"}},
                "spec": {"containers": [{"name": "c", "image": "img", "resources": {"requests": {"cpu": "150m", "memory": "100Mi"}}}],
                         "topologySpreadConstraints": [
                             {"maxSkew": 1, "topologyKey": "example.com/zone", "whenUnsatisfiable": "DoNotSchedule", "labelSelector": sel},
                             {"maxSkew": 2, "topologyKey": "example.com/rack", "whenUnsatisfiable": "DoNotSchedule", "labelSelector": sel},
                             {"maxSkew": 4, "topologyKey": "example.com/region", "whenUnsatisfiable": "DoNotSchedule", "labelSelector": sel}],
                         "affinity": {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
                             {"labelSelector": sel, "topologyKey": "kubernetes.io/hostname"}]}}}}
    return nodes, pods, template


def c5(n=1_000_000, seed=4, n_templates=64):
    """1M nodes (C2 distribution) x 64 distinct podspecs placed round-robin (pod k uses template k % 64;
    report.go:160). NB: the reference supports ONE template (simulator.go:122); this is the roadmap's list-of-pods
    extension (README.md:305-306)."""
    rng = _rng(seed)
    a_cpu, a_mem, a_pods, r_cpu, r_mem, npods = _c2_nodes(n, rng)
    snap = abi.Snapshot(n, a_cpu, a_mem, a_pods, req_cpu=r_cpu, req_mem=r_mem, npods=npods)
    tmpl = []
    for _ in range(n_templates):
        cpu = int(rng.integers(50, 2001))
        mem = int(rng.integers(64, 4097)) * MiB
        tmpl.append(abi.default_template(cpu, mem, fit_only=True))
    return snap, tmpl, []


def closed_form_capacity(snap, t):
    """KA5: single template, node-local predicates, unlimited -> every node fills to its own
    capacity. Returns per-node capacity (numpy int64); sum = instance count."""
    inf = np.iinfo(np.int64).max
    cap = (snap.alloc_pods.astype(np.int64) - snap.npods).clip(min=0)
    if t.req_cpu > 0:
        cap = np.minimum(cap, ((snap.alloc_cpu - snap.req_cpu) // t.req_cpu).clip(min=0))
    if t.req_mem > 0:
        cap = np.minimum(cap, ((snap.alloc_mem - snap.req_mem) // t.req_mem).clip(min=0))
    if t.req_eph > 0:
        cap = np.minimum(cap, ((snap.alloc_eph - snap.req_eph) // t.req_eph).clip(min=0))
    return np.where(cap == inf, 0, cap)
