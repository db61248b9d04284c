"""CPU: the C++ host encoder (libcchost) + the flat C oracle against the object-level Python oracle (oracle/objref.py)
on random small clusters and every supported podspec feature. No GPU: the encoded snapshot is run by the C oracle."""
import importlib
import json

import numpy as np
import pytest

import helpers
from oracle import binding as oracle
from oracle import objref

fw = importlib.import_module("cluster-capacity_b200.framework")
abi = importlib.import_module("cluster-capacity_b200._abi")


def run_both(nodes, pods, tmpl, max_pods=0, exclude=(), variant=None):
    ref = objref.Simulator(tmpl, max_pods, exclude)
    helpers.objref_sync(ref, nodes, pods, variant)
    ref.run()
    cc = fw.New(None, None, tmpl, max_pods, list(exclude))
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods, variant))
    enc = cc.EncodedSnapshot()
    snap, T, ctr, tdict, snames, names = helpers.from_encoded(enc)
    got = oracle.run(snap, T, ctr, max_pods=max_pods)
    seq = [names[i] for i in got.pod_node.tolist()]
    sr = helpers.stop_reason_from_result(got, snap.n, max_pods, tdict, snames, preemption_never=tmpl["spec"].get("preemptionPolicy") == "Never")
    cc.Close()
    return ref, seq, sr


@pytest.mark.parametrize("variant", helpers.TEMPLATE_VARIANTS)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_encoder_matches_object_oracle(built, variant, seed):
    nodes, pods = helpers.random_cluster(seed, n_nodes=30, n_pods=50)
    tmpl = helpers.template(variant, seed)
    ref, seq, sr = run_both(nodes, pods, tmpl, variant=variant)
    assert seq == ref.pods_status
    assert sr == ref.stop_reason


def test_limit_and_exclude(built):
    nodes, pods = helpers.random_cluster(7, n_nodes=25, n_pods=30)
    tmpl = helpers.template("plain")
    ref, seq, sr = run_both(nodes, pods, tmpl, max_pods=17, exclude=("node-03", "node-11"))
    assert seq == ref.pods_status and sr == ref.stop_reason == "LimitReached: Maximum number of pods simulated: 17"
    assert "node-03" not in seq and "node-11" not in seq


def test_readme_demo_objects(built):
    # README.md:44-66 — 4 nodes x 2 CPU / 4 GiB, examples/pod.yaml: 52 instances, 13 per node
    nodes = [helpers.make_node("kube-node-%d" % i, cpu="2", mem="4Gi", pods="110") for i in range(1, 5)]
    tmpl = helpers.make_pod("small-pod", cpu="150m", mem="100Mi")
    ref, seq, sr = run_both(nodes, [], tmpl)
    assert len(seq) == 52 and seq == ref.pods_status
    assert sr == ("Unschedulable: 0/4 nodes are available: 4 Insufficient cpu. preemption: 0/4 nodes are available: "
                  "4 No preemption victims found for incoming pod.")


def test_zone_round_robin_node_order(built):
    # node_tree.go:119-143: zones in first-seen order, round-robin; equal scores -> first node in that order wins
    zs = ["a", "a", "a", "b", "b", "c"]
    nodes = [helpers.make_node("n%d" % i, labels={"topology.kubernetes.io/zone": z}) for i, z in enumerate(zs)]
    ref, seq, _ = run_both(nodes, [], helpers.make_pod("p", cpu="1", mem="1Gi"), max_pods=6)
    assert seq == ref.pods_status == ["n0", "n3", "n5", "n1", "n4", "n2"]


def test_unsupported_features_are_refused(built):
    nodes, pods = helpers.random_cluster(1, n_nodes=5, n_pods=0)
    t = helpers.make_pod("p", cpu="100m")
    t["spec"]["resourceClaims"] = [{"name": "gpu", "resourceClaimName": "claim"}]
    cc = fw.New(None, None, t, 0, [])
    cc.SyncWithClient(fw.ListClient(nodes, pods))
    with pytest.raises(fw.UnsupportedError, match="DynamicResources"):
        cc.EncodedSnapshot()
    t2 = helpers.make_pod("p", cpu="100m")
    t2["spec"]["volumes"] = [{"name": "v", "persistentVolumeClaim": {"claimName": "c"}}]
    cc2 = fw.New(None, None, t2, 0, [])
    cc2.SyncWithClient(fw.ListClient(nodes, pods))
    with pytest.raises(fw.UnsupportedError, match="VolumeBinding"):
        cc2.EncodedSnapshot()


def test_quantities_round_up(built):
    # Quantity.MilliValue()/Value() are ceilings (quantity.go:812-834): 0.1m cpu -> 1 milli, 1.5 bytes -> 2
    nodes = [helpers.make_node("n0", cpu="10m", mem="10", pods="100")]
    t = helpers.make_pod("p", cpu="0.1m", mem="1.5")
    ref, seq, sr = run_both(nodes, [], t)
    assert len(seq) == 5 and seq == ref.pods_status   # memory: floor(10 / 2)


def test_system_default_spreading_needs_a_selecting_service(built):
    """plugin.go:48-59 + helper/spread.go:40-113: a pod without topologySpreadConstraints gets the two system-default soft
    constraints only when a Service of its namespace (or its owning controller) selects it; a Service in another namespace,
    one selecting other labels and one with a nil selector change nothing."""
    nodes, pods = helpers.random_cluster(9, n_nodes=30, n_pods=60)
    tmpl = helpers.template("plain")
    plain, seq_plain, _ = run_both(nodes, pods, tmpl)
    ref, seq, sr = run_both(nodes, pods, tmpl, variant="svc_default_spread")
    assert seq == ref.pods_status and sr == ref.stop_reason
    assert seq != seq_plain                       # spreading changed the placement order
    assert sorted(seq) == sorted(seq_plain)       # ... but not the capacity (scores never change feasibility)
    irrelevant = {"services": helpers.workloads_for("svc_default_spread")["services"][1:]}
    cc = fw.New(None, None, tmpl, 0, [])
    cc.SyncWithClient(fw.ListClient(nodes, pods, **irrelevant))
    enc = cc.EncodedSnapshot()
    assert helpers.from_encoded(enc)[1][0].n_spts == 0
    cc.Close()


def test_go_log_matches_libm_to_an_ulp():
    """The restated math.Log (FreeBSD e_log.c port) agrees with libm within 1 ulp on the sizes the scorer feeds it."""
    import math
    for n in list(range(2, 2000)) + [5000, 100002, 1000002]:
        a, b = objref.go_log(float(n)), math.log(float(n))
        assert abs(a - b) <= math.ulp(b), n


def test_ingest_shapes_and_parallel_item_parsing(built):
    """cc_sync_with_objects takes bare arrays, List objects (whatever the member order) and single objects; big lists are split at
    item boundaries by a string-aware scan and parsed on several threads — same result, same node order."""
    import json
    nodes = [helpers.make_node("n%04d" % i, cpu="2", mem="4Gi", pods="10",
                               labels={"weird": "a]b}\\\"c[{", "topology.kubernetes.io/zone": "z%d" % (i % 3)}) for i in range(2500)]
    pods = [helpers.make_pod("p%04d" % j, cpu="500m", mem="1Gi", node="n%04d" % (j % 2500), labels={"items": "[not a list]"}) for j in range(3000)]
    tmpl = helpers.make_pod("t", cpu="1", mem="1Gi")
    L = fw.lib()

    def encoded(nodes_doc, pods_doc):
        cc = fw.New(None, None, tmpl, 0, [])
        rc = L.cc_sync_with_objects(cc._h, json.dumps(nodes_doc).encode(), json.dumps(pods_doc).encode(), None)
        assert rc == 0, L.cc_last_error(cc._h)
        enc = cc.EncodedSnapshot()
        cc.Close()
        return enc["names"], enc["nodes"]["req_cpu"], enc["nodes"]["npods"]

    want = encoded(nodes, pods)
    assert want[0][:4] == ["n0000", "n0001", "n0002", "n0003"] and sum(want[2]) == 3000
    wrapped_nodes = {"kind": "NodeList", "apiVersion": "v1", "metadata": {"items": ["decoy"], "resourceVersion": "7"}, "items": nodes}
    wrapped_pods = {"items": pods, "kind": "PodList", "metadata": {}}
    assert encoded(wrapped_nodes, wrapped_pods) == want
    one = encoded(nodes[0], [])
    assert one[0] == ["n0000"]
    assert encoded([], [])[0] == []
    cc = fw.New(None, None, tmpl, 0, [])
    assert L.cc_sync_with_objects(cc._h, b'[{"metadata": {"name": "x"}', b"[]", None) != 0      # truncated document: an error, not a crash
    assert b"json" in L.cc_last_error(cc._h)
    cc.Close()


def test_c4_objects_through_the_encoder_equal_the_flat_generator(built):
    """synth.c4_objects (v1.Node / v1.Pod dicts: what bench.py's e2e_objects leg feeds the plugin call) encodes to exactly the
    columns synth.c4 builds directly, and both give the same placement sequence. 6000 nodes: the encoder's per-node loops take
    their multi-threaded path (>= 4096 nodes)."""
    synth = importlib.import_module("cluster-capacity_b200.synth")
    kw = dict(n=6000, n_existing=12000, zones=8, racks=64, regions=4)
    snap, tmpl, ctr = synth.c4(**kw)
    nodes, pods, t = synth.c4_objects(**kw)
    cc = fw.New(None, None, t, 0, [])
    cc.SyncWithClient(fw.ListClient(nodes, pods, ()))
    s2, T2, c2, _, _, names = helpers.from_encoded(cc.EncodedSnapshot())
    for f in ("alloc_cpu", "alloc_mem", "alloc_pods", "req_cpu", "req_mem", "npods", "nz_cpu", "nz_mem"):
        assert np.array_equal(getattr(snap, f), getattr(s2, f)), f
    assert names[:2] == ["node-000000", "node-000001"] and len(c2) == len(ctr)
    a, b = oracle.run(snap, tmpl, ctr, threads=4, memo=True), oracle.run(s2, T2, c2, threads=4, memo=True)
    assert a.placed == b.placed > 100 and np.array_equal(a.pod_node, b.pod_node) and np.array_equal(a.reason_hist, b.reason_hist)


def test_large_topology_dictionaries(built):
    """More distinct topology values than the encoder's dictionaries are first sized for (the flat index grows while ids are being
    handed out in first-seen order): 2500 racks over 40000 nodes; same columns, counters and placement sequence as the flat generator."""
    synth = importlib.import_module("cluster-capacity_b200.synth")
    kw = dict(n=40000, n_existing=25000, zones=50, racks=2500, regions=5, match_frac=0.5)
    snap, tmpl, ctr = synth.c4(**kw)
    nodes, pods, t = synth.c4_objects(**kw)
    cc = fw.New(None, None, t, 0, [])
    cc.SyncWithClient(fw.ListClient(nodes, pods, ()))
    s2, T2, c2, _, _, names = helpers.from_encoded(cc.EncodedSnapshot())
    assert len(c2) == len(ctr) and sorted(c.n_domains for c in c2)[-2] == 2500
    a, b = oracle.run(snap, tmpl, ctr, threads=4, memo=True), oracle.run(s2, T2, c2, threads=4, memo=True)
    assert a.placed == b.placed > 100 and np.array_equal(a.pod_node, b.pod_node) and np.array_equal(a.reason_hist, b.reason_hist)


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_fast_ingest_equals_the_general_parser(built, seed, monkeypatch):
    """fastparse.hpp (no DOM; gives up on pods with init containers / overhead / pod affinity / escapes) against the general parser
    (CCHOST_DOM_ONLY=1) on random clusters whose pods carry all of that: identical encoded snapshots, byte for byte."""
    nodes, pods = helpers.random_cluster(seed, n_nodes=60, n_pods=150)
    pods[3]["metadata"]["labels"]["quote"] = 'a"b' + chr(92) + 'c'                      # escapes: the fast path must hand the item over
    pods[4]["metadata"]["deletionTimestamp"] = "2026-01-01T00:00:00Z"
    pods[5]["status"]["containerStatuses"] = [{"name": "c", "resources": {"requests": {"cpu": "300m"}}, "allocatedResources": {"cpu": "300m", "memory": "64Mi"}}]
    pods[6]["spec"]["containers"][0]["resources"]["requests"] = {"cpu": 2, "memory": "1Gi"}       # a bare JSON number as a quantity
    nodes[2]["status"]["conditions"] = [{"type": "Ready", "status": "True", "message": 'kubelet is posting "ready" status'}]
    nodes[3]["status"]["nodeInfo"] = {"kubeletVersion": "v1.34.1", "architecture": "amd64"}
    out = []
    for dom_only in (False, True):
        if dom_only:
            monkeypatch.setenv("CCHOST_DOM_ONLY", "1")
        for variant in ("plain", "anti_zone", "spread_two", "pref_pod_affinity"):
            cc = fw.New(None, None, helpers.template(variant), 0, [])
            cc.SyncWithClient(helpers.list_client(fw, nodes, pods, variant))
            out.append(cc.EncodedSnapshot())
            cc.Close()
    half = len(out) // 2
    for a, b in zip(out[:half], out[half:]):
        for e in (a, b):       # the template bytes end with a process-local pointer (image_score): compare everything before it
            e["template_hex"] = e["template_hex"][:-16]
            e["templates_hex"] = [x[:-16] for x in e["templates_hex"]]
        assert a == b


def test_parallel_item_location_equals_the_serial_scan(built, monkeypatch):
    """Documents of more than 4 MB have their item spans located on all host cores (cchost.cpp item_spans_parallel: quote parity and
    bracket depth per chunk, then a prefix): same encoded snapshot as the serial string-aware scan, with strings full of brackets,
    quotes and backslashes lying wherever the chunk cuts fall."""
    nodes, pods = helpers.random_cluster(21, n_nodes=300, n_pods=9000)
    nasty = ['}]{[', 'a"b', chr(92), chr(92) * 2 + '"', '"]},{"', chr(92) * 3, '[[[[', '\n\t"', 'x' * 7 + chr(92)]
    for j, p in enumerate(pods):      # annotations are skipped by the ingest: only the scanner sees them
        p["metadata"]["annotations"] = {"note-%d" % q: nasty[(j + q) % len(nasty)] * (1 + (j + 3 * q) % 5) + "#" * ((j * 7 + q) % 90) for q in range(6)}
    for i, n in enumerate(nodes):
        n["metadata"]["annotations"] = {"n": nasty[i % len(nasty)] * 3}
    assert len(json.dumps(pods)) > (4 << 20)
    out = []
    for serial in (False, True):
        if serial:
            monkeypatch.setenv("CCHOST_SERIAL_SPANS", "1")
        cc = fw.New(None, None, helpers.template("spread_two"), 0, [])
        cc.SyncWithClient(helpers.list_client(fw, nodes, pods, "spread_two"))
        e = cc.EncodedSnapshot()
        e["template_hex"] = e["template_hex"][:-16]
        e["templates_hex"] = [x[:-16] for x in e["templates_hex"]]
        out.append(e)
        cc.Close()
    assert out[0] == out[1]
    assert sum(out[0]["nodes"]["npods"]) > 1000


def test_calculate_resource_fast_path_equals_the_general_path(built):
    """CCHOST_CHECK_FAST=1 makes the encoder compute every existing pod's resources twice — the exact-sum fast path for plain pods and
    resourcehelper.PodRequests' list arithmetic (types.go:700-734) — and fail on any difference. The switch is read once per process:
    the object-level cases (init containers, sidecars, overhead, pod-level resources, sub-milli quantities, extended resources) run in
    a child interpreter with it set."""
    import os, subprocess, sys
    env = dict(os.environ, CCHOST_CHECK_FAST="1", CCSIM_NO_REBUILD="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_host_objects.py"), "-q", "-x", "-m", "not gpu",
                        "-k", "encoder_matches_object_oracle or quantities_round_up or c4_objects or readme_demo"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(here), timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_calculate_resource_fast_path_random_request_shapes(built):
    """... and random container shapes aimed at the rounding rules: 0-3 containers, requests missing per resource, sub-milli cpu
    ("0.5m", "1500u", "3n"), fractional memory ("1.5", "129Mi", "1e3"), extended resources, ephemeral storage."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import importlib, json, random, sys
        sys.path.insert(0, %r)
        fw = importlib.import_module("cluster-capacity_b200.framework")
        rnd = random.Random(7)
        cpus = ["0.5m", "1500u", "3n", "100m", "1", "0.25", "2500m", "0"]
        mems = ["1.5", "129Mi", "1e3", "1Gi", "200M", "0", "123456789"]
        def container(k):
            req = {}
            if rnd.random() < 0.7: req["cpu"] = rnd.choice(cpus)
            if rnd.random() < 0.7: req["memory"] = rnd.choice(mems)
            if rnd.random() < 0.2: req["ephemeral-storage"] = rnd.choice(["1Gi", "500M", "1.5"])
            if rnd.random() < 0.2: req["example.com/gpu"] = str(rnd.randint(1, 3))
            c = {"name": "c%%d" %% k, "image": "img"}
            if req or rnd.random() < 0.5: c["resources"] = {"requests": req}
            return c
        nodes = [{"metadata": {"name": "n%%d" %% i, "labels": {"kubernetes.io/hostname": "n%%d" %% i}}, "spec": {},
                  "status": {"allocatable": {"cpu": "64", "memory": "256Gi", "pods": "110", "example.com/gpu": "8", "ephemeral-storage": "1Ti"}}} for i in range(20)]
        pods = [{"metadata": {"name": "p%%d" %% j, "namespace": "default"}, "spec": {"nodeName": "n%%d" %% rnd.randrange(20),
                 "containers": [container(k) for k in range(rnd.randint(0, 3))]}, "status": {"phase": "Running"}} for j in range(400)]
        tmpl = {"metadata": {"name": "t", "namespace": "default"}, "spec": {"containers": [{"name": "c", "image": "img",
                "resources": {"requests": {"cpu": "100m", "memory": "64Mi", "example.com/gpu": "1"}}}]}}
        cc = fw.New(None, None, tmpl, 0, [])
        cc.SyncWithClient(fw.ListClient(nodes=nodes, pods=pods))
        enc = cc.EncodedSnapshot()
        assert sum(enc["nodes"]["npods"]) == 400
        print("ok")
    ''') % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CCHOST_CHECK_FAST="1", CCSIM_NO_REBUILD="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_pending_pods_are_left_out_and_reported(built):
    """Pending pods of the source cluster (no spec.nodeName) are not replayed — the reference would bind them through its binder
    and count them as simulated instances in no defined order (simulator.go:193-200,297-312) — and cc_warnings says how many."""
    nodes, pods = helpers.random_cluster(5, n_nodes=10, n_pods=12)
    tmpl = helpers.template("plain", 5)
    base, seq0, sr0 = run_both(nodes, pods, tmpl)
    pending = [{"metadata": {"name": "pend-%d" % i, "namespace": "default"}, "spec": {"containers": [{"name": "c", "image": "i",
                "resources": {"requests": {"cpu": "4", "memory": "1Gi"}}}]}, "status": {"phase": "Pending"}} for i in range(3)]
    done = [{"metadata": {"name": "done", "namespace": "default"}, "spec": {"containers": []}, "status": {"phase": "Succeeded"}}]
    cc = fw.New(None, None, tmpl, 0, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods + pending + done, None))
    w = cc.Warnings()
    n_pending = sum(1 for p in pods + pending + done if not p["spec"].get("nodeName") and p.get("status", {}).get("phase") not in ("Succeeded", "Failed"))
    assert n_pending >= 3 and len(w) == 1 and w[0].startswith("%d pending pod(s)" % n_pending)
    enc = cc.EncodedSnapshot()
    snap, T, ctr, tdict, snames, names = helpers.from_encoded(enc)
    got = oracle.run(snap, T, ctr, max_pods=0)
    assert [names[i] for i in got.pod_node.tolist()] == seq0      # same result as without them
    cc.Close()
    cc = fw.New(None, None, tmpl, 0, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, [p for p in pods if p["spec"].get("nodeName")], None))
    assert cc.Warnings() == []
    cc.Close()


def test_simd_item_location_equals_scalar_and_serial(built, tmp_path):
    """The AVX2 block-mask passes of item_spans_parallel against the scalar passes and the serial scan, on documents whose strings are
    full of brackets, quotes and backslash runs, for several chunkings (the thread count moves the chunk cuts): one digest for all."""
    import hashlib, os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import hashlib, importlib, json, sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import helpers
        fw = importlib.import_module("cluster-capacity_b200.framework")
        nodes, pods = helpers.random_cluster(33, n_nodes=200, n_pods=7000)
        bs = chr(92)
        nasty = ['}]{[', 'a"b', bs, bs * 2 + '"', '"]},{"', bs * 3, '[[[[', '\\n\\t"', 'x' * 31 + bs, 'y' * 30 + bs * 2, '{' * 33, '"' * 5, bs + '"' + bs, 'z' * 64 + '}' ]
        for j, p in enumerate(pods):
            p["metadata"]["annotations"] = {"k%%d" %% q: nasty[(j + q) %% len(nasty)] * (1 + (j + 3 * q) %% 4) + "#" * ((j * 13 + q * 5) %% 97) for q in range(7)}
        for i, n in enumerate(nodes):
            n["metadata"]["annotations"] = {"n": nasty[i %% len(nasty)] * 2}
        assert len(json.dumps(pods)) > (4 << 20)
        cc = fw.New(None, None, helpers.template("spread_two"), 0, [])
        cc.SyncWithClient(helpers.list_client(fw, nodes, pods, "spread_two"))
        e = cc.EncodedSnapshot()
        e["template_hex"] = e["template_hex"][:-16]
        e["templates_hex"] = [x[:-16] for x in e["templates_hex"]]
        assert sum(e["nodes"]["npods"]) > 1000
        print("DIGEST", hashlib.sha256(json.dumps(e, sort_keys=True).encode()).hexdigest())
    ''') % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for name, env in (("serial", {"CCHOST_SERIAL_SPANS": "1"}), ("scalar-8", {"CCHOST_NO_SIMD": "1", "CCHOST_THREADS": "8"}), ("simd-2", {"CCHOST_THREADS": "2"}),
                      ("simd-3", {"CCHOST_THREADS": "3"}), ("simd-8", {"CCHOST_THREADS": "8"}), ("simd-13", {"CCHOST_THREADS": "13"}), ("simd-64", {"CCHOST_THREADS": "64"})):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CCSIM_NO_REBUILD="1", **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "DIGEST" in r.stdout, name + ": " + r.stdout[-1500:] + r.stderr[-1500:]
        digests[name] = r.stdout.split("DIGEST")[1].split()[0]
    assert len(set(digests.values())) == 1, digests


def test_list_items_that_are_not_objects_are_rejected(built, monkeypatch):
    """A NodeList / PodList whose items are not objects fails like the reference's decoder would (small documents, large ones through
    the parallel item location, and the serial scan)."""
    import ctypes as C
    L = fw.lib()
    tmpl = json.dumps(helpers.template("plain", 1)).encode()
    for doc in (json.dumps({"items": ["x" * 50] * 120000}), json.dumps(["s", 3]), json.dumps({"items": [{"metadata": {"name": "n"}}, 7]})):
        for serial in (False, True):
            if serial:
                monkeypatch.setenv("CCHOST_SERIAL_SPANS", "1")
            else:
                monkeypatch.delenv("CCHOST_SERIAL_SPANS", raising=False)
            h = C.c_void_p()
            assert L.cc_new(None, tmpl, 0, None, 0, C.byref(h)) == 0
            assert L.cc_sync_with_objects(h, doc.encode(), b"[]", None) != 0
            assert b"not an object" in L.cc_last_error(h)
            L.cc_close(h)
