"""The roadmap's "accept a list of pods" (README.md:305-306) through the reference-facing boundary: framework.New with a list,
`cluster-capacity --podspec` repeated / a directory, genpod over several namespaces. Pod k of the run is a clone of podspec
k % T (report.go:160). CPU: C++ encoder (merged snapshot) + C oracle against the object oracle; GPU: the same through CUDA."""
import importlib
import io
import json
import os
from contextlib import redirect_stdout

import pytest

import helpers
from oracle import binding as oracle
from oracle import objref

fw = importlib.import_module("cluster-capacity_b200.framework")
cli = importlib.import_module("cluster-capacity_b200.cli")
genpod = importlib.import_module("cluster-capacity_b200.genpod")

LISTS = {
    "plain3": ["plain", "best_effort", "init_overhead"],
    "selectors": ["selector", "tolerations", "affinity_terms", "gt_lt"],           # static bits of four templates side by side
    "extended": ["extended", "plain", "name_in", "tolerations"],                   # union of extended resources, PreFilter node names
    "never": ["never_preempt", "plain"],
}


def cluster(seed, n_nodes, n_pods):
    """random cluster whose existing pods carry no pod-(anti-)affinity terms: those turn into per-domain score counters for every
    incoming pod they match, and counter-coupled runs are single-podspec (refused by name, tested below)"""
    nodes, pods = helpers.random_cluster(seed, n_nodes=n_nodes, n_pods=n_pods)
    for p in pods:
        p["spec"].pop("affinity", None)
    return nodes, pods


def templates(key):
    out = []
    for i, v in enumerate(LISTS[key]):
        p = helpers.template(v)
        p["metadata"]["name"] = "%s-%d" % (v.replace("_", "-"), i)
        if i == 1 and key == "plain3":
            p["spec"]["containers"][0]["resources"] = {"requests": {"cpu": "900m", "memory": "700Mi"}}
        out.append(p)
    return out


def per_template(seq, names, T):
    rows = []
    for t in range(T):
        order, counts = [], {}
        for n in seq[t::T]:
            if n not in counts:
                order.append(n)
                counts[n] = 0
            counts[n] += 1
        rows.append([{"nodeName": n, "replicas": counts[n]} for n in order])
    return rows


@pytest.mark.parametrize("key", sorted(LISTS))
@pytest.mark.parametrize("limit", [0, 23])
def test_pod_list_encoder_matches_object_oracle(built, key, limit):
    nodes, pods = cluster(31, 30, 50)
    tm = templates(key)
    ref = objref.Simulator(tm, limit)
    ref.sync(nodes, pods)
    ref.run()
    cc = fw.New(None, None, tm, limit, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods))
    snap, T, ctr, tdict, snames, names = helpers.from_encoded(cc.EncodedSnapshot())
    assert len(T) == len(tm) and not ctr
    got = oracle.run(snap, T, ctr, max_pods=limit)
    seq = [names[i] for i in got.pod_node.tolist()]
    assert seq == ref.pods_status
    failed = tm[got.placed % len(tm)]
    sr = helpers.stop_reason_from_result(got, snap.n, limit, tdict, snames, preemption_never=failed["spec"].get("preemptionPolicy") == "Never")
    assert sr == ref.stop_reason
    cc.Close()


def test_pod_list_with_spread_terms_is_refused(built):
    nodes, pods = helpers.random_cluster(3, n_nodes=10, n_pods=10)
    cc = fw.New(None, None, [helpers.template("plain"), helpers.template("spread_zone")], 0, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods))
    with pytest.raises(fw.UnsupportedError, match="several podspecs"):
        cc.EncodedSnapshot()


@pytest.mark.gpu
def test_pod_list_with_normalised_soft_scorer_is_refused_on_gpu(built):
    """preferred nodeAffinity needs multi-pass waves (feasible-set normalisation): single podspec only, refused by name"""
    nodes, pods = cluster(5, 20, 20)
    cc = fw.New(None, None, [helpers.template("plain"), helpers.template("pref_affinity")], 0, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods))
    with pytest.raises(fw.UnsupportedError, match="single template"):
        cc.Run()


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(LISTS))
def test_pod_list_gpu(built, key):
    nodes, pods = cluster(32, 40, 60)
    tm = templates(key)
    ref = objref.Simulator(tm, 0)
    ref.sync(nodes, pods)
    ref.run()
    cc = fw.New(None, None, tm, 0, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods))
    cc.Run()
    assert cc.ScheduledPods() == ref.pods_status and cc.StopReason() == ref.stop_reason
    rep = cc.Report()
    assert [t["metadata"]["name"] for t in rep["spec"]["templates"]] == [t["metadata"]["name"] for t in tm]
    assert [r["podName"] for r in rep["status"]["pods"]] == [t["metadata"]["name"] for t in tm]
    assert [r["replicasOnNodes"] for r in rep["status"]["pods"]] == per_template(ref.pods_status, None, len(tm))
    assert rep["status"]["replicas"] == len(ref.pods_status)
    out = cc.Print(True, "")
    for t in tm:
        assert "The cluster can schedule %d instance(s) of the pod %s." % (len(ref.pods_status[tm.index(t)::len(tm)]), t["metadata"]["name"]) in out


@pytest.mark.gpu
def test_genpod_cli_gpu_64_namespaces(built, tmp_path):
    """BASELINE config C5's front end: genpod over 64 namespaces -> 64 podspec files -> cluster-capacity --podspec DIR on the
    GPU, per-template counts against the object oracle."""
    import yaml
    nodes = [helpers.make_node("n%03d" % i, cpu=str(4 + 4 * (i % 5)), mem="%dGi" % (8 + 8 * (i % 3)), pods="110",
                               labels={"pool": "a" if i % 3 else "b"}) for i in range(120)]
    nss, lrs = [], []
    for k in range(64):
        ann = {"openshift.io/node-selector": "pool=a"} if k % 7 == 0 else {}
        nss.append({"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "team%02d" % k, "annotations": ann}})
        lrs.append({"apiVersion": "v1", "kind": "LimitRange", "metadata": {"name": "lr", "namespace": "team%02d" % k},
                    "spec": {"limits": [{"type": "Pod", "max": {"cpu": "%dm" % (100 + 37 * k), "memory": "%dMi" % (64 + 29 * k)}},
                                        {"type": "Pod", "max": {"cpu": "4", "memory": "8Gi"}}]}})
    snap = tmp_path / "cluster.json"
    snap.write_text(json.dumps({"nodes": nodes, "pods": [], "namespaces": nss, "limitranges": lrs}))
    specs = tmp_path / "specs"
    assert genpod.main(["--namespace", ",".join(n["metadata"]["name"] for n in nss), "--snapshot", str(snap), "--output-dir", str(specs)]) == 0
    files = sorted(os.listdir(specs))
    assert len(files) == 64
    tm = [cli.parse_api_spec(str(specs / f)) for f in files]
    ref = objref.Simulator(tm, 6400)
    ref.sync(nodes, [], nss)
    ref.run()
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert cli.main(["--podspec", str(specs), "--snapshot", str(snap), "--max-limit", "6400", "-o", "json"]) == 0
    rep = json.loads(buf.getvalue().split("\n", 1)[1])
    assert rep["status"]["replicas"] == len(ref.pods_status)
    assert rep["status"]["failReason"]["failType"] == ref.stop_reason.split(":")[0]
    assert [r["replicasOnNodes"] for r in rep["status"]["pods"]] == per_template(ref.pods_status, None, 64)
    assert yaml.safe_load((specs / files[0]).read_text())["spec"]["nodeSelector"] == {"pool": "a"}
