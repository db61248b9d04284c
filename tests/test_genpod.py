"""CPU: genpod (pkg/client/nspod.go:36-131) — LimitRange minimum of Pod-type max + namespace node-selector annotation."""
import importlib

import pytest

genpod = importlib.import_module("cluster-capacity_b200.genpod")


def lr(ns, *items):
    return {"metadata": {"name": "lr", "namespace": ns}, "spec": {"limits": list(items)}}


def test_minimum_over_pod_limits_and_node_selector():
    nss = [{"metadata": {"name": "team-a", "annotations": {"openshift.io/node-selector": "region=east,disk=ssd"}}}, {"metadata": {"name": "other"}}]
    lrs = [lr("team-a", {"type": "Pod", "max": {"cpu": "2", "memory": "1Gi"}}, {"type": "Container", "max": {"cpu": "100m"}}),
           lr("team-a", {"type": "Pod", "max": {"cpu": "1500m", "memory": "2Gi", "nvdia.com/gpu": "1"}}),
           lr("other", {"type": "Pod", "max": {"cpu": "1m"}})]
    pod = genpod.retrieve_namespace_pod(nss, lrs, "team-a")
    res = pod["spec"]["containers"][0]["resources"]
    assert res["requests"] == res["limits"] == {"cpu": "1500m", "memory": "1Gi", "nvdia.com/gpu": "1"}
    assert pod["spec"]["nodeSelector"] == {"region": "east", "disk": "ssd"}
    assert pod["metadata"] == {"name": "cluster-capacity-stub-container", "namespace": "team-a"}
    assert pod["spec"]["containers"][0]["image"] == "gcr.io/google_containers/pause:2.0"


def test_no_limits_no_resources_and_errors():
    nss = [{"metadata": {"name": "ns1"}}]
    pod = genpod.retrieve_namespace_pod(nss, [lr("ns1", {"type": "Pod", "max": {"cpu": "0"}})], "ns1")
    assert "resources" not in pod["spec"]["containers"][0] and "nodeSelector" not in pod["spec"]
    with pytest.raises(LookupError):
        genpod.retrieve_namespace_pod(nss, [], "missing")
    bad = [{"metadata": {"name": "ns2", "annotations": {"openshift.io/node-selector": "a in (b)"}}}]
    with pytest.raises(ValueError, match="Unable to parse"):
        genpod.retrieve_namespace_pod(bad, [], "ns2")


def test_generated_pod_feeds_the_encoder(built):
    fw = importlib.import_module("cluster-capacity_b200.framework")
    import helpers
    from oracle import binding as oracle
    nss = [{"metadata": {"name": "team-a", "annotations": {"openshift.io/node-selector": "disk=ssd"}}}]
    pod = genpod.retrieve_namespace_pod(nss, [lr("team-a", {"type": "Pod", "max": {"cpu": "500m", "memory": "256Mi"}})], "team-a")
    nodes, pods = helpers.random_cluster(5, n_nodes=20, n_pods=10)
    cc = fw.New(None, None, pod, 0, [])
    cc.SyncWithClient(fw.ListClient(nodes, pods, nss))
    snap, T, ctr, *_ = helpers.from_encoded(cc.EncodedSnapshot())
    assert T[0].req_cpu == 500 and T[0].req_mem == 256 << 20
    assert oracle.run(snap, T, ctr).placed > 0
