"""GPU: the reference-facing API (framework.New / SyncWithClient / Run / Report, mirroring pkg/framework) end to end —
objects -> C++ encoder -> CUDA hot path -> report — against the object-level oracle. These read like the reference's
own tests (pkg/framework/simulator_test.go, test/benchmark/pod_colocation_test.go)."""
import importlib
import json

import pytest

import helpers
from oracle import objref

fw = importlib.import_module("cluster-capacity_b200.framework")
pytestmark = pytest.mark.gpu


def analyse(nodes, pods, tmpl, max_pods=0, exclude=(), variant=None):
    cc = fw.New(None, None, tmpl, max_pods, list(exclude))
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods, variant))
    cc.Run()
    ref = objref.Simulator(tmpl, max_pods, exclude)
    helpers.objref_sync(ref, nodes, pods, variant)
    ref.run()
    return cc, ref


def check(cc, ref):
    rep = cc.Report()
    want = ref.report()
    assert cc.ScheduledPods() == ref.pods_status
    assert cc.StopReason() == ref.stop_reason
    assert rep["status"]["replicas"] == want["replicas"]
    assert rep["status"]["failReason"] == {"failType": want["failType"], "failMessage": want["failMessage"]}
    assert rep["status"]["pods"][0]["replicasOnNodes"] == want["replicasOnNodes"]
    assert rep["status"]["pods"][0]["failSummary"] is None
    return rep


@pytest.mark.parametrize("seed", [11, 12, 13])
@pytest.mark.parametrize("variant", helpers.TEMPLATE_VARIANTS)
def test_framework_matches_object_oracle(built, variant, seed):
    nodes, pods = helpers.random_cluster(seed, n_nodes=40 + 9 * (seed - 11), n_pods=70 + 25 * (seed - 11))
    cc, ref = analyse(nodes, pods, helpers.template(variant), variant=variant)
    check(cc, ref)


def _prediction_nodes():
    # pkg/framework/simulator_test.go:103-152
    return [helpers.make_node("test-node-1", cpu="300m", mem="1000000000", pods="3"),
            helpers.make_node("test-node-2", cpu="400m", mem="2000000000", pods="3"),
            helpers.make_node("test-node-3", cpu="1200m", mem="1000000000", pods="3")]


@pytest.mark.parametrize("limit,fail_type", [(6, "LimitReached"), (0, "Unschedulable")])
def test_prediction(built, limit, fail_type):
    # TestPrediction (simulator_test.go:154-259): only FailType is asserted by the reference
    pod = helpers.make_pod("simulated-pod", cpu="100m", mem="5000000")
    pod["spec"]["containers"][0]["resources"]["requests"]["nvdia.com/gpu"] = "0"
    cc, ref = analyse(_prediction_nodes(), [], pod, max_pods=limit)
    rep = check(cc, ref)
    assert rep["status"]["failReason"]["failType"] == fail_type
    assert rep["spec"]["replicas"] == limit


def test_pod_affinity_hard_constraint_single_node(built):
    # TestPodAffinityHardConstraintSingleNode (test/benchmark/pod_colocation_test.go:18-93): all pods on exactly one node
    nodes = [helpers.make_node("node-%d" % i, cpu="1000m", mem="1000", pods="30") for i in range(3)]
    pod = helpers.make_pod("p", cpu="10m", mem="10", labels={"app": "x"})
    pod["spec"]["affinity"] = {"podAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
        {"labelSelector": {"matchLabels": {"app": "x"}}, "topologyKey": "kubernetes.io/hostname"}]}}
    cc, ref = analyse(nodes, [], pod, max_pods=100)
    check(cc, ref)
    assert len(set(cc.ScheduledPods())) == 1 and len(cc.ScheduledPods()) == 30


def test_pod_affinity_hard_constraint_many_nodes(built):
    # TestPodAffinityHardConstraintManyNodes (:95-190): 9 nodes / 3 zones, all pods in exactly one zone
    nodes = [helpers.make_node("node-%d" % i, cpu="1000m", mem="1000", pods="30", labels={"topology.kubernetes.io/zone": "zone-%d" % (i // 3)})
             for i in range(9)]
    pod = helpers.make_pod("p", cpu="10m", mem="10", labels={"app": "x"})
    pod["spec"]["affinity"] = {"podAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
        {"labelSelector": {"matchLabels": {"app": "x"}}, "topologyKey": "topology.kubernetes.io/zone"}]}}
    cc, ref = analyse(nodes, [], pod, max_pods=100)
    check(cc, ref)
    zone = {n["metadata"]["name"]: n["metadata"]["labels"]["topology.kubernetes.io/zone"] for n in nodes}
    assert len({zone[n] for n in cc.ScheduledPods()}) == 1 and len(cc.ScheduledPods()) == 90


def test_readme_output_formats(built):
    nodes = [helpers.make_node("kube-node-%d" % i, cpu="2", mem="4Gi", pods="110") for i in range(1, 5)]
    pod = helpers.make_pod("small-pod", cpu="150m", mem="100Mi")
    cc, ref = analyse(nodes, [], pod)
    check(cc, ref)
    assert cc.Print(False, "") == "52\n"
    verbose = cc.Print(True, "")
    assert verbose.startswith("small-pod pod requirements:\n\t- CPU: 150m\n\t- Memory: 100Mi\n\nThe cluster can schedule 52 instance(s) of the pod small-pod.\n")
    assert "\nTermination reason: Unschedulable: 0/4 nodes are available: 4 Insufficient cpu. preemption:" in verbose
    assert verbose.endswith("Pod distribution among nodes:\nsmall-pod\n\t- kube-node-1: 13 instance(s)\n\t- kube-node-2: 13 instance(s)\n"
                            "\t- kube-node-3: 13 instance(s)\n\t- kube-node-4: 13 instance(s)\n")
    j = json.loads(cc.Print(False, "json"))
    assert j["spec"]["podRequirements"][0]["resources"]["primaryResources"] == {"cpu": "150m", "memory": "100Mi", "nvdia.com/gpu": "0"}
    y = cc.Print(False, "yaml")
    assert "replicas: 52" in y and "- nodeName: kube-node-1" in y
    with pytest.raises(fw.FrameworkError, match="not recognized"):
        cc.Print(False, "xml")


def test_empty_cluster(built):
    cc, ref = analyse([], [], helpers.make_pod("p", cpu="1"))
    assert cc.StopReason() == ref.stop_reason == "Unschedulable: no nodes available to schedule pods"
