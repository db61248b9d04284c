"""Committed fixtures (tests/golden): the reference's own asserted outcomes, and frozen object-oracle sequences.
CPU: the encoder + C oracle must reproduce them; GPU: the framework API must."""
import importlib
import json
import os

import pytest

import helpers
from oracle import binding as oracle
from oracle import objref

fw = importlib.import_module("cluster-capacity_b200.framework")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ASSERTED = json.load(open(os.path.join(GOLD, "reference_asserted.json")))["cases"]
SEQS = json.load(open(os.path.join(GOLD, "objref_sequences.json")))["cases"]


def cpu_path(nodes, pods, tmpl, max_pods, variant=None):
    cc = fw.New(None, None, tmpl, max_pods, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods, variant))
    snap, T, ctr, tdict, snames, names = helpers.from_encoded(cc.EncodedSnapshot())
    got = oracle.run(snap, T, ctr, max_pods=max_pods)
    sr = helpers.stop_reason_from_result(got, snap.n, max_pods, tdict, snames, tmpl["spec"].get("preemptionPolicy") == "Never")
    return [names[i] for i in got.pod_node.tolist()], sr


@pytest.mark.parametrize("case", ASSERTED, ids=[c["name"] for c in ASSERTED])
def test_reference_asserted_outcomes_cpu(built, case):
    seq, sr = cpu_path(case["nodes"], case["pods"], case["template"], case["max_pods"])
    ref = objref.Simulator(case["template"], case["max_pods"])
    ref.sync(case["nodes"], case["pods"])
    ref.run()
    for got_seq, got_sr in ((seq, sr), (ref.pods_status, ref.stop_reason)):
        assert len(got_seq) == case["expect"]["replicas"]
        assert got_sr.split(":")[0] == case["expect"]["failType"]
        if "per_node" in case["expect"]:
            assert {n: got_seq.count(n) for n in set(got_seq)} == case["expect"]["per_node"]


@pytest.mark.parametrize("case", SEQS, ids=[c["name"] for c in SEQS])
def test_frozen_sequences_cpu(built, case):
    nodes, pods = helpers.random_cluster(case["cluster_seed"], n_nodes=24, n_pods=40)
    seq, sr = cpu_path(nodes, pods, helpers.template(case["variant"], case["cluster_seed"]), 0, case["variant"])
    assert seq == case["scheduled"] and sr == case["stop_reason"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ASSERTED + SEQS[::4], ids=[c["name"] for c in ASSERTED + SEQS[::4]])
def test_golden_gpu(built, case):
    if "nodes" in case:
        nodes, pods, tmpl, limit = case["nodes"], case["pods"], case["template"], case["max_pods"]
    else:
        nodes, pods = helpers.random_cluster(case["cluster_seed"], n_nodes=24, n_pods=40)
        tmpl, limit = helpers.template(case["variant"], case["cluster_seed"]), 0
    cc = fw.New(None, None, tmpl, limit, [])
    cc.SyncWithClient(helpers.list_client(fw, nodes, pods, case.get("variant")))
    cc.Run()
    if "expect" in case:
        assert len(cc.ScheduledPods()) == case["expect"]["replicas"] and cc.StopReason().split(":")[0] == case["expect"]["failType"]
    else:
        assert cc.ScheduledPods() == case["scheduled"] and cc.StopReason() == case["stop_reason"]
