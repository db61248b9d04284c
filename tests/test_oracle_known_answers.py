"""CPU: pins the oracle (oracle/ccsim_oracle.c) against the reference's own asserted outcomes and the known-answer
vectors KA1-KA5 of SURVEY.md §8(c). No GPU, no /root/reference access."""
import importlib

import numpy as np
import pytest

abi = importlib.import_module("cluster-capacity_b200._abi")
synth = importlib.import_module("cluster-capacity_b200.synth")
from oracle import binding as oracle  # noqa: E402

GiB, MiB = 1 << 30, 1 << 20


def message(res, n):
    hist = {i: int(c) for i, c in enumerate(res.reason_hist) if c}
    return abi.fit_error_message(n, hist, res.preempt_no_victims, res.preempt_not_helpful, lambda r: abi.REASON_TEXT[r])


def test_ka1_readme_52(built):
    # README.md:44-66: 4 nodes x 2 CPU / 4 GB, small-pod 150m / 100Mi -> 52 instances, 13 per node
    snap, tmpl, ctr = synth.c1()
    r = oracle.run(snap, tmpl, ctr)
    assert r.placed == 52 and r.stop_code == abi.STOP_UNSCHEDULABLE
    assert np.bincount(r.pod_node, minlength=4).tolist() == [13, 13, 13, 13]
    assert message(r, 4) == ("0/4 nodes are available: 4 Insufficient cpu. preemption: 0/4 nodes are available: "
                             "4 No preemption victims found for incoming pod.")


def test_ka1_score_trajectory(built):
    # LeastAllocated / BalancedAllocation / total for k clones already on a node
    snap, tmpl, _ = synth.c1()
    least = [94, 90, 84, 80, 74, 70, 64, 60, 55, 50, 45, 40, 35]
    bal = [97, 94, 92, 89, 87, 84, 82, 79, 77, 74, 72, 69, 67]
    total = [491, 484, 476, 469, 461, 454, 446, 439, 432, 424, 417, 409, 402]
    for k in range(13):
        tot, l, b = oracle.node_score(snap, tmpl[0], 0, k)
        assert (l, b, tot) == (least[k], bal[k], total[k])


def _test_prediction_nodes():
    # pkg/framework/simulator_test.go:103-152: alloc cpu 300m/400m/1200m, mem 1e9/2e9/1e9, pods 3/3/3
    snap = abi.Snapshot(3, np.array([300, 400, 1200]), np.array([10**9, 2 * 10**9, 10**9]), np.array([3, 3, 3]))
    t = abi.default_template(100, 5 * 10**6)   # :190-198 pod 100m / 5e6 B
    return snap, [t]


def test_ka2_testprediction_unlimited(built):
    # simulator_test.go:162-173,250-252 case B: limit 0 -> FailType "Unschedulable"
    snap, tmpl = _test_prediction_nodes()
    r = oracle.run(snap, tmpl)
    assert r.stop_code == abi.STOP_UNSCHEDULABLE and r.placed == 9
    assert np.bincount(r.pod_node, minlength=3).tolist() == [3, 3, 3]
    # node 0 (300m) is also out of cpu after 3 x 100m, and NodeResourcesFit keeps ALL failing reasons
    # (fit.go:519-533), so the histogram has 4 entries for 3 nodes (SURVEY.md §9 gotcha 9; its KA2 text missed this).
    assert message(r, 3) == ("0/3 nodes are available: 1 Insufficient cpu, 3 Too many pods. preemption: 0/3 nodes are "
                             "available: 3 No preemption victims found for incoming pod.")


def test_ka2_testprediction_limit6(built):
    # case A: limit 6 -> FailType "LimitReached", exactly 6 recorded (simulator.go:298-305)
    snap, tmpl = _test_prediction_nodes()
    r = oracle.run(snap, tmpl, max_pods=6)
    assert r.stop_code == abi.STOP_LIMIT_REACHED and r.placed == 6 and r.reason_hist.sum() == 0


def test_e2e_limit5(built):
    # test/e2e/e2e_test.go:36-39,171-173: limit 5 -> LimitReached
    snap, tmpl, ctr = synth.c1()
    r = oracle.run(snap, tmpl, ctr, max_pods=5)
    assert r.stop_code == abi.STOP_LIMIT_REACHED and r.placed == 5


def _colocation(n_nodes, zone_of):
    # test/benchmark/pod_colocation_test.go: nodes cpu=1000m mem=1000 pods=30; pod 10m / 10 B; required pod affinity to itself
    snap = abi.Snapshot(n_nodes, np.full(n_nodes, 1000), np.full(n_nodes, 1000), np.full(n_nodes, 30),
                        topo=[np.asarray(zone_of, np.int32)])
    t = abi.default_template(10, 10)
    t.flags |= abi.TF_AFF_SELF_MATCH_ALL
    t.n_aff = 1
    t.aff_counter[0] = 0
    return snap, [t]


def test_ka3_colocation_single_node(built):
    # pod_colocation_test.go:18-93: hostname affinity to self, 3 nodes, limit 100 -> all pods on exactly 1 node (30)
    snap, tmpl = _colocation(3, [0, 1, 2])
    ctr = [abi.make_counter(0, np.zeros(3, np.int32), inc=1)]
    r = oracle.run(snap, tmpl, ctr, max_pods=100)
    assert r.placed == 30 and r.stop_code == abi.STOP_UNSCHEDULABLE
    assert len(set(r.pod_node.tolist())) == 1
    assert r.reason_hist[abi.R_IPA_AFFINITY] == 2 and r.reason_hist[abi.R_TOO_MANY_PODS] == 1


def test_ka4_colocation_single_zone(built):
    # pod_colocation_test.go:95-190: zone affinity to self, 9 nodes / 3 zones -> all pods in exactly 1 zone (90)
    zone = [0, 0, 0, 1, 1, 1, 2, 2, 2]
    snap, tmpl = _colocation(9, zone)
    ctr = [abi.make_counter(0, np.zeros(3, np.int32), inc=1)]
    r = oracle.run(snap, tmpl, ctr, max_pods=100)
    assert r.placed == 90
    assert len({zone[i] for i in r.pod_node.tolist()}) == 1


@pytest.mark.parametrize("gen,kw", [("c2", dict(n=1500)), ("c3", dict(n=3000)), ("c3", dict(n=3000, prefer_taints=True))])
def test_ka5_closed_form(built, gen, kw):
    # single template + node-local predicates + no limit: count = sum of per-node capacities (order independent)
    snap, tmpl, ctr = getattr(synth, gen)(**kw)
    r = oracle.run(snap, tmpl, ctr)
    cap = synth.closed_form_capacity(snap, tmpl[0])
    if gen == "c3":
        t = tmpl[0]
        tm = snap.taint_mask[0]
        ok = ((snap.static_mask[0] & np.uint64(3)) == np.uint64(3))
        ok &= ((tm >> np.uint64(63)) & np.uint64(1)) == 0
        ok &= (tm & np.uint64(snap.taint_nosched[0]) & ~np.uint64(t.tol_nosched[0])) == 0
        cap = np.where(ok, cap, 0)
    assert r.placed == int(cap.sum())
    assert np.array_equal(np.bincount(r.pod_node, minlength=snap.n), cap)


def test_faithful_mode_same_count(built):
    # adaptive sampling + rotation changes the order, not the closed-form count (KA5)
    snap, tmpl, ctr = synth.c2(n=1200)
    a = oracle.run(snap, tmpl, ctr)
    b = oracle.run(snap, tmpl, ctr, mode=1)
    assert a.placed == b.placed and b.evals < a.evals
    assert np.array_equal(np.bincount(a.pod_node, minlength=snap.n), np.bincount(b.pod_node, minlength=snap.n))


def test_empty_cluster(built):
    snap = abi.Snapshot(0, np.zeros(0), np.zeros(0), np.zeros(0))
    r = oracle.run(snap, [abi.default_template(100, 100)], cap=4)
    assert r.placed == 0 and r.stop_code == abi.STOP_UNSCHEDULABLE


def test_histogram_sorting_as_strings():
    # framework/types.go:816-824: "<count> <reason>" strings are sorted lexically: "10 Insufficient cpu" < "2 Too many pods"
    msg = abi.fit_error_message(12, {abi.R_INSUFFICIENT_CPU: 10, abi.R_TOO_MANY_PODS: 2}, 12, 0, lambda r: abi.REASON_TEXT[r])
    assert msg.startswith("0/12 nodes are available: 10 Insufficient cpu, 2 Too many pods. preemption: 0/12")


@pytest.mark.parametrize("which", ["c3", "c4", "c5"])
def test_memoised_oracle_is_identical(which):
    """oracle.run(memo=True) (used by the full-size GPU parity tests) gives exactly the plain restatement's result, and the
    OpenMP split of filter / score / arg-max does not depend on the thread count."""
    synth = importlib.import_module("cluster-capacity_b200.synth")
    snap, tmpl, ctr = {"c3": lambda: synth.c3(n=6000, prefer_taints=True), "c4": lambda: synth.c4(n=8000, n_existing=16000, zones=8, racks=64, regions=4),
                       "c5": lambda: synth.c5(n=9000, n_templates=7)}[which]()
    a = oracle.run(snap, tmpl, ctr, max_pods=3000, threads=1)
    for kw in (dict(threads=1, memo=True), dict(threads=5), dict(threads=8, memo=True)):
        b = oracle.run(snap, tmpl, ctr, max_pods=3000, **kw)
        assert (a.placed, a.stop_code, a.evals) == (b.placed, b.stop_code, b.evals)
        assert np.array_equal(a.pod_node, b.pod_node) and np.array_equal(a.reason_hist, b.reason_hist)
