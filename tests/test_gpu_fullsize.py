"""GPU: parity at BASELINE.json's FULL sizes — the configurations bench.py times, compared with the oracle to the end.

The oracle runs with all host threads and the memoised node score (oracle/binding.run(memo=True): identical results,
tests/test_oracle_known_answers.py::test_memoised_oracle_is_identical), which keeps each case at 10-30 s of CPU. Next to the
bit-exact comparison every case checks the size-independent properties the domain offers: KA5's closed form for node-local
templates (count = sum of per-node capacities, distribution = capacities), per-template counts, sequence checksums.
Reference contract: pkg/framework/simulator.go:297-354 (bind / limit / stop), schedule_one.go:430-478."""
import importlib
import os

import numpy as np
import pytest

abi = importlib.import_module("cluster-capacity_b200._abi")
synth = importlib.import_module("cluster-capacity_b200.synth")
from oracle import binding as oracle  # noqa: E402

pytestmark = pytest.mark.gpu
THREADS = max(1, min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))


def gpu(snap, tmpl, ctr, max_pods=0, engine_kind=abi.ENGINE_AUTO):
    engine = importlib.import_module("cluster-capacity_b200.engine")
    with engine.Engine(device=0, engine=engine_kind) as eng:
        eng.load_nodes(snap)
        eng.set_templates(tmpl, ctr)
        res = eng.run(max_pods)
        counts = [eng.node_counts(t)[0] for t in range(min(len(tmpl), 4))]
    return res, counts


def same(got, want):
    assert got.placed == want.placed and got.stop_code == want.stop_code
    assert np.array_equal(got.pod_node, want.pod_node), "placement sequence differs from the oracle at pod %d" % int(
        np.nonzero(got.pod_node[:min(len(got.pod_node), len(want.pod_node))] != want.pod_node[:min(len(got.pod_node), len(want.pod_node))])[0][0])
    assert np.array_equal(got.reason_hist, want.reason_hist)
    assert (got.preempt_no_victims, got.preempt_not_helpful) == (want.preempt_no_victims, want.preempt_not_helpful)


def test_c4_full_100k_to_unschedulable(built):
    """The bench workload itself: 100k nodes, 3 spread constraints + hostname anti-affinity, 200k existing pods, unlimited."""
    snap, tmpl, ctr = synth.c4()
    want = oracle.run(snap, tmpl, ctr, threads=THREADS, memo=True)
    got, counts = gpu(snap, tmpl, ctr)
    same(got, want)
    assert want.stop_code == abi.STOP_UNSCHEDULABLE and want.placed > 30000
    assert np.array_equal(counts[0], np.bincount(want.pod_node, minlength=snap.n))
    assert counts[0].max() == 1                      # hostname anti-affinity: one clone per node
    seq, _ = gpu(snap, tmpl, ctr, engine_kind=abi.ENGINE_SEQUENTIAL)
    same(seq, want)
    assert seq.evals == want.evals == (want.placed + 1) * snap.n


def test_c3_full_50k_to_unschedulable(built):
    """50k nodes, nodeSelector + 3 tolerations, full default Filter set, unlimited: bit-exact and KA5's closed form."""
    snap, tmpl, ctr = synth.c3()
    want = oracle.run(snap, tmpl, ctr, threads=THREADS, memo=True)
    got, counts = gpu(snap, tmpl, ctr)
    same(got, want)
    t = tmpl[0]
    ok = ((snap.taint_mask[0] & np.uint64(snap.taint_nosched[0]) & ~np.uint64(t.tol_nosched[0]) & ~np.uint64(1 << 63)) == 0) \
        & ((snap.taint_mask[0] >> np.uint64(63)) == 0) & ((snap.static_mask[0] & np.uint64(t.sel_mask[0])) == np.uint64(t.sel_mask[0]))
    cap = np.where(ok, synth.closed_form_capacity(snap, t), 0)
    assert got.placed == int(cap.sum()) and np.array_equal(counts[0].astype(np.int64), cap)


def test_c5_1m_nodes_64_templates_limit_6400(built):
    """1M nodes x 64 podspecs placed round-robin (report.go:160), --max-limit 6400: the streaming kernel (the tile does not fit
    in shared memory) against the oracle, per-template counts included."""
    snap, tmpl, ctr = synth.c5()
    limit = 6400
    want = oracle.run(snap, tmpl, ctr, max_pods=limit, threads=THREADS, memo=True)
    got, counts = gpu(snap, tmpl, ctr, max_pods=limit)
    same(got, want)
    assert got.placed == limit and got.stop_code == abi.STOP_LIMIT_REACHED
    for t in range(len(counts)):
        assert np.array_equal(counts[t], np.bincount(want.pod_node[t::len(tmpl)], minlength=snap.n))
