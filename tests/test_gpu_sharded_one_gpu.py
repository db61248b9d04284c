"""GPU (ONE device is enough): the node-sharded engines with all ranks on device 0. Every rank is a handle of this process, the
exchange buffers are wired by pointer (ccsim_peer_import_local) and the ranks' persistent kernels run concurrently on different
streams (small clusters: every rank needs only a few SMs), started from one host thread per rank. The same in-kernel exchange
as across GPUs — candidate lines of every CTA into every rank's buffer (multi-commit), winner words (lean, streaming) — only
the stores do not cross NVLink. Results must equal the oracle's, like tests/test_gpu_sharded.py on a multi-GPU box."""
import importlib
import threading

import numpy as np
import pytest

abi = importlib.import_module("cluster-capacity_b200._abi")
synth = importlib.import_module("cluster-capacity_b200.synth")
from oracle import binding as oracle  # noqa: E402

pytestmark = pytest.mark.gpu


def run_sharded(snap, tmpl, ctr, limit, world, kind, runs):
    engine = importlib.import_module("cluster-capacity_b200.engine")
    engs = [engine.Engine(device=0, engine=kind, rank=r, world=world) for r in range(world)]
    for e in engs:
        e.load_nodes(snap)
        e.set_templates(tmpl, ctr)
    engine.Engine.connect_local(engs)
    out = []
    for lim in runs:
        res, errs = [None] * world, []
        for e in engs:              # every rank past its allocations before any rank's kernel starts waiting for its peers
            e.prepare(lim)

        def work(r):
            try:
                res[r] = engs[r].run(lim)
            except Exception as ex:       # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert not errs, errs
        assert all(r is not None for r in res), "a rank did not finish"
        out.append((res, [e.run_stats() for e in engs]))
    for e in engs:
        e.close()
    return out


CASES = {
    "c3": (lambda: synth.c3(n=5001, prefer_taints=True), 0, None),
    "c4": (lambda: synth.c4(n=6000, n_existing=12000, zones=8, racks=64, regions=4), 0, "multi-commit"),
    "c4_limit": (lambda: synth.c4(n=9000, n_existing=15000, zones=16, racks=128, regions=4), 700, "multi-commit"),
    "spread": (lambda: (lambda s, t, c: (s, [_no_anti(t[0])], c[:3]))(*synth.c4(n=7000, n_existing=9000, zones=8, racks=64, regions=4)), 900, "multi-commit"),
    "c5": (lambda: synth.c5(n=9001, n_templates=9), 1200, "streaming"),
}


def _no_anti(t):
    t.n_anti = 0
    return t


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("which", sorted(CASES))
def test_sharded_engines_on_one_gpu_match_the_oracle(built, which, world):
    make, limit, engine_name = CASES[which]
    snap, tmpl, ctr = make()
    runs = [limit, (limit or 0) // 2 + 7, limit]        # several runs per handle: epoch / buffer parity carry over
    wants = [oracle.run(snap, tmpl, ctr, max_pods=lim, threads=4, memo=True) for lim in runs]
    for kind in (abi.ENGINE_AUTO, abi.ENGINE_SEQUENTIAL):
        for (res, stats), w in zip(run_sharded(snap, tmpl, ctr, limit, world, kind, runs), wants):
            for r in res:       # replicated parts: identical on every rank
                assert r.placed == w.placed and r.stop_code == w.stop_code
                assert np.array_equal(r.pod_node, w.pod_node), "rank sequence differs from the oracle at pod %d" % int(
                    np.nonzero(r.pod_node[:min(len(r.pod_node), len(w.pod_node))] != w.pod_node[:min(len(r.pod_node), len(w.pod_node))])[0][0])
            # per-shard parts sum up
            assert np.array_equal(sum(r.reason_hist for r in res), w.reason_hist)
            assert sum(r.preempt_no_victims for r in res) == w.preempt_no_victims
            assert sum(r.preempt_not_helpful for r in res) == w.preempt_not_helpful
            if kind == abi.ENGINE_SEQUENTIAL:
                assert sum(r.evals for r in res) == w.evals
            elif engine_name:
                assert all(engine_name in s["engine"] for s in stats), stats
