"""GPU: randomized differential test of the engines ENGINE_AUTO picks (multi-commit waves, tie-run batching, lean, generic) against the
CPU oracle: spread / anti-affinity templates with random domain counts, skews, self-match flags, missing labels, minDomains, limits."""
import importlib

import numpy as np
import pytest

abi = importlib.import_module("cluster-capacity_b200._abi")
from oracle import binding as oracle  # noqa: E402

pytestmark = pytest.mark.gpu
GiB, MiB = 1 << 30, 1 << 20


def random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([3000, 9000, 40000, 90000]))
    n_topo = int(rng.integers(1, 4))
    doms = [int(rng.choice([3, 8, 40, 300, 2000])) for _ in range(n_topo)]
    topo = []
    for d in doms:
        col = rng.integers(0, d, n).astype(np.int32)
        if rng.random() < 0.4:
            col[rng.random(n) < 0.03] = -1            # nodes without the label
        topo.append(col)
    a_cpu = rng.choice([2000, 4000, 8000, 16000], n)
    npods = rng.integers(0, 20, n).astype(np.int32)
    req_cpu = (rng.random(n) * 0.5 * a_cpu).astype(np.int64) // 10 * 10
    snap = abi.Snapshot(n, a_cpu, a_cpu * (2 * MiB), rng.choice([30, 60, 110], n), req_cpu=req_cpu, req_mem=req_cpu * (1 * MiB),
                        npods=npods, topo=topo)
    ctr, t = [], abi.default_template(int(rng.choice([100, 250, 700])), int(rng.choice([64, 256, 1024])) * MiB)
    n_pts = 0
    for c, d in enumerate(doms):
        init = rng.integers(0, 4, d).astype(np.int32) if rng.random() < 0.7 else np.full(d, int(rng.integers(0, 3)), np.int32)
        self_match = int(rng.random() < 0.85)
        n_present = d if rng.random() < 0.8 else max(1, d - int(rng.integers(1, 3)))
        ctr.append(abi.make_counter(c, init, n_present=n_present, inc=self_match))
        t.pts[n_pts].counter, t.pts[n_pts].max_skew = c, int(rng.choice([1, 1, 2, 5]))
        t.pts[n_pts].self_match, t.pts[n_pts].min_zero = self_match, int(rng.random() < 0.1)
        n_pts += 1
    t.n_pts = n_pts
    kind = rng.random()
    if kind < 0.5:                                    # required anti-affinity on the hostname: node-local counter
        ctr.append(abi.make_counter(-1, (rng.random(n) < 0.1).astype(np.int32), inc=1))
        t.n_anti, t.anti_counter[0] = 1, len(ctr) - 1
    elif kind < 0.7 and doms[0] >= 40:                # ... or on the first topology key (its own counter on the same column)
        ctr.append(abi.make_counter(0, (rng.random(doms[0]) < 0.2).astype(np.int32), inc=1))
        t.n_anti, t.anti_counter[0] = 1, len(ctr) - 1
    limit = int(rng.choice([0, 0, 1, 37, 1500]))
    return snap, [t], ctr, limit


@pytest.mark.parametrize("seed", range(24))
def test_random_coupled_templates(built, seed):
    snap, tmpl, ctr, limit = random_case(seed)
    cap = limit or 4000                               # keep the single-thread oracle in seconds
    want = oracle.run(snap, tmpl, ctr, max_pods=cap, threads=8)
    engine = importlib.import_module("cluster-capacity_b200.engine")
    for kind in (abi.ENGINE_AUTO, abi.ENGINE_SEQUENTIAL):
        with engine.Engine(device=0, engine=kind) as eng:
            eng.load_nodes(snap)
            eng.set_templates(tmpl, ctr)
            got = eng.run(cap)
        assert got.placed == want.placed and got.stop_code == want.stop_code, (seed, kind)
        assert np.array_equal(got.pod_node, want.pod_node), (seed, kind)
        assert np.array_equal(got.reason_hist, want.reason_hist), (seed, kind)
