"""GPU (needs >= 2 devices, skipped otherwise): the node-sharded run — one process per GPU, per-wave exchange of the shard
winners through peer memory inside the persistent kernel — must give exactly the single-GPU / oracle result."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, which, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = importlib.import_module("cluster-capacity_b200.synth")
    engine = importlib.import_module("cluster-capacity_b200.engine")
    sharded = importlib.import_module("cluster-capacity_b200.sharded")
    from oracle import binding as oracle
    if which == "c3":
        snap, tmpl, ctr = synth.c3(n=5001, prefer_taints=True)
        limit = 0
    else:
        snap, tmpl, ctr = synth.c4(n=6000, n_existing=12000, zones=8, racks=64, regions=4)
        limit = 0
    torch.cuda.set_device(rank)
    eng = engine.Engine(device=rank, rank=rank, world=world)
    eng.load_nodes(snap)
    eng.set_templates(tmpl, ctr)
    eng.connect_peers(dist)
    ok = True
    for _ in range(2):          # twice: the run epoch must keep words of the previous run from validating
        dist.barrier()
        res = eng.run(limit)
        m = sharded.merge_results(dist, res)
        want = oracle.run(snap, tmpl, ctr, max_pods=limit, threads=4)
        ok &= (m["placed"] == want.placed and m["stop_code"] == want.stop_code and np.array_equal(m["pod_node"], want.pod_node)
               and np.array_equal(m["reason_hist"], want.reason_hist) and m["preempt_no_victims"] == want.preempt_no_victims
               and m["evals"] == want.evals)
    q.put((rank, bool(ok), int(res.placed)))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("which", ["c3", "c4"])
def test_two_gpu_sharded_matches_oracle(built, which):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, which, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(o[1] for o in out), out
