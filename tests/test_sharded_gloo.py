"""CPU, world_size=2, gloo: the host logic of the node-sharded multi-GPU run — shard bounds match libccsim's split, and
merging per-shard results (histogram / preemption counts / evals summed, replicated parts cross-checked) reproduces
the single-shard oracle result. The per-wave exchange itself lives in the CUDA kernel and is tested on GPUs."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    abi = importlib.import_module("cluster-capacity_b200._abi")
    synth = importlib.import_module("cluster-capacity_b200.synth")
    sharded = importlib.import_module("cluster-capacity_b200.sharded")
    from oracle import binding as oracle
    snap, tmpl, ctr = synth.c3(n=1201, prefer_taints=True)
    full = oracle.run(snap, tmpl, ctr)
    lo, hi = sharded.shard_bounds(snap.n, world, rank)

    class Part:   # what one rank's ccsim_run returns: its shard's slice of the terminal diagnosis, the replicated rest
        pass
    part = Part()
    part.placed, part.stop_code, part.pod_node, part.waves, part.n_nodes = full.placed, full.stop_code, full.pod_node, full.waves, snap.n
    # per-shard diagnosis: rerun the oracle's terminal pass restricted to [lo,hi) by masking the other nodes as "not mine"
    hist = np.zeros_like(full.reason_hist)
    sub = abi.Snapshot(hi - lo, snap.alloc_cpu[lo:hi], snap.alloc_mem[lo:hi], snap.alloc_pods[lo:hi],
                       req_cpu=snap.req_cpu[lo:hi] + np.bincount(full.pod_node, minlength=snap.n)[lo:hi] * tmpl[0].req_cpu,
                       req_mem=snap.req_mem[lo:hi] + np.bincount(full.pod_node, minlength=snap.n)[lo:hi] * tmpl[0].req_mem,
                       npods=snap.npods[lo:hi] + np.bincount(full.pod_node, minlength=snap.n)[lo:hi].astype(np.int32),
                       taint_mask=snap.taint_mask[:, lo:hi], taint_nosched=snap.taint_nosched, taint_prefer=snap.taint_prefer,
                       static_mask=snap.static_mask[:, lo:hi],
                       taint_lists=[snap.taint_list[snap.taint_list_off[i]:snap.taint_list_off[i + 1]].tolist() for i in range(lo, hi)])
    sub_res = oracle.run(sub, tmpl, ctr)          # the shard is full: nothing fits, only the diagnosis runs
    assert sub_res.placed == 0
    part.reason_hist, part.preempt_no_victims, part.evals = sub_res.reason_hist, sub_res.preempt_no_victims, full.waves * (hi - lo)
    part.preempt_not_helpful = sub_res.preempt_not_helpful      # per shard, like the histogram
    merged = sharded.merge_results(dist, part)
    ok = (np.array_equal(merged["reason_hist"], full.reason_hist) and merged["preempt_no_victims"] == full.preempt_no_victims
          and merged["preempt_not_helpful"] == full.preempt_not_helpful and merged["evals"] == full.evals)
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_shard_bounds_cover_and_match_engine_split():
    sharded = importlib.import_module("cluster-capacity_b200.sharded")
    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            b = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            for node in (0, n // 2, n - 1):
                if n:
                    r = sharded.owner_of(node, n, world)
                    assert b[r][0] <= node < b[r][1]


def test_merge_two_ranks_gloo(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(o[1] for o in out), out
