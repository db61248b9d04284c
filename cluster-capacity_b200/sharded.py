"""Node-sharded multi-GPU run: one process per GPU, rank r owns the contiguous node block
[r*ceil(N/W), (r+1)*ceil(N/W)) of the nodeTree order — the same split libccsim applies in ccsim_load_nodes.

The per-wave exchange of shard winners happens inside the persistent kernel over peer memory (engine.connect_peers);
torch.distributed only carries the IPC handles once and the final small reductions below. With the gloo backend the same
host logic runs on CPU (tests/test_sharded_gloo.py)."""
import numpy as np


def shard_bounds(n, world, rank):
    per = (n + world - 1) // world
    lo = min(per * rank, n)
    return lo, min(lo + per, n)


def owner_of(node, n, world):
    per = (n + world - 1) // world
    return min(node // per, world - 1) if per else 0


def merge_results(dist, res):
    """Sum the per-shard parts of a RunResult over the ranks (reason histogram, preemption counts, evals) and check that
    the replicated parts (placed, stop code, pod -> node sequence) agree. Returns a dict."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    vec = np.concatenate([res.reason_hist.astype(np.int64), [res.preempt_not_helpful, res.preempt_no_victims, res.evals]]).astype(np.int64)
    t = torch.from_numpy(vec).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    vec = t.cpu().numpy()
    # replicated parts: a checksum of the sequence must be identical everywhere
    chk = np.array([res.placed, res.stop_code, int(np.asarray(res.pod_node, np.int64).sum()),
                    int((np.asarray(res.pod_node, np.int64) * (np.arange(len(res.pod_node)) % 8191 + 1)).sum())], np.int64)
    lo = torch.from_numpy(chk.copy()).to(dev)
    hi = torch.from_numpy(chk.copy()).to(dev)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not bool((lo == hi).all()):
        raise RuntimeError("sharded run diverged between ranks: %s vs %s" % (lo.tolist(), hi.tolist()))
    return {"placed": int(res.placed), "stop_code": int(res.stop_code), "pod_node": res.pod_node,
            "reason_hist": vec[:-3], "preempt_no_victims": int(vec[-2]), "preempt_not_helpful": int(vec[-3]),   # per-shard parts: they sum up
            "evals": int(vec[-1]), "waves": int(res.waves)}
