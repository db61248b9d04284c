"""CPU: the persistent host thread pool of libcchost (ingest + encoder passes): several host threads driving analyses at once get the
same encoded snapshot as a lone caller, and a forked child starts a pool of its own."""
import ctypes as C
import importlib
import json
import os
import threading

fw = importlib.import_module("cluster-capacity_b200.framework")
synth = importlib.import_module("cluster-capacity_b200.synth")


def _inputs():
    nodes, pods, tmpl = synth.c4_objects(n=6000, n_existing=12000, zones=8, racks=64, regions=4)   # >= 4096 nodes: the parallel paths
    return json.dumps({"items": nodes}).encode(), json.dumps({"items": pods}).encode(), json.dumps(tmpl).encode()


def _one(L, nj, pj, tj):
    h = C.c_void_p()
    assert L.cc_new(None, tj, 0, None, 0, C.byref(h)) == 0
    assert L.cc_sync_with_objects(h, nj, pj, None) == 0
    d = json.loads(L.cc_debug_encoded_snapshot(h))
    L.cc_close(h)
    return d["nodes"]["req_cpu"], d["nodes"]["npods"], [c["init"] for c in d["counters"]]


def test_concurrent_callers_and_fork(built):
    L = fw.lib()
    nj, pj, tj = _inputs()
    ref = _one(L, nj, pj, tj)
    errs = []

    def caller():
        try:
            for _ in range(4):
                assert _one(L, nj, pj, tj) == ref
        except BaseException as e:      # noqa: BLE001 — reported below
            errs.append(e)
    th = [threading.Thread(target=caller) for _ in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs and not any(t.is_alive() for t in th)
    pid = os.fork()
    if pid == 0:                        # the child has none of the parent's worker threads: the pool must notice
        ok = False
        try:
            ok = _one(L, nj, pj, tj) == ref
        finally:
            os._exit(0 if ok else 3)
    import signal, time
    deadline, st = time.time() + 120, None
    while time.time() < deadline:
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            st = status
            break
        time.sleep(0.05)
    if st is None:
        os.kill(pid, signal.SIGKILL)
        os.waitpid(pid, 0)
    assert st is not None and os.WIFEXITED(st) and os.WEXITSTATUS(st) == 0
    assert _one(L, nj, pj, tj) == ref
