"""CPU: the wave model behind the multi-commit kernel's look-ahead (scripts/wave_sim.py). Whatever is published and however the
reference cycles are grouped into waves — strict waves, the shipped look-ahead rule, look-ahead on every spread term — the model must
replay exactly the sequence the C oracle produces one cycle at a time; the look-ahead rule must need fewer waves than strict waves."""
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import binding as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
synth = importlib.import_module("cluster-capacity_b200.synth")
N = 30000
KW = dict(n=N, n_existing=60000, zones=32, racks=256, regions=8)


@pytest.fixture(scope="module")
def oracle_sequence(built):
    snap, tmpl, ctr = synth.c4(**KW)
    r = oracle.run(snap, tmpl, ctr, threads=4, memo=True)
    assert r.placed > 1000
    return r.pod_node[:r.placed]


def run_model(tmp_path, env, args=()):
    out = str(tmp_path / "seq.npy")
    e = dict(os.environ, WAVE_SIM_OUT=out, **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "wave_sim.py"), "--n", str(N), "--existing", str(KW["n_existing"]), "--zones", str(KW["zones"]),
                        "--racks", str(KW["racks"]), "--regions", str(KW["regions"])] + list(args), env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    waves = int(re.search(r"placed \d+ in (\d+) waves", r.stdout).group(1))
    return np.load(out), waves


def test_wave_model_reproduces_the_oracle_sequence(oracle_sequence, tmp_path):
    strict, w_strict = run_model(tmp_path, {})
    assert np.array_equal(strict, oracle_sequence)
    shipped, w_shipped = run_model(tmp_path, {"KNUM": "8", "RMAX": "3", "CF": "1"})
    assert np.array_equal(shipped, oracle_sequence)
    everywhere, w_every = run_model(tmp_path, {}, ["--relax", "1,1,2"])
    assert np.array_equal(everywhere, oracle_sequence)
    assert w_shipped < w_strict
