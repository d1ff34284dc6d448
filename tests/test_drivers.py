"""The three example drivers (drivers/*_Simulation.py) against the reference drivers' contract: the names they import from `src.*`
and the variables they write to the .mat file (tests/golden/mdict_schema.json, frozen from the reference by
tests/golden/make_mdict_schema.py: SingleMassOscillator_Simulation.py:94-124, VehicleSimulation_Simulation.py:105-154,
EMPS_Simulation.py:128-160)."""
import ast
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHEMA = json.load(open(os.path.join(ROOT, "tests", "golden", "mdict_schema.json")))
DRIVERS = sorted(SCHEMA)


@pytest.mark.parametrize("name", DRIVERS)
def test_driver_imports_the_reference_names(name):
    tree = ast.parse(open(os.path.join(ROOT, "drivers", name + ".py")).read())
    got = {f"{n.module}.{a.name}" for n in ast.walk(tree) if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("src")
           for a in n.names}
    assert set(SCHEMA[name]["imports"]) <= got, sorted(set(SCHEMA[name]["imports"]) - got)
    assert os.path.basename(SCHEMA[name]["file"]) in open(os.path.join(ROOT, "drivers", name + ".py")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("name", DRIVERS)
def test_driver_writes_the_reference_mat_schema(name, tmp_path):
    out = tmp_path / "result.mat"
    K, N = 4, 48
    cmd = [sys.executable, os.path.join(ROOT, "drivers", name + ".py"), "--iterations", str(K), "--particles", str(N),
           "--pgas-iterations", "3", "--out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    import scipy.io
    md = scipy.io.loadmat(str(out))
    keys = {k for k in md if not k.startswith("__")}
    assert keys == set(SCHEMA[name]["keys"]), (sorted(set(SCHEMA[name]["keys"]) - keys), sorted(keys - set(SCHEMA[name]["keys"])))
    T = md["time"].size
    # shapes follow src/Algorithm1.py:399-492 and src/Algorithm2.py:106-187: online traces over particles, offline over iterations
    assert md["online_Sigma_X"].shape[:2] == (T, N) and md["offline_Sigma_X"].shape[:2] == (T, K)
    assert md["online_weights"].shape == (T, N) and md["offline_weights"].shape == (T, K)
    assert md["online_log_likelihood"].shape == (T, N) and md["offline_log_likelihood"].shape == (T, K)
    for sfx in ("", "_f", "_r"):
        if "offline_T1" + sfx in md:
            M = md["prior_T1" + sfx].shape[0]
            assert md["offline_T1" + sfx].shape == (K, M, M) and md["online_T1" + sfx].shape == (T, M, M)
            assert md["offline_T0" + sfx].shape[:2] == (K, M) and md["online_T0" + sfx].shape[:2] == (T, M)
    for k in keys:
        assert np.all(np.isfinite(np.asarray(md[k], dtype=np.float64))), k
    if name == "EMPS_Simulation":
        assert md["offline_Sigma_X_PGAS"].shape[:2] == (T, 3)
