import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def table_from_spec(spec: dict) -> np.ndarray:
    from amod_b200 import synth
    if spec["kind"] == "grid":
        return synth.make_road_graph(spec["n_nodes"], seed=spec["seed"]).tt
    return synth.random_table(spec["n_nodes"], spec["seed"], spec["kind"],
                              np.float64 if spec["dtype"] == "f64" else np.float32)


_tables = {}


def load_golden(name: str):
    """-> (tt, [(Epoch, Pool | None), ...]) for one golden family."""
    from amod_b200.marshal import Epoch, Pool
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    spec = json.loads(str(z["spec"]))
    key = json.dumps(spec, sort_keys=True)
    if key not in _tables:
        _tables[key] = table_from_spec(spec)
    cases = []
    for i in range(int(z["n_cases"])):
        ep = Epoch.from_npz_dict(z, f"c{i}_ep_")
        pool = None if f"c{i}_error" in z.files else Pool.from_npz_dict(z, f"c{i}_pool_")
        cases.append((ep, pool))
    return _tables[key], cases


def golden_families(prefixes=("rand_", "sim_", "levels")):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(tuple(prefixes)))
