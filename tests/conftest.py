import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
LAM, GLAM = 500.0, 10.0


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def view_from_fixture(rec):
    from ms_slam_b200 import WindowView
    return WindowView(K=rec["K"], H=rec["H"], feat_ptr=np.array(rec["feat_ptr"], np.int32),
                      feat_mp=np.array(rec["feat_mp"], np.int32), feat_cell=np.array(rec["feat_cell"], np.uint16),
                      mp_nobs=np.array(rec["mp_nobs"], np.int32), mp_obs_ptr=np.array(rec["mp_obs_ptr"], np.int32),
                      mp_obs_kf=np.array(rec["mp_obs_kf"], np.int32), okf_total=np.array(rec["okf_total"], np.int32))


def fixture_key(name, seed, over=None):
    return f"{name}:{seed}:{json.dumps(over or {}, sort_keys=True)}"


@pytest.fixture(scope="session")
def known_answers():
    return load_json("known_answers.json")


@pytest.fixture(scope="session")
def config_bounds():
    return load_json("config_bounds.json")


@pytest.fixture(scope="session")
def emulation_golden():
    return load_json("emulation.json")


@pytest.fixture(scope="session")
def build_native():
    """compile (if stale) and load libmss.so; CPU-only boxes can still dlopen it"""
    import __graft_entry__ as ge
    ge.build()
    from ms_slam_b200 import engine
    return engine.load_library()
