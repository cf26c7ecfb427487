import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """Builds libm2m_b200.so if it is missing or older than its sources (nvcc cross-compiles without a GPU)."""
    from music2midi_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def state_dict():
    from music2midi_b200 import synthetic as syn

    return syn.synthetic_state_dict(0)


@pytest.fixture(scope="session")
def oracle_weights(state_dict):
    from oracle import port

    return port.Weights(state_dict)


def _engine(state_dict, precision):
    import torch

    from music2midi_b200.engine import Engine

    eng = Engine(torch.device("cuda", 0), precision=precision)
    eng.load_state_dict(state_dict)
    return eng


@pytest.fixture(scope="session")
def engine_fp32(state_dict):
    return _engine(state_dict, "fp32")


@pytest.fixture(scope="session")
def engine_bf16(state_dict):
    return _engine(state_dict, "bf16")


@pytest.fixture(scope="session")
def report():
    """Appends one JSON line per measurement to gpurun_out/parity_report.jsonl (travels back)."""
    import json

    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "parity_report.jsonl")

    def _w(**kw):
        with open(path, "a") as f:
            f.write(json.dumps(kw) + "\n")
        print("REPORT", json.dumps(kw))

    return _w
