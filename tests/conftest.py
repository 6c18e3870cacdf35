import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    out = {"params": {}, "grads": {}, "state_in": {}, "state_out": {}}
    for k, v in d.items():
        if k.startswith("param:"):
            out["params"][k[6:]] = torch.from_numpy(v)
        elif k.startswith("grad:"):
            out["grads"][k[5:]] = torch.from_numpy(v)
        elif k.startswith("state_in:"):
            out["state_in"][k[9:]] = torch.from_numpy(v)
        elif k.startswith("state_out:"):
            out["state_out"][k[10:]] = torch.from_numpy(v)
        elif k in ("cfg", "constraints"):
            out[k] = json.loads(str(v))
        elif v.ndim == 0:
            out[k] = v.item()
        else:
            out[k] = torch.from_numpy(v)
    return out


@pytest.fixture
def golden():
    return load_golden
