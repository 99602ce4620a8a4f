import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load_golden(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["name"] = os.path.basename(path)[:-4]
    return d


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def capi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as ge
    ge.build()
    from caffe_escoin_b200 import capi as c
    return c


def geom_from_golden(po, d):
    N, Cin, Cout, H, k, s, p, dil, grp = [int(v) for v in d["geom"]]
    return po.Geom(N, Cin, H, H, Cout, k, s, p, dil, grp)
