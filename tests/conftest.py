import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", params=["outside", "inside", "example", "moving"])
def golden(request):
    return request.param, dict(np.load(os.path.join(GOLDEN, "ref_%s.npz" % request.param)))


def golden_triplet(tag):
    from mrflow_b200 import synth
    kw = dict(outside=dict(shape=(72, 96), seed=7, epipole="outside", flow_noise=0.02),
              inside=dict(shape=(72, 96), seed=8, epipole="inside", moving_objects=True),
              example=dict(shape=(72, 96), seed=9, epipole="outside", flow_noise=0.05),
              moving=dict(shape=(72, 96), seed=11, epipole="outside", moving_objects=True, flow_noise=0.02))[tag]
    t = synth.make_triplet(n_waves=8, **kw)
    if tag == "example":
        # crops of the reference's example/frame1-3.png, stored in the fixture by make_golden.py
        g = np.load(os.path.join(GOLDEN, "ref_example.npz"))
        t["images"] = [g["images"][i] for i in range(3)]
    return t
