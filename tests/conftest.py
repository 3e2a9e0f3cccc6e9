import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
for p in (ROOT, ORACLE):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "needs_reference: needs the unmodified reference under /root/reference")


def pytest_collection_modifyitems(config, items):
    from ref_loader import reference_available
    have_ref = reference_available()
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference not present (GPU box)"))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def matlab():
    return dict(np.load(os.path.join(GOLDEN, "matlab_ldpc.npz")))


@pytest.fixture(scope="session")
def ref_cases():
    z = np.load(os.path.join(GOLDEN, "ref_cases.npz"))
    cases = {}
    for name in z["names"]:
        name = str(name)
        d = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
        m = d["meta"]
        d.update(bg=int(m[0]), A=int(m[1]), qm=int(m[2]), nl=int(m[3]), nref=int(m[4]), g=int(m[5]), nit=int(m[6]),
                 C=int(m[7]), Zc=int(m[8]), iLS=int(m[9]), K=int(m[10]), F=int(m[11]))
        cases[name] = d
    crc = {k[4:]: z[k] for k in z.files if k.startswith("crc/")}
    return cases, crc


MOD_NAME = {1: "BPSK", 2: "QPSK", 4: "16QAM", 6: "64QAM", 8: "256QAM", 10: "1024QAM"}
