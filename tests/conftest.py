import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "tests"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "chain_golden.npz")
    data = np.load(path)
    meta = json.loads(bytes(data["__meta__"]).decode())

    def last_frame(case_name, variant):
        arr = data[f"{case_name}/{variant}"]
        if arr.size == 0:  # export entry identical to the gui one
            arr = data[f"{case_name}/gui"]
        return arr

    return {"meta": meta, "last_frame": last_frame}
