import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def unpack_mask(g):
    n = g["ent"].shape[0]
    return np.unpackbits(g["filter_mask"], axis=1)[:, :n].astype(bool)


def mask_to_csr(mask):
    indptr = np.concatenate([[0], np.cumsum(mask.sum(1))]).astype(np.int64)
    idx = np.nonzero(mask)[1].astype(np.int64)
    return indptr, idx


def strided_neg_idx(g):
    """Rebuild neg_idx with the reference sampler's non-contiguous strides (data.py:78-79)."""
    import torch
    neg = torch.from_numpy(g["neg_idx"])
    b, k, _ = neg.shape
    s = tuple(int(v) for v in g["neg_idx_strides"])
    base = torch.empty(max(1, (b - 1) * s[0] + (k - 1) * s[1] + s[2] + 1), dtype=torch.int64)
    view = base.as_strided((b, k, 2), s)
    view.copy_(neg)
    return view


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
