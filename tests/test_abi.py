"""The C-ABI library loads and exports every symbol include/blp_b200.h declares.
No compute: argument validation paths only (they return before touching CUDA)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "blp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(blp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from blp_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        from blp_b200.build import build
        build()
    return _lib.lib()


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for name in ("blp_score_bcast", "blp_rank_counts", "blp_metrics_from_counts", "blp_metrics_reduce",
                 "blp_eval_rank", "blp_train_loss", "blp_train_workspace_bytes", "blp_pair_loss",
                 "blp_l2_regularization", "blp_scale", "blp_version", "blp_last_error", "blp_device_check"):
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    from blp_b200 import _lib
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/blp_b200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in blp_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(declared_symbols())


def test_version_matches_header(lib):
    ver = int(re.search(r"#define BLP_B200_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.blp_version() == ver


def test_argument_validation_returns_codes(lib):
    # unknown model -> BLP_EINVAL, message names the problem (models.py:26 raises ValueError)
    rc = lib.blp_eval_rank(9, None, 0, 0, 128, None, None, None, 1, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"model" in lib.blp_last_error()
    # odd width for complex -> BLP_EDIM
    rc = lib.blp_score_bcast(2, None, 0, 0, None, 0, 0, None, 0, 0, 1, 1, 7, None, None)
    assert rc == -2
    # null pointers -> BLP_EINVAL
    rc = lib.blp_train_loss(0, 0, None, None, None, 5, None, 0, 0, 0, 4, 4, 128, 0.0, None, None, None, None, None, None, None)
    assert rc == -1
    rc = lib.blp_pair_loss(7, None, None, 0, 1, 1, None, None, None, None)
    assert rc == -1
    assert lib.blp_train_workspace_bytes(64, 512) >= 64


def test_python_wrapper_maps_codes_to_reference_exceptions():
    import blp_b200
    from blp_b200 import ops
    with pytest.raises(ValueError, match="Unknown relational model"):
        blp_b200.LinkPrediction(8, "rotate", "margin", 3, 0)
    with pytest.raises(ValueError, match="Unkown loss function"):
        blp_b200.LinkPrediction(8, "transe", "hinge", 3, 0)
    with pytest.raises(ValueError):
        ops.model_id("nope")


def test_no_cpu_fallback():
    """CPU tensors must raise, not silently run somewhere else."""
    import torch
    import blp_b200
    x = torch.zeros(2, 1, 8)
    with pytest.raises(blp_b200.BlpError):
        blp_b200.transe_score(x, x, x)
    with pytest.raises(blp_b200.BlpError):
        blp_b200.get_metrics(torch.zeros(2, 5), torch.zeros(2, 1, dtype=torch.long), torch.tensor([[1, 3]]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "blp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libblp_oracle" not in text, f
