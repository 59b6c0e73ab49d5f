"""blp_b200.patch() against the UNMODIFIED reference modules (only where /root/reference is mounted: the build
container; skipped on the GPU box).  No compute: checks the seam -- attribute binding by name (models.py:16-36),
inheritance of compute_loss by every model class, utils.get_metrics -- and that CPU tensors are refused, not
silently routed to a fallback."""
import os
import sys
import types

import pytest
import torch

import blp_b200

REF = os.environ.get("BLP_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models.py")), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    """Import reference models / utils with the absent nltk dependency stubbed (data.py is not needed here)."""
    saved = {k: sys.modules.get(k) for k in ("models", "utils")}
    sys.path.insert(0, REF)
    try:
        for k in ("models", "utils"):
            sys.modules.pop(k, None)
        import models  # noqa
        import utils  # noqa
        originals = {name: getattr(models, name) for name in ("transe_score", "distmult_score", "complex_score", "simple_score",
                                                                "margin_loss", "nll_loss", "l2_regularization")}
        originals["compute_loss"] = models.LinkPrediction.compute_loss
        originals["get_metrics"] = utils.get_metrics
        yield models, utils
        for name, fn in originals.items():
            if name == "compute_loss":
                models.LinkPrediction.compute_loss = fn
            elif name == "get_metrics":
                utils.get_metrics = fn
            else:
                setattr(models, name, fn)
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_patch_rebinds_the_hot_path_inside_the_reference_modules(ref):
    models, utils = ref
    before = models.transe_score
    assert blp_b200.patch(models, utils) is models
    assert models.transe_score is blp_b200.transe_score and models.transe_score is not before
    for name in ("distmult_score", "complex_score", "simple_score", "margin_loss", "nll_loss", "l2_regularization"):
        assert getattr(models, name) is getattr(blp_b200.models, name)
    assert utils.get_metrics is blp_b200.get_metrics
    # LinkPrediction.__init__ looks score_fn / loss_fn up by module attribute at construction time (models.py:16-36),
    # so every reference model class built after patch() carries the CUDA functions and the fused compute_loss
    for rel_model, fn in (("transe", blp_b200.transe_score), ("distmult", blp_b200.distmult_score),
                          ("complex", blp_b200.complex_score), ("simple", blp_b200.simple_score)):
        m = models.TransductiveLinkPrediction(128, rel_model, "margin", 10, 3, 0)
        assert m.score_fn is fn and m.loss_fn is blp_b200.margin_loss
        assert type(m).compute_loss is blp_b200.compute_loss
        assert m.normalize_embs == (rel_model == "transe")                       # models.py:18 untouched
    with pytest.raises(ValueError):
        models.TransductiveLinkPrediction(128, "rotate", "margin", 10, 3, 0)     # models.py:26 still the reference's check
    # the inductive classes inherit the same method
    assert models.InductiveLinkPrediction.compute_loss is blp_b200.compute_loss


def test_patched_model_refuses_cpu_tensors(ref):
    """There is no CPU / PyTorch fallback behind the seam: a patched reference model on the CPU raises."""
    models, utils = ref
    blp_b200.patch(models, utils)
    m = models.TransductiveLinkPrediction(128, "transe", "margin", 10, 3, 0)
    pairs, rels = torch.randint(0, 10, (4, 2)), torch.randint(0, 3, (4, 1))
    neg = torch.randint(0, 8, (4, 5, 2))
    with pytest.raises(blp_b200.BlpError):
        m(pairs, rels, neg)
    with pytest.raises(blp_b200.BlpError):
        utils.get_metrics(torch.randn(4, 10), torch.zeros(4, 1, dtype=torch.long), torch.tensor([[1, 3, 10]]))
