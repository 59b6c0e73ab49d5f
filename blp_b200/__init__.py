"""blp_b200 -- B200-native (sm_100a) scoring / loss / ranking hot path of dfdazac/blp.

Public surface (same names as the reference's models.py / utils.py):

    transe_score, distmult_score, complex_score, simple_score   models.py:222-248
    margin_loss, nll_loss, l2_regularization                    models.py:251-266
    LinkPrediction(.compute_loss), InductiveLinkPrediction,
    TransductiveLinkPrediction                                   models.py:7-93, 207-219
    get_metrics                                                  utils.py:86-111
    rank_sweep                                                   train.py:128-171 (fused)
    patch(models, utils)                                         rebinds the path inside the reference's modules

All computation runs in libblp_b200.so (blp_b200/csrc, C ABI in include/blp_b200.h).
"""
from . import _lib, lazy, ops  # noqa: F401
from .ops import store_rows  # noqa: F401
from ._lib import BlpError  # noqa: F401
from .graphs import GraphedLossStep  # noqa: F401
from .evaluate import (AlignedTriples, RankSweepPlan, breakdowns, finalize, gather_rows, rank_sweep, shard_bounds,  # noqa: F401
                       topk_sweep)  # noqa: F401
from .models import (InductiveLinkPrediction, LinkPrediction, TransductiveLinkPrediction,  # noqa: F401
                     complex_score, compute_loss, distmult_score, fused_compute_loss, l2_regularization,
                     margin_loss, nll_loss, simple_score, transe_score)
from .utils import (DeviceFilterIndex, TripleFilterIndex, get_metrics, get_negative_sampling_indices,  # noqa: F401
                    graph_edges, make_ent2idx, split_by_category, split_by_new_position)  # noqa: F401

__version__ = "0.1.0"


def patch(models_module, utils_module=None, lazy_scores=True):
    """Rebind the hot path inside the reference's own modules.

    After `import models, utils; blp_b200.patch(models, utils)` every reference model class
    (BertEmbeddingsLP, BOW, DKRL, ...) built afterwards binds the CUDA score / loss functions
    (models.py:16-24, 31-34), `compute_loss` is the fused kernel, and train.py's
    `utils.get_metrics(...)` / `utils.split_by_new_position` / `utils.split_by_category` calls run on the GPU (one launch
    each instead of per-triple Python loops).  train.py itself is untouched.

    lazy_scores (default on): `score_fn` on the eval broadcast of train.py:146-147 returns a score-matrix HANDLE
    (blp_b200.lazy.LazyScores) that torch.cat, utils.get_metrics and the filter statement train.py:165 consume without
    materialising the (2B, N) matrix -- the unmodified eval loop then runs the fused sweep kernel, one launch per
    batch -- and utils.get_triple_filters returns filter handles backed by a device-resident index built once per
    filtering graph.  Any other use of a handle materialises it with the exact kernels.
    """
    from . import models as _m
    for name in ("transe_score", "distmult_score", "complex_score", "simple_score",
                 "margin_loss", "nll_loss", "l2_regularization"):
        setattr(models_module, name, getattr(_m, name))
    models_module.LinkPrediction.compute_loss = _m.compute_loss
    if utils_module is not None:
        utils_module.get_metrics = get_metrics
        # train.py:173-188: the per-triple Python loops over device tensors (~10 ms per batch of 64 on a GPU) -> one launch each
        utils_module.split_by_new_position = split_by_new_position
        utils_module.split_by_category = split_by_category
        orig = getattr(utils_module.get_triple_filters, "__wrapped__", utils_module.get_triple_filters)
        utils_module.get_triple_filters = lazy.make_get_triple_filters(orig) if lazy_scores else orig
    lazy.enable(lazy_scores)
    return models_module
