"""CUDA-graph replay of the launch-bound training step (models.py:51-70 forward + autograd backward).

At the reference's batch sizes (B = 64, K = 64..512) `compute_loss` is a few tens of microseconds of GPU work
behind several hundred microseconds of host work (autograd bookkeeping, allocator calls, the backward thread
hand-off).  The GPU side is already one memset + one fused kernel + one scale; this module removes the host
side by capturing the whole forward + backward once and replaying it: one graph launch per step.

    step = blp_b200.GraphedLossStep(model, batch_size=64, num_negatives=512)
    loss, grad_ent = step(ent_embs, rels, neg_idx)      # rel_emb.weight.grad is refreshed as well

The captured region is exactly `loss = model.compute_loss(x, rels, neg_idx); loss.backward()` (train.py:344-347
without the optimizer), so every number it produces is the eager path's number.  Inputs are copied into static
buffers that keep the reference sampler's strides for `neg_idx` (data.py:77-79).
"""
import torch


class GraphedLossStep:
    """`compute_loss` forward + backward for fixed (B, K) captured in a CUDA graph.

    model          a LinkPrediction (blp_b200 or patched reference) on a CUDA device
    batch_size     B (positive pairs per step); num_negatives K
    After `loss, grad_ent = step(ent_embs, rels, neg_idx)`:
      loss       0-dim tensor (static: overwritten by the next call)
      grad_ent   d loss / d ent_embs, (B, 2, D) (static); back-propagate into the encoder with
                 `ent_embs.backward(grad_ent)` when ent_embs has a graph behind it
      model.rel_emb.weight.grad   d loss / d rel_emb.weight (static, overwritten -- not accumulated)
    """

    def __init__(self, model, batch_size, num_negatives, warmup=3):
        weight = model.rel_emb.weight
        if not weight.is_cuda:
            raise RuntimeError("GraphedLossStep needs the model on a CUDA device; there is no CPU fallback")
        dev = weight.device
        self.model, self.device = model, dev
        b, k, d = int(batch_size), int(num_negatives), int(model.dim)
        self.ent_embs = torch.zeros((b, 2, d), dtype=torch.float32, device=dev, requires_grad=True)
        self.rels = torch.zeros((b, 1), dtype=torch.int64, device=dev)
        # storage (K, B, 2) viewed as (B, K, 2): the strides of data.get_negative_sampling_indices (data.py:77-79)
        self._neg_storage = torch.zeros((k, b, 2), dtype=torch.int64, device=dev)
        self._neg_storage[:, :, 1] = 1
        self.neg_idx = self._neg_storage.transpose(0, 1)

        def fwd_bwd():
            loss = model.compute_loss(self.ent_embs, self.rels, self.neg_idx)
            loss.backward()
            return loss

        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self.ent_embs.grad = None
                    weight.grad = None
                    fwd_bwd()
            torch.cuda.current_stream().wait_stream(side)
            self.ent_embs.grad = None
            weight.grad = None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = fwd_bwd()
        self.grad_ent = self.ent_embs.grad
        self.grad_rel_weight = weight.grad

    def __call__(self, ent_embs, rels, neg_idx):
        """Copy the step's inputs (device or pinned-host tensors) into the static buffers and replay."""
        with torch.no_grad():
            self.ent_embs.copy_(ent_embs.detach().reshape(self.ent_embs.shape), non_blocking=True)
            self.rels.copy_(rels.reshape(self.rels.shape), non_blocking=True)
            if neg_idx.shape == self.neg_idx.shape:
                self.neg_idx.copy_(neg_idx, non_blocking=True)
            else:                       # the sampler's (K, B, 2) storage passed as is
                self._neg_storage.copy_(neg_idx, non_blocking=True)
        self.graph.replay()
        if self.model.rel_emb.weight.grad is not self.grad_rel_weight:
            self.model.rel_emb.weight.grad = self.grad_rel_weight
        return self.loss, self.grad_ent

    def replay(self):
        """Replay on the inputs already in the static buffers."""
        self.graph.replay()
        return self.loss, self.grad_ent
