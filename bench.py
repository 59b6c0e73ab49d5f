#!/usr/bin/env python
"""bench.py -- triples scored / second on the BLP scoring-and-ranking hot path (BASELINE.json).

One STEP = one pass of the hot path over one batch of synthetic FB15k-237-shaped input:
  * one training step of LinkPrediction.compute_loss forward + backward (models.py:51-70)
    on B positives with K in-batch negatives each              -> B * (K + 1) triples scored
  * one evaluation chunk of E test triples ranked against ALL N entities, heads and tails
    (train.py:128-157)                                         -> 2 * E * N triples scored
`value` = triples scored per second with inputs resident in HBM (CUDA-event timed, max over ranks);
`e2e`   = the same metric through the public API with HOST (pinned) inputs, H2D/D2H inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): every rank runs the same per-GPU work on its own batch
(train replicas are independent exactly like the reference's DataParallel sub-batches; eval
queries are sharded, table replicated: SURVEY.md section 8e "small tables") -> weak scaling, no data-path
collective.  The entity-sharded sweep with its one all-reduce is exercised by `--sharded-sweep`.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    # name: (N entities, R relations, T test triples)
    "fb15k237": (14541, 237, 20480),
    "wn18rr": (40943, 11, 3136),
    "umls": (135, 46, 661),
}
METRIC = "triples scored/sec (train negs + eval full-entity rank)"
UNIT = "triples/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="b200", choices=("b200", "reference"))
    p.add_argument("--dataset", default="fb15k237", choices=sorted(CONFIGS))
    p.add_argument("--model", default="transe", choices=("transe", "distmult", "complex", "simple"))
    p.add_argument("--loss", default="margin", choices=("margin", "nll"))
    p.add_argument("--dim", type=int, default=128)
    p.add_argument("--train-batch", type=int, default=64)
    p.add_argument("--negatives", type=int, default=512)
    p.add_argument("--eval-batch", type=int, default=1024, help="test triples ranked per step (E)")
    p.add_argument("--ref-eval-batch", type=int, default=128, help="E of the bounded CPU sample")
    p.add_argument("--mode", default="exact", choices=("exact", "fast"),
                   help="eval sweep arithmetic: exact = reference fp32 order (bit-exact ranks); fast = tcgen05 split-FP16 "
                        "contraction (distmult / complex / simple only, tolerance-classified)")
    p.add_argument("--eager-train", action="store_true",
                   help="run compute_loss forward+backward eagerly instead of replaying blp_b200.GraphedLossStep")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the Wikidata5M-scale HBM-bound / entity-sharded sweep legs")
    p.add_argument("--wd-entities", type=int, default=4_800_000, help="rows of the Wikidata5M-scale table (whole table)")
    p.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    return p.parse_args()


# ------------------------------------------------------------------ workload ----
def make_workload(args):
    """Synthetic FB15k-237-shaped inputs, fixed seeds; CPU tensors."""
    n, r, t = CONFIGS[args.dataset]
    d, b, k = args.dim, args.train_batch, args.negatives
    g = torch.Generator().manual_seed(0)
    ent = torch.randn(n, d, generator=g)
    if args.model == "transe":
        ent = torch.nn.functional.normalize(ent, dim=-1)                  # models.py:40-41
    g = torch.Generator().manual_seed(1)
    a = (6.0 / (r + d)) ** 0.5
    rel = (torch.rand(r, d, generator=g) * 2 - 1) * a                      # xavier_uniform_, models.py:29
    g = torch.Generator().manual_seed(2)
    triples = torch.stack([torch.randint(0, n, (t,), generator=g), torch.randint(0, n, (t,), generator=g),
                           torch.randint(0, r, (t,), generator=g)], dim=1)
    g = torch.Generator().manual_seed(3)
    pairs = torch.randint(0, n, (b, 2), generator=g)
    rels = torch.randint(0, r, (b, 1), generator=g)
    # in-batch negatives with the structure and the memory layout of data.py:35-81: storage (K, B, 2),
    # returned transposed to (B, K, 2); one column keeps the row's own entity, the other is a random
    # in-batch entity of another row
    own = torch.arange(2 * b).reshape(b, 2)
    neg = own.repeat(k, 1).reshape(k, b, 2).clone()
    w = torch.ones(b, 2 * b)
    w.scatter_(1, own, torch.zeros(b, 2))
    repl = w.multinomial(k, replacement=True, generator=g).t()             # (K, B)
    col = torch.randint(0, 2, (k, b), generator=g)
    neg[torch.arange(k)[:, None], torch.arange(b)[None, :], col] = repl
    return {"ent": ent, "rel": rel, "triples": triples, "pairs": pairs, "rels": rels, "neg_storage": neg,
            "n": n, "r": r, "t": t, "d": d, "b": b, "k": k}


def step_triples(w, e):
    return w["b"] * (w["k"] + 1) + 2 * e * w["n"]


# ----------------------------------------------------------------- clocks ----
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [ln for (ts, ln) in self.lines if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.06)]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower() == "active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------- CPU reference arm ----
def cpu_reference_step(w, args, e, lo):
    """The reference's CPU execution strategy (stock ATen ops, materialised broadcasts) on one bounded
    sample: one compute_loss forward+backward and E test triples in eval batches of 64 (train.py:128-157)."""
    from oracle import torch_port
    ent_embs = w["ent"][w["pairs"]].clone().requires_grad_(True)           # (B,2,D)
    rel_w = w["rel"].clone().requires_grad_(True)
    neg_idx = w["neg_storage"].transpose(0, 1)                             # (B,K,2) non-contiguous view
    loss = torch_port.batch_loss(args.model, args.loss, ent_embs, rel_w[w["rels"][:, 0]], neg_idx, 0.0)
    loss.backward()
    ent_emb = w["ent"].unsqueeze(0)
    k_values = torch.tensor([[1, 3, 10]])
    mrr = 0.0
    for s in range(lo, lo + e, 64):
        tr = w["triples"][torch.arange(s, min(s + 64, lo + e)) % w["t"]]
        out = torch_port.eval_batch(args.model, ent_emb, tr[:, 0:1], tr[:, 1:2], w["rel"][tr[:, 2:3]], k_values)
        mrr += out["recip"].sum().item()
    return float(loss.item()), mrr


def run_cpu_reference(w, args, steps, warmup):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    e = args.ref_eval_batch
    for i in range(warmup):
        cpu_reference_step(w, args, e, (i * e) % w["t"])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_reference_step(w, args, e, ((warmup + i) * e) % w["t"])
    dt = time.perf_counter() - t0
    return {"value": steps * step_triples(w, e) / dt, "seconds": dt, "cores": torch.get_num_threads(),
            "ms_per_step": 1e3 * dt / steps,
            "sample": f"{steps} step(s) of: 1 compute_loss fwd+bwd (B={w['b']}, K={w['k']}) + {e} test triples ranked "
                      f"against all {w['n']} entities in eval batches of 64 (reference eval_batch_size), "
                      f"ATen-op restatement of the reference's CPU path (oracle/torch_port.py), {threads} torch threads"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = make_workload(args)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bound the whole run to a few minutes: one sample step costs ~2 s on 8 cores
    t0 = time.perf_counter()
    cpu_reference_step(w, args, args.ref_eval_batch, 0)
    one = time.perf_counter() - t0
    budget = 150.0
    if (steps + warmup) * one > budget:
        warmup = min(warmup, 1)
        steps = max(1, int(budget / one) - warmup)
    res = run_cpu_reference(w, args, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, w, args.ref_eval_batch, flush=False),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def config_dict(args, w, e, flush):
    return {"workload": f"synthetic {args.dataset} ({w['n']} entities, {w['r']} relations) BLP-{args.model} dim={w['d']}: "
                        f"1 compute_loss fwd+bwd (B={w['b']}, K={w['k']} negatives, {args.loss} loss) + {e} test triples "
                        f"ranked against all entities (heads and tails) per step",
            "entities": w["n"], "relations": w["r"], "dim": w["d"], "rel_model": args.model, "loss": args.loss,
            "train_batch": w["b"], "negatives": w["k"], "eval_triples_per_step": e, "eval_mode": args.mode,
            "test_triples": "20480 synthetic triples, processed in relation order (sorted once per evaluation)",
            "train_step": "eager autograd" if args.eager_train else "CUDA-graph replay of compute_loss forward+backward (blp_b200.GraphedLossStep)",
            "triples_per_step": step_triples(w, e),
            "l2": ("flushed between timed steps (256 MiB write)" if flush else "not flushed (table is L2-resident by design)"),
            "parallelism": f"replicas x{args.gpus}: train sub-batches independent, eval queries sharded, table replicated"}



# ------------------------------------------- Wikidata5M-scale sweep legs ----
def wd_sweep_leg(args, dev, world, rank, hbm_peak):
    """BASELINE configs[4]: synthetic Wikidata5M-scale (4.8 M entities) BLP-TransE eval sweep with the candidate
    axis sharded by rows over the ranks: every rank counts over its shard, ONE all-reduce of the
    int32 counters per sweep.  eval batch 2 is the reference's setting for this dataset and is HBM-bound (each batch
    streams the whole table for 4 queries); eval batch 64 is FP32-bound.  Timed with CUDA events, max over ranks;
    the per-batch launches of one sweep are replayed from a CUDA graph so the host is not the limiter."""
    import torch.distributed as dist

    import blp_b200
    n_total, d, n_rel, T = args.wd_entities, 128, 822, 64
    lo, hi = blp_b200.shard_bounds(n_total, world, rank)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    shard = torch.nn.functional.normalize(torch.randn(hi - lo, d, generator=g, device=dev), dim=-1)   # models.py:40-41
    gc = torch.Generator().manual_seed(7)
    rel = ((torch.rand(n_rel, d, generator=gc) * 2 - 1) * (6.0 / (n_rel + d)) ** 0.5).to(dev)
    rows = torch.stack([torch.randint(0, n_total, (T,), generator=gc), torch.randint(0, n_total, (T,), generator=gc),
                        torch.randint(0, n_rel, (T,), generator=gc)], dim=1).to(dev)
    group = dist.group.WORLD if world > 1 else None
    h_rows = blp_b200.gather_rows(shard, lo, rows[:, 0], group)      # true rows replicated once per sweep
    t_rows = blp_b200.gather_rows(shard, lo, rows[:, 1], group)
    res = {"entities": n_total, "rows_per_rank": hi - lo, "shard_bytes": (hi - lo) * d * 4, "test_triples": T,
           "collective": "one all-reduce of the (2, 2, T) int32 counters per sweep" if world > 1 else "none (1 rank)"}
    for eval_b in (2, 64):
        def sweep(collective=True):
            return blp_b200.rank_sweep("transe", shard, rel, rows, ent_offset=lo, chunk=eval_b, h_rows=h_rows, t_rows=t_rows,
                                       group=group if collective else None, sort_by_relation=False)
        for _ in range(2):
            out = sweep()
        torch.cuda.synchronize()
        # counting part of one sweep as a CUDA graph (T / eval_b batches), the collective stays outside
        graph, static = None, None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                sweep(collective=False)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static = sweep(collective=False)
        except Exception as exc:   # measurement aid only: fall back to eager launches
            graph, static = None, None
            res[f"graph_error_b{eval_b}"] = str(exc)[:120]
            torch.cuda.synchronize()

        def run():
            if graph is None:
                return sweep()
            graph.replay()
            if world > 1:
                cnt = torch.stack([static["gt"], static["ge"]])
                dist.all_reduce(cnt, group=group)
            return static
        reps = 3 if eval_b == 2 else 2
        run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(reps):
            run()
        b_.record()
        torch.cuda.synchronize()
        ms = a_.elapsed_time(b_) / reps
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        batches = T // eval_b
        alg = n_total * d * 4 + eval_b * 3 * d * 4 + 2 * eval_b * 12       # bytes per batch, summed over the ranks
        gbs = alg * batches / (ms * 1e-3) / 1e9
        res[f"eval_batch_{eval_b}"] = {"ms_per_sweep": ms, "ms_per_batch": ms / batches, "scores_per_s": 2 * T * n_total / (ms * 1e-3),
                                      "algorithmic_GBps_all_ranks": gbs, "hbm_frac": gbs / (world * hbm_peak),
                                      "graph": graph is not None}
        del graph, static
    del shard
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------- B200 arm ----
def main_b200(args):
    import torch.distributed as dist

    import blp_b200
    from blp_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    w = make_workload(args)
    e, b, k, d, n, t = args.eval_batch, w["b"], w["k"], w["d"], w["n"], w["t"]
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    flush = not args.no_flush

    # ---- resident inputs (value): everything already in HBM
    ent = w["ent"].to(dev)
    model = blp_b200.TransductiveLinkPrediction(d, args.model, args.loss, 8, w["r"], 0).to(dev)
    with torch.no_grad():
        model.rel_emb.weight.copy_(w["rel"])
    rel_w = model.rel_emb.weight
    # each rank works on its own slice of the test triples / its own training sub-batch
    # the evaluation set is put in relation order once (what rank_sweep(sort_by_relation=True) does per sweep): triples
    # that share a relation let the TransE kernel reuse fl(candidate + r) across head-prediction queries
    test_triples = w["triples"].roll(-rank * e, 0)
    test_triples = test_triples[torch.argsort(test_triples[:, 2], stable=True)].contiguous()
    triples = test_triples.to(dev)
    ent_embs = ent[w["pairs"].to(dev)].contiguous()
    rels = w["rels"].to(dev)
    neg = w["neg_storage"].to(dev).transpose(0, 1)                         # (B,K,2), strides of the reference sampler
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
    # pre-gathered query rows per eval chunk (train.py:141-143 gathers are part of the step: done inside)

    launches = {"n": 0}

    # the launch-bound training step replays a CUDA graph of `loss = model.compute_loss(...); loss.backward()`
    graphed = None if args.eager_train else blp_b200.GraphedLossStep(model, b, k)
    if graphed is not None:
        graphed(ent_embs, rels, neg)                  # resident inputs: loaded into the static buffers once

    def train_step():
        if graphed is not None:
            loss, _ = graphed.replay()
        else:
            x = ent_embs.detach().requires_grad_(True)
            rel_w.grad = None
            loss = model.compute_loss(x, rels, neg)
            loss.backward()
        launches["n"] += 1 + 1            # fused fwd+bwd kernel, one blp_scale in autograd's backward
        return loss

    # the E test triples of step i: consecutive slices of the (rolled) test set, wrapping around
    n_chunks = max(1, t // e)
    chunks = [triples[(torch.arange(c * e, (c + 1) * e, device=dev) % t)].contiguous() for c in range(n_chunks)]

    if args.mode == "fast" and args.model == "transe":
        raise SystemExit("--mode fast covers distmult / complex / simple (TransE is an L1 distance, not a contraction)")
    fast_ws = ops.fast_table(ent) if args.mode == "fast" else None      # split table: built once per entity table

    # pre-validated sweep for E triples per call (same kernels as blp_b200.rank_sweep, ~15 us of host time per call)
    plan = blp_b200.RankSweepPlan(args.model, ent, rel_w, e, mode=args.mode, fast_table=fast_ws)

    def eval_step(i):
        out = plan(chunks[i % n_chunks])
        launches["n"] += out["launches"]
        return out

    def step(i):
        loss = train_step()
        out = eval_step(i)
        return loss, out

    # profile events for the dominant kernel (the eval sweep kernel)
    prof = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a_, b_ in prof + step_ev:          # materialise the cudaEvent_t handles
        a_.record(); b_.record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    launches["n"] = 0
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.12)
    t_wall0 = time.perf_counter()
    host_s = 0.0
    for i in range(steps):
        if flush:
            flush_buf.fill_(i & 0xFF)
        h0 = time.perf_counter()
        step_ev[i][0].record()
        lib.blp_profile_events(1, ctypes.c_void_p(prof[i][0].cuda_event), ctypes.c_void_p(prof[i][1].cuda_event))
        step(warmup + i)
        step_ev[i][1].record()
        host_s += time.perf_counter() - h0
    lib.blp_profile_events(0, None, None)
    barrier()
    t_wall1 = time.perf_counter()
    timed_launches = launches["n"]
    step_ms = [a_.elapsed_time(b_) for a_, b_ in step_ev]
    kern_ms = [a_.elapsed_time(b_) for a_, b_ in prof]
    if os.environ.get("BLP_BENCH_DEBUG"):
        sys.stderr.write(f"[rank {rank}] step_ms min/med/max {min(step_ms):.3f}/{statistics.median(step_ms):.3f}/{max(step_ms):.3f} "
                         f"kern_ms med {statistics.median(kern_ms):.3f} host enqueue {1e3 * host_s / steps:.3f} ms/step "
                         f"cpus {len(os.sched_getaffinity(0))} dev {torch.cuda.current_device()}\n")
    total_ms = sum(step_ms)
    # keep the GPU under the same load until nvidia-smi has a few samples, if the timed region was short
    if t_wall1 - t_wall0 < 1.0:
        t_end = time.perf_counter() + 1.0
        i = 0
        while time.perf_counter() < t_end:
            step(i); i += 1
            if i % 16 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * steps * step_triples(w, e) / (total_ms * 1e-3)

    # ---- e2e: public API, host (pinned) inputs, H2D + D2H inside the timed region
    pin = lambda x: x.contiguous().pin_memory()  # noqa: E731
    h_ent_embs = pin(w["ent"][w["pairs"]])
    h_rels = pin(w["rels"])
    h_neg = pin(w["neg_storage"])
    h_triples = [pin(test_triples[torch.arange(i * e, (i + 1) * e) % t]) for i in range(max(1, t // e))]
    h2d = h_ent_embs.numel() * 4 + h_rels.numel() * 8 + h_neg.numel() * 8 + h_triples[0].numel() * 8
    d2h = 4 + 4 * 8

    # results are read on the host one step behind the GPU (two pinned result slots), like a training loop that logs
    # the previous step's loss while the next step is already queued: every step still does its H2D and its D2H
    host_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_sums = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)]
    landed = [torch.cuda.Event(), torch.cuda.Event()]

    # host -> device prefetch, the way a pinned-memory DataLoader feeds a training loop: the inputs of step i + 1 cross
    # PCIe on a copy stream into one of two device staging slots while step i computes; every step still moves its
    # own h2d bytes inside the timed region
    copy_stream = torch.cuda.Stream()
    stage = [{"ent": torch.empty_like(h_ent_embs, device=dev), "rels": torch.empty_like(h_rels, device=dev),
              "neg": torch.empty_like(h_neg, device=dev), "tr": torch.empty_like(h_triples[0], device=dev)} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_prefetch(i):
        slot = stage[i & 1]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])            # the step that last used this slot has read it
            slot["ent"].copy_(h_ent_embs, non_blocking=True)
            slot["rels"].copy_(h_rels, non_blocking=True)
            slot["neg"].copy_(h_neg, non_blocking=True)
            slot["tr"].copy_(h_triples[i % len(h_triples)], non_blocking=True)
            ready[i & 1].record(copy_stream)

    def e2e_issue(i):
        main = torch.cuda.current_stream()
        if i == e2e_first[0]:
            e2e_prefetch(i)
        e2e_prefetch(i + 1)
        main.wait_event(ready[i & 1])
        slot = stage[i & 1]
        if graphed is not None:
            loss, _ = graphed(slot["ent"], slot["rels"], slot["neg"])   # into the graph's static buffers, one graph launch
        else:
            x = slot["ent"].clone().requires_grad_(True)
            rel_w.grad = None
            loss = model.compute_loss(x, slot["rels"], slot["neg"].transpose(0, 1))
            loss.backward()
        out = plan(slot["tr"])
        consumed[i & 1].record(main)
        # D2H: the loss scalar (train.py:352) + the 4 fp64 metric accumulators (train.py:154-157)
        host_loss[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        host_sums[i & 1].copy_(out["sums"], non_blocking=True)
        landed[i & 1].record()

    def e2e_read(i):
        landed[i & 1].synchronize()
        return float(host_loss[i & 1][0]), {"mrr": float(host_sums[i & 1][0]) / (2 * e)}

    e2e_first = [0]

    def e2e_step(i):
        e2e_first[0] = i
        e2e_issue(i)
        return e2e_read(i)

    for i in range(warmup):
        e2e_step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = steps
    ev0.record()
    e2e_first[0] = 0
    for i in range(e2e_steps):
        e2e_issue(i)
        if i > 0:
            last = e2e_read(i - 1)
    last = e2e_read(e2e_steps - 1)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = world * e2e_steps * step_triples(w, e) / (e2e_ms * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- Wikidata5M-scale sweep: HBM-bound at the reference's eval batch of 2; entity-sharded over the ranks
    wd = None
    if not args.no_extra:
        del flush_buf
        torch.cuda.empty_cache()
        wd = wd_sweep_leg(args, dev, world, rank, hbm_peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (eval sweep kernel), measured live above
    kern_s = statistics.mean(kern_ms) * 1e-3
    alg_bytes = n * d * 4 + e * 3 * d * 4 + 2 * e * 4 + 2 * e * 2 * 4      # table once + query rows + s_true + counters
    kname = f"{'fast_sweep_kernel' if args.mode == 'fast' else 'sweep_kernel'}<{args.model}>"
    traffic = ncu_traffic(f"{kname}|{args.dataset}|E{e}|{args.mode}")
    share = statistics.mean(kern_ms) / statistics.mean(step_ms)
    if args.mode == "fast":
        # tensor-core sweep: S = C (2E x 128) . E^T (128 x N); the split-FP16 form issues three MMAs per algorithmic product
        flops = 2.0 * (2 * e) * n * d
        f16_peak = float(peaks.get("bf16_tflops", 2250.0))
        roofline = {
            "bound": "tensor", "kernel": kname, "achieved": flops / kern_s / 1e12, "peak": f16_peak, "unit": "TFLOP/s",
            "frac": flops / kern_s / 1e12 / f16_peak, "traffic": traffic,
            "peak_source": ("MEASURED_PEAKS.json bf16_tflops (kind::f16 dense rate, cuBLAS 8192^3 burst)" if peaks
                            else "fallback 2250 TFLOP/s"),
            "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_s * 1e3,
            "kernel_share_of_step": share,
            "binding": "tcgen05 kind::f16 on split-FP16 operands (hi*hi + lo*hi + hi*lo, 22 significant bits): three MMAs "
                       "are issued per algorithmic product, so frac <= 1/3 by construction; issued_frac is the tensor pipe's view",
            "issued_tflops": 3 * flops / kern_s / 1e12, "issued_frac": 3 * flops / kern_s / 1e12 / f16_peak,
        }
    else:
        lane_ops_per = {"transe": 2.5, "distmult": 2.5, "complex": 5.0, "simple": 2.5}[args.model] * d  # head/tail mean
        probe = measure_fp32_rate(ops, dev, args.model)
        roofline = {
            "bound": "hbm", "kernel": kname, "achieved": alg_bytes / kern_s / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "frac": alg_bytes / kern_s / 1e9 / hbm_peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s",
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_s * 1e3,
            "kernel_share_of_step": share,
            "binding": "fp32 ALU (exact-order arithmetic): with Q >= ~11 queries per table pass the sweep is bound by "
                       "FP32 lane-ops, not HBM (DESIGN.md); the HBM fraction is reported because the contract asks for it; "
                       "the HBM-bound shape of the same kernel (Wikidata5M-scale table, eval batch 2) is measured in "
                       "wikidata5m_scale_sweep.eval_batch_2.hbm_frac",
            "alu": {"lane_ops_per_launch": 2 * e * n * lane_ops_per, "achieved_tlaneops": 2 * e * n * lane_ops_per / kern_s / 1e12,
                    "peak_tlaneops": probe, "frac": (2 * e * n * lane_ops_per / kern_s / 1e12) / probe if probe else None,
                    "peak_source": "blp_pipe_probe FADD / FADD2 issue rate, measured in this run"},
        }
    if wd is not None:
        wd["traffic_eval_batch_2_1gpu"] = ncu_traffic("sweep_kernel<transe>|wikidata5m|E2|exact")

    cpu = None
    if not args.no_cpu_baseline:
        r = run_cpu_reference(w, args, steps=max(1, int(12.0 / max(0.5, 2.0))), warmup=1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == "exact" else "f16x3 split of f32 (fp32 accumulate; train step f32)", "data": "synthetic",
        "config": config_dict(args, w, e, flush),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / e2e_steps, "last_loss": last[0], "last_mrr": last[1]["mrr"]},
        "gpu_launches": timed_launches,
        "roofline": roofline, "cpu_baseline": cpu, "wikidata5m_scale_sweep": wd,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic(key):
    """dram bytes per launch of a kernel from the committed ncu --set full capture (profiles/rNN_traffic.json), or None."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            v = json.load(open(path)).get(key)
        except (OSError, ValueError):
            continue
        if v is not None:
            return v
    return None


def measure_fp32_rate(ops, dev, model):
    """FP32 pipe peak in T lane-ops/s, measured in this run: best of the FADD and FADD2 issue-rate probes
    (blp_pipe_probe variants 0 and 2; both saturate the same 128 lanes/clk/SM)."""
    best = 0.0
    try:
        for variant in (0, 2):
            for _ in range(3):
                sink = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)   # noqa: F841 (warm allocator)
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                lane_ops, _ = ops.pipe_probe(variant, dev, n_threads=148 * 8 * 256, iters=8192)
                b_.record()
                torch.cuda.synchronize()
                best = max(best, lane_ops / (a_.elapsed_time(b_) * 1e-3) / 1e12)
    except Exception:   # measurement aid only
        return None
    return best


if __name__ == "__main__":
    a = parse()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
