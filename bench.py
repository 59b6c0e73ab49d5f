#!/usr/bin/env python
"""bench.py -- triples scored / second on the BLP scoring-and-ranking hot path (BASELINE.json).

One STEP = one pass of the hot path over one evaluation set of synthetic FB15k-237-shaped input:
  * C = T / 1024 = 20 training steps of LinkPrediction.compute_loss forward + backward (models.py:51-70)
    on B positives with K in-batch negatives each              -> C * B * (K + 1) triples scored
  * the T = 20,480 test triples ranked against ALL N entities, heads and tails (train.py:128-157): exact mode = ONE
    fused launch for the whole sweep (true scores, sweep, counters), triples in relation-aligned order
                                                                -> 2 * T * N triples scored
(the reference arm samples 1 / C of it: one training step + 1,024 test triples in eval batches of 64)
`value` = triples scored per second with inputs resident in HBM (CUDA-event timed, max over ranks);
`e2e`   = the same metric through the public API with HOST (pinned) inputs, H2D/D2H inside the timed region.
`roofline` = the dominant kernel (the eval sweep) against the bound that BINDS it (FP32 pipe for the exact-order FB sweep);
the HBM-bound shape of the same kernel (Wikidata5M-scale table, eval batch 2), the tensor-core DistMult / ComplEx sweeps
and the other BASELINE configs are measured as legs and reported as flat keys of `roofline` plus the `legs` object.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): every rank runs the same per-GPU work on its own batch (train replicas are
independent exactly like the reference's DataParallel sub-batches; FB-size eval queries are sharded, table replicated:
SURVEY.md section 8e "small tables") -> weak scaling, no data-path collective.  The path north_star shards -- the
4.8 M-entity Wikidata5M-scale table split by rows over the ranks, ONE all-reduce of the int32 counters per sweep -- runs
as the `wikidata5m_scale_sweep` leg at every N, and its counters are checked bit-for-bit against a single-rank pass.

--impl reference: the reference's own CPU implementation of the same step (its compute_loss / score_fn / get_metrics,
byte-compiled by oracle/build_ref.py; oracle/torch_port.py when that is absent) on all host cores, each step a bounded
sample of the workload with the same train : eval mix.
"""
import argparse
import ctypes
import hashlib
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
WD_BLOCKS = 8                      # the Wikidata5M-scale table is generated in 8 fixed row blocks (seeds 100..107)
CHECKSUMS = os.path.join(ROOT, "tests", "golden", "wd_sweep_checksum.json")


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=("b200", "reference"))
    p.add_argument("--dataset", default="fb15k237", choices=sorted(CONFIGS))
    p.add_argument("--model", default="transe", choices=("transe", "distmult", "complex", "simple"))
    p.add_argument("--loss", default="margin", choices=("margin", "nll"))
    p.add_argument("--dim", type=int, default=128)
    p.add_argument("--train-batch", type=int, default=64)
    p.add_argument("--negatives", type=int, default=512)
    p.add_argument("--eval-batch", type=int, default=1024, help="test triples ranked per sub-step (E)")
    p.add_argument("--mode", default="exact", choices=("exact", "fast"),
                   help="eval sweep arithmetic: exact = reference fp32 order (bit-exact ranks); fast = tcgen05 split-FP16 "
                        "contraction (distmult / complex / simple only, tolerance-classified)")
    p.add_argument("--eager-train", action="store_true",
                   help="run compute_loss forward+backward eagerly instead of replaying blp_b200.GraphedLossStep")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip every leg (Wikidata5M-scale sweep, other BASELINE configs)")
    p.add_argument("--no-legs", action="store_true", help="skip the single-GPU legs (other BASELINE configs), keep the sharded sweep")
    p.add_argument("--wd-entities", type=int, default=4_800_000, help="rows of the Wikidata5M-scale table (whole table)")
    p.add_argument("--wd-triples", type=int, default=8192, help="test triples of the Wikidata5M-scale sweep (SURVEY 8d)")
    p.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    p.add_argument("--write-checksum", action="store_true", help="record the Wikidata5M-scale sweep checksum (N = 1)")
    return p.parse_args()


# ------------------------------------------------------------------ workload ----
def make_workload(args, dataset=None, model=None):
    """Synthetic inputs of one BASELINE config, fixed seeds; CPU tensors."""
    dataset, model = dataset or args.dataset, model or args.model
    n, r, t = CONFIGS[dataset]
    d, b, k = args.dim, args.train_batch, args.negatives
    g = torch.Generator().manual_seed(0)
    ent = torch.randn(n, d, generator=g)
    if model == "transe":
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
            "n": n, "r": r, "t": t, "d": d, "b": b, "k": k, "dataset": dataset, "model": model}


def substeps(w, e):
    return max(1, w["t"] // e)


def step_triples(w, e):
    return substeps(w, e) * (w["b"] * (w["k"] + 1) + 2 * e * w["n"])


def config_dict(args, w, e, world):
    """Identical in both arms (the reference arm processes a bounded SAMPLE of this step, see cpu_baseline.sample)."""
    c = substeps(w, e)
    return {"workload": f"synthetic {w['dataset']} ({w['n']} entities, {w['r']} relations) BLP-{w['model']} dim={w['d']}: {c} compute_loss "
                        f"fwd+bwd (B={w['b']}, K={w['k']} negatives, {args.loss} loss) + one pass over the {c * e} test triples, each "
                        f"ranked against all entities (heads and tails)",
            "entities": w["n"], "relations": w["r"], "dim": w["d"], "rel_model": w["model"], "loss": args.loss,
            "train_batch": w["b"], "negatives": w["k"], "train_steps_per_step": c, "eval_triples_per_step": c * e,
            "eval_mode": args.mode,
            "test_triples": f"{w['t']} synthetic triples in relation-aligned order (sorted and padded once per evaluation set)",
            "triples_per_step": step_triples(w, e),
            "l2": "flushed between timed steps (256 MiB write); inside a step the 7.4 MB table is L2-resident by design",
            "parallelism": "replicas: train sub-batches independent, eval queries sharded over the GPUs, table replicated "
                           "(weak scaling, no data-path collective); the entity-sharded sweep is the wikidata5m_scale_sweep leg"}


# ----------------------------------------------------------------- clocks ----
class ClockSampler:
    """SM clock + throttle reasons sampled INSIDE the timed region: NVML polled from a thread every ~1 ms (the timed
    region is tens to hundreds of milliseconds); nvidia-smi -lms 20 as a fallback (B200_PROFILING.md recipe)."""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.proc = index, [], False, None, None
        self.kind = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.kind = "nvml"
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except (ValueError, IndexError):
                return index
        return index

    def _poll_nvml(self):
        nv = self.nv
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.handle))
                self.samples.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.001)

    def _poll_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                mask = sum(bit for bit, v in zip((0x8, 0x40, 0x20, 0x4), f[2:6]) if v.lower() == "active")
                self.samples.append((time.perf_counter(), float(f[0]), mask))
                self.max_mhz = float(f[1])
            except (ValueError, IndexError):
                continue

    def start(self):
        if self.kind == "nvml":
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.max_mhz = None
            self.kind = "nvidia-smi"
            self.thread = threading.Thread(target=self._poll_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.kind = None

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.kind is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        inside = [(m, k) for (ts, m, k) in self.samples if t0 <= ts <= t1]
        rows = inside if inside else [(m, k) for (_, m, k) in self.samples]
        mask = 0
        for _, k in rows:
            mask |= k
        return {"sm_mhz": statistics.median(m for m, _ in rows) if rows else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(nm for bit, nm in self.NAMES.items() if mask & bit), "samples": len(rows),
                "samples_inside_timed_region": len(inside), "timed_region_s": t1 - t0,
                "sampler": f"{self.kind}, polled inside the timed region"}


# ------------------------------------------------------- CPU reference arm ----
class CpuArm:
    """The reference's CPU path for one bounded sample of a step: 1 compute_loss forward+backward + E test triples in
    eval batches of 64 (train.py:128-157, eval_batch_size of every small-graph script) = 1 / C of the GPU step with
    the same train : eval mix.  kind 'reference' = the reference's own models.py / utils.py (byte-compiled by
    oracle/build_ref.py); 'port' = oracle/torch_port.py, its ATen-op restatement (asserted equal to the reference by
    tests/test_oracle_golden.py where the reference is mounted)."""

    def __init__(self, w, args):
        self.w, self.args = w, args
        self.kind, self.model = "port", None
        try:
            from oracle import ref_loader
            if ref_loader.available():
                mods = ref_loader.load(("models", "utils"))
                self.ref_models, self.ref_utils = mods["models"], mods["utils"]
                m = self.ref_models.TransductiveLinkPrediction(w["d"], w["model"], args.loss, 8, w["r"], 0)
                with torch.no_grad():
                    m.rel_emb.weight.copy_(w["rel"])
                self.model, self.kind = m, "reference"
        except Exception as exc:       # fall back to the port, say why
            self.why = str(exc)[:120]
        if self.model is None:
            from oracle import torch_port
            self.port = torch_port

    def describe(self, e, threads, steps):
        what = ("the reference's own compute_loss / score_fn / get_metrics (models.py:51-70, 222-258, utils.py:86-111, "
                "byte-compiled oracle/_ref)" if self.kind == "reference"
                else "ATen-op restatement of the reference's CPU path (oracle/torch_port.py)")
        c = substeps(self.w, e)
        return (f"{steps} step(s), each a 1/{c} sample of the GPU step with the same mix: 1 compute_loss fwd+bwd (B={self.w['b']}, "
                f"K={self.w['k']}) + {e} test triples ranked against all {self.w['n']} entities in eval batches of 64; "
                f"{what}, {threads} torch threads")

    def step(self, e, lo):
        w, args = self.w, self.args
        ent_embs = w["ent"][w["pairs"]].clone().requires_grad_(True)           # (B,2,D)
        neg_idx = w["neg_storage"].transpose(0, 1)                             # (B,K,2) non-contiguous view
        k_values = torch.tensor([[1, 3, 10]])
        ent_emb = w["ent"].unsqueeze(0)
        mrr = 0.0
        if self.kind == "reference":
            m = self.model
            m.rel_emb.weight.grad = None
            loss = m.compute_loss(ent_embs, w["rels"], neg_idx)                # models.py:51-70
            loss.backward()
            with torch.no_grad():
                for s in range(lo, lo + e, 64):
                    tr = w["triples"][torch.arange(s, min(s + 64, lo + e)) % w["t"]]
                    heads, tails, rels = torch.chunk(tr, chunks=3, dim=1)
                    head_embs = ent_emb.squeeze()[heads]                       # train.py:141-153
                    tail_embs = ent_emb.squeeze()[tails]
                    rel_embs = m.rel_emb(rels)
                    heads_predictions = m.score_fn(ent_emb, tail_embs, rel_embs)
                    tails_predictions = m.score_fn(head_embs, ent_emb, rel_embs)
                    pred_ents = torch.cat((heads_predictions, tails_predictions))
                    true_ents = torch.cat((heads, tails))
                    reciprocals, hits = self.ref_utils.get_metrics(pred_ents, true_ents, k_values)
                    mrr += reciprocals.sum().item()
        else:
            rel_w = w["rel"].clone().requires_grad_(True)
            loss = self.port.batch_loss(w["model"], args.loss, ent_embs, rel_w[w["rels"][:, 0]], neg_idx, 0.0)
            loss.backward()
            for s in range(lo, lo + e, 64):
                tr = w["triples"][torch.arange(s, min(s + 64, lo + e)) % w["t"]]
                out = self.port.eval_batch(w["model"], ent_emb, tr[:, 0:1], tr[:, 1:2], w["rel"][tr[:, 2:3]], k_values)
                mrr += out["recip"].sum().item()
        return float(loss.item()), mrr

    def run(self, e, steps, warmup):
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        w = self.w
        for i in range(warmup):
            self.step(e, (i * e) % w["t"])
        t0 = time.perf_counter()
        for i in range(steps):
            self.step(e, ((warmup + i) * e) % w["t"])
        dt = time.perf_counter() - t0
        per = w["b"] * (w["k"] + 1) + 2 * e * w["n"]
        return {"value": steps * per / dt, "seconds": dt, "cores": torch.get_num_threads(), "ms_per_step": 1e3 * dt / steps,
                "sample": self.describe(e, threads, steps), "kind": self.kind}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = make_workload(args)
    arm = CpuArm(w, args)
    e = args.eval_batch
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bound the whole run to a few minutes: one sample step costs ~4 s on 16 cores
    t0 = time.perf_counter()
    arm.step(e, 0)
    one = time.perf_counter() - t0
    budget = 200.0
    if (steps + warmup) * one > budget:
        warmup = min(warmup, 1)
        steps = max(1, min(steps, int(budget / one) - warmup))
    res = arm.run(e, steps, warmup)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, w, e, world),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------- helpers ----
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


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


def measure_fp32_rate(ops, dev):
    """FP32 pipe peak in T lane-ops/s, measured in this run: best of the FADD and FADD2 issue-rate probes
    (blp_pipe_probe variants 0 and 2; both saturate the same 128 lanes/clk/SM)."""
    best = 0.0
    try:
        for variant in (0, 2):
            for _ in range(3):
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                lane_ops, _ = ops.pipe_probe(variant, dev, n_threads=148 * 8 * 256, iters=8192)
                b_.record()
                torch.cuda.synchronize()
                best = max(best, lane_ops / (a_.elapsed_time(b_) * 1e-3) / 1e12)
    except Exception:   # measurement aid only
        return None
    return best or None


def time_calls(fn, reps, warm=2):
    """ms per call, CUDA events around `reps` back-to-back calls."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(reps):
        fn()
    b_.record()
    torch.cuda.synchronize()
    return a_.elapsed_time(b_) / reps


# lane-ops per (query, candidate, dim): SURVEY.md section 8(d) algorithmic count / what the exact-order kernel must execute
# algorithmic = the query side hoisted as far as the reference's association allows for TAIL prediction, applied to both roles
# (TransE sub + |.|-accumulate; DistMult mul + add; ComplEx 4 mul + 4 add per complex dim; SimplE 3 mul + 2 add per half dim);
# executed = mean of head and tail prediction (head prediction cannot pre-fold (candidate op relation) bit-exactly)
ALG_OPS = {"transe": 2.0, "distmult": 2.0, "complex": 4.0, "simple": 2.5}
EXEC_OPS = {"transe": 2.5, "distmult": 2.5, "complex": 5.0, "simple": 2.5}
# relation-aligned order: the first rounding of head prediction once per 4 head queries, e.g. TransE ((1 + 4 * 2) / 4 + 2) / 2
EXEC_OPS_ALIGNED = {"transe": 2.125, "distmult": 2.125, "complex": 4.25, "simple": 2.3125}


def sweep_roofline(model, mode, n, d, q_per_launch, kern_s, peaks, fp32_peak, hbm_peak, extra=None, exec_ops=None):
    """Roofline block of one sweep launch ranking q_per_launch queries against n candidates, bound = the binding one."""
    alg_bytes = n * d * 4 + (q_per_launch // 2) * 3 * d * 4 + q_per_launch * 12
    out = {"kernel_ms": kern_s * 1e3, "algorithmic_bytes_per_launch": alg_bytes,
           "hbm_GBps_algorithmic": alg_bytes / kern_s / 1e9, "hbm_frac": alg_bytes / kern_s / 1e9 / hbm_peak}
    if mode.startswith("fast"):
        flops = 2.0 * q_per_launch * n * d
        f16_peak = float(peaks.get("bf16_tflops", 2250.0))
        out.update({"bound": "tensor", "achieved": flops / kern_s / 1e12, "peak": f16_peak, "unit": "TFLOP/s",
                    "frac": flops / kern_s / 1e12 / f16_peak, "issued_frac": 3 * flops / kern_s / 1e12 / f16_peak,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (kind::f16 dense, cuBLAS burst)" if peaks else "fallback 2250 TFLOP/s",
                    "note": "split-FP16 issues 3 MMAs per algorithmic product: frac <= 1/3 by construction, issued_frac is the tensor pipe's view"})
    else:
        ops_alg = q_per_launch * n * d * ALG_OPS[model]
        peak = fp32_peak or 148 * 128 * 1.965e9 / 1e12
        out.update({"bound": "fp32_alu", "achieved": ops_alg / kern_s / 1e12, "peak": peak, "unit": "T lane-op/s",
                    "frac": ops_alg / kern_s / 1e12 / peak,
                    "frac_executed_ops": q_per_launch * n * d * (exec_ops or EXEC_OPS[model]) / kern_s / 1e12 / peak,
                    "peak_source": "blp_pipe_probe FADD / FADD2 issue rate measured in this run" if fp32_peak
                                   else "nominal 148 SM x 128 lanes x 1.965 GHz",
                    "note": f"exact-order fp32: {ALG_OPS[model]:g} lane-ops per (query, candidate, dim) algorithmic (SURVEY 8d), "
                            f"{(exec_ops or EXEC_OPS[model]):g} executed (head prediction cannot pre-fold the relation bit-exactly)"})
    if extra:
        out.update(extra)
    return out


# ------------------------------------------- Wikidata5M-scale sweep legs ----
def wd_blocks(n_total):
    per = -(-n_total // WD_BLOCKS)
    return [(b * per, min(n_total, (b + 1) * per)) for b in range(WD_BLOCKS)]


def wd_make_rows(lo, hi, n_total, d, dev):
    """Rows [lo, hi) of the synthetic Wikidata5M-scale table: fixed blocks with fixed seeds, so the table is the same for
    every world size (L2-normalised like models.py:40-41)."""
    parts = []
    for bi, (blo, bhi) in enumerate(wd_blocks(n_total)):
        if bhi <= lo or blo >= hi:
            continue
        g = torch.Generator(device=dev).manual_seed(100 + bi)
        blk = torch.nn.functional.normalize(torch.randn(bhi - blo, d, generator=g, device=dev), dim=-1)
        parts.append(blk[max(lo, blo) - blo:min(hi, bhi) - blo])
    return torch.cat(parts) if len(parts) != 1 else parts[0].contiguous()


def wd_sweep_leg(args, dev, world, rank, hbm_peak):
    """BASELINE configs[4]: synthetic Wikidata5M-scale (4.8 M entities) BLP-TransE eval sweep with the candidate axis
    sharded by rows over the ranks: every rank counts over its shard, ONE all-reduce of the int32
    counters per sweep.  Table pass size 2 is the reference's eval_batch_size for this dataset (scripts/*wikidata5m*) and
    is HBM-bound (each pass streams the whole table for 4 queries); 32 is FP32-bound.  The whole sweep (all T triples) is
    ONE kernel launch per rank.  Timed with CUDA events including the collective, max over ranks."""
    import torch.distributed as dist

    import blp_b200
    n_total, d, n_rel, T = args.wd_entities, 128, 822, args.wd_triples
    lo, hi = blp_b200.shard_bounds(n_total, world, rank)
    shard = wd_make_rows(lo, hi, n_total, d, dev)
    gc = torch.Generator().manual_seed(7)
    rel = ((torch.rand(n_rel, d, generator=gc) * 2 - 1) * (6.0 / (n_rel + d)) ** 0.5).to(dev)
    rows = torch.stack([torch.randint(0, n_total, (T,), generator=gc), torch.randint(0, n_total, (T,), generator=gc),
                        torch.randint(0, n_rel, (T,), generator=gc)], dim=1).to(dev)
    group = dist.group.WORLD if world > 1 else None
    h_rows = blp_b200.gather_rows(shard, lo, rows[:, 0], group)      # true rows replicated once per sweep
    t_rows = blp_b200.gather_rows(shard, lo, rows[:, 1], group)
    res = {"entities": n_total, "rows_per_rank": hi - lo, "shard_bytes": (hi - lo) * d * 4, "test_triples": T,
           "launches_per_sweep_per_rank": 1,
           "collective": "one all-reduce of the (2, 2, T) int32 counters per sweep" if world > 1 else "none (1 rank)"}
    counters = {}
    for pass_size, reps in ((2, 1), (32, 1)):
        def sweep():
            return blp_b200.rank_sweep("transe", shard, rel, rows, ent_offset=lo, group_triples=pass_size, h_rows=h_rows,
                                       t_rows=t_rows, group=group, sort_by_relation=False)
        out = sweep()                                                 # warm-up (also the result that is checked)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(reps):
            out = sweep()
        b_.record()
        torch.cuda.synchronize()
        ms = a_.elapsed_time(b_) / reps
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        passes = -(-T // pass_size)
        alg = n_total * d * 4 + pass_size * 3 * d * 4 + 2 * pass_size * 12      # bytes per table pass, summed over the ranks
        gbs = alg * passes / (ms * 1e-3) / 1e9
        res[f"table_pass_{pass_size}"] = {"ms_per_sweep": ms, "ms_per_pass": ms / passes, "scores_per_s": 2 * T * n_total / (ms * 1e-3),
                                          "algorithmic_GBps_all_ranks": gbs, "hbm_frac": gbs / (world * hbm_peak),
                                          "mrr": float(out["sums"][0]) / (2 * T)}
        if pass_size == 2:
            res["table_pass_2"]["peak_note"] = ("hbm_frac is against MEASURED_PEAKS.json hbm_gbs, a read + write COPY figure; this kernel "
                                                "only reads (and shards that fit partly in the 126 MB L2 get hits), so > 1 is possible")
        counters[pass_size] = torch.stack([out["gt"], out["ge"]]).cpu().numpy()
    # ---- parity of the sharded sweep (integer counters: must be BIT-equal for every world size and pass size)
    parity = {"pass_sizes_agree": bool(np.array_equal(counters[2], counters[32]))}
    digest = hashlib.sha256(np.ascontiguousarray(counters[2]).tobytes()).hexdigest()
    key = f"transe|n={n_total}|T={T}|blocks={WD_BLOCKS}"
    parity["checksum"] = digest
    try:
        want = json.load(open(CHECKSUMS)).get(key)
    except (OSError, ValueError):
        want = None
    parity["checksum_vs_committed_single_gpu_run"] = (None if want is None or WD_BLOCKS % world else
                                                      "bit-equal" if want == digest else "MISMATCH")
    if args.write_checksum and world == 1 and rank == 0:
        try:
            cur = json.load(open(CHECKSUMS))
        except (OSError, ValueError):
            cur = {}
        cur[key] = digest
        json.dump(cur, open(CHECKSUMS, "w"), indent=1, sort_keys=True)
    if world > 1:
        # direct check: rank 0 holds the WHOLE table and ranks the same triples alone, no collective
        ok = torch.ones(1, device=dev)
        del shard
        torch.cuda.empty_cache()
        if rank == 0:
            full = wd_make_rows(0, n_total, n_total, d, dev)
            solo = blp_b200.rank_sweep("transe", full, rel, rows, group_triples=32, sort_by_relation=False)
            same = np.array_equal(torch.stack([solo["gt"], solo["ge"]]).cpu().numpy(), counters[2])
            ok.fill_(1.0 if same else 0.0)
            del full, solo
        dist.broadcast(ok, 0)
        parity["vs_single_rank_pass_in_this_run"] = "bit-equal" if bool(ok.item()) else "MISMATCH"
    else:
        parity["vs_single_rank_pass_in_this_run"] = "n/a (1 rank)"
    verdicts = [v for v in (parity["checksum_vs_committed_single_gpu_run"], parity["vs_single_rank_pass_in_this_run"])
                if v in ("bit-equal", "MISMATCH")]
    if not parity["pass_sizes_agree"] or "MISMATCH" in verdicts:
        parity["sharded_parity"] = "MISMATCH"
    elif verdicts:
        parity["sharded_parity"] = "bit-equal"
    else:
        parity["sharded_parity"] = "unchecked (no committed checksum for this shape)"
    res["parity"] = parity
    torch.cuda.empty_cache()
    return res


# ------------------------------------------- single-GPU legs: the other BASELINE configs ----
def config_legs(args, dev, peaks, fp32_peak, hbm_peak):
    """BASELINE configs[2] (FB15k-237 DistMult, exact + tensor-core), configs[3] (WN18RR ComplEx full-entity sweep, exact +
    tensor-core), the reference's own eval batch (64 triples per call), BOW widths (D = 300 / 768) and the Wikidata5M
    training batch (B = 1024): each timed alone with CUDA events, each with its own roofline block."""
    import blp_b200
    from blp_b200 import ops
    legs = {}
    sums_of = {}

    def sweep_leg(name, dataset, model, mode, e, overlap=False):
        w = make_workload(args, dataset, model)
        ent, rel = w["ent"].to(dev), w["rel"].to(dev)
        tr = w["triples"][torch.argsort(w["triples"][:, 2], stable=True)].contiguous()
        e = min(e, w["t"])
        state = {"i": 0, "launches": 0}
        exec_ops = None
        if mode == "exact" and e == w["t"]:
            # the whole evaluation set in one launch, relation-aligned order prepared once (like the headline step)
            aligned = blp_b200.AlignedTriples(tr.to(dev))
            exec_ops = EXEC_OPS_ALIGNED.get(model)

            def call():
                out = blp_b200.rank_sweep(model, ent, rel, aligned, sort_by_relation=False)
                state["launches"] = out["launches"]
                return out
        else:
            chunks = [tr[torch.arange(c * e, (c + 1) * e) % w["t"]].contiguous().to(dev) for c in range(max(1, w["t"] // e))]
            ft = ops.fast_table(ent) if mode.startswith("fast") else None
            plan = blp_b200.RankSweepPlan(model, ent, rel, e, mode=mode, fast_table=ft, overlap_calls=overlap)

            def call():
                out = plan(chunks[state["i"] % len(chunks)])
                state["i"] += 1
                state["launches"] = out["launches"]
                return out
        ms = time_calls(call, reps=max(3, min(50, int(200 / max(1, e // 64)))), warm=2)
        # the sweep kernel alone (CUDA events around its launch, blp_profile_events slot 1), median of 7 calls
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
        for a_, b_ in evs:
            a_.record(); b_.record()
        torch.cuda.synchronize()
        lib_ = ops.lib()
        for a_, b_ in evs:
            lib_.blp_profile_events(1, ctypes.c_void_p(a_.cuda_event), ctypes.c_void_p(b_.cuda_event))
            call()
        lib_.blp_profile_events(0, None, None)
        torch.cuda.synchronize()
        kern_ms = sorted(a_.elapsed_time(b_) for a_, b_ in evs)[len(evs) // 2]
        out = call()
        torch.cuda.synchronize()
        leg = {"config": f"synthetic {dataset} ({w['n']} entities) BLP-{model} dim={w['d']}, {mode} mode, {e} test triples per call, "
                         f"ranked against all entities (heads and tails)",
               "ms_per_call": ms, "launches_per_call": state["launches"], "scores_per_s": 2 * e * w["n"] / (ms * 1e-3),
               "sweep_s_all_test_triples": ms * 1e-3 * w["t"] / e, "mrr": float(out["sums"][0]) / (2 * e)}
        leg["roofline"] = sweep_roofline(model, mode, w["n"], w["d"], 2 * e, ms * 1e-3, peaks, fp32_peak, hbm_peak,
                                         {"timed": "whole call (all launches of the call), back to back"}, exec_ops=exec_ops)
        if kern_ms > 0:
            # the dominant kernel of the call on its own (fold / refine / metrics launches excluded)
            kr = sweep_roofline(model, mode, w["n"], w["d"], 2 * e, kern_ms * 1e-3, peaks, fp32_peak, hbm_peak, exec_ops=exec_ops)
            leg["roofline"]["sweep_kernel_ms"] = kern_ms
            leg["roofline"]["sweep_kernel_frac"] = kr["frac"]
            if "issued_frac" in kr:
                leg["roofline"]["sweep_kernel_issued_frac"] = kr["issued_frac"]
        sums_of[name] = out["sums"].detach().cpu()
        if mode == "fast_exact":
            st = out["refine_state"].cpu()
            leg["refine_band_candidates_per_query"] = int(st[0]) / (2 * e)
            leg["refine_overflow"] = bool(int(st[1]))
            leg["note"] = ("tensor-core sweep + exact re-scoring of the candidates inside the a-priori error band around the true "
                           "score: integer ranks bit-identical to the exact mode (DESIGN 4.2b)")
        if overlap:
            leg["note"] = ("consecutive calls launched with programmatic stream serialization (RankSweepPlan(overlap_calls=True), "
                           "blp_plan_set_overlap): batch n + 1 starts on the SMs batch n has left; every chunk is a resident tensor")
        legs[name] = leg

    # one call per evaluation set, like the headline step (FB15k-237: 20,480 test triples, WN18RR: 3,136)
    sweep_leg("fb15k237_distmult_exact", "fb15k237", "distmult", "exact", 20480)
    sweep_leg("fb15k237_distmult_fast", "fb15k237", "distmult", "fast", 20480)
    sweep_leg("wn18rr_complex_exact", "wn18rr", "complex", "exact", 3136)
    sweep_leg("wn18rr_complex_fast", "wn18rr", "complex", "fast", 3136)
    sweep_leg("fb15k237_distmult_fast_exact", "fb15k237", "distmult", "fast_exact", 20480)
    sweep_leg("wn18rr_complex_fast_exact", "wn18rr", "complex", "fast_exact", 3136)
    for ds_model in ("fb15k237_distmult", "wn18rr_complex"):
        # the exact leg ranks the same triples (relation-aligned order only permutes them): identical fp64 metric sums
        # up to the order of the final fp64 additions, so compare the hit COUNTS exactly and the MRR sum to 1e-12
        a_, b_ = sums_of[ds_model + "_exact"], sums_of[ds_model + "_fast_exact"]
        legs[ds_model + "_fast_exact"]["metrics_equal_exact_leg"] = bool(
            torch.equal(a_[1:], b_[1:]) and abs(float(a_[0]) - float(b_[0])) <= 1e-12 * abs(float(a_[0])))
    sweep_leg("fb15k237_transe_eval_batch_64", "fb15k237", "transe", "exact", 64)     # the reference's eval_batch_size
    sweep_leg("fb15k237_transe_eval_batch_64_overlapped", "fb15k237", "transe", "exact", 64, overlap=True)

    # BOW script widths (TransE, D = 300 glove-bow / 768 bert-bow, scripts/test-umls.sh): the D != 128 path
    for d_bow, name in ((300, "fb15k237_transe_d300"), (768, "fb15k237_transe_d768")):
        try:
            n, r = CONFIGS["fb15k237"][:2]
            g = torch.Generator().manual_seed(11)
            ent = torch.nn.functional.normalize(torch.randn(n, d_bow, generator=g), dim=-1).to(dev)
            rel = ((torch.rand(r, d_bow, generator=g) * 2 - 1) * (6.0 / (r + d_bow)) ** 0.5).to(dev)
            e = 64
            tr = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g),
                              torch.randint(0, r, (e,), generator=g)], dim=1).to(dev)
            ms = time_calls(lambda: blp_b200.rank_sweep("transe", ent, rel, tr), reps=5, warm=2)
            ops_alg = 2 * e * n * d_bow * ALG_OPS["transe"]
            peak = fp32_peak or 37.2
            legs[name] = {"config": f"synthetic fb15k237 BLP-transe dim={d_bow} (BOW encoder width), exact mode, {e} test triples per call",
                          "ms_per_call": ms, "scores_per_s": 2 * e * n / (ms * 1e-3),
                          "roofline": {"bound": "fp32_alu", "achieved": ops_alg / (ms * 1e-3) / 1e12, "peak": peak,
                                       "unit": "T lane-op/s", "frac": ops_alg / (ms * 1e-3) / 1e12 / peak}}
        except Exception as exc:
            legs[name] = {"error": str(exc)[:160]}

    # Wikidata5M training batch (scripts/*wikidata5m*: batch_size = 1024) at K = 512 and the reference's K = 64
    lib = ops.lib()
    try:
        # the backward scatter adds one 512-byte gradient row per negative into the (2B, 128) table: the rate of the L2
        # atomic units, measured with the same access pattern and nothing else (blp_atomic_probe)
        atomic_gbs, _ = ops.atomic_probe(dev, rows=2048, iters=256)
    except Exception:
        atomic_gbs = None
    for b, k in ((1024, 512), (1024, 64), (64, 512)):
        g = torch.Generator().manual_seed(0)
        x = torch.nn.functional.normalize(torch.randn(b, 2, 128, generator=g), dim=-1).to(dev)
        rel = ((torch.rand(822, 128, generator=g) * 2 - 1) * 0.08).to(dev)
        rels = torch.randint(0, 822, (b, 1), generator=g).to(dev)
        neg = blp_b200.get_negative_sampling_indices(b, k, device=dev, seed=1)
        reps = 20
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a_, b_ in evs:
            a_.record(); b_.record()
        for _ in range(3):
            ops.train_loss("transe", "margin", x, rel, rels, neg, want_grad=True)
        torch.cuda.synchronize()
        for a_, b_ in evs:
            lib.blp_profile_events(2, ctypes.c_void_p(a_.cuda_event), ctypes.c_void_p(b_.cuda_event))
            ops.train_loss("transe", "margin", x, rel, rels, neg, want_grad=True)
        lib.blp_profile_events(0, None, None)
        torch.cuda.synchronize()
        us = sorted(a_.elapsed_time(b_) * 1e3 for a_, b_ in evs)[reps // 2]
        lane = b * (k + 1) * 128 * 6.0                                  # fwd 3 + bwd 3 lane-ops per (triple, dim)
        alg = 2 * b * 128 * 4 + b * 128 * 4 + b * 8 + b * k * 16 + 3 * b * 128 * 4
        peak = fp32_peak or 37.2
        floor_us = max(lane / (peak * 1e6), alg / (hbm_peak * 1e3))
        bound = "fp32_alu" if lane / (peak * 1e6) >= alg / (hbm_peak * 1e3) else "hbm"
        roof = {"bound": bound, "floor_us": floor_us, "frac": floor_us / us, "algorithmic_bytes": alg, "lane_ops": lane}
        if atomic_gbs:
            # every active negative adds one 128-float gradient row (its corrupting entity) to grad_ent: with the margin
            # loss on random embeddings nearly all B * K negatives are active
            scatter_bytes = b * k * 128 * 4
            atomic_us = scatter_bytes / (atomic_gbs * 1e3)
            roof.update({"l2_atomic_GBps_probe": atomic_gbs, "scatter_bytes": scatter_bytes, "l2_atomic_floor_us": atomic_us,
                         "l2_atomic_frac": atomic_us / us})
            if atomic_us > floor_us:
                roof.update({"bound": "l2_atomic", "floor_us": atomic_us, "frac": atomic_us / us,
                             "note": "the gradient scatter (one fp32 row reduction per negative) is paced by the L2 atomic units: "
                                     "floor = scatter bytes / red.global.add.v4.f32 rate measured in this run with the same access "
                                     "pattern (blp_atomic_probe); fp32_alu / hbm floors kept in lane_ops / algorithmic_bytes"})
        legs[f"train_transe_b{b}_k{k}"] = {
            "config": f"compute_loss forward + backward, BLP-transe dim=128, B={b}, K={k}, margin loss (one fused launch)",
            "us_per_step_kernels": us, "triples_per_s": b * (k + 1) / (us * 1e-6), "roofline": roof}

    # entity-table production (F.normalize + the write into a row shard, train.py:95-123, models.py:38-43):
    # a pure stream, 2 * N * D * 4 bytes, against the measured HBM copy peak
    for rows in (600000, 4800000):
        try:
            raw = torch.randn(rows, 128, device=dev)
            shard = torch.empty_like(raw)
            ms = time_calls(lambda: blp_b200.store_rows(shard, raw, normalize=True), reps=10, warm=3)
            nbytes = 2 * rows * 512
            legs[f"store_rows_{rows}"] = {
                "config": f"blp_store_rows: L2-normalise (ATen order, bit-equal) + store {rows} rows of dim=128 into a row shard",
                "ms_per_call": ms,
                "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": nbytes}}
            del raw, shard
        except Exception as exc:
            legs[f"store_rows_{rows}"] = {"error": str(exc)[:160]}
    torch.cuda.empty_cache()
    return legs


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
    C = substeps(w, e)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    flush = not args.no_flush
    if args.mode == "fast" and args.model == "transe":
        raise SystemExit("--mode fast covers distmult / complex / simple (TransE is an L1 distance, not a contraction)")

    # ---- resident inputs (value): everything already in HBM
    ent = w["ent"].to(dev)
    model = blp_b200.TransductiveLinkPrediction(d, args.model, args.loss, 8, w["r"], 0).to(dev)
    with torch.no_grad():
        model.rel_emb.weight.copy_(w["rel"])
    rel_w = model.rel_emb.weight
    # each rank works on its own slice of the test triples / its own training sub-batch; the evaluation set is put in
    # relation order once (what rank_sweep(sort_by_relation=True) does per sweep): triples that share a relation let
    # the TransE kernel reuse fl(candidate + r) across head-prediction queries
    test_triples = w["triples"].roll(-rank * e, 0)
    test_triples = test_triples[torch.argsort(test_triples[:, 2], stable=True)].contiguous()
    triples = test_triples.to(dev)
    ent_embs = ent[w["pairs"].to(dev)].contiguous()
    rels = w["rels"].to(dev)
    neg = w["neg_storage"].to(dev).transpose(0, 1)                         # (B,K,2), strides of the reference sampler
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
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
        launches["n"] += 2                # fused fwd+bwd kernel + the blp_scale of autograd's backward
        return loss

    chunks = [triples[(torch.arange(c * e, (c + 1) * e, device=dev) % t)].contiguous() for c in range(C)]
    fast_ws = ops.fast_table(ent) if args.mode == "fast" else None
    whole_sweep = args.mode == "exact"
    if whole_sweep:
        # exact mode: the whole evaluation set in ONE launch per step; relation-aligned order, prepared once per set
        eval_set = blp_b200.AlignedTriples(triples[:C * e].contiguous())
        plan = None
    else:
        # tensor-core mode: pre-validated sweep for E triples per call (fold + sweep + metrics launches)
        plan = blp_b200.RankSweepPlan(args.model, ent, rel_w, e, mode=args.mode, fast_table=fast_ws)

    # CUDA events bracketing the dominant kernel (the eval sweep) of every timed launch, on its own stream
    n_prof = steps * (1 if whole_sweep else C)
    prof = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_prof)]
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a_, b_ in prof + step_ev:          # materialise the cudaEvent_t handles
        a_.record(); b_.record()

    def set_prof(idx):
        pa, pb = prof[idx]
        lib.blp_profile_events(1, ctypes.c_void_p(pa.cuda_event), ctypes.c_void_p(pb.cuda_event))

    def step(i, timed_index=None):
        loss = out = None
        if whole_sweep:
            # the sweep is enqueued first: the launch-bound training sub-steps are then issued while the GPU ranks
            if timed_index is not None:
                set_prof(timed_index)
            out = blp_b200.rank_sweep(args.model, ent, rel_w, eval_set, sort_by_relation=False)
            launches["n"] += out["launches"]
        for c in range(C):
            loss = train_step()
            if not whole_sweep:
                if timed_index is not None:
                    set_prof(timed_index * C + c)
                out = plan(chunks[c])
                launches["n"] += out["launches"]
        return loss, out

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
    time.sleep(0.02)
    t_wall0 = time.perf_counter()
    for i in range(steps):
        if flush:
            flush_buf.fill_(i & 0xFF)
        step_ev[i][0].record()
        step(warmup + i, timed_index=i)
        step_ev[i][1].record()
    lib.blp_profile_events(0, None, None)
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    timed_launches = launches["n"]
    step_ms = [a_.elapsed_time(b_) for a_, b_ in step_ev]
    kern_ms = [a_.elapsed_time(b_) for a_, b_ in prof]
    total_ms = sum(step_ms)
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * steps * step_triples(w, e) / (total_ms * 1e-3)

    # ---- e2e: public API, host (pinned) inputs, H2D + D2H inside the timed region, every sub-step
    pin = lambda x: x.contiguous().pin_memory()  # noqa: E731
    h_ent_embs = pin(w["ent"][w["pairs"]])
    h_rels = pin(w["rels"])
    h_neg = pin(w["neg_storage"])
    h_triples = [pin(test_triples[torch.arange(c * e, (c + 1) * e) % t]) for c in range(C)]
    h2d_sub = h_ent_embs.numel() * 4 + h_rels.numel() * 8 + h_neg.numel() * 8 + h_triples[0].numel() * 8
    d2h_sub = 4 + 4 * 8
    # results are read on the host one STEP behind the GPU (two pinned result slots), like a training loop that logs the
    # previous step's losses and metrics while the next step is already queued: every step still does all of its H2D
    # and D2H copies inside the timed region
    host_loss = [torch.empty(C, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_sums = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)]
    landed = [torch.cuda.Event(), torch.cuda.Event()]
    # host -> device prefetch, the way a pinned-memory DataLoader feeds a training loop: the inputs of sub-step j + 1 cross
    # PCIe on a copy stream into one of two device staging slots while sub-step j computes
    copy_stream = torch.cuda.Stream()
    stage = [{"ent": torch.empty_like(h_ent_embs, device=dev), "rels": torch.empty_like(h_rels, device=dev),
              "neg": torch.empty_like(h_neg, device=dev), "tr": torch.empty_like(h_triples[0], device=dev)} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    h_eval = pin(test_triples[:C * e])
    if whole_sweep:
        h2d_step = C * (h_ent_embs.numel() * 4 + h_rels.numel() * 8 + h_neg.numel() * 8) + h_eval.numel() * 8
        d2h_step = C * 4 + 4 * 8
    else:
        h2d_step, d2h_step = C * h2d_sub, C * d2h_sub
    dev_eval = torch.empty_like(h_eval, device=dev)
    eval_consumed = [None]

    # graph replay: two captured steps (ping-pong) whose static input buffers ARE the staging slots, so the pinned host
    # inputs land where the graph reads them (no device-to-device hop)
    graphs2 = None if graphed is None else [graphed, blp_b200.GraphedLossStep(model, b, k)]

    def e2e_prefetch(j):
        slot = stage[j & 1]
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(consumed[j & 1])            # the sub-step that last used this slot has read it
            if graphs2 is not None:
                g_ = graphs2[j & 1]
                g_.ent_embs.copy_(h_ent_embs.reshape(g_.ent_embs.shape), non_blocking=True)
                g_.rels.copy_(h_rels.reshape(g_.rels.shape), non_blocking=True)
                g_._neg_storage.copy_(h_neg.reshape(g_._neg_storage.shape), non_blocking=True)
            else:
                slot["ent"].copy_(h_ent_embs, non_blocking=True)
                slot["rels"].copy_(h_rels, non_blocking=True)
                slot["neg"].copy_(h_neg, non_blocking=True)
            if not whole_sweep:
                slot["tr"].copy_(h_triples[j % C], non_blocking=True)
            ready[j & 1].record(copy_stream)

    def e2e_issue(j, first):
        main = torch.cuda.current_stream()
        if first:
            e2e_prefetch(j)
        e2e_prefetch(j + 1)
        main.wait_event(ready[j & 1])
        slot = stage[j & 1]
        if graphs2 is not None:
            loss, _ = graphs2[j & 1].replay()                           # one graph launch on the inputs that just landed
        else:
            x = slot["ent"].clone().requires_grad_(True)
            rel_w.grad = None
            loss = model.compute_loss(x, slot["rels"], slot["neg"].transpose(0, 1))
            loss.backward()
        out = plan(slot["tr"]) if not whole_sweep else None
        consumed[j & 1].record(main)
        # D2H: the loss scalar (train.py:352) [+ the 4 fp64 metric accumulators (train.py:154-157) per call]
        s_i, c = divmod(j, C)
        host_loss[s_i & 1][c:c + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        if out is not None:
            host_sums[s_i & 1].copy_(out["sums"], non_blocking=True)

    def e2e_read(s_i):
        landed[s_i & 1].synchronize()
        return float(host_loss[s_i & 1][C - 1]), float(host_sums[s_i & 1][0]) / (2 * (C * e if whole_sweep else e))

    def e2e_copy_eval():
        """The evaluation set of the NEXT step crosses PCIe on the copy stream while this step computes."""
        with torch.cuda.stream(copy_stream):
            if eval_consumed[0] is not None:
                copy_stream.wait_event(eval_consumed[0])           # the staging buffer has been gathered from
            dev_eval.copy_(h_eval, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_run(n_steps):
        last = None
        eval_ready = e2e_copy_eval() if (whole_sweep and n_steps) else None
        for s_i in range(n_steps):
            if whole_sweep:
                main = torch.cuda.current_stream()
                main.wait_event(eval_ready)
                ev_set = eval_set.load(dev_eval)
                eval_consumed[0] = torch.cuda.Event()
                eval_consumed[0].record(main)
                out = blp_b200.rank_sweep(args.model, ent, rel_w, ev_set, sort_by_relation=False)
                host_sums[s_i & 1].copy_(out["sums"], non_blocking=True)
                if s_i + 1 < n_steps:
                    eval_ready = e2e_copy_eval()
            for c in range(C):
                j = s_i * C + c
                e2e_issue(j, first=(j == 0))
            landed[s_i & 1].record()
            if s_i > 0:
                last = e2e_read(s_i - 1)
        if n_steps:
            last = e2e_read(n_steps - 1)
        return last

    e2e_run(min(warmup, 3))
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    last = e2e_run(steps)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = world * steps * step_triples(w, e) / (e2e_ms * 1e-3)

    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = measure_fp32_rate(ops, dev)

    # ---- legs: the entity-sharded Wikidata5M-scale sweep (every N) and the other BASELINE configs (N = 1)
    del flush_buf
    torch.cuda.empty_cache()
    wd = legs = None
    if not args.no_extra:
        try:
            wd = wd_sweep_leg(args, dev, world, rank, hbm_peak)
        except Exception as exc:        # a leg must not take the headline down with it
            wd = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            if world > 1:
                raise
        if world == 1 and not args.no_legs:
            try:
                legs = config_legs(args, dev, peaks, fp32_peak, hbm_peak)
            except Exception as exc:
                legs = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the eval sweep kernel), measured live above, against the bound that binds it
    kern_s = statistics.mean(kern_ms) * 1e-3
    kname = f"{'fast_sweep_kernel' if args.mode == 'fast' else 'sweep_kernel'}<{args.model}>"
    share = sum(kern_ms) / sum(step_ms)
    roofline = {"kernel": kname}
    q_launch = 2 * C * e if whole_sweep else 2 * e                 # real queries ranked by one sweep launch
    roofline.update(sweep_roofline(args.model, args.mode, n, d, q_launch, kern_s, peaks, fp32_peak, hbm_peak,
                                   exec_ops=EXEC_OPS_ALIGNED.get(args.model) if whole_sweep else None))
    if whole_sweep:
        roofline["launch"] = (f"{eval_set.num_padded} entries per launch: {eval_set.num_triples} test triples in relation-aligned order + "
                              f"{eval_set.num_padded - eval_set.num_triples} padding entries (computed, not counted)")
    roofline["traffic"] = ncu_traffic(f"{kname}|{args.dataset}|E{C * e if whole_sweep else e}|{args.mode}")
    roofline["kernel_share_of_step"] = share
    roofline["launches_timed"] = len(kern_ms)
    roofline["timed"] = "CUDA events around every sweep launch of the timed region (blp_profile_events, launching stream)"
    # flat copies of the legs' rooflines: the HBM-bound shape of the same kernel, the tensor-core sweeps, the other configs
    if wd and "table_pass_2" in wd:
        p2 = wd["table_pass_2"]
        roofline.update({"wd_bound": "hbm", "wd_achieved_GBps_all_ranks": p2["algorithmic_GBps_all_ranks"], "wd_peak_GBps_per_gpu": hbm_peak,
                         "wd_frac": p2["hbm_frac"], "wd_traffic_1gpu": ncu_traffic("sweep_kernel<transe>|wikidata5m|E2|exact"),
                         "wd_config": f"Wikidata5M-scale TransE sweep, {wd['entities']} rows sharded over {world} GPU(s), table pass size 2, "
                                      f"T={wd['test_triples']}, one launch per rank + one all-reduce",
                         "sharded_parity": wd["parity"]["sharded_parity"]})
    if legs and "error" not in legs:
        for nm, tag in (("fb15k237_distmult_fast", "distmult_fast"), ("fb15k237_distmult_exact", "distmult_exact"),
                        ("fb15k237_distmult_fast_exact", "distmult_fast_exact"), ("wn18rr_complex_fast_exact", "complex_fast_exact"),
                        ("wn18rr_complex_fast", "complex_fast"), ("wn18rr_complex_exact", "complex_exact"),
                        ("fb15k237_transe_eval_batch_64", "transe_e64"), ("fb15k237_transe_eval_batch_64_overlapped", "transe_e64_overlapped"),
                        ("fb15k237_transe_d768", "transe_d768"),
                        ("train_transe_b1024_k512", "train_b1024"), ("store_rows_4800000", "store_rows_4p8m")):
            r = legs.get(nm, {}).get("roofline")
            if r:
                roofline[tag + "_bound"] = r["bound"]
                roofline[tag + "_frac"] = r["frac"]
                if "issued_frac" in r:
                    roofline[tag + "_issued_frac"] = r["issued_frac"]
                if "sweep_kernel_issued_frac" in r:
                    roofline[tag + "_kernel_issued_frac"] = r["sweep_kernel_issued_frac"]

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        arm = CpuArm(w, args)
        r = arm.run(e, steps=3, warmup=1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == "exact" else "f16x3 split of f32 (fp32 accumulate; train step f32)", "data": "synthetic",
        "config": config_dict(args, w, e, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "ms_per_step": e2e_ms / steps, "last_loss": last[0], "last_mrr": last[1]},
        "gpu_launches": timed_launches,
        "roofline": roofline, "cpu_baseline": cpu,
        "sharded_parity": wd["parity"]["sharded_parity"] if wd and "parity" in wd else None,
        "legs": legs, "wikidata5m_scale_sweep": wd,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
