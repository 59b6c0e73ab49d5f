"""Time blp_b200.store_rows (normalise + write into a row shard, blp_store_rows) against a plain device copy:
    [BLP_STORE_STREAM=0] python tools/run_store.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blp_b200  # noqa: E402

dev = torch.device("cuda", 0)


def gpu_ms(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


tag = "register-tile kernel (BLP_STORE_STREAM=0)" if os.environ.get("BLP_STORE_STREAM") == "0" else "TMA stream kernel"
for m in (14541, 600000, 4800000):
    raw = torch.randn(m, 128, device=dev)
    shard = torch.empty_like(raw)
    nbytes = 2 * m * 512
    t_copy = gpu_ms(lambda: shard.copy_(raw))
    t_norm = gpu_ms(lambda: blp_b200.store_rows(shard, raw, normalize=True))
    t_plain = gpu_ms(lambda: blp_b200.store_rows(shard, raw, normalize=False))
    print(f"{tag}: {m} rows: normalise + store {t_norm:.4f} ms = {nbytes / t_norm / 1e6:.0f} GB/s, store only {t_plain:.4f} ms = "
          f"{nbytes / t_plain / 1e6:.0f} GB/s, torch copy_ {t_copy:.4f} ms = {nbytes / t_copy / 1e6:.0f} GB/s")
    del raw, shard
