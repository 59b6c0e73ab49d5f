#!/bin/bash
python - <<'PY'
import torch, blp_b200
dev = torch.device("cuda", 0)
def gpu_ms(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for m in (600000, 4800000):
    raw = torch.randn(m, 128, device=dev); shard = torch.empty_like(raw)
    for norm in (True, False):
        ms = gpu_ms(lambda: blp_b200.store_rows(shard, raw, normalize=norm))
        print(f"store_rows m={m} normalize={norm}: {ms:.3f} ms, {2*m*512/ms/1e6:.0f} GB/s")
    ms = gpu_ms(lambda: shard.copy_(raw))
    print(f"torch copy m={m}: {ms:.3f} ms, {2*m*512/ms/1e6:.0f} GB/s")
    ms = gpu_ms(lambda: torch.nn.functional.normalize(raw, dim=-1, out=shard))
    print(f"torch F.normalize m={m}: {ms:.3f} ms, {2*m*512/ms/1e6:.0f} GB/s")
    del raw, shard
PY
