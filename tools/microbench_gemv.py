"""GEMV microbenchmark at the Llama-3.1-8B decode shapes: achieved HBM bandwidth per kernel (CUDA events, L2 flushed)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from distributed_llama_b200 import ops
from distributed_llama_b200.ops import DeviceQ40


def rand_dev(d, n):
    w = DeviceQ40.empty(d, n)
    w.qs.random_(-2**31, 2**31 - 1)
    w.scales.copy_((torch.rand(d, n // 32, device="cuda") * 0.01).half())
    return w


def bench(name, d, n, pro, epi, nb=1, iters=20, impl="auto"):
    w = rand_dev(d, n)
    x = torch.randn(nb, n, device="cuda")
    nw = torch.ones(n, device="cuda")
    out = torch.zeros(nb, d if epi != ops.EPI_SWIGLU else d // 2, device="cuda")
    flush = torch.ones(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    evs = []
    # queue everything first so the CPU stays ahead of the GPU; a *read* of 512 MB leaves L2 full of clean lines
    for i in range(iters + 3):
        _ = flush.view(torch.int64).sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.gemv_q40(w, x, pro=pro, epi=epi, out=out, norm_w=nw, impl=impl)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    times = [s.elapsed_time(e) * 1e3 for s, e in evs[3:]]
    us = float(np.median(times))
    byts = d * n // 2 + d * n // 32 * 2
    return dict(impl=impl, kernel=name, d=d, n=n, nb=nb, us=round(us, 2), gbs=round(byts / us / 1e3, 1))


if __name__ == "__main__":
    peaks = json.load(open("MEASURED_PEAKS.json")) if len(sys.argv) < 2 else {"hbm_gbs": float(sys.argv[1])}
    rows = []
    for impl in ("ldg", "tma"):
        rows += [bench("qkv", 6144, 4096, ops.PRO_RMSNORM, ops.EPI_STORE, impl=impl),
                 bench("wo", 4096, 4096, ops.PRO_PLAIN, ops.EPI_RESIDUAL, impl=impl),
                 bench("w13", 28672, 4096, ops.PRO_RMSNORM, ops.EPI_SWIGLU, impl=impl),
                 bench("w2", 4096, 14336, ops.PRO_PLAIN, ops.EPI_RESIDUAL, impl=impl),
                 bench("logits", 128256, 4096, ops.PRO_RMSNORM, ops.EPI_STORE, impl=impl),
                 bench("w13_nb4", 28672, 4096, ops.PRO_RMSNORM, ops.EPI_SWIGLU, nb=4, impl=impl),
                 bench("w13_nb8", 28672, 4096, ops.PRO_RMSNORM, ops.EPI_SWIGLU, nb=8, impl=impl)]
    for r in rows:
        r["frac_of_measured_hbm"] = round(r["gbs"] / peaks["hbm_gbs"], 3)
        print(json.dumps(r))
