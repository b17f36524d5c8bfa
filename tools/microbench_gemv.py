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


def bench(name, d, n, pro, epi, nb=1, iters=20):
    w = rand_dev(d, n)
    x = torch.randn(nb, n, device="cuda")
    nw = torch.ones(n, device="cuda")
    out = torch.zeros(nb, d if epi != ops.EPI_SWIGLU else d // 2, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    times = []
    for i in range(iters + 3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.gemv_q40(w, x, pro=pro, epi=epi, out=out, norm_w=nw)
        e.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(s.elapsed_time(e) * 1e3)
    us = float(np.median(times))
    byts = d * n // 2 + d * n // 32 * 2
    return dict(kernel=name, d=d, n=n, nb=nb, us=round(us, 2), gbs=round(byts / us / 1e3, 1))


if __name__ == "__main__":
    peaks = json.load(open("MEASURED_PEAKS.json")) if len(sys.argv) < 2 else {"hbm_gbs": float(sys.argv[1])}
    rows = [bench("qkv", 6144, 4096, ops.PRO_RMSNORM, ops.EPI_STORE),
            bench("wo", 4096, 4096, ops.PRO_PLAIN, ops.EPI_RESIDUAL),
            bench("w13", 28672, 4096, ops.PRO_RMSNORM, ops.EPI_SWIGLU),
            bench("w2", 4096, 14336, ops.PRO_PLAIN, ops.EPI_RESIDUAL),
            bench("logits", 128256, 4096, ops.PRO_RMSNORM, ops.EPI_STORE),
            bench("w13_nb4", 28672, 4096, ops.PRO_RMSNORM, ops.EPI_SWIGLU, nb=4),
            bench("w13_nb8", 28672, 4096, ops.PRO_RMSNORM, ops.EPI_SWIGLU, nb=8)]
    for r in rows:
        r["frac_of_measured_hbm"] = round(r["gbs"] / peaks["hbm_gbs"], 3)
        print(json.dumps(r))
