"""Smallest TMA-variant GEMM (d=256, n=256, T=16) against fp32 — used to bisect pipeline changes under a hard timeout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_llama_b200 import ops
from tools.microbench_gemm import rand_dev
d, n, T = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 256, 16)))
w = rand_dev(d, n)
act = torch.randn(T, n, device="cuda").bfloat16()
out = torch.zeros(T, d, device="cuda")
ops.gemm_q40_tc(w, act, epi=ops.GEPI_STORE_F32, out=out, variant="tma")
torch.cuda.synchronize()
ref = torch.zeros_like(out)
ops.gemm_q40_tc(w, act, epi=ops.GEPI_STORE_F32, out=ref, variant="ldg")
torch.cuda.synchronize()
print("max|tma - ldg| =", (out - ref).abs().max().item(), "max|ref| =", ref.abs().max().item())
