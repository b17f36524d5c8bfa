import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
from distributed_llama_b200 import ops
from distributed_llama_b200.formats import ModelFile
from distributed_llama_b200.models.config import get_config
from distributed_llama_b200.models.loader import load_device_weights
from distributed_llama_b200.models.synthetic import write_synthetic_model
from distributed_llama_b200.models.reference import OracleModel
from distributed_llama_b200.runtime import Engine

name = "tiny-llama31"
path = f"/tmp/dbgp_{name}.m"
write_synthetic_model(path, get_config(name), seed=7)
mf = ModelFile(path); h = mf.header
W = load_device_weights(mf)
orc = OracleModel(mf, act_quant="none", device="cuda")
T = 45
toks = [(7 * i + 3) % 500 + 1 for i in range(T)]
x = orc.embedding[torch.tensor(toks, device="cuda")]
L0, D0 = orc.layers[0], W.layers[0]
y = orc._rms(x, L0["norm_0"], h.norm_epsilon)
xn = ops.rmsnorm_bf16(x.contiguous(), D0.norm0, h.norm_epsilon)
print("rmsnorm", (xn.float() - y).abs().max().item())
qkv_ref = torch.cat([y @ L0["q"].T, y @ L0["k"].T, y @ L0["v"].T], dim=1)
qkv = torch.zeros(T, qkv_ref.shape[1], device="cuda")
ops.gemm_q40_tc(D0.qkv, xn, epi=ops.GEPI_STORE_F32, out=qkv)
print("qkv gemm", (qkv - qkv_ref).abs().max().item(), qkv_ref.abs().max().item())
# full engine: compare TC prefill vs GEMV prefill for several T
for T in (9, 16, 17, 32, 45):
    toks = [(7 * i + 3) % 500 + 1 for i in range(T)]
    e1 = Engine(W); l1 = e1.prefill(toks, 0).clone()
    e2 = Engine(W); e2.use_tc_prefill = False; l2 = e2.prefill(toks, 0).clone()
    orc.reset(); lo = orc.forward(toks, 0)[-1]
    # compare KV caches layer 0
    kd = (e1.k_cache[0][:, :T].float() - e2.k_cache[0][:, :T].float()).abs().max().item()
    vd = (e1.v_cache[0][:, :T].float() - e2.v_cache[0][:, :T].float()).abs().max().item()
    kd1 = (e1.k_cache[1][:, :T].float() - e2.k_cache[1][:, :T].float()).abs().max().item()
    print(T, "tc-vs-gemv", (l1 - l2).abs().max().item(), "tc-vs-oracle", (l1 - lo).abs().max().item(), "gemv-vs-oracle", (l2 - lo).abs().max().item(), "k0", kd, "v0", vd, "k1", kd1)
