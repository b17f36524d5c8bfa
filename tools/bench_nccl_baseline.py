"""Fused in-kernel all-reduce vs the NCCL baseline on the decode step (run under torchrun, one rank per GPU).
Both paths run the same GEMV/attention kernels; the baseline stores the partial WO / W2 products and all-reduces them
with NCCL (captured in a torch CUDA graph, so no host overhead is charged to it). Device-timed, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import ensure_model
from distributed_llama_b200.api import InferenceSession
from distributed_llama_b200.parallel.comm import Communicator

os.environ["NCCL_DEBUG"] = "WARN"
model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
comm = Communicator()
if comm.rank == 0:
    ensure_model(model)
dist.barrier()
m, t = ensure_model(model)
sess = InferenceSession(m, t, max_seq_len=2048, comm=comm)
eng = sess.engine
prompt = [(7 * i + 3) % 1000 + 1 for i in range(64)]
eng.prefill(prompt[:-1], 0, want_logits=False)


def timed(fn):
    dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    tms = torch.tensor([s.elapsed_time(e)], device="cuda", dtype=torch.float64)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    return float(tms.item())

res = {}
for name, mega in (("fused_megakernel", True), ("fused_multi_kernel", False)):
    eng.enable_mega(mega)
    eng.decode_greedy(prompt[-1], 63, 8)
    res[name] = timed(lambda: eng.decode_greedy(prompt[-1], 63, steps)) / steps
# NCCL baseline: graph-captured single step (fixed position: the timing does not depend on it)
eng._set_inputs([prompt[-1]], 63)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        eng.forward_nccl_baseline(1)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng.forward_nccl_baseline(1)
for _ in range(8):
    g.replay()
res["nccl_baseline_graph"] = timed(lambda: [g.replay() for _ in range(steps)]) / steps
if comm.rank == 0:
    out = {"model": model, "n_gpus": comm.world_size, "steps": steps, "ms_per_step": {k: round(v, 4) for k, v in res.items()},
           "speedup_fused_vs_nccl": round(res["nccl_baseline_graph"] / res["fused_megakernel"], 2)}
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
