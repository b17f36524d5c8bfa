"""Phase timeline of the persistent decode kernel (one %globaltimer stamp at every phase start on CTA 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import ensure_model
from distributed_llama_b200.api import InferenceSession

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b"
comm = None
rank = int(os.environ.get("RANK", "0"))
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    from distributed_llama_b200.parallel.comm import Communicator
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
    comm = Communicator()
    if rank == 0:
        ensure_model(model)
    dist.barrier()
m, t = ensure_model(model)
sess = InferenceSession(m, t, max_seq_len=2048, comm=comm)
eng = sess.engine
eng.enable_trace(4096)
eng.enable_mega()
prompt = [(7 * i + 3) % 1000 + 1 for i in range(64)]
eng.prefill(prompt[:-1], 0, want_logits=False)
eng.decode_greedy(prompt[-1], 63, 8)
torch.cuda.synchronize()
eng.trace_buf.zero_()
eng.decode_greedy(prompt[-1], 63, 1)
torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)
st = eng.trace_buf.cpu().numpy().reshape(-1)
st = st[st != 0].astype(np.float64)
st = (st - st[0]) / 1e3
L = eng.w.header.n_layers
names = ["embed"]
for _ in range(L):
    names += ["qkv.pro", "qkv.main", "qkv.epi", "qkv.bar", "attn.run", "attn.bar", "wo.pro", "wo.main", "wo.epi", "wo.bar",
              "w13.pro", "w13.main", "w13.epi", "w13.bar", "w2.pro", "w2.main", "w2.epi", "w2.bar"]
names += ["logits.pro", "logits.main", "logits.epi"]
dur = np.diff(st)
print("total step us:", st[-1], "stamps", len(st), "expected", len(names) + 1)
agg = {}
for i, d in enumerate(dur):
    agg.setdefault(names[i] if i < len(names) else "?", []).append(d)
for nm, d in agg.items():
    print(f"{nm:12s} n={len(d):3d} mean={np.mean(d):7.2f} us  min={np.min(d):7.2f} max={np.max(d):7.2f}")
