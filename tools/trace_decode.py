"""Per-kernel timeline of one decode step under PDL + CUDA graph (device globaltimer stamps)."""
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
prompt = [(7 * i + 3) % 1000 + 1 for i in range(64)]
eng.prefill(prompt[:-1], 0, want_logits=False)
eng.decode_greedy(prompt[-1], 63, 8)
torch.cuda.synchronize()
eng.trace_buf.zero_()
eng.decode_greedy(prompt[-1], 63, 1)
torch.cuda.synchronize()
tr = eng.read_trace().astype(np.float64)
if rank != 0:
    sys.exit(0)
t0 = tr[:, 0].min()
names = ["qkv", "attn", "wo", "w13", "w2"]
print("slot name   entry  waitdone prologue  exit | dur  gap_from_prev_exit  (us)")
prev_exit = None
rows = []
for i, r in enumerate(tr):
    nm = names[i % 5] if i < len(tr) - 1 else "logits"
    e, w, p, x = (r - t0) / 1e3
    if r[2] == 0: p = float('nan')
    gap = (e - prev_exit) if prev_exit is not None else 0.0
    rows.append((nm, x - e, x - w, w - e, gap))
    if i < 12 or i >= len(tr) - 3:
        print(f"{i:4d} {nm:6s} {e:8.2f} {w:8.2f} {p:8.2f} {x:8.2f} | {x-e:6.2f} {gap:7.2f}")
    prev_exit = x
print("total step us:", (tr[:, 3].max() - t0) / 1e3)
for nm in names + ["logits"]:
    sel = [r for r in rows if r[0] == nm]
    if sel:
        a = np.array([[r[1], r[2], r[3], r[4]] for r in sel])
        print(f"{nm:6s} n={len(sel):3d} mean dur={a[:,0].mean():6.2f} after-wait={a[:,1].mean():6.2f} wait={a[:,2].mean():6.2f} entry-gap={a[:,3].mean():6.2f}")
