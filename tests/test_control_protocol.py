"""Root/worker control protocol over a real process group (gloo, 2 ranks, CPU): every forward the root issues must be mirrored
by the worker with the same op / position / tokens, and the stop packet must end the worker loop.
Reference behaviour: RootLlmInference::forward + WorkerLlmInference::tryReadControlPacket (src/app.cpp:168-230)."""
import json
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeEngine:
    def __init__(self, log):
        self.log = log
        self.tokens = torch.zeros(8, dtype=torch.int32)

    def prefill(self, tokens, pos, want_logits=True):
        self.log.append(["prefill", list(tokens), pos])

    def step(self, token, pos):
        self.log.append(["step", token, pos])
        return torch.zeros(4)

    def _set_inputs(self, tokens, pos):
        self.log.append(["set_inputs", list(tokens), pos])
        self.tokens[0] = tokens[0] + 1

    def run_decode_step(self, use_graph=True):
        self.log.append(["decode_step"])


class _FakeSession:
    def __init__(self, log):
        self.engine = _FakeEngine(log)
        self.device = torch.device("cpu")
        self.header = None


def _rank_main(rank, world, init_file, out_dir, use_shm):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from distributed_llama_b200.apps.runtime import RootInference, open_control_channel, worker_loop

    class _Comm:
        world_size = world
        single_node = True
    _Comm.rank = rank
    log = []
    sess = _FakeSession(log)
    chan = open_control_channel(_Comm()) if use_shm else None      # shared-memory packets (default) or torch.distributed broadcasts
    assert (chan is not None) == use_shm
    if rank == 0:
        inf = RootInference(sess, _Comm(), chan)
        inf.prefill([5, 6, 7, 8, 9], 0)
        inf.prefill([], 5)                  # empty chunk: nothing is sent
        inf.forward_logits(11, 5)
        tok = inf.forward_greedy(12, 6)
        assert tok == 13
        inf.finish()
    else:
        worker_loop(sess, _Comm(), chan)
    with open(os.path.join(out_dir, f"log{rank}.json"), "w") as f:
        json.dump(log, f)
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("use_shm", [True, False])
def test_worker_mirrors_root_forwards(use_shm):
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        mp.spawn(_rank_main, args=(2, init_file, d, use_shm), nprocs=2, join=True)
        root = json.load(open(os.path.join(d, "log0.json")))
        worker = json.load(open(os.path.join(d, "log1.json")))
    assert root == worker
    assert root == [["prefill", [5, 6, 7, 8, 9], 0], ["step", 11, 5], ["set_inputs", [12], 6], ["decode_step"]]
