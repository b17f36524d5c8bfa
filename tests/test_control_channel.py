"""Shared-memory control channel (parallel/control.py): packets reach every worker in order, idle workers survive through
heartbeats, and a peer that stops responding surfaces as PeerLost instead of a hang (reference: control packets over TCP +
socket exceptions, src/app.cpp:197-230,356-362)."""
import multiprocessing as mp
import time

import pytest

from distributed_llama_b200.parallel.control import ControlChannel, PeerLost


def _worker(rank, world, name, q, stop_acking_after=None):
    ch = ControlChannel(rank, world, name=name, timeout_s=2.0)
    ch.start_heartbeat(0.05)
    got = []
    while True:
        if stop_acking_after is not None and len(got) >= stop_acking_after:
            time.sleep(30)      # wedged worker: alive, but never consumes another packet
            return
        op, pos, toks = ch.recv()
        got.append((op, pos, toks))
        if op == 0:
            break
    q.put((rank, got))
    ch.close()


def test_packets_reach_all_workers_in_order():
    ctx = mp.get_context("fork")
    root = ControlChannel(0, 3, timeout_s=10.0)
    root.start_heartbeat(0.05)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 3, root.name, q)) for r in (1, 2)]
    for p in procs:
        p.start()
    sent = [(1, 0, list(range(100))), (3, 100, [7]), (3, 101, [9]), (2, 102, [11]), (4, 103, [5, 64]), (0, 0, [])]
    for op, pos, toks in sent:
        root.send(op, pos, toks)
    res = dict(q.get(timeout=30) for _ in procs)
    for p in procs:
        p.join(timeout=10)
    root.close()
    assert res[1] == sent and res[2] == sent


def test_idle_worker_is_not_declared_lost_and_sleeping_poll_wakes_up():
    ctx = mp.get_context("fork")
    root = ControlChannel(0, 2, timeout_s=2.0)
    root.start_heartbeat(0.05)
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(1, 2, root.name, q))
    p.start()
    time.sleep(3.0)                    # longer than the peer timeout and than the busy-poll window
    root.send(3, 5, [42])
    root.send(0, 0, [])
    rank, got = q.get(timeout=30)
    p.join(timeout=10)
    root.close()
    assert got == [(3, 5, [42]), (0, 0, [])]


def test_unresponsive_worker_raises_peer_lost():
    ctx = mp.get_context("fork")
    root = ControlChannel(0, 2, timeout_s=1.5)
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(1, 2, root.name, q, 1))
    p.start()
    root.send(3, 0, [1])               # consumed
    root.send(3, 1, [2])               # published, never acknowledged
    t0 = time.time()
    with pytest.raises(PeerLost):
        root.send(3, 2, [3])
    assert time.time() - t0 < 10
    p.kill()
    p.join(timeout=10)
    root.close()


def test_worker_detects_dead_root():
    root = ControlChannel(0, 2, timeout_s=0.5)
    w = ControlChannel(1, 2, name=root.name, timeout_s=0.5)
    time.sleep(0.8)                    # the root never beats again
    with pytest.raises(PeerLost):
        w.recv()
    w.close()
    root.close()
