"""CLI flag parser, NaiveCache and HTTP plumbing (CPU); end-to-end `dllama inference` / `dllama-api` on a GPU."""
import json
import os
import socket
import subprocess
import sys
import threading
import time

import pytest

from distributed_llama_b200.apps.args import parse_args
from distributed_llama_b200.apps.api_server import HttpRequest, NaiveCache, chunk_json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arg_parser_reference_surface():
    a = parse_args(["inference", "--model", "m.m", "--tokenizer", "t.t", "--prompt", "hi", "--steps", "16", "--buffer-float-type", "q80",
                    "--workers", "10.0.0.2:9998", "10.0.0.3:9998", "--nthreads", "4", "--temperature", "0.5", "--topp", "0.8",
                    "--seed", "7", "--chat-template", "llama3", "--max-seq-len", "2048", "--net-turbo", "0", "--port", "9991",
                    "--host", "127.0.0.1", "--gpu-index", "0", "--gpu-segments", "1:3"], True)
    assert a.mode == "inference" and a.workers == ["10.0.0.2:9998", "10.0.0.3:9998"] and a.steps == 16 and a.nthreads == 4
    assert a.temperature == 0.5 and a.topp == 0.8 and a.seed == 7 and a.chat_template == "llama3" and a.max_seq_len == 2048
    assert not a.net_turbo and a.port == 9991 and a.host == "127.0.0.1" and a.gpu_segments == "1:3" and a.n_batches == 32
    assert parse_args(["chat", "--help"], True).help and parse_args(["-h"], False).help
    for bad in (["x", "--bogus", "1"], ["x", "--workers", "nocolon"], ["x", "--buffer-float-type", "q99"], ["x", "--nthreads", "0"],
                ["x", "--gpu-segments", "3"], ["x", "--chat-template", "nope"]):
        with pytest.raises(ValueError):
            parse_args(bad, True)


def test_naive_cache_semantics():
    c = NaiveCache()
    msgs = [("system", "s"), ("user", "u1")]
    assert c.resolve_delta_prompt(list(msgs)) == (msgs, 0)
    c.push(10, msgs[0]); c.push(10, msgs[1]); c.push(25, ("assistant", "a1"))
    full = msgs + [("assistant", "a1"), ("user", "u2")]
    assert c.resolve_delta_prompt(list(full)) == ([("user", "u2")], 25)
    # same length as the cache -> not an extension -> cleared
    assert c.resolve_delta_prompt(full[:3]) == (full[:3], 0) and c.items == []
    c.push(10, msgs[0])
    assert c.resolve_delta_prompt([("system", "other"), ("user", "x")]) == ([("system", "other"), ("user", "x")], 0)


def test_http_request_parsing_and_responses():
    a, b = socket.socketpair()
    body = json.dumps({"messages": [{"role": "user", "content": "hi"}], "stream": True}).encode()
    a.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: " +
              str(len(body)).encode() + b"\r\n\r\n" + body)
    req = HttpRequest.read(b)
    assert req.method == "POST" and req.path == "/v1/chat/completions" and req.json["stream"] is True
    req.write_stream_start(); req.write_stream_chunk("data: x\r\n\r\n"); req.write_stream_end(); req.write_not_found()
    got = a.recv(65536)
    assert b"Transfer-Encoding: chunked" in got and b"b\r\ndata: x\r\n\r\n\r\n" in got and b"0000\r\n\r\n" in got and b"404 Not Found" in got
    c = json.loads(chunk_json("hey", False)); s = json.loads(chunk_json(None, True))
    assert c["choices"][0]["delta"] == {"role": "assistant", "content": "hey"} and c["id"] == "cmpl-c0"
    assert s["choices"][0]["finish_reason"] == "stop" and "delta" not in s["choices"][0]
    a.close(); b.close()


@pytest.mark.gpu
def test_dllama_inference_cli(tmp_models):
    m, t = tmp_models["tiny-llama31"]
    r = subprocess.run([os.path.join(ROOT, "dllama"), "inference", "--model", m, "--tokenizer", t, "--buffer-float-type", "q80",
                        "--prompt", "Hello world, the model", "--steps", "24", "--temperature", "0", "--seed", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    out = r.stdout
    assert r.returncode == 0, out
    assert "💡 Arch: Llama" in out and "🔷️ Eval" in out and out.count("🔶 Pred") >= 10
    assert "Evaluation" in out and "Prediction" in out and "tokens/s:" in out
    bad = subprocess.run([os.path.join(ROOT, "dllama"), "inference", "--model", m, "--tokenizer", t, "--steps", "4"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert bad.returncode == 1 and "🚨 Critical error: Prompt is required" in bad.stdout


@pytest.mark.gpu
def test_dllama_perplexity_cli(tmp_models):
    m, t = tmp_models["tiny-llama31"]
    r = subprocess.run([os.path.join(ROOT, "dllama"), "perplexity", "--model", m, "--tokenizer", t, "--prompt", "Hello world and the llama"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "perplexity:" in r.stdout and "bitPerToken:" in r.stdout, r.stdout


@pytest.mark.gpu
def test_dllama_api_server(tmp_models):
    import http.client
    m, t = tmp_models["tiny-llama31"]
    port = 19990 + os.getpid() % 1000
    p = subprocess.Popen([os.path.join(ROOT, "dllama-api"), "--model", m, "--tokenizer", t, "--port", str(port), "--host", "127.0.0.1",
                          "--temperature", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        for _ in range(300):
            try:
                socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
                break
            except OSError:
                time.sleep(0.2)
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
        conn.request("GET", "/v1/models")
        models = json.loads(conn.getresponse().read())
        assert models["object"] == "list" and models["data"][0]["id"] == os.path.basename(m)
        body = json.dumps({"messages": [{"role": "user", "content": "Hello"}], "max_tokens": 8})
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=120)
        conn.request("POST", "/v1/chat/completions", body, {"Content-Type": "application/json"})
        resp = json.loads(conn.getresponse().read())
        assert resp["object"] == "chat.completion" and resp["usage"]["completion_tokens"] <= 8
        assert resp["choices"][0]["message"]["role"] == "assistant"
        # follow-up turn extends the history -> naive cache hit; streaming response
        hist = [{"role": "user", "content": "Hello"}, {"role": "assistant", "content": resp["choices"][0]["message"]["content"]},
                {"role": "user", "content": "and the llama"}]
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=120)
        conn.request("POST", "/v1/chat/completions", json.dumps({"messages": hist, "max_tokens": 6, "stream": True}),
                     {"Content-Type": "application/json"})
        raw = conn.getresponse().read().decode()
        assert "data: [DONE]" in raw and '"finish_reason": "stop"' in raw
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
        conn.request("OPTIONS", "/v1/chat/completions")
        assert conn.getresponse().status == 204
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
        conn.request("GET", "/nope")
        assert conn.getresponse().status == 404
    finally:
        p.terminate()
        out = p.communicate(timeout=30)[0]
    assert "🐤 Found naive cache" in out, out[-2000:]


def test_native_cli_arg_surface():
    """dllama-native (C++): usage text, unknown flags / modes and missing files fail with the reference's message + exit code 1."""
    exe = os.path.join(ROOT, "dllama-native")
    r = subprocess.run([exe, "--help"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "Usage: dllama-native" in r.stdout
    for argv, msg in ((["inference", "--bogus", "1"], "Unknown option"), (["train", "--model", "a", "--tokenizer", "b"], "Unsupported mode"),
                      (["inference", "--tokenizer", "b"], "Model is required"),
                      (["inference", "--model", "/nonexistent.m", "--tokenizer", "b", "--workers", "1.2.3.4:9", "5.6.7.8:9", "--nthreads", "4"],
                       "Cannot open model file")):
        r = subprocess.run([exe] + argv, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 1 and "🚨 Critical error" in r.stdout and msg in r.stdout, r.stdout
