"""Model-zoo launcher (tools/launch.py; reference launch.py:82-195) against a local HTTP server: multi-part concatenation,
resume with a Range request after a dropped connection, run-script generation."""
import importlib.util
import os
import sys
import threading
from http.server import BaseHTTPRequestHandler, HTTPServer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PARTS = {"/m_aa": os.urandom(70000), "/m_ab": os.urandom(50000), "/t.t": b"tokenizer-bytes"}
STATE = {"dropped": False, "ranges": []}


class _Handler(BaseHTTPRequestHandler):
    def log_message(self, *a):
        pass

    def do_GET(self):
        path = self.path.split("?")[0]
        data = PARTS.get(path)
        if data is None:
            self.send_error(404)
            return
        start = 0
        rng = self.headers.get("Range")
        if rng:
            STATE["ranges"].append((path, rng))
            start = int(rng.split("=")[1].split("-")[0])
        body = data[start:]
        self.send_response(206 if rng else 200)
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        if path == "/m_ab" and not STATE["dropped"]:
            STATE["dropped"] = True               # first attempt: send a prefix, then drop the connection
            self.wfile.write(body[:20000])
            self.wfile.flush()
            self.connection.close()
            return
        self.wfile.write(body)


def test_download_resume_and_run_script(tmp_path, monkeypatch, capsys):
    srv = HTTPServer(("127.0.0.1", 0), _Handler)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    base = f"http://127.0.0.1:{srv.server_address[1]}"
    spec = importlib.util.spec_from_file_location("launch", os.path.join(ROOT, "tools", "launch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "ROOT", str(tmp_path))
    monkeypatch.setattr(mod.time, "sleep", lambda s: None)
    mod.MODELS["unit_test_model"] = dict(model_urls=[base + "/m_aa?download=true", base + "/m_ab?download=true"], tokenizer_url=base + "/t.t",
                                         buffer_type="q80", mode="chat", extra="--max-seq-len 4096")
    monkeypatch.setattr(sys, "argv", ["launch.py", "-y"])
    try:
        assert mod.main(["unit_test_model", "-skip-run", "-y", "--gpus", "4"]) == 0
    finally:
        srv.shutdown()
    mpath = tmp_path / "models" / "unit_test_model" / "dllama_model_unit_test_model.m"
    tpath = tmp_path / "models" / "unit_test_model" / "dllama_tokenizer_unit_test_model.t"
    got = mpath.read_bytes()
    assert len(got) == 120000, (len(got), STATE)
    assert got == PARTS["/m_aa"] + PARTS["/m_ab"] and tpath.read_bytes() == PARTS["/t.t"]
    assert STATE["dropped"] and STATE["ranges"] and STATE["ranges"][0][0] == "/m_ab" and STATE["ranges"][0][1].startswith("bytes=")
    script = (tmp_path / "run_unit_test_model.sh").read_text()
    assert script.startswith("#!/bin/sh") and " chat --model " in script and "--gpus 4" in script and "--buffer-float-type q80" in script
    assert os.access(tmp_path / "run_unit_test_model.sh", os.X_OK)
    # unknown model -> usage + exit code 1
    assert mod.main(["nope"]) == 1 and "Available models" in capsys.readouterr().out
    assert len(mod.MODELS) >= 11 and len(mod.MODELS["llama3_1_405b_instruct_q40"]["model_urls"]) == 56
