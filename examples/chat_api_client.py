#!/usr/bin/env python
"""Minimal client for `dllama-api` (counterpart of the reference's examples/chat-api-client.js).

    ./dllama-api --model m.m --tokenizer t.t --port 9990 &
    python examples/chat_api_client.py "What is the capital of France?"
"""
import json
import sys
from urllib.request import Request, urlopen

HOST = "http://127.0.0.1:9990"


def chat(messages, max_tokens=128, stream=False):
    req = Request(HOST + "/v1/chat/completions", data=json.dumps({"messages": messages, "temperature": 0.7, "stop": ["<|eot_id|>"],
                                                                  "max_tokens": max_tokens, "stream": stream}).encode(),
                  headers={"Content-Type": "application/json"})
    with urlopen(req) as resp:
        if not stream:
            return json.loads(resp.read())
        for line in resp:
            line = line.decode().strip()
            if line.startswith("data: ") and line != "data: [DONE]":
                choice = json.loads(line[6:])["choices"][0]
                if "delta" in choice:
                    print(choice["delta"]["content"], end="", flush=True)
        print()


if __name__ == "__main__":
    question = sys.argv[1] if len(sys.argv) > 1 else "Hello, who are you?"
    messages = [{"role": "system", "content": "You are an excellent math teacher."}, {"role": "user", "content": question}]
    out = chat(messages)
    print(out["choices"][0]["message"]["content"])
    print(out["usage"])
