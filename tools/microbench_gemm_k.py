"""Prefill GEMM time vs K at fixed d (slope = per-k-block cost, intercept = launch + prologue + epilogue)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_llama_b200 import ops
from tools.microbench_gemm import bench

if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 6144
    for n in (256, 512, 1024, 2048, 4096, 8192 - 256):
        r = bench(f"k{n}", d, n, T, ops.GEPI_STORE_F32, "tma")
        print(json.dumps({k: r[k] for k in ("kernel", "d", "n", "T", "us")}), flush=True)
