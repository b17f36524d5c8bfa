"""ctypes binding of the sm_100a kernel/runtime library (csrc/cuda -> distributed_llama_b200/_cuda.so).

The library is torch-free: tensors cross the boundary as raw device pointers + the current CUDA stream handle.
If the library is missing on a machine with a GPU we build it; failures are loud (no eager fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from .. import _build

_lib: Optional[C.CDLL] = None

u32, u64, i32, f32, vp = C.c_uint32, C.c_uint64, C.c_int, C.c_float, C.c_void_p


class EngineConfig(C.Structure):
    _fields_ = [(n, u32) for n in ("dim", "nLayers", "nHeads", "nKvHeads", "headDim", "ffDim", "vocab", "seqLen",
                                   "nExperts", "nActiveExperts", "maxBatch", "nSplits", "rank", "nRanks", "numSms")] + \
               [("eps", f32), ("usePdl", u32), ("moeFirstExpert", u32), ("moeNumLocal", u32), ("wType", u32), ("hiddenAct", u32)]


class LayerPtrs(C.Structure):
    _fields_ = [(n, vp) for n in ("qkvQs", "qkvSc", "woQs", "woSc", "w13Qs", "w13Sc", "w2Qs", "w2Sc",
                                  "norm0", "norm1", "qNorm", "kNorm", "moeGate", "kCache", "vCache")]


class GlobalPtrs(C.Structure):
    _fields_ = [("embedding", vp), ("embeddingPeers", vp * 8), ("embRowsPerRank", u32), ("finalNorm", vp), ("wclsQs", vp), ("wclsSc", vp), ("rope", vp), ("vocabFull", u32),
                ("tokens", vp), ("pos", vp), ("x", vp), ("qkv", vp), ("z", vp), ("h", vp), ("logits", vp),
                ("attnPartial", vp), ("attnCounters", vp), ("history", vp), ("expertIdx", vp), ("expertWeight", vp),
                ("routerLogits", vp), ("routerCounter", vp), ("moeScratch", vp), ("moeCounters", vp),
                ("maxPrefill", u32), ("pTokens", vp), ("pPos", vp), ("px", vp), ("pqkv", vp), ("pxn", vp), ("pzb", vp), ("phb", vp),
                ("pAttnPartial", vp), ("pAttnCounters", vp),
                ("argVal", vp), ("argIdx", vp), ("argCounter", vp)]


class CommPtrs(C.Structure):
    _fields_ = [("nRanks", u32), ("rank", u32), ("maxCtas", u32), ("slotStride", u32), ("arena", vp * 8), ("mcArena", vp),
                ("slotsOff", u64), ("flagsOff", u64), ("candValOff", u64),
                ("gatherOff", u64), ("prefillSlotsOff", u64), ("prefillSlotStride", u32)]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_cuda()
    L = C.CDLL(str(path))
    L.dl_repack_q40.argtypes = [vp, u64, u64, u32, u32, vp, vp, u32, u32, u32, vp]
    L.dl_repack_q40.restype = i32
    L.dl_dequant_device_q40.argtypes = [vp, vp, u32, u32, vp, vp]
    L.dl_dequant_device_q40.restype = i32
    L.dl_gemv_q40.argtypes = [i32, i32, i32, vp, vp, u32, u32, vp, u32, vp, f32, vp, u32, i32, vp, i32, i32]
    L.dl_gemv_q40.restype = i32
    L.dl_gemv_dense.argtypes = [i32, i32, i32, i32, vp, u32, u32, vp, u32, vp, f32, vp, u32, i32, vp, i32]
    L.dl_gemv_dense.restype = i32
    L.dl_gemm_q40_tc.argtypes = [i32, vp, vp, u32, u32, vp, u32, u32, vp, u32, i32, vp, i32, i32]
    L.dl_gemm_q40_tc.restype = i32
    L.dl_rmsnorm_bf16.argtypes = [vp, u32, vp, vp, u32, u32, f32, u32, vp]
    L.dl_rmsnorm_bf16.restype = i32
    L.dl_attn_prefill_tc.argtypes = [vp, u32, u32, u32, u32, u32, u32, u32, vp, vp, vp, u32, vp]
    L.dl_attn_prefill_tc.restype = i32
    L.dl_engine_create.argtypes = [C.POINTER(EngineConfig)]
    L.dl_engine_create.restype = vp
    L.dl_engine_destroy.argtypes = [vp]
    L.dl_engine_destroy.restype = None
    L.dl_engine_set_layer.argtypes = [vp, u32, C.POINTER(LayerPtrs)]
    L.dl_engine_set_layer.restype = i32
    L.dl_engine_set_globals.argtypes = [vp, C.POINTER(GlobalPtrs)]
    L.dl_engine_set_globals.restype = i32
    L.dl_engine_enable_mega.argtypes = [vp, i32]
    L.dl_engine_enable_mega.restype = i32
    L.dl_engine_set_vocab_limit.argtypes = [vp, u32]
    L.dl_engine_set_vocab_limit.restype = i32
    L.dl_engine_set_comm.argtypes = [vp, C.POINTER(CommPtrs)]
    L.dl_engine_set_comm.restype = i32
    for name, args in (("dl_comm_alloc", [C.c_size_t, C.POINTER(vp)]), ("dl_comm_free", [vp]), ("dl_comm_ipc_handle", [vp, vp]),
                       ("dl_comm_ipc_open", [vp, C.POINTER(vp)]), ("dl_comm_ipc_close", [vp]), ("dl_comm_memset", [vp, i32, C.c_size_t, vp])):
        getattr(L, name).argtypes = args
        getattr(L, name).restype = i32
    L.dl_vmm_supported.argtypes = [C.POINTER(C.c_int)]
    L.dl_vmm_supported.restype = i32
    L.dl_vmm_create.argtypes = [u32, u32, C.c_size_t, C.c_char_p, i32]
    L.dl_vmm_create.restype = vp
    L.dl_vmm_connect.argtypes = [vp]
    L.dl_vmm_connect.restype = i32
    L.dl_vmm_ptr.argtypes = [vp, u32]
    L.dl_vmm_ptr.restype = vp
    L.dl_vmm_mc_ptr.argtypes = [vp]
    L.dl_vmm_mc_ptr.restype = vp
    L.dl_vmm_bytes.argtypes = [vp]
    L.dl_vmm_bytes.restype = C.c_size_t
    L.dl_vmm_destroy.argtypes = [vp]
    L.dl_vmm_destroy.restype = None
    L.dl_vmm_selftest_kernel.argtypes = [vp, u32, i32, vp]
    L.dl_vmm_selftest_kernel.restype = i32
    L.dl_engine_set_trace.argtypes = [vp, vp, u32]
    L.dl_engine_set_trace.restype = i32
    L.dl_engine_sampler_seed.argtypes = [vp, C.c_uint64]
    L.dl_engine_sampler_seed.restype = i32
    L.dl_engine_sample.argtypes = [vp, f32, f32, vp]
    L.dl_engine_sample.restype = i32
    L.dl_sample_logits.argtypes = [vp, vp, u32, f32, f32, vp, vp, vp]
    L.dl_sample_logits.restype = i32
    L.dl_engine_sync_ns.argtypes = [vp]
    L.dl_engine_sync_ns.restype = C.c_uint64
    L.dl_engine_mega_active.argtypes = [vp]
    L.dl_engine_mega_active.restype = i32
    L.dl_engine_aborted.argtypes = [vp]
    L.dl_engine_aborted.restype = i32
    L.dl_engine_set_trace_all.argtypes = [vp, i32]
    L.dl_engine_set_trace_all.restype = i32
    L.dl_engine_num_sms.argtypes = [vp]
    L.dl_engine_num_sms.restype = u32
    L.dl_engine_forward.argtypes = [vp, i32, i32, i32, vp]
    L.dl_engine_forward.restype = i32
    L.dl_engine_forward_part.argtypes = [vp, i32, u32, i32, vp, vp]
    L.dl_engine_forward_part.restype = i32
    L.dl_engine_prefill.argtypes = [vp, u32, u32, i32, vp]
    L.dl_engine_prefill.restype = i32
    L.dl_engine_capture_decode.argtypes = [vp]
    L.dl_engine_capture_decode.restype = i32
    L.dl_engine_decode_graph.argtypes = [vp, i32, vp]
    L.dl_engine_decode_graph.restype = i32
    _lib = L
    return L


_ERRORS = {
    -12: "prompt chunk larger than the prefill all-reduce slots of the peer arena",
    -30: "tensor-parallel slice too narrow for the fused all-reduce GEMV: the per-rank K of WO / W2 (heads/N * headDim, ffDim/N) "
         "must be a multiple of 128 — use fewer ranks, or DL_COLLECTIVES=nccl",
    -31: "mixture-of-experts up-projection shape not covered by the TMA GEMV",
    -32: "mixture-of-experts down-projection shape not covered by the TMA GEMV (per-rank expert ffDim must be a multiple of 128: use moe_mode=ep)",
    -33: "arg-max under tensor parallelism needs the fused logits kernel",
    -36: "mixture-of-experts prefill shape not covered (dim, ffDim multiples of 256, chunk <= 256 tokens)",
}


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed with code {code}" + (f": {_ERRORS[code]}" if code in _ERRORS else ""))


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


PRO_RMSNORM, PRO_PLAIN = 0, 1
EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU = 0, 1, 2
GEPI_STORE_F32, GEPI_RESIDUAL, GEPI_SWIGLU_BF16, GEPI_STORE_BF16 = 0, 1, 2, 3
