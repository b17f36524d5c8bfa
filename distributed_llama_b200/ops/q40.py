"""Device-side q40 matrices: repack from `.m` bytes and the standalone GEMV entry point (tests, microbenchmarks)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import cuda_lib as cl


@dataclass
class DeviceQ40:
    """q40 matrix in the device layout of csrc/cuda/common.cuh: qs u32 [d][n/8], scales f16 [d][n/32]."""
    qs: torch.Tensor
    scales: torch.Tensor
    d: int
    n: int

    @staticmethod
    def empty(d: int, n: int, device="cuda", lead: int = 1) -> "DeviceQ40":
        qs = torch.empty((lead * d, n // 8), dtype=torch.int32, device=device)
        sc = torch.empty((lead * d, n // 32), dtype=torch.float16, device=device)
        return DeviceQ40(qs, sc, d, n)

    def to_f32(self) -> torch.Tensor:
        rows = self.qs.shape[0]
        out = torch.empty((rows, self.n), dtype=torch.float32, device=self.qs.device)
        cl.check(cl.lib().dl_dequant_device_q40(self.qs.data_ptr(), self.scales.data_ptr(), rows, self.n // 32,
                                                 out.data_ptr(), cl.stream_ptr()), "dequant_device_q40")
        return out


@dataclass
class DeviceDense:
    """Row-major f32 / f16 matrix [lead*d][n] for `.m` files that are not q40 (csrc/cuda/gemv_dense.cu). Exposes the same
    `qs` / `scales` attributes as DeviceQ40 so the engine's pointer tables do not care which kind they carry."""
    data: torch.Tensor
    d: int
    n: int

    @property
    def qs(self) -> torch.Tensor:
        return self.data

    @property
    def scales(self):
        return None

    @property
    def wtype(self) -> int:
        return 1 if self.data.dtype == torch.float32 else 2

    def to_f32(self) -> torch.Tensor:
        return self.data.float()


def gemv_dense(w: DeviceDense, x: torch.Tensor, *, pro: int, epi: int, out: torch.Tensor, norm_w: Optional[torch.Tensor] = None,
               eps: float = 1e-5, num_sms: int = 0, pdl: bool = False) -> torch.Tensor:
    """Dense-weight counterpart of gemv_q40 (f32 activations, no q80 round trip)."""
    nb = x.shape[0]
    if num_sms == 0:
        num_sms = torch.cuda.get_device_properties(x.device).multi_processor_count
    cl.check(cl.lib().dl_gemv_dense(w.wtype, pro, epi, nb, w.data.data_ptr(), w.d, w.n, x.data_ptr(), x.stride(0),
                                    norm_w.data_ptr() if norm_w is not None else None, eps, out.data_ptr(), out.stride(0),
                                    num_sms, cl.stream_ptr(), 1 if pdl else 0), "gemv_dense")
    return out


def repack_q40(raw: torch.Tensor, rows: int, n_cols: int, dst: DeviceQ40, *, src_row_pitch: Optional[int] = None,
               src_col_byte_offset: int = 0, dst_row_stride: int = 1, dst_row_offset: int = 0, head_dim: int = 0) -> None:
    """raw: uint8 CUDA tensor holding `rows` source rows of 18-byte blocks (pitch defaults to n_cols/32*18)."""
    assert raw.is_cuda and raw.dtype == torch.uint8
    bpr = n_cols // 32
    pitch = src_row_pitch if src_row_pitch is not None else bpr * 18
    assert dst.n == n_cols
    cl.check(cl.lib().dl_repack_q40(raw.data_ptr(), pitch, src_col_byte_offset, rows, bpr, dst.qs.data_ptr(),
                                    dst.scales.data_ptr(), dst_row_stride, dst_row_offset, head_dim, cl.stream_ptr()),
             "repack_q40")


def gemv_q40(w: DeviceQ40, x: torch.Tensor, *, pro: int, epi: int, out: torch.Tensor, norm_w: Optional[torch.Tensor] = None,
             eps: float = 1e-5, num_sms: int = 0, pdl: bool = False, impl: str = "auto") -> torch.Tensor:
    """x: f32 [nb, n]; out: f32 [nb, d] (STORE / RESIDUAL in place) or [nb, d/2] (SWIGLU)."""
    nb = x.shape[0]
    if num_sms == 0:
        num_sms = torch.cuda.get_device_properties(x.device).multi_processor_count
    cl.check(cl.lib().dl_gemv_q40(pro, epi, nb, w.qs.data_ptr(), w.scales.data_ptr(), w.d, w.n, x.data_ptr(), x.stride(0),
                                  norm_w.data_ptr() if norm_w is not None else None, eps, out.data_ptr(), out.stride(0),
                                  num_sms, cl.stream_ptr(), 1 if pdl else 0, {"auto": 0, "ldg": 1, "tma": 2}[impl]), "gemv_q40")
    return out


def gemm_q40_tc(w: DeviceQ40, act: torch.Tensor, *, epi: int, out: torch.Tensor, num_sms: int = 0, pdl: bool = False,
                variant: str = "auto") -> torch.Tensor:
    """tcgen05 prefill GEMM. act: bf16 [T, n] (T <= 256); out: [T, d] f32 (STORE/RESIDUAL), bf16 [T, d/2] (SWIGLU) or bf16 [T, d]."""
    assert act.dtype == torch.bfloat16 and act.is_cuda and act.stride(1) == 1
    if num_sms == 0:
        num_sms = torch.cuda.get_device_properties(act.device).multi_processor_count
    cl.check(cl.lib().dl_gemm_q40_tc(epi, w.qs.data_ptr(), w.scales.data_ptr(), w.d, w.n, act.data_ptr(), act.stride(0), act.shape[0],
                                     out.data_ptr(), out.stride(0), num_sms, cl.stream_ptr(), 1 if pdl else 0,
                                     {"auto": 0, "ldg": 1, "tma": 2}[variant]), "gemm_q40_tc")
    return out


def rmsnorm_bf16(x: torch.Tensor, w: Optional[torch.Tensor], eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: f32 [T, n] -> bf16 [T, n] = w * x * rsqrt(mean(x^2) + eps); w=None converts only."""
    T, n = x.shape
    if out is None:
        out = torch.empty(T, n, dtype=torch.bfloat16, device=x.device)
    cl.check(cl.lib().dl_rmsnorm_bf16(x.data_ptr(), x.stride(0), w.data_ptr() if w is not None else None, out.data_ptr(), out.stride(0),
                                      n, eps, T, cl.stream_ptr()), "rmsnorm_bf16")
    return out
