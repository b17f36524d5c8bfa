"""CPU tests: q40/q80 codecs, `.m` header/directory/slicing, rope tables. Mirrors the reference's nn-cpu-ops-test
(quant round-trip bounds, src/nn/nn-cpu-ops-test.cpp:77-104), converter/writer-test.py (golden hex) and adds the
slice/split tests the reference lacks."""
import io
import os

import numpy as np
import pytest
import torch

from distributed_llama_b200 import host
from distributed_llama_b200.formats import ModelFile, quants, write_model_header
from distributed_llama_b200.models.config import get_config, PRESETS

GOLDEN_Q40 = ('7e346345a692b89665b2c5790537876e598aaa366d988876a898b8d788a98868ce660c66f6b3a88cba5ce9a871987ba9cc5bcaaa760c1eb556a4455b747b6b9504968828ef2a8d7c1db5c6be3764799e66db6d8e76463126a30e4333cad7a4f645947c6cf97f9de086d468c8d535a6ba7dc799d3d0c657bab6799468cad8bb349eb7d7635c7c798998696bb38e4085a9eb34444ba96a7f8ba7b2b42d746a96cf9660aeb4499d8708ad5c7b9a7558947645f3bbb6b0346a656887ad9a86059baac5c596ab781c703569bb8a4356a4bd58cb78736ba09759bb0e34a6274e827b957d7a67dfa86846955660d234b6d9d78a378094a8a8708a7a774ae92f8a36b8c999a9b77a7d958a69747c807963941235379886d69a7a8767b3a6a4ac71999760')


def test_q40_golden_matches_reference_writer():
    torch.manual_seed(1)
    t = torch.randn(32, 16)
    assert quants.quantize_q40(t.numpy()).tobytes().hex() == GOLDEN_Q40
    assert bytes(np.asarray(host().quantize(quants.F_Q40, t.numpy().reshape(-1)))).hex() == GOLDEN_Q40


@pytest.mark.parametrize("t,bound", [(quants.F_Q40, 0.13), (quants.F_Q80, 0.01), (quants.F_16, 1e-3)])
def test_quant_roundtrip_bounds_and_native_parity(t, bound):
    rng = np.random.default_rng(12345)
    x = (rng.random(32 * 200, dtype=np.float32) * 2 - 1).astype(np.float32)
    a = np.asarray(host().quantize(t, x))
    b = quants.quantize(t, x)
    assert (a == b).all()
    ya = np.asarray(host().dequantize(t, a, x.size))
    yb = quants.dequantize(t, b, x.size)
    assert np.array_equal(ya, yb)
    assert np.abs(ya - x).max() < bound


def test_f16_conversion_all_bit_patterns():
    H = host()
    allh = np.arange(65536, dtype=np.uint16)
    ref = allh.view(np.float16).astype(np.float32)
    got = np.array([H.f16_to_f32(int(v)) for v in allh[::7]], dtype=np.float32)
    assert np.array_equal(np.nan_to_num(got, nan=7.0), np.nan_to_num(ref[::7], nan=7.0))
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(2000).astype(np.float32) * s for s in (1e-6, 1e-3, 1.0, 1e3, 7e4)])
    want = xs.astype(np.float16).view(np.uint16)
    got = np.array([H.f32_to_f16(float(v)) for v in xs], dtype=np.uint16)
    assert np.array_equal(got, want)


def test_header_roundtrip_and_describe(tmp_path):
    cfg = get_config("llama-3.1-8b")
    p = tmp_path / "h.m"
    with open(p, "wb") as f:
        n = write_model_header(f, cfg.header_params(quants.F_Q40))
    h = host().load_model_header(str(p), 4096)
    assert h.header_size == n and h.dim == 4096 and h.n_kv_heads == 8 and h.head_dim == 128 and h.q_dim == 4096
    assert h.seq_len == 4096 and h.orig_seq_len == 131072 and h.rope_type == host().ROPE_LLAMA3_1
    assert abs(h.norm_epsilon - 1e-5) < 1e-9 and h.weight_type == quants.F_Q40
    text = h.describe()
    assert "💡 Arch: Llama" in text and "💡 OrigSeqLen: 131072" in text and "RopeScaling: f=8.0, l=1.0, h=4.0, o=8192" in text


def test_header_rejects_bad_files(tmp_path):
    p = tmp_path / "bad.m"
    p.write_bytes(b"\x00" * 64)
    with pytest.raises(RuntimeError):
        host().load_model_header(str(p), 0)
    p.write_bytes((0xABCD00).to_bytes(4, "little") + b"\x00" * 60)
    with pytest.raises(RuntimeError, match="Old model format"):
        host().load_model_header(str(p), 0)


@pytest.mark.parametrize("name", ["llama-3.1-8b", "qwen3-14b", "qwen3-30b-a3b", "llama-3.3-70b"])
def test_directory_sizes_match_published_file_sizes(name):
    """File sizes the reference README lists for its q40 model zoo (README.md:28-40): 8B 6.32 GB, 14B 10.9 GB, 30B-A3B 17.0 GB, 70B 40 GB."""
    H = host()
    cfg = get_config(name)
    hdr = H.parse_model_header(H.build_model_header([(k, v) for k, v in
                               ((__import__('distributed_llama_b200.formats.model_file', fromlist=['HEADER_KEYS']).HEADER_KEYS[kk], vv)
                                for kk, vv in cfg.header_params(quants.F_Q40).items())]), 0, 0)
    d = H.build_tensor_directory(hdr, False)
    total = d[-1].offset + d[-1].n_bytes
    expect = {"llama-3.1-8b": 6.32e9, "qwen3-14b": 10.9e9, "qwen3-30b-a3b": 17.0e9, "llama-3.3-70b": 40e9}[name]
    # the README mixes GB and GiB
    assert min(abs(total - expect) / expect, abs(total / 2**30 * 1e9 - expect) / expect) < 0.06


def test_slices_cover_tensor_exactly(tmp_models):
    mf = ModelFile(tmp_models["tiny-llama"][0])
    H = host()
    for name in ("block_matmul_q", "block_matmul_wo", "block_matmul_w1", "block_matmul_w2", "final_matmul_logits"):
        e = mf.entry(name, 0)
        full = mf.tensor_f32(e)
        for n in (1, 2):
            parts = [mf.slice_f32(e, r, n) for r in range(n)]
            axis = 0 if e.part == H.PART_ROWS else 1
            assert np.array_equal(np.concatenate(parts, axis=axis), full), (name, n)
    # a column slice must cover whole quant blocks
    e = mf.entry("block_matmul_wo", 0)
    s = H.slice_tensor(e, 1, 2)
    assert s.col_byte_offset == (e.n // 2) // 32 * 18 and s.col_bytes == s.col_byte_offset


def test_rope_table_matches_reference_formula():
    H = host()
    cfg = get_config("tiny-llama31")
    hdr = H.parse_model_header(H.build_model_header([(__import__('distributed_llama_b200.formats.model_file', fromlist=['HEADER_KEYS']).HEADER_KEYS[k], v)
                                                     for k, v in cfg.header_params(quants.F_Q40).items()]), 0, 0)
    tab = np.asarray(H.build_rope_table(hdr, 64))
    hd = hdr.head_dim
    # llama 3.1 scaling (reference nn-core.cpp:326-340)
    for j in (0, 5, hd // 2 - 1):
        freq = 1.0 / (hdr.rope_theta ** (2 * j / hd))
        wave = 2 * np.pi / freq
        orig, f, lo, hi = 8192.0, 8.0, 1.0, 4.0
        if wave < orig / hi:
            pass
        elif wave > orig / lo:
            freq /= f
        else:
            sm = (orig / wave - lo) / (hi - lo)
            freq = (1 - sm) * freq / f + sm * freq
        np.testing.assert_allclose(tab[17, j, 0], np.cos(17 * freq), rtol=0, atol=2e-5)
        np.testing.assert_allclose(tab[17, j, 1], np.sin(17 * freq), rtol=0, atol=2e-5)
