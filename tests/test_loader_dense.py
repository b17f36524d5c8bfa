"""Dense (f32 / f16 / q80) weight files: the loader's tensor-parallel slices must tile the full matrices (CPU only)."""
import pytest
import torch


@pytest.mark.parametrize("name", ["tiny-llama31-f32", "tiny-qwen3-f16", "tiny-llama-q80"])
def test_dense_slices_reassemble(tmp_models, name):
    from distributed_llama_b200.formats import ModelFile
    from distributed_llama_b200.models.loader import load_device_weights
    mf = ModelFile(tmp_models[name][0])
    h = mf.header
    full = load_device_weights(mf, 0, 1, device="cpu")
    parts = [load_device_weights(mf, r, 2, device="cpu") for r in range(2)]
    assert full.weight_kind == (2 if "f16" in name else 1)
    hd = h.head_dim
    for l in range(h.n_layers):
        F = full.layers[l]
        nq, nkv = h.n_heads * hd, h.n_kv_heads * hd
        for lo, hi, per in ((0, nq, parts[0].n_heads * hd), (nq, nq + nkv, parts[0].n_kv_heads * hd),
                            (nq + nkv, nq + 2 * nkv, parts[0].n_kv_heads * hd)):
            off = lo - (0 if lo == 0 else (nq if lo == nq else nq + nkv))
            base = {0: 0, nq: parts[0].n_heads * hd, nq + nkv: parts[0].n_heads * hd + parts[0].n_kv_heads * hd}[lo]
            cat = torch.cat([p.layers[l].qkv.data[base: base + per] for p in parts], 0)
            assert torch.equal(cat, F.qkv.data[lo:hi]) and off == 0
        assert torch.equal(torch.cat([p.layers[l].wo.data for p in parts], 1), F.wo.data)
        assert torch.equal(torch.cat([p.layers[l].w2.data for p in parts], 1), F.w2.data)
        for k in (0, 1):   # gate rows / up rows of the interleaved W1|W3 matrix
            assert torch.equal(torch.cat([p.layers[l].w13.data[k::2] for p in parts], 0), F.w13.data[k::2])
        w1 = torch.from_numpy(mf.tensor_f32(mf.entry("block_matmul_w1", l, 0)))
        assert torch.equal(F.w13.data[0::2].float(), w1.to(F.w13.data.dtype).float())
    assert torch.equal(torch.cat([p.wcls.data for p in parts], 0), full.wcls.data)
    # q80 blocks are dequantised exactly: value = int8 * f16 scale fits f32
    if "q80" in name:
        wq = torch.from_numpy(mf.tensor_f32(mf.entry("block_matmul_wo", 0, 0)))
        assert torch.equal(full.layers[0].wo.data, wq)
