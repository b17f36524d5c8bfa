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


def _device_layout_forward(mf, parts, tokens):
    """Plain PyTorch emulation of what the engine computes from the *device layout* the loader prepares (q|k|v concatenated,
    gate/up rows interleaved, rotary pairs adjacent for every model family, one rope table, per-rank partial sums all-reduced)."""
    h = mf.header
    hd, eps = h.head_dim, h.norm_epsilon
    x_all = []
    n_ranks = len(parts)
    caches = [[([], []) for _ in range(h.n_layers)] for _ in range(n_ranks)]

    def rms(v, w):
        return w * v * torch.rsqrt((v * v).mean(-1, keepdim=True) + eps)

    def rope(v, pos, table):           # v: [heads, hd] with adjacent pairs
        cs = table[pos]                # [hd/2, 2]
        a, b = v[:, 0::2], v[:, 1::2]
        out = torch.empty_like(v)
        out[:, 0::2] = a * cs[:, 0] - b * cs[:, 1]
        out[:, 1::2] = a * cs[:, 1] + b * cs[:, 0]
        return out

    for pos, tok in enumerate(tokens):
        x = parts[0].embedding[tok].float().clone()
        for l in range(h.n_layers):
            partial = torch.zeros_like(x)
            for r, W in enumerate(parts):
                L = W.layers[l]
                nh, nkv = W.n_heads, W.n_kv_heads
                qkv = L.qkv.data.float() @ rms(x, L.norm0)
                q = qkv[: nh * hd].view(nh, hd)
                k = qkv[nh * hd: (nh + nkv) * hd].view(nkv, hd)
                v = qkv[(nh + nkv) * hd:].view(nkv, hd)
                if L.q_norm is not None:
                    q, k = rms(q, L.q_norm), rms(k, L.k_norm)
                q, k = rope(q, pos, W.rope), rope(k, pos, W.rope)
                kc, vc = caches[r][l]
                kc.append(k); vc.append(v)
                K, V = torch.stack(kc, 1), torch.stack(vc, 1)          # [nkv, T, hd]
                z = []
                for i in range(nh):
                    j = i // (nh // nkv)
                    att = torch.softmax((K[j] @ q[i]) / hd ** 0.5, -1)
                    z.append(att @ V[j])
                partial += L.wo.data.float() @ torch.cat(z)
            x = x + partial
            partial = torch.zeros_like(x)
            for W in parts:
                L = W.layers[l]
                gu = L.w13.data.float() @ rms(x, L.norm1)
                partial += L.w2.data.float() @ (torch.nn.functional.silu(gu[0::2]) * gu[1::2])
            x = x + partial
        xn = rms(x, parts[0].final_norm)
        x_all.append(torch.cat([W.wcls.data.float() @ xn for W in parts]))
    return torch.stack(x_all)


@pytest.mark.parametrize("name,n_ranks", [("tiny-llama31-f32", 1), ("tiny-llama31-f32", 4), ("tiny-qwen3-f16", 2), ("kvrep", 4), ("kvrep", 8)])
def test_device_layout_forward_matches_oracle(tmp_models, tmp_path, name, n_ranks):
    """Loader fusions + tensor-parallel slicing (incl. KV-head replication: 4 and 8 ranks over 2 KV heads) against the oracle, on CPU."""
    from distributed_llama_b200.formats import ModelFile, quants
    from distributed_llama_b200.models.config import get_config
    from distributed_llama_b200.models.loader import load_device_weights
    from distributed_llama_b200.models.reference import OracleModel
    from distributed_llama_b200.models.synthetic import write_synthetic_model
    if name == "kvrep":
        path = str(tmp_path / "kvrep.m")
        write_synthetic_model(path, get_config("tiny-llama-kvrep"), weights_float_type=quants.F_32, seed=3)
    else:
        path = tmp_models[name][0]
    mf = ModelFile(path)
    parts = [load_device_weights(mf, r, n_ranks, device="cpu") for r in range(n_ranks)]
    if name == "kvrep":
        assert all(p.n_kv_heads == 1 for p in parts) and parts[0].n_heads == mf.header.n_heads // n_ranks
    toks = [3, 17, 250, 9, 44, 101]
    got = _device_layout_forward(mf, parts, toks)
    ref = OracleModel(mf, act_quant="none").forward(toks, 0)
    assert (got - ref).abs().max().item() < 2e-3
