"""Converter parity: a tiny random HF Llama / Qwen3 checkpoint -> convert_hf -> `.m` -> PyTorch oracle must reproduce the
HuggingFace model's logits (validates tensor order, q/k re-ordering, rope conventions, QK-norm, header mapping)."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _hf_model(kind, tmp_path):
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    if kind in ("llama", "llama31"):
        extra = {}
        if kind == "llama31":   # Llama-3.1 frequency rescaling (reference scaleFrequencyLlama3, src/nn/nn-core.cpp:326-340)
            extra = dict(rope_parameters={"rope_type": "llama3", "rope_theta": 500000.0, "factor": 8.0, "low_freq_factor": 1.0,
                                          "high_freq_factor": 4.0, "original_max_position_embeddings": 64})
        else:
            extra = dict(rope_parameters={"rope_type": "default", "rope_theta": 10000.0})
        cfg = transformers.LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                                       num_key_value_heads=2, vocab_size=300, max_position_embeddings=128, rms_norm_eps=1e-5,
                                       tie_word_embeddings=False, hidden_act="silu", **extra)
        model = transformers.LlamaForCausalLM(cfg)
    elif kind == "qwen3_moe":
        cfg = transformers.Qwen3MoeConfig(hidden_size=128, intermediate_size=256, moe_intermediate_size=64, num_hidden_layers=2,
                                          num_attention_heads=4, num_key_value_heads=2, head_dim=64, vocab_size=300,
                                          max_position_embeddings=128, rms_norm_eps=1e-6, rope_theta=1000000.0, tie_word_embeddings=False,
                                          hidden_act="silu", num_experts=8, num_experts_per_tok=2, norm_topk_prob=True,
                                          decoder_sparse_step=1, mlp_only_layers=[])
        model = transformers.Qwen3MoeForCausalLM(cfg)
    else:
        cfg = transformers.Qwen3Config(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                                       num_key_value_heads=2, head_dim=64, vocab_size=300, max_position_embeddings=128,
                                       rms_norm_eps=1e-6, rope_theta=1000000.0, tie_word_embeddings=True, hidden_act="silu")
        model = transformers.Qwen3ForCausalLM(cfg)
    model = model.float().eval()
    d = tmp_path / kind
    model.save_pretrained(str(d), safe_serialization=True)
    return model, str(d)


@pytest.mark.parametrize("kind", ["llama", "llama31", "qwen3", "qwen3_moe"])
def test_convert_hf_matches_transformers(kind, tmp_path):
    from distributed_llama_b200.formats import ModelFile
    from distributed_llama_b200.models.reference import OracleModel
    model, folder = _hf_model(kind, tmp_path)
    conv = _load_tool("convert_hf")
    out = str(tmp_path / f"{kind}.m")
    conv.convert(folder, "f32", out)
    mf = ModelFile(out)
    assert mf.header.dim == 128 and mf.header.n_layers == 2 and mf.header.vocab_size == 300
    if kind == "llama31":
        assert mf.header.rope_type == 2 and mf.header.rope_scaling_factor == 8.0 and mf.header.rope_scaling_orig_max_seq_len == 64
    if kind == "qwen3_moe":
        assert mf.header.n_experts == 8 and mf.header.n_active_experts == 2 and mf.header.moe_hidden_dim == 64
    toks = [5, 17, 250, 9, 44, 101, 7, 299, 12]
    with torch.no_grad():
        ref = model(torch.tensor([toks])).logits[0]
    got = OracleModel(mf).forward(toks, 0)
    assert (got - ref).abs().max().item() < 2e-3
    # q40 file: same geometry, coarser numerics
    out40 = str(tmp_path / f"{kind}_q40.m")
    conv.convert(folder, "q40", out40)
    got40 = OracleModel(ModelFile(out40)).forward(toks, 0)
    assert (got40 - ref).abs().max().item() < 0.35 * ref.abs().max().item()


def test_llama3_tokenizer_converter(tmp_path):
    import base64
    from distributed_llama_b200 import host
    conv = _load_tool("convert_tokenizer_llama3")
    src = tmp_path / "tokenizer.model"
    pieces = [bytes([b]) for b in range(256)] + [b"he", b"ll", b"hell", b"hello", b" w", b" wo"]
    src.write_text("".join(f"{base64.b64encode(p).decode()} {i}\n" for i, p in enumerate(pieces)))
    out = str(tmp_path / "l3.t")
    conv.convert(str(src), out)
    tk = host().Tokenizer(out)
    assert tk.bos_id == len(pieces) and tk.vocab_size == len(pieces) + 256 and tk.piece(tk.eos_ids[1]) == b"<|eot_id|>"
    ids = tk.encode("hello<|eot_id|>", True, True)
    assert [tk.piece(i) for i in ids] == [b"<|begin_of_text|>", b"hello", b"<|eot_id|>"]


def test_convert_meta_checkpoint_two_shards(tmp_path):
    """tools/convert_llama.py: a 2-way model-parallel Meta checkpoint (consolidated.00/01.pth + params.json), derived from a tiny
    HF Llama, must convert to a `.m` whose oracle logits match the HF model — checks the per-tensor concatenation axes."""
    from distributed_llama_b200.formats import ModelFile
    from distributed_llama_b200.models.reference import OracleModel
    model, _ = _hf_model("llama", tmp_path)
    hf_conv = _load_tool("convert_hf")
    sd = model.state_dict()
    n_heads, n_kv, n_layers = 4, 2, 2
    meta = {"tok_embeddings.weight": sd["model.embed_tokens.weight"], "norm.weight": sd["model.norm.weight"], "output.weight": sd["lm_head.weight"]}
    for l in range(n_layers):
        pre, dst = f"model.layers.{l}.", f"layers.{l}."
        meta[dst + "attention.wq.weight"] = hf_conv.to_interleaved(sd[pre + "self_attn.q_proj.weight"], n_heads)   # Meta keeps rotary pairs adjacent
        meta[dst + "attention.wk.weight"] = hf_conv.to_interleaved(sd[pre + "self_attn.k_proj.weight"], n_kv)
        meta[dst + "attention.wv.weight"] = sd[pre + "self_attn.v_proj.weight"]
        meta[dst + "attention.wo.weight"] = sd[pre + "self_attn.o_proj.weight"]
        meta[dst + "feed_forward.w1.weight"] = sd[pre + "mlp.gate_proj.weight"]
        meta[dst + "feed_forward.w2.weight"] = sd[pre + "mlp.down_proj.weight"]
        meta[dst + "feed_forward.w3.weight"] = sd[pre + "mlp.up_proj.weight"]
        meta[dst + "attention_norm.weight"] = sd[pre + "input_layernorm.weight"]
        meta[dst + "ffn_norm.weight"] = sd[pre + "post_attention_layernorm.weight"]
    folder = tmp_path / "meta"
    folder.mkdir()
    row_parallel = ("wo.weight", "w2.weight", "tok_embeddings.weight")
    for r in range(2):
        shard = {}
        for k, v in meta.items():
            v = v.detach().clone()
            if v.dim() == 1:
                shard[k] = v
            else:
                axis = 1 if k.endswith(row_parallel) else 0
                shard[k] = v.chunk(2, dim=axis)[r].contiguous()
        torch.save(shard, str(folder / f"consolidated.{r:02d}.pth"))
    (folder / "params.json").write_text(json.dumps({"dim": 128, "n_layers": n_layers, "n_heads": n_heads, "n_kv_heads": n_kv, "vocab_size": 300,
                                                    "max_seq_len": 128, "norm_eps": 1e-05, "rope_theta": 10000.0}))
    out = str(tmp_path / "meta.m")
    _load_tool("convert_llama").convert(str(folder), "f32", out)
    mf = ModelFile(out)
    assert mf.header.ff_dim == 256 and mf.header.n_kv_heads == 2 and mf.header.rope_theta == 10000.0
    toks = [5, 17, 250, 9, 44, 101, 7, 299, 12]
    with torch.no_grad():
        ref = model(torch.tensor([toks])).logits[0]
    assert (OracleModel(mf).forward(toks, 0) - ref).abs().max().item() < 2e-3


_CORPUS = ["the quick brown fox jumps over the lazy dog " * 3, "hello world, hello there; héllo wörld ✓ emoji 😀 test",
           "numbers 12345 67890 and symbols !@#$%^&*()", "The model generates tokens one at a time."] * 20


def test_hf_fast_tokenizer_converter_matches_tokenizers(tmp_path):
    """tools/convert_tokenizer_hf.py on a byte-level BPE tokenizer trained in-process: vocabulary bytes (GPT-2 alphabet
    un-mapping), scores (= -id), bos/eos resolution from config.json, chat template; the native encoder reproduces the
    `tokenizers` ids and the streaming decoder reproduces the text."""
    tk = pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    from distributed_llama_b200 import host
    d = tmp_path / "hf_tok"
    d.mkdir()
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    tok.train_from_iterator(_CORPUS, trainers.BpeTrainer(vocab_size=420, special_tokens=[], initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    tok.add_special_tokens(["<|begin_of_text|>", "<|end_of_text|>", "<|eot_id|>"])
    tok.save(str(d / "tokenizer.json"))
    bos, eos = tok.token_to_id("<|begin_of_text|>"), tok.token_to_id("<|end_of_text|>")
    template = "{% for m in messages %}<|start_header_id|>{{m['role']}}<|end_header_id|>{% endfor %}"
    (d / "tokenizer_config.json").write_text(json.dumps({"tokenizer_class": "PreTrainedTokenizerFast", "add_bos_token": True, "chat_template": template}))
    (d / "config.json").write_text(json.dumps({"bos_token_id": bos, "eos_token_id": [eos, eos + 1]}))
    out = str(tmp_path / "hf.t")
    _load_tool("convert_tokenizer_hf").convert(str(d), out)
    T = host().Tokenizer(out)
    assert T.vocab_size == tok.get_vocab_size() and list(T.eos_ids) == [eos, eos + 1] and T.chat_template == template.encode()
    for text in ["the quick brown fox", "hello world", "héllo wörld ✓ 😀", "12345 tokens!", "Theseus unknownword zzz", " leading space"]:
        ids = tok.encode(text).ids
        assert list(T.encode(text, False, False)) == ids, text
        T.reset_decoder()
        assert b"".join(T.decode(i) for i in ids).decode("utf-8") == text
    # with isStart the BOS id is prepended; special tokens are matched literally when asked to
    assert list(T.encode("hello", True, False))[0] == bos
    assert eos in list(T.encode("hello<|end_of_text|>", False, True))


def test_sentencepiece_tokenizer_converters(tmp_path):
    """SentencePiece route of convert_tokenizer_hf.py and convert_tokenizer_llama2.py: `▁` -> space, <0xNN> byte pieces, scores kept."""
    spm = pytest.importorskip("sentencepiece")
    from distributed_llama_b200 import host
    d = tmp_path / "sp_tok"
    d.mkdir()
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(_CORPUS))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(d / "tokenizer"), vocab_size=400, model_type="bpe", byte_fallback=True,
                                   character_coverage=1.0, bos_id=1, eos_id=2, unk_id=0, pad_id=-1, minloglevel=2)
    sp = spm.SentencePieceProcessor(model_file=str(d / "tokenizer.model"))
    (d / "tokenizer_config.json").write_text(json.dumps({"tokenizer_class": "LlamaTokenizer", "add_bos_token": True}))
    out_hf, out_l2 = str(tmp_path / "sp_hf.t"), str(tmp_path / "sp_l2.t")
    _load_tool("convert_tokenizer_hf").convert(str(d), out_hf)
    _load_tool("convert_tokenizer_llama2").convert(str(d), out_l2)
    H = host()
    for path in (out_hf, out_l2):
        T = H.Tokenizer(path)
        assert T.vocab_size == sp.vocab_size() and list(T.eos_ids) == [2]
        for text in ["the quick brown fox", "hello wörld ✓"]:
            ids = sp.encode(text)
            T.reset_decoder()
            got = b"".join(T.decode(i) for i in ids).decode("utf-8")
            assert got.strip() == text            # sentencepiece's dummy prefix becomes a leading space
    # scores survive the conversion (they drive the merge order of the runtime encoder)
    data = H.read_tokenizer_file(out_hf) if hasattr(H, "read_tokenizer_file") else None
    if data is not None:
        assert abs(data.scores[10] - sp.get_score(10)) < 1e-6
