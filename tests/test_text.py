"""CPU tests of the native text stack: tokenizer file round trip, encoder/decoder, sampler, chat templates, stop
detector. The EosDetector / template cases follow the reference's tokenizer-test.cpp:122-303 scenarios."""
import numpy as np
import pytest

from distributed_llama_b200 import host


@pytest.fixture(scope="module")
def tok(tmp_path_factory):
    from distributed_llama_b200.models.synthetic import write_synthetic_tokenizer
    p = str(tmp_path_factory.mktemp("tok") / "t.t")
    write_synthetic_tokenizer(p, 1024)
    return host().Tokenizer(p)


def test_tokenizer_file_roundtrip(tmp_path):
    H = host()
    d = H.TokenizerData()
    d.vocab = [bytes([i]) for i in range(1, 200)] + [b"ab", b"abc", b"<s>", b"</s>"]
    d.scores = [-float(i) for i in range(len(d.vocab))]
    d.bos_id = 201
    d.add_bos = True
    d.eos_ids = [202]
    d.chat_template = b"{{ '[INST]' }}"
    p = str(tmp_path / "x.t")
    H.write_tokenizer_file(p, d)
    r = H.read_tokenizer_file(p)
    assert r.vocab == d.vocab and r.bos_id == 201 and r.eos_ids == [202] and bytes(r.chat_template) == b"{{ '[INST]' }}"
    assert r.max_token_length == 4 and r.add_bos


def test_tokenizer_file_legacy_chat_stop_key(tmp_path):
    """Old `.t` files carry key 8 (CHAT_STOP = length of an ignored stop string stored right after the header). The bytes are
    skipped once the whole header has been read (reference src/tokenizer.cpp:68-86), not in the middle of the key/value table."""
    import struct
    vocab = [b"a", b"b", b"ab", b"<s>", b"</s>"]
    stop, template = b"<|stop|>", b"{{ '[INST]' }}"
    kv = [(3, 3), (0, 1), (8, len(stop)), (1, len(vocab)), (2, 4), (7, len(template)), (9, 1), (10, 1)]   # key 8 sits mid-table
    blob = struct.pack("<ii", 0x567124, 8 + 8 * len(kv)) + b"".join(struct.pack("<ii", k, v) for k, v in kv)
    blob += stop + template + struct.pack("<i", 4)
    for i, t in enumerate(vocab):
        blob += struct.pack("<fI", -float(i), len(t)) + t
    p = tmp_path / "legacy.t"
    p.write_bytes(blob)
    r = host().read_tokenizer_file(str(p))
    assert r.vocab == vocab and r.bos_id == 3 and r.eos_ids == [4] and bytes(r.chat_template) == template and r.add_bos


def test_encode_decode_roundtrip(tok):
    text = "Hello world, the model is a llama! ünïcödé ✓ 😃"
    ids = tok.encode(text, True, True)
    assert ids[0] == tok.bos_id
    tok.reset_decoder()
    out = b"".join(tok.decode(i) for i in ids)
    assert out.decode("utf-8") == text
    # merges happened (fewer tokens than bytes) and special tokens match literally
    assert len(ids) < len(text.encode()) + 1
    ids2 = tok.encode("<|start_header_id|>user<|end_header_id|>hi<|eot_id|>", False, True)
    pieces = [tok.piece(i) for i in ids2]
    assert pieces[0] == b"<|start_header_id|>" and pieces[-1] == b"<|eot_id|>" and tok.is_eos(ids2[-1])
    # without special-token matching the same text is plain bytes
    ids3 = tok.encode("<|eot_id|>", False, False)
    assert all(i < tok.regular_vocab_size for i in ids3)


def test_streaming_utf8_decoder_recovers(tok):
    # emoji split over 4 single-byte tokens: nothing is emitted until the sequence is complete
    e = "😃".encode()
    ids = [tok.encode(bytes([b]), False, False)[0] for b in e]
    tok.reset_decoder()
    outs = [tok.decode(i) for i in ids]
    assert outs[:3] == [b"", b"", b""] and outs[3] == e
    # a broken sequence is replaced by U+FFFD and decoding continues
    tok.reset_decoder()
    a = tok.decode(ids[0])
    b = tok.decode(tok.encode("x", False, False)[0])
    assert a == b"" and b == "�".encode() + b"x"


def test_rng_and_sampler_contract():
    H = host()
    r = H.Rng(12345)
    # xorshift* reference values computed from the published algorithm
    s = 12345
    vals = []
    for _ in range(3):
        s ^= s >> 12
        s ^= (s << 25) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 27
        vals.append(((s * 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF) >> 32)
    assert [r.next_u32() for _ in range(3)] == vals
    logits = np.array([0.1, 3.0, -1.0, 2.9, 0.0], dtype=np.float32)
    assert H.Sampler(5, 0.0, 0.9, 1).sample(logits) == 1
    # multinomial (topp outside (0,1)) follows the CDF of softmax(logits / T)
    smp = H.Sampler(5, 1.0, 1.0, 42)
    counts = np.zeros(5)
    for _ in range(4000):
        counts[smp.sample(logits)] += 1
    p = np.exp(logits - logits.max()); p /= p.sum()
    assert np.abs(counts / 4000 - p).max() < 0.03
    # nucleus sampling never returns tokens outside the top-p set
    smp = H.Sampler(5, 1.0, 0.5, 7)
    assert set(smp.sample(logits) for _ in range(500)) <= {1, 3}
    # same seed -> same sequence
    a = H.Sampler(5, 0.8, 0.9, 99); b = H.Sampler(5, 0.8, 0.9, 99)
    assert [a.sample(logits) for _ in range(50)] == [b.sample(logits) for _ in range(50)]


def test_chat_template_detection_and_rendering():
    H = host()
    mk = lambda tpl, eos=b"<eos>": H.ChatTemplateGenerator(H.TEMPLATE_UNKNOWN, tpl, eos)
    assert mk(b"{% set x %}[INST] foo").type == H.TEMPLATE_LLAMA2
    assert mk(b"<|start_header_id|>...").type == H.TEMPLATE_LLAMA3
    assert mk("...<｜Assistant｜>...".encode()).type == H.TEMPLATE_DEEP_SEEK3
    assert mk(b"<|im_start|>...").type == H.TEMPLATE_CHATML
    with pytest.raises(RuntimeError):
        mk(b"nothing known")
    with pytest.raises(RuntimeError):
        mk(b"")
    g = H.ChatTemplateGenerator(H.TEMPLATE_LLAMA3, b"", b"<|eot_id|>")
    content, public = g.generate([("system", "be brief"), ("user", "hi")], True)
    assert content == (b"<|start_header_id|>system<|end_header_id|>\n\nbe brief<|eot_id|>"
                       b"<|start_header_id|>user<|end_header_id|>\n\nhi<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n")
    assert public == b""
    g = H.ChatTemplateGenerator(H.TEMPLATE_LLAMA2, b"", b"</s>")
    content, _ = g.generate([("system", "S"), ("user", "U"), ("assistant", "A"), ("user", "U2")], True)
    assert content == b"[INST] <<SYS>>\nS\n<</SYS>>\n\nU [/INST]</s>A</s>[INST] U2 [/INST]</s>"
    g = H.ChatTemplateGenerator(H.TEMPLATE_DEEP_SEEK3, b"", b"<eos>")
    content, public = g.generate([("user", "q")], True)
    assert content == "<｜User｜>q<｜Assistant｜><think>\n".encode() and public == b"<think>\n"
    g = H.ChatTemplateGenerator(H.TEMPLATE_CHATML, b"", b"<|im_end|>")
    content, _ = g.generate([("system", "S"), ("user", "U")], True)
    assert content == b"<|im_start|>system\nS<|im_end|>\n<|im_start|>user\nU<|im_end|>\n<|im_start|>assistant\n"


def test_eos_detector_with_padding():
    H = host()
    EOS, MAYBE, NOT = H.EOS, H.MAYBE_EOS, H.NOT_EOS
    d = H.EosDetector([2, 3], [b"<eos>", b"<stop>"], 1, 1)
    # plain text passes through
    assert d.append(10, "x") == NOT and d.get_delta() == b"x"
    d.reset()
    # stop string arriving in pieces, with one byte of left padding
    assert d.append(10, "<") == MAYBE
    assert d.append(11, "eo") == MAYBE
    assert d.append(12, "s>") == EOS and d.get_delta() == b""
    d.reset()
    assert d.append(10, " <") == MAYBE
    assert d.append(11, "stop") == MAYBE
    assert d.append(12, "> ") == EOS and d.get_delta() == b" "
    d.reset()
    # looks like a stop, then diverges: buffered text is released
    assert d.append(10, "<eo") == MAYBE
    assert d.append(11, "lia") == NOT and d.get_delta() == b"<eolia"
    d.reset()
    # EOS by token id wins regardless of text
    assert d.append(10, "abc") == NOT
    d.reset()
    assert d.append(2, "") == EOS and d.get_delta() == b""
    d.reset()
    assert d.append(10, "xy") == NOT
    assert d.append(3, "z") == EOS and d.get_delta() == b"xyz"


def test_eos_detector_without_padding():
    H = host()
    d = H.EosDetector([2], [b"<eos>"], 0, 0)
    assert d.append(10, " <") == H.NOT_EOS
    d.reset()
    assert d.append(10, "<eos") == H.MAYBE_EOS
    assert d.append(11, ">") == H.EOS and d.get_delta() == b""
    d.reset()
    assert d.append(10, "<eos> ") == H.NOT_EOS
