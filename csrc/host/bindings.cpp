// pybind11 surface of the host library (`distributed_llama_b200._host`).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "model_format.hpp"
#include "quants.hpp"
#include "text.hpp"

namespace py = pybind11;
using namespace dl;

using F32Array = py::array_t<float, py::array::c_style | py::array::forcecast>;
using U8Array = py::array_t<uint8_t, py::array::c_style | py::array::forcecast>;

static py::bytes asBytes(const std::string &s) { return py::bytes(s.data(), s.size()); }

PYBIND11_MODULE(_host, m) {
    m.doc() = "distributed_llama_b200 host library: .m/.t formats, q40/q80 codecs, TP slicers, tokenizer, sampler";

    // ---- quants ----
    m.attr("F_32") = (int)F_32;
    m.attr("F_16") = (int)F_16;
    m.attr("F_Q40") = (int)F_Q40;
    m.attr("F_Q80") = (int)F_Q80;
    m.def("parse_float_type", [](const std::string &s) { return (int)parseFloatType(s.c_str()); });
    m.def("float_type_name", [](int t) { return std::string(floatTypeName((FloatType)t)); });
    m.def("tensor_bytes", [](int t, size_t n) { return tensorBytes((FloatType)t, n); });
    m.def("f32_to_f16", &f32ToF16);
    m.def("f16_to_f32", &f16ToF32);
    m.def("quantize", [](int type, F32Array x) {
        const size_t n = (size_t)x.size();
        U8Array out((py::ssize_t)tensorBytes((FloatType)type, n));
        quantize((FloatType)type, x.data(), out.mutable_data(), n);
        return out;
    }, "Pack an f32 array into the byte stream of the given float type");
    m.def("dequantize", [](int type, U8Array in, size_t n) {
        if ((size_t)in.size() < tensorBytes((FloatType)type, n)) throw std::invalid_argument("dequantize: input too short");
        F32Array out((py::ssize_t)n);
        dequantize((FloatType)type, in.data(), out.mutable_data(), n);
        return out;
    });

    // ---- model format ----
    py::class_<ModelHeader>(m, "ModelHeader")
        .def(py::init<>())
        .def_readwrite("header_size", &ModelHeader::headerSize)
        .def_readwrite("file_size", &ModelHeader::fileSize)
        .def_readwrite("version", &ModelHeader::version)
        .def_property("arch", [](const ModelHeader &h) { return (int)h.arch; }, [](ModelHeader &h, int v) { h.arch = (ArchType)v; })
        .def_readwrite("dim", &ModelHeader::dim)
        .def_readwrite("hidden_dim", &ModelHeader::hiddenDim)
        .def_readwrite("moe_hidden_dim", &ModelHeader::moeHiddenDim)
        .def_readwrite("n_layers", &ModelHeader::nLayers)
        .def_readwrite("n_heads", &ModelHeader::nHeads)
        .def_readwrite("n_kv_heads", &ModelHeader::nKvHeads)
        .def_readwrite("head_dim", &ModelHeader::headDim)
        .def_readwrite("n_experts", &ModelHeader::nExperts)
        .def_readwrite("n_active_experts", &ModelHeader::nActiveExperts)
        .def_readwrite("vocab_size", &ModelHeader::vocabSize)
        .def_readwrite("seq_len", &ModelHeader::seqLen)
        .def_readwrite("orig_seq_len", &ModelHeader::origSeqLen)
        .def_readwrite("q_dim", &ModelHeader::qDim)
        .def_readwrite("kv_dim", &ModelHeader::kvDim)
        .def_property("hidden_act", [](const ModelHeader &h) { return (int)h.hiddenAct; }, [](ModelHeader &h, int v) { h.hiddenAct = (HiddenAct)v; })
        .def_property("rope_type", [](const ModelHeader &h) { return (int)h.ropeType; }, [](ModelHeader &h, int v) { h.ropeType = (RopeType)v; })
        .def_readwrite("rope_theta", &ModelHeader::ropeTheta)
        .def_readwrite("rope_scaling_factor", &ModelHeader::ropeScalingFactor)
        .def_readwrite("rope_scaling_low_freq_factor", &ModelHeader::ropeScalingLowFreqFactor)
        .def_readwrite("rope_scaling_high_freq_factor", &ModelHeader::ropeScalingHighFreqFactor)
        .def_readwrite("rope_scaling_orig_max_seq_len", &ModelHeader::ropeScalingOrigMaxSeqLen)
        .def_readwrite("norm_epsilon", &ModelHeader::normEpsilon)
        .def_property("weight_type", [](const ModelHeader &h) { return (int)h.weightType; }, [](ModelHeader &h, int v) { h.weightType = (FloatType)v; })
        .def_property_readonly("ff_dim", &ModelHeader::ffDim)
        .def_property_readonly("qk_norm", &ModelHeader::qkNorm)
        .def_property_readonly("arch_name", [](const ModelHeader &h) { return std::string(archName(h.arch)); })
        .def("describe", &describeModelHeader);
    m.attr("ARCH_LLAMA") = (int)ARCH_LLAMA;
    m.attr("ARCH_QWEN3") = (int)ARCH_QWEN3;
    m.attr("ARCH_QWEN3_MOE") = (int)ARCH_QWEN3_MOE;
    m.attr("ROPE_LLAMA") = (int)ROPE_LLAMA;
    m.attr("ROPE_FALCON") = (int)ROPE_FALCON;
    m.attr("ROPE_LLAMA3_1") = (int)ROPE_LLAMA3_1;
    m.attr("PART_ROOT") = (int)PART_ROOT;
    m.attr("PART_REPLICATE") = (int)PART_REPLICATE;
    m.attr("PART_ROWS") = (int)PART_ROWS;
    m.attr("PART_COLS") = (int)PART_COLS;
    m.def("load_model_header", &loadModelHeader, py::arg("path"), py::arg("max_seq_len") = 0);
    m.def("parse_model_header", [](py::bytes data, uint64_t fileSize, uint32_t maxSeqLen) {
        const std::string s = data;
        return parseModelHeader((const uint8_t *)s.data(), s.size(), fileSize, maxSeqLen);
    }, py::arg("data"), py::arg("file_size"), py::arg("max_seq_len") = 0);
    m.def("build_model_header", [](const std::vector<std::pair<int32_t, int32_t>> &kv) {
        const std::vector<uint8_t> b = buildModelHeader(kv);
        return py::bytes((const char *)b.data(), b.size());
    });

    py::class_<TensorEntry>(m, "TensorEntry")
        .def_readonly("name", &TensorEntry::name)
        .def_readonly("layer", &TensorEntry::layer)
        .def_readonly("expert", &TensorEntry::expert)
        .def_property_readonly("type", [](const TensorEntry &t) { return (int)t.type; })
        .def_readonly("d", &TensorEntry::d)
        .def_readonly("n", &TensorEntry::n)
        .def_readonly("offset", &TensorEntry::offset)
        .def_readonly("n_bytes", &TensorEntry::nBytes)
        .def_property_readonly("part", [](const TensorEntry &t) { return (int)t.part; })
        .def("__repr__", [](const TensorEntry &t) {
            return "<TensorEntry " + t.name + " L" + std::to_string(t.layer) + " E" + std::to_string(t.expert) + " " +
                   floatTypeName(t.type) + " [" + std::to_string(t.d) + "x" + std::to_string(t.n) + "] @" + std::to_string(t.offset) + ">";
        });
    py::class_<SliceRange>(m, "SliceRange")
        .def_readonly("first_row", &SliceRange::firstRow)
        .def_readonly("n_rows", &SliceRange::nRows)
        .def_readonly("row_bytes", &SliceRange::rowBytes)
        .def_readonly("col_byte_offset", &SliceRange::colByteOffset)
        .def_readonly("col_bytes", &SliceRange::colBytes)
        .def_readonly("first_col", &SliceRange::firstCol)
        .def_readonly("n_cols", &SliceRange::nCols)
        .def_property_readonly("total_bytes", &SliceRange::totalBytes);
    m.def("build_tensor_directory", &buildTensorDirectory, py::arg("header"), py::arg("check_file_size") = true);
    m.def("slice_tensor", &sliceTensor);
    m.def("extract_slice", [](const TensorEntry &t, py::buffer file, uint32_t rank, uint32_t nRanks) {
        py::buffer_info info = file.request();
        if ((uint64_t)info.size * (uint64_t)info.itemsize < t.offset + t.nBytes) throw std::invalid_argument("extract_slice: buffer too small");
        const SliceRange s = sliceTensor(t, rank, nRanks);
        U8Array out((py::ssize_t)s.totalBytes());
        extractSlice(t, (const uint8_t *)info.ptr, rank, nRanks, out.mutable_data());
        return out;
    });
    m.def("build_rope_table", [](const ModelHeader &h, uint32_t seqLen) {
        py::array_t<float> out({(py::ssize_t)seqLen, (py::ssize_t)(h.headDim / 2), (py::ssize_t)2});
        buildRopeTable(h, seqLen, out.mutable_data());
        return out;
    });
    m.def("rope_frequency", &ropeFrequency);
    m.def("required_device_bytes", &requiredDeviceBytes);

    // ---- text ----
    py::class_<TokenizerData>(m, "TokenizerData")
        .def(py::init<>())
        .def_property("vocab",
            [](const TokenizerData &d) { py::list l; for (auto &t : d.vocab) l.append(asBytes(t)); return l; },
            [](TokenizerData &d, const std::vector<py::bytes> &v) { d.vocab.clear(); for (auto &b : v) d.vocab.push_back((std::string)b); })
        .def_readwrite("scores", &TokenizerData::scores)
        .def_readwrite("bos_id", &TokenizerData::bosId)
        .def_readwrite("add_bos", &TokenizerData::addBos)
        .def_readwrite("eos_ids", &TokenizerData::eosIds)
        .def_property("chat_template",
            [](const TokenizerData &d) { return asBytes(d.chatTemplate); },
            [](TokenizerData &d, py::bytes b) { d.chatTemplate = (std::string)b; })
        .def_readwrite("max_token_length", &TokenizerData::maxTokenLength);
    m.def("read_tokenizer_file", &readTokenizerFile);
    m.def("write_tokenizer_file", &writeTokenizerFile);

    py::class_<Tokenizer>(m, "Tokenizer")
        .def(py::init<const std::string &>())
        .def(py::init<TokenizerData>())
        .def("encode", [](const Tokenizer &t, py::object text, bool isStart, bool addSpecial) {
            std::string s = py::isinstance<py::bytes>(text) ? (std::string)text.cast<py::bytes>() : text.cast<std::string>();
            return t.encode(s, isStart, addSpecial);
        }, py::arg("text"), py::arg("is_start") = true, py::arg("add_special_tokens") = true)
        .def("decode", [](Tokenizer &t, int32_t token) { return asBytes(t.decode(token)); })
        .def("reset_decoder", &Tokenizer::resetDecoder)
        .def("is_eos", &Tokenizer::isEos)
        .def("describe", &Tokenizer::describe)
        .def("piece", [](const Tokenizer &t, int32_t id) { return asBytes(t.data().vocab.at(id)); })
        .def_property_readonly("vocab_size", &Tokenizer::vocabSize)
        .def_property_readonly("regular_vocab_size", &Tokenizer::regularVocabSize)
        .def_property_readonly("bos_id", [](const Tokenizer &t) { return t.data().bosId; })
        .def_property_readonly("add_bos", [](const Tokenizer &t) { return t.data().addBos; })
        .def_property_readonly("eos_ids", [](const Tokenizer &t) { return t.data().eosIds; })
        .def_property_readonly("chat_template", [](const Tokenizer &t) { return asBytes(t.data().chatTemplate); });

    py::class_<Rng>(m, "Rng")
        .def(py::init<uint64_t>())
        .def("next_u32", &Rng::nextU32)
        .def("next_f32", &Rng::nextF32)
        .def_readwrite("state", &Rng::state);

    py::class_<Sampler>(m, "Sampler")
        .def(py::init<uint32_t, float, float, uint64_t>())
        .def("sample", [](Sampler &s, F32Array logits) {
            if ((uint32_t)logits.size() < s.vocabSize()) throw std::invalid_argument("Sampler: logits shorter than vocab size");
            F32Array copy = logits.attr("copy")().cast<F32Array>();
            return s.sample(copy.mutable_data());
        })
        .def("next_coin", &Sampler::nextCoin)
        .def("set_temperature", &Sampler::setTemperature)
        .def("set_topp", &Sampler::setTopp)
        .def("set_seed", &Sampler::setSeed)
        .def_property_readonly("seed", &Sampler::seed)
        .def_property_readonly("seed_generation", &Sampler::seedGeneration)
        .def_property_readonly("temperature", &Sampler::temperature)
        .def_property_readonly("topp", &Sampler::topp)
        .def_property_readonly("vocab_size", &Sampler::vocabSize);
    m.def("softmax", [](F32Array x) {
        F32Array out = x.attr("copy")().cast<F32Array>();
        softmaxInPlace(out.mutable_data(), (size_t)out.size());
        return out;
    });

    m.attr("TEMPLATE_UNKNOWN") = (int)TEMPLATE_UNKNOWN;
    m.attr("TEMPLATE_LLAMA2") = (int)TEMPLATE_LLAMA2;
    m.attr("TEMPLATE_LLAMA3") = (int)TEMPLATE_LLAMA3;
    m.attr("TEMPLATE_DEEP_SEEK3") = (int)TEMPLATE_DEEP_SEEK3;
    m.attr("TEMPLATE_CHATML") = (int)TEMPLATE_CHATML;
    m.def("parse_chat_template_type", [](const std::string &s) { return (int)parseChatTemplateType(s); });
    py::class_<ChatTemplateGenerator>(m, "ChatTemplateGenerator")
        .def(py::init([](int type, py::bytes tpl, py::bytes eos) {
            return new ChatTemplateGenerator((ChatTemplateType)type, (std::string)tpl, (std::string)eos);
        }))
        .def("generate", [](const ChatTemplateGenerator &g, const std::vector<std::pair<std::string, std::string>> &items, bool gen) {
            std::vector<ChatItem> v;
            for (auto &p : items) v.push_back({p.first, p.second});
            GeneratedChat c = g.generate(v, gen);
            return py::make_tuple(asBytes(c.content), asBytes(c.publicPrompt));
        }, py::arg("items"), py::arg("append_generation_prompt") = true)
        .def_property_readonly("type", [](const ChatTemplateGenerator &g) { return (int)g.type(); })
        .def_property_readonly("type_name", [](const ChatTemplateGenerator &g) { return std::string(chatTemplateTypeName(g.type())); });

    m.attr("MAYBE_EOS") = (int)MAYBE_EOS;
    m.attr("EOS") = (int)EOS;
    m.attr("NOT_EOS") = (int)NOT_EOS;
    py::class_<EosDetector>(m, "EosDetector")
        .def(py::init([](std::vector<int32_t> tokens, const std::vector<py::bytes> &pieces, int padLeft, int padRight) {
            std::vector<std::string> p;
            for (auto &b : pieces) p.push_back((std::string)b);
            return new EosDetector(std::move(tokens), std::move(p), padLeft, padRight);
        }))
        .def("append", [](EosDetector &d, int32_t tokenId, py::object piece) {
            std::string s;
            if (!piece.is_none()) s = py::isinstance<py::bytes>(piece) ? (std::string)piece.cast<py::bytes>() : piece.cast<std::string>();
            return (int)d.append(tokenId, s);
        })
        .def("is_eos", &EosDetector::isEos)
        .def("get_delta", [](const EosDetector &d) { return asBytes(d.getDelta()); })
        .def("reset", &EosDetector::reset);
}
