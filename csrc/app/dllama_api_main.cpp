// `dllama-api-native`: the OpenAI-style HTTP server of the reference (src/dllama-api.cpp:602-632) as a native binary on one
// B200 — api_server.cpp (HTTP, JSON, NaiveCache, templates, stop detection) over NativeEngine. Multi-GPU serving: ./dllama-api --gpus N.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <thread>

#include "api_server.hpp"
#include "native_engine.hpp"

using namespace dl;

namespace {

struct EngineBackend : InferenceBackend {
    NativeEngine &e;
    std::vector<float> tmp;
    explicit EngineBackend(NativeEngine &engine) : e(engine), tmp(engine.header().vocabSize) {}
    uint32_t seqLen() const override { return e.seqLen(); }
    uint32_t vocabSize() const override { return e.header().vocabSize; }
    void setVocabLimit(uint32_t limit) override { e.setVocabLimit(limit); }
    void prefill(const std::vector<int32_t> &tokens, uint32_t pos) override { if (!tokens.empty()) e.prefill(tokens, pos); }
    int32_t next(int32_t token, uint32_t pos, Sampler &sampler) override {
        if (sampler.temperature() == 0.f) return e.stepGreedy(token, pos);
        const float *logits = e.step(token, pos);
        std::memcpy(tmp.data(), logits, tmp.size() * sizeof(float));
        return sampler.sample(tmp.data());
    }
};

const char *kUsage =
    "Usage: dllama-api-native --model <path> --tokenizer <path> [--host <addr>] [--port <p>] [--max-seq-len <n>]\n"
    "        [--temperature <t>] [--topp <p>] [--seed <s>] [--chat-template {llama2|llama3|deepSeek3|chatml}] [--gpu-index <i>]\n"
    "Multi-GPU serving: ./dllama-api ... --gpus N\n";

}  // namespace

int main(int argc, char **argv) {
    std::string model, tokenizer;
    ApiConfig cfg;
    cfg.seed = (uint64_t)std::time(nullptr);
    uint32_t maxSeqLen = 0;
    int gpu = 0;
    try {
        for (int i = 1; i < argc;) {
            const std::string name = argv[i];
            if (name == "--help" || name == "-h" || name == "--usage") { std::fprintf(stderr, "%s", kUsage); return 0; }
            if (name == "--workers") { i++; while (i < argc && argv[i][0] != '-') i++; continue; }
            if (i + 1 >= argc) throw std::runtime_error("Missing value for " + name);
            const std::string v = argv[i + 1];
            if (name == "--model") model = v;
            else if (name == "--tokenizer") tokenizer = v;
            else if (name == "--host") cfg.host = v;
            else if (name == "--port") cfg.port = std::stoi(v);
            else if (name == "--max-seq-len") maxSeqLen = (uint32_t)std::stoul(v);
            else if (name == "--temperature") cfg.temperature = std::stof(v);
            else if (name == "--topp") cfg.topp = std::stof(v);
            else if (name == "--seed") cfg.seed = std::stoull(v);
            else if (name == "--chat-template") cfg.chatTemplate = v;
            else if (name == "--gpu-index") gpu = std::max(0, std::stoi(v));
            else if (name == "--max-requests") cfg.maxRequests = std::stoi(v);
            else if (name == "--buffer-float-type" || name == "--nthreads" || name == "--net-turbo" || name == "--gpu-segments") {}
            else throw std::runtime_error("Unknown option: " + name);
            i += 2;
        }
        if (model.empty()) throw std::runtime_error("Model is required");
        if (tokenizer.empty()) throw std::runtime_error("Tokenizer is required");
    } catch (const std::exception &e) {
        std::printf("🚨 Critical error: %s\n", e.what());
        return 1;
    }
    const size_t slash = model.find_last_of("/\\");
    cfg.modelName = slash == std::string::npos ? model : model.substr(slash + 1);
    // the reference retries its whole inference app every 3 s on errors (dllama-api.cpp:616-628)
    while (true) {
        try {
            NativeEngine engine(model, maxSeqLen, gpu);
            Tokenizer tok(tokenizer);
            std::printf("%s%s💿 Weights loaded\n", tok.describe().c_str(), describeModelHeader(engine.header()).c_str());
            EngineBackend backend(engine);
            ApiServer server(backend, tok, cfg);
            server.serve();
            return 0;
        } catch (const std::exception &e) {
            std::printf("🚨 Inference error: %s\n🔄 Retrying in 3 seconds...\n", e.what());
            std::fflush(stdout);
            std::this_thread::sleep_for(std::chrono::seconds(3));
        }
    }
}
