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
#include <sys/prctl.h>

#include "native_engine.hpp"
#include "tp_job.hpp"

using namespace dl;

namespace {

struct EngineBackend : InferenceBackend {
    NativeEngine &e;
    TpEngine tp;              // announces every engine call to the worker ranks of a --gpus N job (no-op on one GPU)
    std::vector<float> tmp;
    EngineBackend(NativeEngine &engine, TpJob &job) : e(engine), tp{engine, job}, tmp(engine.header().vocabSize) {}
    uint32_t seqLen() const override { return e.seqLen(); }
    uint32_t vocabSize() const override { return e.header().vocabSize; }
    void setVocabLimit(uint32_t limit) override { e.setVocabLimit(limit); }
    void prefill(const std::vector<int32_t> &tokens, uint32_t pos) override { if (!tokens.empty()) tp.prefill(tokens, pos); }
    int32_t next(int32_t token, uint32_t pos, Sampler &sampler) override {
        if (sampler.temperature() == 0.f) return tp.stepGreedy(token, pos);
        if (e.nRanks() > 1) return tp.stepSampled(token, pos, sampler, sampler.topp());   // device sampler: the logits stay sharded
        const float *logits = e.step(token, pos);
        std::memcpy(tmp.data(), logits, tmp.size() * sizeof(float));
        return sampler.sample(tmp.data());
    }
};

const char *kUsage =
    "Usage: dllama-api-native --model <path> --tokenizer <path> [--host <addr>] [--port <p>] [--max-seq-len <n>]\n"
    "        [--temperature <t>] [--topp <p>] [--seed <s>] [--chat-template {llama2|llama3|deepSeek3|chatml}] [--gpu-index <i>]\n"
    "        [--gpus <n>]   tensor parallel over n GPUs: a supervisor forks one process per GPU (rank 0 serves HTTP) and restarts\n"
    "                       the whole job 3 s after any rank fails\n";

// One attempt of rank 0 of a job (or the only process on one GPU): engine + tokenizer + HTTP server until --max-requests is reached.
int serveOnce(const std::string &model, const std::string &tokenizer, uint32_t maxSeqLen, int gpu, const ApiConfig &cfg, TpJob &job) {
    NativeEngine engine(model, maxSeqLen, gpu, 0, job.nRanks, job.tag, [&job] { job.barrier(); });
    Tokenizer tok(tokenizer);
    if (job.nRanks > 1) job.barrier();   // every rank has its weights
    std::printf("%s%s", tok.describe().c_str(), describeModelHeader(engine.header()).c_str());
    if (job.nRanks > 1)
        std::printf("🔗 %u GPUs (one process each), all-reduce inside the kernels over %s\n", job.nRanks,
                    engine.multicast() ? "the NVSwitch multicast mapping (multimem.st)" : "NVLink peer memory");
    std::printf("💿 Weights loaded\n");
    EngineBackend backend(engine, job);
    ApiServer server(backend, tok, cfg);
    server.serve();
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    std::string model, tokenizer;
    ApiConfig cfg;
    cfg.seed = (uint64_t)std::time(nullptr);
    uint32_t maxSeqLen = 0;
    int gpu = 0;
    uint32_t gpus = 1;
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
            else if (name == "--gpus") gpus = (uint32_t)std::max(1, std::stoi(v));
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
    if (gpus <= 1) {
        while (true) {
            try {
                TpJob job;
                return serveOnce(model, tokenizer, maxSeqLen, gpu, cfg, job);
            } catch (const std::exception &e) {
                std::printf("🚨 Inference error: %s\n🔄 Retrying in 3 seconds...\n", e.what());
                std::fflush(stdout);
                std::this_thread::sleep_for(std::chrono::seconds(3));
            }
        }
    }
    // Tensor parallel: this process is a supervisor without a CUDA context, so it can fork a fresh set of ranks for every attempt.
    // Rank 0 serves HTTP, ranks >= 1 mirror its engine calls (tp_job.hpp). The first process that exits ends the attempt: rank 0
    // returning 0 = --max-requests reached; anything else (a rank crashed, a device-side wait timed out, a peer disappeared) tears
    // the job down and a new one is started after 3 s — the reference's root retry loop plus its worker re-listen loop
    // (src/dllama-api.cpp:616-628, src/app.cpp:306-365).
    while (true) {
        TpJob job;
        job.create(gpus);
        const pid_t supervisor = getpid();
        std::vector<pid_t> kids;
        std::fflush(stdout);
        for (uint32_t r = 0; r < job.nRanks; r++) {
            const pid_t pid = fork();
            if (pid < 0) { std::printf("🚨 Critical error: fork failed\n"); return 1; }
            if (pid == 0) {
                prctl(PR_SET_PDEATHSIG, SIGKILL);
                job.rank = r; job.parentPid = supervisor;
                int rc = 1;
                if (r == 0) {
                    try {
                        rc = serveOnce(model, tokenizer, maxSeqLen, gpu, cfg, job);
                        job.sendExit();
                    } catch (const std::exception &e) {
                        job.fail(e.what());
                        std::printf("🚨 Inference error: %s\n", e.what());
                    }
                } else {
                    std::fclose(stdin);
                    rc = tpWorkerMain(model, tokenizer, maxSeqLen, gpu, job);
                }
                std::fflush(stdout);
                _exit(rc);
            }
            kids.push_back(pid);
        }
        int st = 0;
        const pid_t first = waitpid(-1, &st, 0);
        const bool served = first == kids[0] && WIFEXITED(st) && WEXITSTATUS(st) == 0;
        uint32_t who = 0;
        for (uint32_t r = 0; r < kids.size(); r++) if (kids[r] == first) { who = r; kids[r] = 0; }
        job.sendExit();
        if (!served && kids[0] > 0) kill(kids[0], SIGTERM);   // rank 0 may be blocked in accept()
        tpReap(kids);
        const std::string why = job.ctl && job.ctl->failedRank.load() ? std::string(": ") + job.ctl->error : std::string();
        if (job.ctl) munmap(job.ctl, sizeof(TpControl));
        if (served) return 0;
        std::printf("🚨 Inference error: rank %u ended the job%s\n🔄 Retrying in 3 seconds...\n", who, why.c_str());
        std::fflush(stdout);
        std::this_thread::sleep_for(std::chrono::seconds(3));
    }
}
