// `dllama-native {inference|perplexity|chat}` — the reference CLI (src/dllama.cpp:13-285, flag table src/app.cpp:24-135)
// as one native binary: C++ tokenizer / sampler / chat templates (csrc/host), native engine driver (native_engine.cpp), CUDA
// kernels from _cuda.so. No interpreter in the process.
//
// `--gpus N` (tensor parallel over the GPUs of one NVSwitch box) is the reference's root + `dllama worker` processes
// (src/dllama.cpp:260-285, src/app.cpp:306-365) without sockets: the root forks one worker process per extra GPU *before* CUDA is
// initialised; every process loads its slice of the model, joins the peer-memory arena (CUDA VMM handles passed over unix sockets,
// csrc/cuda/comm_vmm.cu) and the workers then mirror the root's engine calls, which reach them through a control block in an
// anonymous shared mapping (the reference's LlmControlPacket, src/app.hpp:46-49). A worker that dies is noticed by the root
// (waitpid) and the root's death by the workers (getppid): nobody waits forever. The all-reduces themselves run inside the
// kernels over NVLink.
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <memory>
#include <new>
#include <csignal>
#include <stdexcept>
#include <string>
#include <vector>

#include "../host/text.hpp"
#include "native_engine.hpp"
#include "tp_job.hpp"

using namespace dl;

namespace {

struct Args {
    std::string mode, model, tokenizer, prompt, chatTemplate, bufferFloatType = "q80";
    bool hasPrompt = false, help = false;
    uint32_t steps = 0, maxSeqLen = 0;
    float temperature = 0.8f, topp = 0.9f;
    uint64_t seed = (uint64_t)std::time(nullptr);
    int gpuIndex = 0;
    uint32_t gpus = 1;
};

const char *kUsage =
    "Usage: dllama-native {inference|perplexity|chat} --model <path> --tokenizer <path>\n"
    "        [--prompt <text>] [--steps <n>] [--temperature <t>] [--topp <p>] [--seed <s>] [--max-seq-len <n>]\n"
    "        [--chat-template {llama2|llama3|deepSeek3|chatml}] [--buffer-float-type q80] [--gpu-index <i>]\n"
    "        [--gpus <n>]   tensor parallel over GPUs gpu-index .. gpu-index + n - 1 of this machine (one process per GPU)\n"
    "        (accepted for drop-in compatibility and ignored: --nthreads --net-turbo --gpu-segments --workers --host --port)\n";

Args parse(int argc, char **argv) {
    Args a;
    int i = 1;
    if (i < argc && argv[i][0] != '-') a.mode = argv[i++];
    while (i < argc) {
        const std::string name = argv[i];
        if (name == "--help" || name == "-h" || name == "--usage") { a.help = true; return a; }
        if (name == "--workers") {   // variadic until the next flag
            i++;
            while (i < argc && argv[i][0] != '-') i++;
            continue;
        }
        if (i + 1 >= argc) throw std::runtime_error("Missing value for " + name);
        const std::string value = argv[i + 1];
        if (name == "--model") a.model = value;
        else if (name == "--tokenizer") a.tokenizer = value;
        else if (name == "--prompt") { a.prompt = value; a.hasPrompt = true; }
        else if (name == "--steps") a.steps = (uint32_t)std::stoul(value);
        else if (name == "--temperature") a.temperature = std::stof(value);
        else if (name == "--topp") a.topp = std::stof(value);
        else if (name == "--seed") a.seed = std::stoull(value);
        else if (name == "--max-seq-len") a.maxSeqLen = (uint32_t)std::stoul(value);
        else if (name == "--chat-template") a.chatTemplate = value;
        else if (name == "--buffer-float-type") a.bufferFloatType = value;
        else if (name == "--gpu-index") a.gpuIndex = std::max(0, std::stoi(value));
        else if (name == "--gpus") a.gpus = (uint32_t)std::max(1, std::stoi(value));
        else if (name == "--nthreads" || name == "--net-turbo" || name == "--gpu-segments" || name == "--host" || name == "--port") {}
        else throw std::runtime_error("Unknown option: " + name);
        i += 2;
    }
    return a;
}

double nowMs() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct App {
    Args args;
    TpJob &job;
    NativeEngine engine;
    Tokenizer tokenizer;
    Sampler sampler;
    TpEngine tp;
    App(const Args &a, TpJob &j)
        : args(a), job(j), engine(a.model, a.maxSeqLen, a.gpuIndex + (int)j.rank, j.rank, j.nRanks, j.tag, [&j] { j.barrier(); }),
          tokenizer(a.tokenizer),
          sampler(std::min<uint32_t>(tokenizer.vocabSize(), engine.header().vocabSize), a.temperature, a.topp, a.seed), tp{engine, j} {
        // sampling ranges over the tokenizer's vocabulary (reference src/app.cpp:243-246); padded embedding rows never win
        engine.setVocabLimit(tokenizer.vocabSize());
    }

    void prefill(const std::vector<int32_t> &tokens, uint32_t pos) { tp.prefill(tokens, pos); }

    int32_t next(int32_t token, uint32_t pos) {
        if (sampler.temperature() == 0.f) return tp.stepGreedy(token, pos);
        if (job.nRanks > 1) return tp.stepSampled(token, pos, sampler, args.topp);   // logits stay sharded on the devices
        const float *logits = engine.step(token, pos);
        std::vector<float> tmp(logits, logits + std::min<uint32_t>(tokenizer.vocabSize(), engine.header().vocabSize));
        return sampler.sample(tmp.data());
    }
};

void inference(App &app) {
    const Args &a = app.args;
    if (!a.hasPrompt) throw std::runtime_error("Prompt is required");
    if (a.steps == 0) throw std::runtime_error("Number of steps is required");
    const ModelHeader &h = app.engine.header();
    const std::vector<int32_t> tokens = app.tokenizer.encode(a.prompt, true, true);
    const uint32_t nIn = (uint32_t)tokens.size();
    if (nIn > h.seqLen) throw std::runtime_error("The number of prompt tokens is greater than the sequence length");
    if (nIn > a.steps) throw std::runtime_error("The number of prompt tokens is greater than the number of steps");
    std::printf("%s\n", a.prompt.c_str());
    double evalMs = 0, predMs = 0, syncMs = 0;
    uint32_t pos = 0;
    const uint32_t chunk = 192;   // tokens per tensor-core prefill launch (the reference feeds 32 per forward)
    while (pos + 1 < nIn) {
        const uint32_t n = std::min(chunk, nIn - 1 - pos);
        const double t0 = nowMs();
        app.prefill(std::vector<int32_t>(tokens.begin() + pos, tokens.begin() + pos + n), pos);
        app.engine.synchronize();
        const double dt = nowMs() - t0;
        evalMs += dt;
        uint64_t sent = 0, recv = 0;
        app.engine.linkBytes(n, sent, recv);
        std::printf("🔷️ Eval%5d ms Sync%5d ms | Sent%6d kB Recv%6d kB | (%u tokens)\n", (int)dt, 0, (int)(sent / 1024), (int)(recv / 1024), n);
        pos += n;
    }
    std::fflush(stdout);
    int32_t token = tokens[pos];
    app.tokenizer.resetDecoder();
    const uint32_t maxPos = std::min(h.seqLen, a.steps);
    uint32_t nPred = 0;
    while (pos < maxPos) {
        const double t0 = nowMs();
        const uint64_t sync0 = app.engine.syncNs();
        token = app.next(token, pos);
        const double dt = nowMs() - t0;
        predMs += dt;
        syncMs += (double)(app.engine.syncNs() - sync0) * 1e-6;
        uint64_t sent = 0, recv = 0;
        app.engine.linkBytes(1, sent, recv);
        const std::string piece = app.tokenizer.decode(token);
        std::printf("🔶 Pred%5d ms Sync%5d ms | Sent%6d kB Recv%6d kB | %s\n", (int)dt, (int)((double)(app.engine.syncNs() - sync0) * 1e-6),
                    (int)(sent / 1024), (int)(recv / 1024), piece.empty() ? "~" : piece.c_str());
        std::fflush(stdout);
        pos++;
        nPred++;
    }
    const uint32_t nEval = nIn - 1;
    std::printf("\nEvaluation\n   nBatches: %u\n    nTokens: %u\n", chunk, nEval);
    if (nEval > 0 && evalMs > 0) std::printf("   tokens/s: %3.2f (%3.2f ms/tok)\n", nEval * 1000.0 / evalMs, evalMs / nEval);
    std::printf("Prediction\n    nTokens: %u\n", nPred);
    if (nPred > 0 && predMs > 0) std::printf("   tokens/s: %3.2f (%3.2f ms/tok)\n", nPred * 1000.0 / predMs, predMs / nPred);
    if (app.job.nRanks > 1 && nPred > 0) std::printf("   syncTime: %3.3f ms/tok waiting for peers inside the fused all-reduces\n", syncMs / nPred);
}

void perplexity(App &app) {
    const Args &a = app.args;
    if (app.job.nRanks > 1) throw std::runtime_error("perplexity needs the full logits on the host: run it on one GPU (or through ./dllama perplexity --gpus N)");
    if (!a.hasPrompt) throw std::runtime_error("Prompt is required");
    const ModelHeader &h = app.engine.header();
    const std::vector<int32_t> tokens = app.tokenizer.encode(a.prompt, true, true);
    const size_t n = tokens.size();
    if (n > h.seqLen) throw std::runtime_error("The number of prompt tokens is greater than the sequence length");
    std::printf("Evaluating %zu tokens...\n", n);
    double total = 0;
    std::vector<float> probs(h.vocabSize);
    for (size_t pos = 0; pos + 1 < n; pos++) {
        const float *logits = app.engine.step(tokens[pos], (uint32_t)pos);
        std::memcpy(probs.data(), logits, (size_t)h.vocabSize * sizeof(float));
        softmaxInPlace(probs.data(), h.vocabSize);
        const double p = probs[tokens[pos + 1]];
        total += std::log(std::max(p, 1e-30));
        std::printf("%5zu / %zu, prob=%f\n", pos + 1, n - 1, p);
    }
    const double avg = total / (double)std::max<size_t>(1, n - 1);
    std::printf("\nResults\n   perplexity: %f (lower = better)\n   avgLogProb: %f\n   bitPerToken: %f\n", std::exp(-avg), avg, -avg / std::log(2.0));
}

bool readLine(const char *prompt, std::string &out) {
    std::printf("%s", prompt);
    std::fflush(stdout);
    return (bool)std::getline(std::cin, out);
}

void chat(App &app) {
    const ModelHeader &h = app.engine.header();
    const TokenizerData &td = app.tokenizer.data();
    std::vector<std::string> stops;
    for (int32_t id : td.eosIds) stops.push_back(td.vocab[(size_t)id]);
    size_t maxStop = 0;
    for (const std::string &s : stops) maxStop = std::max(maxStop, s.size());
    const ChatTemplateType type = app.args.chatTemplate.empty() ? TEMPLATE_UNKNOWN : parseChatTemplateType(app.args.chatTemplate);
    ChatTemplateGenerator gen(type, td.chatTemplate, stops.empty() ? std::string() : stops[0]);
    std::printf("⭐ Chat template: %s\n", chatTemplateTypeName(gen.type()));
    for (const std::string &s : stops) std::printf("🛑 Stop: %s\n", s.c_str());
    EosDetector det(td.eosIds, stops, (int)maxStop, (int)maxStop);

    std::string sysPrompt;
    readLine("💻 System prompt (optional): ", sysPrompt);
    std::vector<ChatItem> items;
    if (!sysPrompt.empty()) items.push_back({"system", sysPrompt});
    uint32_t pos = 0;
    while (pos < h.seqLen) {
        std::string user;
        while (user.empty())
            if (!readLine("\n👱 User\n> ", user)) return;
        items.push_back({"user", user});
        const GeneratedChat g = gen.generate(items, true);
        const std::vector<int32_t> tokens = app.tokenizer.encode(g.content, pos == 0, true);
        const uint32_t end = std::min<uint32_t>(h.seqLen, pos + (uint32_t)tokens.size() - 1);
        const uint32_t n = end - pos;
        app.prefill(std::vector<int32_t>(tokens.begin(), tokens.begin() + n), pos);
        pos += n;
        int32_t token = n < tokens.size() ? tokens[n] : tokens.back();
        app.tokenizer.resetDecoder();
        det.reset();
        std::printf("\n🤖 Assistant\n");
        if (!g.publicPrompt.empty()) std::printf("%s", g.publicPrompt.c_str());
        while (pos < h.seqLen) {
            token = app.next(token, pos);
            const std::string piece = app.tokenizer.decode(token);
            const EosDetectorResult kind = det.append(token, piece);
            if (kind == NOT_EOS || kind == EOS) {
                const std::string delta = det.getDelta();
                if (!delta.empty()) { std::printf("%s", delta.c_str()); std::fflush(stdout); }
                det.reset();
            }
            pos++;
            if (kind == EOS) break;
        }
        items.clear();
    }
    std::printf("(end of context)\n");
}

}  // namespace

int main(int argc, char **argv) {
    try {
        const Args a = parse(argc, argv);
        if (a.help || a.mode.empty()) { std::printf("%s", kUsage); return 0; }
        if (a.mode == "worker")
            throw std::runtime_error("workers are forked by the root: dllama-native <mode> --gpus N");
        if (a.mode != "inference" && a.mode != "perplexity" && a.mode != "chat") throw std::runtime_error("Unsupported mode");
        if (a.model.empty()) throw std::runtime_error("Model is required");
        if (a.tokenizer.empty()) throw std::runtime_error("Tokenizer is required");
        if (a.bufferFloatType != "q80") throw std::runtime_error("This version supports only Q40 weights with Q80 sync type");
        // ---- tensor-parallel job: fork the workers before CUDA exists in this process ----
        TpJob job;
        job.create(a.gpus);
        job.reapChildren = true;
        std::vector<pid_t> children;
        if (job.nRanks > 1) {
            const pid_t rootPid = getpid();
            std::fflush(stdout);
            for (uint32_t r = 1; r < job.nRanks; r++) {
                const pid_t pid = fork();
                if (pid < 0) throw std::runtime_error("fork failed");
                if (pid == 0) {
                    prctl(PR_SET_PDEATHSIG, SIGKILL);
                    job.rank = r; job.reapChildren = false; job.parentPid = rootPid;
                    std::fclose(stdin);
                    _exit(tpWorkerMain(a.model, a.tokenizer, a.maxSeqLen, a.gpuIndex, job));
                }
                children.push_back(pid);
            }
        }
        struct Reaper {   // root: tell the workers to leave and collect them, whatever happens
            TpJob &job; std::vector<pid_t> &children;
            ~Reaper() { if (job.nRanks > 1 && job.rank == 0) { job.sendExit(); tpReap(children); } }
        } reaper{job, children};
        std::unique_ptr<App> appPtr;
        try {
            appPtr.reset(new App(a, job));
            if (job.nRanks > 1) job.barrier();   // every rank has its weights
        } catch (const std::exception &e) {
            job.fail(e.what());
            throw;
        }
        App &app = *appPtr;
        const ModelHeader &h = app.engine.header();
        if (app.tokenizer.vocabSize() != h.vocabSize)
            std::printf("Tokenizer vocab size (%u) does not match the model vocab size (%u)\n", app.tokenizer.vocabSize(), h.vocabSize);
        std::printf("%s", app.tokenizer.describe().c_str());
        std::printf("%s", describeModelHeader(h).c_str());
        std::printf("📀 RequiredMemory: %llu MB\n", (unsigned long long)(requiredDeviceBytes(h, job.nRanks, 2) / (1024 * 1024)));
        if (job.nRanks > 1)
            std::printf("🔗 %u GPUs (one process each), all-reduce inside the kernels over %s\n", job.nRanks,
                        app.engine.multicast() ? "the NVSwitch multicast mapping (multimem.st)" : "NVLink peer memory");
        std::printf("🧠 GPU %d: sm_100a kernels, %s decode; %.2f GB of weights uploaded%s\n", a.gpuIndex,
                    app.engine.persistentKernel() ? "persistent-kernel" : "multi-kernel", app.engine.bytesUploaded() / 1e9,
                    job.nRanks > 1 ? " by this rank" : "");
        std::printf("💿 Weights loaded\n");
        if (a.mode == "inference") inference(app);
        else if (a.mode == "perplexity") perplexity(app);
        else chat(app);
        return 0;
    } catch (const std::exception &e) {
        std::printf("🚨 Critical error: %s\n", e.what());
        return 1;
    }
}
