// Process plumbing of a native tensor-parallel job (`dllama-native --gpus N`, `dllama-api-native --gpus N`): one process per GPU
// of one NVSwitch box, forked before CUDA is initialised. Role in the reference: root + `dllama worker` processes and their
// LlmControlPacket stream (src/app.hpp:46-49, src/app.cpp:168-208,306-365) — here the packets travel through a control block in
// an anonymous shared mapping instead of TCP sockets, and a process that dies is noticed by its peers (waitpid / getppid /
// failedRank) instead of hanging them.
#pragma once
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <csignal>
#include <cstdio>
#include <ctime>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../host/text.hpp"
#include "native_engine.hpp"

namespace dl {

enum : uint32_t { TP_OP_PREFILL = 1, TP_OP_STEP_GREEDY = 2, TP_OP_STEP_SAMPLED = 3, TP_OP_EXIT = 4, TP_OP_SEED = 5 };
constexpr uint32_t kTpMaxRanks = 8, kTpCtrlTokens = 256;

struct TpControl {
    std::atomic<uint32_t> seq;                // bumped by the root for every command
    std::atomic<uint32_t> ack[kTpMaxRanks];   // last command completed by rank r
    std::atomic<uint32_t> arrived;            // bootstrap barrier: monotonic arrival counter
    std::atomic<int32_t> failedRank;          // rank + 1 of the first process that failed, 0 = none
    uint32_t op, n, pos;
    float temperature, topp;
    uint64_t seed;
    int32_t tokens[kTpCtrlTokens];
    char error[240];
};

struct TpJob {   // one process's view of the job
    TpControl *ctl = nullptr;
    uint32_t rank = 0, nRanks = 1, barriers = 0;
    pid_t parentPid = 0;          // the process whose disappearance ends this one (root or supervisor)
    bool reapChildren = false;    // this process forked the other ranks (CLI root): a reaped child = a dead worker
    std::string tag;              // names the bootstrap sockets of the peer-memory arena

    static double nowMs() {
        using namespace std::chrono;
        return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
    }
    // Control block + arena tag; call before forking.
    void create(uint32_t ranks) {
        nRanks = ranks < 1 ? 1 : (ranks > kTpMaxRanks ? kTpMaxRanks : ranks);
        if (nRanks == 1) return;
        void *p = mmap(nullptr, sizeof(TpControl), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) throw std::runtime_error("cannot map the control block");
        ctl = new (p) TpControl();
        static int jobs = 0;
        tag = "dllama-" + std::to_string((long)getpid()) + "-" + std::to_string((long long)std::time(nullptr)) + "-" + std::to_string(jobs++);
    }
    void fail(const std::string &what) {
        if (!ctl) return;
        int32_t none = 0;
        if (ctl->failedRank.compare_exchange_strong(none, (int32_t)rank + 1)) std::snprintf(ctl->error, sizeof(ctl->error), "%s", what.c_str());
    }
    void checkPeers() {
        const int32_t f = ctl->failedRank.load();
        if (f != 0 && f != (int32_t)rank + 1) throw std::runtime_error("rank " + std::to_string(f - 1) + " failed: " + std::string(ctl->error));
        if (reapChildren) {
            int st = 0;
            if (waitpid(-1, &st, WNOHANG) > 0) throw std::runtime_error("a worker process exited unexpectedly");
        } else if (parentPid && getppid() != parentPid) {
            throw std::runtime_error("the parent process is gone");
        }
    }
    // spins while the job is active, sleeps between probes after 1 s of waiting (the reference's "turbo off")
    template <typename Pred> void waitFor(Pred done) {
        const double t0 = nowMs();
        for (uint32_t i = 0; !done(); i++) {
            if ((i & 1023u) == 1023u) {
                checkPeers();
                if (nowMs() - t0 > 1000.0) usleep(200);
            }
        }
    }
    void barrier() {   // all ranks, bootstrap only
        barriers++;
        ctl->arrived.fetch_add(1);
        const uint32_t target = barriers * nRanks;
        waitFor([&] { return ctl->arrived.load() >= target; });
    }
    // root: publish a command once every worker has finished the previous one
    void issue(uint32_t op, const int32_t *tokens, uint32_t n, uint32_t pos, float temperature = 0.f, float topp = 0.f, uint64_t seed = 0) {
        if (nRanks == 1) return;
        const uint32_t cur = ctl->seq.load();
        waitFor([&] { for (uint32_t r = 1; r < nRanks; r++) if (ctl->ack[r].load() != cur) return false; return true; });
        ctl->op = op; ctl->n = n; ctl->pos = pos; ctl->temperature = temperature; ctl->topp = topp; ctl->seed = seed;
        for (uint32_t i = 0; i < n && i < kTpCtrlTokens; i++) ctl->tokens[i] = tokens[i];
        ctl->seq.store(cur + 1, std::memory_order_release);
    }
    void sendExit() {
        if (!ctl) return;
        ctl->op = TP_OP_EXIT;
        ctl->seq.store(ctl->seq.load() + 1, std::memory_order_release);
    }
};

// Root-side engine calls of a job: every call is announced to the workers first, then executed locally (the kernels of all ranks
// meet inside their all-reduces).
struct TpEngine {
    NativeEngine &engine;
    TpJob &job;
    uint64_t seedGen = ~0ull;
    void prefill(const std::vector<int32_t> &tokens, uint32_t pos) {
        for (size_t i = 0; i < tokens.size(); i += 192) {   // one control packet per tensor-core chunk
            const uint32_t n = (uint32_t)std::min<size_t>(192, tokens.size() - i);
            job.issue(TP_OP_PREFILL, tokens.data() + i, n, pos + (uint32_t)i);
            engine.prefill(std::vector<int32_t>(tokens.begin() + i, tokens.begin() + i + n), pos + (uint32_t)i);
        }
    }
    int32_t stepGreedy(int32_t token, uint32_t pos) {
        job.issue(TP_OP_STEP_GREEDY, &token, 1, pos);
        return engine.stepGreedy(token, pos);
    }
    // temperature / top-p on the device (the logits stay sharded); `sampler` only supplies the parameters and the seed
    int32_t stepSampled(int32_t token, uint32_t pos, Sampler &sampler, float topp) {
        if (sampler.seedGeneration() != seedGen) {
            seedGen = sampler.seedGeneration();
            job.issue(TP_OP_SEED, nullptr, 0, 0, 0.f, 0.f, sampler.seed());
            engine.seedSampler(sampler.seed());
        }
        job.issue(TP_OP_STEP_SAMPLED, &token, 1, pos, sampler.temperature(), topp);
        return engine.stepSampled(token, pos, sampler.temperature(), topp);
    }
};

// ranks >= 1: mirror the root's engine calls until TP_OP_EXIT (reference: runWorkerApp, src/app.cpp:306-365). Returns the exit code.
inline int tpWorkerMain(const std::string &model, const std::string &tokenizer, uint32_t maxSeqLen, int gpuIndex, TpJob &job) {
    try {
        NativeEngine engine(model, maxSeqLen, gpuIndex + (int)job.rank, job.rank, job.nRanks, job.tag, [&job] { job.barrier(); });
        {
            Tokenizer tok(tokenizer);      // only for the vocabulary limit of the greedy arg-max (must match the root)
            engine.setVocabLimit(tok.vocabSize());
        }
        job.barrier();                     // "weights loaded" on every rank
        uint32_t mine = 0;
        while (true) {
            job.waitFor([&] { return job.ctl->seq.load(std::memory_order_acquire) != mine; });
            mine++;
            const TpControl &c = *job.ctl;
            if (c.op == TP_OP_EXIT) break;
            if (c.op == TP_OP_PREFILL) { engine.prefill(std::vector<int32_t>(c.tokens, c.tokens + c.n), c.pos); engine.synchronize(); }
            else if (c.op == TP_OP_STEP_GREEDY) engine.stepGreedy(c.tokens[0], c.pos);
            else if (c.op == TP_OP_STEP_SAMPLED) engine.stepSampled(c.tokens[0], c.pos, c.temperature, c.topp);
            else if (c.op == TP_OP_SEED) engine.seedSampler(c.seed);
            job.ctl->ack[job.rank].store(mine, std::memory_order_release);
        }
        job.ctl->ack[job.rank].store(mine, std::memory_order_release);
        return 0;
    } catch (const std::exception &e) {
        job.fail(e.what());
        return 1;
    }
}

// Ends a set of child processes: 3 s of grace after the exit command, then SIGKILL.
inline void tpReap(std::vector<pid_t> &children) {
    for (pid_t c : children) {
        if (c <= 0) continue;
        bool gone = false;
        for (int i = 0; i < 3000 && !gone; i++) {
            int st = 0;
            if (waitpid(c, &st, WNOHANG) != 0) gone = true; else usleep(1000);
        }
        if (!gone) { kill(c, SIGKILL); waitpid(c, nullptr, 0); }
    }
    children.clear();
}

}  // namespace dl
