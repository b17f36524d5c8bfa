// Python-free runtime of one GPU: maps a `.m` file, uploads + re-tiles the q40 weights, owns the device buffers and drives the
// engine in _cuda.so (persistent decode kernel, tcgen05 prefill, CUDA-graph greedy loop).
//
// Role in the reference: loadLlmNetWeight + NnExecutor/NnCpuDevice set-up + RootLlmInference (src/llm.cpp:614-669,
// src/nn/nn-cpu.cpp:41-148, src/app.cpp:168-208) for the single-node case. The tensor-parallel launcher (one process per GPU,
// torch.distributed bootstrap, peer-memory arena) lives in the Python package; this class is what `dllama-native` and embedders
// that do not want an interpreter link against.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../host/model_format.hpp"

namespace dl {

class NativeEngine {
public:
    // maxSeqLen == 0: the model's context length. Throws std::runtime_error on unsupported files / CUDA errors.
    NativeEngine(const std::string &modelPath, uint32_t maxSeqLen, int device);
    ~NativeEngine();
    NativeEngine(const NativeEngine &) = delete;
    NativeEngine &operator=(const NativeEngine &) = delete;

    const ModelHeader &header() const { return h_; }
    uint32_t seqLen() const { return seqLen_; }
    uint64_t bytesUploaded() const { return bytesUploaded_; }
    bool persistentKernel() const { return mega_; }

    // Feeds prompt tokens at positions [pos, pos + n): chunks of up to 192 tokens on the tensor-core path (dense models),
    // power-of-two batches of up to 8 on the GEMV path otherwise. No logits are produced.
    void prefill(const std::vector<int32_t> &tokens, uint32_t pos);
    // One token; returns the logits row (host memory owned by the engine, valid until the next call).
    const float *step(int32_t token, uint32_t pos);
    // One greedy step entirely on the device (graph replay); returns the next token.
    int32_t stepGreedy(int32_t token, uint32_t pos);
    // n greedy steps back to back on the device without host round trips; returns the generated tokens.
    std::vector<int32_t> decodeGreedy(int32_t firstToken, uint32_t pos, uint32_t nSteps);
    void synchronize();
    // Greedy decoding on the device ignores vocabulary rows >= limit (tokenizer vocabulary smaller than the embedding table).
    void setVocabLimit(uint32_t limit);

private:
    struct Impl;
    void *dev(size_t bytes);            // zero-initialised device allocation owned by the engine
    void release();
    void uploadWeights(const uint8_t *file);
    void setInputs(const int32_t *tokens, uint32_t n, uint32_t pos, bool prefillBuffers);
    void forward(uint32_t n, int logitsMode, bool greedyAdvance);

    ModelHeader h_;
    std::vector<TensorEntry> dir_;
    uint32_t seqLen_ = 0, maxBatch_ = 8, maxPrefill_ = 192, nSplits_ = 1, qkvDim_ = 0;
    bool mega_ = false, graphReady_ = false;
    uint64_t bytesUploaded_ = 0;
    Impl *impl_ = nullptr;
};

}  // namespace dl
