// Native `dllama-api`: blocking HTTP/1.1 server with the reference's routes (src/dllama-api.cpp:43-632, src/api-types.hpp):
//   POST /v1/chat/completions (JSON body; JSON response or SSE-style chunked stream), GET /v1/models, OPTIONS * (CORS), else 404.
// One request at a time over one KV sequence with NaiveCache prefix re-use. The inference side is an interface so the HTTP /
// JSON / template / stop-detector plumbing can be exercised without a GPU (tests/native/api_stub_main.cpp).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../host/text.hpp"
#include "json.hpp"

namespace dl {

struct InferenceBackend {
    virtual ~InferenceBackend() {}
    virtual uint32_t seqLen() const = 0;
    virtual uint32_t vocabSize() const = 0;
    virtual void setVocabLimit(uint32_t) {}   // greedy decoding never returns ids >= limit (tokenizer vocabulary)
    virtual void prefill(const std::vector<int32_t> &tokens, uint32_t pos) = 0;
    virtual int32_t next(int32_t token, uint32_t pos, Sampler &sampler) = 0;   // greedy on the device when temperature == 0
};

struct HttpRequest {
    int fd = -1;
    std::string method = "UNKNOWN", path, body;
    std::map<std::string, std::string> headers;   // lower-cased names
    JsonValue json;
    bool hasJson = false;

    static HttpRequest read(int fd);   // throws std::runtime_error on socket / protocol errors
    void writeCors() const;
    void writeNotFound() const;
    void writeJson(const std::string &text) const;
    void writeStreamStart() const;
    void writeStreamChunk(const std::string &data) const;
    void writeStreamEnd() const;
private:
    void send(const std::string &data) const;
};

// (endPos, message) per chat turn; re-used only if *all* cached messages are an exact prefix of the new history.
class NaiveCache {
public:
    void push(uint32_t endPos, const ChatItem &msg) { items_.push_back({endPos, msg}); }
    void clear() { items_.clear(); }
    size_t size() const { return items_.size(); }
    // Returns the position generation restarts from and trims `messages` to the part that still has to be evaluated.
    uint32_t resolveDeltaPrompt(std::vector<ChatItem> &messages);
private:
    struct Item { uint32_t endPos; ChatItem msg; };
    std::vector<Item> items_;
};

std::string chunkJson(const std::string *delta, bool stop);   // one streamed `chat.completion` chunk

struct ApiConfig {
    std::string host = "0.0.0.0", modelName = "model", chatTemplate;
    int port = 9990;
    float temperature = 0.8f, topp = 0.9f;
    uint64_t seed = 0;
    int maxRequests = 0;   // 0 = serve forever
};

class ApiServer {
public:
    ApiServer(InferenceBackend &backend, Tokenizer &tokenizer, const ApiConfig &cfg);
    void complete(HttpRequest &req);
    void models(HttpRequest &req) const;
    void serve();          // accept loop
private:
    InferenceBackend &backend_;
    Tokenizer &tok_;
    ApiConfig cfg_;
    Sampler sampler_;
    NaiveCache cache_;
    std::vector<std::string> stops_;
    ChatTemplateGenerator gen_;
    EosDetector det_;
};

}  // namespace dl
