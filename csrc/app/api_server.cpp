#include "api_server.hpp"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <stdexcept>

namespace dl {

// ---- HTTP plumbing ----------------------------------------------------------------------------------------------------

void HttpRequest::send(const std::string &data) const {
    size_t off = 0;
    while (off < data.size()) {
        const ssize_t n = ::send(fd, data.data() + off, data.size() - off, MSG_NOSIGNAL);
        if (n <= 0) throw std::runtime_error("Error while writing to socket");
        off += (size_t)n;
    }
}

HttpRequest HttpRequest::read(int fd) {
    HttpRequest req;
    req.fd = fd;
    std::string data;
    char buf[8192];
    size_t headEnd = std::string::npos, sepLen = 0;
    while (true) {
        const size_t a = data.find("\r\n\r\n"), b = data.find("\n\n");
        if (a != std::string::npos && (b == std::string::npos || a <= b)) { headEnd = a; sepLen = 4; break; }
        if (b != std::string::npos) { headEnd = b; sepLen = 2; break; }
        const ssize_t n = ::recv(fd, buf, sizeof(buf), 0);
        if (n <= 0) throw std::runtime_error("Error while reading headers from socket");
        data.append(buf, (size_t)n);
        if (data.size() > (1u << 20)) throw std::runtime_error("Request header too large");
    }
    const std::string head = data.substr(0, headEnd);
    std::string rest = data.substr(headEnd + sepLen);
    size_t lineStart = 0;
    bool first = true;
    while (lineStart <= head.size()) {
        size_t lineEnd = head.find('\n', lineStart);
        if (lineEnd == std::string::npos) lineEnd = head.size();
        std::string line = head.substr(lineStart, lineEnd - lineStart);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (first) {
            const size_t s1 = line.find(' ');
            if (s1 != std::string::npos) {
                size_t s2 = line.find(' ', s1 + 1);
                if (s2 == std::string::npos) s2 = line.size();
                req.method = line.substr(0, s1);
                for (char &c : req.method) c = (char)std::toupper((unsigned char)c);
                req.path = line.substr(s1 + 1, s2 - s1 - 1);
            }
            first = false;
        } else {
            const size_t colon = line.find(':');
            if (colon != std::string::npos) {
                std::string k = line.substr(0, colon), v = line.substr(colon + 1);
                for (char &c : k) c = (char)std::tolower((unsigned char)c);
                const auto trim = [](std::string &s) {
                    while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
                    size_t i = 0;
                    while (i < s.size() && std::isspace((unsigned char)s[i])) i++;
                    s.erase(0, i);
                };
                trim(k); trim(v);
                req.headers[k] = v;
            }
        }
        lineStart = lineEnd + 1;
    }
    size_t length = 0;
    const auto it = req.headers.find("content-length");
    if (it != req.headers.end() && !it->second.empty()) length = (size_t)std::strtoull(it->second.c_str(), nullptr, 10);
    if (length > (64u << 20)) throw std::runtime_error("Request body too large");
    if (length > 0 && rest.size() > length) throw std::runtime_error("Received more body data than Content-Length header said");
    while (rest.size() < length) {
        const ssize_t n = ::recv(fd, buf, std::min(sizeof(buf), length - rest.size()), 0);
        if (n <= 0) throw std::runtime_error("Error while reading body from socket");
        rest.append(buf, (size_t)n);
    }
    req.body = rest;
    if (!rest.empty()) {
        req.json = JsonParser(req.body).parse();
        req.hasJson = true;
    }
    return req;
}

void HttpRequest::writeCors() const {
    send("HTTP/1.1 204 No Content\r\nAccess-Control-Allow-Origin: *\r\nAccess-Control-Allow-Methods: GET, POST, PUT, DELETE\r\n"
         "Access-Control-Allow-Headers: Content-Type, Authorization\r\nConnection: close\r\n\r\n");
}
void HttpRequest::writeNotFound() const { send("HTTP/1.1 404 Not Found\r\nConnection: close\r\nContent-Length: 9\r\n\r\nNot Found"); }
void HttpRequest::writeJson(const std::string &text) const {
    send("HTTP/1.1 200 OK\r\nAccess-Control-Allow-Origin: *\r\nContent-Type: application/json; charset=utf-8\r\nConnection: close\r\n"
         "Content-Length: " + std::to_string(text.size()) + "\r\n\r\n" + text);
}
void HttpRequest::writeStreamStart() const {
    send("HTTP/1.1 200 OK\r\nAccess-Control-Allow-Origin: *\r\nContent-Type: text/event-stream; charset=utf-8\r\nConnection: close\r\n"
         "Transfer-Encoding: chunked\r\n\r\n");
}
void HttpRequest::writeStreamChunk(const std::string &data) const {
    char len[32];
    std::snprintf(len, sizeof(len), "%zx\r\n", data.size());
    send(std::string(len) + data + "\r\n");
}
void HttpRequest::writeStreamEnd() const { send("0000\r\n\r\n"); }

// ---- NaiveCache ---------------------------------------------------------------------------------------------------------

uint32_t NaiveCache::resolveDeltaPrompt(std::vector<ChatItem> &messages) {
    const size_t n = items_.size();
    if (n == 0) return 0;
    if (messages.size() > n) {
        bool prefix = true;
        for (size_t i = 0; i < n && prefix; i++)
            prefix = items_[i].msg.role == messages[i].role && items_[i].msg.message == messages[i].message;
        if (prefix) {
            const uint32_t start = items_[n - 1].endPos;
            std::printf("🐤 Found naive cache for %zu messages, pos=%u\n", n, start);
            messages.erase(messages.begin(), messages.begin() + (long)n);
            return start;
        }
    }
    clear();
    return 0;
}

std::string chunkJson(const std::string *delta, bool stop) {
    std::string s = "{\"id\": \"cmpl-c0\", \"object\": \"chat.completion\", \"created\": " + std::to_string((long long)std::time(nullptr)) +
                    ", \"model\": \"Distributed Model\", \"choices\": [{\"index\": 0, \"finish_reason\": " + (stop ? "\"stop\"" : "\"\"");
    if (!stop) s += ", \"delta\": {\"role\": \"assistant\", \"content\": " + jsonEscape(delta ? *delta : std::string()) + "}";
    s += "}]}";
    return s;
}

// ---- completion -----------------------------------------------------------------------------------------------------------

namespace {
std::vector<std::string> stopPieces(const Tokenizer &tok) {
    std::vector<std::string> out;
    for (int32_t id : tok.data().eosIds) out.push_back(tok.data().vocab[(size_t)id]);
    return out;
}
int maxLen(const std::vector<std::string> &v) {
    size_t m = 0;
    for (const std::string &s : v) m = std::max(m, s.size());
    return (int)m;
}
}  // namespace

ApiServer::ApiServer(InferenceBackend &backend, Tokenizer &tokenizer, const ApiConfig &cfg)
    : backend_(backend), tok_(tokenizer), cfg_(cfg), sampler_(std::min<uint32_t>(backend.vocabSize(), tokenizer.vocabSize()), cfg.temperature, cfg.topp, cfg.seed),
      stops_(stopPieces(tokenizer)),
      gen_(cfg.chatTemplate.empty() ? TEMPLATE_UNKNOWN : parseChatTemplateType(cfg.chatTemplate), tokenizer.data().chatTemplate,
           stops_.empty() ? std::string() : stops_[0]),
      det_(tokenizer.data().eosIds, stops_, maxLen(stops_), maxLen(stops_)) {
    backend_.setVocabLimit(tokenizer.vocabSize());
    std::printf("⭐ Chat template: %s\n", chatTemplateTypeName(gen_.type()));
    for (const std::string &s : stops_) std::printf("🛑 Stop: %s\n", s.c_str());
}

void ApiServer::complete(HttpRequest &req) {
    if (!req.hasJson || !req.json.isObject()) throw std::runtime_error("request body must be a JSON object");
    const JsonValue &body = req.json;
    std::vector<ChatItem> messages;
    for (const JsonValue &m : body.at("messages").arr) messages.push_back({m.at("role").str, m.at("content").str});
    const bool stream = body.boolOr("stream", false);
    const int maxTokens = (int)body.numberOr("max_tokens", -1);
    sampler_.setTemperature((float)body.numberOr("temperature", cfg_.temperature));
    sampler_.setTopp((float)body.numberOr("top_p", cfg_.topp));
    if (const JsonValue *s = body.find("seed")) if (s->kind == JsonValue::Number) sampler_.setSeed((uint64_t)s->num);

    const uint32_t seqLen = backend_.seqLen();
    const uint32_t startPos = cache_.resolveDeltaPrompt(messages);
    const GeneratedChat g = gen_.generate(messages, true);
    std::printf("🔹%s🔸", g.content.c_str());
    const std::vector<int32_t> tokens = tok_.encode(g.content, startPos == 0, true);
    const uint32_t nPrompt = (uint32_t)tokens.size();
    const uint32_t promptEnd = std::min(startPos + nPrompt - 1, seqLen);
    const uint32_t maxPred = maxTokens > 0 ? std::min(promptEnd + (uint32_t)maxTokens, seqLen) : seqLen;
    for (const ChatItem &m : messages) cache_.push(promptEnd, m);

    std::string buffer;
    if (stream) req.writeStreamStart();
    if (!g.publicPrompt.empty()) {
        if (stream) req.writeStreamChunk("data: " + chunkJson(&g.publicPrompt, false) + "\r\n\r\n");
        buffer += g.publicPrompt;
    }
    uint32_t pos = startPos;
    const uint32_t n = promptEnd - pos;
    backend_.prefill(std::vector<int32_t>(tokens.begin(), tokens.begin() + n), pos);
    pos += n;
    int32_t token = n < tokens.size() ? tokens[n] : tokens.back();
    tok_.resetDecoder();
    det_.reset();
    while (pos < maxPred) {
        token = backend_.next(token, pos, sampler_);
        const std::string piece = tok_.decode(token);
        const EosDetectorResult kind = det_.append(token, piece);
        if (!piece.empty()) { std::printf("%s", piece.c_str()); std::fflush(stdout); }
        if (kind == NOT_EOS || kind == EOS) {
            const std::string delta = det_.getDelta();
            if (!delta.empty()) {
                if (stream) req.writeStreamChunk("data: " + chunkJson(&delta, false) + "\r\n\r\n");
                buffer += delta;
            }
            det_.reset();
        }
        pos++;
        if (kind == EOS) break;
    }
    if (pos == seqLen) cache_.clear();
    else cache_.push(pos, {"assistant", buffer});
    if (stream) {
        req.writeStreamChunk("data: " + chunkJson(nullptr, true) + "\r\n\r\n");
        req.writeStreamChunk("data: [DONE]");
        req.writeStreamEnd();
    } else {
        const uint32_t nCompletion = pos - promptEnd;
        req.writeJson("{\"id\": \"cmpl-j0\", \"object\": \"chat.completion\", \"created\": " + std::to_string((long long)std::time(nullptr)) +
                      ", \"model\": \"Distributed Model\", \"usage\": {\"completion_tokens\": " + std::to_string(nCompletion) +
                      ", \"prompt_tokens\": " + std::to_string(nPrompt) + ", \"total_tokens\": " + std::to_string(nPrompt + nCompletion) +
                      "}, \"choices\": [{\"index\": 0, \"message\": {\"role\": \"assistant\", \"content\": " + jsonEscape(buffer) +
                      "}, \"finish_reason\": \"stop\"}]}");
    }
    std::printf("🔶\n");
}

void ApiServer::models(HttpRequest &req) const {
    req.writeJson("{\"object\": \"list\", \"data\": [{\"id\": " + jsonEscape(cfg_.modelName) + ", \"object\": \"model\", \"created\": 0, \"owned_by\": \"user\"}]}");
}

void ApiServer::serve() {
    const int srv = ::socket(AF_INET, SOCK_STREAM, 0);
    if (srv < 0) throw std::runtime_error("Cannot create socket");
    int one = 1;
    ::setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)cfg_.port);
    if (::inet_pton(AF_INET, cfg_.host.c_str(), &addr.sin_addr) != 1) { ::close(srv); throw std::runtime_error("Invalid host address: " + cfg_.host); }
    if (::bind(srv, (sockaddr *)&addr, sizeof(addr)) != 0 || ::listen(srv, 8) != 0) {
        ::close(srv);
        throw std::runtime_error("Cannot bind port " + std::to_string(cfg_.port));
    }
    if (cfg_.host == "0.0.0.0" || cfg_.host == "127.0.0.1") std::printf("Server URL: http://localhost:%d/v1/\n", cfg_.port);
    std::fflush(stdout);
    int served = 0;
    while (cfg_.maxRequests == 0 || served < cfg_.maxRequests) {
        const int fd = ::accept(srv, nullptr, nullptr);
        if (fd < 0) continue;
        ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        try {
            HttpRequest req = HttpRequest::read(fd);
            std::printf("🔷 %s %s\n", req.method.c_str(), req.path.c_str());
            if (req.method == "OPTIONS") req.writeCors();
            else if (req.method == "POST" && req.path == "/v1/chat/completions") complete(req);
            else if (req.method == "GET" && req.path == "/v1/models") models(req);
            else req.writeNotFound();
        } catch (const std::exception &e) {
            std::printf("Socket error: %s\n", e.what());
        }
        std::fflush(stdout);
        ::close(fd);
        served++;
    }
    ::close(srv);
}

}  // namespace dl
