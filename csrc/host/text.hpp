// Text stack: `.t` tokenizer file, BPE encoder / streaming decoder, sampler, chat templates, stop detector.
//
// Behavioural parity targets: reference src/tokenizer.cpp:42-164 (file format), :311-390 (encode),
// :224-309 (decode), :405-512 (sampler + xorshift* RNG at :25-36), :522-539 (chat stops),
// :549-637 (templates), :639-724 (EosDetector); converter/tokenizer-writer.py:3-57 (writer).
// The implementation is new: byte strings + hash maps instead of sorted C-string tables, an incremental
// pair-merge loop, std::string based stop detection.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace dl {

constexpr int32_t kTokenizerMagic = 0x567124;
constexpr int32_t kTokenizerMagicOld = 0x567123;

struct TokenizerData {
    std::vector<std::string> vocab;   // raw bytes per token id
    std::vector<float> scores;
    int32_t bosId = -1;
    bool addBos = true;
    std::vector<int32_t> eosIds;
    std::string chatTemplate;         // may be empty
    uint32_t maxTokenLength = 0;
};

TokenizerData readTokenizerFile(const std::string &path);
void writeTokenizerFile(const std::string &path, const TokenizerData &d);

class Tokenizer {
public:
    explicit Tokenizer(const std::string &path);
    explicit Tokenizer(TokenizerData data);

    // addSpecialTokens: match special tokens (id >= bosId) as literal prefixes of the remaining text.
    std::vector<int32_t> encode(const std::string &text, bool isStart, bool addSpecialTokens) const;
    // Streaming decode: returns the longest UTF-8-complete text available after appending `token`.
    // Invalid sequences are replaced by U+FFFD. BOS yields "", EOS flushes whatever is pending.
    std::string decode(int32_t token);
    void resetDecoder() { pending_.clear(); }
    bool isEos(int32_t token) const;
    std::string describe() const;   // the "📄 ..." lines

    const TokenizerData &data() const { return d_; }
    uint32_t vocabSize() const { return (uint32_t)d_.vocab.size(); }
    uint32_t regularVocabSize() const { return regularSize_; }

private:
    void index();
    int32_t findRegular(const std::string &s) const;
    TokenizerData d_;
    uint32_t regularSize_ = 0;
    std::unordered_map<std::string, int32_t> regular_;
    std::vector<int32_t> specials_;   // ids >= bosId in id order
    std::string pending_;
};

// xorshift* generator shared by the host and (by passing the drawn coin) the device sampler.
struct Rng {
    uint64_t state;
    explicit Rng(uint64_t seed) : state(seed) {}
    uint32_t nextU32();
    float nextF32();   // [0, 1)
};

void softmaxInPlace(float *x, size_t n);

class Sampler {
public:
    Sampler(uint32_t vocabSize, float temperature, float topp, uint64_t seed);
    int32_t sample(float *logits);          // modifies logits in place (temperature, softmax)
    float nextCoin() { return rng_.nextF32(); }
    void setTemperature(float t) { temperature_ = t; }
    void setTopp(float p) { topp_ = p; }
    void setSeed(uint64_t s) { rng_.state = s; seed_ = s; seedGeneration_++; }
    uint64_t seed() const { return seed_; }                       // last seed given (constructor or setSeed)
    uint32_t seedGeneration() const { return seedGeneration_; }   // bumped by every setSeed: device-side generators re-seed on change
    float temperature() const { return temperature_; }
    float topp() const { return topp_; }
    uint32_t vocabSize() const { return vocabSize_; }

private:
    uint32_t vocabSize_;
    float temperature_, topp_;
    Rng rng_;
    uint64_t seed_ = 0;
    uint32_t seedGeneration_ = 0;
    std::vector<std::pair<float, int32_t>> candidates_;
};

enum ChatTemplateType : int32_t { TEMPLATE_UNKNOWN = 0, TEMPLATE_LLAMA2 = 1, TEMPLATE_LLAMA3 = 2, TEMPLATE_DEEP_SEEK3 = 3, TEMPLATE_CHATML = 4 };
ChatTemplateType parseChatTemplateType(const std::string &name);   // "llama2" | "llama3" | "deepSeek3" | "chatml"
const char *chatTemplateTypeName(ChatTemplateType t);

struct ChatItem { std::string role, message; };
struct GeneratedChat { std::string content; std::string publicPrompt; };

class ChatTemplateGenerator {
public:
    ChatTemplateGenerator(ChatTemplateType type, const std::string &chatTemplate, const std::string &eos);
    GeneratedChat generate(const std::vector<ChatItem> &items, bool appendGenerationPrompt) const;
    ChatTemplateType type() const { return type_; }
private:
    ChatTemplateType type_;
    std::string eos_;
};

enum EosDetectorResult : int32_t { MAYBE_EOS = 0, EOS = 1, NOT_EOS = 2 };

// Streaming stop detector. Pieces are appended; a stop string may appear after up to `paddingLeft`
// leading bytes and be followed by up to `paddingRight` bytes inside the buffered text.
class EosDetector {
public:
    EosDetector(std::vector<int32_t> tokens, std::vector<std::string> pieces, int paddingLeft, int paddingRight);
    EosDetectorResult append(int32_t tokenId, const std::string &piece);
    bool isEos(int32_t tokenId) const;
    // Text that may be shown to the user; empty when nothing (or only the stop) is buffered.
    std::string getDelta() const;
    void reset();
private:
    std::vector<int32_t> tokens_;
    std::vector<std::string> pieces_;
    int padLeft_, padRight_;
    std::string buffer_;
    int eosPos_ = -1;
};

}  // namespace dl
