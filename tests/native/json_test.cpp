// Unit test of csrc/app/json.hpp (CPU): parser coverage for what OpenAI-style clients send, escaping of what the model emits.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../csrc/app/json.hpp"

using namespace dl;

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)

static bool throws(const char *text) {
    try { const std::string s(text); JsonParser(s).parse(); } catch (const std::exception &) { return true; }
    return false;
}

int main() {
    const std::string doc = R"({"messages": [{"role": "user", "content": "line1\nq\"uote\" \\ \/ tab\t \u00e9 \ud83d\ude00 \u0041"}],
        "stream": true, "max_tokens": 128, "temperature": 0.25, "top_p": 9e-1, "seed": 12345678901, "stop": null, "nested": {"a": [1, [2, {"b": false}]]},
        "empty_o": {}, "empty_a": []})";
    const JsonValue v = JsonParser(doc).parse();
    CHECK(v.isObject());
    const JsonValue &m = v.at("messages");
    CHECK(m.isArray() && m.arr.size() == 1 && m.arr[0].at("role").str == "user");
    CHECK(m.arr[0].at("content").str == std::string("line1\nq\"uote\" \\ / tab\t \xc3\xa9 \xf0\x9f\x98\x80 A"));
    CHECK(v.boolOr("stream", false) && v.numberOr("max_tokens", -1) == 128 && v.numberOr("temperature", 1) == 0.25);
    CHECK(v.numberOr("top_p", 0) == 0.9 && (uint64_t)v.numberOr("seed", 0) == 12345678901ull);
    CHECK(v.at("stop").kind == JsonValue::Null && v.numberOr("missing", 7) == 7 && !v.boolOr("missing", false));
    CHECK(v.at("nested").at("a").arr[1].arr[1].at("b").kind == JsonValue::Bool);
    CHECK(v.at("empty_o").obj.empty() && v.at("empty_a").arr.empty() && v.find("nope") == nullptr);
    CHECK(throws("{\"a\": }") && throws("[1, 2") && throws("{\"a\": 1} x") && throws("\"unterminated") && throws("{\"a\": \"\\x\"}") && throws(""));
    bool missing = false;
    try { v.at("absent"); } catch (const std::exception &) { missing = true; }
    CHECK(missing);
    // lone surrogate -> U+FFFD, escaping round trip through the parser
    const std::string lone = "\"\\ud83d x\"";
    CHECK(JsonParser(lone).parse().str == std::string("\xef\xbf\xbd x"));
    const std::string raw = std::string("a\"b\\c\n\r\t\b\f") + '\x01' + "\xc3\xa9";
    const std::string esc = jsonEscape(raw);
    CHECK(esc.find('\n') == std::string::npos && esc.find("\\u0001") != std::string::npos);
    CHECK(JsonParser(esc).parse().str == raw);
    std::printf("JSON_TEST_OK\n");
    return 0;
}
