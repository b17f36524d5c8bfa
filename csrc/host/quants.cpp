#include "quants.hpp"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

namespace dl {

const char *floatTypeName(FloatType t) {
    switch (t) {
        case F_32: return "f32";
        case F_16: return "f16";
        case F_Q40: return "q40";
        case F_Q80: return "q80";
        default: return "unk";
    }
}

FloatType parseFloatType(const char *s) {
    const std::string v(s);
    if (v == "f32") return F_32;
    if (v == "f16") return F_16;
    if (v == "q40") return F_Q40;
    if (v == "q80") return F_Q80;
    throw std::invalid_argument("Invalid float type: " + v);
}

size_t blockElems(FloatType t) {
    return (t == F_Q40 || t == F_Q80) ? kQBlock : 1;
}

size_t tensorBytes(FloatType t, size_t n) {
    switch (t) {
        case F_32: return n * 4;
        case F_16: return n * 2;
        case F_Q40:
            if (n % kQBlock) throw std::invalid_argument("q40 tensor length must be a multiple of 32");
            return n / kQBlock * kQ40Bytes;
        case F_Q80:
            if (n % kQBlock) throw std::invalid_argument("q80 tensor length must be a multiple of 32");
            return n / kQBlock * kQ80Bytes;
        default: throw std::invalid_argument("Unsupported float type");
    }
}

// ---- IEEE binary16 <-> binary32, pure integer arithmetic -------------------------------------

uint16_t f32ToF16(float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const uint32_t absu = u & 0x7fffffffu;
    if (absu >= 0x7f800000u)  // inf / nan
        return (uint16_t)(sign | 0x7c00u | ((absu > 0x7f800000u) ? (0x200u | ((absu >> 13) & 0x3ffu)) : 0u));
    if (absu >= 0x477ff000u)  // rounds to >= 65520 -> inf
        return (uint16_t)(sign | 0x7c00u);
    if (absu < 0x33000001u)   // <= 2^-25 rounds to zero (ties-to-even at exactly 2^-25)
        return (uint16_t)sign;
    const int32_t exp = (int32_t)(absu >> 23) - 127;
    uint32_t man = (absu & 0x7fffffu) | 0x800000u;  // 24-bit significand
    int shift;                                      // bits dropped from the 24-bit significand
    uint32_t base;
    if (exp < -14) {          // subnormal half: value = man * 2^(exp-23), unit = 2^-24
        shift = -exp - 1;     // 13 + (-14 - exp)
        base = 0;
    } else {
        shift = 13;
        base = (uint32_t)(exp + 15) << 10;
        man &= 0x7fffffu;
    }
    const uint32_t kept = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    uint32_t h = base + kept;
    if (rem > half || (rem == half && (kept & 1u))) h++;  // carries propagate into the exponent correctly
    return (uint16_t)(sign | h);
}

float f16ToF32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t u;
    if (exp == 0) {
        if (man == 0) {
            u = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        u = sign | 0x7f800000u | (man << 13);
    } else {
        u = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

static inline void storeU16(uint8_t *p, uint16_t v) { std::memcpy(p, &v, 2); }
static inline uint16_t loadU16(const uint8_t *p) { uint16_t v; std::memcpy(&v, p, 2); return v; }

// ---- Q80 ---------------------------------------------------------------------------------------

void quantizeQ80(const float *x, uint8_t *out, size_t n) {
    if (n % kQBlock) throw std::invalid_argument("q80: n % 32 != 0");
    for (size_t b = 0; b < n / kQBlock; b++) {
        const float *xb = x + b * kQBlock;
        uint8_t *ob = out + b * kQ80Bytes;
        float amax = 0.f;
        for (int j = 0; j < kQBlock; j++) amax = std::fmax(amax, std::fabs(xb[j]));
        const float d = amax / 127.f;
        const float inv = d != 0.f ? 1.f / d : 0.f;
        storeU16(ob, f32ToF16(d));
        for (int j = 0; j < kQBlock; j++) ob[2 + j] = (uint8_t)(int8_t)std::round(xb[j] * inv);
    }
}

void dequantizeQ80(const uint8_t *in, float *y, size_t n) {
    if (n % kQBlock) throw std::invalid_argument("q80: n % 32 != 0");
    for (size_t b = 0; b < n / kQBlock; b++) {
        const uint8_t *ib = in + b * kQ80Bytes;
        const float d = f16ToF32(loadU16(ib));
        for (int j = 0; j < kQBlock; j++) y[b * kQBlock + j] = (float)(int8_t)ib[2 + j] * d;
    }
}

// ---- Q40 ---------------------------------------------------------------------------------------

void quantizeQ40(const float *x, uint8_t *out, size_t n) {
    if (n % kQBlock) throw std::invalid_argument("q40: n % 32 != 0");
    for (size_t b = 0; b < n / kQBlock; b++) {
        const float *xb = x + b * kQBlock;
        uint8_t *ob = out + b * kQ40Bytes;
        float amax = 0.f, extreme = 0.f;  // the signed value with the largest magnitude maps to nibble 0
        for (int j = 0; j < kQBlock; j++) {
            const float a = std::fabs(xb[j]);
            if (a > amax) { amax = a; extreme = xb[j]; }
        }
        const float d = extreme / -8.f;
        const float inv = d != 0.f ? 1.f / d : 0.f;
        storeU16(ob, f32ToF16(d));
        for (int j = 0; j < kQBlock / 2; j++) {
            int lo = (int)(xb[j] * inv + 8.5f);
            int hi = (int)(xb[j + kQBlock / 2] * inv + 8.5f);
            lo = lo < 0 ? 0 : (lo > 15 ? 15 : lo);
            hi = hi < 0 ? 0 : (hi > 15 ? 15 : hi);
            ob[2 + j] = (uint8_t)(lo | (hi << 4));
        }
    }
}

void dequantizeQ40(const uint8_t *in, float *y, size_t n) {
    if (n % kQBlock) throw std::invalid_argument("q40: n % 32 != 0");
    for (size_t b = 0; b < n / kQBlock; b++) {
        const uint8_t *ib = in + b * kQ40Bytes;
        const float d = f16ToF32(loadU16(ib));
        for (int j = 0; j < kQBlock / 2; j++) {
            y[b * kQBlock + j] = (float)((int)(ib[2 + j] & 0x0f) - 8) * d;
            y[b * kQBlock + j + kQBlock / 2] = (float)((int)(ib[2 + j] >> 4) - 8) * d;
        }
    }
}

void dequantize(FloatType t, const uint8_t *in, float *y, size_t n) {
    switch (t) {
        case F_32: std::memcpy(y, in, n * 4); break;
        case F_16: for (size_t i = 0; i < n; i++) y[i] = f16ToF32(loadU16(in + 2 * i)); break;
        case F_Q40: dequantizeQ40(in, y, n); break;
        case F_Q80: dequantizeQ80(in, y, n); break;
        default: throw std::invalid_argument("Unsupported float type");
    }
}

void quantize(FloatType t, const float *x, uint8_t *out, size_t n) {
    switch (t) {
        case F_32: std::memcpy(out, x, n * 4); break;
        case F_16: for (size_t i = 0; i < n; i++) storeU16(out + 2 * i, f32ToF16(x[i])); break;
        case F_Q40: quantizeQ40(x, out, n); break;
        case F_Q80: quantizeQ80(x, out, n); break;
        default: throw std::invalid_argument("Unsupported float type");
    }
}

}  // namespace dl
