// Quantisation formats of the .m model file and of the activation ("sync") buffers.
//
// Behavioural parity target: reference src/nn/nn-quants.hpp:53-72 (block structs),
// src/nn/nn-quants.cpp:67-246 (codecs), converter/writer.py:29-78 (file writers).
//   Q40 block = fp16 scale d + 16 bytes; byte j holds element j (low nibble) and
//               element j+16 (high nibble); value = (nibble - 8) * d.
//   Q80 block = fp16 scale d + 32 int8;  value = q * d, d = amax/127.
// The implementation here is independent (bit-level f16 conversion, span based API).
#pragma once
#include <cstddef>
#include <cstdint>

namespace dl {

enum FloatType : int32_t { F_UNK = -1, F_32 = 0, F_16 = 1, F_Q40 = 2, F_Q80 = 3 };

constexpr int kQBlock = 32;       // elements per Q40/Q80 block
constexpr int kQ40Bytes = 18;     // 2 (fp16 scale) + 16 (nibbles)
constexpr int kQ80Bytes = 34;     // 2 (fp16 scale) + 32 (int8)

const char *floatTypeName(FloatType t);
FloatType parseFloatType(const char *s);   // "f32" | "f16" | "q40" | "q80"
size_t blockElems(FloatType t);            // 1 for f32/f16, 32 for q40/q80
size_t tensorBytes(FloatType t, size_t nElems);

uint16_t f32ToF16(float v);   // round-to-nearest-even, IEEE binary16
float f16ToF32(uint16_t h);

// All functions operate on n elements (n % 32 == 0) and packed byte streams.
void quantizeQ80(const float *x, uint8_t *out, size_t n);
void dequantizeQ80(const uint8_t *in, float *y, size_t n);
void quantizeQ40(const float *x, uint8_t *out, size_t n);
void dequantizeQ40(const uint8_t *in, float *y, size_t n);

// Generic: convert a packed tensor of type t into f32.
void dequantize(FloatType t, const uint8_t *in, float *y, size_t n);
void quantize(FloatType t, const float *x, uint8_t *out, size_t n);

}  // namespace dl
