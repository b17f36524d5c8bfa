// exp(x) for x <= 0, written with explicitly rounded single operations (fused multiply-adds where fmaf is spelled out, separately
// rounded multiplies / adds elsewhere) so that the host sampler (csrc/host/text.cpp, g++) and the device sampler
// (csrc/cuda/sampler.cu, nvcc --use_fast_math) obtain bit-identical results: both samplers then see the same probabilities and,
// with integer prefix sums, make the same draw for the same seed.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define DL_EXP_HD __host__ __device__ __forceinline__
#define DL_EXP_MUL(a, b) __fmul_rn((a), (b))
#define DL_EXP_ADD(a, b) __fadd_rn((a), (b))
#define DL_EXP_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define DL_EXP_RINT(a) rintf(a)
#elif defined(__CUDACC__)
#define DL_EXP_HD __host__ __device__ inline
#define DL_EXP_MUL(a, b) ((a) * (b))
#define DL_EXP_ADD(a, b) ((a) + (b))
#define DL_EXP_FMA(a, b, c) fmaf((a), (b), (c))
#define DL_EXP_RINT(a) nearbyintf(a)
#else
#define DL_EXP_HD inline
#define DL_EXP_MUL(a, b) ((a) * (b))
#define DL_EXP_ADD(a, b) ((a) + (b))
#define DL_EXP_FMA(a, b, c) std::fmaf((a), (b), (c))
#define DL_EXP_RINT(a) std::nearbyintf(a)
#endif

namespace dl {

DL_EXP_HD float expNeg(float x) {
    if (!(x > -86.0f)) return 0.0f;
    if (x > 0.0f) x = 0.0f;
    const float n = DL_EXP_RINT(DL_EXP_MUL(x, 1.44269504088896341f));
    float r = DL_EXP_FMA(n, -0.693359375f, x);               // ln2 split (Cephes): hi has 9 significant bits, n * hi is exact
    r = DL_EXP_FMA(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = DL_EXP_FMA(p, r, 1.3981999507e-3f);
    p = DL_EXP_FMA(p, r, 8.3334519073e-3f);
    p = DL_EXP_FMA(p, r, 4.1665795894e-2f);
    p = DL_EXP_FMA(p, r, 1.6666665459e-1f);
    p = DL_EXP_FMA(p, r, 5.0000001201e-1f);
    const float z = DL_EXP_MUL(r, r);
    float y = DL_EXP_FMA(p, z, r);
    y = DL_EXP_ADD(y, 1.0f);
    // y * 2^n through the exponent field (n in [-124, 0], y in (0.7, 1.5): the result stays a normal number)
    uint32_t bits;
#if defined(__CUDA_ARCH__)
    bits = __float_as_uint(y);
#else
    std::memcpy(&bits, &y, 4);
#endif
    bits += (uint32_t)((int32_t)n << 23);
#if defined(__CUDA_ARCH__)
    return __uint_as_float(bits);
#else
    float out;
    std::memcpy(&out, &bits, 4);
    return out;
#endif
}

}  // namespace dl
