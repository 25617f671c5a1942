// Low-discrepancy samplers of the reference: Owen-scrambled Sobol and Z-Sobol
// (Ahmed & Wonka 2020), Tracer/Random.cu:L100-437, plus the path tracer's per-slot sampler state.
// Pure integer code; bit-exact with RNGGroupSobol / RNGGroupZSobol (tests/test_gpu_sampler.py).
#pragma once
#include "common.cuh"

namespace mrb
{

constexpr uint32_t SOBOL_MATRIX_WIDTH = 52, SOBOL_DIM_COUNT = 256;
enum : uint32_t { SAMPLER_INDEPENDENT = 0u, SAMPLER_SOBOL = 1u, SAMPLER_ZSOBOL = 2u, SAMPLER_TYPE_MASK = 0xFFu, SAMPLER_REFERENCE_SCRAMBLE = 0x100u };

__device__ __forceinline__ uint64_t MixBits(uint64_t v)
{
    v ^= (v >> 31); v *= 0x7FB5D329728EA185ull;
    v ^= (v >> 27); v *= 0x81DADEF4BC2DD44Dull;
    v ^= (v >> 33);
    return v;
}
// RNGFunctions::HashPCG64::Hash(dim, seed) (Tracer/Random.h:L598-633)
__device__ __forceinline__ uint64_t HashPCG64(uint64_t v)
{
    const uint64_t s = v * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t word = ((s >> ((s >> 59) + 5)) ^ s);
    word *= 12605985483714917081ull;
    return (word >> 43) ^ word;
}
__device__ __forceinline__ uint64_t HashDimSeed(uint32_t dim, uint32_t seed) { return HashPCG64(uint64_t(seed) + HashPCG64(uint64_t(dim))); }

// SobolCommon::ScambleOwenFast (Random.cu:L152-162). The hash acts on the bit-REVERSED value (Laine & Karras
// 2011); the reference returns it without reversing back, so the stratified digits end up in the low bits and the
// points it feeds to ToFloat01 are no better than random (16 scrambled 1-D points occupy the 16 strata as
// [1 0 1 2 1 0 1 0 1 2 0 3 0 4 0 0] instead of once each). `reverseBack` restores the net property; off = the
// reference's numbers bit for bit.
__device__ __forceinline__ uint32_t ScrambleOwenFast(uint32_t v, uint32_t seed, bool reverseBack)
{
    v = __brev(v);
    v ^= v * 0x3A20ADEAu;
    v += seed;
    v *= (seed >> 16) | 1u;
    v ^= v * 0x05526C56u;
    v ^= v * 0x53A22864u;
    return reverseBack ? __brev(v) : v;
}

__device__ __forceinline__ uint32_t SobolMatrixMult(uint64_t a, const uint32_t* __restrict__ m)
{
    uint32_t r = 0;
    for(uint32_t i = 0; i < SOBOL_MATRIX_WIDTH && a != 0ull; i++, a >>= 1)
        if(a & 1ull) r ^= __ldg(m + i);
    return r;
}

__device__ __forceinline__ void ScrambleRequest(uint32_t* s, int count, uint64_t h, bool rb)
{
    if(count == 1) s[0] = ScrambleOwenFast(s[0], uint32_t(h), rb);
    else if(count == 2) { s[0] = ScrambleOwenFast(s[0], uint32_t(h & 0xFFFFFFFFull), rb); s[1] = ScrambleOwenFast(s[1], uint32_t(h >> 32), rb); }
    else
    {
        const uint64_t MASK = (1u << 21) - 1;
        s[0] = ScrambleOwenFast(s[0], uint32_t(h & MASK), rb);
        s[1] = ScrambleOwenFast(s[1], uint32_t((h >> 21) & MASK), rb);
        s[2] = ScrambleOwenFast(s[2], uint32_t(h >> 42), rb);
    }
}

// SobolDetail::Sobol::Next / Next2D / Next3D: `count` numbers of one request at dimension `dim`
__device__ __forceinline__ void SobolNext(const uint32_t* __restrict__ matrices, uint32_t seed, uint32_t sampleIndex,
                                          uint32_t dim, int count, uint32_t* out, bool reverseBack)
{
    dim &= (SOBOL_DIM_COUNT - 1u);   // Sobol::RollDim (256 dimensions)
    for(int k = 0; k < count; k++)
        out[k] = SobolMatrixMult(sampleIndex, matrices + size_t((dim + uint32_t(k)) & (SOBOL_DIM_COUNT - 1u)) * SOBOL_MATRIX_WIDTH);
    ScrambleRequest(out, count, HashDimSeed(dim, seed), reverseBack);
}

// the 24 permutations of {0,1,2,3} (pbrt-v4 ZSobolSampler order) as four 48-bit columns, 2 bits per entry
__device__ __forceinline__ uint32_t ZSobolPermute(uint32_t p, uint32_t digit)
{
    // column `digit` of the permutation table, entry p at bits [2p, 2p+2)
    constexpr uint64_t COL0 = 0xFFFAAA555000ull;                 // 0 x6, 1 x6, 2 x6, 3 x6
    constexpr uint64_t COL1 = 0x0A5F05FA0FA5ull;
    constexpr uint64_t COL2 = 0x6124DC2CE6DEull;
    constexpr uint64_t COL3 = 0x94817383B97Bull;
    const uint64_t col = digit == 0u ? COL0 : (digit == 1u ? COL1 : (digit == 2u ? COL2 : COL3));
    return uint32_t(col >> (2u * p)) & 3u;
}

struct ZSobolGlobals { uint32_t initialMaxSPP, resMaxBits; };

// ZSobolDetail::ZSobol ctor + SampleIndex + Next*: uses the first three Joe-Kuo matrices
__device__ __forceinline__ void ZSobolNext(const uint32_t* __restrict__ matrices, uint32_t seed, uint32_t sampleIndexIn, uint64_t pixelMorton,
                                           ZSobolGlobals g, uint32_t dim, int count, uint32_t* out, bool reverseBack)
{
    uint32_t idx = sampleIndexIn, maxSPP = g.initialMaxSPP;
    for(uint32_t i = 0; i < 32u; i++)
    {
        const uint32_t curSize = g.initialMaxSPP << i;
        if(idx < curSize) break;
        idx -= curSize; maxSPP <<= 1;
    }
    const uint32_t log2SPP = (32u - uint32_t(__clz(int(maxSPP)))) - 1u;
    const uint32_t log4SPP = (log2SPP + 1u) >> 1;
    const int nBase4Digits = int(g.resMaxBits + log4SPP);
    const uint64_t mortonIndex = (pixelMorton << log2SPP) | uint64_t(idx);
    const uint32_t dimMixer = 0x55555555u * dim;
    uint64_t sampleIndex = 0ull;
    const bool pow2 = (log2SPP & 1u) != 0u;
    for(int i = nBase4Digits - 1; i >= (pow2 ? 1 : 0); i--)
    {
        const uint32_t digitShift = uint32_t(2 * i - (pow2 ? 1 : 0));
        uint32_t digit = uint32_t(mortonIndex >> digitShift) & 3u;
        const uint64_t higher = mortonIndex >> (digitShift + 2u);
        // (x >> 24) % 24 on the 40-bit value without a 64-bit division: 2^32 = 16 (mod 24)
        const uint64_t mixed = MixBits(higher ^ uint64_t(dimMixer)) >> 24;
        const uint32_t p = (uint32_t(mixed >> 32) * 16u + uint32_t(mixed & 0xFFFFFFFFull) % 24u) % 24u;
        digit = ZSobolPermute(p, digit);
        sampleIndex |= uint64_t(digit) << digitShift;
    }
    if(pow2)
    {
        const uint32_t digit = uint32_t(mortonIndex & 1ull);
        sampleIndex |= uint64_t(digit ^ uint32_t(MixBits((mortonIndex >> 1) ^ uint64_t(dimMixer)) & 1ull));
    }
    for(int k = 0; k < count; k++) out[k] = SobolMatrixMult(sampleIndex, matrices + size_t(k) * SOBOL_MATRIX_WIDTH);
    ScrambleRequest(out, count, HashDimSeed(dim, seed), reverseBack);
}

// Graphics::MortonCode::Compose2D<uint64_t> (Core/GraphicsFunctions.h:L698-714)
__device__ __forceinline__ uint64_t Morton2D(uint32_t x, uint32_t y)
{
    auto Expand = [](uint32_t v) -> uint64_t
    {
        uint64_t t = v;
        t = (t | (t << 16)) & 0x0000FFFF0000FFFFull; t = (t | (t << 8)) & 0x00FF00FF00FF00FFull;
        t = (t | (t << 4)) & 0x0F0F0F0F0F0F0F0Full; t = (t | (t << 2)) & 0x3333333333333333ull;
        t = (t | (t << 1)) & 0x5555555555555555ull;
        return t;
    };
    return Expand(x) | (Expand(y) << 1);
}

} // namespace mrb
