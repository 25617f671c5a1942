// dist2d.cuh — piecewise-constant 2-D distribution (sampling side) and the skysphere coordinate converters, as device
// functions the shading kernel calls. Restates Tracer/Distributions.h:L108-239 (DistributionPwC<1> / <2>) and
// Tracer/LightsDefault.hpp:L173-310 + Core/GraphicsFunctions.h:L253-383,L447-468 with the reference's operation order
// (explicit _rn intrinsics where nvcc could otherwise contract a multiply into an add).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mrb
{

// Row CDFs (h rows of w values, each ending in 1) and the marginal CDF over rows.
struct Dist2D
{
    const float* cdfX;
    const float* cdfY;
    uint32_t     w, h;
};

// std::lower_bound (Device/CPU/AlgBinarySearchCPU.h): first index with cdf[i] >= v
__device__ __forceinline__ uint32_t DistLowerBound(const float* __restrict__ cdf, uint32_t n, float v)
{
    uint32_t lo = 0, count = n;
    while(count > 0)
    {
        const uint32_t step = count >> 1, mid = lo + step;
        if(__ldg(cdf + mid) < v) { lo = mid + 1; count -= step + 1; }
        else count = step;
    }
    return lo;
}

// DistributionPwC<1>::SampleIndex (Distributions.h:L118-137), including its `indexF < 1` guard as written
__device__ __forceinline__ float DistSampleIndex1D(const float* __restrict__ cdf, uint32_t n, float xi, float& pdf)
{
    const uint32_t index = DistLowerBound(cdf, n, xi);
    const float prev = (index == 0u) ? 0.0f : __ldg(cdf + index - 1);
    const float my = __ldg(cdf + index);
    const float width = __fsub_rn(my, prev);
    const float t = __fdiv_rn(__fsub_rn(xi, prev), width);
    float indexF = __fadd_rn(float(index), t);
    indexF = (indexF < 1.0f) ? indexF : nextafterf(indexF, -3.402823466e+38f);
    pdf = __fmul_rn(width, float(n));
    return indexF;
}

// DistributionPwC<1>::PdfIndex (L147-157)
__device__ __forceinline__ float DistPdfIndex1D(const float* __restrict__ cdf, uint32_t n, float index)
{
    const uint32_t i = uint32_t(index);
    const float prev = (i == 0u) ? 0.0f : __ldg(cdf + i - 1);
    return __fmul_rn(__fsub_rn(__ldg(cdf + i), prev), float(n));
}

// DistributionPwC<2>::SampleUV (L183-213): -> (u, v, pdf)
__device__ __forceinline__ float3 DistSampleUV(const Dist2D& d, float xi0, float xi1)
{
    float pdfY, pdfX;
    const float iy = DistSampleIndex1D(d.cdfY, d.h, xi1, pdfY);
    const uint32_t row = uint32_t(iy);
    const float ix = DistSampleIndex1D(d.cdfX + size_t(row) * d.w, d.w, xi0, pdfX);
    return make_float3(__fmul_rn(ix, __fdiv_rn(1.0f, float(d.w))), __fmul_rn(iy, __fdiv_rn(1.0f, float(d.h))), __fmul_rn(pdfY, pdfX));
}

// DistributionPwC<2>::PdfUV (L215-239)
__device__ __forceinline__ float DistPdfUV(const Dist2D& d, float u, float v)
{
    float fx = __fmul_rn(u, float(d.w)), fy = __fmul_rn(v, float(d.h));
    fx = fminf(fx, float(d.w) - 1.0f); fy = fminf(fy, float(d.h) - 1.0f);
    const uint32_t row = uint32_t(fy);
    const float pm = DistPdfIndex1D(d.cdfY, d.h, fy);
    const float pc = DistPdfIndex1D(d.cdfX + size_t(row) * d.w, d.w, fx);
    return __fmul_rn(pc, pm);
}

// ---- skysphere coordinate converters; mode 1 = SphericalCoordConverter, 2 = CoOctaCoordConverter; Y-up directions ----
constexpr float SKY_PI = 3.14159265358979323846f;

__device__ __forceinline__ float2 SkyDirToUV(uint32_t mode, float dx, float dy, float dz)
{
    const float zx = dz, zy = dx, zz = dy;   // TransformGen::YUpToZUp
    if(mode == 1u)
    {
        const float azimuth = atan2f(zy, zx);
        const float incl = acosf(fminf(fmaxf(zz, -1.0f), 1.0f));
        return make_float2(__fdiv_rn(__fmul_rn(__fadd_rn(azimuth, SKY_PI), 0.5f), SKY_PI),
                           __fsub_rn(1.0f, __fmul_rn(incl, 1.0f / SKY_PI)));
    }
    const float TwoOvrPi = (1.0f / SKY_PI) * 2.0f;
    if(zx == 0.0f && zy == 0.0f) return make_float2(0.0f, 0.0f);
    const float phiPrime = atanf(__fdiv_rn(fabsf(zy), fabsf(zx)));
    const float r1 = __fsub_rn(1.0f, fabsf(zz));
    const float radius = r1 > 0.0f ? __fsqrt_rn(r1) : 0.0f;
    float v = __fmul_rn(__fmul_rn(radius, TwoOvrPi), phiPrime);
    float u = __fsub_rn(radius, v);
    if(zz < 0.0f) { const float up = __fsub_rn(1.0f, v), vp = __fsub_rn(1.0f, u); u = up; v = vp; }
    u = __fmul_rn(u, copysignf(1.0f, zx)); v = __fmul_rn(v, copysignf(1.0f, zy));
    return make_float2(__fmul_rn(__fadd_rn(u, 1.0f), 0.5f), __fmul_rn(__fadd_rn(v, 1.0f), 0.5f));
}

__device__ __forceinline__ float3 SkyUVToDir(uint32_t mode, float u0, float v0)
{
    float zx, zy, zz;
    if(mode == 1u)
    {
        const float theta = __fsub_rn(__fmul_rn(__fmul_rn(u0, SKY_PI), 2.0f), SKY_PI);
        const float phi = __fmul_rn(__fsub_rn(1.0f, v0), SKY_PI);
        float sT, cT, sP, cP;
        sincosf(theta, &sT, &cT); sincosf(phi, &sP, &cP);
        zx = __fmul_rn(cT, sP); zy = __fmul_rn(sT, sP); zz = cP;
    }
    else
    {
        const float PiOvr4 = SKY_PI * 0.25f;
        const float u = __fsub_rn(__fmul_rn(u0, 2.0f), 1.0f), v = __fsub_rn(__fmul_rn(v0, 2.0f), 1.0f);
        const float ua = fabsf(u), va = fabsf(v);
        const float dd = __fsub_rn(1.0f, __fadd_rn(ua, va));
        const float radius = __fsub_rn(1.0f, fabsf(dd));
        float phiPrime = 0.0f;
        if(radius != 0.0f) phiPrime = __fmul_rn(__fadd_rn(__fdiv_rn(__fsub_rn(va, ua), radius), 1.0f), PiOvr4);
        float sinP, cosP; sincosf(phiPrime, &sinP, &cosP);
        const float cosPhi = __fmul_rn(copysignf(1.0f, u), cosP), sinPhi = __fmul_rn(copysignf(1.0f, v), sinP);
        zz = __fmul_rn(copysignf(1.0f, dd), __fsub_rn(1.0f, __fmul_rn(radius, radius)));
        const float xyFactor = __fmul_rn(radius, __fsqrt_rn(__fsub_rn(2.0f, __fmul_rn(radius, radius))));
        zx = __fmul_rn(cosPhi, xyFactor); zy = __fmul_rn(sinPhi, xyFactor);
    }
    return make_float3(zy, zz, zx);   // TransformGen::ZUpToYUp
}

// CoordConverter::ToSolidAnglePdf(pdf, dirYUp) / (pdf, uv)
__device__ __forceinline__ float SkyPdfFromDir(uint32_t mode, float pdf, float dirY)
{
    if(mode != 1u) return __fmul_rn(__fmul_rn(pdf, 0.25f), 1.0f / SKY_PI);
    const float sinPhi = sinf(acosf(fminf(fmaxf(dirY, -1.0f), 1.0f)));
    return (sinPhi <= 0.0f) ? 0.0f : __fdiv_rn(pdf, __fmul_rn(2.0f * (SKY_PI * SKY_PI), sinPhi));
}
__device__ __forceinline__ float SkyPdfFromUV(uint32_t mode, float pdf, float v)
{
    if(mode != 1u) return __fmul_rn(__fmul_rn(pdf, 0.25f), 1.0f / SKY_PI);
    const float sinPhi = sinf(__fmul_rn(__fsub_rn(1.0f, v), SKY_PI));
    return (sinPhi <= 0.0f) ? 0.0f : __fdiv_rn(pdf, __fmul_rn(2.0f * (SKY_PI * SKY_PI), sinPhi));
}

// host side (dist2d.cu)
struct Context;
// DistributionGroupPwC2D::Construct: function (device, h rows of w) -> cdfX (device, w*h), cdfY (device, h); rowTotals: h floats of scratch
void Dist2DBuild(Context& ctx, const float* function, uint32_t w, uint32_t h, float* cdfX, float* cdfY, float* rowTotals);
// parity taps: out[n*4] = {u, v, SampleUV pdf, PdfUV(u, v)}; conv[n*8] per direction (the layout of oracle/ref_build/ref_dist_tap.cpp)
void Dist2DSample(Context& ctx, const Dist2D& d, const float* xi, uint32_t n, float* out);
void SkyConverters(Context& ctx, uint32_t mode, const float* dirs, uint32_t n, float* out);
// KCExtractLuminance for a single-level texture: out[w*h] = Y row of the colour space's ToXYZ matrix . texel rgb
void TextureLuminance(Context& ctx, const void* texels, uint32_t w, uint32_t h, uint32_t channels, uint32_t format, const float yRow[3], float* out);

} // namespace mrb
