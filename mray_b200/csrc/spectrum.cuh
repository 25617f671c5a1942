// Hero-wavelength spectral transport: device functions shared by the standalone
// spectrum entry points and the path tracer's shading kernels.
//   SingleSampleSpectrumWavelength   Tracer/SpectrumContext.cu:L14-135
//   ConvertSpectraToRGBSingle        Tracer/SpectrumContext.cu:L137-171
//   Converter::ConvertAlbedo/Radiance Tracer/SpectrumContext.hpp:L37-150
// The coefficient LUT (9 x 64^3 fp32 = 9.4 MB) stays a plain linear array: it is L2-resident on B200
// (126 MB), and a software trilinear fetch reproduces the reference's host-backend filter
// (Device/CPU/TextureViewCPU.h:L258-376) instead of the 9-bit fixed-point weights of the texture unit,
// so both reference backends are matched to fp32 rounding. A constant albedo / radiance needs its three
// coefficients only once: the renderer hoists the fetch to StartRender and shading is four FMAs + one
// rsqrt per wavelength.
#pragma once
#include "common.cuh"

namespace mrb
{

constexpr int CIE_START = 360, CIE_N = 471;   // Color::CIE_1931_RANGE = [360, 831)

struct SpectrumData
{
    const float*  lut;          // 9 * n^3: block (3 * maxChannel + k) holds coefficient k; index z*n*n + y*n + x
    uint32_t      n;            // 64
    const float4* observer;     // CIE_N entries (x, y, z, 0), normalised by the X/Y/Z integrals
    const float*  illuminant;   // CIE_N entries, normalised
    float         xyzToRGB[9];
    uint32_t      mode;         // WavelengthSampleMode: 0 Uniform, 1 GaussianMIS, 2 HyperbolicPBRT
};

// Math::Lerp, unfused like the host reference: the polynomial multiplies coefficient rounding by lambda^2
__device__ __forceinline__ float SpecLerp(float a, float b, float t)
{ return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, t)), __fmul_rn(b, t)); }

__device__ __forceinline__ float SpecGaussian(float x, float sigma, float mu)
{
    const float invSqrt2Pi = (1.0f / 1.41421356237309504880f) * (1.0f / 1.77245385090551602729f);
    const float sigmaInv = 1.0f / sigma;
    const float pw = (x - mu) * sigmaInv;
    float r = invSqrt2Pi * sigmaInv * expf(-0.5f * pw * pw);
    return (r < 1.17549435e-38f) ? 0.0f : r;   // a denormal pdf would make DivideByPDF overflow
}

// one random number -> 4 stratified rotations -> wavelengths + pdfs
__device__ __forceinline__ void SampleWavelengths(uint32_t mode, float xi0, float w[4], float p[4])
{
    const float START = float(CIE_START), END = float(CIE_START + CIE_N - 1);
    const float offsets[4] = {-0.5f, -0.25f, 0.0f, 0.25f};
    #pragma unroll
    for(int i = 0; i < 4; i++)
    {
        float x = xi0 + offsets[i];
        if(x < 0.0f) x += 1.0f;
        if(x >= 1.0f) x -= 1.0f;
        if(mode == 0u)
        {
            w[i] = __fadd_rn(__fmul_rn(x, END - START), START);   // unfused like the host reference: bit-exact
            p[i] = 1.0f / (END - START);
        }
        else if(mode == 1u)
        {
            const float SIGMA[2] = {25.0f, 48.0f}, MU[2] = {452.0f, 576.0f}, MIS[2] = {0.384615384615f, 0.615384615385f};
            const int si = (x < MIS[0]) ? 0 : 1, oi = si ^ 1;
            float lxi = (x < MIS[0]) ? x / MIS[0] : (x - MIS[0]) / (1.0f - MIS[0]);
            lxi = fminf(lxi, 0.99999994f);
            const float e = erfinvf(2.0f * lxi - 1.0f);
            float v = 1.41421356237309504880f * SIGMA[si] * e + MU[si];
            if(isinf(e)) { const float mm = 3.5f * SIGMA[si]; v = fminf(fmaxf(v, -mm), mm); }
            w[i] = v;
            p[i] = SpecGaussian(v, SIGMA[si], MU[si]) * MIS[si] + SpecGaussian(v, SIGMA[oi], MU[oi]) * MIS[oi];
        }
        else
        {
            const float a = 0.85691062f - 1.82750197f * x;
            w[i] = 538.0f - 138.888889f * atanhf(a);
            const float dn = coshf(0.0072f * (w[i] - 538.0f));
            p[i] = 0.0039398042f / (dn * dn);
        }
    }
}

// linear, clamped fetch position in a 1-D table addressed in texels (TextureViewCPU<1,...>)
__device__ __forceinline__ void SpecInterp1(float x, int size, int& i0, int& i1, float& fr)
{
    const float texel = __fsub_rn(__fmul_rn(__fdiv_rn(x, float(size)), float(size)), 0.5f);
    float base; fr = modff(texel, &base);
    int start = int(base);
    if(fr < 0.0f) { start -= 1; fr = fabsf(fr); }
    i0 = min(max(start, 0), size - 1); i1 = min(max(start + 1, 0), size - 1);
}

__device__ __forceinline__ float FetchIlluminant(const SpectrumData& s, float wave)
{
    int a, b; float f;
    SpecInterp1(wave + (0.5f - float(CIE_START)), CIE_N, a, b, f);
    return SpecLerp(__ldg(s.illuminant + a), __ldg(s.illuminant + b), f);
}

__device__ __forceinline__ float SpecFetch3(const float* tex, int n, const float uv[3])
{
    int st[3]; float fr[3];
    #pragma unroll
    for(int k = 0; k < 3; k++)
    {
        const float texel = __fsub_rn(__fmul_rn(uv[k], float(n)), 0.5f);
        float base; float f = modff(texel, &base);
        st[k] = int(base);
        if(f < 0.0f) { st[k] -= 1; f = fabsf(f); }
        fr[k] = f;
    }
    float pix[8];
    #pragma unroll
    for(int k = 0; k < 2; k++)
    #pragma unroll
    for(int j = 0; j < 2; j++)
    #pragma unroll
    for(int i = 0; i < 2; i++)
    {
        const int x = min(max(st[0] + i, 0), n - 1), y = min(max(st[1] + j, 0), n - 1), z = min(max(st[2] + k, 0), n - 1);
        pix[(k << 2) + (j << 1) + i] = __ldg(tex + (size_t(z) * n + y) * n + x);
    }
    #pragma unroll
    for(int i = 0; i < 4; i++) pix[i] = SpecLerp(pix[2 * i], pix[2 * i + 1], fr[0]);
    #pragma unroll
    for(int i = 0; i < 2; i++) pix[i] = SpecLerp(pix[2 * i], pix[2 * i + 1], fr[1]);
    return SpecLerp(pix[0], pix[1], fr[2]);
}

__device__ __forceinline__ float InvSmoothstep(float y) { return 0.5f - sinf(asinf(1.0f - 2.0f * y) * (1.0f / 3.0f)); }

// Converter::ConvertAlbedo up to (not including) the polynomial: the three coefficients of an RGB value
__device__ __forceinline__ float3 FetchAlbedoCoeffs(const SpectrumData& s, float r, float g, float b)
{
    const float rgb[3] = {r, g, b};
    int maxI = 0; float mx = rgb[0];
    if(rgb[1] > mx) { mx = rgb[1]; maxI = 1; }
    if(rgb[2] > mx) { mx = rgb[2]; maxI = 2; }
    float xyz[3] = {0.f, 0.f, 0.f};
    if(mx > 1.0e-7f)
    {
        const float f = 1.0f / mx;
        xyz[0] = rgb[(maxI + 1) % 3] * f; xyz[1] = rgb[(maxI + 2) % 3] * f;
    }
    const float slice = fminf(fmaxf(mx, 0.0f), 1.0f);
    xyz[2] = InvSmoothstep(InvSmoothstep(slice));
    const float N = float(s.n), A = (N - 1.0f) / N, B = 0.5f / N;
    const float uv[3] = {__fadd_rn(__fmul_rn(xyz[0], A), B), __fadd_rn(__fmul_rn(xyz[1], A), B), __fadd_rn(__fmul_rn(xyz[2], A), B)};
    const size_t n3 = size_t(s.n) * s.n * s.n;
    return make_float3(SpecFetch3(s.lut + (3 * maxI + 0) * n3, int(s.n), uv),
                       SpecFetch3(s.lut + (3 * maxI + 1) * n3, int(s.n), uv),
                       SpecFetch3(s.lut + (3 * maxI + 2) * n3, int(s.n), uv));
}

// EvalPolynomial: sigmoid(c0 l^2 + c1 l + c2)
__device__ __forceinline__ float EvalSpectrum(float3 c, float lambda)
{
    const float t = fmaf(c.x, lambda, c.y);
    const float x = fmaf(t, lambda, c.z);
    const float dr = rsqrtf(fmaf(x, x, 1.0f));
    return fmaf(0.5f * x, dr, 0.5f);
}

// Converter::ConvertRadiance, first half: coefficients of radiance / (2 max) and the scale 2 max
__device__ __forceinline__ float4 FetchRadianceCoeffs(const SpectrumData& s, float r, float g, float b)
{
    const float mx = fmaxf(r, fmaxf(g, b));
    const float scale = mx * 2.0f;
    float3 c;
    if(scale == 0.0f) c = FetchAlbedoCoeffs(s, 0.f, 0.f, 0.f);
    else c = FetchAlbedoCoeffs(s, r / scale, g / scale, b / scale);
    return make_float4(c.x, c.y, c.z, scale);
}

__device__ __forceinline__ float EvalRadiance(const SpectrumData& s, float4 c, float lambda)
{
    float v = EvalSpectrum(make_float3(c.x, c.y, c.z), lambda);
    v *= FetchIlluminant(s, lambda);
    return v * c.w;
}

// ConvertSpectraToRGBSingle. A path that went through a dispersive interface ((Mt)Refract) carries ONE wavelength:
// its secondary waves are SpectrumWaves::DISPERSED_WAVE (-1), only sample 0 counts and the 1/4 weight becomes 1
// (Tracer/SpectrumContext.cu:L137-171, Tracer/TracerTypes.h:L114,L386-401).
__device__ __forceinline__ float3 SpectraToRGB(const SpectrumData& s, const float value[4], const float w[4], const float p[4])
{
    float X = 0.f, Y = 0.f, Z = 0.f;
    const bool dispersed = w[1] == -1.0f;
    #pragma unroll
    for(int i = 0; i < 4; i++)
    {
        if(dispersed && i > 0) break;
        int a, b; float f;
        SpecInterp1(w[i] + (0.5f - float(CIE_START)), CIE_N, a, b, f);
        const float4 oa = __ldg(s.observer + a), ob = __ldg(s.observer + b);
        const float val = (p[i] == 0.0f) ? 0.0f : value[i] / p[i];
        X += SpecLerp(oa.x, ob.x, f) * val; Y += SpecLerp(oa.y, ob.y, f) * val; Z += SpecLerp(oa.z, ob.z, f) * val;
    }
    const float weight = dispersed ? 1.0f : 0.25f;
    X *= weight; Y *= weight; Z *= weight;
    return make_float3(s.xyzToRGB[0] * X + s.xyzToRGB[1] * Y + s.xyzToRGB[2] * Z,
                       s.xyzToRGB[3] * X + s.xyzToRGB[4] * Y + s.xyzToRGB[5] * Z,
                       s.xyzToRGB[6] * X + s.xyzToRGB[7] * Y + s.xyzToRGB[8] * Z);
}

} // namespace mrb

struct mrb_spectrum_t
{
    mrb::SpectrumData d = {};
    mrb::DeviceBlock  mem;
};
