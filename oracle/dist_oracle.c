/* dist_oracle.c — CPU restatement of the reference's piecewise-constant 2-D distribution and of the
 * skysphere light built on it. TEST INFRASTRUCTURE ONLY: the product never links this.
 *
 *   Tracer/Distributions.cu  (MRAY_GPU_BACKEND_CPU kernels, L235-300): KCSegmentedScanPrecise, KCCopyScanY, KCNormalizeXY
 *   Tracer/Distributions.h   L108-239: DistributionPwC<1> / <2>::SampleIndex / SampleUV / PdfIndex / PdfUV
 *   Tracer/LightsDefault.hpp L173-424: Spherical / CoOcta coordinate converters, LightSkysphere
 *   Core/GraphicsFunctions.h L253-383, L447-468: spherical / concentric-octahedral mappings
 *   Tracer/ColorConverter.cu L405-476: KCExtractLuminance
 *
 * Pinned by the unmodified reference: oracle/ref_build/ref_dist_tap.cpp runs DistributionGroupPwC2D (CPU backend)
 * and the converters; oracle/gen_golden_dist.py commits its outputs as tests/golden/dist2d_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

/* ---- construction: DistributionGroupPwC2D::Construct (Distributions.cu:L437-503) ---- */
/* function: h rows of w values. cdfX: w*h (row CDFs), cdfY: h (marginal over rows). */
void orc_dist2d_build(const float* function, uint32_t w, uint32_t h, float* cdfX, float* cdfY)
{
    /* KCSegmentedScanPrecise: per-row inclusive sum of |f| in double, stored as float */
    for(uint32_t y = 0; y < h; y++)
    {
        double sum = 0.0;
        for(uint32_t x = 0; x < w; x++)
        {
            sum += (double)fabsf(function[(size_t)y * w + x]);
            cdfX[(size_t)y * w + x] = (float)sum;
        }
    }
    /* KCCopyScanY: float running sum of the row totals */
    float s = 0.0f;
    for(uint32_t y = 0; y < h; y++)
    {
        s += cdfX[(size_t)y * w + (w - 1)];
        cdfY[y] = s;
    }
    /* KCNormalizeXY: every row, then the marginal, times 1 / last in double */
    for(uint32_t y = 0; y < h; y++)
    {
        float* row = cdfX + (size_t)y * w;
        double recip = 1.0 / (double)row[w - 1];
        for(uint32_t x = 0; x < w; x++) row[x] = (float)((double)row[x] * recip);
    }
    double recip = 1.0 / (double)cdfY[h - 1];
    for(uint32_t y = 0; y < h; y++) cdfY[y] = (float)((double)cdfY[y] * recip);
}

/* std::lower_bound: first index with cdf[i] >= v (n when none) */
static uint32_t lower_bound_f(const float* cdf, uint32_t n, float v)
{
    uint32_t lo = 0, count = n;
    while(count > 0)
    {
        uint32_t step = count / 2, mid = lo + step;
        if(cdf[mid] < v) { lo = mid + 1; count -= step + 1; }
        else count = step;
    }
    return lo;
}

/* DistributionPwC<1>::SampleIndex (Distributions.h:L118-137) */
static float sample_index_1d(const float* cdf, uint32_t n, float xi, float* pdf)
{
    uint32_t index = lower_bound_f(cdf, n, xi);
    float prev = (index == 0) ? 0.0f : cdf[index - 1];
    float my = cdf[index];
    float t = (xi - prev) / (my - prev);
    float indexF = (float)index + t;
    indexF = (indexF < 1.0f) ? indexF : nextafterf(indexF, -3.402823466e+38f);
    *pdf = (my - prev) * (float)n;
    return indexF;
}

/* DistributionPwC<1>::PdfIndex (L147-157) */
static float pdf_index_1d(const float* cdf, uint32_t n, float index)
{
    uint32_t i = (uint32_t)index;
    float prev = (i == 0) ? 0.0f : cdf[i - 1];
    float my = cdf[i];
    return (my - prev) * (float)n;
}

/* DistributionPwC<2>::SampleUV (L183-213): the marginal picks the row with xi[1], the row's CDF the column with xi[0].
 * out = {u, v, pdf} */
void orc_dist2d_sample_uv(const float* cdfX, const float* cdfY, uint32_t w, uint32_t h, float xi0, float xi1, float out[3])
{
    float pdfY, pdfX;
    float iy = sample_index_1d(cdfY, h, xi1, &pdfY);
    uint32_t row = (uint32_t)iy;
    float ix = sample_index_1d(cdfX + (size_t)row * w, w, xi0, &pdfX);
    out[0] = ix * (1.0f / (float)w);
    out[1] = iy * (1.0f / (float)h);
    out[2] = pdfY * pdfX;
}

/* DistributionPwC<2>::PdfUV (L231-239 + PdfIndex L215-228) */
float orc_dist2d_pdf_uv(const float* cdfX, const float* cdfY, uint32_t w, uint32_t h, float u, float v)
{
    float fx = u * (float)w, fy = v * (float)h;
    float mx = (float)w - 1.0f, my = (float)h - 1.0f;
    fx = fx < mx ? fx : mx; fy = fy < my ? fy : my;
    uint32_t row = (uint32_t)fy;
    float pm = pdf_index_1d(cdfY, h, fy);
    float pc = pdf_index_1d(cdfX + (size_t)row * w, w, fx);
    return pc * pm;
}

void orc_dist2d_sample_many(const float* cdfX, const float* cdfY, uint32_t w, uint32_t h, const float* xi, uint32_t n,
                            float* out /* n * 4: u, v, sample pdf, PdfUV(u, v) */)
{
    for(uint32_t i = 0; i < n; i++)
    {
        float o[3];
        orc_dist2d_sample_uv(cdfX, cdfY, w, h, xi[2 * i], xi[2 * i + 1], o);
        out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2];
        out[4 * i + 3] = orc_dist2d_pdf_uv(cdfX, cdfY, w, h, o[0], o[1]);
    }
}

/* ---- coordinate converters (LightsDefault.hpp:L173-310) ---- */
#define ORC_PI 3.14159265358979323846f
static float sign_pm1(float x) { return copysignf(1.0f, x); }
static float clampf(float x, float a, float b) { return x < a ? a : (x > b ? b : x); }

/* mode 1 = SphericalCoordConverter, 2 = CoOctaCoordConverter; directions are Y-up */
void orc_sky_dir_to_uv(int mode, const float d[3], float uv[2])
{
    /* TransformGen::YUpToZUp: (z, x, y) */
    float zx = d[2], zy = d[0], zz = d[1];
    if(mode == 1)
    {
        float azimuth = atan2f(zy, zx);
        float incl = acosf(clampf(zz, -1.0f, 1.0f));
        uv[0] = (azimuth + ORC_PI) * 0.5f / ORC_PI;
        uv[1] = 1.0f - (incl * (1.0f / ORC_PI));
    }
    else
    {
        const float TwoOvrPi = (1.0f / ORC_PI) * 2.0f;
        if(zx == 0.0f && zy == 0.0f) { uv[0] = uv[1] = 0.0f; return; }
        float xAbs = fabsf(zx), yAbs = fabsf(zy);
        float phiPrime = atanf(yAbs / xAbs);
        float r1 = 1.0f - fabsf(zz);
        float radius = r1 > 0.0f ? sqrtf(r1) : 0.0f;
        float v = radius * TwoOvrPi * phiPrime;
        float u = radius - v;
        if(zz < 0.0f) { float up = 1.0f - v, vp = 1.0f - u; u = up; v = vp; }
        u *= sign_pm1(zx); v *= sign_pm1(zy);
        uv[0] = (u + 1.0f) * 0.5f; uv[1] = (v + 1.0f) * 0.5f;
    }
}

void orc_sky_uv_to_dir(int mode, const float uv[2], float d[3])
{
    float zx, zy, zz;
    if(mode == 1)
    {
        float theta = (uv[0] * ORC_PI * 2.0f) - ORC_PI;
        float phi = (1.0f - uv[1]) * ORC_PI;
        float sT = sinf(theta), cT = cosf(theta), sP = sinf(phi), cP = cosf(phi);
        zx = cT * sP; zy = sT * sP; zz = cP;
    }
    else
    {
        const float PiOvr4 = ORC_PI * 0.25f;
        float u = uv[0] * 2.0f - 1.0f, v = uv[1] * 2.0f - 1.0f;
        float ua = fabsf(u), va = fabsf(v);
        float dd = 1.0f - (ua + va);
        float radius = 1.0f - fabsf(dd);
        float phiPrime = 0.0f;
        if(radius != 0.0f) phiPrime = ((va - ua) / radius + 1.0f) * PiOvr4;
        float sinP = sinf(phiPrime), cosP = cosf(phiPrime);
        float cosPhi = sign_pm1(u) * cosP, sinPhi = sign_pm1(v) * sinP;
        zz = sign_pm1(dd) * (1.0f - radius * radius);
        float xyFactor = radius * sqrtf(2.0f - radius * radius);
        zx = cosPhi * xyFactor; zy = sinPhi * xyFactor;
    }
    /* TransformGen::ZUpToYUp: (y, z, x) */
    d[0] = zy; d[1] = zz; d[2] = zx;
}

/* ToSolidAnglePdf(pdf, dirYUp) and ToSolidAnglePdf(pdf, uv): the spherical map divides by 2 pi^2 sin(inclination) */
float orc_sky_pdf_from_dir(int mode, float pdf, const float d[3])
{
    if(mode != 1) return pdf * 0.25f * (1.0f / ORC_PI);
    float incl = acosf(clampf(d[1], -1.0f, 1.0f));   /* z of the Z-up direction = y of the Y-up one */
    float sinPhi = sinf(incl);
    return (sinPhi <= 0.0f) ? 0.0f : pdf / (2.0f * (ORC_PI * ORC_PI) * sinPhi);
}
float orc_sky_pdf_from_uv(int mode, float pdf, const float uv[2])
{
    if(mode != 1) return pdf * 0.25f * (1.0f / ORC_PI);
    float phi = (1.0f - uv[1]) * ORC_PI;
    float sinPhi = sinf(phi);
    return (sinPhi <= 0.0f) ? 0.0f : pdf / (2.0f * (ORC_PI * ORC_PI) * sinPhi);
}

/* KCExtractLuminance (ColorConverter.cu:L405-476): Y of the texel in the tracer's global colour space; yRow = the
 * second row of Color::Colorspace<E>::ToXYZMatrix, applied as Matrix * Vector = Math::Dot, an FMA chain
 * (Core/Matrix.hpp:L273-284, Core/Math.h:L1586-1596). */
void orc_luminance(const float* rgb, uint32_t n, uint32_t stride, const float yRow[3], float* out)
{
    for(uint32_t i = 0; i < n; i++)
    {
        const float* p = rgb + (size_t)i * stride;
        float r = fmaf(yRow[0], p[0], 0.0f);
        r = fmaf(yRow[1], p[1], r);
        out[i] = fmaf(yRow[2], p[2], r);
    }
}
