/* TEST INFRASTRUCTURE — CPU restatement of the reference's hero-wavelength spectral path
 * (SURVEY.md §8a row 14). Never linked into the product. Parity PINNED: tests/test_oracle_spectrum.py
 * checks every function here against outputs of the UNMODIFIED reference
 * (oracle/ref_build/ref_spectrum_tap.cpp running SpectrumContextJakob2019 on its CPU backend),
 * committed as tests/golden/spectrum_mode*.npz.
 *
 * Follows:
 *   SingleSampleSpectrumWavelength     Tracer/SpectrumContext.cu:L14-135
 *   ConvertSpectraToRGBSingle          Tracer/SpectrumContext.cu:L137-171
 *   Converter::ConvertAlbedo/Radiance  Tracer/SpectrumContext.hpp:L37-150
 *   TextureViewCPU linear fetch        Device/CPU/TextureViewCPU.h:L258-376 (software lerp)
 *   RNGFunctions::ToFloat01            Tracer/Random.h:L102-118
 *   Distribution::Common::*            Tracer/DistributionFunctions.h:L627-641,L686-711,L811-818
 *   Math::InvErrFunc / Gaussian / InvSmoothstep   Core/Math.h:L788-840,L1032-1042,L1225-1232
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define CIE_START 360
#define CIE_N 471

typedef struct orc_spectrum_tables
{
    const float* lut;        /* 9 * N^3: table t (max channel) holds c0,c1,c2 blocks at (3t+k) * N^3; index z*N*N + y*N + x */
    uint32_t     n;          /* 64 */
    const float* observer;   /* CIE_N * 3, normalised by the X/Y/Z integrals */
    const float* illuminant; /* CIE_N, normalised */
    float        xyzToRGB[9];
} orc_spectrum_tables;

static float to_float01(uint32_t v)
{
    float f = (float)v * 0x1.p-32f;
    const float prev1 = 0x1.fffffep-1f;
    return f < prev1 ? f : prev1;
}

static float inv_erf(float x)
{
    float t = logf(fmaf(x, 0.0f - x, 1.0f));
    float p;
    if(fabsf(t) > 6.125f)
    {
        p = 3.03697567e-10f;
        p = fmaf(p, t, 2.93243101e-8f); p = fmaf(p, t, 1.22150334e-6f); p = fmaf(p, t, 2.84108955e-5f);
        p = fmaf(p, t, 3.93552968e-4f); p = fmaf(p, t, 3.02698812e-3f); p = fmaf(p, t, 4.83185798e-3f);
        p = fmaf(p, t, -2.64646143e-1f); p = fmaf(p, t, 8.40016484e-1f);
    }
    else
    {
        p = 5.438778320e-9f;
        p = fmaf(p, t, 1.43285448e-7f); p = fmaf(p, t, 1.22774793e-6f); p = fmaf(p, t, 1.12963626e-7f);
        p = fmaf(p, t, -5.61530760e-5f); p = fmaf(p, t, -1.47697632e-4f); p = fmaf(p, t, 2.31468678e-3f);
        p = fmaf(p, t, 1.15392581e-2f); p = fmaf(p, t, -2.32015476e-1f); p = fmaf(p, t, 8.86226892e-1f);
    }
    return p * x;
}

static float gaussian(float x, float sigma, float mu)
{
    const float invSqrt2Pi = (1.0f / 1.41421356237309504880f) * (1.0f / 1.77245385090551602729f);
    float sigmaInv = 1.0f / sigma;
    float result = invSqrt2Pi * sigmaInv;
    float pw = (x - mu) * sigmaInv;
    result *= expf(-0.5f * pw * pw);
    /* the reference's worker threads run flush-to-zero (the golden vectors hold 0 where IEEE gives a
     * denormal); a denormal pdf would turn DivideByPDF into inf */
    if(result < 1.17549435e-38f) result = 0.0f;
    return result;
}

/* mode: 0 Uniform, 1 GaussianMIS, 2 HyperbolicPBRT; one random number per sample */
void orc_sample_wavelengths(int mode, const uint32_t* randoms, uint32_t n, float* waves /* n*4 */, float* pdfs /* n*4 */)
{
    const float START = (float)CIE_START, END = (float)(CIE_START + CIE_N - 1);
    const float offsets[4] = {-0.5f, -0.25f, 0.0f, 0.25f};
    for(uint32_t s = 0; s < n; s++)
    {
        float xi0 = to_float01(randoms[s]);
        float xi[4];
        for(int i = 0; i < 4; i++)
        {
            float x = xi0 + offsets[i];
            if(x < 0.0f) x += 1.0f;
            if(x >= 1.0f) x -= 1.0f;
            xi[i] = x;
        }
        float* w = waves + 4 * (size_t)s; float* p = pdfs + 4 * (size_t)s;
        if(mode == 0)
        {
            for(int i = 0; i < 4; i++) { w[i] = xi[i] * (END - START) + START; p[i] = 1.0f / (END - START); }
        }
        else if(mode == 1)
        {
            const float SIGMA[2] = {25.0f, 48.0f}, MU[2] = {452.0f, 576.0f}, MIS[2] = {0.384615384615f, 0.615384615385f};
            for(int i = 0; i < 4; i++)
            {
                float wgt = MIS[0];
                int si = (xi[i] < wgt) ? 0 : 1;
                float lxi = (xi[i] < wgt) ? xi[i] / wgt : (xi[i] - wgt) / (1.0f - wgt);
                lxi = fminf(lxi, 0x1.fffffep-1f);
                int oi = (si + 1) & 1;
                float x = 1.41421356237309504880f * SIGMA[si];
                float e = inv_erf(2.0f * lxi - 1.0f);
                x = x * e + MU[si];
                if(isinf(e)) { float mm = 3.5f * SIGMA[si]; x = fminf(fmaxf(x, -mm), mm); }
                float pdfS = gaussian(x, SIGMA[si], MU[si]);
                float pdfO = gaussian(x, SIGMA[oi], MU[oi]);
                w[i] = x;
                p[i] = 0.0f + pdfS * MIS[si] + pdfO * MIS[oi];
            }
        }
        else
        {
            for(int i = 0; i < 4; i++)
            {
                float a = 0.85691062f - 1.82750197f * xi[i];
                w[i] = 538.0f - 138.888889f * atanhf(a);
                float d = coshf(0.0072f * (w[i] - 538.0f));
                p[i] = 0.0039398042f / (d * d);
            }
        }
    }
}

static float lerpf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

/* TextureViewCPU<1,...>, unnormalised coordinates, linear, clamp */
static void interp1(float x, uint32_t size, int* i0, int* i1, float* frac)
{
    float uv = x / (float)size;
    float texel = uv * (float)size - 0.5f;
    float base; float fr = modff(texel, &base);
    int start = (int)base;
    if(fr < 0.0f) { start -= 1; fr = fabsf(fr); }
    int a = start, b = start + 1;
    if(a < 0) a = 0; if(a > (int)size - 1) a = (int)size - 1;
    if(b < 0) b = 0; if(b > (int)size - 1) b = (int)size - 1;
    *i0 = a; *i1 = b; *frac = fr;
}

static float fetch_illuminant(const orc_spectrum_tables* t, float x)
{
    int a, b; float f; interp1(x, CIE_N, &a, &b, &f);
    return lerpf(t->illuminant[a], t->illuminant[b], f);
}

/* TextureViewCPU<3,Float>, normalised coordinates, linear, clamp */
static float fetch3(const float* tex, uint32_t n, const float uv[3])
{
    int st[3]; float fr[3];
    for(int d = 0; d < 3; d++)
    {
        float texel = uv[d] * (float)n - 0.5f;
        float base; float f = modff(texel, &base);
        st[d] = (int)base;
        if(f < 0.0f) { st[d] -= 1; f = fabsf(f); }
        fr[d] = f;
    }
    float pix[8];
    for(int k = 0; k < 2; k++) for(int j = 0; j < 2; j++) for(int i = 0; i < 2; i++)
    {
        int x = st[0] + i, y = st[1] + j, z = st[2] + k;
        if(x < 0) x = 0; if(x > (int)n - 1) x = (int)n - 1;
        if(y < 0) y = 0; if(y > (int)n - 1) y = (int)n - 1;
        if(z < 0) z = 0; if(z > (int)n - 1) z = (int)n - 1;
        pix[(k << 2) + (j << 1) + i] = tex[(size_t)z * n * n + (size_t)y * n + (size_t)x];
    }
    for(int pass = 3; pass > 0; pass--)
        for(int i = 0; i < (1 << (pass - 1)); i++)
            pix[i] = lerpf(pix[2 * i], pix[2 * i + 1], fr[3 - pass]);
    return pix[0];
}

static float inv_smoothstep(float y) { return 0.5f - sinf(asinf(1.0f - 2.0f * y) * (1.0f / 3.0f)); }

void orc_convert_albedo(const orc_spectrum_tables* t, const float rgb[3], const float waves[4], float out[4])
{
    int maxI = 0; float mx = rgb[0];
    for(int i = 1; i < 3; i++) if(rgb[i] > mx) { mx = rgb[i]; maxI = i; }
    float xyz[3] = {0.f, 0.f, 0.f};
    if(mx > 1.0e-7f)
    {
        float f = 1.0f / mx;
        xyz[0] = rgb[(maxI + 1) % 3] * f; xyz[1] = rgb[(maxI + 2) % 3] * f;
    }
    float slice = fminf(fmaxf(mx, 0.0f), 1.0f);
    xyz[2] = inv_smoothstep(inv_smoothstep(slice));
    const float N = (float)t->n, A = (N - 1.0f) / N, B = 0.5f / N;
    float uv[3] = {xyz[0] * A + B, xyz[1] * A + B, xyz[2] * A + B};
    const size_t n3 = (size_t)t->n * t->n * t->n;
    float c0 = fetch3(t->lut + (3 * maxI + 0) * n3, t->n, uv);
    float c1 = fetch3(t->lut + (3 * maxI + 1) * n3, t->n, uv);
    float c2 = fetch3(t->lut + (3 * maxI + 2) * n3, t->n, uv);
    for(int i = 0; i < 4; i++)
    {
        float tt = fmaf(c0, waves[i], c1);
        float x = fmaf(tt, waves[i], c2);
        float dr = 1.0f / sqrtf(fmaf(x, x, 1.0f));
        out[i] = fmaf(0.5f * x, dr, 0.5f);
    }
}

void orc_convert_radiance(const orc_spectrum_tables* t, const float radiance[3], const float waves[4], float out[4])
{
    float mx = radiance[0];
    for(int i = 1; i < 3; i++) if(radiance[i] > mx) mx = radiance[i];
    float scale = mx * 2.0f;
    float rgb[3] = {0.f, 0.f, 0.f};
    if(scale != 0.0f) for(int i = 0; i < 3; i++) rgb[i] = radiance[i] / scale;
    orc_convert_albedo(t, rgb, waves, out);
    const float OFFSET = 0.5f - (float)CIE_START;
    for(int i = 0; i < 4; i++) out[i] *= fetch_illuminant(t, waves[i] + OFFSET);
    for(int i = 0; i < 4; i++) out[i] *= scale;
}

void orc_illuminant(const orc_spectrum_tables* t, const float waves[4], float out[4])
{
    const float OFFSET = 0.5f - (float)CIE_START;
    for(int i = 0; i < 4; i++) out[i] = fetch_illuminant(t, waves[i] + OFFSET);
}

/* value (4 spectral samples) -> rgb (out[3] = 0); dispersed waves are not produced by this path */
void orc_spectra_to_rgb(const orc_spectrum_tables* t, const float value[4], const float waves[4], const float pdf[4], float out[4])
{
    const float OFFSET = 0.5f - (float)CIE_START;
    float xyz[3] = {0.f, 0.f, 0.f};
    /* a dispersed path ((Mt)Refract) keeps its first wavelength only: secondary waves are DISPERSED_WAVE = -1, one
     * sample counts and the 1/4 weight becomes 1 (SpectrumContext.cu:L148-157, TracerTypes.h:L114,L386-401) */
    const int dispersed = waves[1] == -1.0f;
    for(int i = 0; i < (dispersed ? 1 : 4); i++)
    {
        int a, b; float f; interp1(waves[i] + OFFSET, CIE_N, &a, &b, &f);
        float val = (pdf[i] == 0.0f) ? 0.0f : value[i] / pdf[i];
        for(int c = 0; c < 3; c++)
            xyz[c] += lerpf(t->observer[3 * a + c], t->observer[3 * b + c], f) * val;
    }
    for(int c = 0; c < 3; c++) xyz[c] *= dispersed ? 1.0f : 0.25f;
    for(int r = 0; r < 3; r++)
        out[r] = t->xyzToRGB[3 * r] * xyz[0] + t->xyzToRGB[3 * r + 1] * xyz[1] + t->xyzToRGB[3 * r + 2] * xyz[2];
    out[3] = 0.0f;
}

/* batched drivers for the tests */
void orc_convert_batch(const orc_spectrum_tables* t, const float rgb[3], const float* waves, const float* pdfs, uint32_t n,
                       float radianceScale, float* albedoSpec, float* radianceSpec, float* rgbAlbedoIllum, float* rgbRadiance)
{
    float rad[3] = {rgb[0] * radianceScale, rgb[1] * radianceScale, rgb[2] * radianceScale};
    for(uint32_t s = 0; s < n; s++)
    {
        const float* w = waves + 4 * (size_t)s; const float* p = pdfs + 4 * (size_t)s;
        orc_convert_albedo(t, rgb, w, albedoSpec + 4 * (size_t)s);
        orc_convert_radiance(t, rad, w, radianceSpec + 4 * (size_t)s);
        float ill[4], v[4];
        orc_illuminant(t, w, ill);
        for(int i = 0; i < 4; i++) v[i] = albedoSpec[4 * (size_t)s + i] * ill[i];
        orc_spectra_to_rgb(t, v, w, p, rgbAlbedoIllum + 4 * (size_t)s);
        orc_spectra_to_rgb(t, radianceSpec + 4 * (size_t)s, w, p, rgbRadiance + 4 * (size_t)s);
    }
}
