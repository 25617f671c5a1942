/* pt_oracle.c — CPU restatement of the reference's path-tracing ESTIMATOR (RGB, Lambert, prim-backed
 * triangle lights, uniform light sampler incl. the boundary light, NEE / NEE+MIS / pure, Russian
 * roulette, stochastic Gaussian film filter), one path at a time instead of a wavefront.
 *
 * TEST INFRASTRUCTURE ONLY (see mray_oracle.c). Parity status: PINNED by reference execution — the unmodified
 * reference (CPU backend, driven through TracerI by oracle/ref_build/tracer_driver.cpp) rendered the Cornell box
 * in all three sample modes, as a two-level scene and with the spectral renderer (oracle/gen_golden_render.py ->
 * tests/golden/render_*.npz); tests/test_oracle_pt.py holds this restatement to those images at the north-star
 * tolerance (relMSE <= 1e-3 on converged images), plus (a) Pure / NEE / NEE+MIS agreeing in expectation and
 * (b) the closed-form direct irradiance under a square light.
 * Statistical parity only (even the reference's own backends differ per sample).
 *
 * Restated from (paths relative to /root/reference/Source):
 *   TracerDLL/PathTracerRendererShaders.h:L201-301  WorkFunction::Call        (BxDF sample, RR)
 *   ...:L352-443                                    WorkFunctionNEE::Call     (light sample, shadow ray, MIS)
 *   ...:L306-347, L448-520                          LightWorkFunction[WithNEE]::Call
 *   TracerDLL/PathTracerRenderer.cu:L11-32          KCAccumulateShadowRaysPT  (depth + 2 <= rrRange[1])
 *   Tracer/MaterialsDefault.hpp:L25-126             LambertMaterial
 *   Tracer/LightsDefault.hpp:L22-168                LightPrim (SampleSolidAngle, PdfSolidAngle, EmitVia*)
 *   Tracer/LightSampler.hpp:L5-119                  SampledRay, DirectLightSamplerUniform
 *   Tracer/PrimitiveDefaultTriangle.hpp:L48-77      Triangle::SampleSurface (Osada)
 *   Tracer/DistributionFunctions.h:L847-871,L943-965 SampleCosDirection, RussianRoulette, BalanceCancelled
 *   Tracer/CamerasDefault.hpp:L8-36,L93-142         CameraPinhole ctor / EvaluateRay
 *   Tracer/Filters.h:L195-227, DistributionFunctions.h:L686-705  Gaussian filter sample / evaluate
 *   Tracer/Random.h:L237-238,L763-812               PermutedCG32
 *   Core/Ray.hpp:L258-301                           Ray::Nudge
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

/* from mray_oracle.c */
void orc_lbvh_trace(const float* pos, const uint32_t* idx, const uint32_t* nodes, const float* boxes,
                    const float* rays, uint32_t nRays, int mode, int cullFace,
                    uint32_t* outPrim, float* outT, float* outBary, uint8_t* outBack);

typedef struct { float x, y, z; } v3;
static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float len(v3 a) { return sqrtf(dot(a, a)); }
static v3 nrm(v3 a) { return mul(a, 1.0f / len(a)); }

typedef struct { uint32_t s; } pcg;
static uint32_t pcg_next(pcg* r)
{
    uint32_t old = r->s;
    r->s = old * 747796405u + 2891336453u;
    uint32_t v = ((old >> ((old >> 28u) + 4u)) ^ old) * 277803737u;
    return (v >> 22u) ^ v;
}
static float pcg_float(pcg* r) { float f = (float)pcg_next(r) * 0x1.p-32f; return f < 0.99999994f ? f : 0.99999994f; }

static v3 nudge(v3 p, v3 n)
{
    const float ORIGIN = 1.0f / 32.0f, FLOAT_SCALE = 1.0f / 65536.0f, INT_SCALE = 256.0f;
    float pin[3] = {p.x, p.y, p.z}, nin[3] = {n.x, n.y, n.z}, out[3];
    for(int k = 0; k < 3; k++)
    {
        int32_t of = (int32_t)(INT_SCALE * nin[k]);
        int32_t pi; memcpy(&pi, &pin[k], 4);
        pi += (pin[k] < 0.0f) ? -of : of;
        float pf; memcpy(&pf, &pi, 4);
        out[k] = (fabsf(pin[k]) < ORIGIN) ? pin[k] + FLOAT_SCALE * nin[k] : pf;
    }
    return V(out[0], out[1], out[2]);
}

typedef struct
{
    const float* pos; const uint32_t* idx; uint32_t nTris;
    const uint32_t* nodes; const float* boxes;          /* binary LBVH of the whole scene */
    const int32_t* triMaterial;                         /* >= 0 Lambert material index, < 0 : light index = -1 - v */
    const float* albedo; const float* radiance; const uint8_t* twoSided;
    const uint32_t* lightTris; uint32_t nLightTris;     /* emissive triangle list (one meta light each) */
    float camPos[3], camGaze[3], camUp[3], fovXY[2], nearFar[2];
    uint32_t width, height, spp, sampleMode, rrLo, rrHi;
    float filterRadius; uint64_t seed;
    /* (R)PathTracerSpectral when non-NULL: tables of spectrum_oracle.c + WavelengthSampleMode */
    const struct orc_spectrum_tables* spectrum; uint32_t wavelengthMode;
    /* textured Lambert albedo (all NULL / 0 = none): per-vertex UV0, texture table, per-material texture index or -1 */
    const float* uv; const struct orc_texture* textures; const int32_t* albedoTexture; uint32_t nTextures;
    /* per material: 0 = (Mt)Lambert, 1 = (Mt)Reflect (NULL = all Lambert) */
    const uint8_t* materialType;
    /* TracerParameters.filmFilter.type: 0 = the default (Gaussian), else FilterType::E + 1 (1 Box, 2 Tent, 3 Gaussian,
     * 4 Mitchell-Netravali); the radius is filterRadius */
    uint32_t filmFilter;
    /* per material 8 floats (NULL = none): materialType 2 = (Mt)Refract {cauchyFront xyz, -, cauchyBack xyz, -},
     * 3 = (Mt)Unreal {roughness, specular, metallic, ...} */
    const float* materialParams;
    /* per vertex world -> tangent-space quaternion (w, x, y, z), the triangle group's NORMAL attribute; NULL = geometric
     * normals */
    const float* vertexTBN;
    /* boundary light surface: 0 = (L)Null, 1 = (L)Skysphere_Spherical, 2 = (L)Skysphere_CoOcta (LightsDefault.hpp:L310-443).
     * boundaryTexture: -1 = constant boundaryRadiance, else an index into `textures` (the radiance map) whose luminance
     * distribution is boundaryCdfX / boundaryCdfY (dist_oracle.c: orc_dist2d_build of orc_luminance). boundaryM / boundaryInvM:
     * linear part of the light surface's transform and its inverse (row-major 3x3). */
    uint32_t boundaryType; int32_t boundaryTexture; float boundaryRadiance[3];
    const float* boundaryCdfX; const float* boundaryCdfY;
    float boundaryM[9], boundaryInvM[9]; float sceneDiameter;
    /* alpha maps (SurfaceParams.alphaMaps): per triangle -1 or an index into `textures` whose first channel is the alpha
     * (NULL = none); uv = the per-vertex UV0 above */
    const int32_t* triAlpha;
    /* normal maps: per material -1 or an index into `textures` holding tangent-space normals (NULL = none); needs vertexTBN */
    const int32_t* normalTexture;
    /* how a textured read turns ray-cone gradients into a mip level (orc_texture_sample_grad): 0 = host backend, 1 = tex2DGrad */
    uint32_t textureLodMode;
} pt_scene;

/* One single-level 2-D texture as the reference's host-backend view reads it (Device/CPU/TextureViewCPU.h):
 * format 0 = fp32, 1 = unorm8; interp 0 = nearest, 1 = linear; edge 0 = wrap, 1 = clamp, 2 = mirror. */
struct orc_texture { const void* data; uint32_t w, h, channels, format, interp, edge, mipCount; };
/* `data` holds mipCount (0 reads as 1) levels back to back, level k at pixel offset TextureMipPixelStart(size, k) with
 * TextureMipSize(size, k) = max(size >> k, 1) texels per axis (Core/GraphicsFunctions.h:L474-525): the layout of the
 * reference's host-backend texture. */
static uint32_t mip_dim(uint32_t n, uint32_t level) { uint32_t v = n >> level; return v ? v : 1u; }
static size_t mip_start(uint32_t w, uint32_t h, uint32_t level)
{ size_t o = 0; for(uint32_t i = 0; i < level; i++) o += (size_t)mip_dim(w, i) * mip_dim(h, i); return o; }
uint32_t orc_texture_mip_count(uint32_t w, uint32_t h)
{ uint32_t m = w > h ? w : h, c = 0; while(m) { c++; m >>= 1; } return c; }   /* Bit::RequiredBitsToRepresent(max dim) */

/* TextureViewCPU::ResolveEdge (TextureViewCPU.h:L196-246); C's / and % truncate like the reference's */
static int tex_edge(int i, int n, uint32_t edge)
{
    if(edge == 1u) return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    if(edge == 2u)
    {
        int dim = i / n;
        i = i % n;
        if(i < 0) i += n;
        if((dim & 1) == 1) i = n - i;
        return i > n - 1 ? n - 1 : i;   /* the reference can produce n here (out of bounds there) */
    }
    i = i % n;
    if(i < 0) i += n;
    return i;
}
/* ReadPixel + Convert (L120-170,L305-340): FromUNorm = v * (1 / 255) */
static void tex_pixel_level(const struct orc_texture* t, uint32_t level, int x, int y, float out[3])
{
    size_t o = (mip_start(t->w, t->h, level) + (size_t)y * mip_dim(t->w, level) + (size_t)x) * t->channels;
    const uint32_t nc = t->channels < 3u ? t->channels : 3u;   /* alpha maps are single-channel: missing channels read 0 */
    out[0] = out[1] = out[2] = 0.0f;
    if(t->format == 0u) { const float* f = (const float*)t->data + o; for(uint32_t k = 0; k < nc; k++) out[k] = f[k]; return; }
    const uint8_t* b = (const uint8_t*)t->data + o;
    const float DELTA = 1.0f / 255.0f;
    for(uint32_t k = 0; k < nc; k++) out[k] = (float)b[k] * DELTA;
}
/* Math::Lerp (Core/Math.h): a * (1 - t) + b * t, unfused */
static float tex_lerp(float a, float b, float t) { volatile float x = a * (1.0f - t); volatile float y = b * t; return x + y; }
/* TracerTexView<2, Vector3>::operator()(uv, dpdx, dpdy) for a texture with ONE mip level: the mip level computed
 * from the gradients clamps to 0 (L429-470), leaving NearestPixel (L248-260) or FindInterpolants +
 * ReadInterpolatedPixel (L262-300,L352-385) on the base level. */
static void texture_sample_level(const struct orc_texture* t, uint32_t level, float u, float v, float out[3])
{
    const uint32_t w = mip_dim(t->w, level), h = mip_dim(t->h, level);
    float tu = u * (float)w, tv = v * (float)h;
    if(t->interp == 0u)
    {
        int x = (int)roundf(tu - 0.5f), y = (int)roundf(tv - 0.5f);
        tex_pixel_level(t, level, tex_edge(x, (int)w, t->edge), tex_edge(y, (int)h, t->edge), out);
        return;
    }
    float bx, by;
    float fx = modff(tu - 0.5f, &bx), fy = modff(tv - 0.5f, &by);
    int x0 = (int)bx, y0 = (int)by;
    if(fx < 0.0f) { x0 -= 1; fx = fabsf(fx); }   /* as the reference: |frac|, not 1 - |frac| */
    if(fy < 0.0f) { y0 -= 1; fy = fabsf(fy); }
    int xa = tex_edge(x0, (int)w, t->edge), xb = tex_edge(x0 + 1, (int)w, t->edge);
    int ya = tex_edge(y0, (int)h, t->edge), yb = tex_edge(y0 + 1, (int)h, t->edge);
    float p00[3], p10[3], p01[3], p11[3];
    tex_pixel_level(t, level, xa, ya, p00); tex_pixel_level(t, level, xb, ya, p10);
    tex_pixel_level(t, level, xa, yb, p01); tex_pixel_level(t, level, xb, yb, p11);
    for(int k = 0; k < 3; k++)
        out[k] = tex_lerp(tex_lerp(p00[k], p10[k], fx), tex_lerp(p01[k], p11[k], fx), fy);
}
void orc_texture_sample(const struct orc_texture* t, float u, float v, float out[3]) { texture_sample_level(t, 0u, u, v, out); }
/* TextureViewCPU::operator()(uv, mipLevel) (TextureViewCPU.h:L422-470): level clamped to [0, mipCount - 1], ModFInt; LINEAR =
 * Math::Lerp of the two levels' bilinear reads; NEAREST = the nearer level's nearest texel (the reference resolves the edge
 * of that read against the BASE size, L446 — an out-of-range index there; here the level's own size). */
void orc_texture_sample_lod(const struct orc_texture* t, float u, float v, float mipLevel, float out[3])
{
    const uint32_t mc = t->mipCount ? t->mipCount : 1u;
    const float top = (float)(mc - 1u);
    mipLevel = mipLevel < 0.0f ? 0.0f : (mipLevel > top ? top : mipLevel);   /* Math::Clamp; NaN (log2 of 0 * ...) cannot occur: -inf clamps */
    if(!(mipLevel >= 0.0f)) mipLevel = 0.0f;
    float ip; const float frac = modff(mipLevel, &ip);
    const uint32_t m0 = (uint32_t)ip, m1 = (m0 + 1u > mc - 1u) ? mc - 1u : m0 + 1u;
    if(t->interp == 0u) { texture_sample_level(t, frac < 0.5f ? m0 : m1, u, v, out); return; }
    float a[3]; texture_sample_level(t, m0, u, v, a);
    if(m0 == m1) { out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; return; }
    float b[3]; texture_sample_level(t, m1, u, v, b);
    for(int k = 0; k < 3; k++) out[k] = tex_lerp(a[k], b[k], frac);
}
/* TextureViewCPU::operator()(uv, dpdx, dpdy) (L405-420): level = 0.5 log2(max |gradient|^2). lodMode 0 = as the reference's host
 * backend reads a normalised-coordinate texture (the gradients stay in UV units); 1 = as the device backends' tex2DGrad
 * (gradients scaled by the base size first). */
void orc_texture_sample_grad(const struct orc_texture* t, float u, float v, const float dpdx[2], const float dpdy[2], uint32_t lodMode, float out[3])
{
    float ax = dpdx[0], ay = dpdx[1], bx = dpdy[0], by = dpdy[1];
    if(lodMode == 1u) { ax *= (float)t->w; ay *= (float)t->h; bx *= (float)t->w; by *= (float)t->h; }
    const float la = fmaf(ay, ay, ax * ax), lb = fmaf(by, by, bx * bx);   /* Math::LengthSqr = Dot: an FMA chain (Core/Math.h:L1586-1596) */
    const float m = la > lb ? la : lb;
    orc_texture_sample_lod(t, u, v, 0.5f * log2f(m), out);
}

static float gauss_pdf_mu(float x, float sig, float mu);
static float lerp_u(float a, float b, float t);
static float mitchell_1d(float x, float rr);
/* <Filter>::Evaluate(duv) of Tracer/Filters.h (L110-117 Box, L153-164 Tent, L202-206 Gaussian, L257-278 Mitchell-Netravali) */
static float filter_evaluate(uint32_t type, float r, float x, float y)
{
    if(type == 0u) { float rr = 1.0f / r; return (fabsf(x) <= r && fabsf(y) <= r) ? 0.25f * rr * rr : 0.0f; }
    if(type == 1u) { float rcp = 1.0f / r, cap = 1.0f / r; return lerp_u(cap, 0.0f, fabsf(x * rcp)) * lerp_u(cap, 0.0f, fabsf(y * rcp)); }
    if(type == 3u) { float rcp = 1.0f / r; return mitchell_1d(x, rcp) * mitchell_1d(y, rcp); }
    const float sigma = r * 0.285714f;
    return gauss_pdf_mu(x, sigma, 0.0f) * gauss_pdf_mu(y, sigma, 0.0f);
}
/* TextureMemory::GenerateMipmaps -> KCGenerateMipmaps (Tracer/TextureFilter.cu:L55-118,L126-198,L1064-1096): every level k >=
 * firstLevel of `chain` (a full chain buffer whose levels < firstLevel are valid) is filtered from level k - 1: 8 x 8 stratified
 * offsets over [-r, r]^2 (FilterMode::ACCUMULATE), weight = Evaluate(offset), the parent texel nearest to the offset pixel centre
 * (ConvertPixelIndices + RoundInt), sum / weight sum. fp32 texels are written as they are, unorm8 texels — filtered as their
 * 0..255 integer values — are rounded and clamped (GenericWrite). */
void orc_texture_generate_mips(void* chain, uint32_t w, uint32_t h, uint32_t channels, uint32_t format, uint32_t firstLevel,
                               uint32_t mipCount, uint32_t filterType, float radius)
{
    for(uint32_t level = firstLevel ? firstLevel : 1u; level < mipCount; level++)
    {
        const uint32_t mw = mip_dim(w, level), mh = mip_dim(h, level), pw = mip_dim(w, level - 1u), ph = mip_dim(h, level - 1u);
        const size_t dst = mip_start(w, h, level), src = mip_start(w, h, level - 1u);
        for(uint32_t y = 0; y < mh; y++) for(uint32_t x = 0; x < mw; x++)
        {
            float acc[4] = {0, 0, 0, 0}, wsum = 0.0f;
            for(uint32_t sy = 0; sy < 8u; sy++) for(uint32_t sx = 0; sx < 8u; sx++)
            {
                const float dxy = 1.0f / 8.0f;
                const float xi0 = dxy * 0.5f + dxy * (float)sx, xi1 = dxy * 0.5f + dxy * (float)sy;
                const float ox = xi0 * 2.0f * radius - radius, oy = xi1 * 2.0f * radius - radius;
                const float wgt = filter_evaluate(filterType, radius, ox, oy);
                /* ConvertPixelIndices(pix + offset, parentRes, mipRes) */
                float rx = ((float)x + ox + 0.5f) * ((float)pw / (float)mw) - 0.5f, ry = ((float)y + oy + 0.5f) * ((float)ph / (float)mh) - 0.5f;
                rx = rx < 0.0f ? 0.0f : (rx > (float)pw - 1.0f ? (float)pw - 1.0f : rx);
                ry = ry < 0.0f ? 0.0f : (ry > (float)ph - 1.0f ? (float)ph - 1.0f : ry);
                const size_t o = (src + (size_t)lroundf(ry) * pw + (size_t)lroundf(rx)) * channels;
                for(uint32_t c = 0; c < channels; c++)
                {
                    const float px = format == 0u ? ((const float*)chain)[o + c] : (float)((const uint8_t*)chain)[o + c];
                    volatile float term = wgt * px;
                    acc[c] += term;
                }
                wsum += wgt;
            }
            const size_t o = (dst + (size_t)y * mw + x) * channels;
            for(uint32_t c = 0; c < channels; c++)
            {
                const float v = acc[c] / wsum;
                if(format == 0u) ((float*)chain)[o + c] = v;
                else { float r = roundf(v); r = r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r); ((uint8_t*)chain)[o + c] = (uint8_t)r; }
            }
        }
    }
}

void orc_pt_filter_sample_typed(uint32_t type, float radius, float xi0, float xi1, float out[5]);
/* TracerParameters.clampedTexRes (TextureMemory::CreateTexture, Tracer/TextureMemory.cpp:L544-583): levels dropped so that the larger
 * side fits `clampRes` = ceil(log2(ceil(maxDim / min(clampRes, maxDim)))) */
uint32_t orc_texture_clamp_levels(uint32_t w, uint32_t h, uint32_t clampRes)
{
    const uint32_t maxDim = w > h ? w : h, c = clampRes < maxDim ? clampRes : maxDim;
    if(c == 0u) return 0u;
    const uint32_t ratio = (maxDim + c - 1u) / c;
    return (uint32_t)(int32_t)ceilf(log2f((float)ratio));
}
/* ClampImageFromBuffer -> KCClampImage (Tracer/TextureFilter.cu:L206-262,L1100-1150): the image `src` (sw x sh) filtered down to
 * dw x dh: per texel 4 x 4 stratified numbers through the mip filter's Sample() (FilterMode::SAMPLING), weight = Evaluate / pdf / 16,
 * the source texel nearest to the offset pixel centre (ConvertPixelIndices + Math::Round), sum / weight sum. */
void orc_texture_clamp(const void* src, uint32_t sw, uint32_t sh, void* dst, uint32_t dw, uint32_t dh, uint32_t channels, uint32_t format,
                       uint32_t filterType, float radius)
{
    for(uint32_t y = 0; y < dh; y++) for(uint32_t x = 0; x < dw; x++)
    {
        float acc[4] = {0, 0, 0, 0}, wsum = 0.0f;
        for(uint32_t sy = 0; sy < 4u; sy++) for(uint32_t sx = 0; sx < 4u; sx++)
        {
            const float dxy = 1.0f / 4.0f, inv = dxy * dxy;
            const float xi0 = dxy * 0.5f + dxy * (float)sx, xi1 = dxy * 0.5f + dxy * (float)sy;
            float fo[5]; orc_pt_filter_sample_typed(filterType, radius, xi0, xi1, fo);
            const float wgt = fo[4], pdf = fo[2];
            float rx = ((float)x + fo[0] + 0.5f) * ((float)sw / (float)dw) - 0.5f, ry = ((float)y + fo[1] + 0.5f) * ((float)sh / (float)dh) - 0.5f;
            rx = rx < 0.0f ? 0.0f : (rx > (float)sw - 1.0f ? (float)sw - 1.0f : rx);
            ry = ry < 0.0f ? 0.0f : (ry > (float)sh - 1.0f ? (float)sh - 1.0f : ry);
            const size_t o = ((size_t)(uint32_t)roundf(ry) * sw + (size_t)(uint32_t)roundf(rx)) * channels;
            for(uint32_t c = 0; c < channels; c++)
            {
                const float px = format == 0u ? ((const float*)src)[o + c] : (float)((const uint8_t*)src)[o + c];
                volatile float t0 = wgt * px; volatile float t1 = t0 * inv; volatile float t2 = t1 / pdf;
                acc[c] += t2;
            }
            volatile float w0 = wgt * inv; volatile float w1 = w0 / pdf;
            wsum += w1;
        }
        const size_t o = ((size_t)y * dw + x) * channels;
        for(uint32_t c = 0; c < channels; c++)
        {
            const float v = acc[c] / wsum;
            if(format == 0u) ((float*)dst)[o + c] = v;
            else { float r = roundf(v); r = r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r); ((uint8_t*)dst)[o + c] = (uint8_t)r; }
        }
    }
}

/* spectrum_oracle.c */
struct orc_spectrum_tables;
void orc_sample_wavelengths(int mode, const uint32_t* randoms, uint32_t n, float* waves, float* pdfs);
void orc_convert_albedo(const struct orc_spectrum_tables* t, const float rgb[3], const float waves[4], float out[4]);
void orc_convert_radiance(const struct orc_spectrum_tables* t, const float radiance[3], const float waves[4], float out[4]);
void orc_spectra_to_rgb(const struct orc_spectrum_tables* t, const float value[4], const float waves[4], const float pdf[4], float out[4]);

/* Spectrum = 4 floats: (r, g, b, 0) for the RGB renderer, 4 hero-wavelength samples for the spectral one */
typedef struct { float v[4]; } s4;
static s4 S(float a, float b, float c, float d) { s4 r = {{a, b, c, d}}; return r; }
static s4 s_mul(s4 a, float k) { return S(a.v[0] * k, a.v[1] * k, a.v[2] * k, a.v[3] * k); }
static s4 s_mulv(s4 a, s4 b) { return S(a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]); }
static s4 s_add(s4 a, s4 b) { return S(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]); }
/* LambertMaterial ctor (MaterialsDefault.hpp:L17-25): albedo = ConvertAlbedo(albedoMap(uv, dpdx, dpdy)) */
void orc_texture_sample_grad(const struct orc_texture* t, float u, float v, const float dpdx[2], const float dpdy[2], uint32_t lodMode, float out[3]);
static s4 albedo_at(const pt_scene* s, uint32_t m, const float waves[4], uint32_t prim, float a, float b, float c, const float dpdx[2], const float dpdy[2])
{
    float rgb[3] = {s->albedo[3 * m], s->albedo[3 * m + 1], s->albedo[3 * m + 2]};
    if(s->albedoTexture && s->albedoTexture[m] >= 0)
    {
        float u = 0.0f, v = 0.0f;
        if(s->uv)
        {
            const uint32_t* ix = s->idx + 3 * (size_t)prim;
            u = s->uv[2 * ix[0]] * a + s->uv[2 * ix[1]] * b + s->uv[2 * ix[2]] * c;
            v = s->uv[2 * ix[0] + 1] * a + s->uv[2 * ix[1] + 1] * b + s->uv[2 * ix[2] + 1] * c;
        }
        orc_texture_sample_grad(s->textures + s->albedoTexture[m], u, v, dpdx, dpdy, s->textureLodMode, rgb);
    }
    if(!s->spectrum) return S(rgb[0], rgb[1], rgb[2], 0.0f);
    s4 o; orc_convert_albedo(s->spectrum, rgb, waves, o.v); return o;
}

static void tri(const pt_scene* s, uint32_t t, v3 p[3])
{
    for(int k = 0; k < 3; k++)
    {
        const float* q = s->pos + 3 * (size_t)s->idx[3 * (size_t)t + k];
        p[k] = V(q[0], q[1], q[2]);
    }
}

typedef int (*orc_leaf_filter)(void* user, uint32_t leaf, const float bary[2]);
void orc_lbvh_trace_filtered(const float* pos, const uint32_t* idx, const uint32_t* nodes, const float* boxes,
                             const float* rays, uint32_t nRays, int mode, int cullFace,
                             uint32_t* outPrim, float* outT, float* outBary, uint8_t* outBack, orc_leaf_filter f, void* user);
void orc_texture_sample(const struct orc_texture* t, float u, float v, float out[3]);
/* IntersectionCheck's stochastic alpha culling (AcceleratorLBVH.hpp:L263-282): alpha = alphaMap(SurfaceParametrization(hit)),
 * the hit is dropped when xi >= alpha; xi comes from the ray's backup generator (here: the path's) */
struct alpha_ctx { const pt_scene* s; pcg* rng; };
static int alpha_filter(void* user, uint32_t leaf, const float bary[2])
{
    struct alpha_ctx* c = (struct alpha_ctx*)user;
    const pt_scene* s = c->s;
    int32_t ti = s->triAlpha[leaf];
    if(ti < 0) return 1;
    const uint32_t* vi = s->idx + 3 * (size_t)leaf;
    float a = bary[0], b = bary[1], cc = 1.0f - a - b;
    float u = s->uv[2 * vi[0]] * a + s->uv[2 * vi[1]] * b + s->uv[2 * vi[2]] * cc;
    float v = s->uv[2 * vi[0] + 1] * a + s->uv[2 * vi[1] + 1] * b + s->uv[2 * vi[2] + 1] * cc;
    float px[3]; orc_texture_sample(&s->textures[ti], u, v, px);
    return pcg_float(c->rng) < px[0];
}
static int trace_rng(const pt_scene* s, pcg* rng, v3 o, v3 d, float tMin, float tMax, int any, uint32_t* prim, float* t, float bary[2])
{
    float ray[8] = {o.x, o.y, o.z, tMin, d.x, d.y, d.z, tMax};
    uint8_t back;
    struct alpha_ctx ctx = {s, rng};
    if(s->triAlpha) orc_lbvh_trace_filtered(s->pos, s->idx, s->nodes, s->boxes, ray, 1, any, 0, prim, t, bary, &back, alpha_filter, &ctx);
    else orc_lbvh_trace(s->pos, s->idx, s->nodes, s->boxes, ray, 1, any, 0, prim, t, bary, &back);
    return *prim != 0xFFFFFFFFu;
}

/* dist_oracle.c */
void orc_dist2d_sample_uv(const float* cdfX, const float* cdfY, uint32_t w, uint32_t h, float xi0, float xi1, float out[3]);
float orc_dist2d_pdf_uv(const float* cdfX, const float* cdfY, uint32_t w, uint32_t h, float u, float v);
void orc_sky_dir_to_uv(int mode, const float d[3], float uv[2]);
void orc_sky_uv_to_dir(int mode, const float uv[2], float d[3]);
float orc_sky_pdf_from_dir(int mode, float pdf, const float d[3]);
float orc_sky_pdf_from_uv(int mode, float pdf, const float uv[2]);

static v3 mul33(const float* m, v3 v) { return V(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z); }
/* LightSkysphere::EmitViaHit / EmitViaSurfacePoint (LightsDefault.hpp:L408-443) */
static s4 sky_emit(const pt_scene* s, v3 wO, const float waves[4])
{
    v3 dir = nrm(mul33(s->boundaryInvM, mul(wO, -1.0f)));
    float rgb[3] = {s->boundaryRadiance[0], s->boundaryRadiance[1], s->boundaryRadiance[2]};
    if(s->boundaryTexture >= 0)
    {
        float dv[3] = {dir.x, dir.y, dir.z}, uv[2];
        orc_sky_dir_to_uv((int)s->boundaryType, dv, uv);
        orc_texture_sample(&s->textures[s->boundaryTexture], uv[0], uv[1], rgb);
    }
    if(!s->spectrum) return S(rgb[0], rgb[1], rgb[2], 0.0f);
    s4 o; orc_convert_radiance(s->spectrum, rgb, waves, o.v); return o;
}
/* LightSkysphere::PdfSolidAngle (L359-370) */
static float sky_pdf(const pt_scene* s, v3 dirWorld)
{
    v3 dY = mul33(s->boundaryInvM, dirWorld), n = nrm(dY);
    float nv[3] = {n.x, n.y, n.z}, dv[3] = {dY.x, dY.y, dY.z}, uv[2];
    orc_sky_dir_to_uv((int)s->boundaryType, nv, uv);
    float pdf = 1.0f;
    if(s->boundaryTexture >= 0)
        pdf = orc_dist2d_pdf_uv(s->boundaryCdfX, s->boundaryCdfY, s->textures[s->boundaryTexture].w, s->textures[s->boundaryTexture].h, uv[0], uv[1]);
    return orc_sky_pdf_from_dir((int)s->boundaryType, pdf, dv);
}
/* LightSkysphere::SampleSolidAngle (L338-357): -> sampled point, solid-angle pdf (not yet divided by the light count) */
static float sky_sample(const pt_scene* s, float x0, float x1, v3 from, v3* lposOut)
{
    float suv[3] = {x0, x1, 1.0f};
    if(s->boundaryTexture >= 0)
        orc_dist2d_sample_uv(s->boundaryCdfX, s->boundaryCdfY, s->textures[s->boundaryTexture].w, s->textures[s->boundaryTexture].h, x0, x1, suv);
    float d[3]; orc_sky_uv_to_dir((int)s->boundaryType, suv, d);
    float pdf = orc_sky_pdf_from_uv((int)s->boundaryType, suv[2], suv);
    v3 worldDir = mul33(s->boundaryM, V(d[0], d[1], d[2]));
    *lposOut = add(from, mul(worldDir, s->sceneDiameter));
    return pdf;
}

static s4 light_emit(const pt_scene* s, uint32_t li, v3 n, v3 wO, const float waves[4])
{
    float NdL = dot(n, wO);
    if(!(s->twoSided && s->twoSided[li]) && NdL <= 0.0f) return S(0, 0, 0, 0);
    if(!s->spectrum) return S(s->radiance[3 * li], s->radiance[3 * li + 1], s->radiance[3 * li + 2], 0.0f);
    s4 o; orc_convert_radiance(s->spectrum, s->radiance + 3 * li, waves, o.v); return o;
}

/* one camera path -> radiance (RGB; spectral paths are converted with their wavelengths at the end, as
 * ConvertSpectrumToRGBIndirect does on dead paths); *filmW receives the filter weight */
/* Distribution::Common::SampleGaussian (DistributionFunctions.h:L686-705): x = sqrt2 sigma erfinv(2 xi - 1), clamped to
 * +-3.5 sigma where erfinv overflows; the inverse error function by Newton iterations on erf in double */
static float gauss_sample(float xi, float sig)
{
    double e = 0;
    double y = 2.0 * xi - 1.0;
    if(y <= -1.0) e = -INFINITY; else if(y >= 1.0) e = INFINITY;
    else { double x = 0; for(int it = 0; it < 60; it++) { double f = erf(x) - y; x -= f / (1.1283791670955126 * exp(-x * x)); } e = x; }
    float x = 1.41421356237f * sig * (float)e;
    if(isinf(e)) { float mm = 3.5f * sig; x = x < -mm ? -mm : (x > mm ? mm : x); }
    return x;
}
/* Math::Gaussian(x, sigma) = PDFGaussian = GaussianFilter::Evaluate per axis (Filters.h:L195-227) */
static float gauss_pdf_mu(float x, float sig, float mu);
static float gauss_pdf(float x, float sig) { return gauss_pdf_mu(x, sig, 0.0f); }
/* GaussianFilter(radius): out = {offset x, offset y, Sample().pdf, Pdf(offset), Evaluate(offset)} */
void orc_pt_filter_sample(float radius, float xi0, float xi1, float out[5])
{
    float sig = radius * 0.285714f;
    out[0] = gauss_sample(xi0, sig); out[1] = gauss_sample(xi1, sig);
    out[2] = gauss_pdf(out[0], sig) * gauss_pdf(out[1], sig);
    out[3] = out[2];
    out[4] = gauss_pdf(out[0], sig) * gauss_pdf(out[1], sig);
}

/* ---- the four film filters of Tracer/Filters.h ---- */
/* Math::Gaussian (Core/Math.h:L1032-1042); InvSqrt2Pi = (1 / Sqrt2<float>) * (1 / SqrtPi<float>) evaluated in float = 0x1.988452p-2 (one ulp under 1 / sqrt(2 pi)) */
static float gauss_pdf_mu(float x, float sig, float mu) { float si = 1.0f / sig, p = (x - mu) * si; return 0x1.988452p-2f * si * expf(-0.5f * p * p); }
static float gauss_sample_mu(float xi, float sig, float mu, float* pdf)
{   /* Common::SampleGaussian(xi, sigma, mu) (DistributionFunctions.h:L686-705) */
    double e, y = 2.0 * xi - 1.0;
    if(y <= -1.0) e = -INFINITY; else if(y >= 1.0) e = INFINITY;
    else { double x = 0; for(int it = 0; it < 60; it++) { double f = erf(x) - y; x -= f / (1.1283791670955126 * exp(-x * x)); } e = x; }
    float x = 1.41421356237f * sig * (float)e + mu;
    if(isinf(e)) { float mm = 3.5f * sig; x = x < -mm ? -mm : (x > mm ? mm : x); }
    *pdf = gauss_pdf_mu(x, sig, mu);
    return x;
}
static float lerp_u(float a, float b, float t) { volatile float x = a * (1.0f - t); volatile float y = b * t; return x + y; }
static const float PREV_ONE = 0.99999994f;
/* Common::SampleTent(xi, -r, r) (DistributionFunctions.h:L785-805): BisectSample2 picks the side, SampleLine(., 1, 0) the
 * distance (L627-643, L745-765) */
static float tent_sample(float xi, float r, float* pdf)
{
    float a = -r, b = r;
    if(b - a < 1.0e-5f) { *pdf = 1.0f / (b - a); return 0.0f; }
    float w = r / (r + r);
    int left = xi < w;
    float lxi = left ? xi / w : (xi - w) / (1.0f - w);
    if(lxi > PREV_ONE) lxi = PREV_ONE;
    if(left) lxi = PREV_ONE - lxi;
    float den = lerp_u(1.0f, 0.0f, lxi); if(den < 0) den = 0;
    float x = lxi / (1.0f + sqrtf(den)); if(x > PREV_ONE) x = PREV_ONE;
    *pdf = 2.0f * lerp_u(1.0f, 0.0f, x) * (1.0f / (b - a));
    return left ? x * a : x * b;
}
/* Common::PDFTent (L807-821) */
static float tent_pdf(float x, float r) { float x01 = (x < 0) ? x / -r : x / r; return (1.0f / (r + r)) * 2.0f * lerp_u(1.0f, 0.0f, x01); }
/* MitchellNetravaliFilter (Filters.h:L234-380), b = c = 0.33333 */
static float mitchell_1d(float x, float rr)
{
    const float B = 0.33333f, Cc = 0.33333f, F = 1.0f / 6.0f;
    x = fabsf(2.0f * x * rr);
    float x2 = x * x, x3 = x2 * x, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if(x < 1.0f) { c0 = F * (12.0f - 9.0f * B - 6.0f * Cc); c1 = F * (-18.0f + 12.0f * B + 6.0f * Cc); c3 = F * (6.0f - 2.0f * B); }
    else if(x < 2.0f) { c0 = F * (-B - 6.0f * Cc); c1 = F * (6.0f * B + 30.0f * Cc); c2 = F * (-12.0f * B - 48.0f * Cc); c3 = F * (8.0f * B + 24.0f * Cc); }
    return (c0 * x3 + c1 * x2 + c2 * x + c3) * 2.0f * rr;
}
static const float MN_MID = 0.960566188838f, MN_SIDES = 0.0197169055809f;
static float mitchell_pdf_1d(float x, float r)
{
    float midS = 0.528f * r * 0.5f, sideS = 0.2f * r * 0.5f, sideM = 1.3f * r * 0.5f;
    return gauss_pdf_mu(x, sideS, -sideM) * MN_SIDES + gauss_pdf_mu(x, midS, 0) * MN_MID + gauss_pdf_mu(x, sideS, sideM) * MN_SIDES;
}
static float mitchell_sample_1d(float xi, float r, float* pdf)
{
    float midS = 0.528f * r * 0.5f, sideS = 0.2f * r * 0.5f, sideM = 1.3f * r * 0.5f;
    int index; float lo, wsel;
    if(xi < MN_SIDES) { index = 0; lo = 0; wsel = MN_SIDES; }
    else if(xi < MN_SIDES + MN_MID) { index = 1; lo = MN_SIDES; wsel = MN_MID; }
    else { index = 2; lo = MN_SIDES + MN_MID; wsel = MN_SIDES; }
    float lxi = (xi - lo) / wsel; if(lxi > PREV_ONE) lxi = PREV_ONE;
    float own, x;
    if(index == 0) x = gauss_sample_mu(lxi, sideS, -sideM, &own);
    else if(index == 1) x = gauss_sample_mu(lxi, midS, 0, &own);
    else x = gauss_sample_mu(lxi, sideS, sideM, &own);
    float p0 = index == 0 ? own : gauss_pdf_mu(x, sideS, -sideM);
    float p1 = index == 1 ? own : gauss_pdf_mu(x, midS, 0);
    float p2 = index == 2 ? own : gauss_pdf_mu(x, sideS, sideM);
    *pdf = p0 * MN_SIDES + p1 * MN_MID + p2 * MN_SIDES;
    return x;
}
/* type = FilterType::E (0 Box, 1 Tent, 2 Gaussian, 3 Mitchell-Netravali):
 * out = {offset x, offset y, Sample().pdf, Pdf(offset), Evaluate(offset)} */
void orc_pt_filter_sample_typed(uint32_t type, float radius, float xi0, float xi1, float out[5])
{
    float px, py;
    if(type == 0u)
    {   /* BoxFilter (Filters.h:L106-146) */
        float range = radius - (-radius), rr = 1.0f / radius;
        out[0] = xi0 * range + (-radius); out[1] = xi1 * range + (-radius);
        px = py = 1.0f / range;
        out[3] = (1.0f / range) * (1.0f / range);
        out[4] = (fabsf(out[0]) <= radius && fabsf(out[1]) <= radius) ? 0.25f * rr * rr : 0.0f;
    }
    else if(type == 1u)
    {   /* TentFilter (Filters.h:L148-193) */
        out[0] = tent_sample(xi0, radius, &px); out[1] = tent_sample(xi1, radius, &py);
        out[3] = tent_pdf(out[0], radius) * tent_pdf(out[1], radius);
        float rcp = 1.0f / radius, cap = 1.0f / radius, tx = fabsf(out[0] * rcp), ty = fabsf(out[1] * rcp);
        out[4] = lerp_u(cap, 0.0f, tx > 1 ? 1 : tx) * lerp_u(cap, 0.0f, ty > 1 ? 1 : ty);
    }
    else if(type == 3u)
    {
        out[0] = mitchell_sample_1d(xi0, radius, &px); out[1] = mitchell_sample_1d(xi1, radius, &py);
        out[3] = mitchell_pdf_1d(out[0], radius) * mitchell_pdf_1d(out[1], radius);
        out[4] = mitchell_1d(out[0], 1.0f / radius) * mitchell_1d(out[1], 1.0f / radius);
    }
    else { orc_pt_filter_sample(radius, xi0, xi1, out); return; }
    out[2] = px * py;
}

/* Distribution::Common::SampleCosDirection (DistributionFunctions.h:L847-871), +Z hemisphere: out = {x, y, z, pdf} */
static void cos_direction(float u0, float u1, float out[4])
{
    float phi = 6.28318530718f * u1, su = sqrtf(u0);
    float lx = su * cosf(phi), ly = su * sinf(phi);
    float lz2 = 1.0f - (lx * lx + ly * ly); float lz = lz2 > 0 ? sqrtf(lz2) : 0;
    out[0] = lx; out[1] = ly; out[2] = lz; out[3] = lz * 0.31830988618f;
}
void orc_pt_sample_cos_direction(float u0, float u1, float out[4]) { cos_direction(u0, u1, out); }

/* LightPrim<Triangle>::SampleSolidAngle (LightsDefault.hpp:L22-46) over Triangle::SampleSurface (Osada,
 * PrimitiveDefaultTriangle.hpp:L48-77): q = the triangle, from = the shaded point; outputs the sampled position, the light's
 * geometric normal, the area and the solid-angle pdf (NOT yet divided by the light count). */
static float light_sample(const v3 q[3], int twoSided, float x0, float x1, v3 from, v3* lposOut, v3* lNOut, v3* sdOut)
{
    float r1 = sqrtf(x0), r2 = x1;
    float la = 1 - r1, lb = (1 - r2) * r1, lc = r1 * r2;
    v3 lpos = add(add(mul(q[0], la), mul(q[1], lb)), mul(q[2], lc));
    v3 le0 = sub(q[1], q[0]), le1 = sub(q[2], q[0]);
    v3 lN = nrm(cross(le0, le1));
    float area = 0.5f * len(cross(le0, le1));
    v3 sd = sub(from, lpos); float distSqr = dot(sd, sd); sd = nrm(sd);
    float NdL = dot(lN, sd);
    NdL = twoSided ? fabsf(NdL) : (NdL > 0 ? NdL : 0);
    float pdfL = (NdL == 0) ? 0 : (1.0f / area) / NdL;
    pdfL *= distSqr;
    *lposOut = lpos; *lNOut = lN; *sdOut = sd;
    return pdfL;
}
/* LightPrim::PdfSolidAngle (LightsDefault.hpp:L48-68) for a hit at `hitPos` on the triangle seen from `from` along dir */
static float light_pdf(const v3 q[3], int twoSided, v3 hitPos, v3 from, v3 dir)
{
    v3 le0 = sub(q[1], q[0]), le1 = sub(q[2], q[0]);
    v3 lN = nrm(cross(le0, le1));
    float area = 0.5f * len(cross(le0, le1));
    float NdL = dot(lN, mul(dir, -1.0f));
    NdL = twoSided ? fabsf(NdL) : (NdL > 0 ? NdL : 0);
    float pdfL = (NdL == 0) ? 0 : (1.0f / area) / NdL;
    v3 dv = sub(from, hitPos);
    return pdfL * dot(dv, dv);
}
/* test entry: tri[9], xi[2], from[3] -> out = {sampled pos xyz, pdf of SampleSolidAngle, PdfSolidAngle of the ray from
 * `from` towards the sample, evaluated at its own intersection with the triangle's plane} */
void orc_pt_light_sample(const float* triv, int twoSided, float x0, float x1, const float* fromv, float out[5])
{
    v3 q[3] = {V(triv[0], triv[1], triv[2]), V(triv[3], triv[4], triv[5]), V(triv[6], triv[7], triv[8])};
    v3 from = V(fromv[0], fromv[1], fromv[2]), lpos, lN, sd;
    out[3] = light_sample(q, twoSided, x0, x1, from, &lpos, &lN, &sd);
    out[0] = lpos.x; out[1] = lpos.y; out[2] = lpos.z;
    v3 dir = nrm(sub(lpos, from));
    /* intersect the plane of the triangle along dir (Triangle::Intersects gives the same point up to rounding) */
    float tt = dot(sub(q[0], from), lN) / dot(dir, lN);
    v3 hit = add(from, mul(dir, tt));
    out[4] = light_pdf(q, twoSided, hit, from, dir);
}

/* ---- (Mt)Refract / (Mt)Unreal (Tracer/MaterialsDefault.hpp:L232-760, Tracer/DistributionFunctions.h:L382-590) ---- */
static float sqrt_max(float x) { return x > 0 ? sqrtf(x) : 0.0f; }
static float pdf_cos_direction(float cosT) { float p = cosT * 0.31830988618f; return p <= 1.0e-5f ? 0.0f : p; }   /* Common::PDFCosDirection */
static float fresnel_dielectric(float cosFront, float etaFront, float etaBack)
{
    float sinFront = sqrt_max(1.0f - cosFront * cosFront), sinBack = etaFront / etaBack * sinFront;
    if(sinFront >= 1.0f) return 1.0f;
    float cosBack = sqrt_max(1.0f - sinBack * sinBack);
    float par = (etaBack * cosFront - etaFront * cosBack) / (etaBack * cosFront + etaFront * cosBack);
    float per = (etaFront * cosFront - etaBack * cosBack) / (etaFront * cosFront + etaBack * cosBack);
    return (par * par + per * per) * 0.5f;
}
static float cauchy_ior(float wavelengthNm, const float* c) { float w = wavelengthNm * 1.0e-3f, w2 = w * w; return c[0] + c[1] / w2 + c[2] / (w2 * w2); }
static float d_ggx(float NdH, float alpha) { float a2 = alpha * alpha, den = NdH * NdH * (a2 - 1.0f) + 1.0f; return a2 / (den * den * 3.14159265358979f); }
static float lambda_smith(v3 v, float alpha) { return (sqrtf(1.0f + alpha * alpha * (v.x * v.x + v.y * v.y) / (v.z * v.z)) - 1.0f) * 0.5f; }
static float g_smith_single(v3 v, float alpha) { return 1.0f / (1.0f + lambda_smith(v, alpha)); }
static float g_smith_correlated(v3 a, v3 b, float alpha) { return 1.0f / (lambda_smith(a, alpha) + lambda_smith(b, alpha) + 1.0f); }
static float vndf_pdf(v3 Vv, v3 H, float alpha)
{
    float VdH = dot(H, Vv); if(VdH < 0) VdH = 0;
    float NdH = H.z > 0 ? H.z : 0, NdV = Vv.z > 0 ? Vv.z : 0;
    if(NdV == 0.0f) return 0.0f;
    return VdH * d_ggx(NdH, alpha) * g_smith_single(Vv, alpha) / NdV;
}
static v3 vndf_sample(v3 Vv, float alpha, float xi0, float xi1, float* pdf)
{
    v3 VH = nrm(V(alpha * Vv.x, alpha * Vv.y, Vv.z));
    float len2 = VH.x * VH.x + VH.y * VH.y;
    v3 T1 = len2 > 0 ? mul(V(-VH.y, VH.x, 0), 1.0f / sqrtf(len2)) : V(1, 0, 0);
    v3 T2 = cross(VH, T1);
    float r = sqrtf(xi0), phi = 6.28318530718f * xi1;
    float t1 = r * cosf(phi), t2 = r * sinf(phi), sh = 0.5f * (1.0f + VH.z);
    t2 = (1.0f - sh) * sqrtf(1.0f - t1 * t1) + sh * t2;
    v3 NH = add(add(mul(T1, t1), mul(T2, t2)), mul(VH, sqrt_max(1.0f - t1 * t1 - t2 * t2)));
    v3 N = V(alpha * NH.x, alpha * NH.y, sqrt_max(NH.z));
    float nl2 = dot(N, N);
    N = nl2 < 1.0e-5f ? V(0, 0, 1) : mul(N, 1.0f / sqrtf(nl2));
    *pdf = vndf_pdf(Vv, N, alpha);
    return N;
}
typedef struct { s4 albedo; float roughness, specular, metallic; } unreal_mat;
static float unreal_mis_ratio(const unreal_mat* u)
{
    float avg = (u->albedo.v[0] + u->albedo.v[1] + u->albedo.v[2] + u->albedo.v[3]) * 0.3333f;
    float integralDiffuse = 2.0f * 3.14159265358979f * avg * (1.0f - u->metallic);
    float specularRatio = u->specular * (1.0f - u->metallic) + avg * u->metallic;
    float total = specularRatio + integralDiffuse;
    return total == 0.0f ? 0.0f : integralDiffuse / total;
}
static s4 unreal_fschlick(const unreal_mat* u, float VdH)
{
    float specOut = u->specular * 0.08f, pw = 1.0f - VdH, pw5 = pw * pw * pw * pw * pw; s4 o;
    for(int k = 0; k < 4; k++) { float f0 = specOut * (1.0f - u->metallic) + u->albedo.v[k] * u->metallic; o.v[k] = (1.0f - f0) * pw5 + f0; }
    return o;
}
static float burley(float NdL, float NdV, float LdH, float roughness)
{
    float fd90 = 0.5f + 2.0f * roughness * LdH * LdH;
    float a = 1.0f - NdL, b = 1.0f - NdV;
    return (1.0f + (fd90 - 1.0f) * a * a * a * a * a) * (1.0f + (fd90 - 1.0f) * b * b * b * b * b);
}
static s4 unreal_terms(const unreal_mat* u, v3 Vt, v3 L, v3 H, int cancelOnBadD, float* pdfSpecOverride)
{
    float alpha = u->roughness * u->roughness;
    float LdH = dot(L, H), VdH = dot(Vt, H); if(LdH < 0) LdH = 0; if(VdH < 0) VdH = 0;
    float NdH = H.z > 0 ? H.z : 0, NdV = Vt.z > 0 ? Vt.z : 0, NdL = L.z > 0 ? L.z : 0;
    float D = d_ggx(NdH, alpha);
    int badD = isnan(D) || isinf(D);
    float G = g_smith_correlated(Vt, L, alpha);
    if(LdH == 0.0f || VdH == 0.0f) G = 0.0f;
    s4 F = unreal_fschlick(u, VdH), spec;
    if(badD && cancelOnBadD) { spec = s_mul(F, G / g_smith_single(Vt, alpha)); if(pdfSpecOverride) *pdfSpecOverride = 1.0f; }
    else
    {
        if(badD) D = 0.0f;
        spec = (NdV == 0.0f) ? S(0, 0, 0, 0) : s_mul(F, D * G * 0.25f / NdV);
    }
    s4 diff = s_mul(u->albedo, NdL * 0.31830988618f * (1.0f - u->metallic) * burley(NdL, NdV, LdH, u->roughness));
    return s_add(diff, spec);
}
static s4 unreal_evaluate(const unreal_mat* u, v3 Vt, v3 L) { return unreal_terms(u, Vt, L, nrm(add(L, Vt)), 0, NULL); }
static float unreal_pdf(const unreal_mat* u, v3 Vt, v3 L)
{
    float alpha = u->roughness * u->roughness, mis = unreal_mis_ratio(u);
    v3 H = nrm(add(L, Vt));
    float NdH = H.z > 0 ? H.z : 0, VdH = dot(Vt, H); if(VdH < 0) VdH = 0;
    float D = d_ggx(NdH, alpha), ps = vndf_pdf(Vt, H, alpha);
    if(isnan(D) || isinf(D)) ps = 0.0f;
    ps = VdH == 0.0f ? 0.0f : ps / (4.0f * VdH);
    return pdf_cos_direction(L.z) * mis + ps * (1.0f - mis);
}
static v3 unreal_sample(const unreal_mat* u, v3 Vt, float sXi, float xi0, float xi1, s4* refl, float* pdfOut)
{
    float alpha = u->roughness * u->roughness, mis = unreal_mis_ratio(u), pd, ps;
    v3 L, H;
    if(sXi < mis)
    {
        float cd[4]; cos_direction(xi0, xi1, cd);
        L = V(cd[0], cd[1], cd[2]); pd = cd[3];
        H = nrm(add(L, Vt));
        float VdH = dot(Vt, H); if(VdH < 0) VdH = 0;
        ps = vndf_pdf(Vt, H, alpha); ps = VdH == 0.0f ? 0.0f : ps / (4.0f * VdH);
    }
    else
    {
        float ph; H = vndf_sample(Vt, alpha, xi0, xi1, &ph);
        L = sub(mul(H, 2.0f * dot(Vt, H)), Vt);
        float VdH = dot(Vt, H); if(VdH < 0) VdH = 0;
        ps = VdH == 0.0f ? 0.0f : ph / (4.0f * VdH);
        pd = pdf_cos_direction(L.z);
    }
    *refl = unreal_terms(u, Vt, L, H, 1, &ps);
    *pdfOut = pd * mis + ps * (1.0f - mis);
    return L;
}

/* Quaternion::SLerp / BarySLerp / OrthoBasisZ (Core/Quaternion.hpp:L256-353): the shading normal of a hit is the Z axis of
 * the barycentric blend of the three vertex frames (Triangle::GenerateSurface, PrimitiveDefaultTriangle.hpp:L463-470) */
static void quat_slerp(const float* a, const float* b, float t, float* o)
{
    float cosT = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3], cf = cosT >= 0 ? cosT : -cosT, s0, s1;
    if(cf < 1.0f - 1.0e-5f) { float ang = acosf(cf), sr = 1.0f / sinf(ang); s0 = sinf(ang * (1.0f - t)) * sr; s1 = sinf(ang * t) * sr; }
    else { s0 = 1.0f - t; s1 = t; }
    if(cosT < 0) s1 = -s1;
    for(int k = 0; k < 4; k++) o[k] = a[k] * s0 + b[k] * s1;
}
static v3 tbn_normal(const float* q0, const float* q1, const float* q2, float a, float b)
{
    float q[4], qab[4];
    if(fabsf(a + b) < 1.0e-5f) memcpy(q, q2, sizeof(q));
    else { quat_slerp(q1, q0, a / (a + b), qab); quat_slerp(qab, q2, 1.0f - a - b, q); }
    float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    return V(2.0f * (x * z - w * y), 2.0f * (y * z + w * x), w * w - x * x - y * y + z * z);
}

/* Shading normal under a normal map (Triangle::GenerateSurface, PrimitiveDefaultTriangle.hpp:L478-491,L571-575): the frame
 * (turned 180 degrees about its tangent on a back-side hit: TANGENT_ROT * tbn) is re-aimed with
 * RotationBetweenZAxis(n).Conjugate() * tbn, so its Z axis in the world is tbn^-1 (n) */
static v3 tbn_normal_mapped(const float* q0, const float* q1, const float* q2, float a, float b, v3 nTS, int backSide)
{
    float q[4], qab[4];
    if(fabsf(a + b) < 1.0e-5f) memcpy(q, q2, sizeof(q));
    else { quat_slerp(q1, q0, a / (a + b), qab); quat_slerp(qab, q2, 1.0f - a - b, q); }
    float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    v3 r0 = V(w * w + x * x - y * y - z * z, 2.0f * (x * y - w * z), 2.0f * (x * z + w * y));
    v3 r1 = V(2.0f * (x * y + w * z), w * w - x * x + y * y - z * z, 2.0f * (y * z - w * x));
    v3 r2 = V(2.0f * (x * z - w * y), 2.0f * (y * z + w * x), w * w - x * x - y * y + z * z);
    float sg = backSide ? -1.0f : 1.0f;
    return add(mul(r0, nTS.x), mul(add(mul(r1, nTS.y), mul(r2, nTS.z)), sg));
}

/* ---- ray cones (Tracer/TracerTypes.h:L48-69,L323-383; RT Gems I ch. 20, Akenine-Moller et al. JCGT 10(1)) ---- */
typedef struct { float aperture, width; } cone_t;
typedef struct { cone_t front, back; float betaN; } cone_surf;
static cone_t cone_advance(cone_t c, float t)
{   /* RayCone::Advance: width + aperture * t, clamped to [Epsilon, 1e6] */
    float w = c.width + c.aperture * t;
    w = w < 1.0e-5f ? 1.0e-5f : (w > 1.0e6f ? 1.0e6f : w);
    cone_t r = {c.aperture, w}; return r;
}
static void cone_project(cone_t c, v3 f, v3 d, v3* a1, v3* a2)
{   /* RayCone::Project (JCGT eq. 8, 9): the two axes of the cone's footprint ellipse on the plane with normal f */
    const float EPS = 1.0e-5f;
    const float fd = dot(f, d);
    if(fabsf(fd + 1.0f) < EPS) d = add(d, V(EPS, EPS, EPS));
    v3 h1 = sub(d, mul(f, fd)), h2 = cross(f, h1);
    const float r = c.width * 0.5f;
    float den1 = len(sub(h1, mul(d, dot(d, h1)))); if(den1 < EPS) den1 = EPS;
    float den2 = len(sub(h2, mul(d, dot(d, h2)))); if(den2 < EPS) den2 = EPS;
    *a1 = mul(h1, r / den1); *a2 = mul(h2, r / den2);
}
/* test tap of cone_project (the inputs of the reference's Tests/Tracer/T_RayCone.cu) */
void orc_ray_cone_project(float aperture, float width, const float f[3], const float d[3], float out[6])
{
    cone_t c = {aperture, width}; v3 a1, a2;
    cone_project(c, V(f[0], f[1], f[2]), V(d[0], d[1], d[2]), &a1, &a2);
    out[0] = a1.x; out[1] = a1.y; out[2] = a1.z; out[3] = a2.x; out[4] = a2.y; out[5] = a2.z;
}
static v3 quat_z(const float* q)
{   /* Quaternion::OrthoBasisZ (Core/Quaternion.hpp:L256-270), q = (w, x, y, z) */
    const float v00 = q[0] * q[0], v01 = q[0] * q[1], v02 = q[0] * q[2], v11 = q[1] * q[1], v13 = q[1] * q[3], v22 = q[2] * q[2], v23 = q[2] * q[3], v33 = q[3] * q[3];
    return V(v13 - v02 + v13 - v02, v23 + v01 + v23 + v01, v00 - v11 - v22 + v33);
}
/* The ray-cone half of Triangle::GenerateSurface (PrimitiveDefaultTriangle.hpp:L495-569): curvature estimate betaN from the
 * vertex normals along the triangle's edges (JCGT eq. 6), and the UV differences dpdx / dpdy over the footprint's axes
 * (listing 1). gN = the geometric normal already flipped towards the ray, dDotN = dot(unflipped normal, dir). */
static void cone_surface(const v3 p[3], const float* q0, const float* q1, const float* q2, const float* uv0, const float* uv1, const float* uv2,
                         float a, float b, v3 pos, v3 gN, float dDotN, v3 dirN, cone_t cone, cone_surf* out, float dpdx[2], float dpdy[2])
{
    v3 a1, a2; cone_project(cone, gN, dirN, &a1, &a2);
    const v3 r0 = nrm(a1), r1 = nrm(a2);
    const v3 e[3] = {sub(p[1], p[0]), sub(p[2], p[0]), sub(p[2], p[1])};
    float k[3] = {0, 0, 0};
    if(q0)
    {
        const v3 n0 = quat_z(q0), n1 = quat_z(q1), n2 = quat_z(q2);
        k[0] = dot(sub(n1, n0), e[0]) / dot(e[0], e[0]);
        k[1] = dot(sub(n2, n0), e[1]) / dot(e[1], e[1]);
        k[2] = dot(sub(n2, n1), e[2]) / dot(e[2], e[2]);
    }
    uint32_t mn = 0, mx = 0;
    for(uint32_t i = 1; i < 3; i++) { if(k[i] < k[mn]) mn = i; if(k[i] > k[mx]) mx = i; }
    const float eMin[2] = {dot(r0, e[mn]), dot(r1, e[mn])}, eMax[2] = {dot(r0, e[mx]), dot(r1, e[mx])};
    const float a1L = len(a1), a2L = len(a2), a1S = a1L * a1L, a2S = a2L * a2L, a12 = a1L * a2L;
    const float l0 = a12 * (1.0f / sqrtf(a1S * eMin[0] * eMin[0] + a2S * eMin[1] * eMin[1]));
    const float l1 = a12 * (1.0f / sqrtf(a1S * eMax[0] * eMax[0] + a2S * eMax[1] * eMax[1]));
    const float lMaxRecip = 1.0f / (l0 > l1 ? l0 : l1);
    const float k0 = k[mn] * l0 * lMaxRecip, k1 = k[mx] * l1 * lMaxRecip;
    const float beta0 = -1.0f * k0 * fabsf(cone.width) / dDotN, beta1 = -1.0f * k1 * fabsf(cone.width) / dDotN;
    out->front = cone; out->back = cone;
    out->betaN = (fabsf(cone.aperture + beta0) >= fabsf(cone.aperture + beta1)) ? beta0 : beta1;
    /* texture gradients: barycentrics of pos + axis, UV there minus UV here */
    const float areaRecip = 1.0f / dot(gN, cross(e[0], e[1]));
    const float c = 1.0f - a - b;
    const float u = uv0[0] * a + uv1[0] * b + uv2[0] * c, v = uv0[1] * a + uv1[1] * b + uv2[1] * c;
    const v3 axes[2] = {a1, a2};
    float* outs[2] = {dpdx, dpdy};
    for(int i = 0; i < 2; i++)
    {
        const v3 eP = add(sub(pos, p[0]), axes[i]);
        const float ba = dot(gN, mul(cross(eP, e[1]), areaRecip)), bb = dot(gN, mul(cross(e[0], eP), areaRecip)), bc = 1.0f - ba - bb;
        outs[i][0] = (uv0[0] * bc + uv1[0] * ba + uv2[0] * bb) - u;
        outs[i][1] = (uv0[1] * bc + uv1[1] * ba + uv2[1] * bb) - v;
    }
}
static cone_t cone_after_scatter(const cone_surf* cs, v3 wI, v3 n)
{   /* RayConeSurface::ConeAfterScatter: reflected cones widen by twice the curvature term, refracted ones take the back cone */
    cone_t f = {cs->front.aperture + 2.0f * cs->betaN, cs->front.width}, b = {cs->back.aperture - cs->betaN, cs->back.width};
    return dot(wI, n) > 0.0f ? f : b;
}
static void rot2(float vx, float vy, float alpha, float u[2], float l[2])
{   /* Rotate2D_UL: v turned by +alpha and by -alpha */
    const float sn = sinf(alpha), cs = cosf(alpha);
    u[0] = vx * cs - vy * sn; u[1] = vx * sn + vy * cs;
    l[0] = vx * cs + vy * sn; l[1] = vx * -sn + vy * cs;
}
static void refract2(const float v[2], const float n[2], float fromEta, float toEta, float out[2])
{   /* Refract2D: Graphics::Refract(n, -v) in the plane; under total internal reflection the tangential direction */
    const float er = fromEta / toEta, cosIn = fmaf(n[1], -v[1], n[0] * -v[0]);   /* Math::Dot(normal, -v): an FMA chain */
    float sinIn2 = 1.0f - cosIn * cosIn; if(sinIn2 < 0) sinIn2 = 0;
    const float sinOut2 = er * er * sinIn2;
    if(sinOut2 >= 1.0f)
    {
        const float nd = n[0] * v[0] + n[1] * v[1];
        float tx = v[0] - n[0] * nd, ty = v[1] - n[1] * nd; const float l = sqrtf(tx * tx + ty * ty);
        out[0] = tx / l; out[1] = ty / l; return;
    }
    const float cosOut = sqrt_max(1.0f - sinOut2);
    out[0] = er * v[0] + (er * cosIn - cosOut) * n[0]; out[1] = er * v[1] + (er * cosIn - cosOut) * n[1];
}
/* RefractMaterial::RefractRayCone (MaterialsDefault.hpp:L355-462, after RT Gems II ch. 10 / Falcor): the back cone of a surface
 * whose refraction exists — the upper and lower edge rays of the cone refracted in the plane of incidence through normals
 * tilted by the curvature term. fromEta / toEta already swapped for a back-side hit; gN flipped towards wO. */
static cone_surf refract_ray_cone(cone_surf in, v3 wO, v3 gN, float fromEta, float toEta)
{
    const float er = fromEta / toEta, cosIn3 = dot(gN, wO);
    float sinIn2 = 1.0f - cosIn3 * cosIn3; if(sinIn2 < 0) sinIn2 = 0;
    if(er * er * sinIn2 >= 1.0f) return in;
    const float cosOut3 = sqrt_max(1.0f - er * er * sinIn2);
    const v3 t3 = add(mul(wO, -er), mul(gN, er * cosIn3 - cosOut3)), d3 = mul(wO, -1.0f);
    const v3 x = nrm(sub(d3, mul(gN, dot(d3, gN)))), y = gN;
    const float d[2] = {dot(x, d3), dot(y, d3)};
    (void)t3;
    const float aperture = in.front.aperture, width = in.front.width;
    const float wSign = width > 0.0f ? 1.0f : 0.0f;
    float du[2], dl[2]; rot2(d[0], d[1], wSign * aperture * 0.5f, du, dl);
    float od[2] = {-d[1] * width * 0.5f, d[0] * width * 0.5f};
    const float uHitX = +od[0] + du[0] * (-od[1] / du[1]), lHitX = -od[0] + dl[0] * (+od[1] / dl[1]);
    const float nSign = uHitX > lHitX ? 1.0f : -1.0f;
    const float dN = -in.betaN * nSign * 0.5f;
    float nu[2], nl[2]; rot2(0.0f, 1.0f, dN, nu, nl);
    float tu[2], tl[2]; refract2(du, nu, fromEta, toEta, tu); refract2(dl, nl, fromEta, toEta, tl);
    od[0] = -d[1]; od[1] = d[0];
    float wl = -uHitX * tu[1]; wl /= od[0] * -tu[1] + od[1] * tu[0];
    float wu = +lHitX * tl[1]; wu /= od[0] * -tl[1] + od[1] * tl[0];
    const float sign = copysignf(1.0f, tu[0] * tl[1] - tu[1] * tl[0]);
    float ct = fmaf(tu[1], tl[1], tu[0] * tl[0]); ct = ct < -1.0f ? -1.0f : (ct > 1.0f ? 1.0f : ct);   /* Math::Dot: FMA chain; acos near 1 is ill-conditioned */
    float ap = acosf(ct) * sign; if(ap < 1.0e-5f) ap = 1.0e-5f;
    cone_surf r = in;
    r.back.aperture = ap + in.betaN; r.back.width = wu + wl;
    return r;
}
/* test tap of refract_ray_cone + cone_after_scatter for a transmitted ray: out = {aperture, width} of the cone that continues */
void orc_refract_ray_cone(float aperture, float width, float betaN, const float wO[3], const float n[3], float fromEta, float toEta, float out[2])
{
    cone_surf cs; cs.front.aperture = aperture; cs.front.width = width; cs.back = cs.front; cs.betaN = betaN;
    const v3 o = V(wO[0], wO[1], wO[2]), nn = V(n[0], n[1], n[2]);
    const cone_surf r = refract_ray_cone(cs, o, nn, fromEta, toEta);
    const cone_t c = cone_after_scatter(&r, mul(nn, -1.0f), nn);   /* any direction below the surface selects the back cone */
    out[0] = c.aperture; out[1] = c.width;
}
static int scene_has_mips(const pt_scene* s)
{ for(uint32_t t = 0; t < s->nTextures; t++) if(s->textures[t].mipCount > 1u) return 1; return 0; }

static s4 path_spectrum(const pt_scene* s, pcg* rng, uint32_t px, uint32_t py, float* filmW, float waves[4], float wavePdf[4]);
static v3 path(const pt_scene* s, pcg* rng, uint32_t px, uint32_t py, float* filmW)
{
    float waves[4] = {0, 0, 0, 0}, wavePdf[4] = {1, 1, 1, 1};
    s4 r = path_spectrum(s, rng, px, py, filmW, waves, wavePdf);
    if(!s->spectrum) return V(r.v[0], r.v[1], r.v[2]);
    float rgb[4];
    orc_spectra_to_rgb(s->spectrum, r.v, waves, wavePdf, rgb);
    return V(rgb[0], rgb[1], rgb[2]);
}

static s4 path_spectrum(const pt_scene* s, pcg* rng, uint32_t px, uint32_t py, float* filmW, float waves[4], float wavePdf[4])
{
    /* camera (CameraPinhole ctor + EvaluateRay with a filter-sampled offset) */
    v3 pos = V(s->camPos[0], s->camPos[1], s->camPos[2]);
    v3 gz = sub(V(s->camGaze[0], s->camGaze[1], s->camGaze[2]), pos), up = V(s->camUp[0], s->camUp[1], s->camUp[2]);
    v3 right = nrm(cross(gz, up)); up = nrm(cross(right, gz)); gz = nrm(cross(up, right));
    float wh = tanf(s->fovXY[0] * 0.5f) * s->nearFar[0], hh = tanf(s->fovXY[1] * 0.5f) * s->nearFar[0];
    v3 bl = add(sub(sub(pos, mul(right, wh)), mul(up, hh)), mul(gz, s->nearFar[0]));
    float off[2];
    if(s->filmFilter == 0u || s->filmFilter == 3u)
    {
        float sig = s->filterRadius * 0.285714f;
        for(int k = 0; k < 2; k++) off[k] = gauss_sample(pcg_float(rng), sig);
        *filmW = 1.0f; /* Evaluate(offset) / pdf(offset): same Gaussian */
    }
    else
    {   /* KCGenerateCamRaysStochastic (RayGenKernels.kt.h:L195-225): weight = Evaluate(offset) / Sample().pdf */
        float fo[5], x0 = pcg_float(rng), x1 = pcg_float(rng);
        orc_pt_filter_sample_typed(s->filmFilter - 1u, s->filterRadius, x0, x1, fo);
        off[0] = fo[0]; off[1] = fo[1];
        *filmW = fo[4] / fo[2];
    }
    float sx = ((float)px + off[0] + 0.5f) * (2.0f * wh / (float)s->width);
    float sy = ((float)py + off[1] + 0.5f) * (2.0f * hh / (float)s->height);
    v3 o = pos, d = nrm(sub(add(add(bl, mul(right, sx)), mul(up, sy)), pos));
    float tMin = s->nearFar[0], tMax = s->nearFar[1];

    if(s->spectrum)
    {   /* one more dimension after the camera sample (PathTracerRendererBase.cu:L139-168) */
        uint32_t rn = pcg_next(rng);
        orc_sample_wavelengths((int)s->wavelengthMode, &rn, 1, waves, wavePdf);
    }
    s4 throughput = S(1, 1, 1, 1), radiance = S(0, 0, 0, 0);
    /* CameraPinhole::EvaluateRay (CamerasDefault.hpp:L134-139): aperture = 2 tan(fovY / 2) / resolution.y, width 0. The cone only
     * matters to textures with more than one mip level, so it is tracked only then. */
    const int useCones = scene_has_mips(s);
    cone_t cone = {2.0f * tanf(s->fovXY[1] * 0.5f) / (float)s->height, 0.0f};
    uint32_t depth = 0; int type = 3; /* CAMERA_RAY */
    float prevPdf = 0;
    const uint32_t nLights = s->nLightTris + 1u; /* + boundary (Null) light, MetaLight.hpp:L514-516 */
    for(;;)
    {
        uint32_t prim; float t, bary[2];
        if(!trace_rng(s, rng, o, d, tMin, tMax, 0, &prim, &t, bary))
        {   /* boundary light (LightWorkFunction[WithNEE] for a light without primitives): (L)Null adds nothing */
            if(s->boundaryType != 0u && !(s->sampleMode == 1u && type != 3 && type != 1))
            {
                s4 thr = throughput;
                if(s->sampleMode == 2u && type == 2)
                {
                    float pdfL = sky_pdf(s, d) * (1.0f / (float)nLights);
                    float mis = prevPdf + pdfL;
                    thr = s_mul(thr, prevPdf);
                    thr = (mis == 0) ? S(0, 0, 0, 0) : s_mul(thr, 1.0f / mis);
                }
                if(depth + 1u <= s->rrHi) radiance = s_add(radiance, s_mulv(sky_emit(s, mul(d, -1.0f), waves), thr));
            }
            break;
        }
        v3 p[3]; tri(s, prim, p);
        float a = bary[0], b = bary[1], c = 1.0f - a - b;
        v3 hitPos = add(add(mul(p[0], a), mul(p[1], b)), mul(p[2], c));
        v3 e0 = sub(p[1], p[0]), e1 = sub(p[2], p[0]);
        v3 gN = nrm(cross(e0, e1));
        int32_t m = s->triMaterial[prim];
        if(m < 0)
        {   /* light hit */
            uint32_t li = (uint32_t)(-1 - m);
            int count = !(s->sampleMode == 1u && type != 3 && type != 1);
            if(count)
            {
                s4 thr = throughput;
                if(s->sampleMode == 2u && type == 2)
                {
                    float pdfL = light_pdf(p, s->twoSided && s->twoSided[li], hitPos, o, d);
                    pdfL *= 1.0f / (float)nLights;
                    float mis = prevPdf + pdfL;
                    thr = s_mul(thr, prevPdf);
                    thr = (mis == 0) ? S(0, 0, 0, 0) : s_mul(thr, 1.0f / mis);
                }
                if(depth + 1u <= s->rrHi) radiance = s_add(radiance, s_mulv(light_emit(s, li, gN, mul(d, -1.0f), waves), thr));
            }
            break;
        }
        const float dDotN = dot(gN, nrm(d));
        const int backSide = dDotN > 0;
        cone_surf cs = {cone, cone, 0.0f}; float dpdx[2] = {0, 0}, dpdy[2] = {0, 0};
        if(useCones)
        {   /* KCRenderWork (RenderWork.kt.h:L65): the cone arrives advanced by the hit distance */
            const uint32_t* vi = s->idx + 3 * (size_t)prim;
            static const float zero2[2] = {0, 0};
            const float* q0 = s->vertexTBN ? s->vertexTBN + 4 * (size_t)vi[0] : NULL; const float* q1 = s->vertexTBN ? s->vertexTBN + 4 * (size_t)vi[1] : NULL;
            const float* q2 = s->vertexTBN ? s->vertexTBN + 4 * (size_t)vi[2] : NULL;
            cone_surface(p, q0, q1, q2, s->uv ? s->uv + 2 * (size_t)vi[0] : zero2, s->uv ? s->uv + 2 * (size_t)vi[1] : zero2, s->uv ? s->uv + 2 * (size_t)vi[2] : zero2,
                         a, b, hitPos, backSide ? mul(gN, -1.0f) : gN, dDotN, nrm(d), cone_advance(cone, t), &cs, dpdx, dpdy);
        }
        v3 sN = gN;   /* shading normal: the interpolated tangent frame's Z axis when the group has a NORMAL attribute */
        int normalMapped = 0;
        if(s->vertexTBN)
        {
            const uint32_t* vi = s->idx + 3 * (size_t)prim;
            const float* q0 = s->vertexTBN + 4 * (size_t)vi[0]; const float* q1 = s->vertexTBN + 4 * (size_t)vi[1]; const float* q2 = s->vertexTBN + 4 * (size_t)vi[2];
            if(s->normalTexture && s->normalTexture[m] >= 0)
            {
                float u = 0, v = 0, px3[3];
                if(s->uv) { u = s->uv[2 * vi[0]] * a + s->uv[2 * vi[1]] * b + s->uv[2 * vi[2]] * c; v = s->uv[2 * vi[0] + 1] * a + s->uv[2 * vi[1] + 1] * b + s->uv[2 * vi[2] + 1] * c; }
                orc_texture_sample_grad(&s->textures[s->normalTexture[m]], u, v, dpdx, dpdy, s->textureLodMode, px3);
                sN = nrm(tbn_normal_mapped(q0, q1, q2, a, b, nrm(V(px3[0], px3[1], px3[2])), backSide));
                normalMapped = 1;
            }
            else sN = tbn_normal(q0, q1, q2, a, b);
        }
        if(backSide) { gN = mul(gN, -1.0f); if(!normalMapped) sN = mul(sN, -1.0f); }
        if(s->materialType && s->materialType[m] == 1u)
        {   /* (Mt)Reflect (MaterialsDefault.hpp:L132-215): a perfect mirror, Specularity() = 1. WorkFunctionNEE samples
             * a light (three random numbers) but casts no shadow ray for a specular material; WorkFunction reflects
             * wO about the shading normal (Graphics::Reflect: 2 (v.n) n - v), reflectance 1, pdf 1, NO Russian roulette,
             * and marks the ray SPECULAR_RAY so that a light it hits counts in full (no MIS, counted under pure NEE). */
            if(s->sampleMode != 0u) { pcg_float(rng); pcg_float(rng); pcg_float(rng); }
            v3 wO = mul(nrm(d), -1.0f);
            v3 wI = nrm(sub(mul(sN, 2.0f * dot(wO, sN)), wO));
            depth += 1;
            if(depth >= s->rrHi) break;
            prevPdf = 1.0f; type = 1; /* SPECULAR_RAY */
            cone = cone_after_scatter(&cs, wI, gN);
            o = nudge(hitPos, gN); d = wI; tMin = 1.0e-4f; tMax = FLT_MAX;
            continue;
        }
        if(s->materialType && s->materialType[m] == 2u)
        {   /* (Mt)Refract (MaterialsDefault.hpp:L246-312): Fresnel-weighted reflection / refraction, reflectance = pdf (cancels),
             * Specularity() = 1 like the mirror; a refraction disperses a spectral path to its first wavelength. The BxDF sample
             * is nudged along the shading normal by the material and again along the (flipped) geometric normal by the
             * work function (PathTracerRendererShaders.h:L277-283). */
            if(s->sampleMode != 0u) { pcg_float(rng); pcg_float(rng); pcg_float(rng); }
            const float* mp = s->materialParams + 8 * (size_t)m;
            int back = backSide;
            float fromEta = s->spectrum ? cauchy_ior(waves[0], mp) : mp[0], toEta = s->spectrum ? cauchy_ior(waves[0], mp + 4) : mp[4];
            if(back) { float tt = fromEta; fromEta = toEta; toEta = tt; }
            v3 wO = mul(nrm(d), -1.0f);
            float cosT = fabsf(dot(wO, sN));
            float f = fresnel_dielectric(cosT, fromEta, toEta);
            int refl = pcg_float(rng) < f;
            v3 wI;
            if(refl) wI = sub(mul(sN, 2.0f * dot(wO, sN)), wO);
            else
            {
                float er = fromEta / toEta, cosIn = dot(sN, wO), sinIn2 = 1.0f - cosIn * cosIn; if(sinIn2 < 0) sinIn2 = 0;
                float cosOut = sqrt_max(1.0f - er * er * sinIn2);
                wI = add(mul(wO, -er), mul(sN, er * cosIn - cosOut));
                if(s->spectrum) { waves[1] = waves[2] = waves[3] = -1.0f; }
            }
            float pdfS = refl ? f : 1.0f - f;
            depth += 1;
            if(depth >= s->rrHi) break;
            throughput = (pdfS == 0) ? S(0, 0, 0, 0) : s_mul(s_mul(throughput, pdfS), 1.0f / pdfS);
            prevPdf = pdfS; type = 1; /* SPECULAR_RAY */
            if(useCones) { cone_surf cr = refract_ray_cone(cs, wO, gN, fromEta, toEta); cone = cone_after_scatter(&cr, nrm(wI), gN); }
            o = nudge(nudge(hitPos, sN), refl ? gN : mul(gN, -1.0f)); d = nrm(wI); tMin = 1.0e-4f; tMax = FLT_MAX;
            continue;
        }
        /* Lambert / Unreal */
        s4 alb = albedo_at(s, (uint32_t)m, waves, prim, a, b, c, dpdx, dpdy);
        v3 hlp = fabsf(sN.x) > 0.9f ? V(0, 1, 0) : V(1, 0, 0);
        v3 tX = nrm(cross(hlp, sN)), tY = cross(sN, tX);
        const int unreal = s->materialType && s->materialType[m] == 3u;
        unreal_mat um; um.albedo = alb; um.roughness = um.specular = um.metallic = 0;
        if(unreal) { const float* mp = s->materialParams + 8 * (size_t)m; um.roughness = mp[0]; um.specular = mp[1]; um.metallic = mp[2]; }
        const int specularMat = unreal && (1.0f - unreal_mis_ratio(&um)) >= 0.95f;   /* MaterialCommon::IsSpecular */
        v3 wOw = mul(nrm(d), -1.0f);
        v3 Vt = V(dot(wOw, tX), dot(wOw, tY), dot(wOw, sN));
        if(s->sampleMode != 0u)
        {   /* NEE */
            float x0 = pcg_float(rng), x1 = pcg_float(rng), xs = pcg_float(rng);
            uint32_t li = (uint32_t)(xs * (float)nLights);
            if(li > nLights - 1u) li = nLights - 1u;
            if((li < s->nLightTris || s->boundaryType != 0u) && !specularMat)
            {
                v3 lpos; float pdfL; s4 em;
                if(li < s->nLightTris)
                {
                    uint32_t lt = s->lightTris[li]; uint32_t lightIdx = (uint32_t)(-1 - s->triMaterial[lt]);
                    v3 q[3]; tri(s, lt, q);
                    v3 lN, sd;
                    pdfL = light_sample(q, s->twoSided && s->twoSided[lightIdx], x0, x1, hitPos, &lpos, &lN, &sd);
                    pdfL *= 1.0f / (float)nLights;
                    em = light_emit(s, lightIdx, lN, sd, waves);
                }
                else
                {   /* the skysphere is the last meta light (MetaLight.hpp:L512-518) */
                    pdfL = sky_sample(s, x0, x1, hitPos, &lpos);
                    pdfL *= 1.0f / (float)nLights;
                    em = sky_emit(s, nrm(sub(hitPos, lpos)), waves);
                }
                v3 wI = nrm(sub(lpos, hitPos));
                v3 lposN = nudge(lpos, mul(wI, -1.0f));
                float length = len(sub(lposN, hitPos));
                float nDotL = dot(sN, wI); if(nDotL < 0) nDotL = 0;
                s4 refl = s_mul(alb, nDotL * 0.31830988618f);
                float pdfBx = pdf_cos_direction(dot(sN, wI));
                if(unreal)
                {
                    v3 Lt = V(dot(wI, tX), dot(wI, tY), dot(wI, sN));
                    refl = unreal_evaluate(&um, Vt, Lt); pdfBx = unreal_pdf(&um, Vt, Lt);
                }
                float pdf = pdfL;
                if(s->sampleMode == 2u) pdf = pdfBx + pdfL;
                s4 sr = s_mulv(s_mulv(throughput, refl), em);
                sr = (pdf == 0) ? S(0, 0, 0, 0) : s_mul(sr, 1.0f / pdf);
                if(depth + 2u <= s->rrHi && (sr.v[0] > 0 || sr.v[1] > 0 || sr.v[2] > 0 || sr.v[3] > 0))
                {
                    v3 so = nudge(hitPos, gN);
                    uint32_t sp; float st, sb[2];
                    if(!trace_rng(s, rng, so, wI, 1.0e-5f, length * (1.0f - 1.0e-4f), 1, &sp, &st, sb)) radiance = s_add(radiance, sr);
                }
            }
        }
        /* BxDF sample + RR */
        float lx, ly, lz, pdfB; s4 reflS;
        if(unreal)
        {
            float sXi = pcg_float(rng), u0 = pcg_float(rng), u1 = pcg_float(rng);
            v3 L = unreal_sample(&um, Vt, sXi, u0, u1, &reflS, &pdfB);
            lx = L.x; ly = L.y; lz = L.z;
        }
        else
        {
            float u0 = pcg_float(rng), u1 = pcg_float(rng);
            float cd[4]; cos_direction(u0, u1, cd);
            lx = cd[0]; ly = cd[1]; lz = cd[2]; pdfB = cd[3];
            reflS = s_mul(alb, lz * 0.31830988618f);
        }
        v3 wI = nrm(add(add(mul(tX, lx), mul(tY, ly)), mul(sN, lz)));
        throughput = s_mulv(throughput, reflS);
        depth += 1;
        int dead = depth >= s->rrHi;
        if(!dead && depth >= s->rrLo && !specularMat)
        {
            float xi = pcg_float(rng);
            float prob = (throughput.v[0] + throughput.v[1] + throughput.v[2] + throughput.v[3]) * (s->spectrum ? 0.25f : 0.33333333f);
            prob = prob < 0.1f ? 0.1f : (prob > 1.0f ? 1.0f : prob);
            if(xi >= prob) dead = 1; else throughput = s_mul(throughput, 1.0f / prob);
        }
        if(dead) break;
        throughput = (pdfB == 0) ? S(0, 0, 0, 0) : s_mul(throughput, 1.0f / pdfB);
        prevPdf = pdfB; type = specularMat ? 1 : 2; /* SPECULAR_RAY after a near-mirror, else PATH_RAY */
        cone = cone_after_scatter(&cs, wI, gN);
        o = nudge(hitPos, gN); d = wI; tMin = 1.0e-4f; tMax = FLT_MAX;
    }
    return radiance;
}

/* Renders rows [y0, y1) into out (planar R,G,B,W sums of the FULL image, row 0 = bottom). Thread safe
 * for disjoint row ranges. RNG: one PCG32 stream per pixel, state = GenerateState(seed32 ^ hash(pixel)). */
void orc_pt_render_rows(const pt_scene* s, uint32_t y0, uint32_t y1, float* out)
{
    size_t plane = (size_t)s->width * s->height;
    uint32_t seed32 = (uint32_t)((s->seed >> 32) ^ (s->seed & 0xFFFFFFFFull));
    for(uint32_t y = y0; y < y1; y++)
    for(uint32_t x = 0; x < s->width; x++)
    {
        pcg rng; uint32_t h = (y * s->width + x) * 2654435761u ^ seed32 ^ 0x9E3779B9u;
        rng.s = 0u * 747796405u + 2891336453u; rng.s += h; rng.s = rng.s * 747796405u + 2891336453u;
        double acc[3] = {0, 0, 0}, w = 0;
        for(uint32_t k = 0; k < s->spp; k++)
        {
            float fw; v3 r = path(s, &rng, x, y, &fw);
            acc[0] += r.x; acc[1] += r.y; acc[2] += r.z; w += fw;
        }
        size_t pix = (size_t)y * s->width + x;
        out[pix] = (float)acc[0]; out[plane + pix] = (float)acc[1]; out[2 * plane + pix] = (float)acc[2]; out[3 * plane + pix] = (float)w;
    }
}
