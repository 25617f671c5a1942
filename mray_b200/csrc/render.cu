// render.cu — wavefront path tracer (RGB, Lambert, prim-backed triangle lights, NEE + MIS, Russian
// roulette, stochastic film filter). One wavefront iteration advances every live path by one bounce:
//
//   KReload   : free slots take the next camera path of the tile (device counter, no host sync),
//               stochastic-filter camera ray, path state init           (PathTracerRendererBase::ReloadPaths,
//               Tracer/PathTracerRendererBase.cu:L69-203; KCGenerateCamRaysStochastic, RayGenKernels.kt.h:L155-227;
//               CameraPinhole::EvaluateRay, CamerasDefault.hpp:L93-142; KCInitPathStateIndirect, L34-57)
//   TraceRays : closest hit                                              (BaseAcceleratorLBVH::CastRays)
//   KShade    : boundary / light hit (MIS re-weighting) / Lambert: NEE shadow ray + BxDF sample + RR
//               (TracerDLL/PathTracerRendererShaders.h:L201-520; LambertMaterial, Tracer/MaterialsDefault.hpp:L25-126;
//               LightPrim, Tracer/LightsDefault.hpp:L22-168; DirectLightSamplerUniform, LightSampler.hpp:L30-119;
//               Triangle::SampleSurface, PrimitiveDefaultTriangle.hpp:L48-77; RussianRoulette, DistributionFunctions.h:L943-953)
//   TraceRays : any hit on the shadow rays                               (CastVisibilityRays)
//   KFinish   : add the pre-multiplied NEE radiance of visible shadow rays (KCAccumulateShadowRaysPT,
//               PathTracerRenderer.cu:L11-32); dead paths go to the film with atomics
//               (KCSetImagePixelsIndirectAtomic + ConvertNaNsToColor, Tracer/TextureFilter.cu:L114-121,L669-697)
//               and free their slot.
//
// The reference runs NEE and BxDF sampling as two partitioned passes with a random-number buffer in
// between; both only depend on the hit, so they are one kernel here and the PCG32 stream is consumed
// in registers. Host synchronisations per iteration: none.
#include "accel.cuh"
#include "spectrum.cuh"
#include "sampler.cuh"
#include "dist2d.cuh"
#include <cfloat>
#include <stdexcept>
#include <cstring>
#include <vector>
#include <map>
#include <set>

namespace mrb
{

void TraceRays(Context& ctx, const mrb_accel_t& acc, bool anyHit, mrb_trace_mode mode,
               mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
               mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount);
void TraceScene(Context& ctx, const SceneData& scn, bool anyHit, mrb_trace_mode mode,
                mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
                mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount);
size_t MultiPartitionTempBytes(uint32_t count, uint32_t batchBits);
void MultiPartition(Context& ctx, uint32_t* keys, uint32_t* indices, uint32_t count,
                    const uint32_t dataBits[2], const uint32_t batchBits[2], bool onlySortForBatches,
                    uint32_t maxPartitions, uint32_t* outCount, uint32_t* outOffsets, uint32_t* outKeys, void* temp);

namespace
{

constexpr int RTPB = 256;
constexpr float PI_F = 3.14159265358979323846f;
constexpr float INV_PI_F = 0.31830988618379067154f;

enum : uint32_t { ST_INVALID = 0u, ST_ALIVE = 1u, ST_DEAD = 2u };
enum : uint32_t { RAY_SHADOW = 0u, RAY_SPECULAR = 1u, RAY_PATH = 2u, RAY_CAMERA = 3u }; // RayType, PathTracerRendererBase.h:L11-17

struct Float3 { float x, y, z; };
__device__ __forceinline__ Float3 F3(float x, float y, float z) { return Float3{x, y, z}; }
__device__ __forceinline__ Float3 operator+(Float3 a, Float3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ Float3 operator-(Float3 a, Float3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ Float3 operator*(Float3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float Dot(Float3 a, Float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Float3 Cross(Float3 a, Float3 b)
{ return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float Length(Float3 a) { return sqrtf(Dot(a, a)); }
__device__ __forceinline__ Float3 Normalize(Float3 a) { return a * (1.0f / Length(a)); }

// Spectrum = Vector<4, Float> (Core/Definitions.h:L238-240): (r, g, b, 0) for the RGB renderer, 4 hero
// wavelength samples for the spectral one
struct Spec { float x, y, z, w; };
__device__ __forceinline__ Spec S4(float x, float y, float z, float w) { return Spec{x, y, z, w}; }
__device__ __forceinline__ Spec S4(float4 v) { return Spec{v.x, v.y, v.z, v.w}; }
__device__ __forceinline__ float4 F4(Spec v) { return make_float4(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ Spec operator*(Spec a, Spec b) { return S4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ Spec operator*(Spec a, float s) { return S4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ Spec operator+(Spec a, Spec b) { return S4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// PermutedCG32 (Tracer/Random.h:L237-238,L763-812): 32-bit state LCG, RXS-M-XS output
struct PCG32
{
    uint32_t s;
    __device__ __forceinline__ uint32_t Next()
    {
        uint32_t old = s;
        s = old * 747796405u + 2891336453u;
        uint32_t r = ((old >> ((old >> 28u) + 4u)) ^ old) * 277803737u;
        return (r >> 22u) ^ r;
    }
    // RNGFunctions::ToFloat01 (Random.h:L102-118)
    __device__ __forceinline__ float NextFloat() { return fminf(float(Next()) * 2.3283064365386963e-10f, 0.99999994f); }
};

// The sample source of one path slot: the independent PCG32 stream, or dimension-indexed (Z-)Sobol points
// (request sizes follow RNRequestList: a 2-D / 3-D request draws correlated dimensions from one scramble hash).
struct SlotSampler
{
    uint32_t type, state, sampleIndex, dim;
    bool reverseBack;
    uint64_t morton;
    const uint32_t* matrices;
    ZSobolGlobals g;
    template<int N>
    __device__ __forceinline__ void Next(float* out)
    {
        if(type == SAMPLER_INDEPENDENT)
        {
            PCG32 pcg{state};
            #pragma unroll
            for(int k = 0; k < N; k++) out[k] = pcg.NextFloat();
            state = pcg.s;
            return;
        }
        uint32_t v[3];
        if(type == SAMPLER_SOBOL) SobolNext(matrices, state, sampleIndex, dim, N, v, reverseBack);
        else ZSobolNext(matrices, state, sampleIndex, morton, g, dim, N, v, reverseBack);
        dim += uint32_t(N);
        #pragma unroll
        for(int k = 0; k < N; k++) out[k] = fminf(float(v[k]) * 2.3283064365386963e-10f, 0.99999994f);
    }
};
// `gpx, gpy`: the pixel in FULL-image coordinates (a tile / region renders the same points as the whole image would)
__device__ __forceinline__ SlotSampler LoadSampler(uint32_t type, uint32_t state, uint2 ss, uint32_t gpx, uint32_t gpy,
                                                   const uint32_t* matrices, ZSobolGlobals g)
{
    SlotSampler s;
    s.reverseBack = (type & SAMPLER_REFERENCE_SCRAMBLE) == 0u; type &= SAMPLER_TYPE_MASK;
    s.type = type; s.state = state; s.sampleIndex = ss.x; s.dim = ss.y; s.matrices = matrices; s.g = g;
    s.morton = (type == SAMPLER_ZSOBOL) ? Morton2D(gpx, gpy) : 0ull;
    return s;
}

// Seeding. The reference seeds one generator per path SLOT from a host std::mt19937 (RNGGroupIndependent,
// Tracer/Random.cu:L661-720) and lets it run on from path to path, so which numbers a sample sees depends on
// which slot happened to pick it up. Here every (pixel, sample index) pair owns its numbers: the PCG32 state of
// a new path (Independent) or the pixel's scramble seed (Sobol / ZSobol) is a hash of (seed, full-image pixel,
// sample index), so a fixed seed gives a fixed image whatever the slot scheduling, the tiling or the number of
// GPUs the sample range is split over (statistical parity with the reference, as SURVEY.md §8a-16 allows).
__device__ __forceinline__ uint64_t Mix64(uint64_t z)
{   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t PixelSeed(uint64_t seed, uint32_t gpix)
{ return uint32_t(Mix64(seed ^ (uint64_t(gpix) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull)) >> 32); }
__device__ __forceinline__ uint32_t PathState(uint64_t seed, uint32_t gpix, uint32_t sample)
{
    const uint32_t xi = uint32_t(Mix64(Mix64(seed + uint64_t(gpix) * 0x9E3779B97F4A7C15ull) ^ (uint64_t(sample) * 0xD1B54A32D192ED03ull)) >> 32);
    uint32_t s = 2891336453u;            // PermutedCG32::GenerateState (Random.h:L763-771): Step(0), += xi, Step
    s += xi;
    return s * 747796405u + 2891336453u;
}

// Ray::Nudge (Core/Ray.hpp:L258-301, after RT Gems I ch. 6)
__device__ __forceinline__ Float3 NudgePos(Float3 p, Float3 n)
{
    const float ORIGIN = 1.0f / 32.0f, FLOAT_SCALE = 1.0f / 65536.0f, INT_SCALE = 256.0f;
    int ox = int(INT_SCALE * n.x), oy = int(INT_SCALE * n.y), oz = int(INT_SCALE * n.z);
    float ix = __int_as_float(__float_as_int(p.x) + ((p.x < 0.f) ? -ox : ox));
    float iy = __int_as_float(__float_as_int(p.y) + ((p.y < 0.f) ? -oy : oy));
    float iz = __int_as_float(__float_as_int(p.z) + ((p.z < 0.f) ? -oz : oz));
    return F3(fabsf(p.x) < ORIGIN ? p.x + FLOAT_SCALE * n.x : ix,
              fabsf(p.y) < ORIGIN ? p.y + FLOAT_SCALE * n.y : iy,
              fabsf(p.z) < ORIGIN ? p.z + FLOAT_SCALE * n.z : iz);
}

// Film filters (Tracer/Filters.h): Sample(xi) -> offset + pdf, Evaluate(offset); the film weight of a camera sample is
// Evaluate / pdf (KCGenerateCamRaysStochastic, RayGenKernels.kt.h:L195-225). Box, Tent and Gaussian are sampled from
// their own shape (weight 1 up to rounding); Mitchell-Netravali is sampled from a 3-Gaussian mixture, so its weight
// varies and can be negative.
enum : uint32_t { FILTER_BOX = 0u, FILTER_TENT = 1u, FILTER_GAUSSIAN = 2u, FILTER_MITCHELL = 3u };   // FilterType::E (Core/TracerEnums.h:L162-173)
constexpr float PREV_ONE = 0.99999994f;

__device__ __forceinline__ float GaussPdf(float x, float sigma, float mu = 0.0f)
{   // Math::Gaussian (Core/Math.h:L1032-1042)
    const float si = 1.0f / sigma, p = (x - mu) * si;
    return 0x1.988452p-2f * si * expf(-0.5f * p * p);   // InvSqrt2Pi = (1 / Sqrt2<float>) * (1 / SqrtPi<float>) as the reference's constexpr evaluates it
}
__device__ __forceinline__ float GaussSample(float xi, float sigma, float mu, float& pdf)
{   // Distribution::Common::SampleGaussian (DistributionFunctions.h:L686-705)
    const float e = erfinvf(2.0f * xi - 1.0f);
    float x = 1.41421356237f * sigma * e + mu;
    if(isinf(e)) x = fminf(fmaxf(x, -3.5f * sigma), 3.5f * sigma);
    pdf = GaussPdf(x, sigma, mu);
    return x;
}
__device__ __forceinline__ float LerpU(float a, float b, float t) { return __fadd_rn(__fmul_rn(a, 1.0f - t), __fmul_rn(b, t)); }   // Math::Lerp
__device__ __forceinline__ float TentSample(float xi, float r, float& pdf)
{   // Distribution::Common::SampleTent(xi, -r, r) (DistributionFunctions.h:L785-805) over BisectSample2 + SampleLine(., 1, 0)
    const float a = -r, b = r;
    if(b - a < 1.0e-4f) { pdf = 1.0f / (b - a); return 0.0f; }   // MathConstants::LargeEpsilon
    const float w = r / (r + r);                                  // weights[0] / weights.Sum()
    const bool left = xi < w;
    float lxi = left ? xi / w : (xi - w) / (1.0f - w);
    lxi = fminf(lxi, PREV_ONE);
    if(left) lxi = PREV_ONE - lxi;
    const float denom = 1.0f + sqrtf(fmaxf(LerpU(1.0f, 0.0f, lxi), 0.0f));
    const float x = fminf(lxi / denom, PREV_ONE);
    pdf = 2.0f * LerpU(1.0f, 0.0f, x) * (1.0f / (b - a));
    return left ? x * a : x * b;
}
__device__ __forceinline__ float Mitchell1D(float x, float radiusRecip)
{   // MitchellNetravaliFilter::Evaluate, b = c = 0.33333 (Filters.h:L258-300)
    const float B = 0.33333f, C = 0.33333f, F = 1.0f / 6.0f;
    x = fabsf(__fmul_rn(__fmul_rn(2.0f, x), radiusRecip));
    const float x2 = __fmul_rn(x, x), x3 = __fmul_rn(x2, x);
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    if(x < 1.0f) { c0 = F * (12.0f - 9.0f * B - 6.0f * C); c1 = F * (-18.0f + 12.0f * B + 6.0f * C); c3 = F * (6.0f - 2.0f * B); }
    else if(x < 2.0f) { c0 = F * (-B - 6.0f * C); c1 = F * (6.0f * B + 30.0f * C); c2 = F * (-12.0f * B - 48.0f * C); c3 = F * (8.0f * B + 24.0f * C); }
    // unfused, left to right, as the reference's host build evaluates it (mip levels filtered with this kernel match bit for bit)
    const float poly = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, x3), __fmul_rn(c1, x2)), __fmul_rn(c2, x)), c3);
    return __fmul_rn(__fmul_rn(poly, 2.0f), radiusRecip);
}
__device__ __forceinline__ float MitchellSampleDim(float xi, float r, float& pdf)
{   // MitchellNetravaliFilter::Sample, one axis (Filters.h:L302-346): balance-heuristic mixture of three Gaussians
    const float MIS_MID = 0.960566188838f, MIS_SIDES = 0.0197169055809f;
    const float midSigma = 0.528f * r * 0.5f, sideSigma = 0.2f * r * 0.5f, sideMean = 1.3f * r * 0.5f;
    // BisectSample<3>(xi, {SIDES, MID, SIDES}, normalised) (DistributionFunctions.h:L645-683)
    uint32_t index; float lo, wsel;
    if(xi < MIS_SIDES) { index = 0u; lo = 0.0f; wsel = MIS_SIDES; }
    else if(xi < MIS_SIDES + MIS_MID) { index = 1u; lo = MIS_SIDES; wsel = MIS_MID; }
    else { index = 2u; lo = MIS_SIDES + MIS_MID; wsel = MIS_SIDES; }
    const float lxi = fminf((xi - lo) / wsel, PREV_ONE);
    float own, x;
    if(index == 0u) x = GaussSample(lxi, sideSigma, -sideMean, own);
    else if(index == 1u) x = GaussSample(lxi, midSigma, 0.0f, own);
    else x = GaussSample(lxi, sideSigma, sideMean, own);
    const float p0 = (index == 0u) ? own : GaussPdf(x, sideSigma, -sideMean);
    const float p1 = (index == 1u) ? own : GaussPdf(x, midSigma);
    const float p2 = (index == 2u) ? own : GaussPdf(x, sideSigma, sideMean);
    pdf = p0 * MIS_SIDES + p1 * MIS_MID + p2 * MIS_SIDES;          // MIS::BalanceCancelled<3>
    return x;
}
// out: offset (x, y), pdf of Sample(), Evaluate(offset)
__device__ __forceinline__ void FilterSample(uint32_t type, float r, float xi0, float xi1, float& ox, float& oy, float& pdf, float& eval)
{
    float px, py;
    if(type == FILTER_BOX)
    {   // Common::SampleUniformRange(xi, -r, r) per axis; Evaluate = 1 / (4 r^2) inside the square (Filters.h:L112-146)
        const float range = r - (-r);
        ox = xi0 * range + (-r); oy = xi1 * range + (-r);
        px = py = 1.0f / range;
        const float rr = 1.0f / r;
        eval = (fabsf(ox) <= r && fabsf(oy) <= r) ? 0.25f * rr * rr : 0.0f;
    }
    else if(type == FILTER_TENT)
    {   // Evaluate = Lerp(1/r, 0, |x| / r) Lerp(1/r, 0, |y| / r) (Filters.h:L160-172)
        ox = TentSample(xi0, r, px); oy = TentSample(xi1, r, py);
        const float rcp = 1.0f / r, cap = 1.0f / r;
        eval = LerpU(cap, 0.0f, fminf(fabsf(ox * rcp), 1.0f)) * LerpU(cap, 0.0f, fminf(fabsf(oy * rcp), 1.0f));
    }
    else if(type == FILTER_MITCHELL)
    {
        ox = MitchellSampleDim(xi0, r, px); oy = MitchellSampleDim(xi1, r, py);
        const float rcp = 1.0f / r;
        eval = Mitchell1D(ox, rcp) * Mitchell1D(oy, rcp);
    }
    else
    {   // GaussianFilter: sigma = 0.285714 r; Evaluate is the sampling density itself (Filters.h:L195-227)
        const float sigma = r * 0.285714f;
        ox = GaussSample(xi0, sigma, 0.0f, px); oy = GaussSample(xi1, sigma, 0.0f, py);
        eval = px * py;
    }
    pdf = px * py;
}

struct Camera // CameraPinhole members (CamerasDefault.hpp:L8-36), tile-local
{
    Float3 position, right, up, bottomLeft;
    float  planeW, planeH, tNear, tFar;
};

// One 2-D texture (single mip level) as the reference's host-backend view sees it (Device/CPU/TextureViewCPU.h)
// `data` holds mipCount levels back to back: level k at texel offset MipStart(k), MipDim(w, k) x MipDim(h, k) texels
// (Graphics::TextureMipSize / TextureMipPixelStart, Core/GraphicsFunctions.h:L474-525 — the layout of the reference's host texture)
struct TexRec { const void* data; uint32_t w, h, channels, format, interp, edge, mipCount; };
__host__ __device__ __forceinline__ uint32_t MipDim(uint32_t n, uint32_t level) { const uint32_t v = n >> level; return v ? v : 1u; }
__host__ __device__ __forceinline__ size_t MipStart(uint32_t w, uint32_t h, uint32_t level)
{ size_t o = 0; for(uint32_t i = 0; i < level; i++) o += size_t(MipDim(w, i)) * MipDim(h, i); return o; }
static uint32_t FullMipCount(uint32_t w, uint32_t h) { uint32_t m = max(w, h), c = 0; while(m) { c++; m >>= 1; } return c; }   // Graphics::TextureMipCount

struct EmissiveTri { float4 p0, e0, e1; float4 radiance; }; // p0.w = area, e0.w = twoSided, radiance.w unused

// Per-instance shading inputs: the primitive group's arrays plus the instance transform
// (PrimitiveC::GenerateSurface in local space + TransformContextSingle::Apply / ApplyN).
struct RenderInstance
{
    const float*    positions;
    const uint32_t* indices;
    const float4*   triPos;          // 3 x float4 per primitive: its vertex positions gathered once per renderer (one load instead of index -> position)
    const float4*   triNrm;          // ... and, where the instance has them, its three vertex normals / tangent frames
    const float2*   triUV;           // ... and UVs
    const float4*   vertexNormals;   // optional: shading normals (xyz), or to-tangent-space quaternions (w, x, y, z) when `tbn` is set; nullptr = geometric
    const uint32_t* lightOfPrim;     // prim index -> emissive triangle index or INVALID
    const float2*   vertexUVs;       // optional UV0 per vertex, nullptr = (0, 0)
    float           transform[12];   // local -> world, row-major 3x4
    float           invTransform[12];
    uint32_t        identity;
    uint32_t        tbn;             // vertexNormals holds quaternions (PrimGroupTriangle's NORMAL attribute as the loader delivers it)
};

struct RenderData
{
    // scene: one record per instance (a single record when rendering one accelerator)
    const RenderInstance* instances;
    uint32_t          sceneMode;       // 1 = hitKeys.accelKey holds the instance index
    const TexRec*     textures;        // textured albedo (ParamVaryingData): texture table ...
    const int32_t*    albedoTex;       // ... and per material index: texture or -1; nullptr = no material is textured
    const int32_t*    normalTex;       // per material index: tangent-space normal map (texture index) or -1; nullptr = none
    float2*           cones;           // per slot ray cone (aperture, width); nullptr unless a texture has more than one mip level
    uint32_t          textureLodMode;  // SampleTextureGrad: 0 = UV-space gradients (reference host backend), 1 = texel-space (tex2DGrad)
    const uint8_t*    materialType;    // per material index: mrb_material_type; nullptr = all Lambert
    const float4*     matParams;       // per material index 2 x float4: Refract (cauchyFront xyz, cauchyBack xyz), Unreal (roughness, specular, metallic)
    const float4*     albedo;          // per material index: (r, g, b, 0), or Jakob coefficients (c0, c1, c2, -) when spectral
    // hero-wavelength spectral transport ((R)PathTracerSpectral); lights carry (c0, c1, c2, scale) in .radiance
    SpectrumData      spec;
    uint32_t          spectral;
    float4*           waves;           // per slot: 4 wavelengths (nm)
    float4*           wavePdf;         // per slot: their pdfs
    const EmissiveTri* lights;         // one per emissive triangle (MetaLight list), world space
    uint32_t          lightCount;      // emissive triangles (+1 boundary light in the sampler)
    // boundary light: 0 = (L)Null, 1 = (L)Skysphere_Spherical, 2 = (L)Skysphere_CoOcta (Tracer/LightsDefault.hpp:L310-443)
    uint32_t          boundaryType;
    int32_t           boundaryTex;     // radiance texture (index into textures) or -1 = constant
    float4            boundaryRadiance;// constant radiance: rgb, or Jakob coefficients + scale when spectral
    Dist2D            boundaryDist;    // luminance distribution of the radiance texture (textured skysphere only)
    float             boundaryM[9], boundaryInvM[9]; // linear part of the light surface's transform and its inverse (ApplyV / InvApplyV)
    uint32_t          boundaryIdentity;
    float             sceneDiameter;
    Camera            cam;
    uint32_t          width, height;   // the tile (region) of the current pass
    uint32_t          fullWidth, fullHeight, regionX, regionY; // the image it is a region of
    uint32_t          filterType;      // FilterType::E: Box, Tent, Gaussian, Mitchell-Netravali (TracerParameters.filmFilter)
    float             filterRadius;
    // options
    uint32_t          rrLo, rrHi, sampleMode; // 0 Pure, 1 NEE, 2 NEE+MIS
    uint64_t          pathLimit;       // camera paths of the current pass (samples x tile pixels)
    uint32_t          sampleBase;      // sample index of the pass's first sample of every pixel
    uint64_t          seed;            // TracerParameters.seed
    // path state (P slots)
    uint32_t          slots;
    mrb_ray_gmem*     rays;
    mrb_ray_gmem*     shadowRays;
    mrb_hit_key_pack* hitKeys;
    mrb_meta_hit*     hits;
    float4*           throughput;
    float4*           radiance;
    float4*           shadowRadiance;
    uint4*            meta;            // x: pathData (depth | status << 8 | type << 16), y: pixel, z: film weight, w: previous bxdf pdf
    uint32_t*         rng;             // Independent: PCG32 state; Sobol / ZSobol: the generator's scramble seed (constant)
    // low-discrepancy samplers (samplerType != 0): one generator per PIXEL (LocalState.seed = PixelSeed());
    // path g renders pixel g % N as that pixel's sample sampleBase + g / N, so the generator state is implicit
    // (the reference keeps its generators per path slot, which only coincides with the pixel while slots and
    // pixels stay aligned; per pixel keeps the stratification of the sequence inside every pixel)
    uint32_t          samplerType;     // 0 Independent (PCG32), 1 Sobol, 2 ZSobol
    uint2*            sampleState;     // per slot: x = sample index of the current path, y = next free dimension
    const uint32_t*   sobolMatrices;   // 256 x 52 Joe-Kuo generator matrices
    ZSobolGlobals     zsobol;
    uint32_t*         visible;
    // material-key ray partitioning (RenderSurfaceWorkHasher + RayPartitioner::MultiPartition)
    uint32_t*         workKeys;        // per slot sort key
    uint32_t*         workIndices;     // slot indices, sorted by work key
    uint32_t*         partTable;       // [0] count, [1..] offsets (count+1), then first keys
    uint32_t          partitionRays;   // 0 = shade in slot order
    uint32_t          matBits;         // data bits of the work key (material index)
    // film (planar R,G,B,W)
    float*            film;
    // u64 counters: [0] next camera path of the pass, [1] completed paths, [2] closest-hit rays cast, [3] shadow rays cast,
    // [4] NEE light samples taken (= the shadow rays the reference casts: it also traces the zero-valued ones)
    unsigned long long* counters;
};

__device__ __forceinline__ uint32_t PackPD(uint32_t depth, uint32_t status, uint32_t type) { return depth | (status << 8) | (type << 16); }

// Reload of one slot; must be called by all 32 lanes of a warp (warp-aggregated claim). `isFree`: the slot
// holds no path (its meta.x status is INVALID).
__device__ __forceinline__ void ReloadClaimed(const RenderData& d, uint32_t i, unsigned long long g);
__device__ __forceinline__ void ReloadSlot(const RenderData& d, uint32_t i, bool inRange, bool isFree)
{
    // block-aggregated claim of new path indices: ONE atomic per block on the global counter (same-address
    // atomics serialise in L2; one per warp was the bottleneck of this kernel), ranks by ballot + a
    // per-warp prefix in shared memory. Must be reached by every thread of the block.
    __shared__ uint32_t sWarpFree[RTPB / 32];
    __shared__ unsigned long long sBlockBase;
    const uint32_t want = __ballot_sync(0xffffffffu, isFree);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if(lane == 0) sWarpFree[warp] = uint32_t(__popc(want));
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t total = 0;
        #pragma unroll
        for(int k = 0; k < RTPB / 32; k++) { const uint32_t c = sWarpFree[k]; sWarpFree[k] = total; total += c; }
        sBlockBase = total ? atomicAdd(&d.counters[0], (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    const unsigned long long base = sBlockBase + sWarpFree[warp];
    if(!inRange) return;
    if(!isFree)
    {
        d.hitKeys[i].primKey = INVALID_U32; // default: boundary (KCSetBoundaryWorkKeysIndirect)
        return;
    }
    ReloadClaimed(d, i, base + __popc(want & ((1u << lane) - 1u)));
}

// The slot's next path: camera path `g` of the pass (or nothing once the pass has handed out all its paths)
__device__ __forceinline__ void ReloadClaimed(const RenderData& d, uint32_t i, unsigned long long g)
{
    if(g >= d.pathLimit)
    {
        // surplus slot: never traced (KCWriteInvalidRaysIndirect)
        d.rays[i].tMin = 1.0f; d.rays[i].tMax = -1.0f;
        d.shadowRays[i].tMin = 1.0f; d.shadowRays[i].tMax = -1.0f;
        d.hitKeys[i].primKey = INVALID_U32;
        return;
    }
    const unsigned long long tilePixels = (unsigned long long)(d.width * d.height);
    const uint32_t pix = uint32_t(g % tilePixels);
    const uint32_t sample = d.sampleBase + uint32_t(g / tilePixels);
    const uint32_t px = pix % d.width, py = pix / d.width;
    const uint32_t gpx = px + d.regionX, gpy = py + d.regionY, gpix = gpy * d.fullWidth + gpx;
    // a new path = sample `sample` of its pixel, dimension 0: numbers depend on (seed, pixel, sample) only
    const bool lowDisc = d.samplerType != SAMPLER_INDEPENDENT;
    const uint2 ss = make_uint2(lowDisc ? sample : 0u, 0u);
    SlotSampler rng = LoadSampler(d.samplerType, lowDisc ? PixelSeed(d.seed, gpix) : PathState(d.seed, gpix, sample), ss, gpx, gpy,
                                  d.sobolMatrices, d.zsobol);
    // stochastic filter sample: offset ~ filter sampler, film weight = Evaluate / pdf
    float xiF[2]; rng.Next<2>(xiF);
    float offx, offy, fPdf, fEval;
    FilterSample(d.filterType, d.filterRadius, xiF[0], xiF[1], offx, offy, fPdf, fEval);
    // Gaussian: Evaluate is the sampling density itself, the quotient is exactly 1
    const float weight = (d.filterType == FILTER_GAUSSIAN) ? 1.0f : fEval / fPdf;
    // the tile may be a region of a larger image (RenderImageParams.regionMin / resolution)
    const float sx = (float(px + d.regionX) + offx + 0.5f) * (d.cam.planeW / float(d.fullWidth));
    const float sy = (float(py + d.regionY) + offy + 0.5f) * (d.cam.planeH / float(d.fullHeight));
    const Float3 point = d.cam.bottomLeft + d.cam.right * sx + d.cam.up * sy;
    const Float3 dir = Normalize(point - d.cam.position);
    float4* rp = reinterpret_cast<float4*>(d.rays + i);
    rp[0] = make_float4(d.cam.position.x, d.cam.position.y, d.cam.position.z, d.cam.tNear);
    rp[1] = make_float4(dir.x, dir.y, dir.z, d.cam.tFar);
    d.shadowRays[i].tMin = 1.0f; d.shadowRays[i].tMax = -1.0f;
    d.hitKeys[i].primKey = INVALID_U32;
    d.throughput[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    d.radiance[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    d.shadowRadiance[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    d.meta[i] = make_uint4(PackPD(0, ST_ALIVE, RAY_CAMERA), pix, __float_as_uint(weight), __float_as_uint(0.0f));
    // CameraPinhole::EvaluateRay (CamerasDefault.hpp:L134-139): aperture = 2 tan(fovY / 2) / resolution.y, width 0
    if(d.cones) d.cones[i] = make_float2(d.cam.planeH / d.cam.tNear / float(d.fullHeight), 0.0f);
    if(d.spectral)
    {
        // one more dimension after the camera sample (PathTracerRendererBase.cu:L139-168)
        float w[4], p[4];
        float xiW[1]; rng.Next<1>(xiW);
        SampleWavelengths(d.spec.mode, xiW[0], w, p);
        d.waves[i] = make_float4(w[0], w[1], w[2], w[3]);
        d.wavePdf[i] = make_float4(p[0], p[1], p[2], p[3]);
    }
    if(!lowDisc) d.rng[i] = rng.state;
    else { d.rng[i] = rng.state; d.sampleState[i] = make_uint2(rng.sampleIndex, rng.dim); }   // rng[i] = the pixel's seed for the bounces
}

__device__ __forceinline__ Float3 ApplyP(const float* m, Float3 p)
{
    return F3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
              m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
// normals move with the inverse transpose (TransformContextSingle::ApplyN)
__device__ __forceinline__ Float3 ApplyN(const float* inv, Float3 n)
{
    return F3(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z,
              inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
              inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
}

__device__ __forceinline__ void LoadTriangle(const RenderInstance& in, uint32_t prim, Float3 p[3], uint32_t vi[3])
{
    // positions come from the renderer's gathered copy (the same floats); the vertex indices are only needed for per-vertex attributes
    const float4* tp = in.triPos + 3 * size_t(prim);
    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);   // (nine scalar loads of the same words: shade 0.142 -> 0.151 ms)
    vi[0] = vi[1] = vi[2] = 0u;   // (per-vertex attributes come from the gathered triNrm / triUV copies too)
    p[0] = F3(a.x, a.y, a.z); p[1] = F3(b.x, b.y, b.z); p[2] = F3(c.x, c.y, c.z);
}

template<class T>
__global__ void __launch_bounds__(256) KGatherTriAttribute(const T* __restrict__ perVertex, const uint32_t* __restrict__ indices, uint32_t tris, T* __restrict__ out)
{
    for(uint32_t j = blockIdx.x * 256u + threadIdx.x; j < 3u * tris; j += gridDim.x * 256u) out[j] = perVertex[indices[j]];
}

// renderer creation: triPos of one accelerator's primitive group
__global__ void __launch_bounds__(256) KGatherTriPositions(const float* __restrict__ positions, const uint32_t* __restrict__ indices, uint32_t tris, float4* __restrict__ out)
{
    for(uint32_t t = blockIdx.x * 256u + threadIdx.x; t < tris; t += gridDim.x * 256u)
    {
        #pragma unroll
        for(int k = 0; k < 3; k++)
        {
            const uint32_t v = indices[3 * size_t(t) + k];
            out[3 * size_t(t) + k] = make_float4(positions[3 * size_t(v)], positions[3 * size_t(v) + 1], positions[3 * size_t(v) + 2], 0.0f);
        }
    }
}

__global__ void __launch_bounds__(RTPB) KReload(RenderData d)
{
    const uint32_t i = blockIdx.x * RTPB + threadIdx.x;
    const bool inRange = i < d.slots;
    const uint32_t pd = inRange ? d.meta[i].x : PackPD(0, ST_ALIVE, 0);
    ReloadSlot(d, i, inRange, ((pd >> 8) & 0xFFu) == ST_INVALID);
}

// TextureViewCPU<2, Vector3> restated (Device/CPU/TextureViewCPU.h): ResolveEdge L196-246 (C++ truncating / and %),
// ReadPixel + Convert L120-170,L305-340, FindInterpolants L262-300 (a negative texel keeps |frac|, as the reference
// does), ReadInterpolatedPixel L352-385 with Math::Lerp unfused, NearestPixel L248-260.
__device__ __forceinline__ int TexResolveEdge(int i, int n, uint32_t edge)
{
    if(edge == 1u) return min(max(i, 0), n - 1);
    if(edge == 2u)
    {
        const int dim = i / n;
        i = i % n;
        if(i < 0) i += n;
        if((dim & 1) == 1) i = n - i;
        return min(i, n - 1);   // the reference's mirror can produce n (out of bounds there)
    }
    i = i % n;
    if(i < 0) i += n;
    return i;
}
__device__ __forceinline__ Float3 TexReadPixel(const TexRec& t, size_t levelStart, uint32_t levelW, int x, int y)
{
    const size_t o = (levelStart + size_t(y) * levelW + size_t(x)) * t.channels;
    if(t.format == 0u)
    {
        const float* f = static_cast<const float*>(t.data) + o;
        return F3(__ldg(f), __ldg(f + 1), __ldg(f + 2));
    }
    const uint8_t* b = static_cast<const uint8_t*>(t.data) + o;
    const float DELTA = 1.0f / 255.0f;
    return F3(__fmul_rn(float(__ldg(b)), DELTA), __fmul_rn(float(__ldg(b + 1)), DELTA), __fmul_rn(float(__ldg(b + 2)), DELTA));
}
__device__ __forceinline__ Float3 TexLerp(Float3 a, Float3 b, float t)
{ return F3(SpecLerp(a.x, b.x, t), SpecLerp(a.y, b.y, t), SpecLerp(a.z, b.z, t)); }
__device__ __forceinline__ Float3 SampleTextureLevel(const TexRec& t, uint32_t level, float u, float v)
{
    const uint32_t w = MipDim(t.w, level), h = MipDim(t.h, level);
    const size_t start = level ? MipStart(t.w, t.h, level) : 0;
    const float tu = __fmul_rn(u, float(w)), tv = __fmul_rn(v, float(h));
    if(t.interp == 0u)
    {
        const int x = int(roundf(tu - 0.5f)), y = int(roundf(tv - 0.5f));
        return TexReadPixel(t, start, w, TexResolveEdge(x, int(w), t.edge), TexResolveEdge(y, int(h), t.edge));
    }
    float bx, by;
    float fx = modff(tu - 0.5f, &bx), fy = modff(tv - 0.5f, &by);
    int x0 = int(bx), y0 = int(by);
    if(fx < 0.0f) { x0 -= 1; fx = fabsf(fx); }
    if(fy < 0.0f) { y0 -= 1; fy = fabsf(fy); }
    const int xa = TexResolveEdge(x0, int(w), t.edge), xb = TexResolveEdge(x0 + 1, int(w), t.edge);
    const int ya = TexResolveEdge(y0, int(h), t.edge), yb = TexResolveEdge(y0 + 1, int(h), t.edge);
    const Float3 p0 = TexLerp(TexReadPixel(t, start, w, xa, ya), TexReadPixel(t, start, w, xb, ya), fx);
    const Float3 p1 = TexLerp(TexReadPixel(t, start, w, xa, yb), TexReadPixel(t, start, w, xb, yb), fx);
    return TexLerp(p0, p1, fy);
}
__device__ __forceinline__ Float3 SampleTexture(const TexRec& t, float u, float v) { return SampleTextureLevel(t, 0u, u, v); }
// TextureViewCPU::operator()(uv, mipLevel) (Device/CPU/TextureViewCPU.h:L422-470): the level clamped to [0, mipCount - 1] and split
// by ModFInt; LINEAR = Math::Lerp of the two levels' bilinear reads, NEAREST = the nearer level's nearest texel (the reference
// resolves the edge of that read against the BASE size, L446 — an out-of-range index; here against the level's own size).
__device__ __noinline__ Float3 SampleTextureLod(const TexRec& t, float u, float v, float mipLevel)
{
    const uint32_t mc = t.mipCount ? t.mipCount : 1u;
    mipLevel = fminf(fmaxf(mipLevel, 0.0f), float(mc - 1u));   // fmaxf(NaN, 0) = 0
    float ip; const float frac = modff(mipLevel, &ip);
    const uint32_t m0 = uint32_t(ip), m1 = min(m0 + 1u, mc - 1u);
    if(t.interp == 0u) return SampleTextureLevel(t, frac < 0.5f ? m0 : m1, u, v);
    const Float3 a = SampleTextureLevel(t, m0, u, v);
    if(m0 == m1) return a;
    return TexLerp(a, SampleTextureLevel(t, m1, u, v), frac);
}
// TextureViewCPU::operator()(uv, dpdx, dpdy) (L405-420): level = 0.5 log2(max |gradient|^2). lodMode 0 keeps the gradients in UV
// units, as the reference's host backend reads a normalised-coordinate texture; 1 scales them by the base size first, which is
// what tex2DGrad does on the reference's device backends.
__device__ __forceinline__ Float3 SampleTextureGrad(const TexRec& t, float u, float v, float2 dpdx, float2 dpdy, uint32_t lodMode)
{
    if(t.mipCount <= 1u) return SampleTextureLevel(t, 0u, u, v);
    if(lodMode == 1u) { dpdx.x *= float(t.w); dpdx.y *= float(t.h); dpdy.x *= float(t.w); dpdy.y *= float(t.h); }
    const float la = __fmaf_rn(dpdx.y, dpdx.y, __fmul_rn(dpdx.x, dpdx.x)), lb = __fmaf_rn(dpdy.y, dpdy.y, __fmul_rn(dpdy.x, dpdy.x));   // Math::LengthSqr = Dot: an FMA chain
    return SampleTextureLod(t, u, v, 0.5f * log2f(fmaxf(la, lb)));
}

// Material / light colour at the path's wavelengths: Converter::ConvertAlbedo / ConvertRadiance with the
// coefficient fetch hoisted to StartRender (constant attributes), or the RGB pass-through converter.
__device__ __forceinline__ Spec AlbedoAt(const RenderData& d, float4 a, float4 w)
{
    if(!d.spectral) return S4(a.x, a.y, a.z, 0.0f);
    const float3 c = make_float3(a.x, a.y, a.z);
    return S4(EvalSpectrum(c, w.x), EvalSpectrum(c, w.y), EvalSpectrum(c, w.z), EvalSpectrum(c, w.w));
}
__device__ __forceinline__ Spec RadianceAt(const RenderData& d, float4 r, float4 w)
{
    if(!d.spectral) return S4(r.x, r.y, r.z, 0.0f);
    return S4(EvalRadiance(d.spec, r, w.x), EvalRadiance(d.spec, r, w.y), EvalRadiance(d.spec, r, w.z), EvalRadiance(d.spec, r, w.w));
}

// ---- (Mt)Refract and (Mt)Unreal building blocks (Tracer/DistributionFunctions.h, Core/GraphicsFunctions.h) ----
enum : uint32_t { MAT_LAMBERT = 0u, MAT_REFLECT = 1u, MAT_REFRACT = 2u, MAT_UNREAL = 3u };   // mrb_material_type
constexpr float SPECULAR_THRESHOLD = 0.95f;   // MaterialCommon::SpecularThreshold (Tracer/MaterialC.h:L11)

__device__ __forceinline__ float SqrtMax(float x) { return sqrtf(fmaxf(x, 0.0f)); }
// Common::PDFCosDirection (DistributionFunctions.h:L882-889): cos / pi, zero at or below MathConstants::Epsilon
__device__ __forceinline__ float PdfCosDirection(float cosTheta) { const float p = cosTheta * INV_PI_F; return (p <= 1.0e-5f) ? 0.0f : p; }
// BxDF::FresnelDielectric (DistributionFunctions.h:L382-403)
__device__ __forceinline__ float FresnelDielectric(float cosFront, float etaFront, float etaBack)
{
    const float sinFront = SqrtMax(1.0f - cosFront * cosFront);
    const float sinBack = etaFront / etaBack * sinFront;
    if(sinFront >= 1.0f) return 1.0f;
    const float cosBack = SqrtMax(1.0f - sinBack * sinBack);
    float parallel = (etaBack * cosFront - etaFront * cosBack) / (etaBack * cosFront + etaFront * cosBack);
    parallel = parallel * parallel;
    float perpendicular = (etaFront * cosFront - etaBack * cosBack) / (etaFront * cosFront + etaBack * cosBack);
    perpendicular = perpendicular * perpendicular;
    return (parallel + perpendicular) * 0.5f;
}
// Graphics::Reflect / Refract (GraphicsFunctions.h:L157-203): v points away from the surface, on the normal's side
__device__ __forceinline__ Float3 ReflectAbout(Float3 n, Float3 v) { return n * (2.0f * Dot(v, n)) - v; }
__device__ __forceinline__ bool RefractThrough(Float3 n, Float3 v, float etaFrom, float etaTo, Float3& out)
{
    const float etaRatio = etaFrom / etaTo;
    const float cosIn = Dot(n, v);
    const float sinInSqr = fmaxf(0.0f, 1.0f - cosIn * cosIn);
    const float sinOutSqr = etaRatio * etaRatio * sinInSqr;
    const float cosOut = SqrtMax(1.0f - sinOutSqr);
    if(sinOutSqr >= 1.0f) return false;
    out = v * -etaRatio + n * (etaRatio * cosIn - cosOut);
    return true;
}
// Medium::WavelengthToIoRCauchy (DistributionFunctions.h:L978-986)
__device__ __forceinline__ float CauchyIoR(float wavelengthNm, float4 c)
{
    const float w = wavelengthNm * 1.0e-3f, w2 = w * w, w4 = w2 * w2;
    return c.x + c.y / w2 + c.z / w4;
}
// GGX microfacet terms in tangent space, Z = shading normal (DistributionFunctions.h:L431-590)
__device__ __forceinline__ float DGGX(float NdH, float alpha)
{
    const float a2 = alpha * alpha;
    float denom = NdH * NdH * (a2 - 1.0f) + 1.0f;
    denom = denom * denom * PI_F;
    return a2 / denom;
}
__device__ __forceinline__ float LambdaSmith(Float3 v, float alpha)
{
    const float inner = alpha * alpha * (v.x * v.x + v.y * v.y) / (v.z * v.z);
    return (sqrtf(1.0f + inner) - 1.0f) * 0.5f;
}
__device__ __forceinline__ float GSmithSingle(Float3 v, float alpha) { return 1.0f / (1.0f + LambdaSmith(v, alpha)); }
__device__ __forceinline__ float GSmithCorrelated(Float3 wO, Float3 wI, float alpha) { return 1.0f / (LambdaSmith(wO, alpha) + LambdaSmith(wI, alpha) + 1.0f); }
__device__ __forceinline__ float VNDFGGXSmithPDF(Float3 V, Float3 H, float alpha)
{
    const float VdH = fmaxf(0.0f, Dot(H, V)), NdH = fmaxf(0.0f, H.z), NdV = fmaxf(0.0f, V.z);
    const float D = DGGX(NdH, alpha), G1 = GSmithSingle(V, alpha);
    if(NdV == 0.0f) return 0.0f;
    return VdH * D * G1 / NdV;
}
__device__ __forceinline__ Float3 VNDFGGXSmithSample(Float3 V, float alpha, float xi0, float xi1, float& pdf)
{   // Heitz 2018, "Sampling the GGX Distribution of Visible Normals" (DistributionFunctions.h:L528-575)
    const Float3 VHemi = Normalize(F3(alpha * V.x, alpha * V.y, V.z));
    const float len2 = VHemi.x * VHemi.x + VHemi.y * VHemi.y;
    const Float3 T1 = (len2 > 0.0f) ? F3(-VHemi.y, VHemi.x, 0.0f) * rsqrtf(len2) : F3(1.f, 0.f, 0.f);
    const Float3 T2 = Cross(VHemi, T1);
    const float r = sqrtf(xi0), phi = 2.0f * PI_F * xi1;
    float sinPhi, cosPhi; sincosf(phi, &sinPhi, &cosPhi);
    const float t1 = r * cosPhi;
    float t2 = r * sinPhi;
    const float s = 0.5f * (1.0f + VHemi.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    const float val = 1.0f - t1 * t1 - t2 * t2;
    const Float3 NHemi = T1 * t1 + T2 * t2 + VHemi * SqrtMax(val);
    Float3 N = F3(alpha * NHemi.x, alpha * NHemi.y, SqrtMax(NHemi.z));
    const float nLen2 = Dot(N, N);
    if(nLen2 < 1.0e-5f) N = F3(0.f, 0.f, 1.f); else N = N * rsqrtf(nLen2);   // MathConstants::Epsilon
    pdf = VNDFGGXSmithPDF(V, N, alpha);
    return N;
}
__device__ __forceinline__ float BurleyDiffuseCorrection(float NdL, float NdV, float LdH, float roughness)
{
    const float Fd90 = 0.5f + 2.0f * roughness * LdH * LdH;
    auto F = [Fd90](float dot) { const float pw = 1.0f - dot, pw2 = pw * pw; return 1.0f + (Fd90 - 1.0f) * (pw2 * pw2 * pw); };
    return F(NdL) * F(NdV);
}
// UnrealMaterial (MaterialsDefault.hpp:L466-760) with constant roughness / specular / metallic; everything in tangent space
struct UnrealBxDF
{
    Spec  albedo;      // at the path's wavelengths (or rgb, 0)
    float roughness, specular, metallic;
    __device__ __forceinline__ float AvgAlbedo() const { return (albedo.x + albedo.y + albedo.z + albedo.w) * 0.3333f; }
    __device__ __forceinline__ float MISRatio() const
    {   // probability of the diffuse lobe
        const float avg = AvgAlbedo();
        const float integralDiffuse = 2.0f * PI_F * avg * (1.0f - metallic);
        const float specularRatio = specular * (1.0f - metallic) + avg * metallic;
        const float total = specularRatio + integralDiffuse;
        return (total == 0.0f) ? 0.0f : integralDiffuse / total;
    }
    __device__ __forceinline__ float Specularity() const { return 1.0f - MISRatio(); }
    __device__ __forceinline__ Spec F0() const
    {
        const float specOut = specular * 0.08f, om = 1.0f - metallic;
        return S4(specOut * om + albedo.x * metallic, specOut * om + albedo.y * metallic, specOut * om + albedo.z * metallic, specOut * om + albedo.w * metallic);
    }
    __device__ __forceinline__ Spec FSchlick(float VdH) const
    {
        const Spec f0 = F0();
        const float pw = 1.0f - VdH, pw2 = pw * pw, pw5 = pw2 * pw2 * pw;
        return S4((1.0f - f0.x) * pw5 + f0.x, (1.0f - f0.y) * pw5 + f0.y, (1.0f - f0.z) * pw5 + f0.z, (1.0f - f0.w) * pw5 + f0.w);
    }
    __device__ __forceinline__ Spec Diffuse(float NdL, float NdV, float LdH) const
    {
        const float k = NdL * INV_PI_F * (1.0f - metallic) * BurleyDiffuseCorrection(NdL, NdV, LdH, roughness);
        return albedo * k;
    }
    // Evaluate (L760+): reflectance of tangent-space directions V (out) and L (in)
    __device__ __forceinline__ Spec Evaluate(Float3 V, Float3 L) const
    {
        const float alpha = roughness * roughness;
        const Float3 H = Normalize(L + V);
        const float LdH = fmaxf(0.f, Dot(L, H)), VdH = fmaxf(0.f, Dot(V, H)), NdH = fmaxf(0.f, H.z), NdV = fmaxf(0.f, V.z);
        float D = DGGX(NdH, alpha);
        D = (isnan(D) || isinf(D)) ? 0.0f : D;
        float G = GSmithCorrelated(V, L, alpha);
        G = (LdH == 0.0f) ? 0.0f : G;
        G = (VdH == 0.0f) ? 0.0f : G;
        Spec specularTerm = FSchlick(VdH) * (D * G * 0.25f / NdV);
        if(NdV == 0.0f) specularTerm = S4(0.f, 0.f, 0.f, 0.f);
        const float NdL = fmaxf(0.f, L.z);
        return Diffuse(NdL, NdV, LdH) + specularTerm;
    }
    __device__ __forceinline__ float Pdf(Float3 V, Float3 L) const
    {
        const float alpha = roughness * roughness;
        const Float3 H = Normalize(L + V);
        const float mis = MISRatio();
        const float pdfD = PdfCosDirection(L.z);
        const float NdH = fmaxf(0.f, H.z), VdH = fmaxf(0.f, Dot(V, H));
        const float D = DGGX(NdH, alpha);
        float pdfS = VNDFGGXSmithPDF(V, H, alpha);
        pdfS = (isnan(D) || isinf(D)) ? 0.0f : pdfS;
        pdfS = (VdH == 0.0f) ? 0.0f : pdfS / (4.0f * VdH);
        return pdfD * mis + pdfS * (1.0f - mis);
    }
    // SampleBxDF (L536-640): sXi picks the lobe, (xi0, xi1) samples it; returns L, reflectance and the mixture pdf
    __device__ __forceinline__ Float3 Sample(Float3 V, float sXi, float xi0, float xi1, Spec& reflectance, float& pdfOut) const
    {
        const float alpha = roughness * roughness;
        const float mis = MISRatio();
        Float3 L, H; float pdfD, pdfS;
        if(sXi < mis)
        {   // Common::SampleCosDirection
            const float phi = 2.0f * PI_F * xi1, su = sqrtf(xi0);
            float sn, cs; sincosf(phi, &sn, &cs);
            const float lx = su * cs, ly = su * sn;
            L = F3(lx, ly, sqrtf(fmaxf(0.0f, 1.0f - (lx * lx + ly * ly))));
            pdfD = L.z * INV_PI_F;                               // the sample's own pdf (SampleCosDirection)
            H = Normalize(L + V);
            const float VdH = fmaxf(0.f, Dot(V, H));
            pdfS = VNDFGGXSmithPDF(V, H, alpha);
            pdfS = (VdH == 0.0f) ? 0.0f : pdfS / (4.0f * VdH);
        }
        else
        {
            float pdfH;
            H = VNDFGGXSmithSample(V, alpha, xi0, xi1, pdfH);
            L = ReflectAbout(H, V);
            const float VdH = fmaxf(0.f, Dot(V, H));
            pdfS = (VdH == 0.0f) ? 0.0f : pdfH / (4.0f * VdH);
            pdfD = PdfCosDirection(L.z);
        }
        const float VdH = fmaxf(0.f, Dot(V, H)), LdH = fmaxf(0.f, Dot(L, H)), NdH = fmaxf(0.f, H.z), NdV = fmaxf(0.f, V.z);
        const float D = DGGX(NdH, alpha);
        float G = GSmithCorrelated(V, L, alpha);
        G = (LdH == 0.0f) ? 0.0f : G;
        G = (VdH == 0.0f) ? 0.0f : G;
        const Spec F = FSchlick(VdH);
        Spec specularTerm = F * (D * G * 0.25f / NdV);
        if(NdV == 0.0f) specularTerm = S4(0.f, 0.f, 0.f, 0.f);
        if(isinf(D) || isnan(D))
        {   // alpha ~ 0: the cancelled form
            specularTerm = F * (G / GSmithSingle(V, alpha));
            pdfS = 1.0f;
        }
        const float NdL = fmaxf(0.f, L.z);
        reflectance = Diffuse(NdL, NdV, LdH) + specularTerm;
        pdfOut = pdfD * mis + pdfS * (1.0f - mis);
        return L;
    }
};

// Shading frame of a triangle hit as the reference builds it (Triangle::GenerateSurface, PrimitiveDefaultTriangle.hpp:L463-470):
// the three vertex quaternions (world -> tangent space rotations, (w, x, y, z)) are blended with Quaternion::BarySLerp
// (Core/Quaternion.hpp:L289-353: slerp(q1, q0, a / (a + b)) then slerp(., q2, 1 - a - b), shortest arc, lerp when nearly
// parallel), normalised, and the shading normal is the frame's Z axis (OrthoBasisZ, L256-270).
__device__ __forceinline__ float4 QuatSLerp(float4 s, float4 e, float t)
{
    const float cosTheta = s.x * e.x + s.y * e.y + s.z * e.z + s.w * e.w;
    const float cosFlipped = (cosTheta >= 0.0f) ? cosTheta : -cosTheta;
    float s0, s1;
    if(cosFlipped < (1.0f - 1.0e-5f))
    {
        const float angle = acosf(cosFlipped), sinRecip = 1.0f / sinf(angle);
        s0 = sinf(angle * (1.0f - t)) * sinRecip;
        s1 = sinf(angle * t) * sinRecip;
    }
    else { s0 = 1.0f - t; s1 = t; }
    s1 = (cosTheta >= 0.0f) ? s1 : -s1;
    return make_float4(s.x * s0 + e.x * s1, s.y * s0 + e.y * s1, s.z * s0 + e.z * s1, s.w * s0 + e.w * s1);
}
__device__ __forceinline__ Float3 ShadingNormalFromTBN(float4 q0, float4 q1, float4 q2, float a, float b)
{
    const float c = 1.0f - a - b;
    float4 q;
    if(fabsf(a + b) < 1.0e-5f) q = q2;
    else q = QuatSLerp(QuatSLerp(q1, q0, a / (a + b)), q2, c);
    const float inv = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float w = q.x * inv, x = q.y * inv, y = q.z * inv, z = q.w * inv;       // stored (w, x, y, z)
    return F3(2.0f * (x * z - w * y), 2.0f * (y * z + w * x), w * w - x * x - y * y + z * z);
}

// The interpolated world -> tangent rotation of a hit as the three rows of its matrix (row 2 = the shading normal above);
// a tangent-space vector v goes back to the world as v.x row0 + v.y row1 + v.z row2 (Quaternion::ApplyInvRotation).
__device__ __forceinline__ void TBNRows(float4 q0, float4 q1, float4 q2, float a, float b, Float3 rows[3])
{
    const float c = 1.0f - a - b;
    float4 q;
    if(fabsf(a + b) < 1.0e-5f) q = q2;
    else q = QuatSLerp(QuatSLerp(q1, q0, a / (a + b)), q2, c);
    const float inv = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float w = q.x * inv, x = q.y * inv, y = q.z * inv, z = q.w * inv;
    rows[0] = F3(w * w + x * x - y * y - z * z, 2.0f * (x * y - w * z), 2.0f * (x * z + w * y));
    rows[1] = F3(2.0f * (x * y + w * z), w * w - x * x + y * y - z * z, 2.0f * (y * z - w * x));
    rows[2] = F3(2.0f * (x * z - w * y), 2.0f * (y * z + w * x), w * w - x * x - y * y + z * z);
}

// ---- ray cones (Tracer/TracerTypes.h:L48-69,L323-383; RT Gems I ch. 20, Akenine-Moller et al. JCGT 10(1)): the footprint a path
// carries so that textured reads can pick a mip level. x = aperture, y = width. ----
struct ConeSurf { float2 front, back; float betaN; };
__device__ __forceinline__ float2 ConeAdvance(float2 c, float t)
{   // RayCone::Advance: width + aperture * t, clamped to [Epsilon, 1e6]
    return make_float2(c.x, fminf(fmaxf(c.y + c.x * t, 1.0e-5f), 1.0e6f));
}
// The ray-cone half of Triangle::GenerateSurface (PrimitiveDefaultTriangle.hpp:L495-569): RayCone::Project gives the two axes of the
// footprint ellipse on the hit plane, the vertex normals along the triangle's edges a curvature estimate (JCGT eq. 6) that will
// widen or narrow the cone at the bounce (betaN), and the UVs one axis away from the hit the texture gradients (listing 1).
// gN = geometric normal flipped towards the ray, dDotN = dot(unflipped normal, dir), dirN = normalised ray direction.
__device__ __noinline__ void ConeSurface(const Float3* p, const Float3* vn, bool hasNormals, float2 t0, float2 t1, float2 t2, float a, float b, Float3 pos,
                                         Float3 gN, float dDotN, Float3 dirN, float2 cone, ConeSurf& out, float2& dpdx, float2& dpdy)
{
    const float EPS = 1.0e-5f;
    Float3 dd = dirN;
    const float fd = Dot(gN, dd);
    if(fabsf(fd + 1.0f) < EPS) dd = dd + F3(EPS, EPS, EPS);
    const Float3 h1 = dd - gN * fd, h2 = Cross(gN, h1);
    const float r = cone.y * 0.5f;
    const Float3 a1 = h1 * (r / fmaxf(EPS, Length(h1 - dd * Dot(dd, h1)))), a2 = h2 * (r / fmaxf(EPS, Length(h2 - dd * Dot(dd, h2))));
    const Float3 r0 = Normalize(a1), r1 = Normalize(a2);
    const Float3 e[3] = {p[1] - p[0], p[2] - p[0], p[2] - p[1]};
    float k[3] = {0.f, 0.f, 0.f};
    if(hasNormals)
    {
        k[0] = Dot(vn[1] - vn[0], e[0]) / Dot(e[0], e[0]);
        k[1] = Dot(vn[2] - vn[0], e[1]) / Dot(e[1], e[1]);
        k[2] = Dot(vn[2] - vn[1], e[2]) / Dot(e[2], e[2]);
    }
    // Vector3::Minimum / Maximum: index of the first smallest / largest
    const uint32_t mn = (k[1] < k[0]) ? ((k[2] < k[1]) ? 2u : 1u) : ((k[2] < k[0]) ? 2u : 0u);
    const uint32_t mx = (k[1] > k[0]) ? ((k[2] > k[1]) ? 2u : 1u) : ((k[2] > k[0]) ? 2u : 0u);
    const Float3 eMn = (mn == 0u) ? e[0] : (mn == 1u ? e[1] : e[2]), eMx = (mx == 0u) ? e[0] : (mx == 1u ? e[1] : e[2]);
    const float kMn = (mn == 0u) ? k[0] : (mn == 1u ? k[1] : k[2]), kMx = (mx == 0u) ? k[0] : (mx == 1u ? k[1] : k[2]);
    const float a1L = Length(a1), a2L = Length(a2), a1S = a1L * a1L, a2S = a2L * a2L, a12 = a1L * a2L;
    const float mnx = Dot(r0, eMn), mny = Dot(r1, eMn), mxx = Dot(r0, eMx), mxy = Dot(r1, eMx);
    const float l0 = a12 * (1.0f / sqrtf(a1S * mnx * mnx + a2S * mny * mny));
    const float l1 = a12 * (1.0f / sqrtf(a1S * mxx * mxx + a2S * mxy * mxy));
    const float lMaxRecip = 1.0f / fmaxf(l0, l1);
    const float beta0 = -1.0f * (kMn * l0 * lMaxRecip) * fabsf(cone.y) / dDotN, beta1 = -1.0f * (kMx * l1 * lMaxRecip) * fabsf(cone.y) / dDotN;
    out.front = cone; out.back = cone;
    out.betaN = (fabsf(cone.x + beta0) >= fabsf(cone.x + beta1)) ? beta0 : beta1;
    const float areaRecip = 1.0f / Dot(gN, Cross(e[0], e[1]));
    const float c = 1.0f - a - b;
    const float u = t0.x * a + t1.x * b + t2.x * c, v = t0.y * a + t1.y * b + t2.y * c;
    #pragma unroll
    for(int i = 0; i < 2; i++)
    {
        const Float3 eP = pos - p[0] + (i == 0 ? a1 : a2);
        const float ba = Dot(gN, Cross(eP, e[1]) * areaRecip), bb = Dot(gN, Cross(e[0], eP) * areaRecip), bc = 1.0f - ba - bb;
        const float2 g = make_float2((t0.x * bc + t1.x * ba + t2.x * bb) - u, (t0.y * bc + t1.y * ba + t2.y * bb) - v);
        if(i == 0) dpdx = g; else dpdy = g;
    }
}
// RayConeSurface::ConeAfterScatter: a reflected cone widens by twice the curvature term, a transmitted one takes the back cone
__device__ __forceinline__ float2 ConeAfterScatter(const ConeSurf& cs, Float3 wI, Float3 n)
{
    return (Dot(wI, n) > 0.0f) ? make_float2(cs.front.x + 2.0f * cs.betaN, cs.front.y) : make_float2(cs.back.x - cs.betaN, cs.back.y);
}
__device__ __forceinline__ float2 Refract2D(float2 v, float2 n, float fromEta, float toEta)
{   // Graphics::Refract(n, -v) in the plane of incidence; under total internal reflection the tangential direction
    const float er = fromEta / toEta, cosIn = -(n.x * v.x + n.y * v.y);
    const float sinOut2 = er * er * fmaxf(0.0f, 1.0f - cosIn * cosIn);
    if(sinOut2 >= 1.0f)
    {
        const float nd = n.x * v.x + n.y * v.y;
        const float tx = v.x - n.x * nd, ty = v.y - n.y * nd, l = 1.0f / sqrtf(tx * tx + ty * ty);
        return make_float2(tx * l, ty * l);
    }
    const float k = er * cosIn - sqrtf(fmaxf(0.0f, 1.0f - sinOut2));
    return make_float2(er * v.x + k * n.x, er * v.y + k * n.y);
}
// RefractMaterial::RefractRayCone (MaterialsDefault.hpp:L355-462, after RT Gems II ch. 10 / Falcor): the upper and lower edge rays
// of the cone refracted in the plane of incidence through normals tilted by the curvature term give the back cone's aperture and
// width. fromEta / toEta already swapped for a back-side hit; gN flipped towards wO.
__device__ __noinline__ void RefractRayCone(ConeSurf& cs, Float3 wO, Float3 gN, float fromEta, float toEta)
{
    const float er = fromEta / toEta, cosIn3 = Dot(gN, wO);
    if(er * er * fmaxf(0.0f, 1.0f - cosIn3 * cosIn3) >= 1.0f) return;
    const Float3 d3 = wO * -1.0f;
    const Float3 x = Normalize(d3 - gN * Dot(d3, gN));
    const float2 d = make_float2(Dot(x, d3), Dot(gN, d3));
    const float aperture = cs.front.x, width = cs.front.y;
    float sn, co; sincosf(((width > 0.0f) ? 1.0f : 0.0f) * aperture * 0.5f, &sn, &co);
    const float2 du = make_float2(d.x * co - d.y * sn, d.x * sn + d.y * co), dl = make_float2(d.x * co + d.y * sn, d.x * -sn + d.y * co);
    const float2 od = make_float2(-d.y * width * 0.5f, d.x * width * 0.5f);
    const float uHitX = +od.x + du.x * (-od.y / du.y), lHitX = -od.x + dl.x * (+od.y / dl.y);
    const float nSign = (uHitX > lHitX) ? 1.0f : -1.0f;
    sincosf(-cs.betaN * nSign * 0.5f, &sn, &co);
    const float2 nu = make_float2(-sn, co), nl = make_float2(sn, co);   // Rotate2D_UL((0, 1), alpha)
    const float2 tu = Refract2D(du, nu, fromEta, toEta), tl = Refract2D(dl, nl, fromEta, toEta);
    const float2 o1 = make_float2(-d.y, d.x);
    const float wl = (-uHitX * tu.y) / (o1.x * -tu.y + o1.y * tu.x), wu = (+lHitX * tl.y) / (o1.x * -tl.y + o1.y * tl.x);
    const float sign = copysignf(1.0f, tu.x * tl.y - tu.y * tl.x);
    const float ap = fmaxf(acosf(fminf(fmaxf(tu.x * tl.x + tu.y * tl.y, -1.0f), 1.0f)) * sign, 1.0e-5f);
    cs.back = make_float2(ap + cs.betaN, wu + wl);
}

// ---- LightSkysphere (Tracer/LightsDefault.hpp:L310-443): the boundary light as an environment sphere ----
// TransformContext::ApplyV / InvApplyV of the light surface's transform (the linear part only: directions)
__device__ __forceinline__ Float3 SkyApply(const float* m, Float3 v)
{
    return F3(Dot(F3(m[0], m[1], m[2]), v), Dot(F3(m[3], m[4], m[5]), v), Dot(F3(m[6], m[7], m[8]), v));
}
// radiance(uv) converted to the path's wavelengths (ParamVaryingData<2, Vector3> behind the SpectrumConverter)
__device__ __forceinline__ Spec SkyRadiance(const RenderData& d, float2 uv, float4 waves)
{
    if(d.boundaryTex < 0) return RadianceAt(d, d.boundaryRadiance, waves);
    const Float3 rgb = SampleTexture(d.textures[d.boundaryTex], uv.x, uv.y);
    if(d.spectral) return RadianceAt(d, FetchRadianceCoeffs(d.spec, rgb.x, rgb.y, rgb.z), waves);
    return S4(rgb.x, rgb.y, rgb.z, 0.0f);
}
// EmitViaHit / EmitViaSurfacePoint (L408-443): the direction alone decides
__device__ __forceinline__ Spec SkyEmit(const RenderData& d, Float3 wO, float4 waves)
{
    Float3 dir = wO * -1.0f;
    if(!d.boundaryIdentity) dir = SkyApply(d.boundaryInvM, dir);
    dir = Normalize(dir);
    const float2 uv = (d.boundaryTex < 0) ? make_float2(0.f, 0.f) : SkyDirToUV(d.boundaryType, dir.x, dir.y, dir.z);
    return SkyRadiance(d, uv, waves);
}
// PdfSolidAngle (L359-370)
__device__ __forceinline__ float SkyPdfSolidAngle(const RenderData& d, Float3 dirWorld)
{
    const Float3 dirYUp = d.boundaryIdentity ? dirWorld : SkyApply(d.boundaryInvM, dirWorld);
    const Float3 n = Normalize(dirYUp);
    const float2 uv = SkyDirToUV(d.boundaryType, n.x, n.y, n.z);
    const float pdf = (d.boundaryTex < 0) ? 1.0f : DistPdfUV(d.boundaryDist, uv.x, uv.y);
    return SkyPdfFromDir(d.boundaryType, pdf, dirYUp.y);
}

// LightPrim::EmitViaHit / EmitViaSurfacePoint for a constant radiance (LightsDefault.hpp:L129-168)
__device__ __forceinline__ Spec Emit(const RenderData& d, const EmissiveTri& l, Float3 n, Float3 wO, float4 waves)
{
    float NdL = Dot(n, wO);
    if(l.e0.w == 0.0f && NdL <= 0.0f) return S4(0.f, 0.f, 0.f, 0.f);
    return RadianceAt(d, l.radiance, waves);
}

// KCGenerateSurfaceWorkKeysIndirect / KCSetBoundaryWorkKeysIndirect (Tracer/RendererCommon.cu:L23-81):
// 32-bit sort key = [work batch : MSBs][data bits]. Work batches here: 0 = Lambert surface work,
// 1 = prim-light work, 2 = boundary (miss) work, 3 = idle slot; data bits = material / light index, so
// rays of one material are contiguous after the sort.
__global__ void __launch_bounds__(RTPB) KGenWorkKeys(RenderData d)
{
    const uint32_t i = blockIdx.x * RTPB + threadIdx.x;
    if(i >= d.slots) return;
    const uint32_t status = (d.meta[i].x >> 8) & 0xFFu;
    const uint4 keys = *reinterpret_cast<const uint4*>(d.hitKeys + i);
    uint32_t batch, data = 0;
    if(status != ST_ALIVE) batch = 3u;
    else if(keys.x == INVALID_U32) batch = 2u;
    else { batch = (keys.y & 0x80000000u) ? 1u : 0u; data = keys.y & ((1u << d.matBits) - 1u); }
    d.workKeys[i] = (batch << d.matBits) | data;
    d.workIndices[i] = i;
}

// Shading of one slot. Returns true when the slot held a live path (= one closest-hit ray was cast for it
// this bounce); castShadow reports an NEE shadow ray.
// GLOSSY (the full-featured kernel): the scene has (Mt)Refract / (Mt)Unreal materials or a skysphere boundary light (their code
// costs 16-30 registers, so scenes without them run the lean kernel)
template<bool GLOSSY>
__device__ __forceinline__ bool ShadeSlot(const RenderData& d, uint32_t i, bool& castShadow, bool& neeSample)
{
    // every per-slot input is requested before the first use, so one round trip covers them all
    const uint4 meta = d.meta[i];
    const float4 r0 = reinterpret_cast<const float4*>(d.rays + i)[0];
    const float4 r1 = reinterpret_cast<const float4*>(d.rays + i)[1];
    const uint4 keys = *reinterpret_cast<const uint4*>(d.hitKeys + i);
    const float2 bary = *reinterpret_cast<const float2*>(d.hits + i);
    const float4 thr4 = d.throughput[i];
    const uint32_t rngState = d.rng[i];
    const uint32_t pd = meta.x;
    if(((pd >> 8) & 0xFFu) != ST_ALIVE) return false;
    uint32_t depth = pd & 0xFFu, type = (pd >> 16) & 0xFFu;
    const Float3 ro = F3(r0.x, r0.y, r0.z), rd = F3(r1.x, r1.y, r1.z);
    Spec throughput = S4(thr4);
    const float4 waves = d.spectral ? d.waves[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    // every live slot gets a well defined (never hitting) shadow ray unless NEE writes one
    d.shadowRays[i].tMin = 1.0f; d.shadowRays[i].tMax = -1.0f;

    if(keys.x == INVALID_U32)
    {
        // boundary light (LightWorkFunction[WithNEE]::Call for a light that is not primitive backed). (L)Null emits
        // nothing; a skysphere emits along the ray's direction, MIS-weighted against its own solid-angle pdf.
        if(GLOSSY && d.boundaryType != 0u)
        {
            bool count = true;
            if(d.sampleMode == 1u && type != RAY_CAMERA && type != RAY_SPECULAR) count = false;
            if(count)
            {
                if(d.sampleMode == 2u && type == RAY_PATH)
                {
                    const float pdfB = __uint_as_float(meta.w);
                    const float pdfL = SkyPdfSolidAngle(d, rd) * (1.0f / float(d.lightCount + 1u));
                    const float mis = pdfB + pdfL;
                    throughput = throughput * pdfB;
                    throughput = (mis == 0.0f) ? S4(0, 0, 0, 0) : throughput * (1.0f / mis);
                }
                const Spec em = SkyEmit(d, rd * -1.0f, waves);
                if(depth + 1u <= d.rrHi)
                    d.radiance[i] = F4(S4(d.radiance[i]) + em * throughput);
            }
        }
        d.meta[i].x = PackPD(depth, ST_DEAD, type);
        return true;
    }
    const uint32_t prim = keys.x & 0x0FFFFFFFu;
    const uint32_t lmKey = keys.y;
    const RenderInstance& in = d.instances[d.sceneMode ? keys.w : 0u];
    Float3 p[3]; uint32_t vi[3];
    LoadTriangle(in, prim, p, vi);
    const float a = bary.x, b = bary.y, c = 1.0f - a - b;
    Float3 pos = p[0] * a + p[1] * b + p[2] * c;
    const Float3 e0 = p[1] - p[0], e1 = p[2] - p[0];
    Float3 geoN = Cross(e0, e1);
    if(!in.identity) { pos = ApplyP(in.transform, pos); geoN = ApplyN(in.invTransform, geoN); }
    geoN = Normalize(geoN);

    if(lmKey & 0x80000000u)
    {
        // ---------------- light hit: LightWorkFunctionWithNEE / LightWorkFunction ----------------
        const EmissiveTri l = d.lights[in.lightOfPrim[prim]];
        bool count = true;
        if(d.sampleMode == 1u && type != RAY_CAMERA && type != RAY_SPECULAR) count = false;
        if(count)
        {
            if(d.sampleMode == 2u && type == RAY_PATH)
            {
                // MIS: undo the bxdf pdf division, divide by (bxdf pdf + light pdf)  (BalanceCancelled)
                float pdfB = __uint_as_float(meta.w);
                float NdL = Dot(geoN, rd * -1.0f);
                NdL = (l.e0.w != 0.0f) ? fabsf(NdL) : fmaxf(0.0f, NdL);
                float pdfL = (NdL == 0.0f) ? 0.0f : (1.0f / l.p0.w) / NdL;
                Float3 dv = ro - pos;
                pdfL *= Dot(dv, dv);
                pdfL *= 1.0f / float(d.lightCount + 1u);
                float mis = pdfB + pdfL;
                throughput = throughput * pdfB;
                throughput = (mis == 0.0f) ? S4(0, 0, 0, 0) : throughput * (1.0f / mis);
            }
            const Spec em = Emit(d, l, geoN, rd * -1.0f, waves);
            if(depth + 1u <= d.rrHi)
                d.radiance[i] = F4(S4(d.radiance[i]) + em * throughput);
        }
        d.meta[i].x = PackPD(depth, ST_DEAD, type);
        return true;
    }

    // ------------------------------- material surface -------------------------------
    SlotSampler rng = LoadSampler(d.samplerType, rngState, d.samplerType != SAMPLER_INDEPENDENT ? d.sampleState[i] : make_uint2(0u, 0u),
                                  meta.y % d.width + d.regionX, meta.y / d.width + d.regionY, d.sobolMatrices, d.zsobol);
    const bool backSide = Dot(geoN, Normalize(rd)) > 0.0f;
    // ray cone at the hit (KCRenderWork, RenderWork.kt.h:L65-85: the cone arrives advanced by the hit distance); only tracked when a
    // texture has more than one mip level (the full kernel)
    float2 dpdx = make_float2(0.f, 0.f), dpdy = make_float2(0.f, 0.f);
    ConeSurf cs; cs.front = cs.back = make_float2(0.f, 0.f); cs.betaN = 0.0f;
    const bool useCones = GLOSSY && d.cones != nullptr;
    if(GLOSSY && useCones)
    {
        const Float3 dirN = Normalize(rd);
        Float3 pw[3] = {p[0], p[1], p[2]}, vn[3] = {geoN, geoN, geoN};
        if(!in.identity) { pw[0] = ApplyP(in.transform, p[0]); pw[1] = ApplyP(in.transform, p[1]); pw[2] = ApplyP(in.transform, p[2]); }
        if(in.vertexNormals)
        {   // Quaternion::OrthoBasisZ of the vertex frames (left in the primitive's local space, as the reference does)
            #pragma unroll
            for(int k = 0; k < 3; k++)
            {
                const float4 q = in.triNrm[3 * size_t(prim) + k];
                vn[k] = in.tbn ? F3(q.y * q.w - q.x * q.z + q.y * q.w - q.x * q.z, q.z * q.w + q.x * q.y + q.z * q.w + q.x * q.y, q.x * q.x - q.y * q.y - q.z * q.z + q.w * q.w)
                               : F3(q.x, q.y, q.z);
            }
        }
        float2 t0 = make_float2(0.f, 0.f), t1 = t0, t2 = t0;
        if(in.vertexUVs) { t0 = in.triUV[3 * size_t(prim) + 0]; t1 = in.triUV[3 * size_t(prim) + 1]; t2 = in.triUV[3 * size_t(prim) + 2]; }
        ConeSurface(pw, vn, in.vertexNormals != nullptr, t0, t1, t2, a, b, pos, backSide ? geoN * -1.0f : geoN, Dot(geoN, dirN), dirN,
                    ConeAdvance(d.cones[i], r1.w), cs, dpdx, dpdy);
    }
    Float3 shadeN = geoN;
    bool normalMapped = false;
    const uint32_t matIndex = lmKey & 0x1FFFFFu;
    if(in.vertexNormals)
    {
        const float4 n0 = in.triNrm[3 * size_t(prim) + 0], n1 = in.triNrm[3 * size_t(prim) + 1], n2 = in.triNrm[3 * size_t(prim) + 2];
        const int32_t normalTex = (GLOSSY && d.normalTex && in.tbn) ? d.normalTex[matIndex] : -1;
        if(GLOSSY && normalTex >= 0)
        {
            // normal map (Triangle::GenerateSurface, PrimitiveDefaultTriangle.hpp:L478-491,L571-575): the hit's frame — turned half
            // a turn about its tangent axis on a back-side hit — is re-aimed so that its Z axis is the texture's normal:
            // tbn' = RotationBetweenZAxis(n).Conjugate() * tbn, i.e. the shading normal is tbn^-1 (n)
            float2 uv = make_float2(0.f, 0.f);
            if(in.vertexUVs)
            {
                const float2 t0 = in.triUV[3 * size_t(prim) + 0], t1 = in.triUV[3 * size_t(prim) + 1], t2 = in.triUV[3 * size_t(prim) + 2];
                uv = make_float2(t0.x * a + t1.x * b + t2.x * c, t0.y * a + t1.y * b + t2.y * c);
            }
            const Float3 nt = Normalize(SampleTextureGrad(d.textures[normalTex], uv.x, uv.y, dpdx, dpdy, d.textureLodMode));
            Float3 rows[3]; TBNRows(n0, n1, n2, a, b, rows);
            const float sgn = backSide ? -1.0f : 1.0f;
            shadeN = rows[0] * nt.x + (rows[1] * nt.y + rows[2] * nt.z) * sgn;
            normalMapped = true;
        }
        else if(in.tbn) shadeN = ShadingNormalFromTBN(n0, n1, n2, a, b);
        else shadeN = F3(n0.x, n0.y, n0.z) * a + F3(n1.x, n1.y, n1.z) * b + F3(n2.x, n2.y, n2.z) * c;
        // (the reference leaves the tangent frame in the primitive's local space under a (T)Single transform — "we can't
        // apply a transform to tbn", PrimitiveDefaultTriangle.hpp:L612 — which tilts its shading normals by the instance's
        // rotation; here the normal follows the instance like the geometry does)
        if(!in.identity) shadeN = ApplyN(in.invTransform, shadeN);
        shadeN = Normalize(shadeN);
    }
    if(backSide) { geoN = geoN * -1.0f; if(!normalMapped) shadeN = shadeN * -1.0f; }
    const uint32_t matType = d.materialType ? uint32_t(d.materialType[matIndex]) : MAT_LAMBERT;
    const Float3 wO = Normalize(rd) * -1.0f;
    if(matType == MAT_REFLECT || (GLOSSY && matType == MAT_REFRACT))
    {
        // Perfectly specular materials. WorkFunctionNEE draws its light sample but casts no shadow ray for them
        // (PathTracerRendererShaders.h:L415-424); WorkFunction samples the BxDF, skips the roulette and marks the next ray
        // SPECULAR_RAY, so a light it hits counts in full in every sample mode (L245-262).
        if(d.sampleMode != 0u) { float skip[3]; rng.Next<3>(skip); }
        Float3 wIr; float pdfS = 1.0f; bool passedThrough = false;
        Float3 originBase = pos;
        if(!GLOSSY || matType == MAT_REFLECT)
        {   // (Mt)Reflect (MaterialsDefault.hpp:L132-215): wO about the shading normal, reflectance 1, pdf 1
            wIr = Normalize(ReflectAbout(shadeN, wO));
        }
        else
        {   // (Mt)Refract (MaterialsDefault.hpp:L246-312): Fresnel-weighted choice between reflection and refraction
            const float4 cf = d.matParams[2 * matIndex], cb = d.matParams[2 * matIndex + 1];
            float fromEta = d.spectral ? CauchyIoR(waves.x, cf) : cf.x;
            float toEta = d.spectral ? CauchyIoR(waves.x, cb) : cb.x;
            if(backSide) { const float t = fromEta; fromEta = toEta; toEta = t; }
            const float cosTheta = fabsf(Dot(wO, shadeN));
            const float f = FresnelDielectric(cosTheta, fromEta, toEta);
            float xiF[1]; rng.Next<1>(xiF);
            const bool doReflection = xiF[0] < f;
            Float3 refr = wO * -1.0f;
            if(!doReflection) RefractThrough(shadeN, wO, fromEta, toEta, refr);   // f = 1 under total internal reflection: never taken then
            wIr = Normalize(doReflection ? ReflectAbout(shadeN, wO) : refr);
            pdfS = doReflection ? f : (1.0f - f);
            passedThrough = !doReflection;
            // BxDFSample.wI is nudged along the shading normal by the material itself (L296), before the work function
            // nudges it again along the (flipped) geometric normal
            originBase = NudgePos(pos, shadeN);
            // reflectance = Spectrum(pdf): cancels against the division by the pdf below
            throughput = throughput * pdfS;
            if(useCones) RefractRayCone(cs, wO, geoN, fromEta, toEta);
            if(passedThrough && d.spectral)
            {   // DisperseWaves + StoreWaves (L229-238): the path keeps its first wavelength only
                d.waves[i] = make_float4(waves.x, -1.0f, -1.0f, -1.0f);
            }
        }
        d.shadowRadiance[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        depth += 1u;
        if(d.samplerType == SAMPLER_INDEPENDENT) d.rng[i] = rng.state;
        else d.sampleState[i].y = rng.dim;
        if(depth < d.rrHi)
        {
            if(GLOSSY && matType == MAT_REFRACT)
            {
                throughput = (pdfS == 0.0f) ? S4(0, 0, 0, 0) : throughput * (1.0f / pdfS);
                d.throughput[i] = F4(throughput);
            }
            d.meta[i].w = __float_as_uint(pdfS);
            const Float3 no = NudgePos(originBase, passedThrough ? geoN * -1.0f : geoN);
            float4* rp = reinterpret_cast<float4*>(d.rays + i);
            rp[0] = make_float4(no.x, no.y, no.z, 1.0e-4f);
            rp[1] = make_float4(wIr.x, wIr.y, wIr.z, FLT_MAX);
            if(GLOSSY && useCones) d.cones[i] = ConeAfterScatter(cs, wIr, geoN);
            d.meta[i].x = PackPD(depth, ST_ALIVE, RAY_SPECULAR);
        }
        else
        {
            d.rays[i].tMin = 1.0f; d.rays[i].tMax = -1.0f;
            d.meta[i].x = PackPD(depth, ST_DEAD, RAY_SPECULAR);
        }
        return true;
    }
    float4 albedoRaw = d.albedo[matIndex];
    const int32_t texIndex = d.albedoTex ? d.albedoTex[matIndex] : -1;
    if(texIndex >= 0)
    {
        // ParamVaryingData<2, Vector3>: albedo texture at the interpolated UV0 (LambertMaterial ctor, MaterialsDefault.hpp:L17-25)
        float2 uv = make_float2(0.f, 0.f);
        if(in.vertexUVs)
        {
            const float2 t0 = in.triUV[3 * size_t(prim) + 0], t1 = in.triUV[3 * size_t(prim) + 1], t2 = in.triUV[3 * size_t(prim) + 2];
            uv = make_float2(t0.x * a + t1.x * b + t2.x * c, t0.y * a + t1.y * b + t2.y * c);
        }
        const Float3 rgb = GLOSSY ? SampleTextureGrad(d.textures[texIndex], uv.x, uv.y, dpdx, dpdy, d.textureLodMode) : SampleTexture(d.textures[texIndex], uv.x, uv.y);
        if(d.spectral) { const float3 cf = FetchAlbedoCoeffs(d.spec, rgb.x, rgb.y, rgb.z); albedoRaw = make_float4(cf.x, cf.y, cf.z, 0.f); }
        else albedoRaw = make_float4(rgb.x, rgb.y, rgb.z, 0.f);
    }
    const Spec albedo = AlbedoAt(d, albedoRaw, waves);
    // (Mt)Unreal (MaterialsDefault.hpp:L466-760): GGX + Burley diffuse; specular enough (>= 0.95) it is treated like a mirror
    const bool unreal = GLOSSY && matType == MAT_UNREAL;
    UnrealBxDF ub; ub.albedo = albedo; ub.roughness = ub.specular = ub.metallic = 0.0f;
    if(unreal) { const float4 mp = d.matParams[2 * matIndex]; ub.roughness = mp.x; ub.specular = mp.y; ub.metallic = mp.z; }
    const bool specularMat = unreal && ub.Specularity() >= SPECULAR_THRESHOLD;
    // orthonormal frame about the shading normal (any frame does: both BxDFs are isotropic)
    const Float3 hlp = (fabsf(shadeN.x) > 0.9f) ? F3(0, 1, 0) : F3(1, 0, 0);
    const Float3 tX = Normalize(Cross(hlp, shadeN));
    const Float3 tY = Cross(shadeN, tX);
    const Float3 V = F3(Dot(wO, tX), Dot(wO, tY), Dot(wO, shadeN));

    // ---- NEE (WorkFunctionNEE::Call) ----
    uint32_t newType = type;
    float4 shadowRad = make_float4(0.f, 0.f, 0.f, 0.f);
    if(d.sampleMode != 0u)
    {
        float xiL[3]; rng.Next<3>(xiL);   // light sample (2-D) + light selection
        const float x0 = xiL[0], x1 = xiL[1], xs = xiL[2];
        const uint32_t nLights = d.lightCount + 1u; // + boundary light
        uint32_t li = min(uint32_t(xs * float(nLights)), nLights - 1u);
        newType = RAY_SHADOW;
        neeSample = !specularMat;
        if((li < d.lightCount || (GLOSSY && d.boundaryType != 0u)) && !specularMat)
        {
            Float3 lpos; float pdfL; Spec em;
            if(!GLOSSY || li < d.lightCount)
            {
                const EmissiveTri l = d.lights[li];
                // Triangle::SampleSurface (Osada)
                const float r1s = sqrtf(x0), r2s = x1;
                const float la = 1.0f - r1s, lb = (1.0f - r2s) * r1s, lc = r1s * r2s;
                const Float3 lp0 = F3(l.p0.x, l.p0.y, l.p0.z), le0 = F3(l.e0.x, l.e0.y, l.e0.z), le1 = F3(l.e1.x, l.e1.y, l.e1.z);
                lpos = lp0 * la + (lp0 + le0) * lb + (lp0 + le1) * lc;
                const Float3 lN = Normalize(Cross(le0, le1));
                // LightPrim::SampleSolidAngle
                Float3 sdir = pos - lpos;
                const float distSqr = Dot(sdir, sdir);
                sdir = Normalize(sdir);
                float NdL = Dot(lN, sdir);
                NdL = (l.e0.w != 0.0f) ? fabsf(NdL) : fmaxf(0.0f, NdL);
                pdfL = (NdL == 0.0f) ? 0.0f : (1.0f / l.p0.w) / NdL;
                pdfL *= distSqr;
                pdfL *= 1.0f / float(nLights);
                em = Emit(d, l, lN, sdir, waves);
            }
            else
            {
                // LightSkysphere::SampleSolidAngle (LightsDefault.hpp:L338-357): a direction from the luminance distribution
                // (uniform in uv for a constant radiance), a point one scene diameter away along it
                const float3 suv = (d.boundaryTex < 0) ? make_float3(x0, x1, 1.0f) : DistSampleUV(d.boundaryDist, x0, x1);
                const float3 dl = SkyUVToDir(d.boundaryType, suv.x, suv.y);
                Float3 worldDir = F3(dl.x, dl.y, dl.z);
                if(!d.boundaryIdentity) worldDir = SkyApply(d.boundaryM, worldDir);
                pdfL = SkyPdfFromUV(d.boundaryType, suv.z, suv.y);
                lpos = pos + worldDir * d.sceneDiameter;
                pdfL *= 1.0f / float(nLights);
                // DirectLightSamplerUniform::SampleLight: EmitViaSurfacePoint(wO = towards the surface)
                em = SkyEmit(d, Normalize(pos - lpos), waves);
            }
            // LightSampleOutput::SampledRay
            const Float3 wI = Normalize(lpos - pos);
            const Float3 lposN = NudgePos(lpos, wI * -1.0f);
            const float len = Length(lposN - pos);
            // Material::Evaluate / Pdf
            Spec refl; float pdfBx;
            if(unreal)
            {
                const Float3 Lt = F3(Dot(wI, tX), Dot(wI, tY), Dot(wI, shadeN));
                refl = ub.Evaluate(V, Lt); pdfBx = ub.Pdf(V, Lt);
            }
            else
            {
                const float nDotL = fmaxf(Dot(shadeN, wI), 0.0f);
                refl = albedo * (nDotL * INV_PI_F);
                pdfBx = fmaxf(PdfCosDirection(Dot(shadeN, wI)), 0.0f);
            }
            float pdf = pdfL;
            if(d.sampleMode == 2u) pdf = pdfBx + pdfL;
            Spec sr = throughput * refl * em;
            sr = (pdf == 0.0f) ? S4(0, 0, 0, 0) : sr * (1.0f / pdf);
            // +2: depth is not incremented yet (KCAccumulateShadowRaysPT)
            if(depth + 2u <= d.rrHi) shadowRad = F4(sr);
            // The reference casts every shadow ray and adds the pre-multiplied estimate if visible; an
            // estimate that is exactly zero (light facing away, surface facing away, depth limit) adds
            // nothing either way, so its visibility ray is not cast at all (about a third of them on the
            // arcade scene). NaNs compare unequal to zero and keep the reference's behaviour.
            if(shadowRad.x != 0.0f || shadowRad.y != 0.0f || shadowRad.z != 0.0f || shadowRad.w != 0.0f)
            {
                const Float3 so = NudgePos(pos, geoN);
                float4* sp = reinterpret_cast<float4*>(d.shadowRays + i);
                sp[0] = make_float4(so.x, so.y, so.z, 1.0e-5f);
                sp[1] = make_float4(wI.x, wI.y, wI.z, len * (1.0f - 1.0e-4f));
                castShadow = true;
            }
        }
    }
    if(!castShadow) newType = type;
    d.shadowRadiance[i] = shadowRad;

    // ---- BxDF sample + Russian roulette (WorkFunction::Call) ----
    Float3 Lt; float pdfB; Spec reflS;
    if(unreal)
    {
        float xiS[1]; rng.Next<1>(xiS);
        float xiB[2]; rng.Next<2>(xiB);
        Lt = ub.Sample(V, xiS[0], xiB[0], xiB[1], reflS, pdfB);
    }
    else
    {   // LambertMaterial::SampleBxDF: cosine-weighted hemisphere
        float xiB[2]; rng.Next<2>(xiB);
        const float phi = 2.0f * PI_F * xiB[1], su = sqrtf(xiB[0]);
        float sn, cs; sincosf(phi, &sn, &cs);
        const float lx = su * cs, ly = su * sn;
        const float lz = sqrtf(fmaxf(0.0f, 1.0f - (lx * lx + ly * ly)));
        Lt = F3(lx, ly, lz);
        pdfB = lz * INV_PI_F;
        reflS = albedo * (fmaxf(lz, 0.0f) * INV_PI_F);
    }
    const Float3 wIw = Normalize(tX * Lt.x + tY * Lt.y + shadeN * Lt.z);
    throughput = throughput * reflS;
    depth += 1u;
    bool dead = depth >= d.rrHi;
    if(!dead && depth >= d.rrLo && !specularMat)
    {
        float xiR[1]; rng.Next<1>(xiR);
        const float rrXi = xiR[0];
        // rrFactor = throughput.Sum() * ChannelCountInv (PathTracerRendererShaders.h:L251-256)
        float prob = (throughput.x + throughput.y + throughput.z + throughput.w) * (d.spectral ? 0.25f : 0.33333333f);
        prob = fminf(fmaxf(prob, 0.1f), 1.0f);
        if(rrXi >= prob) dead = true;
        else throughput = throughput * (1.0f / prob);
    }
    if(d.samplerType == SAMPLER_INDEPENDENT) d.rng[i] = rng.state;
    else d.sampleState[i].y = rng.dim;
    // the NEXT hit's MIS decision: PATH_RAY, or SPECULAR_RAY after a near-mirror; the shadow flag only lives until the finish kernel
    const uint32_t nextType = (specularMat ? RAY_SPECULAR : RAY_PATH) | ((newType == RAY_SHADOW && castShadow) ? 0x80u : 0u);
    if(!dead)
    {
        throughput = (pdfB == 0.0f) ? S4(0, 0, 0, 0) : throughput * (1.0f / pdfB);
        d.throughput[i] = F4(throughput);
        d.meta[i].w = __float_as_uint(pdfB);
        const Float3 no = NudgePos(pos, geoN);
        float4* rp = reinterpret_cast<float4*>(d.rays + i);
        rp[0] = make_float4(no.x, no.y, no.z, 1.0e-4f);
        rp[1] = make_float4(wIw.x, wIw.y, wIw.z, FLT_MAX);
        if(GLOSSY && useCones) d.cones[i] = ConeAfterScatter(cs, wIw, geoN);
        d.meta[i].x = PackPD(depth, ST_ALIVE, nextType);
    }
    else
    {
        d.rays[i].tMin = 1.0f; d.rays[i].tMax = -1.0f;
        d.meta[i].x = PackPD(depth, ST_DEAD, nextType);
    }
    return true;
}

// resident blocks per SM the shading kernels are compiled for: both are bound by the latency of the path-state loads, so occupancy
// pays until the spills start to (lean: 4 blocks = 64 registers, no spills, 0.196 -> 0.158 ms per iteration against 3 blocks / 76
// registers; 5 blocks = 48 registers + spills: 0.167. full: 3 blocks = 80 registers + 58 B of spills, 23.6 ms/spp on the config-4
// material mix; 2 blocks / 125 registers 24.1; unconstrained 130 registers = ONE block: 26.6) — profiles/r2_experiments.md
#ifndef MRB_SHADE_MINBLOCKS
#define MRB_SHADE_MINBLOCKS 4
#endif
#ifndef MRB_GLOSSY_MINBLOCKS
#define MRB_GLOSSY_MINBLOCKS 3
#endif
template<bool GLOSSY>
__global__ void __launch_bounds__(RTPB, GLOSSY ? MRB_GLOSSY_MINBLOCKS : MRB_SHADE_MINBLOCKS) KShade(RenderData d)
{
    const uint32_t tidx = blockIdx.x * RTPB + threadIdx.x;
    bool alive = false, castShadow = false, neeSample = false;
    if(tidx < d.slots) alive = ShadeSlot<GLOSSY>(d, d.partitionRays ? d.workIndices[tidx] : tidx, castShadow, neeSample);
    // ray statistics: one atomic per block per counter
    const int nAlive = __syncthreads_count(alive), nShadow = __syncthreads_count(castShadow), nNee = __syncthreads_count(neeSample);
    if(threadIdx.x == 0)
    {
        if(nAlive) atomicAdd(&d.counters[2], (unsigned long long)nAlive);
        if(nShadow) atomicAdd(&d.counters[3], (unsigned long long)nShadow);
        if(nNee) atomicAdd(&d.counters[4], (unsigned long long)nNee);
    }
}

// KCConvertColor (Tracer/ColorConverter.cu:L306-398): gamma to linear, then the RGB -> RGB matrix into the global colour space
// (Matrix * Vector = Math::Dot FMA chains), in place on the uploaded texels; unorm8 goes through FromUNorm / ToUNorm.
struct ColorConv { float gamma; float m[9]; uint32_t hasMatrix; };
__global__ void __launch_bounds__(256) KConvertTextureColor(TexRec t, ColorConv cc)
{
    const size_t n = MipStart(t.w, t.h, t.mipCount ? t.mipCount : 1u);   // every supplied level (ColorConvParams.validMips)
    for(size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += size_t(gridDim.x) * 256ull)
    {
        float c[3];
        if(t.format == 0u) { const float* f = static_cast<const float*>(t.data) + i * t.channels; c[0] = f[0]; c[1] = f[1]; c[2] = f[2]; }
        else { const uint8_t* b = static_cast<const uint8_t*>(t.data) + i * t.channels; for(int k = 0; k < 3; k++) c[k] = __fmul_rn(float(b[k]), 1.0f / 255.0f); }
        if(cc.gamma != 1.0f) { c[0] = powf(c[0], cc.gamma); c[1] = powf(c[1], cc.gamma); c[2] = powf(c[2], cc.gamma); }
        if(cc.hasMatrix)
        {
            float o[3];
            #pragma unroll
            for(int r = 0; r < 3; r++)
            {
                float d = __fmaf_rn(cc.m[3 * r], c[0], 0.0f);
                d = __fmaf_rn(cc.m[3 * r + 1], c[1], d);
                o[r] = __fmaf_rn(cc.m[3 * r + 2], c[2], d);
            }
            c[0] = o[0]; c[1] = o[1]; c[2] = o[2];
        }
        if(t.format == 0u) { float* f = const_cast<float*>(static_cast<const float*>(t.data)) + i * t.channels; f[0] = c[0]; f[1] = c[1]; f[2] = c[2]; }
        else
        {
            uint8_t* b = const_cast<uint8_t*>(static_cast<const uint8_t*>(t.data)) + i * t.channels;
            for(int k = 0; k < 3; k++) b[k] = uint8_t(fminf(fmaxf(roundf(__fmul_rn(c[k], 255.0f)), 0.0f), 255.0f));   // ToUNorm (clamped: it asserts [0, 1])
        }
    }
}
static bool NeedsColorConversion(const mrb_texture_desc& td) { return (td.gamma != 0.0f && td.gamma != 1.0f) || td.colorMatrix != nullptr; }
static void ConvertTextureColor(Context& ctx, const TexRec& t, const mrb_texture_desc& td)
{
    if(!NeedsColorConversion(td)) return;
    if(t.channels < 3u) throw std::runtime_error("colour conversion needs a texture with at least 3 channels");
    ColorConv cc = {};
    cc.gamma = (td.gamma == 0.0f) ? 1.0f : td.gamma;
    cc.hasMatrix = td.colorMatrix ? 1u : 0u;
    if(td.colorMatrix) memcpy(cc.m, td.colorMatrix, sizeof(cc.m));
    MRB_LAUNCH(ctx, KConvertTextureColor, GridFor(ctx, t.w * t.h, 256u), 256, 0, t, cc);
}

// <Filter>::Evaluate(duv) of Tracer/Filters.h (L110-117 Box, L153-164 Tent, L202-206 Gaussian, L257-278 Mitchell-Netravali)
__device__ __forceinline__ float FilterEvaluate(uint32_t type, float r, float x, float y)
{
    if(type == FILTER_BOX) { const float rr = 1.0f / r; return (fabsf(x) <= r && fabsf(y) <= r) ? 0.25f * rr * rr : 0.0f; }
    if(type == FILTER_TENT) { const float rcp = 1.0f / r; return LerpU(rcp, 0.0f, fabsf(x * rcp)) * LerpU(rcp, 0.0f, fabsf(y * rcp)); }
    if(type == FILTER_MITCHELL) { const float rcp = 1.0f / r; return Mitchell1D(x, rcp) * Mitchell1D(y, rcp); }
    const float sigma = r * 0.285714f;
    return GaussPdf(x, sigma) * GaussPdf(y, sigma);
}
// TextureMemory::GenerateMipmaps -> KCGenerateMipmaps (Tracer/TextureFilter.cu:L55-118,L126-198,L1064-1096): one level from its parent.
// Per texel 8 x 8 stratified offsets over [-r, r]^2 (FilterMode::ACCUMULATE), weight = Evaluate(offset), the parent texel nearest
// to the offset pixel centre (ConvertPixelIndices + RoundInt), sum / weight sum; unorm8 texels are filtered as their 0..255 integer
// values, rounded and clamped (GenericRead / GenericWrite). One thread per texel: the 64 parent reads of neighbouring texels overlap
// in L1, and a full chain is 1/3 of the base level, so the kernel is a small fraction of the upload it follows.
__global__ void __launch_bounds__(256) KGenerateMipLevel(TexRec t, uint32_t level, uint32_t filterType, float radius)
{
    const uint32_t mw = MipDim(t.w, level), mh = MipDim(t.h, level), pw = MipDim(t.w, level - 1u), ph = MipDim(t.h, level - 1u);
    const size_t dst = MipStart(t.w, t.h, level), src = MipStart(t.w, t.h, level - 1u);
    const uint32_t n = mw * mh;
    for(uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u)
    {
        const uint32_t x = i % mw, y = i / mw;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, wsum = 0.0f;
        const float ratioX = float(pw) / float(mw), ratioY = float(ph) / float(mh);
        for(uint32_t sy = 0; sy < 8u; sy++)
        for(uint32_t sx = 0; sx < 8u; sx++)
        {
            const float dxy = 1.0f / 8.0f;
            const float xi0 = __fadd_rn(dxy * 0.5f, __fmul_rn(dxy, float(sx))), xi1 = __fadd_rn(dxy * 0.5f, __fmul_rn(dxy, float(sy)));
            const float ox = __fsub_rn(__fmul_rn(__fmul_rn(xi0, 2.0f), radius), radius), oy = __fsub_rn(__fmul_rn(__fmul_rn(xi1, 2.0f), radius), radius);
            const float wgt = FilterEvaluate(filterType, radius, ox, oy);
            float rx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(float(x), ox), 0.5f), ratioX), 0.5f);
            float ry = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(float(y), oy), 0.5f), ratioY), 0.5f);
            rx = fminf(fmaxf(rx, 0.0f), float(pw) - 1.0f); ry = fminf(fmaxf(ry, 0.0f), float(ph) - 1.0f);
            const size_t o = (src + size_t(lroundf(ry)) * pw + size_t(lroundf(rx))) * t.channels;
            for(uint32_t c = 0; c < t.channels; c++)
            {
                const float px = (t.format == 0u) ? static_cast<const float*>(t.data)[o + c] : float(static_cast<const uint8_t*>(t.data)[o + c]);
                acc[c] = __fadd_rn(acc[c], __fmul_rn(wgt, px));
            }
            wsum = __fadd_rn(wsum, wgt);
        }
        const size_t o = (dst + size_t(y) * mw + x) * t.channels;
        for(uint32_t c = 0; c < t.channels; c++)
        {
            const float v = __fdiv_rn(acc[c], wsum);
            if(t.format == 0u) const_cast<float*>(static_cast<const float*>(t.data))[o + c] = v;
            else const_cast<uint8_t*>(static_cast<const uint8_t*>(t.data))[o + c] = uint8_t(fminf(fmaxf(roundf(v), 0.0f), 255.0f));
        }
    }
}
// ClampImageFromBuffer -> KCClampImage (Tracer/TextureFilter.cu:L206-262,L1100-1150), TracerParameters.clampedTexRes: the pushed image
// `src` (sw x sh, the texture's texel format) filtered down to level 0 of t. Per texel 4 x 4 stratified numbers through the mip filter's
// Sample() (FilterMode::SAMPLING), weight = Evaluate / pdf / 16, the source texel nearest to the offset pixel centre
// (ConvertPixelIndices + Math::Round), sum / weight sum.
__global__ void __launch_bounds__(256) KClampImage(TexRec t, const void* __restrict__ src, uint32_t sw, uint32_t sh, uint32_t filterType, float radius)
{
    const uint32_t n = t.w * t.h;
    const float ratioX = float(sw) / float(t.w), ratioY = float(sh) / float(t.h);
    for(uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u)
    {
        const uint32_t x = i % t.w, y = i / t.w;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, wsum = 0.0f;
        for(uint32_t sy = 0; sy < 4u; sy++)
        for(uint32_t sx = 0; sx < 4u; sx++)
        {
            const float dxy = 0.25f, inv = 0.0625f;
            const float xi0 = __fadd_rn(dxy * 0.5f, __fmul_rn(dxy, float(sx))), xi1 = __fadd_rn(dxy * 0.5f, __fmul_rn(dxy, float(sy)));
            float ox, oy, pdf, wgt;
            FilterSample(filterType, radius, xi0, xi1, ox, oy, pdf, wgt);
            float rx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(float(x), ox), 0.5f), ratioX), 0.5f);
            float ry = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(float(y), oy), 0.5f), ratioY), 0.5f);
            rx = fminf(fmaxf(rx, 0.0f), float(sw) - 1.0f); ry = fminf(fmaxf(ry, 0.0f), float(sh) - 1.0f);
            const size_t o = (size_t(uint32_t(roundf(ry))) * sw + size_t(uint32_t(roundf(rx)))) * t.channels;
            for(uint32_t c = 0; c < t.channels; c++)
            {
                const float px = (t.format == 0u) ? static_cast<const float*>(src)[o + c] : float(static_cast<const uint8_t*>(src)[o + c]);
                acc[c] = __fadd_rn(acc[c], __fdiv_rn(__fmul_rn(__fmul_rn(wgt, px), inv), pdf));
            }
            wsum = __fadd_rn(wsum, __fdiv_rn(__fmul_rn(wgt, inv), pdf));
        }
        const size_t o = size_t(i) * t.channels;
        for(uint32_t c = 0; c < t.channels; c++)
        {
            const float v = __fdiv_rn(acc[c], wsum);
            if(t.format == 0u) const_cast<float*>(static_cast<const float*>(t.data))[o + c] = v;
            else const_cast<uint8_t*>(static_cast<const uint8_t*>(t.data))[o + c] = uint8_t(fminf(fmaxf(roundf(v), 0.0f), 255.0f));
        }
    }
}
// TracerParameters.clampedTexRes (TextureMemory::CreateTexture, Tracer/TextureMemory.cpp:L544-583): the number of levels dropped so that
// the larger side fits clampResolution = ceil(log2(ceil(maxDim / min(clamp, maxDim)))); 0 = no clamp
static uint32_t ClampLevels(const mrb_texture_desc& td)
{
    if(td.clampResolution == 0u) return 0u;
    const uint32_t maxDim = max(td.width, td.height), c = min(td.clampResolution, maxDim);
    return uint32_t(int32_t(ceilf(log2f(float((maxDim + c - 1u) / c)))));
}
// levels a texture ends up with: the supplied ones (less the levels the clamp drops, never below one), or the full chain of the final
// size when mips are generated (TextureMemory::CreateTexture, L556,L588-591)
static uint32_t SuppliedMips(const mrb_texture_desc& td) { return td.mipCount ? td.mipCount : 1u; }
static uint32_t KeptMips(const mrb_texture_desc& td) { const uint32_t s = SuppliedMips(td), r = ClampLevels(td); return r >= s ? 1u : s - r; }
static uint32_t FinalWidth(const mrb_texture_desc& td) { return MipDim(td.width, ClampLevels(td)); }
static uint32_t FinalHeight(const mrb_texture_desc& td) { return MipDim(td.height, ClampLevels(td)); }
static uint32_t FinalMips(const mrb_texture_desc& td) { return td.generateMips ? max(KeptMips(td), FullMipCount(FinalWidth(td), FinalHeight(td))) : KeptMips(td); }
static size_t ChainTexels(uint32_t w, uint32_t h, uint32_t mips) { return MipStart(w, h, mips); }
static void ValidateMips(const mrb_texture_desc& td)
{
    if(SuppliedMips(td) > FullMipCount(td.width, td.height)) throw std::runtime_error("texture mipCount exceeds the full chain of its size");
    if((td.generateMips || (td.clampResolution && td.mipFilterRadius != 0.0f)) && (td.mipFilterType > FILTER_MITCHELL || !(td.mipFilterRadius > 0.0f)))
        throw std::runtime_error("bad mip generation filter");
}
// upload + TextureMemory::Finalize (L809-833): resolution clamp at load, colour conversion of the kept levels, then the missing levels
// of the chain. t = the texture at its FINAL size (FinalWidth x FinalHeight, FinalMips levels allocated).
static void UploadTexture(Context& ctx, TexRec& t, const mrb_texture_desc& td)
{
    const size_t texel = size_t(t.channels) * (t.format == 0u ? 4u : 1u);
    const uint32_t supplied = SuppliedMips(td), drop = ClampLevels(td), kept = KeptMips(td);
    const char* src = static_cast<const char*>(td.data);
    if(drop >= supplied && drop > 0u)
    {   // fewer levels supplied than the clamp drops: the last supplied level is filtered down to the new level 0 (KCClampImage)
        const uint32_t sl = supplied - 1u, sw = MipDim(td.width, sl), sh = MipDim(td.height, sl);
        const size_t bytes = size_t(sw) * sh * texel;
        DeviceBlock stage; stage.Reserve(bytes);
        MRB_CUDA_TRY(cudaMemcpyAsync(stage.Base(), src + MipStart(td.width, td.height, sl) * texel, bytes, cudaMemcpyHostToDevice, ctx.stream));
        MRB_LAUNCH(ctx, KClampImage, GridFor(ctx, t.w * t.h, 256u), 256, 0, t, stage.Base(), sw, sh, td.mipFilterRadius > 0.0f ? td.mipFilterType : uint32_t(FILTER_GAUSSIAN),
                   td.mipFilterRadius > 0.0f ? td.mipFilterRadius : 2.0f);
        MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));   // the staging block dies with this scope
    }
    else
    {   // enough levels supplied: levels drop .. supplied - 1 become levels 0 .. kept - 1 (the reference copies the un-shifted levels
        // here, TextureMemory.cpp:L785-794, which reads the wrong texels; the levels that fit are used instead)
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<void*>(t.data), src + MipStart(td.width, td.height, drop) * texel, ChainTexels(t.w, t.h, kept) * texel,
                                     cudaMemcpyHostToDevice, ctx.stream));
    }
    t.mipCount = kept;
    ConvertTextureColor(ctx, t, td);
    t.mipCount = FinalMips(td);
    for(uint32_t level = kept; level < t.mipCount; level++)
        MRB_LAUNCH(ctx, KGenerateMipLevel, GridFor(ctx, MipDim(t.w, level) * MipDim(t.h, level), 256u), 256, 0, t, level, td.mipFilterType, td.mipFilterRadius);
}

// Converter::ConvertAlbedo / ConvertRadiance, LUT half: RGB attributes -> Jakob coefficients, once per render
__global__ void KPrepareSpectral(SpectrumData s, float4* albedo, uint32_t materialCount, EmissiveTri* lights, uint32_t lightCount,
                                 float4 boundaryRGB, float4* boundaryOut)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i == 0u && boundaryOut) *boundaryOut = FetchRadianceCoeffs(s, boundaryRGB.x, boundaryRGB.y, boundaryRGB.z);
    if(i < materialCount)
    {
        const float4 a = albedo[i];
        const float3 c = FetchAlbedoCoeffs(s, a.x, a.y, a.z);
        albedo[i] = make_float4(c.x, c.y, c.z, 0.0f);
    }
    else if(i < materialCount + lightCount)
    {
        EmissiveTri& l = lights[i - materialCount];
        l.radiance = FetchRadianceCoeffs(s, l.radiance.x, l.radiance.y, l.radiance.z);
    }
}

__global__ void KIndexInstances(InstanceRec* inst, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) inst[i].accelKey = i;
}

// End of a bounce for one slot: add the NEE estimate of an unoccluded shadow ray, put finished paths on the
// film. All inputs are requested up front (one memory round trip). Returns true when the slot is free
// afterwards; `died` reports a path that finished in this call.
// A dead path leaves the pool: its last shadow contribution, ConvertSpectrumToRGBIndirect (PathTracerRendererBase.cu:L228-241),
// ConvertNaNsToColor and the atomic add into the film (planar R,G,B,W).
__device__ __forceinline__ void FilmDeadPath(const RenderData& d, uint32_t i)
{
    const uint4 meta = d.meta[i];
    float4 rad = d.radiance[i];
    const uint32_t visWord = d.visible[i >> 5];
    const float4 sr = d.shadowRadiance[i];
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = w4;
    if(d.spectral) { w4 = d.waves[i]; p4 = d.wavePdf[i]; }
    if((((meta.x >> 16) & 0xFFu) & 0x80u) && ((visWord >> (i & 31u)) & 1u)) { rad.x += sr.x; rad.y += sr.y; rad.z += sr.z; rad.w += sr.w; }
    float w = __uint_as_float(meta.z);
    Float3 v = F3(rad.x, rad.y, rad.z);
    if(d.spectral)
    {
        const float val[4] = {rad.x, rad.y, rad.z, rad.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
        const float3 rgb = SpectraToRGB(d.spec, val, wv, pp);
        v = F3(rgb.x, rgb.y, rgb.z);
    }
    if(!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z))) { v = F3(1e7f, 0.f, 1e7f); w *= 128.0f; }
    const uint32_t pix = meta.y;
    const size_t plane = size_t(d.width) * d.height;
    atomicAdd(d.film + pix, v.x);
    atomicAdd(d.film + plane + pix, v.y);
    atomicAdd(d.film + 2 * plane + pix, v.z);
    atomicAdd(d.film + 3 * plane + pix, w);
    d.meta[i].x = PackPD(0, ST_INVALID, 0);
    d.rays[i].tMin = 1.0f; d.rays[i].tMax = -1.0f;
}

// The end of bounce k (shadow accumulate, film) fused with the reload of bounce k+1: the slot a path just left is refilled in the same pass
// Finish + reload of one block of slots. The expensive branch — a dead path's spectrum -> RGB conversion, its four film atomics and
// the whole camera-ray generation of the slot's next path — concerns about a third of the slots, scattered over the warps (6.7 of 32
// lanes active per instruction when every thread handled its own slot). So the block first settles the live paths, compacts the
// indices of the slots to refill into shared memory (ballot order: the same slot <-> path assignment as before) and then lets
// consecutive threads work through that list with full warps.
#ifndef MRB_FINISH_MINBLOCKS
#define MRB_FINISH_MINBLOCKS 6   // 40 registers; 8 blocks (32 registers + spills) measured slower: 0.129 -> 0.143 ms
#endif
__global__ void __launch_bounds__(RTPB, MRB_FINISH_MINBLOCKS) KFinishReload(RenderData d)
{
    __shared__ uint16_t sList[RTPB];
    __shared__ uint32_t sWarpBase[RTPB / 32];
    __shared__ uint32_t sTotal, sDied;
    __shared__ unsigned long long sBase;
    const uint32_t blockBase = blockIdx.x * RTPB, i = blockBase + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool inRange = i < d.slots;
    // phase A: live paths take their shadow contribution; dead / empty slots are only classified
    bool isFree = false, died = false;
    if(inRange)
    {
        const uint32_t pd = d.meta[i].x;
        const uint32_t status = (pd >> 8) & 0xFFu;
        uint32_t type = (pd >> 16) & 0xFFu;
        if(status == ST_INVALID) isFree = true;
        else if(status == ST_DEAD) { isFree = true; died = true; }
        else
        {
            if(type & 0x80u)
            {   // shadow ray of this bounce: add the pre-multiplied NEE estimate when unoccluded
                if((d.visible[i >> 5] >> (i & 31u)) & 1u)
                {
                    float4 rad = d.radiance[i]; const float4 sr = d.shadowRadiance[i];
                    rad.x += sr.x; rad.y += sr.y; rad.z += sr.z; rad.w += sr.w;
                    d.radiance[i] = rad;
                }
                type &= 0x7Fu;
            }
            d.meta[i].x = PackPD(pd & 0xFFu, ST_ALIVE, type);
            d.hitKeys[i].primKey = INVALID_U32;   // default: boundary (KCSetBoundaryWorkKeysIndirect)
        }
    }
    const uint32_t want = __ballot_sync(0xffffffffu, isFree), dead = __ballot_sync(0xffffffffu, died);
    if(lane == 0) sWarpBase[warp] = uint32_t(__popc(want)) | (uint32_t(__popc(dead)) << 16);
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t total = 0, nDied = 0;
        #pragma unroll
        for(int k = 0; k < RTPB / 32; k++) { const uint32_t c = sWarpBase[k]; sWarpBase[k] = total; total += c & 0xFFFFu; nDied += c >> 16; }
        sTotal = total; sDied = nDied;
        // block-aggregated claim of new path indices: ONE atomic per block on the global counter
        sBase = total ? atomicAdd(&d.counters[0], (unsigned long long)total) : 0ull;
        if(nDied) atomicAdd(&d.counters[1], (unsigned long long)nDied);
    }
    __syncthreads();
    if(isFree) sList[sWarpBase[warp] + __popc(want & ((1u << lane) - 1u))] = uint16_t(threadIdx.x | (died ? 0x8000u : 0u));
    __syncthreads();
    // phase B: thread k refills the k-th free slot of the block
    const uint32_t total = sTotal;
    for(uint32_t k = threadIdx.x; k < total; k += RTPB)
    {
        const uint32_t e = sList[k], slot = blockBase + (e & 0x7FFFu);
        if(e & 0x8000u) FilmDeadPath(d, slot);
        ReloadClaimed(d, slot, sBase + k);
    }
}

} // namespace
} // namespace mrb

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct mrb_renderer_t
{
    mrb::RenderData  d = {};
    mrb::DeviceBlock mem;
    mrb_accel        accel = nullptr;
    mrb_scene        scene = nullptr;
    mrb::SceneData   sceneData;        // scene->d with the renderer's instance records (accelKey = instance index)
    uint64_t         iterations = 0;
    bool             needReload = true;   // the next iteration starts with KReload (first one, or a pass has just begun)
    bool             glossy = false;      // any (Mt)Refract / (Mt)Unreal material, or a skysphere boundary: KShade<true>
    // passes: samples [sampleBase, sampleBase + passSamples) of every pixel of the current region
    uint32_t         maxWidth = 0, maxHeight = 0;      // the film allocation (the largest region a pass may take)
    uint32_t         totalSPP = 0, sampleOffset = 0;
    uint32_t         sppStarted = 0;                   // set_spp_limit bookkeeping: samples per pixel handed out so far
    uint64_t         startedBefore = 0;                // camera paths of the finished passes
    uint64_t         completedTarget = 0;              // counters[1] when every pass begun so far has finished
    // film double buffer + asynchronous hand-off (RenderImage::TransferToHost, Tracer/RenderImage.cpp:L163-219)
    float*           film[2] = {nullptr, nullptr};
    int              cur = 0;
    cudaStream_t     copyStream = nullptr;
    cudaEvent_t      evCompute = nullptr, evCopied[2] = {nullptr, nullptr};
    bool             copied[2] = {false, false};
    // pipelined completion polling of run_pass (pinned host copies of the counters)
    unsigned long long* hCounters = nullptr;           // 2 x 8
    cudaEvent_t      evPoll[2] = {nullptr, nullptr};
    unsigned long long lastCounters[8] = {};           // poll_stats: the latest snapshot that has landed
    bool             pollPending = false;
    ~mrb_renderer_t()
    {
        if(copyStream) { cudaStreamSynchronize(copyStream); cudaStreamDestroy(copyStream); }
        if(evCompute) cudaEventDestroy(evCompute);
        for(int k = 0; k < 2; k++) { if(evCopied[k]) cudaEventDestroy(evCopied[k]); if(evPoll[k]) cudaEventDestroy(evPoll[k]); }
        if(hCounters) cudaFreeHost(hCounters);
    }
};

namespace mrb
{

void CreateRenderer(Context& ctx, mrb_renderer_t& r, const mrb_render_desc& desc)
{
    using namespace mrb;
    RenderData& d = r.d;
    r.accel = desc.accel; r.scene = desc.scene;
    d.sceneMode = desc.scene ? 1u : 0u;
    d.width = desc.width; d.height = desc.height;
    d.fullWidth = desc.fullResolution[0] ? desc.fullResolution[0] : desc.width;
    d.fullHeight = desc.fullResolution[1] ? desc.fullResolution[1] : desc.height;
    d.regionX = desc.regionMin[0]; d.regionY = desc.regionMin[1];
    if(d.regionX + d.width > d.fullWidth || d.regionY + d.height > d.fullHeight) throw std::runtime_error("render region exceeds the image");
    d.rrLo = desc.rrRange[0]; d.rrHi = desc.rrRange[1]; d.sampleMode = desc.sampleMode;
    // a new renderer stands at the start of ONE pass over all of its samples (throughput mode)
    d.pathLimit = uint64_t(desc.totalSPP) * desc.width * desc.height;
    d.sampleBase = desc.sampleOffset; d.seed = desc.seed;
    r.totalSPP = desc.totalSPP; r.sampleOffset = desc.sampleOffset; r.sppStarted = desc.totalSPP;
    r.maxWidth = desc.width; r.maxHeight = desc.height;
    r.completedTarget = d.pathLimit;
    if(desc.filmFilterType > FILTER_MITCHELL) throw std::runtime_error("unknown film filter type");
    if(!(desc.filmFilterRadius > 0.0f)) throw std::runtime_error("film filter radius must be positive");
    d.filterType = desc.filmFilterType; d.filterRadius = desc.filmFilterRadius;
    d.slots = desc.maxPathCount ? desc.maxPathCount : desc.width * desc.height;
    d.partitionRays = desc.partitionRays ? 1u : 0u;
    {   // bits needed for the larger of the material / light tables (Bit::RequiredBitsToRepresent)
        uint32_t m = desc.materialCount > desc.lightCount ? desc.materialCount : desc.lightCount;
        d.matBits = 1; while((1u << d.matBits) < m) d.matBits++;
    }
    // camera (CameraPinhole ctor)
    auto V = [](const float* p) { return make_float3(p[0], p[1], p[2]); };
    auto sub = [](float3 x, float3 y) { return make_float3(x.x - y.x, x.y - y.y, x.z - y.z); };
    auto cross = [](float3 x, float3 y) { return make_float3(x.y * y.z - x.z * y.y, x.z * y.x - x.x * y.z, x.x * y.y - x.y * y.x); };
    auto norm = [](float3 x) { float l = 1.0f / sqrtf(x.x * x.x + x.y * x.y + x.z * x.z); return make_float3(x.x * l, x.y * l, x.z * l); };
    float3 pos = V(desc.camPosition), gazeDir = sub(V(desc.camGaze), pos), up = V(desc.camUp);
    float3 right = norm(cross(gazeDir, up));
    up = norm(cross(right, gazeDir));
    gazeDir = norm(cross(up, right));
    float wh = tanf(desc.fovXY[0] * 0.5f) * desc.nearFar[0], hh = tanf(desc.fovXY[1] * 0.5f) * desc.nearFar[0];
    float3 bl = make_float3(pos.x - right.x * wh - up.x * hh + gazeDir.x * desc.nearFar[0],
                            pos.y - right.y * wh - up.y * hh + gazeDir.y * desc.nearFar[0],
                            pos.z - right.z * wh - up.z * hh + gazeDir.z * desc.nearFar[0]);
    d.cam.position = {pos.x, pos.y, pos.z}; d.cam.right = {right.x, right.y, right.z}; d.cam.up = {up.x, up.y, up.z};
    d.cam.bottomLeft = {bl.x, bl.y, bl.z}; d.cam.planeW = 2.0f * wh; d.cam.planeH = 2.0f * hh;
    d.cam.tNear = desc.nearFar[0]; d.cam.tFar = desc.nearFar[1];

    // instance list: the scene's instances, or one identity instance of the single accelerator
    struct HostInst { const mrb_accel_t* acc; const float* m; const float* inv; bool identity; const float* normals; const float* uvs; const float* tbn; };
    static const float IDENTITY[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    std::vector<HostInst> hinst;
    if(desc.scene)
        for(size_t k = 0; k < desc.scene->hInstances.size(); k++)
        {
            const mrb_instance_desc& id = desc.scene->hInstances[k];
            hinst.push_back({id.accel, id.transform, id.invTransform, id.isIdentity != 0,
                             desc.instanceVertexNormals ? desc.instanceVertexNormals[k] : nullptr,
                             desc.instanceVertexUVs ? desc.instanceVertexUVs[k] : nullptr,
                             desc.instanceVertexTBN ? desc.instanceVertexTBN[k] : nullptr});
        }
    else hinst.push_back({desc.accel, IDENTITY, IDENTITY, true, desc.vertexNormals, desc.vertexUVs, desc.vertexTBN});
    const uint32_t instCount = uint32_t(hinst.size());

    // emissive triangle list (MetaLightArrayT::Construct: one meta light per emissive triangle of every
    // instance, in world space), host side
    std::vector<EmissiveTri> lights;
    std::vector<std::vector<uint32_t>> lightOfPrim(instCount);
    {
        std::vector<float> hpos; std::vector<uint32_t> hidx;
        const mrb_accel_t* loaded = nullptr;
        for(uint32_t k = 0; k < instCount; k++)
        {
            const mrb_accel_t& hacc = *hinst[k].acc;
            const AccelData& a = hacc.d;
            lightOfPrim[k].assign(hacc.triangleCount, INVALID_U32);
            bool anyLight = false;
            // an instance may carry its own keys (mrb_instance_desc.lightOrMatKeys)
            const std::vector<uint32_t>& lmKeys = (desc.scene && !desc.scene->hInstanceKeys[k].empty()) ? desc.scene->hInstanceKeys[k] : hacc.hLmKey;
            for(uint32_t rg = 0; rg < a.ranges.count; rg++)
            {
                anyLight |= (lmKeys[rg] & 0x80000000u) != 0;
                // a material key indexes albedo / materialType / albedoTexture in the shading kernel
                if(!(lmKeys[rg] & 0x80000000u) && (lmKeys[rg] & 0x1FFFFFu) >= desc.materialCount)
                    throw std::runtime_error("material key index exceeds materialCount");
            }
            if(!anyLight) continue;
            if(loaded != &hacc)
            {
                hpos.resize(size_t(hacc.vertexCount) * 3); hidx.resize(size_t(hacc.triangleCount) * 3);
                MRB_CUDA_TRY(cudaMemcpyAsync(hpos.data(), a.positions, hpos.size() * 4, cudaMemcpyDeviceToHost, ctx.stream));
                MRB_CUDA_TRY(cudaMemcpyAsync(hidx.data(), a.indices, hidx.size() * 4, cudaMemcpyDeviceToHost, ctx.stream));
                MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
                loaded = &hacc;
            }
            const float* m = hinst[k].m;
            auto World = [&](const float* p, float* o)
            {
                if(hinst[k].identity) { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; return; }
                for(int rrow = 0; rrow < 3; rrow++)
                    o[rrow] = m[4 * rrow] * p[0] + m[4 * rrow + 1] * p[1] + m[4 * rrow + 2] * p[2] + m[4 * rrow + 3];
            };
            for(uint32_t rg = 0; rg < a.ranges.count; rg++)
            {
                uint32_t key = lmKeys[rg];
                if(!(key & 0x80000000u)) continue;
                uint32_t li = key & 0x1FFFFFu;
                if(li >= desc.lightCount) throw std::runtime_error("light key index exceeds lightCount");
                uint32_t count = hacc.hLeafStart[rg + 1] - hacc.hLeafStart[rg];
                for(uint32_t q = 0; q < count; q++)
                {
                    uint32_t prim = hacc.hPrimBegin[rg] + q;
                    float p0[3], p1[3], p2[3];
                    World(&hpos[3 * size_t(hidx[3 * size_t(prim)])], p0);
                    World(&hpos[3 * size_t(hidx[3 * size_t(prim) + 1])], p1);
                    World(&hpos[3 * size_t(hidx[3 * size_t(prim) + 2])], p2);
                    float e0[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e1[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
                    float cx = e0[1] * e1[2] - e0[2] * e1[1], cy = e0[2] * e1[0] - e0[0] * e1[2], cz = e0[0] * e1[1] - e0[1] * e1[0];
                    float area = 0.5f * sqrtf(cx * cx + cy * cy + cz * cz);
                    EmissiveTri t;
                    t.p0 = make_float4(p0[0], p0[1], p0[2], area);
                    t.e0 = make_float4(e0[0], e0[1], e0[2], (desc.lightTwoSided && desc.lightTwoSided[li]) ? 1.0f : 0.0f);
                    t.e1 = make_float4(e1[0], e1[1], e1[2], 0.0f);
                    t.radiance = make_float4(desc.lightRadiance[3 * li], desc.lightRadiance[3 * li + 1], desc.lightRadiance[3 * li + 2], 0.f);
                    lightOfPrim[k][prim] = uint32_t(lights.size());
                    lights.push_back(t);
                }
            }
        }
    }
    d.lightCount = uint32_t(lights.size());

    // boundary light surface (SetBoundarySurface): (L)Null or one of the skyspheres
    if(desc.boundaryType > 2u) throw std::runtime_error("unknown boundaryType");
    d.boundaryType = desc.boundaryType;
    if(desc.boundaryType != 0u) r.glossy = true;
    d.boundaryTex = -1;
    d.boundaryIdentity = 1u;
    uint32_t skyW = 0, skyH = 0;
    if(desc.boundaryType != 0u)
    {
        if(desc.boundaryTexture >= int32_t(desc.textureCount)) throw std::runtime_error("boundaryTexture index exceeds textureCount");
        d.boundaryTex = desc.boundaryTexture < 0 ? -1 : desc.boundaryTexture;
        if(d.boundaryTex >= 0) { skyW = FinalWidth(desc.textures[d.boundaryTex]); skyH = FinalHeight(desc.textures[d.boundaryTex]); }
        d.boundaryRadiance = make_float4(desc.boundaryRadiance[0], desc.boundaryRadiance[1], desc.boundaryRadiance[2], 0.0f);
        if(desc.boundaryTransform)
        {
            const float* m = desc.boundaryTransform;
            const double a[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
            const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
            if(det == 0.0) throw std::runtime_error("boundaryTransform is singular");
            const double inv[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                                   (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                                   (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det};
            bool ident = true;
            for(int k = 0; k < 9; k++) { d.boundaryM[k] = float(a[k]); d.boundaryInvM[k] = float(inv[k]); ident = ident && a[k] == ((k % 4 == 0) ? 1.0 : 0.0); }
            d.boundaryIdentity = ident ? 1u : 0u;
        }
        // TracerBase::CommitSurfaces (Tracer/TracerBase.cpp:L1653-1664): the diameter over the XZ plane of the scene AABB
        const float* bb = desc.scene ? desc.scene->aabb : desc.accel->info.aabb;
        const float sx = bb[3] - bb[0], sz = bb[5] - bb[2];
        d.sceneDiameter = desc.sceneDiameter > 0.0f ? desc.sceneDiameter : sqrtf(sx * sx + sz * sz);
    }

    std::vector<RenderInstance> hri(instCount);
    std::vector<TexRec> htex(desc.textureCount);
    bool anyMips = false;
    for(uint32_t t = 0; t < desc.textureCount; t++)
    {
        const mrb_texture_desc& td = desc.textures[t];
        if(!td.data || td.width == 0 || td.height == 0 || (td.channels != 3 && td.channels != 4) || td.format > 1u || td.interp > 1u || td.edge > 2u)
            throw std::runtime_error("bad texture descriptor");
        ValidateMips(td);
        htex[t] = TexRec{nullptr, FinalWidth(td), FinalHeight(td), td.channels, td.format, td.interp, td.edge, FinalMips(td)};
        if(FinalMips(td) > 1u) { anyMips = true; r.glossy = true; }   // ray cones + level selection live in the full shading kernel
    }
    if(desc.textureLodMode > 1u) throw std::runtime_error("unknown textureLodMode");
    d.textureLodMode = desc.textureLodMode;
    if(desc.albedoTexture)
        for(uint32_t m = 0; m < desc.materialCount; m++)
            if(desc.albedoTexture[m] >= int32_t(desc.textureCount)) throw std::runtime_error("albedoTexture index exceeds textureCount");
    if(desc.normalTexture)
        for(uint32_t m = 0; m < desc.materialCount; m++)
        {
            if(desc.normalTexture[m] >= int32_t(desc.textureCount)) throw std::runtime_error("normalTexture index exceeds textureCount");
            if(desc.normalTexture[m] >= 0 && !(desc.scene ? desc.instanceVertexTBN != nullptr : desc.vertexTBN != nullptr))
                throw std::runtime_error("normal maps need the per-vertex tangent frames (vertexTBN / instanceVertexTBN)");
        }
    InstanceRec* dSceneInst = nullptr;
    float* skyLuminance = nullptr; float* skyRowTotals = nullptr; float4* dBoundaryCoeffs = nullptr;
    auto Layout = [&](MultiAlloc& ma)
    {
        const uint32_t P = d.slots;
        d.rays = ma.Take<mrb_ray_gmem>(P); d.shadowRays = ma.Take<mrb_ray_gmem>(P);
        d.hitKeys = ma.Take<mrb_hit_key_pack>(P); d.hits = ma.Take<mrb_meta_hit>(P);
        d.throughput = ma.Take<float4>(P); d.radiance = ma.Take<float4>(P); d.shadowRadiance = ma.Take<float4>(P);
        d.meta = ma.Take<uint4>(P); d.rng = ma.Take<uint32_t>(P);
        d.sampleState = (desc.samplerType & 0xFFu) ? ma.Take<uint2>(P) : nullptr;
        d.sobolMatrices = (desc.samplerType & 0xFFu) ? ma.Take<uint32_t>(SOBOL_DIM_COUNT * SOBOL_MATRIX_WIDTH) : nullptr; d.visible = ma.Take<uint32_t>((P + 31) / 32);
        r.film[0] = ma.Take<float>(size_t(4) * d.width * d.height); r.film[1] = ma.Take<float>(size_t(4) * d.width * d.height);
        d.film = r.film[0];
        d.counters = ma.Take<unsigned long long>(8);
        d.albedo = ma.Take<float4>(desc.materialCount ? desc.materialCount : 1);
        d.lights = ma.Take<EmissiveTri>(lights.size() ? lights.size() : 1);
        d.instances = ma.Take<RenderInstance>(instCount);
        dSceneInst = desc.scene ? ma.Take<InstanceRec>(instCount) : nullptr;
        std::map<const mrb_accel_t*, const float4*> triPosOf;   // instances of one accelerator share its gathered positions
        for(uint32_t k = 0; k < instCount; k++)
        {
            auto it = triPosOf.find(hinst[k].acc);
            if(it == triPosOf.end()) it = triPosOf.emplace(hinst[k].acc, ma.Take<float4>(3 * size_t(hinst[k].acc->triangleCount))).first;
            hri[k].triPos = it->second;
            hri[k].lightOfPrim = ma.Take<uint32_t>(hinst[k].acc->triangleCount);
            hri[k].vertexNormals = (hinst[k].normals || hinst[k].tbn) ? ma.Take<float4>(hinst[k].acc->vertexCount) : nullptr;
            hri[k].vertexUVs = hinst[k].uvs ? ma.Take<float2>(hinst[k].acc->vertexCount) : nullptr;
            hri[k].triNrm = hri[k].vertexNormals ? ma.Take<float4>(3 * size_t(hinst[k].acc->triangleCount)) : nullptr;
            hri[k].triUV = hri[k].vertexUVs ? ma.Take<float2>(3 * size_t(hinst[k].acc->triangleCount)) : nullptr;
        }
        d.workKeys = ma.Take<uint32_t>(P); d.workIndices = ma.Take<uint32_t>(P); d.partTable = ma.Take<uint32_t>(32);
        d.albedoTex = desc.albedoTexture ? ma.Take<int32_t>(desc.materialCount ? desc.materialCount : 1) : nullptr;
        d.normalTex = desc.normalTexture ? ma.Take<int32_t>(desc.materialCount ? desc.materialCount : 1) : nullptr;
        d.materialType = desc.materialType ? ma.Take<uint8_t>(desc.materialCount ? desc.materialCount : 1) : nullptr;
        d.matParams = desc.materialParams ? ma.Take<float4>(2 * size_t(desc.materialCount ? desc.materialCount : 1)) : nullptr;
        d.textures = desc.textureCount ? ma.Take<TexRec>(desc.textureCount) : nullptr;
        for(uint32_t t = 0; t < desc.textureCount; t++)
            htex[t].data = ma.Take<char>(ChainTexels(htex[t].w, htex[t].h, htex[t].mipCount) * htex[t].channels * (htex[t].format == 0u ? 4u : 1u));
        d.cones = anyMips ? ma.Take<float2>(P) : nullptr;
        d.waves = desc.spectrum ? ma.Take<float4>(P) : nullptr; d.wavePdf = desc.spectrum ? ma.Take<float4>(P) : nullptr;
        // textured skysphere: the luminance distribution (row CDFs + marginal) and the scratch of its construction
        d.boundaryDist.cdfX = skyW ? ma.Take<float>(size_t(skyW) * skyH) : nullptr;
        d.boundaryDist.cdfY = skyW ? ma.Take<float>(skyH) : nullptr;
        skyLuminance = skyW ? ma.Take<float>(size_t(skyW) * skyH) : nullptr;
        skyRowTotals = skyW ? ma.Take<float>(skyH) : nullptr;
        dBoundaryCoeffs = (desc.spectrum && desc.boundaryType != 0u) ? ma.Take<float4>(1) : nullptr;
    };
    MultiAlloc sz(nullptr); Layout(sz);
    r.mem.Reserve(sz.Total());
    MultiAlloc ma(r.mem.Base()); Layout(ma);
    ctx.persistentBytes += r.mem.Capacity();
    MRB_CUDA_TRY(cudaMemsetAsync(r.mem.Base(), 0, sz.Total(), ctx.stream));

    std::vector<float4> halb(desc.materialCount ? desc.materialCount : 1, make_float4(0, 0, 0, 0));
    for(uint32_t m = 0; m < desc.materialCount; m++)
        halb[m] = make_float4(desc.albedo[3 * m], desc.albedo[3 * m + 1], desc.albedo[3 * m + 2], 0.f);
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float4*>(d.albedo), halb.data(), halb.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx.stream));
    if(desc.materialType)
    {
        for(uint32_t m = 0; m < desc.materialCount; m++)
        {
            if(desc.materialType[m] > MAT_UNREAL) throw std::runtime_error("unknown material type");
            if(desc.materialType[m] >= MAT_REFRACT && !desc.materialParams) throw std::runtime_error("(Mt)Refract / (Mt)Unreal need materialParams");
            if(desc.materialType[m] >= MAT_REFRACT) r.glossy = true;
        }
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(d.materialType), desc.materialType, desc.materialCount, cudaMemcpyHostToDevice, ctx.stream));
    }
    if(desc.materialParams)
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float4*>(d.matParams), desc.materialParams, sizeof(float4) * 2 * desc.materialCount, cudaMemcpyHostToDevice, ctx.stream));
    if(desc.albedoTexture)
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t*>(d.albedoTex), desc.albedoTexture, sizeof(int32_t) * desc.materialCount, cudaMemcpyHostToDevice, ctx.stream));
    if(desc.normalTexture)
    {
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t*>(d.normalTex), desc.normalTexture, sizeof(int32_t) * desc.materialCount, cudaMemcpyHostToDevice, ctx.stream));
        for(uint32_t m = 0; m < desc.materialCount; m++) if(desc.normalTexture[m] >= 0) r.glossy = true;   // normal maps live in the full shading kernel
    }
    for(uint32_t t = 0; t < desc.textureCount; t++) UploadTexture(ctx, htex[t], desc.textures[t]);   // + ConvertColorspaces + GenerateMipmaps
    if(desc.textureCount)
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<TexRec*>(d.textures), htex.data(), htex.size() * sizeof(TexRec), cudaMemcpyHostToDevice, ctx.stream));
    if(!lights.empty())
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<EmissiveTri*>(d.lights), lights.data(), lights.size() * sizeof(EmissiveTri), cudaMemcpyHostToDevice, ctx.stream));
    d.spectral = desc.spectrum ? 1u : 0u;
    if(desc.spectrum)
    {
        d.spec = desc.spectrum->d;
        const uint32_t total = desc.materialCount + uint32_t(lights.size());
        if(total || dBoundaryCoeffs)
            MRB_LAUNCH(ctx, KPrepareSpectral, DivUp(total ? total : 1u, 128u), 128, 0, d.spec, const_cast<float4*>(d.albedo), desc.materialCount,
                       const_cast<EmissiveTri*>(d.lights), uint32_t(lights.size()), d.boundaryRadiance, dBoundaryCoeffs);
        if(dBoundaryCoeffs)
        {   // the constant skysphere radiance as Jakob coefficients + scale, back into the by-value kernel parameters
            MRB_CUDA_TRY(cudaMemcpyAsync(&d.boundaryRadiance, dBoundaryCoeffs, sizeof(float4), cudaMemcpyDeviceToHost, ctx.stream));
            MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
        }
    }
    if(skyW)
    {   // LightGroupSkysphere::Finalize (Tracer/LightsDefault.hpp:L848-893): luminance of the radiance map -> PwC 2-D distribution
        static const float ACES_CG_Y[3] = {0x1.1614ep-2f, 0x1.58e6fep-1f, 0x1.d946e6p-5f};   // Color::Colorspace<MR_ACES_CG>::ToXYZMatrix row 1
        const bool given = desc.luminanceRow[0] != 0.0f || desc.luminanceRow[1] != 0.0f || desc.luminanceRow[2] != 0.0f;
        const TexRec& st = htex[d.boundaryTex];
        TextureLuminance(ctx, st.data, st.w, st.h, st.channels, st.format, given ? desc.luminanceRow : ACES_CG_Y, skyLuminance);
        Dist2DBuild(ctx, skyLuminance, skyW, skyH, const_cast<float*>(d.boundaryDist.cdfX), const_cast<float*>(d.boundaryDist.cdfY), skyRowTotals);
        d.boundaryDist.w = skyW; d.boundaryDist.h = skyH;
    }
    std::vector<float4> hn;
    std::set<const float4*> gathered;
    for(uint32_t k = 0; k < instCount; k++)
    {
        const mrb_accel_t& hacc = *hinst[k].acc;
        RenderInstance& ri = hri[k];
        ri.positions = hacc.d.positions; ri.indices = hacc.d.indices;
        if(hacc.triangleCount && gathered.insert(ri.triPos).second)
            MRB_LAUNCH(ctx, KGatherTriPositions, GridFor(ctx, hacc.triangleCount, 256u), 256, 0, hacc.d.positions, hacc.d.indices, hacc.triangleCount,
                       const_cast<float4*>(ri.triPos));
        memcpy(ri.transform, hinst[k].m, sizeof(ri.transform));
        memcpy(ri.invTransform, hinst[k].inv, sizeof(ri.invTransform));
        ri.identity = hinst[k].identity ? 1u : 0u;
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(ri.lightOfPrim), lightOfPrim[k].data(), lightOfPrim[k].size() * 4, cudaMemcpyHostToDevice, ctx.stream));
        if(hinst[k].uvs)
            MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float2*>(ri.vertexUVs), hinst[k].uvs, size_t(hacc.vertexCount) * sizeof(float2), cudaMemcpyHostToDevice, ctx.stream));
        ri.tbn = hinst[k].tbn ? 1u : 0u;
        if(hinst[k].tbn)
        {
            MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float4*>(ri.vertexNormals), hinst[k].tbn, size_t(hacc.vertexCount) * sizeof(float4), cudaMemcpyHostToDevice, ctx.stream));
            MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
        }
        else if(hinst[k].normals)
        {
            hn.resize(hacc.vertexCount);
            for(uint32_t v = 0; v < hacc.vertexCount; v++)
                hn[v] = make_float4(hinst[k].normals[3 * v], hinst[k].normals[3 * v + 1], hinst[k].normals[3 * v + 2], 0.f);
            MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float4*>(ri.vertexNormals), hn.data(), hn.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx.stream));
            MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream)); // hn is reused
        }
        // per-triangle copies of the per-vertex attributes (stream order: after their uploads above)
        if(ri.triNrm && hacc.triangleCount)
            MRB_LAUNCH(ctx, KGatherTriAttribute<float4>, GridFor(ctx, 3u * hacc.triangleCount, 256u), 256, 0, ri.vertexNormals, hacc.d.indices, hacc.triangleCount,
                       const_cast<float4*>(ri.triNrm));
        if(ri.triUV && hacc.triangleCount)
            MRB_LAUNCH(ctx, KGatherTriAttribute<float2>, GridFor(ctx, 3u * hacc.triangleCount, 256u), 256, 0, ri.vertexUVs, hacc.d.indices, hacc.triangleCount,
                       const_cast<float2*>(ri.triUV));
    }
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<RenderInstance*>(d.instances), hri.data(), hri.size() * sizeof(RenderInstance), cudaMemcpyHostToDevice, ctx.stream));
    if(desc.scene)
    {
        // the scene's instance records with accelKey = instance index, so the shading kernel finds the
        // instance of a hit without a key lookup (world AABBs are filled by the scene build: device copy)
        r.sceneData = desc.scene->d;
        MRB_CUDA_TRY(cudaMemcpyAsync(dSceneInst, desc.scene->d.instances, sizeof(InstanceRec) * instCount, cudaMemcpyDeviceToDevice, ctx.stream));
        MRB_LAUNCH(ctx, KIndexInstances, DivUp(instCount, 256u), 256, 0, dSceneInst, instCount);
        r.sceneData.instances = dSceneInst;
    }
    // samplers: no host-side generator table — every path derives its numbers from (seed, pixel, sample), see PathState()
    const bool lowDisc = (desc.samplerType & 0xFFu) != 0u;
    d.samplerType = lowDisc ? desc.samplerType : 0u;
    if(lowDisc)
    {
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(d.sobolMatrices), desc.sobolMatrices,
                                     sizeof(uint32_t) * SOBOL_DIM_COUNT * SOBOL_MATRIX_WIDTH, cudaMemcpyHostToDevice, ctx.stream));
        // ZSobol globals (Random.cu:L1172-1176): initialMaxSPP = the render's sample budget (of the WHOLE job when the
        // samples are split over GPUs), resMaxBits from the larger side of the full image
        uint32_t maxRes = d.fullWidth > d.fullHeight ? d.fullWidth : d.fullHeight, p2 = 1u, bits = 0u;
        while(p2 < maxRes) { p2 <<= 1; bits++; }
        d.zsobol.initialMaxSPP = desc.jobSPP ? desc.jobSPP : desc.sampleOffset + desc.totalSPP; d.zsobol.resMaxBits = bits;
    }
    MRB_CUDA_TRY(cudaStreamCreateWithFlags(&r.copyStream, cudaStreamNonBlocking));
    MRB_CUDA_TRY(cudaEventCreateWithFlags(&r.evCompute, cudaEventDisableTiming));
    for(int k = 0; k < 2; k++)
    {
        MRB_CUDA_TRY(cudaEventCreateWithFlags(&r.evCopied[k], cudaEventDisableTiming));
        MRB_CUDA_TRY(cudaEventCreateWithFlags(&r.evPoll[k], cudaEventDisableTiming));
    }
    MRB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&r.hCounters), sizeof(unsigned long long) * 16, cudaHostAllocPortable));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
}

// One wavefront iteration = one bounce of every live path. Every iteration ends with KFinishReload, so the slots
// freed by a bounce hold their next camera ray when the following iteration (of this or the next call) starts; the
// separate KReload only runs on the first iteration of a pass.
void RenderIterate(Context& ctx, mrb_renderer_t& r, uint32_t iterations)
{
    using namespace mrb;
    RenderData& d = r.d;
    const uint32_t grid = DivUp(d.slots, RTPB);
    for(uint32_t it = 0; it < iterations; it++)
    {
        ctx.prof.sampleNow = ctx.prof.enabled && (r.iterations % ctx.prof.stride) == 0;
        if(r.needReload) { MRB_LAUNCH(ctx, KReload, grid, RTPB, 0, d); r.needReload = false; }
        // stochastic alpha decisions of this iteration's two casts: a function of (seed, iteration) like every other random
        // number of the renderer, so a fixed seed still gives a fixed image
        {
            uint64_t z = d.seed + 0x9E3779B97F4A7C15ull * (r.iterations + 1ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            ctx.alphaSeed = uint32_t(z) & ~1u;
        }
        if(r.scene) TraceScene(ctx, r.sceneData, false, MRB_TRACE_WIDE, d.hitKeys, d.hits, nullptr, d.rays, nullptr, d.slots);
        else TraceRays(ctx, *r.accel, false, MRB_TRACE_WIDE, d.hitKeys, d.hits, nullptr, d.rays, nullptr, d.slots);
        if(d.partitionRays)
        {
            const uint32_t dataBits[2] = {0u, d.matBits}, batchBits[2] = {d.matBits, d.matBits + 2u};
            ctx.scratch.Reserve(MultiPartitionTempBytes(d.slots, 2));
            MRB_LAUNCH(ctx, KGenWorkKeys, grid, RTPB, 0, d);
            MultiPartition(ctx, d.workKeys, d.workIndices, d.slots, dataBits, batchBits, false, 4,
                           d.partTable, d.partTable + 1, d.partTable + 8, ctx.scratch.Base());
        }
        {
            ProfileScope ps(ctx, PROF_SHADE);
            if(r.glossy) MRB_LAUNCH(ctx, KShade<true>, grid, RTPB, 0, d); else MRB_LAUNCH(ctx, KShade<false>, grid, RTPB, 0, d);
        }
        if(d.sampleMode != 0u)
        {
            MRB_CUDA_TRY(cudaMemsetAsync(d.visible, 0xFF, sizeof(uint32_t) * ((d.slots + 31) / 32), ctx.stream));
            if(r.scene) TraceScene(ctx, r.sceneData, true, MRB_TRACE_WIDE, nullptr, nullptr, d.visible, d.shadowRays, nullptr, d.slots);
            else TraceRays(ctx, *r.accel, true, MRB_TRACE_WIDE, nullptr, nullptr, d.visible, d.shadowRays, nullptr, d.slots);
        }
        { ProfileScope ps(ctx, PROF_FINISH_RELOAD); MRB_LAUNCH(ctx, KFinishReload, grid, RTPB, 0, d); }
        r.iterations++;
    }
    ctx.prof.sampleNow = false;
}


mrb_renderer_t* NewRenderer() { return new mrb_renderer_t(); }

void DestroyRenderer(Context& ctx, mrb_renderer_t* r)
{
    if(!r) return;
    ctx.persistentBytes -= r->mem.Capacity();
    delete r;
}

static void FillStats(const mrb_renderer_t& r, const unsigned long long* h, mrb_render_stats& out)
{
    out.pathsStarted = r.startedBefore + (h[0] < r.d.pathLimit ? h[0] : r.d.pathLimit);
    out.pathsCompleted = h[1]; out.closestRays = h[2]; out.shadowRays = h[3]; out.neeSamples = h[4];
    out.iterations = r.iterations;
    out.finished = (h[1] >= r.completedTarget) ? 1u : 0u;
}

void RendererStats(Context& ctx, mrb_renderer_t& r, mrb_render_stats& out)
{
    unsigned long long h[8];
    MRB_CUDA_TRY(cudaMemcpyAsync(h, r.d.counters, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    FillStats(r, h, out);
}

// Begins the pass "samples [sampleStart, sampleStart + sampleCount) of every pixel of the region": the tile loop of
// PathTracerRendererT::DoLatencyRender (TracerDLL/PathTracerRenderer.cu:L1078-1160) + ImageTiler::NextTile
// (Tracer/RenderImage.cpp:L20-136). Asynchronous: the claim counter is reset in stream order; the caller must have seen
// the previous pass finish (run_pass / get_stats), because paths of two passes cannot share the slots of one region.
void RendererBeginPass(Context& ctx, mrb_renderer_t& r, const uint32_t regionMin[2], const uint32_t regionSize[2],
                       uint32_t sampleStart, uint32_t sampleCount)
{
    RenderData& d = r.d;
    if(regionSize[0] == 0 || regionSize[1] == 0 || regionSize[0] > r.maxWidth || regionSize[1] > r.maxHeight ||
       uint64_t(regionSize[0]) * regionSize[1] > uint64_t(r.maxWidth) * r.maxHeight)
        throw std::runtime_error("pass region exceeds the renderer's tile");
    if(regionMin[0] + regionSize[0] > d.fullWidth || regionMin[1] + regionSize[1] > d.fullHeight)
        throw std::runtime_error("pass region exceeds the image");
    r.startedBefore += d.pathLimit;
    d.regionX = regionMin[0]; d.regionY = regionMin[1]; d.width = regionSize[0]; d.height = regionSize[1];
    d.sampleBase = sampleStart;
    d.pathLimit = uint64_t(sampleCount) * regionSize[0] * regionSize[1];
    r.completedTarget += d.pathLimit;
    MRB_CUDA_TRY(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long), ctx.stream));
    r.needReload = true;
}

// Iterates until the current pass has finished, without draining the stream between polls: `chunk` iterations are
// always in flight while the counters of the previous chunk travel to pinned host memory (the reference reads its dead
// path count on the host after every single iteration, PathTracerRenderer.cu:L1046-1061). The iterations queued
// behind the one that finished the pass find no live path and cost ~0.1 ms each at 1080p.
void RendererRunPass(Context& ctx, mrb_renderer_t& r, uint32_t chunk, mrb_render_stats& out)
{
    if(chunk == 0) chunk = 4;
    int k = 0; bool havePrev = false;
    for(;;)
    {
        RenderIterate(ctx, r, chunk);
        MRB_CUDA_TRY(cudaMemcpyAsync(r.hCounters + 8 * k, r.d.counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, ctx.stream));
        MRB_CUDA_TRY(cudaEventRecord(r.evPoll[k], ctx.stream));
        if(havePrev)
        {
            MRB_CUDA_TRY(cudaEventSynchronize(r.evPoll[k ^ 1]));
            if(r.hCounters[8 * (k ^ 1) + 1] >= r.completedTarget) break;
        }
        havePrev = true; k ^= 1;
    }
    // the chunk still in flight completes no path; its counters equal the ones just read except for the claim counter
    FillStats(r, r.hCounters + 8 * (k ^ 1), out);
    out.iterations = r.iterations;
    memcpy(r.lastCounters, r.hCounters + 8 * (k ^ 1), sizeof(r.lastCounters));
    r.pollPending = false;
}

// Non-blocking counters: returns the latest snapshot that has reached pinned host memory and queues the next one behind
// the work issued so far, so a caller that iterates once per call (throughput mode) learns about the end of the render
// an iteration or two late instead of draining the stream every bounce.
void RendererPollStats(Context& ctx, mrb_renderer_t& r, mrb_render_stats& out)
{
    if(r.pollPending && cudaEventQuery(r.evPoll[0]) == cudaSuccess)
    {
        memcpy(r.lastCounters, r.hCounters, sizeof(r.lastCounters));
        r.pollPending = false;
    }
    cudaGetLastError();   // cudaErrorNotReady is not an error
    if(!r.pollPending)
    {
        MRB_CUDA_TRY(cudaMemcpyAsync(r.hCounters, r.d.counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, ctx.stream));
        MRB_CUDA_TRY(cudaEventRecord(r.evPoll[0], ctx.stream));
        r.pollPending = true;
    }
    FillStats(r, r.lastCounters, out);
}

void RendererReadFilm(Context& ctx, mrb_renderer_t& r, float* out, bool device, bool clear)
{
    size_t bytes = sizeof(float) * 4 * size_t(r.d.width) * r.d.height;
    MRB_CUDA_TRY(cudaMemcpyAsync(out, r.d.film, bytes, device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx.stream));
    if(clear) MRB_CUDA_TRY(cudaMemsetAsync(r.d.film, 0, bytes, ctx.stream));
    if(!device) MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
}

// RenderImage::TransferToHost (Tracer/RenderImage.cpp:L163-219): the film accumulated since the last hand-off travels
// to (pinned) host memory on a copy stream and is cleared there, `onComplete` runs when the copy has landed (the
// reference issues its semaphore release the same way), and rendering continues at once into the other film buffer.
void RendererFilmHandoff(Context& ctx, mrb_renderer_t& r, float* hostDst, void (*onComplete)(void*), void* user)
{
    const size_t bytes = sizeof(float) * 4 * size_t(r.d.width) * r.d.height;
    const int c = r.cur;
    MRB_CUDA_TRY(cudaEventRecord(r.evCompute, ctx.stream));
    MRB_CUDA_TRY(cudaStreamWaitEvent(r.copyStream, r.evCompute, 0));
    MRB_CUDA_TRY(cudaMemcpyAsync(hostDst, r.film[c], bytes, cudaMemcpyDeviceToHost, r.copyStream));
    MRB_CUDA_TRY(cudaMemsetAsync(r.film[c], 0, bytes, r.copyStream));
    MRB_CUDA_TRY(cudaEventRecord(r.evCopied[c], r.copyStream));
    r.copied[c] = true;
    if(onComplete) MRB_CUDA_TRY(cudaLaunchHostFunc(r.copyStream, onComplete, user));
    // continue into the other buffer once ITS previous hand-off (two calls ago) has been copied and cleared
    r.cur = c ^ 1;
    if(r.copied[r.cur]) MRB_CUDA_TRY(cudaStreamWaitEvent(ctx.stream, r.evCopied[r.cur], 0));
    r.d.film = r.film[r.cur];
}

float* RendererFilmPtr(mrb_renderer_t& r) { return r.d.film; }

// Sums the films of `peers` (renderers of OTHER devices with the same tile) into r's film over peer memory and clears
// them: the film reduction of SURVEY.md §8e fused with the "clear tile film" step. Each peer's stream is joined
// through an event; the kernel runs on r's device and reads the peers' HBM over NVLink.
__global__ void KReduceFilms(float4* __restrict__ dst, const float4* const* __restrict__ srcs, uint32_t nSrc, size_t count4)
{
    for(size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count4; i += size_t(gridDim.x) * blockDim.x)
    {
        float4 acc = dst[i];
        for(uint32_t s = 0; s < nSrc; s++)
        {
            float4* sp = const_cast<float4*>(srcs[s]) + i;
            const float4 v = *sp;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            *sp = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        dst[i] = acc;
    }
}

void RendererReducePeers(Context& ctx, mrb_renderer_t& r, Context* const* peerCtx, mrb_renderer_t* const* peers, uint32_t n)
{
    if(n == 0) return;
    if(n > 15) throw std::runtime_error("too many peer renderers");
    const size_t floats = size_t(4) * r.d.width * r.d.height;
    if(floats % 4) throw std::runtime_error("film size must be a multiple of 4 floats");
    std::vector<const float4*> h(n);
    for(uint32_t k = 0; k < n; k++)
    {
        if(peers[k]->d.width != r.d.width || peers[k]->d.height != r.d.height) throw std::runtime_error("peer renderer has a different tile");
        if(peerCtx[k]->device != ctx.device)
        {
            int can = 0;
            MRB_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx.device, peerCtx[k]->device));
            if(!can) throw std::runtime_error("no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(peerCtx[k]->device, 0);
            if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) throw CudaError{e, __FILE__, __LINE__};
            cudaGetLastError();
        }
        // join the peer's stream: record on its device, wait on ours
        MRB_CUDA_TRY(cudaSetDevice(peerCtx[k]->device));
        MRB_CUDA_TRY(cudaEventRecord(peers[k]->evCompute, peerCtx[k]->stream));
        MRB_CUDA_TRY(cudaSetDevice(ctx.device));
        MRB_CUDA_TRY(cudaStreamWaitEvent(ctx.stream, peers[k]->evCompute, 0));
        h[k] = reinterpret_cast<const float4*>(peers[k]->d.film);
    }
    ctx.scratch.Reserve(sizeof(float4*) * 16);
    const float4** dPtrs = static_cast<const float4**>(ctx.scratch.Base());
    MRB_CUDA_TRY(cudaMemcpyAsync(dPtrs, h.data(), sizeof(float4*) * n, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));   // h is a stack vector
    const size_t count4 = floats / 4;
    MRB_LAUNCH(ctx, KReduceFilms, GridFor(ctx, uint32_t(count4), 256, 8), 256, 0, reinterpret_cast<float4*>(r.d.film), dPtrs, n, count4);
    // the peers may only go on once their films have been read and cleared
    MRB_CUDA_TRY(cudaEventRecord(r.evCompute, ctx.stream));
    for(uint32_t k = 0; k < n; k++)
    {
        MRB_CUDA_TRY(cudaSetDevice(peerCtx[k]->device));
        MRB_CUDA_TRY(cudaStreamWaitEvent(peerCtx[k]->stream, r.evCompute, 0));
    }
    MRB_CUDA_TRY(cudaSetDevice(ctx.device));
}

// Latency mode (PathTracerRendererT::DoLatencyRender, TracerDLL/PathTracerRenderer.cu:L1078-1160) on top of passes:
// raising the limit from a to b begins the pass "samples [a, b) of every pixel".
void RendererSetSppLimit(Context& ctx, mrb_renderer_t& r, uint32_t sppLimit)
{
    if(sppLimit > r.totalSPP) throw std::runtime_error("spp limit exceeds totalSPP");
    unsigned long long h[2];
    MRB_CUDA_TRY(cudaMemcpyAsync(h, r.d.counters, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    const unsigned long long started = r.startedBefore + (h[0] < r.d.pathLimit ? h[0] : r.d.pathLimit);
    const uint32_t regionMin[2] = {r.d.regionX, r.d.regionY}, regionSize[2] = {r.d.width, r.d.height};
    if(started == 0 && h[1] == 0)
    {
        // nothing has been started yet: the initial all-samples pass shrinks to [0, sppLimit)
        r.completedTarget = 0; r.d.pathLimit = 0;
        RendererBeginPass(ctx, r, regionMin, regionSize, r.sampleOffset, sppLimit);
        r.sppStarted = sppLimit;
        return;
    }
    if(sppLimit < r.sppStarted) throw std::runtime_error("spp limit below the samples already started");
    if(h[1] < r.completedTarget) throw std::runtime_error("the current pass has not finished");
    RendererBeginPass(ctx, r, regionMin, regionSize, r.sampleOffset + r.sppStarted, sppLimit - r.sppStarted);
    r.sppStarted = sppLimit;
}

// mrb_filter_sample: the film filter on its own (parity tap of FilterSample; Tests/Tracer/T_Filters.cu)
__global__ void KFilterSample(uint32_t type, float radius, const float2* __restrict__ xi, uint32_t n, float4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float ox, oy, pdf, eval;
    FilterSample(type, radius, xi[i].x, xi[i].y, ox, oy, pdf, eval);
    out[i] = make_float4(ox, oy, pdf, eval);
}

void FilterSampleHost(Context& ctx, uint32_t type, float radius, const float* xi, uint32_t n, float* out)
{
    if(type > FILTER_MITCHELL) throw std::runtime_error("unknown film filter type");
    if(!(radius > 0.0f)) throw std::runtime_error("film filter radius must be positive");
    MultiAlloc sz(nullptr); sz.Take<float2>(n); sz.Take<float4>(n);
    ctx.scratch.Reserve(sz.Total());
    MultiAlloc ma(ctx.scratch.Base());
    float2* dXi = ma.Take<float2>(n); float4* dOut = ma.Take<float4>(n);
    MRB_CUDA_TRY(cudaMemcpyAsync(dXi, xi, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx.stream));
    if(n) MRB_LAUNCH(ctx, KFilterSample, DivUp(n, 256u), 256, 0, type, radius, dXi, n, dOut);
    MRB_CUDA_TRY(cudaMemcpyAsync(out, dOut, sizeof(float4) * n, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
}

// mrb_texture_sample: the shading kernel's texture filter on its own (parity tap of SampleTexture)
__global__ void KSampleTexture(TexRec t, const float2* __restrict__ uv, uint32_t n, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const Float3 c = SampleTexture(t, uv[i].x, uv[i].y);
    out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

static void ValidateTapTexture(const mrb_texture_desc& td)
{
    if(!td.data || td.width == 0 || td.height == 0 || (td.channels != 3 && td.channels != 4) || td.format > 1u || td.interp > 1u || td.edge > 2u)
        throw std::runtime_error("bad texture descriptor");
    ValidateMips(td);
}
// the texture as the renderer would hold it, in the context's scratch block (followed by `extra` bytes for the caller)
static TexRec StageTexture(Context& ctx, const mrb_texture_desc& td, size_t extra, char** extraOut)
{
    ValidateTapTexture(td);
    const size_t texBytes = ChainTexels(FinalWidth(td), FinalHeight(td), FinalMips(td)) * td.channels * (td.format == 0u ? 4u : 1u);
    MultiAlloc sz(nullptr); sz.Take<char>(texBytes); sz.Take<char>(extra);
    ctx.scratch.Reserve(sz.Total());
    MultiAlloc ma(ctx.scratch.Base());
    char* dTex = ma.Take<char>(texBytes); char* dExtra = ma.Take<char>(extra);
    if(extraOut) *extraOut = dExtra;
    TexRec t{dTex, FinalWidth(td), FinalHeight(td), td.channels, td.format, td.interp, td.edge, 0u};
    UploadTexture(ctx, t, td);
    return t;
}
size_t TextureChainTexels(uint32_t w, uint32_t h, uint32_t mips) { return ChainTexels(w, h, mips); }
uint32_t TextureFullMipCount(uint32_t w, uint32_t h) { return FullMipCount(w, h); }
void TextureFinalExtent(const mrb_texture_desc& td, uint32_t out[3]) { out[0] = FinalWidth(td); out[1] = FinalHeight(td); out[2] = FinalMips(td); }

void TextureSampleHost(Context& ctx, const mrb_texture_desc& td, const float* uv, uint32_t n, float* rgbOut)
{
    char* extra = nullptr;
    const TexRec t = StageTexture(ctx, td, sizeof(float2) * n + 256 + sizeof(float) * 3 * size_t(n), &extra);
    float2* dUV = reinterpret_cast<float2*>(extra); float* dOut = reinterpret_cast<float*>(extra + ((sizeof(float2) * n + 255) / 256) * 256);
    MRB_CUDA_TRY(cudaMemcpyAsync(dUV, uv, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx.stream));
    if(n) MRB_LAUNCH(ctx, KSampleTexture, DivUp(n, 256u), 256, 0, t, dUV, n, dOut);
    MRB_CUDA_TRY(cudaMemcpyAsync(rgbOut, dOut, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
}

// mrb_texture_sample_lod: SampleTextureLod / SampleTextureGrad on their own
__global__ void KSampleTextureLod(TexRec t, const float2* __restrict__ uv, const float* __restrict__ lod, const float4* __restrict__ grads,
                                  uint32_t lodMode, uint32_t n, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const Float3 c = lod ? SampleTextureLod(t, uv[i].x, uv[i].y, lod[i])
                         : SampleTextureGrad(t, uv[i].x, uv[i].y, make_float2(grads[i].x, grads[i].y), make_float2(grads[i].z, grads[i].w), lodMode);
    out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}
void TextureSampleLodHost(Context& ctx, const mrb_texture_desc& td, const float* uv, const float* lod, const float* grads, uint32_t lodMode, uint32_t n, float* rgbOut)
{
    char* extra = nullptr;
    const size_t slot = ((sizeof(float4) * size_t(n) + 255) / 256) * 256;
    const TexRec t = StageTexture(ctx, td, 4 * slot + 256, &extra);
    float2* dUV = reinterpret_cast<float2*>(extra); float* dLod = reinterpret_cast<float*>(extra + slot);
    float4* dGrad = reinterpret_cast<float4*>(extra + 2 * slot); float* dOut = reinterpret_cast<float*>(extra + 3 * slot);
    MRB_CUDA_TRY(cudaMemcpyAsync(dUV, uv, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx.stream));
    if(lod) MRB_CUDA_TRY(cudaMemcpyAsync(dLod, lod, sizeof(float) * n, cudaMemcpyHostToDevice, ctx.stream));
    else MRB_CUDA_TRY(cudaMemcpyAsync(dGrad, grads, sizeof(float4) * n, cudaMemcpyHostToDevice, ctx.stream));
    if(n) MRB_LAUNCH(ctx, KSampleTextureLod, DivUp(n, 256u), 256, 0, t, dUV, lod ? dLod : nullptr, lod ? nullptr : dGrad, lodMode, n, dOut);
    MRB_CUDA_TRY(cudaMemcpyAsync(rgbOut, dOut, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
}

void TextureMipChainHost(Context& ctx, const mrb_texture_desc& td, void* chainOut, uint32_t* mipCountOut)
{
    const TexRec t = StageTexture(ctx, td, 0, nullptr);
    const size_t texBytes = ChainTexels(t.w, t.h, t.mipCount) * t.channels * (t.format == 0u ? 4u : 1u);
    MRB_CUDA_TRY(cudaMemcpyAsync(chainOut, t.data, texBytes, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    if(mipCountOut) *mipCountOut = t.mipCount;
}

void TextureConvertHost(Context& ctx, const mrb_texture_desc& td, void* texelsOut)
{   // level 0 only: the supplied levels after TextureMemory::ConvertColorspaces
    const TexRec t = StageTexture(ctx, td, 0, nullptr);
    const size_t texBytes = size_t(t.w) * t.h * t.channels * (t.format == 0u ? 4u : 1u);
    MRB_CUDA_TRY(cudaMemcpyAsync(texelsOut, t.data, texBytes, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
}

} // namespace mrb
