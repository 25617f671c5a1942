/* mray_b200.h — C-ABI of the B200-native (sm_100a) MRay hot path.
 *
 * This is the boundary a maintainer of yalcinerbora/mray binds behind the TracerDLL plugin
 * (class Tracer : TracerBase implementing Core/TracerI.h:L150-376): the C++20 host object keeps
 * the scene/group registries and calls these entry points where the reference calls its
 * accelerator / renderer device code. Plain pointers and sizes only; no C++ or torch types.
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/Source). See INTEGRATION.md for the reference-side binding.
 *
 * Conventions
 *   - every function returns MRB_OK (0) or a negative mrb_status; mrb_last_error() gives text.
 *   - "memspace" says whether the array pointers of that call are host or device pointers.
 *     Host variants copy in/out on the context stream and synchronise before returning.
 *   - all device work is issued on the context stream (mrb_context_set_stream); device-pointer
 *     calls are asynchronous with respect to the host unless stated otherwise.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     MRB_ERR_NO_DEVICE.
 *
 * Binary layouts are the reference's own (Tracer/TracerTypes.h):
 *   RayGMem      32 B  { float pos[3]; float tMin; float dir[3]; float tMax; }      L276-283
 *   HitKeyPack   16 B  { u32 primKey; u32 lightOrMatKey; u32 transKey; u32 accelKey } L201-209
 *   MetaHit       8 B  { float a, b }  (triangle barycentrics (1-u-v, u))            Hit.h
 *   PrimitiveKey  u32  batch:4 | index:28     AcceleratorKey u32 batch:12 | index:20  L164-175
 */
#ifndef MRAY_B200_H
#define MRAY_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MRB_API __declspec(dllexport)
#else
#define MRB_API __attribute__((visibility("default")))
#endif

typedef enum mrb_status
{
    MRB_OK = 0,
    MRB_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path refuses to run */
    MRB_ERR_INVALID_ARG = -2,
    MRB_ERR_CUDA = -3,
    MRB_ERR_OUT_OF_MEMORY = -4,
    MRB_ERR_UNSUPPORTED = -5
} mrb_status;

typedef enum mrb_memspace { MRB_MEM_HOST = 0, MRB_MEM_DEVICE = 1 } mrb_memspace;

typedef struct mrb_context_t* mrb_context;
typedef struct mrb_accel_t* mrb_accel;

typedef struct mrb_ray_gmem { float pos[3]; float tMin; float dir[3]; float tMax; } mrb_ray_gmem;
typedef struct mrb_hit_key_pack { uint32_t primKey, lightOrMatKey, transKey, accelKey; } mrb_hit_key_pack;
typedef struct mrb_meta_hit { float a, b; } mrb_meta_hit;

#define MRB_INVALID_KEY 0xFFFFFFFFu

/* ---- library / context ------------------------------------------------------------------ */

/* ABI version of this header (major<<16 | minor); the descriptor struct layouts are part of it. */
#define MRB_ABI_VERSION ((0u << 16) | 8u)
MRB_API uint32_t mrb_abi_version(void);

/* Replaces GPUSystem + GPUQueue ownership inside TracerBase (Device/CUDA/GPUSystemCUDA.cpp:L249-405:
 * the reference always renders on BestDevice() with one queue). One context = one device, one
 * stream, one device arena (the reference's DeviceMemory + MemAlloc::AllocateMultiData,
 * Core/MemAlloc.h:L171-209: 256-byte aligned sub-allocation of a growing arena). */
MRB_API mrb_status mrb_context_create(int device, mrb_context* out);
MRB_API void       mrb_context_destroy(mrb_context ctx);
/* Issue all work on an externally owned cudaStream_t (e.g. torch's current stream). NULL is the
 * CUDA default stream. Until this is called the context uses a private non-blocking stream. */
MRB_API mrb_status mrb_context_set_stream(mrb_context ctx, void* cuda_stream);
MRB_API mrb_status mrb_context_synchronize(mrb_context ctx);
/* TracerI::UsedDeviceMemory / TotalDeviceMemory (Core/TracerI.h:L372-373). */
MRB_API size_t     mrb_context_used_device_memory(mrb_context ctx);
MRB_API size_t     mrb_context_total_device_memory(mrb_context ctx);
/* Number of kernels this library has launched on the context since creation (for the
 * gpu_launches accounting of bench.py). */
MRB_API uint64_t   mrb_context_launch_count(mrb_context ctx);
/* Rays of the last MRB_TRACE_WIDE cast that could not be certified against the reference's box
 * arithmetic by the traversal kernel itself (synchronises). out[4]:
 * [0] total, [1] of which near-tie candidates, [2] of which winner's leaf AABB not certified,
 * [3] rays that needed the full exact binary re-traversal (the others were settled by replaying the
 *     reference's box tests on the candidates' ancestor chains). */
MRB_API mrb_status mrb_context_last_fallback_count(mrb_context ctx, uint32_t* out);
MRB_API const char* mrb_last_error(mrb_context ctx); /* ctx may be NULL: last create error */
/* Sampled kernel timing for measurement (bench.py's roofline): with profiling on, every iterationStride-th wavefront
 * iteration of a renderer records CUDA-event pairs on the context stream around its kernels; mrb_context_get_profile
 * synchronises and returns the accumulated times / sample counts per kind:
 * [0] wide closest-hit traversal kernel, [1] shading kernel, [2] wide any-hit traversal kernel, [3] finish + reload kernel,
 * [4] the exact-resolution tail of a cast (KResolveExact + KTraceBinary). The reference's counterpart is its NVTX ranges. */
typedef struct mrb_kernel_profile { double ms[5]; uint64_t samples[5]; } mrb_kernel_profile;
/* Seed of the stochastic alpha test of the NEXT casts on this context (see mrb_accel_desc.rangeAlphaMap): every cast
 * hashes (seed, ray index, leaf) and then advances the context's seed by one, so consecutive casts decorrelate and a
 * caller that sets the seed replays the same decisions. Renderers derive it from TracerParameters.seed and their
 * iteration counter. */
MRB_API mrb_status mrb_context_set_alpha_seed(mrb_context ctx, uint32_t seed);
MRB_API mrb_status mrb_context_set_profiling(mrb_context ctx, int enabled, uint32_t iterationStride);
MRB_API mrb_status mrb_context_get_profile(mrb_context ctx, mrb_kernel_profile* out);

/* ---- accelerator build ------------------------------------------------------------------ */

typedef enum mrb_build_flags
{
    /* Default: Karras delta with the augmented key (64 + clz(i^j)) on EQUAL Morton codes. This is
     * bit-identical to the reference whenever all codes of an accelerator are distinct
     * (mrb_accel_info.duplicateCodes == 0) and stays a valid tree otherwise. */
    MRB_BUILD_DEFAULT = 0,
    /* Audit mode: the reference's literal equal-code fallback, which compares the raw indices
     * without the +64 offset (AcceleratorLBVH.cu:L85-90) and can emit an ill-formed hierarchy when
     * codes repeat. Implies MRB_BUILD_BINARY_ONLY; the topology arrays still match the reference
     * node for node. */
    MRB_BUILD_REFERENCE_DELTA = 1,
    /* keep only the binary LBVH (skip the wide-BVH collapse) */
    MRB_BUILD_BINARY_ONLY = 2,
    /* Audit mode: collapse with the one-thread-per-node kernel instead of the eight-lanes-per-node one. Both build the same
     * tree (same nodes and records; only the order of allocations inside a level differs). */
    MRB_BUILD_SERIAL_COLLAPSE = 4
} mrb_build_flags;

/* One concrete accelerator over a triangle primitive group.
 * Replaces AcceleratorGroupLBVH<PrimGroupTriangle>::Construct + MultiBuildLBVH
 * (Tracer/AcceleratorLBVH.hpp:L433-899) for one concrete accelerator:
 *   KCGeneratePrimitiveKeys (AcceleratorLinear.cu:L14)  leaf i -> PrimitiveKey
 *   KCGeneratePrimAABBs / KCGenPrimCenters (AcceleratorWork.kt.h:L5-123)
 *   SegmentedTransformReduce(UnionAABB3) (hpp:L764-769)
 *   KCGenMortonCode (AcceleratorLBVH.cu:L96-168), SegmentedRadixSort<true,u64,u32> (hpp:L828-838)
 *   KCConstructLBVHInternalNodes (cu:L170-299), KCUnionLBVHBoundingBoxes (cu:L301-435)
 * then (unless MRB_BUILD_BINARY_ONLY) collapses the bit-exact binary tree into the 8-wide
 * quantised BVH the traversal kernels use. */
typedef struct mrb_accel_desc
{
    const float*    positions;    /* vertexCount * 3 floats (PrimGroupTriangle positions)          */
    uint32_t        vertexCount;
    const uint32_t* indices;      /* triangleCount * 3, already rebased to the group's vertex list */
    uint32_t        triangleCount;/* triangles in the GROUP index list                             */
    mrb_memspace    memspace;     /* of positions / indices                                        */
    uint32_t        primGroupId;  /* batch portion of the PrimitiveKeys (4 bits)                   */
    /* prim ranges of this accelerator inside the group; leaves are numbered range by range. One
     * reference surface has <= 8 (TracerConstants::MaxPrimBatchPerSurface); surfaces sharing the
     * identity transform may be flattened into one accelerator, so any count is accepted.
     * NULL/0 = one range covering the whole group. */
    uint32_t        rangeCount;
    const uint32_t* primRanges;   /* rangeCount * 2 : [begin, end) ; host memory                   */
    const uint32_t* lightOrMatKeys;/* rangeCount ; host ; NULL = 0                                 */
    const uint8_t*  cullBackface; /* rangeCount ; host ; NULL = 0 (two sided)                      */
    uint32_t        flags;        /* mrb_build_flags                                               */
    /* Alpha maps (SurfaceParams.alphaMaps, Core/TracerI.h; AcceleratorLBVH::IntersectionCheck,
     * Tracer/AcceleratorLBVH.hpp:L263-282): rangeAlphaMap[r] = -1, or an index into alphaTextures. A triangle hit of
     * such a range is kept with probability alpha(uv) — uv = the hit's interpolated UV0, alpha = channel 0 of the
     * texture — both in closest-hit and in visibility casts ("stochastic alpha culling"). The reference draws the
     * deciding number from its per-ray backup PCG32; here it is a hash of (cast seed, ray index, leaf), so one
     * (ray, triangle) pair gets one decision however often and in whatever order the traversal meets it.
     * alphaTextures: single-mip textures, 1..4 channels of fp32 / unorm8 (host data, mrb_texture_desc is declared
     * below); vertexUVs: vertexCount * 2 floats in `memspace`. All NULL / 0 = no alpha maps. */
    const float*    vertexUVs;
    uint32_t        alphaTextureCount;
    const struct mrb_texture_desc* alphaTextures;
    const int32_t*  rangeAlphaMap; /* rangeCount ; host ; NULL = none */
} mrb_accel_desc;

MRB_API mrb_status mrb_accel_build(mrb_context ctx, const mrb_accel_desc* desc, mrb_accel* out);
MRB_API void       mrb_accel_destroy(mrb_context ctx, mrb_accel accel);

typedef struct mrb_accel_info
{
    uint32_t leafCount;      /* triangles                                   */
    uint32_t nodeCount;      /* binary LBVH nodes = max(1, leafCount-1)     */
    uint32_t wideNodeCount;  /* 8-wide nodes (0 when BINARY_ONLY)           */
    uint32_t duplicateCodes; /* 1 if the Morton codes were not all distinct */
    float    aabb[6];        /* accelerator AABB (min xyz, max xyz)         */
    float    buildMs;        /* device time of the last build (CUDA events) */
    size_t   deviceBytes;
} mrb_accel_info;
MRB_API mrb_status mrb_accel_get_info(mrb_context ctx, mrb_accel accel, mrb_accel_info* info);

/* Parity tap: copies the binary LBVH artefacts to HOST arrays in the reference's layout
 * (any pointer may be NULL): morton[leaf] in leaf order, sortedMorton/sortedLeaf[leaf] after the
 * stable sort, nodes[node*3] = LBVHNode{left,right,parent} with MSB = leaf flag and ORIGINAL leaf
 * index payload (AcceleratorLBVH.h:L62-67), leafParent[leaf], nodeBoxes[node*6], leafAABBs[leaf*6]. */
MRB_API mrb_status mrb_accel_export_lbvh(mrb_context ctx, mrb_accel accel,
                                         uint64_t* morton, uint64_t* sortedMorton, uint32_t* sortedLeaf,
                                         uint32_t* nodes, uint32_t* leafParent,
                                         float* nodeBoxes, float* leafAABBs);
/* Audit export of the product traversal structure (no reference counterpart: the reference traverses the binary tree):
 * wideNodes = mrb_accel_info.wideNodeCount x 80-byte nodes, triRecords = leafCount x 48-byte records (either may be NULL);
 * layouts in mray_b200/csrc/accel.cuh. Node ORDER inside a level depends on the order of the builder's allocations. */
MRB_API mrb_status mrb_accel_export_wide(mrb_context ctx, mrb_accel accel, void* wideNodes, void* triRecords);

/* ---- ray casting ------------------------------------------------------------------------ */

typedef enum mrb_trace_mode
{
    MRB_TRACE_WIDE = 0,        /* product path: 8-wide quantised BVH                                   */
    MRB_TRACE_BINARY_EXACT = 1,/* audit path: the reference's binary LBVH, left-first, exact slab test */
    /* OR-ed flag: the caller does not care what hitKeys / metaHits / isVisibleBits hold on entry. Rays that miss
     * then get {INVALID, INVALID, INVALID, INVALID} keys, zero barycentrics and a set visibility bit instead of
     * keeping the caller's values, so a host-pointer call uploads the rays only (keys and hits are 43 % of the
     * bytes a closest-hit call sends over PCIe). */
    MRB_TRACE_FRESH_OUTPUTS = 0x100
} mrb_trace_mode;

/* Closest hit. Replaces BaseAcceleratorLBVH::CastRays (Tracer/AcceleratorLBVH.cu:L760-896) +
 * KCLocalRayCast (AcceleratorWork.kt.h:L177-244) + AcceleratorLBVH::ClosestHit
 * (AcceleratorLBVH.hpp:L327-374) for an identity-transform triangle accelerator.
 * For i in [0,rayCount): r = rayIndices ? rayIndices[i] : i ; on a hit writes hitKeys[r]
 * {primKey, lightOrMatKey of the prim range, transKey = 0, accelKey}, metaHits[r] and shrinks
 * rays[r].tMax ; on a miss leaves all three untouched (the caller pre-fills boundary keys,
 * RendererCommon.cu:L68-81). */
MRB_API mrb_status mrb_cast_rays(mrb_context ctx, mrb_accel accel,
                                 mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits,
                                 mrb_ray_gmem* rays, const uint32_t* rayIndices,
                                 uint32_t rayCount, uint32_t totalRayCount,
                                 mrb_memspace memspace, mrb_trace_mode mode);

/* Any hit. Replaces BaseAcceleratorLBVH::CastVisibilityRays (AcceleratorLBVH.cu:L898-1035) +
 * KCVisibilityRayCast (AcceleratorWork.kt.h:L246-296): clears bit r of isVisible (u32 words,
 * Bitspan, Tracer/Bitspan.h:L108) when ray r hits anything in [tMin,tMax); never sets bits. */
MRB_API mrb_status mrb_cast_visibility_rays(mrb_context ctx, mrb_accel accel,
                                            uint32_t* isVisibleBits,
                                            const mrb_ray_gmem* rays, const uint32_t* rayIndices,
                                            uint32_t rayCount, uint32_t totalRayCount,
                                            mrb_memspace memspace, mrb_trace_mode mode);

/* ---- two-level scene ------------------------------------------------------------------- */

typedef struct mrb_scene_t* mrb_scene;

/* One accelerator instance: a concrete accelerator + the transform of its surface
 * (AcceleratorGroup::WriteInstanceKeysAndAABBsInternal, Tracer/AcceleratorCommon.cu:L452-572).
 * transform / invTransform: row-major 3x4 (Matrix3x4), local->world and world->local — both are
 * supplied by the host object, which owns the TransformGroup ((T)Single stores both,
 * Tracer/TransformsDefault.hpp). Rays are taken to local space WITHOUT renormalising the direction
 * (KCLocalRayCast, AcceleratorWork.kt.h:L207-218), so t stays in world units. */
typedef struct mrb_instance_desc
{
    mrb_accel accel;
    float     transform[12];
    float     invTransform[12];
    uint32_t  isIdentity;     /* (T)Identity: no ray transform, world AABB = accelerator AABB */
    uint32_t  transformKey;   /* reported in HitKeyPack.transKey */
    uint32_t  accelKey;       /* reported in HitKeyPack.accelKey (batch:12 | index:20) */
    /* Optional (host pointer, one entry per prim range of `accel`, NULL = the accelerator's own keys):
     * the reference keeps LightOrMatKeys per INSTANCE while surfaces with the same primitive batches and
     * cull flags share one concrete accelerator (AcceleratorGroupLBVH::PreConstruct ->
     * AcceleratorGroup partitions, Tracer/AcceleratorC.h:L780-905: `concreteIndicesOfInstances`,
     * `dAllLeafs` per concrete accelerator, `dMatKeys` per instance), so N instances of one mesh with N
     * different materials cost one BVH. Copied at build time. */
    const uint32_t* lightOrMatKeys;
} mrb_instance_desc;

/* BaseAcceleratorLBVH::InternalConstruct (Tracer/AcceleratorLBVH.cu:L537-740): top-level LBVH over the
 * instances' world AABBs (same Morton / sort / Karras / union chain) + its wide collapse. The
 * accelerators must outlive the scene. */
MRB_API mrb_status mrb_scene_build(mrb_context ctx, const mrb_instance_desc* instances, uint32_t instanceCount, mrb_scene* out);
MRB_API void       mrb_scene_destroy(mrb_context ctx, mrb_scene scene);
/* Parity tap (host arrays, any may be NULL): world AABBs [n*6], scene AABB [6], top-level Morton codes,
 * sorted instance order, binary nodes [max(1,n-1)*3], node boxes [..*6]. */
MRB_API mrb_status mrb_scene_export_tlas(mrb_context ctx, mrb_scene scene, float* instanceAABBs, float* sceneAABB,
                                         uint64_t* morton, uint32_t* sortedInstance, uint32_t* nodes, float* nodeBoxes);
/* BaseAcceleratorLBVH::CastRays / CastVisibilityRays over the whole scene in ONE launch (the reference
 * alternates top-level traversal, a ray sort by instance key, a host read-back and per-instance
 * launches until every ray has left the top-level tree, AcceleratorLBVH.cu:L806-895). Same buffer
 * contract as mrb_cast_rays; hitKeys carry the instance's transKey / accelKey. */
MRB_API mrb_status mrb_scene_cast_rays(mrb_context ctx, mrb_scene scene,
                                       mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits,
                                       mrb_ray_gmem* rays, const uint32_t* rayIndices,
                                       uint32_t rayCount, uint32_t totalRayCount,
                                       mrb_memspace memspace, mrb_trace_mode mode);
MRB_API mrb_status mrb_scene_cast_visibility_rays(mrb_context ctx, mrb_scene scene, uint32_t* isVisibleBits,
                                                  const mrb_ray_gmem* rays, const uint32_t* rayIndices,
                                                  uint32_t rayCount, uint32_t totalRayCount,
                                                  mrb_memspace memspace, mrb_trace_mode mode);

/* ---- hero-wavelength spectral transport ------------------------------------------------- */

typedef struct mrb_spectrum_t* mrb_spectrum;

typedef enum mrb_wavelength_sample_mode
{   /* WavelengthSampleMode (Core/TracerEnums.h:L133-150), TracerParameters.wavelengthSampleMode */
    MRB_WAVELENGTH_UNIFORM = 0, MRB_WAVELENGTH_GAUSSIAN_MIS = 1, MRB_WAVELENGTH_HYPERBOLIC_PBRT = 2
} mrb_wavelength_sample_mode;

/* What SpectrumContextJakob2019's constructor loads (Tracer/SpectrumContext.cu:L354-520), passed as
 * HOST data by the plugin: the payload of "SpectraLUT/<COLORSPACE>.mrspectra" (9 blocks of res^3 fp32:
 * for max-channel table t = 0..2 the coefficients c0, c1, c2; texel (x,y,z) at z*res*res + y*res + x),
 * the CIE-1931 observer divided by its X/Y/Z integrals, the colour space's standard illuminant SPD
 * scaled by CIE_1931_Y_INTEGRAL / SelectIlluminantSPDNormFactor, both 471 entries for 360..830 nm,
 * and the XYZ->RGB matrix (row-major). */
typedef struct mrb_spectrum_desc
{
    const float* lut;
    uint32_t     lutResolution;        /* must be 64 (Jakob2019Detail::Data::N) */
    const float* observerXYZ;          /* 471 * 3 */
    const float* illuminantSPD;        /* 471 */
    float        xyzToRGB[9];
    uint32_t     wavelengthSampleMode; /* mrb_wavelength_sample_mode */
} mrb_spectrum_desc;

MRB_API mrb_status mrb_spectrum_create(mrb_context ctx, const mrb_spectrum_desc* desc, mrb_spectrum* out);
MRB_API void       mrb_spectrum_destroy(mrb_context ctx, mrb_spectrum spectrum);
/* SpectrumContextJakob2019::SampleSpectrumWavelengths (SpectrumContext.cu:L14-135,L173-196): one random
 * number per element -> 4 wavelengths (nm) + 4 pdfs. waves / pdfs: count * 4 floats. */
MRB_API mrb_status mrb_spectrum_sample_wavelengths(mrb_context ctx, mrb_spectrum spectrum, float* waves, float* pdfs,
                                                   const uint32_t* randomNumbers, uint32_t count, mrb_memspace memspace);
/* SpectrumContextJakob2019::ConvertSpectrumToRGB (SpectrumContext.cu:L137-171,L225-254), in place:
 * values[i] (4 spectral samples) -> (r, g, b, 0). */
MRB_API mrb_status mrb_spectrum_convert_to_rgb(mrb_context ctx, mrb_spectrum spectrum, float* values, const float* waves,
                                               const float* pdfs, uint32_t count, mrb_memspace memspace);
/* Jakob2019Detail::Converter::ConvertAlbedo (isRadiance = 0) / ConvertRadiance (1) per element
 * (SpectrumContext.hpp:L37-150). rgb: count * 3 floats, or 3 floats shared by all when rgbIsUniform != 0. */
MRB_API mrb_status mrb_spectrum_upsample(mrb_context ctx, mrb_spectrum spectrum, float* outSpectra, const float* rgb,
                                         int rgbIsUniform, const float* waves, uint32_t count, int isRadiance,
                                         mrb_memspace memspace);

/* ---- samplers --------------------------------------------------------------------------- */

typedef enum mrb_sampler_type
{   /* SamplerType (Core/TracerEnums.h), TracerParameters.samplerType */
    MRB_SAMPLER_INDEPENDENT = 0, MRB_SAMPLER_SOBOL = 1, MRB_SAMPLER_ZSOBOL = 2,
    /* OR-ed flag: scramble exactly like the reference. SobolCommon::ScambleOwenFast (Tracer/Random.cu:L152-162)
     * hashes the bit-reversed value but never reverses it back, which leaves the stratified digits in the low
     * bits: its points are no better than random. Without this flag the final bit reversal is applied (a proper
     * Owen-scrambled net); with it the numbers equal the reference's bit for bit. */
    MRB_SAMPLER_REFERENCE_SCRAMBLE = 0x100
} mrb_sampler_type;

/* RNGGroupSobol / RNGGroupZSobol::GenerateNumbers (Tracer/Random.cu:L1017-1042,L1293-1318 ->
 * KCGenRandomNumbersGeneric + GenerateRNFromList, L439-555) for a width x height grid of generators
 * (generator i = pixel (i % width, i / width), LocalState.seed = generatorSeeds[i]) that all stand at
 * `sampleIndex`: for every request of requestDims[] (each 1, 2 or 3 = Next / Next2D / Next3D) starting at
 * dimension dimensionStart, writes the raw 32-bit numbers dimension-major: out[i + width*height * o].
 * sobolMatrices = SobolDetail::SobolMatrices (256 * 52 u32; ZSobol uses its first three rows).
 * requestDims is a HOST array (<= 32 entries); the other arrays live in `memspace`. Bit-exact. */
MRB_API mrb_status mrb_sampler_generate(mrb_context ctx, uint32_t samplerType, const uint32_t* sobolMatrices,
                                        const uint32_t* generatorSeeds, uint32_t width, uint32_t height,
                                        uint32_t sampleIndex, uint32_t initialMaxSPP, uint32_t dimensionStart,
                                        const uint32_t* requestDims, uint32_t requestCount, uint32_t* numbersOut,
                                        mrb_memspace memspace);

/* ---- wavefront path tracer ------------------------------------------------------------- */

typedef struct mrb_renderer_t* mrb_renderer;

typedef enum mrb_sample_mode
{   /* PathTraceRDetail::SampleModeEnum (TracerDLL/PathTracerRendererShaders.h:L21-35) */
    MRB_SAMPLE_PURE = 0, MRB_SAMPLE_NEE = 1, MRB_SAMPLE_NEE_WITH_MIS = 2
} mrb_sample_mode;

/* Everything PathTracerRendererT::StartRender needs (TracerDLL/PathTracerRenderer.cu:L1182-1355):
 * the committed accelerator, the material / light groups its lightOrMatKeys point into, the
 * camera surface, the render attributes ("totalSPP", "sampleMode", "rrRange" — L1384-1400) and
 * TracerParameters.seed / filmFilter. lightOrMatKey layout (Tracer/TracerTypes.h:L171-175):
 * bit 31 = light flag, bits 0..20 = index into albedo[] (material) or lightRadiance[] (light).
 * Scope: (R)PathTracerRGB / Spectral, (Mt)Lambert (constant or textured albedo), (Mt)Reflect, (L)Prim(P)Triangle with
 * constant radiance, (L)Null boundary, (C)Pinhole, the four film filters, Independent / Sobol / ZSobol samplers. */
typedef enum mrb_material_type
{   /* MatGroupLambert / Reflect / Refract / Unreal (Tracer/MaterialsDefault.h) */
    MRB_MATERIAL_LAMBERT = 0, MRB_MATERIAL_REFLECT = 1, MRB_MATERIAL_REFRACT = 2, MRB_MATERIAL_UNREAL = 3
} mrb_material_type;
/* FilterType::E (Core/TracerEnums.h:L162-173) */
typedef enum mrb_boundary_type { MRB_BOUNDARY_NULL = 0, MRB_BOUNDARY_SKYSPHERE_SPHERICAL = 1, MRB_BOUNDARY_SKYSPHERE_COOCTA = 2 } mrb_boundary_type;
typedef enum mrb_film_filter { MRB_FILTER_BOX = 0, MRB_FILTER_TENT = 1, MRB_FILTER_GAUSSIAN = 2, MRB_FILTER_MITCHELL_NETRAVALI = 3 } mrb_film_filter;

/* One 2-D texture of the renderer (first slice): what TracerI::CreateTexture2D +
 * PushTextureData + CommitTextures hand to TextureMemory (Tracer/TextureMemory.cpp): 3 / 4 channels of fp32 or unorm8, one or more mip levels, already in the global colour space (MRayTextureParameters.colorSpace
 * = MR_DEFAULT, gamma 1: TextureMemory::ConvertColorspaces skips such textures). Sampling restates the reference's
 * host-backend texture view (Device/CPU/TextureViewCPU.h:L172-470): normalised coordinates, texel centres at
 * +0.5, nearest or bilinear with unfused lerps, wrap / clamp / mirror edge resolve; with one mip level the
 * ray-cone gradient of TracerTexView::operator()(uv, dpdx, dpdy) clamps to level 0. */
typedef struct mrb_texture_desc
{   /* (channels: 3 or 4 for colour reads; an alpha map may also have 1 or 2 and only its first channel is read) */
    const void* data;        /* host; row-major, width*height texels, `channels` values each */
    uint32_t    width, height;
    uint32_t    channels;    /* 3 or 4 (a 4th channel is ignored by Vector3 reads) */
    uint32_t    format;      /* 0 = fp32, 1 = unorm8 (NormConversion::FromUNorm: v * (1/255)) */
    uint32_t    interp;      /* MRayTextureInterpEnum: 0 NEAREST, 1 LINEAR */
    uint32_t    edge;        /* MRayTextureEdgeResolveEnum: 0 WRAP, 1 CLAMP, 2 MIRROR */
    /* TextureMemory::ConvertColorspaces (Tracer/TextureMemory.cpp:L434-487) -> KCConvertColor (Tracer/ColorConverter.cu:L306-398),
     * applied to the texels once, on upload: rgb <- pow(rgb, gamma) (OpticalTransferGamma::ToLinear; 0 or 1 = no gamma), then
     * rgb <- colorMatrix . rgb (ColorspaceTransfer<from, global>::RGBToRGBMatrix = FromXYZ(global) ToXYZ(from), row-major 3x3;
     * NULL = the texture is already in the global colour space). unorm8 texels are de-normalised, converted and rounded back
     * (ToUNorm). Needs >= 3 channels. */
    float       gamma;
    const float* colorMatrix;
    /* Mip chain (TracerI::CreateTexture2D(size, mipCount, ...) + PushTextureData(id, mipLevel, ...)): `data` holds mipCount levels
     * (0 reads as 1) back to back, level k = max(width >> k, 1) x max(height >> k, 1) texels (Graphics::TextureMipSize,
     * Core/GraphicsFunctions.h:L474-490). generateMips != 0 = TracerParameters.genMips: the chain is completed down to 1 x 1
     * (TextureMemory::CreateTexture, Tracer/TextureMemory.cpp:L588-591) and the missing levels are filtered from their parents after the
     * colour conversion (TextureMemory::Finalize L809-833 -> KCGenerateMipmaps, Tracer/TextureFilter.cu:L126-198) with
     * TracerParameters.mipGenFilter = {mipFilterType (mrb_film_filter), mipFilterRadius}; the reference's default is Gaussian, 2.
     * A renderer with any multi-level texture carries a ray cone per path (Tracer/TracerTypes.h:L48-69) and reads albedo / normal
     * textures at the level its footprint selects (trilinear under MR_LINEAR). Alpha maps and skysphere radiance are read at
     * level 0, as the reference reads them (no gradients at those call sites). */
    uint32_t    mipCount;
    uint32_t    generateMips;
    uint32_t    mipFilterType;
    float       mipFilterRadius;
    /* TracerParameters.clampedTexRes (0 = none; pass 0 for MRayTextureParameters.ignoreResClamp): a texture whose larger side exceeds it
     * loses ceil(log2(ceil(maxDim / clamp))) levels (TextureMemory::CreateTexture, Tracer/TextureMemory.cpp:L544-583). With fewer levels
     * supplied than that, the last supplied level is filtered down to the new level 0 at upload (ClampImageFromBuffer -> KCClampImage,
     * Tracer/TextureFilter.cu:L206-262: 4 x 4 samples of the mip filter; mipFilterRadius 0 = the default Gaussian, 2); otherwise the
     * levels that fit are kept. Sizes reported by the taps are the clamped ones. */
    uint32_t    clampResolution;
} mrb_texture_desc;

typedef struct mrb_render_desc
{
    mrb_accel       accel;
    uint32_t        vertexCount, triangleCount;   /* of the primitive group the accelerator was built on */
    const float*    vertexNormals;   /* host, vertexCount*3 shading normals, or NULL (geometric normal) */
    uint32_t        materialCount;
    const float*    albedo;          /* host, materialCount*3 */
    uint32_t        lightCount;
    const float*    lightRadiance;   /* host, lightCount*3 */
    const uint8_t*  lightTwoSided;   /* host, lightCount, or NULL */
    float           camPosition[3], camGaze[3], camUp[3];
    float           fovXY[2];        /* radians ("FovAndPlanes") */
    float           nearFar[2];
    uint32_t        width, height;   /* RenderImageParams.resolution (one tile) */
    uint32_t        totalSPP;
    uint32_t        sampleMode;      /* mrb_sample_mode */
    uint32_t        rrRange[2];
    float           filmFilterRadius;/* TracerParameters.filmFilter.radius (default 1); the type is filmFilterType below */
    uint64_t        seed;            /* TracerParameters.seed */
    uint32_t        maxPathCount;    /* paths in flight; 0 = width*height (parallelizationHint tile) */
    uint32_t        partitionRays;   /* 1 = sort live rays by (work batch, material) key before shading
                                      * (RenderSurfaceWorkHasher + RayPartitioner::MultiPartition) */
    /* Two-level scenes: when `scene` is non-NULL it replaces `accel` (which must then be NULL) and
     * vertexCount / triangleCount / vertexNormals are ignored. Surfaces are generated in the instance's
     * local space and moved to world space with the instance transform (TransformContextSingle,
     * Tracer/TransformsDefault.h); emissive triangles are listed per instance in world space.
     * instanceVertexNormals: NULL, or one host pointer per instance (NULL entries = geometric normal),
     * each with as many normals as the instance's accelerator has vertices. */
    mrb_scene       scene;
    const float* const* instanceVertexNormals;
    /* NULL = (R)PathTracerRGB. Non-NULL = (R)PathTracerSpectral: albedo / lightRadiance are upsampled to
     * spectra over 4 hero wavelengths per path (sampled at path start from one extra random number),
     * transport runs on the 4 samples, the film receives ConvertSpectrumToRGB of each finished path
     * (Tracer/PathTracerRendererBase.cu:L139-168,L228-241). The spectrum object must outlive the renderer. */
    mrb_spectrum    spectrum;
    /* TracerParameters.samplerType: mrb_sampler_type. Sobol / ZSobol need sobolMatrices (host, 256 * 52 u32,
     * SobolDetail::SobolMatrices); one generator per PIXEL, seeded like RNGGroupSobol / RNGGroupZSobol
     * (Tracer/Random.cu:L884-1408): the k-th path of a pixel draws sample index k, dimensions advance per
     * request (filter 2-D, wavelength 1-D, then per bounce light 3-D, BxDF 2-D, roulette 1-D);
     * ZSobol's initialMaxSPP = totalSPP. */
    uint32_t        samplerType;
    const uint32_t* sobolMatrices;
    /* Textured Lambert albedo (ParamVaryingData<2, Vector3>, Tracer/ParamVaryingData.h + MaterialsDefault.hpp:L17-25):
     * albedoTexture: NULL, or host int32 per material, -1 = constant `albedo`, else an index into `textures`; the
     * texture is read at the hit's interpolated UV0 (uv0 a + uv1 b + uv2 c, PrimitiveDefaultTriangle.hpp:L472-476)
     * and, in the spectral renderer, upsampled per hit (Converter::ConvertAlbedo). vertexUVs: host, vertexCount*2
     * (single accelerator); instanceVertexUVs: one host pointer per instance (two-level scenes); missing = uv (0,0). */
    uint32_t        textureCount;
    const mrb_texture_desc* textures;
    const int32_t*  albedoTexture;
    const float*    vertexUVs;
    const float* const* instanceVertexUVs;
    /* RenderImageParams{resolution, regionMin, regionMax} (Core/TracerI.h:L38-43): width x height above is the REGION
     * rendered by this renderer; fullResolution (0,0 = the region is the whole image) and regionMin place it in the
     * image the camera spans, so several renderers (tiles, GPUs) can share one image. */
    uint32_t        fullResolution[2];
    uint32_t        regionMin[2];
    /* NULL (every material is (Mt)Lambert), or host u8 per material: mrb_material_type. (Mt)Reflect
     * (Tracer/MaterialsDefault.hpp:L132-215) and (Mt)Refract are perfectly specular: no NEE shadow ray, no Russian
     * roulette, the next ray is a SPECULAR_RAY (PathTracerRendererShaders.h:L245-262,L415-424); their `albedo` entries are
     * ignored. (Mt)Unreal behaves the same way once its specularity (1 - diffuse lobe probability) reaches 0.95. */
    const uint8_t*  materialType;
    /* TracerParameters.filmFilter.type (FilterType::E, Core/TracerEnums.h:L162-173; Tracer/Filters.h): the sampler of the
     * camera sample's sub-pixel offset and its film weight Evaluate / pdf. filmFilterRadius above is its radius. */
    uint32_t        filmFilterType;  /* mrb_film_filter; 0 = Box — set MRB_FILTER_GAUSSIAN for the reference's default */
    /* Sample-range sharding: this renderer produces samples [sampleOffset, sampleOffset + totalSPP) of
     * every pixel. Random numbers are a function of (seed, full-image pixel, sample index) only, so the ranges of several
     * renderers (GPUs, passes, tiles) add up to exactly the image one renderer would produce with jobSPP samples (up to
     * the order of the film's float additions). jobSPP: sample budget of the whole job (ZSobol's initialMaxSPP);
     * 0 = sampleOffset + totalSPP. */
    uint32_t        sampleOffset;
    uint32_t        jobSPP;
    /* NULL, or host materialCount * 8 floats — the constant attributes of the non-Lambert materials:
     *   (Mt)Refract (Tracer/MaterialsDefault.hpp:L232-312): [0..2] cauchyFront, [4..6] cauchyBack — Cauchy coefficients of
     *     the media in front of / behind the surface (index of refraction = c0 + c1 / l^2 + c2 / l^4, l in micrometres, at the
     *     path's first wavelength; c0 alone in the RGB renderer). Refraction disperses a spectral path to that wavelength.
     *   (Mt)Unreal (L466-760): [0] roughness, [1] specular, [2] metallic (constants; `albedo` as for Lambert, may be textured).
     * Other entries are ignored. */
    const float*    materialParams;
    /* PrimGroupTriangle's NORMAL attribute as the scene loader delivers it (Tracer/PrimitiveDefaultTriangle.cu:L169-184): one
     * world -> tangent-space rotation quaternion (w, x, y, z) per vertex. The hit's frame is Quaternion::BarySLerp of the
     * three vertex quaternions and the shading normal its Z axis (Triangle::GenerateSurface,
     * PrimitiveDefaultTriangle.hpp:L463-470) — takes precedence over vertexNormals / instanceVertexNormals, which interpolate
     * plain normals linearly. vertexTBN: host, vertexCount * 4 (single accelerator); instanceVertexTBN: one host pointer per
     * instance or NULL entries (two-level scenes). */
    const float*    vertexTBN;
    const float* const* instanceVertexTBN;
    /* The boundary light surface (TracerI::SetBoundarySurface): what a ray that leaves the scene sees, and one more entry
     * (the last) of the uniform light sampler. MRB_BOUNDARY_NULL = (L)Null. The skyspheres (LightGroupSkysphere<
     * SphericalCoordConverter / CoOctaCoordConverter>, Tracer/LightsDefault.hpp:L173-443) map the Y-up direction to a uv
     * on a latitude-longitude or concentric-octahedral map; boundaryTexture < 0: constant boundaryRadiance (sampled uniformly
     * in uv), else an index into `textures` holding the radiance map, importance-sampled through the piecewise-constant
     * 2-D distribution of its luminance (built at create time, see mrb_dist2d_build). boundaryTransform: NULL = (T)Identity,
     * else the row-major 3x4 local -> world matrix of the light surface's (T)Single transform (directions use its linear
     * part). sceneDiameter: how far away NEE places the sampled point; 0 = the reference's choice, the length of the XZ
     * span of the scene AABB (Tracer/TracerBase.cpp:L1653-1664). luminanceRow: the Y row of the global texture colour
     * space's RGB -> XYZ matrix (KCExtractLuminance, Tracer/ColorConverter.cu:L405-476); all zero = ACES_CG. */
    uint32_t        boundaryType;        /* mrb_boundary_type */
    int32_t         boundaryTexture;
    float           boundaryRadiance[3];
    const float*    boundaryTransform;
    float           sceneDiameter;
    float           luminanceRow[3];
    /* Normal maps (the optional "normalMap" attribute of (Mt)Lambert / (Mt)Unreal; Triangle::GenerateSurface,
     * Tracer/PrimitiveDefaultTriangle.hpp:L478-491,L571-575): NULL, or host int32 per material: -1, or an index into `textures`
     * whose rgb is the TANGENT-SPACE normal (used as is, then normalised). The hit's interpolated tangent frame is re-aimed so
     * that its Z axis is that normal; needs vertexTBN / instanceVertexTBN and the UVs. */
    const int32_t*  normalTexture;
    /* How a textured read turns the ray cone's UV gradients into a mip level (TracerTexView::operator()(uv, dpdx, dpdy)):
     * 0 = level = log2 of the longer gradient in UV units — what the reference's host backend computes for normalised-coordinate
     * textures (Device/CPU/TextureViewCPU.h:L405-420; the mode the parity goldens pin);
     * 1 = gradients scaled by the texture size first — what tex2DGrad does on the reference's device backends. */
    uint32_t        textureLodMode;
} mrb_render_desc;

typedef struct mrb_render_stats
{
    uint64_t pathsStarted, pathsCompleted;    /* camera paths */
    uint64_t closestRays, shadowRays;         /* rays cast (the reference only counts paths) */
    uint64_t iterations;
    uint64_t neeSamples;                      /* NEE light samples taken = the shadow rays the REFERENCE casts (it also
                                               * traces samples whose estimate is exactly zero; shadowRays omits those) */
    uint32_t finished;                        /* 1 when every pass begun so far has completed (triggerSave at the last one) */
} mrb_render_stats;

/* StartRender */
MRB_API mrb_status mrb_renderer_create(mrb_context ctx, const mrb_render_desc* desc, mrb_renderer* out);
/* StopRender + destroy */
MRB_API void       mrb_renderer_destroy(mrb_context ctx, mrb_renderer r);
/* DoRenderWork x iterations (throughput mode: one bounce of every live path per iteration);
 * asynchronous, no host synchronisation inside. */
MRB_API mrb_status mrb_renderer_iterate(mrb_context ctx, mrb_renderer r, uint32_t iterations);
/* Latency render mode (PathTracerRendererT::DoLatencyRender, TracerDLL/PathTracerRenderer.cu:L1078-1160): camera paths
 * are only started up to sppLimit samples per pixel (<= totalSPP; a new renderer starts at totalSPP = throughput mode).
 * The caller lowers it right after create, iterates until stats.finished, reads the film and raises it for the next
 * pass; raising requires the current pass to have finished. Synchronises. */
MRB_API mrb_status mrb_renderer_set_spp_limit(mrb_context ctx, mrb_renderer r, uint32_t sppLimit);
/* Synchronises and reads the counters. */
MRB_API mrb_status mrb_renderer_get_stats(mrb_context ctx, mrb_renderer r, mrb_render_stats* out);
/* Begins the pass "samples [sampleStart, sampleStart + sampleCount) of every pixel of the region regionMin .. regionMin +
 * regionSize" — one step of the tile / burst loop of PathTracerRendererT::DoLatencyRender
 * (TracerDLL/PathTracerRenderer.cu:L1078-1160) with ImageTiler::NextTile (Tracer/RenderImage.cpp:L20-136). The region must
 * fit the width x height the renderer was created with (its film) and the full image. Asynchronous; the previous pass
 * must have finished (mrb_renderer_run_pass / get_stats.finished). The film keeps accumulating: hand it over (or read it
 * with clear) before a pass with a different region. */
MRB_API mrb_status mrb_renderer_begin_pass(mrb_context ctx, mrb_renderer r, const uint32_t regionMin[2], const uint32_t regionSize[2],
                                           uint32_t sampleStart, uint32_t sampleCount);
/* DoRenderWork until the current pass has finished: `chunk` wavefront iterations (0 = 4) stay queued while the path
 * counters of the previous chunk are polled from pinned memory, so the device never idles on the host
 * (the reference synchronises after every iteration to read its dead-path count). Returns the final counters. */
MRB_API mrb_status mrb_renderer_run_pass(mrb_context ctx, mrb_renderer r, uint32_t chunk, mrb_render_stats* out);
/* Non-blocking counters for callers that iterate once per call (throughput mode, one bounce per DoRenderWork): returns
 * the latest snapshot that has landed in pinned host memory (zeros before the first one) and queues the next snapshot
 * behind the work issued so far — `finished` shows up an iteration or two after the last path died, without ever
 * draining the stream. */
MRB_API mrb_status mrb_renderer_poll_stats(mrb_context ctx, mrb_renderer r, mrb_render_stats* out);
/* RenderImage::TransferToHost (Tracer/RenderImage.cpp:L163-219): hands the film accumulated since the last hand-off to
 * hostDst (4 planar planes of the current region; pinned memory — mrb_host_alloc — makes the copy asynchronous) on a
 * copy stream, clears it there and calls onComplete(user) from a CUDA host callback once the data has landed (the
 * reference releases its timeline semaphore this way); rendering continues immediately into a second film buffer.
 * onComplete may be NULL; it must not call back into this library. */
typedef void (*mrb_host_fn)(void* user);
MRB_API mrb_status mrb_renderer_film_handoff(mrb_context ctx, mrb_renderer r, float* hostDst, mrb_host_fn onComplete, void* user);
/* Multi-GPU film reduction inside one process: adds the films of `peers` — renderers with the same
 * region that live on OTHER devices (peerCtx[k] is the context peers[k] was created on) — to r's film and clears them,
 * in one kernel on r's device that reads the peers' HBM over NVLink (peer access is enabled on demand). Stream-ordered on
 * both sides: the peers' pending iterations are joined before the read, their next ones wait for the clear. */
MRB_API mrb_status mrb_renderer_reduce_peers(mrb_context ctx, mrb_renderer r, const mrb_context* peerCtx, const mrb_renderer* peers,
                                             uint32_t peerCount);
/* Page-locked host memory for film staging (the reference's RenderImage staging memory, Tracer/RenderImage.cpp:L233-282). */
MRB_API mrb_status mrb_host_alloc(mrb_context ctx, size_t bytes, void** out);
MRB_API void       mrb_host_free(mrb_context ctx, void* ptr);
/* Parity tap: the film filter on its own (Tests/Tracer/T_Filters.cu). Host pointers: xi[count*2] ->
 * out[count*4] = {offset x, offset y, pdf of Sample(), Evaluate(offset)}. */
MRB_API mrb_status mrb_filter_sample(mrb_context ctx, uint32_t filterType, float radius, const float* xi, uint32_t count, float* out);
/* The film the reference hands over as RenderImageSection (Common/RenderImageStructs.h:L22-37):
 * 4 planar fp32 planes R,G,B,weight of width*height (row 0 = bottom), radiance sums and filter-weight
 * sums accumulated since the last clear. Copies to `out` (host or device); `clear` != 0 zeroes the
 * device film afterwards (the reference's per-iteration delta protocol). */
MRB_API mrb_status mrb_renderer_read_film(mrb_context ctx, mrb_renderer r, float* out, mrb_memspace memspace, int clear);
/* Device pointer of the film planes, for a multi-GPU film reduction (ncclAllReduce over NVLink). */
MRB_API float*     mrb_renderer_film_device_ptr(mrb_renderer r);
/* Parity tap: the renderer's texture filter on its own — TracerTexView<2, Vector3>::operator()(uv, dpdx, dpdy) of a
 * single-level texture (Tracer/TextureView.hpp:L70-85 -> Device/CPU/TextureViewCPU.h:L397-470). Host pointers:
 * uv[count*2] -> rgbOut[count*3]. */
MRB_API mrb_status mrb_texture_sample(mrb_context ctx, const mrb_texture_desc* texture, const float* uv, uint32_t count, float* rgbOut);
/* Parity tap: the upload-time colour conversion of a texture on its own (KCConvertColor): texelsOut (host) receives the texels of
 * `texture` after its gamma / colorMatrix have been applied, in the texture's own format. */
MRB_API mrb_status mrb_texture_convert(mrb_context ctx, const mrb_texture_desc* texture, void* texelsOut);
/* Parity taps of the mip path. mrb_texture_mip_chain: the whole chain of `texture` as the renderer stores it (supplied levels colour
 * converted, generated levels after them), chainOut (host) sized for mrb_texture_chain_texels(...) texels in the texture's own format;
 * *mipCountOut = levels. mrb_texture_sample_lod: TextureViewCPU::operator()(uv, mipLevel) (lod != NULL) or operator()(uv, dpdx, dpdy)
 * (grads = count * 4 floats {dpdx.xy, dpdy.xy}; lodMode as mrb_render_desc.textureLodMode) on that chain. Host pointers. */
MRB_API size_t     mrb_texture_chain_texels(uint32_t width, uint32_t height, uint32_t mipCount);
MRB_API uint32_t   mrb_texture_full_mip_count(uint32_t width, uint32_t height);
MRB_API mrb_status mrb_texture_mip_chain(mrb_context ctx, const mrb_texture_desc* texture, void* chainOut, uint32_t* mipCountOut);
/* {width, height, mip count} the renderer will hold `texture` with (clampResolution / generateMips applied); no device work */
MRB_API mrb_status mrb_texture_final_extent(const mrb_texture_desc* texture, uint32_t extentOut[3]);
MRB_API mrb_status mrb_texture_sample_lod(mrb_context ctx, const mrb_texture_desc* texture, const float* uv, const float* lod, const float* grads,
                                          uint32_t lodMode, uint32_t count, float* rgbOut);

/* ---- spectral LUT generation --------------------------------------------------------------- */

/* GenerateSpectraLUT of the reference's SpectraLUTGen tool (Source/SpectraLUTGen/main.cpp:L81-537, after Jakob & Hanika 2019 /
 * rgb2spec): the payload of SpectraLUT/<COLORSPACE>.mrspectra, computed on the device (one thread per (table, y, x) column,
 * Gauss-Newton in fp64 on the CIELab residual). Host pointers, exactly what the reference's tool reads from Core/ColorFunctions:
 * cieXYZ[471*3] (CIE 1931 observer, 360..830 nm), illuminantSPD[471] with its normalisation factor, the colour space's
 * RGB -> XYZ matrix without white-point adaptation (Color::GenRGBToXYZ(Primaries)) and its inverse (row-major 3x3).
 * lutOut: 9 * resolution^3 floats laid out [table l][coefficient][z][y][x]; whitepointOut (may be NULL): the XYZ white point
 * the residual is measured against. optimizePassCount: the tool uses 15. */
typedef struct mrb_spectra_lut_desc
{
    const float* cieXYZ;
    const float* illuminantSPD;
    float        illuminantNormFactor;
    float        rgbToXYZ[9];
    float        xyzToRGB[9];
    uint32_t     resolution;
    uint32_t     optimizePassCount;
} mrb_spectra_lut_desc;
MRB_API mrb_status mrb_spectra_lut_generate(mrb_context ctx, const mrb_spectra_lut_desc* desc, float* lutOut, double whitepointOut[3]);

/* ---- piecewise-constant 2-D distribution + skysphere maps ---------------------------------- */

/* DistributionGroupPwC2D::Construct (Tracer/Distributions.cu:L437-503): function = height rows of width values (a
 * texture's luminance) -> cdfX (width*height: the normalised inclusive CDF of |f| along every row) and cdfY (height: the
 * normalised CDF of the row sums). Row sums accumulate in fp64 and are stored as fp32, the marginal accumulates in fp32 in
 * row order, both are normalised by 1 / last in fp64 — the arithmetic of the reference's host-backend kernels. */
MRB_API mrb_status mrb_dist2d_build(mrb_context ctx, const float* function, uint32_t width, uint32_t height,
                                    float* cdfX, float* cdfY, mrb_memspace memspace);
/* DistributionPwC<2>::SampleUV + PdfUV (Tracer/Distributions.h:L183-239) for `count` random pairs xi[count*2]:
 * out[count*4] = {u, v, pdf of the sample, PdfUV(u, v)}. (Tests/Tracer/T_Distributions.cu:L17-44 KCSampleDist) */
MRB_API mrb_status mrb_dist2d_sample(mrb_context ctx, const float* cdfX, const float* cdfY, uint32_t width, uint32_t height,
                                     const float* xi, uint32_t count, float* out, mrb_memspace memspace);
/* Parity tap of the skysphere coordinate converters (Tracer/LightsDefault.hpp:L173-310): per unit Y-up direction
 * dirs[count*3] -> out[count*8] = {DirToUV u, v, ToSolidAnglePdf(1, dir), UVToDir(DirToUV(dir)) xyz, ToSolidAnglePdf(1, uv), 0}.
 * converter: MRB_BOUNDARY_SKYSPHERE_SPHERICAL or _COOCTA. */
MRB_API mrb_status mrb_skysphere_convert(mrb_context ctx, uint32_t converter, const float* dirs, uint32_t count, float* out,
                                         mrb_memspace memspace);
/* KCExtractLuminance (Tracer/ColorConverter.cu:L405-476) of a single-level texture: out[width*height] (host) =
 * luminanceRow . rgb of every texel. Host pointers. */
MRB_API mrb_status mrb_texture_luminance(mrb_context ctx, const mrb_texture_desc* texture, const float luminanceRow[3], float* out);

/* ---- device algorithms (Device/GPUAlgRadixSort.h, exposed for parity tests) -------------- */

/* Stable ascending LSD radix sort of (key,value) pairs over bits [bitBegin,bitEnd).
 * Replaces DeviceAlgorithms::RadixSort<true,K,uint32_t> (Device/CUDA/AlgRadixSortCUDA.h:L60-116).
 * In-place on the given arrays (internally double buffered). */
MRB_API mrb_status mrb_radix_sort_pairs_u64(mrb_context ctx, uint64_t* keys, uint32_t* values,
                                            uint32_t count, uint32_t bitBegin, uint32_t bitEnd,
                                            mrb_memspace memspace);
MRB_API mrb_status mrb_radix_sort_pairs_u32(mrb_context ctx, uint32_t* keys, uint32_t* values,
                                            uint32_t count, uint32_t bitBegin, uint32_t bitEnd,
                                            mrb_memspace memspace);

/* RayPartitioner::MultiPartition (Tracer/RayPartitioner.cu:L263-411): sorts (keys, indices) in place —
 * stable, by the data bit range then the batch bit range (batch-only when onlySortForBatches) — and
 * writes the partition table: partitionCount[0], partitionOffsets[count+1] (start of each run of equal
 * batch bits, closed by `count`), partitionKeys[count] (first sorted key of each run). At most
 * maxPartitions entries are written; batch range <= 16 bits. With MRB_MEM_DEVICE nothing synchronises
 * (the reference reads this table on the host every iteration). */
MRB_API mrb_status mrb_multi_partition(mrb_context ctx, uint32_t* keys, uint32_t* indices, uint32_t count,
                                       const uint32_t dataBitRange[2], const uint32_t batchBitRange[2],
                                       int onlySortForBatches, uint32_t maxPartitions,
                                       uint32_t* partitionCount, uint32_t* partitionOffsets, uint32_t* partitionKeys,
                                       mrb_memspace memspace);

/* RayPartitioner::BinaryPartition (Tracer/RayPartitioner.h:L154-200; cub::DevicePartition::If /
 * Device/CPU/AlgBinaryPartitionCPU.h): stable two-way split of an index list. indicesOut =
 * [indices i of indicesIn with flags[i] != 0, in order][the rest, in order]; leftCount[0] = size of the
 * first part. flags is indexed by the VALUE of the index (like IsAliveFunctor over dPathDataPack). */
MRB_API mrb_status mrb_binary_partition(mrb_context ctx, uint32_t* indicesOut, uint32_t* leftCount,
                                        const uint32_t* indicesIn, const uint8_t* flags, uint32_t flagCount,
                                        uint32_t count, mrb_memspace memspace);

#ifdef __cplusplus
}
#endif
#endif /* MRAY_B200_H */
