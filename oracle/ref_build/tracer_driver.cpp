// tracer_driver.cpp — drives ANY TracerDLL (the reference's libTracerDLL_CPU.so, or our
// libTracerDLL_B200.so) through the reference's own plugin interface Core/TracerI.h, exactly as
// MRay's TracerThread / SceneLoaderMRay / RunCommand do (MRay/TracerThread.cpp:L149-300,L855-876,
// SceneLoaderMRay/SceneLoaderMRay.cpp:L335-460,L2136-2260, MRay/RunCommand.cpp:L293-345).
//
// TEST / BASELINE INFRASTRUCTURE ONLY. Compiled against the unmodified reference headers in the
// authoring container; exposes one C entry point so Python (ctypes) can feed it numpy arrays.
#include "Core/TracerI.h"
#include "Core/ThreadPool.h"
#include "Core/TimelineSemaphore.h"
#include "Core/Error.h"
#include "Core/TypeNameGenerators.h"
#include "TransientPool/TransientPool.h"
#include "Core/GraphicsFunctions.h"

#include <dlfcn.h>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>
#include <cstdio>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <ucontext.h>

extern "C"
{

struct DriverScene
{
    // geometry: `batchCount` triangle batches, each with its own vertex list
    uint32_t        batchCount;
    const uint32_t* batchVertexOffsets; // batchCount + 1 (into positions / normals)
    const uint32_t* batchTriOffsets;    // batchCount + 1 (into indices; indices are batch-local)
    const float*    positions;          // V * 3
    const float*    normals;            // V * 3 (unit shading normals)
    const uint32_t* indices;            // T * 3
    const int32_t*  batchMaterial;      // batchCount : Lambert material index, or -1 for an emissive batch
    const int32_t*  batchLight;         // batchCount : light index, or -1
    uint32_t        materialCount;
    const float*    albedo;             // materialCount * 3
    uint32_t        lightCount;
    const float*    radiance;           // lightCount * 3
    // camera (pinhole)
    float           camPos[3], camGaze[3], camUp[3];
    float           fovXY[2];           // radians
    float           nearFar[2];
    // optional: batchCount row-major 3x4 local->world matrices; every batch then gets its own (T)Single
    // transform (positions are local-space). NULL = all surfaces use (T)Identity.
    const float*    batchTransforms;
    // optional: batchCount entries; batchInstanceOf[b] = a >= 0 makes the SURFACE of batch b use the primitive batch
    // of batch a (instancing: same geometry, own material and transform; b's own geometry stays unused). -1 = itself.
    const int32_t*  batchInstanceOf;
    // optional textured albedo: per texture 8 x u32 {width, height, format (0 = MR_RGBA_FLOAT, 1 = MR_RGBA8_UNORM),
    // MRayTextureInterpEnum, MRayTextureEdgeResolveEnum, byte offset into textureBytes, MRayColorSpaceEnum + 1 (0 = MR_DEFAULT),
    // gamma as float bits (0 = 1.0)}; materialTexture: per material
    // -1 or a texture index; uvs: V * 2 (UV0), NULL = zeros
    uint32_t        textureCount;
    const uint32_t* textureInfo;
    const uint8_t*  textureBytes;
    const int32_t*  materialTexture;
    const float*    uvs;
    // optional: per material 0 = (Mt)Lambert, 1 = (Mt)Reflect (its albedo entry is unused); NULL = all Lambert
    const uint8_t*  materialKind;
    // optional: per light the isTwoSided attribute of (L)Prim(P)Triangle; NULL = one-sided
    const uint8_t*  lightTwoSided;
    // optional: per material 8 floats. materialKind 2 = (Mt)Refract {cauchyFront xyz, -, cauchyBack xyz, -},
    // 3 = (Mt)Unreal {roughness, specular, metallic, ...} (albedo from `albedo`, constant)
    const float*    materialParams;
    // optional boundary light surface: 0 = (L)Null, 1 = (L)Skysphere_Spherical, 2 = (L)Skysphere_CoOcta; radiance constant, or
    // (boundaryTexture >= 0) the texture of that index as the radiance map; boundaryTransform: NULL = (T)Identity, else one
    // row-major 3x4 matrix pushed as a (T)Single transform for the light surface
    uint32_t        boundaryType;
    float           boundaryRadiance[3];
    int32_t         boundaryTexture;
    const float*    boundaryTransform;
    // optional alpha maps (SurfaceParams.alphaMaps): per batch -1 or a texture index; such textures have format 2
    // (MR_R_FLOAT) or 3 (MR_R8_UNORM) in textureInfo: single-channel pure data, read as Float (AlphaMap = TracerTexView<2, Float>)
    const int32_t*  batchAlphaMap;
    // optional normal maps: per material -1 or a texture index (RGBA texture read as Vector3): the optional texture-only
    // "normalMap" attribute of (Mt)Lambert / (Mt)Unreal
    const int32_t*  materialNormalMap;
    // optional: per texture the number of mip levels supplied (NULL = 1 each): CreateTexture2D(size, mipCount, ...) and one
    // PushTextureData per level; the levels lie back to back from the texture's byte offset, level k = max(size >> k, 1) texels
    const uint32_t* textureMipCounts;
};

struct DriverRender
{
    const char* rendererName;   // "PathTracerRGB" | "PathTracerSpectral"
    uint32_t    width, height;
    uint32_t    totalSPP;
    const char* sampleMode;     // "Pure" | "WithNextEventEstimation" | "WithNEEAndMIS"
    uint32_t    rrRange[2];
    uint64_t    seed;
    uint32_t    accelMode;      // AcceleratorType: 0 SOFTWARE_NONE(linear), 1 SOFTWARE_BASIC_BVH, 2 HARDWARE
    uint32_t    parallelHint;   // 0 = default (2^21)
    uint32_t    threads;        // host thread pool size (0 = hardware)
    uint32_t    samplerType;    // SamplerType::E: 0 Independent, 1 ZSobol, 2 Sobol
    uint32_t    region[4];      // regionMin.xy, regionMax.xy of RenderImageParams; all 0 = the whole image
    uint32_t    latency;        // 1 = renderMode "Latency" (every DoRenderWork completes burstSize samples per pixel)
    uint32_t    burstSize;      // 0 = 1
    // optional SetCameraTransform exercise: after `camSwitchAfter` DoRenderWork calls (0 = never) the camera is moved to
    // camSwitch{Pos,Gaze,Up} and the accumulation restarts (the driver clears its own accumulator, like Visor does)
    uint32_t    camSwitchAfter;
    float       camSwitch[9];
    // TracerParameters.filmFilter: filmFilter = 0 keeps the default (Gaussian, radius 1), else FilterType::E + 1
    // (1 Box, 2 Tent, 3 Gaussian, 4 Mitchell-Netravali); filmFilterRadius = 0 keeps the default radius
    uint32_t    filmFilter;
    float       filmFilterRadius;
    // TracerParameters.genMips / mipGenFilter: genMips = 1 completes every texture's chain by filtering; mipGenFilter = 0 keeps the
    // default (Gaussian, radius 2), else FilterType::E + 1 with mipGenFilterRadius
    uint32_t    genMips;
    uint32_t    mipGenFilter;
    float       mipGenFilterRadius;
};

struct DriverStats
{
    double commitSeconds;   // CommitSurfaces (BVH build)
    double renderSeconds;   // DoRenderWork loop
    double totalPaths;      // sum of section weights (approx. completed paths)
    uint32_t iterations;
    float  sceneAABB[6];
    double startSeconds;    // StartRender
    double sceneSeconds;    // ConstructTracer + scene upload calls (everything before CommitSurfaces)
    double closeSeconds;    // StopRender + DestroyRenderer + DestroyTracer
    double totalSeconds;    // the whole call
};

static void SegvInfo(int, siginfo_t* si, void* ctx)
{
    ucontext_t* uc = static_cast<ucontext_t*>(ctx);
    void* rip = reinterpret_cast<void*>(uc->uc_mcontext.gregs[REG_RIP]);
    Dl_info info; memset(&info, 0, sizeof(info));
    dladdr(rip, &info);
    char buf[1024];
    int n = snprintf(buf, sizeof(buf), "SEGV addr=%p rip=%p module=%s base=%p off=0x%lx sym=%s\n", si->si_addr, rip,
                     info.dli_fname ? info.dli_fname : "?", info.dli_fbase,
                     (unsigned long)((char*)rip - (char*)info.dli_fbase), info.dli_sname ? info.dli_sname : "?");
    write(2, buf, n);
    // walk a few frames by frame pointer-less heuristic: dump return addresses found on the stack that lie in the module
    unsigned long* sp = reinterpret_cast<unsigned long*>(uc->uc_mcontext.gregs[REG_RSP]);
    for(int i = 0, found = 0; i < 4096 && found < 12; i++)
    {
        Dl_info fi; memset(&fi, 0, sizeof(fi));
        if(dladdr(reinterpret_cast<void*>(sp[i]), &fi) && fi.dli_fbase == info.dli_fbase && fi.dli_sname)
        {
            n = snprintf(buf, sizeof(buf), "  stack[%d] off=0x%lx %s\n", i, (unsigned long)((char*)sp[i] - (char*)fi.dli_fbase), fi.dli_sname);
            write(2, buf, n); found++;
        }
    }
    _exit(98);
}

static void AlarmBacktrace(int)
{
    void* frames[64];
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(99);
}

static void Fail(char* err, size_t n, const std::string& s)
{
    std::snprintf(err, n, "%s", s.c_str());
}

// Returns 0 on success. outRGB: width*height*3 floats (row 0 = bottom row, as the tracer delivers),
// accumulated in double like RunCommand (out = (out*W + in)/(W + w)).
int tracer_driver_render(const char* dllPath, const DriverScene* sc, const DriverRender* rd,
                         float* outRGB, float* outWeight, DriverStats* stats, char* err, size_t errLen)
{
    using namespace std::string_literals;
    const auto callStart = std::chrono::steady_clock::now();
    if(getenv("DRIVER_ALARM"))
    {
        static char altStack[1 << 16];
        stack_t ss; ss.ss_sp = altStack; ss.ss_size = sizeof(altStack); ss.ss_flags = 0;
        sigaltstack(&ss, nullptr);
        struct sigaction sa; memset(&sa, 0, sizeof(sa));
        sa.sa_sigaction = SegvInfo; sa.sa_flags = SA_SIGINFO;
        sigaction(SIGSEGV, &sa, nullptr);
    }
    void* lib = dlopen(dllPath, RTLD_NOW | RTLD_GLOBAL);
    if(!lib) { Fail(err, errLen, "dlopen failed: "s + dlerror()); return 1; }
    using ConstructF = TracerI* (*)(const TracerParameters&);
    using DestroyF = void (*)(TracerI*);
    auto construct = reinterpret_cast<ConstructF>(dlsym(lib, "ConstructTracer"));
    auto destroy = reinterpret_cast<DestroyF>(dlsym(lib, "DestroyTracer"));
    if(!construct || !destroy) { Fail(err, errLen, "ConstructTracer/DestroyTracer not exported"); return 2; }

    TracerI* tracer = nullptr;
    try
    {
        TracerParameters tp;
        tp.seed = rd->seed;
        tp.accelMode = AcceleratorType(rd->accelMode);
        if(rd->parallelHint) tp.parallelizationHint = rd->parallelHint;
        tp.samplerType = SamplerType::E(rd->samplerType);
        if(rd->filmFilter) tp.filmFilter.type = FilterType::E(rd->filmFilter - 1u);
        if(rd->filmFilterRadius > 0.0f) tp.filmFilter.radius = rd->filmFilterRadius;
        tp.genMips = rd->genMips != 0;
        if(rd->mipGenFilter) tp.mipGenFilter.type = FilterType::E(rd->mipGenFilter - 1u);
        if(rd->mipGenFilterRadius > 0.0f) tp.mipGenFilter.radius = rd->mipGenFilterRadius;
        tracer = construct(tp);
        // as MRay/RunCommand.cpp:L1015-1025: worker threads run the tracer's device-init function
        ThreadPool pool;
        auto threadInit = tracer->GetThreadInitFunction();
        pool.RestartThreads(rd->threads ? rd->threads : std::thread::hardware_concurrency(),
                            [threadInit](std::thread::native_handle_type, uint32_t) { threadInit(); });
        tracer->SetThreadPool(pool);
        threadInit();

        // ---- primitives ----
        PrimGroupId pg = tracer->CreatePrimitiveGroup("(P)Triangle");
        std::vector<PrimCount> counts;
        for(uint32_t b = 0; b < sc->batchCount; b++)
            counts.push_back(PrimCount{sc->batchTriOffsets[b + 1] - sc->batchTriOffsets[b],
                                       sc->batchVertexOffsets[b + 1] - sc->batchVertexOffsets[b]});
        PrimBatchIdList batches = tracer->ReservePrimitiveBatches(pg, counts);
        tracer->CommitPrimReservations(pg);
        PrimAttributeInfoList pInfo = tracer->AttributeInfo(pg);
        for(uint32_t b = 0; b < sc->batchCount; b++)
        {
            uint32_t v0 = sc->batchVertexOffsets[b], vN = counts[b].attributeCount;
            uint32_t t0 = sc->batchTriOffsets[b], tN = counts[b].primCount;
            for(uint32_t a = 0; a < pInfo.size(); a++)
            {
                using enum PrimitiveAttributeLogic::E;
                switch(pInfo[a].logic.e)
                {
                    case POSITION:
                    {
                        TransientData d(std::in_place_type_t<Vector3>{}, vN);
                        d.Push(Span<const Vector3>(reinterpret_cast<const Vector3*>(sc->positions) + v0, vN));
                        tracer->PushPrimAttribute(pg, batches[b], a, std::move(d));
                        break;
                    }
                    case NORMAL:
                    {
                        // normals travel as tangent-space rotations (SceneLoaderMRay.cpp:L395-436)
                        TransientData d(std::in_place_type_t<Quaternion>{}, vN);
                        for(uint32_t i = 0; i < vN; i++)
                        {
                            Vector3 n = Math::Normalize(reinterpret_cast<const Vector3*>(sc->normals)[v0 + i]);
                            Vector3 bt = Graphics::OrthogonalVector(n);
                            Vector3 t = Math::Cross(bt, n);
                            Quaternion q = TransformGen::ToSpaceQuat(t, bt, n);
                            d.Push(Span<const Quaternion>(&q, 1));
                        }
                        tracer->PushPrimAttribute(pg, batches[b], a, std::move(d));
                        break;
                    }
                    case UV0:
                    {
                        TransientData d(std::in_place_type_t<Vector2>{}, vN);
                        std::vector<Vector2> z(vN, Vector2::Zero());
                        if(sc->uvs) for(uint32_t i = 0; i < vN; i++) z[i] = Vector2(sc->uvs[2 * size_t(v0 + i)], sc->uvs[2 * size_t(v0 + i) + 1]);
                        d.Push(Span<const Vector2>(z));
                        tracer->PushPrimAttribute(pg, batches[b], a, std::move(d));
                        break;
                    }
                    case INDEX:
                    {
                        TransientData d(std::in_place_type_t<Vector3ui>{}, tN);
                        d.Push(Span<const Vector3ui>(reinterpret_cast<const Vector3ui*>(sc->indices) + t0, tN));
                        tracer->PushPrimAttribute(pg, batches[b], a, std::move(d));
                        break;
                    }
                    default: break;
                }
            }
        }
        // ---- textures (SceneLoaderMRay.cpp:L1040-1105: CreateTexture2D for all -> CommitTextures (allocation) -> PushTextureData) ----
        std::vector<TextureId> texIds;
        for(uint32_t t = 0; t < sc->textureCount; t++)
        {
            const uint32_t* ti = sc->textureInfo + 8 * size_t(t);
            MRayTextureParameters tp;
            tp.pixelType = MRayPixelTypeRT(ti[2] == 0 ? MRayPixelEnum::MR_RGBA_FLOAT : ti[2] == 1 ? MRayPixelEnum::MR_RGBA8_UNORM
                                           : ti[2] == 2 ? MRayPixelEnum::MR_R_FLOAT : MRayPixelEnum::MR_R8_UNORM);
            tp.colorSpace = ti[6] ? MRayColorSpaceEnum(ti[6] - 1u) : MRayColorSpaceEnum::MR_DEFAULT;
            tp.gamma = Float(1);
            if(ti[7]) { float g; std::memcpy(&g, &ti[7], 4); tp.gamma = g; }
            tp.interpolation = MRayTextureInterpEnum(ti[3]); tp.edgeResolve = MRayTextureEdgeResolveEnum(ti[4]);
            if(ti[2] < 2) tp.readMode = MRayTextureReadMode::MR_DROP_1;   // RGBA pixels read as Vector3 (TextureReadMode::TO_3C_FROM_4C): the albedo's view type
            else { tp.readMode = MRayTextureReadMode::MR_PASSTHROUGH; tp.isColor = AttributeIsColor::IS_PURE_DATA; }   // alpha maps
            texIds.push_back(tracer->CreateTexture2D(Vector2ui(ti[0], ti[1]), sc->textureMipCounts ? sc->textureMipCounts[t] : 1u, tp));
        }
        tracer->CommitTextures();
        for(uint32_t t = 0; t < sc->textureCount; t++)
        {
            const uint32_t* ti = sc->textureInfo + 8 * size_t(t);
            const Byte* src = reinterpret_cast<const Byte*>(sc->textureBytes) + ti[5];
            const uint32_t levels = sc->textureMipCounts ? sc->textureMipCounts[t] : 1u;
            for(uint32_t level = 0; level < levels; level++)
            {
                const uint32_t lw = std::max(ti[0] >> level, 1u), lh = std::max(ti[1] >> level, 1u);
                size_t pixels = size_t(lw) * lh;
                // TransientData is typed by the pixel (the reference reads it back with AccessAs<PixelType>)
                if(ti[2] == 0)
                {
                    TransientData d(std::in_place_type_t<Vector4>{}, pixels);
                    d.Push(Span<const Vector4>(reinterpret_cast<const Vector4*>(src), pixels));
                    tracer->PushTextureData(texIds[t], level, std::move(d));
                    src += pixels * sizeof(Vector4);
                }
                else if(ti[2] == 2)
                {
                    TransientData d(std::in_place_type_t<Float>{}, pixels);
                    d.Push(Span<const Float>(reinterpret_cast<const Float*>(src), pixels));
                    tracer->PushTextureData(texIds[t], level, std::move(d));
                    src += pixels * sizeof(Float);
                }
                else if(ti[2] == 3)
                {
                    TransientData d(std::in_place_type_t<uint8_t>{}, pixels);
                    d.Push(Span<const uint8_t>(reinterpret_cast<const uint8_t*>(src), pixels));
                    tracer->PushTextureData(texIds[t], level, std::move(d));
                    src += pixels;
                }
                else
                {
                    TransientData d(std::in_place_type_t<Vector4uc>{}, pixels);
                    d.Push(Span<const Vector4uc>(reinterpret_cast<const Vector4uc*>(src), pixels));
                    tracer->PushTextureData(texIds[t], level, std::move(d));
                    src += pixels * 4;
                }
            }
        }
        // ---- materials: (Mt)Lambert (constant or textured albedo) and, where materialKind says so, (Mt)Reflect ----
        std::vector<uint32_t> lambertOf, reflectOf, refractOf, unrealOf;   // scene material index per group entry
        for(uint32_t m = 0; m < sc->materialCount; m++)
        {
            const uint32_t kind = sc->materialKind ? sc->materialKind[m] : 0u;
            (kind == 1 ? reflectOf : kind == 2 ? refractOf : kind == 3 ? unrealOf : lambertOf).push_back(m);
        }
        MaterialIdList mats(sc->materialCount);
        if(!lambertOf.empty())
        {
            const uint32_t n = uint32_t(lambertOf.size());
            MatGroupId mg = tracer->CreateMaterialGroup("(Mt)Lambert");
            MatAttributeInfoList mInfo = tracer->AttributeInfo(mg);
            std::vector<AttributeCountList> mCounts(n);
            for(auto& c : mCounts) { c = AttributeCountList(StaticVecSize(mInfo.size())); c[0] = 1; for(size_t k = 1; k < mInfo.size(); k++) c[k] = 0; }
            MaterialIdList ids = tracer->ReserveMaterials(mg, mCounts);
            tracer->CommitMatReservations(mg);
            for(uint32_t k = 0; k < n; k++) mats[lambertOf[k]] = ids[k];
            auto range = CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back()));
            std::vector<Vector3> alb(n);
            std::vector<Optional<TextureId>> albedoTex(n, std::nullopt);
            for(uint32_t k = 0; k < n; k++)
            {
                const uint32_t m = lambertOf[k];
                alb[k] = reinterpret_cast<const Vector3*>(sc->albedo)[m];
                if(sc->textureCount && sc->materialTexture && sc->materialTexture[m] >= 0) albedoTex[k] = texIds[size_t(sc->materialTexture[m])];
            }
            TransientData d(std::in_place_type_t<Vector3>{}, n);
            d.Push(Span<const Vector3>(alb.data(), n));
            tracer->PushMatAttribute(mg, range, 0, std::move(d), std::move(albedoTex));
            // Optional texture-only attributes (Lambert: 1 = normalMap) are pushed with an EMPTY TransientData and
            // nullopt ids, as SceneLoaderMRay does (SceneLoaderMRay.cpp:L190-245; TracerBase::PushMatAttribute routes
            // data.IsEmpty() to the texture-only overload, TracerBase.cpp:L843-867). The reference allocates the
            // Optional<TracerTexView> array uninitialised, so leaving this out makes it read garbage.
            for(uint32_t a = 1; a < mInfo.size(); a++)
                if(mInfo[a].isTexturable == AttributeTexturable::MR_TEXTURE_ONLY &&
                   mInfo[a].isOptional == AttributeOptionality::MR_OPTIONAL)
                {
                    TransientData e(std::in_place_type_t<Vector3>{}, 0);
                    std::vector<Optional<TextureId>> nm(n, std::nullopt);
                    if(sc->materialNormalMap && mInfo[a].name == "normalMap")
                        for(uint32_t k = 0; k < n; k++)
                            if(sc->materialNormalMap[lambertOf[k]] >= 0) nm[k] = texIds[size_t(sc->materialNormalMap[lambertOf[k]])];
                    tracer->PushMatAttribute(mg, range, a, std::move(e), std::move(nm));
                }
        }
        if(!reflectOf.empty())
        {
            MatGroupId rg = tracer->CreateMaterialGroup("(Mt)Reflect");      // no attributes
            MatAttributeInfoList rInfoM = tracer->AttributeInfo(rg);
            std::vector<AttributeCountList> rCounts(reflectOf.size(), AttributeCountList(StaticVecSize(rInfoM.size())));
            MaterialIdList ids = tracer->ReserveMaterials(rg, rCounts);
            tracer->CommitMatReservations(rg);
            for(size_t k = 0; k < reflectOf.size(); k++) mats[reflectOf[k]] = ids[k];
        }
        if(!refractOf.empty())
        {   // (Mt)Refract: constant-only Cauchy coefficients, attribute 0 = cauchyBack, 1 = cauchyFront
            const uint32_t n = uint32_t(refractOf.size());
            MatGroupId g = tracer->CreateMaterialGroup("(Mt)Refract");
            MatAttributeInfoList info = tracer->AttributeInfo(g);
            std::vector<AttributeCountList> counts(n);
            for(auto& c : counts) { c = AttributeCountList(StaticVecSize(info.size())); for(size_t k = 0; k < info.size(); k++) c[k] = 1; }
            MaterialIdList ids = tracer->ReserveMaterials(g, counts);
            tracer->CommitMatReservations(g);
            auto range = CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back()));
            for(uint32_t a = 0; a < 2; a++)
            {
                std::vector<Vector3> v(n);
                for(uint32_t k = 0; k < n; k++)
                {
                    const float* mp = sc->materialParams + 8 * size_t(refractOf[k]) + (info[a].name == "cauchyFront" ? 0 : 4);
                    v[k] = Vector3(mp[0], mp[1], mp[2]); mats[refractOf[k]] = ids[k];
                }
                TransientData d(std::in_place_type_t<Vector3>{}, n);
                d.Push(Span<const Vector3>(v.data(), n));
                tracer->PushMatAttribute(g, range, a, std::move(d));
            }
        }
        if(!unrealOf.empty())
        {   // (Mt)Unreal: albedo / roughness / specular / metallic as constants (ParamVarying, no textures), normalMap absent
            const uint32_t n = uint32_t(unrealOf.size());
            MatGroupId g = tracer->CreateMaterialGroup("(Mt)Unreal");
            MatAttributeInfoList info = tracer->AttributeInfo(g);
            std::vector<AttributeCountList> counts(n);
            for(auto& c : counts)
            {
                c = AttributeCountList(StaticVecSize(info.size()));
                for(size_t k = 0; k < info.size(); k++) c[k] = (info[k].isOptional == AttributeOptionality::MR_OPTIONAL) ? 0 : 1;
            }
            MaterialIdList ids = tracer->ReserveMaterials(g, counts);
            tracer->CommitMatReservations(g);
            for(uint32_t k = 0; k < n; k++) mats[unrealOf[k]] = ids[k];
            auto range = CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back()));
            for(uint32_t a = 0; a < info.size(); a++)
            {
                std::vector<Optional<TextureId>> noTex(n, std::nullopt);
                if(info[a].isTexturable == AttributeTexturable::MR_TEXTURE_ONLY)
                {
                    TransientData e(std::in_place_type_t<Vector3>{}, 0);
                    tracer->PushMatAttribute(g, range, a, std::move(e), std::move(noTex));
                }
                else if(info[a].dataType.Name() == MRayDataEnum::MR_VECTOR_3)
                {
                    std::vector<Vector3> v(n);
                    for(uint32_t k = 0; k < n; k++) v[k] = reinterpret_cast<const Vector3*>(sc->albedo)[unrealOf[k]];
                    TransientData d(std::in_place_type_t<Vector3>{}, n);
                    d.Push(Span<const Vector3>(v.data(), n));
                    tracer->PushMatAttribute(g, range, a, std::move(d), std::move(noTex));
                }
                else
                {
                    const uint32_t slot = info[a].name == "roughness" ? 0u : info[a].name == "specular" ? 1u : 2u;
                    std::vector<Float> v(n);
                    for(uint32_t k = 0; k < n; k++) v[k] = sc->materialParams[8 * size_t(unrealOf[k]) + slot];
                    TransientData d(std::in_place_type_t<Float>{}, n);
                    d.Push(Span<const Float>(v.data(), n));
                    tracer->PushMatAttribute(g, range, a, std::move(d), std::move(noTex));
                }
            }
        }
        // ---- lights (prim backed) ----
        LightGroupId lg = sc->lightCount ? tracer->CreateLightGroup("(L)Prim(P)Triangle", pg) : LightGroupId(0);
        LightAttributeInfoList lInfo = tracer->AttributeInfo(lg);
        std::vector<AttributeCountList> lCounts(sc->lightCount);
        for(auto& c : lCounts) { c = AttributeCountList(StaticVecSize(lInfo.size())); for(size_t k = 0; k < lInfo.size(); k++) c[k] = 1; }
        std::vector<PrimBatchId> lightBatches(sc->lightCount);
        for(uint32_t b = 0; b < sc->batchCount; b++)
            if(sc->batchLight[b] >= 0) lightBatches[size_t(sc->batchLight[b])] = batches[b];
        LightIdList lights;
        if(sc->lightCount)
        {
            lights = tracer->ReserveLights(lg, lCounts, lightBatches);
            tracer->CommitLightReservations(lg);
            auto range = CommonIdRange(std::bit_cast<CommonId>(lights.front()), std::bit_cast<CommonId>(lights.back()));
            for(uint32_t a = 0; a < lInfo.size(); a++)
            {
                if(lInfo[a].dataType.Name() == MRayDataEnum::MR_VECTOR_3)
                {
                    TransientData d(std::in_place_type_t<Vector3>{}, sc->lightCount);
                    d.Push(Span<const Vector3>(reinterpret_cast<const Vector3*>(sc->radiance), sc->lightCount));
                    tracer->PushLightAttribute(lg, range, a, std::move(d),
                                               std::vector<Optional<TextureId>>(sc->lightCount, std::nullopt));
                }
                else
                {
                    TransientData d(std::in_place_type_t<bool>{}, sc->lightCount);
                    std::vector<uint8_t> f(sc->lightCount, 0);
                    if(sc->lightTwoSided) for(uint32_t l = 0; l < sc->lightCount; l++) f[l] = sc->lightTwoSided[l] ? 1 : 0;
                    d.Push(Span<const bool>(reinterpret_cast<const bool*>(f.data()), sc->lightCount));
                    tracer->PushLightAttribute(lg, range, a, std::move(d));
                }
            }
        }
        // ---- camera ----
        CameraGroupId cg = tracer->CreateCameraGroup("(C)Pinhole");
        CamAttributeInfoList cInfo = tracer->AttributeInfo(cg);
        AttributeCountList cCount(StaticVecSize(cInfo.size()));
        for(size_t k = 0; k < cInfo.size(); k++) cCount[k] = 1;
        CameraId cam = tracer->ReserveCamera(cg, cCount);
        tracer->CommitCamReservations(cg);
        {
            auto range = CommonIdRange(std::bit_cast<CommonId>(cam), std::bit_cast<CommonId>(cam));
            // attribute order of CameraGroupPinhole: FovAndPlanes, gaze, position, up
            Vector4 fp(sc->fovXY[0], sc->fovXY[1], sc->nearFar[0], sc->nearFar[1]);
            Vector3 v[3] = {Vector3(sc->camGaze[0], sc->camGaze[1], sc->camGaze[2]),
                            Vector3(sc->camPos[0], sc->camPos[1], sc->camPos[2]),
                            Vector3(sc->camUp[0], sc->camUp[1], sc->camUp[2])};
            TransientData d0(std::in_place_type_t<Vector4>{}, 1); d0.Push(Span<const Vector4>(&fp, 1));
            tracer->PushCamAttribute(cg, range, 0, std::move(d0));
            for(uint32_t a = 1; a < 4; a++)
            {
                TransientData d(std::in_place_type_t<Vector3>{}, 1); d.Push(Span<const Vector3>(&v[a - 1], 1));
                tracer->PushCamAttribute(cg, range, a, std::move(d));
            }
        }
        // ---- transforms ----
        std::vector<TransformId> batchTrans(sc->batchCount, TracerConstants::IdentityTransformId);
        TransformId skyT = TracerConstants::IdentityTransformId;   // transform of the boundary light surface
        const bool skyTransform = sc->boundaryType != 0u && sc->boundaryTransform;
        if(sc->batchTransforms || skyTransform)
        {
            TransGroupId tg = tracer->CreateTransformGroup("(T)Single");
            const uint32_t nBatchT = sc->batchTransforms ? sc->batchCount : 0u;
            std::vector<AttributeCountList> tCounts(nBatchT + (skyTransform ? 1u : 0u));
            for(auto& c : tCounts) { c = AttributeCountList(StaticVecSize(1)); c[0] = 1; }
            TransformIdList tids = tracer->ReserveTransformations(tg, tCounts);
            tracer->CommitTransReservations(tg);
            std::vector<Matrix3x4> ms;
            for(uint32_t b = 0; b < nBatchT; b++)
            {
                const float* m = sc->batchTransforms + 12 * size_t(b);
                ms.push_back(Matrix3x4(Vector4(m[0], m[1], m[2], m[3]), Vector4(m[4], m[5], m[6], m[7]), Vector4(m[8], m[9], m[10], m[11])));
                batchTrans[b] = tids[b];
            }
            if(skyTransform)
            {
                const float* m = sc->boundaryTransform;
                ms.push_back(Matrix3x4(Vector4(m[0], m[1], m[2], m[3]), Vector4(m[4], m[5], m[6], m[7]), Vector4(m[8], m[9], m[10], m[11])));
                skyT = tids[nBatchT];
            }
            TransientData d(std::in_place_type_t<Matrix3x4>{}, ms.size());
            d.Push(Span<const Matrix3x4>(ms.data(), ms.size()));
            tracer->PushTransAttribute(tg, CommonIdRange(std::bit_cast<CommonId>(tids.front()), std::bit_cast<CommonId>(tids.back())), 0, std::move(d));
        }
        // ---- surfaces ----
        for(uint32_t b = 0; b < sc->batchCount; b++)
        {
            if(sc->batchMaterial[b] < 0) continue;
            SurfaceParams sp;
            sp.primBatches.push_back(batches[(sc->batchInstanceOf && sc->batchInstanceOf[b] >= 0) ? uint32_t(sc->batchInstanceOf[b]) : b]);
            sp.materials.push_back(mats[size_t(sc->batchMaterial[b])]);
            sp.transformId = batchTrans[b];
            if(sc->batchAlphaMap && sc->batchAlphaMap[b] >= 0) sp.alphaMaps.push_back(texIds[size_t(sc->batchAlphaMap[b])]);
            else sp.alphaMaps.push_back(std::nullopt);
            sp.cullFaceFlags.push_back(false);
            sp.volumes.push_back(TracerConstants::InvalidVolume);
            tracer->CreateSurface(sp);
        }
        for(uint32_t b = 0; b < sc->batchCount; b++)
            if(sc->batchLight[b] >= 0)
                tracer->CreateLightSurface(LightSurfaceParams{lights[size_t(sc->batchLight[b])], batchTrans[b], {}});
        CamSurfaceId camSurf = tracer->CreateCameraSurface(CameraSurfaceParams{cam, TracerConstants::IdentityTransformId, {}});
        if(sc->boundaryType == 0u)
            tracer->SetBoundarySurface(TracerConstants::NullLightId, TracerConstants::IdentityTransformId);
        else
        {   // the skysphere as the scene loader creates it (SceneLoaderMRay.cpp:L1590-1660,L2210-2222): a light group over the
            // empty primitive group, ONE light whose "radiance" is a constant or a texture
            LightGroupId sg = tracer->CreateLightGroup(sc->boundaryType == 1u ? "(L)Skysphere_Spherical" : "(L)Skysphere_CoOcta");
            LightAttributeInfoList sInfo = tracer->AttributeInfo(sg);
            AttributeCountList sCount(StaticVecSize(sInfo.size()));
            for(size_t k = 0; k < sInfo.size(); k++) sCount[k] = 1;
            LightId sky = tracer->ReserveLight(sg, sCount);
            tracer->CommitLightReservations(sg);
            auto range = CommonIdRange(std::bit_cast<CommonId>(sky), std::bit_cast<CommonId>(sky));
            Vector3 rad(sc->boundaryRadiance[0], sc->boundaryRadiance[1], sc->boundaryRadiance[2]);
            TransientData d(std::in_place_type_t<Vector3>{}, 1);
            d.Push(Span<const Vector3>(&rad, 1));
            std::vector<Optional<TextureId>> tex(1, std::nullopt);
            if(sc->boundaryTexture >= 0) tex[0] = texIds[size_t(sc->boundaryTexture)];
            tracer->PushLightAttribute(sg, range, 0, std::move(d), std::move(tex));
            tracer->SetBoundarySurface(sky, skyT);
        }
        VolumeId bVol = tracer->RegisterVolume(VolumeParams{TracerConstants::VacuumMediumId,
                                                            TracerConstants::IdentityTransformId, 0});
        tracer->SetBoundaryVolume(bVol);

        auto c0 = std::chrono::steady_clock::now();
        stats->sceneSeconds = std::chrono::duration<double>(c0 - callStart).count();
        SurfaceCommitResult cr = tracer->CommitSurfaces();
        auto c1 = std::chrono::steady_clock::now();
        stats->commitSeconds = std::chrono::duration<double>(c1 - c0).count();
        for(int k = 0; k < 3; k++) { stats->sceneAABB[k] = cr.aabb.Min()[k]; stats->sceneAABB[3 + k] = cr.aabb.Max()[k]; }

        // ---- renderer ----
        TimelineSemaphore sem(0);
        tracer->SetupRenderEnv(&sem, 4096, 0);
        RendererId rid = tracer->CreateRenderer(std::string("(R)") + rd->rendererName);
        RendererAttributeInfoList rInfo = tracer->AttributeInfo(rid);
        for(uint32_t a = 0; a < rInfo.size(); a++)
        {
            std::string_view name = rInfo[a].name;
            MRayDataTypeRT dt = rInfo[a].dataType;
            auto PushU32 = [&](uint32_t v)
            { TransientData d(std::in_place_type_t<uint32_t>{}, 1); d.Push(Span<const uint32_t>(&v, 1)); tracer->PushRendererAttribute(rid, a, std::move(d)); };
            auto PushStr = [&](std::string_view s)
            {   // as MRay/TracerThread.cpp:L205-216
                TransientData d = AllocateTransientData(dt, s.size());
                d.ReserveAll();
                Span<char> o = d.AccessAsString();
                std::copy(s.cbegin(), s.cend(), o.begin());
                tracer->PushRendererAttribute(rid, a, std::move(d));
            };
            if(name == "totalSPP") PushU32(rd->totalSPP);
            else if(name == "burstSize") PushU32(rd->burstSize ? rd->burstSize : 1u);
            else if(name == "renderMode") PushStr(rd->latency ? "Latency" : "Throughput");
            else if(name == "sampleMode") PushStr(rd->sampleMode);
            else if(name == "rrRange")
            {
                Vector2ui v(rd->rrRange[0], rd->rrRange[1]);
                TransientData d(std::in_place_type_t<Vector2ui>{}, 1); d.Push(Span<const Vector2ui>(&v, 1));
                tracer->PushRendererAttribute(rid, a, std::move(d));
            }
            else if(name == "neeSamplerType") PushStr("Uniform");
            else if(rInfo[a].isOptional == AttributeOptionality::MR_MANDATORY)
                throw MRayError("driver: unknown mandatory renderer attribute {}", name);
            (void)dt;
        }
        RenderImageParams rip{Vector2ui(rd->width, rd->height), Vector2ui(0, 0), Vector2ui(rd->width, rd->height)};
        if(rd->region[2] | rd->region[3]) { rip.regionMin = Vector2ui(rd->region[0], rd->region[1]); rip.regionMax = Vector2ui(rd->region[2], rd->region[3]); }
        if(getenv("DRIVER_VERBOSE")) std::fprintf(stderr, "StartRender...\n");
        if(const char* a = getenv("DRIVER_ALARM")) { signal(SIGALRM, AlarmBacktrace); alarm(unsigned(atoi(a))); }
        auto s0 = std::chrono::steady_clock::now();
        RenderBufferInfo rbi = tracer->StartRender(rid, camSurf, rip, std::nullopt, std::nullopt);
        stats->startSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - s0).count();
        if(getenv("DRIVER_VERBOSE")) std::fprintf(stderr, "StartRender done, buffer %zu bytes\n", rbi.totalSize);

        size_t pix = size_t(rd->width) * rd->height;
        std::vector<double> acc(pix * 4, 0.0);
        auto r0 = std::chrono::steady_clock::now();
        uint32_t iters = 0;
        while(true)
        {
            if(rd->camSwitchAfter && iters == rd->camSwitchAfter)
            {
                CameraTransform ct{Vector3(rd->camSwitch[0], rd->camSwitch[1], rd->camSwitch[2]),
                                   Vector3(rd->camSwitch[3], rd->camSwitch[4], rd->camSwitch[5]),
                                   Vector3(rd->camSwitch[6], rd->camSwitch[7], rd->camSwitch[8])};
                tracer->SetCameraTransform(rid, ct);
                std::fill(acc.begin(), acc.end(), 0.0);
            }
            RendererOutput out = tracer->DoRenderWork();
            iters++;
            if(getenv("DRIVER_VERBOSE")) std::fprintf(stderr, "iter %u img %d save %d\n", iters, int(out.imageOut.has_value()), int(out.triggerSave));
            if(out.imageOut)
            {
                const RenderImageSection& s = *out.imageOut;
                if(!sem.Acquire(s.waitCounter)) break;
                const float* R = reinterpret_cast<const float*>(rbi.data + s.pixStartOffsets[0]);
                const float* G = reinterpret_cast<const float*>(rbi.data + s.pixStartOffsets[1]);
                const float* B = reinterpret_cast<const float*>(rbi.data + s.pixStartOffsets[2]);
                const float* Wt = reinterpret_cast<const float*>(rbi.data + s.weightStartOffset);
                uint32_t w = s.pixelMax[0] - s.pixelMin[0], h = s.pixelMax[1] - s.pixelMin[1];
                for(uint32_t y = 0; y < h; y++)
                for(uint32_t x = 0; x < w; x++)
                {
                    size_t src = size_t(y) * w + x;
                    size_t dst = (size_t(y + s.pixelMin[1]) * rd->width + (x + s.pixelMin[0])) * 4;
                    // RunCommand keeps a running weighted mean; summing numerator and weight is the same value
                    acc[dst + 0] += R[src]; acc[dst + 1] += G[src]; acc[dst + 2] += B[src]; acc[dst + 3] += double(Wt[src]) * s.globalWeight;
                }
                sem.Release();
            }
            if(out.triggerSave) break;
            if(iters > 100000000u) break;
        }
        auto r1 = std::chrono::steady_clock::now();
        tracer->StopRender();
        stats->renderSeconds = std::chrono::duration<double>(r1 - r0).count();
        stats->iterations = iters;
        double tw = 0;
        for(size_t p = 0; p < pix; p++)
        {
            double w = acc[p * 4 + 3];
            tw += w;
            for(int c = 0; c < 3; c++) outRGB[p * 3 + c] = float(w > 0 ? acc[p * 4 + c] / w : 0.0);
            if(outWeight) outWeight[p] = float(w);
        }
        stats->totalPaths = tw;
        tracer->DestroyRenderer(rid);
        destroy(tracer);
        tracer = nullptr;
        const auto callEnd = std::chrono::steady_clock::now();
        stats->closeSeconds = std::chrono::duration<double>(callEnd - r1).count();
        stats->totalSeconds = std::chrono::duration<double>(callEnd - callStart).count();
    }
    catch(const MRayError& e)
    {
        Fail(err, errLen, "MRayError: "s + e.GetError());
        return 3;
    }
    catch(const std::exception& e)
    {
        Fail(err, errLen, "exception: "s + e.what());
        return 4;
    }
    return 0;
}

} // extern "C"
