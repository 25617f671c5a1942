// tracer_b200.cpp — the TracerDLL plugin object: `TracerI` (Core/TracerI.h:L150-376) implemented on
// top of the C-ABI of include/mray_b200.h. Exports the two symbols MRay's TracerThread resolves
// (TracerDLL/EntryPoint.h:L15-18). Host-side C++20, compiled against the reference's own headers so
// that the vtable order and the layouts of TransientData / StaticVector / Optional match.
//
// Scope (everything else throws MRayError, as the reference does for unknown types):
//   (P)Triangle  (Mt)Lambert [constant or textured albedo]  (Mt)Reflect  (L)Prim(P)Triangle [constant radiance,
//   isTwoSided]  (L)Null  (T)Identity  (T)Single  (C)Pinhole  (Md)Vacuum  (R)PathTracerRGB  (R)PathTracerSpectral,
//   single-level RGBA textures, samplerType Independent / Sobol / ZSobol, render regions, Throughput / Latency modes,
//   SetCameraTransform — i.e. BASELINE config 1 / 3 / 4-style scenes.
// Identity-transform surfaces are flattened into ONE accelerator (one prim range per prim-batch / material pair);
// scenes with (T)Single transforms become two-level scenes whose instances share accelerators where their prim
// batches and cull flags agree; ids are Key-typed bit casts like the reference's (Tracer/Key.h).
#include "Core/TracerI.h"
#include "Core/Error.h"
#include "Core/TimelineSemaphore.h"
#include "Core/Quaternion.h"
#include "TransientPool/TransientPool.h"
#include "Core/ColorFunctions.h"
#include "Core/System.h"
#include "mray_b200.h"
#include <dlfcn.h>
#include <filesystem>
#include <fstream>
#include <cstdlib>
#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <array>
#include <cmath>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace
{

using namespace std::string_view_literals;

constexpr uint32_t PRIM_ID_BITS = 28, MAT_ID_BITS = 21, TRANS_ID_BITS = 24, CAM_ID_BITS = 24;
template<class Id> uint32_t Raw(Id id) { return static_cast<uint32_t>(id); }

struct PrimBatch { uint32_t primCount, vertexCount, primOffset, vertexOffset; };
struct PrimGroupB200
{
    std::string type; bool committed = false;
    std::vector<PrimBatch> batches;
    std::vector<Vector3> positions; std::vector<Vector2> uvs; std::vector<Vector3ui> indices;
    std::vector<std::array<float, 4>> tbn;   // NORMAL attribute: world -> tangent-space quaternions (w, x, y, z); all zero = not pushed
    bool hasTBN = false;
    uint32_t primTotal = 0, vertexTotal = 0;
};
struct MatGroupB200
{
    std::string type; bool committed = false; std::vector<Vector3> albedo; std::vector<int32_t> albedoTex; /* TextureId or -1 */ std::vector<int32_t> normalTex; /* TextureId or -1 */
    // constant attributes of (Mt)Refract {cauchyFront xyz, -, cauchyBack xyz, -} and (Mt)Unreal {roughness, specular, metallic, ...}
    std::vector<std::array<float, 8>> params;
};
// One 2-D texture as TracerI::CreateTexture2D / PushTextureData deliver it (first slice: one mip level, 4-channel
// fp32 or unorm8 pixels, already in the global colour space)
struct TextureB200
{
    Vector2ui size; MRayTextureParameters params; uint32_t channels = 4, format = 0;
    std::vector<Byte> pixels; bool loaded = false;   // all levels back to back (Graphics::TextureMipPixelStart order); loaded = level 0 arrived
    float gamma = 1.0f; bool hasColorMatrix = false; float colorMatrix[9] = {};   // TextureMemory::ConvertColorspaces, done on upload
    uint32_t mipCount = 1; std::vector<uint8_t> levelLoaded;   // levels the caller reserved (CreateTexture2D) and pushed so far
    bool genMips = false; uint32_t mipFilterType = 2; float mipFilterRadius = 2.0f;   // TracerParameters.genMips / mipGenFilter
    uint32_t clampRes = 0;                                                             // TracerParameters.clampedTexRes (0 = ignoreResClamp)
    size_t TexelBytes() const { return size_t(channels) * (format == 0 ? 4u : 1u); }
    static uint32_t LevelDim(uint32_t n, uint32_t level) { return std::max(n >> level, 1u); }
    size_t LevelStart(uint32_t level) const
    { size_t o = 0; for(uint32_t i = 0; i < level; i++) o += size_t(LevelDim(size[0], i)) * LevelDim(size[1], i); return o; }
    // the leading run of pushed levels is what the library is given; with genMips it filters the rest itself (KCGenerateMipmaps
    // skips the levels that arrived, Tracer/TextureFilter.cu:L158-167)
    uint32_t SuppliedLevels() const { uint32_t n = 0; while(n < mipCount && levelLoaded[n]) n++; return n; }
    mrb_texture_desc Desc(bool convertColor = true) const
    {
        const uint32_t supplied = SuppliedLevels();
        for(uint32_t l = supplied; l < mipCount; l++)
            if(levelLoaded[l]) throw MRayError("textures: mip level {} was pushed but level {} was not", l, supplied);
        if(!genMips && supplied != mipCount) throw MRayError("textures: {} of {} mip levels have no data", mipCount - supplied, mipCount);
        mrb_texture_desc d = {};
        d.data = pixels.data(); d.width = size[0]; d.height = size[1]; d.channels = channels; d.format = format;
        d.interp = uint32_t(params.interpolation); d.edge = uint32_t(params.edgeResolve);
        d.gamma = convertColor ? gamma : 1.0f; d.colorMatrix = (convertColor && hasColorMatrix) ? colorMatrix : nullptr;
        d.mipCount = supplied; d.generateMips = genMips ? 1u : 0u; d.mipFilterType = mipFilterType; d.mipFilterRadius = mipFilterRadius;
        d.clampResolution = clampRes;
        return d;
    }
};
struct LightGroupB200
{
    std::string type; bool committed = false; uint32_t primGroup = 0;
    std::vector<Vector3> radiance; std::vector<uint8_t> twoSided; std::vector<uint32_t> primBatch;
    std::vector<int32_t> radianceTex;   // skyspheres: TextureId of the radiance map or -1 (constant)
    bool IsSkysphere() const { return type == "(L)Skysphere_Spherical" || type == "(L)Skysphere_CoOcta"; }
};
struct TransGroupB200 { std::string type; bool committed = false; std::vector<Matrix3x4> matrices; };
struct CamGroupB200 { std::string type; bool committed = false; std::vector<Vector4> fovPlanes; std::vector<Vector3> gaze, position, up; };
struct RendererB200
{
    std::string type;
    uint32_t totalSPP = 16384, burstSize = 1, sampleMode = 0; Vector2ui rrRange = Vector2ui(4, 20);
    bool latency = false;      // renderMode "Latency": every DoRenderWork finishes burstSize samples per pixel
};

// Color::Colorspace<E>::ToXYZMatrix of the tracer's global texture colour space (KCExtractLuminance's InvokeAt over the
// ColorspaceList, Tracer/ColorConverter.cu:L70-78,L459-468)
Matrix3x3 LuminanceMatrix(MRayColorSpaceEnum e)
{
    using enum MRayColorSpaceEnum;
    switch(e)
    {
        case MR_ACES2065_1: return Color::Colorspace<MR_ACES2065_1>::ToXYZMatrix;
        case MR_ACES_CG:    return Color::Colorspace<MR_ACES_CG>::ToXYZMatrix;
        case MR_REC_709:    return Color::Colorspace<MR_REC_709>::ToXYZMatrix;
        case MR_REC_2020:   return Color::Colorspace<MR_REC_2020>::ToXYZMatrix;
        case MR_DCI_P3:     return Color::Colorspace<MR_DCI_P3>::ToXYZMatrix;
        case MR_ADOBE_RGB:  return Color::Colorspace<MR_ADOBE_RGB>::ToXYZMatrix;
        default: throw MRayError("skysphere luminance: unsupported global texture colour space");
    }
}

// Inverse of an affine 3x4 matrix by Laplace expansion, the arithmetic of Matrix3x4T::Inverse
// (Core/Matrix.hpp:L853-903) term for term, EXCEPT element (1,2): the reference writes +s1 where the
// cofactor is -s1 (Matrix.hpp:L892), which only cancels when m00*m12 == m02*m10 (axis-aligned scales,
// rotations about z) and otherwise yields a matrix that is not the inverse. The corrected sign is used
// here; for every matrix the reference inverts correctly the result is bit-identical.
std::array<float, 12> InverseAffine(const Matrix3x4& mat)
{
    float m[12];
    for(unsigned r = 0; r < 3; r++) for(unsigned c = 0; c < 4; c++) m[4 * r + c] = mat(r, c);
    auto Det2x2 = [](float m00, float m01, float m10, float m11) { return std::fma(m00, m11, -m01 * m10); };
    float s0 = Det2x2(m[0], m[1], m[4], m[5]), s1 = Det2x2(m[0], m[2], m[4], m[6]), s2 = Det2x2(m[0], m[3], m[4], m[7]);
    float s3 = Det2x2(m[1], m[2], m[5], m[6]), s4 = Det2x2(m[1], m[3], m[5], m[7]), s5 = Det2x2(m[2], m[3], m[6], m[7]);
    float c5 = m[10], c4 = m[9], c2 = m[8];
    float det = ((s0 * c5) - (s1 * c4) + (s3 * c2));
    float detInv = 1.0f / det;
    std::array<float, 12> inv =
    {
        (+m[5] * c5 - m[6] * c4), (-m[1] * c5 + m[2] * c4), s3, (-m[9] * s5 + m[10] * s4 - m[11] * s3),
        (-m[4] * c5 + m[6] * c2), (+m[0] * c5 - m[2] * c2), -s1, (+m[8] * s5 - m[10] * s2 + m[11] * s1),
        (+m[4] * c4 - m[5] * c2), (-m[0] * c4 + m[1] * c2), s0, (-m[8] * s4 + m[9] * s2 - m[11] * s0)
    };
    for(float& v : inv) v *= detInv;
    return inv;
}

// Everything that lives on ONE GPU. The reference renders on GPUSystem::BestDevice() only
// (Device/CUDA/GPUSystemCUDA.cpp:L287-289); here one TracerB200 drives `MRB_DEVICES` GPUs of the box:
// scene buffers, BVHs and LUTs are replicated per device, every pass's sample range is split over the devices, and
// the per-device films are summed into device 0's over NVLink peer memory before the single RenderImageSection
// is handed to the caller.
struct DeviceB200
{
    mrb_context ctx = nullptr;
    mrb_accel accel = nullptr;           // all (T)Identity surfaces
    std::vector<mrb_accel> instAccels;   // two-level scenes: one per distinct (prim ranges, cull flags) set
    mrb_scene scene = nullptr;           // set when any surface is transformed
    mrb_renderer renderer = nullptr;
    mrb_spectrum spectrum = nullptr;     // SpectrumContextJakob2019 of params.globalTextureColorSpace, made on first use
};

class TracerB200 final : public TracerI
{
    TracerParameters params;
    std::vector<DeviceB200> devs;        // devs[0] owns the film hand-off
    mrb_context ctx = nullptr;           // = devs[0].ctx
    uint32_t sceneInstanceCount = 0, uniqueAccelCount = 0;
    float sceneDiameter = 0.0f;          // what TracerBase hands to LightGroup::SetSceneDiameter at CommitSurfaces
    bool committed = false, twoLevel = false;
    std::mutex mtx; // scene-loading calls arrive concurrently from pool threads (TracerBase.h:L97-130)

    std::vector<PrimGroupB200> prims; std::vector<MatGroupB200> mats; std::vector<LightGroupB200> lights;
    std::vector<CamGroupB200> cams; std::vector<RendererB200> renderers; std::vector<TransGroupB200> transforms;
    std::vector<SurfaceParams> surfaces; std::vector<LightSurfaceParams> lightSurfaces;
    std::vector<CameraSurfaceParams> camSurfaces; std::vector<VolumeParams> volumes;
    LightSurfaceParams boundary{};
    // flattened scene
    uint32_t flatPrimGroup = 0; std::vector<float> flatLightRadiance; std::vector<uint8_t> flatLightTwoSided;
    std::vector<float> flatAlbedo;
    std::vector<uint8_t> flatMaterialType;          // per flat material: mrb_material_type
    std::vector<float> flatMaterialParams;          // per flat material: 8 floats (mrb_render_desc.materialParams)
    std::vector<int32_t> flatAlbedoTex;             // per flat material: index into flatTextures or -1
    std::vector<int32_t> flatNormalTex;             // per flat material: normal map (index into flatTextures) or -1
    std::vector<uint32_t> flatTextures;             // TextureIds in use, in first-use order
    std::vector<TextureB200> textures;              // TextureId = index + 1 (0 = InvalidTexture)
    // render hand-off: pinned staging the caller reads between semaphore acquire / release
    TimelineSemaphore* sem = nullptr; uint64_t acquireValue = 0;
    float* staging = nullptr; size_t stagingBytes = 0;
    struct StartArgs { RendererId id; CamSurfaceId camSurf; RenderImageParams rip; Optional<uint32_t> logic0; };
    Optional<StartArgs> lastStart; Optional<CameraTransform> camOverride, pendingCam; bool rebuilding = false;
    uint32_t curRenderer = 0; ThreadPool* pool = nullptr;
    // ImageTiler state (Tracer/RenderImage.h:L56-110): the region is cut into tiles of at most ~parallelizationHint pixels
    Vector2ui fullResolution = Vector2ui::Zero(), regionMin = Vector2ui::Zero(), regionSize = Vector2ui::Zero();
    Vector2ui coveringTile = Vector2ui::Zero(), tileCount = Vector2ui::Zero();
    uint32_t currentTile = 0; std::vector<uint32_t> tileSPPs;
    // the sample range [jobBegin, jobEnd) of totalSPP this process renders (MRB_SPP_SHARD = "rank/world": sample-range
    // sharding across processes, the caller sums the images); a single process renders [0, totalSPP)
    uint32_t jobBegin = 0, jobEnd = 0;
    bool throughputSingle = false, saveImage = true;
    uint64_t completedPaths = 0;
    // one burst pass: samples [tileSPP, tileSPP + count) of tile `tile`
    struct PassPlan { uint32_t tile = 0, tileSPP = 0, count = 0; bool operator==(const PassPlan&) const = default; };
    bool primed = false; PassPlan primedPlan;   // the next pass has been begun (and its first iterations queued) already

    void Check(mrb_status s) const { if(s != MRB_OK) throw MRayError("{}", mrb_last_error(ctx)); }
    static void CheckOn(const DeviceB200& d, mrb_status s) { if(s != MRB_OK) throw MRayError("{}", mrb_last_error(d.ctx)); }
    template<class G> G& Get(std::vector<G>& v, uint32_t id, std::string_view what)
    { if(id >= v.size()) throw MRayError("Unable to find {}({})", what, id); return v[id]; }
    template<class G> const G& Get(const std::vector<G>& v, uint32_t id, std::string_view what) const
    { if(id >= v.size()) throw MRayError("Unable to find {}({})", what, id); return v[id]; }

    public:
    explicit TracerB200(const TracerParameters& p) : params(p)
    {
        if(mrb_abi_version() != MRB_ABI_VERSION)   // descriptor layouts must match the header this object was compiled against
            throw MRayError("mray_b200: libmray_b200.so has ABI {:#x}, the plugin was built for {:#x}", mrb_abi_version(), uint32_t(MRB_ABI_VERSION));
        // MRB_DEVICES = "N" (devices 0..N-1) or a comma list; default: one device, MRB_DEVICE (0)
        std::vector<int> ids;
        if(const char* e = std::getenv("MRB_DEVICES"))
        {
            const std::string v(e);
            if(v.find(',') == std::string::npos) { int n = std::atoi(e); for(int k = 0; k < n; k++) ids.push_back(k); }
            else for(size_t b = 0; b < v.size();) { size_t c = v.find(',', b); if(c == std::string::npos) c = v.size(); ids.push_back(std::atoi(v.substr(b, c - b).c_str())); b = c + 1; }
        }
        if(ids.empty()) { const char* e = std::getenv("MRB_DEVICE"); ids.push_back(e ? std::atoi(e) : 0); }
        for(int id : ids)
        {
            DeviceB200 d;
            mrb_status s = mrb_context_create(id, &d.ctx);
            if(s != MRB_OK)
            {
                const std::string msg = mrb_last_error(nullptr);
                for(DeviceB200& o : devs) mrb_context_destroy(o.ctx);
                throw MRayError("mray_b200: {}", msg); // no CPU fallback
            }
            devs.push_back(d);
        }
        ctx = devs[0].ctx;
        // implicit groups, id 0 of every kind (Core/TracerI.h:L99-124)
        prims.push_back(PrimGroupB200{std::string(TracerConstants::EmptyPrimName), true});
        mats.push_back(MatGroupB200{std::string(TracerConstants::PassthroughMatName), true});
        lights.push_back(LightGroupB200{std::string(TracerConstants::NullLightName), true});
        lights[0].radiance.push_back(Vector3::Zero()); lights[0].twoSided.push_back(0); lights[0].primBatch.push_back(0);
        transforms.push_back(TransGroupB200{std::string(TracerConstants::IdentityTransName), true});
        transforms[0].matrices.push_back(Matrix3x4::Identity());
    }
    // SpectrumContextJakob2019 ctor (Tracer/SpectrumContext.cu:L354-532): LUT file + normalised CIE tables
    void EnsureSpectrum()
    {
        if(devs[0].spectrum) return;
        const MRayColorSpaceEnum cs = params.globalTextureColorSpace;
        const std::string fileName = std::string(MRayColorSpaceStringifier::ToString(cs)) + std::string(Color::LUT_FILE_EXT);
        namespace fs = std::filesystem;
        std::vector<fs::path> candidates;
        if(const char* e = std::getenv("MRB_SPECTRA_LUT_DIR")) candidates.push_back(fs::path(e) / fileName);
        candidates.push_back(fs::path(GetProcessPath()) / fs::path(Color::LUT_FOLDER_NAME) / fileName);   // the reference's location
        Dl_info info;
        if(dladdr(reinterpret_cast<const void*>(&InverseAffine), &info) && info.dli_fname)
            candidates.push_back(fs::path(info.dli_fname).parent_path().parent_path() / "data" / fileName);
        fs::path found;
        for(const fs::path& c : candidates) if(fs::exists(c)) { found = c; break; }
        if(found.empty()) throw MRayError("Unable to open spectra lut file {}!", candidates[candidates.size() > 1 ? 1 : 0].string());
        std::ifstream f(found, std::ios_base::binary);
        char cc[10]; uint32_t res = 0, mode = 0;
        f.read(cc, 10);
        if(!f || std::string_view(cc, 10) != Color::LUT_FILE_CC) throw MRayError("Wrong character code in .mrspectra file!");
        f.read(reinterpret_cast<char*>(&res), 4);
        if(!f || res != 64) throw MRayError("Wrong size ({}), spectra lut size must be {}!", res, 64);
        f.read(reinterpret_cast<char*>(&mode), 4);
        if(!f || mode != 1) throw MRayError("Wrong mode ({}), spectra lut must have mode \"1\" (aka. Float)!", mode);
        std::vector<float> lut(size_t(9) * 64 * 64 * 64);
        if(!f.read(reinterpret_cast<char*>(lut.data()), std::streamsize(lut.size() * sizeof(float)))) throw MRayError("Unable to read sprectum lut file!");
        std::vector<float> observer(Color::CIE_1931_N * 3), illum(Color::CIE_1931_N);
        const auto& spd = Color::SelectIlluminantSPD(cs);
        Float illumNorm = Color::CIE_1931_Y_INTEGRAL;
        illumNorm /= Color::SelectIlluminantSPDNormFactor(cs);
        const Vector3 weight = Vector3(1) / Vector3(Color::CIE_1931_X_INTEGRAL, Color::CIE_1931_Y_INTEGRAL, Color::CIE_1931_Z_INTEGRAL);
        for(uint32_t i = 0; i < Color::CIE_1931_N; i++)
        {
            Vector3 v = Color::CIE_1931_XYZ[i] * weight;
            observer[3 * i] = v[0]; observer[3 * i + 1] = v[1]; observer[3 * i + 2] = v[2];
            illum[i] = spd[i] * illumNorm;
        }
        const Matrix3x3 xyzToRGB = Color::SelectRGBToXYZMatrix(cs).Inverse();
        mrb_spectrum_desc d = {};
        d.lut = lut.data(); d.lutResolution = res; d.observerXYZ = observer.data(); d.illuminantSPD = illum.data();
        for(unsigned i = 0; i < 9; i++) d.xyzToRGB[i] = xyzToRGB[i];
        d.wavelengthSampleMode = uint32_t(params.wavelengthSampleMode.e);
        for(DeviceB200& dv : devs) CheckOn(dv, mrb_spectrum_create(dv.ctx, &d, &dv.spectrum));
    }
    // SobolDetail::SobolMatrices as data (mray_b200/data/sobol_matrices.bin, or $MRB_DATA_DIR)
    std::vector<uint32_t> sobolMatrices;
    void EnsureSobolMatrices()
    {
        if(!sobolMatrices.empty()) return;
        namespace fs = std::filesystem;
        std::vector<fs::path> candidates;
        if(const char* e = std::getenv("MRB_DATA_DIR")) candidates.push_back(fs::path(e) / "sobol_matrices.bin");
        Dl_info info;
        if(dladdr(reinterpret_cast<const void*>(&InverseAffine), &info) && info.dli_fname)
            candidates.push_back(fs::path(info.dli_fname).parent_path().parent_path() / "data" / "sobol_matrices.bin");
        for(const fs::path& c : candidates)
        {
            std::ifstream f(c, std::ios_base::binary);
            if(!f) continue;
            sobolMatrices.resize(size_t(256) * 52);
            if(f.read(reinterpret_cast<char*>(sobolMatrices.data()), std::streamsize(sobolMatrices.size() * 4))) return;
            sobolMatrices.clear();
        }
        throw MRayError("Unable to open the Sobol generator matrices (sobol_matrices.bin)");
    }
    // MRB_PLUGIN_TIMING=1: wall time of the tear-down steps on stderr
    struct ScopeTimer
    {
        const char* what; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        explicit ScopeTimer(const char* w) : what(w) {}
        ~ScopeTimer()
        {
            static const bool on = std::getenv("MRB_PLUGIN_TIMING") != nullptr;
            if(on) std::fprintf(stderr, "[mray_b200 plugin] %s: %.2f ms\n", what, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
    };
    void ReleaseRenderers()
    {
        ScopeTimer t("ReleaseRenderers");
        for(DeviceB200& d : devs) if(d.renderer) { mrb_renderer_destroy(d.ctx, d.renderer); d.renderer = nullptr; }
    }
    void ReleaseAccels()
    {
        for(DeviceB200& d : devs)
        {
            if(d.scene) { mrb_scene_destroy(d.ctx, d.scene); d.scene = nullptr; }
            for(mrb_accel a : d.instAccels) mrb_accel_destroy(d.ctx, a);
            d.instAccels.clear();
            if(d.accel) { mrb_accel_destroy(d.ctx, d.accel); d.accel = nullptr; }
        }
        committed = false; twoLevel = false; sceneInstanceCount = 0; uniqueAccelCount = 0;
    }
    ~TracerB200() override
    {
        ReleaseRenderers();
        { ScopeTimer t("free staging"); if(staging) mrb_host_free(ctx, staging); }
        { ScopeTimer t("free spectrum"); for(DeviceB200& d : devs) if(d.spectrum) mrb_spectrum_destroy(d.ctx, d.spectrum); }
        { ScopeTimer t("ReleaseAccels"); ReleaseAccels(); }
        { ScopeTimer t("context destroy"); for(DeviceB200& d : devs) mrb_context_destroy(d.ctx); }
    }

    // ------------------------------- generic -------------------------------
    TypeNameList PrimitiveGroups() const override { return {"(P)Triangle"sv, "(P)Empty"sv}; }
    TypeNameList MaterialGroups() const override { return {"(Mt)Lambert"sv, "(Mt)Passthrough"sv, "(Mt)Reflect"sv, "(Mt)Refract"sv, "(Mt)Unreal"sv}; }
    TypeNameList TransformGroups() const override { return {"(T)Identity"sv, "(T)Single"sv}; }
    TypeNameList CameraGroups() const override { return {"(C)Pinhole"sv}; }
    TypeNameList MediumGroups() const override { return {"(Md)Vacuum"sv}; }
    TypeNameList LightGroups() const override { return {"(L)Null"sv, "(L)Prim(P)Triangle"sv}; }
    TypeNameList Renderers() const override { return {"(R)PathTracerRGB"sv, "(R)PathTracerSpectral"sv}; }

    PrimAttributeInfoList AttributeInfoPrim(std::string_view name) const override
    {
        using enum MRayDataEnum; using enum AttributeIsArray; using enum AttributeOptionality;
        using L = PrimitiveAttributeLogic;
        if(name != "(P)Triangle"sv) return {};
        // Tracer/PrimitiveDefaultTriangle.cu:L169-184
        return PrimAttributeInfoList
        {
            PrimAttributeInfo(L::POSITION, MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY),
            PrimAttributeInfo(L::NORMAL,   MRayDataTypeRT(MR_QUATERNION), IS_SCALAR, MR_OPTIONAL),
            PrimAttributeInfo(L::UV0,      MRayDataTypeRT(MR_VECTOR_2), IS_SCALAR, MR_OPTIONAL),
            PrimAttributeInfo(L::INDEX,    MRayDataTypeRT(MR_VECTOR_3UI), IS_SCALAR, MR_MANDATORY)
        };
    }
    MatAttributeInfoList AttributeInfoMat(std::string_view name) const override
    {
        using enum MRayDataEnum; using enum AttributeIsArray; using enum AttributeOptionality; using enum AttributeTexturable; using enum AttributeIsColor;
        if(name == "(Mt)Refract"sv)   // Tracer/MaterialsDefault.cpp:L236-251
            return MatAttributeInfoList
            {
                MatAttributeInfo("cauchyBack", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY, MR_CONSTANT_ONLY, IS_PURE_DATA),
                MatAttributeInfo("cauchyFront", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY, MR_CONSTANT_ONLY, IS_PURE_DATA)
            };
        if(name == "(Mt)Unreal"sv)    // Tracer/MaterialsDefault.cpp:L382-403
            return MatAttributeInfoList
            {
                MatAttributeInfo("albedo", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_COLOR),
                MatAttributeInfo("normalMap", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_OPTIONAL, MR_TEXTURE_ONLY, IS_PURE_DATA),
                MatAttributeInfo("roughness", MRayDataTypeRT(MR_FLOAT), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_PURE_DATA),
                MatAttributeInfo("specular", MRayDataTypeRT(MR_FLOAT), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_PURE_DATA),
                MatAttributeInfo("metallic", MRayDataTypeRT(MR_FLOAT), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_PURE_DATA)
            };
        if(name != "(Mt)Lambert"sv) return {};
        return MatAttributeInfoList
        {
            MatAttributeInfo("albedo", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_COLOR),
            MatAttributeInfo("normalMap", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_OPTIONAL, MR_TEXTURE_ONLY, IS_PURE_DATA)
        };
    }
    LightAttributeInfoList AttributeInfoLight(std::string_view name) const override
    {
        using enum MRayDataEnum; using enum AttributeIsArray; using enum AttributeOptionality; using enum AttributeTexturable; using enum AttributeIsColor;
        if(name == "(L)Skysphere_Spherical"sv || name == "(L)Skysphere_CoOcta"sv)
            return LightAttributeInfoList // Tracer/LightsDefault.hpp:L724-738
            {
                LightAttributeInfo("radiance", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_COLOR)
            };
        if(name != "(L)Prim(P)Triangle"sv) return {};
        return LightAttributeInfoList // Tracer/LightsDefault.hpp:L547-562
        {
            LightAttributeInfo("radiance", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY, MR_TEXTURE_OR_CONSTANT, IS_COLOR),
            LightAttributeInfo("isTwoSided", MRayDataTypeRT(MR_BOOL), IS_SCALAR, MR_MANDATORY, MR_CONSTANT_ONLY, IS_PURE_DATA)
        };
    }
    CamAttributeInfoList AttributeInfoCam(std::string_view name) const override
    {
        using enum MRayDataEnum; using enum AttributeIsArray; using enum AttributeOptionality;
        if(name != "(C)Pinhole"sv) return {};
        return CamAttributeInfoList
        {
            CamAttributeInfo("FovAndPlanes", MRayDataTypeRT(MR_VECTOR_4), IS_SCALAR, MR_MANDATORY),
            CamAttributeInfo("gaze", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY),
            CamAttributeInfo("position", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY),
            CamAttributeInfo("up", MRayDataTypeRT(MR_VECTOR_3), IS_SCALAR, MR_MANDATORY)
        };
    }
    MediumAttributeInfoList AttributeInfoMedium(std::string_view) const override { return {}; }
    TransAttributeInfoList AttributeInfoTrans(std::string_view name) const override
    {
        using enum MRayDataEnum; using enum AttributeIsArray; using enum AttributeOptionality;
        if(name != "(T)Single"sv) return {};
        // Tracer/TransformsDefault.cu:L27-37 (declared 4x4; the loader pushes Matrix3x4, SceneLoaderMRay.cpp:L492-576)
        return TransAttributeInfoList{TransAttributeInfo("Transform", MRayDataTypeRT(MR_MATRIX_4x4), IS_SCALAR, MR_MANDATORY)};
    }
    RendererAttributeInfoList AttributeInfoRenderer(std::string_view name) const override
    {
        using enum MRayDataEnum; using enum AttributeIsArray; using enum AttributeOptionality;
        if(name != "(R)PathTracerRGB"sv && name != "(R)PathTracerSpectral"sv) return {};
        return RendererAttributeInfoList // TracerDLL/PathTracerRenderer.cu:L1384-1400
        {
            RendererAttributeInfo("totalSPP", MRayDataTypeRT(MR_UINT32), IS_SCALAR, MR_MANDATORY),
            RendererAttributeInfo("burstSize", MRayDataTypeRT(MR_UINT32), IS_SCALAR, MR_OPTIONAL),
            RendererAttributeInfo("renderMode", MRayDataTypeRT(MR_STRING), IS_SCALAR, MR_MANDATORY),
            RendererAttributeInfo("sampleMode", MRayDataTypeRT(MR_STRING), IS_SCALAR, MR_MANDATORY),
            RendererAttributeInfo("rrRange", MRayDataTypeRT(MR_VECTOR_2UI), IS_SCALAR, MR_MANDATORY),
            RendererAttributeInfo("neeSamplerType", MRayDataTypeRT(MR_STRING), IS_SCALAR, MR_MANDATORY)
        };
    }
    PrimAttributeInfoList AttributeInfo(PrimGroupId id) const override { return AttributeInfoPrim(Get(prims, Raw(id), "PrimitiveGroup").type); }
    CamAttributeInfoList AttributeInfo(CameraGroupId id) const override { return AttributeInfoCam(Get(cams, Raw(id), "CameraGroup").type); }
    MediumAttributeInfoList AttributeInfo(MediumGroupId) const override { return {}; }
    MatAttributeInfoList AttributeInfo(MatGroupId id) const override { return AttributeInfoMat(Get(mats, Raw(id), "MaterialGroup").type); }
    TransAttributeInfoList AttributeInfo(TransGroupId id) const override { return AttributeInfoTrans(Get(transforms, Raw(id), "TransformGroup").type); }
    LightAttributeInfoList AttributeInfo(LightGroupId id) const override { return AttributeInfoLight(Get(lights, Raw(id), "LightGroup").type); }
    RendererAttributeInfoList AttributeInfo(RendererId id) const override { return AttributeInfoRenderer(Get(renderers, Raw(id), "Renderer").type); }
    std::string TypeName(PrimGroupId id) const override { return Get(prims, Raw(id), "PrimitiveGroup").type; }
    std::string TypeName(CameraGroupId id) const override { return Get(cams, Raw(id), "CameraGroup").type; }
    std::string TypeName(MediumGroupId) const override { return std::string(TracerConstants::VacuumMediumName); }
    std::string TypeName(MatGroupId id) const override { return Get(mats, Raw(id), "MaterialGroup").type; }
    std::string TypeName(TransGroupId id) const override { return Get(transforms, Raw(id), "TransformGroup").type; }
    std::string TypeName(LightGroupId id) const override { return Get(lights, Raw(id), "LightGroup").type; }
    std::string TypeName(RendererId id) const override { return Get(renderers, Raw(id), "Renderer").type; }

    // ------------------------------- primitives -------------------------------
    PrimGroupId CreatePrimitiveGroup(std::string typeName) override
    {
        std::lock_guard lk(mtx);
        if(typeName != "(P)Triangle") throw MRayError("Unable to find generator for {}", typeName);
        prims.push_back(PrimGroupB200{typeName});
        return PrimGroupId(uint32_t(prims.size() - 1));
    }
    PrimBatchId ReservePrimitiveBatch(PrimGroupId g, PrimCount c) override { return ReservePrimitiveBatches(g, {c}).front(); }
    PrimBatchIdList ReservePrimitiveBatches(PrimGroupId g, std::vector<PrimCount> counts) override
    {
        std::lock_guard lk(mtx);
        PrimGroupB200& pg = Get(prims, Raw(g), "PrimitiveGroup");
        if(pg.committed) throw MRayError("{}: reservations are already committed", pg.type);
        PrimBatchIdList out;
        for(const PrimCount& c : counts)
        {
            pg.batches.push_back(PrimBatch{c.primCount, c.attributeCount, pg.primTotal, pg.vertexTotal});
            pg.primTotal += c.primCount; pg.vertexTotal += c.attributeCount;
            out.push_back(PrimBatchId((Raw(g) << PRIM_ID_BITS) | uint32_t(pg.batches.size() - 1)));
        }
        return out;
    }
    void CommitPrimReservations(PrimGroupId g) override
    {
        std::lock_guard lk(mtx);
        PrimGroupB200& pg = Get(prims, Raw(g), "PrimitiveGroup");
        pg.positions.assign(pg.vertexTotal, Vector3::Zero());
        pg.tbn.assign(pg.vertexTotal, std::array<float, 4>{1.f, 0.f, 0.f, 0.f});
        pg.uvs.assign(pg.vertexTotal, Vector2::Zero());
        pg.indices.assign(pg.primTotal, Vector3ui::Zero());
        pg.committed = true;
    }
    bool IsPrimCommitted(PrimGroupId g) const override { return Get(prims, Raw(g), "PrimitiveGroup").committed; }
    void PushPrimAttribute(PrimGroupId g, PrimBatchId b, uint32_t attributeIndex, TransientData data) override
    {
        PrimGroupB200& pg = Get(prims, Raw(g), "PrimitiveGroup");
        uint32_t bi = Raw(b) & ((1u << PRIM_ID_BITS) - 1u);
        if(!pg.committed || bi >= pg.batches.size()) throw MRayError("{}: unknown / uncommitted batch {}", pg.type, bi);
        const PrimBatch& pb = pg.batches[bi];
        switch(attributeIndex)
        {
            case 0: { auto s = data.AccessAs<const Vector3>(); if(s.size() != pb.vertexCount) throw MRayError("position count mismatch");
                      std::copy(s.begin(), s.end(), pg.positions.begin() + pb.vertexOffset); break; }
            case 1: { auto s = data.AccessAs<const Quaternion>(); if(s.size() != pb.vertexCount) throw MRayError("normal count mismatch");
                      // the attribute is the to-tangent-space rotation; the renderer blends the three vertex quaternions of a
                      // hit (Quaternion::BarySLerp) and takes the frame's Z axis as the shading normal
                      for(size_t i = 0; i < s.size(); i++)
                      { const Quaternion& q = s[i]; pg.tbn[pb.vertexOffset + i] = {q[0], q[1], q[2], q[3]}; }
                      pg.hasTBN = true; break; }
            case 2: { auto s = data.AccessAs<const Vector2>(); if(s.size() != pb.vertexCount) throw MRayError("uv count mismatch");
                      std::copy(s.begin(), s.end(), pg.uvs.begin() + pb.vertexOffset); break; }
            case 3: { auto s = data.AccessAs<const Vector3ui>(); if(s.size() != pb.primCount) throw MRayError("index count mismatch");
                      // KCAdjustIndices (Tracer/PrimitiveDefaultTriangle.cu:L9): rebase batch-local indices
                      for(size_t i = 0; i < s.size(); i++) pg.indices[pb.primOffset + i] = s[i] + Vector3ui(pb.vertexOffset); break; }
            default: throw MRayError("{}: unknown attribute index {}", pg.type, attributeIndex);
        }
    }
    void PushPrimAttribute(PrimGroupId, PrimBatchId, uint32_t, Vector2ui, TransientData) override
    { throw MRayError("(P)Triangle: sub-batch attribute push is not supported"); }
    void TransformPrimitives(PrimGroupId, std::vector<PrimBatchId>, std::vector<Matrix3x4>) override
    { throw MRayError("TransformPrimitives is not supported"); }

    // ------------------------------- materials -------------------------------
    MatGroupId CreateMaterialGroup(std::string typeName) override
    {
        std::lock_guard lk(mtx);
        // (Mt)Reflect has no attributes (MatGroupReflect::AttributeInfo returns an empty list, MaterialsDefault.cpp:L161-164)
        if(typeName != "(Mt)Lambert" && typeName != "(Mt)Reflect" && typeName != "(Mt)Refract" && typeName != "(Mt)Unreal")
            throw MRayError("Unable to find generator for {}", typeName);
        mats.push_back(MatGroupB200{typeName});
        return MatGroupId(uint32_t(mats.size() - 1));
    }
    MaterialId ReserveMaterial(MatGroupId g, AttributeCountList c) override { return ReserveMaterials(g, {c}).front(); }
    MaterialIdList ReserveMaterials(MatGroupId g, std::vector<AttributeCountList> counts) override
    {
        std::lock_guard lk(mtx);
        MatGroupB200& mg = Get(mats, Raw(g), "MaterialGroup");
        MaterialIdList out;
        for(size_t i = 0; i < counts.size(); i++)
        {
            mg.albedo.push_back(Vector3::Zero()); mg.albedoTex.push_back(-1); mg.normalTex.push_back(-1); mg.params.push_back(std::array<float, 8>{});
            out.push_back(MaterialId((Raw(g) << MAT_ID_BITS) | uint32_t(mg.albedo.size() - 1)));
        }
        return out;
    }
    void CommitMatReservations(MatGroupId g) override { Get(mats, Raw(g), "MaterialGroup").committed = true; }
    bool IsMatCommitted(MatGroupId g) const override { return Get(mats, Raw(g), "MaterialGroup").committed; }
    void PushMatAttribute(MatGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data) override
    {
        MatGroupB200& mg = Get(mats, Raw(g), "MaterialGroup");
        if(mg.type != "(Mt)Refract") throw MRayError("{}: Attribute {:d} is not \"ConstantOnly\", wrong function is called", mg.type, attributeIndex);
        // MatGroupRefract::PushAttribute (Tracer/MaterialsDefault.cpp:L295-314): 0 = cauchyBack, 1 = cauchyFront
        if(attributeIndex > 1) throw MRayError("{:s}: Unkown attribute index {:d}", mg.type, attributeIndex);
        uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
        auto s = data.AccessAs<const Vector3>();
        if(hi >= mg.params.size() || s.size() != hi - lo + 1) throw MRayError("{}: cauchy coefficient range mismatch", mg.type);
        for(size_t k = 0; k < s.size(); k++)
            for(unsigned c = 0; c < 3; c++) mg.params[lo + k][(attributeIndex == 1 ? 0 : 4) + c] = s[k][c];
    }
    void PushMatAttribute(MatGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data,
                          std::vector<Optional<TextureId>> tex) override
    {
        MatGroupB200& mg = Get(mats, Raw(g), "MaterialGroup");
        if(mg.type == "(Mt)Reflect") throw MRayError("{} group does not have any attributes!", mg.type);
        if(mg.type == "(Mt)Refract") throw MRayError("{}: Attribute {:d} is \"ConstantOnly\", wrong function is called", mg.type, attributeIndex);
        if(mg.type == "(Mt)Unreal" && attributeIndex >= 2 && !data.IsEmpty())
        {   // roughness / specular / metallic: ParamVaryingData<2, Float> (MatGroupUnreal::PushTexAttribute, MaterialsDefault.cpp:L433-454)
            if(attributeIndex > 4) throw MRayError("{:s}: Attribute {:d} is not \"ParamVarying\", wrong function is called", mg.type, attributeIndex);
            for(const auto& t : tex) if(t.has_value()) throw MRayError("{}: textured roughness / specular / metallic are not supported yet", mg.type);
            uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
            auto s = data.AccessAs<const Float>();
            if(hi >= mg.params.size() || s.size() != hi - lo + 1 || tex.size() != s.size()) throw MRayError("{}: attribute range mismatch", mg.type);
            for(size_t k = 0; k < s.size(); k++) mg.params[lo + k][attributeIndex - 2] = s[k];
            return;
        }
        // An EMPTY TransientData routes to the optional texture-only overload (TracerBase.cpp:L843-867): the
        // scene loader always pushes Lambert's optional `normalMap` this way, with nullopt where a material has none.
        if(data.IsEmpty())
        {
            if(attributeIndex != 1) throw MRayError("{}: Attribute {:d} is not \"Optional Texture\"", mg.type, attributeIndex);
            uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
            if(hi >= mg.normalTex.size() || tex.size() != hi - lo + 1) throw MRayError("{}: normalMap range mismatch", mg.type);
            for(size_t k = 0; k < tex.size(); k++)
            {
                mg.normalTex[lo + k] = -1;
                if(!tex[k].has_value()) continue;
                const uint32_t tid = Raw(*tex[k]);
                if(tid == 0 || tid > textures.size()) throw MRayError("{}: Given texture({}) is not found", mg.type, tid);
                // the normal map wants a TracerTexView<2, Vector3> (MaterialsDefault.h:L14)
                if(textures[tid - 1].channels != 4 || textures[tid - 1].params.readMode != MRayTextureReadMode::MR_DROP_1)
                    throw MRayError("{}: Given texture({}) does not have a correct type for, Attribute {}", mg.type, tid, attributeIndex);
                mg.normalTex[lo + k] = int32_t(tid);
            }
            return;
        }
        if(attributeIndex != 0) throw MRayError("{}: Attribute {:d} is not \"ParamVarying\"", mg.type, attributeIndex);
        // GenericTexturedGroupT::GenericPushTexAttribute (Tracer/GenericGroup.cpp:L243-288): one Optional<TextureId>
        // AND one constant per material of the range; the constant is used where there is no texture
        uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
        auto s = data.AccessAs<const Vector3>();
        if(hi >= mg.albedo.size() || tex.size() != hi - lo + 1 || s.size() != tex.size()) throw MRayError("{}: albedo range mismatch", mg.type);
        for(size_t k = 0; k < tex.size(); k++)
        {
            mg.albedo[lo + k] = s[k];
            mg.albedoTex[lo + k] = -1;
            if(tex[k].has_value())
            {
                const uint32_t tid = Raw(*tex[k]);
                if(tid == 0 || tid > textures.size()) throw MRayError("{}: Given texture({}) is not found", mg.type, tid);
                // GenericTexturedGroupT::ConvertToView: the albedo wants a TracerTexView<2, Vector3>
                if(textures[tid - 1].channels != 4 || textures[tid - 1].params.readMode != MRayTextureReadMode::MR_DROP_1)
                    throw MRayError("{}: Given texture({}) does not have a correct type for, Attribute {}", mg.type, tid, attributeIndex);
                mg.albedoTex[lo + k] = int32_t(tid);
            }
        }
    }
    void PushMatAttribute(MatGroupId, CommonIdRange, uint32_t, std::vector<TextureId>) override
    { throw MRayError("(Mt)Lambert: texture-only attributes (normalMap) are not supported yet"); }

    // ------------------------------- textures -------------------------------
    // TextureMemory::CreateTexture2D / PushTextureData / CommitTextures (Tracer/TextureMemory.cpp:L655-800), first slice
    TextureId CreateTexture2D(Vector2ui size, uint32_t mipCount, MRayTextureParameters p) override
    {
        std::lock_guard lk(mtx);
        if(size[0] == 0 || size[1] == 0) throw MRayError("textures: empty texture");
        uint32_t fullMips = 0; for(uint32_t m = std::max(size[0], size[1]); m; m >>= 1) fullMips++;   // Graphics::TextureMipCount
        if(mipCount == 0 || mipCount > fullMips) throw MRayError("textures: mipCount {} of a {}x{} texture", mipCount, size[0], size[1]);
        TextureB200 t; t.size = size; t.params = p; t.mipCount = mipCount; t.levelLoaded.assign(mipCount, 0);
        switch(p.pixelType.Name())
        {
            case MRayPixelEnum::MR_RGBA_FLOAT:  t.channels = 4; t.format = 0; break;
            case MRayPixelEnum::MR_RGBA8_UNORM: t.channels = 4; t.format = 1; break;
            case MRayPixelEnum::MR_R_FLOAT:     t.channels = 1; t.format = 0; break;   // alpha maps (TracerTexView<2, Float>)
            case MRayPixelEnum::MR_R8_UNORM:    t.channels = 1; t.format = 1; break;
            default: throw MRayError("textures: only MR_RGBA_FLOAT / MR_RGBA8_UNORM / MR_R_FLOAT / MR_R8_UNORM pixels are supported yet");
        }
        // TextureMemory::ConvertColorspaces leaves a texture alone when it is not a colour, or already global + linear
        const bool needsConversion = p.isColor == AttributeIsColor::IS_COLOR &&
            ((p.colorSpace != MRayColorSpaceEnum::MR_DEFAULT && p.colorSpace != params.globalTextureColorSpace) || p.gamma != Float(1));
        if(needsConversion)
        {   // converted on upload by libmray_b200 (KCConvertColor): gamma to linear, then the RGB -> RGB matrix into the global space
            if(t.channels < 3) throw MRayError("textures: colour conversion of a {}-channel texture is not supported yet", t.channels);
            t.gamma = float(p.gamma);
            if(p.colorSpace != MRayColorSpaceEnum::MR_DEFAULT && p.colorSpace != params.globalTextureColorSpace)
            {
                // ColorspaceTransfer<from, global>::RGBToRGBMatrix = FromXYZ(global) * ToXYZ(from) (Core/ColorFunctions.h:L147-153)
                const Matrix3x3 m = LuminanceMatrix(params.globalTextureColorSpace).Inverse() * LuminanceMatrix(p.colorSpace);
                for(unsigned r = 0; r < 3; r++) for(unsigned c = 0; c < 3; c++) t.colorMatrix[3 * r + c] = m(r, c);
                t.hasColorMatrix = true;
            }
        }
        // the view type follows DetermineReadMode (Tracer/TextureMemory.cpp:L329-396): 4 channels + MR_DROP_1 read as Vector3
        if(p.readMode != MRayTextureReadMode::MR_PASSTHROUGH && p.readMode != MRayTextureReadMode::MR_DROP_1)
            throw MRayError("textures: only MR_PASSTHROUGH / MR_DROP_1 reads are supported yet");
        // TracerParameters.genMips: every (non block-compressed) texture gets its full chain (TextureMemory::CreateTexture L588-591),
        // filtered on upload by libmray_b200 with TracerParameters.mipGenFilter
        t.genMips = params.genMips; t.mipFilterType = uint32_t(params.mipGenFilter.type); t.mipFilterRadius = float(params.mipGenFilter.radius);
        t.pixels.resize(t.LevelStart(mipCount) * t.TexelBytes());
        // TracerParameters.clampedTexRes: applied on upload by libmray_b200 (levels dropped, or the pushed image filtered down)
        t.clampRes = p.ignoreResClamp ? 0u : params.clampedTexRes;
        textures.push_back(std::move(t));
        return TextureId(uint32_t(textures.size()));
    }
    TextureId CreateTexture3D(Vector3ui, uint32_t, MRayTextureParameters) override { throw MRayError("3-D textures are not supported yet"); }
    void CommitTextures() override {}
    void PushTextureData(TextureId id, uint32_t mipLevel, TransientData data) override
    {
        const uint32_t tid = Raw(id);
        if(tid == 0 || tid > textures.size()) throw MRayError("Unable to find texture({})", tid);
        TextureB200& t = textures[tid - 1];
        if(mipLevel >= t.mipCount) throw MRayError("textures: mip level {} of a texture with {} levels", mipLevel, t.mipCount);
        // the TransientData is typed by the pixel (MRayPixelType<E>::Type): Vector4 / Vector4uc here
        const size_t pixels = size_t(TextureB200::LevelDim(t.size[0], mipLevel)) * TextureB200::LevelDim(t.size[1], mipLevel);
        const Byte* src = nullptr; size_t got = 0;
        if(t.channels == 1 && t.format == 0) { auto s = data.AccessAs<const Float>(); src = reinterpret_cast<const Byte*>(s.data()); got = s.size(); }
        else if(t.channels == 1) { auto s = data.AccessAs<const uint8_t>(); src = reinterpret_cast<const Byte*>(s.data()); got = s.size(); }
        else if(t.format == 0) { auto s = data.AccessAs<const Vector4>(); src = reinterpret_cast<const Byte*>(s.data()); got = s.size(); }
        else { auto s = data.AccessAs<const Vector4uc>(); src = reinterpret_cast<const Byte*>(s.data()); got = s.size(); }
        if(got != pixels) throw MRayError("textures: {} pixels pushed, {} expected", got, pixels);
        std::copy(src, src + pixels * t.TexelBytes(), t.pixels.begin() + ptrdiff_t(t.LevelStart(mipLevel) * t.TexelBytes()));
        t.levelLoaded[mipLevel] = 1;
        if(mipLevel != 0) return;
        t.loaded = true;
    }

    // ------------------------------- transforms -------------------------------
    TransGroupId CreateTransformGroup(std::string typeName) override
    {
        std::lock_guard lk(mtx);
        if(typeName != "(T)Single") throw MRayError("Unable to find generator for {}", typeName);
        transforms.push_back(TransGroupB200{typeName});
        return TransGroupId(uint32_t(transforms.size() - 1));
    }
    TransformId ReserveTransformation(TransGroupId g, AttributeCountList c) override { return ReserveTransformations(g, {c}).front(); }
    TransformIdList ReserveTransformations(TransGroupId g, std::vector<AttributeCountList> counts) override
    {
        std::lock_guard lk(mtx);
        TransGroupB200& tg = Get(transforms, Raw(g), "TransformGroup");
        if(Raw(g) == 0 || tg.committed) throw MRayError("{}: reservations are already committed", tg.type);
        TransformIdList out;
        for(size_t i = 0; i < counts.size(); i++)
        {
            tg.matrices.push_back(Matrix3x4::Identity());
            out.push_back(TransformId((Raw(g) << TRANS_ID_BITS) | uint32_t(tg.matrices.size() - 1)));
        }
        return out;
    }
    void CommitTransReservations(TransGroupId g) override { Get(transforms, Raw(g), "TransformGroup").committed = true; }
    bool IsTransCommitted(TransGroupId g) const override { return Get(transforms, Raw(g), "TransformGroup").committed; }
    void PushTransAttribute(TransGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data) override
    {
        TransGroupB200& tg = Get(transforms, Raw(g), "TransformGroup");
        if(Raw(g) == 0) throw MRayError("(T)Identity has no attributes");
        if(attributeIndex != 0) throw MRayError("{:s}: Unknown AttributeIndex {:d}", tg.type, attributeIndex);
        uint32_t lo = range[0] & ((1u << TRANS_ID_BITS) - 1u), hi = range[1] & ((1u << TRANS_ID_BITS) - 1u);
        Span<const Matrix3x4> m = data.AccessAs<const Matrix3x4>();
        if(hi >= tg.matrices.size() || m.size() != size_t(hi - lo + 1)) throw MRayError("{}: transform range / data size mismatch", tg.type);
        for(uint32_t i = lo; i <= hi; i++) tg.matrices[i] = m[i - lo];
    }

    // ------------------------------- lights -------------------------------
    LightGroupId CreateLightGroup(std::string typeName, PrimGroupId pg) override
    {
        std::lock_guard lk(mtx);
        if(typeName != "(L)Prim(P)Triangle" && typeName != "(L)Skysphere_Spherical" && typeName != "(L)Skysphere_CoOcta")
            throw MRayError("Unable to find generator for {}", typeName);
        lights.push_back(LightGroupB200{typeName, false, Raw(pg)});
        return LightGroupId(uint32_t(lights.size() - 1));
    }
    LightId ReserveLight(LightGroupId g, AttributeCountList c, PrimBatchId b) override { return ReserveLights(g, {c}, {b}).front(); }
    LightIdList ReserveLights(LightGroupId g, std::vector<AttributeCountList> counts, std::vector<PrimBatchId> batches) override
    {
        std::lock_guard lk(mtx);
        LightGroupB200& lg = Get(lights, Raw(g), "LightGroup");
        if(!lg.IsSkysphere() && batches.size() != counts.size()) throw MRayError("{}: prim-backed lights need one prim batch each", lg.type);
        LightIdList out;
        for(size_t i = 0; i < counts.size(); i++)
        {
            lg.radiance.push_back(Vector3::Zero()); lg.twoSided.push_back(0); lg.radianceTex.push_back(-1);
            lg.primBatch.push_back(i < batches.size() ? (Raw(batches[i]) & ((1u << PRIM_ID_BITS) - 1u)) : 0u);
            out.push_back(LightId((Raw(g) << MAT_ID_BITS) | uint32_t(lg.radiance.size() - 1)));
        }
        return out;
    }
    void CommitLightReservations(LightGroupId g) override { Get(lights, Raw(g), "LightGroup").committed = true; }
    bool IsLightCommitted(LightGroupId g) const override { return Get(lights, Raw(g), "LightGroup").committed; }
    void PushLightAttribute(LightGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data) override
    {
        LightGroupB200& lg = Get(lights, Raw(g), "LightGroup");
        if(attributeIndex != 1) throw MRayError("{}: Attribute {:d} is not \"ConstantOnly\", wrong function is called", lg.type, attributeIndex);
        uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
        auto s = data.AccessAs<const bool>();
        if(hi >= lg.twoSided.size() || s.size() != hi - lo + 1) throw MRayError("{}: isTwoSided range mismatch", lg.type);
        for(size_t i = 0; i < s.size(); i++) lg.twoSided[lo + i] = s[i] ? 1 : 0;
    }
    void PushLightAttribute(LightGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data,
                            std::vector<Optional<TextureId>> tex) override
    {
        LightGroupB200& lg = Get(lights, Raw(g), "LightGroup");
        if(attributeIndex != 0) throw MRayError("{}: Attribute {:d} is not \"ParamVarying\", wrong function is called", lg.type, attributeIndex);
        uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
        auto s = data.AccessAs<const Vector3>();
        if(hi >= lg.radiance.size() || s.size() != hi - lo + 1) throw MRayError("{}: radiance range mismatch", lg.type);
        std::copy(s.begin(), s.end(), lg.radiance.begin() + lo);
        for(size_t i = 0; i < tex.size() && lo + i <= hi; i++)
        {
            if(!tex[i].has_value()) continue;
            // LightGroupSkysphere::PushTexAttribute (Tracer/LightsDefault.hpp:L779-816): the radiance field of an environment map
            if(!lg.IsSkysphere()) throw MRayError("{}: textured radiance is not supported yet", lg.type);
            const uint32_t tid = Raw(*tex[i]);
            if(tid == 0 || tid > textures.size()) throw MRayError("{:s}: Given texture({:d}) is not found", lg.type, tid);
            lg.radianceTex[lo + i] = int32_t(tid);
        }
    }
    void PushLightAttribute(LightGroupId, CommonIdRange, uint32_t, std::vector<TextureId>) override
    { throw MRayError("textured lights are not supported yet"); }

    // ------------------------------- cameras -------------------------------
    CameraGroupId CreateCameraGroup(std::string typeName) override
    {
        std::lock_guard lk(mtx);
        if(typeName != "(C)Pinhole") throw MRayError("Unable to find generator for {}", typeName);
        cams.push_back(CamGroupB200{typeName});
        return CameraGroupId(uint32_t(cams.size() - 1));
    }
    CameraId ReserveCamera(CameraGroupId g, AttributeCountList c) override { return ReserveCameras(g, {c}).front(); }
    CameraIdList ReserveCameras(CameraGroupId g, std::vector<AttributeCountList> counts) override
    {
        std::lock_guard lk(mtx);
        CamGroupB200& cg = Get(cams, Raw(g), "CameraGroup");
        CameraIdList out;
        for(size_t i = 0; i < counts.size(); i++)
        {
            cg.fovPlanes.push_back(Vector4::Zero()); cg.gaze.push_back(Vector3::Zero());
            cg.position.push_back(Vector3::Zero()); cg.up.push_back(Vector3::YAxis());
            out.push_back(CameraId((Raw(g) << CAM_ID_BITS) | uint32_t(cg.gaze.size() - 1)));
        }
        return out;
    }
    void CommitCamReservations(CameraGroupId g) override { Get(cams, Raw(g), "CameraGroup").committed = true; }
    bool IsCamCommitted(CameraGroupId g) const override { return Get(cams, Raw(g), "CameraGroup").committed; }
    void PushCamAttribute(CameraGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data) override
    {
        CamGroupB200& cg = Get(cams, Raw(g), "CameraGroup");
        uint32_t lo = range[0] & ((1u << CAM_ID_BITS) - 1u);
        if(lo >= cg.gaze.size()) throw MRayError("{}: unknown camera {}", cg.type, lo);
        if(attributeIndex == 0) cg.fovPlanes[lo] = data.AccessAs<const Vector4>().front();
        else if(attributeIndex == 1) cg.gaze[lo] = data.AccessAs<const Vector3>().front();
        else if(attributeIndex == 2) cg.position[lo] = data.AccessAs<const Vector3>().front();
        else if(attributeIndex == 3) cg.up[lo] = data.AccessAs<const Vector3>().front();
        else throw MRayError("{}: unknown attribute index {}", cg.type, attributeIndex);
    }

    // ------------------------------- mediums -------------------------------
    MediumGroupId CreateMediumGroup(std::string typeName) override { throw MRayError("Unable to find generator for {}", typeName); }
    MediumId ReserveMedium(MediumGroupId, AttributeCountList) override { throw MRayError("only (Md)Vacuum exists"); }
    MediumIdList ReserveMediums(MediumGroupId, std::vector<AttributeCountList>) override { throw MRayError("only (Md)Vacuum exists"); }
    void CommitMediumReservations(MediumGroupId) override {}
    bool IsMediumCommitted(MediumGroupId) const override { return true; }
    void PushMediumAttribute(MediumGroupId, CommonIdRange, uint32_t, TransientData) override { throw MRayError("only (Md)Vacuum exists"); }
    void PushMediumAttribute(MediumGroupId, CommonIdRange, uint32_t, TransientData, std::vector<Optional<TextureId>>) override { throw MRayError("only (Md)Vacuum exists"); }
    void PushMediumAttribute(MediumGroupId, CommonIdRange, uint32_t, std::vector<TextureId>) override { throw MRayError("only (Md)Vacuum exists"); }

    // ------------------------------- surfaces -------------------------------
    SurfaceId CreateSurface(SurfaceParams p) override { std::lock_guard lk(mtx); surfaces.push_back(p); return SurfaceId(uint32_t(surfaces.size() - 1)); }
    LightSurfaceId SetBoundarySurface(LightId l, TransformId t) override { boundary = LightSurfaceParams{l, t, {}}; return LightSurfaceId(0xFFFFFFFEu); }
    LightSurfaceId CreateLightSurface(LightSurfaceParams p) override { std::lock_guard lk(mtx); lightSurfaces.push_back(p); return LightSurfaceId(uint32_t(lightSurfaces.size() - 1)); }
    CamSurfaceId CreateCameraSurface(CameraSurfaceParams p) override { std::lock_guard lk(mtx); camSurfaces.push_back(p); return CamSurfaceId(uint32_t(camSurfaces.size() - 1)); }
    VolumeId RegisterVolume(VolumeParams v) override { std::lock_guard lk(mtx); volumes.push_back(v); return VolumeId(uint32_t(volumes.size() - 1)); }
    VolumeIdList RegisterVolumes(std::vector<VolumeParams> v) override { VolumeIdList o; for(auto& x : v) o.push_back(RegisterVolume(x)); return o; }
    void SetBoundaryVolume(VolumeId) override {}

    SurfaceCommitResult CommitSurfaces() override
    {
        // TracerBase::CommitSurfaces (Tracer/TracerBase.cpp:L1529-1674) + BaseAccelerator::Construct:
        // surfaces sharing a transform become the prim ranges of one accelerator; (T)Identity-only scenes
        // are a single accelerator, anything else a two-level scene with one instance per transform.
        if(Raw(boundary.lightId) != 0)
        {   // RendererCommon.cu:L337-345: the boundary light cannot be primitive backed; here it is (L)Null or a skysphere
            const LightGroupB200& blg = Get(lights, Raw(boundary.lightId) >> MAT_ID_BITS, "LightGroup");
            if(!blg.IsSkysphere()) throw MRayError("Primitive-backed light ({}) is requested as a boundary material!", blg.type);
            if((Raw(boundary.lightId) & ((1u << MAT_ID_BITS) - 1u)) >= blg.radiance.size()) throw MRayError("Unable to find Light({})", Raw(boundary.lightId));
        }
        struct Group { uint32_t transformId; std::vector<uint32_t> ranges, lmKeys; std::vector<uint8_t> cull; std::vector<int32_t> alpha; };   // alpha: TextureId or -1
        std::vector<Group> groups;
        auto GroupOf = [&](TransformId t) -> Group&
        {
            for(Group& g : groups) if(g.transformId == Raw(t)) return g;
            const TransGroupB200& tg = Get(transforms, Raw(t) >> TRANS_ID_BITS, "TransformGroup");
            if((Raw(t) & ((1u << TRANS_ID_BITS) - 1u)) >= tg.matrices.size()) throw MRayError("Unable to find Transform({})", Raw(t));
            groups.push_back(Group{Raw(t), {}, {}, {}, {}});
            return groups.back();
        };
        flatAlbedo.clear(); flatAlbedoTex.clear(); flatNormalTex.clear(); flatMaterialType.clear(); flatMaterialParams.clear(); flatTextures.clear();
        flatLightRadiance.clear(); flatLightTwoSided.clear();
        int32_t pgUsed = -1;
        auto UsePrimGroup = [&](uint32_t g)
        {
            if(pgUsed >= 0 && uint32_t(pgUsed) != g) throw MRayError("more than one triangle primitive group per scene is not supported yet");
            pgUsed = int32_t(g);
        };
        // material table: flat index = running index over (group, id) pairs in first-use order
        std::vector<uint32_t> matKeyOf;
        auto FlatMat = [&](MaterialId m)
        {
            for(size_t i = 0; i < matKeyOf.size(); i++) if(matKeyOf[i] == Raw(m)) return uint32_t(i);
            const MatGroupB200& mg = Get(mats, Raw(m) >> MAT_ID_BITS, "MaterialGroup");
            uint32_t idx = Raw(m) & ((1u << MAT_ID_BITS) - 1u);
            if(idx >= mg.albedo.size()) throw MRayError("Unable to find Material({})", Raw(m));
            matKeyOf.push_back(Raw(m));
            flatAlbedo.insert(flatAlbedo.end(), {mg.albedo[idx][0], mg.albedo[idx][1], mg.albedo[idx][2]});
            int32_t ft = -1;
            if(mg.albedoTex[idx] >= 0)
            {
                const uint32_t tid = uint32_t(mg.albedoTex[idx]);
                if(!textures[tid - 1].loaded) throw MRayError("texture({}) has no data", tid);
                auto it = std::find(flatTextures.begin(), flatTextures.end(), tid);
                ft = int32_t(it - flatTextures.begin());
                if(it == flatTextures.end()) flatTextures.push_back(tid);
            }
            flatAlbedoTex.push_back(ft);
            int32_t fn = -1;
            if(mg.normalTex[idx] >= 0)
            {
                const uint32_t tid = uint32_t(mg.normalTex[idx]);
                if(!textures[tid - 1].loaded) throw MRayError("texture({}) has no data", tid);
                auto it = std::find(flatTextures.begin(), flatTextures.end(), tid);
                fn = int32_t(it - flatTextures.begin());
                if(it == flatTextures.end()) flatTextures.push_back(tid);
            }
            flatNormalTex.push_back(fn);
            flatMaterialType.push_back(mg.type == "(Mt)Reflect" ? uint8_t(MRB_MATERIAL_REFLECT) : mg.type == "(Mt)Refract" ? uint8_t(MRB_MATERIAL_REFRACT)
                                       : mg.type == "(Mt)Unreal" ? uint8_t(MRB_MATERIAL_UNREAL) : uint8_t(MRB_MATERIAL_LAMBERT));
            flatMaterialParams.insert(flatMaterialParams.end(), mg.params[idx].begin(), mg.params[idx].end());
            return uint32_t(matKeyOf.size() - 1);
        };
        for(const SurfaceParams& s : surfaces)
        {
            Group& grp = GroupOf(s.transformId);
            for(size_t k = 0; k < s.primBatches.size(); k++)
            {
                uint32_t g = Raw(s.primBatches[k]) >> PRIM_ID_BITS, bi = Raw(s.primBatches[k]) & ((1u << PRIM_ID_BITS) - 1u);
                UsePrimGroup(g);
                const PrimBatch& pb = Get(Get(prims, g, "PrimitiveGroup").batches, bi, "PrimitiveBatch");
                grp.ranges.insert(grp.ranges.end(), {pb.primOffset, pb.primOffset + pb.primCount});
                grp.lmKeys.push_back(FlatMat(s.materials[k]));
                grp.cull.push_back(s.cullFaceFlags[k] ? 1 : 0);
                int32_t am = -1;
                if(k < s.alphaMaps.size() && s.alphaMaps[k].has_value())
                {   // AcceleratorCommon.cu:L293-311: the alpha map must be a single-channel texture (a TracerTexView<2, Float>)
                    const uint32_t tid = Raw(*s.alphaMaps[k]);
                    if(tid == 0 || tid > textures.size()) throw MRayError("Alpha map texture({}) is not found", tid);
                    if(textures[tid - 1].channels != 1) throw MRayError("Alpha map texture({}) is not a single channel texture!", tid);
                    if(!textures[tid - 1].loaded) throw MRayError("texture({}) has no data", tid);
                    am = int32_t(tid);
                }
                grp.alpha.push_back(am);
            }
        }
        for(const LightSurfaceParams& ls : lightSurfaces)
        {
            Group& grp = GroupOf(ls.transformId);
            const LightGroupB200& lg = Get(lights, Raw(ls.lightId) >> MAT_ID_BITS, "LightGroup");
            uint32_t li = Raw(ls.lightId) & ((1u << MAT_ID_BITS) - 1u);
            if(li >= lg.radiance.size()) throw MRayError("Unable to find Light({})", Raw(ls.lightId));
            UsePrimGroup(lg.primGroup);
            const PrimBatch& pb = Get(Get(prims, lg.primGroup, "PrimitiveGroup").batches, lg.primBatch[li], "PrimitiveBatch");
            grp.ranges.insert(grp.ranges.end(), {pb.primOffset, pb.primOffset + pb.primCount});
            grp.lmKeys.push_back(0x80000000u | uint32_t(flatLightTwoSided.size()));
            grp.cull.push_back(0); grp.alpha.push_back(-1);
            flatLightRadiance.insert(flatLightRadiance.end(), {lg.radiance[li][0], lg.radiance[li][1], lg.radiance[li][2]});
            flatLightTwoSided.push_back(lg.twoSided[li]);
        }
        if(pgUsed < 0 || groups.empty()) throw MRayError("empty scene");
        flatPrimGroup = uint32_t(pgUsed);
        const PrimGroupB200& pg = prims[flatPrimGroup];
        ReleaseRenderers();
        ReleaseAccels();
        auto BuildGroup = [&](const DeviceB200& dv, const Group& g)
        {
            mrb_accel_desc d = {};
            d.positions = reinterpret_cast<const float*>(pg.positions.data()); d.vertexCount = pg.vertexTotal;
            d.indices = reinterpret_cast<const uint32_t*>(pg.indices.data()); d.triangleCount = pg.primTotal;
            d.memspace = MRB_MEM_HOST; d.primGroupId = flatPrimGroup;
            d.rangeCount = uint32_t(g.lmKeys.size()); d.primRanges = g.ranges.data();
            d.lightOrMatKeys = g.lmKeys.data(); d.cullBackface = g.cull.data(); d.flags = MRB_BUILD_DEFAULT;
            // alpha maps of this accelerator: its own compact texture table
            std::vector<uint32_t> used; std::vector<int32_t> rangeAlpha(g.alpha.size(), -1); std::vector<mrb_texture_desc> atex;
            for(size_t k = 0; k < g.alpha.size(); k++)
            {
                if(g.alpha[k] < 0) continue;
                auto it = std::find(used.begin(), used.end(), uint32_t(g.alpha[k]));
                rangeAlpha[k] = int32_t(it - used.begin());
                if(it != used.end()) continue;
                used.push_back(uint32_t(g.alpha[k]));
                const TextureB200& t = textures[size_t(g.alpha[k]) - 1];
                atex.push_back(t.Desc(false));   // pure data: no colour conversion; read at level 0 (IntersectionCheck has no gradients)
                atex.back().mipCount = 1; atex.back().generateMips = 0; atex.back().clampResolution = 0;   // alpha maps stay at full resolution
            }
            if(!used.empty())
            {
                d.vertexUVs = reinterpret_cast<const float*>(pg.uvs.data());
                d.alphaTextureCount = uint32_t(atex.size()); d.alphaTextures = atex.data(); d.rangeAlphaMap = rangeAlpha.data();
            }
            mrb_accel a = nullptr;
            CheckOn(dv, mrb_accel_build(dv.ctx, &d, &a));
            return a;
        };
        AABB3 aabb;
        if(groups.size() == 1 && groups[0].transformId == 0)
        {
            // the scene is replicated: every device builds its own copy of the BVH (264 K triangles: 0.5 ms)
            for(DeviceB200& dv : devs) dv.accel = BuildGroup(dv, groups[0]);
            mrb_accel_info info; Check(mrb_accel_get_info(ctx, devs[0].accel, &info));
            aabb = AABB3(Vector3(info.aabb[0], info.aabb[1], info.aabb[2]), Vector3(info.aabb[3], info.aabb[4], info.aabb[5]));
        }
        else
        {
            // Groups with the same prim ranges and cull flags share ONE accelerator and differ only in transform and
            // LightOrMatKeys — the reference's concrete-accelerator / instance split (Tracer/AcceleratorC.h:L780-905).
            for(DeviceB200& dv : devs)
            {
                std::vector<mrb_instance_desc> inst(groups.size());
                std::vector<size_t> builtFor;   // group index each unique accelerator was built from
                for(size_t k = 0; k < groups.size(); k++)
                {
                    mrb_accel a = nullptr;
                    for(size_t u = 0; u < builtFor.size() && !a; u++)
                        if(groups[builtFor[u]].ranges == groups[k].ranges && groups[builtFor[u]].cull == groups[k].cull &&
                           groups[builtFor[u]].alpha == groups[k].alpha) a = dv.instAccels[u];
                    if(!a) { a = BuildGroup(dv, groups[k]); dv.instAccels.push_back(a); builtFor.push_back(k); }
                    const uint32_t tid = groups[k].transformId;
                    const Matrix3x4& m = transforms[tid >> TRANS_ID_BITS].matrices[tid & ((1u << TRANS_ID_BITS) - 1u)];
                    const std::array<float, 12> inv = InverseAffine(m);  // KCInvertTransforms (Tracer/TransformC.h:L119-123)
                    mrb_instance_desc& d = inst[k];
                    d = {};
                    d.accel = a;
                    for(unsigned r = 0; r < 3; r++) for(unsigned c = 0; c < 4; c++)
                    { d.transform[4 * r + c] = m(r, c); d.invTransform[4 * r + c] = inv[4 * r + c]; }
                    d.isIdentity = (tid == 0) ? 1 : 0;
                    d.transformKey = tid; d.accelKey = uint32_t(k);
                    d.lightOrMatKeys = groups[k].lmKeys.data();
                }
                CheckOn(dv, mrb_scene_build(dv.ctx, inst.data(), uint32_t(inst.size()), &dv.scene));
                sceneInstanceCount = uint32_t(inst.size()); uniqueAccelCount = uint32_t(dv.instAccels.size());
            }
            twoLevel = true;
            float box[6];
            Check(mrb_scene_export_tlas(ctx, devs[0].scene, nullptr, box, nullptr, nullptr, nullptr, nullptr));
            aabb = AABB3(Vector3(box[0], box[1], box[2]), Vector3(box[3], box[4], box[5]));
        }
        committed = true;
        {   // TracerBase::CommitSurfaces (Tracer/TracerBase.cpp:L1653-1664): the lights learn the scene's diameter over the XZ plane
            const Vector3 span = aabb.GeomSpan();
            sceneDiameter = Math::Length(Vector2(span[0], span[2]));
        }
        return SurfaceCommitResult
        {
            .aabb = aabb,
            .instanceCount = surfaces.size() + lightSurfaces.size(),
            .acceleratorCount = twoLevel ? uniqueAccelCount : 1u
        };
    }

    CameraTransform GetCamTransform(CamSurfaceId id) const override
    {
        const CameraSurfaceParams& cs = Get(camSurfaces, Raw(id), "CameraSurface");
        const CamGroupB200& cg = Get(cams, Raw(cs.cameraId) >> CAM_ID_BITS, "CameraGroup");
        uint32_t ci = Raw(cs.cameraId) & ((1u << CAM_ID_BITS) - 1u);
        return CameraTransform{cg.position[ci], cg.gaze[ci], cg.up[ci]};
    }

    // ------------------------------- renderers -------------------------------
    RendererId CreateRenderer(std::string typeName) override
    {
        std::lock_guard lk(mtx);
        if(typeName != "(R)PathTracerRGB" && typeName != "(R)PathTracerSpectral") throw MRayError("Unable to find generator for {}", typeName);
        renderers.push_back(RendererB200{typeName});
        return RendererId(uint32_t(renderers.size() - 1));
    }
    void DestroyRenderer(RendererId id) override
    {
        if(Raw(id) >= renderers.size()) throw MRayError("Unable to find renderer ({})", Raw(id));
        if(Raw(id) == curRenderer) ReleaseRenderers();
    }
    void PushRendererAttribute(RendererId id, uint32_t attributeIndex, TransientData dataIn) override
    {
        const TransientData& data = dataIn;
        RendererB200& r = Get(renderers, Raw(id), "Renderer");
        switch(attributeIndex) // PathTracerRendererT::PushAttribute (TracerDLL/PathTracerRenderer.cu)
        {
            case 0: r.totalSPP = data.AccessAs<const uint32_t>().front(); break;
            case 1: r.burstSize = data.AccessAs<const uint32_t>().front(); break;
            case 2: { std::string_view m = data.AccessAsString();
                      if(m != "Throughput"sv && m != "Latency"sv) throw MRayError("Bad enum name"); r.latency = (m == "Latency"sv); break; }
            case 3: { std::string_view m = data.AccessAsString();
                      if(m == "Pure"sv) r.sampleMode = 0; else if(m == "WithNextEventEstimation"sv) r.sampleMode = 1;
                      else if(m == "WithNEEAndMIS"sv) r.sampleMode = 2; else throw MRayError("Bad enum name"); break; }
            case 4: r.rrRange = data.AccessAs<const Vector2ui>().front(); break;
            case 5: { if(data.AccessAsString() != "Uniform"sv) throw MRayError("Bad enum name"); break; }
            default: throw MRayError("{}: unknown attribute index {}", r.type, attributeIndex);
        }
    }

    // ------------------------------- rendering -------------------------------
    void SetupRenderEnv(TimelineSemaphore* s, uint32_t, uint64_t initialAcquireValue) override { sem = s; acquireValue = initialAcquireValue; }

    // ImageTiler::FindOptimumTileSize (Tracer/RenderImage.cpp:L20-52): a tile of about `parallelizationHint` pixels with
    // the region's aspect ratio, adjusted so that an integer number of equal tiles covers each side
    static Vector2ui FindOptimumTileSize(Vector2ui fbSize, uint32_t hint)
    {
        const Float aspect = Float(fbSize[0]) / Float(fbSize[1]);
        const Float factor = std::sqrt(Float(hint) / aspect);
        const Vector2ui tileHint(uint32_t(std::round(aspect * factor)), uint32_t(std::round(factor)));
        auto Adjust = [&](unsigned i) -> uint32_t
        {
            if(fbSize[i] < tileHint[i]) return fbSize[i];
            uint32_t count = uint32_t(std::round(Float(fbSize[i]) / Float(tileHint[i])));
            uint32_t result = fbSize[i] / count, residual = fbSize[i] % count;
            return result + (residual + count - 1) / count;
        };
        return Vector2ui(Adjust(0), Adjust(1));
    }
    Vector2ui TileStart(uint32_t t) const { return Vector2ui((t % tileCount[0]) * coveringTile[0], (t / tileCount[0]) * coveringTile[1]); }
    Vector2ui TileEnd(uint32_t t) const
    {
        Vector2ui e((t % tileCount[0] + 1) * coveringTile[0], (t / tileCount[0] + 1) * coveringTile[1]);
        return Vector2ui(std::min(e[0], regionSize[0]), std::min(e[1], regionSize[1]));
    }
    // [begin, end) of `count` items for part `k` of `n`: contiguous, sizes differ by at most one
    static std::pair<uint32_t, uint32_t> Share(uint32_t count, uint32_t n, uint32_t k)
    {
        uint32_t base = count / n, extra = count % n, b = k * base + std::min(k, extra);
        return {b, b + base + (k < extra ? 1u : 0u)};
    }

    RenderBufferInfo StartRender(RendererId id, CamSurfaceId camSurf, RenderImageParams rip, Optional<uint32_t> logic0, Optional<uint32_t>) override
    {
        if(!sem) throw MRayError("Render environment is not set properly! Please provide a semaphore to the tracer.");
        if(!committed) throw MRayError("CommitSurfaces must be called before StartRender");
        if(!rebuilding) { camOverride.reset(); pendingCam.reset(); }
        lastStart = StartArgs{id, camSurf, rip, logic0};
        const RendererB200& r = Get(renderers, Raw(id), "Renderer");
        const CameraSurfaceParams& cs = Get(camSurfaces, Raw(camSurf), "CameraSurface");
        const CamGroupB200& cg = Get(cams, Raw(cs.cameraId) >> CAM_ID_BITS, "CameraGroup");
        uint32_t ci = Raw(cs.cameraId) & ((1u << CAM_ID_BITS) - 1u);
        const PrimGroupB200& pg = prims[flatPrimGroup];
        ReleaseRenderers();
        if(rip.regionMax[0] > rip.resolution[0] || rip.regionMax[1] > rip.resolution[1] ||
           rip.regionMin[0] >= rip.regionMax[0] || rip.regionMin[1] >= rip.regionMax[1])
            throw MRayError("StartRender: bad render region");
        // TracerParameters.filmFilter (Core/TracerI.h:L45-68): an unknown type is an error, not a Gaussian
        uint32_t filterType;
        switch(params.filmFilter.type)
        {
            case FilterType::BOX: filterType = MRB_FILTER_BOX; break;
            case FilterType::TENT: filterType = MRB_FILTER_TENT; break;
            case FilterType::GAUSSIAN: filterType = MRB_FILTER_GAUSSIAN; break;
            case FilterType::MITCHELL_NETRAVALI: filterType = MRB_FILTER_MITCHELL_NETRAVALI; break;
            default: throw MRayError("Unknown film filter type ({})", uint32_t(params.filmFilter.type));
        }
        // tiles (ImageTiler ctor, Tracer/RenderImage.cpp:L54-76)
        fullResolution = rip.resolution; regionMin = rip.regionMin; regionSize = rip.regionMax - rip.regionMin;
        coveringTile = FindOptimumTileSize(regionSize, std::max<uint32_t>(params.parallelizationHint, 1u));
        tileCount = Vector2ui((regionSize[0] + coveringTile[0] - 1) / coveringTile[0], (regionSize[1] + coveringTile[1] - 1) / coveringTile[1]);
        currentTile = 0; tileSPPs.assign(size_t(tileCount[0]) * tileCount[1], 0u);
        // the samples this process is responsible for
        jobBegin = 0; jobEnd = r.totalSPP;
        if(const char* e = std::getenv("MRB_SPP_SHARD"))
        {
            unsigned rank = 0, world = 1;
            if(std::sscanf(e, "%u/%u", &rank, &world) != 2 || world == 0 || rank >= world) throw MRayError("MRB_SPP_SHARD must be \"rank/world\"");
            auto [b, en] = Share(r.totalSPP, world, rank);
            jobBegin = b; jobEnd = en;
        }
        const uint32_t jobSamples = jobEnd - jobBegin;
        const bool singleTile = tileSPPs.size() == 1;
        // PathTracerRendererBase::DoRender's dispatch (Tracer/PathTracerRendererBase.cu:L515-533)
        throughputSingle = !r.latency && singleTile && r.burstSize <= 1;
        saveImage = true; completedPaths = 0; primed = false;

        mrb_render_desc d = {};
        const float* tbn = pg.hasTBN ? reinterpret_cast<const float*>(pg.tbn.data()) : nullptr;
        std::vector<const float*> instTBN(twoLevel ? sceneInstanceCount : 1u, tbn);
        d.materialCount = uint32_t(flatAlbedo.size() / 3); d.albedo = flatAlbedo.data();
        d.materialType = flatMaterialType.data(); d.materialParams = flatMaterialParams.data();
        std::vector<mrb_texture_desc> texDescs(flatTextures.size());
        std::vector<const float*> instUVs(twoLevel ? sceneInstanceCount : 1u, reinterpret_cast<const float*>(pg.uvs.data()));
        if(!flatTextures.empty())
        {
            for(size_t k = 0; k < flatTextures.size(); k++)
            {
                const TextureB200& t = textures[flatTextures[k] - 1];
                texDescs[k] = t.Desc();
            }
            d.textureCount = uint32_t(texDescs.size()); d.textures = texDescs.data(); d.albedoTexture = flatAlbedoTex.data();
            // mip level from the ray cone's gradients: as the reference's host backend computes it (the mode the goldens pin), or
            // MRB_TEXTURE_LOD=device for what tex2DGrad does on its device backends (gradients scaled by the texture size)
            if(const char* e = std::getenv("MRB_TEXTURE_LOD")) d.textureLodMode = (std::string(e) == "device") ? 1u : 0u;
            if(std::any_of(flatNormalTex.begin(), flatNormalTex.end(), [](int32_t t) { return t >= 0; })) d.normalTexture = flatNormalTex.data();
        }
        // boundary light surface: (L)Null (nothing to do) or a skysphere (LightGroupSkysphere, Tracer/LightsDefault.hpp:L704-893)
        d.boundaryTexture = -1;
        std::array<float, 12> skyMatrix{};
        if(Raw(boundary.lightId) != 0)
        {
            const LightGroupB200& blg = Get(lights, Raw(boundary.lightId) >> MAT_ID_BITS, "LightGroup");
            const uint32_t bi = Raw(boundary.lightId) & ((1u << MAT_ID_BITS) - 1u);
            d.boundaryType = blg.type == "(L)Skysphere_Spherical" ? MRB_BOUNDARY_SKYSPHERE_SPHERICAL : MRB_BOUNDARY_SKYSPHERE_COOCTA;
            for(int k = 0; k < 3; k++) d.boundaryRadiance[k] = blg.radiance[bi][k];
            if(blg.radianceTex[bi] >= 0)
            {
                const uint32_t tid = uint32_t(blg.radianceTex[bi]);
                const TextureB200& t = textures[tid - 1];
                if(!t.loaded) throw MRayError("texture({}) has no data", tid);
                auto it = std::find(flatTextures.begin(), flatTextures.end(), tid);
                if(it != flatTextures.end()) d.boundaryTexture = int32_t(it - flatTextures.begin());
                else
                {   // the radiance map is not an albedo texture of any material: append it to this render's table
                    texDescs.push_back(t.Desc());
                    d.boundaryTexture = int32_t(texDescs.size() - 1);
                    d.textureCount = uint32_t(texDescs.size()); d.textures = texDescs.data();
                    if(flatTextures.empty()) d.albedoTexture = nullptr;
                }
            }
            if(Raw(boundary.transformId) != Raw(TracerConstants::IdentityTransformId))
            {
                const TransGroupB200& tg = Get(transforms, Raw(boundary.transformId) >> TRANS_ID_BITS, "TransformGroup");
                const uint32_t ti = Raw(boundary.transformId) & ((1u << TRANS_ID_BITS) - 1u);
                if(ti >= tg.matrices.size()) throw MRayError("Unable to find Transform({})", Raw(boundary.transformId));
                for(unsigned rr = 0; rr < 3; rr++) for(unsigned cc = 0; cc < 4; cc++) skyMatrix[4 * rr + cc] = tg.matrices[ti](rr, cc);
                d.boundaryTransform = skyMatrix.data();
            }
            // KCExtractLuminance weighs the texel with the Y row of the global texture colour space's RGB -> XYZ matrix
            const Matrix3x3 toXYZ = LuminanceMatrix(params.globalTextureColorSpace);
            for(int k = 0; k < 3; k++) d.luminanceRow[k] = toXYZ(1, k);
            d.sceneDiameter = sceneDiameter;
        }
        d.lightCount = uint32_t(flatLightTwoSided.size()); d.lightRadiance = flatLightRadiance.data(); d.lightTwoSided = flatLightTwoSided.data();
        for(int k = 0; k < 3; k++) { d.camPosition[k] = cg.position[ci][k]; d.camGaze[k] = cg.gaze[ci][k]; d.camUp[k] = cg.up[ci][k]; }
        if(camOverride)   // SetCameraTransform: position / gaze point / up replace the camera's own, fov and planes stay
            for(int k = 0; k < 3; k++) { d.camPosition[k] = camOverride->position[k]; d.camGaze[k] = camOverride->gazePoint[k]; d.camUp[k] = camOverride->up[k]; }
        d.fovXY[0] = cg.fovPlanes[ci][0]; d.fovXY[1] = cg.fovPlanes[ci][1];
        d.nearFar[0] = cg.fovPlanes[ci][2]; d.nearFar[1] = cg.fovPlanes[ci][3];
        // the renderer's film is one (covering) tile; passes move it over the region
        d.width = coveringTile[0]; d.height = coveringTile[1];
        d.fullResolution[0] = rip.resolution[0]; d.fullResolution[1] = rip.resolution[1];
        d.regionMin[0] = rip.regionMin[0]; d.regionMin[1] = rip.regionMin[1];
        // render logic 0 rolls the sample mode like PathTracerRendererT::StartRender (L1192-1200)
        d.sampleMode = (r.sampleMode + logic0.value_or(0)) % 3u;
        d.rrRange[0] = r.rrRange[0]; d.rrRange[1] = r.rrRange[1];
        d.filmFilterRadius = params.filmFilter.radius; d.filmFilterType = filterType; d.seed = params.seed;
        uint64_t pixels = uint64_t(coveringTile[0]) * coveringTile[1];
        d.maxPathCount = uint32_t(std::min<uint64_t>(pixels, std::max<uint32_t>(params.parallelizationHint, 1u)));
        // every material of this plugin is shaded by one fused kernel, so the material-key ray sort
        // (RayPartitioner::MultiPartition) only costs: 8.74 -> 7.74 ms/spp at 1080p without it. It stays
        // available (MRB_PARTITION_RAYS=1) for parity with the reference's per-material work batches.
        const char* pr = std::getenv("MRB_PARTITION_RAYS");
        d.partitionRays = (pr && pr[0] == '1') ? 1u : 0u;
        if(r.type == "(R)PathTracerSpectral") EnsureSpectrum();
        if(params.samplerType.e != SamplerType::INDEPENDENT)
        {
            // the scramble's final bit reversal is applied unless MRB_REFERENCE_SCRAMBLE=1 (see include/mray_b200.h)
            EnsureSobolMatrices();
            d.samplerType = (params.samplerType.e == SamplerType::SOBOL) ? MRB_SAMPLER_SOBOL : MRB_SAMPLER_ZSOBOL;
            const char* rs = std::getenv("MRB_REFERENCE_SCRAMBLE");
            if(rs && rs[0] == '1') d.samplerType |= MRB_SAMPLER_REFERENCE_SCRAMBLE;
            d.sobolMatrices = sobolMatrices.data();
        }
        d.jobSPP = r.totalSPP;
        const uint32_t nDev = uint32_t(devs.size());
        for(uint32_t k = 0; k < nDev; k++)
        {
            DeviceB200& dv = devs[k];
            if(twoLevel) { d.scene = dv.scene; d.accel = nullptr; d.instanceVertexTBN = instTBN.data(); if(!flatTextures.empty()) d.instanceVertexUVs = instUVs.data(); }
            else { d.accel = dv.accel; d.vertexCount = pg.vertexTotal; d.triangleCount = pg.primTotal; d.vertexTBN = tbn; if(!flatTextures.empty()) d.vertexUVs = instUVs[0]; }
            d.spectrum = (r.type == "(R)PathTracerSpectral") ? dv.spectrum : nullptr;
            // throughput mode: the renderer's initial pass is this device's share of the job's samples; pass modes
            // (re)define the range with every pass
            auto [b, e] = Share(jobSamples, nDev, k);
            d.sampleOffset = jobBegin + b; d.totalSPP = std::max(e - b, 1u);
            CheckOn(dv, mrb_renderer_create(dv.ctx, &d, &dv.renderer));
            if(!throughputSingle || e == b)
                CheckOn(dv, mrb_renderer_set_spp_limit(dv.ctx, dv.renderer, 0));   // pass modes start empty (DoRenderWork begins the passes); so does a device without samples
        }
        curRenderer = Raw(id);
        const size_t need = size_t(4) * pixels * sizeof(float);
        if(need > stagingBytes)
        {
            if(staging) { mrb_host_free(ctx, staging); staging = nullptr; stagingBytes = 0; }
            void* ptr = nullptr;
            Check(mrb_host_alloc(ctx, need, &ptr));
            staging = static_cast<float*>(ptr); stagingBytes = need;
        }
        std::memset(staging, 0, need);
        return RenderBufferInfo
        {
            .data = reinterpret_cast<const Byte*>(staging), .totalSize = need,
            .renderColorSpace = params.globalTextureColorSpace, .resolution = rip.resolution,
            .curRenderLogic0 = logic0.value_or(0), .curRenderLogic1 = 0
        };
    }
    // RendererI::SetCameraTransform (Tracer/PathTracerRendererBase.cu:L495-512): taken up by the next DoRenderWork, which
    // restarts the accumulation with position / gaze point / up replaced
    void SetCameraTransform(RendererId id, CameraTransform t) override
    {
        if(Raw(id) >= renderers.size()) throw MRayError("Unable to find Renderer({})", Raw(id));
        pendingCam = t;
    }
    void StopRender() override { ReleaseRenderers(); }

    static void ReleaseSemaphore(void* user) { static_cast<TimelineSemaphore*>(user)->Release(); }

    // the pass DoLatencyRender would run next on the current tile (TracerDLL/PathTracerRenderer.cu:L1100-1112)
    PassPlan PlanPass(const RendererB200& rr, uint32_t jobSamples) const
    {
        const uint32_t passCount = rr.latency ? 1u : std::max(rr.burstSize, 1u);
        PassPlan p; p.tile = currentTile; p.tileSPP = tileSPPs[currentTile];
        uint32_t sppLimit = p.tileSPP + passCount;
        if(saveImage) sppLimit = std::min(sppLimit, jobSamples);
        p.count = sppLimit - p.tileSPP;
        return p;
    }
    // every device moves to the pass's tile (the film layout follows the region) and takes its share of the samples
    void BeginPassOnAll(const PassPlan& plan)
    {
        const Vector2ui ts = TileStart(plan.tile), te = TileEnd(plan.tile);
        const uint32_t rm[2] = {regionMin[0] + ts[0], regionMin[1] + ts[1]}, rs[2] = {te[0] - ts[0], te[1] - ts[1]};
        const uint32_t nDev = uint32_t(devs.size());
        for(uint32_t k = 0; k < nDev; k++)
        {
            auto [b, e] = Share(plan.count, nDev, k);
            CheckOn(devs[k], mrb_renderer_begin_pass(devs[k].ctx, devs[k].renderer, rm, rs, jobBegin + plan.tileSPP + b, e - b));
        }
    }

    // Runs `work(deviceIndex)` for every device concurrently (device 0 on the calling thread) and rethrows the first failure
    template<class F> void OnAllDevices(F&& work)
    {
        const size_t n = devs.size();
        if(n == 1) { work(0u); return; }
        std::vector<std::string> errors(n);
        std::vector<std::thread> threads;
        for(size_t k = 1; k < n; k++)
            threads.emplace_back([&, k]() { try { work(uint32_t(k)); } catch(const MRayError& e) { errors[k] = std::string(e.GetError()); if(errors[k].empty()) errors[k] = "error"; } });
        try { work(0u); } catch(const MRayError& e) { errors[0] = std::string(e.GetError()); if(errors[0].empty()) errors[0] = "error"; }
        for(std::thread& t : threads) t.join();
        for(const std::string& e : errors) if(!e.empty()) throw MRayError("{}", e);
    }

    RendererOutput DoRenderWork() override
    {
        if(!devs[0].renderer) return RendererOutput{};
        if(pendingCam && lastStart)
        {
            camOverride = pendingCam; pendingCam.reset();
            rebuilding = true;
            try { StartRender(lastStart->id, lastStart->camSurf, lastStart->rip, lastStart->logic0, std::nullopt); }
            catch(...) { rebuilding = false; throw; }
            rebuilding = false;
        }
        const auto t0 = std::chrono::steady_clock::now();
        const RendererB200& rr = renderers[curRenderer];
        const uint32_t nDev = uint32_t(devs.size());
        const uint32_t jobSamples = jobEnd - jobBegin;
        Vector2ui secMin, secMax;
        uint64_t passPaths = 0;
        bool triggerSave = false;
        if(throughputSingle)
        {
            // DoThroughputSingleTileRender (TracerDLL/PathTracerRenderer.cu:L1010-1076): one wavefront iteration, then the
            // film delta. The reference synchronises to count the dead paths; here the counters of EARLIER iterations
            // are polled from pinned memory, so triggerSave follows the last path by an iteration or two (empty ones).
            secMin = regionMin; secMax = regionMin + regionSize;
            std::vector<mrb_render_stats> st(nDev);
            OnAllDevices([&](uint32_t k)
            {
                CheckOn(devs[k], mrb_renderer_iterate(devs[k].ctx, devs[k].renderer, 1));
                CheckOn(devs[k], mrb_renderer_poll_stats(devs[k].ctx, devs[k].renderer, &st[k]));
            });
            uint64_t done = 0; bool all = true;
            for(uint32_t k = 0; k < nDev; k++) { done += st[k].pathsCompleted; all = all && st[k].finished; }
            passPaths = done - completedPaths; completedPaths = done;
            triggerSave = saveImage && all;
        }
        else
        {
            // DoLatencyRender(passCount) (TracerDLL/PathTracerRenderer.cu:L1078-1160): passCount more samples of every pixel
            // of the current tile, traced to completion; then the tile's film is handed over and the next tile is up
            const PassPlan plan = PlanPass(rr, jobSamples);
            const Vector2ui ts = TileStart(plan.tile), te = TileEnd(plan.tile);
            secMin = regionMin + ts; secMax = regionMin + te;
            passPaths = uint64_t(plan.count) * (te[0] - ts[0]) * (te[1] - ts[1]);
            if(!(primed && primedPlan == plan)) BeginPassOnAll(plan);
            primed = false;
            OnAllDevices([&](uint32_t k)
            {
                auto [b, e] = Share(plan.count, nDev, k);
                if(e == b) return;
                mrb_render_stats st;
                CheckOn(devs[k], mrb_renderer_run_pass(devs[k].ctx, devs[k].renderer, 4, &st));
            });
            tileSPPs[plan.tile] = plan.tileSPP + plan.count;
            completedPaths += passPaths;
            currentTile = (currentTile + 1) % uint32_t(tileSPPs.size());
            uint64_t sum = 0; for(uint32_t v : tileSPPs) sum += v;
            triggerSave = saveImage && sum == uint64_t(jobSamples) * tileSPPs.size();
        }
        if(triggerSave) saveImage = false;
        // multi-GPU: the peers' films are added to device 0's (and cleared) over NVLink peer memory
        if(nDev > 1)
        {
            std::vector<mrb_context> pc; std::vector<mrb_renderer> prs;
            for(uint32_t k = 1; k < nDev; k++) { pc.push_back(devs[k].ctx); prs.push_back(devs[k].renderer); }
            Check(mrb_renderer_reduce_peers(ctx, devs[0].renderer, pc.data(), prs.data(), uint32_t(prs.size())));
        }
        // RenderImage::TransferToHost (Tracer/RenderImage.cpp:L163-219): wait until the caller has read the previous
        // section, queue the copy into the pinned staging buffer on the copy stream, release the semaphore from a host
        // callback when it has landed; the next DoRenderWork's kernels overlap the copy (second film buffer)
        if(!sem->Acquire(acquireValue)) return RendererOutput{};
        Check(mrb_renderer_film_handoff(ctx, devs[0].renderer, staging, &TracerB200::ReleaseSemaphore, sem));
        acquireValue += 2;
        if(!throughputSingle && !triggerSave)
        {
            // Keep the devices busy while the caller consumes this section: begin the NEXT pass now and queue its first
            // iterations (they accumulate into the second film buffer); the next DoRenderWork picks the pass up.
            const PassPlan next = PlanPass(rr, jobSamples);
            if(next.count > 0)
            {
                BeginPassOnAll(next);
                for(uint32_t k = 0; k < nDev; k++)
                {
                    auto [b, e] = Share(next.count, nDev, k);
                    if(e > b) CheckOn(devs[k], mrb_renderer_iterate(devs[k].ctx, devs[k].renderer, 8));
                }
                primed = true; primedPlan = next;
            }
        }
        const Vector2ui secSize = secMax - secMin;
        size_t plane = size_t(secSize[0]) * secSize[1] * sizeof(float);
        RendererOutput out;
        out.imageOut = RenderImageSection
        {
            .pixelMin = secMin, .pixelMax = secMax, .globalWeight = Float(1),
            .waitCounter = acquireValue - 1,
            .pixStartOffsets = {0, plane, 2 * plane}, .weightStartOffset = 3 * plane
        };
        // CalculateAnalyticDataThroughput / Latency (Tracer/PathTracerRendererBase.cu:L301-360): paths completed by this
        // call over the CPU time of the call
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        size_t usedMem = 0; for(const DeviceB200& dv : devs) usedMem += mrb_context_used_device_memory(dv.ctx);
        out.analytics = RendererAnalyticData
        {
            .throughput = (sec > 0.0) ? double(passPaths) / sec / 1.0e6 : 0.0, .throughputSuffix = "M path/s",
            .workPerPixel = double(completedPaths) / (double(regionSize[0]) * double(regionSize[1])),
            .wppLimit = double(jobSamples), .workPerPixelSuffix = "spp",
            .iterationTimeMS = float(sec * 1.0e3), .renderResolution = fullResolution,
            .outputColorSpace = params.globalTextureColorSpace,
            .usedGPUMemoryBytes = usedMem
        };
        out.triggerSave = triggerSave;
        return out;
    }

    // ------------------------------- misc -------------------------------
    void ClearAll() override
    {
        // TracerBase::ClearAll (Tracer/TracerBase.cpp): back to the freshly constructed state, implicit groups only
        ReleaseRenderers();
        ReleaseAccels();
        prims.resize(1); mats.resize(1); lights.resize(1); transforms.resize(1);
        cams.clear(); renderers.clear();
        surfaces.clear(); lightSurfaces.clear(); camSurfaces.clear(); volumes.clear(); textures.clear();
        boundary = LightSurfaceParams{};
        flatPrimGroup = 0; flatLightRadiance.clear(); flatLightTwoSided.clear(); flatAlbedo.clear();
        flatMaterialType.clear(); flatMaterialParams.clear(); flatAlbedoTex.clear(); flatNormalTex.clear(); flatTextures.clear();
        lastStart.reset(); camOverride.reset(); pendingCam.reset(); tileSPPs.clear(); currentTile = 0;
    }
    void Flush() const override { for(const DeviceB200& d : devs) mrb_context_synchronize(d.ctx); }
    GPUThreadInitFunction GetThreadInitFunction() const override { return []() {}; } // the C-ABI selects its device per call
    void SetThreadPool(ThreadPool& tp) override { pool = &tp; }
    size_t TotalDeviceMemory() const override { return mrb_context_total_device_memory(ctx); }
    size_t UsedDeviceMemory() const override { size_t u = 0; for(const DeviceB200& d : devs) u += mrb_context_used_device_memory(d.ctx); return u; }
    const TracerParameters& Parameters() const override { return params; }
};

} // namespace

extern "C" __attribute__((visibility("default"))) TracerI* ConstructTracer(const TracerParameters& p) { return new TracerB200(p); }
extern "C" __attribute__((visibility("default"))) void DestroyTracer(TracerI* t) { delete t; }
