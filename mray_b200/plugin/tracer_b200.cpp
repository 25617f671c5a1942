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

#include <algorithm>
#include <cmath>
#include <cstring>
#include <array>
#include <cmath>
#include <mutex>
#include <string>
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
    std::vector<Vector3> positions; std::vector<Vector3> normals; std::vector<Vector2> uvs; std::vector<Vector3ui> indices;
    uint32_t primTotal = 0, vertexTotal = 0;
};
struct MatGroupB200 { std::string type; bool committed = false; std::vector<Vector3> albedo; std::vector<int32_t> albedoTex; /* TextureId or -1 */ };
// One 2-D texture as TracerI::CreateTexture2D / PushTextureData deliver it (first slice: one mip level, 4-channel
// fp32 or unorm8 pixels, already in the global colour space)
struct TextureB200
{
    Vector2ui size; MRayTextureParameters params; uint32_t channels = 4, format = 0;
    std::vector<Byte> pixels; bool loaded = false;
};
struct LightGroupB200
{
    std::string type; bool committed = false; uint32_t primGroup = 0;
    std::vector<Vector3> radiance; std::vector<uint8_t> twoSided; std::vector<uint32_t> primBatch;
};
struct TransGroupB200 { std::string type; bool committed = false; std::vector<Matrix3x4> matrices; };
struct CamGroupB200 { std::string type; bool committed = false; std::vector<Vector4> fovPlanes; std::vector<Vector3> gaze, position, up; };
struct RendererB200
{
    std::string type;
    uint32_t totalSPP = 16384, burstSize = 1, sampleMode = 0; Vector2ui rrRange = Vector2ui(4, 20);
    bool latency = false;      // renderMode "Latency": every DoRenderWork finishes burstSize samples per pixel
};

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

class TracerB200 final : public TracerI
{
    TracerParameters params;
    mrb_context ctx = nullptr;
    mrb_accel accel = nullptr;           // all (T)Identity surfaces
    std::vector<mrb_accel> instAccels;   // two-level scenes: one per distinct (prim ranges, cull flags) set
    uint32_t sceneInstanceCount = 0, uniqueAccelCount = 0;
    mrb_scene scene = nullptr;           // set when any surface is transformed
    mrb_renderer renderer = nullptr;
    mrb_spectrum spectrum = nullptr;     // SpectrumContextJakob2019 of params.globalTextureColorSpace, made on first use
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
    std::vector<int32_t> flatAlbedoTex;             // per flat material: index into flatTextures or -1
    std::vector<uint32_t> flatTextures;             // TextureIds in use, in first-use order
    std::vector<TextureB200> textures;              // TextureId = index + 1 (0 = InvalidTexture)
    // render hand-off
    TimelineSemaphore* sem = nullptr; uint64_t acquireValue = 0;
    std::vector<float> staging; Vector2ui resolution = Vector2ui::Zero(), regionMin = Vector2ui::Zero();
    struct StartArgs { RendererId id; CamSurfaceId camSurf; RenderImageParams rip; Optional<uint32_t> logic0; };
    Optional<StartArgs> lastStart; Optional<CameraTransform> camOverride, pendingCam; bool rebuilding = false;
    uint32_t latencySPP = 0;
    uint32_t curRenderer = 0; ThreadPool* pool = nullptr;

    void Check(mrb_status s) const { if(s != MRB_OK) throw MRayError("{}", mrb_last_error(ctx)); }
    template<class G> G& Get(std::vector<G>& v, uint32_t id, std::string_view what)
    { if(id >= v.size()) throw MRayError("Unable to find {}({})", what, id); return v[id]; }
    template<class G> const G& Get(const std::vector<G>& v, uint32_t id, std::string_view what) const
    { if(id >= v.size()) throw MRayError("Unable to find {}({})", what, id); return v[id]; }

    public:
    explicit TracerB200(const TracerParameters& p) : params(p)
    {
        if(mrb_abi_version() != MRB_ABI_VERSION)   // descriptor layouts must match the header this object was compiled against
            throw MRayError("mray_b200: libmray_b200.so has ABI {:#x}, the plugin was built for {:#x}", mrb_abi_version(), uint32_t(MRB_ABI_VERSION));
        mrb_status s = mrb_context_create(0, &ctx);
        if(s != MRB_OK) throw MRayError("mray_b200: {}", mrb_last_error(nullptr)); // no CPU fallback
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
        if(spectrum) return;
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
        Check(mrb_spectrum_create(ctx, &d, &spectrum));
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
    void ReleaseAccels()
    {
        if(scene) { mrb_scene_destroy(ctx, scene); scene = nullptr; }
        for(mrb_accel a : instAccels) mrb_accel_destroy(ctx, a);
        instAccels.clear();
        if(accel) { mrb_accel_destroy(ctx, accel); accel = nullptr; }
    }
    ~TracerB200() override
    {
        if(renderer) mrb_renderer_destroy(ctx, renderer);
        if(spectrum) mrb_spectrum_destroy(ctx, spectrum);
        ReleaseAccels();
        mrb_context_destroy(ctx);
    }

    // ------------------------------- generic -------------------------------
    TypeNameList PrimitiveGroups() const override { return {"(P)Triangle"sv, "(P)Empty"sv}; }
    TypeNameList MaterialGroups() const override { return {"(Mt)Lambert"sv, "(Mt)Passthrough"sv, "(Mt)Reflect"sv}; }
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
        pg.normals.assign(pg.vertexTotal, Vector3::Zero());
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
                      // the attribute is the to-tangent-space rotation; its Z basis is the shading normal
                      for(size_t i = 0; i < s.size(); i++) pg.normals[pb.vertexOffset + i] = s[i].OrthoBasisZ(); break; }
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
        if(typeName != "(Mt)Lambert" && typeName != "(Mt)Reflect") throw MRayError("Unable to find generator for {}", typeName);
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
            mg.albedo.push_back(Vector3::Zero()); mg.albedoTex.push_back(-1);
            out.push_back(MaterialId((Raw(g) << MAT_ID_BITS) | uint32_t(mg.albedo.size() - 1)));
        }
        return out;
    }
    void CommitMatReservations(MatGroupId g) override { Get(mats, Raw(g), "MaterialGroup").committed = true; }
    bool IsMatCommitted(MatGroupId g) const override { return Get(mats, Raw(g), "MaterialGroup").committed; }
    void PushMatAttribute(MatGroupId, CommonIdRange, uint32_t attributeIndex, TransientData) override
    { throw MRayError("(Mt)Lambert: Attribute {:d} is not \"ConstantOnly\", wrong function is called", attributeIndex); }
    void PushMatAttribute(MatGroupId g, CommonIdRange range, uint32_t attributeIndex, TransientData data,
                          std::vector<Optional<TextureId>> tex) override
    {
        MatGroupB200& mg = Get(mats, Raw(g), "MaterialGroup");
        if(mg.type == "(Mt)Reflect") throw MRayError("{} group does not have any attributes!", mg.type);
        // An EMPTY TransientData routes to the optional texture-only overload (TracerBase.cpp:L843-867): the
        // scene loader always pushes Lambert's optional `normalMap` this way, with nullopt where a material has none.
        if(data.IsEmpty())
        {
            if(attributeIndex != 1) throw MRayError("{}: Attribute {:d} is not \"Optional Texture\"", mg.type, attributeIndex);
            for(const auto& t : tex) if(t.has_value()) throw MRayError("{}: normal maps are not supported yet", mg.type);
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
        if(mipCount != 1) throw MRayError("textures: mip chains are not supported yet (mipCount {})", mipCount);
        if(size[0] == 0 || size[1] == 0) throw MRayError("textures: empty texture");
        TextureB200 t; t.size = size; t.params = p;
        switch(p.pixelType.Name())
        {
            case MRayPixelEnum::MR_RGBA_FLOAT:  t.channels = 4; t.format = 0; break;
            case MRayPixelEnum::MR_RGBA8_UNORM: t.channels = 4; t.format = 1; break;
            default: throw MRayError("textures: only MR_RGBA_FLOAT and MR_RGBA8_UNORM pixels are supported yet");
        }
        // TextureMemory::ConvertColorspaces leaves a texture alone when it is not a colour, or already global + linear
        const bool needsConversion = p.isColor == AttributeIsColor::IS_COLOR &&
            ((p.colorSpace != MRayColorSpaceEnum::MR_DEFAULT && p.colorSpace != params.globalTextureColorSpace) || p.gamma != Float(1));
        if(needsConversion) throw MRayError("textures: colour space / gamma conversion is not supported yet");
        // the view type follows DetermineReadMode (Tracer/TextureMemory.cpp:L329-396): 4 channels + MR_DROP_1 read as Vector3
        if(p.readMode != MRayTextureReadMode::MR_PASSTHROUGH && p.readMode != MRayTextureReadMode::MR_DROP_1)
            throw MRayError("textures: only MR_PASSTHROUGH / MR_DROP_1 reads are supported yet");
        if(params.genMips) throw MRayError("textures: genMips is not supported yet");
        if(!p.ignoreResClamp && std::max(size[0], size[1]) > params.clampedTexRes) throw MRayError("textures: clampedTexRes is not supported yet");
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
        if(mipLevel != 0) throw MRayError("textures: mip level {} of a single-level texture", mipLevel);
        // the TransientData is typed by the pixel (MRayPixelType<E>::Type): Vector4 / Vector4uc here
        const size_t pixels = size_t(t.size[0]) * t.size[1];
        const Byte* src = nullptr; size_t got = 0;
        if(t.format == 0) { auto s = data.AccessAs<const Vector4>(); src = reinterpret_cast<const Byte*>(s.data()); got = s.size(); }
        else { auto s = data.AccessAs<const Vector4uc>(); src = reinterpret_cast<const Byte*>(s.data()); got = s.size(); }
        if(got != pixels) throw MRayError("textures: {} pixels pushed, {} expected", got, pixels);
        t.pixels.assign(src, src + pixels * t.channels * (t.format == 0 ? 4u : 1u));
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
        if(typeName != "(L)Prim(P)Triangle") throw MRayError("Unable to find generator for {}", typeName);
        lights.push_back(LightGroupB200{typeName, false, Raw(pg)});
        return LightGroupId(uint32_t(lights.size() - 1));
    }
    LightId ReserveLight(LightGroupId g, AttributeCountList c, PrimBatchId b) override { return ReserveLights(g, {c}, {b}).front(); }
    LightIdList ReserveLights(LightGroupId g, std::vector<AttributeCountList> counts, std::vector<PrimBatchId> batches) override
    {
        std::lock_guard lk(mtx);
        LightGroupB200& lg = Get(lights, Raw(g), "LightGroup");
        if(batches.size() != counts.size()) throw MRayError("{}: prim-backed lights need one prim batch each", lg.type);
        LightIdList out;
        for(size_t i = 0; i < counts.size(); i++)
        {
            lg.radiance.push_back(Vector3::Zero()); lg.twoSided.push_back(0);
            lg.primBatch.push_back(Raw(batches[i]) & ((1u << PRIM_ID_BITS) - 1u));
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
        for(const auto& t : tex) if(t.has_value()) throw MRayError("{}: textured radiance is not supported yet", lg.type);
        uint32_t lo = range[0] & ((1u << MAT_ID_BITS) - 1u), hi = range[1] & ((1u << MAT_ID_BITS) - 1u);
        auto s = data.AccessAs<const Vector3>();
        if(hi >= lg.radiance.size() || s.size() != hi - lo + 1) throw MRayError("{}: radiance range mismatch", lg.type);
        std::copy(s.begin(), s.end(), lg.radiance.begin() + lo);
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
        if(Raw(boundary.lightId) != 0) throw MRayError("boundary lights other than (L)Null are not supported yet");
        struct Group { uint32_t transformId; std::vector<uint32_t> ranges, lmKeys; std::vector<uint8_t> cull; };
        std::vector<Group> groups;
        auto GroupOf = [&](TransformId t) -> Group&
        {
            for(Group& g : groups) if(g.transformId == Raw(t)) return g;
            const TransGroupB200& tg = Get(transforms, Raw(t) >> TRANS_ID_BITS, "TransformGroup");
            if((Raw(t) & ((1u << TRANS_ID_BITS) - 1u)) >= tg.matrices.size()) throw MRayError("Unable to find Transform({})", Raw(t));
            groups.push_back(Group{Raw(t), {}, {}, {}});
            return groups.back();
        };
        flatAlbedo.clear(); flatAlbedoTex.clear(); flatMaterialType.clear(); flatTextures.clear(); flatLightRadiance.clear(); flatLightTwoSided.clear();
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
            flatMaterialType.push_back(mg.type == "(Mt)Reflect" ? uint8_t(MRB_MATERIAL_REFLECT) : uint8_t(MRB_MATERIAL_LAMBERT));
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
            grp.cull.push_back(0);
            flatLightRadiance.insert(flatLightRadiance.end(), {lg.radiance[li][0], lg.radiance[li][1], lg.radiance[li][2]});
            flatLightTwoSided.push_back(lg.twoSided[li]);
        }
        if(pgUsed < 0 || groups.empty()) throw MRayError("empty scene");
        flatPrimGroup = uint32_t(pgUsed);
        const PrimGroupB200& pg = prims[flatPrimGroup];
        ReleaseAccels();
        auto BuildGroup = [&](const Group& g)
        {
            mrb_accel_desc d = {};
            d.positions = reinterpret_cast<const float*>(pg.positions.data()); d.vertexCount = pg.vertexTotal;
            d.indices = reinterpret_cast<const uint32_t*>(pg.indices.data()); d.triangleCount = pg.primTotal;
            d.memspace = MRB_MEM_HOST; d.primGroupId = flatPrimGroup;
            d.rangeCount = uint32_t(g.lmKeys.size()); d.primRanges = g.ranges.data();
            d.lightOrMatKeys = g.lmKeys.data(); d.cullBackface = g.cull.data(); d.flags = MRB_BUILD_DEFAULT;
            mrb_accel a = nullptr;
            Check(mrb_accel_build(ctx, &d, &a));
            return a;
        };
        AABB3 aabb;
        if(groups.size() == 1 && groups[0].transformId == 0)
        {
            accel = BuildGroup(groups[0]);
            mrb_accel_info info; Check(mrb_accel_get_info(ctx, accel, &info));
            aabb = AABB3(Vector3(info.aabb[0], info.aabb[1], info.aabb[2]), Vector3(info.aabb[3], info.aabb[4], info.aabb[5]));
        }
        else
        {
            // Groups with the same prim ranges and cull flags share ONE accelerator and differ only in transform and
            // LightOrMatKeys — the reference's concrete-accelerator / instance split (Tracer/AcceleratorC.h:L780-905).
            std::vector<mrb_instance_desc> inst(groups.size());
            std::vector<size_t> builtFor;   // group index each unique accelerator was built from
            for(size_t k = 0; k < groups.size(); k++)
            {
                mrb_accel a = nullptr;
                for(size_t u = 0; u < builtFor.size() && !a; u++)
                    if(groups[builtFor[u]].ranges == groups[k].ranges && groups[builtFor[u]].cull == groups[k].cull) a = instAccels[u];
                if(!a) { a = BuildGroup(groups[k]); instAccels.push_back(a); builtFor.push_back(k); }
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
            Check(mrb_scene_build(ctx, inst.data(), uint32_t(inst.size()), &scene));
            sceneInstanceCount = uint32_t(inst.size()); uniqueAccelCount = uint32_t(instAccels.size());
            float box[6];
            Check(mrb_scene_export_tlas(ctx, scene, nullptr, box, nullptr, nullptr, nullptr, nullptr));
            aabb = AABB3(Vector3(box[0], box[1], box[2]), Vector3(box[3], box[4], box[5]));
        }
        return SurfaceCommitResult
        {
            .aabb = aabb,
            .instanceCount = surfaces.size() + lightSurfaces.size(),
            .acceleratorCount = scene ? uniqueAccelCount : 1u
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
        if(Raw(id) == curRenderer && renderer) { mrb_renderer_destroy(ctx, renderer); renderer = nullptr; }
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
    RenderBufferInfo StartRender(RendererId id, CamSurfaceId camSurf, RenderImageParams rip, Optional<uint32_t> logic0, Optional<uint32_t>) override
    {
        if(!sem) throw MRayError("Render environment is not set properly! Please provide a semaphore to the tracer.");
        if(!accel && !scene) throw MRayError("CommitSurfaces must be called before StartRender");
        if(!rebuilding) { camOverride.reset(); pendingCam.reset(); }
        lastStart = StartArgs{id, camSurf, rip, logic0};
        const RendererB200& r = Get(renderers, Raw(id), "Renderer");
        const CameraSurfaceParams& cs = Get(camSurfaces, Raw(camSurf), "CameraSurface");
        const CamGroupB200& cg = Get(cams, Raw(cs.cameraId) >> CAM_ID_BITS, "CameraGroup");
        uint32_t ci = Raw(cs.cameraId) & ((1u << CAM_ID_BITS) - 1u);
        const PrimGroupB200& pg = prims[flatPrimGroup];
        if(renderer) { mrb_renderer_destroy(ctx, renderer); renderer = nullptr; }
        Vector2ui tile = rip.regionMax - rip.regionMin;
        if(rip.regionMax[0] > rip.resolution[0] || rip.regionMax[1] > rip.resolution[1] || tile[0] == 0 || tile[1] == 0 ||
           rip.regionMin[0] >= rip.regionMax[0] || rip.regionMin[1] >= rip.regionMax[1])
            throw MRayError("StartRender: bad render region");
        mrb_render_desc d = {};
        bool hasNormals = std::any_of(pg.normals.begin(), pg.normals.end(), [](const Vector3& n) { return n != Vector3::Zero(); });
        const float* normals = hasNormals ? reinterpret_cast<const float*>(pg.normals.data()) : nullptr;
        std::vector<const float*> instNormals(scene ? sceneInstanceCount : 1u, normals);
        if(scene) { d.scene = scene; d.instanceVertexNormals = instNormals.data(); }
        else { d.accel = accel; d.vertexCount = pg.vertexTotal; d.triangleCount = pg.primTotal; d.vertexNormals = normals; }
        d.materialCount = uint32_t(flatAlbedo.size() / 3); d.albedo = flatAlbedo.data();
        d.materialType = flatMaterialType.data();
        std::vector<mrb_texture_desc> texDescs(flatTextures.size());
        std::vector<const float*> instUVs(scene ? sceneInstanceCount : 1u, reinterpret_cast<const float*>(pg.uvs.data()));
        if(!flatTextures.empty())
        {
            for(size_t k = 0; k < flatTextures.size(); k++)
            {
                const TextureB200& t = textures[flatTextures[k] - 1];
                texDescs[k] = mrb_texture_desc{t.pixels.data(), t.size[0], t.size[1], t.channels, t.format,
                                               uint32_t(t.params.interpolation), uint32_t(t.params.edgeResolve)};
            }
            d.textureCount = uint32_t(texDescs.size()); d.textures = texDescs.data(); d.albedoTexture = flatAlbedoTex.data();
            if(scene) d.instanceVertexUVs = instUVs.data(); else d.vertexUVs = instUVs[0];
        }
        d.lightCount = uint32_t(flatLightTwoSided.size()); d.lightRadiance = flatLightRadiance.data(); d.lightTwoSided = flatLightTwoSided.data();
        for(int k = 0; k < 3; k++) { d.camPosition[k] = cg.position[ci][k]; d.camGaze[k] = cg.gaze[ci][k]; d.camUp[k] = cg.up[ci][k]; }
        if(camOverride)   // SetCameraTransform: position / gaze point / up replace the camera's own, fov and planes stay
            for(int k = 0; k < 3; k++) { d.camPosition[k] = camOverride->position[k]; d.camGaze[k] = camOverride->gazePoint[k]; d.camUp[k] = camOverride->up[k]; }
        d.fovXY[0] = cg.fovPlanes[ci][0]; d.fovXY[1] = cg.fovPlanes[ci][1];
        d.nearFar[0] = cg.fovPlanes[ci][2]; d.nearFar[1] = cg.fovPlanes[ci][3];
        d.width = tile[0]; d.height = tile[1]; d.totalSPP = r.totalSPP;
        d.fullResolution[0] = rip.resolution[0]; d.fullResolution[1] = rip.resolution[1];
        d.regionMin[0] = rip.regionMin[0]; d.regionMin[1] = rip.regionMin[1];
        // render logic 0 rolls the sample mode like PathTracerRendererT::StartRender (L1192-1200)
        d.sampleMode = (r.sampleMode + logic0.value_or(0)) % 3u;
        d.rrRange[0] = r.rrRange[0]; d.rrRange[1] = r.rrRange[1];
        d.filmFilterRadius = params.filmFilter.radius; d.seed = params.seed;
        uint64_t pixels = uint64_t(tile[0]) * tile[1];
        d.maxPathCount = uint32_t(std::min<uint64_t>(pixels, std::max<uint32_t>(params.parallelizationHint, 1u)));
        // every material of this plugin is shaded by one fused kernel, so the material-key ray sort
        // (RayPartitioner::MultiPartition) only costs: 8.74 -> 7.74 ms/spp at 1080p without it. It stays
        // available (MRB_PARTITION_RAYS=1) for parity with the reference's per-material work batches.
        const char* pr = std::getenv("MRB_PARTITION_RAYS");
        d.partitionRays = (pr && pr[0] == '1') ? 1u : 0u;
        if(r.type == "(R)PathTracerSpectral") { EnsureSpectrum(); d.spectrum = spectrum; }
        if(params.samplerType.e != SamplerType::INDEPENDENT)
        {
            // the scramble's final bit reversal is applied unless MRB_REFERENCE_SCRAMBLE=1 (see include/mray_b200.h)
            EnsureSobolMatrices();
            d.samplerType = (params.samplerType.e == SamplerType::SOBOL) ? MRB_SAMPLER_SOBOL : MRB_SAMPLER_ZSOBOL;
            const char* rs = std::getenv("MRB_REFERENCE_SCRAMBLE");
            if(rs && rs[0] == '1') d.samplerType |= MRB_SAMPLER_REFERENCE_SCRAMBLE;
            d.sobolMatrices = sobolMatrices.data();
        }
        Check(mrb_renderer_create(ctx, &d, &renderer));
        latencySPP = 0;
        if(r.latency || r.burstSize > 1) Check(mrb_renderer_set_spp_limit(ctx, renderer, 0));   // pass mode, see DoRenderWork
        curRenderer = Raw(id); resolution = tile; regionMin = rip.regionMin;
        staging.assign(size_t(4) * pixels, 0.0f);
        return RenderBufferInfo
        {
            .data = reinterpret_cast<const Byte*>(staging.data()), .totalSize = staging.size() * sizeof(float),
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
    void StopRender() override { if(renderer) { mrb_renderer_destroy(ctx, renderer); renderer = nullptr; } }
    RendererOutput DoRenderWork() override
    {
        if(!renderer) return RendererOutput{};
        if(pendingCam && lastStart)
        {
            camOverride = pendingCam; pendingCam.reset();
            rebuilding = true;
            try { StartRender(lastStart->id, lastStart->camSurf, lastStart->rip, lastStart->logic0, std::nullopt); }
            catch(...) { rebuilding = false; throw; }
            rebuilding = false;
        }
        const RendererB200& rr = renderers[curRenderer];
        mrb_render_stats st;
        if(rr.latency || rr.burstSize > 1)
        {
            // PathTracerRendererBase::DoRender (Tracer/PathTracerRendererBase.cu:L515-533): renderMode Latency runs
            // DoLatencyRender(1), Throughput with burstSize > 1 runs DoLatencyRender(burstSize) — that many more samples
            // of every pixel, traced to completion before the film is handed over (PathTracerRenderer.cu:L1078-1160)
            latencySPP = std::min(rr.totalSPP, latencySPP + (rr.latency ? 1u : rr.burstSize));
            Check(mrb_renderer_set_spp_limit(ctx, renderer, latencySPP));
            do { Check(mrb_renderer_iterate(ctx, renderer, 4)); Check(mrb_renderer_get_stats(ctx, renderer, &st)); } while(!st.finished);
            st.finished = (latencySPP >= rr.totalSPP) ? 1u : 0u;
        }
        else
        {
            // one wavefront iteration (DoThroughputSingleTileRender), then the film delta hand-off of
            // RenderImage::TransferToHost (Tracer/RenderImage.cpp:L163-219): acquire, copy, release, next state
            Check(mrb_renderer_iterate(ctx, renderer, 1));
            Check(mrb_renderer_get_stats(ctx, renderer, &st));
        }
        if(!sem->Acquire(acquireValue)) return RendererOutput{};
        Check(mrb_renderer_read_film(ctx, renderer, staging.data(), MRB_MEM_HOST, 1));
        sem->Release();
        acquireValue += 2;
        size_t plane = size_t(resolution[0]) * resolution[1] * sizeof(float);
        RendererOutput out;
        out.imageOut = RenderImageSection
        {
            .pixelMin = regionMin, .pixelMax = regionMin + resolution, .globalWeight = Float(1),
            .waitCounter = acquireValue - 1,
            .pixStartOffsets = {0, plane, 2 * plane}, .weightStartOffset = 3 * plane
        };
        out.analytics = RendererAnalyticData
        {
            .throughput = 0.0, .throughputSuffix = "M path/s",
            .workPerPixel = double(st.pathsCompleted) / double(resolution[0] * resolution[1]),
            .wppLimit = double(renderers[curRenderer].totalSPP), .workPerPixelSuffix = "spp",
            .iterationTimeMS = 0.0f, .renderResolution = resolution,
            .outputColorSpace = params.globalTextureColorSpace,
            .usedGPUMemoryBytes = mrb_context_used_device_memory(ctx)
        };
        out.triggerSave = st.finished != 0;
        return out;
    }

    // ------------------------------- misc -------------------------------
    void ClearAll() override
    {
        StopRender();
        if(accel) { mrb_accel_destroy(ctx, accel); accel = nullptr; }
        prims.resize(1); mats.resize(1); lights.resize(1); cams.clear(); renderers.clear();
        surfaces.clear(); lightSurfaces.clear(); camSurfaces.clear(); volumes.clear(); textures.clear();
    }
    void Flush() const override { mrb_context_synchronize(ctx); }
    GPUThreadInitFunction GetThreadInitFunction() const override { return []() {}; } // the C-ABI selects its device per call
    void SetThreadPool(ThreadPool& tp) override { pool = &tp; }
    size_t TotalDeviceMemory() const override { return mrb_context_total_device_memory(ctx); }
    size_t UsedDeviceMemory() const override { return mrb_context_used_device_memory(ctx); }
    const TracerParameters& Parameters() const override { return params; }
};

} // namespace

extern "C" __attribute__((visibility("default"))) TracerI* ConstructTracer(const TracerParameters& p) { return new TracerB200(p); }
extern "C" __attribute__((visibility("default"))) void DestroyTracer(TracerI* t) { delete t; }
