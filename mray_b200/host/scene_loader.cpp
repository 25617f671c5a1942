// scene_loader.cpp — libSceneLoaderB200.so: an MRay scene loader (Core/SceneLoaderI.h) for the JSON scene format of
// Docs/markdown/scene/mrayScene.md, written against the reference's public headers like the TracerI plugin. It replaces
// Source/SceneLoaderMRay (SceneLoaderMRay.cpp:L335-2514) for the part of the format the B200 tracer renders — SURVEY.md §8(f)
// rank 4, "callers either side of the path" — and speaks to ANY TracerI (the tests drive the unmodified reference tracer with it).
//
// Like the reference's loader it is ATTRIBUTE DRIVEN: for every group type it asks the tracer for AttributeInfo(group) and
// reads the JSON key of each attribute's name with the attribute's data type, texturability and optionality
// (GenericAttributeLoad / TexturableAttributeLoad, SceneLoaderMRay.cpp:L117-330), so a material or light type the tracer
// advertises loads without this file knowing it. Special cases are the reference's own: Pinhole cameras
// (fov / aspect -> FovAndPlanes, L595-668), Single transforms (matrix / trs layouts, L470-593), triangle primitives given in
// the node (nodeTriangle / nodeTriangleIndexed tags; normals become tangent-space quaternions, L395-436), Primitive lights.
//
// Supported: Cameras (Pinhole), Lights (Null, Primitive, Skysphere_Spherical, Skysphere_CoOcta), Mediums (Vacuum),
// Transforms (Identity, Single), Materials (whatever the tracer advertises with scalar / vector constant or texturable
// attributes), Primitives (Triangle, in-node), Textures (PFM files: "PF" colour, "Pf" single channel — the one image format
// that needs no library), Boundary, Surfaces (alphaMap, cullBackFace), LightSurfaces, CameraSurfaces; arrayed nodes ("id": [...]).
// Not supported (throws): mesh files (assimp), other image formats (OpenImageIO), spheres, media, Multi transforms, volumes
// other than the boundary's vacuum.
#include "Core/SceneLoaderI.h"
#include "Core/TracerI.h"
#include "Core/Error.h"
#include "Core/Timer.h"
#include "Core/GraphicsFunctions.h"
#include "Core/TypeNameGenerators.h"
#include "TransientPool/TransientPool.h"

#include "jsonc.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>

class ThreadPool;

namespace
{
using jsonc::Value;
using namespace std::string_literals;
using namespace std::string_view_literals;

template<class Id> uint32_t Raw(Id id) { return static_cast<uint32_t>(id); }

// One logical item of an (optionally arrayed) JSON node: {"id": [1, 2, 3], "albedo": [[..], [..], [..]]} is three items
struct Item
{
    const Value* node = nullptr;
    uint32_t inner = 0; bool multi = false;
    uint32_t id = 0;
    std::string_view Type() const { return node->At("type").AsString(); }
    // the item's share of a field: field[inner] for arrayed nodes
    const Value* Field(std::string_view key) const
    {
        const Value* v = node->Find(key);
        if(!v) return nullptr;
        if(!multi) return v;
        if(!v->IsArray() || inner >= v->arr.size()) throw MRayError("Arrayed node field \"{}\" does not match its id list", key);
        return &v->arr[inner];
    }
};

std::vector<Item> Expand(const Value& list, std::string_view listName)
{
    std::vector<Item> out;
    if(!list.IsArray()) throw MRayError("\"{}\" must be an array", listName);
    for(const Value& n : list.arr)
    {
        const Value& id = n.At("id");
        if(id.IsArray())
            for(uint32_t i = 0; i < id.arr.size(); i++) out.push_back(Item{&n, i, true, id.arr[i].AsU32()});
        else out.push_back(Item{&n, 0, false, id.AsU32()});
    }
    std::sort(out.begin(), out.end(), [](const Item& a, const Item& b) { return a.id < b.id; });
    for(size_t i = 1; i < out.size(); i++)
        if(out[i].id == out[i - 1].id) throw MRayError("Duplicate id {} in \"{}\"", out[i].id, listName);
    return out;
}

template<unsigned N> Vector<N, Float> ToVec(const Value& v)
{
    if(!v.IsArray() || v.arr.size() != N) throw MRayError("a vector of {} numbers expected", N);
    Vector<N, Float> r;
    for(unsigned i = 0; i < N; i++) r[i] = Float(v.arr[i].AsNumber());
    return r;
}

// Reads one JSON value as the runtime data type `dt` and appends it to `out`
void PushTyped(TransientData& out, const MRayDataTypeRT& dt, const Value& v, std::string_view name)
{
    using enum MRayDataEnum;
    switch(dt.Name())
    {
        case MR_FLOAT:    { Float f = Float(v.AsNumber()); out.Push(Span<const Float>(&f, 1)); break; }
        case MR_VECTOR_2: { Vector2 x = ToVec<2>(v); out.Push(Span<const Vector2>(&x, 1)); break; }
        case MR_VECTOR_3: { Vector3 x = ToVec<3>(v); out.Push(Span<const Vector3>(&x, 1)); break; }
        case MR_VECTOR_4: { Vector4 x = ToVec<4>(v); out.Push(Span<const Vector4>(&x, 1)); break; }
        case MR_BOOL:     { bool b = v.AsBool(); out.Push(Span<const bool>(&b, 1)); break; }
        case MR_UINT32:   { uint32_t u = v.AsU32(); out.Push(Span<const uint32_t>(&u, 1)); break; }
        default: throw MRayError("attribute \"{}\": data type is not supported by this loader", name);
    }
}
void PushZero(TransientData& out, const MRayDataTypeRT& dt, std::string_view name)
{
    using enum MRayDataEnum;
    switch(dt.Name())
    {
        case MR_FLOAT:    { Float f = 0; out.Push(Span<const Float>(&f, 1)); break; }
        case MR_VECTOR_2: { Vector2 x = Vector2::Zero(); out.Push(Span<const Vector2>(&x, 1)); break; }
        case MR_VECTOR_3: { Vector3 x = Vector3::Zero(); out.Push(Span<const Vector3>(&x, 1)); break; }
        case MR_VECTOR_4: { Vector4 x = Vector4::Zero(); out.Push(Span<const Vector4>(&x, 1)); break; }
        case MR_BOOL:     { bool b = false; out.Push(Span<const bool>(&b, 1)); break; }
        case MR_UINT32:   { uint32_t u = 0; out.Push(Span<const uint32_t>(&u, 1)); break; }
        default: throw MRayError("attribute \"{}\": data type is not supported by this loader", name);
    }
}

// {"texture": id [, "channels": ...]} -> scene texture id
bool IsTextureRef(const Value& v) { return v.IsObject() && v.Find("texture"); }

// PFM: "PF" (3 channels) or "Pf" (1), width height, scale (negative = little endian), rows bottom to top
struct Image { uint32_t w = 0, h = 0, channels = 0; std::vector<float> pix; };
Image ReadPFM(const std::filesystem::path& path)
{
    std::ifstream f(path, std::ios::binary);
    if(!f) throw MRayError("Unable to open texture file \"{}\"", path.string());
    std::string magic; int w = 0, h = 0; double scale = 0;
    f >> magic >> w >> h >> scale;
    f.get();   // the single whitespace after the header
    if((magic != "PF" && magic != "Pf") || w <= 0 || h <= 0 || scale == 0.0) throw MRayError("\"{}\" is not a PFM image", path.string());
    Image img; img.w = uint32_t(w); img.h = uint32_t(h); img.channels = magic == "PF" ? 3u : 1u;
    img.pix.resize(size_t(w) * h * img.channels);
    f.read(reinterpret_cast<char*>(img.pix.data()), std::streamsize(img.pix.size() * 4));
    if(!f) throw MRayError("\"{}\": truncated PFM image", path.string());
    if(scale > 0.0)   // big endian file
        for(float& v : img.pix) { uint32_t u; std::memcpy(&u, &v, 4); u = __builtin_bswap32(u); std::memcpy(&v, &u, 4); }
    return img;   // row 0 = bottom row = v 0: the orientation textures are addressed in
}

class SceneLoaderB200 final : public SceneLoaderI
{
    std::string scenePath;

    struct TypeNames
    {
        static std::string Prim(std::string_view t) { return "(P)"s + std::string(t); }
        static std::string Mat(std::string_view t) { return "(Mt)"s + std::string(t); }
        static std::string Trans(std::string_view t) { return "(T)"s + std::string(t); }
        static std::string Cam(std::string_view t) { return "(C)"s + std::string(t); }
        static std::string Light(std::string_view t) { return "(L)"s + std::string(t); }
    };

    // ---- textures (SceneLoaderMRay.cpp:L929-1150): CreateTexture2D for all -> CommitTextures -> PushTextureData ----
    void LoadTextures(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        const Value* list = root.Find("Textures");
        if(!list || list->arr.empty()) return;
        std::vector<Item> items = Expand(*list, "Textures");
        std::vector<Image> images;
        std::vector<TextureId> ids;
        for(const Item& it : items)
        {
            const Value* file = it.Field("file");
            if(!file) throw MRayError("Texture({}) has no \"file\"", it.id);
            std::filesystem::path p = file->AsString();
            if(p.is_relative()) p = std::filesystem::path(scenePath).parent_path() / p;
            if(p.extension() != ".pfm" && p.extension() != ".PFM")
                throw MRayError("Texture({}): only PFM images are readable without an image library (\"{}\")", it.id, p.string());
            images.push_back(ReadPFM(p));
            const Image& img = images.back();
            MRayTextureParameters tp;
            const Value* isColor = it.Field("isColor");
            const bool color = isColor ? isColor->AsBool() : img.channels == 3u;
            // the tracer stores 3-channel images as 4-channel pixels (there is no 3-channel texture format) and reads them
            // back as Vector3 (MR_DROP_1); single-channel images stay single channel
            tp.pixelType = MRayPixelTypeRT(img.channels == 3u ? MRayPixelEnum::MR_RGBA_FLOAT : MRayPixelEnum::MR_R_FLOAT);
            tp.isColor = color ? AttributeIsColor::IS_COLOR : AttributeIsColor::IS_PURE_DATA;
            tp.readMode = img.channels == 3u ? MRayTextureReadMode::MR_DROP_1 : MRayTextureReadMode::MR_PASSTHROUGH;
            if(const Value* v = it.Field("ignoreResClamp")) tp.ignoreResClamp = v->AsBool();
            if(const Value* v = it.Field("gamma")) tp.gamma = Float(v->AsNumber());
            if(const Value* v = it.Field("edgeResolve"))
            {
                tp.edgeResolve = MRayTextureEdgeResolveStringifier::FromString(v->AsString());
                if(tp.edgeResolve == MRayTextureEdgeResolveEnum::MR_ENUM_END) throw MRayError("Unknown edge resolve \"{}\"", v->AsString());
            }
            if(const Value* v = it.Field("interpolation"))
            {
                tp.interpolation = MRayTextureInterpStringifier::FromString(v->AsString());
                if(tp.interpolation == MRayTextureInterpEnum::MR_ENUM_END) throw MRayError("Unknown texture interp \"{}\"", v->AsString());
            }
            if(const Value* v = it.Field("colorSpace"))
            {
                tp.colorSpace = MRayColorSpaceStringifier::FromString(v->AsString());
                if(tp.colorSpace == MRayColorSpaceEnum::MR_ENUM_END) throw MRayError("Unknown color space \"{}\"", v->AsString());
            }
            if(const Value* v = it.Field("isIlluminant"))
                tp.isIlluminant = v->AsBool() ? MRayTextureIsIlluminant::IS_ILLUMINANT : MRayTextureIsIlluminant::IS_ALBEDO;
            ids.push_back(tracer.CreateTexture2D(Vector2ui(img.w, img.h), 1, tp));
            pack.textures.emplace(SceneTexId(it.id), ids.back());
        }
        tracer.CommitTextures();
        for(size_t k = 0; k < images.size(); k++)
        {
            const Image& img = images[k];
            const size_t n = size_t(img.w) * img.h;
            if(img.channels == 3u)
            {
                TransientData d(std::in_place_type_t<Vector4>{}, n);
                std::vector<Vector4> px(n);
                for(size_t i = 0; i < n; i++) px[i] = Vector4(img.pix[3 * i], img.pix[3 * i + 1], img.pix[3 * i + 2], Float(1));
                d.Push(Span<const Vector4>(px));
                tracer.PushTextureData(ids[k], 0, std::move(d));
            }
            else
            {
                TransientData d(std::in_place_type_t<Float>{}, n);
                d.Push(Span<const Float>(img.pix.data(), n));
                tracer.PushTextureData(ids[k], 0, std::move(d));
            }
        }
    }

    // ---- attribute-driven group loading (materials and lights share it) ----
    // Returns per attribute (data, textures); textures is empty for constant-only attributes.
    struct AttrData { TransientData data; std::vector<Optional<TextureId>> textures; bool texturable = false, present = false; };
    template<class InfoList>
    std::vector<AttrData> LoadAttributes(const InfoList& infos, Span<const Item> items, const TracerIdPack& pack,
                                         std::vector<AttributeCountList>& counts)
    {
        std::vector<AttrData> out;
        counts.assign(items.size(), AttributeCountList(StaticVecSize(infos.size())));
        for(size_t a = 0; a < infos.size(); a++)
        {
            const auto& info = infos[a];
            const std::string_view name = info.name;
            if(info.isArray == AttributeIsArray::IS_ARRAY) throw MRayError("attribute \"{}\": array attributes are not supported by this loader", name);
            const bool texOnly = info.isTexturable == AttributeTexturable::MR_TEXTURE_ONLY;
            const bool texOrConst = info.isTexturable == AttributeTexturable::MR_TEXTURE_OR_CONSTANT;
            const bool optional = info.isOptional == AttributeOptionality::MR_OPTIONAL;
            size_t dataCount = 0;
            for(size_t k = 0; k < items.size(); k++)
            {
                const bool has = items[k].Field(name) != nullptr;
                if(!has && !optional) throw MRayError("{}({}) lacks the mandatory attribute \"{}\"", items[k].Type(), items[k].id, name);
                counts[k][a] = (has || !optional) ? 1 : 0;
                if(!texOnly && has) dataCount++;
            }
            AttrData ad{AllocateTransientData(info.dataType, texOnly ? 0 : (optional ? dataCount : items.size())), {}, texOnly || texOrConst, false};
            for(const Item& it : items)
            {
                const Value* v = it.Field(name);
                if(texOnly)
                {
                    Optional<TextureId> t;
                    if(v) { if(!IsTextureRef(*v)) throw MRayError("attribute \"{}\" must be {{\"texture\": id}}", name); t = pack.textures.at(SceneTexId(v->At("texture").AsU32())); }
                    ad.textures.push_back(t);
                }
                else if(texOrConst)
                {
                    if(v && IsTextureRef(*v))
                    {
                        ad.textures.emplace_back(pack.textures.at(SceneTexId(v->At("texture").AsU32())));
                        PushZero(ad.data, info.dataType, name);   // a phony constant rides along, as in the reference (L296-301)
                    }
                    else { ad.textures.emplace_back(std::nullopt); PushTyped(ad.data, info.dataType, *v, name); }
                }
                else if(v) PushTyped(ad.data, info.dataType, *v, name);
                ad.present = ad.present || v != nullptr;
            }
            out.push_back(std::move(ad));
        }
        return out;
    }

    void LoadMaterials(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        const Value* list = root.Find("Materials");
        if(!list) return;
        std::vector<Item> items = Expand(*list, "Materials");
        std::map<std::string, std::vector<Item>, std::less<>> byType;
        for(const Item& it : items) byType[std::string(it.Type())].push_back(it);
        for(auto& [type, group] : byType)
        {
            MatGroupId g = tracer.CreateMaterialGroup(TypeNames::Mat(type));
            MatAttributeInfoList infos = tracer.AttributeInfo(g);
            std::vector<AttributeCountList> counts;
            std::vector<AttrData> attrs = LoadAttributes(infos, Span<const Item>(group), pack, counts);
            MaterialIdList ids = tracer.ReserveMaterials(g, counts);
            tracer.CommitMatReservations(g);
            const auto range = CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back()));
            for(uint32_t a = 0; a < attrs.size(); a++)
            {
                if(attrs[a].texturable) tracer.PushMatAttribute(g, range, a, std::move(attrs[a].data), std::move(attrs[a].textures));
                else if(attrs[a].present || infos[a].isOptional == AttributeOptionality::MR_MANDATORY)
                    tracer.PushMatAttribute(g, range, a, std::move(attrs[a].data));
            }
            for(size_t k = 0; k < group.size(); k++) pack.mats.emplace(group[k].id, Pair<MatGroupId, MaterialId>(g, ids[k]));
        }
    }

    // ---- transforms (SceneLoaderMRay.cpp:L470-593) ----
    static Matrix3x4 TransformOf(const Item& it)
    {
        const Value* layout = it.node->Find("layout");
        if(!layout) throw MRayError("Transform({}) has no \"layout\"", it.id);
        if(layout->AsString() == "matrix")
        {
            const Value* m = it.Field("matrix");
            if(!m || !m->IsArray() || m->arr.size() != 16) throw MRayError("Transform({}): \"matrix\" must hold 16 numbers", it.id);
            Float v[16];
            for(int i = 0; i < 16; i++) v[i] = Float(m->arr[size_t(i)].AsNumber());
            return Matrix3x4(Vector4(v[0], v[1], v[2], v[3]), Vector4(v[4], v[5], v[6], v[7]), Vector4(v[8], v[9], v[10], v[11]));
        }
        if(layout->AsString() == "trs")
        {
            const Value* tv = it.Field("translate"); const Value* rv = it.Field("rotate"); const Value* sv = it.Field("scale");
            const Vector3 t = tv ? ToVec<3>(*tv) : Vector3::Zero(), r = rv ? ToVec<3>(*rv) : Vector3::Zero(), s = sv ? ToVec<3>(*sv) : Vector3(1);
            const Vector3 rRadians = r * MathConstants::DegToRadCoef<Float>();
            Matrix4x4 transform = TransformGen::Scale(s[0], s[1], s[2]);
            transform = TransformGen::Rotate(rRadians[0], Vector3::XAxis()) * transform;
            transform = TransformGen::Rotate(rRadians[1], Vector3::YAxis()) * transform;
            transform = TransformGen::Rotate(rRadians[2], Vector3::ZAxis()) * transform;
            transform = TransformGen::Translate(t) * transform;
            return Matrix3x4(transform);
        }
        throw MRayError("Unkown transform layout");
    }
    void LoadTransforms(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        const Value* list = root.Find("Transforms");
        if(!list) return;
        std::vector<Item> singles;
        for(const Item& it : Expand(*list, "Transforms"))
        {
            if(it.Type() == "Identity")
                pack.transforms.emplace(it.id, Pair<TransGroupId, TransformId>(TracerConstants::IdentityTransGroupId, TracerConstants::IdentityTransformId));
            else if(it.Type() == "Single") singles.push_back(it);
            else throw MRayError("Transform type \"{}\" is not supported by this loader", it.Type());
        }
        if(singles.empty()) return;
        TransGroupId g = tracer.CreateTransformGroup(TypeNames::Trans("Single"));
        std::vector<AttributeCountList> counts(singles.size());
        for(auto& c : counts) { c = AttributeCountList(StaticVecSize(1)); c[0] = 1; }
        TransformIdList ids = tracer.ReserveTransformations(g, counts);
        tracer.CommitTransReservations(g);
        std::vector<Matrix3x4> ms;
        for(const Item& it : singles) ms.push_back(TransformOf(it));
        TransientData d(std::in_place_type_t<Matrix3x4>{}, ms.size());
        d.Push(Span<const Matrix3x4>(ms));
        tracer.PushTransAttribute(g, CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back())), 0, std::move(d));
        for(size_t k = 0; k < singles.size(); k++) pack.transforms.emplace(singles[k].id, Pair<TransGroupId, TransformId>(g, ids[k]));
    }

    // ---- primitives given in the node (MeshLoaderJson.cpp; SceneLoaderMRay.cpp:L335-460) ----
    void LoadPrimitives(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        const Value* list = root.Find("Primitives");
        if(!list) return;
        std::vector<Item> items = Expand(*list, "Primitives");
        if(items.empty()) return;
        for(const Item& it : items)
        {
            if(it.multi) throw MRayError("Primitive({}): arrayed in-node primitives are not supported by this loader", it.id);
            if(it.Type() != "Triangle") throw MRayError("Primitive type \"{}\" is not supported by this loader", it.Type());
            const Value* tag = it.node->Find("tag");
            if(!tag || (tag->AsString() != "nodeTriangle" && tag->AsString() != "nodeTriangleIndexed"))
                throw MRayError("Primitive({}): only in-node triangles (\"nodeTriangle\" / \"nodeTriangleIndexed\") are readable without a mesh library", it.id);
        }
        PrimGroupId pg = tracer.CreatePrimitiveGroup(TypeNames::Prim("Triangle"));
        std::vector<PrimCount> counts;
        for(const Item& it : items)
        {
            const uint32_t vN = uint32_t(it.node->At("position").arr.size());
            const bool indexed = it.node->At("tag").AsString() == "nodeTriangleIndexed";
            const uint32_t tN = indexed ? uint32_t(it.node->At("index").arr.size()) : vN / 3u;
            if(!indexed && vN % 3u != 0) throw MRayError("Primitive({}): \"nodeTriangle\" needs a multiple of three vertices", it.id);
            counts.push_back(PrimCount{tN, vN});
        }
        PrimBatchIdList batches = tracer.ReservePrimitiveBatches(pg, counts);
        tracer.CommitPrimReservations(pg);
        PrimAttributeInfoList infos = tracer.AttributeInfo(pg);
        for(size_t k = 0; k < items.size(); k++)
        {
            const Value& n = *items[k].node;
            const uint32_t vN = counts[k].attributeCount, tN = counts[k].primCount;
            const bool indexed = n.At("tag").AsString() == "nodeTriangleIndexed";
            std::vector<Vector3> pos(vN), nrm(vN, Vector3(0, 0, 1));
            for(uint32_t i = 0; i < vN; i++) pos[i] = ToVec<3>(n.At("position")[i]);
            std::vector<Vector3ui> idx(tN);
            for(uint32_t t = 0; t < tN; t++)
            {
                if(indexed) { const Value& iv = n.At("index")[t]; idx[t] = Vector3ui(iv[0].AsU32(), iv[1].AsU32(), iv[2].AsU32()); }
                else idx[t] = Vector3ui(3 * t, 3 * t + 1, 3 * t + 2);
            }
            if(const Value* nv = n.Find("normal")) { for(uint32_t i = 0; i < vN; i++) nrm[i] = ToVec<3>((*nv)[i]); }
            else
            {   // no normals given: area-weighted face normals
                std::fill(nrm.begin(), nrm.end(), Vector3::Zero());
                for(const Vector3ui& t : idx)
                {
                    const Vector3 fn = Math::Cross(pos[t[1]] - pos[t[0]], pos[t[2]] - pos[t[0]]);
                    for(int c = 0; c < 3; c++) nrm[t[c]] = nrm[t[c]] + fn;
                }
            }
            for(uint32_t a = 0; a < infos.size(); a++)
            {
                using enum PrimitiveAttributeLogic::E;
                switch(infos[a].logic.e)
                {
                    case POSITION:
                    {
                        TransientData d(std::in_place_type_t<Vector3>{}, vN); d.Push(Span<const Vector3>(pos));
                        tracer.PushPrimAttribute(pg, batches[k], a, std::move(d)); break;
                    }
                    case NORMAL:
                    {   // normals travel as world -> tangent-space rotations (SceneLoaderMRay.cpp:L395-436)
                        TransientData d(std::in_place_type_t<Quaternion>{}, vN);
                        for(uint32_t i = 0; i < vN; i++)
                        {
                            const Vector3 nn = Math::Normalize(nrm[i]);
                            const Vector3 bt = Graphics::OrthogonalVector(nn);
                            const Vector3 tg = Math::Cross(bt, nn);
                            const Quaternion q = TransformGen::ToSpaceQuat(tg, bt, nn);
                            d.Push(Span<const Quaternion>(&q, 1));
                        }
                        tracer.PushPrimAttribute(pg, batches[k], a, std::move(d)); break;
                    }
                    case UV0:
                    {
                        std::vector<Vector2> uv(vN, Vector2::Zero());
                        if(const Value* uvv = n.Find("uv")) for(uint32_t i = 0; i < vN; i++) uv[i] = ToVec<2>((*uvv)[i]);
                        TransientData d(std::in_place_type_t<Vector2>{}, vN); d.Push(Span<const Vector2>(uv));
                        tracer.PushPrimAttribute(pg, batches[k], a, std::move(d)); break;
                    }
                    case INDEX:
                    {
                        TransientData d(std::in_place_type_t<Vector3ui>{}, tN); d.Push(Span<const Vector3ui>(idx));
                        tracer.PushPrimAttribute(pg, batches[k], a, std::move(d)); break;
                    }
                    default: break;
                }
            }
            pack.prims.emplace(items[k].id, Pair<PrimGroupId, PrimBatchId>(pg, batches[k]));
        }
    }

    // ---- cameras (SceneLoaderMRay.cpp:L595-668) ----
    void LoadCameras(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        std::vector<Item> items = Expand(root.At("Cameras"), "Cameras");
        if(items.empty()) return;
        for(const Item& it : items) if(it.Type() != "Pinhole") throw MRayError("Camera type \"{}\" is not supported by this loader", it.Type());
        CameraGroupId g = tracer.CreateCameraGroup(TypeNames::Cam("Pinhole"));
        CamAttributeInfoList infos = tracer.AttributeInfo(g);
        std::vector<AttributeCountList> counts(items.size());
        for(auto& c : counts) { c = AttributeCountList(StaticVecSize(infos.size())); for(size_t k = 0; k < infos.size(); k++) c[k] = 1; }
        CameraIdList ids = tracer.ReserveCameras(g, counts);
        tracer.CommitCamReservations(g);
        std::vector<Vector4> fnp; std::vector<Vector3> gaze, pos, up;
        for(const Item& it : items)
        {
            const bool isFovX = it.Field("isFovX")->AsBool();
            const Float fov = Float(it.Field("fov")->AsNumber()), aspect = Float(it.Field("aspect")->AsNumber());
            const Vector2 planes = ToVec<2>(*it.Field("planes"));
            Vector4 f(fov, fov, planes[0], planes[1]);
            if(isFovX) { f[0] *= MathConstants::DegToRadCoef<Float>(); f[1] = Float(2.0) * std::atan(std::tan(f[0] * Float(0.5)) / aspect); }
            else { f[1] *= MathConstants::DegToRadCoef<Float>(); f[0] = Float(2.0) * std::atan(std::tan(f[1] * Float(0.5)) * aspect); }
            fnp.push_back(f);
            gaze.push_back(ToVec<3>(*it.Field("gaze"))); pos.push_back(ToVec<3>(*it.Field("position"))); up.push_back(ToVec<3>(*it.Field("up")));
        }
        const auto range = CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back()));
        // attribute order of CameraGroupPinhole: FovAndPlanes, gaze, position, up
        { TransientData d(std::in_place_type_t<Vector4>{}, fnp.size()); d.Push(Span<const Vector4>(fnp)); tracer.PushCamAttribute(g, range, 0, std::move(d)); }
        const std::vector<Vector3>* v3[3] = {&gaze, &pos, &up};
        for(uint32_t a = 1; a < 4; a++)
        { TransientData d(std::in_place_type_t<Vector3>{}, v3[a - 1]->size()); d.Push(Span<const Vector3>(*v3[a - 1])); tracer.PushCamAttribute(g, range, a, std::move(d)); }
        for(size_t k = 0; k < items.size(); k++) pack.cams.emplace(items[k].id, Pair<CameraGroupId, CameraId>(g, ids[k]));
    }

    // ---- lights (SceneLoaderMRay.cpp:L1577-1720): Null, Primitive (typed by its primitive), anything else by attribute info ----
    void LoadLights(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        std::vector<Item> items = Expand(root.At("Lights"), "Lights");
        std::map<std::string, std::vector<Item>, std::less<>> byType;
        for(const Item& it : items)
        {
            if(it.Type() == "Null") { pack.lights.emplace(it.id, Pair<LightGroupId, LightId>(TracerConstants::NullLightGroupId, TracerConstants::NullLightId)); continue; }
            byType[std::string(it.Type())].push_back(it);
        }
        for(auto& [type, group] : byType)
        {
            const bool primBacked = type == "Primitive";
            LightGroupId g;
            std::vector<PrimBatchId> batches;
            if(primBacked)
            {
                PrimGroupId pgAll = PrimGroupId(0); bool first = true;
                for(const Item& it : group)
                {
                    const Value* p = it.Field("primitive");
                    if(!p) throw MRayError("Light({}) of type Primitive has no \"primitive\"", it.id);
                    const auto& pr = pack.prims.at(p->AsU32());
                    if(!first && Raw(pr.first) != Raw(pgAll)) throw MRayError("Primitive lights over several primitive groups are not supported by this loader");
                    pgAll = pr.first; first = false;
                    batches.push_back(pr.second);
                }
                // PrimLightTypeName: "(L)Prim" + the primitive group's type name (Core/TypeNameGenerators.h)
                g = tracer.CreateLightGroup("(L)Prim"s + tracer.TypeName(pgAll), pgAll);
            }
            else g = tracer.CreateLightGroup(TypeNames::Light(type));
            LightAttributeInfoList infos = tracer.AttributeInfo(g);
            // "isTwoSided" is mandatory in the tracer but commonly left out of scene files: default false
            std::vector<Item> patched = group;
            std::vector<AttributeCountList> counts;
            std::vector<AttrData> attrs;
            {
                LightAttributeInfoList relaxed = infos;
                for(auto& i : relaxed) if(i.name == "isTwoSided"sv) i.isOptional = AttributeOptionality::MR_OPTIONAL;
                attrs = LoadAttributes(relaxed, Span<const Item>(patched), pack, counts);
                for(size_t a = 0; a < infos.size(); a++)
                    if(infos[a].name == "isTwoSided"sv)
                    {
                        TransientData d(std::in_place_type_t<bool>{}, group.size());
                        for(const Item& it : group) { const Value* v = it.Field("isTwoSided"); bool b = v ? v->AsBool() : false; d.Push(Span<const bool>(&b, 1)); }
                        attrs[a].data = std::move(d); attrs[a].present = true;
                        for(auto& c : counts) c[a] = 1;
                    }
            }
            LightIdList ids = primBacked ? tracer.ReserveLights(g, counts, batches) : tracer.ReserveLights(g, counts);
            tracer.CommitLightReservations(g);
            const auto range = CommonIdRange(std::bit_cast<CommonId>(ids.front()), std::bit_cast<CommonId>(ids.back()));
            for(uint32_t a = 0; a < attrs.size(); a++)
            {
                if(attrs[a].texturable) tracer.PushLightAttribute(g, range, a, std::move(attrs[a].data), std::move(attrs[a].textures));
                else tracer.PushLightAttribute(g, range, a, std::move(attrs[a].data));
            }
            for(size_t k = 0; k < group.size(); k++) pack.lights.emplace(group[k].id, Pair<LightGroupId, LightId>(g, ids[k]));
        }
    }

    void LoadMediums(const Value& root, TracerIdPack& pack)
    {
        const Value* list = root.Find("Mediums");
        if(!list) return;
        for(const Item& it : Expand(*list, "Mediums"))
        {
            if(it.Type() != "Vacuum") throw MRayError("Medium type \"{}\" is not supported by this loader", it.Type());
            pack.mediums.emplace(it.id, Pair<MediumGroupId, MediumId>(TracerConstants::VacuumMediumGroupId, TracerConstants::VacuumMediumId));
        }
    }

    TransformId TransformOfSurface(const Value& n, const TracerIdPack& pack) const
    {
        const Value* t = n.Find("transform");
        if(!t) return TracerConstants::IdentityTransformId;
        auto it = pack.transforms.find(t->AsU32());
        if(it == pack.transforms.end()) throw MRayError("Transform({}) is not defined", t->AsU32());
        return it->second.second;
    }

    // ---- surfaces (JsonNode.hpp:L18-122, SceneLoaderMRay.cpp:L2136-2260) ----
    void LoadSurfaces(TracerI& tracer, const Value& root, TracerIdPack& pack)
    {
        uint32_t sIndex = 0;
        for(const Value& n : root.At("Surfaces").arr)
        {
            const Value& mat = n.At("material"); const Value& prim = n.At("primitive");
            if(mat.Size() != prim.Size() || mat.IsArray() != prim.IsArray()) throw MRayError("Material/Primitive pair lists does not match on a surface!");
            const size_t pairs = mat.IsArray() ? mat.arr.size() : 1;
            if(pairs > TracerConstants::MaxPrimBatchPerSurface) throw MRayError("A surface holds at most {} material / primitive pairs", TracerConstants::MaxPrimBatchPerSurface);
            SurfaceParams sp;
            sp.transformId = TransformOfSurface(n, pack);
            auto At = [&](const Value& v, size_t i) -> const Value& { return v.IsArray() ? v.arr[i] : v; };
            for(size_t i = 0; i < pairs; i++)
            {
                auto pIt = pack.prims.find(At(prim, i).AsU32()); auto mIt = pack.mats.find(At(mat, i).AsU32());
                if(pIt == pack.prims.end()) throw MRayError("Primitive({}) is not defined", At(prim, i).AsU32());
                if(mIt == pack.mats.end()) throw MRayError("Material({}) is not defined", At(mat, i).AsU32());
                sp.primBatches.push_back(pIt->second.second);
                sp.materials.push_back(mIt->second.second);
                // defaults of the format: back faces culled, no alpha map (JsonNode.hpp:L31-34)
                bool cull = true;
                if(const Value* c = n.Find("cullBackFace")) cull = At(*c, i).AsBool();
                sp.cullFaceFlags.push_back(cull);
                Optional<TextureId> alpha;
                if(const Value* a = n.Find("alphaMap"))
                {
                    const Value& av = At(*a, i);
                    if(IsTextureRef(av)) alpha = pack.textures.at(SceneTexId(av.At("texture").AsU32()));
                }
                sp.alphaMaps.push_back(alpha);
                sp.volumes.push_back(TracerConstants::InvalidVolume);
            }
            pack.surfaces.emplace_back(sIndex++, tracer.CreateSurface(sp));
        }
        uint32_t lIndex = 0;
        if(const Value* ls = root.Find("LightSurfaces"))
            for(const Value& n : ls->arr)
            {
                auto lIt = pack.lights.find(n.At("light").AsU32());
                if(lIt == pack.lights.end()) throw MRayError("Light({}) is not defined", n.At("light").AsU32());
                pack.lightSurfaces.emplace_back(lIndex++, tracer.CreateLightSurface(LightSurfaceParams{lIt->second.second, TransformOfSurface(n, pack), {}}));
            }
        uint32_t cIndex = 0;
        for(const Value& n : root.At("CameraSurfaces").arr)
        {
            auto cIt = pack.cams.find(n.At("camera").AsU32());
            if(cIt == pack.cams.end()) throw MRayError("Camera({}) is not defined", n.At("camera").AsU32());
            pack.camSurfaces.emplace_back(cIndex++, tracer.CreateCameraSurface(CameraSurfaceParams{cIt->second.second, TransformOfSurface(n, pack), {}}));
        }
        // boundary: light + transform (+ the vacuum it sits in)
        const Value& b = root.At("Boundary");
        const Value* bl = b.Find("light");
        if(!bl) throw MRayError("Boundary light must be set!");
        auto lIt = pack.lights.find(bl->AsU32());
        if(lIt == pack.lights.end()) throw MRayError("Light({}) is not defined", bl->AsU32());
        pack.boundarySurface = Pair<uint32_t, LightSurfaceId>(0u, tracer.SetBoundarySurface(lIt->second.second, TransformOfSurface(b, pack)));
        pack.boundaryVolume = tracer.RegisterVolume(VolumeParams{TracerConstants::VacuumMediumId, TracerConstants::IdentityTransformId, 0});
        tracer.SetBoundaryVolume(pack.boundaryVolume);
    }

    Expected<TracerIdPack> Load(TracerI& tracer, std::string_view text)
    {
        Timer t; t.Start();
        try
        {
            const Value root = jsonc::Parse(text);
            if(!root.IsObject()) throw MRayError("Scene file must hold one JSON object");
            for(std::string_view key : {"Cameras"sv, "Lights"sv, "Boundary"sv, "Surfaces"sv, "CameraSurfaces"sv})
                if(!root.Find(key)) throw MRayError("Scene file does not contain \"{}\"", key);
            TracerIdPack pack;
            // dependency order: textures <- materials / lights; primitives <- lights; everything <- surfaces
            LoadTextures(tracer, root, pack);
            LoadMediums(root, pack);
            LoadTransforms(tracer, root, pack);
            LoadPrimitives(tracer, root, pack);
            LoadMaterials(tracer, root, pack);
            LoadCameras(tracer, root, pack);
            LoadLights(tracer, root, pack);
            LoadSurfaces(tracer, root, pack);
            t.Split();
            pack.loadTimeMS = t.Elapsed<Millisecond>();
            return pack;
        }
        catch(const MRayError& e) { return e; }
        catch(const std::exception& e) { return MRayError("{}", e.what()); }
    }

    public:
    Expected<TracerIdPack> LoadScene(TracerI& tracer, const std::string& filePath) override
    {
        std::ifstream f(filePath, std::ios::binary);
        if(!f) return MRayError("Scene file \"{}\" not found", filePath);
        std::stringstream ss; ss << f.rdbuf();
        scenePath = filePath;
        return Load(tracer, ss.str());
    }
    Expected<TracerIdPack> LoadScene(TracerI& tracer, std::istream& sceneData) override
    {
        std::stringstream ss; ss << sceneData.rdbuf();
        scenePath.clear();
        return Load(tracer, ss.str());
    }
    void ClearScene() override { scenePath.clear(); }
};

} // namespace

// the reference's entry points (SceneLoaderMRay/EntryPoint.h): MRay's TracerThread looks these names up in the library that
// its configuration lists for the ".json" extension (MRay/TracerThread.cpp:L783-810)
extern "C" __attribute__((visibility("default"))) SceneLoaderI* ConstructSceneLoaderMRay(ThreadPool&) { return new SceneLoaderB200(); }
extern "C" __attribute__((visibility("default"))) void DestroySceneLoaderMRay(SceneLoaderI* p) { delete p; }
