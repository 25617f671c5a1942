// Reference taps — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// This TU is compiled against the UNMODIFIED reference headers (-I/root/reference/Source) and
// linked with oracle/_ref/libTracerDLL_CPU.so. It contains no algorithm of its own: every
// function below forwards plain C arrays to the reference's own kernels / device functions
// running on the reference's CPU device backend, so that tests/golden fixtures and the C
// restatement in oracle/ can be pinned against what the reference really computes.
//
//   morton      -> KCGenMortonCode                (Tracer/AcceleratorLBVH.cu:L96-168)
//   sort        -> DeviceAlgorithms::SegmentedRadixSort<true,u64,u32> (AcceleratorLBVH.hpp:L828-838)
//   hierarchy   -> KCConstructLBVHInternalNodes   (AcceleratorLBVH.cu:L170-299)
//   boxes       -> KCUnionLBVHBoundingBoxes       (AcceleratorLBVH.cu:L301-435)
//   prim aabb/centroid -> DefaultTriangleDetail::Triangle::GetAABB/GetCenter
//                                                 (PrimitiveDefaultTriangle.hpp:L100-115)
//   traversal   -> LBVHAccelDetail::TraverseLBVHStack + Ray::IntersectsAABB/IntersectsTriangle
//                  (AcceleratorLBVH.hpp:L109-167, Core/Ray.hpp:L121-219); the leaf functor mirrors
//                  AcceleratorLBVH::ClosestHit / FirstHit / IntersectionCheck (hpp:L225-410)
//                  for a triangle group without alpha maps.
//   texture     -> TextureMemory (CreateTexture2D / PushTextureData per level / Finalize = ConvertColorspaces + GenerateMipmaps,
//                  Tracer/TextureMemory.cpp) and the resulting TracerTexView<2, Vector3>::operator()(uv, mipLevel | dpdx, dpdy)
//                  (Tracer/TextureView.hpp -> Device/CPU/TextureViewCPU.h)
#include "Tracer/AcceleratorLBVH.h"
#include "Tracer/AcceleratorLBVH.hpp"
#include "Tracer/PrimitiveDefaultTriangle.h"
#include "Tracer/PrimitiveDefaultTriangle.hpp"
#include "Tracer/TransformsDefault.h"
#include "Device/GPUSystem.h"
#include "Device/GPUSystem.hpp"
#include "Device/GPUAlgRadixSort.h"
#include "Device/GPUAlgGeneric.h"
#include "Core/GraphicsFunctions.h"
#include "Core/TracerI.h"
#include "Tracer/TextureMemory.h"
#include "Tracer/TextureFilter.h"
#include "Tracer/TextureView.h"
#include "Tracer/TextureView.hpp"
#include "Tracer/GenericGroup.h"
#include "Tracer/SpectrumC.h"
#include "Tracer/MaterialsDefault.h"
#include "Tracer/MaterialsDefault.hpp"

#include <memory>
#include <vector>
#include <cstring>
#include <thread>

static std::unique_ptr<GPUSystem> gSystem;

static const GPUQueue& Queue()
{
    if(!gSystem) gSystem = std::make_unique<GPUSystem>();
    return gSystem->BestDevice().GetComputeQueue(0);
}

using namespace LBVHAccelDetail;

extern "C"
{

int ref_thread_count()
{
    return int(std::thread::hardware_concurrency());
}

// Known-answer helper: Graphics::MortonCode::Compose3D (Core/GraphicsFunctions.h:L586-625)
uint64_t ref_morton_compose64(uint32_t x, uint32_t y, uint32_t z)
{
    return Graphics::MortonCode::Compose3D<uint64_t>(Vector3ui(x, y, z));
}
uint32_t ref_morton_compose32(uint32_t x, uint32_t y, uint32_t z)
{
    return Graphics::MortonCode::Compose3D<uint32_t>(Vector3ui(x, y, z));
}

// Per-triangle AABB and centroid through the reference primitive class.
void ref_tri_aabb_center(const float* positions, uint32_t nVerts,
                         const uint32_t* indices, uint32_t nTris,
                         float* outAABB /*nTris*6*/, float* outCenter /*nTris*3*/)
{
    using namespace DefaultTriangleDetail;
    TriangleData data = {};
    data.positions = Span<const Vector3>(reinterpret_cast<const Vector3*>(positions), nVerts);
    data.indexList = Span<const Vector3ui>(reinterpret_cast<const Vector3ui*>(indices), nTris);
    for(uint32_t i = 0; i < nTris; i++)
    {
        Triangle<TransformContextIdentity> tri(TransformContextIdentity{}, data,
                                               PrimitiveKey::CombinedKey(0, i));
        AABB3 aabb = tri.GetAABB();
        Vector3 c = tri.GetCenter();
        for(int k = 0; k < 3; k++)
        {
            outAABB[i * 6 + k] = aabb.Min()[k];
            outAABB[i * 6 + 3 + k] = aabb.Max()[k];
            outCenter[i * 3 + k] = c[k];
        }
    }
}

// Segmented LBVH build over `nSeg` accelerators, exactly the kernel chain of MultiBuildLBVH
// (AcceleratorLBVH.hpp:L764-895) after the per-primitive AABB/centroid kernels.
//   leafRanges : nSeg+1 prefix offsets into the leaf arrays
//   outputs    : accelAABB[nSeg*6], morton[nLeaf] (unsorted), sortedMorton[nLeaf], sortedIdx[nLeaf],
//                nodes[nNode*3] (left,right,parent), leafParent[nLeaf], nodeBoxes[nNode*6]
// Node ranges follow AcceleratorLBVH.hpp:L456-465 (max(1, leafCount-1) per accelerator).
void ref_lbvh_build(const float* leafAABB, const float* leafCenter,
                    const uint32_t* leafRanges, uint32_t nSeg,
                    float* accelAABB, uint64_t* morton,
                    uint64_t* sortedMorton, uint32_t* sortedIdx,
                    uint32_t* nodes, uint32_t* leafParent, float* nodeBoxes)
{
    static_assert(sizeof(LBVHNode) == 12 && sizeof(LBVHBoundingBox) == 24 && sizeof(AABB3) == 24);
    const GPUQueue& queue = Queue();
    static constexpr uint32_t BLOCK_PER_INSTANCE = 16;
    static constexpr uint32_t TPB = StaticThreadPerBlock1D();
    uint32_t nLeaf = leafRanges[nSeg];
    std::vector<uint32_t> nodeRanges(nSeg + 1, 0);
    for(uint32_t s = 0; s < nSeg; s++)
    {
        uint32_t lc = leafRanges[s + 1] - leafRanges[s];
        nodeRanges[s + 1] = nodeRanges[s] + Math::Max(1u, lc - 1);
    }
    uint32_t nNode = nodeRanges[nSeg];

    Span<const AABB3> dLeafAABBs(reinterpret_cast<const AABB3*>(leafAABB), nLeaf);
    Span<const Vector3> dCenters(reinterpret_cast<const Vector3*>(leafCenter), nLeaf);
    Span<const uint32_t> dLeafSeg(leafRanges, nSeg + 1);
    Span<const uint32_t> dNodeSeg(nodeRanges.data(), nSeg + 1);
    Span<AABB3> dAccelAABBs(reinterpret_cast<AABB3*>(accelAABB), nSeg);

    using namespace DeviceAlgorithms;
    // accel AABB: union of leaf AABBs (the reference uses SegmentedTransformReduce with
    // UnionAABB3Functor, hpp:L764-769; min/max are order independent so a serial fold of the same
    // functor is the same result)
    for(uint32_t s = 0; s < nSeg; s++)
    {
        AABB3 r = AABB3::Negative();
        for(uint32_t i = leafRanges[s]; i < leafRanges[s + 1]; i++)
            r = UnionAABB3Functor()(r, dLeafAABBs[i]);
        dAccelAABBs[s] = r;
    }

    std::vector<uint64_t> codes1(nLeaf);
    std::vector<uint32_t> idx0(nLeaf), idx1(nLeaf);
    std::array<Span<uint64_t>, 2> dCodes = {Span<uint64_t>(morton, nLeaf), Span<uint64_t>(codes1)};
    std::array<Span<uint32_t>, 2> dIdx = {Span<uint32_t>(idx0), Span<uint32_t>(idx1)};

    queue.IssueBlockKernel<KCGenMortonCode>
    (
        "KCGenMortonCodes",
        DeviceBlockIssueParams{.gridSize = nSeg * BLOCK_PER_INSTANCE, .blockSize = TPB},
        dCodes[0], dLeafSeg, ToConstSpan(dAccelAABBs), dCenters, BLOCK_PER_INSTANCE
    );
    queue.Barrier().Wait();
    std::vector<uint64_t> unsorted(morton, morton + nLeaf);

    size_t tmSize = SegmentedRadixSortTMSize<true, uint64_t, uint32_t>(nLeaf, nSeg, queue);
    std::vector<Byte> temp(tmSize + 16);
    SegmentedIota(dIdx[0], dLeafSeg, 0u, queue);
    uint32_t sortedIndex = SegmentedRadixSort<true, uint64_t, uint32_t>
    (
        Span<Span<uint64_t>, 2>(dCodes), Span<Span<uint32_t>, 2>(dIdx),
        Span<Byte>(temp), dLeafSeg, queue
    );
    queue.Barrier().Wait();
    std::memcpy(sortedMorton, dCodes[sortedIndex].data(), nLeaf * sizeof(uint64_t));
    std::memcpy(sortedIdx, dIdx[sortedIndex].data(), nLeaf * sizeof(uint32_t));
    std::memcpy(morton, unsorted.data(), nLeaf * sizeof(uint64_t));

    Span<LBVHNode> dNodes(reinterpret_cast<LBVHNode*>(nodes), nNode);
    Span<uint32_t> dLeafParent(leafParent, nLeaf);
    queue.IssueBlockKernel<KCConstructLBVHInternalNodes>
    (
        "KCConstructLBVHInternalNodes",
        DeviceBlockIssueParams{.gridSize = nSeg * BLOCK_PER_INSTANCE, .blockSize = TPB},
        dNodes, dLeafParent, dLeafSeg, dNodeSeg,
        Span<const uint64_t>(sortedMorton, nLeaf), Span<const uint32_t>(sortedIdx, nLeaf),
        BLOCK_PER_INSTANCE, nSeg
    );
    queue.Barrier().Wait();

    std::vector<uint32_t> counters(nNode, 0u);
    Span<LBVHBoundingBox> dBoxes(reinterpret_cast<LBVHBoundingBox*>(nodeBoxes), nNode);
    queue.IssueBlockKernel<KCUnionLBVHBoundingBoxes>
    (
        "KCUnionLBVHBoundingBoxes",
        DeviceBlockIssueParams{.gridSize = nSeg * BLOCK_PER_INSTANCE, .blockSize = TPB},
        dBoxes, Span<uint32_t>(counters), ToConstSpan(dNodes), ToConstSpan(dLeafParent),
        dLeafSeg, dNodeSeg, dLeafAABBs, BLOCK_PER_INSTANCE, nSeg
    );
    queue.Barrier().Wait();
}

// Closest / any hit over ONE accelerator (identity transform, triangle group, no alpha map):
// TraverseLBVHStack with the leaf functor of AcceleratorLBVH::ClosestHit / FirstHit.
//   rays : n * 8 floats (RayGMem: pos, tMin, dir, tMax — Tracer/TracerTypes.h:L276-283)
//   mode : 0 closest, 1 first (any) hit
//   outPrim[n] : leaf (triangle) index or 0xFFFFFFFF, outT[n], outBary[n*2] (MetaHit a,b), outBack[n]
void ref_lbvh_trace(const float* positions, uint32_t nVerts,
                    const uint32_t* indices, uint32_t nTris,
                    const uint32_t* nodes, const float* nodeBoxes, uint32_t nNode,
                    const float* rays, uint32_t nRays, int mode, int cullFace,
                    uint32_t* outPrim, float* outT, float* outBary, uint8_t* outBack)
{
    using namespace DefaultTriangleDetail;
    TriangleData data = {};
    data.positions = Span<const Vector3>(reinterpret_cast<const Vector3*>(positions), nVerts);
    data.indexList = Span<const Vector3ui>(reinterpret_cast<const Vector3ui*>(indices), nTris);
    Span<const LBVHNode> dNodes(reinterpret_cast<const LBVHNode*>(nodes), nNode);
    Span<const LBVHBoundingBox> dBoxes(reinterpret_cast<const LBVHBoundingBox*>(nodeBoxes), nNode);
    Span<const RayGMem> dRays(reinterpret_cast<const RayGMem*>(rays), nRays);

    auto Work = [&](uint32_t begin, uint32_t end)
    {
        for(uint32_t r = begin; r < end; r++)
        {
            auto [ray, tMM] = RayFromGMem(dRays, r);
            uint32_t bestPrim = 0xFFFFFFFFu; float bestT = tMM[1];
            Vector2 bestHit = Vector2::Zero(); bool bestBack = false;
            BitStack bitStack;
            TraverseLBVHStack<BitStack::MAX_DEPTH>
            (
                bitStack, dNodes, dBoxes, tMM, ray, 0u,
                [&](Vector2& tMinMax, uint32_t leafIndex)
                {
                    Triangle<TransformContextIdentity> prim(TransformContextIdentity{}, data,
                                                            PrimitiveKey::CombinedKey(0, leafIndex));
                    auto isect = prim.Intersects(ray, cullFace != 0);
                    if(!isect) return false;
                    // IsInRange (hpp:L233-236)
                    if(!((isect->t >= tMinMax[0]) && (isect->t < tMinMax[1]))) return false;
                    if(mode == 0)
                    {
                        if(isect->t < tMinMax[1])
                        {
                            bestPrim = leafIndex; bestT = isect->t;
                            bestHit = isect->hit; bestBack = isect->backFace;
                            tMinMax[1] = isect->t;
                        }
                        return false;
                    }
                    bestPrim = leafIndex; bestT = isect->t;
                    bestHit = isect->hit; bestBack = isect->backFace;
                    return true;
                }
            );
            outPrim[r] = bestPrim; outT[r] = bestT;
            outBary[r * 2 + 0] = bestHit[0]; outBary[r * 2 + 1] = bestHit[1];
            outBack[r] = bestBack ? 1 : 0;
        }
    };
    uint32_t nThreads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    uint32_t chunk = (nRays + nThreads - 1) / nThreads;
    for(uint32_t t = 0; t < nThreads; t++)
    {
        uint32_t b = std::min(nRays, t * chunk), e = std::min(nRays, b + chunk);
        if(b < e) pool.emplace_back(Work, b, e);
    }
    for(auto& th : pool) th.join();
}

// Brute force in leaf order — AcceleratorLinear semantics (AcceleratorLinear.hpp:L109-126)
void ref_linear_trace(const float* positions, uint32_t nVerts,
                      const uint32_t* indices, uint32_t nTris,
                      const float* rays, uint32_t nRays, int cullFace,
                      uint32_t* outPrim, float* outT)
{
    using namespace DefaultTriangleDetail;
    TriangleData data = {};
    data.positions = Span<const Vector3>(reinterpret_cast<const Vector3*>(positions), nVerts);
    data.indexList = Span<const Vector3ui>(reinterpret_cast<const Vector3ui*>(indices), nTris);
    Span<const RayGMem> dRays(reinterpret_cast<const RayGMem*>(rays), nRays);
    for(uint32_t r = 0; r < nRays; r++)
    {
        auto [ray, tMM] = RayFromGMem(dRays, r);
        uint32_t best = 0xFFFFFFFFu;
        for(uint32_t i = 0; i < nTris; i++)
        {
            Triangle<TransformContextIdentity> prim(TransformContextIdentity{}, data,
                                                    PrimitiveKey::CombinedKey(0, i));
            auto isect = prim.Intersects(ray, cullFace != 0);
            if(!isect) continue;
            if(!((isect->t >= tMM[0]) && (isect->t < tMM[1]))) continue;
            best = i; tMM[1] = isect->t;
        }
        outPrim[r] = best; outT[r] = tMM[1];
    }
}

// One RGBA texture through the reference's TextureMemory, then `n` reads of its view. format 0 = MR_RGBA_FLOAT, 1 = MR_RGBA8_UNORM;
// `chain` holds suppliedMips levels back to back; genMips = TracerParameters.genMips with mipGenFilter {filterType, filterRadius};
// lod != NULL: view(uv, lod[i]); else view(uv, grads[4 i .. 4 i + 1], grads[4 i + 2 .. 4 i + 3]). clampedTexRes = 0 keeps the default (no
// clamp); sizeOut (optional) receives the texture's final {width, height, mip count}. Returns 0, or -1 on an exception.
int ref_texture_sample(const void* chain, uint32_t w, uint32_t h, uint32_t format, uint32_t interp, uint32_t edge,
                       uint32_t suppliedMips, uint32_t genMips, uint32_t filterType, float filterRadius,
                       const float* uv, const float* lod, const float* grads, uint32_t n, float* out, uint32_t clampedTexRes,
                       uint32_t* sizeOut)
{
    try
    {
        Queue();
        TracerParameters tp;
        if(clampedTexRes) tp.clampedTexRes = clampedTexRes;   // TextureMemory::CreateTexture drops levels / filters the pushed image down
        tp.genMips = genMips != 0;
        tp.mipGenFilter = FilterType{FilterType::E(filterType), filterRadius};
        FilterGeneratorMap fmap;
        fmap.emplace(TextureFilterBox::TypeName, &GenerateType<TextureFilterI, TextureFilterBox, const GPUSystem&, Float>);
        fmap.emplace(TextureFilterTent::TypeName, &GenerateType<TextureFilterI, TextureFilterTent, const GPUSystem&, Float>);
        fmap.emplace(TextureFilterGaussian::TypeName, &GenerateType<TextureFilterI, TextureFilterGaussian, const GPUSystem&, Float>);
        fmap.emplace(TextureFilterMitchellNetravali::TypeName, &GenerateType<TextureFilterI, TextureFilterMitchellNetravali, const GPUSystem&, Float>);
        TextureMemory tm(*gSystem, tp, fmap);
        MRayTextureParameters p;
        p.pixelType = MRayPixelTypeRT(format == 0 ? MRayPixelEnum::MR_RGBA_FLOAT : MRayPixelEnum::MR_RGBA8_UNORM);
        p.colorSpace = MRayColorSpaceEnum::MR_DEFAULT;
        p.gamma = Float(1);
        p.interpolation = MRayTextureInterpEnum(interp); p.edgeResolve = MRayTextureEdgeResolveEnum(edge);
        p.readMode = MRayTextureReadMode::MR_DROP_1;
        TextureId id = tm.CreateTexture2D(Vector2ui(w, h), suppliedMips, p);
        tm.CommitTextures();
        const Byte* src = reinterpret_cast<const Byte*>(chain);
        for(uint32_t level = 0; level < suppliedMips; level++)
        {
            const size_t pixels = size_t(std::max(w >> level, 1u)) * std::max(h >> level, 1u);
            if(format == 0)
            {
                TransientData d(std::in_place_type_t<Vector4>{}, pixels);
                d.Push(Span<const Vector4>(reinterpret_cast<const Vector4*>(src), pixels));
                tm.PushTextureData(id, level, std::move(d));
                src += pixels * sizeof(Vector4);
            }
            else
            {
                TransientData d(std::in_place_type_t<Vector4uc>{}, pixels);
                d.Push(Span<const Vector4uc>(reinterpret_cast<const Vector4uc*>(src), pixels));
                tm.PushTextureData(id, level, std::move(d));
                src += pixels * 4;
            }
        }
        gSystem->SyncAll();   // the clamp filter of PushTextureData runs asynchronously and Finalize releases its staging buffer
        tm.Finalize();
        gSystem->SyncAll();
        if(sizeOut)
        {
            const GenericTexture& gt = tm.Textures().at(id).value().get();
            sizeOut[0] = gt.Extents()[0]; sizeOut[1] = gt.Extents()[1]; sizeOut[2] = gt.MipCount();
        }
        const GenericTextureView& gv = tm.TextureViews().at(id).value().get();
        const auto& view = std::get<TracerTexView<2, Vector3>>(gv);
        for(uint32_t i = 0; i < n; i++)
        {
            Vector2 q(uv[2 * i], uv[2 * i + 1]);
            Vector3 r = lod ? view(q, lod[i])
                            : view(q, Vector2(grads[4 * i], grads[4 * i + 1]), Vector2(grads[4 * i + 2], grads[4 * i + 3]));
            out[3 * i] = r[0]; out[3 * i + 1] = r[1]; out[3 * i + 2] = r[2];
        }
        return 0;
    }
    catch(const MRayError& e) { fprintf(stderr, "ref_texture_sample: %s\n", e.GetError().c_str()); return -1; }
    catch(const std::exception& e) { fprintf(stderr, "ref_texture_sample: %s\n", e.what()); return -1; }
}

// RefractMaterial::RefractRayCone + RayConeSurface::ConeAfterScatter of the reference (Tracer/MaterialsDefault.hpp:L355-462,
// Tracer/TracerTypes.h:L339-352) for n inputs: in[i] = {aperture, width, betaN, wO xyz, geoNormal xyz (already flipped towards wO),
// frontIoR, backIoR, backSide (0 / 1)} (12 floats); out[i] = {aperture, width} of the cone that continues along a TRANSMITTED ray.
void ref_refract_ray_cone(const float* in, uint32_t n, float* out)
{
    using Mat = RefractMatDetail::RefractMaterial<SpectrumContextIdentity>;
    for(uint32_t i = 0; i < n; i++)
    {
        const float* v = in + 12 * size_t(i);
        const Vector3 front(v[9], 0, 0), back(v[10], 0, 0);
        RefractMatDetail::RefractMatData soa{Span<const Vector3>(&front, 1), Span<const Vector3>(&back, 1)};
        DefaultSurface surf{};
        surf.geoNormal = Vector3(v[6], v[7], v[8]);
        surf.shadingTBN = Quaternion::Identity();
        surf.backSide = v[11] != 0.0f;
        SpectrumConverterIdentity conv;
        Mat mat(conv, surf, soa, MaterialKey::CombinedKey(0, 0));
        RayConeSurface rcs{RayCone{v[0], v[1]}, RayCone{v[0], v[1]}, v[2]};
        const Vector3 wO(v[3], v[4], v[5]);
        const RayConeSurface r = mat.RefractRayCone(rcs, wO);
        const RayCone c = r.ConeAfterScatter(-surf.geoNormal, surf.geoNormal);
        out[2 * i] = c.aperture; out[2 * i + 1] = c.width;
    }
}

} // extern "C"
