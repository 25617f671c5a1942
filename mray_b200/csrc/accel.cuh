// accel.cuh — device-side data layout of one triangle accelerator (binary LBVH + 8-wide BVH).
#pragma once
#include "common.cuh"
#include <vector>

namespace mrb
{

// entries of the per-thread stack of the wide traversal kernels (trace.cu); the builders refuse trees that could overflow it
static constexpr uint32_t WIDE_STACK_ENTRIES = 64;

// Reference layouts (Tracer/AcceleratorLBVH.h:L62-77)
struct LBVHNode { uint32_t left, right, parent; };
struct LBVHBox  { float min[3], max[3]; };

// leaf -> primitive mapping of one accelerator: the prim ranges of the surfaces it was built from
// (the reference allows <= 8 per surface, AcceleratorC.h:L281-301; surfaces that share the identity
// transform may be flattened into one accelerator here, so the list lives in device memory and is
// searched by bisection).
// One single-channel 2-D texture read as an alpha map (AlphaMap = TracerTexView<2, Float>, Tracer/AcceleratorC.h:L123):
// channel 0 of `channels`, fp32 or unorm8, one mip level, filtered like every other texture of this library.
struct AlphaTex { const void* data; uint32_t w, h, channels, format, interp, edge, pad; };

struct PrimRanges
{
    uint32_t        count;
    uint32_t        primGroupId;
    const uint32_t* leafStart; // count + 1, prefix sum of range sizes
    const uint32_t* primBegin; // count
    const uint32_t* lmKey;     // count
    const uint32_t* cull;      // count
    // alpha maps (SurfaceParams.alphaMaps, one Optional<TextureId> per prim batch): nullptr = none in this accelerator
    const int32_t*  alphaMap;  // count: index into alphaTex, or -1
    const AlphaTex* alphaTex;
    const float*    uvs;       // vertex UV0 (2 floats per vertex of the primitive group)
};

__host__ __device__ __forceinline__ uint32_t FindRange(const PrimRanges& r, uint32_t leaf)
{
    uint32_t lo = 0, hi = r.count; // invariant: leafStart[lo] <= leaf < leafStart[hi]
    while(hi - lo > 1u)
    {
        uint32_t mid = (lo + hi) >> 1;
        if(leaf >= r.leafStart[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// 8-wide compressed node, 80 bytes = 5 x 128-bit loads (after Ylitie et al. 2017, "CWBVH").
//   q0 : p.x, p.y, p.z (node origin, float bits), {ex, ey, ez, imask} bytes (2^e quantisation
//        step per axis as a biased float exponent; imask = slots holding internal children)
//   q1 : childBase (first internal child node), triBase (first triangle record),
//        meta[0..3], meta[4..7]   meta = 0                      : empty slot
//                                      = 0b001'11sss (0x38|slot): internal child
//                                      = unary(count)<<5 | offset: leaf, `count` (1..3) triangle
//                                        records starting at triBase + offset (offset < 24)
//   q2 : qlo.x[0..7], qlo.y[0..7]   q3 : qlo.z[0..7], qhi.x[0..7]   q4 : qhi.y[0..7], qhi.z[0..7]
struct alignas(16) WideNode { uint4 q[5]; };

// Triangle record, 48 bytes = 3 x 128-bit loads. e0/e1 are the float differences p1-p0, p2-p0
// exactly as Ray::IntersectsTriangle forms them (Core/Ray.hpp:L133-134), so the intersection
// arithmetic stays bit-identical to the reference.
//   v0 : p0.xyz, leaf index (position in the accelerator's leaf list)
//   v1 : e0.xyz, rank (position after the Morton sort = the reference's visit order; tie-break)
//   v2 : e1.xyz, flags (bit0 = cull back face, bit1 = the range has an alpha map) | range index << 8
struct alignas(16) TriRecord { float4 v0, v1, v2; };

struct AccelData
{
    // inputs (device copies owned by the accelerator)
    const float*    positions = nullptr; // V*3
    const uint32_t* indices = nullptr;   // T*3
    PrimRanges      ranges = {};
    uint32_t        leafCount = 0;
    uint32_t        nodeCount = 0;
    // binary LBVH (reference layout)
    float*    leafAABB = nullptr;     // leaf*6
    uint64_t* morton = nullptr;       // leaf (leaf order)
    uint64_t* sortedMorton = nullptr; // leaf (sorted)
    uint32_t* sortedLeaf = nullptr;   // leaf (sorted) -> leaf index
    LBVHNode* nodes = nullptr;        // node
    uint32_t* leafParent = nullptr;   // leaf
    LBVHBox*  boxes = nullptr;        // node
    uint2*    nodeRange = nullptr;    // node : [first,last] sorted positions covered
    uint32_t* accelAABBEnc = nullptr; // 6 order-preserving encoded floats + 2 flags
    // wide BVH
    WideNode*  wideNodes = nullptr;
    TriRecord* tris = nullptr;
    uint32_t   wideNodeCapacity = 0;
    uint32_t   wideNodeCount = 0;
    uint32_t   wideDepth = 0;
    uint32_t   maxLeafSize = 3;       // leaves per wide leaf child (3 triangles; 1 instance in a top-level tree)
    uint32_t*  leafOfSlot = nullptr;  // top-level tree only: leaf record slot -> instance index
};

// One instance of a two-level scene (AcceleratorGroup instance + transform, AcceleratorCommon.cu:L452-572).
struct InstanceRec
{
    float            invTransform[12];   // world -> local, row-major 3x4 (TransformContextSingle::InvApply)
    float            worldAABB[6];       // leaf AABB of the top-level tree
    const WideNode*  wideNodes;
    const TriRecord* tris;
    const float*     leafAABB;
    const float*     positions;          // audit / exact fallback
    const uint32_t*  indices;
    const LBVHNode*  nodes;
    const LBVHBox*   boxes;
    PrimRanges       ranges;
    uint32_t         accelKey, transKey, identity, leafCount;
};

struct SceneData
{
    AccelData          tlas;              // leaves = instances (leafAABB = world AABBs)
    const InstanceRec* instances = nullptr;
    uint32_t           instanceCount = 0;
    uint32_t           hasAlpha = 0;      // any instance's accelerator carries alpha maps: the ALPHA kernels run
};

} // namespace mrb

struct mrb_accel_t
{
    mrb::AccelData   d;
    // host copies of the prim-range table (d.ranges points at the device copy)
    std::vector<uint32_t> hLeafStart, hPrimBegin, hLmKey, hCull;
    std::vector<int32_t> hAlphaMap;        // per range: alpha texture index or -1 (empty = no alpha maps)
    mrb::DeviceBlock mem;
    mrb_accel_info   info = {};
    uint32_t         flags = 0;
    uint32_t         accelKey = 0;
    uint32_t         vertexCount = 0, triangleCount = 0;   // of the primitive group (ranges may cover a subset)
};

struct mrb_scene_t
{
    mrb::SceneData   d;
    mrb::DeviceBlock mem;
    std::vector<mrb_accel> accels;
    std::vector<mrb_instance_desc> hInstances;   // host copy (renderer: forward transforms, keys)
    std::vector<std::vector<uint32_t>> hInstanceKeys; // per instance: LightOrMatKey override per prim range (empty = accelerator's)
    float            aabb[6] = {};
    float            buildMs = 0.f;
};
