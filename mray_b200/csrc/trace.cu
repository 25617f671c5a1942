// trace.cu — closest-hit / any-hit ray casting.
//
//  KTraceWide   : product path. Per-thread traversal of the 8-wide quantised BVH (node = 5 x 128-bit
//                 loads, triangle = 3 x 128-bit loads), octant-ordered child visits, node-group
//                 stack. Box tests are CONSERVATIVE with respect to the reference's slab test on
//                 the exact binary boxes (quantised boxes enclose them; a per-node slack covers the
//                 rounding of both formulations), triangle tests are the reference's Moller-Trumbore
//                 arithmetic bit for bit, so the kernel finds the true minimum over a SUPERSET of the
//                 triangles the reference tests. The reference can differ from that minimum only
//                 when one of its own (exact, binary) box tests rejects the winner numerically, which
//                 needs either a second candidate within rounding distance of the winner or a ray that
//                 grazes the winner's own AABB. Both are detected: culling is relaxed to
//                 t <= tBest * (1 + 2^-16) so every near-tie candidate is seen, and the winner is
//                 certified with the reference's slab test on its leaf AABB (all ancestor boxes
//                 enclose it and the computed slab values are monotone in the box). Certified rays are
//                 provably what BaseAcceleratorLBVH::CastRays returns; the rest (about 1e-6 of the rays
//                 of config 2) are appended to a list and re-traced by KTraceBinary, the reference's
//                 algorithm verbatim.
//  KTraceBinary : audit path. The reference's own algorithm: binary LBVH, left-first stack
//                 traversal (TraverseLBVHStack, AcceleratorLBVH.hpp:L109-167), Ray::IntersectsAABB
//                 (Core/Ray.hpp:L192-219) with identical arithmetic.
#include "accel.cuh"
#include <cstdio>
#include <cfloat>
#include <cstdlib>

namespace mrb
{
namespace
{

constexpr int TRACE_TPB = 128;
// resident blocks per SM the wide kernels are compiled for (register budget 65536 / (128 x blocks)): the exact-resolution
// code is called out of line from the epilogue, and without a cap its register needs would set the kernel's
#ifndef MRB_WIDE_BLOCKS_CLOSEST
#define MRB_WIDE_BLOCKS_CLOSEST 7
#endif
#ifndef MRB_WIDE_BLOCKS_ANY
#define MRB_WIDE_BLOCKS_ANY 8
#endif
constexpr int WIDE_STACK = int(WIDE_STACK_ENTRIES);   // capi.cu refuses trees with 2 * depth (+ 3 + 2 * top-level depth) above it

struct HitRecord
{
    float    t;
    float    u, v;   // Moller-Trumbore u, v
    uint32_t leaf;
    uint32_t rank;
    uint32_t flags;
};

// Ray::IntersectsTriangle (Core/Ray.hpp:L121-167) with Math::Cross / Math::Dot exactly as
// Core/Math.h:L1586-1596,L1627-1633 build them from (true) FMAs.
__device__ __forceinline__ float Dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    float r = __fmaf_rn(ax, bx, 0.0f);
    r = __fmaf_rn(ay, by, r);
    r = __fmaf_rn(az, bz, r);
    return r;
}

__device__ __forceinline__ bool RayTriangle(float ox, float oy, float oz, float dx, float dy, float dz,
                                            const float4& v0, const float4& e0, const float4& e1, bool cull,
                                            float& tOut, float& uOut, float& vOut)
{
    const float eps = 1.0e-7f;
    float px = __fmaf_rn(dy, e1.z, -__fmul_rn(dz, e1.y));
    float py = __fmaf_rn(dz, e1.x, -__fmul_rn(dx, e1.z));
    float pz = __fmaf_rn(dx, e1.y, -__fmul_rn(dy, e1.x));
    float det = Dot3(e0.x, e0.y, e0.z, px, py, pz);
    bool back = det < eps;
    bool parallel = fabsf(det) < eps;
    if((cull && back) || parallel) return false;
    float invDet = __frcp_rn(det);
    float tx = __fsub_rn(ox, v0.x), ty = __fsub_rn(oy, v0.y), tz = __fsub_rn(oz, v0.z);
    float u = __fmul_rn(Dot3(tx, ty, tz, px, py, pz), invDet);
    if(u < 0.0f || u > 1.0f) return false;
    float qx = __fmaf_rn(ty, e0.z, -__fmul_rn(tz, e0.y));
    float qy = __fmaf_rn(tz, e0.x, -__fmul_rn(tx, e0.z));
    float qz = __fmaf_rn(tx, e0.y, -__fmul_rn(ty, e0.x));
    float v = __fmul_rn(Dot3(dx, dy, dz, qx, qy, qz), invDet);
    if(v < 0.0f || __fadd_rn(v, u) > 1.0f) return false;
    float t = __fmul_rn(Dot3(e1.x, e1.y, e1.z, qx, qy, qz), invDet);
    if(t <= eps) return false;
    tOut = t; uOut = u; vOut = v;
    return true;
}

__device__ __forceinline__ void WriteHit(const AccelData& a, uint32_t accelKey, uint32_t r, const HitRecord& h,
                                         mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays)
{
    uint32_t ri = h.flags >> 8;
    uint32_t prim = a.ranges.primBegin[ri] + (h.leaf - a.ranges.leafStart[ri]);
    uint4 keys = make_uint4((a.ranges.primGroupId << 28) | prim, a.ranges.lmKey[ri], 0u, accelKey);
    *reinterpret_cast<uint4*>(hitKeys + r) = keys;
    float w = __fsub_rn(__fsub_rn(1.0f, h.u), h.v);
    *reinterpret_cast<float2*>(metaHits + r) = make_float2(w, h.u); // Vector2(baryCoords) = (1-u-v, u)
    rays[r].tMax = h.t;
}

__device__ __forceinline__ float StdMax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float StdMin(float a, float b) { return (b < a) ? b : a; }


// ------------------------------------------------------------------------------------------------
// Stochastic alpha test (AcceleratorLBVH::IntersectionCheck, Tracer/AcceleratorLBVH.hpp:L263-282)
// ------------------------------------------------------------------------------------------------
// The alpha map's texture view, restating the reference's host-backend view for ONE channel
// (Device/CPU/TextureViewCPU.h: texel centres at +0.5, nearest or bilinear with unfused lerps, wrap / clamp / mirror).
__device__ __forceinline__ int AlphaEdge(int i, int n, uint32_t edge)
{
    if(edge == 1u) return min(max(i, 0), n - 1);
    if(edge == 2u)
    {
        const int dim = i / n;
        i = i % n;
        if(i < 0) i += n;
        if((dim & 1) == 1) i = n - i;
        return min(i, n - 1);
    }
    i = i % n;
    if(i < 0) i += n;
    return i;
}
__device__ __forceinline__ float AlphaTexel(const AlphaTex& t, int x, int y)
{
    const size_t o = (size_t(y) * t.w + size_t(x)) * t.channels;
    if(t.format == 0u) return __ldg(static_cast<const float*>(t.data) + o);
    return __fmul_rn(float(__ldg(static_cast<const uint8_t*>(t.data) + o)), 1.0f / 255.0f);
}
__device__ __forceinline__ float AlphaLerp(float a, float b, float t) { return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, t)), __fmul_rn(b, t)); }
__device__ __forceinline__ float SampleAlpha(const AlphaTex& t, float u, float v)
{
    const float tu = __fmul_rn(u, float(t.w)), tv = __fmul_rn(v, float(t.h));
    if(t.interp == 0u)
    {
        const int x = int(roundf(tu - 0.5f)), y = int(roundf(tv - 0.5f));
        return AlphaTexel(t, AlphaEdge(x, int(t.w), t.edge), AlphaEdge(y, int(t.h), t.edge));
    }
    float bx, by;
    float fx = modff(tu - 0.5f, &bx), fy = modff(tv - 0.5f, &by);
    int x0 = int(bx), y0 = int(by);
    if(fx < 0.0f) { x0 -= 1; fx = fabsf(fx); }
    if(fy < 0.0f) { y0 -= 1; fy = fabsf(fy); }
    const int xa = AlphaEdge(x0, int(t.w), t.edge), xb = AlphaEdge(x0 + 1, int(t.w), t.edge);
    const int ya = AlphaEdge(y0, int(t.h), t.edge), yb = AlphaEdge(y0 + 1, int(t.h), t.edge);
    const float p0 = AlphaLerp(AlphaTexel(t, xa, ya), AlphaTexel(t, xb, ya), fx);
    const float p1 = AlphaLerp(AlphaTexel(t, xa, yb), AlphaTexel(t, xb, yb), fx);
    return AlphaLerp(p0, p1, fy);
}
// Does the hit (u, v = Moller-Trumbore coordinates) on `leaf` of range `ri` survive its alpha map? uv = the hit's
// interpolated UV0 (Triangle::SurfaceParametrization: uv0 a + uv1 b + uv2 c with (a, b) = (1 - u - v, u)); the hit is
// dropped when xi >= alpha. xi: the reference takes the next float of the ray's backup PCG32; here one hash of
// (cast seed, the ray's bits, leaf), so a (ray, triangle) pair always gets the same answer — the wide kernels, the exact
// resolution and the binary fallback may each meet the same triangle. Out of line: only alpha-mapped triangles pay.
__device__ __noinline__ bool AlphaKeepsHit(const PrimRanges& rg, const uint32_t* __restrict__ indices, uint32_t leaf, uint32_t ri,
                                           float u, float v, const mrb_ray_gmem* ray, uint32_t seed)
{
    const AlphaTex t = rg.alphaTex[rg.alphaMap[ri]];
    const uint32_t prim = rg.primBegin[ri] + (leaf - rg.leafStart[ri]);
    const uint32_t i0 = indices[3 * size_t(prim)], i1 = indices[3 * size_t(prim) + 1], i2 = indices[3 * size_t(prim) + 2];
    const float2 t0 = *reinterpret_cast<const float2*>(rg.uvs + 2 * size_t(i0));
    const float2 t1 = *reinterpret_cast<const float2*>(rg.uvs + 2 * size_t(i1));
    const float2 t2 = *reinterpret_cast<const float2*>(rg.uvs + 2 * size_t(i2));
    const float a = __fsub_rn(__fsub_rn(1.0f, u), v), b = u, c = __fsub_rn(__fsub_rn(1.0f, a), b);
    const float tu = __fadd_rn(__fadd_rn(__fmul_rn(t0.x, a), __fmul_rn(t1.x, b)), __fmul_rn(t2.x, c));
    const float tv = __fadd_rn(__fadd_rn(__fmul_rn(t0.y, a), __fmul_rn(t1.y, b)), __fmul_rn(t2.y, c));
    const float alpha = SampleAlpha(t, tu, tv);
    // the "ray" half of the key is the WORLD ray's own bits (as stored in the ray buffer, which no kernel rewrites except
    // tMax): independent of slot assignment and of ray-index indirection, identical in every kernel that meets the pair
    const uint4 w0 = *reinterpret_cast<const uint4*>(ray);
    const uint4 w1 = *(reinterpret_cast<const uint4*>(ray) + 1);
    uint32_t h = seed ^ (leaf * 0x85EBCA6Bu);
    h = (h ^ w0.x) * 0x9E3779B1u; h = (h ^ (h >> 15) ^ w0.y) * 0x85EBCA77u; h = (h ^ (h >> 13) ^ w0.z) * 0xC2B2AE3Du;
    h = (h ^ (h >> 16) ^ w1.x) * 0x27D4EB2Fu; h = (h ^ (h >> 15) ^ w1.y) * 0x165667B1u; h ^= w1.z;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    const float xi = float(h >> 8) * 5.9604644775390625e-08f;   // [0, 1)
    return xi < alpha;
}

template<bool ALPHA>
__device__ __forceinline__ bool AlphaGate(const PrimRanges& rg, const uint32_t* __restrict__ indices, uint32_t flags, uint32_t leaf,
                                          float u, float v, const mrb_ray_gmem* ray, uint32_t seed)
{
    if constexpr(ALPHA) return !(flags & 2u) || AlphaKeepsHit(rg, indices, leaf, flags >> 8, u, v, ray, seed);
    else return true;
}

// Ray::IntersectsAABB (Core/Ray.hpp:L192-219), identical arithmetic (IEEE div / sub / mul, host
// std::min / std::max semantics).
__device__ __forceinline__ bool SlabExact(const float* __restrict__ b, const float o[3], const float invD[3],
                                          float tMin, float tMax)
{
    float o0 = tMin, o1 = tMax;
    #pragma unroll
    for(int k = 0; k < 3; k++)
    {
        float t0 = __fmul_rn(__fsub_rn(b[k], o[k]), invD[k]);
        float t1 = __fmul_rn(__fsub_rn(b[3 + k], o[k]), invD[k]);
        if(invD[k] < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        o0 = StdMax(o0, StdMin(t0, t1));
        o1 = StdMin(o1, StdMax(t0, t1));
    }
    return o1 >= o0;
}

// Would the reference's traversal reach this leaf with upper bound tUpper? Sufficient condition:
// its own slab test passes on the leaf's AABB (every ancestor box encloses it).
__device__ __forceinline__ bool CertifyLeaf(const AccelData& a, uint32_t leaf, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float tMin, float tUpper)
{
    const float2* bp = reinterpret_cast<const float2*>(a.leafAABB + 6 * size_t(leaf));
    float2 b0 = __ldg(bp), b1 = __ldg(bp + 1), b2 = __ldg(bp + 2);
    const float box[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
    const float o[3] = {ox, oy, oz};
    const float invD[3] = {__fdiv_rn(1.0f, dx), __fdiv_rn(1.0f, dy), __fdiv_rn(1.0f, dz)};
    return SlabExact(box, o, invD, tMin, tUpper);
}

constexpr float NEAR_TIE = 1.0000152587890625f; // 1 + 2^-16
// counters of one wide cast: [0] uncertified rays, [1] of which near ties, [2] of which leaf-AABB
// certification failures, [3] dynamic-fetch cursor, [4] rays handed to the full binary fallback

// ------------------------------------------------------------------------------------------------
// Exact binary traversal (audit path, and fallback of the rays the wide path could not certify)
// ------------------------------------------------------------------------------------------------
// One ray through the reference's own algorithm: binary LBVH, left-first stack traversal
// (TraverseLBVHStack, AcceleratorLBVH.hpp:L109-167), Ray::IntersectsAABB arithmetic.
template<bool ANY_HIT, bool ALPHA>
__device__ __forceinline__ void TraceBinaryRayBody(const AccelData& a, uint32_t accelKey,
                                                   mrb_hit_key_pack* __restrict__ hitKeys, mrb_meta_hit* __restrict__ metaHits,
                                                   uint32_t* __restrict__ visibleBits, mrb_ray_gmem* rays, uint32_t r, uint32_t alphaSeed)
{
    const float4 r0 = reinterpret_cast<const float4*>(rays + r)[0];
    const float4 r1 = reinterpret_cast<const float4*>(rays + r)[1];
    const float o[3] = {r0.x, r0.y, r0.z}, d[3] = {r1.x, r1.y, r1.z};
    const float tMin = r0.w; float tMax = r1.w;
    float invD[3];
    #pragma unroll
    for(int k = 0; k < 3; k++) invD[k] = __fdiv_rn(1.0f, d[k]);

    HitRecord best; best.t = tMax; best.rank = 0u; best.leaf = INVALID_U32; best.u = best.v = 0.f; best.flags = 0u;
    uint32_t stack[160];
    int sp = 0;
    stack[sp++] = 0u;
    while(sp > 0)
    {
        uint32_t ni = stack[--sp];
        if(ni == INVALID_U32) continue;
        if(ni & LEAF_FLAG)
        {
            uint32_t leaf = ni & ~LEAF_FLAG;
            const uint32_t ri = (a.ranges.count == 1u) ? 0u : FindRange(a.ranges, leaf);
            uint32_t prim = a.ranges.primBegin[ri] + (leaf - a.ranges.leafStart[ri]);
            uint32_t i0 = a.indices[3 * size_t(prim)], i1 = a.indices[3 * size_t(prim) + 1], i2 = a.indices[3 * size_t(prim) + 2];
            float4 v0, e0, e1;
            const float* p0 = a.positions + 3 * size_t(i0);
            const float* p1 = a.positions + 3 * size_t(i1);
            const float* p2 = a.positions + 3 * size_t(i2);
            v0 = make_float4(p0[0], p0[1], p0[2], 0.f);
            e0 = make_float4(__fsub_rn(p1[0], p0[0]), __fsub_rn(p1[1], p0[1]), __fsub_rn(p1[2], p0[2]), 0.f);
            e1 = make_float4(__fsub_rn(p2[0], p0[0]), __fsub_rn(p2[1], p0[1]), __fsub_rn(p2[2], p0[2]), 0.f);
            float t, u, v;
            if(!RayTriangle(o[0], o[1], o[2], d[0], d[1], d[2], v0, e0, e1, a.ranges.cull[ri] != 0u, t, u, v)) continue;
            if(!(t >= tMin && t < tMax)) continue;
            if(ALPHA && a.ranges.alphaMap && a.ranges.alphaMap[ri] >= 0 && !AlphaKeepsHit(a.ranges, a.indices, leaf, ri, u, v, rays + r, alphaSeed)) continue;
            best.t = t; best.u = u; best.v = v; best.leaf = leaf; best.flags = ri << 8;
            tMax = t;
            if(ANY_HIT) break;
        }
        else
        {
            const float* b = reinterpret_cast<const float*>(a.boxes + ni);
            if(SlabExact(b, o, invD, tMin, tMax))
            {
                LBVHNode nd = a.nodes[ni];
                stack[sp++] = nd.right;
                stack[sp++] = nd.left;
                // a dependent chain of (node, box) loads for a handful of rays: pull the records
                // the next pops will need towards L1 while the current step finishes
                if(!(nd.left & LEAF_FLAG))
                {
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(a.boxes + nd.left));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(a.nodes + nd.left));
                }
                if(!(nd.right & LEAF_FLAG))
                {
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(a.boxes + nd.right));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(a.nodes + nd.right));
                }
            }
        }
    }
    if(best.leaf != INVALID_U32)
    {
        if(ANY_HIT) atomicAnd(&visibleBits[r >> 5], ~(1u << (r & 31u)));
        else WriteHit(a, accelKey, r, best, hitKeys, metaHits, rays);
    }
}

// Out-of-line copies for the wide kernels: a call costs nothing until one of the ~1e-4 rays that need it shows up,
// and keeps the 160-entry stack and the registers of this code out of the hot loop.
__device__ __noinline__ void TraceBinaryRayClosest(const AccelData& a, uint32_t accelKey, mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits,
                                                   mrb_ray_gmem* rays, uint32_t r)
{ TraceBinaryRayBody<false, false>(a, accelKey, hitKeys, metaHits, nullptr, rays, r, 0u); }
__device__ __noinline__ void TraceBinaryRayAny(const AccelData& a, uint32_t* visibleBits, mrb_ray_gmem* rays, uint32_t r)
{ TraceBinaryRayBody<true, false>(a, 0u, nullptr, nullptr, visibleBits, rays, r, 0u); }
// ... and for accelerators with alpha maps (their own copies, so the kernels of scenes without alpha maps compile to exactly
// what they were before alpha maps existed)
__device__ __noinline__ void TraceBinaryRayClosestAlpha(const AccelData& a, uint32_t accelKey, mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits,
                                                        mrb_ray_gmem* rays, uint32_t r, uint32_t alphaSeed)
{ TraceBinaryRayBody<false, true>(a, accelKey, hitKeys, metaHits, nullptr, rays, r, alphaSeed); }
__device__ __noinline__ void TraceBinaryRayAnyAlpha(const AccelData& a, uint32_t* visibleBits, mrb_ray_gmem* rays, uint32_t r, uint32_t alphaSeed)
{ TraceBinaryRayBody<true, true>(a, 0u, nullptr, nullptr, visibleBits, rays, r, alphaSeed); }

template<bool ANY_HIT>
__global__ void __launch_bounds__(TRACE_TPB)
KTraceBinary(const __grid_constant__ AccelData a, uint32_t accelKey,
             mrb_hit_key_pack* __restrict__ hitKeys, mrb_meta_hit* __restrict__ metaHits,
             uint32_t* __restrict__ visibleBits,
             mrb_ray_gmem* rays, const uint32_t* __restrict__ rayIndices, uint32_t rayCount, uint32_t alphaSeed)
{
    for(uint32_t i = blockIdx.x * TRACE_TPB + threadIdx.x; i < rayCount; i += gridDim.x * TRACE_TPB)
        TraceBinaryRayBody<ANY_HIT, true>(a, accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices ? rayIndices[i] : i, alphaSeed);
}

// Exact resolution of a ray the wide path could not certify, WITHOUT a full re-traversal: the
// reference's answer is decided by the (at most two) candidates inside the near-tie window, by the
// order it visits them (Morton rank) and by its box tests on their ancestor chains, evaluated with
// the tMax the reference holds when it first enters each node: the t of an already accepted
// candidate whose rank lies left of the node's range, else anything >= tUpper. A box that fails
// with a real candidate's t is a genuine cull; a box that fails with tUpper is undecidable here and
// the ray goes to the full binary traversal (TraceBinaryRay), as do rays with three or more
// candidates in the window. Called by the lane that owns the ray, straight from the wide kernel's epilogue
// (round 1 ran this and the binary fallback as two more launches per cast: ~17 us of fixed cost each).
template<bool ALPHA>
__device__ __forceinline__ void ResolveExactRayBody(const AccelData& a, uint32_t accelKey,
                                             mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays,
                                             uint32_t* counters, uint32_t r, bool full, HitRecord c0, HitRecord c1, bool hasSecond,
                                             uint32_t alphaSeed)
{
    HitRecord c[2] = {c0, c1};   // by value: the caller's records must stay in registers
    const int n = hasSecond ? 2 : 1;
    const float4 r0 = reinterpret_cast<const float4*>(rays + r)[0];
    const float4 r1 = reinterpret_cast<const float4*>(rays + r)[1];
    const float o[3] = {r0.x, r0.y, r0.z};
    const float invD[3] = {__fdiv_rn(1.0f, r1.x), __fdiv_rn(1.0f, r1.y), __fdiv_rn(1.0f, r1.z)};
    const float tMin = r0.w;
    const float tUpper = fminf(r1.w, c[0].t * NEAR_TIE);
    const int first = (n == 2 && c[1].rank < c[0].rank) ? 1 : 0;
    int accepted = -1;
    for(int s = 0; s < n && !full; s++)
    {
        const HitRecord& z = c[s == 0 ? first : 1 - first];
        // a later-visited candidate only matters when it is strictly closer than the accepted one
        if(accepted >= 0 && !(z.t < c[accepted].t)) continue;
        bool reach = true;
        // every ancestor box encloses the leaf box and the slab values are monotone in the box: if the leaf
        // box passes with tUpper, so does every ancestor tested with tUpper, and once an ancestor contains
        // the accepted candidate (range start <= its rank) all higher ones do too, so the walk can stop at
        // the first ancestor that is not tested with the accepted t (the common ancestor, a few levels up)
        const bool leafPasses = SlabExact(a.leafAABB + 6 * size_t(z.leaf), o, invD, tMin, tUpper);
        uint32_t ni = a.leafParent[z.leaf];
        while(ni != INVALID_U32)
        {
            const bool usePrior = accepted >= 0 && c[accepted].rank < a.nodeRange[ni].x;
            if(!usePrior && leafPasses) break;
            const float tcur = usePrior ? c[accepted].t : tUpper;
            if(!SlabExact(reinterpret_cast<const float*>(a.boxes + ni), o, invD, tMin, tcur))
            {
                if(usePrior) reach = false; else full = true;
                break;
            }
            // a box that passes with tUpper is enclosed by every remaining ancestor, which therefore pass too
            // (edge hits make the flat leaf boxes marginal, not the boxes above them)
            if(!usePrior) break;
            ni = a.nodes[ni].parent;
        }
        if(!full && reach && (accepted < 0 || z.t < c[accepted].t)) accepted = (s == 0 ? first : 1 - first);
    }
    if(full || accepted < 0)
    {
        atomicAdd(counters + 4, 1u);
        if constexpr(ALPHA) TraceBinaryRayClosestAlpha(a, accelKey, hitKeys, metaHits, rays, r, alphaSeed);
        else TraceBinaryRayClosest(a, accelKey, hitKeys, metaHits, rays, r);
    }
    else WriteHit(a, accelKey, r, c[accepted], hitKeys, metaHits, rays);
}

__device__ __noinline__ void ResolveExactRay(const AccelData& a, uint32_t accelKey,
                                             mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays,
                                             uint32_t* counters, uint32_t r, bool full, HitRecord c0, HitRecord c1, bool hasSecond)
{ ResolveExactRayBody<false>(a, accelKey, hitKeys, metaHits, rays, counters, r, full, c0, c1, hasSecond, 0u); }
__device__ __noinline__ void ResolveExactRayAlpha(const AccelData& a, uint32_t accelKey,
                                                  mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays,
                                                  uint32_t* counters, uint32_t r, bool full, HitRecord c0, HitRecord c1, bool hasSecond, uint32_t alphaSeed)
{ ResolveExactRayBody<true>(a, accelKey, hitKeys, metaHits, rays, counters, r, full, c0, c1, hasSecond, alphaSeed); }

// ------------------------------------------------------------------------------------------------
// Wide traversal
// ------------------------------------------------------------------------------------------------
// Quantised byte j of `word` as the float 32768 + q (bits 0x4700qq00): ONE PRMT, no int->float
// conversion. The 32768 offset is folded into the per-node constant (c - 32768 s), whose rounding
// error (<= 2^-9 |s|, 0.002 quantisation cells) is covered by the node slack.
template<int J>
__device__ __forceinline__ float QFI(uint32_t word, uint32_t magic)
{
    uint32_t d; // selector must stay an immediate: {magic.b3, magic.b2, word.bJ, magic.b0}
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(word), "r"(magic), "n"(0x7604 | (J << 4)));
    return __uint_as_float(d);
}
__device__ __forceinline__ float QF(uint32_t word, int j, uint32_t magic)
{
    switch(j)
    {
        case 0: return QFI<0>(word, magic);
        case 1: return QFI<1>(word, magic);
        case 2: return QFI<2>(word, magic);
        default: return QFI<3>(word, magic);
    }
}

struct TraceParams
{
    uint32_t triDiv;    // triangle phase runs when lanesWithTriangles * triDiv >= liveLanes (or no lane has node work)
    uint32_t fetchThr;  // refill idle lanes with new rays when fewer than this many lanes are live
    uint32_t magic;     // 0x47000000 (float 32768): see QF
    uint32_t alphaSeed; // seed of this cast's stochastic alpha decisions (AlphaKeepsHit)
};

// Persistent warps: every lane owns one ray at a time; finished lanes are refilled from a global
// counter (dynamic fetch). Inside a warp the loop is phase-uniform — at most one triangle test and
// one node step per iteration, each executed by all lanes that have that kind of work — so the
// 200-instruction node step and the 60-instruction triangle test always run with as many lanes as
// possible; triangle groups are postponed (kept / pushed on the stack) until enough lanes have one.
// ALPHA: the accelerator has alpha maps. Scenes without them run the instantiation that contains no trace of the alpha test
// (its call sites, live ranges and seed argument cost the 72-register closest-hit kernel 1-3 % otherwise).
template<bool ANY_HIT, bool ALPHA>
__global__ void __launch_bounds__(TRACE_TPB, ANY_HIT ? MRB_WIDE_BLOCKS_ANY : MRB_WIDE_BLOCKS_CLOSEST)
KTraceWide(const __grid_constant__ AccelData a, uint32_t accelKey,
           mrb_hit_key_pack* __restrict__ hitKeys, mrb_meta_hit* __restrict__ metaHits,
           uint32_t* __restrict__ visibleBits,
           mrb_ray_gmem* rays, const uint32_t* __restrict__ rayIndices, uint32_t rayCount,
           uint32_t* __restrict__ counters, TraceParams prm)
{
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint4* __restrict__ nodeBase = reinterpret_cast<const uint4*>(a.wideNodes);
    const float4* __restrict__ triBase = reinterpret_cast<const float4*>(a.tris);

    bool hasRay = false, finished = false, exhausted = false;
    uint32_t r = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, idx = 0, idy = 0, idz = 0;
    float tMin = 0, tMaxOrig = 0, tMax = 0;
    bool overflow = false; // more than two candidates inside the near-tie window
    uint32_t oct = 0;
    bool uncertified = false, done = false;
    HitRecord best; best.t = 0; best.u = best.v = 0; best.leaf = INVALID_U32; best.rank = 0; best.flags = 0;
    HitRecord second = best; // runner-up inside the near-tie window of `best`
    uint2 G = make_uint2(0u, 0u), T = make_uint2(0u, 0u);
    uint2 stack[WIDE_STACK];
    int sp = 0;

    while(true)
    {
        // ------------------------------ epilogue of finished rays ------------------------------
        if(finished)
        {
            finished = false;
            if(ANY_HIT)
            {
                if(done) atomicAnd(&visibleBits[r >> 5], ~(1u << (r & 31u)));
                else if(uncertified)
                {   // a hit whose leaf box the reference's own slab test might reject: ask the reference's algorithm
                    atomicAdd(counters, 1u); atomicAdd(counters + 2, 1u); atomicAdd(counters + 4, 1u);
                    if constexpr(ALPHA) TraceBinaryRayAnyAlpha(a, visibleBits, rays, r, prm.alphaSeed);
                    else TraceBinaryRayAny(a, visibleBits, rays, r);
                }
            }
            else if(best.leaf != INVALID_U32)
            {
                const float tUpper = fminf(tMaxOrig, best.t * NEAR_TIE);
                const bool hasSecond = second.leaf != INVALID_U32 && second.t <= tUpper;
                if(!hasSecond && !overflow && CertifyLeaf(a, best.leaf, ox, oy, oz, dx, dy, dz, tMin, tUpper))
                    WriteHit(a, accelKey, r, best, hitKeys, metaHits, rays);
                else
                {
                    // settle it here and now with the reference's own box arithmetic (out of line, ~1e-4 of the rays)
                    atomicAdd(counters, 1u);
                    atomicAdd(counters + ((hasSecond || overflow) ? 1 : 2), 1u); // [1] near ties, [2] uncertified leaf
                    if constexpr(ALPHA) ResolveExactRayAlpha(a, accelKey, hitKeys, metaHits, rays, counters, r, overflow, best, second, hasSecond, prm.alphaSeed);
                    else ResolveExactRay(a, accelKey, hitKeys, metaHits, rays, counters, r, overflow, best, second, hasSecond);
                }
            }
        }
        // ------------------------------------ dynamic fetch ------------------------------------
        if(!exhausted)
        {
            const uint32_t need = __ballot_sync(FULL, !hasRay);
            if(need)
            {
                const int leader = __ffs(int(need)) - 1;
                uint32_t base = 0;
                if(int(lane) == leader) base = atomicAdd(counters + 3, uint32_t(__popc(need)));
                base = __shfl_sync(FULL, base, leader);
                if(!hasRay)
                {
                    const uint32_t i = base + uint32_t(__popc(need & ltMask));
                    if(i < rayCount)
                    {
                        r = rayIndices ? rayIndices[i] : i;
                        const float4 r0 = reinterpret_cast<const float4*>(rays + r)[0];
                        const float4 r1 = reinterpret_cast<const float4*>(rays + r)[1];
                        ox = r0.x; oy = r0.y; oz = r0.z; tMin = r0.w;
                        dx = r1.x; dy = r1.y; dz = r1.z; tMaxOrig = r1.w;
                        // reciprocal direction; tiny components are clamped so no inf/NaN enters the box tests
                        idx = (fabsf(dx) > 1e-30f) ? 1.0f / dx : copysignf(1e30f, dx);
                        idy = (fabsf(dy) > 1e-30f) ? 1.0f / dy : copysignf(1e30f, dy);
                        idz = (fabsf(dz) > 1e-30f) ? 1.0f / dz : copysignf(1e30f, dz);
                        oct = (dx < 0.0f ? 4u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 1u : 0u);
                        tMax = tMaxOrig; overflow = false;
                        best.t = tMaxOrig; best.rank = 0u; best.leaf = INVALID_U32; best.flags = 0u;
                        second.leaf = INVALID_U32;
                        uncertified = false; done = false;
                        G = make_uint2(0u, 0x80000000u); T = make_uint2(0u, 0u); sp = 0;
                        hasRay = true;
                    }
                }
                if(base + uint32_t(__popc(need)) >= rayCount) exhausted = true;
            }
        }
        uint32_t live = __ballot_sync(FULL, hasRay);
        if(live == 0u) break;
        const int fetchThr = exhausted ? 1 : int(prm.fetchThr);

        // -------------------------------------- traversal --------------------------------------
        do
        {
            const bool triWork = hasRay && (T.y != 0u);
            const uint32_t bT = __ballot_sync(FULL, triWork);
            const uint32_t bN0 = __ballot_sync(FULL, hasRay && ((G.y & 0xFF000000u) != 0u));
#ifdef MRB_TRACE_STATS
            if(lane == 0) { atomicAdd(counters + 12, 1u); atomicAdd(counters + 13, uint32_t(__popc(live))); }
#endif

            // ---- triangle phase: one triangle per lane ----
            if(bT != 0u && (bN0 == 0u || uint32_t(__popc(bT)) * prm.triDiv >= uint32_t(__popc(live))))
            {
#ifdef MRB_TRACE_STATS
                if(lane == 0) { atomicAdd(counters + 10, 1u); atomicAdd(counters + 11, uint32_t(__popc(bT))); }
#endif
                if(triWork)
                {
                    const uint32_t tb = uint32_t(__ffs(int(T.y))) - 1u;
                    T.y &= T.y - 1u;
                    const float4* tp = triBase + size_t(T.x + tb) * 3;
                    const float4 v0 = __ldg(tp + 0), e0 = __ldg(tp + 1), e1 = __ldg(tp + 2);
                    const uint32_t flags = __float_as_uint(e1.w);
                    float t, u, v;
                    if(RayTriangle(ox, oy, oz, dx, dy, dz, v0, e0, e1, (flags & 1u) != 0u, t, u, v) &&
                       (t >= tMin && t < tMaxOrig) && // IsInRange (AcceleratorLBVH.hpp:L233-236)
                       AlphaGate<ALPHA>(a.ranges, a.indices, flags, __float_as_uint(v0.w), u, v, rays + r, prm.alphaSeed))
                    {
                        const uint32_t rank = __float_as_uint(e0.w);
                        if(ANY_HIT)
                        {
                            if(CertifyLeaf(a, __float_as_uint(v0.w), ox, oy, oz, dx, dy, dz, tMin, tMaxOrig))
                            { done = true; hasRay = false; finished = true; }
                            else uncertified = true;
                        }
                        else if(t < best.t || (t == best.t && rank < best.rank))
                        {
                            if(best.leaf != INVALID_U32)
                            {
                                const float win = t * NEAR_TIE;
                                const bool bestIn = best.t <= win;
                                if(bestIn && second.leaf != INVALID_U32 && second.t <= win) overflow = true;
                                if(bestIn) second = best; else second.leaf = INVALID_U32;
                            }
                            best.t = t; best.u = u; best.v = v; best.rank = rank;
                            best.leaf = __float_as_uint(v0.w); best.flags = flags;
                            tMax = fminf(tMaxOrig, t * NEAR_TIE);
                        }
                        else if(t <= best.t * NEAR_TIE)
                        {
                            if(second.leaf != INVALID_U32) overflow = true;
                            else
                            {
                                second.t = t; second.u = u; second.v = v; second.rank = rank;
                                second.leaf = __float_as_uint(v0.w); second.flags = flags;
                            }
                        }
                    }
                }
            }
            // lanes whose triangle group just ran dry and that hold no node group take their next stack entry
            // now, so that they can join this iteration's node phase instead of idling through it
            if(hasRay && (G.y & 0xFF000000u) == 0u && T.y == 0u && sp > 0)
            {
                const uint2 e = stack[--sp];
                if(e.y & 0xFF000000u) G = e; else T = e;
            }
            const bool nodeWork = hasRay && ((G.y & 0xFF000000u) != 0u);
            const uint32_t bN = __ballot_sync(FULL, nodeWork);
            // ---- node phase: one node per lane ----
            if(bN != 0u)
            {
#ifdef MRB_TRACE_STATS
                if(lane == 0) { atomicAdd(counters + 8, 1u); atomicAdd(counters + 9, uint32_t(__popc(bN))); }
#endif
                if(nodeWork && hasRay)
                {
                    if(T.y != 0u) { stack[sp++] = T; T.y = 0u; } // postponed triangle group
                    const uint32_t hitsImask = G.y;
                    const uint32_t bit = 31u - uint32_t(__clz(int(hitsImask)));
                    G.y &= ~(1u << bit);
                    if(G.y & 0xFF000000u) { stack[sp++] = G; }
                    const uint32_t slot = (bit - 24u) ^ oct;
                    const uint32_t rel = __popc(hitsImask & ~(0xFFFFFFFFu << slot));
                    const uint4* np = nodeBase + size_t(G.x + rel) * 5;
                    const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);

                    const float sx = __uint_as_float((n0.w & 0xFFu) << 23) * idx;
                    const float sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idy;
                    const float sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idz;
                    const float cx = (__uint_as_float(n0.x) - ox) * idx;
                    const float cy = (__uint_as_float(n0.y) - oy) * idy;
                    const float cz = (__uint_as_float(n0.z) - oz) * idz;
                    // per-axis slack covering the rounding of this formulation and of the reference's
                    // slab test: 2^-20 (|c| + 255 |s|) + 2^-8 |s|. It must stay per axis: a ray almost
                    // parallel to one axis has huge |s|, |c| there, which says nothing about the others.
                    const float kx = __fmaf_rn(0.0042f, fabsf(sx), 9.5367431640625e-07f * fabsf(cx));
                    const float ky = __fmaf_rn(0.0042f, fabsf(sy), 9.5367431640625e-07f * fabsf(cy));
                    const float kz = __fmaf_rn(0.0042f, fabsf(sz), 9.5367431640625e-07f * fabsf(cz));
                    const float bx = __fmaf_rn(-32768.0f, sx, cx), by = __fmaf_rn(-32768.0f, sy, cy), bz = __fmaf_rn(-32768.0f, sz, cz);
                    const float cnx = bx - kx, cny = by - ky, cnz = bz - kz;
                    const float cfx = bx + kx, cfy = by + ky, cfz = bz + kz;
                    const float tFarLimit = tMax;

                    uint32_t hitmask = 0u;
                    const uint32_t oct4 = oct * 0x01010101u;
                    const uint32_t magic = prm.magic; // 0x47000000, kept opaque so ptxas keeps the PRMT selectors immediate
                    #pragma unroll
                    for(int half = 0; half < 2; half++)
                    {
                        const uint32_t meta4 = half ? n1.w : n1.z;
                        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                        const uint32_t innerMask4 = (isInner4 >> 4) * 0xFFu;
                        const uint32_t bitIndex4 = (meta4 ^ (oct4 & innerMask4)) & 0x1F1F1F1Fu;
                        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
                        const uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z, qlz = half ? n3.y : n3.x;
                        const uint32_t qhx = half ? n3.w : n3.z, qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
                        const uint32_t nx = (dx < 0.0f) ? qhx : qlx, fx = (dx < 0.0f) ? qlx : qhx;
                        const uint32_t ny = (dy < 0.0f) ? qhy : qly, fy = (dy < 0.0f) ? qly : qhy;
                        const uint32_t nz = (dz < 0.0f) ? qhz : qlz, fz = (dz < 0.0f) ? qlz : qhz;
                        #pragma unroll
                        for(int j = 0; j < 4; j++)
                        {
                            const float tnx = __fmaf_rn(QF(nx, j, magic), sx, cnx), tfx = __fmaf_rn(QF(fx, j, magic), sx, cfx);
                            const float tny = __fmaf_rn(QF(ny, j, magic), sy, cny), tfy = __fmaf_rn(QF(fy, j, magic), sy, cfy);
                            const float tnz = __fmaf_rn(QF(nz, j, magic), sz, cnz), tfz = __fmaf_rn(QF(fz, j, magic), sz, cfz);
                            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tMin));
                            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tFarLimit));
                            if(tn <= tf)
                            {
                                const uint32_t cbits = (childBits4 >> (8 * j)) & 0xFFu;
                                const uint32_t bidx = (bitIndex4 >> (8 * j)) & 0xFFu;
                                hitmask |= cbits << bidx;
                            }
                        }
                    }
                    G.x = n1.x;
                    G.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
                    T.x = n1.y;
                    T.y = hitmask & 0x00FFFFFFu;
                }
            }
            // ---- pop / finish ----
            if(hasRay && (G.y & 0xFF000000u) == 0u && T.y == 0u)
            {
                if(sp == 0) { hasRay = false; finished = true; }
                else
                {
                    const uint2 e = stack[--sp];
                    if(e.y & 0xFF000000u) G = e; else T = e;
                }
            }
            live = __ballot_sync(FULL, hasRay);
        } while(__popc(live) >= fetchThr);
    }
}

// ------------------------------------------------------------------------------------------------
// Two-level scenes
// ------------------------------------------------------------------------------------------------
// Matrix3x4::TransformRay (Core/Matrix.hpp:L905-912) with Math::Dot's FMA chains: dir' = M dir,
// pos' = M (pos, 1); no renormalisation.
__device__ __forceinline__ void TransformRayExact(const float* __restrict__ m, float ox, float oy, float oz,
                                                  float dx, float dy, float dz, float lo[3], float ld[3])
{
    #pragma unroll
    for(int r = 0; r < 3; r++)
    {
        float d = __fmaf_rn(m[4 * r + 0], dx, 0.0f);
        d = __fmaf_rn(m[4 * r + 1], dy, d);
        d = __fmaf_rn(m[4 * r + 2], dz, d);
        ld[r] = d;
        float p = __fmaf_rn(m[4 * r + 0], ox, 0.0f);
        p = __fmaf_rn(m[4 * r + 1], oy, p);
        p = __fmaf_rn(m[4 * r + 2], oz, p);
        p = __fmaf_rn(m[4 * r + 3], 1.0f, p);
        lo[r] = p;
    }
}

__device__ __forceinline__ void WriteHit2(const InstanceRec& in, uint32_t r, const HitRecord& h,
                                          mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays)
{
    uint32_t ri = h.flags >> 8;
    uint32_t prim = in.ranges.primBegin[ri] + (h.leaf - in.ranges.leafStart[ri]);
    uint4 keys = make_uint4((in.ranges.primGroupId << 28) | prim, in.ranges.lmKey[ri], in.transKey, in.accelKey);
    *reinterpret_cast<uint4*>(hitKeys + r) = keys;
    float w = __fsub_rn(__fsub_rn(1.0f, h.u), h.v);
    *reinterpret_cast<float2*>(metaHits + r) = make_float2(w, h.u);
    rays[r].tMax = h.t;
}

// Reference reaches (instance, leaf) if its slab test passes on the instance's world AABB with the world
// ray (every top-level ancestor encloses it) and on the leaf AABB with the local ray.
__device__ __forceinline__ bool CertifyLeaf2(const InstanceRec& in, uint32_t leaf, const float wo[3], const float wd[3],
                                             float tMin, float tUpper)
{
    const float winv[3] = {__fdiv_rn(1.0f, wd[0]), __fdiv_rn(1.0f, wd[1]), __fdiv_rn(1.0f, wd[2])};
    if(!SlabExact(in.worldAABB, wo, winv, tMin, tUpper)) return false;
    float lo[3] = {wo[0], wo[1], wo[2]}, ld[3] = {wd[0], wd[1], wd[2]};
    if(!in.identity) TransformRayExact(in.invTransform, wo[0], wo[1], wo[2], wd[0], wd[1], wd[2], lo, ld);
    const float linv[3] = {__fdiv_rn(1.0f, ld[0]), __fdiv_rn(1.0f, ld[1]), __fdiv_rn(1.0f, ld[2])};
    const float* b = in.leafAABB + 6 * size_t(leaf);
    const float box[6] = {b[0], b[1], b[2], b[3], b[4], b[5]};
    return SlabExact(box, lo, linv, tMin, tUpper);
}

// Exact two-level traversal of ONE ray (audit path and fallback): KCIntersectBaseLBVH semantics on the top level
// (internal boxes and the instance leaf AABB slab-tested with the current tMax, left-first), the
// bottom-level ClosestHit / FirstHit in local space, tMax shrinking across instances.
template<bool ANY_HIT, bool ALPHA>
__device__ __forceinline__ void TraceBinary2RayBody(const SceneData& sc,
                                                    mrb_hit_key_pack* __restrict__ hitKeys, mrb_meta_hit* __restrict__ metaHits,
                                                    uint32_t* __restrict__ visibleBits, mrb_ray_gmem* rays, uint32_t r, uint32_t alphaSeed)
{
    const float4 r0 = reinterpret_cast<const float4*>(rays + r)[0];
    const float4 r1 = reinterpret_cast<const float4*>(rays + r)[1];
    const float wo[3] = {r0.x, r0.y, r0.z}, wd[3] = {r1.x, r1.y, r1.z};
    const float winv[3] = {__fdiv_rn(1.0f, wd[0]), __fdiv_rn(1.0f, wd[1]), __fdiv_rn(1.0f, wd[2])};
    const float tMin = r0.w; float tMax = r1.w;
    HitRecord best; best.t = tMax; best.rank = 0u; best.leaf = INVALID_U32; best.u = best.v = 0.f; best.flags = 0u;
    uint32_t bestInst = 0; bool done = false;
    uint32_t tstack[96]; int tsp = 0;
    tstack[tsp++] = 0u;
    while(tsp > 0 && !done)
    {
        uint32_t tn = tstack[--tsp];
        if(tn == INVALID_U32) continue;
        if(!(tn & LEAF_FLAG))
        {
            if(SlabExact(reinterpret_cast<const float*>(sc.tlas.boxes + tn), wo, winv, tMin, tMax))
            {
                LBVHNode nd = sc.tlas.nodes[tn];
                tstack[tsp++] = nd.right; tstack[tsp++] = nd.left;
            }
            continue;
        }
        const uint32_t ii = tn & ~LEAF_FLAG;
        const InstanceRec& in = sc.instances[ii];
        if(!SlabExact(in.worldAABB, wo, winv, tMin, tMax)) continue;
        float o[3] = {wo[0], wo[1], wo[2]}, d[3] = {wd[0], wd[1], wd[2]};
        if(!in.identity) TransformRayExact(in.invTransform, wo[0], wo[1], wo[2], wd[0], wd[1], wd[2], o, d);
        const float invD[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
        uint32_t stack[128]; int sp = 0;
        stack[sp++] = 0u;
        while(sp > 0)
        {
            uint32_t ni = stack[--sp];
            if(ni == INVALID_U32) continue;
            if(ni & LEAF_FLAG)
            {
                uint32_t leaf = ni & ~LEAF_FLAG;
                const uint32_t ri = (in.ranges.count == 1u) ? 0u : FindRange(in.ranges, leaf);
                uint32_t prim = in.ranges.primBegin[ri] + (leaf - in.ranges.leafStart[ri]);
                uint32_t i0 = in.indices[3 * size_t(prim)], i1 = in.indices[3 * size_t(prim) + 1], i2 = in.indices[3 * size_t(prim) + 2];
                const float* p0 = in.positions + 3 * size_t(i0);
                const float* p1 = in.positions + 3 * size_t(i1);
                const float* p2 = in.positions + 3 * size_t(i2);
                float4 v0 = make_float4(p0[0], p0[1], p0[2], 0.f);
                float4 e0 = make_float4(__fsub_rn(p1[0], p0[0]), __fsub_rn(p1[1], p0[1]), __fsub_rn(p1[2], p0[2]), 0.f);
                float4 e1 = make_float4(__fsub_rn(p2[0], p0[0]), __fsub_rn(p2[1], p0[1]), __fsub_rn(p2[2], p0[2]), 0.f);
                float t, u, v;
                if(!RayTriangle(o[0], o[1], o[2], d[0], d[1], d[2], v0, e0, e1, in.ranges.cull[ri] != 0u, t, u, v)) continue;
                if(!(t >= tMin && t < tMax)) continue;
                // (the instance index joins the seed: instances of one accelerator decide independently)
                if(ALPHA && in.ranges.alphaMap && in.ranges.alphaMap[ri] >= 0 && !AlphaKeepsHit(in.ranges, in.indices, leaf, ri, u, v, rays + r, alphaSeed ^ (ii * 0xC2B2AE35u))) continue;
                best.t = t; best.u = u; best.v = v; best.leaf = leaf; best.flags = ri << 8; bestInst = ii;
                tMax = t;
                if(ANY_HIT) { done = true; break; }
            }
            else if(SlabExact(reinterpret_cast<const float*>(in.boxes + ni), o, invD, tMin, tMax))
            {
                LBVHNode nd = in.nodes[ni];
                stack[sp++] = nd.right; stack[sp++] = nd.left;
            }
        }
    }
    if(best.leaf != INVALID_U32)
    {
        if(ANY_HIT) atomicAnd(&visibleBits[r >> 5], ~(1u << (r & 31u)));
        else WriteHit2(sc.instances[bestInst], r, best, hitKeys, metaHits, rays);
    }
}
__device__ __noinline__ void TraceBinary2RayClosest(const SceneData& sc, mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays, uint32_t r)
{ TraceBinary2RayBody<false, false>(sc, hitKeys, metaHits, nullptr, rays, r, 0u); }
__device__ __noinline__ void TraceBinary2RayAny(const SceneData& sc, uint32_t* visibleBits, mrb_ray_gmem* rays, uint32_t r)
{ TraceBinary2RayBody<true, false>(sc, nullptr, nullptr, visibleBits, rays, r, 0u); }
__device__ __noinline__ void TraceBinary2RayClosestAlpha(const SceneData& sc, mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, mrb_ray_gmem* rays, uint32_t r, uint32_t alphaSeed)
{ TraceBinary2RayBody<false, true>(sc, hitKeys, metaHits, nullptr, rays, r, alphaSeed); }
__device__ __noinline__ void TraceBinary2RayAnyAlpha(const SceneData& sc, uint32_t* visibleBits, mrb_ray_gmem* rays, uint32_t r, uint32_t alphaSeed)
{ TraceBinary2RayBody<true, true>(sc, nullptr, nullptr, visibleBits, rays, r, alphaSeed); }

template<bool ANY_HIT>
__global__ void __launch_bounds__(TRACE_TPB)
KTraceBinary2(const __grid_constant__ SceneData sc,
              mrb_hit_key_pack* __restrict__ hitKeys, mrb_meta_hit* __restrict__ metaHits,
              uint32_t* __restrict__ visibleBits,
              mrb_ray_gmem* rays, const uint32_t* __restrict__ rayIndices, uint32_t rayCount, uint32_t alphaSeed)
{
    for(uint32_t i = blockIdx.x * TRACE_TPB + threadIdx.x; i < rayCount; i += gridDim.x * TRACE_TPB)
        TraceBinary2RayBody<ANY_HIT, true>(sc, hitKeys, metaHits, visibleBits, rays, rayIndices ? rayIndices[i] : i, alphaSeed);
}

// Same phase-uniform persistent loop as KTraceWide with one more kind of leaf: in the top-level tree a
// leaf record is an instance; entering it pushes the pending top-level groups and a sentinel, switches the
// lane to the instance's local ray / node arrays, and the sentinel pop switches back.
template<bool ANY_HIT, bool ALPHA>
__global__ void __launch_bounds__(TRACE_TPB)
KTraceWide2(const __grid_constant__ SceneData sc,
            mrb_hit_key_pack* __restrict__ hitKeys, mrb_meta_hit* __restrict__ metaHits,
            uint32_t* __restrict__ visibleBits,
            mrb_ray_gmem* rays, const uint32_t* __restrict__ rayIndices, uint32_t rayCount,
            uint32_t* __restrict__ counters, TraceParams prm)
{
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint4* const tlasNodes = reinterpret_cast<const uint4*>(sc.tlas.wideNodes);

    bool hasRay = false, finished = false, exhausted = false;
    uint32_t r = 0, level = 0, inst = 0;
    const uint4* nodeBase = tlasNodes;
    const float4* triBase = nullptr;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, idx = 0, idy = 0, idz = 0;
    float tMin = 0, tMaxOrig = 0, tMax = 0;
    bool overflow = false, uncertified = false, done = false;
    uint32_t oct = 0, bestInst = 0, secondSeen = 0;
    HitRecord best; best.t = 0; best.u = best.v = 0; best.leaf = INVALID_U32; best.rank = 0; best.flags = 0;
    uint2 G = make_uint2(0u, 0u), T = make_uint2(0u, 0u);
    uint2 stack[WIDE_STACK];
    int sp = 0;

    auto SetRay = [&](float nox, float noy, float noz, float ndx, float ndy, float ndz)
    {
        ox = nox; oy = noy; oz = noz; dx = ndx; dy = ndy; dz = ndz;
        idx = (fabsf(dx) > 1e-30f) ? 1.0f / dx : copysignf(1e30f, dx);
        idy = (fabsf(dy) > 1e-30f) ? 1.0f / dy : copysignf(1e30f, dy);
        idz = (fabsf(dz) > 1e-30f) ? 1.0f / dz : copysignf(1e30f, dz);
        oct = (dx < 0.0f ? 4u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 1u : 0u);
    };

    while(true)
    {
        if(finished)
        {
            finished = false;
            const float4 w0 = reinterpret_cast<const float4*>(rays + r)[0];
            const float4 w1 = reinterpret_cast<const float4*>(rays + r)[1];
            const float wo[3] = {w0.x, w0.y, w0.z}, wd[3] = {w1.x, w1.y, w1.z};
            if(ANY_HIT)
            {
                if(done) atomicAnd(&visibleBits[r >> 5], ~(1u << (r & 31u)));
                else if(uncertified)
                {
                    atomicAdd(counters, 1u); atomicAdd(counters + 2, 1u); atomicAdd(counters + 4, 1u);
                    if constexpr(ALPHA) TraceBinary2RayAnyAlpha(sc, visibleBits, rays, r, prm.alphaSeed);
                    else TraceBinary2RayAny(sc, visibleBits, rays, r);
                }
            }
            else if(best.leaf != INVALID_U32)
            {
                const float tUpper = fminf(tMaxOrig, best.t * NEAR_TIE);
                const InstanceRec& in = sc.instances[bestInst];
                if(!secondSeen && !overflow && CertifyLeaf2(in, best.leaf, wo, wd, tMin, tUpper))
                    WriteHit2(in, r, best, hitKeys, metaHits, rays);
                else
                {
                    // near tie or uncertified leaf: the reference's two-level algorithm decides (out of line, rare)
                    atomicAdd(counters, 1u); atomicAdd(counters + 4, 1u);
                    atomicAdd(counters + ((secondSeen || overflow) ? 1 : 2), 1u);
                    if constexpr(ALPHA) TraceBinary2RayClosestAlpha(sc, hitKeys, metaHits, rays, r, prm.alphaSeed);
                    else TraceBinary2RayClosest(sc, hitKeys, metaHits, rays, r);
                }
            }
        }
        if(!exhausted)
        {
            const uint32_t need = __ballot_sync(FULL, !hasRay);
            if(need)
            {
                const int leader = __ffs(int(need)) - 1;
                uint32_t base = 0;
                if(int(lane) == leader) base = atomicAdd(counters + 3, uint32_t(__popc(need)));
                base = __shfl_sync(FULL, base, leader);
                if(!hasRay)
                {
                    const uint32_t i = base + uint32_t(__popc(need & ltMask));
                    if(i < rayCount)
                    {
                        r = rayIndices ? rayIndices[i] : i;
                        const float4 r0 = reinterpret_cast<const float4*>(rays + r)[0];
                        const float4 r1 = reinterpret_cast<const float4*>(rays + r)[1];
                        SetRay(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                        tMin = r0.w; tMaxOrig = r1.w; tMax = tMaxOrig; overflow = false; secondSeen = 0u;
                        best.t = tMaxOrig; best.rank = 0u; best.leaf = INVALID_U32; best.flags = 0u; bestInst = 0u;
                        uncertified = false; done = false;
                        G = make_uint2(0u, 0x80000000u); T = make_uint2(0u, 0u); sp = 0;
                        level = 0u; nodeBase = tlasNodes; triBase = nullptr;
                        hasRay = true;
                    }
                }
                if(base + uint32_t(__popc(need)) >= rayCount) exhausted = true;
            }
        }
        uint32_t live = __ballot_sync(FULL, hasRay);
        if(live == 0u) break;
        const int fetchThr = exhausted ? 1 : int(prm.fetchThr);
        do
        {
            const bool triWork = hasRay && (T.y != 0u);
            const uint32_t bT = __ballot_sync(FULL, triWork);
            const uint32_t bN0 = __ballot_sync(FULL, hasRay && ((G.y & 0xFF000000u) != 0u));
            if(bT != 0u && (bN0 == 0u || uint32_t(__popc(bT)) * prm.triDiv >= uint32_t(__popc(live))))
            {
                if(triWork)
                {
                    const uint32_t tb = uint32_t(__ffs(int(T.y))) - 1u;
                    T.y &= T.y - 1u;
                    if(level == 0u)
                    {
                        // ---- enter an instance ----
                        const uint32_t ii = sc.tlas.leafOfSlot[T.x + tb];
                        const InstanceRec* in = sc.instances + ii;
                        if(G.y & 0xFF000000u) stack[sp++] = G;
                        if(T.y != 0u) stack[sp++] = T;
                        stack[sp++] = make_uint2(0xFFFFFFFFu, 0u); // sentinel: back to the top level
                        if(!in->identity)
                        {
                            float lo[3], ld[3];
                            TransformRayExact(in->invTransform, ox, oy, oz, dx, dy, dz, lo, ld);
                            SetRay(lo[0], lo[1], lo[2], ld[0], ld[1], ld[2]);
                        }
                        nodeBase = reinterpret_cast<const uint4*>(in->wideNodes);
                        triBase = reinterpret_cast<const float4*>(in->tris);
                        inst = ii; level = 1u;
                        G = make_uint2(0u, 0x80000000u); T = make_uint2(0u, 0u);
                    }
                    else
                    {
                        const float4* tp = triBase + size_t(T.x + tb) * 3;
                        const float4 v0 = __ldg(tp + 0), e0 = __ldg(tp + 1), e1 = __ldg(tp + 2);
                        const uint32_t flags = __float_as_uint(e1.w);
                        float t, u, v;
                        if(RayTriangle(ox, oy, oz, dx, dy, dz, v0, e0, e1, (flags & 1u) != 0u, t, u, v) &&
                           (t >= tMin && t < tMaxOrig) &&
                           AlphaGate<ALPHA>(sc.instances[inst].ranges, sc.instances[inst].indices, flags, __float_as_uint(v0.w), u, v, rays + r,
                                            prm.alphaSeed ^ (inst * 0xC2B2AE35u)))
                        {
                            const uint32_t rank = __float_as_uint(e0.w);
                            if(ANY_HIT)
                            {
                                const float4 w0 = reinterpret_cast<const float4*>(rays + r)[0];
                                const float4 w1 = reinterpret_cast<const float4*>(rays + r)[1];
                                const float wo[3] = {w0.x, w0.y, w0.z}, wd[3] = {w1.x, w1.y, w1.z};
                                if(CertifyLeaf2(sc.instances[inst], __float_as_uint(v0.w), wo, wd, tMin, tMaxOrig))
                                { done = true; hasRay = false; finished = true; }
                                else uncertified = true;
                            }
                            else if(t < best.t)
                            {
                                // any earlier candidate inside the new window makes the ray a near tie
                                if(best.leaf != INVALID_U32 && best.t <= t * NEAR_TIE) secondSeen = 1u;
                                best.t = t; best.u = u; best.v = v; best.rank = rank;
                                best.leaf = __float_as_uint(v0.w); best.flags = flags; bestInst = inst;
                                tMax = fminf(tMaxOrig, t * NEAR_TIE);
                            }
                            else if(t <= best.t * NEAR_TIE) secondSeen = 1u;
                        }
                    }
                }
            }
            // as in KTraceWide: lanes that just ran out of triangles pop now (unless the next entry is the
            // level sentinel, which the regular pop below handles) and join this iteration's node phase
            if(hasRay && (G.y & 0xFF000000u) == 0u && T.y == 0u && sp > 0)
            {
                const uint2 e = stack[sp - 1];
                if(e.y != 0u) { sp--; if(e.y & 0xFF000000u) G = e; else T = e; }
            }
            const bool nodeWork = hasRay && ((G.y & 0xFF000000u) != 0u);
            const uint32_t bN = __ballot_sync(FULL, nodeWork);
            if(bN != 0u)
            {
                if(nodeWork && hasRay && ((G.y & 0xFF000000u) != 0u))
                {
                    if(T.y != 0u) { stack[sp++] = T; T.y = 0u; }
                    const uint32_t hitsImask = G.y;
                    const uint32_t bit = 31u - uint32_t(__clz(int(hitsImask)));
                    G.y &= ~(1u << bit);
                    if(G.y & 0xFF000000u) { stack[sp++] = G; }
                    const uint32_t slot = (bit - 24u) ^ oct;
                    const uint32_t rel = __popc(hitsImask & ~(0xFFFFFFFFu << slot));
                    const uint4* np = nodeBase + size_t(G.x + rel) * 5;
                    const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
                    const float sx = __uint_as_float((n0.w & 0xFFu) << 23) * idx;
                    const float sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idy;
                    const float sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idz;
                    const float cx = (__uint_as_float(n0.x) - ox) * idx;
                    const float cy = (__uint_as_float(n0.y) - oy) * idy;
                    const float cz = (__uint_as_float(n0.z) - oz) * idz;
                    const float kx = __fmaf_rn(0.0042f, fabsf(sx), 9.5367431640625e-07f * fabsf(cx));
                    const float ky = __fmaf_rn(0.0042f, fabsf(sy), 9.5367431640625e-07f * fabsf(cy));
                    const float kz = __fmaf_rn(0.0042f, fabsf(sz), 9.5367431640625e-07f * fabsf(cz));
                    const float bx = __fmaf_rn(-32768.0f, sx, cx), by = __fmaf_rn(-32768.0f, sy, cy), bz = __fmaf_rn(-32768.0f, sz, cz);
                    const float cnx = bx - kx, cny = by - ky, cnz = bz - kz;
                    const float cfx = bx + kx, cfy = by + ky, cfz = bz + kz;
                    const float tFarLimit = tMax;
                    uint32_t hitmask = 0u;
                    const uint32_t oct4 = oct * 0x01010101u;
                    const uint32_t magic = prm.magic;
                    #pragma unroll
                    for(int half = 0; half < 2; half++)
                    {
                        const uint32_t meta4 = half ? n1.w : n1.z;
                        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                        const uint32_t innerMask4 = (isInner4 >> 4) * 0xFFu;
                        const uint32_t bitIndex4 = (meta4 ^ (oct4 & innerMask4)) & 0x1F1F1F1Fu;
                        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
                        const uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z, qlz = half ? n3.y : n3.x;
                        const uint32_t qhx = half ? n3.w : n3.z, qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
                        const uint32_t nx = (dx < 0.0f) ? qhx : qlx, fx = (dx < 0.0f) ? qlx : qhx;
                        const uint32_t ny = (dy < 0.0f) ? qhy : qly, fy = (dy < 0.0f) ? qly : qhy;
                        const uint32_t nz = (dz < 0.0f) ? qhz : qlz, fz = (dz < 0.0f) ? qlz : qhz;
                        #pragma unroll
                        for(int j = 0; j < 4; j++)
                        {
                            const float tnx = __fmaf_rn(QF(nx, j, magic), sx, cnx), tfx = __fmaf_rn(QF(fx, j, magic), sx, cfx);
                            const float tny = __fmaf_rn(QF(ny, j, magic), sy, cny), tfy = __fmaf_rn(QF(fy, j, magic), sy, cfy);
                            const float tnz = __fmaf_rn(QF(nz, j, magic), sz, cnz), tfz = __fmaf_rn(QF(fz, j, magic), sz, cfz);
                            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tMin));
                            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tFarLimit));
                            if(tn <= tf)
                            {
                                const uint32_t cbits = (childBits4 >> (8 * j)) & 0xFFu;
                                const uint32_t bidx = (bitIndex4 >> (8 * j)) & 0xFFu;
                                hitmask |= cbits << bidx;
                            }
                        }
                    }
                    G.x = n1.x;
                    G.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
                    T.x = n1.y;
                    T.y = hitmask & 0x00FFFFFFu;
                }
            }
            if(hasRay && (G.y & 0xFF000000u) == 0u && T.y == 0u)
            {
                if(sp == 0) { hasRay = false; finished = true; }
                else
                {
                    const uint2 e = stack[--sp];
                    if(e.y == 0u)
                    {
                        // sentinel: leave the instance, restore the world ray
                        const float4 r0 = reinterpret_cast<const float4*>(rays + r)[0];
                        const float4 r1 = reinterpret_cast<const float4*>(rays + r)[1];
                        SetRay(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                        level = 0u; nodeBase = tlasNodes;
                    }
                    else if(e.y & 0xFF000000u) G = e; else T = e;
                }
            }
            live = __ballot_sync(FULL, hasRay);
        } while(__popc(live) >= fetchThr);
    }
}

// Scheduling knobs of the wide kernels, shared by single-accelerator and scene casts (MRB_TRI_DIV / MRB_FETCH_THR
// override them for parameter sweeps)
const TraceParams& WideTraceParams(bool anyHit = false)
{
    static const TraceParams prm = []
    {
        TraceParams p{8u, 24u, 0x47000000u, 0u};
        if(const char* e = getenv("MRB_TRI_DIV")) p.triDiv = uint32_t(atoi(e));
        if(const char* e = getenv("MRB_FETCH_THR")) p.fetchThr = uint32_t(atoi(e));
        return p;
    }();
    // the any-hit kernel may take its own knobs (MRB_TRI_DIV_ANY / MRB_FETCH_THR_ANY); unset = the closest-hit kernel's
    static const TraceParams prmAny = []
    {
        TraceParams p = prm;
        p.fetchThr = 28u;   // shadow rays end at their first hit, so lanes free up faster: refilling a little earlier pays (PT 6.72 -> 6.67 ms/spp)
        if(const char* e = getenv("MRB_TRI_DIV_ANY")) p.triDiv = uint32_t(atoi(e));
        if(const char* e = getenv("MRB_FETCH_THR_ANY")) p.fetchThr = uint32_t(atoi(e));
        return p;
    }();
    return anyHit ? prmAny : prm;
}

} // namespace

// L2 residency hint: nodes + triangle records of the accelerator being traced (20 MB at 264 K triangles) are marked
// persisting in the stream's access-policy window, so the ~130 MB ray / hit stream of a cast cannot evict them.
static void SetPersistingWindow(Context& ctx, const void* base, size_t bytes)
{
    if(ctx.persistBase == base && ctx.persistStream == ctx.stream) return;
    static const bool enabled = []{ const char* e = getenv("MRB_L2_PERSIST"); return !(e && e[0] == '0'); }();
    if(!enabled) return;
    int maxWindow = 0, maxPersist = 0;
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, ctx.device);
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx.device);
    ctx.persistBase = base; ctx.persistStream = ctx.stream;
    if(maxWindow <= 0 || maxPersist <= 0) return;
    const size_t win = bytes < size_t(maxWindow) ? bytes : size_t(maxWindow);
    const size_t carve = win < size_t(maxPersist) ? win : size_t(maxPersist);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    cudaStreamAttrValue attr = {};
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : float(double(carve) / double(win));
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(ctx.stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();   // a hint: never fatal
}

void TraceRays(Context& ctx, const mrb_accel_t& acc, bool anyHit, mrb_trace_mode mode,
               mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
               mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount)
{
    if(rayCount == 0) return;
    const uint32_t alphaSeed = ctx.alphaSeed++;   // one seed per cast: consecutive casts decide independently
    const uint32_t grid = DivUp(rayCount, TRACE_TPB);
    if(mode == MRB_TRACE_WIDE)
    {
        // counters of the cast: [0] uncertified rays, [1] near ties, [2] uncertified leaves, [3] fetch cursor,
        // [4] rays re-traced by the reference's binary algorithm
        ctx.traceScratch.Reserve(sizeof(uint32_t) * 64);
        uint32_t* counters = static_cast<uint32_t*>(ctx.traceScratch.Base());
#ifdef MRB_TRACE_STATS
        MRB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 16, ctx.stream));
#else
        MRB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 8, ctx.stream));
#endif
        const bool alpha = acc.d.ranges.alphaMap != nullptr;
        if(!ctx.occWide[0])   // per context: a process may drive several devices
        {
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide[0], KTraceWide<false, false>, TRACE_TPB, 0));
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide[1], KTraceWide<true, false>, TRACE_TPB, 0));
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide[2], KTraceWide<false, true>, TRACE_TPB, 0));
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide[3], KTraceWide<true, true>, TRACE_TPB, 0));
        }
        if(acc.d.wideNodes && acc.d.tris)
            SetPersistingWindow(ctx, acc.d.wideNodes, size_t(reinterpret_cast<const char*>(acc.d.tris + acc.d.leafCount) - reinterpret_cast<const char*>(acc.d.wideNodes)));
        TraceParams prm = WideTraceParams(anyHit);
        prm.alphaSeed = alphaSeed;
        const uint32_t pgrid = min(grid, uint32_t(ctx.smCount) * uint32_t(ctx.occWide[(anyHit ? 1 : 0) + (alpha ? 2 : 0)]));
        {
            ProfileScope ps(ctx, anyHit ? PROF_TRACE_ANY : PROF_TRACE_CLOSEST);
            if(alpha)
            {
                if(anyHit) MRB_LAUNCH(ctx, (KTraceWide<true, true>), pgrid, TRACE_TPB, 0, acc.d, acc.accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
                else       MRB_LAUNCH(ctx, (KTraceWide<false, true>), pgrid, TRACE_TPB, 0, acc.d, acc.accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
            }
            else if(anyHit) MRB_LAUNCH(ctx, (KTraceWide<true, false>), pgrid, TRACE_TPB, 0, acc.d, acc.accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
            else            MRB_LAUNCH(ctx, (KTraceWide<false, false>), pgrid, TRACE_TPB, 0, acc.d, acc.accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
        }
        ctx.lastFallbackCount = counters;
#ifdef MRB_TRACE_STATS
        {   // diagnostic build only: warp-step statistics of this cast
            uint32_t h[16];
            MRB_CUDA_TRY(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
            MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
            fprintf(stderr, "[trace stats] rays %u anyHit %d | loop iters %u (live lanes %.1f) | node steps %u (lanes %.1f, per ray %.2f) | tri steps %u (lanes %.1f, per ray %.2f)\n",
                    rayCount, int(anyHit), h[12], double(h[13]) / double(h[12] ? h[12] : 1), h[8], double(h[9]) / double(h[8] ? h[8] : 1), double(h[9]) / rayCount,
                    h[10], double(h[11]) / double(h[10] ? h[10] : 1), double(h[11]) / rayCount);
        }
#endif
    }
    else
    {
        const uint32_t bgrid = min(grid, uint32_t(ctx.smCount) * 16u);
        if(anyHit) MRB_LAUNCH(ctx, KTraceBinary<true>, bgrid, TRACE_TPB, 0, acc.d, acc.accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, alphaSeed);
        else       MRB_LAUNCH(ctx, KTraceBinary<false>, bgrid, TRACE_TPB, 0, acc.d, acc.accelKey, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, alphaSeed);
        ctx.lastFallbackCount = nullptr;
    }
}


void TraceScene(Context& ctx, const SceneData& scnData, bool anyHit, mrb_trace_mode mode,
                mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
                mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount)
{
    if(rayCount == 0) return;
    const uint32_t alphaSeed = ctx.alphaSeed++;
    const uint32_t grid = DivUp(rayCount, TRACE_TPB);
    if(mode == MRB_TRACE_WIDE)
    {
        ctx.traceScratch.Reserve(sizeof(uint32_t) * 64);
        uint32_t* counters = static_cast<uint32_t*>(ctx.traceScratch.Base());
        MRB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 8, ctx.stream));
        const bool alpha = scnData.hasAlpha != 0u;
        if(!ctx.occWide2[0])
        {
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide2[0], KTraceWide2<false, false>, TRACE_TPB, 0));
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide2[1], KTraceWide2<true, false>, TRACE_TPB, 0));
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide2[2], KTraceWide2<false, true>, TRACE_TPB, 0));
            MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx.occWide2[3], KTraceWide2<true, true>, TRACE_TPB, 0));
        }
        TraceParams prm = WideTraceParams(anyHit);
        prm.alphaSeed = alphaSeed;
        const uint32_t pgrid = min(grid, uint32_t(ctx.smCount) * uint32_t(ctx.occWide2[(anyHit ? 1 : 0) + (alpha ? 2 : 0)]));
        {
            ProfileScope ps(ctx, anyHit ? PROF_TRACE_ANY : PROF_TRACE_CLOSEST);
            if(alpha)
            {
                if(anyHit) MRB_LAUNCH(ctx, (KTraceWide2<true, true>), pgrid, TRACE_TPB, 0, scnData, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
                else       MRB_LAUNCH(ctx, (KTraceWide2<false, true>), pgrid, TRACE_TPB, 0, scnData, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
            }
            else if(anyHit) MRB_LAUNCH(ctx, (KTraceWide2<true, false>), pgrid, TRACE_TPB, 0, scnData, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
            else            MRB_LAUNCH(ctx, (KTraceWide2<false, false>), pgrid, TRACE_TPB, 0, scnData, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, counters, prm);
        }
        ctx.lastFallbackCount = counters;
    }
    else
    {
        const uint32_t bgrid = min(grid, uint32_t(ctx.smCount) * 16u);
        if(anyHit) MRB_LAUNCH(ctx, KTraceBinary2<true>, bgrid, TRACE_TPB, 0, scnData, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, alphaSeed);
        else       MRB_LAUNCH(ctx, KTraceBinary2<false>, bgrid, TRACE_TPB, 0, scnData, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, alphaSeed);
        ctx.lastFallbackCount = nullptr;
    }
}

} // namespace mrb
