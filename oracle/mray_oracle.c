/* mray_oracle.c — CPU restatement of the reference's LBVH build + traversal hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (mray_b200/, include/) may link, load
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do,
 * and only as the checker / reported baseline.
 *
 * Parity status: PINNED. Every function here is checked (tests/test_oracle_vs_reference.py,
 * oracle/gen_golden.py) against the reference itself compiled from /root/reference into
 * oracle/_ref (libTracerDLL_CPU.so + libref_taps.so), and against the known-answer vectors of
 * the reference's own tests (Tests/Core/T_GraphicsFunctions.cpp:L182-307,
 * Tests/Device/T_AlgRadixSort.cu:L76-163). Golden outputs are committed under tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile). The reference's
 * Math::FMA is std::fma (Core/Math.h) so fmaf() below is a *true* fused multiply-add wherever
 * the reference fuses, and nowhere else.
 *
 * Each function cites the reference file:line it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define ORC_INVALID 0xFFFFFFFFu
#define ORC_LEAF_FLAG 0x80000000u /* ChildIndex = KeyT<u32,1,31>, IS_LEAF = 1 (AcceleratorLBVH.h:L46-49) */

/* ---------------------------------------------------------------------------------------------
 * Morton codes — Core/GraphicsFunctions.h:L586-625
 * ------------------------------------------------------------------------------------------- */
static uint64_t expand3d_64(uint32_t v)
{
    uint64_t x = v;
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x001f00000000ffffull;
    x = (x | x << 16) & 0x001f0000ff0000ffull;
    x = (x | x << 8)  & 0x100f00f00f00f00full;
    x = (x | x << 4)  & 0x10c30c30c30c30c3ull;
    x = (x | x << 2)  & 0x1249249249249249ull;
    return x;
}

uint64_t orc_morton_compose64(uint32_t x, uint32_t y, uint32_t z)
{
    return (expand3d_64(x) << 0) | (expand3d_64(y) << 1) | (expand3d_64(z) << 2);
}

static uint32_t expand3d_32(uint32_t x)
{
    x &= 0x000003ffu;
    x = (x ^ (x << 16)) & 0xff0000ffu;
    x = (x ^ (x << 8))  & 0x0300f00fu;
    x = (x ^ (x << 4))  & 0x030c30c3u;
    x = (x ^ (x << 2))  & 0x09249249u;
    return x;
}

uint32_t orc_morton_compose32(uint32_t x, uint32_t y, uint32_t z)
{
    return (expand3d_32(x) << 0) | (expand3d_32(y) << 1) | (expand3d_32(z) << 2);
}

/* ---------------------------------------------------------------------------------------------
 * Per-triangle AABB and centroid — Core/ShapeFunctions.h:L48-57,
 * Tracer/PrimitiveDefaultTriangle.hpp:L100-115 (centroid = p0*k + p1*k + p2*k, k = 0.333333333f,
 * summed left to right, no fusing).
 * ------------------------------------------------------------------------------------------- */
void orc_tri_aabb_center(const float* pos, const uint32_t* idx, uint32_t nTris,
                         float* outAABB, float* outCenter)
{
    const float k = 0.333333333f;
    for(uint32_t i = 0; i < nTris; i++)
    {
        const float* p0 = pos + 3 * (size_t)idx[3 * i + 0];
        const float* p1 = pos + 3 * (size_t)idx[3 * i + 1];
        const float* p2 = pos + 3 * (size_t)idx[3 * i + 2];
        for(int a = 0; a < 3; a++)
        {
            float mn = p0[a], mx = p0[a];
            mn = (p1[a] < mn) ? p1[a] : mn; /* std::min(a,b) = (b<a)?b:a */
            mn = (p2[a] < mn) ? p2[a] : mn;
            mx = (mx < p1[a]) ? p1[a] : mx; /* std::max(a,b) = (a<b)?b:a */
            mx = (mx < p2[a]) ? p2[a] : mx;
            outAABB[6 * i + a] = mn;
            outAABB[6 * i + 3 + a] = mx;
            float c = p0[a] * k;
            c = c + p1[a] * k;
            c = c + p2[a] * k;
            outCenter[3 * i + a] = c;
        }
    }
}

/* Union of leaf AABBs — UnionAABB3Functor seeded with AABB3::Negative()
 * (AcceleratorLBVH.hpp:L764-769, Core/AABB.hpp:L68-72,L140-144). */
void orc_aabb_union(const float* aabbs, uint32_t n, float out[6])
{
    for(int a = 0; a < 3; a++) { out[a] = FLT_MAX; out[3 + a] = -FLT_MAX; }
    for(uint32_t i = 0; i < n; i++)
        for(int a = 0; a < 3; a++)
        {
            float mn = aabbs[6 * i + a], mx = aabbs[6 * i + 3 + a];
            out[a] = (mn < out[a]) ? mn : out[a];
            out[3 + a] = (out[3 + a] < mx) ? mx : out[3 + a];
        }
}

/* ---------------------------------------------------------------------------------------------
 * KCGenMortonCode — Tracer/AcceleratorLBVH.cu:L96-168
 *   maxSide = largest extent; deltaRecip = 2^21 / double(maxSide);
 *   q = clamp(lround(float(double(max(c - min, 0)) * deltaRecip)), 0, 2^21 - 1)
 * ------------------------------------------------------------------------------------------- */
void orc_morton63(const float* centers, uint32_t n, const float aabb[6], uint64_t* out)
{
    float size[3] = {aabb[3] - aabb[0], aabb[4] - aabb[1], aabb[5] - aabb[2]};
    float maxSide = size[0];
    if(size[1] > maxSide) maxSide = size[1];
    if(size[2] > maxSide) maxSide = size[2];
    const double sliceCount = (double)(1ull << 21);
    const uint32_t lastValue = (1u << 21) - 1;
    double deltaRecip = sliceCount / (double)maxSide;
    for(uint32_t i = 0; i < n; i++)
    {
        uint32_t q[3];
        for(int a = 0; a < 3; a++)
        {
            float diff = centers[3 * i + a] - aabb[a];
            diff = (diff < 0.0f) ? 0.0f : diff; /* std::max(diff, 0) */
            float scaled = (float)((double)diff * deltaRecip);
            int32_t r = (int32_t)lroundf(scaled);
            uint32_t u = (uint32_t)r;
            q[a] = (u > lastValue) ? lastValue : u; /* Clamp(xyz, 0u, LastValue) on unsigned */
        }
        out[i] = orc_morton_compose64(q[0], q[1], q[2]);
    }
}

/* ---------------------------------------------------------------------------------------------
 * Stable ascending LSD radix sort of (key,value) pairs over a bit range —
 * semantics of DeviceAlgorithms::(Segmented)RadixSort<true,K,V>
 * (Device/CPU/AlgRadixSortCPU.h:L21-92; CUDA: cub::DeviceRadixSort::SortPairs). Any stable sort
 * gives the same permutation; this one is written as 8-bit LSD passes like the CPU backend.
 * ------------------------------------------------------------------------------------------- */
void orc_radix_sort_u64(uint64_t* keys, uint32_t* vals, uint32_t n, uint32_t bitBegin, uint32_t bitEnd)
{
    if(n == 0) return;
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * n);
    uint32_t* v2 = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint64_t *kin = keys, *kout = k2; uint32_t *vin = vals, *vout = v2;
    for(uint32_t b = bitBegin; b < bitEnd; b += 8)
    {
        uint32_t bits = (bitEnd - b < 8) ? (bitEnd - b) : 8;
        uint32_t mask = (1u << bits) - 1;
        uint32_t count[257]; memset(count, 0, sizeof(count));
        for(uint32_t i = 0; i < n; i++) count[((kin[i] >> b) & mask) + 1]++;
        for(uint32_t i = 0; i < 256; i++) count[i + 1] += count[i];
        for(uint32_t i = 0; i < n; i++)
        {
            uint32_t d = (uint32_t)((kin[i] >> b) & mask);
            uint32_t p = count[d]++;
            kout[p] = kin[i]; vout[p] = vin[i];
        }
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    if(kin != keys) { memcpy(keys, kin, sizeof(uint64_t) * n); memcpy(vals, vin, sizeof(uint32_t) * n); }
    free(k2); free(v2);
}

void orc_radix_sort_u32(uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t bitBegin, uint32_t bitEnd)
{
    if(n == 0) return;
    uint32_t* k2 = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* v2 = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t *kin = keys, *kout = k2; uint32_t *vin = vals, *vout = v2;
    for(uint32_t b = bitBegin; b < bitEnd; b += 8)
    {
        uint32_t bits = (bitEnd - b < 8) ? (bitEnd - b) : 8;
        uint32_t mask = (1u << bits) - 1;
        uint32_t count[257]; memset(count, 0, sizeof(count));
        for(uint32_t i = 0; i < n; i++) count[((kin[i] >> b) & mask) + 1]++;
        for(uint32_t i = 0; i < 256; i++) count[i + 1] += count[i];
        for(uint32_t i = 0; i < n; i++)
        {
            uint32_t d = (kin[i] >> b) & mask;
            uint32_t p = count[d]++;
            kout[p] = kin[i]; vout[p] = vin[i];
        }
        uint32_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    if(kin != keys) { memcpy(keys, kin, sizeof(uint32_t) * n); memcpy(vals, vin, sizeof(uint32_t) * n); }
    free(k2); free(v2);
}

/* ---------------------------------------------------------------------------------------------
 * Delta — Tracer/AcceleratorLBVH.cu:L69-94. NOTE the reference's equal-code fallback compares
 * the raw INDICES (clz64(i ^ j)) without the +64 offset of Karras 2012, so with duplicate codes
 * the index delta competes with code deltas. `robust` != 0 selects the augmented-key variant
 * (64 + clz64(i^j)) which is identical whenever all codes are distinct.
 * ------------------------------------------------------------------------------------------- */
static int32_t delta_fn(const uint64_t* codes, int32_t n, int32_t i, int32_t j, int robust)
{
    if(j < 0 || j >= n) return -1;
    uint64_t l = codes[i], r = codes[j];
    int32_t off = 0;
    if(l == r) { l = (uint64_t)i; r = (uint64_t)j; off = robust ? 64 : 0; }
    uint64_t d = l ^ r;
    int32_t lz = (d == 0) ? 64 : (int32_t)__builtin_clzll(d); /* Bit::CountLZero */
    return lz + off;
}

static int32_t divide_up2(int32_t v) { return (v + 1) / 2; }

/* KCConstructLBVHInternalNodes — Tracer/AcceleratorLBVH.cu:L170-299 (one accelerator).
 * nodes: [n-1][3] = {left, right, parent}; leaf children carry ORC_LEAF_FLAG | ORIGINAL leaf index
 * (through sortedIdx); leafParent is indexed by ORIGINAL leaf index. */
void orc_karras(const uint64_t* codes, const uint32_t* sortedIdx, uint32_t nLeaf,
                uint32_t* nodes, uint32_t* leafParent, int robust)
{
    int32_t totalLeafs = (int32_t)nLeaf;
    if(totalLeafs == 1)
    {
        nodes[0] = ORC_LEAF_FLAG | 0u; /* CombinedKey(IS_LEAF, 0) — index 0, not sortedIdx[0] */
        nodes[1] = ORC_INVALID;
        nodes[2] = ORC_INVALID;
        return;
    }
    int32_t totalNodes = totalLeafs - 1;
    for(int32_t i = 0; i < totalNodes; i++)
    {
        int32_t diff = delta_fn(codes, totalLeafs, i, i + 1, robust) - delta_fn(codes, totalLeafs, i, i - 1, robust);
        int32_t d = (diff < 0) ? -1 : 1;
        int32_t deltaMin = delta_fn(codes, totalLeafs, i, i - d, robust);
        int32_t lMax = 2;
        while(delta_fn(codes, totalLeafs, i, i + lMax * d, robust) > deltaMin) lMax <<= 1;
        int32_t l = 0;
        for(int32_t t = lMax >> 1; t != 0; t >>= 1)
            if(delta_fn(codes, totalLeafs, i, i + (l + t) * d, robust) > deltaMin) l += t;
        int32_t j = i + l * d;
        int32_t s = 0;
        int32_t deltaNode = delta_fn(codes, totalLeafs, i, j, robust);
        for(int32_t t = divide_up2(l); t != 0; t = (t == 1) ? 0 : divide_up2(t))
            if(delta_fn(codes, totalLeafs, i, i + (s + t) * d, robust) > deltaNode) s += t;
        int32_t gamma = i + s * d + ((d < 0) ? d : 0);
        uint32_t* me = nodes + 3 * (size_t)i;
        int32_t mn = (i < j) ? i : j, mx = (i < j) ? j : i;
        if(mn == gamma)
        {
            uint32_t leaf = sortedIdx[gamma];
            me[0] = ORC_LEAF_FLAG | leaf; leafParent[leaf] = (uint32_t)i;
        }
        else { me[0] = (uint32_t)gamma; nodes[3 * (size_t)gamma + 2] = (uint32_t)i; }
        if(mx == gamma + 1)
        {
            uint32_t leaf = sortedIdx[gamma + 1];
            me[1] = ORC_LEAF_FLAG | leaf; leafParent[leaf] = (uint32_t)i;
        }
        else { me[1] = (uint32_t)(gamma + 1); nodes[3 * (size_t)(gamma + 1) + 2] = (uint32_t)i; }
        if(i == 0) me[2] = ORC_INVALID;
    }
}

/* KCUnionLBVHBoundingBoxes — Tracer/AcceleratorLBVH.cu:L301-435. Serial form of the bottom-up
 * "second arriver unions" walk; min/max are exact so any schedule gives the same boxes. */
void orc_union_boxes(const uint32_t* nodes, const uint32_t* leafParent, const float* leafAABB,
                     uint32_t nLeaf, float* boxes)
{
    if(nLeaf == 1) { memcpy(boxes, leafAABB, 6 * sizeof(float)); return; }
    uint32_t nNode = nLeaf - 1;
    uint32_t* counters = (uint32_t*)calloc(nNode, sizeof(uint32_t));
    for(uint32_t i = 0; i < nLeaf; i++)
    {
        uint32_t ni = leafParent[i];
        while(ni != ORC_INVALID)
        {
            if(counters[ni]++ != 1) break;
            const uint32_t* nd = nodes + 3 * (size_t)ni;
            const float* l = (nd[0] & ORC_LEAF_FLAG) ? leafAABB + 6 * (size_t)(nd[0] & ~ORC_LEAF_FLAG)
                                                     : boxes + 6 * (size_t)nd[0];
            const float* r = (nd[1] & ORC_LEAF_FLAG) ? leafAABB + 6 * (size_t)(nd[1] & ~ORC_LEAF_FLAG)
                                                     : boxes + 6 * (size_t)nd[1];
            float* b = boxes + 6 * (size_t)ni;
            for(int a = 0; a < 3; a++)
            {
                b[a] = (r[a] < l[a]) ? r[a] : l[a];                 /* Math::Min(l, r) */
                b[3 + a] = (l[3 + a] < r[3 + a]) ? r[3 + a] : l[3 + a]; /* Math::Max(l, r) */
            }
            ni = nd[2];
        }
    }
    free(counters);
}

/* Whole MultiBuildLBVH chain for ONE accelerator of triangles (AcceleratorLBVH.hpp:L584-899). */
void orc_lbvh_build(const float* pos, const uint32_t* idx, uint32_t nTris, int robust,
                    float* leafAABB, float* accelAABB, uint64_t* morton,
                    uint64_t* sortedMorton, uint32_t* sortedIdx,
                    uint32_t* nodes, uint32_t* leafParent, float* boxes)
{
    float* centers = (float*)malloc(sizeof(float) * 3 * (size_t)nTris);
    orc_tri_aabb_center(pos, idx, nTris, leafAABB, centers);
    orc_aabb_union(leafAABB, nTris, accelAABB);
    orc_morton63(centers, nTris, accelAABB, morton);
    memcpy(sortedMorton, morton, sizeof(uint64_t) * nTris);
    for(uint32_t i = 0; i < nTris; i++) sortedIdx[i] = i; /* SegmentedIota */
    orc_radix_sort_u64(sortedMorton, sortedIdx, nTris, 0, 64);
    orc_karras(sortedMorton, sortedIdx, nTris, nodes, leafParent, robust);
    orc_union_boxes(nodes, leafParent, leafAABB, nTris, boxes);
    free(centers);
}

/* ---------------------------------------------------------------------------------------------
 * Ray::IntersectsAABB — Core/Ray.hpp:L192-219 (host Math::Min/Max = std::min/std::max)
 * ------------------------------------------------------------------------------------------- */
static inline float std_maxf(float a, float b) { return (a < b) ? b : a; }
static inline float std_minf(float a, float b) { return (b < a) ? b : a; }

static int slab_test(const float pos[3], const float dir[3], const float* box, float tMin, float tMax)
{
    float o0 = tMin, o1 = tMax;
    for(int i = 0; i < 3; i++)
    {
        float invD = 1.0f / dir[i];
        float t0 = (box[i] - pos[i]) * invD;
        float t1 = (box[3 + i] - pos[i]) * invD;
        if(invD < 0) { float t = t0; t0 = t1; t1 = t; }
        o0 = std_maxf(o0, std_minf(t0, t1));
        o1 = std_minf(o1, std_maxf(t0, t1));
    }
    return o1 >= o0;
}

/* Ray::IntersectsTriangle (Möller–Trumbore) — Core/Ray.hpp:L121-167, with Math::Cross / Math::Dot
 * built from FMA exactly as Core/Math.h:L1586-1596,L1627-1633:
 *   Dot   : r = fma(a0,b0,0); r = fma(a1,b1,r); r = fma(a2,b2,r)
 *   Cross : (fma(a1,b2,-(a2*b1)), fma(a2,b0,-(a0*b2)), fma(a0,b1,-(a1*b0)))
 * Returns 1 on hit; bary = (1-u-v, u) as the reference stores MetaHit (Vector2(baryCoords)). */
static inline float dot3(const float a[3], const float b[3])
{
    float r = fmaf(a[0], b[0], 0.0f);
    r = fmaf(a[1], b[1], r);
    r = fmaf(a[2], b[2], r);
    return r;
}
static inline void cross3(float out[3], const float a[3], const float b[3])
{
    out[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    out[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    out[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

int orc_ray_triangle(const float pos[3], const float dir[3],
                     const float* p0, const float* p1, const float* p2, int cullFace,
                     float* tOut, float bary[2], int* backFace)
{
    const float eps = 1.0e-7f; /* MathConstants::SmallEpsilon */
    float e0[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    float e1[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    float p[3]; cross3(p, dir, e1);
    float det = dot3(e0, p);
    int back = det < eps;
    int parallel = fabsf(det) < eps;
    if((cullFace && back) || parallel) return 0;
    float invDet = 1.0f / det;
    float tv[3] = {pos[0] - p0[0], pos[1] - p0[1], pos[2] - p0[2]};
    float u = dot3(tv, p) * invDet;
    if(u < 0 || u > 1) return 0;
    float q[3]; cross3(q, tv, e0);
    float v = dot3(dir, q) * invDet;
    if(v < 0 || (v + u) > 1) return 0;
    float t = dot3(e1, q) * invDet;
    if(t <= eps) return 0;
    float w = 1 - u - v;
    *tOut = t; bary[0] = w; bary[1] = u; *backFace = back;
    return 1;
}

/* Closest / first hit over one accelerator — TraverseLBVHStack (AcceleratorLBVH.hpp:L109-167)
 * with the leaf functors of ClosestHit / FirstHit (hpp:L327-410) and IntersectionCheck's range
 * test (hpp:L233-236,L270-271). rays: n*8 floats in RayGMem order (pos,tMin,dir,tMax). */
/* leafFilter (may be NULL): IntersectionCheck's stochastic alpha test (hpp:L263-282) — called for a leaf whose triangle was
 * hit inside the range with the hit's barycentrics; returning 0 drops the hit. */
typedef int (*orc_leaf_filter)(void* user, uint32_t leaf, const float bary[2]);
void orc_lbvh_trace_filtered(const float* pos, const uint32_t* idx,
                             const uint32_t* nodes, const float* boxes,
                             const float* rays, uint32_t nRays, int mode, int cullFace,
                             uint32_t* outPrim, float* outT, float* outBary, uint8_t* outBack,
                             orc_leaf_filter leafFilter, void* user)
{
    for(uint32_t r = 0; r < nRays; r++)
    {
        const float* ray = rays + 8 * (size_t)r;
        const float* rp = ray; const float* rd = ray + 4;
        float tMin = ray[3], tMax = ray[7];
        uint32_t best = ORC_INVALID; float bb[2] = {0, 0}; int bback = 0;
        uint32_t stack[160]; int sp = 0;
        stack[sp++] = 0;
        while(sp > 0)
        {
            uint32_t ni = stack[--sp];
            if(ni == ORC_INVALID) continue;
            if(ni & ORC_LEAF_FLAG)
            {
                uint32_t leaf = ni & ~ORC_LEAF_FLAG;
                const float* p0 = pos + 3 * (size_t)idx[3 * leaf + 0];
                const float* p1 = pos + 3 * (size_t)idx[3 * leaf + 1];
                const float* p2 = pos + 3 * (size_t)idx[3 * leaf + 2];
                float t, b2[2]; int back;
                if(!orc_ray_triangle(rp, rd, p0, p1, p2, cullFace, &t, b2, &back)) continue;
                if(!(t >= tMin && t < tMax)) continue;
                if(leafFilter && !leafFilter(user, leaf, b2)) continue;
                best = leaf; tMax = t; bb[0] = b2[0]; bb[1] = b2[1]; bback = back;
                if(mode == 1) break;
            }
            else if(slab_test(rp, rd, boxes + 6 * (size_t)ni, tMin, tMax))
            {
                stack[sp++] = nodes[3 * (size_t)ni + 1];
                stack[sp++] = nodes[3 * (size_t)ni + 0];
            }
        }
        outPrim[r] = best; outT[r] = tMax;
        outBary[2 * r] = bb[0]; outBary[2 * r + 1] = bb[1]; outBack[r] = (uint8_t)bback;
    }
}

void orc_lbvh_trace(const float* pos, const uint32_t* idx,
                    const uint32_t* nodes, const float* boxes,
                    const float* rays, uint32_t nRays, int mode, int cullFace,
                    uint32_t* outPrim, float* outT, float* outBary, uint8_t* outBack)
{
    orc_lbvh_trace_filtered(pos, idx, nodes, boxes, rays, nRays, mode, cullFace, outPrim, outT, outBary, outBack, NULL, NULL);
}

/* Topology-independent closest hit: brute force over all triangles, winner = min (t, rank) where
 * rank[leaf] is the leaf's position in the Morton-sorted order (= the reference's left-first visit
 * order). Equals orc_lbvh_trace whenever no ancestor box test of the winner fails numerically. */
void orc_brute_trace(const float* pos, const uint32_t* idx, uint32_t nTris, const uint32_t* rank,
                     const float* rays, uint32_t nRays, int cullFace,
                     uint32_t* outPrim, float* outT)
{
    for(uint32_t r = 0; r < nRays; r++)
    {
        const float* ray = rays + 8 * (size_t)r;
        float tMin = ray[3], tMax = ray[7];
        uint32_t best = ORC_INVALID; float bestT = tMax; uint32_t bestRank = ORC_INVALID;
        for(uint32_t i = 0; i < nTris; i++)
        {
            float t, b2[2]; int back;
            if(!orc_ray_triangle(ray, ray + 4, pos + 3 * (size_t)idx[3 * i], pos + 3 * (size_t)idx[3 * i + 1],
                                 pos + 3 * (size_t)idx[3 * i + 2], cullFace, &t, b2, &back)) continue;
            if(!(t >= tMin && t < tMax)) continue;
            uint32_t rk = rank ? rank[i] : i;
            if(t < bestT || (t == bestT && rk < bestRank)) { best = i; bestT = t; bestRank = rk; }
        }
        outPrim[r] = best; outT[r] = bestT;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Two-level accelerator (BaseAcceleratorLBVH, Tracer/AcceleratorLBVH.cu:L537-1035)
 * ------------------------------------------------------------------------------------------- */

/* Top-level LBVH over instance AABBs — InternalConstruct (AcceleratorLBVH.cu:L537-740): scene AABB =
 * union, centres = AABB::Centroid() = min + (max - min) * 0.5 (Core/AABB.hpp:L62-65, KCGenAABBCenters
 * L497-510), then the same Morton / sort / Karras / union chain as a bottom-level accelerator. */
void orc_tlas_build(const float* instAABB, uint32_t n, int robust,
                    float* sceneAABB, uint64_t* morton, uint64_t* sortedMorton, uint32_t* sortedIdx,
                    uint32_t* nodes, uint32_t* leafParent, float* boxes)
{
    float* centers = (float*)malloc(sizeof(float) * 3 * (size_t)n);
    for(uint32_t i = 0; i < n; i++)
        for(int a = 0; a < 3; a++)
        {
            float span = instAABB[6 * i + 3 + a] - instAABB[6 * i + a];
            centers[3 * i + a] = instAABB[6 * i + a] + span * 0.5f;
        }
    orc_aabb_union(instAABB, n, sceneAABB);
    orc_morton63(centers, n, sceneAABB, morton);
    memcpy(sortedMorton, morton, sizeof(uint64_t) * n);
    for(uint32_t i = 0; i < n; i++) sortedIdx[i] = i;
    orc_radix_sort_u64(sortedMorton, sortedIdx, n, 0, 64);
    orc_karras(sortedMorton, sortedIdx, n, nodes, leafParent, robust);
    orc_union_boxes(nodes, leafParent, instAABB, n, boxes);
    free(centers);
}

/* Matrix3x4::TransformAABB (Core/Matrix.hpp:L915-932) with the homogeneous coordinate held at 1 for all
 * eight corners. (The reference reassigns its Vector4 `vertex` from an operator* whose 4th lane is never
 * written — L777-787 — so for corners 1..7 its w is indeterminate; identity-transform instances do not go
 * through this function.) Dot = Math::Dot FMA chain. m: row-major 3x4. */
void orc_transform_aabb(const float m[12], const float aabb[6], float out[6])
{
    for(int a = 0; a < 3; a++) { out[a] = FLT_MAX; out[3 + a] = -FLT_MAX; }
    for(unsigned i = 0; i < 8; i++)
    {
        float v[4];
        for(unsigned j = 0; j < 3; j++) v[j] = ((i >> j) & 1u) ? aabb[3 + j] : aabb[j];
        v[3] = 1.0f;
        for(int r = 0; r < 3; r++)
        {
            float d = fmaf(m[4 * r + 0], v[0], 0.0f);
            d = fmaf(m[4 * r + 1], v[1], d);
            d = fmaf(m[4 * r + 2], v[2], d);
            d = fmaf(m[4 * r + 3], v[3], d);
            out[r] = (d < out[r]) ? d : out[r];
            out[3 + r] = (out[3 + r] < d) ? d : out[3 + r];
        }
    }
}

/* Matrix3x4::TransformRay (Core/Matrix.hpp:L905-912): dir' = M * dir (3-dot), pos' = M * (pos, 1) (4-dot);
 * no renormalisation, so t stays in world units (KCLocalRayCast, AcceleratorWork.kt.h:L207-218). */
static void transform_ray(const float m[12], const float* ray, float* out)
{
    for(int r = 0; r < 3; r++)
    {
        float d = fmaf(m[4 * r + 0], ray[4], 0.0f);
        d = fmaf(m[4 * r + 1], ray[5], d);
        d = fmaf(m[4 * r + 2], ray[6], d);
        out[4 + r] = d;
        float p = fmaf(m[4 * r + 0], ray[0], 0.0f);
        p = fmaf(m[4 * r + 1], ray[1], p);
        p = fmaf(m[4 * r + 2], ray[2], p);
        p = fmaf(m[4 * r + 3], 1.0f, p);
        out[r] = p;
    }
    out[3] = ray[3]; out[7] = ray[7];
}

typedef struct
{
    const float* pos; const uint32_t* idx; const uint32_t* nodes; const float* boxes; /* bottom-level accelerator */
    float invTransform[12];
    int identity;
} orc_instance;

/* CastRays / CastVisibilityRays over a two-level scene: left-first traversal of the top-level tree
 * (KCIntersectBaseLBVH: internal boxes AND the instance leaf AABB are slab-tested with the current
 * [tMin,tMax]); every reached instance runs the bottom-level ClosestHit / FirstHit in its local space and
 * shrinks tMax (the reference interleaves this per top-level round over all rays — per ray the visit order
 * is the same). mode 0 closest, 1 any. outInst = instance index or ORC_INVALID. */
void orc_scene_trace(const orc_instance* inst, const float* instAABB,
                     const uint32_t* tNodes, const float* tBoxes, uint32_t nInst,
                     const float* rays, uint32_t nRays, int mode, int cullFace,
                     uint32_t* outInst, uint32_t* outPrim, float* outT, float* outBary)
{
    for(uint32_t r = 0; r < nRays; r++)
    {
        const float* ray = rays + 8 * (size_t)r;
        float tMin = ray[3], tMax = ray[7];
        uint32_t bestI = ORC_INVALID, bestP = ORC_INVALID; float bb[2] = {0, 0};
        uint32_t stack[160]; int sp = 0; int done = 0;
        stack[sp++] = 0;
        while(sp > 0 && !done)
        {
            uint32_t ni = stack[--sp];
            if(ni == ORC_INVALID) continue;
            if(ni & ORC_LEAF_FLAG)
            {
                uint32_t ii = ni & ~ORC_LEAF_FLAG;
                if(!slab_test(ray, ray + 4, instAABB + 6 * (size_t)ii, tMin, tMax)) continue;
                float local[8];
                if(inst[ii].identity) memcpy(local, ray, sizeof(local)); else transform_ray(inst[ii].invTransform, ray, local);
                local[3] = tMin; local[7] = tMax;
                uint32_t prim; float t, b2[2]; uint8_t back;
                orc_lbvh_trace(inst[ii].pos, inst[ii].idx, inst[ii].nodes, inst[ii].boxes, local, 1, mode, cullFace, &prim, &t, b2, &back);
                if(prim != ORC_INVALID)
                {
                    bestI = ii; bestP = prim; tMax = t; bb[0] = b2[0]; bb[1] = b2[1];
                    if(mode == 1) done = 1;
                }
            }
            else if(slab_test(ray, ray + 4, tBoxes + 6 * (size_t)ni, tMin, tMax))
            {
                stack[sp++] = tNodes[3 * (size_t)ni + 1];
                stack[sp++] = tNodes[3 * (size_t)ni + 0];
            }
        }
        outInst[r] = bestI; outPrim[r] = bestP; outT[r] = tMax; outBary[2 * r] = bb[0]; outBary[2 * r + 1] = bb[1];
    }
}
