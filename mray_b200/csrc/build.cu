// build.cu — accelerator construction on the device.
//
//  Stage 1 (bit-exact with the reference; AcceleratorLBVH.hpp:L584-899):
//     KLeafAABB    <- KCGeneratePrimitiveKeys + KCGeneratePrimAABBs + SegmentedTransformReduce
//     KMorton      <- KCGenPrimCenters + KCGenMortonCode      (centroid recomputed, never stored)
//     RadixSort    <- SegmentedIota + SegmentedRadixSort<true,u64,u32>
//     KKarras      <- KCConstructLBVHInternalNodes
//     KUnionBoxes  <- KCUnionLBVHBoundingBoxes
//  Stage 2 (new): KCollapse + KFillTris turn the binary tree into the 8-wide quantised BVH.
//
// Floating point: every expression that feeds an integer artefact (Morton code) is written with
// explicit round-to-nearest intrinsics so that nvcc cannot contract it into an FMA.
#include "accel.cuh"
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <cfloat>
#include <cstring>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace mrb
{
namespace
{

constexpr int TPB = 256;

__device__ __forceinline__ uint32_t EncodeOrdered(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float DecodeOrdered(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__device__ __forceinline__ uint32_t LeafToPrim(const PrimRanges& r, uint32_t leaf, uint32_t& rangeIdx)
{
    const uint32_t k = (r.count == 1u) ? 0u : FindRange(r, leaf);
    rangeIdx = k;
    return r.primBegin[k] + (leaf - r.leafStart[k]);
}

__device__ __forceinline__ void LoadTri(const float* __restrict__ pos, const uint32_t* __restrict__ idx,
                                        uint32_t prim, float p[3][3])
{
    uint32_t i0 = idx[3 * size_t(prim) + 0], i1 = idx[3 * size_t(prim) + 1], i2 = idx[3 * size_t(prim) + 2];
    #pragma unroll
    for(int a = 0; a < 3; a++)
    {
        p[0][a] = pos[3 * size_t(i0) + a];
        p[1][a] = pos[3 * size_t(i1) + a];
        p[2][a] = pos[3 * size_t(i2) + a];
    }
}

// Per-leaf AABB (Shape::Triangle::BoundingBox, Core/ShapeFunctions.h:L48-57) + accelerator AABB.
__global__ void __launch_bounds__(TPB)
KLeafAABB(AccelData a)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(uint32_t leaf = blockIdx.x * TPB + threadIdx.x; leaf < a.leafCount; leaf += gridDim.x * TPB)
    {
        uint32_t ri; uint32_t prim = LeafToPrim(a.ranges, leaf, ri);
        float p[3][3]; LoadTri(a.positions, a.indices, prim, p);
        float lo[3], hi[3];
        #pragma unroll
        for(int k = 0; k < 3; k++)
        {
            float l = p[0][k], h = p[0][k];
            l = (p[1][k] < l) ? p[1][k] : l; l = (p[2][k] < l) ? p[2][k] : l;
            h = (h < p[1][k]) ? p[1][k] : h; h = (h < p[2][k]) ? p[2][k] : h;
            lo[k] = l; hi[k] = h;
            mn[k] = fminf(mn[k], l); mx[k] = fmaxf(mx[k], h);
        }
        float2* out = reinterpret_cast<float2*>(a.leafAABB + 6 * size_t(leaf));
        out[0] = make_float2(lo[0], lo[1]); out[1] = make_float2(lo[2], hi[0]); out[2] = make_float2(hi[1], hi[2]);
    }
    #pragma unroll
    for(int k = 0; k < 3; k++)
    {
        for(int o = 16; o > 0; o >>= 1)
        {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    // one atomic per block and bound: same-address atomics serialise in L2 (~1 ns each; one set per WARP cost this kernel ~15 of its
    // 37 us at 264 K triangles)
    __shared__ float sBox[TPB / 32][6];
    if((threadIdx.x & 31) == 0)
    {
        #pragma unroll
        for(int k = 0; k < 3; k++) { sBox[threadIdx.x >> 5][k] = mn[k]; sBox[threadIdx.x >> 5][3 + k] = mx[k]; }
    }
    __syncthreads();
    if(threadIdx.x < 6)
    {
        float v = sBox[0][threadIdx.x];
        for(int w = 1; w < TPB / 32; w++) v = (threadIdx.x < 3) ? fminf(v, sBox[w][threadIdx.x]) : fmaxf(v, sBox[w][threadIdx.x]);
        if(threadIdx.x < 3) atomicMin(&a.accelAABBEnc[threadIdx.x], EncodeOrdered(v));
        else atomicMax(&a.accelAABBEnc[threadIdx.x], EncodeOrdered(v));
    }
}

// Interleave — Graphics::MortonCode::Compose3D<uint64_t> (Core/GraphicsFunctions.h:L606-625)
__device__ __forceinline__ uint64_t Expand3D(uint32_t v)
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

// KCGenMortonCode (AcceleratorLBVH.cu:L96-168) on the centroid of Triangle::GetCenter
// (PrimitiveDefaultTriangle.hpp:L109-115). Also seeds the identity permutation (SegmentedIota).
__global__ void __launch_bounds__(TPB)
KMorton(AccelData a)
{
    float bl[3], sz[3];
    #pragma unroll
    for(int k = 0; k < 3; k++)
    {
        bl[k] = DecodeOrdered(a.accelAABBEnc[k]);
        sz[k] = __fsub_rn(DecodeOrdered(a.accelAABBEnc[3 + k]), bl[k]);
    }
    float maxSide = fmaxf(sz[0], fmaxf(sz[1], sz[2]));
    const double deltaRecip = __ddiv_rn(2097152.0, double(maxSide));
    const uint32_t lastValue = (1u << 21) - 1u;
    const float third = 0.333333333f;
    for(uint32_t leaf = blockIdx.x * TPB + threadIdx.x; leaf < a.leafCount; leaf += gridDim.x * TPB)
    {
        uint32_t ri; uint32_t prim = LeafToPrim(a.ranges, leaf, ri);
        float p[3][3]; LoadTri(a.positions, a.indices, prim, p);
        uint32_t q[3];
        #pragma unroll
        for(int k = 0; k < 3; k++)
        {
            float c = __fmul_rn(p[0][k], third);
            c = __fadd_rn(c, __fmul_rn(p[1][k], third));
            c = __fadd_rn(c, __fmul_rn(p[2][k], third));
            float diff = __fsub_rn(c, bl[k]);
            diff = (diff < 0.0f) ? 0.0f : diff;
            float scaled = __double2float_rn(__dmul_rn(double(diff), deltaRecip));
            int32_t r = int32_t(lroundf(scaled));
            uint32_t u = uint32_t(r);
            q[k] = (u > lastValue) ? lastValue : u;
        }
        uint64_t code = Expand3D(q[0]) | (Expand3D(q[1]) << 1) | (Expand3D(q[2]) << 2);
        a.morton[leaf] = code;
        a.sortedMorton[leaf] = code;
        a.sortedLeaf[leaf] = leaf;
    }
}

// Delta (AcceleratorLBVH.cu:L69-94); robust != 0 adds the +64 of Karras' augmented key.
__device__ __forceinline__ int32_t Delta(const uint64_t* __restrict__ codes, int32_t n, int32_t i, uint64_t ci,
                                         int32_t j, int robust)
{
    if(j < 0 || j >= n) return -1;
    uint64_t cj = codes[j];
    uint64_t l = ci, r = cj;
    int32_t off = 0;
    if(l == r) { l = uint64_t(i); r = uint64_t(j); off = robust ? 64 : 0; }
    return __clzll(l ^ r) + off;
}

// KCConstructLBVHInternalNodes (AcceleratorLBVH.cu:L170-299), one accelerator. Additionally
// records the sorted range each node covers (used by the wide collapse) and whether any two
// neighbouring codes are equal.
__global__ void __launch_bounds__(TPB)
KKarras(AccelData a, int robust, uint32_t* dupFlag, uint32_t* __restrict__ sortedLeafParent)
{
    const int32_t totalLeafs = int32_t(a.leafCount);
    const uint64_t* __restrict__ codes = a.sortedMorton;
    if(totalLeafs == 1)
    {
        if(blockIdx.x == 0 && threadIdx.x == 0)
        {
            a.nodes[0] = LBVHNode{LEAF_FLAG | 0u, INVALID_U32, INVALID_U32};
            a.nodeRange[0] = make_uint2(0, 0);
            a.leafParent[0] = 0u; // (the reference leaves it unset; KResolveExact walks it)
        }
        return;
    }
    const int32_t totalNodes = totalLeafs - 1;
    for(int32_t i = blockIdx.x * TPB + threadIdx.x; i < totalNodes; i += gridDim.x * TPB)
    {
        uint64_t ci = codes[i];
        if(ci == codes[i + 1]) *dupFlag = 1u;
        int32_t diff = Delta(codes, totalLeafs, i, ci, i + 1, robust) - Delta(codes, totalLeafs, i, ci, i - 1, robust);
        int32_t d = (diff < 0) ? -1 : 1;
        int32_t deltaMin = Delta(codes, totalLeafs, i, ci, i - d, robust);
        int32_t lMax = 2;
        while(Delta(codes, totalLeafs, i, ci, i + lMax * d, robust) > deltaMin) lMax <<= 1;
        int32_t l = 0;
        for(int32_t t = lMax >> 1; t != 0; t >>= 1)
            if(Delta(codes, totalLeafs, i, ci, i + (l + t) * d, robust) > deltaMin) l += t;
        int32_t j = i + l * d;
        int32_t s = 0;
        int32_t deltaNode = Delta(codes, totalLeafs, i, ci, j, robust);
        for(int32_t t = (l + 1) / 2; t != 0; t = (t == 1) ? 0 : (t + 1) / 2)
            if(Delta(codes, totalLeafs, i, ci, i + (s + t) * d, robust) > deltaNode) s += t;
        int32_t gamma = i + s * d + min(d, 0);
        int32_t lo = min(i, j), hi = max(i, j);
        uint32_t left, right;
        if(lo == gamma)
        {
            uint32_t leaf = a.sortedLeaf[gamma];
            left = LEAF_FLAG | leaf; a.leafParent[leaf] = uint32_t(i); sortedLeafParent[gamma] = uint32_t(i);
        }
        else { left = uint32_t(gamma); a.nodes[gamma].parent = uint32_t(i); }
        if(hi == gamma + 1)
        {
            uint32_t leaf = a.sortedLeaf[gamma + 1];
            right = LEAF_FLAG | leaf; a.leafParent[leaf] = uint32_t(i); sortedLeafParent[gamma + 1] = uint32_t(i);
        }
        else { right = uint32_t(gamma + 1); a.nodes[gamma + 1].parent = uint32_t(i); }
        a.nodes[i].left = left;
        a.nodes[i].right = right;
        if(i == 0) a.nodes[i].parent = INVALID_U32;
        a.nodeRange[i] = make_uint2(uint32_t(lo), uint32_t(hi));
    }
}

// KCUnionLBVHBoundingBoxes (AcceleratorLBVH.cu:L301-435): the second arriver at a node unions
// its children's boxes. Boxes cross threads through L2 (st.cg / ld.cg) with fences around the
// counter atomic.
__global__ void __launch_bounds__(TPB)
KUnionBoxes(AccelData a, uint32_t* counters, const uint32_t* __restrict__ sortedLeafParent)
{
    const uint32_t totalLeafs = a.leafCount;
    if(totalLeafs == 1)
    {
        if(blockIdx.x == 0 && threadIdx.x == 0)
            for(int k = 0; k < 3; k++) { a.boxes[0].min[k] = a.leafAABB[k]; a.boxes[0].max[k] = a.leafAABB[3 + k]; }
        return;
    }
    // threads walk up from the leaves in SORTED order: Karras' node i sits next to sorted position i, so the
    // counters, node records and boxes a warp touches are neighbours even when the input order is random
    for(uint32_t i = blockIdx.x * TPB + threadIdx.x; i < totalLeafs; i += gridDim.x * TPB)
    {
        uint32_t ni = sortedLeafParent[i];
        // the node record of a level is fetched one level ahead (the topology is read-only here), so a level of the walk costs the
        // counter's round trip + the two child boxes', not a third one for the record
        LBVHNode nd = a.nodes[ni];
        while(ni != INVALID_U32)
        {
            uint32_t prev = atomicAdd(&counters[ni], 1u);
            if(prev != 1u) break;
            __threadfence();
            LBVHNode up = nd;
            if(nd.parent != INVALID_U32) up = a.nodes[nd.parent];
            float l[6], r[6];
            const float* lp = (nd.left & LEAF_FLAG) ? a.leafAABB + 6 * size_t(nd.left & ~LEAF_FLAG)
                                                    : reinterpret_cast<const float*>(a.boxes + nd.left);
            const float* rp = (nd.right & LEAF_FLAG) ? a.leafAABB + 6 * size_t(nd.right & ~LEAF_FLAG)
                                                     : reinterpret_cast<const float*>(a.boxes + nd.right);
            #pragma unroll
            for(int k = 0; k < 6; k++) { l[k] = __ldcg(lp + k); r[k] = __ldcg(rp + k); }
            float* bp = reinterpret_cast<float*>(a.boxes + ni);
            #pragma unroll
            for(int k = 0; k < 3; k++)
            {
                __stcg(bp + k, (r[k] < l[k]) ? r[k] : l[k]);                     // Math::Min(l, r)
                __stcg(bp + 3 + k, (l[3 + k] < r[3 + k]) ? r[3 + k] : l[3 + k]); // Math::Max(l, r)
            }
            __threadfence();
            ni = nd.parent; nd = up;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Wide collapse
// ------------------------------------------------------------------------------------------------
struct CollapseState
{
    uint32_t head;     // unused
    uint32_t created;  // wide nodes allocated so far (root = 1)
    uint32_t done;     // wide nodes finished
    uint32_t triCount; // triangle records allocated
    uint32_t maxDepth;
    uint32_t error;
};

struct ChildRef { uint32_t lo, hi, node; float area; };

__device__ __forceinline__ float HalfArea(const float* b)
{
    float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ ChildRef MakeRef(const AccelData& a, uint32_t child, uint32_t leafPos)
{
    ChildRef r;
    if(child & LEAF_FLAG) { r.lo = r.hi = leafPos; r.node = INVALID_U32; r.area = -1.0f; }
    else
    {
        uint2 rg = a.nodeRange[child];
        r.lo = rg.x; r.hi = rg.y; r.node = child;
        r.area = HalfArea(reinterpret_cast<const float*>(a.boxes + child));
    }
    return r;
}

// One wide node. Everything per-child lives in registers: the loops over the (at most 8) children and slots
// are fully unrolled with predicates and the child <-> slot maps are packed nibbles, because a lone thread
// indexing local-memory arrays dynamically made this routine cost 60-140 K cycles per node — and the level
// loop below pays one node latency per level.
__device__ void CollapseNode(const AccelData& a, CollapseState* st, unsigned long long* queue,
                             uint32_t* triRank, uint32_t wideIdx, uint32_t binNode, uint32_t depth)
{
    const uint32_t MAX_LEAF = a.maxLeafSize;
    ChildRef refs[8];
    #pragma unroll
    for(int c = 0; c < 8; c++) refs[c] = ChildRef{0u, 0u, INVALID_U32, -1.0f};
    uint32_t n = 0;
    if(a.leafCount == 1) { n = 1; }
    else
    {
        LBVHNode nd = a.nodes[binNode];
        uint2 rg = a.nodeRange[binNode];
        refs[0] = MakeRef(a, nd.left, rg.x);
        refs[1] = MakeRef(a, nd.right, rg.y);
        n = 2;
        // phase 1: open (largest surface area first) only subtrees that cannot be a leaf child, so slots go to
        // shrinking the internal children; phase 2: left-over slots split multi-triangle leaves for tighter boxes
        for(int phase = 0; phase < 2 && n < 8; phase++)
        {
            while(n < 8)
            {
                int best = -1; float bestArea = -1.0f;
                #pragma unroll
                for(int c = 0; c < 8; c++)
                {
                    if(uint32_t(c) >= n || refs[c].node == INVALID_U32) continue;
                    const uint32_t size = refs[c].hi - refs[c].lo + 1;
                    if(phase == 0 && size <= MAX_LEAF) continue;
                    if(refs[c].area > bestArea) { bestArea = refs[c].area; best = c; }
                }
                if(best < 0) break;
                ChildRef o = refs[0];
                #pragma unroll
                for(int c = 1; c < 8; c++) if(c == best) o = refs[c];
                const LBVHNode on = a.nodes[o.node];
                const ChildRef l = MakeRef(a, on.left, o.lo), r = MakeRef(a, on.right, o.hi);
                #pragma unroll
                for(int c = 0; c < 8; c++) { if(c == best) refs[c] = l; if(uint32_t(c) == n) refs[c] = r; }
                n++;
            }
        }
    }
    // child boxes, node bounds
    float cb[8][6];
    float nb[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t innerMask = 0u, numInner = 0, numTris = 0;
    #pragma unroll
    for(int c = 0; c < 8; c++)
    {
        if(uint32_t(c) >= n) continue;
        const float* src = (refs[c].node != INVALID_U32)
                         ? reinterpret_cast<const float*>(a.boxes + refs[c].node)
                         : a.leafAABB + 6 * size_t(a.sortedLeaf[refs[c].lo]);
        #pragma unroll
        for(int k = 0; k < 6; k++) cb[c][k] = src[k];
    }
    #pragma unroll
    for(int c = 0; c < 8; c++)
    {
        if(uint32_t(c) >= n) continue;
        #pragma unroll
        for(int k = 0; k < 3; k++) { nb[k] = fminf(nb[k], cb[c][k]); nb[3 + k] = fmaxf(nb[3 + k], cb[c][3 + k]); }
        const uint32_t size = refs[c].hi - refs[c].lo + 1;
        if(size > MAX_LEAF) { innerMask |= 1u << c; numInner++; } else numTris += size;
    }
    // slot assignment: greedy on cost[c][s] = dot(centroid_c - centre, dir_s), dir_s component is
    // -1 where the slot bit is set (bit2 = x, bit1 = y, bit0 = z). A ray whose negative-direction
    // octant is r visits slot s with priority (s ^ r): slot r^7 first (nearest), slot r last.
    uint32_t slotOfChild = 0xFFFFFFFFu;   // nibble c = slot of child c (0xF = none yet)
    uint32_t childOfSlot = 0xFFFFFFFFu;   // nibble s = child in slot s (0xF = empty)
    uint32_t slotFree = 0xFFu, innerLeft = innerMask;
    const float cen[3] = {0.5f * (nb[0] + nb[3]), 0.5f * (nb[1] + nb[4]), 0.5f * (nb[2] + nb[5])};
    float ddx[8], ddy[8], ddz[8];
    #pragma unroll
    for(int c = 0; c < 8; c++)
    {
        ddx[c] = 0.5f * (cb[c][0] + cb[c][3]) - cen[0];
        ddy[c] = 0.5f * (cb[c][1] + cb[c][4]) - cen[1];
        ddz[c] = 0.5f * (cb[c][2] + cb[c][5]) - cen[2];
    }
    for(uint32_t it = 0; it < numInner; it++)
    {
        float bestCost = -FLT_MAX; uint32_t bc = 0, bs = 0;
        #pragma unroll
        for(int c = 0; c < 8; c++)
        {
            if(!((innerLeft >> c) & 1u)) continue;
            #pragma unroll
            for(int sl = 0; sl < 8; sl++)
            {
                if(!((slotFree >> sl) & 1u)) continue;
                const float cost = ((sl & 4) ? -ddx[c] : ddx[c]) + ((sl & 2) ? -ddy[c] : ddy[c]) + ((sl & 1) ? -ddz[c] : ddz[c]);
                if(cost > bestCost) { bestCost = cost; bc = uint32_t(c); bs = uint32_t(sl); }
            }
        }
        innerLeft &= ~(1u << bc); slotFree &= ~(1u << bs);
        slotOfChild = (slotOfChild & ~(0xFu << (4u * bc))) | (bs << (4u * bc));
        childOfSlot = (childOfSlot & ~(0xFu << (4u * bs))) | (bc << (4u * bs));
    }
    #pragma unroll
    for(int c = 0; c < 8; c++)
    {
        if(uint32_t(c) >= n || ((innerMask >> c) & 1u)) continue;
        const uint32_t sl = uint32_t(__ffs(int(slotFree))) - 1u;   // leaf children take the free slots in order
        slotFree &= ~(1u << sl);
        slotOfChild = (slotOfChild & ~(0xFu << (4u * uint32_t(c)))) | (sl << (4u * uint32_t(c)));
        childOfSlot = (childOfSlot & ~(0xFu << (4u * sl))) | (uint32_t(c) << (4u * sl));
    }
    // allocation
    uint32_t childBase = numInner ? atomicAdd(&st->created, numInner) : 0u;
    uint32_t triBase = numTris ? atomicAdd(&st->triCount, numTris) : 0u;
    if(numInner && childBase + numInner > a.wideNodeCapacity) { st->error = 2u; return; }
    // quantisation frame: per axis the smallest power-of-two cell such that 255 cells cover the node
    int ex[3]; double scale[3], invScale[3];
    #pragma unroll
    for(int k = 0; k < 3; k++)
    {
        float ext = nb[3 + k] - nb[k];
        int e = -126;
        if(ext > 0.0f)
        {
            int fe; float m = frexpf(ext / 255.0f, &fe);
            e = (m == 0.5f) ? fe - 1 : fe;
            while(ldexp(255.0, e) < double(nb[3 + k]) - double(nb[k])) e++;
            e = max(-126, min(127, e));
        }
        ex[k] = e; scale[k] = ldexp(1.0, e); invScale[k] = ldexp(1.0, -e);
    }
    // child boxes on the grid. (lo - p) and the scaling by a power of two are exact in fp64, so floor / ceil give
    // the tightest enclosing cells directly; the comparisons only guard the clamped ends.
    unsigned long long qlo[3] = {~0ull, ~0ull, ~0ull}, qhi[3] = {0ull, 0ull, 0ull}, meta = 0ull;   // empty slots: inverted box
    uint32_t imask = 0;
    #pragma unroll
    for(int c = 0; c < 8; c++)
    {
        if(uint32_t(c) >= n) continue;
        const uint32_t sl = (slotOfChild >> (4u * uint32_t(c))) & 0xFu, sh = 8u * sl;
        #pragma unroll
        for(int k = 0; k < 3; k++)
        {
            const double p = double(nb[k]), lo = double(cb[c][k]), hi = double(cb[c][3 + k]);
            int ql = __double2int_rd((lo - p) * invScale[k]);
            ql = max(0, min(255, ql));
            if(ql > 0 && p + ql * scale[k] > lo) ql--;
            int qh = __double2int_ru((hi - p) * invScale[k]);
            qh = max(0, min(255, qh));
            if(qh < 255 && p + qh * scale[k] < hi) qh++;
            qlo[k] = (qlo[k] & ~(0xFFull << sh)) | ((unsigned long long)(uint32_t(ql)) << sh);
            qhi[k] = (qhi[k] & ~(0xFFull << sh)) | ((unsigned long long)(uint32_t(qh)) << sh);
        }
        if((innerMask >> c) & 1u) { meta |= (unsigned long long)(0x20u | (24u + sl)) << sh; imask |= 1u << sl; }
    }
    // triangle groups take their record offsets in slot order
    uint32_t triOff = 0;
    #pragma unroll
    for(int sl = 0; sl < 8; sl++)
    {
        const uint32_t c = (childOfSlot >> (4u * uint32_t(sl))) & 0xFu;
        if(c == 0xFu || ((innerMask >> c) & 1u)) continue;
        uint32_t lo = 0, size = 0;
        #pragma unroll
        for(int cc = 0; cc < 8; cc++) if(uint32_t(cc) == c) { lo = refs[cc].lo; size = refs[cc].hi - refs[cc].lo + 1; }
        meta |= (unsigned long long)((((1u << size) - 1u) << 5) | triOff) << (8u * uint32_t(sl));
        for(uint32_t j = 0; j < size; j++) triRank[triBase + triOff + j] = lo + j;
        triOff += size;
    }
    WideNode w;
    w.q[0] = make_uint4(__float_as_uint(nb[0]), __float_as_uint(nb[1]), __float_as_uint(nb[2]),
                        uint32_t(ex[0] + 127) | (uint32_t(ex[1] + 127) << 8) | (uint32_t(ex[2] + 127) << 16) | (imask << 24));
    w.q[1] = make_uint4(childBase, triBase, uint32_t(meta), uint32_t(meta >> 32));
    w.q[2] = make_uint4(uint32_t(qlo[0]), uint32_t(qlo[0] >> 32), uint32_t(qlo[1]), uint32_t(qlo[1] >> 32));
    w.q[3] = make_uint4(uint32_t(qlo[2]), uint32_t(qlo[2] >> 32), uint32_t(qhi[0]), uint32_t(qhi[0] >> 32));
    w.q[4] = make_uint4(uint32_t(qhi[1]), uint32_t(qhi[1] >> 32), uint32_t(qhi[2]), uint32_t(qhi[2] >> 32));
    a.wideNodes[wideIdx] = w;
    // enqueue internal children in slot order
    uint32_t rel = 0;
    #pragma unroll
    for(int sl = 0; sl < 8; sl++)
    {
        if(!((imask >> sl) & 1u)) continue;
        const uint32_t c = (childOfSlot >> (4u * uint32_t(sl))) & 0xFu;
        uint32_t node = 0;
        #pragma unroll
        for(int cc = 0; cc < 8; cc++) if(uint32_t(cc) == c) node = refs[cc].node;
        queue[childBase + rel] = (unsigned long long)(node) | ((unsigned long long)(depth + 1) << 32);
        rel++;
    }
    atomicMax(&st->maxDepth, depth);
}

#include "collapse_group.cuh"   // CollapseNodeGroup + KCollapseGroups: the same node built by eight lanes (the product path)

// Level-synchronous collapse in ONE cooperative launch: wide nodes [begin,end) form the current
// level, their internal children are appended behind `end` (atomic bump of st->created), a grid-wide
// barrier separates levels. No spinning, no inter-thread waiting other than the barrier.
__global__ void __launch_bounds__(64)
KCollapse(AccelData a, CollapseState* st, unsigned long long* queue, uint32_t* triRank)
{
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gsize = gridDim.x * blockDim.x;
    uint32_t begin = 0, end = 1;
    while(begin < end)
    {
        for(uint32_t k = begin + gtid; k < end; k += gsize)
        {
            unsigned long long item = queue[k];
            CollapseNode(a, st, queue, triRank, k, uint32_t(item & 0xFFFFFFFFull), uint32_t(item >> 32));
        }
        grid.sync();
        begin = end;
        end = min(*reinterpret_cast<volatile uint32_t*>(&st->created), a.wideNodeCapacity);
        if(*reinterpret_cast<volatile uint32_t*>(&st->error)) break;
    }
}

// One thread per triangle record.
__global__ void __launch_bounds__(TPB)
KFillTris(AccelData a, const uint32_t* __restrict__ triRank)
{
    for(uint32_t s = blockIdx.x * TPB + threadIdx.x; s < a.leafCount; s += gridDim.x * TPB)
    {
        uint32_t rank = triRank[s];
        uint32_t leaf = a.sortedLeaf[rank];
        uint32_t ri; uint32_t prim = LeafToPrim(a.ranges, leaf, ri);
        float p[3][3]; LoadTri(a.positions, a.indices, prim, p);
        TriRecord t;
        t.v0 = make_float4(p[0][0], p[0][1], p[0][2], __uint_as_float(leaf));
        t.v1 = make_float4(__fsub_rn(p[1][0], p[0][0]), __fsub_rn(p[1][1], p[0][1]), __fsub_rn(p[1][2], p[0][2]),
                           __uint_as_float(rank));
        t.v2 = make_float4(__fsub_rn(p[2][0], p[0][0]), __fsub_rn(p[2][1], p[0][1]), __fsub_rn(p[2][2], p[0][2]),
                           __uint_as_float((a.ranges.cull[ri] ? 1u : 0u) | ((a.ranges.alphaMap && a.ranges.alphaMap[ri] >= 0) ? 2u : 0u) | (ri << 8)));
        a.tris[s] = t;
    }
}

} // namespace

namespace
{

// Matrix3x4::TransformAABB (Core/Matrix.hpp:L915-932; Math::Dot FMA chains) with the homogeneous
// coordinate held at 1 for all eight corners: the reference reassigns its Vector4 from an operator*
// (L777-787) that never writes the 4th lane, so its w is indeterminate after the first corner.
// Identity instances copy the accelerator AABB (TransformContextIdentity::Apply) and stay bit exact.
struct InstanceBuildIn { float transform[12]; const uint32_t* accelAABBEnc; uint32_t identity; uint32_t pad; };

__global__ void __launch_bounds__(TPB)
KInstanceAABB(AccelData a, const InstanceBuildIn* __restrict__ in, InstanceRec* __restrict__ recs)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(uint32_t i = blockIdx.x * TPB + threadIdx.x; i < a.leafCount; i += gridDim.x * TPB)
    {
        float box[6];
        #pragma unroll
        for(int k = 0; k < 6; k++) box[k] = DecodeOrdered(in[i].accelAABBEnc[k]);
        float lo[3], hi[3];
        if(in[i].identity) { for(int k = 0; k < 3; k++) { lo[k] = box[k]; hi[k] = box[3 + k]; } }
        else
        {
            for(int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
            const float* m = in[i].transform;
            for(uint32_t c = 0; c < 8; c++)
            {
                float v[3];
                for(uint32_t j = 0; j < 3; j++) v[j] = ((c >> j) & 1u) ? box[3 + j] : box[j];
                #pragma unroll
                for(int r = 0; r < 3; r++)
                {
                    float d = __fmaf_rn(m[4 * r + 0], v[0], 0.0f);
                    d = __fmaf_rn(m[4 * r + 1], v[1], d);
                    d = __fmaf_rn(m[4 * r + 2], v[2], d);
                    d = __fmaf_rn(m[4 * r + 3], 1.0f, d);
                    lo[r] = (d < lo[r]) ? d : lo[r];
                    hi[r] = (hi[r] < d) ? d : hi[r];
                }
            }
        }
        for(int k = 0; k < 3; k++)
        {
            a.leafAABB[6 * size_t(i) + k] = lo[k]; a.leafAABB[6 * size_t(i) + 3 + k] = hi[k];
            recs[i].worldAABB[k] = lo[k]; recs[i].worldAABB[3 + k] = hi[k];
            mn[k] = fminf(mn[k], lo[k]); mx[k] = fmaxf(mx[k], hi[k]);
        }
    }
    for(int k = 0; k < 3; k++)
        for(int o = 16; o > 0; o >>= 1)
        {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if((threadIdx.x & 31) == 0)
        for(int k = 0; k < 3; k++)
        {
            atomicMin(&a.accelAABBEnc[k], EncodeOrdered(mn[k]));
            atomicMax(&a.accelAABBEnc[3 + k], EncodeOrdered(mx[k]));
        }
}

// KCGenAABBCenters (AABB::Centroid = min + (max - min) * 0.5) + KCGenMortonCode over the scene AABB.
__global__ void __launch_bounds__(TPB)
KMortonBoxes(AccelData a)
{
    float bl[3], sz[3];
    for(int k = 0; k < 3; k++)
    {
        bl[k] = DecodeOrdered(a.accelAABBEnc[k]);
        sz[k] = __fsub_rn(DecodeOrdered(a.accelAABBEnc[3 + k]), bl[k]);
    }
    float maxSide = fmaxf(sz[0], fmaxf(sz[1], sz[2]));
    const double deltaRecip = __ddiv_rn(2097152.0, double(maxSide));
    const uint32_t lastValue = (1u << 21) - 1u;
    for(uint32_t leaf = blockIdx.x * TPB + threadIdx.x; leaf < a.leafCount; leaf += gridDim.x * TPB)
    {
        uint32_t q[3];
        for(int k = 0; k < 3; k++)
        {
            float lo = a.leafAABB[6 * size_t(leaf) + k], hi = a.leafAABB[6 * size_t(leaf) + 3 + k];
            float c = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), 0.5f));
            float diff = __fsub_rn(c, bl[k]);
            diff = (diff < 0.0f) ? 0.0f : diff;
            float scaled = __double2float_rn(__dmul_rn(double(diff), deltaRecip));
            uint32_t u = uint32_t(int32_t(lroundf(scaled)));
            q[k] = (u > lastValue) ? lastValue : u;
        }
        uint64_t code = Expand3D(q[0]) | (Expand3D(q[1]) << 1) | (Expand3D(q[2]) << 2);
        a.morton[leaf] = code; a.sortedMorton[leaf] = code; a.sortedLeaf[leaf] = leaf;
    }
}

__global__ void __launch_bounds__(TPB)
KFillInstanceSlots(AccelData a, const uint32_t* __restrict__ slotRank)
{
    for(uint32_t s = blockIdx.x * TPB + threadIdx.x; s < a.leafCount; s += gridDim.x * TPB)
        a.leafOfSlot[s] = a.sortedLeaf[slotRank[s]];
}

} // namespace

// BaseAcceleratorLBVH::InternalConstruct (Tracer/AcceleratorLBVH.cu:L537-740): top-level LBVH over the
// instances' world AABBs, then the same wide collapse as a bottom-level tree (one instance per leaf).
void BuildScene(Context& ctx, mrb_scene_t& sc, const mrb_instance_desc* inst, uint32_t n)
{
    SceneData& s = sc.d;
    AccelData& d = s.tlas;
    d = AccelData{};
    d.leafCount = n; d.nodeCount = n > 1 ? n - 1 : 1; d.maxLeafSize = 1;
    d.ranges = PrimRanges{};
    // per-instance LightOrMatKey overrides (instances of one accelerator with different materials)
    size_t overrideKeyCount = 0;
    for(uint32_t i = 0; i < n; i++) if(inst[i].lightOrMatKeys) overrideKeyCount += inst[i].accel->d.ranges.count;
    uint32_t* dInstKeys = nullptr;
    auto Layout = [&](MultiAlloc& ma)
    {
        d.leafAABB = ma.Take<float>(size_t(n) * 6);
        d.morton = ma.Take<uint64_t>(n); d.sortedMorton = ma.Take<uint64_t>(n); d.sortedLeaf = ma.Take<uint32_t>(n);
        d.nodes = ma.Take<LBVHNode>(d.nodeCount); d.leafParent = ma.Take<uint32_t>(n);
        d.boxes = ma.Take<LBVHBox>(d.nodeCount); d.nodeRange = ma.Take<uint2>(d.nodeCount);
        d.accelAABBEnc = ma.Take<uint32_t>(8);
        d.wideNodeCapacity = n + 2;
        d.wideNodes = ma.Take<WideNode>(d.wideNodeCapacity);
        d.leafOfSlot = ma.Take<uint32_t>(n);
        s.instances = ma.Take<InstanceRec>(n);
        dInstKeys = ma.Take<uint32_t>(overrideKeyCount);
    };
    MultiAlloc sz(nullptr); Layout(sz);
    sc.mem.Reserve(sz.Total());
    MultiAlloc ma(sc.mem.Base()); Layout(ma);
    ctx.persistentBytes += sc.mem.Capacity();
    s.instanceCount = n;

    std::vector<InstanceRec> recs(n);
    std::vector<InstanceBuildIn> bin(n);
    sc.accels.assign(n, nullptr);
    sc.hInstances.assign(inst, inst + n);
    sc.hInstanceKeys.assign(n, {});
    std::vector<uint32_t> flatKeys; flatKeys.reserve(overrideKeyCount);
    for(uint32_t i = 0; i < n; i++)
    {
        const mrb_accel_t& a = *inst[i].accel;
        sc.accels[i] = inst[i].accel;
        InstanceRec& r = recs[i];
        memcpy(r.invTransform, inst[i].invTransform, sizeof(r.invTransform));
        r.wideNodes = a.d.wideNodes; r.tris = a.d.tris; r.leafAABB = a.d.leafAABB;
        r.positions = a.d.positions; r.indices = a.d.indices; r.nodes = a.d.nodes; r.boxes = a.d.boxes;
        if(a.d.ranges.alphaMap) s.hasAlpha = 1u;
        r.ranges = a.d.ranges; r.accelKey = inst[i].accelKey; r.transKey = inst[i].transformKey;
        r.identity = inst[i].isIdentity ? 1u : 0u; r.leafCount = a.d.leafCount;
        if(inst[i].lightOrMatKeys)
        {
            r.ranges.lmKey = dInstKeys + flatKeys.size();
            sc.hInstanceKeys[i].assign(inst[i].lightOrMatKeys, inst[i].lightOrMatKeys + a.d.ranges.count);
            flatKeys.insert(flatKeys.end(), sc.hInstanceKeys[i].begin(), sc.hInstanceKeys[i].end());
            sc.hInstances[i].lightOrMatKeys = nullptr;   // the caller's array need not outlive the call
        }
        memcpy(bin[i].transform, inst[i].transform, sizeof(bin[i].transform));
        bin[i].accelAABBEnc = a.d.accelAABBEnc; bin[i].identity = r.identity; bin[i].pad = 0;
    }
    // scratch: instance inputs | sort temp | counters | collapse state | queue | slot ranks
    size_t sortBytes = RadixSortTempBytes(n, 8);
    MultiAlloc ssz(nullptr);
    ssz.Take<InstanceBuildIn>(n); ssz.Take<char>(sortBytes); ssz.Take<uint32_t>(d.nodeCount + 8); ssz.Take<CollapseState>(1);
    ssz.Take<unsigned long long>(d.wideNodeCapacity + 1); ssz.Take<uint32_t>(n);
    ctx.scratch.Reserve(ssz.Total());
    MultiAlloc sm(ctx.scratch.Base());
    InstanceBuildIn* dIn = sm.Take<InstanceBuildIn>(n);
    void* sortTemp = sm.Take<char>(sortBytes);
    uint32_t* counters = sm.Take<uint32_t>(d.nodeCount + 8);
    CollapseState* cst = sm.Take<CollapseState>(1);
    unsigned long long* queue = sm.Take<unsigned long long>(d.wideNodeCapacity + 1);
    uint32_t* slotRank = sm.Take<uint32_t>(n);
    uint32_t* dupFlag = counters + d.nodeCount;

    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<InstanceRec*>(s.instances), recs.data(), sizeof(InstanceRec) * n, cudaMemcpyHostToDevice, ctx.stream));
    if(!flatKeys.empty())
        MRB_CUDA_TRY(cudaMemcpyAsync(dInstKeys, flatKeys.data(), sizeof(uint32_t) * flatKeys.size(), cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(dIn, bin.data(), sizeof(InstanceBuildIn) * n, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaEventRecord(ctx.ev0, ctx.stream));
    const uint32_t initEnc[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
    MRB_CUDA_TRY(cudaMemcpyAsync(d.accelAABBEnc, initEnc, sizeof(initEnc), cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * (d.nodeCount + 8), ctx.stream));
    const uint32_t grid = GridFor(ctx, n, TPB, 8);
    MRB_LAUNCH(ctx, KInstanceAABB, grid, TPB, 0, d, dIn, const_cast<InstanceRec*>(s.instances));
    MRB_LAUNCH(ctx, KMortonBoxes, grid, TPB, 0, d);
    RadixSortPairs(ctx, d.sortedMorton, d.sortedLeaf, n, 0, 64, sortTemp);
    MRB_LAUNCH(ctx, KKarras, grid, TPB, 0, d, 1, dupFlag, slotRank);   // slotRank doubles as the sorted-position -> parent table until the collapse
    MRB_LAUNCH(ctx, KUnionBoxes, grid, TPB, 0, d, counters, slotRank);
    {
        CollapseState init = {0u, 1u, 0u, 0u, 0u, 0u};
        MRB_CUDA_TRY(cudaMemcpyAsync(cst, &init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
        unsigned long long rootItem = 0ull;
        MRB_CUDA_TRY(cudaMemcpyAsync(queue, &rootItem, sizeof(rootItem), cudaMemcpyHostToDevice, ctx.stream));
        int perSM = 0;
        // eight lanes per node (KCollapseGroups); MRB_COLLAPSE_SERIAL=1 / MRB_BUILD_SERIAL_COLLAPSE select the one-thread-per-node audit kernel
        static const bool serialCollapse = []{ const char* e = getenv("MRB_COLLAPSE_SERIAL"); return e && e[0] == '1'; }();
        const void* ck = serialCollapse ? (const void*)KCollapse : (const void*)KCollapseGroups;
        const uint32_t ctpb = serialCollapse ? 64u : COLLAPSE_TPB;
        MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, ck, int(ctpb), 0));
        uint32_t cgrid = min(uint32_t(ctx.smCount) * uint32_t(max(1, min(perSM, serialCollapse ? 16 : 4))),
                             max(1u, DivUp(d.wideNodeCapacity, serialCollapse ? 64u : COLLAPSE_TPB / 8u)));
        void* cargs[] = {(void*)&d, (void*)&cst, (void*)&queue, (void*)&slotRank};
        MRB_CUDA_TRY(cudaLaunchCooperativeKernel(ck, dim3(cgrid), dim3(ctpb), cargs, 0, ctx.stream));
        ctx.launches++;
        MRB_LAUNCH(ctx, KFillInstanceSlots, grid, TPB, 0, d, slotRank);
    }
    MRB_CUDA_TRY(cudaEventRecord(ctx.ev1, ctx.stream));
    uint32_t enc[8]; CollapseState hst = {};
    MRB_CUDA_TRY(cudaMemcpyAsync(enc, d.accelAABBEnc, sizeof(uint32_t) * 6, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(&hst, cst, sizeof(hst), cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    MRB_CUDA_TRY(cudaEventElapsedTime(&sc.buildMs, ctx.ev0, ctx.ev1));
    if(hst.error || hst.triCount != n) throw CudaError{cudaErrorUnknown, "top-level collapse failed", int(hst.error)};
    d.wideNodeCount = hst.created; d.wideDepth = hst.maxDepth + 1;
    for(int k = 0; k < 6; k++)
    {
        uint32_t u = enc[k]; uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
        memcpy(&sc.aabb[k], &b, 4);
    }
}

// Sizes the persistent block of one accelerator (sub-arrays are 256-byte aligned).
static void LayoutAccel(MultiAlloc& ma, AccelData& d, uint32_t vertexCount, uint32_t triCount, bool ownInputs, bool wide,
                        const mrb_accel_desc* alphaDesc = nullptr, std::vector<AlphaTex>* alphaTex = nullptr)
{
    if(alphaDesc)
    {   // alpha maps: per-range table, texture records + texels, vertex UVs
        d.ranges.alphaMap = ma.Take<int32_t>(d.ranges.count);
        d.ranges.alphaTex = ma.Take<AlphaTex>(alphaDesc->alphaTextureCount);
        d.ranges.uvs = ma.Take<float>(size_t(vertexCount) * 2);
        for(uint32_t t = 0; t < alphaDesc->alphaTextureCount; t++)
        {
            const mrb_texture_desc& td = alphaDesc->alphaTextures[t];
            (*alphaTex)[t] = AlphaTex{ma.Take<char>(size_t(td.width) * td.height * td.channels * (td.format == 0u ? 4u : 1u)),
                                      td.width, td.height, td.channels, td.format, td.interp, td.edge, 0u};
        }
    }
    if(ownInputs)
    {
        d.positions = ma.Take<float>(size_t(vertexCount) * 3);
        d.indices = ma.Take<uint32_t>(size_t(triCount) * 3);
    }
    d.leafAABB = ma.Take<float>(size_t(d.leafCount) * 6);
    d.morton = ma.Take<uint64_t>(d.leafCount);
    d.sortedMorton = ma.Take<uint64_t>(d.leafCount);
    d.sortedLeaf = ma.Take<uint32_t>(d.leafCount);
    d.nodes = ma.Take<LBVHNode>(d.nodeCount);
    d.leafParent = ma.Take<uint32_t>(d.leafCount);
    d.boxes = ma.Take<LBVHBox>(d.nodeCount);
    d.nodeRange = ma.Take<uint2>(d.nodeCount);
    d.accelAABBEnc = ma.Take<uint32_t>(8);
    d.ranges.leafStart = ma.Take<uint32_t>(d.ranges.count + 1);
    d.ranges.primBegin = ma.Take<uint32_t>(d.ranges.count);
    d.ranges.lmKey = ma.Take<uint32_t>(d.ranges.count);
    d.ranges.cull = ma.Take<uint32_t>(d.ranges.count);
    if(wide)
    {
        d.wideNodeCapacity = d.leafCount / 3 + 2;
        d.wideNodes = ma.Take<WideNode>(d.wideNodeCapacity);
        d.tris = ma.Take<TriRecord>(d.leafCount);
    }
}

void BuildAccel(Context& ctx, mrb_accel_t& acc, const mrb_accel_desc& desc)
{
    AccelData& d = acc.d;
    const bool wide = !(desc.flags & (MRB_BUILD_BINARY_ONLY | MRB_BUILD_REFERENCE_DELTA));
    const int robust = (desc.flags & MRB_BUILD_REFERENCE_DELTA) ? 0 : 1;
    // ranges (host table, then device copy inside the accelerator block)
    PrimRanges& r = d.ranges;
    r = PrimRanges{};
    r.primGroupId = desc.primGroupId;
    const bool whole = (desc.rangeCount == 0 || desc.primRanges == nullptr);
    r.count = whole ? 1u : desc.rangeCount;
    acc.hLeafStart.assign(r.count + 1, 0u); acc.hPrimBegin.assign(r.count, 0u);
    acc.vertexCount = desc.vertexCount; acc.triangleCount = desc.triangleCount;
    acc.hLmKey.assign(r.count, 0u); acc.hCull.assign(r.count, 0u);
    for(uint32_t i = 0; i < r.count; i++)
    {
        uint32_t b = whole ? 0u : desc.primRanges[2 * i], e = whole ? desc.triangleCount : desc.primRanges[2 * i + 1];
        acc.hPrimBegin[i] = b;
        acc.hLeafStart[i + 1] = acc.hLeafStart[i] + (e - b);
        acc.hLmKey[i] = desc.lightOrMatKeys ? desc.lightOrMatKeys[i] : 0u;
        acc.hCull[i] = (desc.cullBackface && desc.cullBackface[i]) ? 1u : 0u;
    }
    d.leafCount = acc.hLeafStart[r.count];
    d.nodeCount = d.leafCount > 1 ? d.leafCount - 1 : 1;
    // alpha maps (SurfaceParams.alphaMaps): validated here, uploaded below
    acc.hAlphaMap.clear();
    bool anyAlpha = false;
    if(desc.rangeAlphaMap)
        for(uint32_t i = 0; i < r.count; i++) anyAlpha = anyAlpha || desc.rangeAlphaMap[i] >= 0;
    std::vector<AlphaTex> alphaTex(anyAlpha ? desc.alphaTextureCount : 0u);
    if(anyAlpha)
    {
        if(!desc.vertexUVs || !desc.alphaTextures) throw std::runtime_error("alpha maps need vertexUVs and alphaTextures");
        acc.hAlphaMap.assign(desc.rangeAlphaMap, desc.rangeAlphaMap + r.count);
        for(int32_t m : acc.hAlphaMap) if(m >= int32_t(desc.alphaTextureCount)) throw std::runtime_error("rangeAlphaMap index exceeds alphaTextureCount");
        for(uint32_t t = 0; t < desc.alphaTextureCount; t++)
        {
            const mrb_texture_desc& td = desc.alphaTextures[t];
            if(!td.data || td.width == 0 || td.height == 0 || td.channels < 1 || td.channels > 4 || td.format > 1u || td.interp > 1u || td.edge > 2u)
                throw std::runtime_error("bad alpha texture descriptor");
        }
    }
    const mrb_accel_desc* alphaDesc = anyAlpha ? &desc : nullptr;

    MultiAlloc sizing(nullptr);
    LayoutAccel(sizing, d, desc.vertexCount, desc.triangleCount, true, wide, alphaDesc, &alphaTex);
    acc.mem.Reserve(sizing.Total());
    MultiAlloc ma(acc.mem.Base());
    LayoutAccel(ma, d, desc.vertexCount, desc.triangleCount, true, wide, alphaDesc, &alphaTex);
    ctx.persistentBytes += acc.mem.Capacity();

    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(r.leafStart), acc.hLeafStart.data(), 4 * (r.count + 1), cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(r.primBegin), acc.hPrimBegin.data(), 4 * r.count, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(r.lmKey), acc.hLmKey.data(), 4 * r.count, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(r.cull), acc.hCull.data(), 4 * r.count, cudaMemcpyHostToDevice, ctx.stream));
    cudaMemcpyKind kind = (desc.memspace == MRB_MEM_HOST) ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if(anyAlpha)
    {
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t*>(r.alphaMap), acc.hAlphaMap.data(), 4 * r.count, cudaMemcpyHostToDevice, ctx.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(r.uvs), desc.vertexUVs, sizeof(float) * 2 * size_t(desc.vertexCount), kind, ctx.stream));
        for(uint32_t t = 0; t < desc.alphaTextureCount; t++)
            MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<void*>(alphaTex[t].data), desc.alphaTextures[t].data,
                                         size_t(alphaTex[t].w) * alphaTex[t].h * alphaTex[t].channels * (alphaTex[t].format == 0u ? 4u : 1u),
                                         cudaMemcpyHostToDevice, ctx.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<AlphaTex*>(r.alphaTex), alphaTex.data(), sizeof(AlphaTex) * alphaTex.size(), cudaMemcpyHostToDevice, ctx.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));   // alphaTex (host vector) is read by the copy above
    }
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(d.positions), desc.positions,
                                 sizeof(float) * 3 * size_t(desc.vertexCount), kind, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(const_cast<uint32_t*>(d.indices), desc.indices,
                                 sizeof(uint32_t) * 3 * size_t(desc.triangleCount), kind, ctx.stream));

    // scratch: sort temp | union counters | collapse state | queue | triRank
    size_t sortBytes = RadixSortTempBytes(d.leafCount, 8);
    MultiAlloc ssz(nullptr);
    ssz.Take<char>(sortBytes); ssz.Take<uint32_t>(d.nodeCount + 8); ssz.Take<CollapseState>(1);
    ssz.Take<unsigned long long>(d.wideNodeCapacity + 1); ssz.Take<uint32_t>(d.leafCount);
    ctx.scratch.Reserve(ssz.Total());
    MultiAlloc sm(ctx.scratch.Base());
    void* sortTemp = sm.Take<char>(sortBytes);
    uint32_t* counters = sm.Take<uint32_t>(d.nodeCount + 8);
    CollapseState* cst = sm.Take<CollapseState>(1);
    unsigned long long* queue = sm.Take<unsigned long long>(d.wideNodeCapacity + 1);
    uint32_t* triRank = sm.Take<uint32_t>(d.leafCount);
    uint32_t* dupFlag = counters + d.nodeCount;

    MRB_CUDA_TRY(cudaEventRecord(ctx.ev0, ctx.stream));
    const uint32_t initEnc[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
    MRB_CUDA_TRY(cudaMemcpyAsync(d.accelAABBEnc, initEnc, sizeof(initEnc), cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * (d.nodeCount + 8), ctx.stream));

    const uint32_t grid = GridFor(ctx, d.leafCount, TPB, 8);
    MRB_LAUNCH(ctx, KLeafAABB, grid, TPB, 0, d);
    MRB_LAUNCH(ctx, KMorton, grid, TPB, 0, d);
    RadixSortPairs(ctx, d.sortedMorton, d.sortedLeaf, d.leafCount, 0, 64, sortTemp);
    MRB_LAUNCH(ctx, KKarras, grid, TPB, 0, d, robust, dupFlag, triRank);   // triRank doubles as the sorted-position -> parent table until the collapse
    MRB_LAUNCH(ctx, KUnionBoxes, grid, TPB, 0, d, counters, triRank);
    if(wide)
    {
        CollapseState init = {0u, 1u, 0u, 0u, 0u, 0u};
        MRB_CUDA_TRY(cudaMemcpyAsync(cst, &init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
        unsigned long long rootItem = 0ull; // binary node 0, depth 0
        MRB_CUDA_TRY(cudaMemcpyAsync(queue, &rootItem, sizeof(rootItem), cudaMemcpyHostToDevice, ctx.stream));
        int perSM = 0;
        // eight lanes per node (KCollapseGroups); MRB_COLLAPSE_SERIAL=1 / MRB_BUILD_SERIAL_COLLAPSE select the one-thread-per-node audit kernel
        static const bool serialEnv = []{ const char* e = getenv("MRB_COLLAPSE_SERIAL"); return e && e[0] == '1'; }();
        const bool serialCollapse = serialEnv || (desc.flags & MRB_BUILD_SERIAL_COLLAPSE);
        const void* ck = serialCollapse ? (const void*)KCollapse : (const void*)KCollapseGroups;
        const uint32_t ctpb = serialCollapse ? 64u : COLLAPSE_TPB;
        MRB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, ck, int(ctpb), 0));
        uint32_t cgrid = min(uint32_t(ctx.smCount) * uint32_t(max(1, min(perSM, serialCollapse ? 16 : 4))),
                             max(1u, DivUp(d.wideNodeCapacity, serialCollapse ? 64u : COLLAPSE_TPB / 8u)));
        void* cargs[] = {(void*)&d, (void*)&cst, (void*)&queue, (void*)&triRank};
        MRB_CUDA_TRY(cudaLaunchCooperativeKernel(ck, dim3(cgrid), dim3(ctpb), cargs, 0, ctx.stream));
        ctx.launches++;
        MRB_LAUNCH(ctx, KFillTris, grid, TPB, 0, d, triRank);
    }
    MRB_CUDA_TRY(cudaEventRecord(ctx.ev1, ctx.stream));

    // results the host needs
    uint32_t enc[8]; CollapseState hst = {};
    MRB_CUDA_TRY(cudaMemcpyAsync(enc, d.accelAABBEnc, sizeof(uint32_t) * 6, cudaMemcpyDeviceToHost, ctx.stream));
    uint32_t hDup = 0;
    MRB_CUDA_TRY(cudaMemcpyAsync(&hDup, dupFlag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    if(wide) MRB_CUDA_TRY(cudaMemcpyAsync(&hst, cst, sizeof(hst), cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    float ms = 0.f; MRB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx.ev0, ctx.ev1));

    acc.info.leafCount = d.leafCount;
    acc.info.nodeCount = d.nodeCount;
    acc.info.duplicateCodes = hDup;
    acc.info.buildMs = ms;
    acc.info.deviceBytes = acc.mem.Capacity();
    for(int k = 0; k < 6; k++)
    {
        uint32_t u = enc[k];
        uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
        float f; memcpy(&f, &b, 4); acc.info.aabb[k] = f;
    }
    if(wide)
    {
        if(hst.error || hst.triCount != d.leafCount)
            throw CudaError{cudaErrorUnknown, "wide collapse failed (queue stall / capacity)", int(hst.error)};
        d.wideNodeCount = hst.created;
        d.wideDepth = hst.maxDepth + 1;
        acc.info.wideNodeCount = hst.created;
    }
}

} // namespace mrb
