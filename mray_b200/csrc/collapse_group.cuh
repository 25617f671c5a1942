// collapse_group.cuh — CollapseNode of build.cu done by EIGHT LANES (one per child). Included by build.cu after CollapseNode /
// ChildRef / MakeRef / CollapseState, inside its anonymous namespace.
//
// The serial CollapseNode is a 15-20 K-instruction dependent chain on a lone thread (~40 us), and the level loop of the collapse
// pays it once per level (8 levels at 264 K triangles: 300 of the build's 560 us). Here the per-child work — reference loads,
// boxes, quantisation, slot costs — runs side by side, and the sequential parts (the greedy opening, the greedy slot assignment)
// reduce over the 8-lane group with shuffles. Decisions and tie-breaks are the serial routine's (first maximum in child order,
// then in slot order), so both build the same tree; MRB_COLLAPSE_SERIAL=1 keeps the serial kernel available as the audit path.

constexpr uint32_t COLLAPSE_TPB = 256;

__device__ void CollapseNodeGroup(const AccelData& a, CollapseState* st, unsigned long long* queue, uint32_t* triRank,
                                  uint32_t wideIdx, uint32_t binNode, uint32_t depth, uint32_t c, uint32_t gmask, uint32_t gbase, bool active,
                                  uint32_t* sAlloc)
{
    const uint32_t MAX_LEAF = a.maxLeafSize;
    // group-relative shuffles: lane index inside the warp = gbase + child
    auto ShflU = [&](uint32_t v, uint32_t child) { return __shfl_sync(gmask, v, int(gbase + child)); };
    ChildRef ref = ChildRef{0u, 0u, INVALID_U32, -1.0f};
    uint32_t refL = INVALID_U32, refR = INVALID_U32;   // children of ref.node, fetched together with the reference itself: an
                                                       // opening then costs ONE dependent round trip (its two new references)
    auto Fetch = [&](uint32_t child, uint32_t leafPos)
    {
        ref = MakeRef(a, child, leafPos);
        if(!(child & LEAF_FLAG)) { const LBVHNode cn = a.nodes[child]; refL = cn.left; refR = cn.right; }
    };
    uint32_t n = 0;
    if(!active) { n = 0; }
    else if(a.leafCount == 1) { n = 1; }
    else
    {
        const LBVHNode nd = a.nodes[binNode];
        const uint2 rg = a.nodeRange[binNode];
        if(c == 0u) Fetch(nd.left, rg.x);
        if(c == 1u) Fetch(nd.right, rg.y);
        n = 2;
        for(int phase = 0; phase < 2 && n < 8; phase++)
        {
            while(n < 8)
            {
                // candidate key of this lane's child; the winner is the largest area, the lowest child on ties
                const uint32_t sz = ref.hi - ref.lo + 1u;
                const bool cand = c < n && ref.node != INVALID_U32 && !(phase == 0 && sz <= MAX_LEAF);
                float bestArea = cand ? ref.area : -2.0f;
                uint32_t best = c;
                #pragma unroll
                for(int o = 4; o > 0; o >>= 1)
                {
                    const float oa = __shfl_xor_sync(gmask, bestArea, o);
                    const uint32_t ob = __shfl_xor_sync(gmask, best, o);
                    if(oa > bestArea || (oa == bestArea && ob < best)) { bestArea = oa; best = ob; }
                }
                // (the serial loop starts from bestArea = -1 with a strict '>': a candidate needs area > -1, i.e. any real area)
                if(!(bestArea > -1.0f)) break;
                const uint32_t oLeft = ShflU(refL, best), oRight = ShflU(refR, best), oLo = ShflU(ref.lo, best), oHi = ShflU(ref.hi, best);
                if(c == best) Fetch(oLeft, oLo);
                if(c == n) Fetch(oRight, oHi);
                n++;
            }
        }
    }
    const bool valid = c < n;
    // child box, node bounds
    float cb[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    if(valid)
    {
        const float* src = (ref.node != INVALID_U32) ? reinterpret_cast<const float*>(a.boxes + ref.node)
                                                     : a.leafAABB + 6 * size_t(a.sortedLeaf[ref.lo]);
        #pragma unroll
        for(int k = 0; k < 6; k++) cb[k] = src[k];
    }
    float nb[6];
    #pragma unroll
    for(int k = 0; k < 6; k++)
    {
        float v = cb[k];
        #pragma unroll
        for(int o = 4; o > 0; o >>= 1) { const float w = __shfl_xor_sync(gmask, v, o); v = (k < 3) ? fminf(v, w) : fmaxf(v, w); }
        nb[k] = v;
    }
    const uint32_t size = valid ? (ref.hi - ref.lo + 1u) : 0u;
    const bool inner = valid && size > MAX_LEAF;
    const uint32_t innerMask = (__ballot_sync(gmask, inner) >> gbase) & 0xFFu;
    const uint32_t leafMask = (__ballot_sync(gmask, valid && !inner) >> gbase) & 0xFFu;
    const uint32_t numInner = __popc(innerMask);
    uint32_t numTris = inner ? 0u : size;
    #pragma unroll
    for(int o = 4; o > 0; o >>= 1) numTris += __shfl_xor_sync(gmask, numTris, o);
    // slot assignment (see the serial routine): greedy on cost[c][s], first maximum in (child, slot) order
    const float cen[3] = {0.5f * (nb[0] + nb[3]), 0.5f * (nb[1] + nb[4]), 0.5f * (nb[2] + nb[5])};
    const float ddx = 0.5f * (cb[0] + cb[3]) - cen[0], ddy = 0.5f * (cb[1] + cb[4]) - cen[1], ddz = 0.5f * (cb[2] + cb[5]) - cen[2];
    uint32_t slotFree = 0xFFu, innerLeft = innerMask, mySlot = 0xFu;
    for(uint32_t it = 0; it < numInner; it++)
    {
        float bestCost = -FLT_MAX; uint32_t bs = 0, bc = c;
        const bool mine = (innerLeft >> c) & 1u;
        if(mine)
        {
            #pragma unroll
            for(int sl = 0; sl < 8; sl++)
            {
                if(!((slotFree >> sl) & 1u)) continue;
                const float cost = ((sl & 4) ? -ddx : ddx) + ((sl & 2) ? -ddy : ddy) + ((sl & 1) ? -ddz : ddz);
                if(cost > bestCost) { bestCost = cost; bs = uint32_t(sl); }
            }
        }
        // lanes without a candidate must lose whatever their cost field holds: carry a flag, not a sentinel cost
        uint32_t has = mine ? 1u : 0u;
        #pragma unroll
        for(int o = 4; o > 0; o >>= 1)
        {
            const float oc = __shfl_xor_sync(gmask, bestCost, o);
            const uint32_t obs = __shfl_xor_sync(gmask, bs, o), obc = __shfl_xor_sync(gmask, bc, o), oh = __shfl_xor_sync(gmask, has, o);
            const bool take = oh && (!has || oc > bestCost || (oc == bestCost && obc < bc));
            if(take) { bestCost = oc; bs = obs; bc = obc; has = 1u; }
        }
        // (a cost that is not above -FLT_MAX — NaN boxes — never wins the serial loop's '>': it keeps child 0 / slot 0)
        if(!(bestCost > -FLT_MAX)) { bc = 0u; bs = 0u; }
        innerLeft &= ~(1u << bc); slotFree &= ~(1u << bs);
        if(c == bc) mySlot = bs;
    }
    if(valid && !inner)
    {   // leaf children take the free slots in child order
        const uint32_t r = __popc(leafMask & ((1u << c) - 1u));
        mySlot = __fns(slotFree, 0u, int(r) + 1);
    }
    // allocation, aggregated over the BLOCK: two same-address atomics per node serialise in L2 (~1 ns each: 33 us on a level of
    // 16.7 K nodes, measured), so every block (COLLAPSE_TPB / 8 nodes) takes one range per counter and hands out sub-ranges.
    // Every thread of the block reaches this point in every round (inactive groups carry zero requests).
    constexpr uint32_t GROUPS = COLLAPSE_TPB / 8u;
    const uint32_t gib = threadIdx.x >> 3;
    if(c == 0u) { sAlloc[gib] = numInner; sAlloc[GROUPS + gib] = numTris; }
    __syncthreads();
    if(threadIdx.x < 2u)
    {   // thread 0: wide nodes, thread 1: triangle records — exclusive scan in place, then one atomic for the block
        uint32_t* v = sAlloc + threadIdx.x * GROUPS;
        uint32_t sum = 0u;
        for(uint32_t g = 0; g < GROUPS; g++) { const uint32_t x = v[g]; v[g] = sum; sum += x; }
        sAlloc[2u * GROUPS + threadIdx.x] = sum ? atomicAdd(threadIdx.x == 0u ? &st->created : &st->triCount, sum) : 0u;
    }
    __syncthreads();
    const uint32_t childBase = sAlloc[2u * GROUPS] + sAlloc[gib], triBase = sAlloc[2u * GROUPS + 1u] + sAlloc[GROUPS + gib];
    __syncthreads();   // the next round overwrites sAlloc
    if(!active) return;
    if(numInner && childBase + numInner > a.wideNodeCapacity) { if(c == 0u) st->error = 2u; return; }
    // quantisation frame (every lane computes it: three scalars)
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
    // this lane's child on the grid, then the node's byte rows by OR-reduction over the group
    const uint32_t sh = 8u * (mySlot & 7u);
    unsigned long long lo64[3] = {0ull, 0ull, 0ull}, hi64[3] = {0ull, 0ull, 0ull}, occ = 0ull, meta = 0ull;
    uint32_t imaskBit = 0u;
    if(valid)
    {
        #pragma unroll
        for(int k = 0; k < 3; k++)
        {
            const double p = double(nb[k]), lo = double(cb[k]), hi = double(cb[3 + k]);
            int ql = __double2int_rd((lo - p) * invScale[k]);
            ql = max(0, min(255, ql));
            if(ql > 0 && p + ql * scale[k] > lo) ql--;
            int qh = __double2int_ru((hi - p) * invScale[k]);
            qh = max(0, min(255, qh));
            if(qh < 255 && p + qh * scale[k] < hi) qh++;
            lo64[k] = (unsigned long long)(uint32_t(ql)) << sh;
            hi64[k] = (unsigned long long)(uint32_t(qh)) << sh;
        }
        occ = 0xFFull << sh;
        if(inner) { meta = (unsigned long long)(0x20u | (24u + mySlot)) << sh; imaskBit = 1u << mySlot; }
    }
    // triangle groups take their record offsets in slot order
    {
        const uint32_t leafSize = (valid && !inner) ? size : 0u;
        uint32_t triOff = 0u;
        #pragma unroll
        for(uint32_t j = 0; j < 8u; j++)
        {
            const uint32_t js = ShflU(mySlot, j), jsz = ShflU(leafSize, j);
            if(jsz && js < mySlot) triOff += jsz;
        }
        if(leafSize)
        {
            meta = (unsigned long long)((((1u << leafSize) - 1u) << 5) | triOff) << sh;
            for(uint32_t j = 0; j < leafSize; j++) triRank[triBase + triOff + j] = ref.lo + j;
        }
    }
    auto OrReduce64 = [&](unsigned long long v)
    {
        uint32_t l = uint32_t(v), h = uint32_t(v >> 32);
        #pragma unroll
        for(int o = 4; o > 0; o >>= 1) { l |= __shfl_xor_sync(gmask, l, o); h |= __shfl_xor_sync(gmask, h, o); }
        return (unsigned long long)l | ((unsigned long long)h << 32);
    };
    #pragma unroll
    for(int k = 0; k < 3; k++) { lo64[k] = OrReduce64(lo64[k]); hi64[k] = OrReduce64(hi64[k]); }
    occ = OrReduce64(occ); meta = OrReduce64(meta);
    uint32_t imask = imaskBit;
    #pragma unroll
    for(int o = 4; o > 0; o >>= 1) imask |= __shfl_xor_sync(gmask, imask, o);
    #pragma unroll
    for(int k = 0; k < 3; k++) lo64[k] |= ~occ;   // empty slots: inverted box (lo = 255, hi = 0)
    if(c == 0u)
    {
        WideNode w;
        w.q[0] = make_uint4(__float_as_uint(nb[0]), __float_as_uint(nb[1]), __float_as_uint(nb[2]),
                            uint32_t(ex[0] + 127) | (uint32_t(ex[1] + 127) << 8) | (uint32_t(ex[2] + 127) << 16) | (imask << 24));
        w.q[1] = make_uint4(childBase, triBase, uint32_t(meta), uint32_t(meta >> 32));
        w.q[2] = make_uint4(uint32_t(lo64[0]), uint32_t(lo64[0] >> 32), uint32_t(lo64[1]), uint32_t(lo64[1] >> 32));
        w.q[3] = make_uint4(uint32_t(lo64[2]), uint32_t(lo64[2] >> 32), uint32_t(hi64[0]), uint32_t(hi64[0] >> 32));
        w.q[4] = make_uint4(uint32_t(hi64[1]), uint32_t(hi64[1] >> 32), uint32_t(hi64[2]), uint32_t(hi64[2] >> 32));
        a.wideNodes[wideIdx] = w;
        atomicMax(&st->maxDepth, depth);
    }
    // internal children are enqueued in slot order
    if(inner)
    {
        const uint32_t rel = __popc(imask & ((1u << mySlot) - 1u));
        queue[childBase + rel] = (unsigned long long)(ref.node) | ((unsigned long long)(depth + 1) << 32);
    }
}

// The level loop of KCollapse with eight lanes per wide node (COLLAPSE_TPB / 8 nodes per block).
__global__ void __launch_bounds__(COLLAPSE_TPB)
KCollapseGroups(AccelData a, CollapseState* st, unsigned long long* queue, uint32_t* triRank)
{
    __shared__ uint32_t sAlloc[2u * (COLLAPSE_TPB / 8u) + 2u];
    cg::grid_group grid = cg::this_grid();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = lane & 7u, gbase = lane & 24u, gmask = 0xFFu << gbase;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, groups = (gridDim.x * blockDim.x) >> 3;
    uint32_t begin = 0, end = 1;
#ifdef MRB_COLLAPSE_TIMING   // diagnostic build: where the level loop's time goes
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned long long stampWork[24], stampSync[24]; uint32_t widths[24]; int lv = 0;
#endif
    while(begin < end)
    {
        // every group runs every round (the block-wide allocation inside needs all threads), idle ones with active = false
        for(uint32_t base = begin; base < end; base += groups)
        {
            const uint32_t k = base + group;
            const bool active = k < end;
            const unsigned long long item = active ? queue[k] : 0ull;
            CollapseNodeGroup(a, st, queue, triRank, k, uint32_t(item & 0xFFFFFFFFull), uint32_t(item >> 32), c, gmask, gbase, active, sAlloc);
        }
#ifdef MRB_COLLAPSE_TIMING
        if(lv < 24) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stampWork[lv])); widths[lv] = end - begin; }
#endif
        grid.sync();
#ifdef MRB_COLLAPSE_TIMING
        if(lv < 24) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stampSync[lv])); lv++; }
#endif
        begin = end;
        end = min(*reinterpret_cast<volatile uint32_t*>(&st->created), a.wideNodeCapacity);
        if(*reinterpret_cast<volatile uint32_t*>(&st->error)) break;
    }
#ifdef MRB_COLLAPSE_TIMING
    if(blockIdx.x == 0 && threadIdx.x == 0)
    {
        unsigned long long prev = t0;
        for(int l = 0; l < lv; l++)
        {
            printf("[collapse] level %d width %u: work %llu ns, sync %llu ns\n", l, widths[l], stampWork[l] - prev, stampSync[l] - stampWork[l]);
            prev = stampSync[l];
        }
        printf("[collapse] grid %u blocks, total %llu ns\n", gridDim.x, prev - t0);
    }
#endif
}
