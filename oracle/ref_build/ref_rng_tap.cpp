// TEST INFRASTRUCTURE (oracle side). Runs the UNMODIFIED reference samplers RNGGroupSobol / RNGGroupZSobol
// (Tracer/Random.cu:L884-1408) on the reference's CPU backend and dumps what they generate, plus the
// Joe-Kuo generator matrices the Sobol sampler reads (SobolDetail::SobolMatrices, Tracer/SobolMatrices.cpp).
//
// usage: ref_rng_tap <type: 1 Sobol | 2 ZSobol> <W> <H> <seed> <initialMaxSPP> <increments> <out.bin>
//   out: u32 seeds[W*H] (LocalState.seed of every generator, = the mt19937(seed32) draws),
//        then for each sample increment k = 0..increments-1 (IncrementSampleId, then GenerateNumbers):
//          config A: dimensionStart  0, requests {2,1,3,2,1} -> u32 [9 * W*H]  (dimension-major, as the reference lays them out)
//          config B: dimensionStart 37, requests {3,2,1}     -> u32 [6 * W*H]
//          config C: dimensionStart 249, requests {2,2,3}    -> u32 [7 * W*H]  (dims 249..255: the last ones of the table)
//          config D: dimensionStart 300, requests {1,2}      -> u32 [3 * W*H]  (Sobol::RollDim wrap-around)
//        for type 1 additionally, at the very end: u32 matrices[256 * 52]
#include "Core/TracerI.h"
#include "Tracer/Random.h"
#include "Tracer/SobolMatrices.h"
#include "Device/GPUSystem.h"
#include "Device/GPUSystem.hpp"
#include "Core/ThreadPool.h"

#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

template<class Group>
static int Run(uint32_t W, uint32_t H, uint64_t seed, uint32_t maxSPP, uint32_t increments, FILE* out, bool dumpMatrices)
{
    GPUSystem system;
    const GPUQueue& queue = system.BestDevice().GetComputeQueue(0);
    ThreadPool pool;
    pool.RestartThreads(4, [](std::thread::native_handle_type, uint32_t) {});
    RenderImageParams rip{Vector2ui(W, H), Vector2ui(0, 0), Vector2ui(W, H)};
    Group group(rip, Vector2ui(W, H), maxSPP, seed, system, pool);
    group.SetupRange(Vector2ui(0, 0), Vector2ui(W, H), queue);
    queue.Barrier().Wait();
    const uint32_t n = W * H;
    for(uint32_t i = 0; i < n; i++) { uint32_t s = group.hMainStatesAll[i].seed; fwrite(&s, 4, 1, out); }
    const RNRequestList listA = GenRNRequestList<2, 1, 3, 2, 1>();
    const RNRequestList listB = GenRNRequestList<3, 2, 1>();
    const RNRequestList listC = GenRNRequestList<2, 2, 3>();
    const RNRequestList listD = GenRNRequestList<1, 2>();
    struct Cfg { uint16_t dimStart; RNRequestList list; };
    const Cfg cfgs[4] = {{0, listA}, {37, listB}, {249, listC}, {300, listD}};
    for(uint32_t k = 0; k < increments; k++)
    {
        group.IncrementSampleId(queue);
        for(const Cfg& c : cfgs)
        {
            std::vector<RandomNumber> numbers(size_t(c.list.TotalRNCount()) * n);
            group.GenerateNumbers(Span<RandomNumber>(numbers), c.dimStart, c.list, queue);
            queue.Barrier().Wait();
            fwrite(numbers.data(), 4, numbers.size(), out);
        }
    }
    if(dumpMatrices) fwrite(SobolDetail::SobolMatrices.data(), 4, SobolDetail::SobolMatrices.size(), out);
    return 0;
}

int main(int argc, char** argv)
{
    if(argc != 8) { fprintf(stderr, "usage: type W H seed maxSPP increments out\n"); return 1; }
    const int type = atoi(argv[1]);
    const uint32_t W = uint32_t(atoi(argv[2])), H = uint32_t(atoi(argv[3]));
    const uint64_t seed = strtoull(argv[4], nullptr, 10);
    const uint32_t maxSPP = uint32_t(atoi(argv[5])), increments = uint32_t(atoi(argv[6]));
    FILE* out = fopen(argv[7], "wb");
    if(!out) return 2;
    int rc;
    try
    {
        rc = (type == 1) ? Run<RNGGroupSobol>(W, H, seed, maxSPP, increments, out, true)
                         : Run<RNGGroupZSobol>(W, H, seed, maxSPP, increments, out, false);
    }
    catch(const MRayError& e) { fprintf(stderr, "MRayError: %s\n", e.GetError().c_str()); rc = 3; }
    fclose(out);
    return rc;
}
