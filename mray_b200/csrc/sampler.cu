// Standalone sampler entry point: KCGenRandomNumbersGeneric of the reference for Sobol / Z-Sobol
// (Tracer/Random.cu:L439-555). The path tracer calls the device functions of sampler.cuh inside its
// own kernels instead.
#include "sampler.cuh"

namespace mrb
{
namespace
{
__global__ void __launch_bounds__(256) KSamplerGenerate(uint32_t type, const uint32_t* __restrict__ matrices, const uint32_t* __restrict__ seeds,
                                                        uint32_t width, uint32_t n, uint32_t sampleIndex, ZSobolGlobals g, uint32_t dimStart,
                                                        uint64_t requests, uint32_t requestCount, uint32_t* __restrict__ out)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if(i >= n) return;
    const uint32_t seed = seeds[i];
    const bool rb = (type & SAMPLER_REFERENCE_SCRAMBLE) == 0u;
    type &= SAMPLER_TYPE_MASK;
    const uint64_t morton = Morton2D(i % width, i / width);
    uint32_t o = 0;
    for(uint32_t r = 0; r < requestCount; r++)
    {
        const int dims = int((requests >> (2u * r)) & 3ull);
        uint32_t v[3];
        if(type == SAMPLER_SOBOL) SobolNext(matrices, seed, sampleIndex, dimStart + o, dims, v, rb);
        else ZSobolNext(matrices, seed, sampleIndex, morton, g, dimStart + o, dims, v, rb);
        for(int k = 0; k < dims; k++) out[i + size_t(n) * (o + uint32_t(k))] = v[k];
        o += uint32_t(dims);
    }
}
} // namespace

void SamplerGenerate(Context& ctx, uint32_t type, const uint32_t* matrices, const uint32_t* seeds, uint32_t width, uint32_t height,
                     uint32_t sampleIndex, uint32_t initialMaxSPP, uint32_t dimStart, uint64_t requests, uint32_t requestCount, uint32_t* out)
{
    uint32_t maxRes = width > height ? width : height, p2 = 1u, bits = 0u;
    while(p2 < maxRes) { p2 <<= 1; bits++; }
    const ZSobolGlobals g{initialMaxSPP, bits};
    const uint32_t n = width * height;
    MRB_LAUNCH(ctx, KSamplerGenerate, DivUp(n, 256u), 256, 0, type, matrices, seeds, width, n, sampleIndex, g, dimStart, requests, requestCount, out);
}
} // namespace mrb
