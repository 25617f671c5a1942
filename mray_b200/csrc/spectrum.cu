// SpectrumContextJakob2019 on the device: table upload + the reference's public spectrum operations
// (Tracer/SpectrumContext.h:L76-155) as standalone kernels. The path tracer uses the device functions
// of spectrum.cuh inside its own kernels instead of launching these.
#include "spectrum.cuh"
#include <vector>

namespace mrb
{
namespace
{
constexpr int STPB = 256;

// KCSampleSpectrumWavelengths (SpectrumContext.cu:L173-196): RNGDispenser::NextFloat<0> of one number per path
__global__ void __launch_bounds__(STPB) KSampleWavelengths(SpectrumData s, const uint32_t* __restrict__ randoms, uint32_t n,
                                                           float4* __restrict__ waves, float4* __restrict__ pdfs)
{
    const uint32_t i = blockIdx.x * STPB + threadIdx.x;
    if(i >= n) return;
    const float xi = fminf(float(randoms[i]) * 2.3283064365386963e-10f, 0.99999994f);
    float w[4], p[4];
    SampleWavelengths(s.mode, xi, w, p);
    waves[i] = make_float4(w[0], w[1], w[2], w[3]);
    pdfs[i] = make_float4(p[0], p[1], p[2], p[3]);
}

// KCConvertSpectrumToRGB (SpectrumContext.cu:L225-254), in place; .w = 0
__global__ void __launch_bounds__(STPB) KSpectraToRGB(SpectrumData s, float4* __restrict__ values, const float4* __restrict__ waves,
                                                      const float4* __restrict__ pdfs, uint32_t n)
{
    const uint32_t i = blockIdx.x * STPB + threadIdx.x;
    if(i >= n) return;
    const float4 v4 = values[i], w4 = waves[i], p4 = pdfs[i];
    const float v[4] = {v4.x, v4.y, v4.z, v4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w};
    const float3 rgb = SpectraToRGB(s, v, w, p);
    values[i] = make_float4(rgb.x, rgb.y, rgb.z, 0.0f);
}

// Converter::ConvertAlbedo / ConvertRadiance per element (SpectrumContext.hpp:L37-150)
__global__ void __launch_bounds__(STPB) KUpsample(SpectrumData s, const float* __restrict__ rgb, uint32_t rgbStride,
                                                  const float4* __restrict__ waves, uint32_t n, int isRadiance, float4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * STPB + threadIdx.x;
    if(i >= n) return;
    const float* c = rgb + size_t(rgbStride) * i;
    const float4 w4 = waves[i];
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
    float o[4];
    if(isRadiance)
    {
        const float4 k = FetchRadianceCoeffs(s, c[0], c[1], c[2]);
        #pragma unroll
        for(int j = 0; j < 4; j++) o[j] = EvalRadiance(s, k, w[j]);
    }
    else
    {
        const float3 k = FetchAlbedoCoeffs(s, c[0], c[1], c[2]);
        #pragma unroll
        for(int j = 0; j < 4; j++) o[j] = EvalSpectrum(k, w[j]);
    }
    out[i] = make_float4(o[0], o[1], o[2], o[3]);
}
} // namespace

void CreateSpectrum(Context& ctx, mrb_spectrum_t& sp, const mrb_spectrum_desc& desc)
{
    SpectrumData& d = sp.d;
    const size_t n3 = size_t(desc.lutResolution) * desc.lutResolution * desc.lutResolution;
    float* lut = nullptr; float4* obs = nullptr; float* ill = nullptr;
    auto Layout = [&](MultiAlloc& ma) { lut = ma.Take<float>(9 * n3); obs = ma.Take<float4>(CIE_N); ill = ma.Take<float>(CIE_N); };
    MultiAlloc sz(nullptr); Layout(sz);
    sp.mem.Reserve(sz.Total());
    MultiAlloc ma(sp.mem.Base()); Layout(ma);
    ctx.persistentBytes += sp.mem.Capacity();
    std::vector<float4> hobs(CIE_N);
    for(int i = 0; i < CIE_N; i++) hobs[i] = make_float4(desc.observerXYZ[3 * i], desc.observerXYZ[3 * i + 1], desc.observerXYZ[3 * i + 2], 0.f);
    MRB_CUDA_TRY(cudaMemcpyAsync(lut, desc.lut, sizeof(float) * 9 * n3, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(obs, hobs.data(), sizeof(float4) * CIE_N, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(ill, desc.illuminantSPD, sizeof(float) * CIE_N, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    d.lut = lut; d.n = desc.lutResolution; d.observer = obs; d.illuminant = ill;
    for(int i = 0; i < 9; i++) d.xyzToRGB[i] = desc.xyzToRGB[i];
    d.mode = desc.wavelengthSampleMode;
}

void SpectrumSampleWavelengths(Context& ctx, const mrb_spectrum_t& sp, const uint32_t* randoms, uint32_t n, float* waves, float* pdfs)
{
    MRB_LAUNCH(ctx, KSampleWavelengths, DivUp(n, uint32_t(STPB)), STPB, 0, sp.d, randoms, n,
               reinterpret_cast<float4*>(waves), reinterpret_cast<float4*>(pdfs));
}

void SpectrumToRGB(Context& ctx, const mrb_spectrum_t& sp, float* values, const float* waves, const float* pdfs, uint32_t n)
{
    MRB_LAUNCH(ctx, KSpectraToRGB, DivUp(n, uint32_t(STPB)), STPB, 0, sp.d, reinterpret_cast<float4*>(values),
               reinterpret_cast<const float4*>(waves), reinterpret_cast<const float4*>(pdfs), n);
}

void SpectrumUpsample(Context& ctx, const mrb_spectrum_t& sp, const float* rgb, uint32_t rgbStride, const float* waves, uint32_t n,
                      bool isRadiance, float* out)
{
    MRB_LAUNCH(ctx, KUpsample, DivUp(n, uint32_t(STPB)), STPB, 0, sp.d, rgb, rgbStride,
               reinterpret_cast<const float4*>(waves), n, isRadiance ? 1 : 0, reinterpret_cast<float4*>(out));
}

} // namespace mrb
