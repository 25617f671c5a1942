// TEST INFRASTRUCTURE (oracle side). A small executable that runs the UNMODIFIED reference spectral
// code (SpectrumContextJakob2019 on its CPU backend) and dumps inputs/outputs as flat binary, the
// way Tests/Tracer/T_Spectrum.cu drives it. It is an executable rather than a library because the
// reference looks for "SpectraLUT/<COLORSPACE>.mrspectra" next to the PROCESS image
// (SpectrumContext.cu:L298-312), i.e. oracle/_ref/SpectraLUT/ when this binary lives in oracle/_ref/.
//
// usage: ref_spectrum_tap <in.bin> <out.bin> <mode: 0 Uniform | 1 GaussianMIS | 2 HyperbolicPBRT>
//   in : u32 nSamples, u32 nColors, u32 randoms[nSamples], f32 colors[nColors*3]
//   out: f32 observerXYZ[471*3] (normalised, texel centres), f32 illuminant[471], f32 xyzToRGB[9],
//        f32 waves[n*4], f32 pdfs[n*4],
//        per colour c: f32 albedoSpec[n*4], f32 radianceSpec[n*4] (of colour*4.5),
//                      f32 rgbOfAlbedoTimesIlluminant[n*4] (ConvertSpectraToRGB), f32 rgbOfRadiance[n*4]
#include "Tracer/SpectrumContext.h"
#include "Tracer/SpectrumContext.hpp"
#include "Tracer/Random.h"
#include "Device/GPUSystem.h"
#include "Device/GPUSystem.hpp"
#include "Core/ColorFunctions.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

static std::vector<char> ReadAll(const char* path)
{
    FILE* f = fopen(path, "rb");
    if(!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> b(size_t(n), 0);
    if(fread(b.data(), 1, size_t(n), f) != size_t(n)) exit(2);
    fclose(f);
    return b;
}

int main(int argc, char** argv)
{
    if(argc != 4) { fprintf(stderr, "usage: in out mode\n"); return 1; }
    std::vector<char> in = ReadAll(argv[1]);
    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(in.data());
    const uint32_t n = hdr[0], nColors = hdr[1];
    const uint32_t* randoms = hdr + 2;
    const float* colors = reinterpret_cast<const float*>(randoms + n);
    WavelengthSampleMode mode(WavelengthSampleMode::E(atoi(argv[3])));

    GPUSystem system;
    const GPUQueue& queue = system.BestDevice().GetComputeQueue(0);
    try
    {
        SpectrumContextJakob2019 ctx(MRayColorSpaceEnum::MR_ACES_CG, mode, system);
        Jakob2019Detail::Data data = ctx.GetData();
        FILE* out = fopen(argv[2], "wb");
        auto Write = [&](const void* p, size_t bytes) { fwrite(p, 1, bytes, out); };
        // tables, sampled at the texel centres
        for(uint32_t i = 0; i < Color::CIE_1931_N; i++) { Vector3 v = data.spdObserverXYZ(Float(i) + Float(0.5)); Write(&v, 12); }
        for(uint32_t i = 0; i < Color::CIE_1931_N; i++) { Float v = data.spdIlluminant(Float(i) + Float(0.5)); Write(&v, 4); }
        for(uint32_t i = 0; i < 9; i++) { Float v = data.XYZToRGB[i]; Write(&v, 4); }

        std::vector<SpectrumWaves> waves(n);
        std::vector<Spectrum> pdfs(n);
        std::vector<RandomNumber> rn(randoms, randoms + n);
        ctx.SampleSpectrumWavelengths(Span<SpectrumWaves>(waves), Span<Spectrum>(pdfs), Span<const RandomNumber>(rn), queue);
        queue.Barrier().Wait();
        for(uint32_t i = 0; i < n; i++) for(uint32_t k = 0; k < 4; k++) { Float v = waves[i][k]; Write(&v, 4); }
        for(uint32_t i = 0; i < n; i++) Write(&pdfs[i], 16);

        for(uint32_t c = 0; c < nColors; c++)
        {
            Vector3 col(colors[3 * c], colors[3 * c + 1], colors[3 * c + 2]);
            std::vector<Spectrum> alb(n), rad(n), rgbA(n), rgbR(n);
            for(uint32_t i = 0; i < n; i++)
            {
                SpectrumWaves w = waves[i];
                Jakob2019Detail::Converter conv(w, data);
                alb[i] = conv.ConvertAlbedo(col);
                rad[i] = conv.ConvertRadiance(col * Float(4.5));
                Spectrum s = alb[i];
                for(uint32_t k = 0; k < 4; k++) s[k] *= data.spdIlluminant(w[k] + Float(0.5) - Float(Color::CIE_1931_RANGE[0]));
                rgbA[i] = s; rgbR[i] = rad[i];
            }
            ctx.ConvertSpectrumToRGB(Span<Spectrum>(rgbA), Span<const SpectrumWaves>(waves), Span<const Spectrum>(pdfs), queue);
            ctx.ConvertSpectrumToRGB(Span<Spectrum>(rgbR), Span<const SpectrumWaves>(waves), Span<const Spectrum>(pdfs), queue);
            queue.Barrier().Wait();
            Write(alb.data(), 16 * size_t(n)); Write(rad.data(), 16 * size_t(n));
            Write(rgbA.data(), 16 * size_t(n)); Write(rgbR.data(), 16 * size_t(n));
        }
        fclose(out);
    }
    catch(const MRayError& e) { fprintf(stderr, "MRayError: %s\n", e.GetError().c_str()); return 3; }
    return 0;
}
