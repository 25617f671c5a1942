// TEST INFRASTRUCTURE (oracle side). Runs the UNMODIFIED reference DistributionGroupPwC2D (Tracer/Distributions.cu, CPU
// backend kernels) and the skysphere coordinate converters (Tracer/LightsDefault.hpp:L173-310) and dumps what they produce.
//
// usage: ref_dist_tap <in.bin> <out.bin>
//   in : u32 w, h, nSamples, nDirs; f32 function[w*h]; f32 xi[nSamples*2]; f32 dirs[nDirs*3] (unit, Y-up)
//   out: f32 cdfX[w*h], cdfY[h]; per sample f32 {u, v, SampleUV pdf, PdfUV(u, v)};
//        per converter (Spherical, CoOcta) and direction: f32 {DirToUV u, v, ToSolidAnglePdf(1, dir), UVToDir(DirToUV) xyz,
//        ToSolidAnglePdf(1, uv)}  (8 floats)
#include "Core/TracerI.h"
#include "Tracer/Distributions.h"
#include "Tracer/LightsDefault.h"
#include "Tracer/LightsDefault.hpp"
#include "Device/GPUSystem.h"
#include "Device/GPUSystem.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

template<class CC>
static void Converters(const std::vector<float>& dirs, FILE* out)
{
    for(size_t i = 0; i < dirs.size() / 3; i++)
    {
        Vector3 d(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        Vector2 uv = CC::DirToUV(d);
        Float pd = CC::ToSolidAnglePdf(Float(1), d);
        Vector3 back = CC::UVToDir(uv);
        Float pu = CC::ToSolidAnglePdf(Float(1), uv);
        float o[8] = {uv[0], uv[1], pd, back[0], back[1], back[2], pu, 0.0f};
        fwrite(o, 4, 8, out);
    }
}

int main(int argc, char** argv)
{
    if(argc < 3) return 1;
    FILE* in = fopen(argv[1], "rb");
    if(!in) return 2;
    uint32_t hdr[4];
    if(fread(hdr, 4, 4, in) != 4) return 3;
    const uint32_t w = hdr[0], h = hdr[1], nS = hdr[2], nD = hdr[3];
    std::vector<float> f(size_t(w) * h), xi(size_t(nS) * 2), dirs(size_t(nD) * 3);
    if(fread(f.data(), 4, f.size(), in) != f.size()) return 3;
    if(fread(xi.data(), 4, xi.size(), in) != xi.size()) return 3;
    if(nD && fread(dirs.data(), 4, dirs.size(), in) != dirs.size()) return 3;
    fclose(in);

    using namespace Distribution;
    GPUSystem system;
    const GPUQueue& queue = system.BestDevice().GetComputeQueue(0);
    DistributionGroupPwC2D group(system);
    uint32_t id = group.Reserve(Vector2ui(w, h));
    group.Commit();
    // the CPU backend's "device" memory is host memory: the function can be handed over directly
    group.Construct(id, Span<const Float>(f.data(), f.size()), queue);
    queue.Barrier().Wait();
    auto mem = group.DistMemory(id);
    auto dists = group.DeviceDistributions();

    FILE* out = fopen(argv[2], "wb");
    fwrite(mem.dCDFsX.data(), 4, mem.dCDFsX.size(), out);
    fwrite(mem.dCDFsY.data(), 4, mem.dCDFsY.size(), out);
    for(uint32_t i = 0; i < nS; i++)
    {
        SampleT<Vector2> s = dists[id].SampleUV(Vector2(xi[2 * i], xi[2 * i + 1]));
        Float p = dists[id].PdfUV(s.value);
        float o[4] = {s.value[0], s.value[1], s.pdf, p};
        fwrite(o, 4, 4, out);
    }
    Converters<LightDetail::SphericalCoordConverter>(dirs, out);
    Converters<LightDetail::CoOctaCoordConverter>(dirs, out);
    fclose(out);
    return 0;
}
