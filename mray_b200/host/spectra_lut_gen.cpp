// spectra_lut_gen.cpp — `mray_b200_spectra_lut_gen <resolution> <colorspace> <output folder> [--dump-inputs file]`: the command
// line of the reference's SpectraLUTGen tool (Source/SpectraLUTGen/main.cpp:L539-634) with the optimisation running on the B200
// (mrb_spectra_lut_generate, csrc/spectra_lut.cu). Writes <folder>/<COLORSPACE>.mrspectra: "MR_SPECTRA", u32 resolution, u32 mode
// (1 = fp32), 9 * resolution^3 floats — the file SpectrumContextJakob2019 loads (Tracer/SpectrumContext.cu:L298-352).
// The colour tables (CIE 1931 observer, illuminant SPDs, primaries) are the reference's own, read from its headers / Core library.
#include "Core/Definitions.h"
#include "Core/Vector.h"
#include "Core/ColorFunctions.h"
#include "mray_b200.h"

#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <vector>

template<MRayColorSpaceEnum E>
static void Matrices(mrb_spectra_lut_desc& d)
{
    // FetchConversionMatrices (main.cpp:L58-79): the primaries' own RGB <-> XYZ, without the common-white-point adaptation
    static constexpr Matrix3x3 toXYZ = Color::GenRGBToXYZ(Color::Colorspace<E>::Prims);
    static constexpr Matrix3x3 fromXYZ = toXYZ.Inverse();
    for(unsigned r = 0; r < 3; r++) for(unsigned c = 0; c < 3; c++) { d.rgbToXYZ[3 * r + c] = toXYZ(r, c); d.xyzToRGB[3 * r + c] = fromXYZ(r, c); }
}

int main(int argc, const char* argv[])
{
    if(argc != 4 && argc != 6) { std::fprintf(stderr, "Wrong Argument Count(%d)\n", argc); return 1; }
    const uint32_t resolution = uint32_t(std::strtoul(argv[1], nullptr, 10));
    if(resolution == 0) { std::fprintf(stderr, "1st arg is not a number. (%s)\n", argv[1]); return 1; }
    const std::string csName = argv[2];
    const MRayColorSpaceEnum cs = MRayColorSpaceStringifier::FromString(csName);
    if(cs == MRayColorSpaceEnum::MR_ENUM_END) { std::fprintf(stderr, "Unknown color space name (%s)\n", csName.c_str()); return 1; }
    namespace fs = std::filesystem;
    std::error_code ec;
    fs::create_directories(argv[3], ec);
    if(ec) { std::fprintf(stderr, "Unable to create path towards \"%s\"\n", argv[3]); return 1; }

    mrb_spectra_lut_desc d = {};
    using enum MRayColorSpaceEnum;
    switch(cs)
    {
        case MR_ACES2065_1: Matrices<MR_ACES2065_1>(d); break;
        case MR_ACES_CG:    Matrices<MR_ACES_CG>(d); break;
        case MR_REC_709:    Matrices<MR_REC_709>(d); break;
        case MR_REC_2020:   Matrices<MR_REC_2020>(d); break;
        case MR_DCI_P3:     Matrices<MR_DCI_P3>(d); break;
        case MR_ADOBE_RGB:  Matrices<MR_ADOBE_RGB>(d); break;
        default: std::fprintf(stderr, "Unknown color space name (%s)\n", csName.c_str()); return 1;
    }
    static_assert(sizeof(Vector3) == 12 && Color::CIE_1931_N == 471);
    d.cieXYZ = reinterpret_cast<const float*>(Color::CIE_1931_XYZ.data());
    d.illuminantSPD = Color::SelectIlluminantSPD(cs).data();
    d.illuminantNormFactor = Color::SelectIlluminantSPDNormFactor(cs);
    d.resolution = resolution; d.optimizePassCount = 15u;
    if(argc == 6 && std::strcmp(argv[4], "--dump-inputs") == 0)
    {   // what the device routine is fed, for test fixtures: f32 cie[471*3], spd[471], norm, rgbToXYZ[9], xyzToRGB[9]
        std::ofstream f(argv[5], std::ios::binary);
        f.write(reinterpret_cast<const char*>(d.cieXYZ), 471 * 12); f.write(reinterpret_cast<const char*>(d.illuminantSPD), 471 * 4);
        f.write(reinterpret_cast<const char*>(&d.illuminantNormFactor), 4);
        f.write(reinterpret_cast<const char*>(d.rgbToXYZ), 36); f.write(reinterpret_cast<const char*>(d.xyzToRGB), 36);
        return f ? 0 : 1;
    }
    mrb_context ctx = nullptr;
    if(mrb_context_create(0, &ctx) != MRB_OK) { std::fprintf(stderr, "%s\n", mrb_last_error(nullptr)); return 1; }
    std::vector<float> lut(size_t(9) * resolution * resolution * resolution);
    double wp[3];
    if(mrb_spectra_lut_generate(ctx, &d, lut.data(), wp) != MRB_OK) { std::fprintf(stderr, "%s\n", mrb_last_error(ctx)); mrb_context_destroy(ctx); return 1; }
    mrb_context_destroy(ctx);
    const fs::path out = fs::absolute(fs::path(argv[3]) / (csName + std::string(Color::LUT_FILE_EXT)));
    std::ofstream f(out, std::ios::binary);
    if(!f) { std::fprintf(stderr, "Unable to open %s\n", out.string().c_str()); return 1; }
    f << Color::LUT_FILE_CC;
    f.write(reinterpret_cast<const char*>(&resolution), 4);
    const uint32_t mode = 1;
    f.write(reinterpret_cast<const char*>(&mode), 4);
    f.write(reinterpret_cast<const char*>(lut.data()), std::streamsize(lut.size() * 4));
    return f ? 0 : 1;
}
