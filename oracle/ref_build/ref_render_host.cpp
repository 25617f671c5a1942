// ref_render_host.cpp — a process to host libtracer_driver.so in (TEST / BASELINE INFRASTRUCTURE ONLY).
//
// The reference's spectral renderer loads "SpectraLUT/<COLORSPACE>.mrspectra" from the directory of the
// running EXECUTABLE (GetProcessPath(), Tracer/SpectrumContext.cu:L302-308); under ctypes that is the Python
// interpreter's directory, which this repository must not touch. This tiny host lives in oracle/_ref/ next to
// SpectraLUT/, reads a scene blob written by tests/oracle_lib.py::driver_render(host_exe=True), calls
// tracer_driver_render() and writes the image back. No reference headers are needed here.
//
//   ref_render_host <in.blob> <out.bin>
// blob: u64 sectionCount, then per section u64 byteCount + bytes (padded to 8). Section order: see below.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <unistd.h>
#include <libgen.h>
#include <climits>

struct DriverScene
{
    uint32_t batchCount; const uint32_t* batchVertexOffsets; const uint32_t* batchTriOffsets;
    const float* positions; const float* normals; const uint32_t* indices;
    const int32_t* batchMaterial; const int32_t* batchLight;
    uint32_t materialCount; const float* albedo; uint32_t lightCount; const float* radiance;
    float camPos[3], camGaze[3], camUp[3]; float fovXY[2]; float nearFar[2];
    const float* batchTransforms; const int32_t* batchInstanceOf;
    uint32_t textureCount; const uint32_t* textureInfo; const uint8_t* textureBytes; const int32_t* materialTexture; const float* uvs; const uint8_t* materialKind; const uint8_t* lightTwoSided; const float* materialParams;
    uint32_t boundaryType; float boundaryRadiance[3]; int32_t boundaryTexture; const float* boundaryTransform;
    const int32_t* batchAlphaMap; const int32_t* materialNormalMap; const uint32_t* textureMipCounts;
};
struct DriverRender
{
    const char* rendererName; uint32_t width, height; uint32_t totalSPP; const char* sampleMode;
    uint32_t rrRange[2]; uint64_t seed; uint32_t accelMode; uint32_t parallelHint; uint32_t threads; uint32_t samplerType; uint32_t region[4]; uint32_t latency; uint32_t burstSize; uint32_t camSwitchAfter; float camSwitch[9];
    uint32_t filmFilter; float filmFilterRadius; uint32_t genMips; uint32_t mipGenFilter; float mipGenFilterRadius;
};
struct DriverStats { double commitSeconds, renderSeconds, totalPaths; uint32_t iterations; float sceneAABB[6]; double startSeconds, sceneSeconds, closeSeconds, totalSeconds; };
using RenderF = int (*)(const char*, const DriverScene*, const DriverRender*, float*, float*, DriverStats*, char*, size_t);

int main(int argc, char** argv)
{
    if(argc != 3) { std::fprintf(stderr, "usage: %s in.blob out.bin\n", argv[0]); return 64; }
    FILE* f = std::fopen(argv[1], "rb");
    if(!f) { std::perror("blob"); return 65; }
    uint64_t n = 0;
    if(std::fread(&n, 8, 1, f) != 1 || (n < 23 || n > 28)) { std::fprintf(stderr, "bad blob\n"); return 66; }
    std::vector<std::vector<uint64_t>> sec(n);     // 8-byte aligned storage
    std::vector<uint64_t> bytes(n);
    for(uint64_t i = 0; i < n; i++)
    {
        if(std::fread(&bytes[i], 8, 1, f) != 1) return 66;
        sec[i].resize((bytes[i] + 7) / 8 + 1, 0);
        if(bytes[i] && std::fread(sec[i].data(), 1, (bytes[i] + 7) / 8 * 8, f) != (bytes[i] + 7) / 8 * 8) return 66;
    }
    std::fclose(f);
    auto P = [&](int i) { return bytes[i] ? reinterpret_cast<const void*>(sec[i].data()) : nullptr; };
    // sections: 0 dll path, 1 renderer name, 2 sample mode (NUL-terminated by the zero padding),
    // 3 u32[30] {..., filmFilter, filmFilterRadius (float bits)} after: {batchCount, materialCount, lightCount, width, height, spp, rr0, rr1, accelMode, parallelHint, threads, sampler, region[4],
    //   latency, burstSize, camSwitchAfter, camSwitch[9] (float bits)},
    // 4 u64 seed, 5 f32[13] camera {pos, gaze, up, fovXY, nearFar}, 6 vertexOffsets, 7 triOffsets, 8 positions,
    // 9 normals, 10 indices, 11 batchMaterial, 12 batchLight, 13 albedo, 14 radiance, 15 batchTransforms (may be empty), 16 batchInstanceOf (may be empty),
    // 17 textureInfo (8 u32 per texture; may be empty), 18 textureBytes, 19 materialTexture, 20 uvs (may be empty), 21 materialKind (may be empty), 22 lightTwoSided (may be empty)
    const uint32_t* u = static_cast<const uint32_t*>(P(3));
    const float* cam = static_cast<const float*>(P(5));
    DriverScene sc{};
    sc.batchCount = u[0]; sc.materialCount = u[1]; sc.lightCount = u[2];
    sc.batchVertexOffsets = static_cast<const uint32_t*>(P(6)); sc.batchTriOffsets = static_cast<const uint32_t*>(P(7));
    sc.positions = static_cast<const float*>(P(8)); sc.normals = static_cast<const float*>(P(9));
    sc.indices = static_cast<const uint32_t*>(P(10));
    sc.batchMaterial = static_cast<const int32_t*>(P(11)); sc.batchLight = static_cast<const int32_t*>(P(12));
    sc.albedo = static_cast<const float*>(P(13)); sc.radiance = static_cast<const float*>(P(14));
    sc.batchTransforms = static_cast<const float*>(P(15)); sc.batchInstanceOf = static_cast<const int32_t*>(P(16));
    sc.textureCount = uint32_t(bytes[17] / 32); sc.textureInfo = static_cast<const uint32_t*>(P(17));
    sc.textureBytes = static_cast<const uint8_t*>(P(18)); sc.materialTexture = static_cast<const int32_t*>(P(19));
    sc.uvs = static_cast<const float*>(P(20)); sc.materialKind = static_cast<const uint8_t*>(P(21));
    sc.lightTwoSided = static_cast<const uint8_t*>(P(22));
    if(n > 23) sc.materialParams = static_cast<const float*>(P(23));   // 23 materialParams (8 floats per material; may be empty)
    if(n > 24 && bytes[24] >= 20)
    {   // 24 boundary light: u32 type, f32 radiance[3], i32 texture, then (optional) f32[12] transform
        const uint32_t* b = static_cast<const uint32_t*>(P(24));
        sc.boundaryType = b[0]; std::memcpy(sc.boundaryRadiance, b + 1, 12); std::memcpy(&sc.boundaryTexture, b + 4, 4);
        if(bytes[24] >= 20 + 48) sc.boundaryTransform = reinterpret_cast<const float*>(b + 5);
    }
    if(n > 25) sc.batchAlphaMap = static_cast<const int32_t*>(P(25));   // 25 batchAlphaMap (i32 per batch; may be empty)
    if(n > 26) sc.materialNormalMap = static_cast<const int32_t*>(P(26));   // 26 materialNormalMap (i32 per material; may be empty)
    if(n > 27) sc.textureMipCounts = static_cast<const uint32_t*>(P(27));   // 27 textureMipCounts (u32 per texture; may be empty)
    std::memcpy(sc.camPos, cam, 12); std::memcpy(sc.camGaze, cam + 3, 12); std::memcpy(sc.camUp, cam + 6, 12);
    std::memcpy(sc.fovXY, cam + 9, 8); std::memcpy(sc.nearFar, cam + 11, 8);
    DriverRender rd{};
    rd.rendererName = static_cast<const char*>(P(1)); rd.sampleMode = static_cast<const char*>(P(2));
    rd.width = u[3]; rd.height = u[4]; rd.totalSPP = u[5]; rd.rrRange[0] = u[6]; rd.rrRange[1] = u[7];
    rd.seed = *static_cast<const uint64_t*>(P(4));
    rd.accelMode = u[8]; rd.parallelHint = u[9]; rd.threads = u[10]; rd.samplerType = u[11];
    for(int k = 0; k < 4; k++) rd.region[k] = u[12 + k];
    rd.latency = u[16]; rd.burstSize = u[17]; rd.camSwitchAfter = u[18];
    std::memcpy(rd.camSwitch, u + 19, sizeof(rd.camSwitch));
    if(bytes[3] >= 30 * 4) { rd.filmFilter = u[28]; std::memcpy(&rd.filmFilterRadius, u + 29, 4); }
    if(bytes[3] >= 33 * 4) { rd.genMips = u[30]; rd.mipGenFilter = u[31]; std::memcpy(&rd.mipGenFilterRadius, u + 32, 4); }   // TracerParameters.genMips / mipGenFilter

    char self[PATH_MAX]; ssize_t k = readlink("/proc/self/exe", self, sizeof(self) - 1);
    if(k <= 0) return 67;
    self[k] = 0;
    std::string drv = std::string(dirname(self)) + "/libtracer_driver.so";
    void* lib = dlopen(drv.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if(!lib) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 68; }
    auto render = reinterpret_cast<RenderF>(dlsym(lib, "tracer_driver_render"));
    if(!render) return 69;
    size_t pix = size_t(rd.width) * rd.height;
    std::vector<float> rgb(pix * 3), wgt(pix);
    DriverStats st{}; char err[1024] = {0};
    int rc = render(static_cast<const char*>(P(0)), &sc, &rd, rgb.data(), wgt.data(), &st, err, sizeof(err));
    if(rc != 0) { std::fprintf(stderr, "tracer driver failed (%d): %s\n", rc, err); return 70; }
    FILE* o = std::fopen(argv[2], "wb");
    if(!o) return 71;
    std::fwrite(rgb.data(), 4, rgb.size(), o); std::fwrite(wgt.data(), 4, wgt.size(), o);
    double s[4] = {st.commitSeconds, st.renderSeconds, st.totalPaths, double(st.iterations)};
    std::fwrite(s, 8, 4, o); std::fwrite(st.sceneAABB, 4, 6, o); std::fwrite(&st.startSeconds, 8, 1, o);
    std::fclose(o);
    return 0;
}
