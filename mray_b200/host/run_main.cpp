// run_main.cpp — `mray_b200_run`: the head-less "run" command (MRay/RunCommand.cpp:L143-700, MRay/TracerThread.cpp:L149-300,
// L855-876) rebuilt around the B200 plugin: load a TracerDLL and a scene loader by path, load the scene, commit, create the
// renderer, push its attributes, StartRender, loop DoRenderWork while accumulating the RenderImageSections it hands over
// (timeline-semaphore protocol; running weighted mean in fp64 like Accum::AccumulateScanline, RunCommand.cpp:L293-345), and
// save the image when the tracer says so. SURVEY.md §8(f) rank 4. Works with ANY TracerI (the tests also run the unmodified
// reference tracer through it). Output: PFM (rows bottom to top, exactly as the tracer delivers them) — the reference writes
// EXR through OpenImageIO, which this image does not have.
//
//   mray_b200_run --tracer libTracerDLL_B200.so --loader libSceneLoaderB200.so --scene scene.json -r 512x512 --spp 64
//                 [--renderer PathTracerRGB|PathTracerSpectral] [--sampleMode Pure|WithNextEventEstimation|WithNEEAndMIS]
//                 [--rr 3,8] [--seed 0] [--sampler Independent|Sobol|ZSobol] [--renderMode Throughput|Latency] [--burst 1]
//                 [--filter Gaussian,1.0] [--genMips [Gaussian,2.0]] [--hint 2097152] [-t threads] [--camera 0] --out image.pfm
#include "Core/TracerI.h"
#include "Core/SceneLoaderI.h"
#include "Core/ThreadPool.h"
#include "Core/TimelineSemaphore.h"
#include "Core/Error.h"
#include "TransientPool/TransientPool.h"

#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace
{
using namespace std::string_literals;

struct Options
{
    std::string tracer, loader, scene, out = "out.pfm";
    std::string renderer = "PathTracerRGB", sampleMode = "WithNEEAndMIS", sampler = "Independent", renderMode = "Throughput", filter = "Gaussian";
    uint32_t width = 512, height = 512, spp = 64, rrLo = 3, rrHi = 8, burst = 1, threads = 0, hint = 0, camera = 0;
    float filterRadius = 1.0f;
    uint64_t seed = 0;
    bool genMips = false; std::string mipFilter = "Gaussian"; float mipFilterRadius = 2.0f;   // tracer config "genMipmaps" / "mipGenFilter"
};

bool Parse(int argc, char** argv, Options& o, std::string& err)
{
    auto Need = [&](int& i) -> const char* { if(i + 1 >= argc) { err = "missing value of "s + argv[i]; return nullptr; } return argv[++i]; };
    for(int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        const char* v = nullptr;
        if(a == "--tracer") { if(!(v = Need(i))) return false; o.tracer = v; }
        else if(a == "--loader") { if(!(v = Need(i))) return false; o.loader = v; }
        else if(a == "--scene" || a == "-s") { if(!(v = Need(i))) return false; o.scene = v; }
        else if(a == "--out" || a == "-o") { if(!(v = Need(i))) return false; o.out = v; }
        else if(a == "--renderer") { if(!(v = Need(i))) return false; o.renderer = v; }
        else if(a == "--sampleMode") { if(!(v = Need(i))) return false; o.sampleMode = v; }
        else if(a == "--sampler") { if(!(v = Need(i))) return false; o.sampler = v; }
        else if(a == "--renderMode") { if(!(v = Need(i))) return false; o.renderMode = v; }
        else if(a == "--resolution" || a == "-r") { if(!(v = Need(i))) return false; if(std::sscanf(v, "%ux%u", &o.width, &o.height) != 2) { err = "resolution must be WxH"; return false; } }
        else if(a == "--spp") { if(!(v = Need(i))) return false; o.spp = uint32_t(std::strtoul(v, nullptr, 10)); }
        else if(a == "--rr") { if(!(v = Need(i))) return false; if(std::sscanf(v, "%u,%u", &o.rrLo, &o.rrHi) != 2) { err = "rr must be lo,hi"; return false; } }
        else if(a == "--seed") { if(!(v = Need(i))) return false; o.seed = std::strtoull(v, nullptr, 10); }
        else if(a == "--burst") { if(!(v = Need(i))) return false; o.burst = uint32_t(std::strtoul(v, nullptr, 10)); }
        else if(a == "--threads" || a == "-t") { if(!(v = Need(i))) return false; o.threads = uint32_t(std::strtoul(v, nullptr, 10)); }
        else if(a == "--hint") { if(!(v = Need(i))) return false; o.hint = uint32_t(std::strtoul(v, nullptr, 10)); }
        else if(a == "--camera") { if(!(v = Need(i))) return false; o.camera = uint32_t(std::strtoul(v, nullptr, 10)); }
        else if(a == "--filter")
        {
            if(!(v = Need(i))) return false;
            std::string s = v; size_t c = s.find(',');
            o.filter = s.substr(0, c);
            if(c != std::string::npos) o.filterRadius = std::strtof(s.c_str() + c + 1, nullptr);
        }
        else if(a == "--genMips")
        {   // TracerParameters.genMips, optionally with mipGenFilter as NAME,RADIUS (default Gaussian,2 as the reference)
            o.genMips = true;
            if(i + 1 < argc && argv[i + 1][0] != '-')
            {
                std::string s = argv[++i]; size_t c = s.find(',');
                o.mipFilter = s.substr(0, c);
                if(c != std::string::npos) o.mipFilterRadius = std::strtof(s.c_str() + c + 1, nullptr);
            }
        }
        else { err = "unknown option " + a; return false; }
    }
    if(o.tracer.empty() || o.loader.empty() || o.scene.empty()) { err = "--tracer, --loader and --scene are required"; return false; }
    return true;
}

bool WritePFM(const std::string& path, const std::vector<double>& rgbw, uint32_t w, uint32_t h)
{
    std::ofstream f(path, std::ios::binary);
    if(!f) return false;
    f << "PF\n" << w << " " << h << "\n-1.0\n";
    std::vector<float> row(size_t(w) * 3);
    for(uint32_t y = 0; y < h; y++)
    {
        for(uint32_t x = 0; x < w; x++)
            for(int c = 0; c < 3; c++) row[size_t(x) * 3 + c] = float(rgbw[(size_t(y) * w + x) * 4 + c]);
        f.write(reinterpret_cast<const char*>(row.data()), std::streamsize(row.size() * 4));
    }
    return bool(f);
}

} // namespace

int main(int argc, char** argv)
{
    Options o; std::string err;
    if(!Parse(argc, argv, o, err)) { std::fprintf(stderr, "mray_b200_run: %s\n", err.c_str()); return 64; }
    void* tlib = dlopen(o.tracer.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if(!tlib) { std::fprintf(stderr, "mray_b200_run: %s\n", dlerror()); return 65; }
    void* llib = dlopen(o.loader.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if(!llib) { std::fprintf(stderr, "mray_b200_run: %s\n", dlerror()); return 65; }
    using ConstructT = TracerI* (*)(const TracerParameters&);
    using DestroyT = void (*)(TracerI*);
    using ConstructL = SceneLoaderI* (*)(ThreadPool&);
    using DestroyL = void (*)(SceneLoaderI*);
    auto constructT = reinterpret_cast<ConstructT>(dlsym(tlib, "ConstructTracer"));
    auto destroyT = reinterpret_cast<DestroyT>(dlsym(tlib, "DestroyTracer"));
    auto constructL = reinterpret_cast<ConstructL>(dlsym(llib, "ConstructSceneLoaderMRay"));
    auto destroyL = reinterpret_cast<DestroyL>(dlsym(llib, "DestroySceneLoaderMRay"));
    if(!constructT || !destroyT || !constructL || !destroyL) { std::fprintf(stderr, "mray_b200_run: entry points not exported\n"); return 66; }

    TracerI* tracer = nullptr; SceneLoaderI* loader = nullptr;
    int rc = 0;
    try
    {
        const auto t0 = std::chrono::steady_clock::now();
        // tracer configuration (the "Parameters" object of the tracer-config JSON, TracerThread.cpp:L79-147)
        TracerParameters tp;
        tp.seed = o.seed;
        tp.accelMode = AcceleratorType::SOFTWARE_BASIC_BVH;
        if(o.hint) tp.parallelizationHint = o.hint;
        static const std::map<std::string, SamplerType::E> samplers = {{"Independent", SamplerType::INDEPENDENT}, {"Sobol", SamplerType::SOBOL}, {"ZSobol", SamplerType::Z_SOBOL}};
        static const std::map<std::string, FilterType::E> filters = {{"Box", FilterType::BOX}, {"Tent", FilterType::TENT}, {"Gaussian", FilterType::GAUSSIAN},
                                                                      {"Mitchell-Netravali", FilterType::MITCHELL_NETRAVALI}};
        if(!samplers.count(o.sampler)) throw MRayError("unknown sampler \"{}\"", o.sampler);
        if(!filters.count(o.filter)) throw MRayError("unknown film filter \"{}\"", o.filter);
        tp.samplerType = samplers.at(o.sampler);
        tp.filmFilter.type = filters.at(o.filter); tp.filmFilter.radius = o.filterRadius;
        tp.genMips = o.genMips; tp.mipGenFilter.type = filters.at(o.mipFilter); tp.mipGenFilter.radius = o.mipFilterRadius;
        tracer = constructT(tp);
        ThreadPool pool;
        auto threadInit = tracer->GetThreadInitFunction();
        pool.RestartThreads(o.threads ? o.threads : std::thread::hardware_concurrency(),
                            [threadInit](std::thread::native_handle_type, uint32_t) { threadInit(); });
        tracer->SetThreadPool(pool);
        threadInit();

        loader = constructL(pool);
        Expected<TracerIdPack> loaded = loader->LoadScene(*tracer, o.scene);
        if(loaded.has_error()) throw loaded.error();
        const TracerIdPack& pack = loaded.value();
        const auto t1 = std::chrono::steady_clock::now();
        SurfaceCommitResult cr = tracer->CommitSurfaces();
        const auto t2 = std::chrono::steady_clock::now();
        if(o.camera >= pack.camSurfaces.size()) throw MRayError("camera surface {} does not exist", o.camera);
        const CamSurfaceId camSurf = pack.camSurfaces[o.camera].second;

        // renderer + its attributes (the render-config JSON's "initialName" / attribute object, TracerThread.cpp:L160-231)
        TimelineSemaphore sem(0);
        tracer->SetupRenderEnv(&sem, 4096, 0);
        RendererId rid = tracer->CreateRenderer("(R)"s + o.renderer);
        RendererAttributeInfoList rInfo = tracer->AttributeInfo(rid);
        for(uint32_t a = 0; a < rInfo.size(); a++)
        {
            const std::string_view name = rInfo[a].name;
            auto PushU32 = [&](uint32_t v)
            { TransientData d(std::in_place_type_t<uint32_t>{}, 1); d.Push(Span<const uint32_t>(&v, 1)); tracer->PushRendererAttribute(rid, a, std::move(d)); };
            auto PushStr = [&](std::string_view s)
            {
                TransientData d = AllocateTransientData(rInfo[a].dataType, s.size());
                d.ReserveAll();
                Span<char> out = d.AccessAsString();
                std::copy(s.cbegin(), s.cend(), out.begin());
                tracer->PushRendererAttribute(rid, a, std::move(d));
            };
            if(name == "totalSPP") PushU32(o.spp);
            else if(name == "burstSize") PushU32(o.burst ? o.burst : 1u);
            else if(name == "renderMode") PushStr(o.renderMode);
            else if(name == "sampleMode") PushStr(o.sampleMode);
            else if(name == "neeSamplerType") PushStr("Uniform");
            else if(name == "rrRange")
            {
                Vector2ui v(o.rrLo, o.rrHi);
                TransientData d(std::in_place_type_t<Vector2ui>{}, 1); d.Push(Span<const Vector2ui>(&v, 1));
                tracer->PushRendererAttribute(rid, a, std::move(d));
            }
            else if(rInfo[a].isOptional == AttributeOptionality::MR_MANDATORY) throw MRayError("unknown mandatory renderer attribute {}", name);
        }
        RenderImageParams rip{Vector2ui(o.width, o.height), Vector2ui(0, 0), Vector2ui(o.width, o.height)};
        RenderBufferInfo rbi = tracer->StartRender(rid, camSurf, rip, std::nullopt, std::nullopt);
        const auto t3 = std::chrono::steady_clock::now();

        // accumulation: out = (out W + in) / (W + w), W += w (Accum::AccumulateScanline)
        const size_t pix = size_t(o.width) * o.height;
        std::vector<double> img(pix * 4, 0.0);
        uint64_t iterations = 0; double paths = 0.0;
        for(;;)
        {
            RendererOutput out = tracer->DoRenderWork();
            iterations++;
            if(out.imageOut)
            {
                const RenderImageSection& s = *out.imageOut;
                if(!sem.Acquire(s.waitCounter)) break;
                const float* R = reinterpret_cast<const float*>(rbi.data + s.pixStartOffsets[0]);
                const float* G = reinterpret_cast<const float*>(rbi.data + s.pixStartOffsets[1]);
                const float* B = reinterpret_cast<const float*>(rbi.data + s.pixStartOffsets[2]);
                const float* W = reinterpret_cast<const float*>(rbi.data + s.weightStartOffset);
                const uint32_t w = s.pixelMax[0] - s.pixelMin[0], h = s.pixelMax[1] - s.pixelMin[1];
                for(uint32_t y = 0; y < h; y++)
                for(uint32_t x = 0; x < w; x++)
                {
                    const size_t src = size_t(y) * w + x;
                    double* d = img.data() + (size_t(y + s.pixelMin[1]) * o.width + (x + s.pixelMin[0])) * 4;
                    const double wIn = double(W[src]) * s.globalWeight, wTot = d[3] + wIn;
                    if(wTot != 0.0)
                    {
                        d[0] = (d[0] * d[3] + R[src]) / wTot; d[1] = (d[1] * d[3] + G[src]) / wTot; d[2] = (d[2] * d[3] + B[src]) / wTot;
                    }
                    d[3] = wTot; paths += wIn;
                }
                sem.Release();
            }
            if(out.triggerSave) break;
        }
        const auto t4 = std::chrono::steady_clock::now();
        tracer->StopRender();
        if(!WritePFM(o.out, img, o.width, o.height)) throw MRayError("unable to write \"{}\"", o.out);
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::printf("{\"scene_load_ms\": %.2f, \"commit_ms\": %.2f, \"start_ms\": %.2f, \"render_ms\": %.2f, \"iterations\": %llu, \"paths\": %.0f, "
                    "\"surfaces\": %zu, \"instances\": %zu, \"accelerators\": %u, \"aabb\": [%g, %g, %g, %g, %g, %g], \"out\": \"%s\"}\n",
                    ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), (unsigned long long)iterations, paths, pack.surfaces.size(), size_t(cr.instanceCount),
                    unsigned(cr.acceleratorCount), cr.aabb.Min()[0], cr.aabb.Min()[1], cr.aabb.Min()[2], cr.aabb.Max()[0], cr.aabb.Max()[1], cr.aabb.Max()[2],
                    o.out.c_str());
        tracer->DestroyRenderer(rid);
    }
    catch(const MRayError& e) { std::fprintf(stderr, "mray_b200_run: %s\n", e.GetError().c_str()); rc = 70; }
    catch(const std::exception& e) { std::fprintf(stderr, "mray_b200_run: %s\n", e.what()); rc = 71; }
    if(loader) destroyL(loader);
    if(tracer) destroyT(tracer);
    return rc;
}
