// capi.cu — the C-ABI of include/mray_b200.h. Thin: argument checks, host<->device staging for the
// MRB_MEM_HOST variants, exception -> status translation. No CPU fallback anywhere.
#include "accel.cuh"
#include "spectrum.cuh"
#include "dist2d.cuh"
#include <cstring>
#include <initializer_list>
#include <vector>
#include <cstdio>
#include <new>
#include <stdexcept>
#include <algorithm>

namespace mrb
{
void BuildAccel(Context& ctx, mrb_accel_t& acc, const mrb_accel_desc& desc);
void TraceRays(Context& ctx, const mrb_accel_t& acc, bool anyHit, mrb_trace_mode mode,
               mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
               mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount);
}

struct mrb_render_data_fwd;
namespace mrb
{
void CreateRenderer(Context& ctx, mrb_renderer_t& r, const mrb_render_desc& desc);
void RenderIterate(Context& ctx, mrb_renderer_t& r, uint32_t iterations);
void DestroyRenderer(Context& ctx, mrb_renderer_t* r);
mrb_renderer_t* NewRenderer();
void RendererStats(Context& ctx, mrb_renderer_t& r, mrb_render_stats& out);
void RendererReadFilm(Context& ctx, mrb_renderer_t& r, float* out, bool device, bool clear);
float* RendererFilmPtr(mrb_renderer_t& r);
void RendererSetSppLimit(Context& ctx, mrb_renderer_t& r, uint32_t sppLimit);
void TextureSampleHost(Context& ctx, const mrb_texture_desc& td, const float* uv, uint32_t n, float* rgbOut);
void RendererBeginPass(Context& ctx, mrb_renderer_t& r, const uint32_t regionMin[2], const uint32_t regionSize[2],
                       uint32_t sampleStart, uint32_t sampleCount);
void RendererRunPass(Context& ctx, mrb_renderer_t& r, uint32_t chunk, mrb_render_stats& out);
void RendererPollStats(Context& ctx, mrb_renderer_t& r, mrb_render_stats& out);
void RendererFilmHandoff(Context& ctx, mrb_renderer_t& r, float* hostDst, void (*onComplete)(void*), void* user);
void RendererReducePeers(Context& ctx, mrb_renderer_t& r, Context* const* peerCtx, mrb_renderer_t* const* peers, uint32_t n);
void FilterSampleHost(Context& ctx, uint32_t type, float radius, const float* xi, uint32_t n, float* out);
}

namespace mrb
{
size_t MultiPartitionTempBytes(uint32_t count, uint32_t batchBits);
void MultiPartition(Context& ctx, uint32_t* keys, uint32_t* indices, uint32_t count,
                    const uint32_t dataBits[2], const uint32_t batchBits[2], bool onlySortForBatches,
                    uint32_t maxPartitions, uint32_t* outCount, uint32_t* outOffsets, uint32_t* outKeys, void* temp);
size_t BinaryPartitionTempBytes(uint32_t count);
void BinaryPartition(Context& ctx, uint32_t* indicesOut, uint32_t* leftCount, const uint32_t* indicesIn,
                     const uint8_t* flags, uint32_t count, void* temp);
}

namespace mrb
{
void BuildScene(Context& ctx, mrb_scene_t& sc, const mrb_instance_desc* inst, uint32_t n);
void SamplerGenerate(Context& ctx, uint32_t type, const uint32_t* matrices, const uint32_t* seeds, uint32_t width, uint32_t height,
                     uint32_t sampleIndex, uint32_t initialMaxSPP, uint32_t dimStart, uint64_t requests, uint32_t requestCount, uint32_t* out);
void CreateSpectrum(Context& ctx, mrb_spectrum_t& sp, const mrb_spectrum_desc& desc);
void SpectrumSampleWavelengths(Context& ctx, const mrb_spectrum_t& sp, const uint32_t* randoms, uint32_t n, float* waves, float* pdfs);
void SpectrumToRGB(Context& ctx, const mrb_spectrum_t& sp, float* values, const float* waves, const float* pdfs, uint32_t n);
void SpectrumUpsample(Context& ctx, const mrb_spectrum_t& sp, const float* rgb, uint32_t rgbStride, const float* waves, uint32_t n,
                      bool isRadiance, float* out);
void TextureConvertHost(Context& ctx, const mrb_texture_desc& td, void* texelsOut);
void TextureMipChainHost(Context& ctx, const mrb_texture_desc& td, void* chainOut, uint32_t* mipCountOut);
void TextureSampleLodHost(Context& ctx, const mrb_texture_desc& td, const float* uv, const float* lod, const float* grads, uint32_t lodMode, uint32_t n, float* rgbOut);
size_t TextureChainTexels(uint32_t w, uint32_t h, uint32_t mips);
uint32_t TextureFullMipCount(uint32_t w, uint32_t h);
void TextureFinalExtent(const mrb_texture_desc& td, uint32_t out[3]);
void GenerateSpectraLUT(Context& ctx, const float* cieXYZ, const float* illuminantSPD, float illuminantNorm, const float rgbToXYZ[9],
                        const float xyzToRGB[9], uint32_t res, uint32_t passes, float* lutOut, double whitepointOut[3]);
void TraceScene(Context& ctx, const SceneData& scn, bool anyHit, mrb_trace_mode mode,
                mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
                mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount);
}

struct mrb_context_t { mrb::Context c; };

static thread_local std::string gCreateError;

template<class F>
static mrb_status Guard(mrb_context ctx, F&& f)
{
    if(!ctx) return MRB_ERR_INVALID_ARG;
    try
    {
        cudaError_t e = cudaSetDevice(ctx->c.device);
        if(e != cudaSuccess) throw mrb::CudaError{e, __FILE__, __LINE__};
        return f(ctx->c);
    }
    catch(const mrb::CudaError& e)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d", int(e.code), cudaGetErrorString(e.code), e.file, e.line);
        ctx->c.error = buf;
        cudaGetLastError();
        return (e.code == cudaErrorMemoryAllocation) ? MRB_ERR_OUT_OF_MEMORY : MRB_ERR_CUDA;
    }
    catch(const std::bad_alloc&) { ctx->c.error = "host allocation failed"; return MRB_ERR_OUT_OF_MEMORY; }
    catch(const std::exception& e) { ctx->c.error = e.what(); return MRB_ERR_INVALID_ARG; }   // descriptor validation of the builders
}

static mrb_status Fail(mrb::Context& c, mrb_status s, const char* msg) { c.error = msg; return s; }

extern "C"
{

uint32_t mrb_abi_version(void) { return MRB_ABI_VERSION; }   // 8: texture mip chains + ray cones (mrb_texture_desc / mrb_render_desc fields); 7: texture colour conversion fields; 6: normal maps; 5: alpha maps in mrb_accel_desc; 4: boundary light (skysphere) fields of mrb_render_desc, mrb_dist2d_*

mrb_status mrb_context_create(int device, mrb_context* out)
{
    if(!out) return MRB_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if(e != cudaSuccess || count == 0)
    {
        gCreateError = std::string("no CUDA device: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return MRB_ERR_NO_DEVICE; // the product path refuses to run without a GPU
    }
    if(device < 0 || device >= count) { gCreateError = "device index out of range"; return MRB_ERR_INVALID_ARG; }
    mrb_context ctx = new(std::nothrow) mrb_context_t();
    if(!ctx) return MRB_ERR_OUT_OF_MEMORY;
    ctx->c.device = device;
    try
    {
        MRB_CUDA_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        MRB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        ctx->c.smCount = prop.multiProcessorCount;
        ctx->c.totalMem = prop.totalGlobalMem;
        MRB_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->c.ownStream, cudaStreamNonBlocking));
        ctx->c.stream = ctx->c.ownStream;
        MRB_CUDA_TRY(cudaEventCreate(&ctx->c.ev0));
        MRB_CUDA_TRY(cudaEventCreate(&ctx->c.ev1));
    }
    catch(const mrb::CudaError& err)
    {
        gCreateError = std::string("context creation failed: ") + cudaGetErrorString(err.code);
        delete ctx;
        return MRB_ERR_CUDA;
    }
    *out = ctx;
    return MRB_OK;
}

void mrb_context_destroy(mrb_context ctx)
{
    if(!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    ctx->c.scratch.Free();
    ctx->c.traceScratch.Free();
    if(ctx->c.copyIn) cudaStreamDestroy(ctx->c.copyIn);
    if(ctx->c.copyOut) cudaStreamDestroy(ctx->c.copyOut);
    if(ctx->c.evStart) cudaEventDestroy(ctx->c.evStart);
    for(int k = 0; k < mrb::Context::PIPE_CHUNKS; k++)
    {
        if(ctx->c.evIn[k]) cudaEventDestroy(ctx->c.evIn[k]);
        if(ctx->c.evDone[k]) cudaEventDestroy(ctx->c.evDone[k]);
    }
    for(int k = 0; k < ctx->c.prof.created; k++) { cudaEventDestroy(ctx->c.prof.ev[k][0]); cudaEventDestroy(ctx->c.prof.ev[k][1]); }
    if(ctx->c.ev0) cudaEventDestroy(ctx->c.ev0);
    if(ctx->c.ev1) cudaEventDestroy(ctx->c.ev1);
    if(ctx->c.ownStream) cudaStreamDestroy(ctx->c.ownStream);
    delete ctx;
}

mrb_status mrb_context_set_stream(mrb_context ctx, void* cuda_stream)
{
    if(!ctx) return MRB_ERR_INVALID_ARG;
    // pass-through: NULL is the CUDA (legacy) default stream, exactly what torch hands out as its
    // default "current stream"; the context's own stream is only used until this is called
    ctx->c.stream = static_cast<cudaStream_t>(cuda_stream);
    return MRB_OK;
}

mrb_status mrb_context_synchronize(mrb_context ctx)
{
    return Guard(ctx, [&](mrb::Context& c) { MRB_CUDA_TRY(cudaStreamSynchronize(c.stream)); return MRB_OK; });
}

size_t mrb_context_used_device_memory(mrb_context ctx) { return ctx ? ctx->c.persistentBytes + ctx->c.scratch.Capacity() + ctx->c.traceScratch.Capacity() : 0; }
size_t mrb_context_total_device_memory(mrb_context ctx) { return ctx ? ctx->c.totalMem : 0; }
uint64_t mrb_context_launch_count(mrb_context ctx) { return ctx ? ctx->c.launches : 0; }

mrb_status mrb_context_last_fallback_count(mrb_context ctx, uint32_t* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        out[0] = out[1] = out[2] = out[3] = 0;
        if(!c.lastFallbackCount) return MRB_OK;
        uint32_t h[5];
        MRB_CUDA_TRY(cudaMemcpyAsync(h, c.lastFallbackCount, sizeof(uint32_t) * 5, cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; out[3] = h[4];
        return MRB_OK;
    });
}

const char* mrb_last_error(mrb_context ctx) { return ctx ? ctx->c.error.c_str() : gCreateError.c_str(); }

mrb_status mrb_context_set_alpha_seed(mrb_context ctx, uint32_t seed)
{
    return Guard(ctx, [&](mrb::Context& c) { c.alphaSeed = seed; return MRB_OK; });
}

mrb_status mrb_context_set_profiling(mrb_context ctx, int enabled, uint32_t iterationStride)
{
    if(!ctx) return MRB_ERR_INVALID_ARG;
    ctx->c.prof.enabled = enabled != 0;
    ctx->c.prof.stride = iterationStride ? iterationStride : 16u;
    return MRB_OK;
}

mrb_status mrb_context_get_profile(mrb_context ctx, mrb_kernel_profile* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::KernelProfile& p = c.prof;
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        for(int k = 0; k < p.used; k++)
        {
            float ms = 0.f;
            if(cudaEventElapsedTime(&ms, p.ev[k][0], p.ev[k][1]) == cudaSuccess) { p.ms[p.kind[k]] += ms; p.samples[p.kind[k]]++; }
        }
        cudaGetLastError();
        p.used = 0;
        for(int k = 0; k < mrb::PROF_KINDS; k++) { out->ms[k] = p.ms[k]; out->samples[k] = p.samples[k]; }
        return MRB_OK;
    });
}

mrb_status mrb_accel_build(mrb_context ctx, const mrb_accel_desc* desc, mrb_accel* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!desc || !out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if(!desc->positions || !desc->indices || desc->triangleCount == 0 || desc->vertexCount == 0)
            return Fail(c, MRB_ERR_INVALID_ARG, "empty primitive group");
        if(desc->rangeCount > (1u << 20)) return Fail(c, MRB_ERR_INVALID_ARG, "too many prim ranges");
        if(desc->primGroupId > 15) return Fail(c, MRB_ERR_INVALID_ARG, "primGroupId exceeds PrimitiveKey batch bits (4)");
        uint64_t leafs = 0;
        if(desc->rangeCount && desc->primRanges)
        {
            for(uint32_t i = 0; i < desc->rangeCount; i++)
            {
                uint32_t b = desc->primRanges[2 * i], e = desc->primRanges[2 * i + 1];
                if(e <= b || e > desc->triangleCount) return Fail(c, MRB_ERR_INVALID_ARG, "bad prim range");
                leafs += e - b;
            }
        }
        else leafs = desc->triangleCount;
        if(leafs >= (1ull << 28)) return Fail(c, MRB_ERR_INVALID_ARG, "leaf count exceeds PrimitiveKey index bits (28)");
        mrb_accel acc = new mrb_accel_t();
        acc->flags = desc->flags;
        try { mrb::BuildAccel(c, *acc, *desc); }
        catch(...) { c.persistentBytes -= acc->mem.Capacity(); delete acc; throw; }
        // The wide traversal keeps at most TWO stack entries per level of the path it is on (the postponed triangle
        // group and the rest of the node group, trace.cu) on a 64-entry stack: a tree of wideDepth levels needs
        // 2 * wideDepth entries. Deeper (pathologically skewed) trees are refused here instead of corrupting the stack.
        if(2u * acc->d.wideDepth > mrb::WIDE_STACK_ENTRIES)
        {
            c.persistentBytes -= acc->mem.Capacity(); delete acc;
            return Fail(c, MRB_ERR_UNSUPPORTED, "wide BVH deeper than the traversal stack (2 * depth > 64)");
        }
        *out = acc;
        return MRB_OK;
    });
}

void mrb_accel_destroy(mrb_context ctx, mrb_accel accel)
{
    if(!ctx || !accel) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    ctx->c.persistentBytes -= accel->mem.Capacity();
    delete accel;
}

mrb_status mrb_accel_get_info(mrb_context ctx, mrb_accel accel, mrb_accel_info* info)
{
    if(!ctx || !accel || !info) return MRB_ERR_INVALID_ARG;
    *info = accel->info;
    return MRB_OK;
}

mrb_status mrb_accel_export_lbvh(mrb_context ctx, mrb_accel accel,
                                 uint64_t* morton, uint64_t* sortedMorton, uint32_t* sortedLeaf,
                                 uint32_t* nodes, uint32_t* leafParent, float* nodeBoxes, float* leafAABBs)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!accel) return Fail(c, MRB_ERR_INVALID_ARG, "null accelerator");
        const mrb::AccelData& d = accel->d;
        auto D2H = [&](void* dst, const void* src, size_t bytes)
        {
            if(dst) MRB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
        };
        D2H(morton, d.morton, sizeof(uint64_t) * d.leafCount);
        D2H(sortedMorton, d.sortedMorton, sizeof(uint64_t) * d.leafCount);
        D2H(sortedLeaf, d.sortedLeaf, sizeof(uint32_t) * d.leafCount);
        D2H(nodes, d.nodes, sizeof(mrb::LBVHNode) * d.nodeCount);
        D2H(leafParent, d.leafParent, sizeof(uint32_t) * d.leafCount);
        D2H(nodeBoxes, d.boxes, sizeof(mrb::LBVHBox) * d.nodeCount);
        D2H(leafAABBs, d.leafAABB, sizeof(float) * 6 * d.leafCount);
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        return MRB_OK;
    });
}

mrb_status mrb_accel_export_wide(mrb_context ctx, mrb_accel accel, void* wideNodes, void* triRecords)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!accel) return Fail(c, MRB_ERR_INVALID_ARG, "null accelerator");
        const mrb::AccelData& d = accel->d;
        if(!d.wideNodes) return Fail(c, MRB_ERR_UNSUPPORTED, "accelerator was built BINARY_ONLY");
        if(wideNodes) MRB_CUDA_TRY(cudaMemcpyAsync(wideNodes, d.wideNodes, sizeof(mrb::WideNode) * d.wideNodeCount, cudaMemcpyDeviceToHost, c.stream));
        if(triRecords) MRB_CUDA_TRY(cudaMemcpyAsync(triRecords, d.tris, sizeof(mrb::TriRecord) * d.leafCount, cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        return MRB_OK;
    });
}

} // extern "C"

// Host-pointer casts (the boundary a TracerI host uses before its buffers live on the device): the rays are
// cut into chunks; chunk k+1 uploads and chunk k-1 downloads while chunk k is traced, on three streams, so
// the call costs max(H2D, D2H, trace) instead of their sum (PCIe is full duplex). Indirect casts
// (rayIndices) need every ray resident and take the unpipelined route.
template<class TraceF>
static mrb_status CastGeneric(mrb_context ctx, bool anyHit,
                              mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
                              mrb_ray_gmem* rays, const uint32_t* rayIndices,
                              uint32_t rayCount, uint32_t totalRayCount, mrb_memspace memspace, bool fresh, TraceF&& trace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!rays) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(anyHit ? !visibleBits : (!hitKeys || !metaHits)) return Fail(c, MRB_ERR_INVALID_ARG, "null output");
        if(totalRayCount < rayCount && !rayIndices) return Fail(c, MRB_ERR_INVALID_ARG, "totalRayCount < rayCount");
        if(rayCount == 0) return MRB_OK;
        const size_t words = (size_t(totalRayCount) + 31) / 32;
        // MRB_TRACE_FRESH_OUTPUTS: outputs are initialised on the device instead of being read from the caller
        auto Fresh = [&](mrb_hit_key_pack* k, mrb_meta_hit* h, uint32_t* b, size_t rayN, size_t wordN, cudaStream_t st)
        {
            if(b) MRB_CUDA_TRY(cudaMemsetAsync(b, 0xFF, sizeof(uint32_t) * wordN, st));
            if(k) MRB_CUDA_TRY(cudaMemsetAsync(k, 0xFF, sizeof(mrb_hit_key_pack) * rayN, st));
            if(h) MRB_CUDA_TRY(cudaMemsetAsync(h, 0, sizeof(mrb_meta_hit) * rayN, st));
        };
        if(memspace == MRB_MEM_DEVICE)
        {
            if(fresh) Fresh(anyHit ? nullptr : hitKeys, anyHit ? nullptr : metaHits, anyHit ? visibleBits : nullptr, totalRayCount, words, c.stream);
            trace(c, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount);
            return MRB_OK;
        }
        mrb::MultiAlloc sz(nullptr);
        sz.Take<mrb_ray_gmem>(totalRayCount); sz.Take<uint32_t>(rayIndices ? rayCount : 0);
        if(anyHit) sz.Take<uint32_t>(words); else { sz.Take<mrb_hit_key_pack>(totalRayCount); sz.Take<mrb_meta_hit>(totalRayCount); }
        c.scratch.Reserve(sz.Total());
        mrb::MultiAlloc ma(c.scratch.Base());
        mrb_ray_gmem* dRays = ma.Take<mrb_ray_gmem>(totalRayCount);
        uint32_t* dIdx = ma.Take<uint32_t>(rayIndices ? rayCount : 0);
        uint32_t* dBits = anyHit ? ma.Take<uint32_t>(words) : nullptr;
        mrb_hit_key_pack* dKeys = anyHit ? nullptr : ma.Take<mrb_hit_key_pack>(totalRayCount);
        mrb_meta_hit* dHits = anyHit ? nullptr : ma.Take<mrb_meta_hit>(totalRayCount);

        constexpr uint32_t MIN_CHUNK = 1u << 16;   // rays; chunk starts stay multiples of 32 (visibility words)
        if(!rayIndices && rayCount >= 2 * MIN_CHUNK)
        {
            if(!c.copyIn)
            {
                MRB_CUDA_TRY(cudaStreamCreateWithFlags(&c.copyIn, cudaStreamNonBlocking));
                MRB_CUDA_TRY(cudaStreamCreateWithFlags(&c.copyOut, cudaStreamNonBlocking));
                MRB_CUDA_TRY(cudaEventCreateWithFlags(&c.evStart, cudaEventDisableTiming));
                for(int k = 0; k < mrb::Context::PIPE_CHUNKS; k++)
                {
                    MRB_CUDA_TRY(cudaEventCreateWithFlags(&c.evIn[k], cudaEventDisableTiming));
                    MRB_CUDA_TRY(cudaEventCreateWithFlags(&c.evDone[k], cudaEventDisableTiming));
                }
            }
            uint32_t chunks = rayCount / MIN_CHUNK;
            if(chunks > uint32_t(mrb::Context::PIPE_CHUNKS)) chunks = uint32_t(mrb::Context::PIPE_CHUNKS);
            const uint32_t per = ((rayCount + chunks - 1) / chunks + 1023u) & ~1023u;
            // whatever the caller queued on the context stream (a build, a previous cast) comes first
            MRB_CUDA_TRY(cudaEventRecord(c.evStart, c.stream));
            MRB_CUDA_TRY(cudaStreamWaitEvent(c.copyIn, c.evStart, 0));
            MRB_CUDA_TRY(cudaStreamWaitEvent(c.copyOut, c.evStart, 0));
            uint32_t k = 0;
            for(uint32_t b = 0; b < rayCount; b += per, k++)
            {
                const uint32_t n = (rayCount - b < per) ? rayCount - b : per;
                const size_t w0 = b / 32, wn = (size_t(n) + 31) / 32;
                MRB_CUDA_TRY(cudaMemcpyAsync(dRays + b, rays + b, sizeof(mrb_ray_gmem) * n, cudaMemcpyHostToDevice, c.copyIn));
                if(fresh) Fresh(anyHit ? nullptr : dKeys + b, anyHit ? nullptr : dHits + b, anyHit ? dBits + w0 : nullptr, n, wn, c.copyIn);
                else if(anyHit) MRB_CUDA_TRY(cudaMemcpyAsync(dBits + w0, visibleBits + w0, sizeof(uint32_t) * wn, cudaMemcpyHostToDevice, c.copyIn));
                else
                {
                    MRB_CUDA_TRY(cudaMemcpyAsync(dKeys + b, hitKeys + b, sizeof(mrb_hit_key_pack) * n, cudaMemcpyHostToDevice, c.copyIn));
                    MRB_CUDA_TRY(cudaMemcpyAsync(dHits + b, metaHits + b, sizeof(mrb_meta_hit) * n, cudaMemcpyHostToDevice, c.copyIn));
                }
                MRB_CUDA_TRY(cudaEventRecord(c.evIn[k], c.copyIn));
                MRB_CUDA_TRY(cudaStreamWaitEvent(c.stream, c.evIn[k], 0));
                trace(c, anyHit ? nullptr : dKeys + b, anyHit ? nullptr : dHits + b, anyHit ? dBits + w0 : nullptr, dRays + b, nullptr, n);
                MRB_CUDA_TRY(cudaEventRecord(c.evDone[k], c.stream));
                MRB_CUDA_TRY(cudaStreamWaitEvent(c.copyOut, c.evDone[k], 0));
                if(anyHit) MRB_CUDA_TRY(cudaMemcpyAsync(visibleBits + w0, dBits + w0, sizeof(uint32_t) * wn, cudaMemcpyDeviceToHost, c.copyOut));
                else
                {
                    MRB_CUDA_TRY(cudaMemcpyAsync(hitKeys + b, dKeys + b, sizeof(mrb_hit_key_pack) * n, cudaMemcpyDeviceToHost, c.copyOut));
                    MRB_CUDA_TRY(cudaMemcpyAsync(metaHits + b, dHits + b, sizeof(mrb_meta_hit) * n, cudaMemcpyDeviceToHost, c.copyOut));
                    MRB_CUDA_TRY(cudaMemcpyAsync(rays + b, dRays + b, sizeof(mrb_ray_gmem) * n, cudaMemcpyDeviceToHost, c.copyOut));
                }
            }
            MRB_CUDA_TRY(cudaStreamSynchronize(c.copyOut));
            MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
            return MRB_OK;
        }

        MRB_CUDA_TRY(cudaMemcpyAsync(dRays, rays, sizeof(mrb_ray_gmem) * totalRayCount, cudaMemcpyHostToDevice, c.stream));
        if(rayIndices) MRB_CUDA_TRY(cudaMemcpyAsync(dIdx, rayIndices, sizeof(uint32_t) * rayCount, cudaMemcpyHostToDevice, c.stream));
        if(fresh) Fresh(dKeys, dHits, dBits, totalRayCount, words, c.stream);
        if(anyHit)
        {
            if(!fresh) MRB_CUDA_TRY(cudaMemcpyAsync(dBits, visibleBits, sizeof(uint32_t) * words, cudaMemcpyHostToDevice, c.stream));
            trace(c, nullptr, nullptr, dBits, dRays, rayIndices ? dIdx : nullptr, rayCount);
            MRB_CUDA_TRY(cudaMemcpyAsync(visibleBits, dBits, sizeof(uint32_t) * words, cudaMemcpyDeviceToHost, c.stream));
        }
        else
        {
            if(!fresh)
            {
                MRB_CUDA_TRY(cudaMemcpyAsync(dKeys, hitKeys, sizeof(mrb_hit_key_pack) * totalRayCount, cudaMemcpyHostToDevice, c.stream));
                MRB_CUDA_TRY(cudaMemcpyAsync(dHits, metaHits, sizeof(mrb_meta_hit) * totalRayCount, cudaMemcpyHostToDevice, c.stream));
            }
            trace(c, dKeys, dHits, nullptr, dRays, rayIndices ? dIdx : nullptr, rayCount);
            MRB_CUDA_TRY(cudaMemcpyAsync(hitKeys, dKeys, sizeof(mrb_hit_key_pack) * totalRayCount, cudaMemcpyDeviceToHost, c.stream));
            MRB_CUDA_TRY(cudaMemcpyAsync(metaHits, dHits, sizeof(mrb_meta_hit) * totalRayCount, cudaMemcpyDeviceToHost, c.stream));
            MRB_CUDA_TRY(cudaMemcpyAsync(rays, dRays, sizeof(mrb_ray_gmem) * totalRayCount, cudaMemcpyDeviceToHost, c.stream));
        }
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        return MRB_OK;
    });
}

static mrb_status CastCommon(mrb_context ctx, mrb_accel accel, bool anyHit,
                             mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits, uint32_t* visibleBits,
                             mrb_ray_gmem* rays, const uint32_t* rayIndices,
                             uint32_t rayCount, uint32_t totalRayCount,
                             mrb_memspace memspace, mrb_trace_mode mode)
{
    const bool fresh = (uint32_t(mode) & uint32_t(MRB_TRACE_FRESH_OUTPUTS)) != 0u;
    mode = mrb_trace_mode(uint32_t(mode) & 0xFFu);
    if(ctx && !accel) return Guard(ctx, [&](mrb::Context& c) { return Fail(c, MRB_ERR_INVALID_ARG, "null argument"); });
    if(ctx && mode == MRB_TRACE_WIDE && !accel->d.wideNodes)
        return Guard(ctx, [&](mrb::Context& c) { return Fail(c, MRB_ERR_UNSUPPORTED, "accelerator was built BINARY_ONLY"); });
    return CastGeneric(ctx, anyHit, hitKeys, metaHits, visibleBits, rays, rayIndices, rayCount, totalRayCount, memspace, fresh,
                       [&](mrb::Context& c, mrb_hit_key_pack* k, mrb_meta_hit* h, uint32_t* b, mrb_ray_gmem* r, const uint32_t* idx, uint32_t n)
                       { mrb::TraceRays(c, *accel, anyHit, mode, k, h, b, r, idx, n); });
}

extern "C"
{

mrb_status mrb_cast_rays(mrb_context ctx, mrb_accel accel, mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits,
                         mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount, uint32_t totalRayCount,
                         mrb_memspace memspace, mrb_trace_mode mode)
{
    return CastCommon(ctx, accel, false, hitKeys, metaHits, nullptr, rays, rayIndices, rayCount, totalRayCount, memspace, mode);
}

mrb_status mrb_cast_visibility_rays(mrb_context ctx, mrb_accel accel, uint32_t* isVisibleBits,
                                    const mrb_ray_gmem* rays, const uint32_t* rayIndices,
                                    uint32_t rayCount, uint32_t totalRayCount,
                                    mrb_memspace memspace, mrb_trace_mode mode)
{
    return CastCommon(ctx, accel, true, nullptr, nullptr, isVisibleBits, const_cast<mrb_ray_gmem*>(rays), rayIndices,
                      rayCount, totalRayCount, memspace, mode);
}

} // extern "C"

template<class K>
static mrb_status SortCommon(mrb_context ctx, K* keys, uint32_t* values, uint32_t count,
                             uint32_t bitBegin, uint32_t bitEnd, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(bitEnd > sizeof(K) * 8 || bitBegin > bitEnd) return Fail(c, MRB_ERR_INVALID_ARG, "bad bit range");
        if(count == 0) return MRB_OK;
        if(!keys || !values) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(count >= (1u << 30)) return Fail(c, MRB_ERR_UNSUPPORTED, "count >= 2^30");
        size_t tempBytes = mrb::RadixSortTempBytes(count, sizeof(K));
        mrb::MultiAlloc sz(nullptr);
        sz.Take<char>(tempBytes); sz.Take<K>(count); sz.Take<uint32_t>(count);
        c.scratch.Reserve(sz.Total());
        mrb::MultiAlloc ma(c.scratch.Base());
        void* temp = ma.Take<char>(tempBytes);
        if(memspace == MRB_MEM_DEVICE) { mrb::RadixSortPairs(c, keys, values, count, bitBegin, bitEnd, temp); return MRB_OK; }
        K* dk = ma.Take<K>(count); uint32_t* dv = ma.Take<uint32_t>(count);
        MRB_CUDA_TRY(cudaMemcpyAsync(dk, keys, sizeof(K) * count, cudaMemcpyHostToDevice, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(dv, values, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, c.stream));
        mrb::RadixSortPairs(c, dk, dv, count, bitBegin, bitEnd, temp);
        MRB_CUDA_TRY(cudaMemcpyAsync(keys, dk, sizeof(K) * count, cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(values, dv, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        return MRB_OK;
    });
}

extern "C"
{

mrb_status mrb_radix_sort_pairs_u64(mrb_context ctx, uint64_t* keys, uint32_t* values, uint32_t count,
                                    uint32_t bitBegin, uint32_t bitEnd, mrb_memspace memspace)
{
    return SortCommon<uint64_t>(ctx, keys, values, count, bitBegin, bitEnd, memspace);
}

mrb_status mrb_radix_sort_pairs_u32(mrb_context ctx, uint32_t* keys, uint32_t* values, uint32_t count,
                                    uint32_t bitBegin, uint32_t bitEnd, mrb_memspace memspace)
{
    return SortCommon<uint32_t>(ctx, keys, values, count, bitBegin, bitEnd, memspace);
}


mrb_status mrb_spectrum_create(mrb_context ctx, const mrb_spectrum_desc* desc, mrb_spectrum* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!desc || !out || !desc->lut || !desc->observerXYZ || !desc->illuminantSPD) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if(desc->lutResolution != 64) return Fail(c, MRB_ERR_INVALID_ARG, "Wrong size, spectra lut size must be 64!");
        if(desc->wavelengthSampleMode > 2) return Fail(c, MRB_ERR_INVALID_ARG, "Unkown wavelength sample mode!");
        mrb_spectrum sp = new mrb_spectrum_t();
        try { mrb::CreateSpectrum(c, *sp, *desc); }
        catch(...) { c.persistentBytes -= sp->mem.Capacity(); delete sp; throw; }
        *out = sp;
        return MRB_OK;
    });
}

void mrb_spectrum_destroy(mrb_context ctx, mrb_spectrum spectrum)
{
    if(!ctx || !spectrum) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    ctx->c.persistentBytes -= spectrum->mem.Capacity();
    delete spectrum;
}

} // extern "C"

// Runs `work(devicePointers...)` on device copies of host arrays (or directly on device arrays).
struct StagedArray { void* user; size_t bytes; bool in, out; void* dev; };
template<class F>
static mrb_status Staged(mrb::Context& c, mrb_memspace memspace, std::initializer_list<StagedArray> list, F&& work)
{
    std::vector<StagedArray> a(list);
    if(memspace == MRB_MEM_DEVICE) { for(auto& x : a) x.dev = x.user; work(a); return MRB_OK; }
    mrb::MultiAlloc sz(nullptr);
    for(auto& x : a) sz.Take<char>(mrb::AlignUp(x.bytes, 256));
    c.scratch.Reserve(sz.Total());
    mrb::MultiAlloc ma(c.scratch.Base());
    for(auto& x : a)
    {
        x.dev = ma.Take<char>(mrb::AlignUp(x.bytes, 256));
        if(x.in && x.bytes) MRB_CUDA_TRY(cudaMemcpyAsync(x.dev, x.user, x.bytes, cudaMemcpyHostToDevice, c.stream));
    }
    work(a);
    for(auto& x : a) if(x.out && x.bytes) MRB_CUDA_TRY(cudaMemcpyAsync(x.user, x.dev, x.bytes, cudaMemcpyDeviceToHost, c.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
    return MRB_OK;
}

extern "C" {

mrb_status mrb_spectrum_sample_wavelengths(mrb_context ctx, mrb_spectrum spectrum, float* waves, float* pdfs,
                                           const uint32_t* randomNumbers, uint32_t count, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!spectrum || !waves || !pdfs || !randomNumbers) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(count == 0) return MRB_OK;
        return Staged(c, memspace, {{waves, size_t(count) * 16, false, true, nullptr}, {pdfs, size_t(count) * 16, false, true, nullptr},
                                    {const_cast<uint32_t*>(randomNumbers), size_t(count) * 4, true, false, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      { mrb::SpectrumSampleWavelengths(c, *spectrum, static_cast<const uint32_t*>(a[2].dev), count,
                                                       static_cast<float*>(a[0].dev), static_cast<float*>(a[1].dev)); });
    });
}

mrb_status mrb_spectrum_convert_to_rgb(mrb_context ctx, mrb_spectrum spectrum, float* values, const float* waves,
                                       const float* pdfs, uint32_t count, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!spectrum || !values || !waves || !pdfs) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(count == 0) return MRB_OK;
        return Staged(c, memspace, {{values, size_t(count) * 16, true, true, nullptr},
                                    {const_cast<float*>(waves), size_t(count) * 16, true, false, nullptr},
                                    {const_cast<float*>(pdfs), size_t(count) * 16, true, false, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      { mrb::SpectrumToRGB(c, *spectrum, static_cast<float*>(a[0].dev), static_cast<const float*>(a[1].dev),
                                           static_cast<const float*>(a[2].dev), count); });
    });
}

mrb_status mrb_spectrum_upsample(mrb_context ctx, mrb_spectrum spectrum, float* outSpectra, const float* rgb,
                                 int rgbIsUniform, const float* waves, uint32_t count, int isRadiance, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!spectrum || !outSpectra || !rgb || !waves) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(count == 0) return MRB_OK;
        const size_t rgbBytes = rgbIsUniform ? 12 : size_t(count) * 12;
        return Staged(c, memspace, {{outSpectra, size_t(count) * 16, false, true, nullptr},
                                    {const_cast<float*>(rgb), rgbBytes, true, false, nullptr},
                                    {const_cast<float*>(waves), size_t(count) * 16, true, false, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      { mrb::SpectrumUpsample(c, *spectrum, static_cast<const float*>(a[1].dev), rgbIsUniform ? 0u : 3u,
                                              static_cast<const float*>(a[2].dev), count, isRadiance != 0, static_cast<float*>(a[0].dev)); });
    });
}

mrb_status mrb_sampler_generate(mrb_context ctx, uint32_t samplerType, const uint32_t* sobolMatrices,
                                const uint32_t* generatorSeeds, uint32_t width, uint32_t height,
                                uint32_t sampleIndex, uint32_t initialMaxSPP, uint32_t dimensionStart,
                                const uint32_t* requestDims, uint32_t requestCount, uint32_t* numbersOut, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!sobolMatrices || !generatorSeeds || !requestDims || !numbersOut) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        const uint32_t baseType = samplerType & 0xFFu;
        if((baseType != MRB_SAMPLER_SOBOL && baseType != MRB_SAMPLER_ZSOBOL) || (samplerType & ~0x1FFu))
            return Fail(c, MRB_ERR_INVALID_ARG, "sampler type must be Sobol or ZSobol");
        if(requestCount > 32) return Fail(c, MRB_ERR_INVALID_ARG, "RNRequestList holds at most 32 requests");
        if(baseType == MRB_SAMPLER_ZSOBOL && initialMaxSPP == 0) return Fail(c, MRB_ERR_INVALID_ARG, "initialMaxSPP must be positive");
        uint64_t packed = 0; uint32_t total = 0;
        for(uint32_t r = 0; r < requestCount; r++)
        {
            if(requestDims[r] < 1 || requestDims[r] > 3) return Fail(c, MRB_ERR_INVALID_ARG, "a request draws 1, 2 or 3 dimensions");
            packed |= uint64_t(requestDims[r]) << (2u * r); total += requestDims[r];
        }
        const size_t n = size_t(width) * height;
        if(n == 0 || total == 0) return MRB_OK;
        return Staged(c, memspace, {{numbersOut, n * total * 4, false, true, nullptr},
                                    {const_cast<uint32_t*>(sobolMatrices), size_t(256) * 52 * 4, true, false, nullptr},
                                    {const_cast<uint32_t*>(generatorSeeds), n * 4, true, false, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      { mrb::SamplerGenerate(c, samplerType, static_cast<const uint32_t*>(a[1].dev), static_cast<const uint32_t*>(a[2].dev), width, height,
                                             sampleIndex, initialMaxSPP, dimensionStart, packed, requestCount, static_cast<uint32_t*>(a[0].dev)); });
    });
}

mrb_status mrb_renderer_create(mrb_context ctx, const mrb_render_desc* desc, mrb_renderer* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!desc || !out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if((desc->accel != nullptr) == (desc->scene != nullptr))
            return Fail(c, MRB_ERR_INVALID_ARG, "exactly one of accel / scene must be set");
        if(desc->accel && !desc->accel->d.wideNodes) return Fail(c, MRB_ERR_UNSUPPORTED, "accelerator was built BINARY_ONLY");
        if(desc->width == 0 || desc->height == 0 || desc->totalSPP == 0) return Fail(c, MRB_ERR_INVALID_ARG, "empty render");
        if(desc->sampleMode > 2) return Fail(c, MRB_ERR_INVALID_ARG, "unknown sampleMode");
        if(desc->rrRange[1] > 255) return Fail(c, MRB_ERR_INVALID_ARG, "rrRange[1] exceeds PathDataPack depth (u8)");
        if((desc->samplerType & 0xFFu) > 2 || (desc->samplerType & ~0x1FFu)) return Fail(c, MRB_ERR_INVALID_ARG, "unknown samplerType");
        if((desc->samplerType & 0xFFu) && !desc->sobolMatrices) return Fail(c, MRB_ERR_INVALID_ARG, "Sobol / ZSobol need sobolMatrices");
        if((desc->materialCount && !desc->albedo) || (desc->lightCount && !desc->lightRadiance))
            return Fail(c, MRB_ERR_INVALID_ARG, "missing material / light attributes");
        mrb_renderer r = mrb::NewRenderer();
        try { mrb::CreateRenderer(c, *r, *desc); }
        catch(...) { mrb::DestroyRenderer(c, r); throw; }
        *out = r;
        return MRB_OK;
    });
}

void mrb_renderer_destroy(mrb_context ctx, mrb_renderer r)
{
    if(!ctx || !r) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    mrb::DestroyRenderer(ctx->c, r);
}

mrb_status mrb_renderer_iterate(mrb_context ctx, mrb_renderer r, uint32_t iterations)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r) return Fail(c, MRB_ERR_INVALID_ARG, "null renderer");
        mrb::RenderIterate(c, *r, iterations);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_set_spp_limit(mrb_context ctx, mrb_renderer r, uint32_t sppLimit)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r) return Fail(c, MRB_ERR_INVALID_ARG, "null renderer");
        mrb::RendererSetSppLimit(c, *r, sppLimit);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_get_stats(mrb_context ctx, mrb_renderer r, mrb_render_stats* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || !out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::RendererStats(c, *r, *out);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_read_film(mrb_context ctx, mrb_renderer r, float* out, mrb_memspace memspace, int clear)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || !out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::RendererReadFilm(c, *r, out, memspace == MRB_MEM_DEVICE, clear != 0);
        return MRB_OK;
    });
}

float* mrb_renderer_film_device_ptr(mrb_renderer r) { return r ? mrb::RendererFilmPtr(*r) : nullptr; }

mrb_status mrb_renderer_begin_pass(mrb_context ctx, mrb_renderer r, const uint32_t regionMin[2], const uint32_t regionSize[2],
                                   uint32_t sampleStart, uint32_t sampleCount)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || !regionMin || !regionSize) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::RendererBeginPass(c, *r, regionMin, regionSize, sampleStart, sampleCount);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_run_pass(mrb_context ctx, mrb_renderer r, uint32_t chunk, mrb_render_stats* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || !out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::RendererRunPass(c, *r, chunk, *out);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_poll_stats(mrb_context ctx, mrb_renderer r, mrb_render_stats* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || !out) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::RendererPollStats(c, *r, *out);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_film_handoff(mrb_context ctx, mrb_renderer r, float* hostDst, mrb_host_fn onComplete, void* user)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || !hostDst) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::RendererFilmHandoff(c, *r, hostDst, onComplete, user);
        return MRB_OK;
    });
}

mrb_status mrb_renderer_reduce_peers(mrb_context ctx, mrb_renderer r, const mrb_context* peerCtx, const mrb_renderer* peers, uint32_t peerCount)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!r || (peerCount && (!peerCtx || !peers))) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        std::vector<mrb::Context*> pc(peerCount); std::vector<mrb_renderer_t*> pr(peerCount);
        for(uint32_t k = 0; k < peerCount; k++)
        {
            if(!peerCtx[k] || !peers[k]) return Fail(c, MRB_ERR_INVALID_ARG, "null peer");
            pc[k] = &peerCtx[k]->c; pr[k] = peers[k];
        }
        try { mrb::RendererReducePeers(c, *r, pc.data(), pr.data(), peerCount); }
        catch(...) { cudaSetDevice(c.device); throw; }
        return MRB_OK;
    });
}

mrb_status mrb_host_alloc(mrb_context ctx, size_t bytes, void** out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!out || bytes == 0) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        MRB_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
        return MRB_OK;
    });
}

void mrb_host_free(mrb_context ctx, void* ptr)
{
    if(!ctx || !ptr) return;
    cudaSetDevice(ctx->c.device);
    cudaFreeHost(ptr);
}

mrb_status mrb_filter_sample(mrb_context ctx, uint32_t filterType, float radius, const float* xi, uint32_t count, float* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(count && (!xi || !out)) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::FilterSampleHost(c, filterType, radius, xi, count, out);
        return MRB_OK;
    });
}

mrb_status mrb_texture_sample(mrb_context ctx, const mrb_texture_desc* texture, const float* uv, uint32_t count, float* rgbOut)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!texture || (count && (!uv || !rgbOut))) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::TextureSampleHost(c, *texture, uv, count, rgbOut);
        return MRB_OK;
    });
}


mrb_status mrb_dist2d_build(mrb_context ctx, const float* function, uint32_t width, uint32_t height, float* cdfX, float* cdfY, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!function || !cdfX || !cdfY) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(width == 0 || height == 0) return Fail(c, MRB_ERR_INVALID_ARG, "empty distribution");
        const size_t n = size_t(width) * height;
        if(n > 0xFFFFFFFFull / 4) return Fail(c, MRB_ERR_INVALID_ARG, "distribution too large");
        return Staged(c, memspace, {{const_cast<float*>(function), n * 4, true, false, nullptr}, {cdfX, n * 4, false, true, nullptr},
                                    {cdfY, size_t(height) * 4, false, true, nullptr}, {nullptr, 0, false, false, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      {
                          // row totals: scratch of the build (the staging block when staged, the trace scratch for device arrays)
                          c.traceScratch.Reserve(size_t(height) * 4);
                          mrb::Dist2DBuild(c, static_cast<const float*>(a[0].dev), width, height, static_cast<float*>(a[1].dev),
                                           static_cast<float*>(a[2].dev), static_cast<float*>(c.traceScratch.Base()));
                      });
    });
}

mrb_status mrb_dist2d_sample(mrb_context ctx, const float* cdfX, const float* cdfY, uint32_t width, uint32_t height,
                             const float* xi, uint32_t count, float* out, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!cdfX || !cdfY || (count && (!xi || !out))) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(width == 0 || height == 0) return Fail(c, MRB_ERR_INVALID_ARG, "empty distribution");
        if(count == 0) return MRB_OK;
        const size_t n = size_t(width) * height;
        return Staged(c, memspace, {{const_cast<float*>(cdfX), n * 4, true, false, nullptr}, {const_cast<float*>(cdfY), size_t(height) * 4, true, false, nullptr},
                                    {const_cast<float*>(xi), size_t(count) * 8, true, false, nullptr}, {out, size_t(count) * 16, false, true, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      {
                          const mrb::Dist2D d{static_cast<const float*>(a[0].dev), static_cast<const float*>(a[1].dev), width, height};
                          mrb::Dist2DSample(c, d, static_cast<const float*>(a[2].dev), count, static_cast<float*>(a[3].dev));
                      });
    });
}

mrb_status mrb_skysphere_convert(mrb_context ctx, uint32_t converter, const float* dirs, uint32_t count, float* out, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(converter != MRB_BOUNDARY_SKYSPHERE_SPHERICAL && converter != MRB_BOUNDARY_SKYSPHERE_COOCTA)
            return Fail(c, MRB_ERR_INVALID_ARG, "converter must be one of the skysphere types");
        if(count && (!dirs || !out)) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(count == 0) return MRB_OK;
        return Staged(c, memspace, {{const_cast<float*>(dirs), size_t(count) * 12, true, false, nullptr}, {out, size_t(count) * 32, false, true, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      { mrb::SkyConverters(c, converter, static_cast<const float*>(a[0].dev), count, static_cast<float*>(a[1].dev)); });
    });
}

mrb_status mrb_texture_luminance(mrb_context ctx, const mrb_texture_desc* texture, const float luminanceRow[3], float* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!texture || !luminanceRow || !out || !texture->data) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(texture->width == 0 || texture->height == 0 || (texture->channels != 3 && texture->channels != 4) || texture->format > 1u)
            return Fail(c, MRB_ERR_INVALID_ARG, "bad texture descriptor");
        const size_t n = size_t(texture->width) * texture->height;
        const size_t texBytes = n * texture->channels * (texture->format == 0u ? 4u : 1u);
        return Staged(c, MRB_MEM_HOST, {{const_cast<void*>(texture->data), texBytes, true, false, nullptr}, {out, n * 4, false, true, nullptr}},
                      [&](std::vector<StagedArray>& a)
                      { mrb::TextureLuminance(c, a[0].dev, texture->width, texture->height, texture->channels, texture->format, luminanceRow,
                                              static_cast<float*>(a[1].dev)); });
    });
}

mrb_status mrb_spectra_lut_generate(mrb_context ctx, const mrb_spectra_lut_desc* desc, float* lutOut, double whitepointOut[3])
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!desc || !lutOut || !desc->cieXYZ || !desc->illuminantSPD) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(!(desc->illuminantNormFactor > 0.0f)) return Fail(c, MRB_ERR_INVALID_ARG, "illuminantNormFactor must be positive");
        mrb::GenerateSpectraLUT(c, desc->cieXYZ, desc->illuminantSPD, desc->illuminantNormFactor, desc->rgbToXYZ, desc->xyzToRGB,
                                desc->resolution, desc->optimizePassCount ? desc->optimizePassCount : 15u, lutOut, whitepointOut);
        return MRB_OK;
    });
}

mrb_status mrb_texture_convert(mrb_context ctx, const mrb_texture_desc* texture, void* texelsOut)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!texture || !texelsOut) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::TextureConvertHost(c, *texture, texelsOut);
        return MRB_OK;
    });
}

size_t mrb_texture_chain_texels(uint32_t width, uint32_t height, uint32_t mipCount) { return mrb::TextureChainTexels(width, height, mipCount); }
uint32_t mrb_texture_full_mip_count(uint32_t width, uint32_t height) { return mrb::TextureFullMipCount(width, height); }
mrb_status mrb_texture_final_extent(const mrb_texture_desc* texture, uint32_t extentOut[3])
{
    if(!texture || !extentOut || texture->width == 0 || texture->height == 0) return MRB_ERR_INVALID_ARG;
    mrb::TextureFinalExtent(*texture, extentOut);
    return MRB_OK;
}
mrb_status mrb_texture_mip_chain(mrb_context ctx, const mrb_texture_desc* texture, void* chainOut, uint32_t* mipCountOut)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!texture || !chainOut) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        mrb::TextureMipChainHost(c, *texture, chainOut, mipCountOut);
        return MRB_OK;
    });
}
mrb_status mrb_texture_sample_lod(mrb_context ctx, const mrb_texture_desc* texture, const float* uv, const float* lod, const float* grads,
                                  uint32_t lodMode, uint32_t count, float* rgbOut)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!texture || (count && (!uv || !rgbOut || (!lod && !grads)))) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(lodMode > 1u) return Fail(c, MRB_ERR_INVALID_ARG, "unknown lodMode");
        mrb::TextureSampleLodHost(c, *texture, uv, lod, grads, lodMode, count, rgbOut);
        return MRB_OK;
    });
}

mrb_status mrb_multi_partition(mrb_context ctx, uint32_t* keys, uint32_t* indices, uint32_t count,
                               const uint32_t dataBitRange[2], const uint32_t batchBitRange[2],
                               int onlySortForBatches, uint32_t maxPartitions,
                               uint32_t* partitionCount, uint32_t* partitionOffsets, uint32_t* partitionKeys,
                               mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!dataBitRange || !batchBitRange || !partitionCount || !partitionOffsets || !partitionKeys)
            return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(count && (!keys || !indices)) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        if(batchBitRange[1] <= batchBitRange[0] || batchBitRange[1] > 32 || batchBitRange[1] - batchBitRange[0] > 16)
            return Fail(c, MRB_ERR_INVALID_ARG, "batch bit range must be 1..16 bits");
        if(dataBitRange[1] < dataBitRange[0] || dataBitRange[1] > 32) return Fail(c, MRB_ERR_INVALID_ARG, "bad data bit range");
        if(maxPartitions == 0) return Fail(c, MRB_ERR_INVALID_ARG, "maxPartitions == 0");
        const uint32_t bb = batchBitRange[1] - batchBitRange[0];
        size_t tempBytes = mrb::MultiPartitionTempBytes(count, bb);
        mrb::MultiAlloc sz(nullptr);
        sz.Take<char>(tempBytes); sz.Take<uint32_t>(count); sz.Take<uint32_t>(count);
        sz.Take<uint32_t>(4); sz.Take<uint32_t>(maxPartitions + 1); sz.Take<uint32_t>(maxPartitions);
        c.scratch.Reserve(sz.Total());
        mrb::MultiAlloc ma(c.scratch.Base());
        void* temp = ma.Take<char>(tempBytes);
        const bool batchOnly = onlySortForBatches != 0 || dataBitRange[1] == dataBitRange[0];
        if(memspace == MRB_MEM_DEVICE)
        {
            mrb::MultiPartition(c, keys, indices, count, dataBitRange, batchBitRange, batchOnly, maxPartitions,
                                partitionCount, partitionOffsets, partitionKeys, temp);
            return MRB_OK;
        }
        uint32_t* dk = ma.Take<uint32_t>(count); uint32_t* di = ma.Take<uint32_t>(count);
        uint32_t* dc = ma.Take<uint32_t>(4); uint32_t* dofs = ma.Take<uint32_t>(maxPartitions + 1); uint32_t* dpk = ma.Take<uint32_t>(maxPartitions);
        MRB_CUDA_TRY(cudaMemcpyAsync(dk, keys, 4 * size_t(count), cudaMemcpyHostToDevice, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(di, indices, 4 * size_t(count), cudaMemcpyHostToDevice, c.stream));
        mrb::MultiPartition(c, dk, di, count, dataBitRange, batchBitRange, batchOnly, maxPartitions, dc, dofs, dpk, temp);
        MRB_CUDA_TRY(cudaMemcpyAsync(keys, dk, 4 * size_t(count), cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(indices, di, 4 * size_t(count), cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(partitionCount, dc, 4, cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        uint32_t n = partitionCount[0];
        MRB_CUDA_TRY(cudaMemcpyAsync(partitionOffsets, dofs, 4 * size_t(n + 1), cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(partitionKeys, dpk, 4 * size_t(n), cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        return MRB_OK;
    });
}

mrb_status mrb_binary_partition(mrb_context ctx, uint32_t* indicesOut, uint32_t* leftCount,
                                const uint32_t* indicesIn, const uint8_t* flags, uint32_t flagCount,
                                uint32_t count, mrb_memspace memspace)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!leftCount || (count && (!indicesOut || !indicesIn || !flags))) return Fail(c, MRB_ERR_INVALID_ARG, "null argument");
        size_t tempBytes = mrb::BinaryPartitionTempBytes(count);
        mrb::MultiAlloc sz(nullptr);
        sz.Take<char>(tempBytes); sz.Take<uint32_t>(count); sz.Take<uint32_t>(count); sz.Take<uint8_t>(flagCount); sz.Take<uint32_t>(4);
        c.scratch.Reserve(sz.Total());
        mrb::MultiAlloc ma(c.scratch.Base());
        void* temp = ma.Take<char>(tempBytes);
        if(memspace == MRB_MEM_DEVICE) { mrb::BinaryPartition(c, indicesOut, leftCount, indicesIn, flags, count, temp); return MRB_OK; }
        uint32_t* dout = ma.Take<uint32_t>(count); uint32_t* din = ma.Take<uint32_t>(count);
        uint8_t* df = ma.Take<uint8_t>(flagCount); uint32_t* dc = ma.Take<uint32_t>(4);
        MRB_CUDA_TRY(cudaMemcpyAsync(din, indicesIn, 4 * size_t(count), cudaMemcpyHostToDevice, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(df, flags, flagCount, cudaMemcpyHostToDevice, c.stream));
        mrb::BinaryPartition(c, dout, dc, din, df, count, temp);
        MRB_CUDA_TRY(cudaMemcpyAsync(indicesOut, dout, 4 * size_t(count), cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaMemcpyAsync(leftCount, dc, 4, cudaMemcpyDeviceToHost, c.stream));
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        return MRB_OK;
    });
}


mrb_status mrb_scene_build(mrb_context ctx, const mrb_instance_desc* instances, uint32_t instanceCount, mrb_scene* out)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!instances || !out || instanceCount == 0) return Fail(c, MRB_ERR_INVALID_ARG, "empty scene");
        *out = nullptr;
        if(instanceCount >= (1u << 20)) return Fail(c, MRB_ERR_INVALID_ARG, "instance count exceeds AcceleratorKey index bits (20)");
        for(uint32_t i = 0; i < instanceCount; i++)
        {
            if(!instances[i].accel) return Fail(c, MRB_ERR_INVALID_ARG, "null accelerator in instance list");
            if(!instances[i].accel->d.wideNodes) return Fail(c, MRB_ERR_UNSUPPORTED, "instance of a BINARY_ONLY accelerator");
        }
        mrb_scene sc = new mrb_scene_t();
        try { mrb::BuildScene(c, *sc, instances, instanceCount); }
        catch(...) { c.persistentBytes -= sc->mem.Capacity(); delete sc; throw; }
        // two-level traversal: 2 entries per top-level level + the 3 entries pushed on entering an instance (pending node
        // group, pending instance group, sentinel) + 2 per level of the deepest instance
        uint32_t deepest = 0;
        for(uint32_t i = 0; i < instanceCount; i++) deepest = std::max(deepest, instances[i].accel->d.wideDepth);
        if(2u * sc->d.tlas.wideDepth + 3u + 2u * deepest > mrb::WIDE_STACK_ENTRIES)
        { c.persistentBytes -= sc->mem.Capacity(); delete sc; return Fail(c, MRB_ERR_UNSUPPORTED, "scene deeper than the traversal stack (2 * top + 3 + 2 * bottom > 64)"); }
        *out = sc;
        return MRB_OK;
    });
}

void mrb_scene_destroy(mrb_context ctx, mrb_scene scene)
{
    if(!ctx || !scene) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    ctx->c.persistentBytes -= scene->mem.Capacity();
    delete scene;
}

mrb_status mrb_scene_export_tlas(mrb_context ctx, mrb_scene scene, float* instanceAABBs, float* sceneAABB,
                                 uint64_t* morton, uint32_t* sortedInstance, uint32_t* nodes, float* nodeBoxes)
{
    return Guard(ctx, [&](mrb::Context& c)
    {
        if(!scene) return Fail(c, MRB_ERR_INVALID_ARG, "null scene");
        const mrb::AccelData& d = scene->d.tlas;
        auto D2H = [&](void* dst, const void* src, size_t bytes)
        { if(dst) MRB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream)); };
        D2H(instanceAABBs, d.leafAABB, sizeof(float) * 6 * d.leafCount);
        D2H(morton, d.morton, sizeof(uint64_t) * d.leafCount);
        D2H(sortedInstance, d.sortedLeaf, sizeof(uint32_t) * d.leafCount);
        D2H(nodes, d.nodes, sizeof(mrb::LBVHNode) * d.nodeCount);
        D2H(nodeBoxes, d.boxes, sizeof(mrb::LBVHBox) * d.nodeCount);
        MRB_CUDA_TRY(cudaStreamSynchronize(c.stream));
        if(sceneAABB) memcpy(sceneAABB, scene->aabb, sizeof(float) * 6);
        return MRB_OK;
    });
}

mrb_status mrb_scene_cast_rays(mrb_context ctx, mrb_scene scene, mrb_hit_key_pack* hitKeys, mrb_meta_hit* metaHits,
                               mrb_ray_gmem* rays, const uint32_t* rayIndices, uint32_t rayCount, uint32_t totalRayCount,
                               mrb_memspace memspace, mrb_trace_mode mode)
{
    if(!scene) return MRB_ERR_INVALID_ARG;
    const bool fresh = (uint32_t(mode) & uint32_t(MRB_TRACE_FRESH_OUTPUTS)) != 0u;
    mode = mrb_trace_mode(uint32_t(mode) & 0xFFu);
    return CastGeneric(ctx, false, hitKeys, metaHits, nullptr, rays, rayIndices, rayCount, totalRayCount, memspace, fresh,
        [&](mrb::Context& c, mrb_hit_key_pack* k, mrb_meta_hit* h, uint32_t* b, mrb_ray_gmem* r, const uint32_t* idx, uint32_t n)
        { mrb::TraceScene(c, scene->d, false, mode, k, h, b, r, idx, n); });
}

mrb_status mrb_scene_cast_visibility_rays(mrb_context ctx, mrb_scene scene, uint32_t* isVisibleBits,
                                          const mrb_ray_gmem* rays, const uint32_t* rayIndices,
                                          uint32_t rayCount, uint32_t totalRayCount, mrb_memspace memspace, mrb_trace_mode mode)
{
    if(!scene) return MRB_ERR_INVALID_ARG;
    const bool fresh = (uint32_t(mode) & uint32_t(MRB_TRACE_FRESH_OUTPUTS)) != 0u;
    mode = mrb_trace_mode(uint32_t(mode) & 0xFFu);
    return CastGeneric(ctx, true, nullptr, nullptr, isVisibleBits, const_cast<mrb_ray_gmem*>(rays), rayIndices, rayCount, totalRayCount, memspace, fresh,
        [&](mrb::Context& c, mrb_hit_key_pack* k, mrb_meta_hit* h, uint32_t* b, mrb_ray_gmem* r, const uint32_t* idx, uint32_t n)
        { mrb::TraceScene(c, scene->d, true, mode, k, h, b, r, idx, n); });
}

} // extern "C"
