// dist2d.cu — construction of the piecewise-constant 2-D distribution behind the skysphere light
// (DistributionGroupPwC2D::Construct, Tracer/Distributions.cu:L437-503) and the parity taps of dist2d.cuh.
//
// The reference issues three kernels (segmented scan of |f| in fp64, marginal scan, normalisation) and a fourth that
// patches pointers into per-row Distribution1D objects. Here:
//   KDistRows : one block per row — fp64 inclusive scan of |f| (warp shuffles, 1 024 texels per step), CDF stored as fp32,
//               row total kept aside, then the row is normalised in place with 1 / total in fp64 while it is still in L2.
//               Algorithmic traffic 8 B / texel (read f, write CDF); the normalisation pass re-reads from L2.
//   KDistMarginal : one block — the marginal is the fp32 running sum of the row totals IN ROW ORDER (the reference's CPU
//               kernel accumulates in float, so the order of the additions is part of the result), then normalised.
// The distribution object is two pointers and two sizes (Dist2D), so no pointer-patching kernel is needed.
#include "common.cuh"
#include "dist2d.cuh"

namespace mrb
{
namespace
{
constexpr uint32_t DTPB = 256, DITEMS = 4, DCHUNK = DTPB * DITEMS;

__global__ void __launch_bounds__(DTPB) KDistRows(const float* __restrict__ f, uint32_t w, uint32_t h, float* __restrict__ cdfX,
                                                  float* __restrict__ rowTotals)
{
    __shared__ double warpSum[DTPB / 32];
    __shared__ double carry;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for(uint32_t row = blockIdx.x; row < h; row += gridDim.x)
    {
        const float* __restrict__ in = f + size_t(row) * w;
        float* __restrict__ out = cdfX + size_t(row) * w;
        if(threadIdx.x == 0) carry = 0.0;
        __syncthreads();
        for(uint32_t base = 0; base < w; base += DCHUNK)
        {
            const uint32_t i0 = base + threadIdx.x * DITEMS;
            double v[DITEMS];
            #pragma unroll
            for(uint32_t k = 0; k < DITEMS; k++) v[k] = (i0 + k < w) ? double(fabsf(in[i0 + k])) : 0.0;
            #pragma unroll
            for(uint32_t k = 1; k < DITEMS; k++) v[k] += v[k - 1];
            // warp inclusive scan of the per-thread totals
            double tot = v[DITEMS - 1], incl = tot;
            #pragma unroll
            for(uint32_t o = 1; o < 32u; o <<= 1)
            {
                const double n = __shfl_up_sync(0xffffffffu, incl, o);
                if(lane >= o) incl += n;
            }
            if(lane == 31u) warpSum[warp] = incl;
            __syncthreads();
            double before = carry;
            for(uint32_t k = 0; k < warp; k++) before += warpSum[k];
            before += incl - tot;
            #pragma unroll
            for(uint32_t k = 0; k < DITEMS; k++) if(i0 + k < w) out[i0 + k] = float(before + v[k]);
            __syncthreads();
            if(threadIdx.x == DTPB - 1u) carry = before + v[DITEMS - 1];
            __syncthreads();
        }
        // KCNormalizeXY on this row: the unnormalised total is the row's last STORED value (written above by another thread
        // of this block; the barrier that closed the loop makes it visible)
        const float total = out[w - 1];
        __syncthreads();
        if(threadIdx.x == 0) rowTotals[row] = total;
        const double recip = 1.0 / double(total);
        for(uint32_t i = threadIdx.x; i < w; i += DTPB) out[i] = float(double(out[i]) * recip);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(DTPB) KDistMarginal(const float* __restrict__ rowTotals, uint32_t h, float* __restrict__ cdfY)
{
    __shared__ float buf[DCHUNK];
    __shared__ float running;
    if(threadIdx.x == 0) running = 0.0f;
    for(uint32_t base = 0; base < h; base += DCHUNK)
    {
        const uint32_t n = min(DCHUNK, h - base);
        for(uint32_t i = threadIdx.x; i < n; i += DTPB) buf[i] = rowTotals[base + i];
        __syncthreads();
        if(threadIdx.x == 0)
        {   // KCCopyScanY (CPU backend): sum += rowTotal; cdfY[i] = sum — float additions in row order
            float s = running;
            for(uint32_t i = 0; i < n; i++) { s = __fadd_rn(s, buf[i]); buf[i] = s; }
            running = s;
        }
        __syncthreads();
        for(uint32_t i = threadIdx.x; i < n; i += DTPB) cdfY[base + i] = buf[i];
        __syncthreads();
    }
    const double recip = 1.0 / double(running);
    for(uint32_t i = threadIdx.x; i < h; i += DTPB) cdfY[i] = float(double(cdfY[i]) * recip);
}

__global__ void __launch_bounds__(256) KDistSample(Dist2D d, const float2* __restrict__ xi, uint32_t n, float4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if(i >= n) return;
    const float3 s = DistSampleUV(d, xi[i].x, xi[i].y);
    out[i] = make_float4(s.x, s.y, s.z, DistPdfUV(d, s.x, s.y));
}

__global__ void __launch_bounds__(256) KSkyConverters(uint32_t mode, const float* __restrict__ dirs, uint32_t n, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if(i >= n) return;
    const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
    const float2 uv = SkyDirToUV(mode, dx, dy, dz);
    const float3 back = SkyUVToDir(mode, uv.x, uv.y);
    float* o = out + 8 * size_t(i);
    o[0] = uv.x; o[1] = uv.y; o[2] = SkyPdfFromDir(mode, 1.0f, dy);
    o[3] = back.x; o[4] = back.y; o[5] = back.z; o[6] = SkyPdfFromUV(mode, 1.0f, uv.y); o[7] = 0.0f;
}

__global__ void __launch_bounds__(256) KTextureLuminance(const void* __restrict__ texels, uint32_t n, uint32_t channels, uint32_t format,
                                                         float y0, float y1, float y2, float* __restrict__ out)
{
    for(uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u)
    {
        float c[3] = {0.f, 0.f, 0.f};
        for(uint32_t k = 0; k < 3u && k < channels; k++)
            c[k] = format == 0u ? static_cast<const float*>(texels)[size_t(i) * channels + k]
                                : __fmul_rn(float(static_cast<const uint8_t*>(texels)[size_t(i) * channels + k]), 1.0f / 255.0f);
        float r = __fmaf_rn(y0, c[0], 0.0f);
        r = __fmaf_rn(y1, c[1], r);
        out[i] = __fmaf_rn(y2, c[2], r);
    }
}
} // namespace

void Dist2DBuild(Context& ctx, const float* function, uint32_t w, uint32_t h, float* cdfX, float* cdfY, float* rowTotals)
{
    MRB_LAUNCH(ctx, KDistRows, GridFor(ctx, h * DTPB, DTPB), DTPB, 0, function, w, h, cdfX, rowTotals);
    MRB_LAUNCH(ctx, KDistMarginal, 1, DTPB, 0, rowTotals, h, cdfY);
}

void Dist2DSample(Context& ctx, const Dist2D& d, const float* xi, uint32_t n, float* out)
{
    if(n) MRB_LAUNCH(ctx, KDistSample, DivUp(n, 256u), 256, 0, d, reinterpret_cast<const float2*>(xi), n, reinterpret_cast<float4*>(out));
}

void SkyConverters(Context& ctx, uint32_t mode, const float* dirs, uint32_t n, float* out)
{
    if(n) MRB_LAUNCH(ctx, KSkyConverters, DivUp(n, 256u), 256, 0, mode, dirs, n, out);
}

void TextureLuminance(Context& ctx, const void* texels, uint32_t w, uint32_t h, uint32_t channels, uint32_t format, const float yRow[3], float* out)
{
    const uint32_t n = w * h;
    if(n) MRB_LAUNCH(ctx, KTextureLuminance, GridFor(ctx, n, 256u), 256, 0, texels, n, channels, format, yRow[0], yRow[1], yRow[2], out);
}
} // namespace mrb
