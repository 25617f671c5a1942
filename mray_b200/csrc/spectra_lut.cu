// spectra_lut.cu — the Jakob-Hanika 2019 coefficient LUT generator on the GPU: what the reference's SpectraLUTGen tool computes
// on host threads (Source/SpectraLUTGen/main.cpp, after mitsuba's rgb2spec) and ships as SpectraLUT/<COLORSPACE>.mrspectra — an
// INPUT FILE of the spectral hot path.
//
// For every cell of three res^3 tables (one per "largest channel" l) a 3-coefficient polynomial c is fitted so that the spectrum
// sigmoid(c0 t^2 + c1 t + c2) integrates (CIE 1931 observer x standard illuminant, Simpson 3/8) to the cell's RGB, by Gauss-Newton
// in fp64 on the CIELab residual: 15 passes at most, central-difference Jacobian (7 residual evaluations of 471 samples per pass),
// a 3x3 LU solve. A cell's initial guess is its neighbour's solution along the brightness axis k (from k = res / 5 outwards), so
// one THREAD owns one (l, j, i) column and walks its res cells in sequence: 3 res^2 threads (12 288 at res 64), each ~64 x 15 x 7
// x 471 dependent fp64 steps. The integration weights (471 x 3 doubles, 11 KB) sit in shared memory.
//
// Arithmetic follows the reference's host build operation by operation (explicit _rn intrinsics: its x86-64 build has no FMA
// contraction; Matrix * Vector is Math::Dot = true fma chains), so the result differs from the reference's file only through
// cbrt() (1 ulp) and the order of its white-point atomics: ~1e-7 relative on the stored fp32 coefficients.
#include "common.cuh"
#include <cmath>
#include <stdexcept>
#include <vector>

namespace mrb
{
namespace
{
constexpr uint32_t CIE_N = 471;   // Color::CIE_1931_N (Core/ColorFunctions.h:L66-68)

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 D3Sub(D3 a, D3 b) { return D3{__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
// Matrix<3> * Vector<3> = Math::Dot per row: fma(a0, b0, 0) -> fma(a1, b1, .) -> fma(a2, b2, .)
__device__ __forceinline__ D3 MatVec(const double* m, D3 v)
{
    D3 r;
    r.x = __fma_rn(m[2], v.z, __fma_rn(m[1], v.y, __fma_rn(m[0], v.x, 0.0)));
    r.y = __fma_rn(m[5], v.z, __fma_rn(m[4], v.y, __fma_rn(m[3], v.x, 0.0)));
    r.z = __fma_rn(m[8], v.z, __fma_rn(m[7], v.y, __fma_rn(m[6], v.x, 0.0)));
    return r;
}
// Color::XYZToCIELab (Core/ColorFunctions.h:L260-285)
__device__ __forceinline__ double LabF(double t)
{
    const double D = 6.0 / 29.0, DCube = D * D * D, Case2Factor = 1.0 / (D * D * 3.0), C = 4.0 / 29.0;
    return (t > DCube) ? cbrt(t) : __dadd_rn(__dmul_rn(t, Case2Factor), C);
}
__device__ __forceinline__ D3 XYZToLab(D3 xyz, D3 wp)
{
    const double xN = LabF(__ddiv_rn(xyz.x, wp.x)), yN = LabF(__ddiv_rn(xyz.y, wp.y)), zN = LabF(__ddiv_rn(xyz.z, wp.z));
    return D3{__dsub_rn(__dmul_rn(116.0, yN), 16.0), __dmul_rn(500.0, __dsub_rn(xN, yN)), __dmul_rn(200.0, __dsub_rn(yN, zN))};
}

struct LutParams
{
    double   rgbToXYZ[9];
    D3       whitepoint;
    uint32_t res, passes;
};

// EvaluateResidual (main.cpp:L158-186)
__device__ D3 Residual(const D3* __restrict__ sSpectraToRGB, const LutParams& p, D3 rgbLab, D3 c)
{
    D3 acc{0.0, 0.0, 0.0};
    const double NORM = 1.0 / double(CIE_N);
    for(uint32_t i = 0; i < CIE_N; i++)
    {
        const double lambdaN = __dmul_rn(__dmul_rn(double(i), 1.0), NORM);
        double x = c.x;
        x = __dadd_rn(__dmul_rn(x, lambdaN), c.y);
        x = __dadd_rn(__dmul_rn(x, lambdaN), c.z);
        double s = __dmul_rn(0.5, x);
        s = __ddiv_rn(s, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(x, x))));
        s = __dadd_rn(s, 0.5);
        const D3 w = sSpectraToRGB[i];
        acc.x = __dadd_rn(acc.x, __dmul_rn(w.x, s)); acc.y = __dadd_rn(acc.y, __dmul_rn(w.y, s)); acc.z = __dadd_rn(acc.z, __dmul_rn(w.z, s));
    }
    return D3Sub(rgbLab, XYZToLab(MatVec(p.rgbToXYZ, acc), p.whitepoint));
}

// LinearAlg::LUDecompose + SolveWithLU (Core/LinearAlg.h) for N = 3; returns false on a degenerate pivot
__device__ bool SolveLU3(double LU[3][3], D3 y, D3& out)
{
    int P[3] = {0, 1, 2};
    #pragma unroll
    for(int i = 0; i < 3; i++)
    {
        double maxVal = 0.0; int maxI = i;
        for(int k = i; k < 3; k++) { const double a = fabs(LU[k][i]); if(a > maxVal) { maxVal = a; maxI = k; } }
        if(maxVal < 1e-16) return false;
        if(maxI != i)
        {
            const int t = P[i]; P[i] = P[maxI]; P[maxI] = t;
            for(int x = 0; x < 3; x++) { const double v = LU[i][x]; LU[i][x] = LU[maxI][x]; LU[maxI][x] = v; }
        }
        // (the reference writes Float(1) / LU(i, i): a float one, promoted — the same value)
        const double diag = __ddiv_rn(1.0, LU[i][i]);
        for(int j = i + 1; j < 3; j++)
        {
            LU[j][i] = __dmul_rn(LU[j][i], diag);
            for(int k = i + 1; k < 3; k++) LU[j][k] = __dsub_rn(LU[j][k], __dmul_rn(LU[j][i], LU[i][k]));
        }
    }
    const double yy[3] = {y.x, y.y, y.z};
    double x[3];
    for(int i = 0; i < 3; i++)
    {
        x[i] = yy[P[i]];
        for(int k = 0; k < i; k++) x[i] = __dsub_rn(x[i], __dmul_rn(LU[i][k], x[k]));
    }
    for(int i = 2; i >= 0; i--)
    {
        for(int k = i + 1; k < 3; k++) x[i] = __dsub_rn(x[i], __dmul_rn(LU[i][k], x[k]));
        x[i] = __ddiv_rn(x[i], LU[i][i]);
    }
    out = D3{x[0], x[1], x[2]};
    return true;
}

// OptimizePolynomial (main.cpp:L139-262)
__device__ D3 Optimize(const D3* __restrict__ sW, const LutParams& p, D3 rgb, D3 guess)
{
    const D3 rgbLab = XYZToLab(MatVec(p.rgbToXYZ, rgb), p.whitepoint);
    D3 c = guess;
    const double EPS = 1e-4, FACTOR = 0.5 / EPS;
    for(uint32_t pass = 0; pass < p.passes; pass++)
    {
        const D3 residual = Residual(sW, p, rgbLab, c);
        double J[3][3];
        #pragma unroll
        for(int i = 0; i < 3; i++)
        {
            D3 a = c, b = c;
            if(i == 0) { a.x = __dsub_rn(a.x, EPS); b.x = __dadd_rn(b.x, EPS); }
            if(i == 1) { a.y = __dsub_rn(a.y, EPS); b.y = __dadd_rn(b.y, EPS); }
            if(i == 2) { a.z = __dsub_rn(a.z, EPS); b.z = __dadd_rn(b.z, EPS); }
            const D3 r0 = Residual(sW, p, rgbLab, a), r1 = Residual(sW, p, rgbLab, b);
            J[0][i] = __dmul_rn(__dsub_rn(r1.x, r0.x), FACTOR);
            J[1][i] = __dmul_rn(__dsub_rn(r1.y, r0.y), FACTOR);
            J[2][i] = __dmul_rn(__dsub_rn(r1.z, r0.z), FACTOR);
        }
        D3 step;
        if(SolveLU3(J, residual, step)) c = D3Sub(c, step);
        else c = D3Sub(c, D3{nan(""), nan(""), nan("")});   // the reference asserts; release builds run on with garbage
        const double mx = fmax(c.x, fmax(c.y, c.z));
        if(mx > 200.0) { const double k = __ddiv_rn(200.0, mx); c.x = __dmul_rn(c.x, k); c.y = __dmul_rn(c.y, k); c.z = __dmul_rn(c.z, k); }
        const double err = __dadd_rn(__dadd_rn(__dmul_rn(residual.x, residual.x), __dmul_rn(residual.y, residual.y)), __dmul_rn(residual.z, residual.z));
        if(err < 1.0e-7) break;    // MathConstants::SmallEpsilon<double>() (Core/MathConstants.h:L26)
    }
    return c;
}

// PassGenerateSpectrumLUT (main.cpp:L264-430): thread = one (l, j, i) column
__global__ void __launch_bounds__(128) KSpectraLUT(LutParams p, const D3* __restrict__ spectraToRGB, float* __restrict__ lut, uint32_t* __restrict__ badCells)
{
    __shared__ D3 sW[CIE_N];
    for(uint32_t i = threadIdx.x; i < CIE_N; i += blockDim.x) sW[i] = spectraToRGB[i];
    __syncthreads();
    const uint32_t res = p.res, xy = res * res, xyz = xy * res;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if(tid >= 3u * xy) return;
    const uint32_t l = tid / xy, j = (tid % xy) / res, i = tid % res;
    const uint32_t l1 = (l + 1u) % 3u, l2 = (l1 + 1u) % 3u;
    const uint32_t mid = res / 5u;
    auto Cell = [&](uint32_t k, D3 guess) -> D3
    {
        const double den = double(res - 1u);
        const double x = __ddiv_rn(double(i), den), y = __ddiv_rn(double(j), den), z = __ddiv_rn(double(k), den);
        auto SmoothStep = [](double t) { return __dmul_rn(__dmul_rn(t, t), __dsub_rn(3.0, __dmul_rn(2.0, t))); };
        const double b = SmoothStep(SmoothStep(z));
        double rgb[3];
        rgb[l] = b; rgb[l1] = __dmul_rn(x, b); rgb[l2] = __dmul_rn(y, b);
        const D3 c = Optimize(sW, p, D3{rgb[0], rgb[1], rgb[2]}, guess);
        if(!(isfinite(c.x) && isfinite(c.y) && isfinite(c.z))) atomicAdd(badCells, 1u);
        // from the normalised wavelength t in [0, 1] to nanometres: p0 = 360, p1 = 1 / (N - 1)
        const double p0 = 360.0, p1 = 1.0 / double(CIE_N - 1u);
        const double o0 = __dmul_rn(__dmul_rn(c.x, p1), p1);
        const double o1 = __dsub_rn(__dmul_rn(c.y, p1), __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.x), p0), p1), p1));
        const double o2 = __dadd_rn(__dsub_rn(c.z, __dmul_rn(__dmul_rn(c.y, p0), p1)), __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(c.x, p0), p1), p0), p1));
        const size_t off = size_t(l) * xyz * 3u + size_t(k) * xy + size_t(j) * res + i;
        lut[off] = float(o0); lut[off + xyz] = float(o1); lut[off + 2u * size_t(xyz)] = float(o2);
        return c;
    };
    D3 cur{0.0, 0.0, 0.0}, middle{0.0, 0.0, 0.0};
    for(uint32_t k = mid; k < res; k++)
    {
        cur = Cell(k, cur);
        if(k == mid) middle = cur;
    }
    cur = middle;
    for(int32_t k = int32_t(mid) - 1; k >= 0; k--) cur = Cell(uint32_t(k), cur);
}
} // namespace

// GenerateSpectraLUT (main.cpp:L464-537). Host inputs exactly as the reference holds them: the CIE 1931 observer (Vector3 float),
// the illuminant SPD (float) with its normalisation factor, and the colour space's RGB <-> XYZ matrices WITHOUT white-point
// adaptation (Color::GenRGBToXYZ(Primaries) and its inverse, float). lutOut: host, 9 * res^3 floats (table l, coefficient, z, y, x).
void GenerateSpectraLUT(Context& ctx, const float* cieXYZ, const float* illuminantSPD, float illuminantNorm, const float rgbToXYZ[9],
                        const float xyzToRGB[9], uint32_t res, uint32_t passes, float* lutOut, double whitepointOut[3])
{
    if(res < 5u || res > 256u) throw std::runtime_error("LUT resolution must be in [5, 256]");
    // PassGenSpectraToRGB (main.cpp:L81-137): Simpson 3/8 weights x illuminant x (XYZ -> RGB) x observer; white point = the sum
    std::vector<double> w(size_t(CIE_N) * 3);
    double wp[3] = {0.0, 0.0, 0.0};
    double M[9]; for(int k = 0; k < 9; k++) M[k] = double(xyzToRGB[k]);
    for(uint32_t i = 0; i < CIE_N; i++)
    {
        const double W = 3.0 / 8.0 * 1.0;
        const bool edge = (i == CIE_N - 1u || i == 0u);
        const double weight = edge ? W : (((i - 1u) % 3u == 2u) ? W * 2.0 : W * 3.0);
        const double I = double(illuminantSPD[i]) / double(illuminantNorm);
        const double x = double(cieXYZ[3 * i]), y = double(cieXYZ[3 * i + 1]), z = double(cieXYZ[3 * i + 2]);
        const double rgb[3] = {std::fma(M[2], z, std::fma(M[1], y, std::fma(M[0], x, 0.0))),
                               std::fma(M[5], z, std::fma(M[4], y, std::fma(M[3], x, 0.0))),
                               std::fma(M[8], z, std::fma(M[7], y, std::fma(M[6], x, 0.0)))};
        for(int c = 0; c < 3; c++) w[3 * size_t(i) + c] = rgb[c] * I * weight;
        wp[0] += x * I * weight; wp[1] += y * I * weight; wp[2] += z * I * weight;
    }
    LutParams p = {};
    for(int k = 0; k < 9; k++) p.rgbToXYZ[k] = double(rgbToXYZ[k]);
    p.whitepoint = D3{wp[0], wp[1], wp[2]};
    p.res = res; p.passes = passes;
    const size_t cells = size_t(9) * res * res * res;
    MultiAlloc sz(nullptr); sz.Take<D3>(CIE_N); sz.Take<float>(cells); sz.Take<uint32_t>(1);
    ctx.scratch.Reserve(sz.Total());
    MultiAlloc ma(ctx.scratch.Base());
    D3* dW = ma.Take<D3>(CIE_N); float* dLut = ma.Take<float>(cells); uint32_t* dBad = ma.Take<uint32_t>(1);
    MRB_CUDA_TRY(cudaMemcpyAsync(dW, w.data(), sizeof(double) * 3 * CIE_N, cudaMemcpyHostToDevice, ctx.stream));
    MRB_CUDA_TRY(cudaMemsetAsync(dBad, 0, 4, ctx.stream));
    const uint32_t threads = 3u * res * res;
    MRB_LAUNCH(ctx, KSpectraLUT, DivUp(threads, 128u), 128, 0, p, dW, dLut, dBad);
    uint32_t bad = 0;
    MRB_CUDA_TRY(cudaMemcpyAsync(lutOut, dLut, sizeof(float) * cells, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaMemcpyAsync(&bad, dBad, 4, cudaMemcpyDeviceToHost, ctx.stream));
    MRB_CUDA_TRY(cudaStreamSynchronize(ctx.stream));
    if(bad) throw std::runtime_error("Unable to optimize polynomial for " + std::to_string(bad) + " LUT cells");
    if(whitepointOut) { whitepointOut[0] = wp[0]; whitepointOut[1] = wp[1]; whitepointOut[2] = wp[2]; }
}

} // namespace mrb
