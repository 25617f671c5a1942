/* spectra_lut_oracle.c — CPU restatement of the reference's SpectraLUTGen (Source/SpectraLUTGen/main.cpp: PassGenSpectraToRGB
 * L81-137, OptimizePolynomial L139-262, PassGenerateSpectrumLUT L264-430; LinearAlg::LUDecompose / SolveWithLU, Core/LinearAlg.h;
 * Color::XYZToCIELab, Core/ColorFunctions.h:L260-285). TEST INFRASTRUCTURE ONLY.
 * Pinned by the reference's own output: tests/test_oracle_spectra_lut.py compares columns of the table against the
 * ACES_CG.mrspectra file the unmodified tool wrote in the authoring container. */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define CIE_N 471u
typedef struct { double x, y, z; } d3;

static d3 matvec(const double* m, d3 v)
{
    d3 r;
    r.x = fma(m[2], v.z, fma(m[1], v.y, fma(m[0], v.x, 0.0)));
    r.y = fma(m[5], v.z, fma(m[4], v.y, fma(m[3], v.x, 0.0)));
    r.z = fma(m[8], v.z, fma(m[7], v.y, fma(m[6], v.x, 0.0)));
    return r;
}
static double lab_f(double t)
{
    const double D = 6.0 / 29.0, DCube = D * D * D, Case2Factor = 1.0 / (D * D * 3.0), C = 4.0 / 29.0;
    return (t > DCube) ? cbrt(t) : t * Case2Factor + C;
}
static d3 xyz_to_lab(d3 xyz, d3 wp)
{
    double xN = lab_f(xyz.x / wp.x), yN = lab_f(xyz.y / wp.y), zN = lab_f(xyz.z / wp.z);
    d3 r = {116.0 * yN - 16.0, 500.0 * (xN - yN), 200.0 * (yN - zN)};
    return r;
}
typedef struct { double rgbToXYZ[9]; d3 wp; const d3* w; uint32_t passes; } lut_ctx;

static d3 residual(const lut_ctx* c, d3 rgbLab, d3 k)
{
    d3 acc = {0, 0, 0};
    const double NORM = 1.0 / (double)CIE_N;
    for(uint32_t i = 0; i < CIE_N; i++)
    {
        double lambdaN = (double)i * 1.0 * NORM;
        double x = k.x;
        x = x * lambdaN + k.y;
        x = x * lambdaN + k.z;
        double s = 0.5 * x;
        s /= sqrt(1.0 + x * x);
        s += 0.5;
        acc.x += c->w[i].x * s; acc.y += c->w[i].y * s; acc.z += c->w[i].z * s;
    }
    d3 lab = xyz_to_lab(matvec(c->rgbToXYZ, acc), c->wp);
    d3 r = {rgbLab.x - lab.x, rgbLab.y - lab.y, rgbLab.z - lab.z};
    return r;
}
static int solve_lu3(double LU[3][3], d3 y, d3* out)
{
    int P[3] = {0, 1, 2};
    for(int i = 0; i < 3; i++)
    {
        double maxVal = 0.0; int maxI = i;
        for(int k = i; k < 3; k++) { double a = fabs(LU[k][i]); if(a > maxVal) { maxVal = a; maxI = k; } }
        if(maxVal < 1e-16) return 0;
        if(maxI != i)
        {
            int t = P[i]; P[i] = P[maxI]; P[maxI] = t;
            for(int x = 0; x < 3; x++) { double v = LU[i][x]; LU[i][x] = LU[maxI][x]; LU[maxI][x] = v; }
        }
        double diag = 1.0 / LU[i][i];
        for(int j = i + 1; j < 3; j++)
        {
            LU[j][i] *= diag;
            for(int k = i + 1; k < 3; k++) LU[j][k] -= LU[j][i] * LU[i][k];
        }
    }
    double yy[3] = {y.x, y.y, y.z}, x[3];
    for(int i = 0; i < 3; i++) { x[i] = yy[P[i]]; for(int k = 0; k < i; k++) x[i] -= LU[i][k] * x[k]; }
    for(int i = 2; i >= 0; i--) { for(int k = i + 1; k < 3; k++) x[i] -= LU[i][k] * x[k]; x[i] /= LU[i][i]; }
    out->x = x[0]; out->y = x[1]; out->z = x[2];
    return 1;
}
static d3 optimize(const lut_ctx* c, d3 rgb, d3 guess)
{
    d3 rgbLab = xyz_to_lab(matvec(c->rgbToXYZ, rgb), c->wp);
    d3 k = guess;
    const double EPS = 1e-4, FACTOR = 0.5 / EPS;
    for(uint32_t pass = 0; pass < c->passes; pass++)
    {
        d3 r = residual(c, rgbLab, k);
        double J[3][3];
        for(int i = 0; i < 3; i++)
        {
            d3 a = k, b = k;
            if(i == 0) { a.x -= EPS; b.x += EPS; } else if(i == 1) { a.y -= EPS; b.y += EPS; } else { a.z -= EPS; b.z += EPS; }
            d3 r0 = residual(c, rgbLab, a), r1 = residual(c, rgbLab, b);
            J[0][i] = (r1.x - r0.x) * FACTOR; J[1][i] = (r1.y - r0.y) * FACTOR; J[2][i] = (r1.z - r0.z) * FACTOR;
        }
        d3 step;
        if(!solve_lu3(J, r, &step)) { k.x = k.y = k.z = NAN; return k; }
        k.x -= step.x; k.y -= step.y; k.z -= step.z;
        double mx = fmax(k.x, fmax(k.y, k.z));
        if(mx > 200.0) { double f = 200.0 / mx; k.x *= f; k.y *= f; k.z *= f; }
        double err = r.x * r.x + r.y * r.y + r.z * r.z;
        if(err < 1.0e-7) break;
    }
    return k;
}

/* One (l, j, i) column of the LUT: out[res * 3] = the three stored coefficients of cells k = 0 .. res - 1.
 * inputs: the 7 612-byte block mray_b200_spectra_lut_gen --dump-inputs writes (f32 cie[471*3], spd[471], norm, rgbToXYZ[9], xyzToRGB[9]). */
void orc_spectra_lut_column(const float* inputs, uint32_t res, uint32_t passes, uint32_t l, uint32_t j, uint32_t i, float* out)
{
    const float* cie = inputs; const float* spd = inputs + CIE_N * 3; float norm = inputs[CIE_N * 4];
    const float* r2x = inputs + CIE_N * 4 + 1; const float* x2r = r2x + 9;
    static d3 w[CIE_N];
    lut_ctx c; c.w = w; c.passes = passes;
    double M[9]; for(int k = 0; k < 9; k++) { M[k] = (double)x2r[k]; c.rgbToXYZ[k] = (double)r2x[k]; }
    d3 wp = {0, 0, 0};
    for(uint32_t n = 0; n < CIE_N; n++)
    {
        const double W = 3.0 / 8.0 * 1.0;
        int edge = (n == CIE_N - 1u || n == 0u);
        double weight = edge ? W : (((n - 1u) % 3u == 2u) ? W * 2.0 : W * 3.0);
        double I = (double)spd[n] / (double)norm;
        d3 xyz = {(double)cie[3 * n], (double)cie[3 * n + 1], (double)cie[3 * n + 2]};
        d3 rgb = matvec(M, xyz);
        w[n].x = rgb.x * I * weight; w[n].y = rgb.y * I * weight; w[n].z = rgb.z * I * weight;
        wp.x += xyz.x * I * weight; wp.y += xyz.y * I * weight; wp.z += xyz.z * I * weight;
    }
    c.wp = wp;
    const uint32_t l1 = (l + 1u) % 3u, l2 = (l1 + 1u) % 3u, mid = res / 5u;
    d3 cur = {0, 0, 0}, middle = {0, 0, 0};
    for(int dir = 0; dir < 2; dir++)
    {
        int32_t k = dir == 0 ? (int32_t)mid : (int32_t)mid - 1;
        if(dir == 1) cur = middle;
        for(; dir == 0 ? k < (int32_t)res : k >= 0; k += dir == 0 ? 1 : -1)
        {
            double den = (double)(res - 1u);
            double x = (double)i / den, y = (double)j / den, z = (double)k / den;
            double b = z * z * (3.0 - 2.0 * z); b = b * b * (3.0 - 2.0 * b);
            double rgb[3]; rgb[l] = b; rgb[l1] = x * b; rgb[l2] = y * b;
            d3 in = {rgb[0], rgb[1], rgb[2]};
            d3 co = optimize(&c, in, cur);
            const double p0 = 360.0, p1 = 1.0 / (double)(CIE_N - 1u);
            out[3 * k + 0] = (float)(co.x * p1 * p1);
            out[3 * k + 1] = (float)(co.y * p1 - 2.0 * co.x * p0 * p1 * p1);
            out[3 * k + 2] = (float)(co.z - co.y * p0 * p1 + co.x * p0 * p1 * p0 * p1);
            if(dir == 0 && k == (int32_t)mid) middle = co;
            cur = co;
        }
    }
}
