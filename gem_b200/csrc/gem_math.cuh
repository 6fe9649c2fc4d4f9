// gem_math.cuh -- arithmetic definition of the GEM hot path for sm_100a.
//
// The whole library is compiled with -fmad=false (no FMA contraction) and without fast
// math, so `a*b+c` is two IEEE roundings, `/` is __fdiv_rn / __ddiv_rn and sqrtf is
// __fsqrt_rn.  That makes the fp32 pipeline bit-identical to a plain C evaluation of the
// reference source (DESIGN.md "Arithmetic definition").
//
// The reference's feature kernel calls atan2f/sinf/cosf/acosf (gpu.cu:116-120, :650-652);
// CUDA's implementations of those are accurate to 1-2 ulp but their last bit is not
// specified.  We define them as the float rounding of a double evaluation built only from
// + - * / sqrt floor, which is reproducible on any IEEE machine.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gem {

// float -> int with the semantics of the instruction the reference compiles to
// (cvt.rzi.s32.f32: truncate toward zero, saturate, NaN -> 0).
__device__ __forceinline__ int f2i(float f) { return __float2int_rz(f); }
__device__ __forceinline__ int d2i(double d) { return __double2int_rz(d); }

namespace detail {
__device__ __forceinline__ double sin_k(double r)
{
    const double r2 = r * r;
    double p = -2.81145725434552076320e-15;
    p = p * r2 + 7.64716373181981647590e-13;
    p = p * r2 + -1.60590438368216145994e-10;
    p = p * r2 + 2.50521083854417187751e-08;
    p = p * r2 + -2.75573192239858906526e-06;
    p = p * r2 + 1.98412698412698412698e-04;
    p = p * r2 + -8.33333333333333333333e-03;
    p = p * r2 + 1.66666666666666666667e-01;
    return r - (r * r2) * p;
}
__device__ __forceinline__ double cos_k(double r)
{
    const double r2 = r * r;
    double p = -1.56192069685862264622e-16;
    p = p * r2 + 4.77947733238738529744e-14;
    p = p * r2 + -1.14707455977297247139e-11;
    p = p * r2 + 2.08767569878680989792e-09;
    p = p * r2 + -2.75573192239858906526e-07;
    p = p * r2 + 2.48015873015873015873e-05;
    p = p * r2 + -1.38888888888888888889e-03;
    p = p * r2 + 4.16666666666666666667e-02;
    p = p * r2 + -0.5;
    return 1.0 + r2 * p;
}
__device__ __forceinline__ void sincos_d(double x, double &s, double &c)
{
    if (!(x == x) || x - x != 0.0) {
        s = x - x;
        c = x - x;
        return;
    }
    const double kd = floor(x * 0.63661977236758134308 + 0.5);
    const double r = (x - kd * 1.57079632673412561417e+00) - kd * 6.07710050650619224932e-11;
    const long long k = (long long)kd;
    const double sr = sin_k(r), cr = cos_k(r);
    switch (k & 3) {
    case 0: s = sr; c = cr; break;
    case 1: s = cr; c = -sr; break;
    case 2: s = -sr; c = -cr; break;
    default: s = -cr; c = sr; break;
    }
}
__device__ __forceinline__ double atan_poly(double t)
{
    // odd Maclaurin series through t^43, Horner in t^2; the coefficients 1/n are
    // correctly rounded quotients (compile-time constant folding == runtime IEEE division)
    const double t2 = t * t;
    double p = 0.0;
#pragma unroll
    for (int n = 43; n >= 3; n -= 2) {
        double coef = 1.0 / (double)n;
        if (((n - 1) / 2) & 1) coef = -coef;
        p = (p + coef) * t2;
    }
    return t + t * p;
}
__device__ __forceinline__ double atan_core(double z)
{
    if (z > 0.41421356237309503) {
        const double t = (z - 1.0) / (z + 1.0);
        return 0.78539816339744830962 + atan_poly(t);
    }
    return atan_poly(z);
}
__device__ __forceinline__ double atan2_d(double y, double x)
{
    if (y != y || x != x) return y + x;
    const double ax = fabs(x), ay = fabs(y);
    double a;
    if (ax == 0.0 && ay == 0.0)
        a = 0.0;
    else if (ay <= ax)
        a = atan_core(ay / ax);
    else
        a = 1.57079632679489661923 - atan_core(ax / ay);
    if (x < 0.0) a = 3.14159265358979323846 - a;
    if (y < 0.0) a = -a;
    return a;
}
} // namespace detail

__device__ __forceinline__ void sincosf_det(float a, float &s, float &c)
{
    double sd, cd;
    detail::sincos_d((double)a, sd, cd);
    s = (float)sd;
    c = (float)cd;
}
__device__ __forceinline__ float atan2f_det(float y, float x)
{
    return (float)detail::atan2_d((double)y, (double)x);
}
__device__ __forceinline__ float acosf_det(float xf)
{
    const double x = (double)xf;
    if (!(x >= -1.0 && x <= 1.0)) return __int_as_float(0x7fc00000);
    return (float)detail::atan2_d(sqrt((1.0 - x) * (1.0 + x)), x);
}

} // namespace gem
