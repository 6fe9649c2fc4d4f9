/*
 * gem_oracle.c -- CPU ORACLE (test infrastructure, see gem_oracle.h header comment).
 *
 * PARITY STATUS: pinned against the reference's own kernels (gpu_process.cu compiled unmodified
 * from /root/reference by oracle/build_ref.py and run on a B200; tests/test_reference_pin.py,
 * tests/golden/gem_golden_v1.npz) -- see gem_oracle.h.  This file restates
 * /root/reference/elevation_mapping/elevation_mapping/cuda/gpu_process.cu ("gpu.cu").
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  x86-64 SSE2
 * gives true IEEE binary32 for `float` expressions (FLT_EVAL_METHOD == 0).
 *
 * Arithmetic conventions (DESIGN.md "Arithmetic definition"):
 *   - fp32 expressions are evaluated left to right exactly as the C++ source spells them,
 *     with NO fused multiply-add (the reference binary's FMA contraction is unknowable:
 *     its CMake passes no -fmad flag and the binary cannot be built here);
 *   - sub-expressions that C++ promotes to double (double literals such as 0.5, 0.0001,
 *     5, 0.6) are evaluated in double;
 *   - float->int casts follow the GPU instruction cvt.rzi.s32.f32 (truncate, saturate,
 *     NaN -> 0), because the reference code runs on the GPU;
 *   - Eigen reductions (norm(), 3-term inner products) are summed left to right.
 */
#include "gem_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* casts with GPU semantics                                                             */
/* ------------------------------------------------------------------------------------ */
static int f2i_rz(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}
static int d2i_rz(double d)
{
    if (d != d) return 0;
    if (d >= 2147483648.0) return INT_MAX;
    if (d <= -2147483649.0) return INT_MIN;
    return (int)d;
}

/* ------------------------------------------------------------------------------------ */
/* Deterministic float trig.                                                            */
/* The reference calls CUDA's atan2f/sinf/cosf/acosf (gpu.cu:116-120, :650-652).  Their */
/* last-bit behaviour is implementation specific, so the oracle DEFINES them as the     */
/* float rounding of a double-precision evaluation built only from + - * / sqrt, which  */
/* the product kernel mirrors bit for bit.  tests/ check these against libm (<= 1 ulp). */
/* ------------------------------------------------------------------------------------ */
#define ORC_PI 3.14159265358979323846
#define ORC_PIO2 1.57079632679489661923
#define ORC_PIO4 0.78539816339744830962
#define ORC_PIO2_HI 1.57079632673412561417e+00 /* first 33 bits of pi/2 */
#define ORC_PIO2_LO 6.07710050650619224932e-11 /* pi/2 - PIO2_HI         */
#define ORC_2OPI 0.63661977236758134308

static double sin_kernel(double r)
{ /* |r| <= pi/4, Taylor through r^17 */
    double r2 = r * r;
    double p = -2.81145725434552076320e-15;             /* -1/17! */
    p = p * r2 + 7.64716373181981647590e-13;            /*  1/15! */
    p = p * r2 + -1.60590438368216145994e-10;           /* -1/13! */
    p = p * r2 + 2.50521083854417187751e-08;            /*  1/11! */
    p = p * r2 + -2.75573192239858906526e-06;           /* -1/9!  */
    p = p * r2 + 1.98412698412698412698e-04;            /*  1/7!  */
    p = p * r2 + -8.33333333333333333333e-03;           /* -1/5!  */
    p = p * r2 + 1.66666666666666666667e-01;            /*  1/3!  */
    /* sin r = r - r^3 * (1/6 - r^2/120 + ...) */
    return r - (r * r2) * p;
}
static double cos_kernel(double r)
{ /* |r| <= pi/4, Taylor through r^18 */
    double r2 = r * r;
    double p = -1.56192069685862264622e-16;             /* -1/18! */
    p = p * r2 + 4.77947733238738529744e-14;            /*  1/16! */
    p = p * r2 + -1.14707455977297247139e-11;           /* -1/14! */
    p = p * r2 + 2.08767569878680989792e-09;            /*  1/12! */
    p = p * r2 + -2.75573192239858906526e-07;           /* -1/10! */
    p = p * r2 + 2.48015873015873015873e-05;            /*  1/8!  */
    p = p * r2 + -1.38888888888888888889e-03;           /* -1/6!  */
    p = p * r2 + 4.16666666666666666667e-02;            /*  1/4!  */
    p = p * r2 + -0.5;
    return 1.0 + r2 * p;
}
static void sincos_d(double x, double *s, double *c)
{
    /* |x| is at most a few pi on this path; NaN/Inf propagate as NaN */
    if (!(x == x) || x - x != 0.0) {
        *s = x - x;
        *c = x - x;
        return;
    }
    double kd = floor(x * ORC_2OPI + 0.5);
    double r = (x - kd * ORC_PIO2_HI) - kd * ORC_PIO2_LO;
    long k = (long)kd;
    double sr = sin_kernel(r), cr = cos_kernel(r);
    switch (k & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}
static double atan_poly(double t)
{ /* |t| <= tan(pi/8): odd Maclaurin series through t^43 */
    double t2 = t * t;
    double p = 0.0;
    int n;
    for (n = 43; n >= 3; n -= 2) {
        double coef = 1.0 / (double)n;
        if (((n - 1) / 2) & 1) coef = -coef;
        p = (p + coef) * t2;
    }
    return t + t * p;
}
static double atan_core(double z)
{ /* 0 <= z <= 1 */
    if (z > 0.41421356237309503) {
        double t = (z - 1.0) / (z + 1.0);
        return ORC_PIO4 + atan_poly(t);
    }
    return atan_poly(z);
}
static double atan2_d(double y, double x)
{
    if (y != y || x != x) return y + x;
    double ax = fabs(x), ay = fabs(y), a;
    if (ax == 0.0 && ay == 0.0)
        a = 0.0;
    else if (ay <= ax)
        a = atan_core(ay / ax);
    else
        a = ORC_PIO2 - atan_core(ax / ay);
    if (x < 0.0) a = ORC_PI - a;
    if (y < 0.0) a = -a;
    return a;
}
float orc_sinf(float a)
{
    double s, c;
    sincos_d((double)a, &s, &c);
    return (float)s;
}
float orc_cosf(float a)
{
    double s, c;
    sincos_d((double)a, &s, &c);
    return (float)c;
}
float orc_atan2f(float y, float x) { return (float)atan2_d((double)y, (double)x); }
float orc_acosf(float xf)
{
    double x = (double)xf;
    if (!(x >= -1.0 && x <= 1.0)) return NAN; /* like acosf */
    return (float)atan2_d(sqrt((1.0 - x) * (1.0 + x)), x);
}

/* ------------------------------------------------------------------------------------ */
/* create / destroy: Init_GPU_elevationmap gpu.cu:940-994, G_Init_map gpu.cu:198-214    */
/* ------------------------------------------------------------------------------------ */
orc_map *orc_create(int length, float resolution, float mahalanobis, float obstacle_threshold)
{
    orc_map *m = (orc_map *)calloc(1, sizeof(orc_map));
    size_t C = (size_t)length * (size_t)length, i;
    m->L = length;
    m->res = resolution;
    m->mahalanobis = mahalanobis;
    m->obstacle_threshold = obstacle_threshold;
    m->lowest = (float *)malloc(C * sizeof(float));
    m->elevation = (float *)malloc(C * sizeof(float));
    m->variance = (float *)malloc(C * sizeof(float));
    m->intensity = (float *)malloc(C * sizeof(float));
    m->traver = (float *)malloc(C * sizeof(float));
    m->colorR = (int *)malloc(C * sizeof(int));
    m->colorG = (int *)malloc(C * sizeof(int));
    m->colorB = (int *)malloc(C * sizeof(int));
    for (i = 0; i < C; i++) { /* gpu.cu:203-210 */
        m->intensity[i] = 0;
        m->elevation[i] = -10;
        m->variance[i] = -10;
        m->lowest[i] = 100;
        m->traver[i] = -10;
        m->colorR[i] = 0;
        m->colorG[i] = 0;
        m->colorB[i] = 0;
    }
    m->centre[0] = m->centre[1] = 0; /* gpu.cu:942,972 */
    m->start[0] = m->start[1] = 0;   /* gpu.cu:943,973 */
    m->sensorZ = 0;
    m->compat_box_filter = 1;
    return m;
}
void orc_destroy(orc_map *m)
{
    if (!m) return;
    free(m->lowest); free(m->elevation); free(m->variance); free(m->intensity);
    free(m->traver); free(m->colorR); free(m->colorG); free(m->colorB);
    free(m);
}

/* ------------------------------------------------------------------------------------ */
/* Move: gpu.cu:1004-1083                                                               */
/* ------------------------------------------------------------------------------------ */
static int index_to_range(int index, int L)
{ /* gpu.cu:916-921 */
    if (index < 0) index += ((-index / L) + 1) * L;
    index = index % L;
    return index;
}
static float position_to_range(float p, float shift, float resolution)
{ /* gpu.cu:996-1002: int p_index = round(p / resolution) (float division, roundf) */
    int p_index = f2i_rz(roundf(p / resolution));
    int shift_index = f2i_rz(roundf(shift / resolution));
    int current_index = p_index + shift_index;
    return (float)current_index * resolution;
}
static void clear_cell(orc_map *m, size_t c)
{ /* G_Clear_map gpu.cu:260-273: traver and lowest are NOT reset */
    m->intensity[c] = 0;
    m->elevation[c] = -10;
    m->variance[c] = -10;
    m->colorR[c] = 0;
    m->colorG[c] = 0;
    m->colorB[c] = 0;
}
static void clear_rows(orc_map *m, int start, int n)
{ /* Clear_regionrow gpu.cu:923-929 -> G_Clear_map(start, n, true) :258-266 */
    int L = m->L, i;
    for (i = 0; i < L * n; i++) clear_cell(m, (size_t)start * L + i);
}
static void clear_cols(orc_map *m, int start, int n)
{ /* Clear_regioncol gpu.cu:931-938 -> G_Clear_map(start, n, false) :267-274 */
    int L = m->L, i;
    for (i = 0; i < L * n; i++) clear_cell(m, (size_t)(i / n) * L + i % n + start);
}
static void clear_all(orc_map *m)
{ /* G_Clear_allmap gpu.cu:216-230: also resets traver, not lowest */
    size_t C = (size_t)m->L * m->L, c;
    for (c = 0; c < C; c++) {
        clear_cell(m, c);
        m->traver[c] = -10;
    }
}
void orc_move(orc_map *m, const float pos[3], float centre_out[2], int start_out[2],
              float shift_out[2])
{
    int L = m->L, i;
    int indexShift[2];
    float positionShift[2], aligned[2];
    m->sensorZ = pos[2]; /* gpu.cu:1011-1012 */
    positionShift[0] = pos[0] - m->centre[0];
    positionShift[1] = pos[1] - m->centre[1];
    for (i = 0; i < 2; i++) { /* gpu.cu:893-902: float/float + double -> int truncation */
        double v = (double)(positionShift[i] / m->res) + 0.5 * (positionShift[i] > 0 ? 1 : -1);
        indexShift[i] = d2i_rz(v);
        aligned[i] = (float)indexShift[i] * m->res; /* gpu.cu:904-914 */
    }
    for (i = 0; i < 2; i++) {
        if (indexShift[i] != 0) {
            /* gpu.cu:1033 tests only indexShift >= length; a shift <= -length would run the
             * partial-clear branch with nCells > length and write out of bounds (reference
             * bug).  ORACLE DEFINITION: |shift| >= length clears all. */
            if (indexShift[i] >= L || indexShift[i] <= -L) {
                clear_all(m);
            } else {
                int sign = (indexShift[i] > 0 ? 1 : -1);
                int startIndex = m->start[i] - (sign > 0 ? 1 : 0);
                int endIndex = startIndex + sign - indexShift[i];
                int nCells = abs(indexShift[i]);
                int index = (sign < 0 ? startIndex : endIndex);
                index = index_to_range(index, L);
                if (index + nCells <= L) {
                    if (i == 0) clear_rows(m, index, nCells);
                    else clear_cols(m, index, nCells);
                } else {
                    int firstnCells = L - index;
                    int secondnCells = nCells - firstnCells;
                    if (i == 0) {
                        clear_rows(m, index, firstnCells);
                        clear_rows(m, 0, secondnCells);
                    } else {
                        clear_cols(m, index, firstnCells);
                        clear_cols(m, 0, secondnCells);
                    }
                }
            }
        }
        m->start[i] -= indexShift[i];
        m->start[i] = index_to_range(m->start[i], L);
        m->centre[i] = position_to_range(m->centre[i], aligned[i], m->res);
    }
    if (centre_out) { centre_out[0] = m->centre[0]; centre_out[1] = m->centre[1]; }
    if (start_out) { start_out[0] = m->start[0]; start_out[1] = m->start[1]; }
    if (shift_out) { shift_out[0] = aligned[0]; shift_out[1] = aligned[1]; }
}

/* ------------------------------------------------------------------------------------ */
/* index functions: gpu.cu:309-358                                                      */
/* ------------------------------------------------------------------------------------ */
int orc_points_to_index(const orc_map *m, float px, float py, int *storage)
{
    int L = m->L, ix, iy;
    float sx = px - m->centre[0];
    float sy = py - m->centre[1];
    if (L % 2 == 0) { /* gpu.cu:316-317 */
        ix = f2i_rz((float)(L / 2) - sx / m->res);
        iy = f2i_rz((float)(L / 2) - sy / m->res);
    } else { /* gpu.cu:321-322: float quotient + double 0.5*sign, cast from double */
        ix = L / 2 - d2i_rz((double)(sx / m->res) + 0.5 * (sx > 0 ? 1 : -1));
        iy = L / 2 - d2i_rz((double)(sy / m->res) + 0.5 * (sy > 0 ? 1 : -1));
    }
    if (ix >= 0 && ix < L && iy >= 0 && iy < L) {
        if (storage) { /* gpu.cu:350-353 */
            int stx = (ix + m->start[0]) % L;
            int sty = (iy + m->start[1]) % L;
            *storage = stx * L + sty;
        }
        return ix * L + iy;
    }
    if (storage) *storage = -1;
    return -1;
}

/* ------------------------------------------------------------------------------------ */
/* per-point transform + filter + variance: G_pointsprocess gpu.cu:384-455              */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int accepted;
    float h, hv, xt, yt;
    int key, geo;
} pt_result;

static float sensor_variances(const orc_sensor *s, float x, float y, float z, float *vN)
{ /* returns varianceLateral, writes varianceNormal */
    if (s->type == ORC_SENSOR_STRUCTURED_LIGHT) {
        /* StructuredLightSensorProcessor.cpp:129-139; the parameters are doubles
         * (sensorParameters_ is a map<string,double>), pow() is double */
        float d = z; /* :130 measurementDistance = pointVector.z() */
        /* pow(x, 1.0) == x exactly in IEEE libms; spelled out so that the GPU path need not
         * match a libm pow bit for bit (realsense_d435.yaml: normal_factor_e = 1) */
        double pw = (s->nf_e == 1.0) ? (double)d : pow((double)d, s->nf_e);
        float devN = (float)(s->nf_a + s->nf_b * ((double)d - s->nf_c) * ((double)d - s->nf_c) + s->nf_d * pw);
        float devL = (float)(s->lateral * (double)d);
        *vN = devN * devN;
        return devL * devL;
    } else {
        /* gpu.cu:407 Eigen norm(), :410-411 pow(.,2) == exact square */
        float d = sqrtf((x * x + y * y) + z * z);
        float b = s->beam_c + s->beam_a * d;
        *vN = s->min_r * s->min_r;
        return b * b;
    }
}

int orc_clean_point_cloud(const orc_sensor *s, int n, float *xyzi, unsigned char *rgba)
{
    int i, k = 0;
    float lo = (float)s->cutoff_min, hi = (float)s->cutoff_max; /* setFilterLimits(const float&, const float&) */
    for (i = 0; i < n; i++) {
        const float x = xyzi[4 * i], y = xyzi[4 * i + 1], z = xyzi[4 * i + 2];
        if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue; /* both filters */
        if (s->type == ORC_SENSOR_STRUCTURED_LIGHT && (z < lo || z > hi)) continue; /* SL.cpp:58 */
        if (k != i) {
            memcpy(xyzi + 4 * k, xyzi + 4 * i, 4 * sizeof(float));
            if (rgba) memcpy(rgba + 4 * k, rgba + 4 * i, 4);
        }
        k++;
    }
    return k;
}

static void process_one(const orc_map *m, float x, float y, float z, const float T[16],
                        double relLower, double relUpper, const orc_sensor *sensor,
                        const float sJ[3], const float rotVar[9], const float C_SB_T[9],
                        const float P[3], const float B_skew[9], pt_result *r)
{
    /* gpu.cu:389 */
    float h = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
    int flag = 0;
    if (m->compat_box_filter) { /* gpu.cu:393 */
        if ((x > -1.5 && x < 1.5 && y > -1.5 && y < 1.5) || (y > -1 && y < 1) || y > 0) flag = 1;
    }
    r->accepted = 0;
    r->h = -1; r->hv = -1; r->xt = -1; r->yt = -1; r->key = -1; r->geo = -1; /* :443-450 */
    if (((double)h > relLower && (double)h < relUpper) && flag == 0) { /* gpu.cu:397 */
        float vN, vL;
        float q[3], S[9], rotJ[3], A1[3], B1[3], SV[9];
        float term1, term2, hv;
        int j;
        r->accepted = 1;
        r->xt = ((T[0] * x + T[1] * y) + T[2] * z) + T[3]; /* :399 */
        r->yt = ((T[4] * x + T[5] * y) + T[6] * z) + T[7]; /* :400 */
        r->h = h;
        vL = sensor_variances(sensor, x, y, z, &vN);
        /* rotation Jacobian gpu.cu:417-418 */
        for (j = 0; j < 3; j++) q[j] = (C_SB_T[3 * j] * x + C_SB_T[3 * j + 1] * y) + C_SB_T[3 * j + 2] * z;
        S[0] = 0 + B_skew[0];     S[1] = -q[2] + B_skew[1]; S[2] = q[1] + B_skew[2];
        S[3] = q[2] + B_skew[3];  S[4] = 0 + B_skew[4];     S[5] = -q[0] + B_skew[5];
        S[6] = -q[1] + B_skew[6]; S[7] = q[0] + B_skew[7];  S[8] = 0 + B_skew[8];
        for (j = 0; j < 3; j++) rotJ[j] = (P[0] * S[j] + P[1] * S[3 + j]) + P[2] * S[6 + j];
        /* cuda_computer gpu.cu:295-300 */
        for (j = 0; j < 3; j++) A1[j] = (rotJ[0] * rotVar[j] + rotJ[1] * rotVar[3 + j]) + rotJ[2] * rotVar[6 + j];
        term1 = (A1[0] * rotJ[0] + A1[1] * rotJ[1]) + A1[2] * rotJ[2];
        memset(SV, 0, sizeof SV);
        SV[0] = vL; SV[4] = vL; SV[8] = vN; /* :413-414 */
        for (j = 0; j < 3; j++) B1[j] = (sJ[0] * SV[j] + sJ[1] * SV[3 + j]) + sJ[2] * SV[6 + j];
        term2 = (B1[0] * sJ[0] + B1[1] * sJ[1]) + B1[2] * sJ[2];
        hv = term1;   /* :422 */
        hv += term2;  /* :425 */
        r->hv = hv;
        r->geo = orc_points_to_index(m, r->xt, r->yt, &r->key); /* :430-431 */
    }
}

void orc_process_points(orc_map *m, int n, const float *x, const float *y, const float *z,
                        const float T[16], double relLower, double relUpper,
                        const orc_sensor *sensor, const float sJ[3], const float rotVar[9],
                        const float C_SB_T[9], const float P[3], const float B_skew[9],
                        int *key, float *var, float *x_ts, float *y_ts, float *z_ts)
{
    size_t C = (size_t)m->L * m->L;
    float *minh = (float *)malloc(C * sizeof(float));
    int *argmin = (int *)malloc(C * sizeof(int));
    float *hvs = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    int *touched = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    int nt = 0, i;
    size_t c;
    for (c = 0; c < C; c++) argmin[c] = -1;
    for (i = 0; i < n; i++) {
        pt_result r;
        process_one(m, x[i], y[i], z[i], T, relLower, relUpper, sensor, sJ, rotVar, C_SB_T, P,
                    B_skew, &r);
        if (key) key[i] = r.key;
        if (var) var[i] = r.hv;
        if (x_ts) x_ts[i] = r.xt;
        if (y_ts) y_ts[i] = r.yt;
        if (z_ts) z_ts[i] = r.h;
        hvs[i] = r.hv;
        if (r.accepted && r.geo != -1) {
            /* first index attaining the minimum wins */
            if (argmin[r.geo] < 0) {
                argmin[r.geo] = i;
                minh[r.geo] = r.h;
                touched[nt++] = r.geo;
            } else if (r.h < minh[r.geo]) {
                argmin[r.geo] = i;
                minh[r.geo] = r.h;
            }
        }
    }
    /* gpu.cu:432-438.  The reference does atomicMin(lowest, h) and then a NON-atomic
     * "if (h == lowest) lowest += 3*hv" -- a data race.  ORACLE DEFINITION:
     * per call and per geographic cell, with m = min h and i* the first index attaining it,
     *      if (m <= lowest_old) lowest = m + 3*hv[i*];  else unchanged.                   */
    for (i = 0; i < nt; i++) {
        int g = touched[i];
        if (minh[g] <= m->lowest[g]) m->lowest[g] = minh[g] + 3 * hvs[argmin[g]];
    }
    free(minh); free(argmin); free(hvs); free(touched);
}

/* ------------------------------------------------------------------------------------ */
/* G_fuse gpu.cu:477-537                                                                */
/* ------------------------------------------------------------------------------------ */
static void fuse_one(orc_map *m, int c, int R, int G, int B, float I, float h, float v)
{
    int colour_ok = (R != 0 && G != 0 && B != 0 && I != 0); /* :488,508,520 */
    if (m->elevation[c] == -10) { /* :484-495 */
        m->elevation[c] = h;
        m->variance[c] = v;
        if (colour_ok) { m->intensity[c] = I; m->colorR[c] = R; m->colorG[c] = G; m->colorB[c] = B; }
    } else {
        float mah;
        if ((double)m->variance[c] < 0.0001) m->variance[c] = (float)0.0001; /* :500-501 */
        mah = fabsf(h - m->elevation[c]) / sqrtf(m->variance[c]);            /* :502 */
        if ((double)mah > 5) {                                                /* :504 */
            if (m->elevation[c] < h) {                                        /* :505-515 */
                m->elevation[c] = h;
                m->variance[c] = v;
                if (colour_ok) { m->intensity[c] = I; m->colorR[c] = R; m->colorG[c] = G; m->colorB[c] = B; }
            }
        } else { /* :518-526; elevation first (old variance), then variance */
            float ov = m->variance[c], oe = m->elevation[c];
            m->elevation[c] = (ov * h + v * oe) / (ov + v);
            m->variance[c] = (v * ov) / (v + ov);
            if (colour_ok) { m->intensity[c] = I; m->colorR[c] = R; m->colorG[c] = G; m->colorB[c] = B; }
        }
    }
}
static void fuse_floor(orc_map *m)
{ /* gpu.cu:533-534, applied to EVERY cell (empty cells go from -10 to 1e-4) */
    size_t C = (size_t)m->L * m->L, c;
    for (c = 0; c < C; c++)
        if ((double)m->variance[c] < 0.0001) m->variance[c] = (float)0.0001;
}
void orc_fuse(orc_map *m, int n, const int *key, const int *R, const int *G, const int *B,
              const float *intensity, const float *h, const float *var)
{
    long C = (long)m->L * m->L;
    int i;
    for (i = 0; i < n; i++) {
        if (key[i] < 0 || key[i] >= C) continue; /* no thread has that map_index */
        if (h[i] == -1) continue;                 /* :482 */
        fuse_one(m, key[i], R ? R[i] : 0, G ? G[i] : 0, B ? B[i] : 0, intensity ? intensity[i] : 0,
                 h[i], var[i]);
    }
    fuse_floor(m);
}
void orc_fuse_literal(orc_map *m, int n, const int *key, const int *R, const int *G,
                      const int *B, const float *intensity, const float *h, const float *var)
{
    int C = m->L * m->L, c, i;
    for (c = 0; c < C; c++) {
        for (i = 0; i < n; i++) {
            if (key[i] != c || h[i] == -1) continue;
            fuse_one(m, c, R ? R[i] : 0, G ? G[i] : 0, B ? B[i] : 0, intensity ? intensity[i] : 0,
                     h[i], var[i]);
        }
        if ((double)m->variance[c] < 0.0001) m->variance[c] = (float)0.0001;
    }
}

/* G_Mapvar_update gpu.cu:540-547 */
void orc_var_update(orc_map *m, float dv)
{
    size_t C = (size_t)m->L * m->L, c;
    for (c = 0; c < C; c++)
        if (m->variance[c] != -10) m->variance[c] += dv;
}

/* ------------------------------------------------------------------------------------ */
/* computerEigenvalue gpu.cu:66-187 (Jacobi, nDim == 3)                                 */
/* ------------------------------------------------------------------------------------ */
static void jacobi_min_eigvec(float *pM, float *vec_out, float dbEps, int nJt)
{
    const int nDim = 3;
    float V[9];
    int i, j, nCount = 0;
    for (i = 0; i < nDim; i++) {
        V[i * nDim + i] = 1.0f;
        for (j = 0; j < nDim; j++)
            if (i != j) V[i * nDim + j] = 0.0f;
    }
    for (;;) {
        float dbMax = pM[1]; /* :85 signed initial value */
        int nRow = 0, nCol = 1;
        for (i = 0; i < nDim; i++)
            for (j = 0; j < nDim; j++) {
                float d = fabsf(pM[i * nDim + j]);
                if ((i != j) && (d > dbMax)) { dbMax = d; nRow = i; nCol = j; }
            }
        if (dbMax < dbEps) break; /* :103 */
        if (nCount > nJt) break;  /* :106 */
        nCount++;
        {
            float dbApp = pM[nRow * nDim + nRow];
            float dbApq = pM[nRow * nDim + nCol];
            float dbAqq = pM[nCol * nDim + nCol];
            /* :116  0.5*atan2f(...) : double product of an exact halving, stored to float */
            float dbAngle = (float)(0.5 * (double)orc_atan2f(-2 * dbApq, dbAqq - dbApp));
            float s = orc_sinf(dbAngle);
            float c = orc_cosf(dbAngle);
            float s2 = orc_sinf(2 * dbAngle);
            float c2 = orc_cosf(2 * dbAngle);
            /* :122-127 */
            pM[nRow * nDim + nRow] = (dbApp * c * c + dbAqq * s * s) + 2 * dbApq * c * s;
            pM[nCol * nDim + nCol] = (dbApp * s * s + dbAqq * c * c) - 2 * dbApq * c * s;
            pM[nRow * nDim + nCol] =
                (float)(0.5 * (double)(dbAqq - dbApp) * (double)s2 + (double)(dbApq * c2));
            pM[nCol * nDim + nRow] = pM[nRow * nDim + nCol];
            for (i = 0; i < nDim; i++) { /* :129-139 */
                if ((i != nCol) && (i != nRow)) {
                    int u = i * nDim + nRow, w = i * nDim + nCol;
                    float t = pM[u];
                    pM[u] = pM[w] * s + t * c;
                    pM[w] = pM[w] * c - t * s;
                }
            }
            for (j = 0; j < nDim; j++) { /* :141-151 */
                if ((j != nCol) && (j != nRow)) {
                    int u = nRow * nDim + j, w = nCol * nDim + j;
                    float t = pM[u];
                    pM[u] = pM[w] * s + t * c;
                    pM[w] = pM[w] * c - t * s;
                }
            }
            for (i = 0; i < nDim; i++) { /* :154-161 */
                int u = i * nDim + nRow, w = i * nDim + nCol;
                float t = V[u];
                V[u] = V[w] * s + t * c;
                V[w] = V[w] * c - t * s;
            }
        }
    }
    {
        int min_id = 0;
        float minEig = pM[0];
        for (i = 1; i < nDim; i++)
            if (minEig > pM[i * nDim + i]) { minEig = pM[i * nDim + i]; min_id = i; }
        for (i = 0; i < nDim; i++) vec_out[i] = V[min_id + nDim * i];
    }
}

/* G_Mapfeature gpu.cu:549-670 */
void orc_map_feature(orc_map *m, float *o_elev, float *o_var, int *o_R, int *o_G, int *o_B,
                     float *o_rough, float *o_slope, float *o_traver, float *o_intensity)
{
    int L = m->L, idx;
    /* the kernel reads neighbours' map_elevation while other threads only write map_traver,
     * so a sequential sweep is equivalent */
    for (idx = 0; idx < L * L; idx++) {
        float px[25], py[25], pz[25];
        float px_mean = 0, py_mean = 0, pz_mean = 0;
        int cell_x = idx / L, cell_y = idx % L, p_n = 0, i, j;
        if (o_elev) o_elev[idx] = m->elevation[idx];
        if (o_R) o_R[idx] = m->colorR[idx];
        if (o_G) o_G[idx] = m->colorG[idx];
        if (o_B) o_B[idx] = m->colorB[idx];
        if (o_intensity) o_intensity[idx] = m->intensity[idx];
        if (o_var) o_var[idx] = m->variance[idx];
        if (m->elevation[idx] == -10) { /* :581 early return; outputs uninitialised in ref */
            if (o_rough) o_rough[idx] = 0;
            if (o_slope) o_slope[idx] = 0;
            if (o_traver) o_traver[idx] = -10;
            continue;
        }
        for (i = -2; i < 3; i++)
            for (j = -2; j < 3; j++) {
                int Ele_x = (cell_x + L - m->start[0]) % L + i;
                int Ele_y = (cell_y + L - m->start[1]) % L + j;
                if (Ele_x >= 0 && Ele_x < L && Ele_y >= 0 && Ele_y < L) {
                    int qx = (cell_x + i + L) % L, qy = (cell_y + j + L) % L;
                    float s_z = m->elevation[qx * L + qy];
                    if (s_z != -10) {
                        px[p_n] = (float)qx * m->res; /* :606 int * float */
                        py[p_n] = (float)qy * m->res;
                        pz[p_n] = s_z;
                        px_mean = px_mean + px[p_n];
                        py_mean = py_mean + py[p_n];
                        pz_mean = pz_mean + pz[p_n];
                        p_n++;
                    }
                }
            }
        if (p_n > 7) {
            float M[9] = {0}, nv[3], Slope, Rough, Traver, height;
            px_mean = px_mean / (float)p_n;
            py_mean = py_mean / (float)p_n;
            pz_mean = pz_mean / (float)p_n;
            for (i = 0; i < p_n; i++) { /* :626-637 */
                float dx = px[i] - px_mean, dy = py[i] - py_mean, dz = pz[i] - pz_mean;
                M[0] = M[0] + dx * dx;
                M[4] = M[4] + dy * dy;
                M[8] = M[8] + dz * dz;
                M[1] = M[1] + dx * dy;
                M[2] = M[2] + dx * dz;
                M[5] = M[5] + dy * dz;
                M[3] = M[1]; M[6] = M[2]; M[7] = M[5];
            }
            jacobi_min_eigvec(M, nv, 0.01f, 30);
            height = m->elevation[idx];
            if (nv[2] > 0) Slope = orc_acosf(nv[2]);
            else Slope = orc_acosf(-nv[2]);
            Rough = fabsf(height - pz_mean);
            /* :655 all-double expression stored to float */
            Traver = (float)(0.5 * (1.0 - (double)Slope / 0.6) + 0.5 * (1.0 - ((double)Rough / 0.2)));
            if (o_slope) o_slope[idx] = Slope;
            if (o_rough) o_rough[idx] = Rough;
            if (o_traver) o_traver[idx] = Traver;
            m->traver[idx] = Traver;
        } else {
            if (o_slope) o_slope[idx] = 0;
            if (o_rough) o_rough[idx] = 0;
            if (o_traver) o_traver[idx] = -10;
            m->traver[idx] = -10;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* G_Raytracing gpu.cu:708-891                                                          */
/* ------------------------------------------------------------------------------------ */
static int p_is_valid(const orc_map *m, int cx, int cy)
{ /* :682-690 */
    return m->lowest[cx * m->L + cy] != 10;
}
static float d_min_elevation(const orc_map *m, int cx, int cy, int ox, float robot)
{ /* :692-706 */
    float x1 = (float)(cx - ox);
    float x2 = (float)cx - robot;
    float low = m->lowest[cx * m->L + cy];
    float h2 = m->sensorZ - low;
    return low + h2 / x2 * x1;
}
void orc_raytracing(orc_map *m)
{
    int L = m->L, i;
    size_t C = (size_t)L * L, c;
    /* threads only write their own map_elevation[i] and read traver/lowest/variance of
     * other cells, never elevation of other cells => sequential sweep is equivalent */
    for (i = 0; i < L * L; i++) {
        int cell_x, cell_y, robot_index, ob[2], cur[2], inc_x, inc_y;
        float obstacle_ele, inc[2], restrict_ele, max_ele;
        if (!(m->traver[i] < m->obstacle_threshold && m->elevation[i] != -10)) continue;
        cell_x = i / L;
        cell_y = i % L;
        ob[0] = (cell_x + L - m->start[0]) % L; /* :672-675 */
        ob[1] = (cell_y + L - m->start[1]) % L;
        obstacle_ele = m->elevation[i];
        cur[0] = ob[0];
        cur[1] = ob[1];
        if (L % 2 == 0) robot_index = f2i_rz((float)((double)(L / 2) - 0.5)); /* :733 */
        else robot_index = f2i_rz((float)(L / 2));                            /* :739 */
        inc[0] = (float)(ob[0] - robot_index);
        inc[1] = (float)(ob[1] - robot_index);
        inc_x = inc[0] > 0 ? 1 : (inc[0] == 0 ? 0 : -1);
        inc_y = inc[1] > 0 ? 1 : (inc[1] == 0 ? 0 : -1);
        restrict_ele = obstacle_ele;
        if (inc_x == 0 && inc_y == 0) continue;
        /* :762-793 axis-aligned rays compute the restriction and then `return` before the
         * removal test -- no observable effect, so nothing to do */
        if (inc_x == 0 || inc_y == 0) continue;
        {
            float dis = sqrtf(inc[0] * inc[0] + inc[1] * inc[1]);
            float dir[2], threshold, dnx, dny, bx, by, later = 0;
            dir[0] = inc[0] / dis;
            dir[1] = inc[1] / dis;
            if (fabsf(inc[0]) > fabsf(inc[1])) { /* :801-804 double expression */
                double t = 0.5 / (double)inc[0] * (double)inc[1];
                threshold = (float)sqrt(0.5 * 0.5 + t * t);
            } else {
                double t = 0.5 / (double)inc[1] * (double)inc[0];
                threshold = (float)sqrt(0.5 * 0.5 + t * t);
            }
            bx = (float)inc_x / 2;
            by = (float)inc_y / 2;
            dnx = bx / dir[0];
            dny = by / dir[1];
            while (cur[0] >= 0 && cur[0] < L && cur[1] >= 0 && cur[1] < L) {
                if (dnx > dny) {
                    if (dny - later > threshold && cur[0] != ob[0] && cur[1] != ob[1]) {
                        if (p_is_valid(m, cur[0], cur[1])) {
                            max_ele = d_min_elevation(m, cur[0], cur[1], ob[0], (float)robot_index);
                            if (max_ele < restrict_ele) restrict_ele = max_ele;
                        }
                    }
                    cur[1] += inc_y;
                    by += (float)inc_y;
                    later = dny;
                    dny = by / dir[1];
                } else if (dnx < dny) {
                    if (dnx - later > threshold && cur[0] != ob[0] && cur[1] != ob[1]) {
                        if (p_is_valid(m, cur[0], cur[1])) {
                            max_ele = d_min_elevation(m, cur[0], cur[1], ob[0], (float)robot_index);
                            if (max_ele < restrict_ele) restrict_ele = max_ele;
                        }
                    }
                    cur[0] += inc_x;
                    bx += (float)inc_x;
                    later = dnx;
                    dnx = bx / dir[0];
                } else {
                    if (dnx - later > threshold && cur[0] != ob[0] && cur[1] != ob[1]) {
                        if (p_is_valid(m, cur[0], cur[1])) {
                            max_ele = d_min_elevation(m, cur[0], cur[1], ob[0], (float)robot_index);
                            if (max_ele < restrict_ele) restrict_ele = max_ele;
                        }
                    }
                    cur[0] += inc_x;
                    cur[1] += inc_y;
                    bx += (float)inc_x;
                    by += (float)inc_y;
                    later = dnx;
                    dnx = bx / dir[0];
                    dny = by / dir[1];
                }
            }
            if (obstacle_ele - 3 * sqrtf(m->variance[i]) > restrict_ele) m->elevation[i] = -10; /* :885 */
        }
    }
    for (c = 0; c < C; c++) m->lowest[c] = 10; /* G_Clear_maplowest :232-239 */
}

/* ------------------------------------------------------------------------------------ */
/* Map_optmove gpu.cu:1215-1233, Map_closeloop gpu.cu:1235-1254                         */
/* ------------------------------------------------------------------------------------ */
static void update_mapheight(orc_map *m, float dz)
{ /* G_update_mapheight :1195-1202 */
    size_t C = (size_t)m->L * m->L, c;
    for (c = 0; c < C; c++)
        if (m->elevation[c] != -10) m->elevation[c] += dz;
}
void orc_optmove(orc_map *m, const float opt_p[2], float height_update, float aligned_out[2])
{
    int i;
    for (i = 0; i < 2; i++) { /* alignedPosition :1203-1213 */
        float ps = opt_p[i] - m->centre[i];
        int is = d2i_rz((double)(ps / m->res) + 0.5 * (ps > 0 ? 1 : -1));
        float a = m->centre[i] + m->res * (float)is;
        if (aligned_out) aligned_out[i] = a;
        m->centre[i] = a;
    }
    update_mapheight(m, height_update);
}
void orc_closeloop(orc_map *m, const float up[2], float height_update)
{
    int i;
    for (i = 0; i < 2; i++) {
        float ps = up[i] - m->centre[i];
        int is = d2i_rz((double)(ps / m->res) + 0.5 * (ps > 0 ? 1 : -1));
        float aligned = (float)is * m->res;
        m->centre[i] = position_to_range(m->centre[i], aligned, m->res);
    }
    update_mapheight(m, height_update);
}

/* ------------------------------------------------------------------------------------ */
/* ElevationMap::show, ElevationMap.cpp:85-149: orthomosaic + visual cloud               */
/* ------------------------------------------------------------------------------------ */
/* Inputs are Map_feature's output arrays, as show() receives them.  Cells are visited in GridMapIterator order
 * (linear index i -> index_x = i % L, index_y = i / L: grid_map's column-major buffer order).
 * The image part is plain index arithmetic (:123-125).  The point position comes from grid_map::GridMap::getPosition,
 * an un-vendored dependency (ANYbotics/grid_map GridMapMath.cpp getPositionFromIndex, version not pinned by the
 * reference's package.xml); ORACLE DEFINITION restating its published algorithm:
 *   position = mapPosition + length/2 - resolution/2 - resolution * unwrappedIndex   (double), stored as float,
 * with mapPosition / startIndex = what Move returned (ElevationMap.cpp:172-177). */
void orc_show(const orc_map *m, double grid_res, const float *elevation, const float *traver, const int *R,
              const int *G, const int *B, unsigned char *bgr, float *xyz, unsigned char *rgb, int *count)
{
    const int L = m->L;
    /* the node's resolution_ is a double (ElevationMapping.hpp:314) and grid_map computes with it; 0 = the float */
    const double res = grid_res > 0.0 ? grid_res : (double)m->res, half = 0.5 * ((double)L * res) - 0.5 * res;
    int n = 0, i;
    if (bgr) memset(bgr, 0, (size_t)L * L * 3);                      /* :87 cv::Scalar(0,0,0) */
    for (i = 0; i < L * L; i++) {
        const int ix = i % L, iy = i / L;
        const int index = ix * L + iy;                                /* :99 */
        if (elevation[index] != -10 && traver[index] != -10 && !isnan(traver[index])) { /* :101 */
            /* :108-110 the int colour becomes a float layer value, :118-120,123-125 then an unsigned char */
            const unsigned char r = (unsigned char)(float)R[index], g = (unsigned char)(float)G[index],
                                b = (unsigned char)(float)B[index];
            const int ux = (ix + L - m->start[0]) % L, uy = (iy + L - m->start[1]) % L;
            if (bgr) {
                unsigned char *px = bgr + 3 * ((size_t)ux * L + uy);
                px[0] = b; px[1] = g; px[2] = r;
            }
            if (xyz) {
                xyz[3 * n + 0] = (float)((double)m->centre[0] + half - res * (double)ux);
                xyz[3 * n + 1] = (float)((double)m->centre[1] + half - res * (double)uy);
                xyz[3 * n + 2] = elevation[index];
            }
            if (rgb) { rgb[3 * n + 0] = r; rgb[3 * n + 1] = g; rgb[3 * n + 2] = b; }
            n++;
        }
    }
    if (count) *count = n;
}

/* ------------------------------------------------------------------------------------ */
/* scroll-out harvest into the submap store, ElevationMapping.cpp:716-765                */
/* ------------------------------------------------------------------------------------ */
/* prevMap_ is the visualMap_ of the previous frame (:422): the Map_feature outputs of that frame masked by show()
 * (ElevationMap.cpp:101, NaN elsewhere) with that frame's centre / start index.  Every cell with
 * elevation != -10 && traver >= 0 (:725) whose centre lies outside the CURRENT window on the side(s) selected by the
 * signs of the last position shift (:726-733) is emitted in GridMapIterator order as one PointXYZRGBICT record
 * {x, y, elevation, 1 | bgra, variance, intensity, traver} (:748-759; GridPointData :736-737 holds the same values).
 * Cell-centre positions: grid_map arithmetic as in orc_show (ORACLE DEFINITION).  data[3] = 1 and a = 255 are
 * defined here; the reference leaves them uninitialised (PointXYZRGBICT.hpp:35). */
void orc_harvest(int L, double grid_res, const float centre_prev[2], const int start_prev[2], const float *elevation,
                 const float *variance, const float *traver, const int *R, const int *G, const int *B,
                 const float *intensity, const float current[2], const float shift[2], float *out, int *count)
{
    const double res = grid_res, half = 0.5 * ((double)L * res) - 0.5 * res;
    const double halfwin = (double)L * res / 2;                       /* length_ * resolution_ / 2 */
    const double lox = (double)current[0] - halfwin, hix = (double)current[0] + halfwin;
    const double loy = (double)current[1] - halfwin, hiy = (double)current[1] + halfwin;
    const float dx = shift[0], dy = shift[1];
    int n = 0, i;
    for (i = 0; i < L * L; i++) {
        const int ix = i % L, iy = i / L, index = ix * L + iy;
        const int shown = elevation[index] != -10 && traver[index] != -10 && !isnan(traver[index]);
        double x, y;
        if (!shown || !(traver[index] >= 0.0)) continue;              /* :725 on the NaN-masked layers */
        x = (double)centre_prev[0] + half - res * (double)((ix + L - start_prev[0]) % L);
        y = (double)centre_prev[1] + half - res * (double)((iy + L - start_prev[1]) % L);
        if (((x < lox || y < loy) && (dx > 0 && dy > 0)) || ((x > hix || y > hiy) && (dx < 0 && dy < 0)) ||
            ((x < lox || y > hiy) && (dx > 0 && dy < 0)) || ((x > hix || y < loy) && (dx < 0 && dy > 0)) ||
            ((x < lox) && (dx > 0 && dy == 0)) || ((x > hix) && (dx < 0 && dy == 0)) ||
            ((y < loy) && (dy > 0 && dx == 0)) || ((y > hiy) && (dy < 0 && dx == 0))) {
            if (out) {
                float *o = out + 8 * (size_t)n;
                const unsigned r = (unsigned char)(float)R[index], g = (unsigned char)(float)G[index],
                               b = (unsigned char)(float)B[index];
                const unsigned bgra = b | (g << 8) | (r << 16) | 0xff000000u;
                o[0] = (float)x; o[1] = (float)y; o[2] = elevation[index]; o[3] = 1.0f;
                memcpy(&o[4], &bgra, 4);
                o[5] = variance[index]; o[6] = intensity[index]; o[7] = traver[index];
            }
            n++;
        }
    }
    if (count) *count = n;
}

/* ------------------------------------------------------------------------------------ */
/* colourisation, ElevationMapping.cpp:331-381 (CPU loop of the ROS node)               */
/* ------------------------------------------------------------------------------------ */
void orc_colourise(float *xyzi, int n, const double Tc[12], const double Tl[16], const unsigned char *bgr, int width,
                   int height, int row_stride, unsigned char *rgba_out)
{
    double P[12];
    int i, j, k;
    for (i = 0; i < 3; i++) /* :347 P_lidar2img = Tcamera * TLidar */
        for (j = 0; j < 4; j++) {
            double a = Tc[4 * i + 0] * Tl[0 + j];
            for (k = 1; k < 4; k++) a = a + Tc[4 * i + k] * Tl[4 * k + j];
            P[4 * i + j] = a;
        }
    for (i = 0; i < n; i++) {
        double x = (double)xyzi[4 * i], y = (double)xyzi[4 * i + 1], z = (double)xyzi[4 * i + 2];
        double X = ((P[0] * x + P[1] * y) + P[2] * z) + P[3] * 1.0; /* :351-355 */
        double Y = ((P[4] * x + P[5] * y) + P[6] * z) + P[7] * 1.0;
        double Z = ((P[8] * x + P[9] * y) + P[10] * z) + P[11] * 1.0;
        float Px = (float)(X / Z), Py = (float)(Y / Z); /* :359-360 */
        int mx = f2i_rz(Px), my = f2i_rz(Py);           /* :364-365 */
        unsigned char *o = rgba_out + 4 * (size_t)i;
        if (mx > 0 && mx < width && my > 0 && my < height && Z > 0) { /* :368 */
            const unsigned char *px = bgr + (size_t)my * row_stride + 3 * (size_t)mx;
            o[0] = px[2]; o[1] = px[1]; o[2] = px[0]; o[3] = 255;
        } else { /* :376-381 */
            o[0] = o[1] = o[2] = o[3] = 0;
            xyzi[4 * i + 3] = 0;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* multi-threaded CPU baseline (bench.py cpu_baseline / --impl reference only)          */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    orc_map *m;
    int n, tid, nthreads;
    const float *xyzi;
    const unsigned char *rgba;
    const float *T;
    double lo, hi;
    const orc_sensor *sensor;
    const float *sJ;
    int *key, *geo;
    float *h, *hv;
    pthread_barrier_t *bar;
} mt_ctx;

static void *mt_worker(void *arg)
{
    mt_ctx *c = (mt_ctx *)arg;
    orc_map *m = c->m;
    static const float Z9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    static const float Z3[3] = {0, 0, 0};
    int L = m->L, i;
    int lo_i = (int)((long)c->n * c->tid / c->nthreads);
    int hi_i = (int)((long)c->n * (c->tid + 1) / c->nthreads);
    int row_lo = (int)((long)L * c->tid / c->nthreads);
    int row_hi = (int)((long)L * (c->tid + 1) / c->nthreads);
    /* phase 1: per-point transform, contiguous slices */
    for (i = lo_i; i < hi_i; i++) {
        pt_result r;
        process_one(m, c->xyzi[4 * i], c->xyzi[4 * i + 1], c->xyzi[4 * i + 2], c->T, c->lo, c->hi,
                    c->sensor, c->sJ, Z9, Z9, Z3, Z9, &r);
        c->key[i] = r.key;
        c->geo[i] = r.accepted ? r.geo : -1;
        c->h[i] = r.h;
        c->hv[i] = r.hv;
    }
    pthread_barrier_wait(c->bar);
    /* phase 2: every thread owns a band of rows; scanning points in index order keeps the
     * per-cell order of G_fuse.  lowest (geographic rows) and layers (storage rows). */
    {
        /* lowest: two sweeps (min, then first index attaining it) using the layer itself is
         * not possible without scratch, so use the ORACLE DEFINITION incrementally:
         * track per-cell (min, argmin) in thread-local scratch for the owned band. */
        int band = row_hi - row_lo;
        float *minh = (float *)malloc((size_t)(band > 0 ? band : 1) * L * sizeof(float));
        int *amin = (int *)malloc((size_t)(band > 0 ? band : 1) * L * sizeof(int));
        long k, bn = (long)band * L;
        for (k = 0; k < bn; k++) amin[k] = -1;
        for (i = 0; i < c->n; i++) {
            int g = c->geo[i];
            if (g >= 0) {
                int gr = g / L;
                if (gr >= row_lo && gr < row_hi) {
                    long q = g - (long)row_lo * L;
                    if (amin[q] < 0 || c->h[i] < minh[q]) { amin[q] = i; minh[q] = c->h[i]; }
                }
            }
            {
                int s = c->key[i];
                if (s >= 0 && c->h[i] != -1) {
                    int sr = s / L;
                    if (sr >= row_lo && sr < row_hi) {
                        int R = 0, G = 0, B = 0;
                        if (c->rgba) { R = c->rgba[4 * i]; G = c->rgba[4 * i + 1]; B = c->rgba[4 * i + 2]; }
                        fuse_one(m, s, R, G, B, c->xyzi[4 * i + 3], c->h[i], c->hv[i]);
                    }
                }
            }
        }
        for (k = 0; k < bn; k++)
            if (amin[k] >= 0) {
                long g = k + (long)row_lo * L;
                if (minh[k] <= m->lowest[g]) m->lowest[g] = minh[k] + 3 * c->hv[amin[k]];
            }
        /* variance floor over the owned band (gpu.cu:533-534) */
        for (k = (long)row_lo * L; k < (long)row_hi * L; k++)
            if ((double)m->variance[k] < 0.0001) m->variance[k] = (float)0.0001;
        free(minh);
        free(amin);
    }
    return NULL;
}

void orc_add_points_mt(orc_map *m, int n, const float *xyzi, const unsigned char *rgba,
                       const float T[16], double relLower, double relUpper,
                       const orc_sensor *sensor, const float sJ[3], int nthreads)
{
    pthread_t *th;
    mt_ctx *ctx;
    pthread_barrier_t bar;
    int t;
    int *key = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    int *geo = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    float *h = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    float *hv = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    if (nthreads < 1) nthreads = 1;
    if (nthreads > m->L) nthreads = m->L;
    th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    ctx = (mt_ctx *)malloc(sizeof(mt_ctx) * nthreads);
    pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
    for (t = 0; t < nthreads; t++) {
        mt_ctx c = {m, n, t, nthreads, xyzi, rgba, T, relLower, relUpper, sensor, sJ, key, geo, h, hv, &bar};
        ctx[t] = c;
        if (t > 0) pthread_create(&th[t], NULL, mt_worker, &ctx[t]);
    }
    mt_worker(&ctx[0]);
    for (t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    free(th); free(ctx); free(key); free(geo); free(h); free(hv);
}

/* ------------------------------------------------------------------------------------ */
/* pooled multi-threaded CPU baseline: persistent workers, per-band point lists          */
/* (bench.py cpu_baseline / --impl reference only).  Same results as orc_add_points_mt   */
/* and as the single-threaded process_points + fuse (tests/test_oracle_kats.py).         */
/*   phase 1  every thread transforms a contiguous slice of the cloud and counts, per     */
/*            band of map rows, its points (storage rows for the fold, geographic rows    */
/*            for the lowest layer);                                                      */
/*   phase 2  the slice's point indices are written into the bands' lists: band-major,    */
/*            then thread-major, then slice order = ascending point index inside a band   */
/*            = the per-cell visiting order of G_fuse;                                    */
/*   phase 3  bands are drawn from a shared counter (more bands than threads: the points  */
/*            of a frame sit near the sensor) and folded by ONE thread each.              */
/* ------------------------------------------------------------------------------------ */
#include <stdatomic.h>
struct orc_pool {
    int nthreads, nbands;
    pthread_t *th;
    pthread_barrier_t bar;
    int quit;
    /* the job */
    orc_map *m; int n;
    const float *xyzi; const unsigned char *rgba; const float *T; double lo, hi; const orc_sensor *sensor; const float *sJ;
    /* scratch */
    int cap; int *key, *geo, *idxS, *idxG; float *h, *hv;
    long *cntS, *cntG;   /* [nbands + 1][nthreads] -> offsets */
    size_t cells; float *minh; int *amin; unsigned *seen; unsigned call_id;
    atomic_int next_band;
};
typedef struct { orc_pool *p; int tid; } pool_arg;

static int band_of(int row, int L, int nbands) { return (int)((long)row * nbands / L); }
static void pool_job(orc_pool *p, int tid)
{
    static const float Z9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    static const float Z3[3] = {0, 0, 0};
    orc_map *m = p->m;
    const int L = m->L, nt = p->nthreads, nb = p->nbands;
    const int lo_i = (int)((long)p->n * tid / nt), hi_i = (int)((long)p->n * (tid + 1) / nt);
    int i, b;
    for (b = 0; b < nb; b++) { p->cntS[(long)b * nt + tid] = 0; p->cntG[(long)b * nt + tid] = 0; }
    for (i = lo_i; i < hi_i; i++) { /* phase 1 */
        pt_result r;
        process_one(m, p->xyzi[4 * i], p->xyzi[4 * i + 1], p->xyzi[4 * i + 2], p->T, p->lo, p->hi, p->sensor, p->sJ, Z9, Z9, Z3, Z9, &r);
        p->key[i] = (r.key >= 0 && r.h != -1) ? r.key : -1;
        p->geo[i] = r.accepted ? r.geo : -1;
        p->h[i] = r.h;
        p->hv[i] = r.hv;
        if (p->key[i] >= 0) p->cntS[(long)band_of(p->key[i] / L, L, nb) * nt + tid]++;
        if (p->geo[i] >= 0) p->cntG[(long)band_of(p->geo[i] / L, L, nb) * nt + tid]++;
    }
    pthread_barrier_wait(&p->bar);
    if (tid == 0) { /* exclusive prefix in (band, thread) order */
        long accS = 0, accG = 0, k, tot = (long)nb * nt;
        for (k = 0; k < tot; k++) {
            long c = p->cntS[k]; p->cntS[k] = accS; accS += c;
            c = p->cntG[k]; p->cntG[k] = accG; accG += c;
        }
        p->cntS[tot] = accS; p->cntG[tot] = accG;
        atomic_store(&p->next_band, 0);
    }
    pthread_barrier_wait(&p->bar);
    { /* phase 2: this thread's slots of every band */
        long *wS = (long *)malloc(sizeof(long) * nb), *wG = (long *)malloc(sizeof(long) * nb);
        for (b = 0; b < nb; b++) { wS[b] = p->cntS[(long)b * nt + tid]; wG[b] = p->cntG[(long)b * nt + tid]; }
        for (i = lo_i; i < hi_i; i++) {
            if (p->key[i] >= 0) p->idxS[wS[band_of(p->key[i] / L, L, nb)]++] = i;
            if (p->geo[i] >= 0) p->idxG[wG[band_of(p->geo[i] / L, L, nb)]++] = i;
        }
        free(wS); free(wG);
    }
    pthread_barrier_wait(&p->bar);
    for (;;) { /* phase 3 */
        long k, k0, k1, c;
        int row_lo, row_hi;
        b = atomic_fetch_add(&p->next_band, 1);
        if (b >= nb) break;
        k0 = p->cntS[(long)b * nt]; k1 = p->cntS[(long)(b + 1) * nt];
        for (k = k0; k < k1; k++) {
            int R = 0, G = 0, B = 0;
            i = p->idxS[k];
            if (p->rgba) { R = p->rgba[4 * i]; G = p->rgba[4 * i + 1]; B = p->rgba[4 * i + 2]; }
            fuse_one(m, p->key[i], R, G, B, p->xyzi[4 * i + 3], p->h[i], p->hv[i]);
        }
        k0 = p->cntG[(long)b * nt]; k1 = p->cntG[(long)(b + 1) * nt];
        for (k = k0; k < k1; k++) { /* lowest: minimum of the call and the first index attaining it (ORACLE DEFINITION) */
            int g;
            i = p->idxG[k]; g = p->geo[i];
            if (p->seen[g] != p->call_id) { p->seen[g] = p->call_id; p->amin[g] = i; p->minh[g] = p->h[i]; }
            else if (p->h[i] < p->minh[g]) { p->amin[g] = i; p->minh[g] = p->h[i]; }
        }
        for (k = k0; k < k1; k++) {
            int g;
            i = p->idxG[k]; g = p->geo[i];
            if (p->amin[g] == i && p->minh[g] <= m->lowest[g]) m->lowest[g] = p->minh[g] + 3 * p->hv[i];
        }
        /* rows r with band_of(r) == b: [ceil(b L / nb), ceil((b + 1) L / nb)) */
        row_lo = (int)(((long)b * L + nb - 1) / nb); row_hi = (int)(((long)(b + 1) * L + nb - 1) / nb);
        for (c = (long)row_lo * L; c < (long)row_hi * L; c++) /* variance floor over the band (gpu.cu:533-534) */
            if ((double)m->variance[c] < 0.0001) m->variance[c] = (float)0.0001;
    }
}
static void *pool_worker(void *a)
{
    pool_arg *pa = (pool_arg *)a;
    orc_pool *p = pa->p;
    const int tid = pa->tid;
    free(pa);
    for (;;) {
        pthread_barrier_wait(&p->bar); /* job posted */
        if (p->quit) break;
        pool_job(p, tid);
        pthread_barrier_wait(&p->bar); /* job done */
    }
    return NULL;
}
orc_pool *orc_pool_create(int nthreads)
{
    orc_pool *p = (orc_pool *)calloc(1, sizeof(orc_pool));
    int t;
    if (nthreads < 1) nthreads = 1;
    p->nthreads = nthreads;
    p->th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    pthread_barrier_init(&p->bar, NULL, (unsigned)nthreads);
    for (t = 1; t < nthreads; t++) {
        pool_arg *a = (pool_arg *)malloc(sizeof(pool_arg));
        a->p = p; a->tid = t;
        pthread_create(&p->th[t], NULL, pool_worker, a);
    }
    return p;
}
void orc_pool_destroy(orc_pool *p)
{
    int t;
    if (!p) return;
    p->quit = 1;
    pthread_barrier_wait(&p->bar);
    for (t = 1; t < p->nthreads; t++) pthread_join(p->th[t], NULL);
    pthread_barrier_destroy(&p->bar);
    free(p->th); free(p->key); free(p->geo); free(p->idxS); free(p->idxG); free(p->h); free(p->hv);
    free(p->cntS); free(p->cntG); free(p->minh); free(p->amin); free(p->seen);
    free(p);
}
void orc_add_points_pool(orc_pool *p, orc_map *m, int n, const float *xyzi, const unsigned char *rgba, const float T[16],
                         double relLower, double relUpper, const orc_sensor *sensor, const float sJ[3])
{
    const size_t cells = (size_t)m->L * m->L;
    int nb = 8 * p->nthreads;
    if (nb > m->L) nb = m->L;
    if (n > p->cap) {
        p->cap = n + n / 8 + 1024;
        free(p->key); free(p->geo); free(p->idxS); free(p->idxG); free(p->h); free(p->hv);
        p->key = (int *)malloc(sizeof(int) * p->cap); p->geo = (int *)malloc(sizeof(int) * p->cap);
        p->idxS = (int *)malloc(sizeof(int) * p->cap); p->idxG = (int *)malloc(sizeof(int) * p->cap);
        p->h = (float *)malloc(sizeof(float) * p->cap); p->hv = (float *)malloc(sizeof(float) * p->cap);
    }
    if (nb != p->nbands || !p->cntS) {
        p->nbands = nb;
        free(p->cntS); free(p->cntG);
        p->cntS = (long *)malloc(sizeof(long) * ((size_t)nb * p->nthreads + 1));
        p->cntG = (long *)malloc(sizeof(long) * ((size_t)nb * p->nthreads + 1));
    }
    if (cells != p->cells) {
        p->cells = cells;
        free(p->minh); free(p->amin); free(p->seen);
        p->minh = (float *)malloc(sizeof(float) * cells); p->amin = (int *)malloc(sizeof(int) * cells);
        p->seen = (unsigned *)calloc(cells, sizeof(unsigned));
        p->call_id = 0;
    }
    p->call_id++;
    if (p->call_id == 0) { memset(p->seen, 0, sizeof(unsigned) * cells); p->call_id = 1; }
    p->m = m; p->n = n; p->xyzi = xyzi; p->rgba = rgba; p->T = T; p->lo = relLower; p->hi = relUpper; p->sensor = sensor; p->sJ = sJ;
    pthread_barrier_wait(&p->bar); /* post */
    pool_job(p, 0);
    pthread_barrier_wait(&p->bar); /* done */
}

/* ------------------------------------------------------------------------------------ */
/* loop-closure submap re-fusion: ElevationMapping::updateGlobalMap, ElevationMapping.cpp:773-905                      */
/* (PARITY UNPINNED: needs PCL / ROS; restated from the source, with the DEFINITIONS of DESIGN.md "f4" where the        */
/* reference depends on unordered_map iteration order or uninitialised fields).  Points are 8 floats:                   */
/* {x, y, z, w, bgra bits, covariance, intensity, travers} (PointXYZRGBICT.hpp:26-48).                                   */
/* ------------------------------------------------------------------------------------ */
void orc_transform_cloud(float *pts, int n, const float T[16])
{ /* :805 pcl::transformPointCloud, scalar form, left to right */
    int i;
    for (i = 0; i < n; i++) {
        float x = pts[8 * i], y = pts[8 * i + 1], z = pts[8 * i + 2];
        pts[8 * i] = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        pts[8 * i + 1] = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        pts[8 * i + 2] = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
    }
}
typedef struct { float rx, ry; int idx; } cell_ref;
static int cell_cmp(const void *a, const void *b)
{
    const cell_ref *p = (const cell_ref *)a, *q = (const cell_ref *)b;
    if (p->rx != q->rx) return p->rx < q->rx ? -1 : 1;
    if (p->ry != q->ry) return p->ry < q->ry ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}
/* the cells of a cloud: sorted (cell, first index) table of the non-NaN points; keep[i] = 1 for the first point of every
 * cell (and for NaN-positioned points, which equal nothing) */
static int build_cells(const float *pts, int n, double res, cell_ref *tab, unsigned char *keep, float *rxs, float *rys)
{
    int i, m = 0, u = 0;
    for (i = 0; i < n; i++) { /* pointCloudtoHash :1183-1184 */
        float rx = (float)(ceil((double)pts[8 * i] / res) * res - res / 2.0);
        float ry = (float)(ceil((double)pts[8 * i + 1] / res) * res - res / 2.0);
        rxs[i] = rx; rys[i] = ry;
        keep[i] = 0;
        if (rx != rx || ry != ry) { keep[i] = 1; continue; }
        tab[m].rx = rx; tab[m].ry = ry; tab[m].idx = i; m++;
    }
    qsort(tab, (size_t)m, sizeof(cell_ref), cell_cmp);
    for (i = 0; i < m; i++) /* insert() keeps the first point of a cell */
        if (i == 0 || tab[i].rx != tab[i - 1].rx || tab[i].ry != tab[i - 1].ry) { keep[tab[i].idx] = 1; tab[u++] = tab[i]; }
    return u;
}
static int find_cell(const cell_ref *tab, int m, float rx, float ry)
{
    int lo = 0, hi = m - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if (tab[mid].rx == rx && tab[mid].ry == ry) return tab[mid].idx;
        if (tab[mid].rx < rx || (tab[mid].rx == rx && tab[mid].ry < ry)) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}
static int compact_kept(float *pts, int n, const unsigned char *keep, const float *rxs, const float *rys)
{
    int i, k = 0;
    for (i = 0; i < n; i++)
        if (keep[i]) {
            if (k != i) memcpy(pts + 8 * k, pts + 8 * i, 32);
            pts[8 * k] = rxs[i]; pts[8 * k + 1] = rys[i]; pts[8 * k + 3] = 1.0f; /* localHashtoPointCloud :1129-1130 */
            k++;
        }
    return k;
}
int orc_refuse_submaps(float *pn, int *n_new, float *po, int *n_old, double res, int compat)
{
    const int nn = *n_new, no = *n_old;
    cell_ref *tn = (cell_ref *)malloc(sizeof(cell_ref) * (size_t)(nn + 1)), *to = (cell_ref *)malloc(sizeof(cell_ref) * (size_t)(no + 1));
    unsigned char *kn = (unsigned char *)malloc((size_t)nn + 1), *ko = (unsigned char *)malloc((size_t)no + 1);
    float *rxn = (float *)malloc(4 * (size_t)(nn + 1)), *ryn = (float *)malloc(4 * (size_t)(nn + 1));
    float *rxo = (float *)malloc(4 * (size_t)(no + 1)), *ryo = (float *)malloc(4 * (size_t)(no + 1));
    const int mo = build_cells(po, no, res, to, ko, rxo, ryo);
    int i, count = 0;
    (void)build_cells(pn, nn, res, tn, kn, rxn, ryn);
    for (i = 0; i < nn; i++) { /* every cell of the new map, once (:847) */
        int j;
        if (!kn[i] || rxn[i] != rxn[i] || ryn[i] != ryn[i]) continue;
        j = find_cell(to, mo, rxn[i], ryn[i]);
        if (j >= 0 && po[8 * j + 5] > 0 && po[8 * j + 5] < 1) { /* :857 */
            const float vo = po[8 * j + 5], eo = po[8 * j + 2], vn = pn[8 * i + 5], en = pn[8 * i + 2];
            const double vn2 = pow((double)vn, 2), vo2 = pow((double)vo, 2);
            float ef, vf;
            if (compat) { /* :862-863, as written */
                ef = (float)(vn2 * eo + vo2 * en / vo2 + vn2);
                vf = (float)(vo2 * vn2 / vo2 + vn2);
            } else {
                ef = (float)((vn2 * eo + vo2 * en) / (vo2 + vn2));
                vf = (float)(vo2 * vn2 / (vo2 + vn2));
            }
            pn[8 * i + 2] = ef; pn[8 * i + 5] = vf;
            memcpy(po + 8 * j, pn + 8 * i, 32); /* tmp_data is built from the NEW map's entry (:856) and inserted into both */
            count++;
        }
    }
    *n_new = compact_kept(pn, nn, kn, rxn, ryn);
    *n_old = compact_kept(po, no, ko, rxo, ryo);
    free(tn); free(to); free(kn); free(ko); free(rxn); free(ryn); free(rxo); free(ryo);
    return count;
}
