/*
 * oracle/ef_oracle.c -- TEST INFRASTRUCTURE ONLY (see ef_oracle.h).
 *
 * CPU restatement of the ElasticFusion dense tracker.  Each function cites the reference
 * file:line it follows (paths relative to elasticfusionpublic/Core/src/).
 *
 * Float arithmetic notes: the reference is compiled with nvcc's default --fmad=true, so
 * `a*b + c` chains contract into FMAs; where that matters for integer-valued outputs
 * (intensity, derivatives, pyramids) this file spells the FMA out with fmaf() so that the
 * rounding sequence is the same.  Approximate GPU ops (div.full, rsqrt.approx,
 * sqrt.approx, rcp.approx) are replaced by IEEE ops.
 */
#include "ef_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* small float3 / mat33 helpers: Cuda/operators.cuh:55-91                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } f3;

static inline f3 f3_make(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 f3_sub(f3 a, f3 b) { return f3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 f3_add(f3 a, f3 b) { return f3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
/* operators.cuh:72-75, contracted the way nvcc 12.9 contracts a*b + c*d + e*f for sm_100a (checked in
 * SASS): t = c*d (FMUL); t = fma(a, b, t); t = fma(e, f, t) */
static inline float f3_dot(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
/* operators.cuh:67-70 */
static inline f3 f3_cross(f3 a, f3 b)
{
    return f3_make(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline float f3_norm(f3 a) { return sqrtf(f3_dot(a, a)); } /* operators.cuh:77-80 */
static inline f3 f3_normalized(f3 a)                             /* operators.cuh:82-86 */
{
    const float rn = 1.0f / sqrtf(f3_dot(a, a));
    return f3_make(a.x * rn, a.y * rn, a.z * rn);
}
/* operators.cuh:88-91 ; m row-major */
static inline f3 m33_mul(const float * m, f3 a)
{
    return f3_make(f3_dot(f3_make(m[0], m[1], m[2]), a), f3_dot(f3_make(m[3], m[4], m[5]), a),
                   f3_dot(f3_make(m[6], m[7], m[8]), a));
}

static inline float qnan_f(void)
{
    union { uint32_t u; float f; } c;
    c.u = 0x7fffffffu; /* cudafuncs.cu:130 */
    return c.f;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* __float2int_rn: round to nearest even, saturating; NaN -> 0 (PTX cvt.rni.s32.f32) */
static inline int float2int_rn(float v)
{
    if(isnan(v)) return 0;
    if(v >= 2147483648.0f) return 2147483647;
    if(v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)nearbyintf(v); /* default rounding mode = to nearest even */
}

/* float -> int truncation as PTX cvt.rzi.s32.f32 (saturating, NaN -> 0) */
static inline int float2int_rz(float v)
{
    if(isnan(v)) return 0;
    if(v >= 2147483648.0f) return 2147483647;
    if(v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:57-107  pyrDownGaussKernel / pyrDown                                         */
/* ------------------------------------------------------------------------------------------ */
void efo_pyr_down_u16(const uint16_t * src, int srows, int scols, uint16_t * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
    const float sigma_color = 30.0f;                       /* :103 */
    const float weights[3] = {0.375f, 0.25f, 0.0625f};     /* :78 */
    const int D = 5;

#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
    {
        for(int x = 0; x < dcols; x++)
        {
            int center = src[(size_t)(2 * y) * scols + 2 * x];            /* :67 */
            int x_mi = imax(0, 2 * x - D / 2) - 2 * x;                    /* :69-73 */
            int y_mi = imax(0, 2 * y - D / 2) - 2 * y;
            int x_ma = imin(scols, 2 * x - D / 2 + D) - 2 * x;
            int y_ma = imin(srows, 2 * y - D / 2 + D) - 2 * y;

            float sum = 0, wall = 0;
            for(int yi = y_mi; yi < y_ma; ++yi)
                for(int xi = x_mi; xi < x_ma; ++xi)
                {
                    int val = src[(size_t)(2 * y + yi) * scols + 2 * x + xi];
                    if((float)abs(val - center) < 3 * sigma_color)        /* :85 */
                    {
                        /* :87-88 ; every product here is exact in binary32 */
                        sum = fmaf((float)val * weights[abs(xi)], weights[abs(yi)], sum);
                        wall = fmaf(weights[abs(xi)], weights[abs(yi)], wall);
                    }
                }
            dst[(size_t)y * dcols + x] = (uint16_t)float2int_rz(sum / wall); /* :93 */
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:109-149  computeVmapKernel / createVMap                                      */
/* ------------------------------------------------------------------------------------------ */
void efo_create_vmap(const uint16_t * depth, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff,
                     float * vmap)
{
    const float fx_inv = 1.f / fx, fy_inv = 1.f / fy; /* :147, host IEEE division */
    const size_t plane = (size_t)rows * cols;

#pragma omp parallel for schedule(static)
    for(int v = 0; v < rows; v++)
        for(int u = 0; u < cols; u++)
        {
            /* :116 `depth / 1000.f` compiles to a multiply by 0x3A83126F under --prec-div=false */
            float z = (float)depth[(size_t)v * cols + u] * 0.001f;
            if(z != 0 && z < cutoff)
            {
                vmap[(size_t)v * cols + u] = z * ((float)u - cx) * fx_inv;              /* :120 */
                vmap[plane + (size_t)v * cols + u] = z * ((float)v - cy) * fy_inv;      /* :121 */
                vmap[2 * plane + (size_t)v * cols + u] = z;
            }
            else
            {
                vmap[(size_t)v * cols + u] = qnan_f();                                   /* :130 x plane only */
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:151-204  computeNmapKernel / createNMap                                      */
/* ------------------------------------------------------------------------------------------ */
void efo_create_nmap(const float * vmap, int rows, int cols, float * nmap)
{
    const size_t plane = (size_t)rows * cols;

#pragma omp parallel for schedule(static)
    for(int v = 0; v < rows; v++)
        for(int u = 0; u < cols; u++)
        {
            size_t i = (size_t)v * cols + u;
            if(u == cols - 1 || v == rows - 1) /* :159 */
            {
                nmap[i] = qnan_f();
                continue;
            }
            f3 v00, v01, v10;
            v00.x = vmap[i];
            v01.x = vmap[i + 1];
            v10.x = vmap[i + cols];
            if(!isnan(v00.x) && !isnan(v01.x) && !isnan(v10.x))
            {
                v00.y = vmap[plane + i];
                v01.y = vmap[plane + i + 1];
                v10.y = vmap[plane + i + cols];
                v00.z = vmap[2 * plane + i];
                v01.z = vmap[2 * plane + i + 1];
                v10.z = vmap[2 * plane + i + cols];
                f3 r = f3_normalized(f3_cross(f3_sub(v01, v00), f3_sub(v10, v00))); /* :180 */
                nmap[i] = r.x;
                nmap[plane + i] = r.y;
                nmap[2 * plane + i] = r.z;
            }
            else
                nmap[i] = qnan_f();
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:206-268  tranformMapsKernel / tranformMaps (called in place)                 */
/* ------------------------------------------------------------------------------------------ */
void efo_transform_maps(const float * vsrc, const float * nsrc, int rows, int cols, const float * R, const float * t,
                        float * vdst, float * ndst)
{
    const size_t plane = (size_t)rows * cols;
    const f3 tv = f3_make(t[0], t[1], t[2]);

#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            size_t i = (size_t)y * cols + x;
            f3 s, d = f3_make(qnan_f(), qnan_f(), qnan_f());
            s.x = vsrc[i];
            if(!isnan(s.x))
            {
                s.y = vsrc[plane + i];
                s.z = vsrc[2 * plane + i];
                d = f3_add(m33_mul(R, s), tv); /* :223 */
                vdst[plane + i] = d.y;
                vdst[2 * plane + i] = d.z;
            }
            vdst[i] = d.x;

            f3 n, nd = f3_make(qnan_f(), qnan_f(), qnan_f());
            n.x = nsrc[i];
            if(!isnan(n.x))
            {
                n.y = nsrc[plane + i];
                n.z = nsrc[2 * plane + i];
                nd = m33_mul(R, n); /* :240 */
                ndst[plane + i] = nd.y;
                ndst[2 * plane + i] = nd.z;
            }
            ndst[i] = nd.x;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:270-330  copyMapsKernel / copyMaps                                           */
/* ------------------------------------------------------------------------------------------ */
void efo_copy_maps(const float * vs, const float * ns, int rows, int cols, float * vmap, float * nmap)
{
    const size_t plane = (size_t)rows * cols;

#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            size_t i = (size_t)y * cols + x;
            const float * v = vs + i * 4;
            const float * n = ns + i * 4;
            const int valid = !(v[2] == 0); /* :285 and :301 both key on the VERTEX z */
            vmap[i] = valid ? v[0] : qnan_f();
            vmap[plane + i] = valid ? v[1] : qnan_f();
            vmap[2 * plane + i] = valid ? v[2] : qnan_f();
            nmap[i] = valid ? n[0] : qnan_f();
            nmap[plane + i] = valid ? n[1] : qnan_f();
            nmap[2 * plane + i] = valid ? n[2] : qnan_f();
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:365-444  resizeMapKernel / resizeVMap / resizeNMap                           */
/* ------------------------------------------------------------------------------------------ */
void efo_resize_map(const float * in, int srows, int scols, float * out, int normalize)
{
    const int drows = srows / 2, dcols = scols / 2;
    const size_t splane = (size_t)srows * scols, dplane = (size_t)drows * dcols;

#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            const size_t s = (size_t)(2 * y) * scols + 2 * x;
            const size_t d = (size_t)y * dcols + x;
            float x00 = in[s], x01 = in[s + 1], x10 = in[s + scols], x11 = in[s + scols + 1];
            if(isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11)) /* :384 */
            {
                out[d] = qnan_f();
                continue;
            }
            f3 n;
            n.x = (x00 + x01 + x10 + x11) / 4; /* :393 */
            n.y = (in[splane + s] + in[splane + s + 1] + in[splane + s + scols] + in[splane + s + scols + 1]) / 4;
            n.z = (in[2 * splane + s] + in[2 * splane + s + 1] + in[2 * splane + s + scols] +
                   in[2 * splane + s + scols + 1]) / 4;
            if(normalize) n = f3_normalized(n); /* :409 */
            out[d] = n.x;
            out[dplane + d] = n.y;
            out[2 * dplane + d] = n.z;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:526-546  verticesToDepthKernel / verticesToDepth                             */
/* ------------------------------------------------------------------------------------------ */
void efo_vertices_to_depth(const float * vs, int rows, int cols, float cutoff, float * dst)
{
#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            float z = vs[((size_t)y * cols + x) * 4 + 2];
            dst[(size_t)y * cols + x] = (z > cutoff || z <= 0) ? qnan_f() : z; /* :536 */
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:332-363, 446-468  pyrDownKernelGaussF / pyrDownGaussF                        */
/* ------------------------------------------------------------------------------------------ */
static const float k_gauss5x5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};

void efo_pyr_down_gauss_f32(const float * src, int srows, int scols, float * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
    const int D = 5;

#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            int tx = imin(2 * x - D / 2 + D, scols - 1); /* :344-346 upper clamp is cols-1 */
            int ty = imin(2 * y - D / 2 + D, srows - 1);
            float sum = 0;
            int count = 0;
            for(int cy = imax(0, 2 * y - D / 2); cy < ty; ++cy)
                for(int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx)
                {
                    float s = src[(size_t)cy * scols + cx];
                    if(!isnan(s))
                    {
                        float k = k_gauss5x5[(ty - cy - 1) * 5 + (tx - cx - 1)]; /* :357 */
                        sum = fmaf(s, k, sum);
                        count = (int)((float)count + k); /* :358 int += float */
                    }
                }
            dst[(size_t)y * dcols + x] = sum / (float)count; /* :362 ; 0/0 -> NaN */
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:470-524  pyrDownKernelIntensityGauss / pyrDownUcharGauss                     */
/* ------------------------------------------------------------------------------------------ */
void efo_pyr_down_gauss_u8(const uint8_t * src, int srows, int scols, uint8_t * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
    const int D = 5;

#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            int tx = imin(2 * x - D / 2 + D, scols - 1);
            int ty = imin(2 * y - D / 2 + D, srows - 1);
            float sum = 0;
            int count = 0;
            for(int cy = imax(0, 2 * y - D / 2); cy < ty; ++cy)
                for(int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx)
                {
                    uint8_t s = src[(size_t)cy * scols + cx];
                    if(s > 0) /* :493 */
                    {
                        float k = k_gauss5x5[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum = fmaf((float)s, k, sum);
                        count = (int)((float)count + k);
                    }
                }
            float q = sum / (float)count; /* :499 ; NaN converts to 0 on the GPU */
            dst[(size_t)y * dcols + x] = (uint8_t)float2int_rz(q);
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:548-577  bgr2IntensityKernel / imageBGRToIntensity                           */
/* ------------------------------------------------------------------------------------------ */
void efo_bgr_to_intensity(const uint8_t * rgba, int rows, int cols, uint8_t * dst)
{
#pragma omp parallel for schedule(static)
    for(int i = 0; i < rows * cols; i++)
    {
        const uint8_t * p = rgba + (size_t)i * 4;
        /* :560  x*0.114f + y*0.299f + z*0.587f ; nvcc: t = y*.299; t = fma(x,.114,t); t = fma(z,.587,t) */
        float v = fmaf((float)p[2], 0.587f, fmaf((float)p[0], 0.114f, (float)p[1] * 0.299f));
        dst[i] = (uint8_t)float2int_rz(v);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:580-639  applyKernel / computeDerivativeImages                               */
/* ------------------------------------------------------------------------------------------ */
void efo_derivative_images(const uint8_t * src, int rows, int cols, int16_t * dx, int16_t * dy)
{
    /* :615-621 (double literals narrowed to float) */
    const float gsx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    const float gsy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};

#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            float dxv = 0, dyv = 0;
            int k = 8; /* :594 counts down over VISITED taps only */
            for(int j = imax(y - 1, 0); j <= imin(y + 1, rows - 1); j++)
                for(int i = imax(x - 1, 0); i <= imin(x + 1, cols - 1); i++)
                {
                    float s = (float)src[(size_t)j * cols + i];
                    dxv = fmaf(s, gsx[k], dxv);
                    dyv = fmaf(s, gsy[k], dyv);
                    --k;
                }
            int a = float2int_rz(dxv), b = float2int_rz(dyv); /* :605-606 float -> short */
            a = a > 32767 ? 32767 : (a < -32768 ? -32768 : a);
            b = b > 32767 ? 32767 : (b < -32768 ? -32768 : b);
            dx[(size_t)y * cols + x] = (int16_t)a;
            dy[(size_t)y * cols + x] = (int16_t)b;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* cudafuncs.cu:641-674  projectPointsKernel / projectToPointCloud                           */
/* ------------------------------------------------------------------------------------------ */
void efo_project_point_cloud(const float * depth, int rows, int cols, float fx, float fy, float cx, float cy, float * cloud)
{
    const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy; /* :671 host division */
#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            size_t i = (size_t)y * cols + x;
            float z = depth[i];
            cloud[i * 3 + 0] = ((float)x - cx) * z * inv_fx; /* :656 */
            cloud[i * 3 + 1] = ((float)y - cy) * z * inv_fy;
            cloud[i * 3 + 2] = z;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* 29- and 11-float accumulators: Cuda/types.cuh:101-181                                     */
/* ------------------------------------------------------------------------------------------ */
static inline void products_se3(const float * row, float inl, double * acc)
{
    int k = 0;
    for(int i = 0; i < 7; i++)
        for(int j = i; j < 7; j++)
        {
            if(i == 6 && j == 6) break;
            acc[k++] += (double)(row[i] * row[j]); /* reduce.cu:350-381 float products */
        }
    acc[27] += (double)(row[6] * row[6]);          /* :383 residual */
    acc[28] += (double)inl;                        /* :384 inliers */
}

static inline void products_so3(const float * row, float inl, double * acc)
{
    int k = 0;
    for(int i = 0; i < 4; i++)
        for(int j = i; j < 4; j++)
        {
            if(i == 3 && j == 3) break;
            acc[k++] += (double)(row[i] * row[j]); /* reduce.cu:1039-1049 */
        }
    acc[9] += (double)(row[3] * row[3]);
    acc[10] += (double)inl;
}

/* deterministic row-partial reduction: per-row double partials summed in row order */
static void reduce_rows(const double * partial, int rows, int n, float * out)
{
    for(int k = 0; k < n; k++)
    {
        double s = 0;
        for(int r = 0; r < rows; r++) s += partial[(size_t)r * n + k];
        out[k] = (float)s;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* reduce.cu:257-490  ICPReduction / icpKernel / icpStep                                     */
/* ------------------------------------------------------------------------------------------ */
void efo_icp_step(const float * Rcurr, const float * tcurr, const float * vmap_curr, const float * nmap_curr,
                  const float * Rprev_inv, const float * tprev, float fx, float fy, float cx, float cy,
                  const float * vmap_g_prev, const float * nmap_g_prev, float dist_thresh, float angle_thresh, int rows,
                  int cols, float * out29)
{
    const size_t plane = (size_t)rows * cols;
    const f3 tc = f3_make(tcurr[0], tcurr[1], tcurr[2]);
    const f3 tp = f3_make(tprev[0], tprev[1], tprev[2]);
    double * partial = (double *)calloc((size_t)rows * 29, sizeof(double));

#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
    {
        double * acc = partial + (size_t)y * 29;
        for(int x = 0; x < cols; x++)
        {
            const size_t i = (size_t)y * cols + x;
            float row[7] = {0, 0, 0, 0, 0, 0, 0};
            int found = 0;

            /* search(): :282-325 */
            f3 vcurr = f3_make(vmap_curr[i], vmap_curr[plane + i], vmap_curr[2 * plane + i]);
            f3 vcurr_g = f3_add(m33_mul(Rcurr, vcurr), tc);
            f3 vcurr_cp = m33_mul(Rprev_inv, f3_sub(vcurr_g, tp));
            int ux = float2int_rn(vcurr_cp.x * fx / vcurr_cp.z + cx); /* :294 */
            int uy = float2int_rn(vcurr_cp.y * fy / vcurr_cp.z + cy);

            if(!(ux < 0 || uy < 0 || ux >= cols || uy >= rows || vcurr_cp.z < 0)) /* :297 */
            {
                const size_t j = (size_t)uy * cols + ux;
                f3 vprev_g = f3_make(vmap_g_prev[j], vmap_g_prev[plane + j], vmap_g_prev[2 * plane + j]);
                f3 ncurr = f3_make(nmap_curr[i], nmap_curr[plane + i], nmap_curr[2 * plane + i]);
                f3 ncurr_g = m33_mul(Rcurr, ncurr);
                f3 nprev_g = f3_make(nmap_g_prev[j], nmap_g_prev[plane + j], nmap_g_prev[2 * plane + j]);

                float dist = f3_norm(f3_sub(vprev_g, vcurr_g));      /* :317 */
                float sine = f3_norm(f3_cross(ncurr_g, nprev_g));    /* :318 */

                found = (sine < angle_thresh && dist <= dist_thresh && !isnan(ncurr.x) && !isnan(nprev_g.x)); /* :324 */
                if(found)
                {
                    /* getProducts(): :341-347 */
                    f3 s_cp = m33_mul(Rprev_inv, f3_sub(vcurr_g, tp));
                    f3 d_cp = m33_mul(Rprev_inv, f3_sub(vprev_g, tp));
                    f3 n_cp = m33_mul(Rprev_inv, nprev_g);
                    f3 c = f3_cross(s_cp, n_cp);
                    row[0] = n_cp.x; row[1] = n_cp.y; row[2] = n_cp.z;
                    row[3] = c.x; row[4] = c.y; row[5] = c.z;
                    row[6] = f3_dot(n_cp, f3_sub(s_cp, d_cp));
                }
            }
            if(found) products_se3(row, 1.0f, acc);
        }
    }
    reduce_rows(partial, rows, 29, out29);
    free(partial);
}

/* ------------------------------------------------------------------------------------------ */
/* reduce.cu:739-936  RGBResidual / residualKernel / computeRgbResidual                      */
/* ------------------------------------------------------------------------------------------ */
void efo_rgb_residual(float min_scale, const int16_t * dIdx, const int16_t * dIdy, const float * last_depth,
                      const float * next_depth, const uint8_t * last_image, const uint8_t * next_image,
                      efo_data_term * corres, float max_depth_delta, const float * kt, const float * K, int rows, int cols,
                      int * sigma_sum, int * count)
{
    const int border = 16; /* :779 */
    uint32_t total_count = 0, total_sigma = 0; /* int2 sums wrap like the GPU's int adds */

#pragma omp parallel for schedule(static) reduction(+ : total_count, total_sigma)
    for(int i = 0; i < rows; i++)
        for(int j0 = 0; j0 < cols; j0++)
        {
            efo_data_term c;
            memset(&c, 0, sizeof(c));
            const size_t k = (size_t)i * cols + j0;

            if(i >= border && i < rows - border && j0 >= border && j0 < cols - border && j0 < cols - 5 && i < rows - 1)
            {
                int valid = 1;
                for(int u = imax(i - 2, 0); u < imin(i + 2, rows); u++)      /* :787-793 */
                    for(int v = imax(j0 - 2, 0); v < imin(j0 + 2, cols); v++)
                        valid = valid && (next_image[(size_t)u * cols + v] > 0);

                if(valid)
                {
                    int valx = dIdx[k], valy = dIdy[k];
                    float mTwo = (float)(valx * valx + valy * valy);        /* :802 */
                    if(mTwo >= min_scale)
                    {
                        const int y = i, x = j0;
                        float d1 = next_depth[k];
                        if(!isnan(d1))
                        {
                            const float xf = (float)x, yf = (float)y;
                            /* :813-815 */
                            float td1 = fmaf(d1, fmaf(K[6], xf, K[7] * yf) + K[8], kt[2]);
                            int u0 = float2int_rn(fmaf(d1, fmaf(K[0], xf, K[1] * yf) + K[2], kt[0]) / td1);
                            int v0 = float2int_rn(fmaf(d1, fmaf(K[3], xf, K[4] * yf) + K[5], kt[1]) / td1);
                            if(u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows)
                            {
                                const size_t k0 = (size_t)v0 * cols + u0;
                                float d0 = last_depth[k0];
                                if(d0 > 0 && fabsf(td1 - d0) <= max_depth_delta && last_image[k0] != 0) /* :821 */
                                {
                                    c.zero_x = (int16_t)u0; c.zero_y = (int16_t)v0;
                                    c.one_x = (int16_t)x; c.one_y = (int16_t)y;
                                    c.diff = (float)next_image[k] - (float)last_image[k0];
                                    c.valid = 1;
                                    total_count += 1u;
                                    total_sigma += (uint32_t)float2int_rz(c.diff * c.diff); /* :830 */
                                }
                            }
                        }
                    }
                }
            }
            corres[k] = c; /* :839 written for every pixel */
        }

    *count = (int)total_count;
    *sigma_sum = (int)total_sigma;
}

/* ------------------------------------------------------------------------------------------ */
/* reduce.cu:494-678  RGBReduction / rgbKernel / rgbStep                                     */
/* ------------------------------------------------------------------------------------------ */
void efo_rgb_step(const efo_data_term * corres, float sigma, const float * cloud, float fx, float fy, const int16_t * dIdx,
                  const int16_t * dIdy, float sobel_scale, int rows, int cols, float * out29)
{
    double * partial = (double *)calloc((size_t)rows * 29, sizeof(double));

#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
    {
        double * acc = partial + (size_t)y * 29;
        for(int x = 0; x < cols; x++)
        {
            const efo_data_term * c = &corres[(size_t)y * cols + x]; /* :515 linear index */
            if(!c->valid) continue;

            float row[7];
            float w = sigma + fabsf(c->diff);                 /* :523 */
            w = w > FLT_EPSILON ? 1.0f / w : 1.0f;            /* :525 */
            if(sigma == -1) w = 1;                            /* :528 */
            row[6] = -w * c->diff;

            const float * cp = cloud + ((size_t)c->zero_y * cols + c->zero_x) * 3;
            const float X = cp[0], Y = cp[1], Z = cp[2];
            float invz = (float)(1.0 / (double)Z);            /* :539 double reciprocal */
            float dI_dx = w * sobel_scale * (float)dIdx[(size_t)c->one_y * cols + c->one_x];
            float dI_dy = w * sobel_scale * (float)dIdy[(size_t)c->one_y * cols + c->one_x];
            float v0 = dI_dx * fx * invz;
            float v1 = dI_dy * fy * invz;
            float v2 = -(fmaf(v0, X, v1 * Y)) * invz;         /* :544 */
            row[0] = v0; row[1] = v1; row[2] = v2;
            row[3] = fmaf(-Z, v1, Y * v2);                    /* :549-551 */
            row[4] = fmaf(Z, v0, -(X * v2));
            row[5] = fmaf(-Y, v0, X * v1);
            products_se3(row, 1.0f, acc);
        }
    }
    reduce_rows(partial, rows, 29, out29);
    free(partial);
}

/* ------------------------------------------------------------------------------------------ */
/* reduce.cu:938-1141  SO3Reduction / so3Kernel / so3Step                                    */
/* ------------------------------------------------------------------------------------------ */
static inline void so3_gradient(const uint8_t * img, int cols, int x, int y, float * gx, float * gy)
{
    /* :955-969 */
    float actu = (float)img[(size_t)y * cols + x];
    float back = (float)img[(size_t)y * cols + x - 1];
    float fore = (float)img[(size_t)y * cols + x + 1];
    *gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    back = (float)img[(size_t)(y - 1) * cols + x];
    fore = (float)img[(size_t)(y + 1) * cols + x];
    *gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

void efo_so3_step(const uint8_t * last_image, const uint8_t * next_image, const float * H, const float * kinv,
                  const float * krlr, int rows, int cols, float * out11)
{
    double * partial = (double *)calloc((size_t)rows * 11, sizeof(double));

#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
    {
        double * acc = partial + (size_t)y * 11;
        for(int x = 0; x < cols; x++)
        {
            f3 p = f3_make((float)x, (float)y, 1.0f);              /* :980 */
            f3 wp = m33_mul(H, p);                                 /* :982 */
            int wx = float2int_rn(wp.x / wp.z);                    /* :984-985 */
            int wy = float2int_rn(wp.y / wp.z);

            if(!(wx >= 1 && wx < cols - 1 && wy >= 1 && wy < rows - 1 && x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1))
                continue;                                          /* :987-997 */

            float gnx, gny, glx, gly;
            so3_gradient(next_image, cols, wx, wy, &gnx, &gny);
            so3_gradient(last_image, cols, x, y, &glx, &gly);
            float gx = (gnx + glx) / 2.0f, gy = (gny + gly) / 2.0f; /* :1007-1008 */

            f3 point = m33_mul(kinv, p);
            float z2 = point.z * point.z;
            float a = krlr[0], b = krlr[1], c = krlr[2], d = krlr[3], e = krlr[4], f = krlr[5], g = krlr[6], h = krlr[7],
                  ii = krlr[8];
            const float xf = (float)x, yf = (float)y;
            /* :1027-1029 */
            f3 lp = f3_make(((point.z * (d * gy + a * gx)) - (gy * g * yf) - (gx * g * xf)) / z2,
                            ((point.z * (e * gy + b * gx)) - (gy * h * yf) - (gx * h * xf)) / z2,
                            ((point.z * (f * gy + c * gx)) - (gy * ii * yf) - (gx * ii * xf)) / z2);
            f3 jr = f3_cross(lp, point);
            float row[4];
            row[0] = jr.x; row[1] = jr.y; row[2] = jr.z;
            row[3] = -((float)next_image[(size_t)wy * cols + wx] - (float)last_image[(size_t)y * cols + x]); /* :1036 */
            products_so3(row, 1.0f, acc);
        }
    }
    reduce_rows(partial, rows, 11, out11);
    free(partial);
}

/* reduce.cu:475-489 */
void efo_unpack_se3(const float * h, float * A, float * b, float * residual)
{
    int shift = 0;
    for(int i = 0; i < 6; ++i)
        for(int j = i; j < 7; ++j)
        {
            float value = h[shift++];
            if(j == 6) b[i] = value;
            else A[j * 6 + i] = A[i * 6 + j] = value;
        }
    residual[0] = h[27];
    residual[1] = h[28];
}

/* reduce.cu:1126-1140 */
void efo_unpack_so3(const float * h, float * A, float * b, float * residual)
{
    int shift = 0;
    for(int i = 0; i < 3; ++i)
        for(int j = i; j < 4; ++j)
        {
            float value = h[shift++];
            if(j == 3) b[i] = value;
            else A[j * 3 + i] = A[i * 3 + j] = value;
        }
    residual[0] = h[9];
    residual[1] = h[10];
}

/* ------------------------------------------------------------------------------------------ */
/* Host math standing in for the Eigen calls of RGBDOdometry.cpp (Eigen itself is an         */
/* un-vendored, unpinned dependency: README.md:15).                                          */
/* ------------------------------------------------------------------------------------------ */

/* OdometryProvider.h:35-71 */
void efo_rodrigues(const double * src, double * R)
{
    for(int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    double rx = src[0], ry = src[1], rz = src[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if(theta >= DBL_EPSILON)
    {
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        double c = cos(theta), s = sin(theta), c1 = 1. - c;
        double itheta = theta ? 1. / theta : 0.;
        rx *= itheta; ry *= itheta; rz *= itheta;
        double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
        double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        for(int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
    }
}

/* Eigen's LDLT (A.ldlt().solve(b)): symmetric-pivoted LDL^T.  Returns 0 on success. */
int efo_ldlt_solve_f64(const double * A_in, const double * b, int n, double * x)
{
    double A[36], y[6];
    int perm[6];
    if(n > 6) return -1;
    for(int i = 0; i < n * n; i++) A[i] = A_in[i];
    for(int i = 0; i < n; i++) perm[i] = i;

    for(int k = 0; k < n; k++)
    {
        /* pivot: largest |diagonal| of the trailing block */
        int p = k;
        double best = fabs(A[k * n + k]);
        for(int i = k + 1; i < n; i++)
            if(fabs(A[i * n + i]) > best) { best = fabs(A[i * n + i]); p = i; }
        if(p != k)
        {
            for(int j = 0; j < n; j++) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for(int i = 0; i < n; i++) { double t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        double d = A[k * n + k];
        if(d == 0.0) continue;
        for(int i = k + 1; i < n; i++) A[i * n + k] /= d;
        for(int i = k + 1; i < n; i++)
            for(int j = k + 1; j <= i; j++)
            {
                A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
                A[j * n + i] = A[i * n + j];
            }
    }
    /* solve P^T L D L^T P x = b */
    for(int i = 0; i < n; i++) y[i] = b[perm[i]];
    for(int i = 0; i < n; i++)
        for(int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
    for(int i = 0; i < n; i++) y[i] = (A[i * n + i] != 0.0) ? y[i] / A[i * n + i] : 0.0;
    for(int i = n - 1; i >= 0; i--)
        for(int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
    for(int i = 0; i < n; i++) x[perm[i]] = y[i];
    return 0;
}

/* float 3x3 LDLT for the SO(3) step (RGBDOdometry.cpp:368) */
void efo_ldlt_solve3_f32(const float * A_in, const float * b, float * x)
{
    float A[9], y[3];
    int perm[3] = {0, 1, 2};
    const int n = 3;
    for(int i = 0; i < 9; i++) A[i] = A_in[i];
    for(int k = 0; k < n; k++)
    {
        int p = k;
        float best = fabsf(A[k * n + k]);
        for(int i = k + 1; i < n; i++)
            if(fabsf(A[i * n + i]) > best) { best = fabsf(A[i * n + i]); p = i; }
        if(p != k)
        {
            for(int j = 0; j < n; j++) { float t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for(int i = 0; i < n; i++) { float t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        float d = A[k * n + k];
        if(d == 0.0f) continue;
        for(int i = k + 1; i < n; i++) A[i * n + k] /= d;
        for(int i = k + 1; i < n; i++)
            for(int j = k + 1; j <= i; j++)
            {
                A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
                A[j * n + i] = A[i * n + j];
            }
    }
    for(int i = 0; i < n; i++) y[i] = b[perm[i]];
    for(int i = 0; i < n; i++)
        for(int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
    for(int i = 0; i < n; i++) y[i] = (A[i * n + i] != 0.0f) ? y[i] / A[i * n + i] : 0.0f;
    for(int i = n - 1; i >= 0; i--)
        for(int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
    for(int i = 0; i < n; i++) x[perm[i]] = y[i];
}

void efo_mul33_f64(const double * A, const double * B, double * C)
{
    double T[9];
    for(int i = 0; i < 3; i++)
        for(int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    memcpy(C, T, sizeof(T));
}

/* cofactor inverse (what Eigen uses for fixed 3x3) */
void efo_inverse3_f64(const double * m, double * o)
{
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

void efo_inverse3_f32(const float * m, float * o)
{
    float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    float id = 1.0f / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

/* general 4x4 inverse by Gauss-Jordan with partial pivoting */
void efo_inverse4_f64(const double * M, double * Minv)
{
    double a[4][8];
    for(int i = 0; i < 4; i++)
        for(int j = 0; j < 4; j++)
        {
            a[i][j] = M[i * 4 + j];
            a[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for(int c = 0; c < 4; c++)
    {
        int p = c;
        for(int r = c + 1; r < 4; r++)
            if(fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if(p != c)
            for(int j = 0; j < 8; j++) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        double inv = 1.0 / a[c][c];
        for(int j = 0; j < 8; j++) a[c][j] *= inv;
        for(int r = 0; r < 4; r++)
            if(r != c)
            {
                double f = a[r][c];
                if(f != 0.0)
                    for(int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
            }
    }
    for(int i = 0; i < 4; i++)
        for(int j = 0; j < 4; j++) Minv[i * 4 + j] = a[i][4 + j];
}

/* ------------------------------------------------------------------------------------------ */
/* Tracker object: RGBDOdometry.h:31-134, RGBDOdometry.cpp:21-111                             */
/* ------------------------------------------------------------------------------------------ */
#define NUM_PYRS 3

struct efo_tracker
{
    int width, height;
    float cx, cy, fx, fy;
    float dist_thresh, angle_thresh;
    float sobel_scale, max_depth_delta_rgb, max_depth_rgb;
    float min_grad[NUM_PYRS];

    uint16_t * depth_tmp[NUM_PYRS];
    float * vmaps_tmp;  /* rgba32f staging, RGBDOdometry.h:82 */
    float * nmaps_tmp;
    float * vmaps_g_prev[NUM_PYRS], * nmaps_g_prev[NUM_PYRS];
    float * vmaps_curr[NUM_PYRS], * nmaps_curr[NUM_PYRS];
    float * last_depth[NUM_PYRS], * next_depth[NUM_PYRS];
    uint8_t * last_image[NUM_PYRS], * next_image[NUM_PYRS], * last_next_image[NUM_PYRS];
    int16_t * dIdx[NUM_PYRS], * dIdy[NUM_PYRS];
    efo_data_term * corres[NUM_PYRS];
    float * cloud[NUM_PYRS];

    efo_stats st;
};

static void level_intr(const efo_tracker * t, int level, float * fx, float * fy, float * cx, float * cy)
{
    int div = 1 << level; /* types.cuh:94-98 */
    *fx = t->fx / div; *fy = t->fy / div; *cx = t->cx / div; *cy = t->cy / div;
}

efo_tracker * efo_tracker_create(int width, int height, float cx, float cy, float fx, float fy, float dist_thresh,
                                 float angle_thresh)
{
    efo_tracker * t = (efo_tracker *)calloc(1, sizeof(efo_tracker));
    t->width = width; t->height = height;
    t->cx = cx; t->cy = cy; t->fx = fx; t->fy = fy;
    t->dist_thresh = dist_thresh; t->angle_thresh = angle_thresh;
    t->sobel_scale = (float)(1.0 / pow(2.0, 3));   /* :34-35 */
    t->max_depth_delta_rgb = 0.07f;                /* :36 */
    t->max_depth_rgb = 6.0f;                       /* :37 */
    t->min_grad[0] = 5; t->min_grad[1] = 3; t->min_grad[2] = 1; /* :107-110 */
    t->st.last_icp_count = t->st.last_rgb_count = t->st.last_so3_count = (float)(width * height); /* :26-31 */

    t->vmaps_tmp = (float *)calloc((size_t)width * height * 4, sizeof(float));
    t->nmaps_tmp = (float *)calloc((size_t)width * height * 4, sizeof(float));
    for(int i = 0; i < NUM_PYRS; i++)
    {
        size_t n = (size_t)(width >> i) * (height >> i);
        t->depth_tmp[i] = (uint16_t *)calloc(n, sizeof(uint16_t));
        t->vmaps_g_prev[i] = (float *)calloc(n * 3, sizeof(float));
        t->nmaps_g_prev[i] = (float *)calloc(n * 3, sizeof(float));
        t->vmaps_curr[i] = (float *)calloc(n * 3, sizeof(float));
        t->nmaps_curr[i] = (float *)calloc(n * 3, sizeof(float));
        t->last_depth[i] = (float *)calloc(n, sizeof(float));
        t->next_depth[i] = (float *)calloc(n, sizeof(float));
        t->last_image[i] = (uint8_t *)calloc(n, 1);
        t->next_image[i] = (uint8_t *)calloc(n, 1);
        t->last_next_image[i] = (uint8_t *)calloc(n, 1);
        t->dIdx[i] = (int16_t *)calloc(n, sizeof(int16_t));
        t->dIdy[i] = (int16_t *)calloc(n, sizeof(int16_t));
        t->corres[i] = (efo_data_term *)calloc(n, sizeof(efo_data_term));
        t->cloud[i] = (float *)calloc(n * 3, sizeof(float));
    }
    return t;
}

void efo_tracker_destroy(efo_tracker * t)
{
    if(!t) return;
    free(t->vmaps_tmp); free(t->nmaps_tmp);
    for(int i = 0; i < NUM_PYRS; i++)
    {
        free(t->depth_tmp[i]); free(t->vmaps_g_prev[i]); free(t->nmaps_g_prev[i]); free(t->vmaps_curr[i]);
        free(t->nmaps_curr[i]); free(t->last_depth[i]); free(t->next_depth[i]); free(t->last_image[i]);
        free(t->next_image[i]); free(t->last_next_image[i]); free(t->dIdx[i]); free(t->dIdy[i]); free(t->corres[i]);
        free(t->cloud[i]);
    }
    free(t);
}

const void * efo_tracker_buffer(const efo_tracker * t, const char * name, int level)
{
    if(level < 0 || level >= NUM_PYRS) return NULL;
    if(!strcmp(name, "vmap_curr")) return t->vmaps_curr[level];
    if(!strcmp(name, "nmap_curr")) return t->nmaps_curr[level];
    if(!strcmp(name, "vmap_g_prev")) return t->vmaps_g_prev[level];
    if(!strcmp(name, "nmap_g_prev")) return t->nmaps_g_prev[level];
    if(!strcmp(name, "last_depth")) return t->last_depth[level];
    if(!strcmp(name, "next_depth")) return t->next_depth[level];
    if(!strcmp(name, "last_image")) return t->last_image[level];
    if(!strcmp(name, "next_image")) return t->next_image[level];
    if(!strcmp(name, "last_next_image")) return t->last_next_image[level];
    if(!strcmp(name, "dIdx")) return t->dIdx[level];
    if(!strcmp(name, "dIdy")) return t->dIdy[level];
    if(!strcmp(name, "depth_tmp")) return t->depth_tmp[level];
    return NULL;
}

/* RGBDOdometry.cpp:118-142 */
void efo_init_icp_depth(efo_tracker * t, const uint16_t * depth, float cutoff)
{
    memcpy(t->depth_tmp[0], depth, (size_t)t->width * t->height * sizeof(uint16_t));
    for(int i = 1; i < NUM_PYRS; ++i)
        efo_pyr_down_u16(t->depth_tmp[i - 1], t->height >> (i - 1), t->width >> (i - 1), t->depth_tmp[i]);
    for(int i = 0; i < NUM_PYRS; ++i)
    {
        float fx, fy, cx, cy;
        level_intr(t, i, &fx, &fy, &cx, &cy);
        efo_create_vmap(t->depth_tmp[i], t->height >> i, t->width >> i, fx, fy, cx, cy, cutoff, t->vmaps_curr[i]);
        efo_create_nmap(t->vmaps_curr[i], t->height >> i, t->width >> i, t->nmaps_curr[i]);
    }
}

/* RGBDOdometry.cpp:144-167 */
void efo_init_icp_maps(efo_tracker * t, const float * v4, const float * n4, float cutoff)
{
    (void)cutoff; /* unused by the reference too */
    size_t bytes = (size_t)t->width * t->height * 4 * sizeof(float);
    memcpy(t->vmaps_tmp, v4, bytes);
    memcpy(t->nmaps_tmp, n4, bytes);
    efo_copy_maps(t->vmaps_tmp, t->nmaps_tmp, t->height, t->width, t->vmaps_curr[0], t->nmaps_curr[0]);
    for(int i = 1; i < NUM_PYRS; ++i)
    {
        efo_resize_map(t->vmaps_curr[i - 1], t->height >> (i - 1), t->width >> (i - 1), t->vmaps_curr[i], 0);
        efo_resize_map(t->nmaps_curr[i - 1], t->height >> (i - 1), t->width >> (i - 1), t->nmaps_curr[i], 1);
    }
}

/* RGBDOdometry.cpp:169-206 ; pose16 row-major 4x4 */
void efo_init_icp_model(efo_tracker * t, const float * v4, const float * n4, float cutoff, const float * pose)
{
    (void)cutoff;
    size_t bytes = (size_t)t->width * t->height * 4 * sizeof(float);
    memcpy(t->vmaps_tmp, v4, bytes);
    memcpy(t->nmaps_tmp, n4, bytes);
    efo_copy_maps(t->vmaps_tmp, t->nmaps_tmp, t->height, t->width, t->vmaps_g_prev[0], t->nmaps_g_prev[0]);
    for(int i = 1; i < NUM_PYRS; ++i)
    {
        efo_resize_map(t->vmaps_g_prev[i - 1], t->height >> (i - 1), t->width >> (i - 1), t->vmaps_g_prev[i], 0);
        efo_resize_map(t->nmaps_g_prev[i - 1], t->height >> (i - 1), t->width >> (i - 1), t->nmaps_g_prev[i], 1);
    }
    float R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
    float tv[3] = {pose[3], pose[7], pose[11]};
    for(int i = 0; i < NUM_PYRS; ++i)
        efo_transform_maps(t->vmaps_g_prev[i], t->nmaps_g_prev[i], t->height >> i, t->width >> i, R, tv,
                           t->vmaps_g_prev[i], t->nmaps_g_prev[i]);
}

/* RGBDOdometry.cpp:208-235 */
static void populate_rgbd(efo_tracker * t, const uint8_t * rgba, float ** depths, uint8_t ** images)
{
    efo_vertices_to_depth(t->vmaps_tmp, t->height, t->width, t->max_depth_rgb, depths[0]);
    for(int i = 0; i + 1 < NUM_PYRS; i++)
        efo_pyr_down_gauss_f32(depths[i], t->height >> i, t->width >> i, depths[i + 1]);
    efo_bgr_to_intensity(rgba, t->height, t->width, images[0]);
    for(int i = 0; i + 1 < NUM_PYRS; i++)
        efo_pyr_down_gauss_u8(images[i], t->height >> i, t->width >> i, images[i + 1]);
}

void efo_init_rgb_model(efo_tracker * t, const uint8_t * rgba) { populate_rgbd(t, rgba, t->last_depth, t->last_image); }
void efo_init_rgb(efo_tracker * t, const uint8_t * rgba) { populate_rgbd(t, rgba, t->next_depth, t->next_image); }

/* RGBDOdometry.cpp:249-265 */
void efo_init_first_rgb(efo_tracker * t, const uint8_t * rgba)
{
    efo_bgr_to_intensity(rgba, t->height, t->width, t->last_next_image[0]);
    for(int i = 0; i + 1 < NUM_PYRS; i++)
        efo_pyr_down_gauss_u8(t->last_next_image[i], t->height >> i, t->width >> i, t->last_next_image[i + 1]);
}

static void cast9(const double * d, float * f) { for(int i = 0; i < 9; i++) f[i] = (float)d[i]; }

/* RGBDOdometry.cpp:267-603 */
void efo_get_incremental_transformation(efo_tracker * t, float * trans, float * rot, int rgb_only, float icp_weight,
                                        int pyramid, int fast_odom, int so3, efo_stats * stats)
{
    const int icp = !rgb_only && icp_weight > 0;   /* :275 */
    const int rgb = rgb_only || icp_weight < 100;  /* :276 */

    float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
    memcpy(Rprev, rot, sizeof(Rprev)); memcpy(tprev, trans, sizeof(tprev));
    memcpy(Rcurr, rot, sizeof(Rcurr)); memcpy(tcurr, trans, sizeof(tcurr));

    if(rgb) /* :284-290 */
        for(int i = 0; i < NUM_PYRS; i++)
            efo_derivative_images(t->next_image[i], t->height >> i, t->width >> i, t->dIdx[i], t->dIdy[i]);

    double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    t->st.so3_iterations = 0;
    t->st.se3_iterations[0] = t->st.se3_iterations[1] = t->st.se3_iterations[2] = 0;

    if(so3) /* :294-382 */
    {
        const int lvl = 2;
        const int rows = t->height >> lvl, cols = t->width >> lvl;
        float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        float fx, fy, cx, cy;
        level_intr(t, lvl, &fx, &fy, &cx, &cy);
        double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1}, K_inv[9];
        efo_inverse3_f64(K, K_inv);

        float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
        double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

        for(int i = 0; i < 10; i++)
        {
            double tmp[9], Hd[9], KRd[9];
            efo_mul33_f64(K, resultR, tmp);
            efo_mul33_f64(tmp, K_inv, Hd);              /* :318 homography = K R K^-1 */
            memcpy(KRd, tmp, sizeof(tmp));          /* :327 K * resultR */
            float H[9], kinv[9], krlr[9];
            cast9(Hd, H); cast9(K_inv, kinv); cast9(KRd, krlr);

            float out11[11], jtj[9], jtr[3], residual[2];
            efo_so3_step(t->last_next_image[lvl], t->next_image[lvl], H, kinv, krlr, rows, cols, out11);
            efo_unpack_so3(out11, jtj, jtr, residual);
            t->st.so3_iterations++;

            t->st.last_so3_error = sqrtf(residual[0]) / residual[1]; /* :348 */
            t->st.last_so3_count = residual[1];

            if(t->st.last_so3_error < lastError && lastCount == t->st.last_so3_count) break; /* :352 */
            else if(t->st.last_so3_error > lastError + 0.001)                                 /* :356 */
            {
                t->st.last_so3_error = lastError;
                t->st.last_so3_count = lastCount;
                memcpy(resultR, lastResultR, sizeof(resultR));
                break;
            }
            lastError = t->st.last_so3_error;
            lastCount = t->st.last_so3_count;
            memcpy(lastResultR, resultR, sizeof(resultR));

            float delta[3];
            efo_ldlt_solve3_f32(jtj, jtr, delta);       /* :368 */
            double dd[3] = {delta[0], delta[1], delta[2]}, rotUpdate[9];
            efo_rodrigues(dd, rotUpdate);
            float ru[9], nr[9];
            cast9(rotUpdate, ru);
            for(int r = 0; r < 3; r++)              /* :372 R_lr = rotUpdate.cast<float>() * R_lr */
                for(int c = 0; c < 3; c++)
                    nr[r * 3 + c] = ru[r * 3] * R_lr[c] + ru[r * 3 + 1] * R_lr[3 + c] + ru[r * 3 + 2] * R_lr[6 + c];
            memcpy(R_lr, nr, sizeof(nr));
            for(int k = 0; k < 9; k++) resultR[k] = R_lr[k];
        }
    }

    int iterations[NUM_PYRS] = {fast_odom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0}; /* :384-386 */

    float Rprev_inv[9];
    efo_inverse3_f32(Rprev, Rprev_inv); /* :388 */

    double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if(so3)
        for(int x = 0; x < 3; x++)
            for(int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

    for(int i = NUM_PYRS - 1; i >= 0; i--) /* :405 */
    {
        const int rows = t->height >> i, cols = t->width >> i;
        float fx, fy, cx, cy;
        level_intr(t, i, &fx, &fy, &cx, &cy);

        if(rgb) efo_project_point_cloud(t->last_depth[i], rows, cols, fx, fy, cx, cy, t->cloud[i]); /* :409 */

        double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1}, K_inv[9];
        efo_inverse3_f64(K, K_inv);

        t->st.last_rgb_error = FLT_MAX; /* :420 */

        for(int j = 0; j < iterations[i]; j++)
        {
            double Rt[16];
            efo_inverse4_f64(resultRt, Rt);                              /* :424 */
            double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
            double tmp[9], KRK_inv[9];
            efo_mul33_f64(K, R, tmp);
            efo_mul33_f64(tmp, K_inv, KRK_inv);                              /* :428 */
            float krkInv[9];
            cast9(KRK_inv, krkInv);
            double tv[3] = {Rt[3], Rt[7], Rt[11]};
            float kt[3];
            for(int r = 0; r < 3; r++) kt[r] = (float)(K[r * 3] * tv[0] + K[r * 3 + 1] * tv[1] + K[r * 3 + 2] * tv[2]);

            int sigma = 0, rgbSize = 0;
            if(rgb)
            {
                float minScale = (float)(pow(t->min_grad[i], 2.0) / pow(t->sobel_scale, 2.0)); /* :442 */
                efo_rgb_residual(minScale, t->dIdx[i], t->dIdy[i], t->last_depth[i], t->next_depth[i], t->last_image[i],
                                 t->next_image[i], t->corres[i], t->max_depth_delta_rgb, kt, krkInv, rows, cols, &sigma,
                                 &rgbSize);
            }

            /* :461 precedence quirk: ((float)sigma / rgbSize == 0) ? 1 : rgbSize */
            float sigmaVal = (float)sqrt((double)(((float)sigma / (float)rgbSize == 0) ? 1 : rgbSize));
            float rgbError = (float)(sqrt((double)sigma) / (rgbSize == 0 ? 1 : rgbSize)); /* :462 */

            if(rgb_only && rgbError > t->st.last_rgb_error) break; /* :464 */

            t->st.last_rgb_error = rgbError;
            t->st.last_rgb_count = (float)rgbSize;
            if(rgb_only) sigmaVal = -1; /* :472 */

            float A_icp[36] = {0}, b_icp[6] = {0}, A_rgb[36] = {0}, b_rgb[6] = {0}, residual[2];
            float out29[29];
            if(icp)
            {
                efo_icp_step(Rcurr, tcurr, t->vmaps_curr[i], t->nmaps_curr[i], Rprev_inv, tprev, fx, fy, cx, cy,
                             t->vmaps_g_prev[i], t->nmaps_g_prev[i], t->dist_thresh, t->angle_thresh, rows, cols, out29);
                efo_unpack_se3(out29, A_icp, b_icp, residual);
                /* :515-516 (the reference reads `residual` uninitialised when !icp; we leave the
                 * previous values in place in that case) */
                t->st.last_icp_error = sqrtf(residual[0]) / residual[1];
                t->st.last_icp_count = residual[1];
            }
            if(rgb)
            {
                efo_rgb_step(t->corres[i], sigmaVal, t->cloud[i], fx, fy, t->dIdx[i], t->dIdy[i], t->sobel_scale, rows,
                             cols, out29);
                float r2[2];
                efo_unpack_se3(out29, A_rgb, b_rgb, r2);
            }

            double * lastA = t->st.last_A, * lastb = t->st.last_b, result[6];
            if(icp && rgb) /* :547-553 */
            {
                double w = icp_weight;
                for(int k = 0; k < 36; k++) lastA[k] = (double)A_rgb[k] + w * w * (double)A_icp[k];
                for(int k = 0; k < 6; k++) lastb[k] = (double)b_rgb[k] + w * (double)b_icp[k];
            }
            else if(icp)
            {
                for(int k = 0; k < 36; k++) lastA[k] = A_icp[k];
                for(int k = 0; k < 6; k++) lastb[k] = b_icp[k];
            }
            else
            {
                for(int k = 0; k < 36; k++) lastA[k] = A_rgb[k];
                for(int k = 0; k < 6; k++) lastb[k] = b_rgb[k];
            }
            efo_ldlt_solve_f64(lastA, lastb, 6, result);
            t->st.se3_iterations[i]++;

            /* OdometryProvider.h:73-93 computeUpdateSE3 */
            double Rupd[9], Rtupd[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, newRt[16];
            efo_rodrigues(result + 3, Rupd);
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++) Rtupd[r * 4 + c] = Rupd[r * 3 + c];
                Rtupd[r * 4 + 3] = result[r];
            }
            for(int r = 0; r < 4; r++)
                for(int c = 0; c < 4; c++)
                {
                    double s = 0;
                    for(int k = 0; k < 4; k++) s += Rtupd[r * 4 + k] * resultRt[k * 4 + c];
                    newRt[r * 4 + c] = s;
                }
            memcpy(resultRt, newRt, sizeof(newRt));

            float oR[9], ot[3]; /* rgbOdom (Isometry3f) */
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++) oR[r * 3 + c] = (float)resultRt[r * 4 + c];
                ot[r] = (float)resultRt[r * 4 + 3];
            }
            /* :575-583 currentT = [Rprev|tprev] * rgbOdom.inverse()  (Isometry inverse = transpose) */
            float iR[9], it[3];
            for(int r = 0; r < 3; r++)
                for(int c = 0; c < 3; c++) iR[r * 3 + c] = oR[c * 3 + r];
            for(int r = 0; r < 3; r++) it[r] = -(iR[r * 3] * ot[0] + iR[r * 3 + 1] * ot[1] + iR[r * 3 + 2] * ot[2]);
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++)
                    Rcurr[r * 3 + c] = Rprev[r * 3] * iR[c] + Rprev[r * 3 + 1] * iR[3 + c] + Rprev[r * 3 + 2] * iR[6 + c];
                tcurr[r] = Rprev[r * 3] * it[0] + Rprev[r * 3 + 1] * it[1] + Rprev[r * 3 + 2] * it[2] + tprev[r];
            }
        }
    }

    if(rgb) /* :587-591 */
    {
        float d[3] = {tcurr[0] - tprev[0], tcurr[1] - tprev[1], tcurr[2] - tprev[2]};
        if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
        {
            memcpy(Rcurr, Rprev, sizeof(Rcurr));
            memcpy(tcurr, tprev, sizeof(tcurr));
        }
    }

    if(so3) /* :593-599 */
        for(int i = 0; i < NUM_PYRS; i++)
        {
            uint8_t * tmp = t->last_next_image[i];
            t->last_next_image[i] = t->next_image[i];
            t->next_image[i] = tmp;
        }

    memcpy(trans, tcurr, sizeof(tcurr));
    memcpy(rot, Rcurr, sizeof(Rcurr));
    if(stats) *stats = t->st;
}

/* RGBDOdometry.cpp:605-608  lastA.lu().inverse() */
void efo_get_covariance(const efo_tracker * t, double * cov)
{
    double a[6][12];
    for(int i = 0; i < 6; i++)
        for(int j = 0; j < 6; j++)
        {
            a[i][j] = t->st.last_A[i * 6 + j];
            a[i][6 + j] = (i == j) ? 1.0 : 0.0;
        }
    for(int c = 0; c < 6; c++)
    {
        int p = c;
        for(int r = c + 1; r < 6; r++)
            if(fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if(p != c)
            for(int j = 0; j < 12; j++) { double tt = a[c][j]; a[c][j] = a[p][j]; a[p][j] = tt; }
        double inv = 1.0 / a[c][c];
        for(int j = 0; j < 12; j++) a[c][j] *= inv;
        for(int r = 0; r < 6; r++)
            if(r != c)
            {
                double f = a[r][c];
                for(int j = 0; j < 12; j++) a[r][j] -= f * a[c][j];
            }
    }
    for(int i = 0; i < 6; i++)
        for(int j = 0; j < 6; j++) cov[i * 6 + j] = a[i][6 + j];
}

/* ------------------------------------------------------------------------------------------------
 * ElasticFusion::filterDepth -- Shaders/depth_bilateral.frag:30-76 (GLSL in the reference; there is no reference CUDA
 * for it).  Restated with exact texel fetches at (cx, cy) -- the shader samples at float(cx)/cols, the left EDGE of the
 * texel, where the nearest-texel choice of a GL implementation is not pinned -- IEEE float arithmetic in the shader's
 * order without contraction, libm expf for exp(), round-half-away for round().  Parity therefore is unpinned against
 * the reference's GL output; it is pinned against this restatement.
 * ------------------------------------------------------------------------------------------------ */
void efo_depth_bilateral(const uint16_t * src, int rows, int cols, float max_depth_m, uint16_t * dst)
{
    const unsigned max_mm = (unsigned)(max_depth_m * 1000.0f); /* :36 uint(maxD * 1000.0f) */
    const float sigma_space2_inv_half = 0.024691358f;          /* :42 */
    const float sigma_color2_inv_half = 0.000555556f;          /* :43 */
    const int R = 6, D = R * 2 + 1;                            /* :45-46 */
#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            const unsigned value = src[(size_t)y * cols + x];
            if(value > max_mm || value < 300u)
            {
                dst[(size_t)y * cols + x] = 0;
                continue;
            }
            const int tx = (x - D / 2 + D) < cols ? (x - D / 2 + D) : cols; /* :48 */
            const int ty = (y - D / 2 + D) < rows ? (y - D / 2 + D) : rows; /* :49 */
            float sum1 = 0.f, sum2 = 0.f;
            for(int cy = (y - D / 2 > 0 ? y - D / 2 : 0); cy < ty; ++cy)
                for(int cx = (x - D / 2 > 0 ? x - D / 2 : 0); cx < tx; ++cx)
                {
                    const unsigned tmp = src[(size_t)cy * cols + cx];
                    const float dx = (float)x - (float)cx, dy = (float)y - (float)cy;
                    const float space2 = dx * dx + dy * dy;                               /* :61 */
                    const float dc = (float)value - (float)tmp;
                    const float color2 = dc * dc;                                         /* :62 */
                    const float weight = expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half)); /* :64 */
                    sum1 += (float)tmp * weight;                                          /* :66 */
                    sum2 += weight;
                }
            dst[(size_t)y * cols + x] = (uint16_t)(unsigned)roundf(sum1 / sum2);          /* :71 */
        }
}

/* ElasticFusion::metriciseDepth -- Shaders/depth_metric.frag:28-40 */
void efo_depth_metric(const uint16_t * src, int rows, int cols, float max_depth_m, float * dst)
{
    const unsigned max_mm = (unsigned)(max_depth_m * 1000.0f);
    for(size_t i = 0; i < (size_t)rows * cols; i++)
    {
        const unsigned value = src[i];
        dst[i] = (value > max_mm || value < 300u) ? 0.f : (float)value / 1000.0f;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* the step around the tracker (OpenGL in the reference): model prediction + fill-in          */
/* IndexMap::combinedPredict IndexMap.cpp:468-575, Shaders/splat.vert:19-88,                  */
/* Shaders/combo_splat.frag:19-67; FillIn: Shaders/fill_vertex.frag, fill_normal.frag,        */
/* fill_rgb.frag, geometry.glsl:49-59.                                                         */
/* PARITY UNPINNED against a GL driver (none in this image): what GL leaves to the             */
/* implementation is fixed as follows and the CUDA kernels are compared with THIS statement:  */
/*  - a sprite of size s centred on window (xw, yw) covers the fragments whose centres        */
/*    (i + 0.5, j + 0.5) lie in [xw - s/2, xw + s/2) x [yw - s/2, yw + s/2); s is clamped to  */
/*    [1, 64]; a sprite whose centre is outside the viewport is clipped;                        */
/*  - depth buffer: 24-bit unsigned normalised, GL_LESS, surfels submitted in index order     */
/*    (so the lower index wins a tie);                                                        */
/*  - normalize(v) = v / sqrt(dot(v, v)), dot summed left to right, IEEE operations, no FMA   */
/*    (this file is built with -ffp-contract=off).                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } p3;
static inline float p3_dot(p3 a, p3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline p3 p3_normalize(p3 a)
{
    const float len = sqrtf(p3_dot(a, a));
    p3 r = {a.x / len, a.y / len, a.z / len};
    return r;
}
static inline p3 p3_cross(p3 a, p3 b)
{
    p3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}

typedef struct
{
    const float * t; /* rows 0..2 of t_inv */
    float cx, cy, fx, fy;
    int rows, cols;
    float max_depth, conf_threshold;
    int time, max_time, time_delta;
} splat_params;

typedef struct { p3 pos, nrm; float conf, rad, xw, yw, size; } surfel_view;

static inline void splat_project(const splat_params * P, float x, float y, float z, float * u, float * v)
{
    *u = (P->fx * x) / z + P->cx; /* splat.vert:40-45 */
    *v = (P->fy * y) / z + P->cy;
}

/* splat.vert:47-88 */
static int splat_vertex_stage(const splat_params * P, const float * s, surfel_view * S)
{
    const float * t = P->t;
    const float px = s[0], py = s[1], pz = s[2], conf = s[3], tstamp = s[7];
    S->pos.x = ((t[0] * px + t[1] * py) + t[2] * pz) + t[3];
    S->pos.y = ((t[4] * px + t[5] * py) + t[6] * pz) + t[7];
    S->pos.z = ((t[8] * px + t[9] * py) + t[10] * pz) + t[11];
    if(S->pos.z > P->max_depth || S->pos.z < 0 || conf < P->conf_threshold || (float)P->time - tstamp > (float)P->time_delta ||
       tstamp > (float)P->max_time)
        return 0;
    splat_project(P, S->pos.x, S->pos.y, S->pos.z, &S->xw, &S->yw);
    if(!(S->xw >= 0.f && S->xw <= (float)P->cols && S->yw >= 0.f && S->yw <= (float)P->rows)) return 0;
    S->conf = conf;
    S->rad = s[11];
    {
        const float nx = s[8], ny = s[9], nz = s[10];
        p3 rn = {(t[0] * nx + t[1] * ny) + t[2] * nz, (t[4] * nx + t[5] * ny) + t[6] * nz, (t[8] * nx + t[9] * ny) + t[10] * nz};
        S->nrm = p3_normalize(rn);
    }
    {
        p3 xa = {S->nrm.y - S->nrm.z, -S->nrm.x, S->nrm.x};
        const p3 xd = p3_normalize(xa);
        const float scale = S->rad * 1.41421356f;
        const p3 x1 = {xd.x * scale, xd.y * scale, xd.z * scale};
        const p3 y1 = p3_cross(S->nrm, x1);
        float u[4], v[4];
        splat_project(P, S->pos.x + x1.x, S->pos.y + x1.y, S->pos.z + x1.z, &u[0], &v[0]);
        splat_project(P, S->pos.x + y1.x, S->pos.y + y1.y, S->pos.z + y1.z, &u[1], &v[1]);
        splat_project(P, S->pos.x - y1.x, S->pos.y - y1.y, S->pos.z - y1.z, &u[2], &v[2]);
        splat_project(P, S->pos.x - x1.x, S->pos.y - x1.y, S->pos.z - x1.z, &u[3], &v[3]);
        const float xmin = fminf(u[0], fminf(u[1], fminf(u[2], u[3]))), xmax = fmaxf(u[0], fmaxf(u[1], fmaxf(u[2], u[3])));
        const float ymin = fminf(v[0], fminf(v[1], fminf(v[2], v[3]))), ymax = fmaxf(v[0], fmaxf(v[1], fmaxf(v[2], v[3])));
        const float size = fmaxf(0.f, fmaxf(fabsf(xmax - xmin), fabsf(ymax - ymin)));
        if(!(size == size)) return 0;
        S->size = fminf(fmaxf(size, 1.f), 64.f);
    }
    return 1;
}

/* combo_splat.frag:37-67 */
static int splat_fragment_stage(const splat_params * P, const surfel_view * S, int px, int py, float * z, unsigned * depth24)
{
    const float fxc = (float)px + 0.5f, fyc = (float)py + 0.5f;
    p3 la = {(fxc - P->cx) / P->fx, (fyc - P->cy) / P->fy, 1.f};
    const p3 l = p3_normalize(la);
    const float k = p3_dot(S->pos, S->nrm) / p3_dot(l, S->nrm);
    const p3 c = {k * l.x, k * l.y, k * l.z};
    const p3 d = {c.x - S->pos.x, c.y - S->pos.y, c.z - S->pos.z};
    if(!(p3_dot(d, d) <= S->rad * S->rad)) return 0;
    *z = c.z;
    {
        const float depth = c.z / (2.f * P->max_depth) + 0.5f;
        if(!(depth >= 0.f && depth <= 1.f)) return 0;
        *depth24 = (unsigned)lrintf(depth * 16777215.f);
    }
    return 1;
}

void efo_splat_predict(const float * surfels, int stride_floats, int count, const float * t_inv16, float cx, float cy, float fx, float fy, int rows,
                       int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta, uint8_t * image_rgba8,
                       float * vertex_rgba32f, float * normal_rgba32f, uint16_t * time_u16)
{
    const splat_params P = {t_inv16, cx, cy, fx, fy, rows, cols, max_depth, conf_threshold, time, max_time, time_delta};
    const size_t n = (size_t)rows * cols;
    unsigned * zbuf = (unsigned *)malloc(n * sizeof(unsigned));
    int * owner = (int *)malloc(n * sizeof(int));
    for(size_t i = 0; i < n; i++) { zbuf[i] = 0xffffffffu; owner[i] = -1; }
    /* the draw: surfels in submission order, GL_LESS */
    for(int i = 0; i < count; i++)
    {
        surfel_view S;
        if(!splat_vertex_stage(&P, surfels + (size_t)i * stride_floats, &S)) continue;
        const float h = S.size * 0.5f;
        int x0 = (int)ceilf((S.xw - h) - 0.5f), x1 = (int)ceilf((S.xw + h) - 0.5f);
        int y0 = (int)ceilf((S.yw - h) - 0.5f), y1 = (int)ceilf((S.yw + h) - 0.5f);
        if(x0 < 0) x0 = 0;
        if(y0 < 0) y0 = 0;
        if(x1 > cols) x1 = cols;
        if(y1 > rows) y1 = rows;
        for(int y = y0; y < y1; y++)
            for(int x = x0; x < x1; x++)
            {
                float z;
                unsigned d;
                if(splat_fragment_stage(&P, &S, x, y, &z, &d) && d < zbuf[(size_t)y * cols + x])
                {
                    zbuf[(size_t)y * cols + x] = d;
                    owner[(size_t)y * cols + x] = i;
                }
            }
    }
    /* the render targets, cleared to zero */
    for(size_t i = 0; i < n; i++)
    {
        float v[4] = {0, 0, 0, 0}, nn[4] = {0, 0, 0, 0};
        uint8_t img[4] = {0, 0, 0, 0};
        uint16_t tm = 0;
        if(owner[i] >= 0)
        {
            const float * s = surfels + (size_t)owner[i] * stride_floats;
            const int y = (int)(i / cols), x = (int)(i - (size_t)y * cols);
            surfel_view S;
            float z;
            unsigned d;
            splat_vertex_stage(&P, s, &S);
            splat_fragment_stage(&P, &S, x, y, &z, &d);
            {
                const float fxc = (float)x + 0.5f, fyc = (float)y + 0.5f;
                const int rgb = (int)s[4]; /* color.glsl:27-34 */
                img[0] = (rgb >> 16) & 0xff; img[1] = (rgb >> 8) & 0xff; img[2] = rgb & 0xff; img[3] = 255;
                v[0] = ((fxc - cx) * z) * (1.f / fx); v[1] = ((fyc - cy) * z) * (1.f / fy); v[2] = z; v[3] = S.conf; /* :61 */
                nn[0] = S.nrm.x; nn[1] = S.nrm.y; nn[2] = S.nrm.z; nn[3] = S.rad;
                tm = (uint16_t)(unsigned)s[6]; /* :65 */
            }
        }
        if(image_rgba8) memcpy(image_rgba8 + 4 * i, img, 4);
        memcpy(vertex_rgba32f + 4 * i, v, sizeof(v));
        memcpy(normal_rgba32f + 4 * i, nn, sizeof(nn));
        if(time_u16) time_u16[i] = tm;
    }
    free(zbuf);
    free(owner);
}

/* geometry.glsl:42-47 on the raw depth; texture fetches clamp to the edge */
static p3 fill_raw_vertex(const uint16_t * depth, int rows, int cols, float cx, float cy, float inv_fx, float inv_fy, int x, int y)
{
    const int xs = x < 0 ? 0 : (x > cols - 1 ? cols - 1 : x), ys = y < 0 ? 0 : (y > rows - 1 ? rows - 1 : y);
    const float z = (float)depth[(size_t)ys * cols + xs] / 1000.0f;
    p3 r = {(((float)x - cx) * z) * inv_fx, (((float)y - cy) * z) * inv_fy, z};
    return r;
}

/* Shaders/fill_vertex.frag:19-53 */
void efo_fill_vertex(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                     float * out)
{
    const float ifx = 1.f / fx, ify = 1.f / fy;
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            const size_t i = (size_t)y * cols + x;
            if(predicted[4 * i + 2] == 0 || passthrough == 1)
            {
                const p3 v = fill_raw_vertex(depth, rows, cols, cx, cy, ifx, ify, x, y);
                out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = 1.f;
            }
            else
                memcpy(out + 4 * i, predicted + 4 * i, 16);
        }
}

/* Shaders/fill_normal.frag:19-55, geometry.glsl:49-59 */
void efo_fill_normal(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                     float * out)
{
    const float ifx = 1.f / fx, ify = 1.f / fy;
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            const size_t i = (size_t)y * cols + x;
            if(predicted[4 * i + 2] == 0 || passthrough == 1)
            {
                const p3 v = fill_raw_vertex(depth, rows, cols, cx, cy, ifx, ify, x, y);
                const p3 vx = fill_raw_vertex(depth, rows, cols, cx, cy, ifx, ify, x + 1, y);
                const p3 vy = fill_raw_vertex(depth, rows, cols, cx, cy, ifx, ify, x, y + 1);
                const p3 dx = {vx.x - v.x, vx.y - v.y, vx.z - v.z}, dy = {vy.x - v.x, vy.y - v.y, vy.z - v.z};
                const p3 n = p3_normalize(p3_cross(dx, dy));
                out[4 * i] = n.x; out[4 * i + 1] = n.y; out[4 * i + 2] = n.z; out[4 * i + 3] = 1.f;
            }
            else
                memcpy(out + 4 * i, predicted + 4 * i, 16);
        }
}

/* Shaders/fill_rgb.frag:19-37 */
void efo_fill_rgb(const uint8_t * predicted, const uint8_t * raw, int rows, int cols, int passthrough, uint8_t * out)
{
    for(size_t i = 0; i < (size_t)rows * cols; i++)
    {
        const uint8_t * s = predicted + 4 * i;
        memcpy(out + 4 * i, ((int)s[0] + s[1] + s[2] == 0 || passthrough == 1) ? raw + 4 * i : s, 4);
    }
}
