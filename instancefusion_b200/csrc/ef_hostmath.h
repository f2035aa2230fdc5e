// ef_hostmath.h -- host-side (and, where marked, device-side) small dense linear algebra of the tracker
// driver: the operations RGBDOdometry.cpp delegates to Eigen (an un-vendored dependency of the
// reference): fixed-size inverses, symmetric-pivoted LDL^T solves, Rodrigues' formula
// (OdometryProvider.h:35-71) and the SE(3) update (OdometryProvider.h:73-93).
//
// Everything is templated on the scalar and written without FMA-sensitive tricks so the SAME source
// runs on the host (EF_SOLVE_HOST) and inside the persistent kernel's single solver thread
// (EF_SOLVE_DEVICE).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define EF_HD __host__ __device__ __forceinline__
#else
#define EF_HD inline
#endif

namespace ef
{
namespace hm
{

template<class T> EF_HD T absT(T v) { return v < T(0) ? -v : v; }

// C = A * B, 3x3 row-major (C may alias A or B)
template<class T> EF_HD void mul33(const T * A, const T * B, T * C)
{
    T r[9];
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
#pragma unroll
        for(int j = 0; j < 3; j++) r[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
    }
#pragma unroll
    for(int i = 0; i < 9; i++) C[i] = r[i];
}

// adjugate / determinant inverse (what Eigen uses for fixed 3x3)
template<class T> EF_HD void inverse33(const T * m, T * o)
{
    const T c00 = m[4] * m[8] - m[5] * m[7];
    const T c01 = m[5] * m[6] - m[3] * m[8];
    const T c02 = m[3] * m[7] - m[4] * m[6];
    const T det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const T id = T(1) / det;
    T r[9];
    r[0] = c00 * id; r[1] = (m[2] * m[7] - m[1] * m[8]) * id; r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r[3] = c01 * id; r[4] = (m[0] * m[8] - m[2] * m[6]) * id; r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r[6] = c02 * id; r[7] = (m[1] * m[6] - m[0] * m[7]) * id; r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
#pragma unroll
    for(int i = 0; i < 9; i++) o[i] = r[i];
}

// general NxN inverse, Gauss-Jordan with partial pivoting (N = 4 for resultRt, 6 for the covariance)
template<class T, int N> EF_HD void inverseNN(const T * M, T * Minv)
{
    T a[N][2 * N];
    for(int i = 0; i < N; i++)
        for(int j = 0; j < N; j++)
        {
            a[i][j] = M[i * N + j];
            a[i][N + j] = (i == j) ? T(1) : T(0);
        }
    for(int c = 0; c < N; c++)
    {
        int p = c;
        for(int r = c + 1; r < N; r++)
            if(absT(a[r][c]) > absT(a[p][c])) p = r;
        if(p != c)
            for(int j = 0; j < 2 * N; j++)
            {
                const T t = a[c][j];
                a[c][j] = a[p][j];
                a[p][j] = t;
            }
        const T inv = T(1) / a[c][c];
        for(int j = 0; j < 2 * N; j++) a[c][j] *= inv;
        for(int r = 0; r < N; r++)
            if(r != c)
            {
                const T f = a[r][c];
                if(f != T(0))
                    for(int j = 0; j < 2 * N; j++) a[r][j] -= f * a[c][j];
            }
    }
    for(int i = 0; i < N; i++)
        for(int j = 0; j < N; j++) Minv[i * N + j] = a[i][N + j];
}

// x = A^-1 b through a symmetric-pivoted LDL^T (Eigen's A.ldlt().solve(b): pivot = largest remaining
// |diagonal|).  Every loop bound and array index is a compile-time constant -- the pivot exchange is a chain
// of predicated swaps -- so on the device the whole factorisation lives in registers (dynamic indexing
// would put A in local memory and cost the single solver thread tens of microseconds).
template<class T> EF_HD void swap_if(bool c, T & a, T & b)
{
    const T ta = a, tb = b;
    a = c ? tb : ta;
    b = c ? ta : tb;
}

template<class T, int N> EF_HD void ldlt_solve(const T * A_in, const T * b, T * x)
{
    T A[N * N], y[N];
    int perm[N];
#pragma unroll
    for(int i = 0; i < N * N; i++) A[i] = A_in[i];
#pragma unroll
    for(int i = 0; i < N; i++) perm[i] = i;
#pragma unroll
    for(int k = 0; k < N; k++)
    {
        int p = k;
        T best = absT(A[k * N + k]);
#pragma unroll
        for(int i = k + 1; i < N; i++)
        {
            const T d = absT(A[i * N + i]);
            const bool g = d > best;
            best = g ? d : best;
            p = g ? i : p;
        }
#pragma unroll
        for(int i = k + 1; i < N; i++)
        {
            const bool sw = (p == i);
#pragma unroll
            for(int j = 0; j < N; j++) swap_if(sw, A[k * N + j], A[i * N + j]); // rows k <-> i
#pragma unroll
            for(int j = 0; j < N; j++) swap_if(sw, A[j * N + k], A[j * N + i]); // cols k <-> i
            swap_if(sw, perm[k], perm[i]);
        }
        const T d = A[k * N + k];
        if(d != T(0))
        {
#pragma unroll
            for(int i = k + 1; i < N; i++) A[i * N + k] /= d;
#pragma unroll
            for(int i = k + 1; i < N; i++)
            {
#pragma unroll
                for(int j = k + 1; j <= i; j++)
                {
                    A[i * N + j] -= A[i * N + k] * d * A[j * N + k];
                    A[j * N + i] = A[i * N + j];
                }
            }
        }
    }
    // y = P b
#pragma unroll
    for(int i = 0; i < N; i++)
    {
        T v = T(0);
#pragma unroll
        for(int j = 0; j < N; j++) v = (perm[i] == j) ? b[j] : v;
        y[i] = v;
    }
#pragma unroll
    for(int i = 0; i < N; i++)
    {
#pragma unroll
        for(int j = 0; j < i; j++) y[i] -= A[i * N + j] * y[j];
    }
#pragma unroll
    for(int i = 0; i < N; i++) y[i] = (A[i * N + i] != T(0)) ? y[i] / A[i * N + i] : T(0);
#pragma unroll
    for(int i = N - 1; i >= 0; i--)
    {
#pragma unroll
        for(int j = i + 1; j < N; j++) y[i] -= A[j * N + i] * y[j];
    }
    // x = P^T y
#pragma unroll
    for(int j = 0; j < N; j++)
    {
        T v = T(0);
#pragma unroll
        for(int i = 0; i < N; i++) v = (perm[i] == j) ? y[i] : v;
        x[j] = v;
    }
}

// inverse of an affine 4x4 [A t; 0 0 0 1] (row-major): [A^-1, -A^-1 t; 0 0 0 1].  This IS the general
// inverse for such a matrix (no orthogonality assumed); resultRt keeps an exact (0,0,0,1) last row because
// every update multiplies by a matrix with that last row (OdometryProvider.h:79-88).
template<class T> EF_HD void inverse_affine44(const T * M, T * Minv)
{
    const T A[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
    T Ai[9];
    inverse33(A, Ai);
#pragma unroll
    for(int r = 0; r < 3; r++)
    {
#pragma unroll
        for(int c = 0; c < 3; c++) Minv[r * 4 + c] = Ai[r * 3 + c];
        Minv[r * 4 + 3] = -(Ai[r * 3 + 0] * M[3] + Ai[r * 3 + 1] * M[7] + Ai[r * 3 + 2] * M[11]);
    }
    Minv[12] = T(0); Minv[13] = T(0); Minv[14] = T(0); Minv[15] = T(1);
}

// OdometryProvider.h:35-71
EF_HD void rodrigues(const double * src, double * R)
{
#pragma unroll
    for(int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    double rx = src[0], ry = src[1], rz = src[2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if(theta >= 2.2204460492503131e-16)
    {
        const double c = cos(theta), s = sin(theta), c1 = 1. - c;
        const double itheta = theta ? 1. / theta : 0.;
        rx *= itheta; ry *= itheta; rz *= itheta;
        const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
        const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
#pragma unroll
        for(int k = 0; k < 9; k++) R[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * r_x[k];
    }
}

// OdometryProvider.h:73-93: resultRt = [rodrigues(x[3:6]) | x[0:3]] * resultRt (row-major 4x4 double).
// Both factors are affine (last row exactly 0 0 0 1), so the terms the reference's full 4x4 product adds
// for that row are exact zeros / ones: this evaluates the same sums in the same order.
EF_HD void update_se3(double * resultRt, const double * x)
{
    double R[9], N[16];
    rodrigues(x + 3, R);
#pragma unroll
    for(int r = 0; r < 3; r++)
    {
#pragma unroll
        for(int c = 0; c < 3; c++)
            N[r * 4 + c] = R[r * 3 + 0] * resultRt[0 * 4 + c] + R[r * 3 + 1] * resultRt[1 * 4 + c] + R[r * 3 + 2] * resultRt[2 * 4 + c];
        N[r * 4 + 3] = R[r * 3 + 0] * resultRt[3] + R[r * 3 + 1] * resultRt[7] + R[r * 3 + 2] * resultRt[11] + x[r];
    }
#pragma unroll
    for(int i = 0; i < 12; i++) resultRt[i] = N[i];
    resultRt[12] = 0.0; resultRt[13] = 0.0; resultRt[14] = 0.0; resultRt[15] = 1.0;
}

// Unpivoted LDL^T solve of a symmetric positive-definite 6x6 system given as the 27 leading entries of the
// JtJJtrSE3 accumulator (types.cuh:101-152: for i < 6, for j = i..6: [J|r]_i * [J|r]_j, so row i holds A(i,i..5)
// followed by b(i)) -- used by the single device solver thread, where Eigen-style pivoting would cost hundreds
// of predicated swaps.  For the SPD normal equations of the tracker both factorisations give the same solution
// to ~1e-15 relative; a zero pivot (no correspondences) yields x = 0 like the pivoted routine.
EF_HD int acc_index(int i, int j) { return i * 7 - i * (i - 1) / 2 + (j - i); } // j >= i, j == 6 selects b(i)

EF_HD void ldlt_solve_spd6_acc(double * a, double * x)
{
    // In place on the accumulator (destroyed), right-looking: pivot j scales its row once (ONE reciprocal per pivot)
    // and updates the trailing rows including the right-hand side column, so the forward substitution rides along
    // and all updates of a pivot are independent.  Afterwards a(j,i) = L(i,j) and a(j,6) = (L^-1 b)(j).
    double inv[6];
#pragma unroll
    for(int j = 0; j < 6; j++)
    {
        const double dj = a[acc_index(j, j)];
        inv[j] = (dj != 0.0) ? 1.0 / dj : 0.0;
        double l[6];
#pragma unroll
        for(int i = j + 1; i < 6; i++) l[i] = a[acc_index(j, i)] * inv[j];
#pragma unroll
        for(int i = j + 1; i < 6; i++)
        {
#pragma unroll
            for(int k = i; k < 7; k++) a[acc_index(i, k)] -= l[i] * a[acc_index(j, k)];
        }
#pragma unroll
        for(int i = j + 1; i < 6; i++) a[acc_index(j, i)] = l[i];
    }
#pragma unroll
    for(int i = 5; i >= 0; i--)
    {
        double v = a[acc_index(i, 6)] * inv[i];
#pragma unroll
        for(int k = i + 1; k < 6; k++) v -= a[acc_index(i, k)] * x[k];
        x[i] = v;
    }
}

// RGBDOdometry.cpp:571-583: [Rcurr|tcurr] = [Rprev|tprev] * (float(resultRt))^-1, with the Isometry3f
// inverse (transpose rotation, -R^T t), all in float.
EF_HD void compose_pose(const double * resultRt, const float * Rprev, const float * tprev, float * Rcurr, float * tcurr)
{
    float oR[9], ot[3], iR[9], it[3];
#pragma unroll
    for(int r = 0; r < 3; r++)
    {
#pragma unroll
        for(int c = 0; c < 3; c++) oR[r * 3 + c] = (float)resultRt[r * 4 + c];
        ot[r] = (float)resultRt[r * 4 + 3];
    }
#pragma unroll
    for(int r = 0; r < 3; r++)
#pragma unroll
        for(int c = 0; c < 3; c++) iR[r * 3 + c] = oR[c * 3 + r];
#pragma unroll
    for(int r = 0; r < 3; r++) it[r] = -(iR[r * 3] * ot[0] + iR[r * 3 + 1] * ot[1] + iR[r * 3 + 2] * ot[2]);
#pragma unroll
    for(int r = 0; r < 3; r++)
    {
#pragma unroll
        for(int c = 0; c < 3; c++) Rcurr[r * 3 + c] = Rprev[r * 3] * iR[c] + Rprev[r * 3 + 1] * iR[3 + c] + Rprev[r * 3 + 2] * iR[6 + c];
        tcurr[r] = Rprev[r * 3] * it[0] + Rprev[r * 3 + 1] * it[1] + Rprev[r * 3 + 2] * it[2] + tprev[r];
    }
}

// the same from the 3x4 block alone (row-major, 12 entries)
EF_HD void compose_pose_affine12(const double * Rt12, const float * Rprev, const float * tprev, float * Rcurr, float * tcurr)
{
    const double M[16] = {Rt12[0], Rt12[1], Rt12[2], Rt12[3], Rt12[4], Rt12[5], Rt12[6], Rt12[7], Rt12[8], Rt12[9], Rt12[10], Rt12[11], 0, 0, 0, 1};
    compose_pose(M, Rprev, tprev, Rcurr, tcurr);
}

// RGBDOdometry.cpp:424-434: from resultRt build krkInv = K R K^-1 and kt = K t of Rt = resultRt^-1
EF_HD void rgb_warp_params(const double * resultRt, const double * K, const double * K_inv, float * krkinv9, float * kt3)
{
    double Rt[16];
    inverse_affine44(resultRt, Rt);
    const double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
    double tmp[9], KRK[9];
    mul33(K, R, tmp);
    mul33(tmp, K_inv, KRK);
#pragma unroll
    for(int i = 0; i < 9; i++) krkinv9[i] = (float)KRK[i];
    const double tv[3] = {Rt[3], Rt[7], Rt[11]};
#pragma unroll
    for(int r = 0; r < 3; r++) kt3[r] = (float)(K[r * 3] * tv[0] + K[r * 3 + 1] * tv[1] + K[r * 3 + 2] * tv[2]);
}

// K = [fx 0 cx; 0 fy cy; 0 0 1] and its inverse have five structural zeros: K R K^-1 and K t with the zero terms
// dropped.  Dropping  + 0 * x  terms is exact, so this equals the dense products of rgb_warp_params / the SO(3)
// homography (RGBDOdometry.cpp:318-329, :424-434) evaluated with fused multiply-adds; used by the device solver
// thread, where every double-precision instruction is on the critical path of the iteration.
EF_HD void krk_sparse(const double * R, double fx, double fy, double cx, double cy, const double * K_inv, double * KR, double * KRK)
{
#pragma unroll
    for(int j = 0; j < 3; j++)
    {
        KR[j] = fma(cx, R[6 + j], fx * R[j]);
        KR[3 + j] = fma(cy, R[6 + j], fy * R[3 + j]);
        KR[6 + j] = R[6 + j];
    }
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
        KRK[i * 3 + 0] = KR[i * 3 + 0] * K_inv[0];
        KRK[i * 3 + 1] = KR[i * 3 + 1] * K_inv[4];
        KRK[i * 3 + 2] = fma(KR[i * 3 + 2], K_inv[8], fma(KR[i * 3 + 1], K_inv[5], KR[i * 3 + 0] * K_inv[2]));
    }
}

EF_HD void rgb_warp_params_sparse(const double * resultRt, double fx, double fy, double cx, double cy, const double * K_inv, float * krkinv9,
                                  float * kt3)
{
    double Rt[16];
    inverse_affine44(resultRt, Rt);
    const double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
    double KR[9], KRK[9];
    krk_sparse(R, fx, fy, cx, cy, K_inv, KR, KRK);
#pragma unroll
    for(int i = 0; i < 9; i++) krkinv9[i] = (float)KRK[i];
    kt3[0] = (float)fma(cx, Rt[11], fx * Rt[3]);
    kt3[1] = (float)fma(cy, Rt[11], fy * Rt[7]);
    kt3[2] = (float)Rt[11];
}

// unpack the 29-float accumulator into A (6x6 row-major), b, residual[2]: reduce.cu:475-489
template<class TA> EF_HD void unpack_se3(const float * h, TA * A, TA * b, float * residual)
{
    int shift = 0;
#pragma unroll
    for(int i = 0; i < 6; ++i)
#pragma unroll
        for(int j = i; j < 7; ++j)
        {
            const float value = h[shift++];
            if(j == 6) b[i] = (TA)value;
            else A[j * 6 + i] = A[i * 6 + j] = (TA)value;
        }
    if(residual)
    {
        residual[0] = h[27];
        residual[1] = h[28];
    }
}

// reduce.cu:1126-1140
EF_HD void unpack_so3(const float * h, float * A, float * b, float * residual)
{
    int shift = 0;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = i; j < 4; ++j)
        {
            const float value = h[shift++];
            if(j == 3) b[i] = value;
            else A[j * 3 + i] = A[i * 3 + j] = value;
        }
    residual[0] = h[9];
    residual[1] = h[10];
}

} // namespace hm
} // namespace ef
