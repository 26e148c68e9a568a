// Device-side element mathematics of the ERT path (FP64, sm_100a).
//
// Restates what the reference evaluates per cell on the CPU:
//   ElementMatrix<double>::ux2uy2uz2  core/src/elementmatrix.cpp:798-1122  (int grad N_i . grad N_j)
//   ElementMatrix<double>::u2         core/src/elementmatrix.cpp:683-796   (int N_i N_j = size * Uhat)
// for Tri3 / Tri6 / Tet4 / Tet10 on the straight-sided simplex of the corner nodes
// (shape.h:295-297).  The reference integrates P2 stiffness with degree-2 rules
// (triWeights(2) / tetWeights(2)) and mass with degree-4 rules -- exact for these
// integrands -- so closed forms in barycentric coordinates give the same numbers up
// to rounding.  Bessel K0 uses the same Abramowitz-Stegun polynomials as
// core/src/numericbase.h:80-180 (parity trap: ~1e-7 accurate by construction).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pgb {

enum Elem : int { TRI3 = 0, TRI6 = 1, TET4 = 2, TET10 = 3 };

template <int E> struct ElemTraits;
template <> struct ElemTraits<TRI3>  { static constexpr int DIM = 2, NV = 3, NL = 3,  ORDER = 1; };
template <> struct ElemTraits<TRI6>  { static constexpr int DIM = 2, NV = 3, NL = 6,  ORDER = 2; };
template <> struct ElemTraits<TET4>  { static constexpr int DIM = 3, NV = 4, NL = 4,  ORDER = 1; };
template <> struct ElemTraits<TET10> { static constexpr int DIM = 3, NV = 4, NL = 10, ORDER = 2; };

// local edge -> corner pair for the mid-side nodes (Tri6: (0-1),(1-2),(2-0);
// Tet10 Zienkiewicz: (0-1),(0-2),(0-3),(1-2),(2-3),(3-1); meshentities.h:907-912)
__device__ __forceinline__ void mid_corners(int dim, int m, int &a, int &b) {
    if (dim == 2) { a = m; b = (m + 1) % 3; }
    else {
        const int ea[6] = {0, 0, 0, 1, 2, 3};
        const int eb[6] = {1, 2, 3, 2, 3, 1};
        a = ea[m]; b = eb[m];
    }
}

// size (area / volume) and Gram matrix G[a][b] = grad(lambda_a) . grad(lambda_b)
template <int DIM>
__device__ __forceinline__ void simplex_gram(const double (&X)[DIM + 1][3], double &size, double (&G)[DIM + 1][DIM + 1]) {
    double g[DIM + 1][3];
    if (DIM == 2) {
        const double x0 = X[0][0], y0 = X[0][1], x1 = X[1][0], y1 = X[1][1], x2 = X[2][0], y2 = X[2][1];
        const double det = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);
        size = 0.5 * fabs(det);
        const double inv = 1.0 / det;
        g[0][0] = (y1 - y2) * inv; g[0][1] = (x2 - x1) * inv; g[0][2] = 0.0;
        g[1][0] = (y2 - y0) * inv; g[1][1] = (x0 - x2) * inv; g[1][2] = 0.0;
        g[2][0] = (y0 - y1) * inv; g[2][1] = (x1 - x0) * inv; g[2][2] = 0.0;
    } else {
        double e1[3], e2[3], e3[3];
#pragma unroll
        for (int d = 0; d < 3; d++) { e1[d] = X[1][d] - X[0][d]; e2[d] = X[2][d] - X[0][d]; e3[d] = X[DIM][d] - X[0][d]; }
        // rows of the inverse of [e1 e2 e3]^T are the gradients of lambda_1..3
        double c1[3] = { e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0] };
        double c2[3] = { e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0] };
        double c3[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
        const double det = e1[0] * c1[0] + e1[1] * c1[1] + e1[2] * c1[2];
        size = fabs(det) / 6.0;
        const double inv = 1.0 / det;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            g[1][d] = c1[d] * inv; g[2][d] = c2[d] * inv; g[DIM][d] = c3[d] * inv;
            g[0][d] = -(g[1][d] + g[2][d] + g[DIM][d]);
        }
    }
#pragma unroll
    for (int a = 0; a <= DIM; a++)
#pragma unroll
        for (int b = 0; b <= DIM; b++)
            G[a][b] = g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2];
}

// integral over the simplex of lambda_a * lambda_b divided by its size: (1 + delta_ab) / ((d+1)(d+2))
template <int DIM> __device__ __forceinline__ double lam2(int a, int b) {
    return (a == b ? 2.0 : 1.0) / double((DIM + 1) * (DIM + 2));
}

// stiffness entry (i, j) divided by nothing: int grad N_i . grad N_j over the cell
//   P1: size * G_ij
//   P2: grad N_corner(i) = (4 L_i - 1) grad L_i,  grad N_mid(a,b) = 4 (L_a grad L_b + L_b grad L_a)
//       integrated exactly with  int L_a = size/(d+1),  int L_a L_b = size * lam2(a,b).
template <int E>
__device__ __forceinline__ double stiff_entry(int i, int j, double size,
                                              const double (&G)[ElemTraits<E>::NV][ElemTraits<E>::NV]) {
    constexpr int DIM = ElemTraits<E>::DIM, NV = ElemTraits<E>::NV;
    if (ElemTraits<E>::ORDER == 1) return size * G[i][j];
    const double I1 = 1.0 / double(DIM + 1);
    if (i < NV && j < NV) {
        // (4Li-1)(4Lj-1) G_ij
        return size * G[i][j] * (16.0 * lam2<DIM>(i, j) - 8.0 * I1 + 1.0);
    }
    if (i < NV || j < NV) {
        const int c = i < NV ? i : j;
        const int m = (i < NV ? j : i) - NV;
        int a, b; mid_corners(DIM, m, a, b);
        // (4Lc-1) * 4 (La G_cb + Lb G_ca)
        return size * 4.0 * (G[c][b] * (4.0 * lam2<DIM>(c, a) - I1) + G[c][a] * (4.0 * lam2<DIM>(c, b) - I1));
    }
    int a, b, c, d; mid_corners(DIM, i - NV, a, b); mid_corners(DIM, j - NV, c, d);
    // 16 (La grad Lb + Lb grad La).(Lc grad Ld + Ld grad Lc)
    return size * 16.0 * (lam2<DIM>(a, c) * G[b][d] + lam2<DIM>(a, d) * G[b][c] +
                          lam2<DIM>(b, c) * G[a][d] + lam2<DIM>(b, d) * G[a][c]);
}

// unit mass matrix entry Uhat_ij = (int N_i N_j) / size  (exact rational values)
template <int E>
__device__ __forceinline__ double mass_unit(int i, int j) {
    constexpr int DIM = ElemTraits<E>::DIM, NV = ElemTraits<E>::NV;
    if (ElemTraits<E>::ORDER == 1) return lam2<DIM>(i, j);
    if (DIM == 2) {
        if (i < NV && j < NV) return (i == j ? 6.0 : -1.0) / 180.0;
        if (i < NV || j < NV) {
            const int c = i < NV ? i : j, m = (i < NV ? j : i) - NV;
            int a, b; mid_corners(2, m, a, b);
            return (c == a || c == b) ? 0.0 : -4.0 / 180.0;
        }
        return (i == j ? 32.0 : 16.0) / 180.0;
    } else {
        if (i < NV && j < NV) return (i == j ? 6.0 : 1.0) / 420.0;
        if (i < NV || j < NV) {
            const int c = i < NV ? i : j, m = (i < NV ? j : i) - NV;
            int a, b; mid_corners(3, m, a, b);
            return ((c == a || c == b) ? -4.0 : -6.0) / 420.0;
        }
        if (i == j) return 32.0 / 420.0;
        int a, b, c, d; mid_corners(3, i - NV, a, b); mid_corners(3, j - NV, c, d);
        const bool share = (a == c || a == d || b == c || b == d);
        return (share ? 16.0 : 8.0) / 420.0;
    }
}

// ---- Abramowitz & Stegun 9.8.1 / 9.8.5 / 9.8.6 -------------------------------------
__device__ __forceinline__ double as_bessel_i0(double x) {
    const double ax = fabs(x);
    if (ax < 3.75) {
        double y = x / 3.75; y = y * y;
        return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    }
    const double y = 3.75 / ax;
    return (exp(ax) / sqrt(ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
           y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
}
__device__ __forceinline__ double as_bessel_k0(double x) {
    if (x <= 2.0) {
        const double y = x * x / 4.0;
        return (-log(x / 2.0) * as_bessel_i0(x)) + (-0.57721566 + y * (0.42278420 + y * (0.23069756 + y * (0.3488590e-1 +
               y * (0.262698e-2 + y * (0.10750e-3 + y * 0.74e-5))))));
    }
    const double y = 2.0 / x;
    return (exp(-x) / sqrt(x)) * (1.25331414 + y * (-0.7832358e-1 + y * (0.2189568e-1 + y * (-0.1062446e-1 +
           y * (0.587872e-2 + y * (-0.251540e-2 + y * 0.53208e-3))))));
}

} // namespace pgb
