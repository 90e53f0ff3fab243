// pfc_bristle.cuh -- per-patch 6x6 stiffness scaling and K^(-1/2) for the bristle friction model.
//
// Reference: decompose_K! / calc_K̄_sqrt_inv / calc_trace_12
// (/root/reference/src/contact_algorithms_friction.jl:85-117).  The reference calls LAPACK's
// symmetric eigensolver (Float64) or GenericLinearAlgebra (Dual); only V f(L) V' is consumed, so
// the eigenvector sign/order convention is irrelevant.  Here: cyclic Jacobi rotations on the
// scaled matrix, run redundantly by every lane of the warp that owns the instruction (all lanes
// hold identical inputs after the butterfly reduction, so no broadcast is needed).  In Dual mode
// the partials of V f(L) V' follow from the Daleckii-Krein formula.
#pragma once
#include "pfc_math.cuh"

namespace pfc {

// A: symmetric 6x6 (full storage, row-major), destroyed.  V: eigenvectors in columns.
static __device__ __noinline__ void jacobi6(double* A, double* V, double* lam) {
    for (int i = 0; i < 36; ++i) V[i] = 0.0;
    for (int i = 0; i < 6; ++i) V[7 * i] = 1.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < 6; ++i) {
            dia += A[7 * i] * A[7 * i];
            for (int j = i + 1; j < 6; ++j) off += A[6 * i + j] * A[6 * i + j];
        }
        if (off == 0.0 || off <= 1.0e-42 * dia) break;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                const double apq = A[6 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[7 * q] - A[7 * p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < 6; ++k) {
                    const double akp = A[6 * k + p], akq = A[6 * k + q];
                    A[6 * k + p] = c * akp - s * akq;
                    A[6 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 6; ++k) {
                    const double apk = A[6 * p + k], aqk = A[6 * q + k];
                    A[6 * p + k] = c * apk - s * aqk;
                    A[6 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 6; ++k) {
                    const double vkp = V[6 * k + p], vkq = V[6 * k + q];
                    V[6 * k + p] = c * vkp - s * vkq;
                    V[6 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 6; ++i) lam[i] = A[7 * i];
}

// a21: K11 upper (0..5), K12 row-major (6..14), K22 upper (15..20), already multiplied by k_bar.
// Outputs: Sinv (6) and Khalf = K̄^(-1/2) (36, row-major).  scratch: 108 doubles of (shared) memory.
static __device__ __noinline__ void decompose_K(const double* a21, double magic, double* Sinv, double* Khalf, double* scratch) {
    double* K = scratch;
    double* A = scratch + 36;
    double* V = scratch + 72;
    {
        // K11
        K[0] = a21[0]; K[1] = a21[1]; K[2] = a21[2]; K[7] = a21[3]; K[8] = a21[4]; K[14] = a21[5];
        K[6] = K[1]; K[12] = K[2]; K[13] = K[8];
        // K12 (rows 0..2, cols 3..5) and its transpose
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { K[6 * i + 3 + j] = a21[6 + 3 * i + j]; K[6 * (3 + j) + i] = a21[6 + 3 * i + j]; }
        // K22
        K[21] = a21[15]; K[22] = a21[16]; K[23] = a21[17]; K[28] = a21[18]; K[29] = a21[19]; K[35] = a21[20];
        K[27] = K[22]; K[33] = K[23]; K[34] = K[29];
    }
    const double t1 = K[0] + K[7] + K[14];
    const double t2 = K[21] + K[28] + K[35];
    const double s1 = 1.0 / sqrt(t1), s2 = 1.0 / sqrt(t2);
    for (int k = 0; k < 3; ++k) { Sinv[k] = s1 * magic; Sinv[3 + k] = s2; }
    double lam[6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A[6 * i + j] = Sinv[i] * K[6 * i + j] * Sinv[j];
    jacobi6(A, V, lam);
    double mx = lam[0];
    for (int k = 1; k < 6; ++k) mx = fmax(mx, lam[k]);
    double sig[6];
    for (int k = 0; k < 6; ++k) sig[k] = 1.0 / sqrt(fmax(lam[k], mx * 1.0e-16));
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += (V[6 * i + k] * sig[k]) * V[6 * j + k];
            Khalf[6 * i + j] = acc;
        }
}

}  // namespace pfc
