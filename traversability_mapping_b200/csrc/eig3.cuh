// eig3.cuh -- smallest eigenpair of a symmetric 3x3 matrix in fp64.
//
// Replaces Eigen::SelfAdjointEigenSolver<Matrix3d> at TraversabilityMetrics.cpp:85-88, where only
// eigenvalues()(0) and eigenvectors().col(0) are consumed.  Not Eigen's QR iteration: closed-form
// (trigonometric) eigenvalues of the scaled, trace-shifted matrix, the eigenvector from the
// best-conditioned cross product of rows of (A - lambda I), then one Rayleigh-quotient update of
// lambda.  Accuracy ~1e-15*|A| on lambda and ~1e-15*|A|/gap on the vector, which is what the
// 1e-4 parity bar needs (tests/eig3_check.cpp (run by tests/test_abi.py::test_eig3_closed_form_vs_oracle_qr) compares against the oracle's QR restatement).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TMB_HD __host__ __device__ __forceinline__
#else
#define TMB_HD inline
#endif

namespace tmb
{
// a = (a00, a01, a02, a11, a12, a22).  Returns lambda_min; v = unit eigenvector (sign arbitrary).
TMB_HD double eig3_min(double a00, double a01, double a02, double a11, double a12, double a22, double v[3])
{
    double scale = fmax(fmax(fabs(a00), fabs(a11)), fabs(a22));
    scale = fmax(scale, fmax(fmax(fabs(a01), fabs(a02)), fabs(a12)));
    if (!(scale > 0.0))
    {
        // zero matrix: Eigen leaves Q = I, so column 0 = e_x with eigenvalue 0
        v[0] = 1.0; v[1] = 0.0; v[2] = 0.0;
        return 0.0;
    }
    const double is = 1.0 / scale;
    a00 *= is; a01 *= is; a02 *= is; a11 *= is; a12 *= is; a22 *= is;
    const double m = (a00 + a11 + a22) * (1.0 / 3.0);
    const double k00 = a00 - m, k11 = a11 - m, k22 = a22 - m;
    const double p = (k00 * k00 + k11 * k11 + k22 * k22 + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12)) * (1.0 / 6.0);
    if (!(p > 1e-300))
    {
        v[0] = 1.0; v[1] = 0.0; v[2] = 0.0;
        return m * scale;
    }
    const double q = 0.5 * (k00 * (k11 * k22 - a12 * a12) - a01 * (a01 * k22 - a12 * a02) + a02 * (a01 * a12 - k11 * a02));
    const double sp = sqrt(p);
    double disc = p * p * p - q * q;
    if (disc < 0.0) disc = 0.0;
    const double phi = atan2(sqrt(disc), q) * (1.0 / 3.0);  // [0, pi/3]
    double sn, cs;
#if defined(__CUDA_ARCH__)
    sincos(phi, &sn, &cs);
#else
    sn = sin(phi); cs = cos(phi);
#endif
    double lam = m - sp * (cs + 1.7320508075688772 * sn);

    // eigenvector: rows of B = A - lam I; the largest pairwise cross product spans null(B)
    for (int it = 0; it < 2; ++it)
    {
        const double b00 = a00 - lam, b11 = a11 - lam, b22 = a22 - lam;
        const double c0x = a01 * a12 - a02 * b11, c0y = a02 * a01 - b00 * a12, c0z = b00 * b11 - a01 * a01;  // r0 x r1
        const double c1x = a01 * b22 - a02 * a12, c1y = a02 * a02 - b00 * b22, c1z = b00 * a12 - a01 * a02;  // r0 x r2
        const double c2x = b11 * b22 - a12 * a12, c2y = a12 * a02 - a01 * b22, c2z = a01 * a12 - b11 * a02;  // r1 x r2
        const double n0 = c0x * c0x + c0y * c0y + c0z * c0z;
        const double n1 = c1x * c1x + c1y * c1y + c1z * c1z;
        const double n2 = c2x * c2x + c2y * c2y + c2z * c2z;
        double x, y, z, nn;
        if (n0 >= n1 && n0 >= n2) { x = c0x; y = c0y; z = c0z; nn = n0; }
        else if (n1 >= n2) { x = c1x; y = c1y; z = c1z; nn = n1; }
        else { x = c2x; y = c2y; z = c2z; nn = n2; }
        if (!(nn > 1e-60))
        {
            // lam is (numerically) a repeated eigenvalue: null(B) is a plane; take any unit vector
            // orthogonal to the dominant row of B.
            const double r0 = b00 * b00 + a01 * a01 + a02 * a02;
            const double r1 = a01 * a01 + b11 * b11 + a12 * a12;
            const double r2 = a02 * a02 + a12 * a12 + b22 * b22;
            double rx, ry, rz;
            if (r0 >= r1 && r0 >= r2) { rx = b00; ry = a01; rz = a02; }
            else if (r1 >= r2) { rx = a01; ry = b11; rz = a12; }
            else { rx = a02; ry = a12; rz = b22; }
            if (fabs(rx) <= fabs(ry) && fabs(rx) <= fabs(rz)) { x = 0.0; y = -rz; z = ry; }
            else if (fabs(ry) <= fabs(rz)) { x = -rz; y = 0.0; z = rx; }
            else { x = -ry; y = rx; z = 0.0; }
            nn = x * x + y * y + z * z;
            if (!(nn > 0.0)) { x = 1.0; y = 0.0; z = 0.0; nn = 1.0; }
        }
        const double inv = 1.0 / sqrt(nn);
        x *= inv; y *= inv; z *= inv;
        v[0] = x; v[1] = y; v[2] = z;
        // Rayleigh quotient: second-order accurate eigenvalue for the vector just found
        const double ax = a00 * x + a01 * y + a02 * z;
        const double ay = a01 * x + a11 * y + a12 * z;
        const double az = a02 * x + a12 * y + a22 * z;
        lam = x * ax + y * ay + z * az;
    }
    return lam * scale;
}
}  // namespace tmb
