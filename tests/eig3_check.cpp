// eig3_check.cpp -- compares the product's closed-form smallest-eigenpair routine (csrc/eig3.cuh,
// compiled for the host) against the oracle's restatement of Eigen's QR solver.
// Prints max relative eigenvalue error and max normal deviation over well-separated cases.
#include "../oracle/tmap_oracle.cpp"
#include "../traversability_mapping_b200/csrc/eig3.cuh"
#include <cstdio>
#include <random>

int main()
{
    std::mt19937_64 rng(123);
    std::normal_distribution<double> g(0.0, 1.0);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    double max_rel_lam = 0, max_vec = 0, max_rel_lam_all = 0;
    int n_sep = 0, n_all = 0, n_bad = 0;
    for (int it = 0; it < 400000; ++it)
    {
        // terrain-like covariance: points in a disc of radius ~0.2..1 m on a tilted plane with noise,
        // a fraction with step-like vertical spread, absolute z up to 100 m (cancellation in Q/N - mu mu^T)
        const int npts = 40 + (int)(u(rng) * 200);
        const double R = 0.1 + u(rng) * 0.9, tiltx = g(rng) * 0.3, tilty = g(rng) * 0.3;
        const double sig = std::pow(10.0, -3.5 + 3.0 * u(rng)), z0 = (u(rng) < 0.3 ? 100.0 : 2.0) * g(rng);
        const double stepamp = u(rng) < 0.25 ? u(rng) * 0.5 : 0.0;
        double S[3] = {0, 0, 0}, Q[3][3] = {{0}};
        for (int k = 0; k < npts; ++k)
        {
            const double x = (2 * u(rng) - 1) * R, y = (2 * u(rng) - 1) * R;
            const double z = z0 + tiltx * x + tilty * y + sig * g(rng) + (x > 0 ? stepamp : 0.0);
            const double p[3] = {x, y, z};
            for (int i = 0; i < 3; ++i) { S[i] += p[i]; for (int j = 0; j < 3; ++j) Q[i][j] += p[i] * p[j]; }
        }
        double A[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = Q[i][j] / npts - (S[i] / npts) * (S[j] / npts);
        const Eig3 e = eigh3(A);
        double v[3];
        const double lam = tmb::eig3_min(A[0][0], A[0][1], A[0][2], A[1][1], A[1][2], A[2][2], v);
        const double rel = std::abs(lam - e.w[0]) / std::max(std::abs(e.w[0]), 1e-300);
        ++n_all;
        max_rel_lam_all = std::max(max_rel_lam_all, rel);
        const double gap = (e.w[1] - e.w[0]) / std::max(e.w[2], 1e-300);
        if (gap > 1e-6)
        {
            ++n_sep;
            if (e.w[0] > 1e-12 * e.w[2]) max_rel_lam = std::max(max_rel_lam, rel);
            double dot = v[0] * e.v[0][0] + v[1] * e.v[1][0] + v[2] * e.v[2][0];
            double dev = 1.0 - std::abs(dot);
            max_vec = std::max(max_vec, dev);
            if (dev > 1e-9 || (rel > 1e-6 && e.w[0] > 1e-12 * e.w[2])) ++n_bad;
        }
    }
    std::printf("cases %d separated %d max_rel_lambda %.3e max_rel_lambda_all %.3e max_1-|dot| %.3e bad %d\n", n_all, n_sep,
                max_rel_lam, max_rel_lam_all, max_vec, n_bad);
    return n_bad != 0;
}
