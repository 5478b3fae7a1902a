/* tmap_oracle.cpp -- CPU ORACLE (test infrastructure only; see tmap_oracle.h).
 *
 * Dependency-free restatement of the reference's CPU path.  Every function names the
 * reference file:line it follows (paths relative to /root/reference/traversability_mapping).
 * Built with `g++ -O2 -ffp-contract=off` (x86-64 baseline, no FMA, no fast-math) to match the
 * reference's Release build semantics (CMakeLists.txt:4-6,15-16).
 */
#include "tmap_oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <unordered_set>
#include <vector>
#include <thread>
#include <atomic>

namespace tmo_detail
{
constexpr float kNaNf = std::numeric_limits<float>::quiet_NaN();
constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();

/* ---- Moments.hpp:40-88, Moments.cpp:18-58 ------------------------------------------- */
struct Moments
{
    unsigned int N = 0;
    double sx = 0, sy = 0, sz = 0, sx2 = 0, sy2 = 0, sz2 = 0, sxy = 0, sxz = 0, syz = 0;
    /* extension (not in the reference): extreme map-frame z of the inserted points */
    float zmin = kNaNf, zmax = kNaNf;

    void insert(double x, double y, double z)  /* Moments.hpp:53-59 */
    {
        ++N;
        sx += x;  sy += y;  sz += z;
        sx2 += x * x;  sy2 += y * y;  sz2 += z * z;
        sxy += x * y;  sxz += x * z;  syz += y * z;
    }
    void fuse(const Moments &o)  /* Moments.hpp:63-69 */
    {
        N += o.N;
        sx += o.sx;  sy += o.sy;  sz += o.sz;
        sx2 += o.sx2;  sy2 += o.sy2;  sz2 += o.sz2;
        sxy += o.sxy;  sxz += o.sxz;  syz += o.syz;
    }
    /* Moments.cpp:49-58: S' = S + N d ; Q' = ((Q + S d^T) + d S^T) + (N d) d^T, element-wise */
    void shift(double dx, double dy, double dz)
    {
        const double s[3] = {sx, sy, sz}, d[3] = {dx, dy, dz};
        const double q[3][3] = {{sx2, sxy, sxz}, {sxy, sy2, syz}, {sxz, syz, sz2}};
        const double n = static_cast<double>(N);
        double q2[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                q2[i][j] = ((q[i][j] + s[i] * d[j]) + d[i] * s[j]) + (n * d[i]) * d[j];
        sx = s[0] + n * d[0];  sy = s[1] + n * d[1];  sz = s[2] + n * d[2];
        sx2 = q2[0][0];  sy2 = q2[1][1];  sz2 = q2[2][2];
        sxy = q2[0][1];  sxz = q2[0][2];  syz = q2[1][2];
    }
};

/* ---- Moments.hpp:90-133 -------------------------------------------------------------- */
struct Lattice
{
    double x0 = 0, y0 = 0, res = 0.25;
    void cellOf(double x, double y, int &ci, int &cj) const
    {
        ci = static_cast<int>(std::lround((x - x0) / res));
        cj = static_cast<int>(std::lround((y - y0) / res));
    }
    double cx(int ci) const { return x0 + ci * res; }
    double cy(int cj) const { return y0 + cj * res; }
    static std::uint64_t key(int ci, int cj)
    {
        return (static_cast<std::uint64_t>(static_cast<std::uint32_t>(ci)) << 32) |
               static_cast<std::uint64_t>(static_cast<std::uint32_t>(cj));
    }
    static void unkey(std::uint64_t k, int &ci, int &cj)
    {
        ci = static_cast<int>(static_cast<std::uint32_t>(k >> 32));
        cj = static_cast<int>(static_cast<std::uint32_t>(k & 0xffffffffu));
    }
};

/* Eigen::Affine3f * Vector3f (KeyFrame.cpp:57, System.cpp:105).  Eigen 3.4 evaluates the
 * 3x3*3x1 block coefficient-wise, r_i = t_i + ((R_i0 x + R_i1 y) + R_i2 z), every product and
 * sum rounded to float (no FMA on the reference's x86-64 baseline build).  T is row-major 3x4. */
inline void xform(const float T[12], float x, float y, float z, float out[3])
{
    for (int i = 0; i < 3; ++i)
    {
        const float a = T[4 * i + 0] * x;
        const float b = T[4 * i + 1] * y;
        const float c = T[4 * i + 2] * z;
        const float s = (a + b) + c;
        out[i] = s + T[4 * i + 3];
    }
}

/* Helpers.cpp:156-169 */
std::vector<std::pair<int, int>> discOffsets(double radius, double res)
{
    const double rc = std::max(1.0, radius / res);
    const int r = static_cast<int>(std::ceil(rc));
    const double rc2 = rc * rc;
    std::vector<std::pair<int, int>> out;
    for (int di = -r; di <= r; ++di)
        for (int dj = -r; dj <= r; ++dj)
            if (static_cast<double>(di * di + dj * dj) <= rc2)
                out.emplace_back(di, dj);
    return out;
}

/* ---- Eigen 3.4 SelfAdjointEigenSolver<Matrix3d>::compute, restated ------------------- *
 * (TraversabilityMetrics.cpp:85-88 call site).  Steps: scale by max|a_ij| of the lower
 * triangle; 3x3 tridiagonalisation (Tridiagonalization.h, fixed-size-3 selector);
 * implicit symmetric QR with Wilkinson shift (tridiagonal_qr_step) with the 3.4 deflation
 * test; ascending selection sort of eigenvalues with eigenvector columns.               */
struct Eig3
{
    double w[3];
    double v[3][3]; /* v[row][col]; column k is the eigenvector of w[k] */
};

inline void makeGivens(double p, double q, double &c, double &s)
{
    if (q == 0.0) { c = p < 0.0 ? -1.0 : 1.0; s = 0.0; }
    else if (p == 0.0) { c = 0.0; s = q < 0.0 ? 1.0 : -1.0; }
    else if (std::abs(p) > std::abs(q))
    {
        const double t = q / p;
        double u = std::sqrt(1.0 + t * t);
        if (p < 0.0) u = -u;
        c = 1.0 / u;
        s = -t * c;
    }
    else
    {
        const double t = p / q;
        double u = std::sqrt(1.0 + t * t);
        if (q < 0.0) u = -u;
        s = -1.0 / u;
        c = -t * s;
    }
}

Eig3 eigh3(const double A[3][3])
{
    Eig3 r;
    double m[3][3];
    double scale = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j <= i; ++j)
            scale = std::max(scale, std::abs(A[i][j]));
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j <= i; ++j)
            m[i][j] = A[i][j] / scale;

    double diag[3], sub[2];
    double Q[3][3];
    const double tol = std::numeric_limits<double>::min();
    diag[0] = m[0][0];
    const double v1norm2 = m[2][0] * m[2][0];
    if (v1norm2 <= tol)
    {
        diag[1] = m[1][1];
        diag[2] = m[2][2];
        sub[0] = m[1][0];
        sub[1] = m[2][1];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Q[i][j] = i == j ? 1.0 : 0.0;
    }
    else
    {
        const double beta = std::sqrt(m[1][0] * m[1][0] + v1norm2);
        const double invBeta = 1.0 / beta;
        const double m01 = m[1][0] * invBeta;
        const double m02 = m[2][0] * invBeta;
        const double q = 2.0 * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
        diag[1] = m[1][1] + m02 * q;
        diag[2] = m[2][2] - m02 * q;
        sub[0] = beta;
        sub[1] = m[2][1] - m01 * q;
        const double Qi[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
        std::memcpy(Q, Qi, sizeof(Q));
    }

    const int n = 3;
    int end = n - 1, start = 0, iter = 0;
    const int maxIter = 30;
    const double considerAsZero = std::numeric_limits<double>::min();
    const double precision_inv = 1.0 / std::numeric_limits<double>::epsilon();
    while (end > 0)
    {
        for (int i = start; i < end; ++i)
        {
            if (std::abs(sub[i]) < considerAsZero)
                sub[i] = 0.0;
            else
            {
                const double scaled = precision_inv * sub[i];
                if (scaled * scaled <= (std::abs(diag[i]) + std::abs(diag[i + 1])))
                    sub[i] = 0.0;
            }
        }
        while (end > 0 && sub[end - 1] == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > maxIter * n) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0) start--;

        /* tridiagonal_qr_step */
        const double td = (diag[end - 1] - diag[end]) * 0.5;
        const double e = sub[end - 1];
        double mu = diag[end];
        if (td == 0.0)
            mu -= std::abs(e);
        else if (e != 0.0)
        {
            const double e2 = e * e;
            const double h = std::hypot(td, e);
            if (e2 == 0.0)
                mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
            else
                mu -= e2 / (td + (td > 0.0 ? h : -h));
        }
        double x = diag[start] - mu;
        double z = sub[start];
        for (int k = start; k < end && z != 0.0; ++k)
        {
            double c, s;
            makeGivens(x, z, c, s);
            const double sdk = s * diag[k] + c * sub[k];
            const double dkp1 = s * sub[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            sub[k] = c * sdk - s * dkp1;
            if (k > start)
                sub[k - 1] = c * sub[k - 1] - s * z;
            x = sub[k];
            if (k < end - 1)
            {
                z = -s * sub[k + 1];
                sub[k + 1] = c * sub[k + 1];
            }
            for (int i = 0; i < 3; ++i) /* Q = Q * G on columns k, k+1 */
            {
                const double xi = Q[i][k], yi = Q[i][k + 1];
                Q[i][k] = c * xi - s * yi;
                Q[i][k + 1] = s * xi + c * yi;
            }
        }
    }
    for (int i = 0; i < n - 1; ++i) /* ascending selection sort */
    {
        int k = 0;
        double mn = diag[i];
        for (int j = 1; j < n - i; ++j)
            if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
        if (k > 0)
        {
            std::swap(diag[i], diag[k + i]);
            for (int rr = 0; rr < 3; ++rr) std::swap(Q[rr][i], Q[rr][k + i]);
        }
    }
    for (int i = 0; i < 3; ++i)
    {
        r.w[i] = diag[i] * scale;
        for (int j = 0; j < 3; ++j) r.v[i][j] = Q[i][j];
    }
    return r;
}

/* ---- TraversabilityMetrics.cpp:23-124 ------------------------------------------------ */
enum { HAZ_OVERALL = 0, HAZ_BORDER, HAZ_ELEVATION, HAZ_SLOPE, HAZ_STEP, HAZ_ROUGHNESS,
       HAZ_NORMAL_X, HAZ_NORMAL_Y, HAZ_NORMAL_Z, HAZ_COUNT };

struct CellMoment
{
    Moments data;
    double c[3] = {0, 0, 0};
};

std::array<double, HAZ_COUNT> computeGoodness(const CellMoment &query,
                                              const std::vector<CellMoment> &occupied,
                                              int vicinity_cell_count, double ground_clearance,
                                              double max_slope, unsigned int min_points)
{
    std::array<double, HAZ_COUNT> haz;
    haz.fill(kNaN);
    if (query.data.N < 1)
        return haz;
    haz[HAZ_BORDER] = std::min(static_cast<double>(query.data.N) /
                                   static_cast<double>(std::max(1u, min_points)),
                               1.0);
    bool border = query.data.N < min_points ||
                  static_cast<int>(occupied.size()) < vicinity_cell_count;
    for (const auto &c : occupied)
        if (c.data.N < min_points) { border = true; break; }
    if (border)
        return haz;

    haz[HAZ_ELEVATION] = query.data.sz / static_cast<double>(query.data.N) + query.c[2];

    Moments fused;
    std::vector<std::array<double, 3>> bary;
    bary.reserve(occupied.size());
    for (const auto &c : occupied)
    {
        Moments m = c.data;
        m.shift(c.c[0] - query.c[0], c.c[1] - query.c[1], c.c[2] - query.c[2]);
        const double n = static_cast<double>(m.N);
        bary.push_back({m.sx / n, m.sy / n, m.sz / n});
        fused.fuse(m);
    }
    const double nf = static_cast<double>(fused.N);
    const double mu[3] = {fused.sx / nf, fused.sy / nf, fused.sz / nf};
    const double Qf[3][3] = {{fused.sx2, fused.sxy, fused.sxz},
                             {fused.sxy, fused.sy2, fused.syz},
                             {fused.sxz, fused.syz, fused.sz2}};
    double cov[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            cov[i][j] = Qf[i][j] / nf - mu[i] * mu[j];

    const Eig3 es = eigh3(cov);
    double nrm[3] = {es.v[0][0], es.v[1][0], es.v[2][0]};
    const double lambda_min = es.w[0];
    if (nrm[2] < 0.0)
        for (double &v : nrm) v = -v;
    const double z2 = (nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2];
    if (z2 > 0.0)
    {
        const double l = std::sqrt(z2);
        for (double &v : nrm) v /= l;
    }
    haz[HAZ_NORMAL_X] = nrm[0];
    haz[HAZ_NORMAL_Y] = nrm[1];
    haz[HAZ_NORMAL_Z] = nrm[2];

    const double slope = std::acos(std::min(1.0, std::abs(nrm[2])));
    haz[HAZ_SLOPE] = std::min(slope / max_slope, 1.0);
    const double roughness = std::sqrt(std::max(0.0, lambda_min));
    haz[HAZ_ROUGHNESS] = std::min(roughness / ground_clearance, 1.0);

    double min_d = std::numeric_limits<double>::infinity();
    double max_d = -std::numeric_limits<double>::infinity();
    for (const auto &b : bary)
    {
        const double d = ((b[0] - mu[0]) * nrm[0] + (b[1] - mu[1]) * nrm[1]) + (b[2] - mu[2]) * nrm[2];
        min_d = std::min(min_d, d);
        max_d = std::max(max_d, d);
    }
    const double step = max_d - min_d;
    haz[HAZ_STEP] = std::min(step / ground_clearance, 1.0);
    haz[HAZ_OVERALL] = std::max(haz[HAZ_SLOPE], std::max(haz[HAZ_STEP], haz[HAZ_ROUGHNESS]));
    return haz;
}

/* ---- grid (Helpers.cpp:35-136) -------------------------------------------------------- *
 * grid_map storage restated as dense float planes addressed by absolute cell index; the grid
 * is origin-pinned with 2k+1 cells per axis, so a lattice cell is inside iff |ci|<=kx,
 * |cj|<=ky (grid_map::GridMap::isInside/getIndex on a cell centre).                         */
struct Grid
{
    double res = 0;
    int kx = 0, ky = 0, W = 0, H = 0;
    std::vector<float> L[TMO_NUM_LAYERS];

    void make(double res_, double halfX, double halfY)  /* Helpers.cpp:35-51 */
    {
        res = res_;
        kx = static_cast<int>(std::ceil(halfX / res));
        ky = static_cast<int>(std::ceil(halfY / res));
        W = 2 * kx + 1;
        H = 2 * ky + 1;
        for (auto &l : L) l.assign(static_cast<size_t>(W) * H, kNaNf);
    }
    bool inside(int ci, int cj) const { return ci >= -kx && ci <= kx && cj >= -ky && cj <= ky; }
    size_t at(int ci, int cj) const { return static_cast<size_t>(cj + ky) * W + (ci + kx); }
    double lengthX() const { return W * res; }
    double lengthY() const { return H * res; }
};

struct KeyFrame
{
    std::uint64_t id = 0;
    std::vector<std::array<float, 3>> cloud;  /* pruned, base frame */
    float pose[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    bool pending = false;
    std::map<std::uint64_t, Moments> partials;  /* reference: unordered_map; order is irrelevant */
    bool hasCloud() const { return !cloud.empty(); }
};

/* KeyFrame.cpp:52-67 */
void rebin(const Lattice &lat, const std::vector<std::array<float, 3>> &cloud, const float pose[12],
           std::map<std::uint64_t, Moments> &partials)
{
    partials.clear();
    for (const auto &pb : cloud)
    {
        float pm[3];
        xform(pose, pb[0], pb[1], pb[2], pm);
        int ci, cj;
        lat.cellOf(pm[0], pm[1], ci, cj);
        const double cx = lat.cx(ci), cy = lat.cy(cj);
        Moments &m = partials[Lattice::key(ci, cj)];
        m.insert(static_cast<double>(pm[0]) - cx, static_cast<double>(pm[1]) - cy,
                 static_cast<double>(pm[2]));
        /* extension: z extremes of the float map-frame z */
        m.zmin = std::isnan(m.zmin) ? pm[2] : std::min(m.zmin, pm[2]);
        m.zmax = std::isnan(m.zmax) ? pm[2] : std::max(m.zmax, pm[2]);
    }
}
}  // namespace tmo_detail
using namespace tmo_detail;

struct tmo_map
{
    tmo_params P;
    Lattice lat;
    Grid grid;
    std::vector<std::pair<int, int>> disc, inflOffsets;
    struct Tap { int di, dj; float decay; };
    std::vector<Tap> inflKernel;
    float Tbv[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    std::map<std::uint64_t, std::shared_ptr<KeyFrame>> kfs;
    std::deque<std::shared_ptr<KeyFrame>> queue;
    std::set<std::uint64_t> queued;
    std::set<std::uint64_t> changed;

    explicit tmo_map(const tmo_params &p) : P(p)
    {
        /* System.cpp:27-30, LocalMap.cpp:41-74 */
        lat.x0 = p.grid_center_x;
        lat.y0 = p.grid_center_y;
        lat.res = p.resolution;
        disc = discOffsets(p.robot_radius, p.resolution);
        if (p.inflation_enabled)
        {
            inflOffsets = discOffsets(p.inflation_radius, p.resolution);
            for (const auto &o : inflOffsets)
            {
                const double dist_m = std::hypot(static_cast<double>(o.first), static_cast<double>(o.second)) * p.resolution;
                inflKernel.push_back({o.first, o.second,
                                      static_cast<float>(std::exp(-p.cost_scaling_factor * dist_m))});
            }
        }
        grid.make(p.resolution, p.half_size, p.half_size);
    }

    /* Helpers.cpp:53-84 + LocalMap.cpp:91-106 */
    void growToIncludeCells(const std::map<std::uint64_t, Moments> &partials)
    {
        if (partials.empty())
            return;
        double minx = 1e18, maxx = -1e18, miny = 1e18, maxy = -1e18;
        for (const auto &kv : partials)
        {
            int ci, cj;
            Lattice::unkey(kv.first, ci, cj);
            const double cx = lat.cx(ci), cy = lat.cy(cj);
            minx = std::min(minx, cx); maxx = std::max(maxx, cx);
            miny = std::min(miny, cy); maxy = std::max(maxy, cy);
        }
        const double margin = grid.res;
        const double curHalfX = grid.lengthX() / 2.0, curHalfY = grid.lengthY() / 2.0;
        const double needX = std::max(std::abs(maxx - lat.x0), std::abs(minx - lat.x0)) + margin;
        const double needY = std::max(std::abs(maxy - lat.y0), std::abs(miny - lat.y0)) + margin;
        if (needX <= curHalfX && needY <= curHalfY)
            return;
        double newHalfX = curHalfX, newHalfY = curHalfY;
        while (newHalfX < needX) newHalfX += P.extend_length;
        while (newHalfY < needY) newHalfY += P.extend_length;
        Grid old = std::move(grid);
        grid = Grid();
        grid.make(old.res, newHalfX, newHalfY);
        for (int cj = -old.ky; cj <= old.ky; ++cj)
            for (int ci = -old.kx; ci <= old.kx; ++ci)
            {
                if (!grid.inside(ci, cj))
                    continue;
                const size_t a = old.at(ci, cj), b = grid.at(ci, cj);
                for (int l = 0; l < TMO_NUM_LAYERS; ++l)
                    if (!std::isnan(old.L[l][a]))
                        grid.L[l][b] = old.L[l][a];
            }
    }

    /* Helpers.cpp:93-112 */
    bool readCell(int ci, int cj, Moments &out) const
    {
        if (!grid.inside(ci, cj))
            return false;
        const size_t a = grid.at(ci, cj);
        const float n = grid.L[TMO_L_N][a];
        if (std::isnan(n) || n < 1.f)
            return false;
        out.N = static_cast<unsigned int>(std::lround(n));
        out.sx = grid.L[TMO_L_SX][a];   out.sy = grid.L[TMO_L_SY][a];   out.sz = grid.L[TMO_L_SZ][a];
        out.sx2 = grid.L[TMO_L_SX2][a]; out.sy2 = grid.L[TMO_L_SY2][a]; out.sz2 = grid.L[TMO_L_SZ2][a];
        out.sxy = grid.L[TMO_L_SXY][a]; out.sxz = grid.L[TMO_L_SXZ][a]; out.syz = grid.L[TMO_L_SYZ][a];
        return true;
    }

    /* LocalMap.cpp:110-126 + Helpers.cpp:114-127 */
    void addPartialToGrid(std::uint64_t id, const Moments &m, double sign)
    {
        int ci, cj;
        Lattice::unkey(id, ci, cj);
        if (!grid.inside(ci, cj))
            return;
        const size_t a = grid.at(ci, cj);
        const auto acc = [a](std::vector<float> &layer, double v)
        {
            float &cell = layer[a];
            if (std::isnan(cell)) cell = 0.f;
            cell += static_cast<float>(v);
        };
        acc(grid.L[TMO_L_N], sign * m.N);
        acc(grid.L[TMO_L_SX], sign * m.sx);   acc(grid.L[TMO_L_SY], sign * m.sy);   acc(grid.L[TMO_L_SZ], sign * m.sz);
        acc(grid.L[TMO_L_SX2], sign * m.sx2); acc(grid.L[TMO_L_SY2], sign * m.sy2); acc(grid.L[TMO_L_SZ2], sign * m.sz2);
        acc(grid.L[TMO_L_SXY], sign * m.sxy); acc(grid.L[TMO_L_SXZ], sign * m.sxz); acc(grid.L[TMO_L_SYZ], sign * m.syz);
        if (sign > 0)
        {
            /* extension layers: running extremes since the cell was last blank; a subtraction
             * cannot undo an extreme, so they are left alone until the cell blanks. */
            float &zn = grid.L[TMO_L_ZMIN][a], &zx = grid.L[TMO_L_ZMAX][a];
            zn = std::isnan(zn) ? m.zmin : std::min(zn, m.zmin);
            zx = std::isnan(zx) ? m.zmax : std::max(zx, m.zmax);
        }
        if (sign < 0 && std::lround(grid.L[TMO_L_N][a]) <= 0)
        {
            for (auto &l : grid.L) l[a] = kNaNf;  /* blankCell, Helpers.cpp:86-91 */
            changed.insert(id);
        }
    }

    /* LocalMap.cpp:130-189 */
    bool recomputeCell(std::uint64_t id)
    {
        int ci, cj;
        Lattice::unkey(id, ci, cj);
        if (!grid.inside(ci, cj))
            return false;
        CellMoment query;
        if (!readCell(ci, cj, query.data))
            return false;
        query.c[0] = lat.cx(ci); query.c[1] = lat.cy(cj); query.c[2] = 0.0;
        std::vector<CellMoment> occupied;
        occupied.reserve(disc.size());
        for (const auto &o : disc)
        {
            const int i = ci + o.first, j = cj + o.second;
            CellMoment c;
            if (readCell(i, j, c.data))
            {
                c.c[0] = lat.cx(i); c.c[1] = lat.cy(j); c.c[2] = 0.0;
                occupied.push_back(c);
            }
        }
        const auto haz = computeGoodness(query, occupied, static_cast<int>(disc.size()),
                                         P.ground_clearance, P.max_slope,
                                         static_cast<unsigned int>(P.min_points_per_grid));
        const size_t a = grid.at(ci, cj);
        if (!std::isnan(haz[HAZ_ELEVATION]))
            grid.L[TMO_L_ELEVATION][a] = static_cast<float>(haz[HAZ_ELEVATION]);
        if (std::isnan(haz[HAZ_OVERALL]))
        {
            grid.L[TMO_L_HAZARD][a] = kNaNf;
            grid.L[TMO_L_SLOPE][a] = kNaNf;
            grid.L[TMO_L_STEP][a] = kNaNf;
            grid.L[TMO_L_ROUGHNESS][a] = kNaNf;
            grid.L[TMO_L_BORDER][a] = static_cast<float>(haz[HAZ_BORDER]);
            grid.L[TMO_L_NORMAL_X][a] = kNaNf;
            grid.L[TMO_L_NORMAL_Y][a] = kNaNf;
            grid.L[TMO_L_NORMAL_Z][a] = kNaNf;
            return true;
        }
        grid.L[TMO_L_HAZARD][a] = static_cast<float>(haz[HAZ_OVERALL]);
        grid.L[TMO_L_SLOPE][a] = static_cast<float>(haz[HAZ_SLOPE]);
        grid.L[TMO_L_STEP][a] = static_cast<float>(haz[HAZ_STEP]);
        grid.L[TMO_L_ROUGHNESS][a] = static_cast<float>(haz[HAZ_ROUGHNESS]);
        grid.L[TMO_L_BORDER][a] = static_cast<float>(haz[HAZ_BORDER]);
        grid.L[TMO_L_NORMAL_X][a] = static_cast<float>(haz[HAZ_NORMAL_X]);
        grid.L[TMO_L_NORMAL_Y][a] = static_cast<float>(haz[HAZ_NORMAL_Y]);
        grid.L[TMO_L_NORMAL_Z][a] = static_cast<float>(haz[HAZ_NORMAL_Z]);
        return true;
    }

    /* Helpers.cpp:171-184 */
    static std::unordered_set<std::uint64_t> dilate(const std::unordered_set<std::uint64_t> &touched,
                                                    const std::vector<std::pair<int, int>> &offsets)
    {
        std::unordered_set<std::uint64_t> out;
        out.reserve(touched.size() * offsets.size());
        for (auto id : touched)
        {
            int ci, cj;
            Lattice::unkey(id, ci, cj);
            for (const auto &o : offsets)
                out.insert(Lattice::key(ci + o.first, cj + o.second));
        }
        return out;
    }

    /* LocalMap.cpp:201-241 */
    void inflateDirty(const std::unordered_set<std::uint64_t> &dirty)
    {
        const auto window = dilate(dirty, inflOffsets);
        const float lethal = static_cast<float>(P.lethal_step_threshold);
        for (auto id : window)
        {
            int ci, cj;
            Lattice::unkey(id, ci, cj);
            if (!grid.inside(ci, cj))
                continue;
            const size_t a = grid.at(ci, cj);
            if (!(grid.L[TMO_L_N][a] >= 1.f))
                continue;
            float best = 0.f;
            for (const auto &k : inflKernel)
            {
                if (!grid.inside(ci + k.di, cj + k.dj))
                    continue;
                const float sv = grid.L[TMO_L_STEP][grid.at(ci + k.di, cj + k.dj)];
                if (sv >= lethal)
                    best = std::max(best, sv * k.decay);
            }
            float &cur = grid.L[TMO_L_STEP_INFLATED][a];
            if (!(cur == best))
            {
                cur = best;
                changed.insert(id);
            }
        }
    }

    /* LocalMap.cpp:191-199.  threads > 1: the archived node's variant -- a parallel loop over the dirty cells, one cell per
     * task, disjoint writes, flags merged serially afterwards
     * (traversability_mapping_ros/archive/local_traversability_monolithic.cpp:349-362).  Same results either way:
     * recomputeCell reads moment layers only and writes its own cell's derived layers. */
    int threads = 1;
    void recomputeDirty(const std::unordered_set<std::uint64_t> &dirty)
    {
        if (threads > 1)
        {
            const std::vector<std::uint64_t> cells(dirty.begin(), dirty.end());
            const long n = static_cast<long>(cells.size());
            std::vector<char> wrote(cells.size(), 0);
            /* `#pragma omp parallel for schedule(dynamic, 64)` restated with std::thread (this image has no libgomp
             * spec file, so -fopenmp does not link): workers draw chunks of 64 cells from a shared counter */
            std::atomic<long> next{0};
            auto work = [&]() {
                for (;;)
                {
                    const long k0 = next.fetch_add(64);
                    if (k0 >= n) break;
                    const long k1 = std::min(n, k0 + 64);
                    for (long k = k0; k < k1; ++k)
                        wrote[k] = recomputeCell(cells[k]) ? 1 : 0;
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < threads; ++t) pool.emplace_back(work);
            work();
            for (auto &t : pool) t.join();
            for (long k = 0; k < n; ++k)
                if (wrote[k])
                    changed.insert(cells[k]);
        }
        else
        {
            for (auto id : dirty)
                if (recomputeCell(id))
                    changed.insert(id);
        }
        if (P.inflation_enabled)
            inflateDirty(dirty);
    }

    /* LocalMap.cpp:245-315, steps 1-3 (fold); returns false when skipped */
    bool foldKeyFrame(const std::shared_ptr<KeyFrame> &kf, const float pose[12],
                      std::unordered_set<std::uint64_t> &touched)
    {
        if (kfs.find(kf->id) == kfs.end())
            return false;
        if (!kf->hasCloud())
            return false;
        for (auto &kv : kf->partials)
        {
            addPartialToGrid(kv.first, kv.second, -1.0);
            touched.insert(kv.first);
        }
        rebin(lat, kf->cloud, pose, kf->partials);
        growToIncludeCells(kf->partials);
        for (auto &kv : kf->partials)
        {
            addPartialToGrid(kv.first, kv.second, +1.0);
            touched.insert(kv.first);
        }
        return true;
    }

    bool recomputeKeyFrame(const std::shared_ptr<KeyFrame> &kf, const float pose[12])
    {
        std::unordered_set<std::uint64_t> touched;
        if (!foldKeyFrame(kf, pose, touched))
            return false;
        recomputeDirty(dilate(touched, disc));
        if (!P.is_kf_optimization_enabled)
        {
            kf->cloud.clear();
            kf->cloud.shrink_to_fit();
        }
        return true;
    }
};

extern "C" {

void tmo_default_params(tmo_params *p)
{
    std::memset(p, 0, sizeof(*p));
    p->resolution = 0.05; p->grid_center_x = 0; p->grid_center_y = 0; p->half_size = 7.5;
    p->extend_length = 30.0; p->robot_radius = 0.20; p->ground_clearance = 0.15; p->max_slope = 0.4;
    p->min_points_per_grid = 3; p->is_kf_optimization_enabled = 1; p->num_local_keyframes = 10;
    p->global_adjustment_sleep = 0; p->inflation_enabled = 0; p->inflation_radius = 1.5;
    p->cost_scaling_factor = 2.0; p->lethal_step_threshold = 0.95; p->robot_height = 3.5;
    p->max_range = 6.0; p->min_range = 1.0; p->use_pointcloud_buffer = 0; p->use_ros_buffer = 0;
    p->publish_rate_hz = 1.0;
}

tmo_map *tmo_create(const tmo_params *p) { return new tmo_map(*p); }
void tmo_destroy(tmo_map *m) { delete m; }

void tmo_set_extrinsics(tmo_map *m, const float T_bv[12]) { std::memcpy(m->Tbv, T_bv, sizeof(m->Tbv)); }
void tmo_set_threads(tmo_map *m, int32_t n) { m->threads = n < 1 ? 1 : n; }

/* System.cpp:93-114 */
size_t tmo_prune(const tmo_map *m, const float *xyz, size_t n, size_t stride, const float T[12],
                 double max_range, double min_range, float *out)
{
    size_t k = 0;
    const float rh = static_cast<float>(m->P.robot_height);
    const float mx = static_cast<float>(max_range), mn = static_cast<float>(min_range);
    for (size_t i = 0; i < n; ++i)
    {
        const float x = xyz[i * stride], y = xyz[i * stride + 1], z = xyz[i * stride + 2];
        if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z))
            continue;
        if (x == 0.f && y == 0.f)
            continue;
        float pb[3];
        xform(T, x, y, z, pb);
        if (pb[2] > rh)
            continue;
        /* Eigen Vector3f::norm(): sqrt(((x^2 + y^2) + z^2)) in float */
        const float range = std::sqrt((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
        if (range > mx || range < mn)
            continue;
        out[3 * k] = pb[0]; out[3 * k + 1] = pb[1]; out[3 * k + 2] = pb[2];
        ++k;
    }
    return k;
}

int tmo_add_keyframe_base(tmo_map *m, uint64_t kf_id, const float *xyz, size_t n)
{
    /* System.cpp:121-151 (duplicate / empty guards) + LocalMap.cpp:319-328 */
    if (m->kfs.count(kf_id))
        return 0;
    auto kf = std::make_shared<KeyFrame>();
    kf->id = kf_id;
    kf->cloud.resize(n);
    for (size_t i = 0; i < n; ++i)
        kf->cloud[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    m->kfs[kf_id] = kf;
    return 1;
}

int tmo_add_keyframe(tmo_map *m, uint64_t kf_id, const float *xyz, size_t n, size_t stride)
{
    if (kf_id > static_cast<uint64_t>(std::numeric_limits<int64_t>::max()))
        return -1;  /* System.cpp:163-168 throws */
    if (m->kfs.count(kf_id))
        return 0;
    std::vector<float> base(n * 3 + 3);
    const size_t k = tmo_prune(m, xyz, n, stride, m->Tbv, m->P.max_range, m->P.min_range, base.data());
    if (k == 0)
        return 0;  /* System.cpp:137-141 */
    return tmo_add_keyframe_base(m, kf_id, base.data(), k);
}

int tmo_update_pose(tmo_map *m, uint64_t kf_id, const float T_mb[12])
{
    auto it = m->kfs.find(kf_id);
    if (it == m->kfs.end())
        return 0;
    std::memcpy(it->second->pose, T_mb, sizeof(float) * 12);  /* KeyFrame.cpp:29-34 */
    it->second->pending = true;
    if (m->queued.insert(kf_id).second)  /* LocalMap.cpp:336-338 */
        m->queue.push_back(it->second);
    return 1;
}

int tmo_process(tmo_map *m)
{
    int updates = 0;
    while (!m->queue.empty())
    {
        auto kf = m->queue.front();
        m->queue.pop_front();
        m->queued.erase(kf->id);
        if (!kf->pending)  /* KeyFrame.cpp:36-44 */
            continue;
        kf->pending = false;
        float pose[12];
        std::memcpy(pose, kf->pose, sizeof(pose));
        if (m->recomputeKeyFrame(kf, pose))
            ++updates;
    }
    return updates;
}

int tmo_process_batched(tmo_map *m)
{
    int updates = 0;
    std::unordered_set<std::uint64_t> touched;
    std::vector<std::shared_ptr<KeyFrame>> done;
    while (!m->queue.empty())
    {
        auto kf = m->queue.front();
        m->queue.pop_front();
        m->queued.erase(kf->id);
        if (!kf->pending)
            continue;
        kf->pending = false;
        float pose[12];
        std::memcpy(pose, kf->pose, sizeof(pose));
        if (m->foldKeyFrame(kf, pose, touched))
        {
            ++updates;
            done.push_back(kf);
        }
    }
    m->recomputeDirty(tmo_map::dilate(touched, m->disc));
    if (!m->P.is_kf_optimization_enabled)
        for (auto &kf : done) { kf->cloud.clear(); kf->cloud.shrink_to_fit(); }
    return updates;
}

int tmo_delete_keyframe(tmo_map *m, uint64_t kf_id)
{
    auto it = m->kfs.find(kf_id);
    if (it == m->kfs.end())
        return 0;
    auto kf = it->second;
    std::unordered_set<std::uint64_t> touched;
    for (auto &kv : kf->partials)
    {
        m->addPartialToGrid(kv.first, kv.second, -1.0);
        touched.insert(kv.first);
    }
    m->recomputeDirty(tmo_map::dilate(touched, m->disc));
    kf->partials.clear();
    m->kfs.erase(it);
    m->queued.erase(kf_id);
    return 1;
}

void tmo_clear_entire_map(tmo_map *m)
{
    size_t n = tmo_occupied_cell_keys(m, nullptr, 0);
    std::vector<uint64_t> occ(n);
    tmo_occupied_cell_keys(m, occ.data(), n);
    for (auto k : occ) m->changed.insert(k);
    for (auto &l : m->grid.L) std::fill(l.begin(), l.end(), kNaNf);
    for (auto &kv : m->kfs)
    {
        kv.second->partials.clear();
        kv.second->pending = true;
        if (m->queued.insert(kv.first).second)
            m->queue.push_back(kv.second);
    }
}

size_t tmo_num_keyframes(const tmo_map *m) { return m->kfs.size(); }
size_t tmo_queue_depth(const tmo_map *m) { return m->queue.size(); }

size_t tmo_take_changed_cells(tmo_map *m, uint64_t *out, size_t cap)
{
    const size_t n = m->changed.size();
    if (!out)
        return n;
    size_t i = 0;
    for (auto k : m->changed)
        if (i < cap) out[i++] = k;
    m->changed.clear();
    return n;
}

size_t tmo_occupied_cell_keys(const tmo_map *m, uint64_t *out, size_t cap)
{
    /* Helpers.cpp:138-154; returned ascending by key here (reference: grid iteration order) */
    std::vector<uint64_t> keys;
    const Grid &g = m->grid;
    for (int cj = -g.ky; cj <= g.ky; ++cj)
        for (int ci = -g.kx; ci <= g.kx; ++ci)
        {
            const float n = g.L[TMO_L_N][g.at(ci, cj)];
            if (std::isnan(n) || n < 1.f)
                continue;
            keys.push_back(Lattice::key(ci, cj));
        }
    std::sort(keys.begin(), keys.end());
    if (out)
        for (size_t i = 0; i < keys.size() && i < cap; ++i) out[i] = keys[i];
    return keys.size();
}

void tmo_read_layers(const tmo_map *m, const uint64_t *keys, size_t n, const int32_t *layer_ids,
                     size_t n_layers, float *out)
{
    for (size_t i = 0; i < n; ++i)
    {
        int ci, cj;
        Lattice::unkey(keys[i], ci, cj);
        const bool in = m->grid.inside(ci, cj);
        for (size_t l = 0; l < n_layers; ++l)
            out[i * n_layers + l] = in ? m->grid.L[layer_ids[l]][m->grid.at(ci, cj)] : kNaNf;
    }
}

size_t tmo_fill_sparse_update(const tmo_map *m, const int32_t *layer_ids, size_t n_layers, const uint64_t *keys, size_t n,
                              uint64_t *out_keys, float *out_values)
{
    /* tmap_ros::fillSparseUpdate (traversability_mapping_ros/src/conversions.cpp:28-55): keys whose cell centre is
     * outside the grid are skipped; the rest keep their order, values[i * n_layers + l]. */
    size_t k = 0;
    for (size_t i = 0; i < n; ++i)
    {
        int ci, cj;
        Lattice::unkey(keys[i], ci, cj);
        if (!m->grid.inside(ci, cj))
            continue;
        out_keys[k] = keys[i];
        for (size_t l = 0; l < n_layers; ++l)
            out_values[k * n_layers + l] = m->grid.L[layer_ids[l]][m->grid.at(ci, cj)];
        ++k;
    }
    return k;
}

void tmo_grid_extent(const tmo_map *m, int32_t *kx, int32_t *ky) { *kx = m->grid.kx; *ky = m->grid.ky; }

void tmo_export_dense(const tmo_map *m, int32_t layer, int32_t ci0, int32_t cj0, int32_t w, int32_t h, float *out)
{
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i)
        {
            const int ci = ci0 + i, cj = cj0 + j;
            out[static_cast<size_t>(j) * w + i] = m->grid.inside(ci, cj) ? m->grid.L[layer][m->grid.at(ci, cj)] : kNaNf;
        }
}

void tmo_export_occupancy(const tmo_map *m, int32_t ci0, int32_t cj0, int32_t w, int32_t h, int8_t *out)
{
    /* grid_map_ros GridMapRosConverter::toOccupancyGrid(grid, "hazard", 0.0, 1.0, og), call
     * sites global_traversability_node.cpp:374, local_traversability_node.cpp:236:
     * v = (x - min)/(max - min); NaN -> -1; else 0 + clamp(v,0,1)*100 assigned to int8. */
    const float dataMin = 0.f, dataMax = 1.f, cellMin = 0.f, cellRange = 100.f;
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i)
        {
            const int ci = ci0 + i, cj = cj0 + j;
            float value = m->grid.inside(ci, cj) ? m->grid.L[TMO_L_HAZARD][m->grid.at(ci, cj)] : kNaNf;
            value = (value - dataMin) / (dataMax - dataMin);
            if (std::isnan(value))
                value = -1.f;
            else
                value = cellMin + std::min(std::max(0.0f, value), 1.0f) * cellRange;
            out[static_cast<size_t>(j) * w + i] = static_cast<int8_t>(value);
        }
}

void tmo_export_completeness(const tmo_map *m, int32_t ci0, int32_t cj0, int32_t w, int32_t h, double publish_res,
                             int8_t *out, int32_t *out_w, int32_t *out_h)
{
    /* GlobalTraversabilityNode::publishCompleteness (traversability_mapping_ros/src/global_traversability_node.cpp:381-459):
     * toOccupancyGrid(grid, "border_haz", 0, 1) -> min-pool to publish_res (unobserved fine cells skipped) ->
     * RViz costmap-palette bytes.  nav_msgs MapMetaData::resolution is float32. */
    const float dataMin = 0.f, dataMax = 1.f, cellMin = 0.f, cellRange = 100.f;
    std::vector<int8_t> fine(static_cast<size_t>(w) * h);
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i)
        {
            const int ci = ci0 + i, cj = cj0 + j;
            float value = m->grid.inside(ci, cj) ? m->grid.L[TMO_L_BORDER][m->grid.at(ci, cj)] : kNaNf;
            value = (value - dataMin) / (dataMax - dataMin);
            if (std::isnan(value))
                value = -1.f;
            else
                value = cellMin + std::min(std::max(0.0f, value), 1.0f) * cellRange;
            fine[static_cast<size_t>(j) * w + i] = static_cast<int8_t>(value);
        }
    const float fine_res = static_cast<float>(m->P.resolution);
    int f = static_cast<int>(std::lround(publish_res / fine_res));
    if (f < 1)
        f = 1;
    const unsigned fw = w, fh = h;
    const unsigned cw = (fw + f - 1) / f, ch = (fh + f - 1) / f;
    *out_w = static_cast<int32_t>(cw);
    *out_h = static_cast<int32_t>(ch);
    for (unsigned cy = 0; cy < ch; ++cy)
        for (unsigned cx = 0; cx < cw; ++cx)
        {
            int best = 101;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx)
                {
                    const unsigned fx = cx * f + dx, fy = cy * f + dy;
                    if (fx >= fw || fy >= fh)
                        continue;
                    const int v = fine[static_cast<size_t>(fy) * fw + fx];
                    if (v >= 0 && v < best)
                        best = v;
                }
            int8_t d = (best <= 100) ? static_cast<int8_t>(best) : static_cast<int8_t>(-1);
            if (d >= 0)
            {
                const double c = static_cast<double>(d) / 100.0;
                const int byte = (d >= 100) ? 113 : 128 + static_cast<int>(std::lround(c * (254 - 128)));
                d = static_cast<int8_t>(byte);
            }
            out[static_cast<size_t>(cy) * cw + cx] = d;
        }
}

size_t tmo_voxel_grid(const float *xyz, size_t n, float lx, float ly, float lz, float *out, size_t cap);

size_t tmo_stitched_cloud(const tmo_map *m, float vx, float vy, float vz, float *out, size_t cap)
{
    /* LocalMap::getStitchedPointCloud (LocalMap.cpp:416-448): pose * retained cloud of every keyframe (ascending id
     * here; the reference walks an unordered_map), then pcl::VoxelGrid when all three leaf sizes are positive.
     * out == NULL: the unfiltered size. */
    std::vector<float> all;
    for (const auto &kv : m->kfs)
    {
        const KeyFrame &kf = *kv.second;
        if (!kf.hasCloud())
            continue;
        for (const auto &pb : kf.cloud)
        {
            float pm[3];
            xform(kf.pose, pb[0], pb[1], pb[2], pm);
            all.insert(all.end(), pm, pm + 3);
        }
    }
    const size_t n = all.size() / 3;
    if (!out)
        return n;
    if (vx > 0.f && vy > 0.f && vz > 0.f && n)
        return tmo_voxel_grid(all.data(), n, vx, vy, vz, out, cap);
    for (size_t i = 0; i < std::min(n, cap) * 3; ++i) out[i] = all[i];
    return n;
}

size_t tmo_voxel_grid(const float *xyz, size_t n, float lx, float ly, float lz, float *out, size_t cap)
{
    /* pcl::VoxelGrid<pcl::PointXYZ>::applyFilter as LocalMap::getStitchedPointCloud uses it (LocalMap.cpp:438-446;
     * PCL 1.12 filters/impl/voxel_grid.hpp:214-426 restated, PCL itself is not in /root/reference): dense cloud, no
     * filter field, min_points_per_voxel 0.  PCL's std::sort / spreadsort on the voxel index leaves the order INSIDE
     * a voxel unspecified; this restatement sorts stably, so centroids are PCL's up to float summation order. */
    if (n == 0) return 0;
    const float leaf[3] = {lx, ly, lz};
    float inv[3], mn[3], mx[3];
    for (int k = 0; k < 3; ++k)
    {
        inv[k] = 1.0f / leaf[k];
        mn[k] = std::numeric_limits<float>::max();
        mx[k] = -std::numeric_limits<float>::max();
    }
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
        {
            mn[k] = std::min(mn[k], xyz[3 * i + k]);
            mx[k] = std::max(mx[k], xyz[3 * i + k]);
        }
    const std::int64_t dx = static_cast<std::int64_t>((mx[0] - mn[0]) * inv[0]) + 1;
    const std::int64_t dy = static_cast<std::int64_t>((mx[1] - mn[1]) * inv[1]) + 1;
    const std::int64_t dz = static_cast<std::int64_t>((mx[2] - mn[2]) * inv[2]) + 1;
    if ((dx * dy * dz) > static_cast<std::int64_t>(std::numeric_limits<std::int32_t>::max()))
    {
        for (size_t i = 0; i < std::min(n, cap) * 3; ++i) out[i] = xyz[i];  // "Leaf size is too small": output = input
        return n;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int k = 0; k < 3; ++k)
    {
        min_b[k] = static_cast<int>(std::floor(mn[k] * inv[k]));
        max_b[k] = static_cast<int>(std::floor(mx[k] * inv[k]));
        div_b[k] = max_b[k] - min_b[k] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    std::vector<std::pair<unsigned int, unsigned int>> index_vector;
    index_vector.reserve(n);
    for (size_t i = 0; i < n; ++i)
    {
        const int ijk0 = static_cast<int>(std::floor(xyz[3 * i] * inv[0]) - static_cast<float>(min_b[0]));
        const int ijk1 = static_cast<int>(std::floor(xyz[3 * i + 1] * inv[1]) - static_cast<float>(min_b[1]));
        const int ijk2 = static_cast<int>(std::floor(xyz[3 * i + 2] * inv[2]) - static_cast<float>(min_b[2]));
        const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
        index_vector.emplace_back(static_cast<unsigned int>(idx), static_cast<unsigned int>(i));
    }
    std::stable_sort(index_vector.begin(), index_vector.end(),
                     [](const std::pair<unsigned int, unsigned int> &a, const std::pair<unsigned int, unsigned int> &b) { return a.first < b.first; });
    size_t nv = 0;
    for (size_t i = 0; i < n;)
    {
        size_t j = i;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        while (j < n && index_vector[j].first == index_vector[i].first)
        {
            const size_t p = 3 * static_cast<size_t>(index_vector[j].second);
            sx += xyz[p]; sy += xyz[p + 1]; sz += xyz[p + 2];
            ++j;
        }
        const float c = static_cast<float>(j - i);
        if (nv < cap) { out[3 * nv] = sx / c; out[3 * nv + 1] = sy / c; out[3 * nv + 2] = sz / c; }
        ++nv;
        i = j;
    }
    return nv;
}

void tmo_cell_of(double x0, double y0, double res, double x, double y, int32_t *ci, int32_t *cj)
{
    Lattice l; l.x0 = x0; l.y0 = y0; l.res = res;
    int a, b;
    l.cellOf(x, y, a, b);
    *ci = a; *cj = b;
}

uint64_t tmo_key(int32_t ci, int32_t cj) { return Lattice::key(ci, cj); }

size_t tmo_disc_offsets(double radius, double res, int32_t *out, size_t cap)
{
    const auto d = discOffsets(radius, res);
    for (size_t i = 0; i < d.size() && i < cap; ++i)
    {
        out[2 * i] = d[i].first;
        out[2 * i + 1] = d[i].second;
    }
    return d.size();
}

size_t tmo_rebin(double x0, double y0, double res, const float *xyz, size_t n, const float T[12],
                 uint64_t *out_keys, double *out_m)
{
    Lattice l; l.x0 = x0; l.y0 = y0; l.res = res;
    std::vector<std::array<float, 3>> cloud(n);
    for (size_t i = 0; i < n; ++i) cloud[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    std::map<std::uint64_t, Moments> partials;
    rebin(l, cloud, T, partials);
    size_t k = 0;
    for (const auto &kv : partials)
    {
        out_keys[k] = kv.first;
        const Moments &m = kv.second;
        const double v[10] = {static_cast<double>(m.N), m.sx, m.sy, m.sz, m.sx2, m.sy2, m.sz2, m.sxy, m.sxz, m.syz};
        std::memcpy(out_m + 10 * k, v, sizeof(v));
        ++k;
    }
    return k;
}

static CellMoment unpackCell(const double *c)
{
    CellMoment cm;
    cm.data.N = static_cast<unsigned int>(c[0]);
    cm.data.sx = c[1]; cm.data.sy = c[2]; cm.data.sz = c[3];
    cm.data.sx2 = c[4]; cm.data.sy2 = c[5]; cm.data.sz2 = c[6];
    cm.data.sxy = c[7]; cm.data.sxz = c[8]; cm.data.syz = c[9];
    cm.c[0] = c[10]; cm.c[1] = c[11]; cm.c[2] = c[12];
    return cm;
}

void tmo_compute_goodness(const double *query, const double *cells, size_t n_occ, int32_t vicinity,
                          double ground_clearance, double max_slope, uint32_t min_points, double *out9)
{
    const CellMoment q = unpackCell(query);
    std::vector<CellMoment> occ(n_occ);
    for (size_t i = 0; i < n_occ; ++i) occ[i] = unpackCell(cells + 13 * i);
    const auto h = computeGoodness(q, occ, vicinity, ground_clearance, max_slope, min_points);
    for (int i = 0; i < 9; ++i) out9[i] = h[i];
}

void tmo_eigh3(const double a[9], double w[3], double v[9])
{
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            A[i][j] = a[3 * i + j];
    const Eig3 e = eigh3(A);
    for (int k = 0; k < 3; ++k)
    {
        w[k] = e.w[k];
        for (int i = 0; i < 3; ++i) v[3 * k + i] = e.v[i][k];
    }
}

}  // extern "C"
