// kernels_classify.cu -- halo-tile stencil classification, step inflation, exports.
//
// Replaces (reference paths relative to traversability_mapping/):
//   LocalMap::recomputeCell / recomputeDirty   src/LocalMap.cpp:130-199    -> k_gate + k_classify
//   computeGoodness                            src/TraversabilityMetrics.cpp:23-124
//   discOffsets / dilate                       src/Helpers.cpp:156-184     (disc in __constant__,
//                                              dilation evaluated per cell from the touched bits)
//   LocalMap::inflateDirty                     src/LocalMap.cpp:201-241    -> k_inflate
//   takeChangedCells / allOccupiedKeys         src/LocalMap.cpp:401-414, src/Helpers.cpp:138-154
//   fillSparseUpdate                           traversability_mapping_ros/src/conversions.cpp:28-55
//   GridMapRosConverter::toOccupancyGrid       (grid_map_ros; call site global_traversability_node.cpp:374)
//
//   publishCompleteness                        traversability_mapping_ros/src/global_traversability_node.cpp:381-459
//
// k_gate decides, from the flag bytes alone, which cells of every candidate 16x16 bin are dirty and which of
// those pass the footprint gate; k_classify fits the gated-in cells of the bins k_gate listed.  The (16+2r)^2 x
// 48-byte halo tile of moment records arrives with ONE 3-D TMA box load (cp.async.bulk.tensor, mbarrier
// completion, out-of-bounds cells filled with NaN = unobserved) and is widened to fp64 once; each thread then
// classifies one cell: the disc moments are fused about the query centre as 21 lattice-weighted sums (the shift
// of Moments.cpp:49-58 expanded for offsets that are integer multiples of the resolution), followed by the 3x3
// covariance, the smallest eigenpair (eig3.cuh) and a second pass over the disc for the step spread.  All in fp64.
#include "tmb_common.cuh"
#include "eig3.cuh"
#include "tmb_async.cuh"
#include <cuda.h>
#include <math_constants.h>
#include <cstdlib>
#include <mutex>

namespace tmb
{
__constant__ int c_disc_off[MAX_DISC];       // smem cell offset dj*TW + di
__constant__ double2 c_disc_d[MAX_DISC];     // (di, dj) as doubles

enum : uint8_t { H_OBS = 1, H_OK = 2, H_TOUCHED = 4 };
constexpr int CLS_THREADS = 256;
constexpr int FLAG_BOX_W = BIN + 32;  // flag-byte box: starts 16 columns left of the bin (TMA needs a 16-byte aligned start)

// ---------------------------------------------------------------------------------------------
// k_gate: the cheap half of recomputeCell for every candidate bin, from the FLAG BYTES alone (1 B/cell):
// which cells are dirty (in dilate(touched, disc)), which of those pass the footprint gate.  Gated-out dirty
// cells get their outputs here (border value, NaN hazards); gated-in ones are marked F_FIT and their bin is
// listed for the TMA-staged fit kernel.  Flag box via 2-D TMA (double-buffered), row masks via ballots.
// ---------------------------------------------------------------------------------------------
constexpr int GATE_THREADS = 256;
__global__ void __launch_bounds__(GATE_THREADS) k_gate(const __grid_constant__ CUtensorMap tmap_flags, BatchView b, GridView g,
                                                        const ClassifyParams c_cp)
{
    __shared__ __align__(128) uint8_t s_fl[2][(BIN + 16) * FLAG_BOX_W];   // r <= 8 -> at most 32 rows of 48 bytes
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_t[32], s_o[32], s_e[32];
    const int r = c_cp.r, TW = BIN + 2 * r;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_cand = b.counters[1];
    const uint32_t minp = c_cp.min_points;
    const size_t pl = g.plane();
    uint32_t n_done = 0;

    auto issue = [&](int stage, uint32_t it)  // thread 0 only
    {
        const uint32_t gt = b.cand_tiles[it];
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);
        mbar_arrive_expect_tx(&s_bar[stage], (uint32_t)(TW * FLAG_BOX_W));
        tma_load_2d(&s_fl[stage][0], &tmap_flags, &s_bar[stage], bx * BIN - g.ci0 - 16, by * BIN - g.cj0 - r);
    };
    if (tid == 0)
    {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t stride = gridDim.x;
    uint32_t cur = blockIdx.x;
    if (tid == 0 && cur < n_cand)
        issue(0, cur);
    int stage = 0;
    uint32_t parity0 = 0, parity1 = 0;
    while (cur < n_cand)
    {
        const uint32_t nxt = cur + stride;
        __syncthreads();  // previous bin consumed (masks and the other stage's bytes)
        if (tid == 0 && nxt < n_cand)
            issue(stage ^ 1, nxt);
        const uint32_t gt = b.cand_tiles[cur];
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);
        const int col0 = bx * BIN - g.ci0, row0 = by * BIN - g.cj0;
        if (stage == 0) { mbar_wait(&s_bar[0], parity0); parity0 ^= 1; }
        else { mbar_wait(&s_bar[1], parity1); parity1 ^= 1; }
        const uint8_t *fl = &s_fl[stage][0];
        // one warp per halo row: touched / observed / enough bits as 32-bit row masks
        int any_touched = 0;
        for (int hy = warp; hy < TW; hy += GATE_THREADS / 32)
        {
            const uint8_t f = (int)lane < TW ? fl[hy * FLAG_BOX_W + ((int)lane - r + 16)] : (uint8_t)0;
            const uint32_t mt = __ballot_sync(0xffffffffu, (f & F_TOUCHED) != 0);
            const uint32_t mo = __ballot_sync(0xffffffffu, (f & F_OBSERVED) != 0);
            const uint32_t me = __ballot_sync(0xffffffffu, (f & F_ENOUGH) != 0);
            any_touched |= mt != 0;
            if (lane == 0) { s_t[hy] = mt; s_o[hy] = mo; s_e[hy] = me; }
        }
        int any_fit = 0;
        if (__syncthreads_or(any_touched))  // else nothing in this bin's footprint halo changed: no cell is dirty
        {
            const int x = tid & (BIN - 1), y = tid >> BIN_SHIFT;
            // LocalMap.cpp:138-140: an unobserved query cell is skipped, nothing written
            if ((s_o[y + r] >> (x + r)) & 1u)
            {
                // dirty = query in dilate(touched, disc); gate = every disc cell holds >= min_points:
                // per footprint row, a contiguous run of 2w+1 bits of the row masks
                bool dirty = false, gate = true;
                for (int dj = -r; dj <= r; ++dj)
                {
                    const int w = c_cp.row_w[dj + r];
                    if (w < 0) continue;
                    const uint32_t wm = (2u << (2 * w)) - 1u;
                    const int sh = x + r - w;
                    dirty |= ((s_t[y + r + dj] >> sh) & wm) != 0u;
                    gate &= ((s_e[y + r + dj] >> sh) & wm) == wm;
                }
                if (dirty)
                {
                    ++n_done;
                    const size_t cell = (size_t)(row0 + y) * g.W + (col0 + x);
                    if (gate)
                    {
                        g.flags[cell] |= (uint8_t)(F_CHANGED | F_FIT);  // recomputeCell returns true (LocalMap.cpp:194-196)
                        any_fit = 1;
                    }
                    else
                    {
                        g.flags[cell] |= F_CHANGED;
                        float border = 1.f;  // N >= min_points
                        if (!((s_e[y + r] >> (x + r)) & 1u))
                        {
                            const double nq = (double)rintf(g.mom[cell * REC + R_N]);
                            border = (float)fmin(nq / (double)max(1u, minp), 1.0);
                        }
                        g.der[D_HAZARD * pl + cell] = CUDART_NAN_F;
                        g.der[D_SLOPE * pl + cell] = CUDART_NAN_F;
                        g.der[D_STEP * pl + cell] = CUDART_NAN_F;
                        g.der[D_ROUGHNESS * pl + cell] = CUDART_NAN_F;
                        g.der[D_BORDER * pl + cell] = border;
                        g.der[D_NX * pl + cell] = CUDART_NAN_F;
                        g.der[D_NY * pl + cell] = CUDART_NAN_F;
                        g.der[D_NZ * pl + cell] = CUDART_NAN_F;
                    }
                }
            }
        }
        if (__syncthreads_or(any_fit) && tid == 0)
            b.fit_tiles[atomicAdd(b.counters + 12, 1u)] = gt;
        cur = nxt;
        stage ^= 1;
    }
    for (int o = 16; o > 0; o >>= 1)
        n_done += __shfl_xor_sync(0xffffffffu, n_done, o);
    if (lane == 0 && n_done)
        atomicAdd(b.counters + 7, n_done);
}

struct ClsSmem
{
    size_t rec_bytes, flag_bytes, stage_bytes, mom_off, inv_off, total;
    __host__ __device__ explicit ClsSmem(int r)
    {
        const size_t TW = BIN + 2 * r, ncell = TW * TW;
        rec_bytes = ncell * REC * 4;
        flag_bytes = TW * FLAG_BOX_W;
        stage_bytes = ((rec_bytes + 127) & ~(size_t)127) + ((flag_bytes + 127) & ~(size_t)127);
        mom_off = stage_bytes;                 // double [ncell][10]: N, sx, sy, sz, sxx, syy, szz, sxy, sxz, syz
        inv_off = mom_off + ncell * 10 * 8;    // double [ncell]: 1/N of observed cells, else 0
        total = inv_off + ncell * 8;
    }
};

// k_classify: fit of the cells k_gate marked (F_FIT), one 16x16 bin per iteration.
//   1. the bin's halo tile of moment records (3-D TMA box) and flag bytes (2-D box) land in shared memory;
//   2. pre-pass, once per halo cell: the ten float32 moments are widened to fp64 and 1/N is formed
//      (MomentLayers::read, Helpers.cpp:101-112).  The conversion unit is 4x narrower than the FP64 pipe on
//      sm_100 (profiles/ubench/pipes.cu) and every halo cell is read by |disc| query cells, so the widening
//      does not sit in the footprint loops;
//   3. the float staging area is free again: the NEXT bin's boxes are requested and arrive during step 4;
//   4. pass A fuses the footprint about the query centre (fp64 only), covariance + smallest eigenpair,
//      pass B the spread of the barycentres along the normal.
// (A variant that replaced pass A's per-cell multiply-adds by differences of row prefix sums was measured and
//  dropped: it halves the fp64 work but needs 30 shared-memory doubles per footprint row and two more barriers
//  per bin, and came out 10-20 % slower.)
__global__ void __launch_bounds__(CLS_THREADS) k_classify(const __grid_constant__ CUtensorMap tmap,
                                                             const __grid_constant__ CUtensorMap tmap_flags, BatchView b, GridView g,
                                                          const ClassifyParams c_cp)
{
    extern __shared__ __align__(1024) unsigned char s_raw[];
    const int r = c_cp.r;
    const int TW = BIN + 2 * r;
    const int ncell = TW * TW;
    // shared-memory layout (same arithmetic as ClsSmem on the host)
    const uint32_t rec_bytes = (uint32_t)ncell * REC * 4, flag_bytes = (uint32_t)TW * FLAG_BOX_W;
    const uint32_t rec_pad = (rec_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = rec_pad + ((flag_bytes + 127u) & ~127u);
    const float *s_rec = reinterpret_cast<const float *>(s_raw);
    const uint8_t *s_fraw = s_raw + rec_pad;
    double *s_mom = reinterpret_cast<double *>(s_raw + stage_bytes);                          // [ncell][10]
    double *s_inv = reinterpret_cast<double *>(s_raw + stage_bytes + (size_t)ncell * 80);     // [ncell]
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t tid = threadIdx.x;
    const uint32_t n_cand = b.counters[12];  // bins with at least one cell to fit (listed by k_gate)
    const size_t pl = g.plane();
    const double rs = c_cp.res;

    auto issue = [&](uint32_t it)  // thread 0 only
    {
        const uint32_t gt = b.fit_tiles[it];
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);
        const int col0 = bx * BIN - g.ci0, row0 = by * BIN - g.cj0;
        mbar_arrive_expect_tx(&s_bar, rec_bytes + flag_bytes);
        tma_load_3d(s_raw, &tmap, &s_bar, 0, col0 - r, row0 - r);
        tma_load_2d(s_raw + rec_pad, &tmap_flags, &s_bar, col0 - 16, row0 - r);
    };

    if (tid == 0)
    {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // static round-robin over the list (its order is arbitrary): no ticket atomics in the loop
    const uint32_t stride = gridDim.x;
    uint32_t cur = blockIdx.x;
    if (tid == 0 && cur < n_cand)
        issue(cur);
    uint32_t parity = 0;
    const int x = tid & (BIN - 1), y = tid >> BIN_SHIFT;
    const int idx = (y + r) * TW + (x + r);
    const int nd = c_cp.n_disc;

    while (cur < n_cand)
    {
        const uint32_t nxt = cur + stride;
        const uint32_t gt = b.fit_tiles[cur];
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);  // global bin
        const int col0 = bx * BIN - g.ci0, row0 = by * BIN - g.cj0;
        mbar_wait(&s_bar, parity);
        parity ^= 1;

        // ---- pre-pass: widen once, 1/N ----
        for (int i = tid; i < ncell; i += CLS_THREADS)
        {
            const float4 *p = reinterpret_cast<const float4 *>(s_rec + (size_t)i * REC);
            const float4 a0 = p[0], a1 = p[1], a2 = p[2];
            double *m = s_mom + (size_t)i * 10;
            reinterpret_cast<double2 *>(m)[0] = make_double2((double)a0.x, (double)a0.y);  // N is integral by construction
            reinterpret_cast<double2 *>(m)[1] = make_double2((double)a0.z, (double)a0.w);
            reinterpret_cast<double2 *>(m)[2] = make_double2((double)a1.x, (double)a1.y);
            reinterpret_cast<double2 *>(m)[3] = make_double2((double)a1.z, (double)a1.w);
            reinterpret_cast<double2 *>(m)[4] = make_double2((double)a2.x, (double)a2.y);
            s_inv[i] = (!isnan(a0.x) && a0.x >= 1.f) ? __drcp_rn((double)rintf(a0.x)) : 0.0;
        }
        const bool fit = (s_fraw[(y + r) * FLAG_BOX_W + (x + 16)] & F_FIT) != 0;  // marked by k_gate
        const double nq = (double)rintf(s_rec[(size_t)idx * REC + R_N]);
        const double szq = (double)s_rec[(size_t)idx * REC + R_SZ];
        __syncthreads();  // fp64 tile complete, float staging area consumed
        if (tid == 0 && nxt < n_cand)
            issue(nxt);   // arrives while this bin is fitted

        if (fit)
        {
            const size_t cell = (size_t)(row0 + y) * g.W + (col0 + x);
            g.flags[cell] &= (uint8_t)~F_FIT;
            // ---- pass A: fuse the disc about the query centre ----
            double Nf = 0, Nx = 0, Ny = 0, Nxx = 0, Nyy = 0, Nxy = 0;
            double Sx0 = 0, Sxi = 0, Sxj = 0, Sy0 = 0, Syi = 0, Syj = 0, Sz0 = 0, Szi = 0, Szj = 0;
            double Qxx = 0, Qyy = 0, Qzz = 0, Qxy = 0, Qxz = 0, Qyz = 0;
            for (int o = 0; o < nd; ++o)
            {
                const int j = idx + c_disc_off[o];
                const double2 *m = reinterpret_cast<const double2 *>(s_mom + (size_t)j * 10);
                const double2 m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4];
                const double2 d = c_disc_d[o];
                const double n = m0.x, sx = m0.y, sy = m1.x, sz = m1.y;
                const double ni = n * d.x, nj = n * d.y;
                Nf += n; Nx += ni; Ny += nj;
                Nxx = fma(ni, d.x, Nxx); Nyy = fma(nj, d.y, Nyy); Nxy = fma(ni, d.y, Nxy);
                Sx0 += sx; Sxi = fma(sx, d.x, Sxi); Sxj = fma(sx, d.y, Sxj);
                Sy0 += sy; Syi = fma(sy, d.x, Syi); Syj = fma(sy, d.y, Syj);
                Sz0 += sz; Szi = fma(sz, d.x, Szi); Szj = fma(sz, d.y, Szj);
                Qxx += m2.x; Qyy += m2.y; Qzz += m3.x;
                Qxy += m3.y; Qxz += m4.x; Qyz += m4.y;
            }
            const double rs2 = rs * rs;
            const double SX = fma(rs, Nx, Sx0), SY = fma(rs, Ny, Sy0), SZ = Sz0;
            const double QXX = Qxx + 2.0 * rs * Sxi + rs2 * Nxx;
            const double QYY = Qyy + 2.0 * rs * Syj + rs2 * Nyy;
            const double QZZ = Qzz;
            const double QXY = Qxy + rs * (Sxj + Syi) + rs2 * Nxy;
            const double QXZ = fma(rs, Szi, Qxz);
            const double QYZ = fma(rs, Szj, Qyz);
            const double inf = 1.0 / Nf;
            const double mx = SX * inf, my = SY * inf, mz = SZ * inf;
            const double cxx = QXX * inf - mx * mx, cyy = QYY * inf - my * my, czz = QZZ * inf - mz * mz;
            const double cxy = QXY * inf - mx * my, cxz = QXZ * inf - mx * mz, cyz = QYZ * inf - my * mz;
            double nv[3];
            const double lam0 = eig3_min(cxx, cxy, cxz, cyy, cyz, czz, nv);
            if (nv[2] < 0.0) { nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2]; }
            // ---- pass B: spread of the per-cell barycentres along the normal.  The barycentre of cell j about the
            // query centre is (d.x*rs + sx/N, d.y*rs + sy/N, sz/N); the footprint mean is a common offset of every
            // projection and cancels in max - min. ----
            double dmin = CUDART_INF, dmax = -CUDART_INF;
            const double rnx = rs * nv[0], rny = rs * nv[1];
            for (int o = 0; o < nd; ++o)
            {
                const int j = idx + c_disc_off[o];
                const double2 *m = reinterpret_cast<const double2 *>(s_mom + (size_t)j * 10);
                const double2 m0 = m[0], m1 = m[1];
                const double2 d = c_disc_d[o];
                const double within = fma(m0.y, nv[0], fma(m1.x, nv[1], m1.y * nv[2])) * s_inv[j];
                const double dd = fma(d.x, rnx, fma(d.y, rny, within));
                dmin = fmin(dmin, dd);
                dmax = fmax(dmax, dd);
            }
            const double slope = acos(fmin(1.0, fabs(nv[2])));
            const double slope_h = fmin(slope / c_cp.max_slope, 1.0);
            const double rough_h = fmin(sqrt(fmax(0.0, lam0)) / c_cp.ground_clearance, 1.0);
            const double step_h = fmin((dmax - dmin) / c_cp.ground_clearance, 1.0);
            const double overall = fmax(slope_h, fmax(step_h, rough_h));
            g.der[D_ELEVATION * pl + cell] = (float)(szq / nq);
            g.der[D_HAZARD * pl + cell] = (float)overall;
            g.der[D_SLOPE * pl + cell] = (float)slope_h;
            g.der[D_STEP * pl + cell] = (float)step_h;
            g.der[D_ROUGHNESS * pl + cell] = (float)rough_h;
            g.der[D_BORDER * pl + cell] = 1.0f;  // N_q >= min_points: min(N/min_points, 1) == 1
            g.der[D_NX * pl + cell] = (float)nv[0];
            g.der[D_NY * pl + cell] = (float)nv[1];
            g.der[D_NZ * pl + cell] = (float)nv[2];
        }
        __syncthreads();  // fp64 tile consumed before the next pre-pass overwrites it
        cur = nxt;
    }
}

// clear the per-batch touched / blanked bits of the active tiles
__global__ void k_clear_touched(BatchView b, GridView g)
{
    const uint32_t n_heavy = b.counters[10], n_active = n_heavy + b.counters[11];
    for (uint32_t it = blockIdx.x; it < n_active; it += gridDim.x)
    {
        const uint32_t gt = (it < n_heavy ? b.active_tiles[it] : b.active_tiles[(uint32_t)(b.btw * b.bth) - 1 - (it - n_heavy)]) & 0x00ffffffu;
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);
        const int col0 = bx * BIN - g.ci0, row0 = by * BIN - g.cj0;
        for (int q = threadIdx.x; q < BIN_CELLS; q += blockDim.x)
        {
            const size_t cell = (size_t)(row0 + (q >> BIN_SHIFT)) * g.W + (col0 + (q & (BIN - 1)));
            g.flags[cell] &= (uint8_t)~(F_TOUCHED | F_BLANKED);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_inflate (LocalMap.cpp:201-241): for observed cells, best = max over kernel taps of
// step*decay where step >= lethal; written (and flagged changed) only when the value differs.
// Recomputing a cell outside the reference's window is a no-op (pure function of step_haz, N),
// so whole candidate tiles are processed.  Per tile: one 2-D TMA box load of the step_haz halo
// tile, a summed-area table of the seed mask (cells whose footprint holds no seed finish with
// four lookups), then the taps in order of decreasing decay with an exact early exit: step <= 1,
// so once decay_k <= best no later tap can raise the maximum.  max() is order independent, so
// the result is bit-identical to the reference's loop.
// ---------------------------------------------------------------------------------------------
constexpr int INF_THREADS = 1024;  // one cell per thread (measured: 256 threads x 4 cells is 2.3x slower)
struct InflTap { int off; float decay; };

__global__ void __launch_bounds__(INF_THREADS) k_inflate(const __grid_constant__ CUtensorMap tmap_step, BatchView b, GridView g,
                                                         const InflTap *__restrict__ taps /*sorted by decay, descending*/,
                                                         const ClassifyParams c_cp)
{
    extern __shared__ __align__(1024) unsigned char s_raw[];
    const int R = c_cp.infl_r, TW = TILE + 2 * R, KW = 2 * R + 1;
    const int ntap = c_cp.n_infl;
    const size_t stage_bytes = ((size_t)TW * TW * 4 + 127) & ~(size_t)127;
    unsigned long long *s_seedrow = reinterpret_cast<unsigned long long *>(s_raw + 2 * stage_bytes);  // [TW] seed bit per column
    InflTap *s_tap = reinterpret_cast<InflTap *>(s_raw + 2 * stage_bytes + (((size_t)(TW <= 64 ? TW : 0) * 8 + 15) & ~(size_t)15));
    __shared__ __align__(8) uint64_t s_bar[2];
    const float lethal = c_cp.lethal;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_tiles = b.counters[2];
    const size_t pl = g.plane();

    auto issue = [&](int stage, uint32_t it)  // thread 0 only
    {
        const uint32_t t32 = b.infl_tiles[it];  // grid-relative 32x32 tile
        const int col0 = (int)(t32 % g.tw) * TILE, row0 = (int)(t32 / g.tw) * TILE;
        mbar_arrive_expect_tx(&s_bar[stage], (uint32_t)(TW * TW * 4));
        tma_load_2d(s_raw + (size_t)stage * stage_bytes, &tmap_step, &s_bar[stage], col0 - R, row0 - R);
    };

    if (tid == 0)
    {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < ntap; i += INF_THREADS)
        s_tap[i] = taps[i];
    __syncthreads();
    // static round-robin over the tile list, next tile's box issued one iteration ahead
    const uint32_t stride = gridDim.x;
    uint32_t curt = blockIdx.x;
    if (tid == 0 && curt < n_tiles)
        issue(0, curt);
    int stage = 0;
    uint32_t parity0 = 0, parity1 = 0;
    while (curt < n_tiles)
    {
        const uint32_t nxt = curt + stride;
        __syncthreads();  // previous tile consumed
        if (tid == 0 && nxt < n_tiles)
            issue(stage ^ 1, nxt);
        const uint32_t t32 = b.infl_tiles[curt];
        const int col0 = (int)(t32 % g.tw) * TILE, row0 = (int)(t32 / g.tw) * TILE;
        const float *s_step = reinterpret_cast<const float *>(s_raw + (size_t)stage * stage_bytes);
        // this thread's cells: N (observed?) and the current inflated value, in flight during the TMA
        float nobs[TILE_CELLS / INF_THREADS], cur[TILE_CELLS / INF_THREADS];
#pragma unroll
        for (int k = 0; k < TILE_CELLS / INF_THREADS; ++k)
        {
            const int q = tid + k * INF_THREADS;
            const size_t cell = (size_t)(row0 + (q >> 5)) * g.W + (col0 + (q & 31));
            nobs[k] = (g.flags[cell] & F_OBSERVED) ? 1.f : 0.f;  // == (N >= 1), maintained by k_fold
            cur[k] = g.der[D_INFLATED * pl + cell];
        }
        if (stage == 0) { mbar_wait(&s_bar[0], parity0); parity0 ^= 1; }
        else { mbar_wait(&s_bar[1], parity1); parity1 ^= 1; }
        // seed mask per halo row (64-bit, when the halo tile is at most 64 columns wide): one warp per row,
        // two ballots.  Wider kernels skip the per-cell box pre-test and rely on the taps' early exit.
        int any_seed = 0;
        const bool use_masks = TW <= 64;
        if (use_masks)
        {
            const uint32_t lane = tid & 31, warp = tid >> 5;
            for (int hy = warp; hy < TW; hy += INF_THREADS / 32)
            {
                const float *row = s_step + hy * TW;
                const uint32_t lo = __ballot_sync(0xffffffffu, (int)lane < TW && row[lane] >= lethal);
                const uint32_t hi = __ballot_sync(0xffffffffu, (int)lane + 32 < TW && row[min((int)lane + 32, TW - 1)] >= lethal);
                any_seed |= (lo | hi) != 0u;
                if (lane == 0)
                    s_seedrow[hy] = ((unsigned long long)hi << 32) | lo;
            }
        }
        else
        {
            for (int i = tid; i < TW * TW; i += INF_THREADS)
                any_seed |= s_step[i] >= lethal;
        }
        const int have_seed = __syncthreads_or(any_seed);
#pragma unroll
        for (int k = 0; k < TILE_CELLS / INF_THREADS; ++k)
        {
            const int q = tid + k * INF_THREADS;
            const int x = q & 31, y = q >> 5;
            // any seed inside the (2R+1)^2 box of a cell?  A warp owns one tile row (INF_THREADS == 32 x 32), so
            // the OR of that row's KW halo-row masks is formed once per warp (two warp reductions, all lanes
            // take part) and each cell only tests its KW-bit window of it.
            unsigned long long seeds = ~0ull;
            if (have_seed && use_masks)
            {
                const uint32_t lane = tid & 31;
                unsigned long long rm = (int)lane < KW ? s_seedrow[y + lane] : 0ull;
                if (KW > 32 && lane == 0) rm |= s_seedrow[y + 32];  // KW <= 33 here
                const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)rm);
                const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(rm >> 32));
                const unsigned long long km = (1ull << KW) - 1ull;
                seeds = ((((unsigned long long)hi << 32) | lo) >> x) & km;
            }
            if (!(nobs[k] >= 1.f))
                continue;  // observed cells only (LocalMap.cpp:220-221)
            float best = 0.f;
            if (have_seed)
            {
                if (seeds != 0ull)
                {
                    const int base = (y + R) * TW + (x + R);
                    for (int t = 0; t < ntap; ++t)
                    {
                        const InflTap tp = s_tap[t];
                        if (tp.decay <= best)
                            break;
                        const float sv = s_step[base + tp.off];
                        if (sv >= lethal)
                            best = fmaxf(best, __fmul_rn(sv, tp.decay));
                    }
                }
            }
            if (!(cur[k] == best))
            {
                const size_t cell = (size_t)(row0 + y) * g.W + (col0 + x);
                g.der[D_INFLATED * pl + cell] = best;
                g.flags[cell] |= F_CHANGED;
            }
        }
        curt = nxt;
        stage ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// queries / exports
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(int ci, int cj)
{
    return ((unsigned long long)(uint32_t)ci << 32) | (unsigned long long)(uint32_t)cj;
}

// mode 0: keys of cells with the CHANGED bit (optionally clearing it); mode 1: keys of observed
// cells (N >= 1).  Cells outside the logical grid |ci|<=kx, |cj|<=ky never hold data.
__global__ void k_collect_keys(GridView g, int mode, int clear, unsigned long long *out, uint32_t cap, uint32_t *count)
{
    const size_t n = g.plane();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        bool hit;
        if (mode == 0)
        {
            const uint8_t f = g.flags[i];
            hit = (f & F_CHANGED) != 0;
            if (hit && clear) g.flags[i] = f & (uint8_t)~F_CHANGED;
        }
        else
        {
            const float nn = g.mom[i * REC + R_N];
            hit = !isnan(nn) && nn >= 1.f;
        }
        if (hit)
        {
            const uint32_t p = atomicAdd(count, 1u);
            if (out && p < cap)
                out[p] = make_key(g.ci0 + (int)(i % g.W), g.cj0 + (int)(i / g.W));
        }
    }
}

__device__ __forceinline__ float read_layer(const GridView &g, int layer, int ci, int cj)
{
    const int c = ci - g.ci0, r = cj - g.cj0;
    if (c < 0 || c >= g.W || r < 0 || r >= g.H)
        return CUDART_NAN_F;
    const size_t cell = (size_t)r * g.W + c;
    if (layer < 10)
        return g.mom[cell * REC + layer];
    if (layer < 20)
        return g.der[(size_t)(layer - 10) * g.plane() + cell];
    return g.mom[cell * REC + (layer - 20 + R_ZMIN)];
}

__global__ void k_read_layers(GridView g, const unsigned long long *keys, size_t n, const int *layer_ids, int n_layers,
                              int kx, int ky, float *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_layers)
        return;
    const size_t k = i / n_layers;
    const int l = (int)(i % n_layers);
    const int ci = (int)(uint32_t)(keys[k] >> 32), cj = (int)(uint32_t)(keys[k] & 0xffffffffu);
    float v = CUDART_NAN_F;
    if (abs(ci) <= kx && abs(cj) <= ky)
        v = read_layer(g, layer_ids[l], ci, cj);
    out[i] = v;
}

__global__ void k_export_dense(GridView g, int layer, int ci0, int cj0, int w, int h, int kx, int ky, float *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)w * h)
        return;
    const int ci = ci0 + (int)(i % w), cj = cj0 + (int)(i / w);
    float v = CUDART_NAN_F;
    if (abs(ci) <= kx && abs(cj) <= ky)
        v = read_layer(g, layer, ci, cj);
    out[i] = v;
}

// grid_map_ros toOccupancyGrid(grid, "hazard", 0, 1): v=(x-0)/(1-0); NaN -> -1; else
// 0 + min(max(0,v),1)*100 assigned to int8 (truncation)
__global__ void k_export_occupancy(GridView g, int ci0, int cj0, int w, int h, int kx, int ky, signed char *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)w * h)
        return;
    const int ci = ci0 + (int)(i % w), cj = cj0 + (int)(i / w);
    float v = CUDART_NAN_F;
    if (abs(ci) <= kx && abs(cj) <= ky)
        v = read_layer(g, 10, ci, cj);
    float value = __fdiv_rn(__fsub_rn(v, 0.f), __fsub_rn(1.f, 0.f));
    if (isnan(value))
        value = -1.f;
    else
        value = __fadd_rn(0.f, __fmul_rn(fminf(fmaxf(0.0f, value), 1.0f), 100.f));
    out[i] = (signed char)value;
}

// publishCompleteness (traversability_mapping_ros/src/global_traversability_node.cpp:381-459): per coarse cell the
// MINIMUM of toOccupancyGrid(grid, "border_haz", 0, 1) over its f x f fine cells (unobserved ones skipped), then
// the RViz costmap-palette byte (100 -> 113, 0..99 -> 128 + lround(c * 126), unobserved -> -1).
__global__ void k_export_completeness(GridView g, int ci0, int cj0, int w, int h, int kx, int ky, int f, int cw, int ch,
                                      signed char *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)cw * ch)
        return;
    const int cx = (int)(i % cw), cy = (int)(i / cw);
    int best = 101;
    for (int dy = 0; dy < f; ++dy)
        for (int dx = 0; dx < f; ++dx)
        {
            const int fx = cx * f + dx, fy = cy * f + dy;
            if (fx >= w || fy >= h)
                continue;  // partial edge block
            const int ci = ci0 + fx, cj = cj0 + fy;
            float v = CUDART_NAN_F;
            if (abs(ci) <= kx && abs(cj) <= ky)
                v = read_layer(g, 15, ci, cj);  // border_haz
            float value = __fdiv_rn(__fsub_rn(v, 0.f), __fsub_rn(1.f, 0.f));
            if (isnan(value))
                continue;  // -1: carries no value
            value = __fadd_rn(0.f, __fmul_rn(fminf(fmaxf(0.0f, value), 1.0f), 100.f));
            const int iv = (int)(signed char)value;
            if (iv >= 0 && iv < best)
                best = iv;
        }
    signed char d = -1;
    if (best <= 100)
    {
        const double c = (double)best / 100.0;
        const int byte = best >= 100 ? 113 : 128 + (int)llround(c * (254 - 128));
        d = (signed char)byte;  // 128..254 wrap to negative int8 on purpose
    }
    out[i] = d;
}

// clearEntireMap (LocalMap.cpp:377-383): every occupied cell becomes 'changed'; per-batch bits drop
__global__ void k_mark_observed_changed(GridView g)
{
    const size_t n = g.plane();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const float nn = g.mom[i * REC + R_N];
        const uint8_t f = g.flags[i];
        g.flags[i] = (uint8_t)((f & F_CHANGED) | ((!isnan(nn) && nn >= 1.f) ? F_CHANGED : 0));
    }
}

// pose * cloud for the stitched cloud (LocalMap.cpp:427-437)
__global__ void k_transform_cloud(const float4 *cloud, uint32_t n, const float *T12, float *out_xyz)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float4 p = cloud[i];
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = T12[k];
    out_xyz[3 * i + 0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], p.x), __fmul_rn(T[1], p.y)), __fmul_rn(T[2], p.z)), T[3]);
    out_xyz[3 * i + 1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], p.x), __fmul_rn(T[5], p.y)), __fmul_rn(T[6], p.z)), T[7]);
    out_xyz[3 * i + 2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], p.x), __fmul_rn(T[9], p.y)), __fmul_rn(T[10], p.z)), T[11]);
}

// ---------------------------------------------------------------------------------------------
// host-callable wrappers
// ---------------------------------------------------------------------------------------------
size_t classify_smem_bytes(int r) { return ClsSmem(r).total + 128; }
size_t inflate_smem_bytes(int R, int n_taps)
{
    const size_t TW = TILE + 2 * R;
    return 2 * ((TW * TW * 4 + 127) & ~(size_t)127) + (((TW <= 64 ? TW : 0) * 8 + 15) & ~(size_t)15) + (size_t)n_taps * 8 + 128;
}

cudaError_t upload_classify_constants(const int *disc_off, const double *disc_d /*pairs*/, int n_disc)
{
    cudaError_t e = cudaMemcpyToSymbol(c_disc_off, disc_off, sizeof(int) * n_disc);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_disc_d, disc_d, sizeof(double) * 2 * n_disc);
}

// The dynamic shared-memory limit of a kernel is a per-device attribute and several maps (threads, devices, footprint
// radii) share this process: the limit only ever grows, per device, under a lock.
static void raise_smem_limit(const void *fn, int which, size_t bytes)
{
    static std::mutex m;
    static size_t limit[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) { cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); return; }
    std::lock_guard<std::mutex> l(m);
    if (bytes > limit[which][dev])
    {
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        limit[which][dev] = bytes;
    }
}

cudaError_t launch_classify(cudaStream_t st, const CUtensorMap &tmap, const CUtensorMap &tmap_flags, const BatchView &b, const GridView &g,
                            const ClassifyParams &cp, int n_sm, cudaEvent_t ev_after_gate)
{
    const int r = cp.r;
    const size_t smem = classify_smem_bytes(r);
    raise_smem_limit(reinterpret_cast<const void *>(k_classify), 0, smem);
    const size_t nb = (size_t)b.btw * b.bth;
    // gate pass over every candidate bin (flag bytes only), then the TMA-staged fit over the bins it listed
    // CTAs per SM; measured flat between 4 and 8 (profiles/sweep_grids.sh overrides them through the environment)
    static const int ggate = getenv("TMB_GATE_GRID") ? atoi(getenv("TMB_GATE_GRID")) : 6;
    k_gate<<<(int)min((size_t)(ggate * n_sm), nb), GATE_THREADS, 0, st>>>(tmap_flags, b, g, cp);
    if (ev_after_gate) cudaEventRecord(ev_after_gate, st);
    static const int gmul = getenv("TMB_CLS_GRID") ? atoi(getenv("TMB_CLS_GRID")) : 6;
    const int n_blocks = (int)min((size_t)(gmul * n_sm), nb);
    k_classify<<<n_blocks, CLS_THREADS, smem, st>>>(tmap, tmap_flags, b, g, cp);
    return cudaGetLastError();
}

cudaError_t launch_clear_touched(cudaStream_t st, const BatchView &b, const GridView &g, int n_blocks)
{
    k_clear_touched<<<n_blocks, 256, 0, st>>>(b, g);
    return cudaGetLastError();
}

cudaError_t launch_inflate(cudaStream_t st, const CUtensorMap &tmap_step, const BatchView &b, const GridView &g,
                           const ClassifyParams &cp, const void *d_taps, int n_sm)
{
    const int R = cp.infl_r;
    const size_t smem = inflate_smem_bytes(R, cp.n_infl);
    raise_smem_limit(reinterpret_cast<const void *>(k_inflate), 1, smem);
    static const int gmul = getenv("TMB_INF_GRID") ? atoi(getenv("TMB_INF_GRID")) : 4;
    const int n_blocks = (int)min((size_t)(gmul * n_sm), (size_t)g.tw * g.th);
    k_inflate<<<n_blocks, INF_THREADS, smem, st>>>(tmap_step, b, g, static_cast<const InflTap *>(d_taps), cp);
    return cudaGetLastError();
}

cudaError_t launch_collect_keys(cudaStream_t st, const GridView &g, int mode, int clear, unsigned long long *d_out, uint32_t cap,
                                uint32_t *d_count)
{
    cudaMemsetAsync(d_count, 0, 4, st);
    const size_t n = g.plane();
    const unsigned blocks = (unsigned)min((size_t)148 * 16, (n + 255) / 256);
    k_collect_keys<<<blocks ? blocks : 1, 256, 0, st>>>(g, mode, clear, d_out, cap, d_count);
    return cudaGetLastError();
}

cudaError_t launch_read_layers(cudaStream_t st, const GridView &g, const unsigned long long *d_keys, size_t n, const int *d_layers,
                               int n_layers, int kx, int ky, float *d_out)
{
    const size_t tot = n * n_layers;
    if (!tot) return cudaSuccess;
    k_read_layers<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, d_keys, n, d_layers, n_layers, kx, ky, d_out);
    return cudaGetLastError();
}

cudaError_t launch_export_dense(cudaStream_t st, const GridView &g, int layer, int ci0, int cj0, int w, int h, int kx, int ky,
                                float *d_out)
{
    const size_t tot = (size_t)w * h;
    if (!tot) return cudaSuccess;
    k_export_dense<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, layer, ci0, cj0, w, h, kx, ky, d_out);
    return cudaGetLastError();
}

cudaError_t launch_export_occupancy(cudaStream_t st, const GridView &g, int ci0, int cj0, int w, int h, int kx, int ky,
                                    signed char *d_out)
{
    const size_t tot = (size_t)w * h;
    if (!tot) return cudaSuccess;
    k_export_occupancy<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, ci0, cj0, w, h, kx, ky, d_out);
    return cudaGetLastError();
}

cudaError_t launch_export_completeness(cudaStream_t st, const GridView &g, int ci0, int cj0, int w, int h, int kx, int ky, int f,
                                       int cw, int ch, signed char *d_out)
{
    const size_t tot = (size_t)cw * ch;
    if (!tot) return cudaSuccess;
    k_export_completeness<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, ci0, cj0, w, h, kx, ky, f, cw, ch, d_out);
    return cudaGetLastError();
}

cudaError_t launch_mark_observed_changed(cudaStream_t st, const GridView &g)
{
    const size_t n = g.plane();
    const unsigned blocks = (unsigned)min((size_t)148 * 16, (n + 255) / 256);
    k_mark_observed_changed<<<blocks ? blocks : 1, 256, 0, st>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_transform_cloud(cudaStream_t st, const float4 *d_cloud, uint32_t n, const float *d_T, float *d_out)
{
    if (!n) return cudaSuccess;
    k_transform_cloud<<<(n + 255) / 256, 256, 0, st>>>(d_cloud, n, d_T, d_out);
    return cudaGetLastError();
}

}  // namespace tmb
