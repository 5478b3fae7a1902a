// kernels_bin.cu -- prune, transform + cell key, tile-bucket sort, per-tile segmented fold.
//
// Replaces (reference paths relative to traversability_mapping/):
//   System::prune                 src/System.cpp:93-114        -> k_prune_count / k_prune_scatter (also the
//                                 PointCloud2 decode of conversions.cpp:19-26: field offsets, endianness)
//   KeyFrame::rebin               src/KeyFrame.cpp:52-67       -> k_hist + k_scatter + k_fold3
//   Lattice::cellOf / centerOf    include/.../Moments.hpp:105-116
//   NodeMetaData::insert          include/.../Moments.hpp:53-59
//   MomentLayers::add             src/Helpers.cpp:114-127      -> k_fold3 (float32 fold, blank rule
//   LocalMap::addPartialToGrid    src/LocalMap.cpp:110-126        of LocalMap.cpp:121-125)
//
// Bit-exactness contract: the transform is r_i = ((R_i0 x + R_i1 y) + R_i2 z) + t_i
// in float32 with every product and sum rounded (no FMA); cell index = lround((double(x)-x0)/res);
// per (keyframe, cell) the nine double sums run in CLOUD ORDER with separately rounded products
// and sums; each keyframe's partial is cast to float32 and added to the float32 layer in op
// order.  This file is compiled with -fmad=false and uses the _rn intrinsics on those paths.
//
// Pipeline for a batch of ops (op = keyframe cloud x pose x sign); buckets are 16x16-cell bins:
//   k_hist      : per chunk of 256*PPT*SUB points (512 .. 8 192), transform+key, shared-memory histogram over
//                 the op's bin window (warp-aggregated shared atomics; integer counts are exact)
//   k_scan_*    : exclusive prefixes over chunks, ops and bins -> every (bin, op, chunk) slot,
//                 plus the work lists (bins to fold / gate, tiles to inflate)
//   k_scatter   : re-derive the keys, STABLE rank inside the chunk (per-(bin, warp) counters + one
//                 match_any per point), write (x_m, y_m, z_m, sign|op|cell) records bucketed by bin --
//                 on a sharded map straight into the owning rank's receive buffer (peer-mapped pointer)
//   k_fold3     : one CTA per touched bin: stable counting sort of a 2 048-record chunk on the 8-bit cell key
//                 into shared memory, cells handed to walker threads by descending record count, each
//                 walker follows its cell's run in (op, point) order accumulating doubles and folds float32
//                 into the cell's 48-byte record (blank rule included); cell state lives in shared memory
//   k_fold4     : experiment (TMB_FOLD=4): sorts run heads instead of records; parity-green, slower
//   k_fold_atomic_experiment : the shared-atomic alternative the sort is "chosen against" (profiles/README.md 6)
#include "tmb_common.cuh"
#include "tmb_async.cuh"
#include <math_constants.h>
#include <cstdlib>
#include <cstdio>

namespace tmb
{
__device__ __forceinline__ void xform_rn(const float *__restrict__ T, float x, float y, float z,
                                         float &xm, float &ym, float &zm)
{
    xm = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
    ym = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
    zm = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
}

// Lattice::cellOf: (int)lround((x - x0)/res), round half away from zero, in double.
// Fast path: multiply by 1/res (relative error <= 2^-52 against the true quotient) and round to nearest with the
// 2^52 + 2^51 add/subtract (two FP64-pipe adds; the result's low word is the integer, so no conversion unit
// is involved).  That can only disagree with the divided value when the quotient sits within the multiply's error
// (< 5e-7 cells for |q| < 1e9) of a rounding boundary k + 1/2, so exactly those points -- and anything outside
// the +-1e9 range, including NaN -- take the IEEE division.  Result == the reference's.
__device__ __forceinline__ int cell_of(float xm, double x0, double res, double inv_res)
{
    const double MAGIC = 6755399441055744.0;       // 2^52 + 2^51
    const double d = __dsub_rn((double)xm, x0);
    const double q = __dmul_rn(d, inv_res);
    const double t = __dadd_rn(q, MAGIC);
    const double r = __dsub_rn(t, MAGIC);          // nearest integer (ties to even)
    const double f = fabs(__dsub_rn(q, r));        // distance to it, exact, <= 0.5
    if (!(f < 0.499999) || !(fabs(q) < 1.0e9))
        return (int)llround(__ddiv_rn(d, res));    // the reference's own arithmetic, half away from zero
    return __double2loint(t);                      // away from boundaries every rounding rule agrees
}

// ---------------------------------------------------------------------------------------------
// prune (System.cpp:93-114): finite, not (x==0&&y==0), base = T*p, z <= robot_height,
// min_range <= |base| <= max_range; stable order.  One warp owns 32*PPT consecutive points.
// ---------------------------------------------------------------------------------------------
constexpr int PRUNE_PPT = 8;

__device__ __forceinline__ bool prune_point(const float *__restrict__ T, float x, float y, float z, float rh,
                                            float mx, float mn, float &xb, float &yb, float &zb, float &range)
{
    if (!isfinite(x) || !isfinite(y) || !isfinite(z))
        return false;
    if (x == 0.f && y == 0.f)
        return false;
    xform_rn(T, x, y, z, xb, yb, zb);
    if (zb > rh)
        return false;
    range = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xb, xb), __fmul_rn(yb, yb)), __fmul_rn(zb, zb)));
    if (range > mx || range < mn)
        return false;
    return true;
}

struct PruneArgs
{
    const unsigned char *src;  // device, stride bytes apart
    size_t n, stride;
    float T[12];
    float rh, mx, mn;
    int xo, yo, zo;            // position of x, y, z inside a point, in floats (PointCloud2 field offsets / 4)
    int swap;                  // 1: big-endian floats (sensor_msgs/PointCloud2::is_bigendian)
};

__device__ __forceinline__ void load_xyz(const PruneArgs &a, size_t i, float &x, float &y, float &z)
{
    const float *p = reinterpret_cast<const float *>(a.src + i * a.stride);
    x = p[a.xo]; y = p[a.yo]; z = p[a.zo];
    if (a.swap)
    {
        x = __uint_as_float(__byte_perm(__float_as_uint(x), 0, 0x0123));
        y = __uint_as_float(__byte_perm(__float_as_uint(y), 0, 0x0123));
        z = __uint_as_float(__byte_perm(__float_as_uint(z), 0, 0x0123));
    }
}

// The count pass also forms max |p_base| of the kept points (the keyframe's window radius), so the host has both
// the exact size of the keyframe store and the window bound after ONE round trip, before the scatter runs.
__global__ void k_prune_count(PruneArgs a, uint32_t *__restrict__ warp_count, uint32_t *__restrict__ out_total)
{
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const size_t base = (size_t)gw * 32 * PRUNE_PPT;
    if (base >= a.n)
        return;
    uint32_t cnt = 0;
    float mxr = 0.f;
#pragma unroll
    for (int r = 0; r < PRUNE_PPT; ++r)
    {
        const size_t i = base + r * 32 + lane;
        bool ok = false;
        if (i < a.n)
        {
            float x, y, z, xb, yb, zb, rg;
            load_xyz(a, i, x, y, z);
            ok = prune_point(a.T, x, y, z, a.rh, a.mx, a.mn, xb, yb, zb, rg);
            if (ok) mxr = fmaxf(mxr, rg);
        }
        cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        mxr = fmaxf(mxr, __shfl_xor_sync(0xffffffffu, mxr, o));
    if (lane == 0)
    {
        warp_count[gw] = cnt;
        atomicMax(out_total + 1, __float_as_uint(mxr));  // non-negative floats order like their bit patterns
    }
}

// single-block exclusive scan over n_warps counts; total -> out_total[0]; maxnorm slot zeroed
__global__ void k_prune_scan(uint32_t *__restrict__ warp_count, uint32_t n_warps, uint32_t *__restrict__ out_total)
{
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_run;
    if (threadIdx.x == 0)
        s_run = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b = 0; b < n_warps; b += blockDim.x)
    {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < n_warps ? warp_count[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            uint32_t w = lane < (blockDim.x >> 5) ? s_w[lane] : 0;
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_w[lane] = winc - w;
            if (lane == 31) s_w[32] = winc;
        }
        __syncthreads();
        const uint32_t run = s_run;
        if (i < n_warps)
            warp_count[i] = run + s_w[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0)
            s_run = run + s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        out_total[0] = s_run;
}

__global__ void k_prune_scatter(PruneArgs a, const uint32_t *__restrict__ warp_base, float4 *__restrict__ out)
{
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const size_t base = (size_t)gw * 32 * PRUNE_PPT;
    if (base >= a.n)
        return;
    uint32_t pos = warp_base[gw];
#pragma unroll
    for (int r = 0; r < PRUNE_PPT; ++r)
    {
        const size_t i = base + r * 32 + lane;
        bool ok = false;
        float xb = 0, yb = 0, zb = 0, rg = 0;
        if (i < a.n)
        {
            float x, y, z;
            load_xyz(a, i, x, y, z);
            ok = prune_point(a.T, x, y, z, a.rh, a.mx, a.mn, xb, yb, zb, rg);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok)
            out[pos + __popc(m & ((1u << lane) - 1u))] = make_float4(xb, yb, zb, 0.f);
        pos += __popc(m);
    }
}

// packed xyz (host-provided base cloud) -> float4 store + max norm
__global__ void k_pack_base(const float *__restrict__ xyz, size_t n, float4 *__restrict__ out, uint32_t *__restrict__ out_total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float rg = 0.f;
    if (i < n)
    {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        out[i] = make_float4(x, y, z, 0.f);
        rg = sqrtf(x * x + y * y + z * z);
        if (!(rg >= 0.f)) rg = CUDART_INF_F;  // NaN coordinates: force the widest window -> rejected on the host
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        rg = fmaxf(rg, __shfl_xor_sync(0xffffffffu, rg, o));
    if ((threadIdx.x & 31) == 0)
        atomicMax(out_total + 1, __float_as_uint(rg));
}

// ---------------------------------------------------------------------------------------------
// k_hist: transform + key + per-chunk bin histogram.
// A chunk is 256 * PPT * SUB consecutive points of one op, processed as SUB sub-rounds of 256 * PPT points that
// accumulate into ONE shared-memory histogram: the per-chunk fixed costs (op descriptor fetch, zeroing and
// writing the T-entry histogram row, the row's later trips through k_scan_chunks and k_scatter) are paid once per
// 8 192 points at loop-closure scale (PPT 4, SUB 8) instead of once per 2 048.  The next sub-round's points are
// requested before the current one is processed, so a CTA always has loads in flight.
// ---------------------------------------------------------------------------------------------
template <int PPT, int SUB>
__global__ void __launch_bounds__(BIN_THREADS, 5) k_hist(BatchView b, LatticeD lat)
{
    extern __shared__ uint32_t s_hist[];
    __shared__ OpDesc s_op;
    const uint32_t chunk = blockIdx.x;
    const uint32_t opi = b.chunk_op[chunk];
    if (threadIdx.x < sizeof(OpDesc) / 4)
        reinterpret_cast<uint32_t *>(&s_op)[threadIdx.x] = reinterpret_cast<const uint32_t *>(b.ops + opi)[threadIdx.x];
    __syncthreads();
    const uint32_t T = s_op.T_tiles;
    const double inv_res = 1.0 / lat.res;
    for (uint32_t i = threadIdx.x; i < T; i += BIN_THREADS)
        s_hist[i] = 0;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cl = chunk - s_op.chunk0;
    const uint32_t chunk_base = cl * (BIN_THREADS * PPT * SUB);
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
    bool bad = false;
    // op constants live in registers (the shared copy would be re-read around every shared atomic)
    const uint32_t n_pts = s_op.n;
    const float4 *cloud = s_op.cloud;
    const int wtx0 = s_op.wtx0, wty0 = s_op.wty0, wtw = s_op.wtw, wth = s_op.wth;
    float Tm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) Tm[k] = s_op.T[k];
    float4 nxt[PPT];
#pragma unroll
    for (int r = 0; r < PPT; ++r)
    {
        const uint32_t i = chunk_base + warp * (32 * PPT) + r * 32 + lane;
        nxt[r] = i < n_pts ? __ldg(cloud + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int sub = 0; sub < SUB; ++sub)
    {
        const uint32_t sub_base = chunk_base + sub * (BIN_THREADS * PPT);
        if (sub_base >= n_pts)
            break;  // uniform over the CTA
        const uint32_t base = sub_base + warp * (32 * PPT);
        float4 pts[PPT];
#pragma unroll
        for (int r = 0; r < PPT; ++r) pts[r] = nxt[r];
        if (sub + 1 < SUB)
        {
#pragma unroll
            for (int r = 0; r < PPT; ++r)
            {
                const uint32_t i = base + BIN_THREADS * PPT + r * 32 + lane;
                nxt[r] = i < n_pts ? __ldg(cloud + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < PPT; ++r)
        {
            const uint32_t i = base + r * 32 + lane;
            uint32_t tile = 0xffffffffu;
            if (i < n_pts)
            {
                const float4 p = pts[r];
                const float xm = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tm[0], p.x), __fmul_rn(Tm[1], p.y)), __fmul_rn(Tm[2], p.z)), Tm[3]);
                const float ym = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tm[4], p.x), __fmul_rn(Tm[5], p.y)), __fmul_rn(Tm[6], p.z)), Tm[7]);
                const int ci = cell_of(xm, lat.x0, lat.res, inv_res), cj = cell_of(ym, lat.y0, lat.res, inv_res);
                const int tx = (ci >> BIN_SHIFT) - wtx0, ty = (cj >> BIN_SHIFT) - wty0;
                if (tx >= 0 && tx < wtw && ty >= 0 && ty < wth)
                {
                    tile = ty * wtw + tx;
                    mnx = min(mnx, ci); mxx = max(mxx, ci); mny = min(mny, cj); mxy = max(mxy, cj);
                }
                else
                    bad = true;
            }
            const uint32_t m = __match_any_sync(0xffffffffu, tile);
            if (tile != 0xffffffffu && (m >> lane) == 1u)  // the group's highest lane adds the group size
                atomicAdd(&s_hist[tile], __popc(m));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0 && mnx != INT_MAX)
    {
        int *bb = reinterpret_cast<int *>(b.op_bbox + opi);
        atomicMin(bb + 0, mnx); atomicMax(bb + 1, mxx); atomicMin(bb + 2, mny); atomicMax(bb + 3, mxy);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0)
        atomicOr(b.counters + 3, (uint32_t)ERR_OUT_OF_WINDOW);
    __syncthreads();
    uint32_t *row = b.hist + s_op.hist_off + (size_t)cl * T;
    for (uint32_t i = threadIdx.x; i < T; i += BIN_THREADS)
        row[i] = s_hist[i];
}

// per (op, bin): exclusive prefix over the op's chunks, in place; total -> op_total.
// A CTA owns 32 adjacent columns (bins): a warp reads the 32 counters of a chunk row as one 128-byte line.
// The chunk rows are split into 8 contiguous ranges, one per warp; each warp sums its range (loads issued
// SCAN_U at a time), the range sums are exchanged through shared memory, then the range is re-read (L1/L2
// hits) and replaced by the running prefix.
constexpr int SCAN_U = 8;
__global__ void __launch_bounds__(256) k_scan_chunks(BatchView b)
{
    __shared__ uint32_t s_sum[8][32];
    const uint32_t opi = blockIdx.y;
    const OpDesc *op = b.ops + opi;
    const uint32_t T = __ldg(&op->T_tiles);
    if (blockIdx.x * 32 >= T)
        return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * 32 + lane;
    const bool ok = t < T;
    const uint32_t nch = __ldg(&op->nchunks);
    const uint32_t L = (nch + 7) / 8;
    const uint32_t c_begin = min(warp * L, nch), c_end = min(c_begin + L, nch);
    uint32_t *col = b.hist + __ldg(&op->hist_off) + t;
    uint32_t sum = 0;
    if (ok)
        for (uint32_t c0 = c_begin; c0 < c_end; c0 += SCAN_U)
        {
            uint32_t v[SCAN_U];
#pragma unroll
            for (int k = 0; k < SCAN_U; ++k)
                v[k] = c0 + k < c_end ? col[(size_t)(c0 + k) * T] : 0u;
#pragma unroll
            for (int k = 0; k < SCAN_U; ++k)
                sum += v[k];
        }
    s_sum[warp][lane] = sum;
    __syncthreads();
    uint32_t run = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
    {
        const uint32_t v = s_sum[w][lane];
        if (w < (int)warp) run += v;
        total += v;
    }
    if (!ok)
        return;
    for (uint32_t c0 = c_begin; c0 < c_end; c0 += SCAN_U)
    {
        uint32_t v[SCAN_U];
#pragma unroll
        for (int k = 0; k < SCAN_U; ++k)
            v[k] = c0 + k < c_end ? col[(size_t)(c0 + k) * T] : 0u;
#pragma unroll
        for (int k = 0; k < SCAN_U; ++k)
        {
            if (c0 + k < c_end)
                col[(size_t)(c0 + k) * T] = run;
            run += v[k];
        }
    }
    if (warp == 0)
        b.op_total[__ldg(&op->tot_off) + t] = total;
}

constexpr uint32_t HEAVY_BIN = 32768;  // records; bins at least this heavy are scheduled first

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_w /*[33]*/, uint32_t &total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        const uint32_t w = lane < nw ? s_w[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_w[lane] = winc - w;
        if (lane == 31) s_w[32] = winc;
    }
    __syncthreads();
    const uint32_t r = s_w[warp] + inc - v;
    total = s_w[32];
    __syncthreads();
    return r;
}

// per bin of the batch bbox: exclusive prefix over ops (op order = fold order) -> op_start; returns the bin count.
// Ops are visited in order; groups of 32 whose union window misses this bin are skipped whole.
__device__ __forceinline__ uint32_t scan_ops_bin(const BatchView &b, uint32_t g)
{
    const int tx = b.btx0 + (int)(g % b.btw), ty = b.bty0 + (int)(g / b.btw);
    uint32_t run = 0;
    for (uint32_t o0 = 0; o0 < b.n_ops; o0 += 32)
    {
        const int4 gw = __ldg(b.op_groups + (o0 >> 5));  // x0, y0, x1, y1 (inclusive) of the group's windows
        if (tx < gw.x || tx > gw.z || ty < gw.y || ty > gw.w)
            continue;
        const uint32_t o1 = min(o0 + 32u, b.n_ops);
        // 8 ops at a time: their descriptors and totals are requested together (read-only path, so the loads
        // are not ordered behind the op_start stores), then the running prefix is applied in op order
        for (uint32_t ob = o0; ob < o1; ob += 8)
        {
            uint32_t t[8], v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                t[k] = 0xffffffffu;
                v[k] = 0;
                if (ob + k < o1)
                {
                    const OpDesc *op = b.ops + ob + k;
                    const int4 w = __ldg(reinterpret_cast<const int4 *>(&op->wtx0));  // wtx0, wty0, wtw, wth
                    const int lx = tx - w.x, ly = ty - w.y;
                    if (lx >= 0 && lx < w.z && ly >= 0 && ly < w.w)
                    {
                        t[k] = __ldg(&op->tot_off) + (uint32_t)(ly * w.z + lx);
                        v[k] = __ldg(b.op_total + t[k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (t[k] != 0xffffffffu)
                {
                    b.op_start[t[k]] = run;
                    run += v[k];
                }
        }
    }
    return run;
}

// work lists for bin i holding v records: bins to fold (heavy first / split), bins to classify (active dilated
// by dc bins) and 32x32 tiles to inflate (within di bins of an active bin; dedup through epoch-stamped flags).
// List order is irrelevant: bins / tiles are independent work items.
// warp-aggregated append of `c` (0..K) slots per lane: every lane of the warp calls it; returns the lane's
// first slot.  One atomic round trip per warp and list.
__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, uint32_t c)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += u;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t base = 0;
    if (lane == 31 && total)
        base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + inc - c;
}

// Work lists for the K bins i[k] (count v[k], valid[k]) held by this thread; called by every thread of the warp.
//   bins to fold: light ones fill the list from the back, heavy ones go to the front so the fold's ticket loop
//     starts the long sequential chains first (longest-processing-time-first).  A bin above a fold CTA's fair
//     share is additionally split into P work items by cell-row group (rows r with r % P == part): cells are
//     independent, so P CTAs walk the same bucket and each keeps only its rows' records.  Split only as far as
//     load balance needs it: each part re-streams the bucket.
//   bins to gate / classify: active bins dilated by dc bins; 32x32 tiles to inflate: within di bins of an active
//     bin, dedup through epoch-stamped flags.  List order is irrelevant: bins / tiles are independent work items.
template <int K>
__device__ __forceinline__ void emit_lists(const BatchView &b, const GridView &g, const uint32_t (&i)[K], const uint32_t (&v)[K],
                                           const bool (&valid)[K], int dc, int di)
{
    const uint32_t nb = (uint32_t)(b.btw * b.bth);
    bool light[K], nc[K], ni[K], first[K];
    uint32_t t32[K], n_light = 0, n_nc = 0, n_first = 0, vmax = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        const bool act = valid[k] && v[k] > 0;
        light[k] = act && v[k] < HEAVY_BIN;
        n_light += light[k];
        if (act) vmax = max(vmax, v[k]);
        if (act && !light[k])
        {
            uint32_t lg = 0;
            while (lg < 4 && (v[k] >> lg) > b.split_thresh) ++lg;
            const uint32_t P = 1u << lg;
            const uint32_t at = atomicAdd(b.counters + 10, P);
            for (uint32_t part = 0; part < P; ++part)
                b.active_tiles[at + part] = i[k] | (part << 24) | (lg << 28);
        }
        nc[k] = ni[k] = false;
        t32[k] = 0;
        if (valid[k])
        {
            const int lx = i[k] % b.btw, ly = i[k] / b.btw;
            const int bx = b.btx0 + lx, by = b.bty0 + ly;  // global bin
            // inside the allocated grid, and in the rows this process owns?
            if (bx >= g.tx0 * 2 && bx < (g.tx0 + g.tw) * 2 && by >= g.ty0 * 2 && by < (g.ty0 + g.th) * 2 && ly >= b.cand_y0 &&
                ly < b.cand_y1)
            {
                const int dm = max(dc, di);
                for (int dy = -dm; dy <= dm; ++dy)
                    for (int dx = -dm; dx <= dm; ++dx)
                    {
                        const int x = lx + dx, y = ly + dy;
                        if (x < 0 || x >= b.btw || y < 0 || y >= b.bth)
                            continue;
                        if (b.nbr_count[y * b.btw + x])
                        {
                            const int d = max(abs(dx), abs(dy));
                            nc[k] |= d <= dc;
                            ni[k] |= d <= di;
                        }
                    }
                t32[k] = (uint32_t)(((by >> 1) - g.ty0) * g.tw + ((bx >> 1) - g.tx0));
            }
        }
        n_nc += nc[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        first[k] = ni[k] && atomicExch(b.infl_flags + t32[k], b.epoch) != b.epoch;
        n_first += first[k];
    }
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, 16));
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, 8));
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, 4));
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, 2));
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, 1));
    if ((threadIdx.x & 31) == 0 && vmax)
        atomicMax(b.counters + 9, vmax);  // heaviest bin of the batch (statistics)
    uint32_t sl = warp_append(b.counters + 11, n_light);
    uint32_t sc = warp_append(b.counters + 1, n_nc);
    uint32_t sf = warp_append(b.counters + 2, n_first);
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        if (light[k]) b.active_tiles[nb - 1 - sl++] = i[k];
        if (nc[k]) b.cand_tiles[sc++] = i[k];
        if (first[k]) b.infl_tiles[sf++] = t32[k];
    }
}

// per 256-bin block: bin counts + block sum
__global__ void __launch_bounds__(256) k_scan_ops(BatchView b)
{
    __shared__ uint32_t s_w[33];
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    const uint32_t nb = (uint32_t)(b.btw * b.bth);
    uint32_t run = 0;
    if (g < nb)
    {
        run = scan_ops_bin(b, g);
        b.tile_count[g] = run;
    }
    uint32_t tot;
    block_excl_scan(run, s_w, tot);
    if (threadIdx.x == 0)
        b.block_sums[blockIdx.x] = tot;
}

// single CTA: exclusive scan of the per-block sums (in place)
__global__ void __launch_bounds__(1024) k_scan_blocks(BatchView b, uint32_t n_blocks)
{
    __shared__ uint32_t s_w[33];
    uint32_t run = 0;
    for (uint32_t base = 0; base < n_blocks; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? b.block_sums[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, s_w, tot);
        if (i < n_blocks)
            b.block_sums[i] = run + ex;
        run += tot;
    }
}

// per bin: bucket start + work lists
__global__ void __launch_bounds__(256) k_tile_starts(BatchView b, GridView g, int dc, int di)
{
    __shared__ uint32_t s_w[33];
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t nb = (uint32_t)(b.btw * b.bth);
    const uint32_t v = i < nb ? b.tile_count[i] : 0;
    uint32_t tot;
    const uint32_t ex = block_excl_scan(v, s_w, tot);
    if (i < nb)
        b.tile_start[i] = b.block_sums[blockIdx.x] + ex;
    const uint32_t ii[1] = {i}, vv[1] = {v};
    const bool ok[1] = {i < nb};
    emit_lists<1>(b, g, ii, vv, ok, dc, di);
}

// the three kernels above in ONE single-CTA launch, for batches whose bbox has at most 4 096 bins (a
// stream tick): saves two dependent launches on the latency-bound path
__global__ void __launch_bounds__(1024) k_scan_small(BatchView b)
{
    __shared__ uint32_t s_w[33];
    const uint32_t nb = (uint32_t)(b.btw * b.bth);
    // a thread owns 4 consecutive bins, so one block scan of the threads' sums orders all of them
    uint32_t cnt[4], ii[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        ii[k] = threadIdx.x * 4 + k;
        ok[k] = ii[k] < nb;
        cnt[k] = 0;
    }
    if (b.n_ops == 1)
    {
        // a stream tick: the bin count is the single op's total; the four lookups are independent loads
        const OpDesc *op = b.ops;
        const int4 w = __ldg(reinterpret_cast<const int4 *>(&op->wtx0));
        const uint32_t tot_off = __ldg(&op->tot_off);
        uint32_t t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int lx = b.btx0 + (int)(ii[k] % b.btw) - w.x, ly = b.bty0 + (int)(ii[k] / b.btw) - w.y;
            t[k] = (ok[k] && lx >= 0 && lx < w.z && ly >= 0 && ly < w.w) ? tot_off + (uint32_t)(ly * w.z + lx) : 0xffffffffu;
            if (t[k] != 0xffffffffu)
                cnt[k] = __ldg(b.op_total + t[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t[k] != 0xffffffffu)
                b.op_start[t[k]] = 0;
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (ok[k])
                cnt[k] = scan_ops_bin(b, ii[k]);
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(cnt[0] + cnt[1] + cnt[2] + cnt[3], s_w, tot);
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        if (ok[k])
        {
            b.tile_count[ii[k]] = cnt[k];
            b.tile_start[ii[k]] = ex;
        }
        ex += cnt[k];
    }
    // the work lists are emitted by the first CTAs of k_scatter (emit_small_lists): one bin per thread there
    // instead of four per thread of this single CTA
}

// ---------------------------------------------------------------------------------------------
// k_scatter: stable bucket scatter by global bin, same chunking as k_hist (SUB sub-rounds of 256 * PPT points).
// Per sub-round: keys + per-warp u16 counter rows (ONE match.any per point, its rank and group leader kept in
// a register for the placement), a prefix over the warps' rows per TOUCHED bin, placement.  The absolute base of
// a bin (bin start + op offset + this chunk's prefix: three global reads) is fetched the first time a sub-round
// touches the bin -- a chunk of consecutive LiDAR returns touches a few dozen of the op's ~800 window bins.
// Record order inside a bucket is (op, point): sub-rounds, warps, rounds and lanes all advance with the index.
// ---------------------------------------------------------------------------------------------
template <int PPT, int SUB, bool EMIT>
__global__ void __launch_bounds__(BIN_THREADS, 4) k_scatter(BatchView b, LatticeD lat, GridView g, int dc, int di)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ OpDesc s_op;
    constexpr int NW = BIN_THREADS / 32;
    if (EMIT)
    {
        // small batches (k_scan_small path): the fold / gate / inflate work lists, one bin per thread, spread
        // over this kernel's CTAs (the bin counts are final since the previous launch)
        const uint32_t nb = (uint32_t)(b.btw * b.bth);
        for (uint32_t base = blockIdx.x * BIN_THREADS; base < nb; base += gridDim.x * BIN_THREADS)
        {
            const uint32_t ii[1] = {base + threadIdx.x};
            const bool ok[1] = {ii[0] < nb};
            const uint32_t vv[1] = {ok[0] ? b.tile_count[ii[0]] : 0u};
            emit_lists<1>(b, g, ii, vv, ok, dc, di);
        }
    }
    const uint32_t chunk = blockIdx.x;
    const uint32_t opi = b.chunk_op[chunk];
    if (threadIdx.x < sizeof(OpDesc) / 4)
        reinterpret_cast<uint32_t *>(&s_op)[threadIdx.x] = reinterpret_cast<const uint32_t *>(b.ops + opi)[threadIdx.x];
    __syncthreads();
    const uint32_t T = s_op.T_tiles;
    const double inv_res = 1.0 / lat.res;
    static_assert(NW == 8, "a bin's eight per-warp u16 counters are read and written as one 16-byte word");
    uint16_t *s_cnt = reinterpret_cast<uint16_t *>(s_dyn);                // [T][NW]  per (bin, warp) counters / offsets
    uint32_t *s_base = reinterpret_cast<uint32_t *>(s_cnt + (size_t)T * NW);  // [T]  absolute slot of the chunk's first record per bin
    uint16_t *s_run = reinterpret_cast<uint16_t *>(s_base + T);           // [T]  records of this chunk already placed per bin
    uint8_t *s_own = reinterpret_cast<uint8_t *>(s_run + T);              // [T]  rank that owns the bin (sharded scatter; else 0)
    __shared__ unsigned long long s_rptr[32];                             // destination array per owner
    __shared__ uint32_t s_roff[32];
    __shared__ int s_ry1[32];
    for (uint32_t i = threadIdx.x; i < T; i += BIN_THREADS)
    {
        reinterpret_cast<uint4 *>(s_cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
        s_run[i] = 0;
    }
    if (threadIdx.x < 32)
    {
        const bool routed = b.rt_world > 0 && (int)threadIdx.x < b.rt_world;
        s_rptr[threadIdx.x] = routed ? b.rt_ptr[threadIdx.x] : (unsigned long long)b.sorted;
        s_roff[threadIdx.x] = routed ? b.rt_off[threadIdx.x] : 0u;
        s_ry1[threadIdx.x] = routed ? b.rt_y1[threadIdx.x] : 0x7fffffff;
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t cl = chunk - s_op.chunk0;
    const uint32_t chunk_base = cl * (BIN_THREADS * PPT * SUB);
    uint16_t *my = s_cnt + warp;  // this warp's counter of bin t: my[t * NW]
    const uint32_t sign_bit = s_op.sign < 0.f ? META_SIGN : 0u;
    const uint32_t n_pts = s_op.n;
    const float4 *cloud = s_op.cloud;
    const int wtx0 = s_op.wtx0, wty0 = s_op.wty0, wtw = s_op.wtw, wth = s_op.wth;
    const uint32_t meta_op = sign_bit | ((opi + b.op_base) << 8);
    float Tm[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Tm[k] = s_op.T[k];
    const uint32_t *row = b.hist + s_op.hist_off + (size_t)cl * T;
    float4 nxt[PPT];
#pragma unroll
    for (int r = 0; r < PPT; ++r)
    {
        const uint32_t i = chunk_base + warp * (32 * PPT) + r * 32 + lane;
        nxt[r] = i < n_pts ? __ldg(cloud + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // absolute base of this chunk's run in every bin of the op's window = bin start + op offset + chunk prefix: three
    // global reads per bin, requested here (with the first points) so that none of them sits between two barriers
    for (uint32_t t = threadIdx.x; t < T; t += BIN_THREADS)
    {
        const int gx = wtx0 + (int)(t % wtw) - b.btx0, gy = wty0 + (int)(t / wtw) - b.bty0;
        uint32_t own = 0;
        while (gy >= s_ry1[own]) ++own;  // slab of the bin's row (single GPU: s_ry1[0] = INT_MAX)
        s_own[t] = (uint8_t)own;
        s_base[t] = __ldg(b.tile_start + gy * b.btw + gx) + __ldg(b.op_start + s_op.tot_off + t) + __ldg(row + t) + s_roff[own];
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < SUB; ++sub)
    {
        const uint32_t sub_base = chunk_base + sub * (BIN_THREADS * PPT);
        if (sub_base >= n_pts)
            break;  // uniform over the CTA
        const uint32_t base = sub_base + warp * (32 * PPT);
        float px[PPT], py[PPT], pz[PPT];
        uint32_t tk[PPT];  // bin in window (bits 0..15, 0xffff = none) | rank in group << 16 | leader << 31
        uint32_t cc[PPT];
        {
            float4 pts[PPT];
#pragma unroll
            for (int r = 0; r < PPT; ++r) pts[r] = nxt[r];
            if (sub + 1 < SUB)
            {
#pragma unroll
                for (int r = 0; r < PPT; ++r)
                {
                    const uint32_t i = base + BIN_THREADS * PPT + r * 32 + lane;
                    nxt[r] = i < n_pts ? __ldg(cloud + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int r = 0; r < PPT; ++r)
            {
                const uint32_t i = base + r * 32 + lane;
                uint32_t tl = 0xffffu;
                cc[r] = 0;
                px[r] = py[r] = pz[r] = 0.f;
                if (i < n_pts)
                {
                    const float4 p = pts[r];
                    xform_rn(Tm, p.x, p.y, p.z, px[r], py[r], pz[r]);
                    const int ci = cell_of(px[r], lat.x0, lat.res, inv_res), cj = cell_of(py[r], lat.y0, lat.res, inv_res);
                    const int tx = (ci >> BIN_SHIFT) - wtx0, ty = (cj >> BIN_SHIFT) - wty0;
                    if (tx >= 0 && tx < wtw && ty >= 0 && ty < wth)
                    {
                        tl = ty * wtw + tx;
                        cc[r] = ((cj & (BIN - 1)) << BIN_SHIFT) | (ci & (BIN - 1));
                    }
                }
                const uint32_t m = __match_any_sync(0xffffffffu, tl);
                const bool lead = (m >> lane) == 1u;
                if (tl != 0xffffu && lead)
                    my[tl * NW] += (uint16_t)__popc(m);   // column `warp` is private to this warp
                tk[r] = tl | ((uint32_t)__popc(m & lt) << 16) | (lead ? 0x80000000u : 0u);
                __syncwarp();
            }
        }
        __syncthreads();
        // per touched bin: the warps' exclusive offsets, continuing from the earlier sub-rounds of this chunk.  The eight
        // u16 counters of a bin are one 16-byte word; (warp, bin) pairs without points keep a zero for the next sub-round.
        for (uint32_t t = threadIdx.x; t < T; t += BIN_THREADS)
        {
            uint4 v = reinterpret_cast<const uint4 *>(s_cnt)[t];
            if ((v.x | v.y | v.z | v.w) == 0u)
                continue;
            uint32_t run = s_run[t];
            uint32_t w32[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                const uint32_t c0 = w32[q] & 0xffffu, c1 = w32[q] >> 16;
                const uint32_t o0 = run;
                run += c0;
                const uint32_t o1 = run;
                run += c1;
                w32[q] = (c0 ? o0 : 0u) | ((c1 ? o1 : 0u) << 16);
            }
            reinterpret_cast<uint4 *>(s_cnt)[t] = make_uint4(w32[0], w32[1], w32[2], w32[3]);
            s_run[t] = (uint16_t)run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PPT; ++r)
        {
            const uint32_t tl = tk[r] & 0xffffu;
            uint32_t slot = 0;
            if (tl != 0xffffu)
            {
                slot = my[tl * NW] + ((tk[r] >> 16) & 0x7fffu);
                // single GPU: the bucket array; sharded: the owner's receive buffer (a peer-mapped pointer -> NVLink store)
                reinterpret_cast<float4 *>(s_rptr[s_own[tl]])[s_base[tl] + slot] =
                    make_float4(px[r], py[r], pz[r], __uint_as_float(meta_op | cc[r]));
            }
            __syncwarp();
            if (tl != 0xffffu && (tk[r] & 0x80000000u))  // highest lane of the group: its slot + 1 is the group's end
                my[tl * NW] = (uint16_t)(slot + 1u);
            __syncwarp();
        }
        if (sub + 1 < SUB)
        {
            // this warp's counters back to zero for the next sub-round (they are warp-private: no CTA barrier needed
            // here; the other warps read them again only after the next sub-round's first barrier)
#pragma unroll
            for (int r = 0; r < PPT; ++r)
                if ((tk[r] & 0xffffu) != 0xffffu) my[(tk[r] & 0xffffu) * NW] = 0;
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_fold: per touched 16x16 bin -- one-pass stable counting sort by cell in shared memory, ordered
// double accumulation per (op, cell), float32 fold into the grid records, blank rule.
// 256 threads = 256 cells; ~45 KB shared memory, so several CTAs share an SM and hide each other's
// barriers and global-load latency.
// ---------------------------------------------------------------------------------------------
struct CellAcc
{
    uint32_t n;
    double sx, sy, sz, sx2, sy2, sz2, sxy, sxz, syz;
    float zmin, zmax;
    __device__ __forceinline__ void reset()
    {
        n = 0;
        sx = sy = sz = sx2 = sy2 = sz2 = sxy = sxz = syz = 0.0;
        zmin = CUDART_INF_F;
        zmax = -CUDART_INF_F;
    }
};

// Inside k_fold a cell's record is held in "zero form": a blank cell (all NaN in memory) is N = 0, sums = 0,
// zmin = +inf, zmax = -inf.  That is MomentLayers::add's NaN -> 0 (Helpers.cpp:114-127) done once at load
// instead of at every fold; N is a sum of integers, so N == 0 identifies a blank cell exactly and the store
// turns it back into NaNs.
__device__ __forceinline__ void rec_to_zero_form(float *rec)
{
    if (isnan(rec[R_N]))  // a cell is either entirely NaN or entirely numeric
    {
#pragma unroll
        for (int k = 0; k < 10; ++k) rec[k] = 0.f;
        rec[R_ZMIN] = CUDART_INF_F;
        rec[R_ZMAX] = -CUDART_INF_F;
    }
}

__device__ __forceinline__ void fold_acc(float *rec, const CellAcc &a, bool subtract, uint32_t &fl)
{
    // rec += (float)(sign * partial); (float)(-v) == -(float)v, so the sign is applied to the rounded float
    const float f0 = __double2float_rn((double)a.n), f1 = __double2float_rn(a.sx), f2 = __double2float_rn(a.sy),
                f3 = __double2float_rn(a.sz), f4 = __double2float_rn(a.sx2), f5 = __double2float_rn(a.sy2),
                f6 = __double2float_rn(a.sz2), f7 = __double2float_rn(a.sxy), f8 = __double2float_rn(a.sxz),
                f9 = __double2float_rn(a.syz);
    fl |= F_TOUCHED;
    if (!subtract)
    {
        rec[0] = __fadd_rn(rec[0], f0); rec[1] = __fadd_rn(rec[1], f1); rec[2] = __fadd_rn(rec[2], f2);
        rec[3] = __fadd_rn(rec[3], f3); rec[4] = __fadd_rn(rec[4], f4); rec[5] = __fadd_rn(rec[5], f5);
        rec[6] = __fadd_rn(rec[6], f6); rec[7] = __fadd_rn(rec[7], f7); rec[8] = __fadd_rn(rec[8], f8);
        rec[9] = __fadd_rn(rec[9], f9);
        rec[R_ZMIN] = fminf(rec[R_ZMIN], a.zmin);
        rec[R_ZMAX] = fmaxf(rec[R_ZMAX], a.zmax);
    }
    else
    {
        rec[0] = __fsub_rn(rec[0], f0); rec[1] = __fsub_rn(rec[1], f1); rec[2] = __fsub_rn(rec[2], f2);
        rec[3] = __fsub_rn(rec[3], f3); rec[4] = __fsub_rn(rec[4], f4); rec[5] = __fsub_rn(rec[5], f5);
        rec[6] = __fsub_rn(rec[6], f6); rec[7] = __fsub_rn(rec[7], f7); rec[8] = __fsub_rn(rec[8], f8);
        rec[9] = __fsub_rn(rec[9], f9);
        if (rec[R_N] < 0.5f)
        {
            // LocalMap.cpp:121-125: a subtraction that empties the cell (lround(N) <= 0) blanks all of its layers
#pragma unroll
            for (int k = 0; k < 10; ++k) rec[k] = 0.f;
            rec[R_ZMIN] = CUDART_INF_F;
            rec[R_ZMAX] = -CUDART_INF_F;
            fl |= F_BLANKED | F_CHANGED;
        }
    }
}

constexpr int FOLD_NW = FOLD_THREADS / 32;

constexpr int FOLD_R = FOLD_CH / FOLD_THREADS;  // records per thread and chunk

// ---------------------------------------------------------------------------------------------
// k_fold3: the fold with per-chunk load balancing of the walk.
// A LiDAR bucket is very uneven over the 256 cells of a bin (cells on a ring hold tens of records per chunk, cells
// between rings none): with one fixed thread per cell a warp walks for as long as its heaviest cell and 9 of 32
// lanes are busy (ncu, round 2).  Here the cell state lives in shared memory -- the 48-byte record in zero form, the
// open (keyframe, cell) segment's double sums -- so ANY thread can walk ANY cell, and after the counting sort of a
// chunk the non-empty cells are handed to threads in descending order of their record count (a 32-class counting
// sort of the counts): warps get cells of similar weight.  Cells are independent and a cell is walked whole by one
// thread per chunk, in (keyframe, point) order, so the assignment does not touch the arithmetic.
// The bucket is read straight into registers (8 coalesced 16-byte loads per thread and chunk); the next chunk -- or
// the next bin's first chunk, whose work item is resolved two bins ahead -- is pulled into L2 by one
// cp.async.bulk.prefetch while the current chunk is sorted and walked.
// ---------------------------------------------------------------------------------------------
struct Fold3Smem
{
    static constexpr size_t sorted_off = 0;                                   // float4 [FOLD_CH]
    static constexpr size_t cnt_off = sorted_off + (size_t)FOLD_CH * 16;      // u16 [8][256]
    static constexpr size_t cell_off = cnt_off + (size_t)FOLD_NW * BIN_CELLS * 2;   // float [256][12], zero form
    static constexpr size_t cd_off = cell_off + (size_t)BIN_CELLS * REC * 4;  // double [9][256] open-segment sums
    static constexpr size_t cm_off = cd_off + (size_t)9 * BIN_CELLS * 8;      // uint4 [256]: sign|op, n, zmin, zmax of the open segment
    static constexpr size_t run_off = cm_off + (size_t)BIN_CELLS * 16;        // u32 [256]: run start | count << 16
    static constexpr size_t perm_off = run_off + (size_t)BIN_CELLS * 4;       // u16 [256]: walker -> cell
    static constexpr size_t fl_off = perm_off + (size_t)BIN_CELLS * 2;        // u8 [256]: F_TOUCHED / F_BLANKED / F_CHANGED of the item
    static constexpr size_t class_off = fl_off + (size_t)BIN_CELLS;           // u32 [2][32]: class histogram, class cursors
    static constexpr size_t total = class_off + 256;
};
static_assert(Fold3Smem::total <= 73 * 1024, "three fold CTAs per SM");

__device__ __forceinline__ void l2_prefetch(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

template <bool TIMING, bool SHARD>
__global__ void __launch_bounds__(FOLD_THREADS, 3) k_fold3(BatchView b, GridView g, LatticeD lat, unsigned long long *__restrict__ tim)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    float4 *s_sorted = reinterpret_cast<float4 *>(s_dyn + Fold3Smem::sorted_off);
    uint16_t(*s_cnt)[BIN_CELLS] = reinterpret_cast<uint16_t(*)[BIN_CELLS]>(s_dyn + Fold3Smem::cnt_off);
    float4 *s_cell = reinterpret_cast<float4 *>(s_dyn + Fold3Smem::cell_off);      // [256][3]
    double(*s_cd)[BIN_CELLS] = reinterpret_cast<double(*)[BIN_CELLS]>(s_dyn + Fold3Smem::cd_off);
    uint4 *s_cm = reinterpret_cast<uint4 *>(s_dyn + Fold3Smem::cm_off);
    uint32_t *s_run = reinterpret_cast<uint32_t *>(s_dyn + Fold3Smem::run_off);
    uint16_t *s_perm = reinterpret_cast<uint16_t *>(s_dyn + Fold3Smem::perm_off);
    uint8_t *s_fl = s_dyn + Fold3Smem::fl_off;
    uint32_t *s_class = reinterpret_cast<uint32_t *>(s_dyn + Fold3Smem::class_off);  // [0..31] histogram, [32..63] cursors
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_item[2][4];  // two published work items: packed item, bucket start, bucket count, valid
    // SHARD: per published item, the sources' segments of the bin inside the receive buffer (cumulative counts, first records)
    __shared__ uint32_t s_segcum[2][33], s_segbase[2][32];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_heavy = b.counters[10];
    const uint32_t n_active = n_heavy + b.counters[11];  // heavy items (front of the list) + light bins (back)
    const uint32_t nb_total = (uint32_t)(b.btw * b.bth);
    long long t_mark = 0;
    unsigned long long t_acc[6] = {0, 0, 0, 0, 0, 0};  // load, count, scan, place, walk, store (thread 0, TIMING only)
    auto lap = [&](int k) {
        if (TIMING && tid == 0) { const long long t = clock64(); t_acc[k] += (unsigned long long)(t - t_mark); t_mark = t; }
    };

    // ---- work items: ticket counter, heaviest bins first, resolved TWO items ahead by thread 0 (see k_fold) ----
    uint32_t nn_ticket = 0, nn_item = 0xffffffffu, nn_start = 0, nn_cnt = 0;
    int nn_stage = 2;
    auto list_at = [&](uint32_t t) -> uint32_t {
        if (t >= n_active) return 0xffffffffu;
        return t < n_heavy ? b.active_tiles[t] : b.active_tiles[nb_total - 1 - (t - n_heavy)];
    };
    auto range_of = [&](uint32_t item, uint32_t &start, uint32_t &cnt) {
        start = cnt = 0;
        if (item == 0xffffffffu) return;
        const uint32_t gtn = item & 0x00ffffffu;
        start = b.tile_start[gtn];
        cnt = b.tile_count[gtn];
        const int bx = b.btx0 + (int)(gtn % b.btw), by = b.bty0 + (int)(gtn / b.btw);
        const int col = bx * BIN - g.ci0, row = by * BIN - g.cj0;
        if (col < 0 || col + BIN > g.W || row < 0 || row + BIN > g.H)
        {
            atomicOr(b.counters + 3, (uint32_t)ERR_OUT_OF_GRID);  // reported by the host; the item is skipped
            cnt = 0;
        }
    };
    auto publish = [&](int slot, uint32_t item, uint32_t start, uint32_t cnt) {
        s_item[slot][0] = item; s_item[slot][1] = start; s_item[slot][2] = cnt; s_item[slot][3] = item != 0xffffffffu;
    };
    // SHARD, warp 0: lane s resolves source s's segment of bin `gtn` (two independent loads), a warp scan orders them
    auto publish_segments = [&](int slot, uint32_t item) {
        const uint32_t nbs = (uint32_t)(b.btw * b.bth);
        uint32_t c = 0, base = 0;
        if (item != 0xffffffffu && lane < (uint32_t)b.sh_world)
        {
            const uint32_t gtn = item & 0x00ffffffu;
            c = b.sh_counts[(size_t)lane * nbs + gtn];
            base = (uint32_t)b.sh_src_base[lane] + (b.sh_starts[(size_t)lane * nbs + gtn] - b.sh_starts[(size_t)lane * nbs + b.sh_first]);
        }
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += u;
        }
        s_segcum[slot][lane + 1] = inc;
        if (lane == 0) s_segcum[slot][0] = 0;
        s_segbase[slot][lane] = base;
    };
    auto prefetch_chunk_of = [&](int slot, uint32_t start, uint32_t pos, uint32_t cnt) {  // thread 0
        const uint32_t n = min((uint32_t)FOLD_CH, cnt - pos);
        if (!SHARD)
        {
            l2_prefetch(b.sorted + start + pos, n * 16u);
            return;
        }
        // the chunk [pos, pos + n) of the logical bucket, piece by piece over the sources' segments it overlaps
        const uint32_t lo = pos, hi = pos + n;
        for (int sg = 0; sg < b.sh_world; ++sg)
        {
            const uint32_t c0 = s_segcum[slot][sg], c1 = s_segcum[slot][sg + 1];
            const uint32_t x0 = max(lo, c0), x1 = min(hi, c1);
            if (x0 < x1)
                l2_prefetch(b.sh_recv + s_segbase[slot][sg] + (x0 - c0), (x1 - x0) * 16u);
        }
    };
    auto prefetch_first_of = [&](int slot) {
        if (s_item[slot][3] && s_item[slot][2])
            prefetch_chunk_of(slot, s_item[slot][1], 0u, s_item[slot][2]);
    };
    auto nn_step = [&](int upto) {  // thread 0
        if (nn_stage < 1 && upto >= 1) { nn_item = list_at(nn_ticket); nn_stage = 1; }
        if (nn_stage < 2 && upto >= 2) { range_of(nn_item, nn_start, nn_cnt); nn_stage = 2; }
    };
    if (tid == 0)
    {
        for (int slot = 0; slot < 2; ++slot)
        {
            const uint32_t item = list_at(atomicAdd(b.counters + 4, 1u));
            uint32_t st, cn;
            range_of(item, st, cn);
            publish(slot, item, st, cn);
        }
        nn_ticket = atomicAdd(b.counters + 4, 1u);
        nn_stage = 0;
    }
#pragma unroll
    for (int w = 0; w < FOLD_NW; ++w)
        s_cnt[w][tid] = 0;
    if (tid < 64) s_class[tid] = 0;
    if (SHARD)
    {
        __syncthreads();
        if (warp == 0)
        {
            publish_segments(0, s_item[0][0]);
            publish_segments(1, s_item[1][0]);
        }
    }
    __syncthreads();
    if (TIMING && tid == 0) t_mark = clock64();

    for (uint32_t it = 0;; ++it)
    {
        const int slot = (int)(it & 1u);
        const uint32_t item = s_item[slot][0], start = s_item[slot][1], cnt = s_item[slot][2], valid = s_item[slot][3];
        if (!valid)
            break;
        const uint32_t gt = item & 0x00ffffffu, part = (item >> 24) & 15u, pmask = (1u << (item >> 28)) - 1u;
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);  // global bin
        const uint32_t nchunks = (cnt + FOLD_CH - 1) / FOLD_CH;
        uint32_t segc = 0;  // SHARD: source segment the current chunk starts in
        // thread = cell for everything that is per cell and not the walk: the record's trip from / to global memory
        const int my_ci = bx * BIN + (int)(tid & (BIN - 1)), my_cj = by * BIN + (int)(tid >> BIN_SHIFT);
        const size_t my_cell = (size_t)(my_cj - g.cj0) * g.W + (size_t)(my_ci - g.ci0);  // dereferenced only when nchunks > 0
        if (nchunks == 0)
            __syncthreads();  // everyone holds this item's fields before its slot is republished below
        else
        {
            // the whole bin's records arrive while the first chunk is sorted (zero form: Helpers.cpp:114-127's NaN -> 0)
            const float4 *src = reinterpret_cast<const float4 *>(g.mom + my_cell * REC);
            const float4 a0 = src[0], a1 = src[1], a2 = src[2];
            float rec[REC] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
            rec_to_zero_form(rec);
            s_cell[tid * 3 + 0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
            s_cell[tid * 3 + 1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
            s_cell[tid * 3 + 2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
            s_cm[tid] = make_uint4(0xffffffffu, 0u, 0u, 0u);
            s_fl[tid] = 0;
        }
        for (uint32_t c = 0; c < nchunks; ++c)
        {
            const uint32_t pos = c * FOLD_CH;
            const uint32_t n = min((uint32_t)FOLD_CH, cnt - pos);
            const uint32_t R = (n + FOLD_THREADS - 1) / FOLD_THREADS;  // rounds per thread
            const uint32_t SL = R * 32;                                // records per warp slice
            // ---- this thread's records -> registers; warp w owns the contiguous slice [w*SL, (w+1)*SL) ----
            float4 rr[FOLD_R];
            uint32_t pk[FOLD_R];  // per record: cell key (bits 0..8, 256 = not mine) | group leader (bit 15) | rank in group << 16
            const float4 *bucket = b.sorted + start + pos;
            bool straddles = false;
            if (SHARD)
            {
                // the source segment the chunk starts in (uniform over the CTA; segments are long, so nearly every chunk
                // lies inside one of them and is read exactly like a single-GPU bucket)
                while (pos >= s_segcum[slot][segc + 1]) ++segc;
                straddles = pos + n > s_segcum[slot][segc + 1];
                bucket = b.sh_recv + s_segbase[slot][segc] + (pos - s_segcum[slot][segc]);
            }
            if (!straddles)
            {
#pragma unroll
                for (int r = 0; r < FOLD_R; ++r)
                {
                    const uint32_t i = warp * SL + r * 32 + lane;
                    pk[r] = 256u;
                    if ((uint32_t)r < R && i < n)
                        rr[r] = bucket[i];
                }
            }
            else
            {
                uint32_t seg = segc;  // source segment of the last record fetched (indices only grow with r)
#pragma unroll
                for (int r = 0; r < FOLD_R; ++r)
                {
                    const uint32_t i = warp * SL + r * 32 + lane;
                    pk[r] = 256u;
                    if ((uint32_t)r < R && i < n)
                    {
                        const uint32_t gi = pos + i;  // index inside the bin's logical bucket
                        while (gi >= s_segcum[slot][seg + 1]) ++seg;
                        rr[r] = b.sh_recv[s_segbase[slot][seg] + (gi - s_segcum[slot][seg])];
                    }
                }
            }
            if (tid == 0)
            {
                if (c + 1 < nchunks) prefetch_chunk_of(slot, start, pos + FOLD_CH, cnt);  // into L2 while this chunk is sorted and walked
                else prefetch_first_of(slot ^ 1);                                // ... or the next bin's first chunk
            }
            const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
            for (int r = 0; r < FOLD_R; ++r)
            {
                const uint32_t i = warp * SL + r * 32 + lane;
                if ((uint32_t)r < R)
                {
                    uint32_t key = 256u;
                    if (i < n)
                    {
                        const uint32_t kk = __float_as_uint(rr[r].w) & 255u;
                        key = (((kk >> BIN_SHIFT) & pmask) == part) ? kk : 256u;  // pmask == 0: every record is mine
                    }
                    const uint32_t m = __match_any_sync(0xffffffffu, key);
                    const bool lead = (m >> lane) == 1u;  // the group's highest lane adds the group size
                    if (key < 256u && lead)
                        s_cnt[warp][key] += (uint16_t)__popc(m);
                    pk[r] = key | (lead ? 0x8000u : 0u) | ((uint32_t)__popc(m & lt) << 16);
                    __syncwarp();
                }
            }
            lap(0);
            __syncthreads();  // B1: counts complete (and, on the first chunk, the cell records are in shared memory)
            lap(1);
            if (tid == 0) nn_step(1);
            // ---- thread = cell: exclusive prefix over the warps' counts, then over cells; weight class of the cell ----
            uint32_t tot = 0, wc[FOLD_NW];
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
            {
                wc[w] = tot;
                tot += s_cnt[w][tid];
            }
            // 32 weight classes of width max(1, R/2) records (R = the chunk's mean per cell), heaviest = class 31
            const uint32_t cls = min(31u, tot / max(1u, R >> 1));
            if (tot) atomicAdd(&s_class[cls], 1u);
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += u;
            }
            if (lane == 31) s_w[warp] = inc;
            __syncthreads();  // B2
            uint32_t cstart = inc - tot;
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
                if (w < (int)warp) cstart += s_w[w];
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
                s_cnt[w][tid] = (uint16_t)(cstart + wc[w]);  // next slot of (warp, cell) in the sorted order
            s_run[tid] = cstart | (tot << 16);
            {
                // walkers are numbered heaviest class first; inside a class the order is whatever the cursor hands out
                // (it only decides WHICH thread walks a cell, never what is computed).  Cells in heavier classes = an
                // exclusive suffix sum over the 32 class counts, formed once per warp with shuffles (lane = class).
                uint32_t suf = s_class[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t u = __shfl_down_sync(0xffffffffu, suf, o);
                    if (lane + (uint32_t)o < 32u) suf += u;
                }
                const uint32_t heavier = __shfl_sync(0xffffffffu, suf, (cls + 1u) & 31u);  // inclusive suffix of class cls + 1
                const uint32_t before = cls == 31u ? 0u : heavier;
                if (tot)
                    s_perm[before + atomicAdd(&s_class[32 + cls], 1u)] = (uint16_t)tid;
            }
            if (tid == 0) nn_step(2);
            const uint32_t n_walk = __syncthreads_count(tot != 0);  // B3
            lap(2);
            // ---- stable placement of the records themselves ----
#pragma unroll
            for (int r = 0; r < FOLD_R; ++r)
            {
                if ((uint32_t)r < R)  // uniform over the CTA
                {
                    const uint32_t cc = pk[r] & 0x1ffu;
                    uint32_t sl = 0;
                    if (cc < 256u)
                    {
                        sl = s_cnt[warp][cc] + (pk[r] >> 16);
                        s_sorted[sl] = rr[r];
                    }
                    __syncwarp();
                    if (cc < 256u && (pk[r] & 0x8000u))  // highest lane of the group: its slot + 1 is the group's end
                        s_cnt[warp][cc] = (uint16_t)(sl + 1u);
                    __syncwarp();
                }
            }
            __syncthreads();  // B4
            lap(3);
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
                s_cnt[w][tid] = 0;  // for the next chunk's count (everyone is past the placement)
            if (tid < 64) s_class[tid] = 0;
            // ---- walk: thread `tid` takes the tid-th heaviest non-empty cell, segment by segment (see k_fold) ----
            {
                const bool walker = tid < n_walk;
                uint32_t cellk = 0, k = 0, k1 = 0;
                if (walker)
                {
                    cellk = s_perm[tid];
                    const uint32_t run = s_run[cellk];
                    k = run & 0xffffu;
                    k1 = k + (run >> 16);
                }
                const int ci = bx * BIN + (int)(cellk & (BIN - 1)), cj = by * BIN + (int)(cellk >> BIN_SHIFT);
                const double cx = __dadd_rn(lat.x0, __dmul_rn((double)ci, lat.res));  // Lattice::centerOf
                const double cy = __dadd_rn(lat.y0, __dmul_rn((double)cj, lat.res));
                CellAcc acc;
                acc.reset();
                uint32_t cur = 0xffffffffu, fl = 0;
                bool had_open = false;
                if (walker)
                {
                    const uint4 cm = s_cm[cellk];
                    cur = cm.x;
                    if (cur != 0xffffffffu)  // the segment the previous chunk left open
                    {
                        had_open = true;
                        acc.n = cm.y; acc.zmin = __uint_as_float(cm.z); acc.zmax = __uint_as_float(cm.w);
                        acc.sx = s_cd[0][cellk]; acc.sy = s_cd[1][cellk]; acc.sz = s_cd[2][cellk];
                        acc.sx2 = s_cd[3][cellk]; acc.sy2 = s_cd[4][cellk]; acc.sz2 = s_cd[5][cellk];
                        acc.sxy = s_cd[6][cellk]; acc.sxz = s_cd[7][cellk]; acc.syz = s_cd[8][cellk];
                    }
                }
                float rec[REC];
                bool loaded = false;
                auto load_rec = [&]()
                {
                    if (loaded) return;
                    const float4 a0 = s_cell[cellk * 3 + 0], a1 = s_cell[cellk * 3 + 1], a2 = s_cell[cellk * 3 + 2];
                    rec[0] = a0.x; rec[1] = a0.y; rec[2] = a0.z; rec[3] = a0.w;
                    rec[4] = a1.x; rec[5] = a1.y; rec[6] = a1.z; rec[7] = a1.w;
                    rec[8] = a2.x; rec[9] = a2.y; rec[10] = a2.z; rec[11] = a2.w;
                    loaded = true;
                };
                if (k < k1 && cur != 0xffffffffu)
                {
                    const uint32_t so0 = __float_as_uint(s_sorted[k].w) >> 8;
                    if (so0 != cur)
                    {
                        load_rec();
                        fold_acc(rec, acc, (cur >> 23) != 0, fl);
                        cur = 0xffffffffu;
                    }
                }
                for (;;)
                {
                    const bool has = k < k1;
                    if (!__any_sync(0xffffffffu, has))
                        break;
                    if (has)
                    {
                        float4 p = s_sorted[k];
                        const uint32_t so = __float_as_uint(p.w) >> 8;  // sign | op
                        double keep = (cur == so) ? 1.0 : 0.0;          // continuing the segment left open by the previous chunk?
                        if (cur != so)
                        {
                            acc.n = 0;
                            acc.zmin = CUDART_INF_F;
                            acc.zmax = -CUDART_INF_F;
                        }
                        cur = so;
                        for (;;)
                        {
                            // KeyFrame.cpp:62-65 + NodeMetaData::insert (Moments.hpp:53-59); a new segment restarts the sums
                            // through the fma's multiplier (acc * 0 + x == x, acc * 1 + x == acc + x, both exactly)
                            const double lx = __dsub_rn((double)p.x, cx), ly = __dsub_rn((double)p.y, cy), lz = (double)p.z;
                            const float z = p.z;
                            ++k;
                            bool more = false;
                            if (k < k1)
                            {
                                p = s_sorted[k];  // next record: in flight during this one's arithmetic
                                more = (__float_as_uint(p.w) >> 8) == so;
                            }
                            acc.n += 1u;
                            acc.sx = __fma_rn(acc.sx, keep, lx); acc.sy = __fma_rn(acc.sy, keep, ly); acc.sz = __fma_rn(acc.sz, keep, lz);
                            acc.sx2 = __fma_rn(acc.sx2, keep, __dmul_rn(lx, lx));
                            acc.sy2 = __fma_rn(acc.sy2, keep, __dmul_rn(ly, ly));
                            acc.sz2 = __fma_rn(acc.sz2, keep, __dmul_rn(lz, lz));
                            acc.sxy = __fma_rn(acc.sxy, keep, __dmul_rn(lx, ly));
                            acc.sxz = __fma_rn(acc.sxz, keep, __dmul_rn(lx, lz));
                            acc.syz = __fma_rn(acc.syz, keep, __dmul_rn(ly, lz));
                            acc.zmin = fminf(acc.zmin, z);
                            acc.zmax = fmaxf(acc.zmax, z);
                            keep = 1.0;
                            if (!more)
                                break;
                        }
                    }
                    if (has && k < k1)  // a following record exists => this segment is complete
                    {
                        load_rec();
                        fold_acc(rec, acc, (cur >> 23) != 0, fl);
                        cur = 0xffffffffu;
                    }
                }
                if (walker)
                {
                    if (cur != 0xffffffffu)  // the cell's last segment of this chunk stays open
                    {
                        s_cm[cellk] = make_uint4(cur, acc.n, __float_as_uint(acc.zmin), __float_as_uint(acc.zmax));
                        s_cd[0][cellk] = acc.sx; s_cd[1][cellk] = acc.sy; s_cd[2][cellk] = acc.sz;
                        s_cd[3][cellk] = acc.sx2; s_cd[4][cellk] = acc.sy2; s_cd[5][cellk] = acc.sz2;
                        s_cd[6][cellk] = acc.sxy; s_cd[7][cellk] = acc.sxz; s_cd[8][cellk] = acc.syz;
                    }
                    else if (had_open)
                        s_cm[cellk].x = 0xffffffffu;
                    if (loaded)
                    {
                        s_cell[cellk * 3 + 0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
                        s_cell[cellk * 3 + 1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
                        s_cell[cellk * 3 + 2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
                        s_fl[cellk] = (uint8_t)(s_fl[cellk] | fl);
                    }
                }
            }
            lap(4);
            __syncthreads();  // B5: s_sorted, the counters and the cell state are free for the next chunk / the store below
        }
        // ---- thread = cell: fold the segment still open after the last chunk, then the record goes back ----
        uint32_t touched = 0;
        if (nchunks)
        {
            uint32_t fl = s_fl[tid];
            const uint4 cm = s_cm[tid];
            if (cm.x != 0xffffffffu || (fl & F_TOUCHED))
            {
                const float4 a0 = s_cell[tid * 3 + 0], a1 = s_cell[tid * 3 + 1], a2 = s_cell[tid * 3 + 2];
                float rec[REC] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
                if (cm.x != 0xffffffffu)
                {
                    CellAcc acc;
                    acc.n = cm.y; acc.zmin = __uint_as_float(cm.z); acc.zmax = __uint_as_float(cm.w);
                    acc.sx = s_cd[0][tid]; acc.sy = s_cd[1][tid]; acc.sz = s_cd[2][tid];
                    acc.sx2 = s_cd[3][tid]; acc.sy2 = s_cd[4][tid]; acc.sz2 = s_cd[5][tid];
                    acc.sxy = s_cd[6][tid]; acc.sxz = s_cd[7][tid]; acc.syz = s_cd[8][tid];
                    fold_acc(rec, acc, (cm.x >> 23) != 0, fl);
                }
                touched = 1;
                float4 *dstp = reinterpret_cast<float4 *>(g.mom + my_cell * REC);
                if (rec[R_N] == 0.f)  // zero form of a blank cell -> NaN layers
                {
#pragma unroll
                    for (int k = 0; k < REC; ++k) rec[k] = CUDART_NAN_F;
                }
                dstp[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
                dstp[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
                dstp[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
                // the flag byte also mirrors "observed" (N >= 1) so the inflation kernel need not touch the records
                uint8_t obs = 0;
                if (!isnan(rec[R_N]) && rec[R_N] >= 1.f)
                    obs = (uint8_t)(F_OBSERVED | (lroundf(rec[R_N]) >= (long)b.min_points ? F_ENOUGH : 0));
                g.flags[my_cell] = (uint8_t)((g.flags[my_cell] & ~(F_OBSERVED | F_ENOUGH)) | fl | obs);
                if (fl & F_BLANKED)
                {
                    const size_t pl = g.plane();
#pragma unroll
                    for (int k = 0; k < NUM_DER; ++k) g.der[k * pl + my_cell] = CUDART_NAN_F;
                }
            }
        }
        if (SHARD && warp == 0)
        {
            if (tid == 0) nn_step(2);
            publish_segments(slot, __shfl_sync(0xffffffffu, nn_item, 0));
        }
        if (tid == 0)
        {
            nn_step(2);                                  // whatever is still outstanding of item it + 2
            publish(slot, nn_item, nn_start, nn_cnt);    // this item's slot becomes item it + 2
            nn_ticket = atomicAdd(b.counters + 4, 1u);   // consumed during the next item
            nn_stage = 0;
        }
        const uint32_t nt = __syncthreads_count(touched);  // item barrier; also publishes s_item
        if (tid == 0 && nt)
            atomicAdd(b.counters + 8, nt);
        lap(5);
    }
    if (TIMING && tid == 0)
    {
#pragma unroll
        for (int k = 0; k < 6; ++k) atomicAdd(tim + k, t_acc[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// k_fold4: the fold sorts RUNS, not records.
// A LiDAR bucket arrives in (keyframe, point) order and consecutive points mostly fall into the same cell
// (3.1 records per run on the loop-closure workload, 10-25 in the bins under the sensor).  Here
//   * the TMA unit lands a chunk of 3 072 records in shared memory and the records never move again;
//   * run heads are found with one shuffle + ballot per 32 records; only the HEADS go through the stable
//     counting sort on the 8-bit cell key (match.any over the head lanes, per-warp counters), and what is
//     placed is a 4-byte descriptor (first record | length << 12) -- a third of the sort work of k_fold3 and no
//     16-byte record traffic inside shared memory;
//   * cells are handed to walker threads in EXACT descending order of their record count in the chunk (a
//     256-class counting sort of the counts; the class scan runs in warp 0 beside the offset scan), so the
//     lanes of a warp carry near-equal work;
//   * a walker follows its cell's runs in (keyframe, point) order; inside a run there is no key test at all.
// Cell state (48-byte record in zero form, the open (keyframe, cell) segment's double sums) lives in shared
// memory between chunks as in k_fold3, so any thread can walk any cell.  Arithmetic and order of operations
// per (keyframe, cell) and per cell are exactly k_fold's.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t FOLD4_CH_MASK = 4095u;          // keeps a speculative record address inside shared memory
template <int FOLD4_CH>
struct Fold4Smem
{
    static constexpr size_t raw_off = 0;                                           // float4 [FOLD4_CH]
    static constexpr size_t runs_off = raw_off + (size_t)FOLD4_CH * 16;            // u32 [FOLD4_CH]: first record | length << 12, sorted by (cell, keyframe, point)
    static constexpr size_t cnt_off = runs_off + (size_t)FOLD4_CH * 4;             // u16 [8][256]: runs per (warp, cell) -> next slot
    static constexpr size_t tot_off = cnt_off + (size_t)FOLD_NW * BIN_CELLS * 2;   // u32 [256]: records per cell in this chunk
    static constexpr size_t hist_off = tot_off + (size_t)BIN_CELLS * 4;            // u32 [256]: cells per count class -> walker cursors
    static constexpr size_t cell_off = hist_off + (size_t)BIN_CELLS * 4;           // float [256][12], zero form
    static constexpr size_t cd_off = cell_off + (size_t)BIN_CELLS * REC * 4;       // double [9][256] open-segment sums
    static constexpr size_t cm_off = cd_off + (size_t)9 * BIN_CELLS * 8;           // uint4 [256]: sign|op, n, zmin, zmax of the open segment
    static constexpr size_t rrun_off = cm_off + (size_t)BIN_CELLS * 16;            // u32 [256]: first run slot | runs << 16
    static constexpr size_t perm_off = rrun_off + (size_t)BIN_CELLS * 4;           // u16 [256]: walker -> cell
    static constexpr size_t fl_off = perm_off + (size_t)BIN_CELLS * 2;             // u8 [256]
    static constexpr size_t total = fl_off + BIN_CELLS;
};

template <int FOLD4_CH, int PER_SM, bool TIMING>
__global__ void __launch_bounds__(FOLD_THREADS, PER_SM) k_fold4(BatchView b, GridView g, LatticeD lat, unsigned long long *__restrict__ tim)
{
    constexpr int FOLD4_R = FOLD4_CH / FOLD_THREADS;  // rows of 32 records per warp slice
    static_assert(FOLD4_CH <= 4096 && FOLD4_CH % FOLD_THREADS == 0, "run descriptors keep the first record in 12 bits");
    using Fold4Smem = tmb::Fold4Smem<FOLD4_CH>;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    float4 *s_raw = reinterpret_cast<float4 *>(s_dyn + Fold4Smem::raw_off);
    uint32_t *s_runs = reinterpret_cast<uint32_t *>(s_dyn + Fold4Smem::runs_off);
    uint16_t(*s_cnt)[BIN_CELLS] = reinterpret_cast<uint16_t(*)[BIN_CELLS]>(s_dyn + Fold4Smem::cnt_off);
    uint32_t *s_tot = reinterpret_cast<uint32_t *>(s_dyn + Fold4Smem::tot_off);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_dyn + Fold4Smem::hist_off);
    float4 *s_cell = reinterpret_cast<float4 *>(s_dyn + Fold4Smem::cell_off);      // [256][3]
    double(*s_cd)[BIN_CELLS] = reinterpret_cast<double(*)[BIN_CELLS]>(s_dyn + Fold4Smem::cd_off);
    uint4 *s_cm = reinterpret_cast<uint4 *>(s_dyn + Fold4Smem::cm_off);
    uint32_t *s_rrun = reinterpret_cast<uint32_t *>(s_dyn + Fold4Smem::rrun_off);
    uint16_t *s_perm = reinterpret_cast<uint16_t *>(s_dyn + Fold4Smem::perm_off);
    uint8_t *s_fl = s_dyn + Fold4Smem::fl_off;
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_item[2][4];  // two published work items: packed item, bucket start, bucket count, valid
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_heavy = b.counters[10];
    const uint32_t n_active = n_heavy + b.counters[11];  // heavy items (front of the list) + light bins (back)
    const uint32_t nb_total = (uint32_t)(b.btw * b.bth);
    long long t_mark = 0;
    unsigned long long t_acc[6] = {0, 0, 0, 0, 0, 0};  // wait, heads, scan, place, walk, store (thread 0, TIMING only)
    auto lap = [&](int k) {
        if (TIMING && tid == 0) { const long long t = clock64(); t_acc[k] += (unsigned long long)(t - t_mark); t_mark = t; }
    };

    // ---- work items: ticket counter, heaviest bins first, resolved TWO items ahead by thread 0 (see k_fold) ----
    uint32_t nn_ticket = 0, nn_item = 0xffffffffu, nn_start = 0, nn_cnt = 0;
    int nn_stage = 2;
    auto list_at = [&](uint32_t t) -> uint32_t {
        if (t >= n_active) return 0xffffffffu;
        return t < n_heavy ? b.active_tiles[t] : b.active_tiles[nb_total - 1 - (t - n_heavy)];
    };
    auto range_of = [&](uint32_t item, uint32_t &start, uint32_t &cnt) {
        start = cnt = 0;
        if (item == 0xffffffffu) return;
        const uint32_t gtn = item & 0x00ffffffu;
        start = b.tile_start[gtn];
        cnt = b.tile_count[gtn];
        const int bx = b.btx0 + (int)(gtn % b.btw), by = b.bty0 + (int)(gtn / b.btw);
        const int col = bx * BIN - g.ci0, row = by * BIN - g.cj0;
        if (col < 0 || col + BIN > g.W || row < 0 || row + BIN > g.H)
        {
            atomicOr(b.counters + 3, (uint32_t)ERR_OUT_OF_GRID);  // reported by the host; the item is skipped
            cnt = 0;
        }
    };
    auto publish = [&](int slot, uint32_t item, uint32_t start, uint32_t cnt) {
        s_item[slot][0] = item; s_item[slot][1] = start; s_item[slot][2] = cnt; s_item[slot][3] = item != 0xffffffffu;
    };
    auto issue_chunk = [&](uint32_t start, uint32_t pos, uint32_t cnt) {  // thread 0: chunk at `pos` of a bucket -> s_raw
        const uint32_t n = min((uint32_t)FOLD4_CH, cnt - pos);
        mbar_arrive_expect_tx(&s_bar, n * 16u);
        bulk_load(s_raw, b.sorted + start + pos, n * 16u, &s_bar);
    };
    auto issue_first_of = [&](int slot) {  // thread 0: first chunk of a published item (if it has one)
        if (s_item[slot][3] && s_item[slot][2])
            issue_chunk(s_item[slot][1], 0u, s_item[slot][2]);
    };
    auto prefetch_first_of = [&](int slot) {  // thread 0: pull a published item's first chunk into L2
        if (s_item[slot][3] && s_item[slot][2])
            l2_prefetch(b.sorted + s_item[slot][1], min((uint32_t)FOLD4_CH, s_item[slot][2]) * 16u);
    };
    auto nn_step = [&](int upto) {  // thread 0
        if (nn_stage < 1 && upto >= 1) { nn_item = list_at(nn_ticket); nn_stage = 1; }
        if (nn_stage < 2 && upto >= 2) { range_of(nn_item, nn_start, nn_cnt); nn_stage = 2; }
    };
    if (tid == 0)
    {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        for (int slot = 0; slot < 2; ++slot)
        {
            const uint32_t item = list_at(atomicAdd(b.counters + 4, 1u));
            uint32_t st, cn;
            range_of(item, st, cn);
            publish(slot, item, st, cn);
        }
        issue_first_of(0);
        nn_ticket = atomicAdd(b.counters + 4, 1u);
        nn_stage = 0;
    }
#pragma unroll
    for (int w = 0; w < FOLD_NW; ++w)
        s_cnt[w][tid] = 0;
    s_tot[tid] = 0;
    s_hist[tid] = 0;
    __syncthreads();
    if (TIMING && tid == 0) t_mark = clock64();

    uint32_t phase = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t it = 0;; ++it)
    {
        const int slot = (int)(it & 1u);
        const uint32_t item = s_item[slot][0], start = s_item[slot][1], cnt = s_item[slot][2], valid = s_item[slot][3];
        if (!valid)
            break;
        const uint32_t gt = item & 0x00ffffffu, part = (item >> 24) & 15u, pmask = (1u << (item >> 28)) - 1u;
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);  // global bin
        const uint32_t nchunks = (cnt + FOLD4_CH - 1) / FOLD4_CH;
        // thread = cell for everything that is per cell and not the walk: the record's trip from / to global memory
        const int my_ci = bx * BIN + (int)(tid & (BIN - 1)), my_cj = by * BIN + (int)(tid >> BIN_SHIFT);
        const size_t my_cell = (size_t)(my_cj - g.cj0) * g.W + (size_t)(my_ci - g.ci0);  // dereferenced only when nchunks > 0
        if (nchunks == 0)
        {
            __syncthreads();  // everyone holds this item's fields before its slot is republished below
            if (tid == 0) issue_first_of(slot ^ 1);
        }
        else
        {
            // the whole bin's records arrive while the first chunk lands (zero form: Helpers.cpp:114-127's NaN -> 0)
            const float4 *src = reinterpret_cast<const float4 *>(g.mom + my_cell * REC);
            const float4 a0 = src[0], a1 = src[1], a2 = src[2];
            float rec[REC] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
            rec_to_zero_form(rec);
            s_cell[tid * 3 + 0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
            s_cell[tid * 3 + 1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
            s_cell[tid * 3 + 2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
            s_cm[tid] = make_uint4(0xffffffffu, 0u, 0u, 0u);
            s_fl[tid] = 0;
        }
        for (uint32_t c = 0; c < nchunks; ++c)
        {
            const uint32_t pos = c * FOLD4_CH;
            const uint32_t n = min((uint32_t)FOLD4_CH, cnt - pos);
            const uint32_t R = (n + FOLD_THREADS - 1) / FOLD_THREADS;  // rows per warp slice
            const uint32_t SL = R * 32;                                // records per warp slice
            if (tid == 0)
            {
                // the chunk after this one (or the next bin's first) starts its trip to L2 now; the TMA copy itself
                // is issued when s_raw is free again, after the walk
                if (c + 1 < nchunks) l2_prefetch(b.sorted + start + pos + FOLD4_CH, min((uint32_t)FOLD4_CH, cnt - pos - FOLD4_CH) * 16u);
                else prefetch_first_of(slot ^ 1);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
            lap(0);
            // ---- run heads of this warp's slice [warp*SL, (warp+1)*SL): count runs per (warp, cell), records per cell ----
            uint32_t pk[FOLD4_R];  // per record: cell (0..7) | head (8) | group leader (9) | rank in group (10..14) | run length (16..21)
            uint32_t keys[FOLD4_R];
#pragma unroll
            for (int r = 0; r < FOLD4_R; ++r)
            {
                const uint32_t i = warp * SL + r * 32 + lane;
                keys[r] = ((uint32_t)r < R && i < n) ? __float_as_uint(s_raw[i].w) : 0xffffffffu;
            }
#pragma unroll
            for (int r = 0; r < FOLD4_R; ++r)
            {
                pk[r] = 0u;
                if ((uint32_t)r < R)  // uniform over the CTA
                {
                    const uint32_t i = warp * SL + r * 32 + lane;
                    const uint32_t key = keys[r];
                    const uint32_t cellk = key & 255u;
                    const bool mine = i < n && (((cellk >> BIN_SHIFT) & pmask) == part);  // pmask == 0: every record is mine
                    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
                    const bool head = mine && (lane == 0 || key != prev);
                    const uint32_t hm = __ballot_sync(0xffffffffu, head);
                    const uint32_t stopm = hm | ~__ballot_sync(0xffffffffu, mine);
                    if (head)
                    {
                        const uint32_t above = lane == 31 ? 0u : (stopm >> (lane + 1));
                        const uint32_t len = above ? (uint32_t)__ffs(above) : 32u - lane;
                        const uint32_t m = __match_any_sync(hm, cellk);
                        if ((m >> lane) == 1u)  // the group's highest lane books the group's runs
                            s_cnt[warp][cellk] += (uint16_t)__popc(m);  // row `warp` is private to this warp
                        atomicAdd(&s_tot[cellk], len);
                        pk[r] = cellk | 0x100u | (((m >> lane) == 1u) ? 0x200u : 0u) | ((uint32_t)__popc(m & lt) << 10) | (len << 16);
                    }
                    __syncwarp();
                }
            }
            lap(1);
            __syncthreads();  // B1: counts complete (and, on the first chunk, the cell records are in shared memory)
            if (tid == 0) nn_step(1);
            // ---- thread = cell: run offsets (exclusive over warps, then over cells) and the cell's count class ----
            uint32_t tot = 0, wc[FOLD_NW];
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
            {
                wc[w] = tot;
                tot += s_cnt[w][tid];
            }
            const uint32_t nrec = s_tot[tid];
            const uint32_t cls = min(255u, nrec);
            if (nrec) atomicAdd(&s_hist[cls], 1u);
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += u;
            }
            if (lane == 31) s_w[warp] = inc;
            __syncthreads();  // B2: warp totals, class histogram
            uint32_t cstart = inc - tot;
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
                if (w < (int)warp) cstart += s_w[w];
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
                s_cnt[w][tid] = (uint16_t)(cstart + wc[w]);  // next slot of (warp, cell) in the sorted run list
            s_rrun[tid] = cstart | (tot << 16);
            s_tot[tid] = 0;  // for the next chunk
            if (warp == 0)
            {
                // walkers are numbered by descending record count: s_hist[k] <- number of cells in classes above k
                uint32_t h[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { h[k] = s_hist[lane * 8 + k]; sum += h[k]; }
                uint32_t suf = sum;  // inclusive suffix over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t u = __shfl_down_sync(0xffffffffu, suf, o);
                    if (lane + (uint32_t)o < 32u) suf += u;
                }
                uint32_t run = suf - sum;  // cells in the classes of higher lanes
#pragma unroll
                for (int k = 7; k >= 0; --k) { s_hist[lane * 8 + k] = run; run += h[k]; }
            }
            if (tid == 0) nn_step(2);
            __syncthreads();  // B3: run offsets and walker cursors
            lap(2);
            if (nrec)
                s_perm[atomicAdd(&s_hist[cls], 1u)] = (uint16_t)tid;  // order inside a class only decides WHO walks a cell
            // ---- stable placement of the run descriptors ----
#pragma unroll
            for (int r = 0; r < FOLD4_R; ++r)
            {
                if ((uint32_t)r < R)  // uniform over the CTA
                {
                    const uint32_t cc = pk[r] & 0xffu;
                    const bool head = (pk[r] & 0x100u) != 0u;
                    uint32_t sl = 0;
                    if (head)
                    {
                        sl = s_cnt[warp][cc] + ((pk[r] >> 10) & 31u);
                        s_runs[sl] = (warp * SL + r * 32 + lane) | ((pk[r] >> 16) << 12);
                    }
                    __syncwarp();
                    if (head && (pk[r] & 0x200u))  // highest lane of the group: its slot + 1 is the group's end
                        s_cnt[warp][cc] = (uint16_t)(sl + 1u);
                    __syncwarp();
                }
            }
            const uint32_t n_walk = __syncthreads_count(nrec != 0);  // B4
            lap(3);
#pragma unroll
            for (int w = 0; w < FOLD_NW; ++w)
                s_cnt[w][tid] = 0;  // for the next chunk's count (everyone is past the placement)
            s_hist[tid] = 0;
            // ---- walk: thread `tid` takes the tid-th heaviest cell of the chunk, run by run ----
            {
                const bool walker = tid < n_walk;
                uint32_t cellk = 0, rk = 0, rend = 0;
                if (walker)
                {
                    cellk = s_perm[tid];
                    const uint32_t rr = s_rrun[cellk];
                    rk = rr & 0xffffu;
                    rend = rk + (rr >> 16);
                }
                const int ci = bx * BIN + (int)(cellk & (BIN - 1)), cj = by * BIN + (int)(cellk >> BIN_SHIFT);
                double cx = __dadd_rn(lat.x0, __dmul_rn((double)ci, lat.res));  // Lattice::centerOf
                double cy = __dadd_rn(lat.y0, __dmul_rn((double)cj, lat.res));
                asm volatile("" : "+d"(cx), "+d"(cy));  // opaque: keeps the compiler from recomputing them inside the walk loop
                CellAcc acc;
                acc.reset();
                uint32_t cur = 0xffffffffu, fl = 0;
                bool had_open = false;
                if (walker)
                {
                    const uint4 cm = s_cm[cellk];
                    cur = cm.x;
                    if (cur != 0xffffffffu)  // the segment the previous chunk left open
                    {
                        had_open = true;
                        acc.n = cm.y; acc.zmin = __uint_as_float(cm.z); acc.zmax = __uint_as_float(cm.w);
                        acc.sx = s_cd[0][cellk]; acc.sy = s_cd[1][cellk]; acc.sz = s_cd[2][cellk];
                        acc.sx2 = s_cd[3][cellk]; acc.sy2 = s_cd[4][cellk]; acc.sz2 = s_cd[5][cellk];
                        acc.sxy = s_cd[6][cellk]; acc.sxz = s_cd[7][cellk]; acc.syz = s_cd[8][cellk];
                    }
                }
                float rec[REC];
                bool loaded = false;
                auto load_rec = [&]()
                {
                    if (loaded) return;
                    const float4 a0 = s_cell[cellk * 3 + 0], a1 = s_cell[cellk * 3 + 1], a2 = s_cell[cellk * 3 + 2];
                    rec[0] = a0.x; rec[1] = a0.y; rec[2] = a0.z; rec[3] = a0.w;
                    rec[4] = a1.x; rec[5] = a1.y; rec[6] = a1.z; rec[7] = a1.w;
                    rec[8] = a2.x; rec[9] = a2.y; rec[10] = a2.z; rec[11] = a2.w;
                    loaded = true;
                };
                // ONE flat loop, one record per iteration and lane: the lane's position runs through its cell's runs in
                // (keyframe, point) order; the next record's address is known as soon as the position has advanced, so
                // its load is in flight during this record's arithmetic, and the next run's descriptor is fetched one run
                // ahead.  A change of keyframe folds the finished (keyframe, cell) segment into the cell's record (the only
                // divergent block); the cell's last segment of the chunk stays open for the next chunk.
                uint32_t pos = 0, end = 0, nd = 0;
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rk < rend)
                {
                    const uint32_t d = s_runs[rk];
                    pos = d & 0xfffu;
                    end = pos + (d >> 12);
                    p = s_raw[pos];
                    if (rk + 1 < rend) nd = s_runs[rk + 1];
                }
                while (__any_sync(0xffffffffu, rk < rend))
                {
                    if (rk < rend)
                    {
                        // advance to the following record (possibly the next run's first) and request it
                        ++pos;
                        if (pos == end)
                        {
                            ++rk;
                            pos = nd & 0xfffu;
                            end = pos + (nd >> 12);
                            if (rk + 1 < rend) nd = s_runs[rk + 1];
                        }
                        const float4 pn = s_raw[pos & (FOLD4_CH_MASK)];  // unused after the cell's last record
                        const uint32_t so = __float_as_uint(p.w) >> 8;   // sign | op
                        double keep = 1.0;
                        if (so != cur)
                        {
                            if (cur != 0xffffffffu)
                            {
                                load_rec();
                                fold_acc(rec, acc, (cur >> 23) != 0, fl);
                            }
                            // a new segment restarts the sums through the fma's multiplier (acc * 0 + x == x exactly)
                            keep = 0.0;
                            acc.n = 0;
                            acc.zmin = CUDART_INF_F;
                            acc.zmax = -CUDART_INF_F;
                            cur = so;
                        }
                        // KeyFrame.cpp:62-65 + NodeMetaData::insert (Moments.hpp:53-59)
                        const double lx = __dsub_rn((double)p.x, cx), ly = __dsub_rn((double)p.y, cy), lz = (double)p.z;
                        acc.n += 1u;
                        acc.sx = __fma_rn(acc.sx, keep, lx); acc.sy = __fma_rn(acc.sy, keep, ly); acc.sz = __fma_rn(acc.sz, keep, lz);
                        acc.sx2 = __fma_rn(acc.sx2, keep, __dmul_rn(lx, lx));
                        acc.sy2 = __fma_rn(acc.sy2, keep, __dmul_rn(ly, ly));
                        acc.sz2 = __fma_rn(acc.sz2, keep, __dmul_rn(lz, lz));
                        acc.sxy = __fma_rn(acc.sxy, keep, __dmul_rn(lx, ly));
                        acc.sxz = __fma_rn(acc.sxz, keep, __dmul_rn(lx, lz));
                        acc.syz = __fma_rn(acc.syz, keep, __dmul_rn(ly, lz));
                        acc.zmin = fminf(acc.zmin, p.z);
                        acc.zmax = fmaxf(acc.zmax, p.z);
                        p = pn;
                    }
                }
                if (walker)
                {
                    if (cur != 0xffffffffu)  // the cell's last segment of this chunk stays open
                    {
                        s_cm[cellk] = make_uint4(cur, acc.n, __float_as_uint(acc.zmin), __float_as_uint(acc.zmax));
                        s_cd[0][cellk] = acc.sx; s_cd[1][cellk] = acc.sy; s_cd[2][cellk] = acc.sz;
                        s_cd[3][cellk] = acc.sx2; s_cd[4][cellk] = acc.sy2; s_cd[5][cellk] = acc.sz2;
                        s_cd[6][cellk] = acc.sxy; s_cd[7][cellk] = acc.sxz; s_cd[8][cellk] = acc.syz;
                    }
                    else if (had_open)
                        s_cm[cellk].x = 0xffffffffu;
                    if (loaded)
                    {
                        s_cell[cellk * 3 + 0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
                        s_cell[cellk * 3 + 1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
                        s_cell[cellk * 3 + 2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
                        s_fl[cellk] = (uint8_t)(s_fl[cellk] | fl);
                    }
                }
            }
            lap(4);
            __syncthreads();  // B5: s_raw, the run list, the counters and the cell state are free
            if (tid == 0)
            {
                if (c + 1 < nchunks) issue_chunk(start, pos + FOLD4_CH, cnt);
                else issue_first_of(slot ^ 1);
            }
        }
        // ---- thread = cell: fold the segment still open after the last chunk, then the record goes back ----
        uint32_t touched = 0;
        if (nchunks)
        {
            uint32_t fl = s_fl[tid];
            const uint4 cm = s_cm[tid];
            if (cm.x != 0xffffffffu || (fl & F_TOUCHED))
            {
                const float4 a0 = s_cell[tid * 3 + 0], a1 = s_cell[tid * 3 + 1], a2 = s_cell[tid * 3 + 2];
                float rec[REC] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
                if (cm.x != 0xffffffffu)
                {
                    CellAcc acc;
                    acc.n = cm.y; acc.zmin = __uint_as_float(cm.z); acc.zmax = __uint_as_float(cm.w);
                    acc.sx = s_cd[0][tid]; acc.sy = s_cd[1][tid]; acc.sz = s_cd[2][tid];
                    acc.sx2 = s_cd[3][tid]; acc.sy2 = s_cd[4][tid]; acc.sz2 = s_cd[5][tid];
                    acc.sxy = s_cd[6][tid]; acc.sxz = s_cd[7][tid]; acc.syz = s_cd[8][tid];
                    fold_acc(rec, acc, (cm.x >> 23) != 0, fl);
                }
                touched = 1;
                float4 *dstp = reinterpret_cast<float4 *>(g.mom + my_cell * REC);
                if (rec[R_N] == 0.f)  // zero form of a blank cell -> NaN layers
                {
#pragma unroll
                    for (int k = 0; k < REC; ++k) rec[k] = CUDART_NAN_F;
                }
                dstp[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
                dstp[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
                dstp[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
                // the flag byte also mirrors "observed" (N >= 1) so the inflation kernel need not touch the records
                uint8_t obs = 0;
                if (!isnan(rec[R_N]) && rec[R_N] >= 1.f)
                    obs = (uint8_t)(F_OBSERVED | (lroundf(rec[R_N]) >= (long)b.min_points ? F_ENOUGH : 0));
                g.flags[my_cell] = (uint8_t)((g.flags[my_cell] & ~(F_OBSERVED | F_ENOUGH)) | fl | obs);
                if (fl & F_BLANKED)
                {
                    const size_t pl = g.plane();
#pragma unroll
                    for (int k = 0; k < NUM_DER; ++k) g.der[k * pl + my_cell] = CUDART_NAN_F;
                }
            }
        }
        if (tid == 0)
        {
            nn_step(2);                                  // whatever is still outstanding of item it + 2
            publish(slot, nn_item, nn_start, nn_cnt);    // this item's slot becomes item it + 2
            nn_ticket = atomicAdd(b.counters + 4, 1u);   // consumed during the next item
            nn_stage = 0;
        }
        const uint32_t nt = __syncthreads_count(touched);  // item barrier; also publishes s_item
        if (tid == 0 && nt)
            atomicAdd(b.counters + 8, nt);
        lap(5);
    }
    if (TIMING && tid == 0)
    {
#pragma unroll
        for (int k = 0; k < 6; ++k) atomicAdd(tim + k, t_acc[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// EXPERIMENT (not on the product path; enabled by TMB_EXPERIMENT_ATOMIC_FOLD=1, see profiles/README.md):
// the alternative the north_star asks the sort to be "chosen against" -- privatised shared-memory atomic
// accumulation.  One CTA per bin streams the bin's records in bucket order and atomicAdd()s the ten sums per
// cell in shared memory (fp64 atomics), then casts once to float32.  It needs no sort, but (i) the order of
// the fp64 additions is whatever the hardware serialises, and (ii) all keyframes of a batch are summed before
// the single float32 cast, whereas the reference rounds each keyframe's partial to float32 and adds those
// (Helpers.cpp:114-127) -- so its moment layers are not the reference's.  The kernel writes to a scratch
// record array; `mismatch` counts cells whose ten float32 moments differ from the sorted fold's.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FOLD_THREADS) k_fold_atomic_experiment(BatchView b, GridView g, LatticeD lat, float *scratch /*[H][W][12]*/,
                                                                         uint32_t *mismatch)
{
    __shared__ double s_acc[BIN_CELLS][10];
    const uint32_t tid = threadIdx.x;
    const uint32_t nb = (uint32_t)(b.btw * b.bth);
    for (uint32_t gt = blockIdx.x; gt < nb; gt += gridDim.x)
    {
        const uint32_t cnt = b.tile_count[gt];
        if (!cnt)
            continue;
        const uint32_t start = b.tile_start[gt];
        const int bx = b.btx0 + (int)(gt % b.btw), by = b.bty0 + (int)(gt / b.btw);
#pragma unroll
        for (int k = 0; k < 10; ++k) s_acc[tid][k] = 0.0;
        __syncthreads();
        for (uint32_t i = tid; i < cnt; i += FOLD_THREADS)
        {
            const float4 p = b.sorted[start + i];
            const uint32_t meta = __float_as_uint(p.w);
            const uint32_t c = meta & 255u;
            const double sg = (meta & META_SIGN) ? -1.0 : 1.0;
            const int ci = bx * BIN + (int)(c & (BIN - 1)), cj = by * BIN + (int)(c >> BIN_SHIFT);
            const double lx = (double)p.x - (lat.x0 + ci * lat.res), ly = (double)p.y - (lat.y0 + cj * lat.res), lz = (double)p.z;
            double *a = s_acc[c];
            atomicAdd(a + 0, sg);
            atomicAdd(a + 1, sg * lx); atomicAdd(a + 2, sg * ly); atomicAdd(a + 3, sg * lz);
            atomicAdd(a + 4, sg * lx * lx); atomicAdd(a + 5, sg * ly * ly); atomicAdd(a + 6, sg * lz * lz);
            atomicAdd(a + 7, sg * lx * ly); atomicAdd(a + 8, sg * lx * lz); atomicAdd(a + 9, sg * ly * lz);
        }
        __syncthreads();
        const int ci = bx * BIN + (int)(tid & (BIN - 1)), cj = by * BIN + (int)(tid >> BIN_SHIFT);
        const int col = ci - g.ci0, row = cj - g.cj0;
        uint32_t bad = 0;
        if (col >= 0 && col < g.W && row >= 0 && row < g.H && s_acc[tid][0] != 0.0)
        {
            const size_t cell = (size_t)row * g.W + col;
            for (int k = 0; k < 10; ++k)
            {
                const float v = (float)s_acc[tid][k];
                scratch[cell * REC + k] = v;
                bad |= (v != g.mom[cell * REC + k]);
            }
        }
        bad = __syncthreads_count(bad);
        if (tid == 0 && bad)
            atomicAdd(mismatch, bad);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// sharded rebuild (one process per GPU; bins are owned in slabs of bin rows)
//   counts_all / starts_all : [world][nb] every rank's per-bin record count and exclusive start
//   k_shard_counts : my bins' merged count (sum over sources inside my slab, 0 outside), the global
//                    total per bin (candidate-list neighbour test) and the 256-bin block sums
//   (no merge copy: k_fold3<.., SHARD> reads a bin's bucket in place as the sources' segments in rank order
//    = global op order, because ranks hold contiguous keyframe-id blocks)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_shard_counts(BatchView b, const uint32_t *__restrict__ counts_all, int world, int y0, int y1,
                                                      uint32_t *__restrict__ glob)
{
    __shared__ uint32_t s_w[33];
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t nb = (uint32_t)(b.btw * b.bth);
    uint32_t mine = 0;
    if (i < nb)
    {
        uint32_t tot = 0;
        for (int s = 0; s < world; ++s)
            tot += counts_all[(size_t)s * nb + i];
        glob[i] = tot;
        const int ly = i / b.btw;
        mine = (ly >= y0 && ly < y1) ? tot : 0;
        b.tile_count[i] = mine;
    }
    uint32_t tot;
    block_excl_scan(mine, s_w, tot);
    if (threadIdx.x == 0)
        b.block_sums[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------
// host-callable launch wrappers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_prune(cudaStream_t st, const void *d_src, size_t n, size_t stride, const float T[12], float rh,
                         float mx, float mn, uint32_t *d_warp_cnt, uint32_t *d_total, float4 *d_out, int phase,
                         const int field_floats[3], int big_endian)
{
    PruneArgs a;
    a.src = static_cast<const unsigned char *>(d_src);
    a.n = n; a.stride = stride;
    a.xo = field_floats ? field_floats[0] : 0;
    a.yo = field_floats ? field_floats[1] : 1;
    a.zo = field_floats ? field_floats[2] : 2;
    a.swap = big_endian;
    for (int i = 0; i < 12; ++i) a.T[i] = T[i];
    a.rh = rh; a.mx = mx; a.mn = mn;
    const uint32_t n_warps = (uint32_t)((n + 32 * PRUNE_PPT - 1) / (32 * PRUNE_PPT));
    const uint32_t blocks = (n_warps + 7) / 8;
    if (phase == 0)
    {
        cudaMemsetAsync(d_total, 0, 8, st);
        k_prune_count<<<blocks, 256, 0, st>>>(a, d_warp_cnt, d_total);
        k_prune_scan<<<1, 1024, 0, st>>>(d_warp_cnt, n_warps, d_total);
    }
    else
        k_prune_scatter<<<blocks, 256, 0, st>>>(a, d_warp_cnt, d_out);
    return cudaGetLastError();
}

cudaError_t launch_pack_base(cudaStream_t st, const float *d_xyz, size_t n, float4 *d_out, uint32_t *d_total)
{
    cudaMemsetAsync(d_total, 0, 8, st);
    k_pack_base<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyz, n, d_out, d_total);
    return cudaGetLastError();
}

static size_t scatter_smem(uint32_t maxT) { return (size_t)maxT * 16 + (size_t)maxT * 4 + (size_t)maxT * 2 + (size_t)maxT + 32; }

// TMB_FOLD: 3 = k_fold3 (default), 4 = k_fold4 (run sort; measured slower, kept for A/B runs: profiles/README.md R2.4);
// TMB_FOLD_TIMING=1: per-phase clock64 sums of thread 0 of every CTA, printed after the launch (diagnostics only)
static void launch_fold(cudaStream_t st, const BatchView &b, const GridView &g, const LatticeD &lat, int n_sm)
{
    static const int variant = getenv("TMB_FOLD") ? atoi(getenv("TMB_FOLD")) : 3;
    static const bool timing = getenv("TMB_FOLD_TIMING") != nullptr;
    static const int per_sm = getenv("TMB_FOLD_GRID") ? atoi(getenv("TMB_FOLD_GRID")) : 3;
    const uint32_t nb = b.btw * b.bth;
    const uint32_t blocks = (uint32_t)min((uint32_t)(n_sm * per_sm), nb);
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_done[64] = {false};
    if (dev < 64 && !attr_done[dev])
    {
        cudaFuncSetAttribute(k_fold3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold3Smem::total);
        cudaFuncSetAttribute(k_fold3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold3Smem::total);
        cudaFuncSetAttribute(k_fold3<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold3Smem::total);
        cudaFuncSetAttribute(k_fold4<3072, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold4Smem<3072>::total);
        cudaFuncSetAttribute(k_fold4<3072, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold4Smem<3072>::total);
        cudaFuncSetAttribute(k_fold4<1536, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold4Smem<1536>::total);
        cudaFuncSetAttribute(k_fold4<1536, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fold4Smem<1536>::total);
        attr_done[dev] = true;
    }
    if (b.sh_world > 0)  // sharded: the bucket is read in place from the sources' segments (k_fold3 only)
    {
        k_fold3<false, true><<<blocks, FOLD_THREADS, Fold3Smem::total, st>>>(b, g, lat, nullptr);
        return;
    }
    static unsigned long long *d_tim = nullptr;
    if (timing)
    {
        if (!d_tim) cudaMalloc(&d_tim, 64);
        cudaMemsetAsync(d_tim, 0, 64, st);
    }
    uint32_t launched = blocks;
    if (variant == 4)
    {
        // chunk size / CTAs per SM: 3072 x 2 or 1536 x 3 (TMB_FOLD_CH)
        static const int ch = getenv("TMB_FOLD_CH") ? atoi(getenv("TMB_FOLD_CH")) : 1536;
        static const int per_sm4 = getenv("TMB_FOLD_GRID") ? atoi(getenv("TMB_FOLD_GRID")) : (ch == 3072 ? 2 : 3);
        launched = (uint32_t)min((uint32_t)(n_sm * per_sm4), nb);
        if (ch == 3072)
        {
            if (timing) k_fold4<3072, 2, true><<<launched, FOLD_THREADS, Fold4Smem<3072>::total, st>>>(b, g, lat, d_tim);
            else k_fold4<3072, 2, false><<<launched, FOLD_THREADS, Fold4Smem<3072>::total, st>>>(b, g, lat, nullptr);
        }
        else
        {
            if (timing) k_fold4<1536, 3, true><<<launched, FOLD_THREADS, Fold4Smem<1536>::total, st>>>(b, g, lat, d_tim);
            else k_fold4<1536, 3, false><<<launched, FOLD_THREADS, Fold4Smem<1536>::total, st>>>(b, g, lat, nullptr);
        }
    }
    else if (timing)
        k_fold3<true, false><<<blocks, FOLD_THREADS, Fold3Smem::total, st>>>(b, g, lat, d_tim);
    else
        k_fold3<false, false><<<blocks, FOLD_THREADS, Fold3Smem::total, st>>>(b, g, lat, nullptr);
    if (!timing)
        return;
    unsigned long long h[6];
    cudaMemcpyAsync(h, d_tim, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    double tot = 0;
    for (int k = 0; k < 6; ++k) tot += (double)h[k];
    if (tot > 0)
        fprintf(stderr, "[fold phases, %% of thread-0 cycles over %u CTAs] load/wait %.1f count/heads %.1f scan %.1f place %.1f walk %.1f store %.1f | %.0f cycles per CTA\n",
                launched, 100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, 100 * h[5] / tot, tot / launched);
}

cudaError_t launch_bin(cudaStream_t st, const BatchView &b, const GridView &g, const LatticeD &lat, int ppt, uint32_t maxT,
                       int dc, int di, int n_sm, cudaEvent_t *ev /*[5] or null*/, int phases)
{
    // phases: bit0 = hist + scans + scatter, bit1 = fold, bit2 (with bit0) = stop before the scatter, bit3 = the scatter alone
    // chunk tiers (planBatch picks the largest one that still gives every SM a few chunks): points per thread and chunk
    //   2  = 256 x 2 x 1 (a stream tick), 8 = 256 x 4 x 2, 32 = 256 x 4 x 8 (loop-closure scale)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        static bool attr_done[64] = {false};
        if (dev < 64 && !attr_done[dev])
        {
            const int sm = (int)scatter_smem(MAX_WINDOW_BINS);
            cudaFuncSetAttribute(k_scatter<2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(k_scatter<2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(k_scatter<4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(k_scatter<4, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(k_scatter<4, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(k_scatter<4, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            attr_done[dev] = true;
        }
    }
    const uint32_t nb = b.btw * b.bth;
    const uint32_t nblk = (nb + 255) / 256;
    if (phases & 1)
    {
    // counters and per-op bboxes arrive initialised with the batch's single host-to-device copy
    if (ev) cudaEventRecord(ev[0], st);
    // (a rank of a sharded pass may hold no keyframe at all: nothing to key, but its per-bin counts must still be zeroed)
    if (b.n_chunks == 0) { }
    else if (ppt == 2)
        k_hist<2, 1><<<b.n_chunks, BIN_THREADS, maxT * 4, st>>>(b, lat);
    else if (ppt == 8)
        k_hist<4, 2><<<b.n_chunks, BIN_THREADS, maxT * 4, st>>>(b, lat);
    else
        k_hist<4, 8><<<b.n_chunks, BIN_THREADS, maxT * 4, st>>>(b, lat);
    if (ev) cudaEventRecord(ev[1], st);
    if (b.n_ops)
        k_scan_chunks<<<dim3((maxT + 31) / 32, b.n_ops), 256, 0, st>>>(b);
    if (nb <= 4096)
        k_scan_small<<<1, 1024, 0, st>>>(b);
    else
    {
        k_scan_ops<<<nblk, 256, 0, st>>>(b);
        k_scan_blocks<<<1, 1024, 0, st>>>(b, nblk);
        k_tile_starts<<<nblk, 256, 0, st>>>(b, g, dc, di);
    }
    if (ev) cudaEventRecord(ev[2], st);
    }
    if (((phases & 1) && !(phases & 4)) || (phases & 8))
    {
    const bool emit = nb <= 4096;
    const size_t ssm = scatter_smem(maxT);
    if (b.n_chunks == 0) { }
    else if (ppt == 2)
    {
        if (emit) k_scatter<2, 1, true><<<b.n_chunks, BIN_THREADS, ssm, st>>>(b, lat, g, dc, di);
        else k_scatter<2, 1, false><<<b.n_chunks, BIN_THREADS, ssm, st>>>(b, lat, g, dc, di);
    }
    else if (ppt == 8)
    {
        if (emit) k_scatter<4, 2, true><<<b.n_chunks, BIN_THREADS, ssm, st>>>(b, lat, g, dc, di);
        else k_scatter<4, 2, false><<<b.n_chunks, BIN_THREADS, ssm, st>>>(b, lat, g, dc, di);
    }
    else
    {
        if (emit) k_scatter<4, 8, true><<<b.n_chunks, BIN_THREADS, ssm, st>>>(b, lat, g, dc, di);
        else k_scatter<4, 8, false><<<b.n_chunks, BIN_THREADS, ssm, st>>>(b, lat, g, dc, di);
    }
    if (ev) cudaEventRecord(ev[3], st);
    }
    if (!(phases & 2))
        return cudaGetLastError();
    launch_fold(st, b, g, lat, n_sm);
    if (ev) cudaEventRecord(ev[4], st);
    return cudaGetLastError();
}

cudaError_t launch_shard_merge(cudaStream_t st, const BatchView &b, const GridView &g, const float4 *d_recv, const uint32_t *d_counts_all,
                               const uint32_t *d_starts_all, const unsigned long long *d_src_base, int world, int y0, int y1,
                               uint32_t *d_glob, int dc, int di)
{
    const uint32_t nb = b.btw * b.bth;
    const uint32_t nblk = (nb + 255) / 256;
    cudaMemsetAsync(b.counters, 0, 64, st);
    k_shard_counts<<<nblk, 256, 0, st>>>(b, d_counts_all, world, y0, y1, d_glob);
    k_scan_blocks<<<1, 1024, 0, st>>>(b, nblk);
    k_tile_starts<<<nblk, 256, 0, st>>>(b, g, dc, di);
    // the merge copy (k_shard_merge) is gone: the fold reads each bin's bucket in place from the sources' segments
    (void)d_recv; (void)d_starts_all; (void)d_src_base;
    return cudaGetLastError();
}

// per bin ROW of the batch bbox: records this process holds in it (the slab planner's input; one warp per row)
__global__ void __launch_bounds__(256) k_row_totals(BatchView b, uint32_t *__restrict__ rows)
{
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= (uint32_t)b.bth)
        return;
    uint32_t s = 0;
    for (uint32_t x = lane; x < (uint32_t)b.btw; x += 32)
        s += b.tile_count[(size_t)row * b.btw + x];
#pragma unroll
    for (int o = 16; o; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0)
        rows[row] = s;
}

// sharded rebuild: this process's cell bounding box (union of its ops' boxes) and its error flags -> out[0..4]
__global__ void __launch_bounds__(256) k_ops_bbox(BatchView b, int32_t *__restrict__ out)
{
    __shared__ int s_v[4][8];
    int v0 = INT_MAX, v1 = INT_MIN, v2 = INT_MAX, v3 = INT_MIN;
    for (uint32_t i = threadIdx.x; i < b.n_ops; i += 256)
    {
        const int4 q = b.op_bbox[i];
        if (q.x != INT_MAX) { v0 = min(v0, q.x); v1 = max(v1, q.y); v2 = min(v2, q.z); v3 = max(v3, q.w); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
    {
        v0 = min(v0, __shfl_xor_sync(0xffffffffu, v0, o)); v1 = max(v1, __shfl_xor_sync(0xffffffffu, v1, o));
        v2 = min(v2, __shfl_xor_sync(0xffffffffu, v2, o)); v3 = max(v3, __shfl_xor_sync(0xffffffffu, v3, o));
    }
    if ((threadIdx.x & 31) == 0) { s_v[0][threadIdx.x >> 5] = v0; s_v[1][threadIdx.x >> 5] = v1; s_v[2][threadIdx.x >> 5] = v2; s_v[3][threadIdx.x >> 5] = v3; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < 8; ++w) { v0 = min(v0, s_v[0][w]); v1 = max(v1, s_v[1][w]); v2 = min(v2, s_v[2][w]); v3 = max(v3, s_v[3][w]); }
        out[0] = v0; out[1] = v1; out[2] = v2; out[3] = v3;
        out[4] = (int32_t)b.counters[3];
    }
}

cudaError_t launch_ops_bbox(cudaStream_t st, const BatchView &b, int32_t *d_out)
{
    k_ops_bbox<<<1, 256, 0, st>>>(b, d_out);
    return cudaGetLastError();
}

cudaError_t launch_row_totals(cudaStream_t st, const BatchView &b, uint32_t *d_rows)
{
    const uint32_t warps = (uint32_t)b.bth;
    k_row_totals<<<(warps + 7) / 8, 256, 0, st>>>(b, d_rows);
    return cudaGetLastError();
}

cudaError_t launch_fold_atomic_experiment(cudaStream_t st, const BatchView &b, const GridView &g, const LatticeD &lat, float *d_scratch,
                                          uint32_t *d_mismatch, int n_sm)
{
    const uint32_t nb = b.btw * b.bth;
    k_fold_atomic_experiment<<<min((uint32_t)n_sm * 6u, nb), FOLD_THREADS, 0, st>>>(b, g, lat, d_scratch, d_mismatch);
    return cudaGetLastError();
}

}  // namespace tmb
