// tmb_common.cuh -- shared device/host structures of the B200 traversability hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace tmb
{
constexpr int TILE = 32;              // grid allocation / inflation tile edge (32x32 cells)
constexpr int TILE_SHIFT = 5;
constexpr int TILE_CELLS = TILE * TILE;
constexpr int BIN = 16;               // binning + classification tile edge: buckets are 16x16 cells
constexpr int BIN_SHIFT = 4;
constexpr int BIN_CELLS = BIN * BIN;  // 256: one fold / classify thread per cell
constexpr int REC = 12;               // floats per moment record
constexpr int NUM_DER = 10;           // derived planes
constexpr int MAX_WINDOW_BINS = 6144; // bins in one keyframe's window (scatter shared-memory limit)
constexpr int MAX_DISC = 1024;        // footprint offsets held in __constant__ memory
constexpr int BIN_THREADS = 256;
constexpr int FOLD_THREADS = 256;
constexpr int FOLD_CH = 2048;         // records staged per fold iteration
constexpr uint32_t META_SIGN = 0x80000000u;  // record meta: bit31 = subtract, bits 8..30 = op, bits 0..7 = cell in bin

// moment record field order (one 48-byte AoS record per cell)
enum { R_N = 0, R_SX, R_SY, R_SZ, R_SX2, R_SY2, R_SZ2, R_SXY, R_SXZ, R_SYZ, R_ZMIN, R_ZMAX };
// derived plane order
enum { D_HAZARD = 0, D_ELEVATION, D_SLOPE, D_STEP, D_ROUGHNESS, D_BORDER, D_NX, D_NY, D_NZ, D_INFLATED };
// per-cell flag bits
// OBSERVED mirrors N >= 1 and ENOUGH mirrors N >= min_points (both kept by k_fold); FIT marks a dirty cell whose
// footprint passed the gate (set by k_gate, consumed by k_classify)
enum : uint8_t { F_TOUCHED = 1, F_CHANGED = 2, F_BLANKED = 4, F_OBSERVED = 8, F_ENOUGH = 16, F_FIT = 32 };

struct LatticeD { double x0, y0, res; };

// Device grid: a dense, tile-aligned window of the absolute cell lattice.
//   mom   [H][W][12]  float   AoS moment records (N, sx..syz, zmin, zmax)
//   der   [10][H][W]  float   derived planes
//   flags [H][W]      uint8   touched / changed / blanked bits
// Column c <-> cell index ci = ci0 + c, row r <-> cj = cj0 + r; ci0, cj0, W, H multiples of 32.
struct GridView
{
    float *mom;
    float *der;
    uint8_t *flags;
    int W, H;
    int ci0, cj0;
    int tx0, ty0, tw, th;  // same window in tile units
    __host__ __device__ size_t plane() const { return (size_t)W * H; }
};

// One binning op = one keyframe cloud folded with a sign at a pose.
struct alignas(16) OpDesc
{
    const float4 *cloud;   // base-frame points (x,y,z,unused)
    uint32_t n;            // points
    float sign;            // +1 add, -1 subtract
    float T[12];           // map <- base, row-major 3x4
    int wtx0, wty0, wtw, wth;  // window in BIN units (global bin indices) covering every point
    uint32_t T_tiles;      // wtw*wth bins
    uint32_t chunk0;       // first chunk id of this op
    uint32_t nchunks;
    uint32_t tot_off;      // offset of this op's per-tile arrays (total / op_start)
    unsigned long long hist_off;  // offset (u32 units) of this op's [nchunks][T_tiles] histogram
};
static_assert(sizeof(OpDesc) == 112 && offsetof(OpDesc, wtx0) == 64, "window fields are read as one aligned int4");

struct BatchView
{
    const OpDesc *ops;
    uint32_t n_ops;
    const int4 *op_groups;     // per 32 consecutive ops: union window (x0, y0, x1, y1) in bins, inclusive
    const uint32_t *chunk_op;  // chunk -> op
    uint32_t n_chunks;
    uint32_t *hist;            // per (op, chunk, tile) counts -> exclusive prefix over chunks
    uint32_t *op_total;        // per (op, tile) point count
    uint32_t *op_start;        // per (op, tile) offset inside the global tile bucket
    // batch bounding box in BIN units (global bin indices) and per-bin arrays over it
    int btx0, bty0, btw, bth;
    uint32_t *tile_count;      // [btw*bth] records per bin
    uint32_t *tile_start;      // [btw*bth] exclusive scan of tile_count
    uint32_t *block_sums;      // [ceil(btw*bth/256)] scan scratch
    uint32_t *active_tiles;    // compact list of bins with count>0 (index into bbox)
    uint32_t *cand_tiles;      // bins to gate (active dilated by the footprint)
    uint32_t *fit_tiles;       // bins holding at least one cell to fit (k_gate -> k_classify)
    uint32_t min_points;       // traversability/min_points_per_grid (k_fold keeps the ENOUGH bit)
    uint32_t *infl_tiles;      // 32x32 tiles to inflate: (ty32 - g.ty0) * g.tw + (tx32 - g.tx0)
    uint32_t *infl_flags;      // [g.tw*g.th] dedup stamps for infl_tiles: == epoch once the tile is listed
    uint32_t epoch;            // batch serial (never 0)
    uint32_t *counters;        // [0] unused [1] n_cand [2] n_infl [3] error flags [4] fold ticket
                               // [7] cells classified [8] cells touched [9] heaviest bin [10] heavy items
                               // [11] light items [12] n_fit
    int4 *op_bbox;             // per op: min ci, max ci, min cj, max cj of its points
    float4 *sorted;            // bucket-sorted records (xm, ym, zm, meta)
    uint32_t total_points;
    uint32_t op_base;          // added to the op index written into record meta (sharded runs: globally unique ops)
    const uint32_t *nbr_count; // per-bin counts used for the candidate-list neighbour test (sharded: global totals)
    uint32_t split_thresh;     // bins with more records than this are split across CTAs by cell-row group
    int cand_y0, cand_y1;      // bbox-local bin rows [y0, y1) this process classifies / inflates (sharded: its slab)
    // sharded fold: a bin's bucket is the concatenation, in source-rank order, of the segments the ranks sent -- read in
    // place from the receive buffer (sh_world == 0: the bucket is b.sorted[tile_start .. + tile_count), one GPU)
    const float4 *sh_recv;              // received records, grouped by source rank
    const uint32_t *sh_counts;          // [world][nb] every rank's per-bin record count
    const uint32_t *sh_starts;          // [world][nb] every rank's exclusive per-bin start in ITS bucket array
    const unsigned long long *sh_src_base;  // [world] first record of each source inside sh_recv
    int sh_world;
    uint32_t sh_first;                  // first bin (bbox-local index) of this process's slab
    // sharded scatter: a record goes straight into the receive buffer of the rank that owns its bin row, through a
    // peer-mapped pointer (NVLink store); rt_world == 0: into b.sorted
    const unsigned long long *rt_ptr;   // [world] receive buffer of every rank as mapped into THIS process
    const uint32_t *rt_off;             // [world] added (mod 2^32) to the local bucket slot -> slot in the owner's buffer
    const int *rt_y1;                   // [world] exclusive last bin row (bbox-local) of every rank's slab
    int rt_world;
};

struct ClassifyParams
{
    double res, ground_clearance, max_slope;
    uint32_t min_points;
    int r;           // disc bounding radius in cells
    int n_disc;      // number of footprint offsets
    int infl_r;      // inflation radius in cells (multiple of 4), 0 = disabled
    int n_infl;
    float lethal;
    signed char row_w[36];  // footprint half-width per row dj+r (-1: empty row); rows are contiguous runs
};

enum { ERR_OUT_OF_WINDOW = 1, ERR_OUT_OF_GRID = 2 };

}  // namespace tmb
