// runtime.cu -- host runtime (System / LocalMap mirror) and the C ABI of libtmap_b200.so.
//
// Mirrors the reference's orchestration layer (paths relative to traversability_mapping/):
//   System      include/.../System.hpp:53-212, src/System.cpp   -> class System
//   LocalMap    include/.../LocalMap.hpp:55-231, src/LocalMap.cpp -> class LocalMap
//   KeyFrame    include/.../KeyFrame.hpp, src/KeyFrame.cpp:29-50 (pose state machine) -> struct KeyFrame
// What differs by design: clouds live in device memory; a keyframe's contribution is removed by
// re-binning its cloud at the pose it was integrated with (bit-identical to the stored partials the
// reference subtracts, since the binning is deterministic) instead of keeping 76-byte partials per
// (keyframe, cell); the grid is a dense tile-aligned device window that grows by reallocation.
// No CPU fallback exists: every arithmetic step of the path runs in the CUDA kernels.
#include "../../include/tmap.h"
#include "tmb_kernels.h"
#include "tmb_nccl.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace tmb
{
thread_local std::string g_last_error;

struct CudaError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};
struct Unsupported : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e__));                     \
    } while (0)

static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_h2d{0}, g_d2h{0};  // bytes copied across PCIe by this library

struct DevBuf
{
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes)
    {
        if (bytes <= cap)
            return;
        if (p) cudaFree(p);
        p = nullptr;
        const size_t want = std::max(bytes, cap + cap / 2);
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() { if (p) cudaFree(p); }
};
struct PinBuf
{
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes)
    {
        if (bytes <= cap)
            return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        const size_t want = std::max(bytes, cap + cap / 2);
        CK(cudaMallocHost(&p, want));
        cap = want;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

static inline int floordiv32(int a) { return a >> TILE_SHIFT; }
static inline int floordiv16(int a) { return a >> BIN_SHIFT; }

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn)
    {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess)
            throw CudaError("cuTensorMapEncodeTiled entry point unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// -------------------------------------------------------------------------------------------
// Device grid
// -------------------------------------------------------------------------------------------
struct DeviceGrid
{
    GridView v{};
    CUtensorMap tmap_mom{}, tmap_step{}, tmap_flags{};
    int r = 0, R = 0;  // classify halo, inflation halo (0 = none)
    uint64_t reallocs = 0;

    ~DeviceGrid() { release(nullptr); }
    // The planes come from the device's stream-ordered pool (release threshold raised by System): growing the
    // window costs no device-wide synchronisation and reuses pooled memory.  st == nullptr: synchronous free.
    void release(cudaStream_t st)
    {
        float *ptrs[2] = {v.mom, v.der};
        for (float *p : ptrs)
            if (p) { if (st) cudaFreeAsync(p, st); else cudaFree(p); }
        if (v.flags) { if (st) cudaFreeAsync(v.flags, st); else cudaFree(v.flags); }
        v = GridView{};
    }
    static void alloc_view(GridView &n, int tx0, int ty0, int tw, int th, cudaStream_t st)
    {
        n = GridView{};
        n.tx0 = tx0; n.ty0 = ty0; n.tw = tw; n.th = th;
        n.W = tw * TILE; n.H = th * TILE;
        n.ci0 = tx0 * TILE; n.cj0 = ty0 * TILE;
        const size_t pl = n.plane();
        CK(cudaMallocAsync(reinterpret_cast<void **>(&n.mom), pl * REC * sizeof(float), st));
        CK(cudaMallocAsync(reinterpret_cast<void **>(&n.der), pl * NUM_DER * sizeof(float), st));
        CK(cudaMallocAsync(reinterpret_cast<void **>(&n.flags), pl, st));
        CK(cudaMemsetAsync(n.mom, 0xFF, pl * REC * sizeof(float), st));   // all-ones = NaN
        CK(cudaMemsetAsync(n.der, 0xFF, pl * NUM_DER * sizeof(float), st));
        CK(cudaMemsetAsync(n.flags, 0, pl, st));
    }
    void encode()
    {
        EncodeTiledFn enc = get_encode();
        {
            const cuuint64_t dims[3] = {(cuuint64_t)REC, (cuuint64_t)v.W, (cuuint64_t)v.H};
            const cuuint64_t strides[2] = {(cuuint64_t)REC * 4, (cuuint64_t)v.W * REC * 4};
            const cuuint32_t box[3] = {(cuuint32_t)REC, (cuuint32_t)(BIN + 2 * r), (cuuint32_t)(BIN + 2 * r)};
            const cuuint32_t es[3] = {1, 1, 1};
            CUresult rc = enc(&tmap_mom, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.mom, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
            if (rc != CUDA_SUCCESS)
                throw CudaError("cuTensorMapEncodeTiled(moments) failed: " + std::to_string((int)rc));
        }
        {
            // flag bytes: box starts 16 columns left of the bin so the inner start stays 16-byte aligned
            const cuuint64_t dims[2] = {(cuuint64_t)v.W, (cuuint64_t)v.H};
            const cuuint64_t strides[1] = {(cuuint64_t)v.W};
            const cuuint32_t box[2] = {(cuuint32_t)(BIN + 32), (cuuint32_t)(BIN + 2 * r)};
            const cuuint32_t es[2] = {1, 1};
            CUresult rc = enc(&tmap_flags, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, v.flags, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS)
                throw CudaError("cuTensorMapEncodeTiled(flags) failed: " + std::to_string((int)rc));
        }
        if (R > 0)
        {
            const cuuint64_t dims[2] = {(cuuint64_t)v.W, (cuuint64_t)v.H};
            const cuuint64_t strides[1] = {(cuuint64_t)v.W * 4};
            const cuuint32_t box[2] = {(cuuint32_t)(TILE + 2 * R), (cuuint32_t)(TILE + 2 * R)};
            const cuuint32_t es[2] = {1, 1};
            CUresult rc = enc(&tmap_step, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, v.der + (size_t)D_STEP * v.plane(), dims, strides,
                              box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
            if (rc != CUDA_SUCCESS)
                throw CudaError("cuTensorMapEncodeTiled(step) failed: " + std::to_string((int)rc));
        }
    }
    // make the window cover tiles [tx0,tx1) x [ty0,ty1); existing content is kept
    void ensure(cudaStream_t st, int tx0, int ty0, int tx1, int ty1)
    {
        if (v.mom && tx0 >= v.tx0 && ty0 >= v.ty0 && tx1 <= v.tx0 + v.tw && ty1 <= v.ty0 + v.th)
            return;
        int nx0 = tx0, ny0 = ty0, nx1 = tx1, ny1 = ty1;
        if (!v.mom)
        {
            // first allocation: spare capacity on every side (a reallocation maps fresh device memory and copies
            // the whole window -- milliseconds -- so a moving robot should not trigger one every few metres)
            const int pad = std::min(16, std::max(4, std::max(tx1 - tx0, ty1 - ty0) / 2));
            nx0 -= pad; ny0 -= pad; nx1 += pad; ny1 += pad;
        }
        else
        {
            // grow geometrically in the direction that is needed so repeated extension amortises
            // (at least 8 tiles, at most 32: a 1 km map must not double its footprint for one more keyframe)
            const int gx = std::min(32, std::max(8, v.tw / 2)), gy = std::min(32, std::max(8, v.th / 2));
            if (tx0 < v.tx0) nx0 = std::min(tx0, v.tx0 - gx); else nx0 = v.tx0;
            if (ty0 < v.ty0) ny0 = std::min(ty0, v.ty0 - gy); else ny0 = v.ty0;
            if (tx1 > v.tx0 + v.tw) nx1 = std::max(tx1, v.tx0 + v.tw + gx); else nx1 = v.tx0 + v.tw;
            if (ty1 > v.ty0 + v.th) ny1 = std::max(ty1, v.ty0 + v.th + gy); else ny1 = v.ty0 + v.th;
        }
        GridView n;
        alloc_view(n, nx0, ny0, nx1 - nx0, ny1 - ny0, st);
        if (v.mom)
        {
            const size_t ox = (size_t)(v.tx0 - nx0) * TILE, oy = (size_t)(v.ty0 - ny0) * TILE;
            CK(cudaMemcpy2DAsync(n.mom + (oy * n.W + ox) * REC, (size_t)n.W * REC * 4, v.mom, (size_t)v.W * REC * 4,
                                 (size_t)v.W * REC * 4, v.H, cudaMemcpyDeviceToDevice, st));
            for (int k = 0; k < NUM_DER; ++k)
                CK(cudaMemcpy2DAsync(n.der + k * n.plane() + oy * n.W + ox, (size_t)n.W * 4, v.der + k * v.plane(),
                                     (size_t)v.W * 4, (size_t)v.W * 4, v.H, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpy2DAsync(n.flags + oy * n.W + ox, n.W, v.flags, v.W, v.W, v.H, cudaMemcpyDeviceToDevice, st));
            release(st);  // stream-ordered: freed once the copies above have run
        }
        v = n;
        ++reallocs;
        encode();
    }
    void blank_all(cudaStream_t st)
    {
        const size_t pl = v.plane();
        CK(cudaMemsetAsync(v.mom, 0xFF, pl * REC * sizeof(float), st));
        CK(cudaMemsetAsync(v.der, 0xFF, pl * NUM_DER * sizeof(float), st));
    }
};

// -------------------------------------------------------------------------------------------
// KeyFrame (KeyFrame.hpp): device cloud + pose state machine + the pose it is integrated at
// -------------------------------------------------------------------------------------------
struct KeyFrame
{
    uint64_t id = 0;
    double timestamp = 0;
    uint64_t mapId = 0;
    float4 *d_cloud = nullptr;
    uint32_t n = 0;
    float maxnorm = 0.f;
    bool hasCloud = true;  // false once "dropped" (is_kf_optimization_enabled == false)
    std::mutex poseMutex;
    float pose[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    bool pending = false;
    bool integrated = false;  // has a contribution in the grid (== reference partials non-empty)
    float intPose[12] = {0};
    cudaStream_t allocStream = nullptr;  // stream-ordered allocation: freed without a device-wide sync
    cudaEvent_t ready = nullptr;         // recorded behind the prune's compaction on the ingest stream
    std::atomic<bool> waited{false};
    ~KeyFrame()
    {
        if (ready) cudaEventDestroy(ready);
        if (!d_cloud) return;
        if (allocStream) cudaFreeAsync(d_cloud, allocStream);
        else cudaFree(d_cloud);
    }
    // the first consumer orders its stream behind the compaction; later uses are ordered behind that one
    // (every consumer of one map runs on that map's stream, and a re-parented keyframe has long been complete)
    void waitReady(cudaStream_t consumer)
    {
        if (!ready || waited.load(std::memory_order_acquire)) return;
        cudaStreamWaitEvent(consumer, ready, 0);
        if (cudaEventQuery(ready) == cudaSuccess) waited.store(true, std::memory_order_release);
    }

    void setPose(const float p[12])  // KeyFrame.cpp:29-34
    {
        std::lock_guard<std::mutex> l(poseMutex);
        std::memcpy(pose, p, sizeof(pose));
        pending = true;
    }
    bool getPendingPose(float out[12])  // KeyFrame.cpp:36-44
    {
        std::lock_guard<std::mutex> l(poseMutex);
        if (!pending) return false;
        std::memcpy(out, pose, sizeof(pose));
        pending = false;
        return true;
    }
    void getPose(float out[12])
    {
        std::lock_guard<std::mutex> l(poseMutex);
        std::memcpy(out, pose, sizeof(pose));
    }
};

struct Op
{
    std::shared_ptr<KeyFrame> kf;
    float pose[12];
    float sign;
};

static std::vector<std::pair<int, int>> discOffsets(double radius, double res)  // Helpers.cpp:156-169
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

// -------------------------------------------------------------------------------------------
// LocalMap
// -------------------------------------------------------------------------------------------
class LocalMap
{
  public:
    LocalMap(uint64_t mapID, const tmap_params &P, tmap_update_cb cb, void *user) : mapID_(mapID), P_(P), cb_(cb), user_(user)
    {
        CK(cudaSetDevice(P.device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, P.device));
        nSM_ = prop.multiProcessorCount;
        lat_ = LatticeD{P.grid_center_x, P.grid_center_y, P.resolution};
        disc_ = discOffsets(P.robot_radius, P.resolution);
        if ((int)disc_.size() > MAX_DISC)
            throw Unsupported("robot_radius/resolution gives a footprint of " + std::to_string(disc_.size()) + " cells (limit " +
                              std::to_string(MAX_DISC) + ")");
        r_ = 0;
        for (auto &o : disc_) r_ = std::max(r_, std::max(std::abs(o.first), std::abs(o.second)));
        if (classify_smem_bytes(r_) > 226 * 1024 || r_ > 8)
            throw Unsupported("footprint radius of " + std::to_string(r_) + " cells exceeds the classifier's 32-column halo-row limit (r <= 8)");
        R_ = 0;
        if (P.inflation_enabled)
        {
            auto io = discOffsets(P.inflation_radius, P.resolution);
            int rr = 0;
            for (auto &o : io) rr = std::max(rr, std::max(std::abs(o.first), std::abs(o.second)));
            R_ = (rr + 3) & ~3;  // multiple of 4 cells: the TMA box must start 16-byte aligned in the inner (x) dimension
            nInfl_ = (int)io.size();
            if (inflate_smem_bytes(R_, nInfl_) > 200 * 1024 || TILE + 2 * R_ > 256)
                throw Unsupported("inflation radius of " + std::to_string(rr) + " cells exceeds the halo-tile budget");
            // taps (LocalMap.cpp:57-67) sorted by decreasing decay for the kernel's early exit
            struct Tap { int off; float decay; };
            const int TWi = TILE + 2 * R_;
            std::vector<Tap> taps;
            for (auto &o : io)
            {
                const double dist_m = std::hypot((double)o.first, (double)o.second) * P.resolution;
                taps.push_back({o.second * TWi + o.first, static_cast<float>(std::exp(-P.cost_scaling_factor * dist_m))});
            }
            std::stable_sort(taps.begin(), taps.end(), [](const Tap &a, const Tap &b) { return a.decay > b.decay; });
            decay_.ensure(taps.size() * sizeof(Tap));
            CK(cudaMemcpy(decay_.p, taps.data(), taps.size() * sizeof(Tap), cudaMemcpyHostToDevice));
        }
        static std::atomic<uint64_t> next_serial{1};
        serial_ = next_serial.fetch_add(1);
        cp_.res = P.resolution;
        cp_.ground_clearance = P.ground_clearance;
        cp_.max_slope = P.max_slope;
        cp_.min_points = (uint32_t)P.min_points_per_grid;
        cp_.r = r_;
        cp_.n_disc = (int)disc_.size();
        cp_.infl_r = R_;
        cp_.n_infl = nInfl_;
        cp_.lethal = (float)P.lethal_step_threshold;
        for (auto &w : cp_.row_w) w = -1;
        for (auto &o : disc_)  // rows (dj) of the footprint are contiguous runs of di: keep the half-widths
            cp_.row_w[o.second + r_] = (signed char)std::max<int>(cp_.row_w[o.second + r_], std::abs(o.first));
        CK(cudaStreamCreateWithFlags(&ownStream_, cudaStreamNonBlocking));
        stream_ = ownStream_;
        for (auto &e : ev_) CK(cudaEventCreate(&e));
        grid_.r = r_;
        grid_.R = R_;
        // logical (reference) extent: makeGridMap(half, half) (LocalMap.cpp:73-74, Helpers.cpp:42-47)
        kx_ = static_cast<int>(std::ceil(P.half_size / P.resolution));
        ky_ = kx_;
        grid_.ensure(stream_, floordiv32(-kx_), floordiv32(-ky_), floordiv32(kx_) + 1, floordiv32(ky_) + 1);
        running_ = true;
        worker_ = std::thread(&LocalMap::run, this);
    }
    ~LocalMap()
    {
        {
            std::lock_guard<std::mutex> l(qMutex_);
            running_ = false;
        }
        qCv_.notify_all();
        if (worker_.joinable()) worker_.join();
        cudaStreamSynchronize(stream_);
        shardCommDestroy();
        for (auto &e : ev_) cudaEventDestroy(e);
        cudaStreamDestroy(ownStream_);
    }

    uint64_t id() const { return mapID_; }
    std::recursive_mutex &gridMutex() { return gridMutex_; }

    void addAlreadyDeclaredKF(const std::shared_ptr<KeyFrame> &kf)  // LocalMap.cpp:319-328
    {
        std::lock_guard<std::mutex> l(kfMutex_);
        kf->mapId = mapID_;
        kfs_[kf->id] = kf;
    }
    void setKeyFramePose(const std::shared_ptr<KeyFrame> &kf, const float pose[12])  // LocalMap.cpp:330-341
    {
        if (!kf) return;
        kf->setPose(pose);
        {
            std::lock_guard<std::mutex> l(qMutex_);
            if (queued_.insert(kf->id).second)
                queue_.push_back(kf);
        }
        qCv_.notify_one();
    }
    std::shared_ptr<KeyFrame> detachKeyFrame(uint64_t id)  // LocalMap.cpp:343-373
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        std::shared_ptr<KeyFrame> kf;
        {
            std::lock_guard<std::mutex> l(kfMutex_);
            auto it = kfs_.find(id);
            if (it == kfs_.end()) return nullptr;
            kf = it->second;
        }
        if (kf->integrated)
        {
            std::vector<Op> ops(1);
            ops[0].kf = kf;
            std::memcpy(ops[0].pose, kf->intPose, sizeof(kf->intPose));
            ops[0].sign = -1.f;
            runBatch(ops);
            kf->integrated = false;
        }
        {
            std::lock_guard<std::mutex> l(kfMutex_);
            kfs_.erase(id);
        }
        {
            std::lock_guard<std::mutex> l(qMutex_);
            queued_.erase(id);
        }
        return kf;
    }
    void clearEntireMap()  // LocalMap.cpp:375-397
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        CK(launch_mark_observed_changed(stream_, grid_.v));
        ++g_launches;
        grid_.blank_all(stream_);
        CK(cudaStreamSynchronize(stream_));
        std::lock_guard<std::mutex> l(kfMutex_);
        {
            std::lock_guard<std::mutex> ql(qMutex_);
            for (auto &kv : kfs_)  // ascending kf id (the reference iterates an unordered_map)
            {
                kv.second->integrated = false;
                float p[12];
                kv.second->getPose(p);
                kv.second->setPose(p);
                if (queued_.insert(kv.first).second)
                    queue_.push_back(kv.second);
            }
        }
        qCv_.notify_one();
    }
    size_t numKeyFrames()
    {
        std::lock_guard<std::mutex> l(kfMutex_);
        return kfs_.size();
    }
    size_t queueDepth()
    {
        std::lock_guard<std::mutex> l(qMutex_);
        return queue_.size();
    }
    // Block until the queue is drained.  The caller does not just sleep: while it would wait it takes
    // pending keyframes itself (same locks, same per-pass batch limit as the worker), which removes two
    // thread hand-offs per pass.  on_update then fires on the flushing thread for those keyframes.
    void flush()
    {
        for (;;)
        {
            {
                std::unique_lock<std::mutex> l(qMutex_);
                idleCv_.wait(l, [&] { return !busy_ || !running_; });
                if (!running_ || queue_.empty())
                    return;
                busy_ = true;
            }
            runClaimed();
        }
    }
    void setStream(void *s)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaStreamSynchronize(stream_));
        stream_ = s ? static_cast<cudaStream_t>(s) : ownStream_;
    }

    // ---- queries (caller may hold gridMutex via tmap_lock; the mutex is recursive) ----
    // `cap`: room in the caller's buffer.  A drain that would not fit is refused whole (nothing cleared, the
    // count is still reported), so no changed cell is ever lost to a short buffer.
    std::vector<uint64_t> collectKeys(int mode, bool drain, size_t cap, size_t *total)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        scratchCount_.ensure(16);
        uint32_t *d_count = scratchCount_.as<uint32_t>();
        // first pass: count; second pass: write (+clear)
        CK(launch_collect_keys(stream_, grid_.v, mode, 0, nullptr, 0, d_count));
        uint32_t n = 0;
        CK(cudaMemcpyAsync(&n, d_count, 4, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(4);
        CK(cudaStreamSynchronize(stream_));
        g_launches += 1;
        *total = n;
        if (drain && cap < n)
            return {};
        std::vector<uint64_t> keys(n);
        if (n)
        {
            scratchKeys_.ensure((size_t)n * 8);
            CK(launch_collect_keys(stream_, grid_.v, mode, drain ? 1 : 0, scratchKeys_.as<unsigned long long>(), n, d_count));
            CK(cudaMemcpyAsync(keys.data(), scratchKeys_.p, (size_t)n * 8, cudaMemcpyDeviceToHost, stream_));
            g_d2h += (uint64_t)((size_t)n * 8);
            CK(cudaStreamSynchronize(stream_));
            g_launches += 1;
            std::sort(keys.begin(), keys.end());
        }
        return keys;
    }
    size_t countKeys(int mode)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        scratchCount_.ensure(16);
        uint32_t *d_count = scratchCount_.as<uint32_t>();
        CK(launch_collect_keys(stream_, grid_.v, mode, 0, nullptr, 0, d_count));
        uint32_t n = 0;
        CK(cudaMemcpyAsync(&n, d_count, 4, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(4);
        CK(cudaStreamSynchronize(stream_));
        g_launches += 1;
        return n;
    }
    void readLayers(const uint64_t *keys, size_t n, const int32_t *layers, size_t nl, float *out)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        if (!n || !nl) return;
        scratchKeys_.ensure(n * 8);
        scratchLayers_.ensure(nl * 4);
        scratchOut_.ensure(n * nl * 4);
        CK(cudaMemcpyAsync(scratchKeys_.p, keys, n * 8, cudaMemcpyHostToDevice, stream_));
        g_h2d += (uint64_t)(n * 8);
        CK(cudaMemcpyAsync(scratchLayers_.p, layers, nl * 4, cudaMemcpyHostToDevice, stream_));
        g_h2d += (uint64_t)(nl * 4);
        CK(launch_read_layers(stream_, grid_.v, scratchKeys_.as<unsigned long long>(), n, scratchLayers_.as<int>(), (int)nl, kx_, ky_,
                              scratchOut_.as<float>()));
        CK(cudaMemcpyAsync(out, scratchOut_.p, n * nl * 4, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(n * nl * 4);
        CK(cudaStreamSynchronize(stream_));
        g_launches += 1;
    }
    // fillSparseUpdate (conversions.cpp:28-55): keys outside the grid are skipped (order kept), then one gather
    size_t fillSparseUpdate(const int32_t *layers, size_t nl, const uint64_t *keys, size_t n, uint64_t *out_keys, float *out_values)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        size_t k = 0;
        for (size_t i = 0; i < n; ++i)
        {
            const int ci = (int)(uint32_t)(keys[i] >> 32), cj = (int)(uint32_t)(keys[i] & 0xffffffffu);
            if (std::abs((long)ci) <= kx_ && std::abs((long)cj) <= ky_) out_keys[k++] = keys[i];
        }
        readLayers(out_keys, k, layers, nl, out_values);
        return k;
    }
    void exportDense(int layer, int ci0, int cj0, int w, int h, float *out)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        const size_t tot = (size_t)w * h;
        if (!tot) return;
        scratchOut_.ensure(tot * 4);
        CK(launch_export_dense(stream_, grid_.v, layer, ci0, cj0, w, h, kx_, ky_, scratchOut_.as<float>()));
        CK(cudaMemcpyAsync(out, scratchOut_.p, tot * 4, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(tot * 4);
        CK(cudaStreamSynchronize(stream_));
        g_launches += 1;
    }
    void exportOccupancy(int ci0, int cj0, int w, int h, int8_t *out)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        const size_t tot = (size_t)w * h;
        if (!tot) return;
        scratchOut_.ensure(tot);
        CK(launch_export_occupancy(stream_, grid_.v, ci0, cj0, w, h, kx_, ky_, scratchOut_.as<signed char>()));
        CK(cudaMemcpyAsync(out, scratchOut_.p, tot, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(tot);
        CK(cudaStreamSynchronize(stream_));
        g_launches += 1;
    }
    // publishCompleteness (global_traversability_node.cpp:381-459); out holds ceil(w/f) x ceil(h/f) bytes
    void exportCompleteness(int ci0, int cj0, int w, int h, double publish_res, int8_t *out, size_t cap, int32_t *cw_out, int32_t *ch_out)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        const float fine_res = (float)P_.resolution;  // nav_msgs MapMetaData::resolution is float32
        int f = (int)std::lround(publish_res / fine_res);
        if (f < 1) f = 1;
        const int cw = (w + f - 1) / f, ch = (h + f - 1) / f;
        *cw_out = cw; *ch_out = ch;
        const size_t tot = (size_t)cw * ch;
        if (!out || !tot) return;
        if (cap < tot) throw std::invalid_argument("completeness output buffer too small");
        scratchOut_.ensure(tot);
        CK(launch_export_completeness(stream_, grid_.v, ci0, cj0, w, h, kx_, ky_, f, cw, ch, scratchOut_.as<signed char>()));
        CK(cudaMemcpyAsync(out, scratchOut_.p, tot, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(tot);
        CK(cudaStreamSynchronize(stream_));
        g_launches += 1;
    }
    void extent(int32_t *kx, int32_t *ky)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        *kx = kx_;
        *ky = ky_;
    }
    size_t stitched(float *out, size_t cap)  // LocalMap.cpp:416-437 (no voxel filter)
    {
        std::vector<std::shared_ptr<KeyFrame>> snap;
        {
            std::lock_guard<std::mutex> l(kfMutex_);
            for (auto &kv : kfs_) snap.push_back(kv.second);
        }
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        size_t total = 0;
        for (auto &kf : snap)
            if (kf->hasCloud) total += kf->n;
        if (!out) return total;
        size_t off = 0;
        scratchLayers_.ensure(48);
        for (auto &kf : snap)
        {
            if (!kf->hasCloud || off + kf->n > cap) continue;
            float p[12];
            kf->getPose(p);
            kf->waitReady(stream_);
            scratchOut_.ensure((size_t)kf->n * 12);
            CK(cudaMemcpyAsync(scratchLayers_.p, p, 48, cudaMemcpyHostToDevice, stream_));
            g_h2d += (uint64_t)(48);
            CK(launch_transform_cloud(stream_, kf->d_cloud, kf->n, scratchLayers_.as<float>(), scratchOut_.as<float>()));
            CK(cudaMemcpyAsync(out + off * 3, scratchOut_.p, (size_t)kf->n * 12, cudaMemcpyDeviceToHost, stream_));
            g_d2h += (uint64_t)kf->n * 12;
            CK(cudaStreamSynchronize(stream_));
            g_launches += 1;
            off += kf->n;
        }
        return total;
    }
    // getStitchedPointCloud(vx, vy, vz) with positive leaf sizes (LocalMap.cpp:438-446): pcl::VoxelGrid centroids
    size_t stitchedVoxel(float lx, float ly, float lz, float *out, size_t cap)
    {
        if (!(lx > 0.f && ly > 0.f && lz > 0.f)) return stitched(out, cap);
        std::vector<std::shared_ptr<KeyFrame>> snap;
        {
            std::lock_guard<std::mutex> l(kfMutex_);
            for (auto &kv : kfs_) snap.push_back(kv.second);
        }
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        size_t total = 0;
        for (auto &kf : snap)
            if (kf->hasCloud) total += kf->n;
        if (!out || total == 0) return total;  // upper bound for sizing
        if (total >= (1ull << 31)) throw Unsupported("stitched cloud exceeds 2^31 points");
        const uint32_t n = (uint32_t)total;
        voxXyz_.ensure((size_t)n * 12);
        scratchLayers_.ensure(48);
        size_t off = 0;
        for (auto &kf : snap)
        {
            if (!kf->hasCloud) continue;
            float p[12];
            kf->getPose(p);
            kf->waitReady(stream_);
            CK(cudaMemcpyAsync(scratchLayers_.p, p, 48, cudaMemcpyHostToDevice, stream_));
            g_h2d += (uint64_t)(48);
            CK(launch_transform_cloud(stream_, kf->d_cloud, kf->n, scratchLayers_.as<float>(), voxXyz_.as<float>() + off * 3));
            CK(cudaStreamSynchronize(stream_));  // the pose staging buffer is reused by the next keyframe
            g_launches += 1;
            off += kf->n;
        }
        voxIdx_.ensure((size_t)n * 4 * 6 + 64);
        uint32_t *keys = voxIdx_.as<uint32_t>(), *vals = keys + 2 * (size_t)n, *head = vals + 2 * (size_t)n, *ordinal = head + n;
        uint32_t *mm = ordinal + n;
        CK(vox_minmax(stream_, voxXyz_.as<float>(), n, mm));
        uint32_t mmh[6];
        CK(cudaMemcpyAsync(mmh, mm, sizeof(mmh), cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(sizeof(mmh));
        CK(cudaStreamSynchronize(stream_));
        // PCL 1.12 voxel_grid.hpp:246-291 in its own float / int arithmetic
        const float leaf[3] = {lx, ly, lz};
        float inv[3], mn[3], mx[3];
        int min_b[3], max_b[3], div_b[3], mul[3];
        int64_t d[3];
        for (int k = 0; k < 3; ++k)
        {
            inv[k] = 1.0f / leaf[k];
            mn[k] = vox_ordered_to_float(mmh[k]);
            mx[k] = vox_ordered_to_float(mmh[3 + k]);
            d[k] = (int64_t)((mx[k] - mn[k]) * inv[k]) + 1;
        }
        size_t n_out;
        if (d[0] * d[1] * d[2] > (int64_t)std::numeric_limits<int32_t>::max())
        {
            // "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input
            n_out = n;
            CK(cudaMemcpyAsync(out, voxXyz_.p, std::min<size_t>(n, cap) * 12, cudaMemcpyDeviceToHost, stream_));
            g_d2h += (uint64_t)(std::min<size_t>(n, cap) * 12);
            CK(cudaStreamSynchronize(stream_));
            return n_out;
        }
        for (int k = 0; k < 3; ++k)
        {
            min_b[k] = (int)std::floor(mn[k] * inv[k]);
            max_b[k] = (int)std::floor(mx[k] * inv[k]);
            div_b[k] = max_b[k] - min_b[k] + 1;
        }
        mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
        const size_t tb = vox_temp_bytes(n);
        voxTemp_.ensure(tb + 16);
        voxOut_.ensure((size_t)n * 12);
        uint32_t nv = 0;
        CK(vox_filter(stream_, voxXyz_.as<float>(), n, inv, min_b, mul, keys, vals, head, ordinal, voxTemp_.p, tb, voxOut_.as<float>(), &nv));
        g_launches += 5;
        CK(cudaMemcpyAsync(out, voxOut_.p, std::min<size_t>(nv, cap) * 12, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(std::min<size_t>(nv, cap) * 12);
        CK(cudaStreamSynchronize(stream_));
        return nv;
    }
    void getStats(tmap_stats *s)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        *s = stats_;
        s->kernel_launches = g_launches.load();
        s->h2d_bytes = g_h2d.load();
        s->d2h_bytes = g_d2h.load();
        s->grid_reallocs = grid_.reallocs;
    }

  private:
    // ---- the worker (LocalMap::RunLocalKeyFrames, LocalMap.cpp:452-485) ----
    void popLocked(std::vector<std::shared_ptr<KeyFrame>> &take)  // qMutex_ held
    {
        const size_t lim = P_.max_batch <= 0 ? queue_.size() : std::min<size_t>(queue_.size(), (size_t)P_.max_batch);
        for (size_t i = 0; i < lim; ++i)
        {
            take.push_back(queue_.front());
            queue_.pop_front();
            queued_.erase(take.back()->id);
        }
    }
    // busy_ already raised by the caller.  The keyframes are taken from the queue only once the grid mutex is
    // held: whatever a caller queued while it held the mutex (a PGO message, informLoopClosure's re-enqueue of
    // every keyframe) is then folded in ONE pass instead of being split at whatever the worker had popped before
    // it blocked on the mutex.  Same outcome as the reference's pop-then-lock (LocalMap.cpp:452-485) -- latest
    // pose wins, every pending keyframe is recomputed once -- with fewer passes.
    void runClaimed()
    {
        int changed = 0;
        try
        {
            std::lock_guard<std::recursive_mutex> g(gridMutex_);
            std::vector<std::shared_ptr<KeyFrame>> take;
            {
                std::lock_guard<std::mutex> l(qMutex_);
                popLocked(take);
            }
            changed = processTaken(take);
        }
        catch (const std::exception &e)
        {
            std::lock_guard<std::mutex> l(qMutex_);
            workerError_ = e.what();
        }
        for (int i = 0; i < changed; ++i)
            if (cb_) cb_(user_);  // onUpdate_, grid lock released (LocalMap.cpp:479-481)
        {
            std::lock_guard<std::mutex> l(qMutex_);
            busy_ = false;
        }
        idleCv_.notify_all();
        qCv_.notify_one();
    }
    void run()
    {
        cudaSetDevice(P_.device);
        for (;;)
        {
            {
                std::unique_lock<std::mutex> l(qMutex_);
                qCv_.wait(l, [&] { return (!queue_.empty() && !busy_) || !running_; });
                if (!running_) break;
                busy_ = true;
            }
            runClaimed();
        }
        idleCv_.notify_all();
    }
    int processTaken(const std::vector<std::shared_ptr<KeyFrame>> &take)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        std::vector<Op> ops;
        std::vector<std::shared_ptr<KeyFrame>> done;
        for (auto &kf : take)
        {
            float pose[12];
            if (!kf->getPendingPose(pose)) continue;  // LocalMap.cpp:473-475
            {
                std::lock_guard<std::mutex> l(kfMutex_);  // deleted since enqueued? (LocalMap.cpp:248-256)
                if (kfs_.find(kf->id) == kfs_.end()) continue;
            }
            if (!kf->hasCloud) continue;  // LocalMap.cpp:260-267
            if (kf->integrated)
            {
                Op s;
                s.kf = kf;
                std::memcpy(s.pose, kf->intPose, sizeof(s.pose));
                s.sign = -1.f;
                ops.push_back(s);
            }
            Op a;
            a.kf = kf;
            std::memcpy(a.pose, pose, sizeof(a.pose));
            a.sign = +1.f;
            ops.push_back(a);
            done.push_back(kf);
        }
        if (ops.empty()) return 0;
        runBatch(ops);
        for (auto &op : ops)
            if (op.sign > 0)
            {
                op.kf->integrated = true;
                std::memcpy(op.kf->intPose, op.pose, sizeof(op.pose));
            }
        if (!P_.is_kf_optimization_enabled)
            for (auto &kf : done) kf->hasCloud = false;  // dropCloud (LocalMap.cpp:311-312)
        stats_.keyframes_processed += done.size();
        return (int)done.size();
    }

    // ---- one device pass over a list of ops (LocalMap::recomputeKeyFrame steps 1-4) ----
    std::chrono::steady_clock::time_point t_planned_{};
    struct Plan
    {
        BatchView b{};
        uint32_t n_ops = 0, maxT = 0;
        int ppt = 2, dc = 0, di = -1, dm = 0;
        uint64_t total_pts = 0;
        int ux0 = 0, uy0 = 0, ux1 = 0, uy1 = 0;  // union of op windows, bins, inclusive
    };

    // op windows: |x_m - t_x| <= |row_x| * max|p_base| (Cauchy-Schwarz) + float-rounding slack, in bins
    void opWindow(const Op &op, int &bx0, int &by0, int &bx1, int &by1) const
    {
        const double res = P_.resolution;
        for (int k = 0; k < 12; ++k)
            if (!std::isfinite(op.pose[k])) throw std::invalid_argument("non-finite pose");
        const double mn = (double)op.kf->maxnorm;
        double lo[2], hi[2];
        for (int a = 0; a < 2; ++a)
        {
            const float *rw = op.pose + 4 * a;
            const double rn = std::sqrt((double)rw[0] * rw[0] + (double)rw[1] * rw[1] + (double)rw[2] * rw[2]);
            const double t = rw[3];
            const double rad = rn * mn;
            const double slack = 1e-5 * (std::abs(t) + rad) + 2.0 * res;
            lo[a] = t - rad - slack;
            hi[a] = t + rad + slack;
        }
        const double qx0 = std::floor((lo[0] - lat_.x0) / res) - 1, qx1 = std::ceil((hi[0] - lat_.x0) / res) + 1;
        const double qy0 = std::floor((lo[1] - lat_.y0) / res) - 1, qy1 = std::ceil((hi[1] - lat_.y0) / res) + 1;
        if (!(qx0 > -1e9 && qx1 < 1e9 && qy0 > -1e9 && qy1 < 1e9))
            throw std::invalid_argument("keyframe window out of the 32-bit cell range");
        bx0 = floordiv16((int)qx0); by0 = floordiv16((int)qy0);
        bx1 = floordiv16((int)qx1); by1 = floordiv16((int)qy1);
    }

    // host side of a pass: op descriptors, chunk table, scratch buffers, uploads.  `forced` (bins, inclusive,
    // already including the candidate margin) pins the per-bin arrays to a bbox shared by all ranks.
    // a bin is split across CTAs (by cell-row group) once it exceeds 1/splitDiv of a resident fold CTA's fair share
    static int splitDiv()
    {
        static const int d = std::getenv("TMB_SPLIT_DIV") ? std::max(1, std::atoi(std::getenv("TMB_SPLIT_DIV"))) : 1;
        return d;
    }
    Plan planBatch(const std::vector<Op> &ops, const int *forced, uint32_t op_base, bool ensure_grid = true)
    {
        Plan pl;
        const uint32_t n_ops = (uint32_t)ops.size();
        pl.n_ops = n_ops;
        uint64_t total_pts = 0, chunks8 = 0, chunks32 = 0;
        for (auto &op : ops)
        {
            total_pts += op.kf->n;
            chunks8 += (op.kf->n + BIN_THREADS * 8 - 1) / (BIN_THREADS * 8);
            chunks32 += (op.kf->n + BIN_THREADS * 32 - 1) / (BIN_THREADS * 32);
        }
        pl.total_pts = total_pts;
        // points per thread and chunk: the largest tier that still hands every SM several chunks (small batches use
        // small chunks so the grid fills the GPU; big ones amortise the per-chunk histogram row over 8 192 points)
        pl.ppt = chunks32 >= (uint64_t)8 * nSM_ ? 32 : chunks8 >= (uint64_t)2 * nSM_ ? 8 : 2;
        const uint32_t CH = BIN_THREADS * pl.ppt;
        if (total_pts >= (1ull << 32))
            throw Unsupported("batch exceeds 2^32 points");
        if (n_ops > 65535 || (uint64_t)n_ops + op_base >= (1u << 23))
            throw Unsupported("batch exceeds 65535 ops");

        // one pinned staging block per pass: [counters 64 B][per-op bbox][op descriptors][chunk table][op groups]
        // -> ONE host-to-device copy (it also zeroes the counters and initialises the bboxes), and the first
        //    two sections come back with ONE device-to-host copy at the end of the pass
        uint32_t n_chunks_all = 0;
        for (auto &op : ops) n_chunks_all += (op.kf->n + CH - 1) / CH;
        const uint32_t n_groups = (n_ops + 31) / 32;
        const size_t off_bbox = 64;
        const size_t off_ops = (off_bbox + (size_t)n_ops * sizeof(int4) + 15) & ~(size_t)15;
        const size_t ops_end = off_ops + (size_t)n_ops * sizeof(OpDesc);
        const size_t off_chunk = (ops_end + 15) & ~(size_t)15;
        const size_t off_grp = (off_chunk + (size_t)n_chunks_all * 4 + 15) & ~(size_t)15;
        const size_t meta_bytes = off_grp + (size_t)n_groups * sizeof(int4);
        metaH_.ensure(meta_bytes + 64);
        OpDesc *od = reinterpret_cast<OpDesc *>(metaH_.as<unsigned char>() + off_ops);
        uint32_t n_chunks = 0, tot_off = 0, maxT = 0;
        unsigned long long hist_off = 0;
        int ux0 = INT_MAX, uy0 = INT_MAX, ux1 = INT_MIN, uy1 = INT_MIN;
        for (uint32_t i = 0; i < n_ops; ++i)
        {
            const Op &op = ops[i];
            OpDesc &d = od[i];
            std::memset(&d, 0, sizeof(d));
            op.kf->waitReady(stream_);
            d.cloud = op.kf->d_cloud;
            d.n = op.kf->n;
            d.sign = op.sign;
            std::memcpy(d.T, op.pose, sizeof(d.T));
            int bx0, by0, bx1, by1;
            opWindow(op, bx0, by0, bx1, by1);
            d.wtx0 = bx0; d.wty0 = by0;
            d.wtw = bx1 - bx0 + 1; d.wth = by1 - by0 + 1;
            const uint64_t T = (uint64_t)d.wtw * d.wth;
            if (T > MAX_WINDOW_BINS)
                throw Unsupported("keyframe spans " + std::to_string(T) + " bins of 16x16 cells (sensor range / resolution too large; limit " +
                                  std::to_string(MAX_WINDOW_BINS) + ")");
            d.T_tiles = (uint32_t)T;
            maxT = std::max(maxT, d.T_tiles);
            d.chunk0 = n_chunks;
            d.nchunks = (d.n + CH - 1) / CH;
            n_chunks += d.nchunks;
            d.tot_off = tot_off;
            tot_off += d.T_tiles;
            d.hist_off = hist_off;
            hist_off += (unsigned long long)d.nchunks * d.T_tiles;
            ux0 = std::min(ux0, bx0); uy0 = std::min(uy0, by0);
            ux1 = std::max(ux1, bx1); uy1 = std::max(uy1, by1);
        }
        pl.maxT = maxT;
        pl.dc = (r_ + BIN - 1) / BIN;                       // candidate dilation, in bins
        pl.di = R_ > 0 ? (r_ + R_ + BIN - 1) / BIN : -1;
        pl.dm = std::max(pl.dc, pl.di);
        BatchView &b = pl.b;
        if (forced)
        {
            if (n_ops && (ux0 < forced[0] || uy0 < forced[1] || ux1 > forced[2] || uy1 > forced[3]))
                throw std::invalid_argument("shard bbox does not cover this rank's keyframe windows");
            b.btx0 = forced[0]; b.bty0 = forced[1];
            b.btw = forced[2] - forced[0] + 1; b.bth = forced[3] - forced[1] + 1;
            ux0 = forced[0]; uy0 = forced[1]; ux1 = forced[2]; uy1 = forced[3];
        }
        else
        {
            b.btx0 = ux0 - pl.dm; b.bty0 = uy0 - pl.dm;
            b.btw = ux1 - ux0 + 1 + 2 * pl.dm; b.bth = uy1 - uy0 + 1 + 2 * pl.dm;
        }
        pl.ux0 = ux0; pl.uy0 = uy0; pl.ux1 = ux1; pl.uy1 = uy1;
        if (ensure_grid)
            grid_.ensure(stream_, ux0 >> 1, uy0 >> 1, (ux1 >> 1) + 1, (uy1 >> 1) + 1);  // bins -> 32x32 tiles

        b.n_ops = n_ops;
        b.n_chunks = n_chunks;
        b.op_base = op_base;
        const size_t nb = (size_t)b.btw * b.bth;
        if (nb >= (1u << 24))
            throw Unsupported("batch spans more than 2^24 bins of 16x16 cells");
        b.total_points = (uint32_t)total_pts;
        b.cand_y0 = 0; b.cand_y1 = b.bth;
        // a bin is split across CTAs only when it exceeds the fair share of one resident fold CTA (3 per SM)
        b.split_thresh = (uint32_t)std::max<uint64_t>(32768, total_pts / (uint64_t)(3 * nSM_ * splitDiv()));

        // chunk -> op table, followed by the per-32-op union windows used by the scan over ops
        std::memset(metaH_.p, 0, 64);
        int4 *bb0 = reinterpret_cast<int4 *>(metaH_.as<unsigned char>() + off_bbox);
        for (uint32_t i = 0; i < n_ops; ++i) bb0[i] = make_int4(INT_MAX, INT_MIN, INT_MAX, INT_MIN);
        uint32_t *ch = reinterpret_cast<uint32_t *>(metaH_.as<unsigned char>() + off_chunk);
        int4 *grp = reinterpret_cast<int4 *>(metaH_.as<unsigned char>() + off_grp);
        for (uint32_t gi = 0; gi < n_groups; ++gi)
        {
            int4 w = make_int4(INT_MAX, INT_MAX, INT_MIN, INT_MIN);
            for (uint32_t i = gi * 32; i < std::min(n_ops, gi * 32 + 32); ++i)
            {
                w.x = std::min(w.x, od[i].wtx0); w.y = std::min(w.y, od[i].wty0);
                w.z = std::max(w.z, od[i].wtx0 + od[i].wtw - 1); w.w = std::max(w.w, od[i].wty0 + od[i].wth - 1);
            }
            grp[gi] = w;
        }
        for (uint32_t i = 0; i < n_ops; ++i)
            for (uint32_t c = 0; c < od[i].nchunks; ++c) ch[od[i].chunk0 + c] = i;

        metaD_.ensure(meta_bytes + 64);
        histD_.ensure((size_t)hist_off * 4 + 4);
        totD_.ensure((size_t)tot_off * 8 + 8);
        const size_t n_t32 = (size_t)grid_.v.tw * grid_.v.th;
        tileD_.ensure((nb * 6 + (nb + 255) / 256 + n_t32) * 4 + 256);
        sortedD_.ensure((size_t)total_pts * 16 + 16);
        {
            // inflation-tile dedup stamps: own buffer (persistent across passes), zeroed when it grows
            const size_t need = n_t32 * 4 + 64;
            if (need > inflFlagsD_.cap)
            {
                inflFlagsD_.ensure(need);
                CK(cudaMemsetAsync(inflFlagsD_.p, 0, inflFlagsD_.cap, stream_));
                epoch_ = 0;
            }
            if (++epoch_ == 0)
            {
                CK(cudaMemsetAsync(inflFlagsD_.p, 0, inflFlagsD_.cap, stream_));
                epoch_ = 1;
            }
        }
        {
            CK(cudaMemcpyAsync(metaD_.p, metaH_.p, meta_bytes, cudaMemcpyHostToDevice, stream_));
            g_h2d += (uint64_t)(meta_bytes);
        }
        unsigned char *md = metaD_.as<unsigned char>();
        b.ops = reinterpret_cast<const OpDesc *>(md + off_ops);
        b.chunk_op = reinterpret_cast<const uint32_t *>(md + off_chunk);
        b.op_groups = reinterpret_cast<const int4 *>(md + off_grp);
        b.hist = histD_.as<uint32_t>();
        b.op_total = totD_.as<uint32_t>();
        b.op_start = totD_.as<uint32_t>() + tot_off;
        b.tile_count = tileD_.as<uint32_t>();
        b.tile_start = b.tile_count + nb;
        b.active_tiles = b.tile_start + nb;
        b.cand_tiles = b.active_tiles + nb;
        b.fit_tiles = b.cand_tiles + nb;
        b.min_points = (uint32_t)std::max(0, P_.min_points_per_grid);
        globD_ = b.fit_tiles + nb;                  // sharded runs: global per-bin totals
        b.block_sums = globD_ + nb;
        b.infl_tiles = b.block_sums + (nb + 255) / 256;
        b.infl_flags = inflFlagsD_.as<uint32_t>();
        b.epoch = epoch_;
        b.nbr_count = b.tile_count;
        b.counters = reinterpret_cast<uint32_t *>(md);
        b.op_bbox = reinterpret_cast<int4 *>(md + off_bbox);
        b.sorted = sortedD_.as<float4>();
        return pl;
    }

    // classify + inflate + clear of the bins listed by the bin phase, statistics, logical growth
    void finishBatch(const Plan &pl, const std::vector<Op> &ops, const std::chrono::steady_clock::time_point &t_begin,
                     const std::function<void()> &before_inflate = nullptr)
    {
        const BatchView &b = pl.b;
        const size_t nb = (size_t)b.btw * b.bth;
        const bool prof = P_.profile != 0;
        {
            // the footprint tables live in __constant__ memory (one copy per device, shared by every map of the process):
            // the check / upload and the enqueue of the kernels that read them are one critical section, so another map
            // with a different footprint cannot slip its tables in between (its upload first waits for the device)
            std::lock_guard<std::mutex> cl(constMutex());
            uploadConstantsLocked();
            CK(launch_classify(stream_, grid_.tmap_mom, grid_.tmap_flags, b, grid_.v, cp_, nSM_, prof ? ev_[8] : nullptr));
        }
        g_launches += 2;  // k_gate + k_classify
        if (prof) CK(cudaEventRecord(ev_[5], stream_));
        if (R_ > 0 && before_inflate)
            before_inflate();  // sharded map: the neighbours' step values R cells beyond the slab arrive here
        if (R_ > 0)
        {
            CK(launch_inflate(stream_, grid_.tmap_step, b, grid_.v, cp_, decay_.p, nSM_));
            g_launches += 1;
        }
        if (prof) CK(cudaEventRecord(ev_[6], stream_));
        CK(launch_clear_touched(stream_, b, grid_.v, (int)std::min<size_t>((size_t)nSM_ * 8, nb)));
        g_launches += 1;
        if (prof) CK(cudaEventRecord(ev_[7], stream_));

        // results the host needs: counters + per-op bounding boxes
        const uint32_t n_ops = pl.n_ops;
        const size_t res_bytes = 64 + (size_t)n_ops * sizeof(int4);
        resH_.ensure(res_bytes + 64);
        const auto t_launched = std::chrono::steady_clock::now();
        CK(cudaMemcpyAsync(resH_.p, metaD_.p, res_bytes, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(res_bytes);
        CK(cudaStreamSynchronize(stream_));
        const auto t_synced = std::chrono::steady_clock::now();
        const uint32_t *cnt = resH_.as<uint32_t>();
        if (cnt[3])
            throw CudaError("binning consistency check failed (flags " + std::to_string(cnt[3]) + ")");

        // logical grid growth, one keyframe at a time (growToIncludeCells, LocalMap.cpp:91-106 +
        // growGridToInclude, Helpers.cpp:53-84)
        const int4 *bb = reinterpret_cast<const int4 *>(resH_.as<unsigned char>() + 64);
        for (uint32_t i = 0; i < n_ops && i < ops.size(); ++i)
            if (ops[i].sign > 0 && bb[i].x != INT_MAX)
                growLogical(bb[i].x, bb[i].y, bb[i].z, bb[i].w);

        stats_.host_ms_batch = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        if (t_planned_ >= t_begin)  // single-GPU batches (the sharded phases are timed by their caller)
        {
            stats_.sum_host_ms_plan += std::chrono::duration<double, std::milli>(t_planned_ - t_begin).count();
            stats_.sum_host_ms_launch += std::chrono::duration<double, std::milli>(t_launched - t_planned_).count();
            stats_.sum_host_ms_wait += std::chrono::duration<double, std::milli>(t_synced - t_launched).count();
        }
        stats_.sum_host_ms_batch += stats_.host_ms_batch;
        stats_.batches += 1;
        stats_.points_binned += pl.total_pts;
        stats_.cells_classified += cnt[7];
        stats_.last_batch_points = pl.total_pts;
        stats_.last_batch_tiles = cnt[10] + cnt[11];
        stats_.last_batch_cells_touched = cnt[8];
        stats_.last_batch_cells_classified = cnt[7];
        stats_.last_batch_max_bin = cnt[9];
        if (prof)
        {
            float t;
            cudaEventElapsedTime(&t, ev_[0], ev_[1]); stats_.ms_hist = t;
            cudaEventElapsedTime(&t, ev_[1], ev_[2]); stats_.ms_scan = t;
            cudaEventElapsedTime(&t, ev_[2], ev_[3]); stats_.ms_scatter = t;
            cudaEventElapsedTime(&t, ev_[3], ev_[4]); stats_.ms_fold = t;
            cudaEventElapsedTime(&t, ev_[4], ev_[5]); stats_.ms_classify = t;
            cudaEventElapsedTime(&t, ev_[4], ev_[8]); stats_.ms_gate = t; stats_.sum_ms_gate += t;
            cudaEventElapsedTime(&t, ev_[5], ev_[6]); stats_.ms_inflate = t;
            cudaEventElapsedTime(&t, ev_[0], ev_[7]); stats_.ms_total = t;
            stats_.sum_ms_hist += stats_.ms_hist; stats_.sum_ms_scan += stats_.ms_scan; stats_.sum_ms_scatter += stats_.ms_scatter;
            stats_.sum_ms_fold += stats_.ms_fold; stats_.sum_ms_classify += stats_.ms_classify;
            stats_.sum_ms_inflate += stats_.ms_inflate; stats_.sum_ms_total += stats_.ms_total;
        }
        stats_.sum_infl_tiles += cnt[2];
        stats_.sum_cand_bins += cnt[1];
        stats_.sum_tiles += cnt[10] + cnt[11];
        stats_.sum_cells_touched += cnt[8];
    }

    // ---- one device pass over a list of ops (LocalMap::recomputeKeyFrame steps 1-4) ----
    void runBatch(const std::vector<Op> &ops)
    {
        const auto t_begin = std::chrono::steady_clock::now();
        CK(cudaSetDevice(P_.device));
        Plan pl = planBatch(ops, nullptr, 0);
        t_planned_ = std::chrono::steady_clock::now();
        CK(launch_bin(stream_, pl.b, grid_.v, lat_, pl.ppt, pl.maxT, pl.dc, pl.di, nSM_, P_.profile ? ev_ : nullptr, 3));
        g_launches += ((size_t)pl.b.btw * pl.b.bth <= 4096) ? 5 : 7;  // hist, scan kernels, scatter, fold
        static const bool atomic_experiment = std::getenv("TMB_EXPERIMENT_ATOMIC_FOLD") != nullptr;  // profiles/README.md §6
        if (atomic_experiment)
            runAtomicExperiment(pl);
        finishBatch(pl, ops, t_begin);
    }

    // profiles/README.md "sort vs shared atomics": times the atomic alternative on the batch just folded and counts
    // the cells whose float32 moments differ from the product's (meaningful after a rebuild from a blank map)
    void runAtomicExperiment(const Plan &pl)
    {
        expScratch_.ensure(grid_.v.plane() * REC * 4 + 64);
        expCount_.ensure(64);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaMemsetAsync(expCount_.p, 0, 4, stream_));
        CK(cudaEventRecord(e0, stream_));
        CK(launch_fold_atomic_experiment(stream_, pl.b, grid_.v, lat_, expScratch_.as<float>(), expCount_.as<uint32_t>(), nSM_));
        CK(cudaEventRecord(e1, stream_));
        uint32_t bad = 0;
        CK(cudaMemcpyAsync(&bad, expCount_.p, 4, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)(4);
        CK(cudaStreamSynchronize(stream_));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        std::fprintf(stderr, "[tmap experiment] atomic fold: %.3f ms for %llu records, %u cells differ from the sorted fold\n", ms,
                     (unsigned long long)pl.total_pts, bad);
    }

  public:
    // ---- sharded rebuild: this process owns a slab of bin rows of a global map -------------
    void shardBlankRows(int cj0, int n_rows)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        const GridView &v = grid_.v;
        if (!v.mom) return;
        const long r0 = std::max<long>(0, (long)cj0 - v.cj0), r1 = std::min<long>(v.H, (long)cj0 + n_rows - v.cj0);
        if (r1 <= r0) return;
        const size_t off = (size_t)r0 * v.W, cnt = (size_t)(r1 - r0) * v.W;
        CK(cudaMemsetAsync(v.mom + off * REC, 0xFF, cnt * REC * sizeof(float), stream_));
        for (int k = 0; k < NUM_DER; ++k)
            CK(cudaMemsetAsync(v.der + (size_t)k * v.plane() + off, 0xFF, cnt * sizeof(float), stream_));
        CK(cudaMemsetAsync(v.flags + off, 0, cnt, stream_));
    }
    // key pass + scans of this rank's ops over the shared bbox: per-bin counts and every record's slot in this rank's
    // (virtual) bucket array; nothing is written yet and the grid is not touched
    void shardKeyPass(const std::vector<Op> &ops, uint32_t op_base, const int32_t bbox[4])
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        shardOps_ = ops;
        const int forced[4] = {bbox[0], bbox[1], bbox[2], bbox[3]};
        shardPlan_ = planBatch(shardOps_, forced, op_base, false);
        CK(launch_bin(stream_, shardPlan_.b, grid_.v, lat_, shardPlan_.ppt, std::max(shardPlan_.maxT, 1u), shardPlan_.dc, shardPlan_.di,
                      nSM_, nullptr, 5));
        g_launches += 6;
    }
    // fold the records received from every rank for bin rows [y0, y1) (bbox-local): per bin, the sources' segments in rank order
    void shardFold(const void *d_recv, uint64_t n_recv, const void *d_counts_all, const void *d_starts_all, const uint64_t *src_base,
                   int world, int y0, int y1, bool sync = true)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        CK(cudaSetDevice(P_.device));
        BatchView &b = shardPlan_.b;
        if (y0 < 0 || y1 > b.bth || y0 > y1) throw std::invalid_argument("bad slab rows");
        if (world > 32) throw Unsupported("sharded fold: at most 32 ranks");
        srcBaseD_.ensure((size_t)world * 8);
        CK(cudaMemcpyAsync(srcBaseD_.p, src_base, (size_t)world * 8, cudaMemcpyHostToDevice, stream_));
        g_h2d += (uint64_t)((size_t)world * 8);
        b.split_thresh = (uint32_t)std::max<uint64_t>(32768, n_recv / (uint64_t)(3 * nSM_ * splitDiv()));
        b.n_ops = 0;  // per-op bboxes were consumed after the key pass
        b.nbr_count = globD_;
        b.cand_y0 = y0; b.cand_y1 = y1;
        // a bin's bucket = the sources' segments in rank order, read in place from the receive buffer by the fold
        b.sh_recv = static_cast<const float4 *>(d_recv);
        b.sh_counts = static_cast<const uint32_t *>(d_counts_all);
        b.sh_starts = static_cast<const uint32_t *>(d_starts_all);
        b.sh_src_base = srcBaseD_.as<unsigned long long>();
        b.sh_world = world;
        b.sh_first = (uint32_t)y0 * (uint32_t)b.btw;
        if (R_ > 0)
        {
            // the key pass / scatter of this pass may already have stamped inflation tiles with the batch's epoch (their
            // work lists are discarded): the merged counts below list the tiles afresh
            if (++epoch_ == 0)
            {
                CK(cudaMemsetAsync(inflFlagsD_.p, 0, inflFlagsD_.cap, stream_));
                epoch_ = 1;
            }
            b.epoch = epoch_;
        }
        CK(launch_shard_merge(stream_, b, grid_.v, static_cast<const float4 *>(d_recv), static_cast<const uint32_t *>(d_counts_all),
                              static_cast<const uint32_t *>(d_starts_all), srcBaseD_.as<unsigned long long>(), world, y0, y1, globD_,
                              shardPlan_.dc, shardPlan_.di));
        CK(launch_bin(stream_, b, grid_.v, lat_, shardPlan_.ppt, 1, shardPlan_.dc, shardPlan_.di, nSM_, nullptr, 2));
        g_launches += 6;
        if (sync) CK(cudaStreamSynchronize(stream_));
    }
    // ---- the whole sharded rebuild behind ONE call, collectives issued from here (NCCL on the map's stream) ----
    struct ShardInfo
    {
        int32_t bbox[4];            // btx0, bty0, btw, bth (bins) of the job-wide bbox
        int32_t slab_y0, slab_y1;   // this rank's bin rows [y0, y1), bbox-local
        int32_t cj0, cj1;           // the same in absolute cell rows
        uint64_t local_records, sent_records, recv_records;
        uint64_t halo_bytes, grid_bytes;
        float ms[8];                // bin, plan (gathers + host), alltoall, fold, halo, classify, total, -
        int32_t n_slabs;
        int32_t slabs[2 * 64];
    };
    static void ncclCheck(ncclResult_t r, const char *what)
    {
        if (r != ncclSuccess)
            throw CudaError(std::string(what) + ": " + NcclApi::get().GetErrorString(r));
    }
    void shardCommInit(int rank, int world, const uint8_t *id128)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        NcclApi &N = NcclApi::get();
        if (!N.ok()) throw Unsupported(N.error);
        if (world < 1 || world > 64 || rank < 0 || rank >= world) throw std::invalid_argument("bad rank / world size");
        CK(cudaSetDevice(P_.device));
        shardCommDestroy();
        ncclUniqueId id;
        static_assert(sizeof(id) == 128, "tmap_shard_unique_id hands out 128 bytes");
        std::memcpy(&id, id128, sizeof(id));
        ncclComm_t c = nullptr;
        ncclCheck(N.CommInitRank(&c, world, id, rank), "ncclCommInitRank");
        comm_ = c;
        rank_ = rank;
        world_ = world;
        for (auto &e : shardEv_) CK(cudaEventCreate(&e));
    }
    void shardCommDestroy()
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        if (!comm_) return;
        cudaStreamSynchronize(stream_);
        for (size_t k = 0; k < peerPtr_.size(); ++k)
            if ((int)k != rank_ && peerPtr_[k]) cudaIpcCloseMemHandle(peerPtr_[k]);
        peerPtr_.clear();
        peerGen_.clear();
        for (auto &gr : recvGrave_) cudaFree(gr.second);
        recvGrave_.clear();
        NcclApi::get().CommDestroy(static_cast<ncclComm_t>(comm_));
        comm_ = nullptr;
        for (auto &e : shardEv_) { if (e) cudaEventDestroy(e); e = nullptr; }
    }
    // Slab planner (pure host function, also behind tmap_shard_plan for the CPU tests): cut n_rows bin rows into
    // `world` contiguous slabs of ~equal record count, interior boundaries on even GLOBAL bin rows (32-cell tile
    // rows), every slab >= 2 rows; then this rank's send / receive counts per peer and the receive offsets.
    static void shardPlanRows(const uint32_t *rows_all /*[world][n_rows]*/, int world, int n_rows, int first_global_row, int rank,
                              int32_t *slabs /*[2*world]*/, uint64_t *send, uint64_t *recv, uint64_t *src_base)
    {
        if (n_rows < 2 * world + 2)
            throw std::invalid_argument("map too small to shard: " + std::to_string(n_rows) + " bin rows over " + std::to_string(world) +
                                        " ranks");
        std::vector<double> cum((size_t)n_rows + 1, 0.0);
        for (int y = 0; y < n_rows; ++y)
        {
            double t = 0;
            for (int s = 0; s < world; ++s) t += (double)rows_all[(size_t)s * n_rows + y];
            cum[y + 1] = cum[y] + t;
        }
        const double total = cum[n_rows];
        std::vector<int> cuts{0};
        for (int k = 1; k < world; ++k)
        {
            const double target = total * k / world;
            int y = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
            if ((first_global_row + y) % 2) ++y;
            y = std::max(y, cuts.back() + 2);
            y = std::min(y, n_rows - 2 * (world - k));
            if ((first_global_row + y) % 2) --y;
            cuts.push_back(y);
        }
        cuts.push_back(n_rows);
        for (int k = 0; k < world; ++k) { slabs[2 * k] = cuts[k]; slabs[2 * k + 1] = cuts[k + 1]; }
        for (int k = 0; k < world; ++k)
        {
            uint64_t t = 0;
            for (int y = cuts[k]; y < cuts[k + 1]; ++y) t += rows_all[(size_t)rank * n_rows + y];
            send[k] = t;
        }
        uint64_t run = 0;
        for (int s = 0; s < world; ++s)
        {
            uint64_t t = 0;
            for (int y = cuts[rank]; y < cuts[rank + 1]; ++y) t += rows_all[(size_t)s * n_rows + y];
            recv[s] = t;
            src_base[s] = run;
            run += t;
        }
    }
    // update == false: rebuild from blank (LocalMap::clearEntireMap + re-integration, LocalMap.cpp:375-397), slabs re-planned.
    // update == true : the listed keyframes move (LocalMap::recomputeKeyFrame for each of them, LocalMap.cpp:245-315, as ONE
    //                  batch): subtract at the pose each is integrated at, add at the new one; the slabs of the last rebuild
    //                  stay (cells keep their owners), nothing is blanked, only dirty cells are re-classified.
    void shardRebuild(const uint64_t *ids, size_t n, const float *poses, ShardInfo *info, bool update = false)
    {
        std::lock_guard<std::recursive_mutex> g(gridMutex_);
        if (!comm_) throw std::invalid_argument("tmap_shard_comm_init has not been called on this map");
        NcclApi &N = NcclApi::get();
        ncclComm_t comm = static_cast<ncclComm_t>(comm_);
        const int W = world_, rank = rank_;
        if (W > 32) throw Unsupported("sharded rebuild: at most 32 ranks");
        CK(cudaSetDevice(P_.device));
        const auto t_begin = std::chrono::steady_clock::now();
        static const bool dbg = std::getenv("TMB_SHARD_DEBUG") != nullptr;
        // TMB_SHARD_ROUTE=0: bucket locally and route with grouped ncclSend / ncclRecv instead of the fused scatter (A/B runs)
        static const bool fused = !(std::getenv("TMB_SHARD_ROUTE") && std::atoi(std::getenv("TMB_SHARD_ROUTE")) == 0);
        auto trace = [&](const char *what) { if (dbg) { std::fprintf(stderr, "[shard %d/%d] %s\n", rank, W, what); std::fflush(stderr); } };
        CK(cudaEventRecord(shardEv_[0], stream_));
        // 1. job-wide op numbering and bbox: every rank's keyframe count and window in ONE all-gather of 8 int64
        // A rank that finds its own arguments invalid must not leave the others waiting in a collective: the verdict
        // travels with the first all-gather and every rank throws together.
        std::string local_error;
        if (update && shardCuts_.size() != (size_t)W + 1)
            local_error = "tmap_shard_update needs a map built by tmap_shard_rebuild";
        std::vector<Op> ops;
        std::vector<std::shared_ptr<KeyFrame>> moved;
        {
            std::lock_guard<std::mutex> l(kfMutex_);
            for (size_t i = 0; i < n && local_error.empty(); ++i)
            {
                auto it = kfs_.find(ids[i]);
                if (it == kfs_.end()) { local_error = "unknown keyframe id in shard op list"; break; }
                if (i && ids[i] <= ids[i - 1]) { local_error = "keyframe ids must be ascending"; break; }
                for (int q = 0; q < 12; ++q)
                    if (!std::isfinite(poses[12 * i + q])) local_error = "non-finite pose";
                if (!local_error.empty()) break;
                if (update && it->second->integrated)
                {
                    Op sub;
                    sub.kf = it->second;
                    std::memcpy(sub.pose, it->second->intPose, sizeof(sub.pose));
                    sub.sign = -1.f;
                    ops.push_back(sub);
                }
                Op add;
                add.kf = it->second;
                std::memcpy(add.pose, poses + 12 * i, sizeof(add.pose));
                add.sign = +1.f;
                ops.push_back(add);
                moved.push_back(it->second);
            }
        }
        // the limits planBatch enforces, checked here so that a rank-local refusal is agreed on before the data collectives
        int32_t win[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN};
        {
            uint64_t pts = 0;
            for (auto &op : ops)
            {
                int x0, y0w, x1, y1w;
                opWindow(op, x0, y0w, x1, y1w);
                if ((uint64_t)(x1 - x0 + 1) * (uint64_t)(y1w - y0w + 1) > MAX_WINDOW_BINS)
                    local_error = "a keyframe spans more than " + std::to_string(MAX_WINDOW_BINS) + " bins of 16x16 cells";
                win[0] = std::min(win[0], x0); win[1] = std::min(win[1], y0w); win[2] = std::max(win[2], x1); win[3] = std::max(win[3], y1w);
                pts += op.kf->n;
            }
            if (pts >= (1ull << 32)) local_error = "batch exceeds 2^32 points";
            if (ops.size() > 65535) local_error = "batch exceeds 65535 ops";
        }
        if (!local_error.empty()) { ops.clear(); moved.clear(); }
        const size_t n_ops_local = ops.size();
        shardH_.ensure(8192 + (size_t)W * 256);
        shardD_.ensure(8192 + (size_t)W * 256);
        long long *mh = shardH_.as<long long>();
        mh[0] = (long long)n_ops_local; mh[1] = win[0]; mh[2] = win[1]; mh[3] = win[2]; mh[4] = win[3];
        mh[5] = local_error.empty() ? 0 : 1; mh[6] = mh[7] = 0;
        long long *md = shardD_.as<long long>();
        CK(cudaMemcpyAsync(md, mh, 64, cudaMemcpyHostToDevice, stream_));
        ncclCheck(N.AllGather(md, md + 8, 8, ncclInt64, comm, stream_), "ncclAllGather(meta)");
        CK(cudaMemcpyAsync(mh + 8, md + 8, (size_t)W * 64, cudaMemcpyDeviceToHost, stream_));
        g_h2d += 64; g_d2h += (uint64_t)W * 64;
        CK(cudaStreamSynchronize(stream_));
        uint64_t op_base = 0;
        int32_t u[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN};
        for (int s = 0; s < W; ++s)
            if (mh[8 + 8 * s + 5])
                throw std::invalid_argument(local_error.empty() ? "sharded pass refused: rank " + std::to_string(s) + " reported invalid arguments"
                                                                : local_error);
        for (int s = 0; s < W; ++s)
        {
            const long long *m = mh + 8 + 8 * s;
            if (s < rank) op_base += (uint64_t)m[0];
            if (m[0])
            {
                u[0] = std::min<long long>(u[0], m[1]); u[1] = std::min<long long>(u[1], m[2]);
                u[2] = std::max<long long>(u[2], m[3]); u[3] = std::max<long long>(u[3], m[4]);
            }
        }
        if (u[0] == INT_MAX) throw std::invalid_argument("sharded rebuild without keyframes");
        {
            uint64_t all_ops = 0;
            for (int s = 0; s < W; ++s) all_ops += (uint64_t)mh[8 + 8 * s];
            if (all_ops >= (1u << 23)) throw Unsupported("sharded pass exceeds 2^23 ops over the job");
        }
        if (!update)  // a rebuild drops every contribution: keyframes not listed are no longer integrated
        {
            std::lock_guard<std::mutex> l(kfMutex_);
            for (auto &kv : kfs_) kv.second->integrated = false;
        }
        const int margin = std::max(1, (r_ + BIN - 1) / BIN);
        int32_t bbox[4] = {u[0] - margin, u[1] - margin, u[2] + margin, u[3] + margin};
        if (update)
        {
            // keep every slab boundary of the last rebuild inside the bbox so that each rank has rows to own
            bbox[1] = std::min(bbox[1], shardCuts_[0]);
            bbox[3] = std::max(bbox[3], shardCuts_[W] - 1);
        }
        // 2. key pass + scans over the shared bbox for this rank's keyframes: per-bin counts and every record's slot in
        //    this rank's (virtual) bucket array.  The grid is not touched yet and no record has moved.
        trace("meta exchanged");
        shardKeyPass(ops, (uint32_t)op_base, bbox);
        BatchView &b = shardPlan_.b;
        const size_t nb = (size_t)b.btw * b.bth;
        const int bth = b.bth, btw = b.btw;
        // 3. per-row totals + this rank's cell bbox + its error flags of every rank -> host (slab plan); per-bin counts and
        //    starts of every rank stay on the device (the fold's segment tables)
        const size_t rw = (size_t)bth + 8;  // u32 per rank
        rowsD_.ensure((rw + (size_t)W * rw) * 4 + 64);
        rowsH_.ensure((size_t)W * rw * 4 + 64);
        uint32_t *rowsMine = rowsD_.as<uint32_t>(), *rowsAll = rowsMine + rw;
        CK(launch_row_totals(stream_, b, rowsMine));
        CK(launch_ops_bbox(stream_, b, reinterpret_cast<int32_t *>(rowsMine + bth)));
        g_launches += 2;
        ncclCheck(N.AllGather(rowsMine, rowsAll, rw, ncclUint32, comm, stream_), "ncclAllGather(rows)");
        CK(cudaMemcpyAsync(rowsH_.p, rowsAll, (size_t)W * rw * 4, cudaMemcpyDeviceToHost, stream_));
        g_d2h += (uint64_t)W * rw * 4;
        countsAllD_.ensure((size_t)W * nb * 4 + 64);
        startsAllD_.ensure((size_t)W * nb * 4 + 64);
        ncclCheck(N.GroupStart(), "ncclGroupStart");
        ncclCheck(N.AllGather(b.tile_count, countsAllD_.p, nb, ncclUint32, comm, stream_), "ncclAllGather(counts)");
        ncclCheck(N.AllGather(b.tile_start, startsAllD_.p, nb, ncclUint32, comm, stream_), "ncclAllGather(starts)");
        ncclCheck(N.GroupEnd(), "ncclGroupEnd");
        CK(cudaEventRecord(shardEv_[1], stream_));
        CK(cudaStreamSynchronize(stream_));
        std::vector<uint32_t> rows((size_t)W * bth);
        int32_t cb[4] = {INT_MAX, INT_MIN, INT_MAX, INT_MIN};
        uint32_t err_flags = 0;
        for (int s = 0; s < W; ++s)
        {
            const uint32_t *src = rowsH_.as<uint32_t>() + (size_t)s * rw;
            std::memcpy(rows.data() + (size_t)s * bth, src, (size_t)bth * 4);
            int32_t c5[5];
            std::memcpy(c5, src + bth, 20);
            if (c5[0] != INT_MAX)
            {
                cb[0] = std::min(cb[0], c5[0]); cb[1] = std::max(cb[1], c5[1]);
                cb[2] = std::min(cb[2], c5[2]); cb[3] = std::max(cb[3], c5[3]);
            }
            err_flags |= (uint32_t)c5[4];
        }
        trace("rows gathered");
        if (err_flags)
            throw CudaError("binning consistency check failed on some rank (flags " + std::to_string(err_flags) + ")");
        std::vector<int32_t> slabs(2 * (size_t)W);
        std::vector<uint64_t> send(W), recv(W), src_base(W);
        if (!update)
        {
            shardPlanRows(rows.data(), W, bth, b.bty0, rank, slabs.data(), send.data(), recv.data(), src_base.data());
            shardCuts_.assign((size_t)W + 1, 0);  // slab boundaries as GLOBAL bin rows: they survive a change of bbox
            for (int kq = 0; kq < W; ++kq) shardCuts_[kq] = b.bty0 + slabs[2 * kq];
            shardCuts_[W] = b.bty0 + bth;
        }
        else
        {
            // the owners stay: the cuts of the last rebuild, clamped into this pass's bbox rows (the first / last rank own
            // whatever lies beyond the old extent)
            for (int kq = 0; kq < W; ++kq)
            {
                const int lo = kq == 0 ? 0 : std::min(std::max(shardCuts_[kq] - b.bty0, 0), bth);
                const int hi = kq == W - 1 ? bth : std::min(std::max(shardCuts_[kq + 1] - b.bty0, 0), bth);
                slabs[2 * kq] = lo;
                slabs[2 * kq + 1] = std::max(hi, lo);
            }
            uint64_t run = 0;
            for (int kq = 0; kq < W; ++kq)
            {
                uint64_t t = 0;
                for (int y = slabs[2 * kq]; y < slabs[2 * kq + 1]; ++y) t += rows[(size_t)rank * bth + y];
                send[kq] = t;
            }
            for (int sq = 0; sq < W; ++sq)
            {
                uint64_t t = 0;
                for (int y = slabs[2 * rank]; y < slabs[2 * rank + 1]; ++y) t += rows[(size_t)sq * bth + y];
                recv[sq] = t;
                src_base[sq] = run;
                run += t;
            }
        }
        const int y0 = slabs[2 * rank], y1 = slabs[2 * rank + 1];
        const int cj0 = (b.bty0 + y0) * BIN, cj1 = (b.bty0 + y1) * BIN;
        uint64_t n_recv = 0, n_sent = 0, n_local = 0;
        for (int s = 0; s < W; ++s) { n_recv += recv[s]; n_local += send[s]; if (s != rank) n_sent += send[s]; }
        // the window of THIS rank: its slab rows + halo, the bbox columns (the other slabs never exist here)
        {
            const int mt = (std::max(r_, R_) + r_ + TILE - 1) / TILE;  // footprint (+ inflation) halo, in 32-cell tiles
            const int ty0 = ((b.bty0 + y0) >> 1) - mt, ty1 = ((b.bty0 + y1 + 1) >> 1) + mt;
            grid_.ensure(stream_, b.btx0 >> 1, ty0, ((b.btx0 + btw - 1) >> 1) + 1, ty1);
        }
        if (R_ > 0)
        {
            // the inflation work list and its dedup stamps are sized by the window, which has only just been fixed
            const size_t n_t32 = (size_t)grid_.v.tw * grid_.v.th;
            const size_t need = n_t32 * 4 + 64;
            if (need > inflFlagsD_.cap)
            {
                inflFlagsD_.ensure(need);
                CK(cudaMemsetAsync(inflFlagsD_.p, 0, inflFlagsD_.cap, stream_));
                epoch_ = 0;
            }
            if (++epoch_ == 0)
            {
                CK(cudaMemsetAsync(inflFlagsD_.p, 0, inflFlagsD_.cap, stream_));
                epoch_ = 1;
            }
            shardInflD_.ensure(need);
            b.infl_flags = inflFlagsD_.as<uint32_t>();
            b.epoch = epoch_;
            b.infl_tiles = shardInflD_.as<uint32_t>();
        }
        // 4. receive buffers: one per rank, mapped into every peer (CUDA IPC); handles travel only when a buffer was
        //    reallocated.  rt_off[k] turns a slot of my bucket array into a slot of rank k's buffer: my records of k's slab
        //    land behind those of the lower ranks, in my bucket order.
        for (auto it = recvGrave_.begin(); it != recvGrave_.end();)  // buffers replaced two or more rebuilds ago: unmapped everywhere
        {
            if (it->first + 2 <= shardStep_) { cudaFree(it->second); it = recvGrave_.erase(it); }
            else ++it;
        }
        ++shardStep_;
        if (recvD_.cap < (size_t)n_recv * 16 + 16)
        {
            if (recvD_.p) recvGrave_.emplace_back(shardStep_, recvD_.p);  // the peers still map it until they see the new handle
            recvD_.p = nullptr;
            recvD_.cap = 0;
            recvD_.ensure((size_t)n_recv * 16 + (size_t)n_recv * 4 + 4096);  // 25 % spare: slab loads move a little from rebuild to rebuild
            ++recvGen_;
        }
        if (fused)
        {
            struct HandleMsg { unsigned long long gen; cudaIpcMemHandle_t h; unsigned char pad[128 - 8 - sizeof(cudaIpcMemHandle_t)]; };
            static_assert(sizeof(HandleMsg) == 128, "handle message");
            HandleMsg *hm = reinterpret_cast<HandleMsg *>(shardH_.as<unsigned char>() + 4096);
            HandleMsg *hd = reinterpret_cast<HandleMsg *>(shardD_.as<unsigned char>() + 4096);
            std::memset(hm, 0, sizeof(HandleMsg));
            hm[0].gen = recvGen_;
            if (W > 1) CK(cudaIpcGetMemHandle(&hm[0].h, recvD_.p));
            CK(cudaMemcpyAsync(hd, hm, sizeof(HandleMsg), cudaMemcpyHostToDevice, stream_));
            ncclCheck(N.AllGather(hd, hd + 1, sizeof(HandleMsg), ncclUint8, comm, stream_), "ncclAllGather(handles)");
            CK(cudaMemcpyAsync(hm + 1, hd + 1, (size_t)W * sizeof(HandleMsg), cudaMemcpyDeviceToHost, stream_));
            CK(cudaStreamSynchronize(stream_));
            trace("handles gathered");
            peerPtr_.resize(W, nullptr);
            peerGen_.resize(W, 0);
            for (int k = 0; k < W; ++k)
            {
                if (k == rank) { peerPtr_[k] = recvD_.p; continue; }
                const HandleMsg &m = hm[1 + k];
                if (peerGen_[k] != m.gen || !peerPtr_[k])
                {
                    if (peerPtr_[k]) CK(cudaIpcCloseMemHandle(peerPtr_[k]));
                    void *p = nullptr;
                    CK(cudaIpcOpenMemHandle(&p, m.h, cudaIpcMemLazyEnablePeerAccess));
                    peerPtr_[k] = p;
                    peerGen_[k] = m.gen;
                }
            }
            trace("peer buffers mapped");
            // routing tables: [ptr u64 x W][off u32 x W][y1 i32 x W] in one small upload
            unsigned char *th = shardH_.as<unsigned char>() + 2048;
            unsigned long long *tp = reinterpret_cast<unsigned long long *>(th);
            uint32_t *to = reinterpret_cast<uint32_t *>(th + 8 * W);
            int32_t *ty = reinterpret_cast<int32_t *>(th + 12 * W);
            // my records sent to k start, in MY bucket array, at the first slot of k's slab = sum of my rows below it
            uint64_t my_first = 0;
            for (int k = 0; k < W; ++k)
            {
                uint64_t base_at_k = 0;  // records rank k receives from the ranks below me
                for (int s = 0; s < rank; ++s)
                    for (int y = slabs[2 * k]; y < slabs[2 * k + 1]; ++y) base_at_k += rows[(size_t)s * bth + y];
                tp[k] = reinterpret_cast<unsigned long long>(peerPtr_[k]);
                to[k] = (uint32_t)(base_at_k - my_first);  // modulo 2^32: slots stay below 2^32 on both sides
                ty[k] = slabs[2 * k + 1];
                my_first += send[k];
            }
            unsigned char *td = shardD_.as<unsigned char>() + 2048;
            CK(cudaMemcpyAsync(td, th, (size_t)16 * W, cudaMemcpyHostToDevice, stream_));
            b.rt_ptr = reinterpret_cast<const unsigned long long *>(td);
            b.rt_off = reinterpret_cast<const uint32_t *>(td + 8 * W);
            b.rt_y1 = reinterpret_cast<const int *>(td + 12 * W);
            b.rt_world = W;
        }
        CK(cudaEventRecord(shardEv_[2], stream_));
        if (fused)
        {
            // 5. the scatter IS the all-to-all: every record is written once, straight into its owner's receive buffer
            //    (NVLink stores through the peer mapping; the own slab's records into the own buffer).  A tiny all-gather
            //    afterwards is the "every peer has finished writing into my buffer" barrier.
            CK(launch_bin(stream_, b, grid_.v, lat_, shardPlan_.ppt, std::max(shardPlan_.maxT, 1u), shardPlan_.dc, shardPlan_.di, nSM_,
                          nullptr, 8));
            g_launches += 1;
            b.rt_world = 0;
            uint32_t *fl = reinterpret_cast<uint32_t *>(shardD_.as<unsigned char>() + 3072);
            CK(cudaMemcpyAsync(fl, b.counters + 3, 4, cudaMemcpyDeviceToDevice, stream_));
            ncclCheck(N.AllGather(fl, fl + 1, 1, ncclUint32, comm, stream_), "ncclAllGather(barrier)");
        }
        else
        {
            // 5'. bucket locally, then grouped ncclSend / ncclRecv: a slab's records are one contiguous range of the sender's
            //     bucket array
            CK(launch_bin(stream_, b, grid_.v, lat_, shardPlan_.ppt, std::max(shardPlan_.maxT, 1u), shardPlan_.dc, shardPlan_.di, nSM_,
                          nullptr, 8));
            g_launches += 1;
            const float4 *sorted = b.sorted;
            float4 *rb = recvD_.as<float4>();
            uint64_t soff = 0;
            ncclCheck(N.GroupStart(), "ncclGroupStart");
            for (int p = 0; p < W; ++p)
            {
                if (p != rank)
                {
                    if (send[p]) ncclCheck(N.Send(sorted + soff, (size_t)send[p] * 4, ncclFloat, p, comm, stream_), "ncclSend(records)");
                    if (recv[p]) ncclCheck(N.Recv(rb + src_base[p], (size_t)recv[p] * 4, ncclFloat, p, comm, stream_), "ncclRecv(records)");
                }
                else if (send[p])
                    CK(cudaMemcpyAsync(rb + src_base[p], sorted + soff, (size_t)send[p] * 16, cudaMemcpyDeviceToDevice, stream_));
                soff += send[p];
            }
            ncclCheck(N.GroupEnd(), "ncclGroupEnd");
        }
        CK(cudaEventRecord(shardEv_[3], stream_));
        trace("records routed (enqueued)");
        // 6. blank the slab + footprint halo; fold: a bin's bucket = the sources' segments in rank order, read in place
        if (!update)
            shardBlankRows(cj0 - 2 * BIN, (y1 - y0) * BIN + 4 * BIN);
        shardFold(recvD_.p, n_recv, countsAllD_.p, startsAllD_.p, src_base.data(), W, y0, y1, false);
        CK(cudaEventRecord(shardEv_[4], stream_));
        // 7. halo exchange, in place: r cell rows of moment records and flag bytes either side of each slab boundary, over
        //    the bbox columns (row segments are contiguous in the row-major window)
        uint64_t halo_bytes = 0;
        if (W > 1)
        {
            const GridView &v = grid_.v;
            const int col0 = b.btx0 * BIN - v.ci0, wcols = btw * BIN;
            if (col0 < 0 || col0 + wcols > v.W) throw CudaError("sharded window does not cover the bbox columns");
            auto mom_row = [&](int cj) { return v.mom + ((size_t)(cj - v.cj0) * v.W + col0) * REC; };
            auto flag_row = [&](int cj) { return v.flags + (size_t)(cj - v.cj0) * v.W + col0; };
            auto in_grid = [&](int cj) { return cj >= v.cj0 && cj < v.cj0 + v.H; };
            ncclCheck(N.GroupStart(), "ncclGroupStart");
            for (int k = 0; k < r_; ++k)
            {
                if (rank > 0)
                {
                    if (!in_grid(cj0 + k) || !in_grid(cj0 - r_ + k)) throw CudaError("halo rows outside the allocated window");
                    ncclCheck(N.Send(mom_row(cj0 + k), (size_t)wcols * REC, ncclFloat, rank - 1, comm, stream_), "ncclSend(halo)");
                    ncclCheck(N.Send(flag_row(cj0 + k), (size_t)wcols, ncclUint8, rank - 1, comm, stream_), "ncclSend(halo)");
                    ncclCheck(N.Recv(mom_row(cj0 - r_ + k), (size_t)wcols * REC, ncclFloat, rank - 1, comm, stream_), "ncclRecv(halo)");
                    ncclCheck(N.Recv(flag_row(cj0 - r_ + k), (size_t)wcols, ncclUint8, rank - 1, comm, stream_), "ncclRecv(halo)");
                    halo_bytes += (uint64_t)wcols * 49;
                }
                if (rank < W - 1)
                {
                    if (!in_grid(cj1 - r_ + k) || !in_grid(cj1 + k)) throw CudaError("halo rows outside the allocated window");
                    ncclCheck(N.Send(mom_row(cj1 - r_ + k), (size_t)wcols * REC, ncclFloat, rank + 1, comm, stream_), "ncclSend(halo)");
                    ncclCheck(N.Send(flag_row(cj1 - r_ + k), (size_t)wcols, ncclUint8, rank + 1, comm, stream_), "ncclSend(halo)");
                    ncclCheck(N.Recv(mom_row(cj1 + k), (size_t)wcols * REC, ncclFloat, rank + 1, comm, stream_), "ncclRecv(halo)");
                    ncclCheck(N.Recv(flag_row(cj1 + k), (size_t)wcols, ncclUint8, rank + 1, comm, stream_), "ncclRecv(halo)");
                    halo_bytes += (uint64_t)wcols * 49;
                }
            }
            ncclCheck(N.GroupEnd(), "ncclGroupEnd");
        }
        CK(cudaEventRecord(shardEv_[5], stream_));
        // 8. classify the slab; the received halo rows are not ours: drop their touched / changed bits afterwards
        if (cb[0] != INT_MAX)
            growLogical(cb[0], cb[1], cb[2], cb[3]);
        {
            std::vector<Op> none;
            Plan pl = shardPlan_;
            pl.n_ops = 0;
            // step inflation (LocalMap::inflateDirty, LocalMap.cpp:201-241) reads the step layer R cells around a cell: the R
            // rows of it either side of each slab boundary are exchanged after the classification, before the inflation
            auto step_halo = [&]()
            {
                if (W == 1) return;
                const GridView &v = grid_.v;
                const int col0 = b.btx0 * BIN - v.ci0, wcols = btw * BIN;
                float *step = v.der + (size_t)D_STEP * v.plane();
                auto row = [&](int cj) { return step + (size_t)(cj - v.cj0) * v.W + col0; };
                auto in_grid = [&](int cj) { return cj >= v.cj0 && cj < v.cj0 + v.H; };
                ncclCheck(N.GroupStart(), "ncclGroupStart");
                for (int q = 0; q < R_; ++q)
                {
                    if (rank > 0)
                    {
                        if (!in_grid(cj0 + q) || !in_grid(cj0 - R_ + q)) throw CudaError("inflation halo rows outside the allocated window");
                        ncclCheck(N.Send(row(cj0 + q), (size_t)wcols, ncclFloat, rank - 1, comm, stream_), "ncclSend(step halo)");
                        ncclCheck(N.Recv(row(cj0 - R_ + q), (size_t)wcols, ncclFloat, rank - 1, comm, stream_), "ncclRecv(step halo)");
                        halo_bytes += (uint64_t)wcols * 4;
                    }
                    if (rank < W - 1)
                    {
                        if (!in_grid(cj1 - R_ + q) || !in_grid(cj1 + q)) throw CudaError("inflation halo rows outside the allocated window");
                        ncclCheck(N.Send(row(cj1 - R_ + q), (size_t)wcols, ncclFloat, rank + 1, comm, stream_), "ncclSend(step halo)");
                        ncclCheck(N.Recv(row(cj1 + q), (size_t)wcols, ncclFloat, rank + 1, comm, stream_), "ncclRecv(step halo)");
                        halo_bytes += (uint64_t)wcols * 4;
                    }
                }
                ncclCheck(N.GroupEnd(), "ncclGroupEnd");
            };
            finishBatch(pl, none, t_begin, step_halo);
        }
        if (W > 1)
        {
            const GridView &v = grid_.v;
            const int col0 = b.btx0 * BIN - v.ci0, wcols = btw * BIN;
            if (rank > 0)
                CK(cudaMemset2DAsync(v.flags + (size_t)(cj0 - r_ - v.cj0) * v.W + col0, v.W, 0, wcols, r_, stream_));
            if (rank < W - 1)
                CK(cudaMemset2DAsync(v.flags + (size_t)(cj1 - v.cj0) * v.W + col0, v.W, 0, wcols, r_, stream_));
        }
        CK(cudaEventRecord(shardEv_[6], stream_));
        CK(cudaStreamSynchronize(stream_));
        for (size_t i = 0; i < moved.size(); ++i)
        {
            moved[i]->integrated = true;
            std::memcpy(moved[i]->intPose, poses + 12 * i, sizeof(moved[i]->intPose));
        }
        trace("rebuild done");
        if (info)
        {
            std::memset(info, 0, sizeof(*info));
            info->bbox[0] = b.btx0; info->bbox[1] = b.bty0; info->bbox[2] = btw; info->bbox[3] = bth;
            info->slab_y0 = y0; info->slab_y1 = y1; info->cj0 = cj0; info->cj1 = cj1;
            info->local_records = n_local; info->sent_records = n_sent; info->recv_records = n_recv;
            info->halo_bytes = halo_bytes;
            info->grid_bytes = (uint64_t)grid_.v.plane() * (REC * 4 + NUM_DER * 4 + 1);
            for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&info->ms[k], shardEv_[k], shardEv_[k + 1]);
            cudaEventElapsedTime(&info->ms[6], shardEv_[0], shardEv_[6]);
            info->n_slabs = W;
            for (int k = 0; k < 2 * W; ++k) info->slabs[k] = slabs[k];
        }
    }
    int rOf() const { return r_; }

  private:
    void growLogical(int cimin, int cimax, int cjmin, int cjmax)
    {
        const double res = P_.resolution;
        const double minx = lat_.x0 + cimin * res, maxx = lat_.x0 + cimax * res;
        const double miny = lat_.y0 + cjmin * res, maxy = lat_.y0 + cjmax * res;
        const double margin = res;
        const double curHalfX = ((2 * kx_ + 1) * res) / 2.0, curHalfY = ((2 * ky_ + 1) * res) / 2.0;
        const double needX = std::max(std::abs(maxx - lat_.x0), std::abs(minx - lat_.x0)) + margin;
        const double needY = std::max(std::abs(maxy - lat_.y0), std::abs(miny - lat_.y0)) + margin;
        if (needX <= curHalfX && needY <= curHalfY)
            return;
        double newHalfX = curHalfX, newHalfY = curHalfY;
        while (newHalfX < needX) newHalfX += P_.extend_length;
        while (newHalfY < needY) newHalfY += P_.extend_length;
        kx_ = static_cast<int>(std::ceil(newHalfX / res));
        ky_ = static_cast<int>(std::ceil(newHalfY / res));
    }

    static std::mutex &constMutex()
    {
        static std::mutex m;
        return m;
    }
    void uploadConstantsLocked()  // caller holds constMutex()
    {
        // re-upload whenever a different map instance (identified by a never-reused serial) or another device used the
        // tables last
        static uint64_t owner = 0;
        static int owner_dev = -1;
        if (owner == serial_ && owner_dev == P_.device) return;
        const int TW = BIN + 2 * r_;
        std::vector<int> off(disc_.size());
        std::vector<double> dd(disc_.size() * 2);
        for (size_t i = 0; i < disc_.size(); ++i)
        {
            off[i] = disc_[i].second * TW + disc_[i].first;
            dd[2 * i] = disc_[i].first;
            dd[2 * i + 1] = disc_[i].second;
        }
        CK(cudaDeviceSynchronize());
        CK(upload_classify_constants(off.data(), dd.data(), (int)disc_.size()));
        owner = serial_;
        owner_dev = P_.device;
    }

    uint64_t mapID_;
    uint64_t serial_ = 0;
    ClassifyParams cp_{};
    tmap_params P_;
    tmap_update_cb cb_;
    void *user_;
    int nSM_ = 148;
    LatticeD lat_;
    std::vector<std::pair<int, int>> disc_;
    int r_ = 0, R_ = 0, nInfl_ = 0;
    int kx_ = 0, ky_ = 0;  // logical (reference) grid half extents in cells

    std::recursive_mutex gridMutex_;  // masterGridMapMutex_
    DeviceGrid grid_;
    cudaStream_t stream_ = nullptr, ownStream_ = nullptr;
    cudaEvent_t ev_[9];  // [8]: between k_gate and k_classify
    tmap_stats stats_{};

    std::mutex kfMutex_;  // keyFramesMapMutex_
    std::map<uint64_t, std::shared_ptr<KeyFrame>> kfs_;

    std::mutex qMutex_;  // poseQueueMutex_
    std::condition_variable qCv_, idleCv_;
    std::deque<std::shared_ptr<KeyFrame>> queue_;
    std::unordered_set<uint64_t> queued_;
    bool running_ = false, busy_ = false;
    std::string workerError_;
    std::thread worker_;

    PinBuf metaH_, resH_;
    DevBuf metaD_, histD_, totD_, tileD_, sortedD_, inflFlagsD_, decay_;
    DevBuf voxXyz_, voxIdx_, voxTemp_, voxOut_;
    uint32_t epoch_ = 0;
    DevBuf scratchKeys_, scratchLayers_, scratchOut_, scratchCount_, mergedD_, srcBaseD_, expScratch_, expCount_;
    uint32_t *globD_ = nullptr;
    Plan shardPlan_;
    std::vector<Op> shardOps_;
    void *comm_ = nullptr;  // ncclComm_t of the sharded job (tmap_shard_comm_init)
    int rank_ = 0, world_ = 1;
    cudaEvent_t shardEv_[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    PinBuf shardH_, rowsH_;
    DevBuf shardD_, rowsD_, countsAllD_, startsAllD_, recvD_, shardInflD_;
    std::vector<std::pair<uint64_t, void *>> recvGrave_;  // replaced receive buffers, freed once no peer can still map them
    uint64_t shardStep_ = 0;
    std::vector<int> shardCuts_;  // slab boundaries of the last sharded rebuild, global bin rows, world + 1 entries
    unsigned long long recvGen_ = 1;       // bumped whenever recvD_ is reallocated: peers re-open its IPC handle
    std::vector<void *> peerPtr_;          // every rank's receive buffer as mapped into this process
    std::vector<unsigned long long> peerGen_;

  public:
    std::string takeWorkerError()
    {
        std::lock_guard<std::mutex> l(qMutex_);
        std::string e;
        e.swap(workerError_);
        return e;
    }
};

// -------------------------------------------------------------------------------------------
// System
// -------------------------------------------------------------------------------------------
static void compose(const float A[12], const float B[12], float out[12])
{
    // Eigen Affine3f * Affine3f: linear = A.linear*B.linear, translation = A.linear*B.translation + A.translation,
    // coefficient-wise ((a0 b0 + a1 b1) + a2 b2), float
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j)
            out[4 * i + j] = (A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j];
        out[4 * i + 3] = ((A[4 * i] * B[3] + A[4 * i + 1] * B[7]) + A[4 * i + 2] * B[11]) + A[4 * i + 3];
    }
}

class System
{
  public:
    explicit System(const tmap_params &P) : P_(P)
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev <= 0 || P.device >= ndev || P.device < 0)
            throw std::runtime_error("NO_DEVICE");
        CK(cudaSetDevice(P.device));
        if (!(P.resolution > 0)) throw std::invalid_argument("resolution must be positive");
        CK(cudaStreamCreateWithFlags(&ingest_, cudaStreamNonBlocking));
        // keep freed keyframe clouds cached in the stream-ordered pool (rolling windows recycle them)
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, P.device));
        uint64_t thr = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    ~System()
    {
        current_.reset();  // every owner of a KeyFrame must go before the allocation stream does
        maps_.clear();
        kfs_.clear();
        for (auto &b : cloudBuf_) cudaFreeAsync(b.d, ingest_);
        cloudBuf_.clear();
        cudaStreamSynchronize(ingest_);
        cudaStreamDestroy(ingest_);
    }

    int setExtrinsics(const float Tsv[12], const float Tbs[12])  // System.cpp:45-51
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        compose(Tbs, Tsv, Tbv_);
        return TMAP_OK;
    }
    int setObstacleParams(const float T[12], double mx, double mn)  // System.cpp:53-60
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        std::memcpy(Tb_obs_, T, sizeof(Tb_obs_));
        obsMax_ = mx; obsMin_ = mn; obsSet_ = true;
        return TMAP_OK;
    }
    int addMap(uint64_t id, tmap_update_cb cb, void *user)  // System.cpp:62-72
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        if (maps_.count(id)) return TMAP_IGNORED_DUPLICATE;
        auto mp = std::make_shared<LocalMap>(id, P_, cb, user);
        {
            std::lock_guard<std::mutex> r(mapsM_);  // writers hold m_ AND mapsM_; map() reads under mapsM_ alone
            maps_[id] = mp;
        }
        current_ = mp;
        currentId_.store(id);
        return TMAP_OK;
    }
    int setCurrentMap(uint64_t id)  // System.cpp:74-82
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        auto it = maps_.find(id);
        if (it == maps_.end()) return TMAP_IGNORED_UNKNOWN_MAP;
        current_ = it->second;
        currentId_.store(id);
        return TMAP_OK;
    }
    uint64_t currentMapId() const { return currentId_.load(); }
    // Registry lookup for the reader / query entry points.  It must NOT take m_ (localMapMutex_): a reader
    // holds the map's grid mutex across take + read (LocalMap.hpp:102-104) while deleteKeyFrame /
    // updateKFMap / informLoopClosure go m_ -> grid mutex; the reference's readers call LocalMap methods
    // directly and never touch localMapMutex_ either.
    std::shared_ptr<LocalMap> map(uint64_t id)
    {
        std::lock_guard<std::mutex> r(mapsM_);
        auto it = maps_.find(id);
        return it == maps_.end() ? nullptr : it->second;
    }

    // System::prune on the device; src may be host or device memory
    std::shared_ptr<KeyFrame> pruneToKeyFrame(const void *src, bool on_device, size_t n, size_t stride, const float T[12], double mx,
                                              double mn, const int *field_floats = nullptr, int big_endian = 0)
    {
        if (stride < 12 || stride % 4) throw std::invalid_argument("stride_bytes must be a multiple of 4 and >= 12");
        std::lock_guard<std::mutex> l(ingestMutex_);
        CK(cudaSetDevice(P_.device));
        auto kf = std::make_shared<KeyFrame>();
        if (n == 0) return kf;
        const void *d_src = src;
        if (!on_device)
        {
            raw_.ensure(n * stride);
            CK(cudaMemcpyAsync(raw_.p, src, n * stride, cudaMemcpyHostToDevice, ingest_));
            g_h2d += (uint64_t)(n * stride);
            d_src = raw_.p;
        }
        const size_t n_warps = (n + 255) / 256;
        warpCnt_.ensure(n_warps * 4 + 64);
        total_.ensure(16);
        totalH_.ensure(16);
        // Two phases around ONE host round trip: count (+ max norm) -> the store is allocated for exactly the points
        // that survive the prune (a LiDAR frame loses 40-50 % to misses / range / height) -> compaction.  The
        // compaction is not waited for here: the keyframe carries an event the consuming stream waits on.
        CK(launch_prune(ingest_, d_src, n, stride, T, (float)P_.robot_height, (float)mx, (float)mn, warpCnt_.as<uint32_t>(),
                        total_.as<uint32_t>(), nullptr, 0, field_floats, big_endian));
        CK(cudaMemcpyAsync(totalH_.p, total_.p, 8, cudaMemcpyDeviceToHost, ingest_));
        g_d2h += (uint64_t)(8);
        CK(cudaStreamSynchronize(ingest_));  // the caller's host buffer has been consumed by now as well
        kf->n = totalH_.as<uint32_t>()[0];
        std::memcpy(&kf->maxnorm, totalH_.as<uint32_t>() + 1, 4);
        g_launches += 2;
        if (kf->n == 0) return kf;
        CK(cudaMallocAsync(reinterpret_cast<void **>(&kf->d_cloud), (size_t)kf->n * sizeof(float4), ingest_));
        kf->allocStream = ingest_;
        CK(launch_prune(ingest_, d_src, n, stride, T, (float)P_.robot_height, (float)mx, (float)mn, warpCnt_.as<uint32_t>(),
                        total_.as<uint32_t>(), kf->d_cloud, 1, field_floats, big_endian));
        g_launches += 1;
        CK(cudaEventCreateWithFlags(&kf->ready, cudaEventDisableTiming));
        CK(cudaEventRecord(kf->ready, ingest_));
        if (on_device)
            CK(cudaStreamSynchronize(ingest_));  // a caller-owned device source may be reused as soon as we return
        return kf;
    }

    int registerKeyFrame(double ts, uint64_t kfID, uint64_t mapID, std::shared_ptr<KeyFrame> kf)  // System.cpp:121-151
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        if (kfs_.count(kfID)) return TMAP_IGNORED_DUPLICATE;
        auto mit = maps_.find(mapID);
        if (mit == maps_.end()) return TMAP_IGNORED_UNKNOWN_MAP;
        if (kf->n == 0) return TMAP_IGNORED_EMPTY;
        kf->id = kfID;
        kf->timestamp = ts;
        kf->mapId = mapID;
        kfs_[kfID] = kf;
        kfMap_[kfID] = mapID;
        mit->second->addAlreadyDeclaredKF(kf);
        return TMAP_OK;
    }
    static double nsToSec(uint64_t ns) { return (double)(ns / 1000000000ULL) + (double)(ns % 1000000000ULL) * 1e-9; }

    int addKeyFrame(uint64_t ts_ns, uint64_t kfID, uint64_t mapID, const void *xyz, bool on_device, size_t n, size_t stride,
                    const uint32_t *field_offsets = nullptr, int big_endian = 0)
    {
        if (kfID > (uint64_t)std::numeric_limits<int64_t>::max()) return TMAP_ERR_NEGATIVE_ID;  // System.cpp:163-168
        if (!xyz) return TMAP_IGNORED_NULL;
        if (stride < 12 || stride % 4) throw std::invalid_argument("stride_bytes must be a multiple of 4 and >= 12");
        int ff[3] = {0, 1, 2};
        if (field_offsets)
            for (int k = 0; k < 3; ++k)
            {
                if (field_offsets[k] % 4 || (size_t)field_offsets[k] + 4 > stride)
                    throw Unsupported("PointCloud2 x/y/z fields must be 4-byte aligned FLOAT32 inside point_step");
                ff[k] = (int)(field_offsets[k] / 4);
            }
        {
            // cheap pre-checks before spending a prune pass (same outcomes as registerKeyFrame)
            std::lock_guard<std::recursive_mutex> l(m_);
            if (kfs_.count(kfID)) return TMAP_IGNORED_DUPLICATE;
            if (!maps_.count(mapID)) return TMAP_IGNORED_UNKNOWN_MAP;
        }
        float T[12];
        {
            std::lock_guard<std::recursive_mutex> l(m_);
            std::memcpy(T, Tbv_, sizeof(T));
        }
        auto kf = pruneToKeyFrame(xyz, on_device, n, stride, T, P_.max_range, P_.min_range, ff, big_endian);
        return registerKeyFrame(nsToSec(ts_ns), kfID, mapID, kf);
    }
    // System::pushToBuffer (System.cpp:339-344) + PointCloudBuffer::addPointCloud (PointCloudBuffer.cpp:26-32)
    int pushToBuffer(double timestamp, const void *xyz, size_t n, size_t stride)
    {
        if (!P_.use_pointcloud_buffer) return TMAP_OK;  // no buffer was created (System.cpp:38-42): silently nothing
        if (!xyz) return TMAP_IGNORED_NULL;
        if (stride < 12 || stride % 4) throw std::invalid_argument("stride_bytes must be a multiple of 4 and >= 12");
        std::lock_guard<std::recursive_mutex> l(bufMutex_);
        Buffered b{timestamp, nullptr, n, stride};
        {
            std::lock_guard<std::mutex> il(ingestMutex_);
            CK(cudaSetDevice(P_.device));
            CK(cudaMallocAsync(&b.d, std::max<size_t>(n * stride, 16), ingest_));
            if (n) { CK(cudaMemcpyAsync(b.d, xyz, n * stride, cudaMemcpyHostToDevice, ingest_)); g_h2d += (uint64_t)(n * stride); }
            CK(cudaStreamSynchronize(ingest_));  // the caller's buffer may go away after the call returns
        }
        cloudBuf_.push_back(b);
        if (std::abs(cloudBuf_.front().t - cloudBuf_.back().t) > 25.0)
        {
            // deletePointsBefore(timestamp - 5.0) (PointCloudBuffer.cpp:48-53)
            std::lock_guard<std::mutex> il(ingestMutex_);
            while (!cloudBuf_.empty() && cloudBuf_.front().t < timestamp - 5.0)
            {
                cudaFreeAsync(cloudBuf_.front().d, ingest_);
                cloudBuf_.pop_front();
            }
        }
        return TMAP_OK;
    }
    // System::addNewKeyFrameTsDouble (System.cpp:174-186) + getClosestPointCloud (PointCloudBuffer.cpp:34-46)
    int addKeyFrameTs(double timestamp, uint64_t kfID, uint64_t mapID)
    {
        std::lock_guard<std::recursive_mutex> l(bufMutex_);
        if (!P_.use_pointcloud_buffer || cloudBuf_.empty()) return TMAP_IGNORED_NULL;  // "buffer is empty" -> nullptr -> return
        // std::min_element: the FIRST element with the smallest |t - query|
        size_t best = 0;
        for (size_t i = 1; i < cloudBuf_.size(); ++i)
            if (std::abs(cloudBuf_[i].t - timestamp) < std::abs(cloudBuf_[best].t - timestamp)) best = i;
        const Buffered &b = cloudBuf_[best];  // (no id-range check here: System.cpp:174-186 has none)
        {
            std::lock_guard<std::recursive_mutex> ml(m_);
            if (kfs_.count(kfID)) return TMAP_IGNORED_DUPLICATE;
            if (!maps_.count(mapID)) return TMAP_IGNORED_UNKNOWN_MAP;
        }
        float T[12];
        {
            std::lock_guard<std::recursive_mutex> ml(m_);
            std::memcpy(T, Tbv_, sizeof(T));
        }
        auto kf = pruneToKeyFrame(b.d, true, b.n, b.stride, T, P_.max_range, P_.min_range);
        return registerKeyFrame(timestamp, kfID, mapID, kf);
    }
    size_t bufferedClouds()
    {
        std::lock_guard<std::recursive_mutex> l(bufMutex_);
        return cloudBuf_.size();
    }
    int addKeyFrameBase(uint64_t kfID, uint64_t mapID, const float *xyz, size_t n)
    {
        if (n && !xyz) return TMAP_IGNORED_NULL;
        auto kf = std::make_shared<KeyFrame>();
        if (n)
        {
            std::lock_guard<std::mutex> l(ingestMutex_);
            CK(cudaSetDevice(P_.device));
            raw_.ensure(n * 12);
            total_.ensure(16);
            totalH_.ensure(16);
            CK(cudaMemcpyAsync(raw_.p, xyz, n * 12, cudaMemcpyHostToDevice, ingest_));
            g_h2d += (uint64_t)(n * 12);
            CK(cudaMallocAsync(reinterpret_cast<void **>(&kf->d_cloud), n * sizeof(float4), ingest_));
            kf->allocStream = ingest_;
            CK(launch_pack_base(ingest_, raw_.as<float>(), n, kf->d_cloud, total_.as<uint32_t>()));
            g_launches += 1;
            CK(cudaMemcpyAsync(totalH_.p, total_.p, 8, cudaMemcpyDeviceToHost, ingest_));
            g_d2h += (uint64_t)(8);
            CK(cudaStreamSynchronize(ingest_));
            kf->n = (uint32_t)n;
            std::memcpy(&kf->maxnorm, totalH_.as<uint32_t>() + 1, 4);
            if (!std::isfinite(kf->maxnorm)) throw std::invalid_argument("non-finite point in base cloud");
        }
        return registerKeyFrame(0.0, kfID, mapID, kf);
    }
    int updateKeyFrame(uint64_t kfID, const float pose[12])  // System.cpp:195-224
    {
        for (int k = 0; k < 12; ++k)
            if (!std::isfinite(pose[k])) throw std::invalid_argument("non-finite pose");
        std::shared_ptr<LocalMap> mp;
        std::shared_ptr<KeyFrame> kf;
        {
            std::lock_guard<std::recursive_mutex> l(m_);
            auto rit = kfMap_.find(kfID);
            if (rit == kfMap_.end()) return TMAP_IGNORED_UNKNOWN_KF;
            auto mit = maps_.find(rit->second);
            if (mit == maps_.end()) return TMAP_IGNORED_UNKNOWN_MAP;
            mp = mit->second;
            auto kit = kfs_.find(kfID);
            if (kit != kfs_.end()) kf = kit->second;
        }
        mp->setKeyFramePose(kf, pose);
        return TMAP_OK;
    }
    // The registry lock is dropped while the map subtracts the keyframe (grid mutex + a GPU pass) and re-taken
    // to erase the entries: holding m_ across detachKeyFrame would order m_ -> grid mutex against any caller
    // that holds the grid mutex and enters the library again.
    int deleteKeyFrame(uint64_t kfID)  // System.cpp:234-248
    {
        std::shared_ptr<LocalMap> mp;
        {
            std::lock_guard<std::recursive_mutex> l(m_);
            auto rit = kfMap_.find(kfID);
            if (rit == kfMap_.end()) return TMAP_IGNORED_UNKNOWN_KF;
            auto mit = maps_.find(rit->second);
            if (mit != maps_.end()) mp = mit->second;
        }
        if (mp) mp->detachKeyFrame(kfID);
        std::lock_guard<std::recursive_mutex> l(m_);
        kfs_.erase(kfID);
        kfMap_.erase(kfID);
        return TMAP_OK;
    }
    int keyFramePoints(uint64_t kfID, size_t *n_out)  // KeyFrame.hpp:85
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        auto it = kfs_.find(kfID);
        if (it == kfs_.end()) return TMAP_IGNORED_UNKNOWN_KF;
        *n_out = it->second->hasCloud ? (size_t)it->second->n : 0;
        return TMAP_OK;
    }
    int updateKFMap(uint64_t kfID, uint64_t mapID)  // System.cpp:250-276
    {
        std::shared_ptr<LocalMap> from, to;
        {
            std::lock_guard<std::recursive_mutex> l(m_);
            auto rit = kfMap_.find(kfID);
            if (rit == kfMap_.end()) return TMAP_IGNORED_UNKNOWN_KF;
            auto tit = maps_.find(mapID);
            if (tit == maps_.end()) return TMAP_IGNORED_UNKNOWN_MAP;
            if (rit->second == mapID) return TMAP_OK;
            auto fit = maps_.find(rit->second);
            if (fit == maps_.end()) return TMAP_OK;
            from = fit->second;
            to = tit->second;
        }
        auto kf = from->detachKeyFrame(kfID);  // grid mutex of the old map; m_ not held
        if (!kf) return TMAP_OK;
        to->addAlreadyDeclaredKF(kf);
        float p[12];
        kf->getPose(p);
        to->setKeyFramePose(kf, p);
        std::lock_guard<std::recursive_mutex> l(m_);
        auto rit = kfMap_.find(kfID);
        if (rit != kfMap_.end()) rit->second = mapID;
        return TMAP_OK;
    }
    int informLoopClosure()  // System.cpp:278-284
    {
        auto mp = map(0);
        if (mp) mp->clearEntireMap();
        return TMAP_OK;
    }
    int addObstacleKeyFrame(uint64_t ts_ns, const float *xyz, size_t n, size_t stride, uint64_t *id_out)  // System.cpp:300-323
    {
        std::lock_guard<std::recursive_mutex> l(m_);
        *id_out = 0;
        if (!obsSet_) return TMAP_IGNORED_NO_OBSTACLE_PARAMS;
        if (!xyz) return TMAP_IGNORED_NULL;
        auto kf = pruneToKeyFrame(xyz, false, n, stride, Tb_obs_, obsMax_, obsMin_);
        if (kf->n == 0) return TMAP_IGNORED_EMPTY;
        uint64_t id;
        if (!freeObs_.empty()) { id = freeObs_.back(); freeObs_.pop_back(); }
        else id = nextObs_--;
        const int rc = registerKeyFrame(nsToSec(ts_ns), id, 0, kf);
        if (rc != TMAP_OK) { freeObs_.push_back(id); return rc; }
        *id_out = id;
        return TMAP_OK;
    }
    int deleteObstacleKeyFrame(uint64_t id)  // System.cpp:325-335
    {
        {
            std::lock_guard<std::recursive_mutex> l(m_);
            if (!kfMap_.count(id)) return TMAP_IGNORED_UNKNOWN_KF;
        }
        if (deleteKeyFrame(id) != TMAP_OK) return TMAP_IGNORED_UNKNOWN_KF;  // lost a race with another delete
        std::lock_guard<std::recursive_mutex> l(m_);
        freeObs_.push_back(id);
        return TMAP_OK;
    }
    // System::getGlobalPointCloud (System.cpp:370-387): every map's stitched (optionally voxel-filtered) cloud,
    // concatenated.  The reference walks an unordered_map; here maps are visited in ascending id.
    size_t globalPointCloud(float vx, float vy, float vz, float *out, size_t cap)
    {
        std::vector<std::shared_ptr<LocalMap>> ms;
        {
            std::lock_guard<std::mutex> r(mapsM_);
            for (auto &kv : maps_) ms.push_back(kv.second);
        }
        size_t total = 0;
        for (auto &mp : ms)
        {
            const size_t room = out && cap > total ? cap - total : 0;
            total += mp->stitchedVoxel(vx, vy, vz, room ? out + 3 * total : nullptr, room);
        }
        return total;
    }
    int flush()
    {
        std::vector<std::shared_ptr<LocalMap>> ms;
        {
            std::lock_guard<std::recursive_mutex> l(m_);
            for (auto &kv : maps_) ms.push_back(kv.second);
        }
        std::string err;
        for (auto &mp : ms)
        {
            mp->flush();
            std::string e = mp->takeWorkerError();
            if (!e.empty()) err = e;
        }
        if (!err.empty()) throw CudaError("worker: " + err);
        return TMAP_OK;
    }

  private:
    tmap_params P_;
    std::recursive_mutex m_;  // localMapMutex_
    std::mutex mapsM_;        // leaf lock of the maps_ container alone (see map())
    std::map<uint64_t, std::shared_ptr<LocalMap>> maps_;
    std::shared_ptr<LocalMap> current_;
    std::atomic<uint64_t> currentId_{0};
    std::unordered_map<uint64_t, std::shared_ptr<KeyFrame>> kfs_;
    std::unordered_map<uint64_t, uint64_t> kfMap_;  // allKeyFramesSet_
    float Tbv_[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    float Tb_obs_[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    double obsMax_ = 0, obsMin_ = 0;
    bool obsSet_ = false;
    uint64_t nextObs_ = (uint64_t)std::numeric_limits<int64_t>::max();
    std::vector<uint64_t> freeObs_;
    std::mutex ingestMutex_;
    cudaStream_t ingest_ = nullptr;
    DevBuf raw_, warpCnt_, total_;
    PinBuf totalH_;
    // PointCloudBuffer (PointCloudBuffer.cpp): clouds by acquisition time, kept in DEVICE memory (stream-ordered pool)
    struct Buffered
    {
        double t;
        void *d;
        size_t n, stride;
    };
    std::deque<Buffered> cloudBuf_;
    std::recursive_mutex bufMutex_;
};
}  // namespace tmb

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
struct tmap_system
{
    std::unique_ptr<tmb::System> sys;
};

#define TMAP_API_BEGIN try {
#define TMAP_API_END                                                                              \
    }                                                                                             \
    catch (const tmb::Unsupported &e) { tmb::g_last_error = e.what(); return TMAP_ERR_UNSUPPORTED; } \
    catch (const tmb::CudaError &e) { tmb::g_last_error = e.what(); return TMAP_ERR_CUDA; }        \
    catch (const std::invalid_argument &e) { tmb::g_last_error = e.what(); return TMAP_ERR_INVALID; } \
    catch (const std::exception &e)                                                                \
    {                                                                                             \
        tmb::g_last_error = e.what();                                                              \
        return std::string(e.what()) == "NO_DEVICE" ? TMAP_ERR_NO_DEVICE : TMAP_ERR_CUDA;          \
    }

#define NEED(s) if (!(s) || !(s)->sys) { tmb::g_last_error = "null handle"; return TMAP_ERR_INVALID; }
#define NEED_MAP(mp, s, id)                                                                       \
    auto mp = (s)->sys->map(id);                                                                  \
    if (!mp) { tmb::g_last_error = "unknown map"; return TMAP_IGNORED_UNKNOWN_MAP; }

extern "C" {

void tmap_default_params(tmap_params *p)
{
    std::memset(p, 0, sizeof(*p));
    p->resolution = 0.05; p->grid_center_x = 0; p->grid_center_y = 0; p->half_size = 7.5;
    p->extend_length = 30.0; p->robot_radius = 0.20; p->ground_clearance = 0.15; p->max_slope = 0.4;
    p->min_points_per_grid = 3; p->is_kf_optimization_enabled = 1; p->num_local_keyframes = 10;
    p->global_adjustment_sleep = 0; p->inflation_enabled = 0; p->inflation_radius = 1.5;
    p->cost_scaling_factor = 2.0; p->lethal_step_threshold = 0.95; p->robot_height = 3.5;
    p->max_range = 6.0; p->min_range = 1.0; p->use_pointcloud_buffer = 0; p->use_ros_buffer = 0;
    p->publish_rate_hz = 1.0;
    p->device = 0; p->max_batch = 1; p->profile = 0;
}

const char *tmap_last_error(void) { return tmb::g_last_error.c_str(); }

int tmap_create(const tmap_params *p, tmap_system **out)
{
    if (!p || !out) { tmb::g_last_error = "null argument"; return TMAP_ERR_INVALID; }
    *out = nullptr;
    TMAP_API_BEGIN
    auto s = new tmap_system;
    try { s->sys.reset(new tmb::System(*p)); }
    catch (...) { delete s; throw; }
    *out = s;
    return TMAP_OK;
    TMAP_API_END
}

void tmap_destroy(tmap_system *s) { delete s; }

int tmap_set_extrinsics(tmap_system *s, const float T_sv[12], const float T_bs[12])
{
    NEED(s) TMAP_API_BEGIN return s->sys->setExtrinsics(T_sv, T_bs); TMAP_API_END
}
int tmap_add_map(tmap_system *s, uint64_t map_id, tmap_update_cb cb, void *user)
{
    NEED(s) TMAP_API_BEGIN return s->sys->addMap(map_id, cb, user); TMAP_API_END
}
int tmap_set_current_map(tmap_system *s, uint64_t map_id)
{
    NEED(s) TMAP_API_BEGIN return s->sys->setCurrentMap(map_id); TMAP_API_END
}
int tmap_add_keyframe(tmap_system *s, uint64_t ts, uint64_t kf_id, uint64_t map_id, const float *xyz, size_t n, size_t stride)
{
    NEED(s) TMAP_API_BEGIN return s->sys->addKeyFrame(ts, kf_id, map_id, xyz, false, n, stride); TMAP_API_END
}
int tmap_add_keyframe_device(tmap_system *s, uint64_t ts, uint64_t kf_id, uint64_t map_id, const void *d_xyz, size_t n, size_t stride)
{
    NEED(s) TMAP_API_BEGIN return s->sys->addKeyFrame(ts, kf_id, map_id, d_xyz, true, n, stride); TMAP_API_END
}
int tmap_add_keyframe_pc2(tmap_system *s, uint64_t ts_ns, uint64_t kf_id, uint64_t map_id, const void *data, size_t n_points,
                          size_t point_step, uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, int is_bigendian)
{
    NEED(s)
    TMAP_API_BEGIN
    const uint32_t off[3] = {x_offset, y_offset, z_offset};
    return s->sys->addKeyFrame(ts_ns, kf_id, map_id, data, false, n_points, point_step, off, is_bigendian ? 1 : 0);
    TMAP_API_END
}
int tmap_push_to_buffer(tmap_system *s, double timestamp_s, const float *xyz, size_t n, size_t stride_bytes)
{
    NEED(s) TMAP_API_BEGIN return s->sys->pushToBuffer(timestamp_s, xyz, n, stride_bytes); TMAP_API_END
}
int tmap_add_keyframe_ts(tmap_system *s, double timestamp_s, uint64_t kf_id, uint64_t map_id)
{
    NEED(s) TMAP_API_BEGIN return s->sys->addKeyFrameTs(timestamp_s, kf_id, map_id); TMAP_API_END
}
int tmap_buffered_clouds(tmap_system *s, size_t *n_out)
{
    NEED(s)
    if (!n_out) { tmb::g_last_error = "null n_out"; return TMAP_ERR_INVALID; }
    TMAP_API_BEGIN *n_out = s->sys->bufferedClouds(); return TMAP_OK; TMAP_API_END
}
int tmap_add_keyframe_base(tmap_system *s, uint64_t kf_id, uint64_t map_id, const float *xyz, size_t n)
{
    NEED(s) TMAP_API_BEGIN return s->sys->addKeyFrameBase(kf_id, map_id, xyz, n); TMAP_API_END
}
int tmap_update_pose(tmap_system *s, uint64_t kf_id, const float T[12])
{
    NEED(s) TMAP_API_BEGIN return s->sys->updateKeyFrame(kf_id, T); TMAP_API_END
}
int tmap_update_poses(tmap_system *s, const uint64_t *kf_ids, const float *poses12, size_t n)
{
    NEED(s)
    if (n && (!kf_ids || !poses12)) { tmb::g_last_error = "null ids / poses"; return TMAP_ERR_INVALID; }
    TMAP_API_BEGIN
    for (size_t i = 0; i < n; ++i)
    {
        const int rc = s->sys->updateKeyFrame(kf_ids[i], poses12 + 12 * i);
        if (rc < 0) return rc;
    }
    return TMAP_OK;
    TMAP_API_END
}
int tmap_delete_keyframe(tmap_system *s, uint64_t kf_id)
{
    NEED(s) TMAP_API_BEGIN return s->sys->deleteKeyFrame(kf_id); TMAP_API_END
}
int tmap_update_kf_map(tmap_system *s, uint64_t kf_id, uint64_t map_id)
{
    NEED(s) TMAP_API_BEGIN return s->sys->updateKFMap(kf_id, map_id); TMAP_API_END
}
int tmap_inform_loop_closure(tmap_system *s)
{
    NEED(s) TMAP_API_BEGIN return s->sys->informLoopClosure(); TMAP_API_END
}
int tmap_set_obstacle_params(tmap_system *s, const float T[12], double mx, double mn)
{
    NEED(s) TMAP_API_BEGIN return s->sys->setObstacleParams(T, mx, mn); TMAP_API_END
}
int tmap_add_obstacle_keyframe(tmap_system *s, uint64_t ts, const float *xyz, size_t n, size_t stride, uint64_t *id_out)
{
    NEED(s)
    if (!id_out) { tmb::g_last_error = "null id_out"; return TMAP_ERR_INVALID; }
    TMAP_API_BEGIN return s->sys->addObstacleKeyFrame(ts, xyz, n, stride, id_out); TMAP_API_END
}
int tmap_delete_obstacle_keyframe(tmap_system *s, uint64_t kf_id)
{
    NEED(s) TMAP_API_BEGIN return s->sys->deleteObstacleKeyFrame(kf_id); TMAP_API_END
}
int tmap_flush(tmap_system *s)
{
    NEED(s) TMAP_API_BEGIN return s->sys->flush(); TMAP_API_END
}
int tmap_lock(tmap_system *s, uint64_t map_id)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->gridMutex().lock(); return TMAP_OK; TMAP_API_END
}
int tmap_unlock(tmap_system *s, uint64_t map_id)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->gridMutex().unlock(); return TMAP_OK; TMAP_API_END
}
static int keys_out(std::vector<uint64_t> &keys, size_t total, uint64_t *out, size_t cap, size_t *n_out)
{
    if (n_out) *n_out = total;
    for (size_t i = 0; i < keys.size() && i < cap; ++i) out[i] = keys[i];
    return TMAP_OK;
}
int tmap_take_changed_cells(tmap_system *s, uint64_t map_id, uint64_t *out, size_t cap, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    if (!out) { if (n_out) *n_out = mp->countKeys(0); return TMAP_OK; }
    size_t total = 0;
    auto keys = mp->collectKeys(0, true, cap, &total);  // cap < total: nothing is drained, *n_out = total
    return keys_out(keys, total, out, cap, n_out);
    TMAP_API_END
}
int tmap_occupied_cell_keys(tmap_system *s, uint64_t map_id, uint64_t *out, size_t cap, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    if (!out) { if (n_out) *n_out = mp->countKeys(1); return TMAP_OK; }
    size_t total = 0;
    auto keys = mp->collectKeys(1, false, cap, &total);
    return keys_out(keys, total, out, cap, n_out);
    TMAP_API_END
}
int tmap_read_layers(tmap_system *s, uint64_t map_id, const uint64_t *keys, size_t n, const int32_t *layer_ids, size_t n_layers,
                     float *out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    for (size_t l = 0; l < n_layers; ++l)
        if (layer_ids[l] < 0 || layer_ids[l] >= TMAP_NUM_LAYERS) throw std::invalid_argument("bad layer id");
    mp->readLayers(keys, n, layer_ids, n_layers, out);
    return TMAP_OK;
    TMAP_API_END
}
int tmap_fill_sparse_update(tmap_system *s, uint64_t map_id, const int32_t *layer_ids, size_t n_layers, const uint64_t *keys,
                            size_t n_keys, uint64_t *out_keys, float *out_values, size_t *n_out)
{
    NEED(s)
    if (!n_out || (n_keys && (!keys || !out_keys || !out_values)) || (n_layers && !layer_ids))
    {
        tmb::g_last_error = "null sparse-update argument";
        return TMAP_ERR_INVALID;
    }
    TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    *n_out = mp->fillSparseUpdate(layer_ids, n_layers, keys, n_keys, out_keys, out_values);
    return TMAP_OK;
    TMAP_API_END
}
int tmap_grid_extent(tmap_system *s, uint64_t map_id, int32_t *kx, int32_t *ky)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->extent(kx, ky); return TMAP_OK; TMAP_API_END
}
int tmap_export_dense(tmap_system *s, uint64_t map_id, int32_t layer, int32_t ci0, int32_t cj0, int32_t w, int32_t h, float *out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    if (layer < 0 || layer >= TMAP_NUM_LAYERS || w < 0 || h < 0) throw std::invalid_argument("bad layer id or size");
    mp->exportDense(layer, ci0, cj0, w, h, out);
    return TMAP_OK;
    TMAP_API_END
}
int tmap_export_occupancy(tmap_system *s, uint64_t map_id, int32_t ci0, int32_t cj0, int32_t w, int32_t h, int8_t *out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    if (w < 0 || h < 0) throw std::invalid_argument("bad size");
    mp->exportOccupancy(ci0, cj0, w, h, out);
    return TMAP_OK;
    TMAP_API_END
}
int tmap_export_completeness(tmap_system *s, uint64_t map_id, int32_t ci0, int32_t cj0, int32_t w, int32_t h, double publish_res,
                             int8_t *out, size_t cap, int32_t *out_w, int32_t *out_h)
{
    NEED(s)
    if (!out_w || !out_h || w < 0 || h < 0) { tmb::g_last_error = "bad completeness arguments"; return TMAP_ERR_INVALID; }
    TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->exportCompleteness(ci0, cj0, w, h, publish_res, out, cap, out_w, out_h); return TMAP_OK; TMAP_API_END
}
int tmap_num_keyframes(tmap_system *s, uint64_t map_id, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) *n_out = mp->numKeyFrames(); return TMAP_OK; TMAP_API_END
}
int tmap_keyframe_points(tmap_system *s, uint64_t kf_id, size_t *n_out)
{
    NEED(s)
    if (!n_out) { tmb::g_last_error = "null n_out"; return TMAP_ERR_INVALID; }
    TMAP_API_BEGIN return s->sys->keyFramePoints(kf_id, n_out); TMAP_API_END
}
int tmap_queue_depth(tmap_system *s, uint64_t map_id, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) *n_out = mp->queueDepth(); return TMAP_OK; TMAP_API_END
}
int tmap_stitched_cloud(tmap_system *s, uint64_t map_id, float *out_xyz, size_t cap, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    const size_t n = mp->stitched(out_xyz, cap);
    if (n_out) *n_out = n;
    return TMAP_OK;
    TMAP_API_END
}
int tmap_stitched_cloud_voxel(tmap_system *s, uint64_t map_id, float leaf_x, float leaf_y, float leaf_z, float *out_xyz,
                              size_t cap, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    const size_t n = mp->stitchedVoxel(leaf_x, leaf_y, leaf_z, out_xyz, cap);
    if (n_out) *n_out = n;
    return TMAP_OK;
    TMAP_API_END
}
int tmap_current_map(tmap_system *s, uint64_t *map_id_out)
{
    NEED(s)
    if (!map_id_out) { tmb::g_last_error = "null map_id_out"; return TMAP_ERR_INVALID; }
    TMAP_API_BEGIN
    *map_id_out = s->sys->currentMapId();
    return s->sys->map(*map_id_out) ? TMAP_OK : TMAP_IGNORED_UNKNOWN_MAP;
    TMAP_API_END
}
int tmap_global_point_cloud(tmap_system *s, float leaf_x, float leaf_y, float leaf_z, float *out_xyz, size_t cap, size_t *n_out)
{
    NEED(s) TMAP_API_BEGIN
    const size_t n = s->sys->globalPointCloud(leaf_x, leaf_y, leaf_z, out_xyz, cap);
    if (n_out) *n_out = n;
    return TMAP_OK;
    TMAP_API_END
}
int tmap_get_stats(tmap_system *s, uint64_t map_id, tmap_stats *out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->getStats(out); return TMAP_OK; TMAP_API_END
}
int tmap_set_stream(tmap_system *s, uint64_t map_id, void *stream)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->setStream(stream); return TMAP_OK; TMAP_API_END
}

/* ---- sharded global map ---- */
int tmap_shard_unique_id(uint8_t out_id[128])
{
    TMAP_API_BEGIN
    tmb::NcclApi &N = tmb::NcclApi::get();
    if (!N.ok()) throw tmb::Unsupported(N.error);
    ncclUniqueId id;
    ncclResult_t r = N.GetUniqueId(&id);
    if (r != ncclSuccess) throw tmb::CudaError(std::string("ncclGetUniqueId: ") + N.GetErrorString(r));
    std::memcpy(out_id, &id, 128);
    return TMAP_OK;
    TMAP_API_END
}
int tmap_shard_comm_init(tmap_system *s, uint64_t map_id, int32_t rank, int32_t world, const uint8_t id[128])
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->shardCommInit(rank, world, id); return TMAP_OK; TMAP_API_END
}
int tmap_shard_comm_destroy(tmap_system *s, uint64_t map_id)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id) mp->shardCommDestroy(); return TMAP_OK; TMAP_API_END
}
int tmap_shard_rebuild(tmap_system *s, uint64_t map_id, const uint64_t *kf_ids, size_t n, const float *poses, tmap_shard_info *out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    static_assert(sizeof(tmap_shard_info) == sizeof(tmb::LocalMap::ShardInfo), "layout");
    mp->shardRebuild(kf_ids, n, poses, reinterpret_cast<tmb::LocalMap::ShardInfo *>(out));
    return TMAP_OK;
    TMAP_API_END
}
int tmap_shard_update(tmap_system *s, uint64_t map_id, const uint64_t *kf_ids, size_t n, const float *poses, tmap_shard_info *out)
{
    NEED(s) TMAP_API_BEGIN NEED_MAP(mp, s, map_id)
    mp->shardRebuild(kf_ids, n, poses, reinterpret_cast<tmb::LocalMap::ShardInfo *>(out), true);
    return TMAP_OK;
    TMAP_API_END
}
int tmap_shard_plan(const uint32_t *rows_all, int32_t world, int32_t n_rows, int32_t first_global_row, int32_t rank, int32_t *slabs,
                    uint64_t *send, uint64_t *recv, uint64_t *src_base)
{
    TMAP_API_BEGIN
    if (!rows_all || !slabs || !send || !recv || !src_base || world < 1 || rank < 0 || rank >= world)
        throw std::invalid_argument("tmap_shard_plan: bad arguments");
    tmb::LocalMap::shardPlanRows(rows_all, world, n_rows, first_global_row, rank, slabs, send, recv, src_base);
    return TMAP_OK;
    TMAP_API_END
}

}  // extern "C"
