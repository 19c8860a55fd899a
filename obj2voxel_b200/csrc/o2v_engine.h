// Host-side driver of the B200 voxelizer: owns the device buffers of one GPU and runs the kernel pipeline of
// o2v_kernels.cuh on a caller-provided stream.  Mirrors what src/obj2voxel.cpp:467-520 (voxelize_specialized) does on the
// host for the reference: bounds -> transform -> (chunk sort | tile binning) -> voxelize -> sink.
#ifndef O2V_ENGINE_H
#define O2V_ENGINE_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "o2v_kernels.cuh"

namespace o2v {

/// src/obj2voxel.cpp:370-402 (computeMeshTransform), exact binary32, host side (o2v_host_math.cpp, -ffp-contract=off).
void computeMeshTransform(const float meshMin[3], const float meshMax[3], uint32_t sampleResolution, const int unit[9],
                          float out[12]);

struct EngineParams {
    uint32_t resolution = 0;
    uint32_t supersampling = 1;
    uint8_t strategy = kMax;
    bool boundsKnown = false;
    float bounds[6] = {0, 0, 0, 0, 0, 0};
    int unitTransform[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    uint32_t slabZ0 = 0, slabZ1 = 0;  // owned sample-space z range, multiples of 8; 0,0 = whole grid
    int variant = -1;                 // -1 = default kernel variant
    int prefilter = 1;
    int occupancyPath = 1;            // 1 = all-MATERIALLESS meshes take the occupancy-only path; 0 = always fold weights
    bool slabFiltered = false;        // the mesh is the output of filterSlab() for this very slab: skip the filter pass
    bool bitmapResult = false;        // occupancy-only path: stop at the bitmaps (no records on the device); the caller
                                      // downloads them and expands them on the host (see Engine::bitmapResult)
    bool floatRecords = false;        // parity / debug: also keep every voxel's float (weight, r, g, b) as the fold left it,
                                      // before the ARGB8 truncation (forces the weighted pipeline; see floatRecords())
    bool packedResult = false;        // occupancy-only path: packed positions instead of 16-byte records (every voxel is
                                      // white): 4 bytes per voxel while the output grid fits 10 bits per axis, else 8
                                      // (see Engine::packedBits); a host-to-host job expands them on the host
    int accumulate = 0;               // occupancy-only path, a job whose triangles arrive in pieces: 1 = first piece (every
                                      // chunk of the slab gets a bitmap, cleared), 2 = a later piece (the bitmaps stay);
                                      // a piece's result = the voxels no earlier piece has set (Engine::accumulateBytes)
};

/// What a run with EngineParams::bitmapResult leaves on the device: `chunks` bitmaps of kChunkWords 64-bit words
/// (layout: OccupancyView::bits, OUTPUT space), the chunk id of each (cx + C * (cy + C * (cz - chunkZ0))) and its number
/// of occupied voxels.  Every occupied voxel is white (0xFFFFFFFF).
struct BitmapResult {
    const unsigned long long *bits = nullptr;
    const uint32_t *chunkIds = nullptr;
    const uint32_t *chunkCounts = nullptr;
    uint32_t chunks = 0;
    uint32_t chunksPerAxis = 0, chunkZ0 = 0;
};

struct RunStats {
    RunCounters counters;
    float transform[12];
    float msTotal = 0;     // device time of the whole pipeline (CUDA events on the run stream)
    float msSetup = 0;     // bounds + count + scans + emit + sort
    float msVoxelize = 0;  // clip + fold + heavy tiles
    float msClip = 0;      // the exact-clip kernel alone (sparseClipKernel / occupancyClipKernel)
    float msClassify = 0;  // occupancy-only path: the SAT classification kernel alone
    float msFilter = 0;    // occupancy-only path: undecided ranges -> clip queue
    float msExpand = 0;    // occupancy-only path: bitmap -> records
    int voxelizeLaunches = 0;
    int kernelLaunches = 0;
    unsigned long long outCapacity = 0;
    bool occupancyPath = false;  // this run took the occupancy-only path
    unsigned long long slabTriangles = 0;  // occupancy-only path: triangles the passes worked on (= kept by the slab filter)
    unsigned long long downloadBytes = 0;  // host-to-host jobs (o2v_job.cpp): bytes copied device -> host
};

class DeviceBuffer {
public:
    DeviceBuffer() = default;
    ~DeviceBuffer();
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    /// Grow-only; returns false (and leaves the old allocation) when cudaMalloc fails.
    bool ensure(size_t bytes);
    template <typename T>
    T *as() const
    {
        return static_cast<T *>(ptr_);
    }
    size_t size() const { return size_; }
    void swap(DeviceBuffer &other)
    {
        std::swap(ptr_, other.ptr_);
        std::swap(size_, other.size_);
    }

private:
    void *ptr_ = nullptr;
    size_t size_ = 0;
};

class Engine {
public:
    /// Fails (returns nullptr, message in *error) when no CUDA device is usable: there is no CPU fallback.
    static Engine *create(int device, std::string *error);
    ~Engine();

    int device() const { return device_; }
    int smCount() const { return smCount_; }

    /// Voxelizes a device-resident mesh.  textures: HOST array of TextureView whose pixel pointers are DEVICE pointers.
    /// Result stays on the device (deviceVoxels / voxelCount) until the next call.  Returns 0 or a negative error.
    /// Bytes of the two bitmap sets a job with EngineParams::accumulate keeps for this grid / slab (0: grid too large).
    static size_t accumulateBytes(const EngineParams &params);
    int voxelize(const MeshView &meshIn, const TextureView *textures, uint32_t textureCount, const EngineParams &params,
                 cudaStream_t stream, RunStats *stats);

    const VoxelRecord *deviceVoxels() const { return out_.as<VoxelRecord>(); }
    /// float4 (weight, r, g, b) per record of the last run, index-aligned with deviceVoxels(); null unless the run had
    /// EngineParams::floatRecords.
    const float *floatRecords() const { return floatValid_ ? floatOut_.as<float>() : nullptr; }
    unsigned long long voxelCount() const { return voxelCount_; }
    const std::string &lastError() const { return error_; }

    /// Two output buffers: after the swap the result of the last run stays where it is (the caller keeps the pointer
    /// it got from deviceVoxels) while the next run writes the other buffer.
    void swapOutputBuffers() { out_.swap(outSpare_); }

    /// Copies the result of the last run to host memory (count * 16 bytes) on `stream` and synchronises it.
    int download(void *hostDst, cudaStream_t stream);

    /// 0: deviceVoxels() holds Voxel32 records; 32 / 64: the last run (packedResult, occupancy-only path) left packed
    /// positions there instead — x | y << 10 | z << 20 as u32, or x | y << 21 | z << 42 as u64 — voxelCount() of them.
    int packedBits() const { return packedBits_; }
    /// true if the last run ended with bitmaps instead of records (bitmapResult was asked for and the run took the
    /// occupancy-only path); voxelCount() is valid either way.
    bool hasBitmapResult() const { return bitmapValid_; }
    BitmapResult bitmapResult() const { return bitmap_; }
    /// Like swapOutputBuffers, for the bitmaps: the result of the last run stays put while the next run writes others.
    void swapBitmapBuffers()
    {
        tileBits_.swap(tileBitsSpare_);
        chunkList_.swap(chunkListSpare_);
        chunkCounts_.swap(chunkCountsSpare_);
    }

    /// Bounds of a device-resident mesh (src/obj2voxel.cpp:180-200 findMeshBounds) — one device's share of a job.
    int meshBounds(const MeshView &mesh, cudaStream_t stream, float outMin[3], float outMax[3]);

    /// Triangles of `mesh` per row of `unit` sample-space z layers (rows <= 128 rows from z = 0): the work estimate a job
    /// over several devices balances its Z-slabs by.  params must carry the mesh bounds.  Synchronises `stream`.
    int zRowHistogram(const MeshView &mesh, const EngineParams &params, uint32_t unit, uint32_t rows, cudaStream_t stream,
                      unsigned long long *outHistogram);

    /// Region `source` (of `sources`) of this device's receive buffer for a multi-device job: room for `capacity`
    /// triangles per region.  Peers write into it directly (SlabScatter).
    float *receiveRegion(uint32_t source, uint32_t sources, unsigned long long capacity);
    /// Copies the first counts[r] triangles of every region into one dense array; returns it (nullptr: out of memory).
    const float *packReceived(const unsigned long long *counts, uint32_t sources, unsigned long long capacity,
                              cudaStream_t stream, unsigned long long *total);
    /// Bins `mesh` (this device's share) by Z-slab and stores every triangle into scatter.dest of the slabs it can reach;
    /// sentCounts[s] = triangles sent to slab s.  params must carry the mesh bounds.  Synchronises `stream`.
    int scatterToSlabs(const MeshView &mesh, const EngineParams &params, SlabScatter scatter, cudaStream_t stream,
                       unsigned long long *sentCounts);

    /// The ingest step of a multi-GPU job on the occupancy-only path: copies the triangles of `mesh` whose z range can
    /// reach the slab of `params` into an engine-owned dense array (*kept, valid until the next filterSlab).  A later
    /// voxelize() on that array with EngineParams::slabFiltered never reads the rest of the mesh.  Order is arbitrary.
    int filterSlab(const MeshView &mesh, const EngineParams &params, cudaStream_t stream, const float **kept,
                   unsigned long long *keptCount);

    /// Order-independent 64-bit checksum of the last run's records, computed on the device (see launchRecordHash):
    /// parts of a job and ranks of a multi-GPU job add theirs, and the sum is comparable with the reference's.
    int resultHash(cudaStream_t stream, unsigned long long *out);

    /// Engine-owned pinned host buffer `slot` (0 or 1) of at least `bytes` (grow-only); nullptr if pinning fails.
    void *pinnedStaging(int slot, size_t bytes);

private:
    Engine() = default;
    int fail(int code, const std::string &message);
    /// The occupancy-only pipeline (o2v_occupancy.cu) for an all-MATERIALLESS mesh; kOccupancyFallback if its bitmaps do
    /// not fit device memory (the caller then runs the weighted pipeline).
    int setupGrid(const MeshView &mesh, const EngineParams &params, cudaStream_t stream, RunStats &st, GridView &grid,
                  bool *emptySlab);
    int voxelizeOccupancy(const MeshView &mesh, const EngineParams &params, const GridView &grid, cudaStream_t stream,
                          RunStats &st);
    DeviceBuffer emittedBits_;
    static constexpr int kOccupancyFallback = 1;
    /// The count pass met huge triangles (o2v_device.cuh) without room to list them all: the run starts over with the list
    /// sized for them.  Sticky until a run meets none.
    static constexpr int kRetryHugeWalk = 2;
    bool walkHuge_ = false;
    unsigned long long hugeExpected_ = 0;
    DeviceBuffer hugeList_, hugeSubtree_;
    /// The list of this attempt (capacity 0 while no run has met a huge triangle).
    bool hugeWork(HugeWork &work);
    int voxelizeOnce(const MeshView &meshIn, const TextureView *textures, uint32_t textureCount, const EngineParams &params,
                     cudaStream_t stream, RunStats *stats);

    int device_ = 0;
    int smCount_ = 0;
    size_t totalMemory_ = 0;
    std::string error_;
    unsigned long long voxelCount_ = 0;
    bool bitmapValid_ = false;
    int packedBits_ = 0;
    bool floatValid_ = false;
    BitmapResult bitmap_;

    void *staging_[2] = {nullptr, nullptr};  // pinned, for host sinks
    size_t stagingBytes_[2] = {0, 0};
    RunCounters *hostCounters_ = nullptr;  // pinned + mapped: the device publishes the counters into it
    RunCounters *hostCountersDevice_ = nullptr;  // its device-side address
    RunCounters *hostCountersInit_ = nullptr;  // pinned template
    cudaEvent_t evStart_ = nullptr, evSetup_ = nullptr, evVoxStart_ = nullptr, evVoxEnd_ = nullptr;
    cudaEvent_t evClipStart_ = nullptr, evClipEnd_ = nullptr, evClassifyStart_ = nullptr, evFilterStart_ = nullptr;

    DeviceBuffer hash_, floatOut_, counters_, leafCount_, leafOffset_, tileCount_, tileStart_, tileFill_, tileCand_, activeTiles_, lightTiles_, bigLightTiles_,
        scratch_;
    DeviceBuffer leaves_, leafUvs_, tileList_, out_, outSpare_, textures_;
    DeviceBuffer allTiles_, longTiles_, pairTile_, pairSurvivors_, pairOffset_, pairMask_, pairBox_, entries_, contribUvs_;
    DeviceBuffer chunkFlag_, chunkSlot_, chunkList_, tileBits_, occQueue_, occRanges_, bigLeaves_, slabVerts_, slabKept_, extraLeaves_;
    DeviceBuffer tileBitsSpare_, chunkListSpare_, chunkCounts_, chunkCountsSpare_, received_, receivedPacked_, scatterCounts_;  // occupancy-only path
};

// error codes of Engine::voxelize / the additive C-ABI (include/obj2voxel_b200.h)
enum EngineError {
    kErrOk = 0,
    kErrCuda = -1,
    kErrBadParams = -2,
    kErrOutOfMemory = -3,
    kErrTooLarge = -4,
};

}  // namespace o2v

#endif  // O2V_ENGINE_H
