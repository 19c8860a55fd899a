// Device-side data layout and kernel launch wrappers of the B200 voxelizer.  See DESIGN.md §3 for the pipeline:
//
//   bounds -> [host: mesh transform] -> countLeaves -> scan -> emitLeaves(+tile lists) -> sortTileLists
//          -> voxelizeTiles (the hot kernel) -> compact Voxel32 records
//
// Everything here is plain CUDA (no torch types); the C-ABI in o2v_capi.cpp sits on top of o2v::Engine (o2v_engine.h).
#ifndef O2V_KERNELS_CUH
#define O2V_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "o2v_exact.cuh"

namespace o2v {

constexpr uint32_t kTileEdge = 8;  // voxels per tile edge (sample space); 8^3 = 512 voxels = one thread block
constexpr uint32_t kTileVoxels = kTileEdge * kTileEdge * kTileEdge;
constexpr uint32_t kLeafBatch = 32;  // leaves staged in shared memory per round
constexpr uint32_t kWarpFoldMax = 512;  // light tiles up to this many candidates are folded by a thread or a warp, above by a block
constexpr uint32_t kLightMaxCandidates = 4096;  // tiles up to this many candidate voxels take the staged sparse path

/// Input mesh, model space.  All pointers are device pointers.
struct MeshView {
    const float *verts;        // 9 floats per triangle
    const float *uvs;          // 6 floats per triangle or null
    const uint8_t *types;      // TriangleType per triangle or null (= textured if uvs && textures else materialless)
    const float *colors;       // 3 floats per triangle or null (kUntextured)
    const uint32_t *textureIds;  // per-triangle index into textures or null (= 0)
    uint64_t count;
};

/// Sample-space grid and this rank's Z-slab of it.
struct GridView {
    float xf[12];           // mesh -> voxel affine (3x3 row-major, translation)
    uint32_t sampleRes;     // S = resolution * supersampling
    uint32_t tilesPerAxis;  // tiles per axis of the chunk grid: ceil(S / 64) * 8
    uint32_t tileShift;     // log2(tilesPerAxis) when that is a power of two (tile id -> origin by shifts), else 32
    uint32_t gridExtent;    // ceil(S / 64) * 64: the reference voxelizes whole 64^3 chunks (src/obj2voxel.cpp:245-252,580-581)
    uint32_t slabZ0, slabZ1;        // owned voxel z range [z0, z1), multiples of 8
    uint32_t slabTileZ0, slabTileZCount;
    uint32_t supersampling;  // 1 or 2
    uint32_t strategy;       // ColorStrategy
};

/// One leaf of the subdivision (voxel space).  48 bytes, 16-byte aligned: three float4 loads.
struct __align__(16) LeafRecord {
    float v[9];
    uint32_t tri;  // input triangle index (fold order key together with the leaf's position in the leaf array)
    float area;    // area of the WHOLE input triangle
    uint32_t flags;  // kLeafNeedsCull: the leaf's computed normal is not provably robust (slivers) -> exact distance cull
};

struct __align__(16) LeafUv {
    float t[6];
    float pad[2];
};

struct __align__(16) VoxelRecord {  // voxelio Voxel32, types.hpp:59-70: {i32 x, y, z; u32 argb}
    int32_t x, y, z;
    uint32_t argb;
};

/// Counters produced on the device and read back once per run.
struct RunCounters {
    unsigned long long leaves;          // total leaf records
    unsigned long long pairs;           // total (leaf, tile) pairs
    unsigned long long candidateVoxels; // sum of leaf AABB volumes inside the slab (upper bound on contributions)
    unsigned long long activeTiles;     // light + heavy
    unsigned long long bigLightTiles;   // light tiles with more than kWarpFoldMax candidates (block-level fold)
    unsigned long long lightTiles;      // tiles on the staged sparse path (<= kLightMaxCandidates candidate voxels)
    unsigned long long heavyTiles;      // tiles voxelized block-per-tile
    unsigned long long longTiles;       // tiles with more than 32 leaves (block-level list sort)
    unsigned long long survivors;       // sparse path: candidates that passed the SAT prefilter (= exact clips there)
    unsigned long long voxels;          // emitted voxels
    unsigned long long contributions;   // (triangle, voxel) merges: N_contrib of SURVEY §8
    unsigned long long clipCalls;       // exact clips executed (prefilter survivors)
    unsigned long long depthOverflow;   // triangles that hit kMaxSubdivisionDepth
    unsigned long long hugeTriangles;   // triangles listed for the device-wide subdivision walk (o2v_device.cuh)
    unsigned long long droppedTriangles;  // zero-area / non-finite input triangles
    unsigned long long outputOverflow;  // voxels that did not fit the output buffer
    float boundsMin[3];
    float boundsMax[3];
    unsigned int boundsMinBits[3];
    unsigned int boundsMaxBits[3];
    unsigned int tileCursor;
    unsigned int pad;
    // occupancy-only path
    unsigned long long bigLeaves;       // leaves with more than kOccBigVolume candidate voxels
    unsigned long long bigBoxes;        // their 16^3 boxes
    unsigned long long bigTicket;       // emit pass: (table row << 40) | first box, handed out by one atomic
    unsigned long long slabTriangles;   // triangles the slab filter kept (only when the slab is a part of the grid)
    unsigned long long extraLeaves;     // leaves beyond the first of their triangle (the emit pass writes those)
    unsigned long long scanTotal;       // total of the exclusive scan over the per-triangle extra-leaf counts
    unsigned long long ranges;          // rows with voxels the SAT left undecided = entries of OccupancyView::ranges
};

/// Where the count pass lists the huge triangles of a run (o2v_device.cuh) and what the huge passes keep per
/// (triangle, subtree) item.
constexpr uint32_t kHugeSubtreesPerTriangle = 256;  // = kHugeSubtrees (o2v_device.cuh)
struct HugeWork {
    uint32_t *list;     // triangle indices, in the order the count pass met them (any)
    uint32_t *subtree;  // [capacity * 256]: leaves of the subtree; after hugeScanKernel: leaves of the subtrees before it
    uint32_t capacity;  // triangles the two arrays hold
};

/// Descriptor of a light tile: everything the warp needs in one 16-byte load.
struct __align__(16) LightTile {
    uint32_t tile;        // slab-local tile id
    uint32_t listStart;   // offset into tileList
    uint32_t leafCount;   // <= candidates
    uint32_t candidates;  // sum of leaf AABB volumes clipped to the tile, <= kLightMaxCandidates
};

struct TileWork {
    const uint32_t *allTiles;     // slab-local ids of every non-empty tile (list sorting)
    uint32_t allCount;
    const uint32_t *longTiles;    // the subset whose list is longer than one warp (> 32 leaves)
    uint32_t longCount;
    const uint32_t *activeTiles;  // slab-local ids of the HEAVY tiles (block-per-tile kernel)
    const uint32_t *tileStart;    // exclusive scan of tileCount (slab-local tile id -> list offset)
    const uint32_t *tileCount;
    const uint32_t *tileList;     // leaf indices, ascending within a tile after sortTileLists
    uint32_t activeCount;
};

// ---- launch wrappers (o2v_kernels.cu) ----

void launchBounds(const MeshView &mesh, RunCounters *counters, cudaStream_t stream);
void launchFinishBounds(RunCounters *counters, cudaStream_t stream);
void launchPublishCounters(const RunCounters *counters, RunCounters *hostMapped, cudaStream_t stream);
/// *sum += the order-independent 64-bit checksum of `count` records (o2v_kernels.cu: recordHashKernel).
void launchRecordHash(const VoxelRecord *records, unsigned long long count, unsigned long long *sum, int smCount,
                      cudaStream_t stream);

void launchCountLeaves(const MeshView &mesh, const GridView &grid, uint32_t *leafCount, uint32_t *tileCount,
                       uint32_t *tileCandidates, RunCounters *counters, const HugeWork &work, cudaStream_t stream);
/// The listed huge triangles (o2v_device.cuh): leaves per (triangle, subtree) item, then per triangle the scan of its
/// subtrees and its leaf count into perTriangle[tri] (occupancy: leaves beyond the first, and the leaf tallies).
/// `expected` sizes the grids (the number of huge triangles the previous attempt of the run met).
void launchHugeSubtreeScan(const MeshView &mesh, const GridView &grid, const HugeWork &work, RunCounters *counters,
                           uint32_t *perTriangle, bool occupancy, unsigned long long expected, cudaStream_t stream);
void launchHugeCountTiles(const MeshView &mesh, const GridView &grid, const HugeWork &work, uint32_t *tileCount,
                          uint32_t *tileCandidates, RunCounters *counters, unsigned long long expected, cudaStream_t stream);

/// Exclusive scan of n u32 values; total (u64) is written to *total.  scratch must hold scanScratchElems(n) u32.
size_t scanScratchElems(size_t n);
void launchExclusiveScan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *scratch, unsigned long long *total,
                         cudaStream_t stream);

/// Splits the non-empty tiles into light descriptors and the heavy id list (order irrelevant: tiles are independent).
void launchCompactActiveTiles(const uint32_t *tileCount, const uint32_t *tileCandidates, const uint32_t *tileStart,
                              uint32_t tileTotal, uint32_t *allTiles, uint32_t *longTiles, uint32_t *heavyTiles,
                              LightTile *lightTiles, LightTile *bigLightTiles, RunCounters *counters,
                              cudaStream_t stream);

void launchEmitLeaves(const MeshView &mesh, const GridView &grid, const uint32_t *leafOffset, const uint32_t *tileStart,
                      uint32_t *tileFill, LeafRecord *leaves, LeafUv *leafUvs, uint32_t *tileList, uint32_t *pairTile,
                      RunCounters *counters, const HugeWork &work, unsigned long long hugeExpected, cudaStream_t stream);

void launchSortTileLists(const TileWork &work, uint32_t *tileList, cudaStream_t stream);

/// Buffers of the staged sparse path (o2v_sparse.cu).  A "pair" is one (leaf, tile) entry of tileList.
struct SparseView {
    const uint32_t *pairTile;        // per pair: slab-local tile id
    const uint32_t *tileCandidates;  // per tile: candidate voxel count (classifies the tile)
    uint32_t pairCount;
    uint32_t *pairSurvivors;         // per pair (+1): SAT survivors, scanned into pairOffset
    unsigned long long *pairMask;    // per pair with <= 64 candidates: survivor bit per candidate (z, y, x order) ...
    uint32_t *pairBox;               // ... and its tile-local AABB, so the write pass does not redo the SAT
    uint32_t *pairOffset;            // pairCount + 1 entries
    uint4 *entries;                  // per survivor, one 16-byte record: {pair index, tile-local voxel} written by the
                                     // survivors pass, {clip weight bits (0 = no contribution), triangle index} by the clip
    float2 *uvs;                     // per survivor (textured meshes only)
};

constexpr uint32_t kChunkEdge = 64;     // voxels per chunk edge: the unit the occupancy bitmaps are allocated in
constexpr uint32_t kChunkWords = 4096;  // 64-bit words of one chunk bitmap: 512 tiles x 8 layers (32 KB)
constexpr uint32_t kOccBigVolume = 4096;  // leaves with more candidate voxels are classified box by box ...
constexpr uint32_t kOccBoxEdge = 16;      // ... in 16^3 boxes
constexpr uint32_t kOccDirectCandidates = 8;  // see launchOccupancyClassify
constexpr uint32_t kLeafEmpty = 4u;       // LeafRecord::flags on this path: the triangle of this slot has no leaf in the slab

constexpr uint32_t kMaxSlabs = 16;  // devices one job is spread over

/// Multi-device ingest on the occupancy-only path: a device bins its share of the triangles by Z-slab and writes each
/// triangle straight into the memory of the device(s) owning the slab(s) it can reach (peer stores over NVLink).
/// Device s receives what source r sends in region r of its receive buffer (dest[s] points at that region), so no
/// counter is shared between devices: count[s] lives on the source.
struct SlabScatter {
    uint32_t slabs;
    uint32_t bound[kMaxSlabs + 1];    // slab s owns sample-space z in [bound[s], bound[s + 1])
    float *dest[kMaxSlabs];           // region of this source in the receive buffer of slab s's device
    unsigned long long capacity;      // triangles a region holds
    unsigned long long *count;        // [slabs], on the source device: triangles sent to slab s so far
};

/// Buffers of the occupancy-only path (o2v_occupancy.cu): meshes whose every triangle is MATERIALLESS voxelize white
/// whatever the weights are (src/triangle.hpp:186; BLEND of equal colours is exact, MAX keeps a colour), so only the
/// occupancy has to be decided — an order-independent OR into per-chunk bitmaps.
struct OccupancyView {
    uint32_t *chunkFlag;             // one bit per chunk (64^3 OUTPUT voxels) of the slab: some leaf's box reaches it
    uint32_t *chunkSlot;             // per chunk: index of its bitmap
    uint32_t *chunkList;             // per bitmap: its chunk
    uint32_t chunksPerAxis, chunkZ0, chunkTotal;  // chunk id = cx + C * (cy + C * (cz - chunkZ0)), OUTPUT space
    uint32_t shift;                  // log2(supersampling): output voxel = sample voxel >> shift
    uint32_t activeChunks;           // bitmaps in use (host copy of the device count)
    unsigned long long *bits;        // kChunkWords per bitmap: word = tile (x | y << 3 | z << 6) * 8 + layer z,
                                     // bit (x + 8 y) = OUTPUT voxel (x, y, z) of that tile is occupied
    const unsigned long long *emitted;  // null, or the bits earlier pieces of the job have delivered (EngineParams::
                                     // accumulate): the expand kernel leaves those voxels out
    uint4 *ranges;                   // per row with undecided voxels: {leaf, x | y << 16, z | gap << 16, lenA | lenB << 16}
    unsigned long long rangeCapacity;
    uint4 *queue;                    // {leaf, x | y << 16, z, -} of the voxels neither the SAT nor the bitmap decided
    unsigned long long queueCapacity;
    const LeafRecord *extraLeaves;   // leaf firstLeaves + k: the leaves beyond the first of their triangle
    uint32_t firstLeaves;            // = triangles: leaf i < firstLeaves is the first leaf of triangle i (VoxelizeArgs::leaves)
    uint2 *bigLeaves;                // {leaf, first box} of the leaves with more than kOccBigVolume candidates
    uint32_t bigCapacity;
};

struct VoxelizeArgs {
    GridView grid;
    TileWork work;
    const LeafRecord *leaves;
    const LeafUv *leafUvs;  // null when the mesh has no uvs
    MeshView mesh;
    const TextureView *textures;
    uint32_t textureCount;
    VoxelRecord *out;
    float4 *floatOut;  // null, or per record the float (weight, r, g, b) before quantisation (weighted pipeline only)
    unsigned long long outCapacity;
    RunCounters *counters;
    const LightTile *lightTiles;     // <= kWarpFoldMax candidates
    uint32_t lightCount;
    const LightTile *bigLightTiles;  // kWarpFoldMax < candidates <= kLightMaxCandidates
    uint32_t bigLightCount;
    SparseView sparse;
    OccupancyView occ;
    int variant;  // reserved for kernel A/B experiments
    int prefilter;  // 0 disables the conservative SAT prefilter (debug / validation)
    int packedBits;  // occupancy-only path: 0 = `out` receives Voxel32 records; 32 / 64 = it receives packed positions
                     // x | y << 10 | z << 20 (u32) / x | y << 21 | z << 42 (u64) of the all-white voxels instead
    float certainMargin;  // occupancy-only path: shrink of the voxel box for `certain` (certainMarginFor(sample resolution))
};

/// Heavy tiles: one 512-thread block per tile, thread = voxel.
void launchVoxelizeTiles(const VoxelizeArgs &args, int smCount, cudaStream_t stream);
/// Staged sparse path for light tiles.  Every stage is dense over its own work items, with stream compaction in HBM
/// between the stages: (1) thread per (leaf, tile) pair counts / (2) writes the candidate voxels that survive the SAT
/// prefilter, (3) thread per survivor runs the exact clip, (4) warp per tile sorts its contributions by (voxel, list
/// position) and replays the reference's fold order, then writes the Voxel32 records.
void launchSparseSurvivors(const VoxelizeArgs &args, bool write, cudaStream_t stream);
void launchSparseClip(const VoxelizeArgs &args, int smCount, cudaStream_t stream);
void launchSparseFold(const VoxelizeArgs &args, int smCount, cudaStream_t stream);

/// Occupancy-only path (see OccupancyView and the header of o2v_occupancy.cu): count / emit leaves per triangle, classify
/// every candidate voxel with the three-way SAT of o2v_sat.cuh (`certain` -> bitmap, `uncertain` -> queue), exact clip
/// for the queue, bitmap -> Voxel32 records.
void launchOccupancySlabFilter(const MeshView &mesh, const GridView &grid, float *kept, RunCounters *counters,
                               int smCount, cudaStream_t stream);
void launchOccupancyCount(const MeshView &mesh, const GridView &grid, const OccupancyView &occ, uint32_t *extraCount,
                          LeafRecord *firstLeaves, RunCounters *counters, bool countFromFilter, const HugeWork &work,
                          unsigned long long hugeExpected, int smCount, cudaStream_t stream);
void launchOccupancyAssignChunks(const OccupancyView &occ, RunCounters *counters, cudaStream_t stream);
/// Every chunk of the slab gets a bitmap: slot = chunk (jobs that accumulate pieces, EngineParams::accumulate).
void launchOccupancyAllChunks(const OccupancyView &occ, cudaStream_t stream);
void launchOccupancySlabScatter(const MeshView &mesh, const GridView &grid, const SlabScatter &scatter, int smCount,
                                cudaStream_t stream);
/// histogram[r] += triangles whose z range reaches row r (rows of `unit` sample-space layers, rows <= 128).
void launchOccupancyZHistogram(const MeshView &mesh, const GridView &grid, uint32_t unit, uint32_t rows,
                               unsigned long long *histogram, int smCount, cudaStream_t stream);
/// chunkCounts[slot] = occupied voxels of bitmap `slot`; their sum is added to counters->voxels.
void launchOccupancyChunkCount(const OccupancyView &occ, uint32_t *chunkCounts, RunCounters *counters, int smCount,
                               cudaStream_t stream);
void launchOccupancyEmit(const MeshView &mesh, const GridView &grid, const OccupancyView &occ,
                         const uint32_t *leafOffset, LeafRecord *leaves, RunCounters *counters, const HugeWork &work,
                         unsigned long long hugeExpected, int smCount, cudaStream_t stream);
/// microLeaves: the mesh averages at most kOccDirectCandidates candidate voxels per leaf — classified thread = leaf
/// (occupancyClassifyDirectKernel) instead of block = 64 leaves.
void launchOccupancyClassify(const VoxelizeArgs &args, unsigned long long leafTotal, bool microLeaves, uint32_t bigCount,
                             unsigned long long boxTotal, int smCount, cudaStream_t stream);
void launchOccupancyFilterQueue(const VoxelizeArgs &args, int smCount, cudaStream_t stream);
void launchOccupancyClip(const VoxelizeArgs &args, int smCount, cudaStream_t stream);
void launchOccupancyExpand(const VoxelizeArgs &args, int smCount, cudaStream_t stream);

}  // namespace o2v

#endif  // O2V_KERNELS_CUH
