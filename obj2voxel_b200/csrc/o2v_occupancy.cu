// Occupancy-only path (see OccupancyView in o2v_kernels.cuh).
//
// A mesh whose every triangle is MATERIALLESS (any STL, any OBJ without materials, obj2voxel_set_triangle_basic /
// _colored — SURVEY fact 8) voxelizes to 0xFFFFFFFF wherever a voxel is occupied: colorAt_f returns white
// (src/triangle.hpp:186), BLEND of equal colours is (w1 + w2) / (w1 + w2) = 1 exactly, MAX keeps one of two whites, the
// 2x downscale combines whites.  The weights — the only thing the fold order and the piece count feed — never reach the
// output, so the result is the set { voxel : some leaf's exact clip has >= 1 piece }, an order-independent OR.  No tile
// lists, no sort, no fold:
//
//   filter    thread = triangle           only for a rank that owns a part of the grid: the triangles whose z range can
//                                          reach the slab, copied into a dense array (everything below reads that)
//   count     thread = triangle           transform, subdivision DFS (exact), statistics, touched 64^3 chunks — and the
//                                          triangle's first leaf goes straight into leaf slot i
//   emit      thread = triangle           only if some triangle subdivides or some leaf is big: the other leaves; leaves
//                                          with more than 4096 candidates enter the big-leaf table (16^3 boxes)
//   classify  thread = row segment        block = 64 leaves (or one 16^3 box): SAT constants staged in shared memory, the
//                                          batch's row segments (<= 8 voxels of one row in one tile column) as ONE flat
//                                          index space (balanced whatever the box sizes); three-way SAT per voxel with one
//                                          multiply-add per axis (o2v_sat.cuh): `certain` -> one RED per segment into the
//                                          chunk bitmap, `uncertain` -> queue (unless the bitmap already decides it)
//   clip      persistent lanes            the bit-exact six-plane clip (WarpClipper) for queued voxels only
//   expand    lane = tile, then = record  bitmap (OR-reduced 2x2x2 when supersampling) -> compacted Voxel32 records,
//                                          rank / select in shared memory so that a warp stores 512 contiguous bytes
//
// filter, count and emit stream the triangle array through shared memory with bulk-async copies (cp.async.bulk + mbarrier,
// the TMA engine: streamTriangles in o2v_device.cuh).
//
// Exactness: `miss` and `certain` are proofs about the reference's result (o2v_sat.cuh header; fuzzed by
// tests/test_sat_classifier.py), everything else runs the reference arithmetic.  prefilter = 0 sends every candidate
// through the exact clip (validation).
#include "o2v_device.cuh"

namespace o2v {

namespace {

constexpr int kOccSetupThreads = 128;
#ifndef O2V_OCC_BATCH
#define O2V_OCC_BATCH 64
#endif
#ifndef O2V_OCC_THREADS
#define O2V_OCC_THREADS 128
#endif
constexpr int kOccBatch = O2V_OCC_BATCH;      // leaves per classify block
constexpr int kOccThreads = O2V_OCC_THREADS;  // threads per classify block
constexpr int kOccClipThreads = 128;
constexpr int kOccExpandThreads = 128;
constexpr int kOccRefillThreshold = 8;
constexpr uint32_t kOccMaybeCap = 1024;  // undecided voxels a block buffers before filtering them against the bitmap

// ---------------------------------------------------------------------------------------------------------------------
// addressing

/// Clips a leaf's voxel AABB to the chunk grid and this rank's slab (the candidate set of voxelizeSubTriangle,
/// src/voxelization.cpp:440-447, restricted to the voxels this rank owns).  false if nothing is left.
__device__ __forceinline__ bool leafBoxInSlab(const float *v, const GridView &grid, uint32_t lo[3], uint32_t hi[3])
{
    triVoxelBounds(v, lo, hi);
    // voxels beyond the chunk grid belong to chunks the reference never dispatches (obj2voxel.cpp:503-505)
    hi[0] = min(hi[0], grid.gridExtent);
    hi[1] = min(hi[1], grid.gridExtent);
    lo[2] = max(lo[2], grid.slabZ0);
    hi[2] = min(hi[2], min(grid.slabZ1, grid.gridExtent));
    return lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2];
}

/// Index of the 64-bit bitmap word (layer z of the voxel's tile) and the voxel's bit in it.
__device__ __forceinline__ size_t bitmapWord(const OccupancyView &occ, uint32_t x, uint32_t y, uint32_t z)
{
    const uint32_t chunk = (x >> 6) + occ.chunksPerAxis * ((y >> 6) + occ.chunksPerAxis * ((z >> 6) - occ.chunkZ0));
    const uint32_t tileLocal = ((x >> 3) & 7u) | (((y >> 3) & 7u) << 3) | (((z >> 3) & 7u) << 6);
    return (size_t) __ldg(occ.chunkSlot + chunk) * kChunkWords + tileLocal * kTileEdge + (z & 7u);
}

__device__ __forceinline__ unsigned long long bitmapBit(uint32_t x, uint32_t y)
{
    return 1ull << ((x & 7u) + 8u * (y & 7u));
}

/// Bits of `m` (layout x + 8 y) smeared over their 2x2 xy blocks: the footprint of the parents that already have a child.
__device__ __forceinline__ unsigned long long smear2x2(unsigned long long m)
{
    const unsigned long long evenX = (m | (m >> 1)) & 0x5555555555555555ull;
    const unsigned long long pairX = evenX | (evenX << 1);
    const unsigned long long evenY = (pairX | (pairX >> 8)) & 0x00ff00ff00ff00ffull;
    return evenY | (evenY << 8);
}

/// true if the bitmap already decides voxel (x, y, z): its bit is set or — when downscaling — its parent has a child.
__device__ __forceinline__ bool alreadyDecided(const OccupancyView &occ, bool downscale, uint32_t x, uint32_t y,
                                               uint32_t z)
{
    const size_t word = bitmapWord(occ, x, y, z);
    unsigned long long known = __ldcg(occ.bits + word);
    if (downscale) {
        known = smear2x2(known | __ldcg(occ.bits + (word ^ 1u)));  // z ^ 1 is the neighbouring word of the same tile
    }
    return (known & bitmapBit(x, y)) != 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// count / emit: thread per triangle

__device__ __forceinline__ uint32_t boxCountOf(const uint32_t *lo, const uint32_t *hi)
{
    return ((hi[0] - lo[0] + kOccBoxEdge - 1) / kOccBoxEdge) * ((hi[1] - lo[1] + kOccBoxEdge - 1) / kOccBoxEdge) *
           ((hi[2] - lo[2] + kOccBoxEdge - 1) / kOccBoxEdge);
}

/// Only when the rank's slab is a part of the grid: copies the triangles whose z range can reach the slab into a dense
/// array, so that the count and emit passes neither read nor diverge on the others (with N ranks, (N - 1) / N of the
/// mesh).  A block collects what it keeps in shared memory and appends it to the array kOccSetupThreads or more
/// triangles at a time: one atomic and one coalesced copy per append.  The order of the array is arbitrary; the
/// occupancy result is an OR, order-free.
__global__ void __launch_bounds__(kOccSetupThreads)
occupancySlabFilterKernel(MeshView mesh, GridView grid, float *__restrict__ kept, RunCounters *counters)
{
    __shared__ TriangleBatch<kOccSetupThreads> batch;
    __shared__ float keep[2 * kOccSetupThreads * 9];
    __shared__ uint32_t warpCount[kOccSetupThreads / 32];
    __shared__ unsigned long long appendAt;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t held = 0;  // block-uniform: triangles waiting in `keep`, < kOccSetupThreads between batches

    auto append = [&]() {  // all threads; `held` is block-uniform
        if (tid == 0) {
            appendAt = atomicAdd(&counters->slabTriangles, (unsigned long long) held);
        }
        __syncthreads();
        float *dst = kept + appendAt * 9;
        for (uint32_t k = tid; k < held * 9u; k += kOccSetupThreads) {
            dst[k] = keep[k];
        }
        __syncthreads();
        held = 0;
    };

    streamTriangles<kOccSetupThreads>(mesh.verts, mesh.count, batch,
                                      [&](unsigned long long, const float in[9], bool valid) {
        const bool mine = valid && !triangleMissesSlab(grid, in);
        const unsigned int votes = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) {
            warpCount[warp] = (uint32_t) __popc(votes);
        }
        __syncthreads();
        uint32_t slot = held + (uint32_t) __popc(votes & ((1u << lane) - 1u)), added = 0;
#pragma unroll
        for (uint32_t w = 0; w < kOccSetupThreads / 32; ++w) {
            slot += w < warp ? warpCount[w] : 0u;
            added += warpCount[w];
        }
        if (mine) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                keep[slot * 9 + k] = in[k];  // stride 9 words: conflict-free
            }
        }
        __syncthreads();
        held += added;
        if (held >= (uint32_t) kOccSetupThreads) {
            append();
        }
    });
    if (held != 0) {
        append();
    }
}

constexpr uint32_t kOccBlockChunkWords = 2048;  // chunk bits a block collects in shared memory: 65536 chunks (8 KB)

/// The one pass every triangle takes: transform, subdivision DFS, statistics, chunk marks — and the triangle's first leaf
/// goes straight into leaf slot i, so that a mesh whose triangles are all leaves themselves (anything fine relative to
/// the grid) needs no second pass.  extraCount[i] = the triangle's leaves beyond the first.
__global__ void __launch_bounds__(kOccSetupThreads)
occupancyCountKernel(MeshView mesh, GridView grid, OccupancyView occ, uint32_t *__restrict__ extraCount,
                     LeafRecord *__restrict__ firstLeaves, RunCounters *counters, bool countFromFilter)
{
    if (countFromFilter) {
        // the array is what the slab filter kept: its length is still on the device (no host round trip in between)
        mesh.count = counters->slabTriangles;
    }
    __shared__ TriangleBatch<kOccSetupThreads> batch;
    // Millions of leaves mark a few thousand chunks: up to 65536 chunks per slab (any grid up to 2560^3, and slabs of
    // larger ones) a block ORs its marks into shared memory and publishes each word once when it is done, so that the
    // hot words see a few reads per block instead of a read per leaf.
    __shared__ uint32_t chunkBits[kOccBlockChunkWords];
    const uint32_t chunkWords = (occ.chunkTotal + 31u) / 32u;
    const bool collect = chunkWords <= kOccBlockChunkWords;
    if (collect) {
        for (uint32_t w = threadIdx.x; w < chunkWords; w += kOccSetupThreads) {
            chunkBits[w] = 0;
        }
    }
    // (streamTriangles starts with a barrier)
    unsigned long long candidates = 0, dropped = 0, overflow = 0, bigLeaves = 0, bigBoxes = 0, leafTally = 0,
                       extraTally = 0;
    streamTriangles<kOccSetupThreads>(mesh.verts, mesh.count, batch,
                                      [&](unsigned long long i, const float in[9], bool valid) {
        if (!valid) {
            return;
        }
        Tri<false> root;
        float area;
        uint32_t leaves = 0;
        if (setupTriangle<false>(grid, in, root, area)) {
            const bool ok = traverseLeaves<false>(root, grid, [&](const Tri<false> &leaf, const uint32_t *lo,
                                                                   const uint32_t *hi) {
                if (leaves == 0) {
                    LeafRecord rec;
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        rec.v[k] = leaf.v[k];
                    }
                    rec.tri = static_cast<uint32_t>(i);  // position in the array this pass reads (unused on this path)
                    rec.area = area;
                    rec.flags = leafFlagsOf(leaf.v);
                    firstLeaves[i] = rec;
                }
                ++leaves;
                const unsigned long long volume =
                    (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
                candidates += volume;
                if (volume > kOccBigVolume) {
                    ++bigLeaves;
                    bigBoxes += boxCountOf(lo, hi);
                }
                for (uint32_t cz = lo[2] >> 6; cz <= (hi[2] - 1) >> 6; ++cz) {
                    for (uint32_t cy = lo[1] >> 6; cy <= (hi[1] - 1) >> 6; ++cy) {
                        for (uint32_t cx = lo[0] >> 6; cx <= (hi[0] - 1) >> 6; ++cx) {
                            const uint32_t chunk = cx + occ.chunksPerAxis * (cy + occ.chunksPerAxis * (cz - occ.chunkZ0));
                            const uint32_t bit = 1u << (chunk & 31u);
                            if (collect) {
                                if ((chunkBits[chunk >> 5] & bit) == 0) {
                                    atomicOr(&chunkBits[chunk >> 5], bit);
                                }
                            }
                            else if ((__ldcg(occ.chunkFlag + (chunk >> 5)) & bit) == 0) {  // look before the atomic
                                atomicOr(occ.chunkFlag + (chunk >> 5), bit);
                            }
                        }
                    }
                }
            });
            overflow += ok ? 0 : 1;
        }
        else {
            ++dropped;
        }
        if (leaves == 0) {  // only the flags of an empty slot are ever read
            reinterpret_cast<float4 *>(firstLeaves + i)[2] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(kLeafEmpty));
        }
        extraCount[i] = leaves > 1u ? leaves - 1u : 0u;
        leafTally += leaves;
        extraTally += leaves > 1u ? leaves - 1u : 0u;
    });
    if (collect) {
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < chunkWords; w += kOccSetupThreads) {
            const uint32_t marks = chunkBits[w];
            if (marks != 0 && (marks & ~__ldcg(occ.chunkFlag + w)) != 0) {
                atomicOr(occ.chunkFlag + w, marks);
            }
        }
    }
    warpTally(&counters->leaves, leafTally);
    warpTally(&counters->extraLeaves, extraTally);
    warpTally(&counters->candidateVoxels, candidates);
    warpTally(&counters->droppedTriangles, dropped);
    warpTally(&counters->depthOverflow, overflow);
    warpTally(&counters->bigLeaves, bigLeaves);
    warpTally(&counters->bigBoxes, bigBoxes);
}

__global__ void occupancyAssignChunksKernel(OccupancyView occ, RunCounters *counters)
{
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const bool touched = chunk < occ.chunkTotal && ((occ.chunkFlag[chunk >> 5] >> (chunk & 31u)) & 1u) != 0;
    const unsigned int ballot = __ballot_sync(0xffffffffu, touched);
    if (ballot == 0) {
        return;
    }
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(&counters->activeTiles, (unsigned long long) __popc(ballot));  // here: active chunks
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (touched) {
        const uint32_t slot = (uint32_t) base + __popc(ballot & ((1u << lane) - 1u));
        occ.chunkSlot[chunk] = slot;
        occ.chunkList[slot] = chunk;
    }
}

/// Second pass, only for meshes that need it (some triangle subdivides, or some leaf is big): writes the leaves beyond
/// the first of each triangle to extraLeaves[extraOffset[i] ...] (leaf index firstLeaves + that) and enters the leaves
/// with more than kOccBigVolume candidates — first leaves included — into the big-leaf table.
__global__ void __launch_bounds__(kOccSetupThreads)
occupancyEmitKernel(MeshView mesh, GridView grid, OccupancyView occ, const uint32_t *__restrict__ extraOffset,
                    LeafRecord *__restrict__ extraLeaves, RunCounters *counters)
{
    __shared__ TriangleBatch<kOccSetupThreads> batch;
    streamTriangles<kOccSetupThreads>(mesh.verts, mesh.count, batch,
                                      [&](unsigned long long i, const float in[9], bool valid) {
        Tri<false> root;
        float area;
        if (!valid || !setupTriangle<false>(grid, in, root, area)) {
            return;
        }
        uint32_t seen = 0;
        const uint32_t extraAt = extraOffset[i];
        traverseLeaves<false>(root, grid, [&](const Tri<false> &leaf, const uint32_t *lo, const uint32_t *hi) {
            const uint32_t index = seen == 0 ? static_cast<uint32_t>(i) : occ.firstLeaves + extraAt + (seen - 1u);
            if (seen != 0) {
                LeafRecord rec;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    rec.v[k] = leaf.v[k];
                }
                rec.tri = static_cast<uint32_t>(i);
                rec.area = area;
                rec.flags = leafFlagsOf(leaf.v);
                extraLeaves[extraAt + (seen - 1u)] = rec;
            }
            const unsigned long long volume = (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
            if (volume > kOccBigVolume) {
                // one atomic hands out the table row (high 24 bits) and the first box number (low 40 bits) together, so
                // the rows are sorted by first box: the box kernel finds its leaf by binary search
                const uint32_t boxes = boxCountOf(lo, hi);
                const unsigned long long ticket = atomicAdd(&counters->bigTicket, (1ull << 40) | boxes);
                const uint32_t row = (uint32_t) (ticket >> 40);
                if (row < occ.bigCapacity) {
                    occ.bigLeaves[row] = make_uint2(index, (uint32_t) (ticket & ((1ull << 40) - 1ull)));
                }
            }
            ++seen;
        });
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// classify: thread per candidate voxel

/// Batch entry in shared memory: the SAT constants plus where the entry's box sits.  The unit of work is a *row
/// segment*: the voxels of one row (fixed y, z) of the box that share a tile column (x >> 3) — at most 8 voxels, all in
/// one 32-bit half of one bitmap word.  60 words, 16-byte aligned.
struct alignas(16) BatchEntry {
    PairSat sat;
    uint32_t leaf;                // leaf index (queue entries name it)
    uint32_t x0, y0, z0;          // box min corner, voxel space
    uint32_t dx;                  // box extent in x
    uint32_t segs, segsDy;        // segments per row, per layer (= segs * extent in y)
    uint32_t magicSegs, magicSegsDy;  // n / d == __umulhi(n, magic) for d > 1, n * d <= 2^24 (d == 1: n itself)
    uint32_t slot;                // bitmap of the box's chunk, or kNoSlot if the box spans several chunks
    uint32_t pad[2];
};

constexpr uint32_t kNoSlot = 0xffffffffu;

struct ClassifyShared {
    BatchEntry entry[kOccBatch];
    uint32_t prefix[kOccBatch + 1];   // exclusive scan of the entries' row-segment counts
    uint32_t maybe[kOccMaybeCap];     // undecided voxels: entry << 15 | segment << 3 | x & 7; top bit: survived the filter
    uint32_t warpSums[kOccThreads / 32];
    uint32_t maybeCount;
    unsigned long long queueBase;
};

__device__ __forceinline__ uint32_t magicOf(uint32_t d)
{
    // floor(2^32 / d) + 1 for d that do not divide 2^32, 2^32 / d (exact quotients) for those that do; 32-bit division
    return d > 1u ? 0xffffffffu / d + 1u : 0u;
}

__device__ __forceinline__ uint32_t divideBy(uint32_t n, uint32_t d, uint32_t magic)
{
    return d > 1u ? __umulhi(n, magic) : n;
}

/// Bitmap word of voxel (x, y, z) for an entry: the chunk lookup is skipped when the entry's box lies in one chunk.
__device__ __forceinline__ size_t entryWord(const OccupancyView &occ, const BatchEntry &e, uint32_t x, uint32_t y,
                                            uint32_t z)
{
    if (e.slot == kNoSlot) {
        return bitmapWord(occ, x, y, z);
    }
    const uint32_t tileLocal = ((x >> 3) & 7u) | (((y >> 3) & 7u) << 3) | (((z >> 3) & 7u) << 6);
    return (size_t) e.slot * kChunkWords + tileLocal * kTileEdge + (z & 7u);
}

/// Fills one batch entry for `leaf` restricted to the box [lo, hi) (at most kOccBigVolume voxels).  Returns the number
/// of row segments (0 = skip).
__device__ __forceinline__ uint32_t stageBatchEntry(BatchEntry &e, const OccupancyView &occ, uint32_t leafIndex,
                                                    const uint32_t lo[3], const uint32_t hi[3], LeafStage &s)
{
    const bool oneChunk = (lo[0] >> 6) == ((hi[0] - 1) >> 6) && (lo[1] >> 6) == ((hi[1] - 1) >> 6) &&
                          (lo[2] >> 6) == ((hi[2] - 1) >> 6);
    e.slot = oneChunk ? __ldg(occ.chunkSlot + (lo[0] >> 6) +
                              occ.chunksPerAxis * ((lo[1] >> 6) + occ.chunksPerAxis * ((lo[2] >> 6) - occ.chunkZ0)))
                      : kNoSlot;
    const float origin[3] = {(float) lo[0], (float) lo[1], (float) lo[2]};
    buildPrefilter(s, origin);
    buildPairSat(e.sat, s, origin);
    e.leaf = leafIndex;
    e.x0 = lo[0];
    e.y0 = lo[1];
    e.z0 = lo[2];
    e.dx = hi[0] - lo[0];
    e.segs = ((hi[0] - 1u) >> 3) - (lo[0] >> 3) + 1u;
    e.segsDy = e.segs * (hi[1] - lo[1]);
    e.magicSegs = magicOf(e.segs);
    e.magicSegsDy = magicOf(e.segsDy);
    return e.segsDy * (hi[2] - lo[2]);
}

/// Leaf `index`: slots below firstLeaves hold the first leaf of triangle `index`, the others follow in extraLeaves.
__device__ __forceinline__ const LeafRecord *leafAt(const VoxelizeArgs &args, uint32_t index)
{
    return index < args.occ.firstLeaves ? args.leaves + index : args.occ.extraLeaves + (index - args.occ.firstLeaves);
}

__device__ __forceinline__ void loadLeafVertices(LeafStage &s, const VoxelizeArgs &args, uint32_t leafIndex)
{
    const float4 *src = reinterpret_cast<const float4 *>(leafAt(args, leafIndex));
    const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w;
    s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
    s.v[8] = c.x;
    s.flags = __float_as_uint(c.w);
}

/// Row (yi, zi, box-relative) and voxel range [xs, xe) (absolute x) of row segment `unit` of an entry.
__device__ __forceinline__ void segmentOf(const BatchEntry &e, uint32_t unit, uint32_t &yi, uint32_t &zi, uint32_t &xs,
                                          uint32_t &xe)
{
    zi = divideBy(unit, e.segsDy, e.magicSegsDy);
    const uint32_t inLayer = unit - zi * e.segsDy;
    yi = divideBy(inLayer, e.segs, e.magicSegs);
    const uint32_t column = (e.x0 >> 3) + (inLayer - yi * e.segs);
    xs = max(e.x0, column << 3);
    xe = min(e.x0 + e.dx, (column + 1u) << 3);
}

/// Voxel named by an entry of ClassifyShared::maybe.
__device__ __forceinline__ void maybeVoxel(const BatchEntry &e, uint32_t m, uint32_t &x, uint32_t &y, uint32_t &z)
{
    uint32_t yi, zi, xs, xe;
    segmentOf(e, (m >> 3) & 4095u, yi, zi, xs, xe);
    x = (xs & ~7u) | (m & 7u);
    y = e.y0 + yi;
    z = e.z0 + zi;
}

/// The block-wide part shared by the leaf-batch and the box kernels.  sh.entry[0 .. count) are staged and
/// sh.prefix[0 .. kOccBatch] holds the exclusive scan of their candidate counts (entries >= count contribute 0).
/// `noSat[p]` (bit p of a per-entry flag kept in entry.sat.planeLimit < 0) marks kLeafNoPrefilter leaves.
__device__ __forceinline__ void classifyBatch(ClassifyShared &sh, const VoxelizeArgs &args)
{
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t total = sh.prefix[kOccBatch];
    uint32_t *bits32 = reinterpret_cast<uint32_t *>(occ.bits);

    // ---- verdicts: lane = row segment, the flat segment space of the batch shared out warp by warp.  The row tests
    // (o2v_sat.cuh: yz edge functions, the plane at both ends of the segment) skip most segments of a thin triangle in a
    // fat box; the others walk their <= 8 voxels with one multiply-add per axis and voxel. ----
    for (uint32_t base = warp * 32u; base < total; base += kOccThreads) {  // warp-uniform trip count
        const uint32_t i = base + lane;
        uint32_t sure = 0, open = 0, p = 0, code = 0, x8 = 0, y = 0, z = 0;  // bit x & 7: certain / undecided
        if (i < total) {
#pragma unroll
            for (uint32_t step = kOccBatch / 2; step > 0; step >>= 1) {  // last entry with prefix[p] <= i
                p += sh.prefix[p + step] <= i ? step : 0u;
            }
            const BatchEntry &e = sh.entry[p];
            const uint32_t unit = i - sh.prefix[p];
            code = (p << 15) | (unit << 3);
            uint32_t yi, zi, xs, xe;
            segmentOf(e, unit, yi, zi, xs, xe);
            x8 = xs & ~7u;
            y = e.y0 + yi;
            z = e.z0 + zi;
            // planeLimit < 0 marks a leaf whose normal is too noisy for the SAT (kLeafNoPrefilter): all undecided
            if (args.prefilter && e.sat.planeLimit >= 0.0f) {
                RowSat row;
                buildRowSat(e.sat, (float) yi, (float) zi, row);
                if (!rowPlaneSpanMisses(e.sat, row, (float) (xs - e.x0), (float) (xe - 1u - e.x0))) {
                    for (uint32_t x = xs; x < xe; ++x) {
                        const int verdict = classifyInRow(e.sat, row, (float) (x - e.x0));
                        sure |= verdict == kSatCertain ? 1u << (x & 7u) : 0u;
                        open |= verdict == kSatUncertain ? 1u << (x & 7u) : 0u;
                    }
                }
            }
            else {
                open = ((1u << (xe - x8)) - 1u) & ~((1u << (xs - x8)) - 1u);
            }
        }
        if (sure != 0) {
            // one RED per segment on its 32-bit half word (4 rows of one tile layer); nothing waits for the result
            const size_t half = entryWord(occ, sh.entry[p], x8, y, z) * 2u + ((y & 7u) >> 2);
            atomicOr(bits32 + half, sure << (8u * (y & 3u)));
        }
        if (open != 0) {
            // undecided voxels are rare (a few per warp round): each lane reserves its own slots
            uint32_t slot = atomicAdd(&sh.maybeCount, (uint32_t) __popc(open));
            while (open != 0) {
                const uint32_t bit = (uint32_t) __ffs((int) open) - 1u;
                open &= open - 1u;
                if (slot < kOccMaybeCap) {
                    sh.maybe[slot] = code | bit;
                }
                else {  // buffer full (dense batch): straight to the queue, unfiltered
                    const unsigned long long index = atomicAdd(&args.counters->survivors, 1ull);
                    if (index < occ.queueCapacity) {
                        occ.queue[index] = make_uint4(sh.entry[p].leaf, (x8 | bit) | (y << 16), z, 0u);
                    }
                }
                ++slot;
            }
        }
    }
    __syncthreads();

    // ---- filter the undecided voxels by what the bitmap shows now (this block's own `certain` bits included) ----
    const uint32_t buffered = min(sh.maybeCount, kOccMaybeCap);
    uint32_t count = 0;
    for (uint32_t k = tid; k < buffered; k += kOccThreads) {
        const uint32_t m = sh.maybe[k];
        uint32_t x, y, z;
        maybeVoxel(sh.entry[m >> 15], m, x, y, z);
        const size_t word = entryWord(occ, sh.entry[m >> 15], x, y, z);
        unsigned long long known = __ldcg(occ.bits + word);
        if (downscale) {
            known = smear2x2(known | __ldcg(occ.bits + (word ^ 1u)));
        }
        if ((known & bitmapBit(x, y)) == 0) {  // a stale read only costs a redundant clip
            sh.maybe[k] = m | 0x80000000u;
            ++count;
        }
    }
    uint32_t inclusive = count;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(full, inclusive, o);
        inclusive += lane >= (uint32_t) o ? up : 0u;
    }
    if (lane == 31) {
        sh.warpSums[warp] = inclusive;
    }
    __syncthreads();
    uint32_t blockTotal = 0, warpBase = 0;
    for (uint32_t w = 0; w < kOccThreads / 32; ++w) {
        warpBase += w < warp ? sh.warpSums[w] : 0u;
        blockTotal += sh.warpSums[w];
    }
    if (blockTotal != 0) {  // block-uniform
        if (tid == 0) {
            sh.queueBase = atomicAdd(&args.counters->survivors, (unsigned long long) blockTotal);
        }
        __syncthreads();
        unsigned long long index = sh.queueBase + warpBase + (inclusive - count);
        for (uint32_t k = tid; k < buffered; k += kOccThreads) {
            const uint32_t m = sh.maybe[k];
            if ((m & 0x80000000u) != 0) {
                uint32_t x, y, z;
                const BatchEntry &e = sh.entry[(m >> 15) & 0xffffu];
                maybeVoxel(e, m, x, y, z);
                if (index < occ.queueCapacity) {  // beyond: counted only; the engine grows the queue and reruns
                    occ.queue[index] = make_uint4(e.leaf, x | (y << 16), z, 0u);
                }
                ++index;
            }
        }
    }
}

/// One block = kOccBatch consecutive leaves (big leaves are left to the box kernel).
__global__ void __launch_bounds__(kOccThreads)
occupancyClassifyKernel(const VoxelizeArgs args, uint32_t leafTotal)
{
    __shared__ ClassifyShared sh;
    const unsigned int full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t leafIndex = blockIdx.x * kOccBatch + tid;

    uint32_t volume = 0;
    if (tid < kOccBatch && leafIndex < leafTotal) {
        LeafStage s;
        loadLeafVertices(s, args, leafIndex);
        uint32_t lo[3], hi[3];
        if ((s.flags & kLeafEmpty) == 0 && leafBoxInSlab(s.v, args.grid, lo, hi)) {
            const unsigned long long v64 = (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
            if (v64 <= kOccBigVolume) {
                volume = stageBatchEntry(sh.entry[tid], args.occ, leafIndex, lo, hi, s);
                if ((s.flags & kLeafNoPrefilter) != 0) {
                    sh.entry[tid].sat.planeLimit = -1.0f;
                }
            }
        }
    }
    if (tid == 0) {
        sh.maybeCount = 0;
        sh.prefix[0] = 0;
    }
    uint32_t inclusive = volume;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(full, inclusive, o);
        inclusive += lane >= (uint32_t) o ? up : 0u;
    }
    if (lane == 31) {
        sh.warpSums[warp] = inclusive;
    }
    __syncthreads();
    uint32_t warpBase = 0;
    for (uint32_t w = 0; w < warp; ++w) {
        warpBase += sh.warpSums[w];
    }
    if (tid < kOccBatch) {
        sh.prefix[tid + 1] = warpBase + inclusive;
    }
    __syncthreads();
    if (sh.prefix[kOccBatch] == 0) {
        return;
    }
    classifyBatch(sh, args);
}

constexpr int kOccDirectThreads = 128;
constexpr uint32_t kOccDirectMaybeCap = 1024;

struct DirectShared {
    uint4 maybe[kOccDirectMaybeCap];  // undecided voxels as queue entries; .w = 1: survived the filter
    uint32_t warpSums[kOccDirectThreads / 32];
    uint32_t maybeCount;
    unsigned long long queueBase;
};

/// Thread = leaf, for meshes of micro-triangles (on average at most kOccDirectCandidates candidate voxels per leaf:
/// BASELINE config 5).  Sharing a leaf's few voxels out over a block costs more than testing them: the SAT constants
/// stay in registers and the thread walks its own rows.  Same verdict functions, same bitmap and queue as
/// occupancyClassifyKernel; big leaves are left to the box kernel.
__global__ void __launch_bounds__(kOccDirectThreads)
occupancyClassifyDirectKernel(const VoxelizeArgs args, uint32_t leafTotal)
{
    __shared__ DirectShared sh;
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool downscale = args.grid.supersampling == 2;
    uint32_t *bits32 = reinterpret_cast<uint32_t *>(occ.bits);
    if (tid == 0) {
        sh.maybeCount = 0;
    }
    __syncthreads();

    const uint32_t leafIndex = blockIdx.x * kOccDirectThreads + tid;
    if (leafIndex < leafTotal) {
        LeafStage s;
        loadLeafVertices(s, args, leafIndex);
        uint32_t lo[3], hi[3];
        if ((s.flags & kLeafEmpty) == 0 && leafBoxInSlab(s.v, args.grid, lo, hi) &&
            (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]) <= kOccBigVolume) {
            const bool oneChunk = (lo[0] >> 6) == ((hi[0] - 1) >> 6) && (lo[1] >> 6) == ((hi[1] - 1) >> 6) &&
                                  (lo[2] >> 6) == ((hi[2] - 1) >> 6);
            BatchEntry where;  // only .slot is used (entryWord)
            where.slot = oneChunk ? __ldg(occ.chunkSlot + (lo[0] >> 6) +
                                          occ.chunksPerAxis * ((lo[1] >> 6) + occ.chunksPerAxis * ((lo[2] >> 6) - occ.chunkZ0)))
                                  : kNoSlot;
            const float origin[3] = {(float) lo[0], (float) lo[1], (float) lo[2]};
            buildPrefilter(s, origin);
            PairSat sat;
            buildPairSat(sat, s, origin);
            // a leaf whose normal is too noisy for the SAT (kLeafNoPrefilter): all undecided
            const bool useSat = args.prefilter && (s.flags & kLeafNoPrefilter) == 0;
            for (uint32_t z = lo[2]; z < hi[2]; ++z) {
                for (uint32_t y = lo[1]; y < hi[1]; ++y) {
                    RowSat row;
                    buildRowSat(sat, (float) (y - lo[1]), (float) (z - lo[2]), row);
                    for (uint32_t x8 = lo[0] & ~7u; x8 < hi[0]; x8 += 8u) {
                        const uint32_t xs = max(lo[0], x8), xe = min(hi[0], x8 + 8u);
                        uint32_t sure = 0, open = 0;  // bit x & 7: certain / undecided
                        if (!useSat) {
                            open = ((1u << (xe - x8)) - 1u) & ~((1u << (xs - x8)) - 1u);
                        }
                        else if (!rowPlaneSpanMisses(sat, row, (float) (xs - lo[0]), (float) (xe - 1u - lo[0]))) {
                            for (uint32_t x = xs; x < xe; ++x) {
                                const int verdict = classifyInRow(sat, row, (float) (x - lo[0]));
                                sure |= verdict == kSatCertain ? 1u << (x & 7u) : 0u;
                                open |= verdict == kSatUncertain ? 1u << (x & 7u) : 0u;
                            }
                        }
                        if (sure != 0) {
                            const size_t half = entryWord(occ, where, x8, y, z) * 2u + ((y & 7u) >> 2);
                            atomicOr(bits32 + half, sure << (8u * (y & 3u)));
                        }
                        if (open != 0) {
                            uint32_t slot = atomicAdd(&sh.maybeCount, (uint32_t) __popc(open));
                            while (open != 0) {
                                const uint32_t bit = (uint32_t) __ffs((int) open) - 1u;
                                open &= open - 1u;
                                const uint4 entry = make_uint4(leafIndex, (x8 | bit) | (y << 16), z, 0u);
                                if (slot < kOccDirectMaybeCap) {
                                    sh.maybe[slot] = entry;
                                }
                                else {  // buffer full: straight to the queue, unfiltered
                                    const unsigned long long index = atomicAdd(&args.counters->survivors, 1ull);
                                    if (index < occ.queueCapacity) {
                                        occ.queue[index] = entry;
                                    }
                                }
                                ++slot;
                            }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- filter the undecided voxels by what the bitmap shows now, then one reservation in the queue per block ----
    const uint32_t buffered = min(sh.maybeCount, kOccDirectMaybeCap);
    if (buffered == 0) {
        return;
    }
    uint32_t count = 0;
    for (uint32_t k = tid; k < buffered; k += kOccDirectThreads) {
        const uint4 m = sh.maybe[k];
        if (!alreadyDecided(occ, downscale, m.y & 0xffffu, m.y >> 16, m.z)) {  // a stale read only costs a redundant clip
            sh.maybe[k].w = 1u;
            ++count;
        }
    }
    uint32_t inclusive = count;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(full, inclusive, o);
        inclusive += lane >= (uint32_t) o ? up : 0u;
    }
    if (lane == 31) {
        sh.warpSums[warp] = inclusive;
    }
    __syncthreads();
    uint32_t blockTotal = 0, warpBase = 0;
    for (uint32_t w = 0; w < kOccDirectThreads / 32; ++w) {
        warpBase += w < warp ? sh.warpSums[w] : 0u;
        blockTotal += sh.warpSums[w];
    }
    if (blockTotal != 0) {  // block-uniform
        if (tid == 0) {
            sh.queueBase = atomicAdd(&args.counters->survivors, (unsigned long long) blockTotal);
        }
        __syncthreads();
        unsigned long long index = sh.queueBase + warpBase + (inclusive - count);
        for (uint32_t k = tid; k < buffered; k += kOccDirectThreads) {
            const uint4 m = sh.maybe[k];
            if (m.w != 0) {
                if (index < occ.queueCapacity) {  // beyond: counted only; the engine grows the queue and reruns
                    occ.queue[index] = make_uint4(m.x, m.y, m.z, 0u);
                }
                ++index;
            }
        }
    }
}

/// Persistent blocks over the 16^3 boxes of the big leaves (axis-aligned triangles the reference does not subdivide).
__global__ void __launch_bounds__(kOccThreads)
occupancyClassifyBoxesKernel(const VoxelizeArgs args, uint32_t bigCount, unsigned long long boxTotal)
{
    __shared__ ClassifyShared sh;
    const uint32_t tid = threadIdx.x;
    for (unsigned long long box = blockIdx.x; box < boxTotal; box += gridDim.x) {
        __syncthreads();  // the previous round is done with the shared state
        if (tid == 0) {
            // last table row whose first box is <= box (rows are sorted by first box, see occupancyEmitKernel)
            uint32_t lo = 0, hi = bigCount;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (args.occ.bigLeaves[mid].y <= box) {
                    lo = mid;
                }
                else {
                    hi = mid;
                }
            }
            const uint2 row = args.occ.bigLeaves[lo];
            LeafStage s;
            loadLeafVertices(s, args, row.x);
            uint32_t leafLo[3], leafHi[3];
            leafBoxInSlab(s.v, args.grid, leafLo, leafHi);
            const uint32_t nbx = (leafHi[0] - leafLo[0] + kOccBoxEdge - 1) / kOccBoxEdge;
            const uint32_t nby = (leafHi[1] - leafLo[1] + kOccBoxEdge - 1) / kOccBoxEdge;
            const uint32_t b = (uint32_t) (box - row.y);
            const uint32_t bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
            uint32_t lo3[3] = {leafLo[0] + bx * kOccBoxEdge, leafLo[1] + by * kOccBoxEdge, leafLo[2] + bz * kOccBoxEdge};
            uint32_t hi3[3] = {min(lo3[0] + kOccBoxEdge, leafHi[0]), min(lo3[1] + kOccBoxEdge, leafHi[1]),
                               min(lo3[2] + kOccBoxEdge, leafHi[2])};
            const uint32_t volume = stageBatchEntry(sh.entry[0], args.occ, row.x, lo3, hi3, s);
            if ((s.flags & kLeafNoPrefilter) != 0) {
                sh.entry[0].sat.planeLimit = -1.0f;
            }
            sh.maybeCount = 0;
            sh.prefix[0] = 0;
            for (uint32_t k = 1; k <= kOccBatch; ++k) {
                sh.prefix[k] = volume;
            }
        }
        __syncthreads();
        classifyBatch(sh, args);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// clip: the exact six-plane clip for the queued voxels

/// Same persistent-lane scheme as sparseClipKernel (o2v_sparse.cu), fed from the queue; a surviving piece sets the bit.
__global__ void __launch_bounds__(kOccClipThreads)
occupancyClipKernel(const VoxelizeArgs args)
{
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const bool downscale = args.grid.supersampling == 2;
    unsigned long long total = args.counters->survivors;
    total = total < occ.queueCapacity ? total : occ.queueCapacity;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t below = (1u << lane) - 1u;
    const unsigned long long warpsTotal = (unsigned long long) gridDim.x * (blockDim.x >> 5);
    const unsigned long long warpIndex = (unsigned long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned long long chunk = ((total + warpsTotal - 1) / warpsTotal + 31ull) & ~31ull;
    unsigned long long cursor = warpIndex * chunk;
    const unsigned long long end = cursor + chunk < total ? cursor + chunk : total;

    __shared__ uint8_t caseTable[64];
    fillClipCaseTable(caseTable);
    __syncthreads();
    WarpClipper<false> clipper;
    ClipStack<false> stack;
    clipper.idle();
    clipper.r.pieces = 0;
    clipper.r.weight = clipper.r.u = clipper.r.v = 0.0f;
    bool hasEntry = false;
    unsigned long long *word = nullptr;
    unsigned long long bit = 0;

    for (;;) {
        const unsigned int idle = __ballot_sync(full, clipper.done);
        const bool moreWork = cursor < end;
        if (idle == full && !moreWork) {
            break;
        }
        if (moreWork && (__popc(idle) >= kOccRefillThreshold || idle == full)) {
            if (clipper.done) {
                if (hasEntry) {
                    if (clipper.r.pieces != 0) {
                        atomicOr(word, bit);
                    }
                    hasEntry = false;
                }
                const unsigned long long e = cursor + __popc(idle & below);
                if (e < end) {
                    const uint4 entry = occ.queue[e];
                    const uint32_t x = entry.y & 0xffffu, y = entry.y >> 16, z = entry.z;
                    if (!alreadyDecided(occ, downscale, x, y, z)) {
                        const float4 *src = reinterpret_cast<const float4 *>(leafAt(args, entry.x));
                        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
                        Tri<false> leaf;
                        leaf.v[0] = a.x; leaf.v[1] = a.y; leaf.v[2] = a.z; leaf.v[3] = a.w;
                        leaf.v[4] = b.x; leaf.v[5] = b.y; leaf.v[6] = b.z; leaf.v[7] = b.w;
                        leaf.v[8] = c.x;
                        clipper.begin(leaf, x, y, z, c.z);
                        if ((__float_as_uint(c.w) & kLeafNeedsCull) != 0 && planeDistanceCulled(leaf.v, x, y, z)) {
                            clipper.done = true;  // voxelization.cpp:451-458 (slivers only)
                        }
                        hasEntry = true;
                        word = occ.bits + bitmapWord(occ, x, y, z);
                        bit = bitmapBit(x, y);
                    }
                }
            }
            cursor += __popc(idle);
        }
        clipper.round(stack, caseTable);
    }
    if (hasEntry && clipper.r.pieces != 0) {
        atomicOr(word, bit);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// expand: thread per tile of every active chunk

/// Position of the r-th (0-based) set bit of w; r < popcount(w).
__device__ __forceinline__ uint32_t selectBit64(unsigned long long w, uint32_t r)
{
    uint32_t word = (uint32_t) w, pos = 0;
    const uint32_t inLow = (uint32_t) __popc(word);
    if (r >= inLow) {
        r -= inLow;
        word = (uint32_t) (w >> 32);
        pos = 32;
    }
#pragma unroll
    for (uint32_t step = 16; step > 0; step >>= 1) {
        const uint32_t c = (uint32_t) __popc(word & ((1u << step) - 1u));
        const bool up = r >= c;
        r -= up ? c : 0u;
        word = up ? word >> step : word;
        pos += up ? step : 0u;
    }
    return pos;
}

/// What the record-writing lanes need to know about one of the warp's 32 tiles.
struct alignas(16) ExpandTile {
    uint32_t first;         // records of the warp's earlier tiles
    uint16_t origin[3];     // output-space min corner
    uint16_t wordFirst[9];  // records of the tile's earlier layers; [8] = the tile's total
};

struct ExpandShared {
    unsigned long long mask[kOccExpandThreads / 32][kTileEdge][32];  // [layer][tile]: conflict-free to fill
    ExpandTile tile[kOccExpandThreads / 32][32];
};

/// Two steps per warp and 32 tiles: (1) lane = tile: load the 8 layer words, fold 2x2x2 when downscaling, count;
/// (2) lane = output record: records are numbered across the warp's tiles, record j finds its tile, layer and bit by
/// rank/select in shared memory — all lanes busy whatever the fill of the tiles, and a warp stores 512 contiguous bytes.
__global__ void __launch_bounds__(kOccExpandThreads)
occupancyExpandKernel(const VoxelizeArgs args)
{
    __shared__ ExpandShared sh;
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t shift = downscale ? 1u : 0u;
    const unsigned long long tiles = (unsigned long long) occ.activeChunks * (kChunkWords / kTileEdge);
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    unsigned long long overflow = 0;

    for (unsigned long long base = (unsigned long long) blockIdx.x * blockDim.x + (threadIdx.x - lane); base < tiles;
         base += stride) {
        const unsigned long long t = base + lane;
        unsigned long long m[kTileEdge];
#pragma unroll
        for (int z = 0; z < (int) kTileEdge; ++z) {
            m[z] = 0;
        }
        uint32_t origin[3] = {0, 0, 0};
        if (t < tiles) {
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(occ.bits + t * kTileEdge);
#pragma unroll
            for (int k = 0; k < (int) kTileEdge / 2; ++k) {
                const ulonglong2 w = __ldcs(src + k);  // read once
                m[2 * k] = w.x;
                m[2 * k + 1] = w.y;
            }
        }
        if (downscale) {
            // parent (qx, qy, qz) = OR of its 8 children; kept at bit (2 qx + 16 qy) of word qz
#pragma unroll
            for (int q = 0; q < (int) kTileEdge / 2; ++q) {
                unsigned long long w = m[2 * q] | m[2 * q + 1];
                w = (w | (w >> 1)) & 0x5555555555555555ull;
                w = (w | (w >> 8)) & 0x00ff00ff00ff00ffull;
                m[q] = w;
            }
#pragma unroll
            for (int q = (int) kTileEdge / 2; q < (int) kTileEdge; ++q) {
                m[q] = 0;
            }
        }
        uint32_t count = 0;
        uint16_t wordFirst[kTileEdge + 1];
#pragma unroll
        for (int z = 0; z < (int) kTileEdge; ++z) {
            wordFirst[z] = (uint16_t) count;
            count += (uint32_t) __popcll(m[z]);
        }
        wordFirst[kTileEdge] = (uint16_t) count;
        uint32_t inclusive = count;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(full, inclusive, o);
            inclusive += lane >= (uint32_t) o ? up : 0u;
        }
        const uint32_t warpCount = __shfl_sync(full, inclusive, 31);
        if (warpCount == 0) {
            continue;
        }
        if (count != 0) {
            const uint32_t chunk = occ.chunkList[t >> 9], tileLocal = (uint32_t) t & 511u;
            const uint32_t C = occ.chunksPerAxis;
            origin[0] = ((chunk % C) * kChunkEdge + (tileLocal & 7u) * kTileEdge) >> shift;
            origin[1] = (((chunk / C) % C) * kChunkEdge + ((tileLocal >> 3) & 7u) * kTileEdge) >> shift;
            origin[2] = ((chunk / (C * C) + occ.chunkZ0) * kChunkEdge + (tileLocal >> 6) * kTileEdge) >> shift;
        }
        __syncwarp();  // the previous round's readers are done
        ExpandTile &mine = sh.tile[warp][lane];
        mine.first = inclusive - count;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mine.origin[a] = (uint16_t) origin[a];
        }
#pragma unroll
        for (int z = 0; z <= (int) kTileEdge; ++z) {
            mine.wordFirst[z] = wordFirst[z];
        }
#pragma unroll
        for (int z = 0; z < (int) kTileEdge; ++z) {
            sh.mask[warp][z][lane] = m[z];
        }
        unsigned long long index = 0;
        if (lane == 31) {
            index = atomicAdd(&args.counters->voxels, (unsigned long long) warpCount);
        }
        index = __shfl_sync(full, index, 31);
        __syncwarp();

        for (uint32_t j = lane; j < warpCount; j += 32u) {
            uint32_t tile = 0;
#pragma unroll
            for (uint32_t step = 16; step > 0; step >>= 1) {  // last tile with first <= j (empty tiles share a first)
                tile += sh.tile[warp][tile + step].first <= j ? step : 0u;
            }
            const ExpandTile &info = sh.tile[warp][tile];
            const uint32_t r = j - info.first;
            uint32_t z = 0;
#pragma unroll
            for (uint32_t step = kTileEdge / 2; step > 0; step >>= 1) {  // last layer with wordFirst <= r
                z += info.wordFirst[z + step] <= r ? step : 0u;
            }
            const uint32_t b = selectBit64(sh.mask[warp][z][tile], r - info.wordFirst[z]);
            if (index + j < args.outCapacity) {
                VoxelRecord rec;
                rec.x = (int32_t) (info.origin[0] + ((b & 7u) >> shift));
                rec.y = (int32_t) (info.origin[1] + ((b >> 3) >> shift));
                rec.z = (int32_t) (info.origin[2] + z);
                rec.argb = 0xFFFFFFFFu;  // quantizeArgb(1, 1, 1)
                __stcs(reinterpret_cast<int4 *>(args.out + index + j), *reinterpret_cast<const int4 *>(&rec));
            }
            else {
                ++overflow;
            }
        }
    }
    if (overflow != 0) {
        atomicAdd(&args.counters->outputOverflow, overflow);
    }
}

template <typename Kernel>
unsigned occupancyPersistentBlocks(Kernel kernel, int threads, int smCount)
{
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, 0);
    perSm = perSm < 1 ? 1 : perSm;
    return (unsigned) smCount * (unsigned) perSm;  // a multiple of the SM count
}

/// Persistent grid of the triangle passes: as many blocks as are resident at once (a multiple of the SM count; no tail
/// wave, since a block's batches are dealt round-robin), fewer if there are fewer batches.
template <typename Kernel>
unsigned setupBlocks(Kernel kernel, unsigned long long n, int smCount)
{
    unsigned long long batches = (n + kOccSetupThreads - 1) / kOccSetupThreads;
    batches = batches < 1 ? 1 : batches;
    const unsigned long long resident = occupancyPersistentBlocks(kernel, kOccSetupThreads, smCount);
    return (unsigned) (batches < resident ? batches : resident);
}

}  // namespace

void launchOccupancySlabFilter(const MeshView &mesh, const GridView &grid, float *kept, RunCounters *counters,
                               int smCount, cudaStream_t stream)
{
    occupancySlabFilterKernel<<<setupBlocks(occupancySlabFilterKernel, mesh.count, smCount), kOccSetupThreads, 0,
                                stream>>>(mesh, grid, kept, counters);
}

void launchOccupancyCount(const MeshView &mesh, const GridView &grid, const OccupancyView &occ, uint32_t *extraCount,
                          LeafRecord *firstLeaves, RunCounters *counters, bool countFromFilter, int smCount,
                          cudaStream_t stream)
{
    // countFromFilter: mesh.count is only an upper bound (the grid is sized by it; blocks without a batch leave at once)
    occupancyCountKernel<<<setupBlocks(occupancyCountKernel, mesh.count, smCount), kOccSetupThreads, 0, stream>>>(
        mesh, grid, occ, extraCount, firstLeaves, counters, countFromFilter);
}

void launchOccupancyAssignChunks(const OccupancyView &occ, RunCounters *counters, cudaStream_t stream)
{
    occupancyAssignChunksKernel<<<(occ.chunkTotal + 255) / 256, 256, 0, stream>>>(occ, counters);
}

void launchOccupancyEmit(const MeshView &mesh, const GridView &grid, const OccupancyView &occ,
                         const uint32_t *leafOffset, LeafRecord *leaves, RunCounters *counters, int smCount,
                         cudaStream_t stream)
{
    occupancyEmitKernel<<<setupBlocks(occupancyEmitKernel, mesh.count, smCount), kOccSetupThreads, 0, stream>>>(
        mesh, grid, occ, leafOffset, leaves, counters);
}

void launchOccupancyClassify(const VoxelizeArgs &args, unsigned long long leafTotal, bool microLeaves, uint32_t bigCount,
                             unsigned long long boxTotal, int smCount, cudaStream_t stream)
{
    if (leafTotal != 0 && microLeaves) {
        const unsigned blocks = (unsigned) ((leafTotal + kOccDirectThreads - 1) / kOccDirectThreads);
        occupancyClassifyDirectKernel<<<blocks, kOccDirectThreads, 0, stream>>>(args, (uint32_t) leafTotal);
    }
    else if (leafTotal != 0) {
        const unsigned blocks = (unsigned) ((leafTotal + kOccBatch - 1) / kOccBatch);
        occupancyClassifyKernel<<<blocks, kOccThreads, 0, stream>>>(args, (uint32_t) leafTotal);
    }
    if (bigCount != 0 && boxTotal != 0) {
        unsigned long long blocks = occupancyPersistentBlocks(occupancyClassifyBoxesKernel, kOccThreads, smCount);
        blocks = blocks < boxTotal ? blocks : boxTotal;
        occupancyClassifyBoxesKernel<<<(unsigned) blocks, kOccThreads, 0, stream>>>(args, bigCount, boxTotal);
    }
}

void launchOccupancyClip(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    // the queue length lives on the device (RunCounters::survivors)
    occupancyClipKernel<<<occupancyPersistentBlocks(occupancyClipKernel, kOccClipThreads, smCount), kOccClipThreads, 0,
                          stream>>>(args);
}

void launchOccupancyExpand(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    if (args.occ.activeChunks == 0) {
        return;
    }
    unsigned long long blocks = occupancyPersistentBlocks(occupancyExpandKernel, kOccExpandThreads, smCount);
    const unsigned long long needed =
        ((unsigned long long) args.occ.activeChunks * (kChunkWords / kTileEdge) + kOccExpandThreads - 1) /
        kOccExpandThreads;
    blocks = blocks < needed ? blocks : needed;
    occupancyExpandKernel<<<(unsigned) blocks, kOccExpandThreads, 0, stream>>>(args);
}

}  // namespace o2v
