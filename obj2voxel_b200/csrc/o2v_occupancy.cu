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
#include <algorithm>

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
#ifndef O2V_OCC_MIN_BLOCKS
#define O2V_OCC_MIN_BLOCKS 8
#endif
constexpr int kOccBatch = O2V_OCC_BATCH;      // leaves per classify block
constexpr int kOccThreads = O2V_OCC_THREADS;  // threads per classify block
constexpr int kOccClipThreads = 128;
constexpr int kOccExpandThreads = 128;
constexpr int kOccRefillThreshold = 8;
constexpr uint32_t kOccRangeCap = 128;     // undecided x ranges a block buffers before appending them to the range list
constexpr uint32_t kOccOwnerCap = 2048;    // rows of a batch whose owner (batch entry) is looked up in a table
constexpr int kOccFilterThreads = 256;
constexpr uint32_t kFilterEntries = 4;     // range entries a filter thread handles per round (their probes overlap)
constexpr uint32_t kFilterProbe = 2;       // voxels per run probed up front (ordinary runs are 1 - 2 long)

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

/// The bitmaps live in OUTPUT space: with 2x supersampling a sample voxel (x, y, z) sets the bit of its parent
/// (x >> 1, y >> 1, z >> 1) — the downscale of an all-white model is the OR of the children, so the
/// children never need bits of their own.  One eighth of the memory to clear, classify into and expand, and a bitmap
/// that stays in L2 at 1024^3.
/// Index of the 64-bit bitmap word (layer oz of the output voxel's tile) and the voxel's bit in it.
__device__ __forceinline__ size_t bitmapWord(const OccupancyView &occ, uint32_t ox, uint32_t oy, uint32_t oz)
{
    const uint32_t chunk = (ox >> 6) + occ.chunksPerAxis * ((oy >> 6) + occ.chunksPerAxis * ((oz >> 6) - occ.chunkZ0));
    const uint32_t tileLocal = ((ox >> 3) & 7u) | (((oy >> 3) & 7u) << 3) | (((oz >> 3) & 7u) << 6);
    return (size_t) __ldg(occ.chunkSlot + chunk) * kChunkWords + tileLocal * kTileEdge + (oz & 7u);
}

__device__ __forceinline__ unsigned long long bitmapBit(uint32_t ox, uint32_t oy)
{
    return 1ull << ((ox & 7u) + 8u * (oy & 7u));
}

/// true if the bitmap already decides sample voxel (x, y, z): the bit of its output voxel is set (by this voxel or, when
/// downscaling, by any of its siblings).
__device__ __forceinline__ bool alreadyDecided(const OccupancyView &occ, uint32_t x, uint32_t y, uint32_t z)
{
    const uint32_t ox = x >> occ.shift, oy = y >> occ.shift, oz = z >> occ.shift;
    return (__ldcg(occ.bits + bitmapWord(occ, ox, oy, oz)) & bitmapBit(ox, oy)) != 0;
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

/// Multi-device ingest (SlabScatter): every triangle of this device's share goes to the device(s) whose slab its z
/// range can reach.  Per batch of kOccSetupThreads triangles the block counts its triangles per slab (ballots), reserves
/// their places with one LOCAL atomic per slab, and the threads store their 36 bytes to the peer's memory — posted
/// writes over NVLink, nothing comes back.  A triangle with a negative z goes to slab 0, which drops and counts it.
__global__ void __launch_bounds__(kOccSetupThreads)
occupancySlabScatterKernel(MeshView mesh, GridView grid, SlabScatter scatter)
{
    __shared__ TriangleBatch<kOccSetupThreads> batch;
    __shared__ uint32_t warpCount[kOccSetupThreads / 32][kMaxSlabs];
    __shared__ unsigned long long slabBase[kMaxSlabs];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t gridEnd = scatter.bound[scatter.slabs];

    streamTriangles<kOccSetupThreads>(mesh.verts, mesh.count, batch,
                                      [&](unsigned long long, const float in[9], bool valid) {
        uint32_t first = 1, last = 0;  // slabs [first, last] receive the triangle
        if (valid) {
            float zlo, zhi;
            triangleZRange(grid, in, zlo, zhi);
            if (zlo < 0.0f) {
                first = last = 0;
            }
            else if (toU32(zlo) < gridEnd) {
                const uint32_t z0 = toU32(zlo), z1 = min(toU32(zhi), gridEnd - 1u);
                first = last = 0;
                for (uint32_t s = 1; s < scatter.slabs; ++s) {
                    first += scatter.bound[s] <= z0 ? 1u : 0u;
                    last += scatter.bound[s] <= z1 ? 1u : 0u;
                }
            }
        }
        for (uint32_t s = 0; s < scatter.slabs; ++s) {
            const unsigned int votes = __ballot_sync(0xffffffffu, first <= s && s <= last);
            if (lane == 0) {
                warpCount[warp][s] = (uint32_t) __popc(votes);
            }
        }
        __syncthreads();
        if (tid < scatter.slabs) {
            uint32_t total = 0;
#pragma unroll
            for (uint32_t w = 0; w < kOccSetupThreads / 32; ++w) {
                total += warpCount[w][tid];
            }
            slabBase[tid] = total != 0 ? atomicAdd(scatter.count + tid, (unsigned long long) total) : 0ull;
        }
        __syncthreads();
        for (uint32_t s = 0; s < scatter.slabs; ++s) {  // block-uniform loop; a triangle sends to one slab as a rule
            const bool mine = first <= s && s <= last;
            const unsigned int votes = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                unsigned long long slot = slabBase[s] + __popc(votes & below);
                for (uint32_t w = 0; w < warp; ++w) {
                    slot += warpCount[w][s];
                }
                if (slot < scatter.capacity) {  // (a region holds the source's whole share: always true)
                    float *dst = scatter.dest[s] + slot * 9;
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        dst[k] = in[k];
                    }
                }
            }
        }
        __syncthreads();  // warpCount / slabBase are rewritten by the next batch
    });
}

/// Triangles per row of `unit` sample-space z layers (rows [0, rows) of the grid): what a job over several devices
/// balances its slabs by.  A triangle counts in every row its z range reaches.
__global__ void __launch_bounds__(kOccSetupThreads)
occupancyZHistogramKernel(MeshView mesh, GridView grid, uint32_t unit, uint32_t rows, unsigned long long *histogram)
{
    __shared__ TriangleBatch<kOccSetupThreads> batch;
    __shared__ uint32_t local[128];
    for (uint32_t r = threadIdx.x; r < 128u; r += kOccSetupThreads) {
        local[r] = 0;
    }
    // (streamTriangles starts with a barrier)
    streamTriangles<kOccSetupThreads>(mesh.verts, mesh.count, batch,
                                      [&](unsigned long long, const float in[9], bool valid) {
        if (!valid) {
            return;
        }
        float zlo, zhi;
        triangleZRange(grid, in, zlo, zhi);
        if (zlo < 0.0f || toU32(zlo) >= rows * unit) {
            return;  // dropped as a whole, or beyond the grid
        }
        const uint32_t r0 = toU32(zlo) / unit, r1 = min(toU32(zhi) / unit, rows - 1u);
        for (uint32_t r = r0; r <= r1; ++r) {
            atomicAdd(&local[r], 1u);
        }
    });
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < rows; r += kOccSetupThreads) {
        if (local[r] != 0) {
            atomicAdd(histogram + r, (unsigned long long) local[r]);
        }
    }
}

/// Block = bitmap: the occupied voxels of each chunk, for a host that expands downloaded bitmaps itself.
__global__ void __launch_bounds__(256)
occupancyChunkCountKernel(OccupancyView occ, uint32_t *__restrict__ chunkCounts, RunCounters *counters)
{
    __shared__ uint32_t warpSum[8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t slot = blockIdx.x; slot < occ.activeChunks; slot += gridDim.x) {
        const ulonglong2 *words = reinterpret_cast<const ulonglong2 *>(occ.bits + (size_t) slot * kChunkWords);
        uint32_t sum = 0;
        for (uint32_t k = threadIdx.x; k < kChunkWords / 2; k += 256u) {
            const ulonglong2 w = words[k];
            sum += (uint32_t) (__popcll(w.x) + __popcll(w.y));
        }
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
        }
        if (lane == 0) {
            warpSum[warp] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
            for (int w = 0; w < 8; ++w) {
                total += warpSum[w];
            }
            chunkCounts[slot] = total;
            if (total != 0) {
                atomicAdd(&counters->voxels, (unsigned long long) total);
            }
        }
        __syncthreads();
    }
}

constexpr uint32_t kOccBlockChunkWords = 2048;  // chunk bits a block collects in shared memory: 65536 chunks (8 KB)

struct OccCountTally {
    unsigned long long candidates, bigLeaves, bigBoxes;
};

/// What the count pass does with one leaf: statistics and the marks of the chunks its box reaches.  `chunkBits` = the
/// block's marks in shared memory (when the slab's chunk flags fit there), else the marks go to occ.chunkFlag.
__device__ __forceinline__ void occCountLeaf(const OccupancyView &occ, uint32_t *chunkBits, const uint32_t *lo,
                                             const uint32_t *hi, OccCountTally &tally)
{
    const unsigned long long volume = (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    tally.candidates += volume;
    if (volume > kOccBigVolume) {
        ++tally.bigLeaves;
        tally.bigBoxes += boxCountOf(lo, hi);
    }
    const uint32_t cs = 6u + occ.shift;  // a chunk is 64^3 OUTPUT voxels
    auto mark = [&](uint32_t cx, uint32_t cy, uint32_t cz) {
        const uint32_t chunk = cx + occ.chunksPerAxis * (cy + occ.chunksPerAxis * (cz - occ.chunkZ0));
        const uint32_t bit = 1u << (chunk & 31u);
        if (chunkBits != nullptr) {
            if ((chunkBits[chunk >> 5] & bit) == 0) {
                atomicOr(&chunkBits[chunk >> 5], bit);
            }
        }
        else if ((__ldcg(occ.chunkFlag + (chunk >> 5)) & bit) == 0) {  // look before the atomic
            atomicOr(occ.chunkFlag + (chunk >> 5), bit);
        }
    };
    const uint32_t cx0 = lo[0] >> cs, cy0 = lo[1] >> cs, cz0 = lo[2] >> cs;
    const uint32_t cx1 = (hi[0] - 1) >> cs, cy1 = (hi[1] - 1) >> cs, cz1 = (hi[2] - 1) >> cs;
    mark(cx0, cy0, cz0);
    if (cx0 != cx1 || cy0 != cy1 || cz0 != cz1) {  // rare: the box straddles a chunk boundary
        for (uint32_t cz = cz0; cz <= cz1; ++cz) {
            for (uint32_t cy = cy0; cy <= cy1; ++cy) {
                for (uint32_t cx = cx0; cx <= cx1; ++cx) {
                    mark(cx, cy, cz);
                }
            }
        }
    }
}

__device__ __forceinline__ void storeLeafRecord(LeafRecord *slot, const float *v, uint32_t tri, float area)
{
    LeafRecord rec;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        rec.v[k] = v[k];
    }
    rec.tri = tri;
    rec.area = area;
    rec.flags = leafFlagsOf(v);
    *slot = rec;
}

/// The one pass every triangle takes: transform, subdivision DFS, statistics, chunk marks — and the triangle's first leaf
/// goes straight into leaf slot i, so that a mesh whose triangles are all leaves themselves (anything fine relative to
/// the grid) needs no second pass.  extraCount[i] = the triangle's leaves beyond the first.
/// A huge triangle is only listed here (`work`, o2v_device.cuh; its slot stays empty): occHugeCountKernel does for it
/// what this kernel does for the others, spread over the device.
__global__ void __launch_bounds__(kOccSetupThreads, 9)  // 56 registers, as before huge triangles were listed here
occupancyCountKernel(MeshView mesh, GridView grid, OccupancyView occ, uint32_t *__restrict__ extraCount,
                     LeafRecord *__restrict__ firstLeaves, RunCounters *counters, bool countFromFilter, HugeWork work)
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
    uint32_t *const marks = collect ? chunkBits : nullptr;
    // (streamTriangles starts with a barrier)
    OccCountTally tally{0ull, 0ull, 0ull};
    unsigned long long dropped = 0, overflow = 0, leafTally = 0, extraTally = 0;
    streamTriangles<kOccSetupThreads>(mesh.verts, mesh.count, batch,
                                      [&](unsigned long long i, const float in[9], bool valid) {
        if (!valid) {
            return;
        }
        Tri<false> root;
        float area;
        uint32_t leaves = 0;
        if (setupTriangle<false>(grid, in, root, area)) {
            bool huge = false;
            const bool ok = traverseLeaves<false>(root, grid, [&](const Tri<false> &leaf, const uint32_t *lo,
                                                                   const uint32_t *hi) {
                if (leaves == 0) {
                    // tri = position in the array this pass reads (unused on this path)
                    storeLeafRecord(firstLeaves + i, leaf.v, static_cast<uint32_t>(i), area);
                }
                ++leaves;
                occCountLeaf(occ, marks, lo, hi, tally);
            }, &huge);
            overflow += ok ? 0 : 1;
            if (huge) {
                listHugeTriangle(work, counters, i);
            }
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
            const uint32_t bits = chunkBits[w];
            if (bits != 0 && (bits & ~__ldcg(occ.chunkFlag + w)) != 0) {
                atomicOr(occ.chunkFlag + w, bits);
            }
        }
    }
    warpTally(&counters->leaves, leafTally);
    warpTally(&counters->extraLeaves, extraTally);
    warpTally(&counters->candidateVoxels, tally.candidates);
    warpTally(&counters->droppedTriangles, dropped);
    warpTally(&counters->depthOverflow, overflow);
    warpTally(&counters->bigLeaves, tally.bigLeaves);
    warpTally(&counters->bigBoxes, tally.bigBoxes);
}

constexpr int kOccHugeThreads = 32;  // a thread walks whole subtrees, latency-bound: many small blocks over all SMs

/// The count pass for the listed huge triangles, after hugeSubtreeCountKernel / hugeScanKernel (o2v_kernels.cu) placed
/// the subtrees: leaf 0 of a triangle goes into its slot, every leaf marks its chunks and enters the statistics.
__global__ void __launch_bounds__(kOccHugeThreads)
occHugeCountKernel(MeshView mesh, GridView grid, OccupancyView occ, HugeWork work, LeafRecord *__restrict__ firstLeaves,
                   RunCounters *counters)
{
    OccCountTally tally{0ull, 0ull, 0ull};
    const unsigned long long overflow = forEachHugeLeaf<false>(
        mesh, grid, work, counters, true,
        [&](uint32_t tri, float area, const Tri<false> &leaf, const uint32_t *lo, const uint32_t *hi, uint32_t seq) {
            if (seq == 0) {
                storeLeafRecord(firstLeaves + tri, leaf.v, tri, area);
            }
            occCountLeaf(occ, nullptr, lo, hi, tally);
        });
    warpTally(&counters->candidateVoxels, tally.candidates);
    warpTally(&counters->bigLeaves, tally.bigLeaves);
    warpTally(&counters->bigBoxes, tally.bigBoxes);
    warpTally(&counters->depthOverflow, overflow);
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

__global__ void occupancyAllChunksKernel(OccupancyView occ)
{
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk < occ.chunkTotal) {
        occ.chunkSlot[chunk] = chunk;
        occ.chunkList[chunk] = chunk;
    }
}

/// What the second pass does with leaf number `seq` of triangle `tri`: leaves beyond the first are written to
/// extraLeaves, and a leaf with more than kOccBigVolume candidates — first leaves included — enters the big-leaf table.
__device__ __forceinline__ void occEmitLeaf(const OccupancyView &occ, const uint32_t *extraOffset, LeafRecord *extraLeaves,
                                            RunCounters *counters, uint32_t tri, float area, uint32_t seq,
                                            const Tri<false> &leaf, const uint32_t *lo, const uint32_t *hi)
{
    uint32_t index = tri;
    if (seq != 0) {
        const uint32_t extraAt = extraOffset[tri] + (seq - 1u);
        index = occ.firstLeaves + extraAt;
        storeLeafRecord(extraLeaves + extraAt, leaf.v, tri, area);
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
        bool huge = false;  // listed: occHugeEmitKernel
        traverseLeaves<false>(root, grid, [&](const Tri<false> &leaf, const uint32_t *lo, const uint32_t *hi) {
            occEmitLeaf(occ, extraOffset, extraLeaves, counters, static_cast<uint32_t>(i), area, seen, leaf, lo, hi);
            ++seen;
        }, &huge);
    });
}

__global__ void __launch_bounds__(kOccHugeThreads)
occHugeEmitKernel(MeshView mesh, GridView grid, OccupancyView occ, HugeWork work, const uint32_t *__restrict__ extraOffset,
                  LeafRecord *__restrict__ extraLeaves, RunCounters *counters)
{
    forEachHugeLeaf<false>(mesh, grid, work, counters, true,
                           [&](uint32_t tri, float area, const Tri<false> &leaf, const uint32_t *lo, const uint32_t *hi,
                               uint32_t seq) {
                               occEmitLeaf(occ, extraOffset, extraLeaves, counters, tri, area, seq, leaf, lo, hi);
                           });
}

// ---------------------------------------------------------------------------------------------------------------------
// classify: lane = row of a leaf's box

/// Batch entry in shared memory: the row-interval SAT constants (o2v_sat.cuh, SpanSat) plus where the entry's box sits.
/// The unit of work is a *row* of the box (fixed y, z): classifySpan solves the thirteen axes for the x interval that is
/// not a `miss` and the sub-interval that is `certain`, so a row costs the same whatever its length and every lane of a
/// warp does the same work.  84 words: lanes reading the same field of different entries with 128-bit loads hit distinct
/// banks (84 mod 32 = 20).
struct alignas(16) BatchEntry {
    SpanSat sat;
    uint32_t leaf;        // leaf index (queue entries name it)
    uint32_t x0, y0, z0;  // box min corner, sample space
    uint32_t dx, dy;      // box extent in x and y (rows = dy * extent in z)
    uint32_t magicDy;     // n / dy == __umulhi(n, magic) for dy > 1, n * dy <= 2^24
    uint32_t slot;        // bitmap of the box's chunk, or kNoSlot if the box spans several chunks
    uint32_t noSat;       // kLeafNoPrefilter: the leaf's normal is too noisy for the SAT, every candidate is undecided
    uint32_t pad[3];
};
static_assert(sizeof(BatchEntry) == 84 * 4, "bank-conflict-free stride");
static_assert(kOccBatch <= 128, "ClassifyShared::maybe keeps the entry in 7 bits");

constexpr uint32_t kNoSlot = 0xffffffffu;

/// Undecided voxels leave the classifiers as the two x runs at the ends of a row's interval, one entry per row:
/// {leaf, x | y << 16, z | gap << 16, lengthA | lengthB << 16} = voxels x .. x + lengthA - 1 and, `gap` voxels further,
/// lengthB more.  A block collects its entries in shared memory and appends them to the global list with one atomic.
/// Whether the bitmap already decides a voxel is asked later, when every `certain` bit of the run is in place and the
/// probes of millions of entries can be in flight together (occupancyFilterQueueKernel): inside the classifier the same
/// probes cost a quarter of its run time in exposed latency (profiles/r02_history.md).
struct RangeBuffer {
    uint4 slot[kOccRangeCap];
    uint32_t count;
    unsigned long long base;
};

__device__ __forceinline__ void pushRange(RangeBuffer &rb, const VoxelizeArgs &args, uint32_t leaf, uint32_t x, uint32_t y,
                                          uint32_t z, uint32_t lengthA, uint32_t gap, uint32_t lengthB)
{
    const uint4 entry = make_uint4(leaf, x | (y << 16), z | (gap << 16), lengthA | (lengthB << 16));
    const uint32_t slot = atomicAdd(&rb.count, 1u);
    if (slot < kOccRangeCap) {
        rb.slot[slot] = entry;
    }
    else {  // buffer full (a dense batch): straight to the list
        const unsigned long long index = atomicAdd(&args.counters->ranges, 1ull);
        if (index < args.occ.rangeCapacity) {  // beyond: counted only; the engine grows the list and reruns
            args.occ.ranges[index] = entry;
        }
    }
}

/// All threads of the block, after a barrier that follows the last pushRange.
__device__ __forceinline__ void flushRanges(RangeBuffer &rb, const VoxelizeArgs &args)
{
    const uint32_t buffered = min(rb.count, kOccRangeCap);
    if (buffered == 0) {  // block-uniform
        return;
    }
    if (threadIdx.x == 0) {
        rb.base = atomicAdd(&args.counters->ranges, (unsigned long long) buffered);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < buffered; k += blockDim.x) {
        const unsigned long long index = rb.base + k;
        if (index < args.occ.rangeCapacity) {
            args.occ.ranges[index] = rb.slot[k];
        }
    }
}

/// The same buffer holding single voxels {leaf, x | y << 16, z, 0} that go to the clip queue as they are: on a mesh of
/// micro-triangles hardly any undecided voxel is decided by another leaf (BASELINE config 5: 17.2 of 17.7 million
/// survive), so the check against the bitmap is left to the clip kernel's own.
__device__ __forceinline__ void pushVoxel(RangeBuffer &rb, const VoxelizeArgs &args, uint32_t leaf, uint32_t x, uint32_t y,
                                          uint32_t z)
{
    const uint4 entry = make_uint4(leaf, x | (y << 16), z, 0u);
    const uint32_t slot = atomicAdd(&rb.count, 1u);
    if (slot < kOccRangeCap) {
        rb.slot[slot] = entry;
    }
    else {
        const unsigned long long index = atomicAdd(&args.counters->survivors, 1ull);
        if (index < args.occ.queueCapacity) {  // beyond: counted only; the engine grows the queue and reruns
            args.occ.queue[index] = entry;
        }
    }
}

__device__ __forceinline__ void flushVoxels(RangeBuffer &rb, const VoxelizeArgs &args)
{
    const uint32_t buffered = min(rb.count, kOccRangeCap);
    if (buffered == 0) {  // block-uniform
        return;
    }
    if (threadIdx.x == 0) {
        rb.base = atomicAdd(&args.counters->survivors, (unsigned long long) buffered);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < buffered; k += blockDim.x) {
        const unsigned long long index = rb.base + k;
        if (index < args.occ.queueCapacity) {
            args.occ.queue[index] = rb.slot[k];
        }
    }
}

struct ClassifyShared {
    BatchEntry entry[kOccBatch];
    uint32_t prefix[kOccBatch + 1];   // exclusive scan of the entries' row counts
    uint32_t warpSums[kOccThreads / 32];
    RangeBuffer ranges;
    uint8_t owner[kOccOwnerCap];      // batch entry of row i, for batches of at most kOccOwnerCap rows
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

/// Bitmap word of OUTPUT voxel (ox, oy, oz) for an entry: the chunk lookup is skipped when the entry's box lies in one
/// chunk.
__device__ __forceinline__ size_t entryWord(const OccupancyView &occ, uint32_t slot, uint32_t ox, uint32_t oy, uint32_t oz)
{
    if (slot == kNoSlot) {
        return bitmapWord(occ, ox, oy, oz);
    }
    const uint32_t tileLocal = ((ox >> 3) & 7u) | (((oy >> 3) & 7u) << 3) | (((oz >> 3) & 7u) << 6);
    return (size_t) slot * kChunkWords + tileLocal * kTileEdge + (oz & 7u);
}

/// Bitmap slot of the box [lo, hi) (sample space) if it lies in one chunk, else kNoSlot.
__device__ __forceinline__ uint32_t boxSlot(const OccupancyView &occ, const uint32_t lo[3], const uint32_t hi[3])
{
    const uint32_t cs = 6u + occ.shift;
    const bool oneChunk = (lo[0] >> cs) == ((hi[0] - 1) >> cs) && (lo[1] >> cs) == ((hi[1] - 1) >> cs) &&
                          (lo[2] >> cs) == ((hi[2] - 1) >> cs);
    return oneChunk ? __ldg(occ.chunkSlot + (lo[0] >> cs) +
                            occ.chunksPerAxis * ((lo[1] >> cs) + occ.chunksPerAxis * ((lo[2] >> cs) - occ.chunkZ0)))
                    : kNoSlot;
}

/// Fills one batch entry for `leaf` restricted to the box [lo, hi) (at most kOccBigVolume voxels).  Returns the number
/// of rows (0 = skip).
__device__ __forceinline__ uint32_t stageBatchEntry(BatchEntry &e, const OccupancyView &occ, uint32_t leafIndex,
                                                    const uint32_t lo[3], const uint32_t hi[3], LeafStage &s,
                                                    float certainMargin)
{
    e.slot = boxSlot(occ, lo, hi);
    const float origin[3] = {(float) lo[0], (float) lo[1], (float) lo[2]};
    buildPrefilter(s, origin);
    PairSat pair;
    buildPairSat(pair, s, origin, certainMargin);
    buildSpanSat(e.sat, pair);
    e.leaf = leafIndex;
    e.x0 = lo[0];
    e.y0 = lo[1];
    e.z0 = lo[2];
    e.dx = hi[0] - lo[0];
    e.dy = hi[1] - lo[1];
    e.magicDy = magicOf(e.dy);
    e.noSat = (s.flags & kLeafNoPrefilter) != 0 ? 1u : 0u;
    return e.dy * (hi[2] - lo[2]);
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

/// ORs the output voxels ox0 .. ox1 of output row (oy, oz) into the bitmap: one RED per tile column on the 32-bit half
/// word that holds the row (4 rows of one tile layer); nothing waits for the result.  A run of up to nine voxels — every
/// run of an ordinary leaf — touches at most two columns: two predicated REDs, no loop.
__device__ __forceinline__ void setRun(const OccupancyView &occ, uint32_t slot, uint32_t ox0, uint32_t ox1, uint32_t oy,
                                       uint32_t oz)
{
    uint32_t *bits32 = reinterpret_cast<uint32_t *>(occ.bits);
    const uint32_t rowShift = 8u * (oy & 3u);
    const uint32_t first = ox0 & 7u, length = ox1 - ox0 + 1u;
    if (slot != kNoSlot && first + length <= 16u) {
        // the whole box lies in one chunk: half word of tile column c = rowBase + 16 c
        uint32_t *column = bits32 + ((size_t) slot * (kChunkWords * 2u) +
                                     ((((ox0 >> 3) & 7u) | (((oy >> 3) & 7u) << 3) | (((oz >> 3) & 7u) << 6)) * kTileEdge +
                                      (oz & 7u)) * 2u +
                                     ((oy & 7u) >> 2));
        const uint32_t mask = ((1u << length) - 1u) << first;  // 16 bits: this column and the next
        atomicOr(column, (mask & 0xffu) << rowShift);
        if ((mask >> 8) != 0) {
            atomicOr(column + 16, (mask >> 8) << rowShift);
        }
        return;
    }
    for (uint32_t column = ox0 >> 3; column <= (ox1 >> 3); ++column) {
        const uint32_t from = max(ox0, column << 3) & 7u, to = min(ox1, (column << 3) + 7u) & 7u;
        const uint32_t mask = ((2u << to) - 1u) & ~((1u << from) - 1u);
        const size_t half = entryWord(occ, slot, column << 3, oy, oz) * 2u + ((oy & 7u) >> 2);
        atomicOr(bits32 + half, mask << rowShift);
    }
}

/// The block-wide part shared by the leaf-batch and the box kernels.  sh.entry[0 .. count) are staged and
/// sh.prefix[0 .. kOccBatch] holds the exclusive scan of their row counts (entries >= count contribute 0); the caller
/// has reset sh.ranges.count and synchronised.
__device__ __forceinline__ void classifyBatch(ClassifyShared &sh, const VoxelizeArgs &args)
{
    const OccupancyView &occ = args.occ;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t shift = occ.shift;
    const uint32_t total = sh.prefix[kOccBatch];

    // ---- who owns row i: a table for batches of ordinary leaves (two threads fill the rows of one entry), a binary
    // search over the prefix sums otherwise ----
    const bool table = total <= kOccOwnerCap;  // block-uniform
    if (table) {
        for (uint32_t t = tid; t < 2u * kOccBatch; t += kOccThreads) {
            const uint32_t p = t >> 1;
            const uint32_t first = sh.prefix[p], rows = sh.prefix[p + 1] - first, half = (rows + 1u) >> 1;
            const uint32_t from = (t & 1u) != 0 ? half : 0u, to = (t & 1u) != 0 ? rows : half;
            for (uint32_t r = from; r < to; ++r) {
                sh.owner[first + r] = (uint8_t) p;
            }
        }
        __syncthreads();
    }

    // ---- verdicts: lane = row, the flat row space of the batch shared out warp by warp ----
    for (uint32_t base = warp * 32u; base < total; base += kOccThreads) {  // warp-uniform trip count
        const uint32_t i = base + lane;
        if (i < total) {
            uint32_t p = 0;
            if (table) {
                p = sh.owner[i];
            }
            else {
#pragma unroll
                for (uint32_t step = kOccBatch / 2; step > 0; step >>= 1) {  // last entry with prefix[p] <= i
                    p += sh.prefix[p + step] <= i ? step : 0u;
                }
            }
            const BatchEntry &e = sh.entry[p];
            const uint32_t row = i - sh.prefix[p];
            const uint32_t zi = divideBy(row, e.dy, e.magicDy);
            const uint32_t yi = row - zi * e.dy;
            int i0 = 0, i1 = (int) e.dx - 1, j0 = 0, j1 = -1;
            if (args.prefilter && e.noSat == 0) {
                classifySpan(e.sat, (float) yi, (float) zi, (float) (e.dx - 1u), i0, i1, j0, j1);
            }
            if (i0 <= i1) {
                const uint32_t y = e.y0 + yi, z = e.z0 + zi;
                // undecided voxels: [i0, j0) and (j1, i1] — all of [i0, i1] without a certain run — minus the siblings of
                // this row's own certain voxels
                int a1 = i1, b0 = i1 + 1;
                if (j0 <= j1) {
                    const uint32_t ox0 = (e.x0 + (uint32_t) j0) >> shift, ox1 = (e.x0 + (uint32_t) j1) >> shift;
                    setRun(occ, e.slot, ox0, ox1, y >> shift, z >> shift);
                    a1 = min(j0 - 1, (int) (ox0 << shift) - (int) e.x0 - 1);
                    b0 = max(j1 + 1, (int) ((ox1 + 1u) << shift) - (int) e.x0);
                }
                const int lengthA = max(a1 - i0 + 1, 0), lengthB = max(i1 - b0 + 1, 0);
                if (lengthA + lengthB != 0) {
                    pushRange(sh.ranges, args, e.leaf, e.x0 + (uint32_t) i0, y, z, (uint32_t) lengthA,
                              (uint32_t) (b0 - i0 - lengthA), (uint32_t) lengthB);
                }
            }
        }
    }
    __syncthreads();
    flushRanges(sh.ranges, args);
}

/// One block = kOccBatch consecutive leaves (big leaves are left to the box kernel).
__global__ void __launch_bounds__(kOccThreads, O2V_OCC_MIN_BLOCKS)
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
                volume = stageBatchEntry(sh.entry[tid], args.occ, leafIndex, lo, hi, s, args.certainMargin);
            }
        }
    }
    if (tid == 0) {
        sh.ranges.count = 0;
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

/// Thread = leaf, for meshes of micro-triangles (on average at most kOccDirectCandidates candidate voxels per leaf:
/// BASELINE config 5).  Sharing a leaf's one or two rows out over a block costs more than testing its few voxels: the
/// per-voxel form of the SAT with its constants in registers, the thread walks its own rows.  Same bitmap as
/// occupancyClassifyKernel, undecided voxels straight to the clip queue; big leaves are left to the box kernel.
__global__ void __launch_bounds__(kOccDirectThreads)
occupancyClassifyDirectKernel(const VoxelizeArgs args, uint32_t leafTotal)
{
    __shared__ RangeBuffer ranges;
    const OccupancyView &occ = args.occ;
    const uint32_t tid = threadIdx.x;
    const uint32_t shift = occ.shift;
    uint32_t *bits32 = reinterpret_cast<uint32_t *>(occ.bits);
    if (tid == 0) {
        ranges.count = 0;
    }
    __syncthreads();

    const uint32_t leafIndex = blockIdx.x * kOccDirectThreads + tid;
    if (leafIndex < leafTotal) {
        LeafStage s;
        loadLeafVertices(s, args, leafIndex);
        uint32_t lo[3], hi[3];
        if ((s.flags & kLeafEmpty) == 0 && leafBoxInSlab(s.v, args.grid, lo, hi) &&
            (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]) <= kOccBigVolume) {
            const uint32_t slot = boxSlot(occ, lo, hi);
            const float origin[3] = {(float) lo[0], (float) lo[1], (float) lo[2]};
            buildPrefilter(s, origin);
            PairSat sat;
            buildPairSat(sat, s, origin, args.certainMargin);
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
                            // the segment's output voxels: itself, or its 4 parents (bits 2k and 2k + 1 -> bit k)
                            uint32_t mask = sure;
                            if (shift != 0) {
                                mask = (sure | (sure >> 1)) & 0x55u;
                                mask = (mask | (mask >> 1)) & 0x33u;
                                mask = ((mask | (mask >> 2)) & 0x0fu) << ((x8 >> 1) & 4u);
                                // siblings of a certain voxel need no clip
                                const uint32_t pairs = (sure | (sure >> 1)) & 0x55u;
                                open &= ~(pairs | (pairs << 1));
                            }
                            const uint32_t oy = y >> shift;
                            const size_t half = entryWord(occ, slot, x8 >> shift, oy, z >> shift) * 2u + ((oy & 7u) >> 2);
                            atomicOr(bits32 + half, mask << (8u * (oy & 3u)));
                        }
                        while (open != 0) {
                            const uint32_t bit = (uint32_t) __ffs((int) open) - 1u;
                            open &= open - 1u;
                            pushVoxel(ranges, args, leafIndex, x8 + bit, y, z);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    flushVoxels(ranges, args);
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
            const uint32_t volume = stageBatchEntry(sh.entry[0], args.occ, row.x, lo3, hi3, s, args.certainMargin);
            sh.ranges.count = 0;
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
// filter: undecided ranges -> queue of the voxels the bitmap has not decided

/// Appends `count` voxels per lane, {leaf, xy, z} from the lane's arrays, to the clip queue: one atomic for the warp
/// (all 32 lanes call).
template <uint32_t N>
__device__ __forceinline__ void appendSurvivors(const VoxelizeArgs &args, const uint32_t (&leaf)[N],
                                                const uint32_t (&xy)[N], const uint32_t (&z)[N], uint32_t count)
{
    const unsigned int full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t inclusive = count;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(full, inclusive, o);
        inclusive += lane >= (uint32_t) o ? up : 0u;
    }
    const uint32_t warpTotal = __shfl_sync(full, inclusive, 31);
    if (warpTotal == 0) {
        return;
    }
    unsigned long long index = 0;
    if (lane == 31) {
        index = atomicAdd(&args.counters->survivors, (unsigned long long) warpTotal);
    }
    index = __shfl_sync(full, index, 31) + (inclusive - count);
#pragma unroll
    for (uint32_t k = 0; k < N; ++k) {
        if (k < count && index + k < args.occ.queueCapacity) {  // beyond: counted only; the engine grows the queue and reruns
            args.occ.queue[index + k] = make_uint4(leaf[k], xy[k], z[k], 0u);
        }
    }
}

/// Thread = kFilterEntries range entries per round.  Runs after every classifier of the run: all `certain` bits are in
/// place, so a voxel whose output voxel is already set (by another leaf, or by a sibling when downscaling) needs no
/// clip — about two thirds of the undecided voxels on BASELINE config 4.  The kernel is a gather: what matters is how
/// many probes are in flight, so a thread loads its entries, then issues the probes of the first kFilterProbe voxels of
/// all their runs back to back, and the warp reserves queue slots with one atomic per round; the rest of a long run
/// (slivers: whole rows) follows in a plain loop.
__global__ void __launch_bounds__(kOccFilterThreads)
occupancyFilterQueueKernel(const VoxelizeArgs args)
{
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long total = args.counters->ranges;
    total = total < occ.rangeCapacity ? total : occ.rangeCapacity;
    const unsigned long long perRound = (unsigned long long) gridDim.x * blockDim.x;
    constexpr uint32_t kSlots = kFilterEntries * 2 * kFilterProbe;
    for (unsigned long long base = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x - lane;
         base < total; base += perRound * kFilterEntries) {  // warp-uniform
        uint4 r[kFilterEntries];
#pragma unroll
        for (uint32_t e = 0; e < kFilterEntries; ++e) {
            const unsigned long long i = base + e * perRound + lane;
            // entries beyond the end probe the last entry's first voxel (a valid address) and keep nothing
            r[e] = __ldcs(occ.ranges + (i < total ? i : total - 1));
            r[e].w = i < total ? r[e].w : 0u;
        }
        uint32_t px[kSlots];
        bool wanted[kSlots], decided[kSlots];
#pragma unroll
        for (uint32_t e = 0; e < kFilterEntries; ++e) {
            const uint32_t x0 = r[e].y & 0xffffu, lengthA = r[e].w & 0xffffu, lengthB = r[e].w >> 16;
            const uint32_t xB = x0 + lengthA + (r[e].z >> 16);
#pragma unroll
            for (uint32_t k = 0; k < kFilterProbe; ++k) {
                wanted[(e * 2) * kFilterProbe + k] = k < lengthA;
                wanted[(e * 2 + 1) * kFilterProbe + k] = k < lengthB;
                px[(e * 2) * kFilterProbe + k] = k < lengthA ? x0 + k : x0;  // (x0 is always inside the grid)
                px[(e * 2 + 1) * kFilterProbe + k] = k < lengthB ? xB + k : x0;
            }
        }
#pragma unroll
        for (uint32_t s = 0; s < kSlots; ++s) {  // every probe loads: no branch between them
            const uint32_t e = s / (2 * kFilterProbe);
            decided[s] = alreadyDecided(occ, px[s], r[e].y >> 16, r[e].z & 0xffffu);
        }
        uint32_t leaf[kSlots], xy[kSlots], z[kSlots], count = 0;
#pragma unroll
        for (uint32_t s = 0; s < kSlots; ++s) {
            const uint32_t e = s / (2 * kFilterProbe);
            if (wanted[s] && !decided[s]) {
                leaf[count] = r[e].x;
                xy[count] = px[s] | (r[e].y & 0xffff0000u);
                z[count] = r[e].z & 0xffffu;
                ++count;
            }
        }
        appendSurvivors<kSlots>(args, leaf, xy, z, count);
        // what is left of long runs
#pragma unroll
        for (uint32_t e = 0; e < kFilterEntries; ++e) {
            const uint32_t x0 = r[e].y & 0xffffu, y = r[e].y >> 16, zz = r[e].z & 0xffffu;
            const uint32_t lengthA = r[e].w & 0xffffu, lengthB = r[e].w >> 16;
            const uint32_t xB = x0 + lengthA + (r[e].z >> 16);
            const uint32_t longest = __reduce_max_sync(full, max(lengthA, lengthB));
            for (uint32_t k = kFilterProbe; k < longest; ++k) {  // warp-uniform trip count
                uint32_t l2[2], xy2[2], z2[2], n = 0;
                if (k < lengthA && !alreadyDecided(occ, x0 + k, y, zz)) {
                    l2[n] = r[e].x; xy2[n] = (x0 + k) | (y << 16); z2[n] = zz; ++n;
                }
                if (k < lengthB && !alreadyDecided(occ, xB + k, y, zz)) {
                    l2[n] = r[e].x; xy2[n] = (xB + k) | (y << 16); z2[n] = zz; ++n;
                }
                appendSurvivors<2>(args, l2, xy2, z2, n);
            }
        }
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
                    if (!alreadyDecided(occ, x, y, z)) {
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
                        word = occ.bits + bitmapWord(occ, x >> occ.shift, y >> occ.shift, z >> occ.shift);
                        bit = bitmapBit(x >> occ.shift, y >> occ.shift);
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

/// Two steps per warp and 32 tiles: (1) lane = tile: load the 8 layer words (already in output space), count;
/// (2) lane = output record: records are numbered across the warp's tiles, record j finds its tile, layer and bit by
/// rank/select in shared memory — all lanes busy whatever the fill of the tiles, and a warp stores 512 contiguous bytes.
__global__ void __launch_bounds__(kOccExpandThreads)
occupancyExpandKernel(const VoxelizeArgs args)
{
    __shared__ ExpandShared sh;
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
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
            if (occ.emitted == nullptr) {
#pragma unroll
                for (int k = 0; k < (int) kTileEdge / 2; ++k) {
                    const ulonglong2 w = __ldcs(src + k);  // read once
                    m[2 * k] = w.x;
                    m[2 * k + 1] = w.y;
                }
            }
            else {  // a piece of a job that accumulates: only what no earlier piece has delivered (block-uniform)
                const ulonglong2 *old = reinterpret_cast<const ulonglong2 *>(occ.emitted + t * kTileEdge);
#pragma unroll
                for (int k = 0; k < (int) kTileEdge / 2; ++k) {
                    const ulonglong2 w = __ldcg(src + k), e = __ldcg(old + k);
                    m[2 * k] = w.x & ~e.x;
                    m[2 * k + 1] = w.y & ~e.y;
                }
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
            origin[0] = (chunk % C) * kChunkEdge + (tileLocal & 7u) * kTileEdge;
            origin[1] = ((chunk / C) % C) * kChunkEdge + ((tileLocal >> 3) & 7u) * kTileEdge;
            origin[2] = (chunk / (C * C) + occ.chunkZ0) * kChunkEdge + (tileLocal >> 6) * kTileEdge;
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
                const uint32_t x = info.origin[0] + (b & 7u), y = info.origin[1] + (b >> 3), zz = info.origin[2] + z;
                if (args.packedBits == 32) {  // (block-uniform) positions only: a host expands them (o2v_job.cpp)
                    __stcs(reinterpret_cast<uint32_t *>(args.out) + index + j, x | (y << 10) | (zz << 20));
                }
                else if (args.packedBits == 64) {
                    __stcs(reinterpret_cast<unsigned long long *>(args.out) + index + j,
                           (unsigned long long) x | ((unsigned long long) y << 21) | ((unsigned long long) zz << 42));
                }
                else {
                    VoxelRecord rec;
                    rec.x = (int32_t) x;
                    rec.y = (int32_t) y;
                    rec.z = (int32_t) zz;
                    rec.argb = 0xFFFFFFFFu;  // quantizeArgb(1, 1, 1)
                    __stcs(reinterpret_cast<int4 *>(args.out + index + j), *reinterpret_cast<const int4 *>(&rec));
                }
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

static int occHugeBlocks(unsigned long long expected)
{
    const unsigned long long blocks = (expected * kHugeSubtrees + kOccHugeThreads - 1) / kOccHugeThreads;
    return (int) std::min<unsigned long long>(std::max<unsigned long long>(blocks, 1), 148ull * 32);
}

void launchOccupancyCount(const MeshView &mesh, const GridView &grid, const OccupancyView &occ, uint32_t *extraCount,
                          LeafRecord *firstLeaves, RunCounters *counters, bool countFromFilter, const HugeWork &work,
                          unsigned long long hugeExpected, int smCount, cudaStream_t stream)
{
    // countFromFilter: mesh.count is only an upper bound (the grid is sized by it; blocks without a batch leave at once)
    occupancyCountKernel<<<setupBlocks(occupancyCountKernel, mesh.count, smCount), kOccSetupThreads, 0, stream>>>(
        mesh, grid, occ, extraCount, firstLeaves, counters, countFromFilter, work);
    if (work.capacity != 0) {
        launchHugeSubtreeScan(mesh, grid, work, counters, extraCount, true, hugeExpected, stream);
        occHugeCountKernel<<<occHugeBlocks(hugeExpected), kOccHugeThreads, 0, stream>>>(mesh, grid, occ, work, firstLeaves,
                                                                                        counters);
    }
}

void launchOccupancyAllChunks(const OccupancyView &occ, cudaStream_t stream)
{
    occupancyAllChunksKernel<<<(occ.chunkTotal + 255) / 256, 256, 0, stream>>>(occ);
}

void launchOccupancySlabScatter(const MeshView &mesh, const GridView &grid, const SlabScatter &scatter, int smCount,
                                cudaStream_t stream)
{
    occupancySlabScatterKernel<<<setupBlocks(occupancySlabScatterKernel, mesh.count, smCount), kOccSetupThreads, 0,
                                 stream>>>(mesh, grid, scatter);
}

void launchOccupancyZHistogram(const MeshView &mesh, const GridView &grid, uint32_t unit, uint32_t rows,
                               unsigned long long *histogram, int smCount, cudaStream_t stream)
{
    occupancyZHistogramKernel<<<setupBlocks(occupancyZHistogramKernel, mesh.count, smCount), kOccSetupThreads, 0, stream>>>(
        mesh, grid, unit, rows, histogram);
}

void launchOccupancyChunkCount(const OccupancyView &occ, uint32_t *chunkCounts, RunCounters *counters, int smCount,
                               cudaStream_t stream)
{
    if (occ.activeChunks == 0) {
        return;
    }
    const unsigned blocks = occ.activeChunks < (unsigned) smCount * 8u ? occ.activeChunks : (unsigned) smCount * 8u;
    occupancyChunkCountKernel<<<blocks, 256, 0, stream>>>(occ, chunkCounts, counters);
}

void launchOccupancyAssignChunks(const OccupancyView &occ, RunCounters *counters, cudaStream_t stream)
{
    occupancyAssignChunksKernel<<<(occ.chunkTotal + 255) / 256, 256, 0, stream>>>(occ, counters);
}

void launchOccupancyEmit(const MeshView &mesh, const GridView &grid, const OccupancyView &occ,
                         const uint32_t *leafOffset, LeafRecord *leaves, RunCounters *counters, const HugeWork &work,
                         unsigned long long hugeExpected, int smCount, cudaStream_t stream)
{
    occupancyEmitKernel<<<setupBlocks(occupancyEmitKernel, mesh.count, smCount), kOccSetupThreads, 0, stream>>>(
        mesh, grid, occ, leafOffset, leaves, counters);
    if (work.capacity != 0) {
        occHugeEmitKernel<<<occHugeBlocks(hugeExpected), kOccHugeThreads, 0, stream>>>(mesh, grid, occ, work, leafOffset,
                                                                                       leaves, counters);
    }
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

void launchOccupancyFilterQueue(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    // the length of the range list lives on the device (RunCounters::ranges)
    occupancyFilterQueueKernel<<<occupancyPersistentBlocks(occupancyFilterQueueKernel, kOccFilterThreads, smCount),
                                 kOccFilterThreads, 0, stream>>>(args);
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
