// Hand-written sm_100a kernels of the B200 voxelizer (see o2v_kernels.cuh / DESIGN.md §3).
//
// Work decomposition (B200-first, not the reference's 64^3 chunks + worker threads):
//   * triangles -> leaves of the reference's subdivision (exact arithmetic, deterministic order = (triangle, DFS order))
//   * leaves are binned into 8^3-voxel tiles; every tile owns an ascending list of leaf indices
//   * one thread block voxelizes one tile: thread = voxel, the tile's leaves stream through shared memory in list order,
//     so the per-voxel fold order (ascending triangle index, DFS order inside a triangle) of the reference
//     (src/obj2voxel.cpp:226-243,270-272; src/voxelization.cpp:56-63,513-526) is reproduced without atomics on voxel
//     data, without a hash table and without a sort of contributions: accumulators live in registers and only the final
//     16-byte Voxel32 records are written to HBM (coalesced, block-compacted).
//   * candidate voxels are culled by a conservative triangle/box SAT (FMA allowed, margin kPrefilterMargin); survivors run
//     the bit-exact six-plane clip of o2v_exact.cuh.
#include "o2v_kernels.cuh"
#include <algorithm>

#include "o2v_device.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace o2v {

namespace {

constexpr int kSetupThreads = 128;

__device__ __forceinline__ unsigned int orderedBits(float f)
{
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float fromOrderedBits(unsigned int u)
{
    const unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(b);
}

// ---------------------------------------------------------------------------------------------------------------------
// mesh bounds: src/obj2voxel.cpp:180-200 (min/max are exact, any reduction order gives the reference's result)

__global__ void boundsKernel(const float *__restrict__ verts, unsigned long long vertexCount, RunCounters *counters)
{
    float mn[3] = {INFINITY, INFINITY, INFINITY};
    float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < vertexCount;
         i += stride) {
        for (int a = 0; a < 3; ++a) {
            const float c = verts[i * 3 + a];
            mn[a] = fminf(mn[a], c);
            mx[a] = fmaxf(mx[a], c);
        }
    }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; ++a) {
            atomicMin(&counters->boundsMinBits[a], orderedBits(mn[a]));
            atomicMax(&counters->boundsMaxBits[a], orderedBits(mx[a]));
        }
    }
}

__global__ void finishBoundsKernel(RunCounters *counters)
{
    if (threadIdx.x < 3) {
        counters->boundsMin[threadIdx.x] = fromOrderedBits(counters->boundsMinBits[threadIdx.x]);
        counters->boundsMax[threadIdx.x] = fromOrderedBits(counters->boundsMaxBits[threadIdx.x]);
    }
}

__device__ __forceinline__ uint32_t localTileId(const GridView &grid, uint32_t tx, uint32_t ty, uint32_t tz)
{
    return ((tz - grid.slabTileZ0) * grid.tilesPerAxis + ty) * grid.tilesPerAxis + tx;
}

/// What the count pass does with one leaf: the tiles its (clamped) box reaches get one more leaf and its candidates.
__device__ __forceinline__ void countLeafTiles(const GridView &grid, const uint32_t *lo, const uint32_t *hi,
                                               uint32_t *__restrict__ tileCount, uint32_t *__restrict__ tileCandidates,
                                               unsigned long long &candidates)
{
    candidates += (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    for (uint32_t tz = lo[2] / kTileEdge; tz <= (hi[2] - 1) / kTileEdge; ++tz) {
        const uint32_t dz = min(hi[2], (tz + 1) * kTileEdge) - max(lo[2], tz * kTileEdge);
        for (uint32_t ty = lo[1] / kTileEdge; ty <= (hi[1] - 1) / kTileEdge; ++ty) {
            const uint32_t dy = min(hi[1], (ty + 1) * kTileEdge) - max(lo[1], ty * kTileEdge);
            for (uint32_t tx = lo[0] / kTileEdge; tx <= (hi[0] - 1) / kTileEdge; ++tx) {
                const uint32_t dx = min(hi[0], (tx + 1) * kTileEdge) - max(lo[0], tx * kTileEdge);
                const uint32_t tile = localTileId(grid, tx, ty, tz);
                atomicAdd(&tileCount[tile], 1u);
                atomicAdd(&tileCandidates[tile], dx * dy * dz);
            }
        }
    }
}

/// A huge triangle is only listed here (`work`; its leaf count stays 0): hugeSubtreeCountKernel .. hugeCountTilesKernel
/// do for it what this kernel does for the others, spread over the device.
template <bool UV>
__global__ void __launch_bounds__(kSetupThreads)
countLeavesKernel(MeshView mesh, GridView grid, uint32_t *__restrict__ leafCount, uint32_t *__restrict__ tileCount,
                  uint32_t *__restrict__ tileCandidates, RunCounters *counters, HugeWork work)
{
    unsigned long long candidates = 0, dropped = 0, overflow = 0;
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < mesh.count;
         i += stride) {
        Tri<UV> root;
        float area;
        uint32_t leaves = 0;
        if (loadTriangle<UV>(mesh, grid, i, root, area)) {
            bool huge = false;
            const bool ok = traverseLeaves<UV>(root, grid, [&](const Tri<UV> &, const uint32_t *lo, const uint32_t *hi) {
                ++leaves;
                countLeafTiles(grid, lo, hi, tileCount, tileCandidates, candidates);
            }, &huge);
            overflow += ok ? 0 : 1;
            if (huge) {
                listHugeTriangle(work, counters, i);
            }
        }
        else {
            ++dropped;
        }
        leafCount[i] = leaves;
    }
    warpTally(&counters->candidateVoxels, candidates);
    warpTally(&counters->droppedTriangles, dropped);
    warpTally(&counters->depthOverflow, overflow);
}

// ---- the huge triangles of a run: (triangle, subtree) items over the whole device ----

constexpr int kHugeThreads = 32;  // a thread walks whole subtrees, latency-bound: many small blocks over all SMs

/// subtree[item] = the leaves of the item's subtree that the passes visit (geometry only: no texture coordinates).
__global__ void __launch_bounds__(kHugeThreads)
hugeSubtreeCountKernel(MeshView mesh, GridView grid, HugeWork work, const RunCounters *counters)
{
    const unsigned long long seen = counters->hugeTriangles;
    const unsigned long long items = (seen < work.capacity ? seen : work.capacity) * kHugeSubtrees;
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long item = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; item < items;
         item += stride) {
        Tri<false> root;
        float area = 0.0f;
        loadTriangle<false>(mesh, grid, work.list[item / kHugeSubtrees], root, area);
        uint32_t leaves = 0;
        forEachLeafOfSubtree<false>(root, kHugeSplitDepth, (uint32_t) (item % kHugeSubtrees), [&](const Tri<false> &leaf) {
            visitClamped<false>(leaf, grid, [&](const Tri<false> &, const uint32_t *, const uint32_t *) { ++leaves; });
        });
        work.subtree[item] = leaves;
    }
}

/// Block = listed triangle: subtree[] becomes its exclusive scan (the place of every subtree in the triangle's leaf
/// sequence) and the triangle's leaf count goes where the count pass would have put it — leafCount[tri] on the weighted
/// pipeline; on the occupancy pipeline extraCount[tri] = leaves beyond the first, and the leaf tallies.
__global__ void __launch_bounds__(kHugeSubtrees)
hugeScanKernel(HugeWork work, RunCounters *counters, uint32_t *__restrict__ perTriangle, bool occupancy)
{
    __shared__ uint32_t warpSums[kHugeSubtrees / 32];
    const unsigned long long seen = counters->hugeTriangles;
    const uint32_t listed = (uint32_t) (seen < work.capacity ? seen : work.capacity);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t h = blockIdx.x; h < listed; h += gridDim.x) {
        uint32_t *mine = work.subtree + (size_t) h * kHugeSubtrees + threadIdx.x;
        const uint32_t value = *mine;
        uint32_t inclusive = value;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inclusive, o);
            inclusive += lane >= (uint32_t) o ? up : 0u;
        }
        if (lane == 31) {
            warpSums[warp] = inclusive;
        }
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t w = 0; w < kHugeSubtrees / 32; ++w) {
            before += w < warp ? warpSums[w] : 0u;
            total += warpSums[w];
        }
        *mine = before + inclusive - value;
        if (threadIdx.x == 0) {
            const uint32_t tri = work.list[h];
            if (occupancy) {
                const uint32_t extra = total > 1u ? total - 1u : 0u;
                perTriangle[tri] = extra;
                atomicAdd(&counters->leaves, (unsigned long long) total);
                atomicAdd(&counters->extraLeaves, (unsigned long long) extra);
            }
            else {
                perTriangle[tri] = total;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kHugeThreads)
hugeCountTilesKernel(MeshView mesh, GridView grid, HugeWork work, uint32_t *__restrict__ tileCount,
                     uint32_t *__restrict__ tileCandidates, RunCounters *counters)
{
    unsigned long long candidates = 0;
    const unsigned long long overflow = forEachHugeLeaf<false>(
        mesh, grid, work, counters, false,
        [&](uint32_t, float, const Tri<false> &, const uint32_t *lo, const uint32_t *hi, uint32_t) {
            countLeafTiles(grid, lo, hi, tileCount, tileCandidates, candidates);
        });
    warpTally(&counters->candidateVoxels, candidates);
    warpTally(&counters->depthOverflow, overflow);
}

struct EmitTargets {
    const uint32_t *leafOffset, *tileStart;
    uint32_t *tileFill;
    LeafRecord *leaves;
    LeafUv *leafUvs;
    uint32_t *tileList, *pairTile;
};

/// What the emit pass does with one leaf: record `index`, and an entry in the list of every tile its box reaches.
template <bool UV>
__device__ __forceinline__ void emitLeaf(const GridView &grid, const EmitTargets &out, uint32_t index, uint32_t tri,
                                         float area, const Tri<UV> &leaf, const uint32_t *lo, const uint32_t *hi)
{
    LeafRecord rec;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        rec.v[k] = leaf.v[k];
    }
    rec.tri = tri;
    rec.area = area;
    rec.flags = leafFlagsOf(leaf.v);
    out.leaves[index] = rec;
    if (UV) {
        LeafUv uv;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            uv.t[k] = leaf.t[k];
        }
        uv.pad[0] = uv.pad[1] = 0.0f;
        out.leafUvs[index] = uv;
    }
    for (uint32_t tz = lo[2] / kTileEdge; tz <= (hi[2] - 1) / kTileEdge; ++tz) {
        for (uint32_t ty = lo[1] / kTileEdge; ty <= (hi[1] - 1) / kTileEdge; ++ty) {
            for (uint32_t tx = lo[0] / kTileEdge; tx <= (hi[0] - 1) / kTileEdge; ++tx) {
                const uint32_t tile = localTileId(grid, tx, ty, tz);
                const uint32_t slot = atomicAdd(&out.tileFill[tile], 1u);
                out.tileList[out.tileStart[tile] + slot] = index;
                out.pairTile[out.tileStart[tile] + slot] = tile;
            }
        }
    }
}

/// Huge triangles are skipped here (their leaves: hugeEmitLeavesKernel).
template <bool UV>
__global__ void __launch_bounds__(kSetupThreads)
emitLeavesKernel(MeshView mesh, GridView grid, const uint32_t *__restrict__ leafOffset,
                 const uint32_t *__restrict__ tileStart, uint32_t *__restrict__ tileFill,
                 LeafRecord *__restrict__ leaves, LeafUv *__restrict__ leafUvs, uint32_t *__restrict__ tileList,
                 uint32_t *__restrict__ pairTile)
{
    const EmitTargets out{leafOffset, tileStart, tileFill, leaves, leafUvs, tileList, pairTile};
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < mesh.count;
         i += stride) {
        Tri<UV> root;
        float area;
        if (!loadTriangle<UV>(mesh, grid, i, root, area)) {
            continue;
        }
        uint32_t index = leafOffset[i];
        bool huge = false;
        traverseLeaves<UV>(root, grid, [&](const Tri<UV> &leaf, const uint32_t *lo, const uint32_t *hi) {
            emitLeaf<UV>(grid, out, index, static_cast<uint32_t>(i), area, leaf, lo, hi);
            ++index;
        }, &huge);
    }
}

/// A leaf's index = its triangle's offset + its position in the triangle's (reference-order) leaf sequence.
template <bool UV>
__global__ void __launch_bounds__(kHugeThreads)
hugeEmitLeavesKernel(MeshView mesh, GridView grid, HugeWork work, EmitTargets out, const RunCounters *counters)
{
    forEachHugeLeaf<UV>(mesh, grid, work, counters, true,
                        [&](uint32_t tri, float area, const Tri<UV> &leaf, const uint32_t *lo, const uint32_t *hi,
                            uint32_t seq) { emitLeaf<UV>(grid, out, out.leafOffset[tri] + seq, tri, area, leaf, lo, hi); });
}

// ---------------------------------------------------------------------------------------------------------------------
// exclusive scan (u32 in, u32 out, u64 total): reduce -> single-block scan of block sums -> apply

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long blockReduceU64(unsigned long long v, unsigned long long *smem)
{
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        smem[warp] = v;
    }
    __syncthreads();
    unsigned long long total = 0;
    if (warp == 0) {
        total = lane < (blockDim.x >> 5) ? smem[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) {
            total += __shfl_xor_sync(0xffffffffu, total, o);
        }
    }
    return total;  // valid in warp 0
}

__global__ void __launch_bounds__(kScanThreads)
scanReduceKernel(const uint32_t *__restrict__ in, size_t n, unsigned long long *__restrict__ blockSums)
{
    __shared__ unsigned long long smem[32];
    const size_t base = (size_t) blockIdx.x * kScanBlock;
    unsigned long long sum = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const size_t i = base + (size_t) k * kScanThreads + threadIdx.x;
        sum += i < n ? in[i] : 0u;
    }
    const unsigned long long total = blockReduceU64(sum, smem);
    if (threadIdx.x == 0) {
        blockSums[blockIdx.x] = total;
    }
}

__global__ void __launch_bounds__(1024)
scanTopKernel(unsigned long long *__restrict__ blockSums, size_t blocks, unsigned long long *__restrict__ total)
{
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) {
        carry = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (size_t base = 0; base < blocks; base += 1024) {
        const size_t i = base + threadIdx.x;
        const unsigned long long v = i < blocks ? blockSums[i] : 0;
        unsigned long long inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o);
            inc += lane >= o ? up : 0;
        }
        if (lane == 31) {
            warpSums[warp] = inc;
        }
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warpSums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, w, o);
                w += lane >= o ? up : 0;
            }
            warpSums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const unsigned long long before = carry + (warp > 0 ? warpSums[warp - 1] : 0) + (inc - v);
        if (i < blocks) {
            blockSums[i] = before;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry += warpSums[31];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = carry;
    }
}

__global__ void __launch_bounds__(kScanThreads)
scanApplyKernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n,
                const unsigned long long *__restrict__ blockSums)
{
    __shared__ uint32_t warpSums[kScanThreads / 32];
    const size_t base = (size_t) blockIdx.x * kScanBlock + (size_t) threadIdx.x * kScanItems;
    uint32_t items[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        items[k] = base + k < n ? in[base + k] : 0u;
        sum += items[k];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        inc += lane >= o ? up : 0;
    }
    if (lane == 31) {
        warpSums[warp] = inc;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? warpSums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, w, o);
            w += lane >= o ? up : 0;
        }
        if (lane < kScanThreads / 32) {
            warpSums[lane] = w;
        }
    }
    __syncthreads();
    uint32_t running = static_cast<uint32_t>(blockSums[blockIdx.x]) + (warp > 0 ? warpSums[warp - 1] : 0) + (inc - sum);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) {
            out[base + k] = running;
        }
        running += items[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// active tile compaction (order irrelevant: tiles are independent)

__global__ void compactActiveTilesKernel(const uint32_t *__restrict__ tileCount,
                                         const uint32_t *__restrict__ tileCandidates,
                                         const uint32_t *__restrict__ tileStart, uint32_t tileTotal,
                                         uint32_t *__restrict__ allTiles, uint32_t *__restrict__ longTiles,
                                         uint32_t *__restrict__ heavyTiles, LightTile *__restrict__ lightTiles,
                                         LightTile *__restrict__ bigLightTiles, RunCounters *counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = i < tileTotal ? tileCount[i] : 0u;
    const uint32_t candidates = count != 0 ? tileCandidates[i] : 0u;
    const bool sparse = count != 0 && candidates <= kLightMaxCandidates;
    const bool light = sparse && candidates <= kWarpFoldMax;
    const bool bigLight = sparse && !light;
    const bool heavy = count != 0 && !sparse;
    const int lane = threadIdx.x & 31;
    const unsigned int below = (1u << lane) - 1u;

    const unsigned int lightBallot = __ballot_sync(0xffffffffu, light);
    if (lightBallot != 0) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(&counters->lightTiles, (unsigned long long) __popc(lightBallot));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (light) {
            LightTile d;
            d.tile = i;
            d.listStart = tileStart[i];
            d.leafCount = count;
            d.candidates = candidates;
            lightTiles[base + __popc(lightBallot & below)] = d;
        }
    }
    const unsigned int bigBallot = __ballot_sync(0xffffffffu, bigLight);
    if (bigBallot != 0) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(&counters->bigLightTiles, (unsigned long long) __popc(bigBallot));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (bigLight) {
            LightTile d;
            d.tile = i;
            d.listStart = tileStart[i];
            d.leafCount = count;
            d.candidates = candidates;
            bigLightTiles[base + __popc(bigBallot & below)] = d;
        }
    }
    const unsigned int heavyBallot = __ballot_sync(0xffffffffu, heavy);
    if (heavyBallot != 0) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(&counters->heavyTiles, (unsigned long long) __popc(heavyBallot));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (heavy) {
            heavyTiles[base + __popc(heavyBallot & below)] = i;
        }
    }
    const unsigned int longBallot = __ballot_sync(0xffffffffu, count > 32u);
    if (longBallot != 0) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(&counters->longTiles, (unsigned long long) __popc(longBallot));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (count > 32u) {
            longTiles[base + __popc(longBallot & below)] = i;
        }
    }
    const unsigned int anyBallot = lightBallot | bigBallot | heavyBallot;
    if (anyBallot != 0) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(&counters->activeTiles, (unsigned long long) __popc(anyBallot));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (count != 0) {
            allTiles[base + __popc(anyBallot & below)] = i;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-tile list sort (ascending leaf index == ascending (triangle, DFS order)); the atomic fill order is arbitrary

__global__ void __launch_bounds__(256)
sortSmallListsKernel(TileWork work, uint32_t *__restrict__ tileList)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < work.allCount; w += warpsPerGrid) {
        const uint32_t tile = work.allTiles[w];
        const uint32_t n = work.tileCount[tile];
        if (n < 2 || n > 32) {
            continue;
        }
        uint32_t *list = tileList + work.tileStart[tile];
        const uint32_t mine = lane < n ? list[lane] : 0xffffffffu;
        uint32_t rank = 0;
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t other = __shfl_sync(0xffffffffu, mine, k);
            rank += other < mine ? 1u : 0u;
        }
        __syncwarp();
        if (lane < n) {
            list[rank] = mine;
        }
    }
}

constexpr uint32_t kSortSmemCap = 4096;

__global__ void __launch_bounds__(512)
sortLargeListsKernel(TileWork work, uint32_t *__restrict__ tileList)
{
    __shared__ uint32_t keys[kSortSmemCap];
    for (uint32_t w = blockIdx.x; w < work.longCount; w += gridDim.x) {
        const uint32_t tile = work.longTiles[w];
        const uint32_t n = work.tileCount[tile];
        if (n <= 32) {
            continue;
        }
        uint32_t *list = tileList + work.tileStart[tile];
        uint32_t *a = n <= kSortSmemCap ? keys : list;  // long lists are sorted in place (L2 resident)
        if (n <= kSortSmemCap) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                keys[i] = list[i];
            }
        }
        __syncthreads();
        uint32_t padded = 1;
        while (padded < n) {
            padded <<= 1;
        }
        // ascending-only bitonic network: virtual +inf padding beyond n never moves
        for (uint32_t k = 2; k <= padded; k <<= 1) {
            for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x) {
                const uint32_t l = i ^ (k - 1);
                if (l > i && l < n) {
                    const uint32_t x = a[i], y = a[l];
                    if (x > y) {
                        a[i] = y;
                        a[l] = x;
                    }
                }
            }
            __syncthreads();
            for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x) {
                    const uint32_t l = i ^ j;
                    if (l > i && l < n) {
                        const uint32_t x = a[i], y = a[l];
                        if (x > y) {
                            a[i] = y;
                            a[l] = x;
                        }
                    }
                }
                __syncthreads();
            }
        }
        if (n <= kSortSmemCap) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                list[i] = keys[i];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the hot kernel: one block = one 8^3 tile, thread = voxel

constexpr int kTileThreads = 512;

struct TileShared {
    LeafStage stage[kLeafBatch];
    uint32_t tileSlot;
    uint32_t warpCounts[kTileThreads / 32];
    unsigned long long outBase;
    // supersampling fold (8^3 children -> 4^3 parents)
    float dsW[kTileVoxels], dsR[kTileVoxels], dsG[kTileVoxels], dsB[kTileVoxels];
    uint8_t dsHas[kTileVoxels];
    uint8_t caseTable[64];  // clipCaseOf lookup for the warp-synchronous clip
};

template <bool UV>
__global__ void __launch_bounds__(kTileThreads, 1)
voxelizeTilesKernel(const VoxelizeArgs args)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    TileShared &sh = *reinterpret_cast<TileShared *>(smemRaw);

    const uint32_t tid = threadIdx.x;
    const uint32_t lx = tid & 7u, ly = (tid >> 3) & 7u, lz = tid >> 6;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    unsigned long long clipCalls = 0, contributions = 0;
    fillClipCaseTable(sh.caseTable);

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sh.tileSlot = atomicAdd(&args.counters->tileCursor, 1u);
        }
        __syncthreads();
        const uint32_t slot = sh.tileSlot;
        if (slot >= args.work.activeCount) {
            break;
        }
        const uint32_t tile = args.work.activeTiles[slot];
        const uint32_t listStart = args.work.tileStart[tile];
        const uint32_t listCount = args.work.tileCount[tile];
        const uint32_t T = args.grid.tilesPerAxis;
        const uint32_t tileOrigin[3] = {(tile % T) * kTileEdge, ((tile / T) % T) * kTileEdge,
                                        (tile / (T * T) + args.grid.slabTileZ0) * kTileEdge};
        const uint32_t px = tileOrigin[0] + lx, py = tileOrigin[1] + ly, pz = tileOrigin[2] + lz;

        VoxelAccumulator acc;
        acc.hasPartial = false;
        acc.hasVoxel = false;
        acc.partialTri = 0;
        acc.contributions = 0;
        acc.partial.w = acc.partial.u = acc.partial.v = 0.0f;
        acc.voxel.w = acc.voxel.r = acc.voxel.g = acc.voxel.b = 0.0f;

        for (uint32_t base = 0; base < listCount; base += kLeafBatch) {
            const uint32_t count = min(kLeafBatch, listCount - base);
            __syncthreads();
            if (tid < count) {
                stageLeaf<UV>(sh.stage[tid], args, args.work.tileList[listStart + base + tid], tileOrigin);
            }
            __syncthreads();

            for (uint32_t j = 0; j < count; ++j) {
                const LeafStage &s = sh.stage[j];
                if (acc.hasPartial && acc.partialTri != s.tri) {
                    flushPartial(acc, args);  // the list is ordered by triangle: this triangle's uv buffer is complete
                }
                const uint32_t box = s.box;
                const bool inside = lx >= (box & 15u) && ly >= ((box >> 4) & 15u) && lz >= ((box >> 8) & 15u) &&
                                    lx < ((box >> 12) & 15u) && ly < ((box >> 16) & 15u) && lz < ((box >> 20) & 15u);
                bool hit = inside && (!args.prefilter || prefilterPass(s, (float) lx, (float) ly, (float) lz));
                if (hit && (s.flags & kLeafNeedsCull) != 0) {
                    hit = !planeDistanceCulled(s.v, px, py, pz);  // voxelization.cpp:451-458, slivers only
                }
                if (!__any_sync(0xffffffffu, hit)) {
                    continue;  // warp-uniform: none of this warp's 32 voxels can touch the leaf
                }
                Tri<UV> leaf;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    leaf.v[k] = s.v[k];
                }
                if (UV) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        leaf.t[k] = s.t[k];
                    }
                }
                // warp-synchronous exact clip: the lanes that hit run the classify / split phases in lockstep
                const ClipResult r = clipLeafInVoxelWarp<UV>(hit, leaf, px, py, pz, s.area, sh.caseTable);
                if (hit) {
                    ++clipCalls;
                    if (r.pieces != 0) {
                        addContribution(acc, s.tri, r.weight, r.u, r.v);
                    }
                }
            }
        }
        flushPartial(acc, args);
        contributions += acc.contributions;

        // ---- output: optional 2x downscale (intended semantics), quantise, block-compact, store ----
        bool emit = acc.hasVoxel;
        int32_t ox = (int32_t) px, oy = (int32_t) py, oz = (int32_t) pz;
        WeightedColor result = acc.voxel;
        if (args.grid.supersampling == 2) {
            sh.dsW[tid] = acc.voxel.w;
            sh.dsR[tid] = acc.voxel.r;
            sh.dsG[tid] = acc.voxel.g;
            sh.dsB[tid] = acc.voxel.b;
            sh.dsHas[tid] = acc.hasVoxel ? 1 : 0;
            __syncthreads();
            emit = false;
            if (tid < 64) {
                const uint32_t qx = tid & 3u, qy = (tid >> 2) & 3u, qz = tid >> 4;
                // children in ascending Morton order (x is the most significant bit of the triple, ileave.hpp:243-246)
                for (uint32_t child = 0; child < 8; ++child) {
                    const uint32_t cx = 2 * qx + ((child >> 2) & 1u), cy = 2 * qy + ((child >> 1) & 1u),
                                   cz = 2 * qz + (child & 1u);
                    const uint32_t ci = cx + 8 * cy + 64 * cz;
                    if (!sh.dsHas[ci]) {
                        continue;
                    }
                    if (!emit) {
                        emit = true;
                        result.w = sh.dsW[ci];
                        result.r = sh.dsR[ci];
                        result.g = sh.dsG[ci];
                        result.b = sh.dsB[ci];
                    }
                    else {
                        combineColorInto(result, sh.dsW[ci], sh.dsR[ci], sh.dsG[ci], sh.dsB[ci],
                                         args.grid.strategy == kBlend);
                    }
                }
                ox = (int32_t) (tileOrigin[0] / 2 + qx);
                oy = (int32_t) (tileOrigin[1] / 2 + qy);
                oz = (int32_t) (tileOrigin[2] / 2 + qz);
            }
        }

        const unsigned int ballot = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) {
            sh.warpCounts[warp] = __popc(ballot);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t total = 0;
            for (int w = 0; w < kTileThreads / 32; ++w) {
                const uint32_t c = sh.warpCounts[w];
                sh.warpCounts[w] = total;
                total += c;
            }
            sh.outBase = total != 0 ? atomicAdd(&args.counters->voxels, (unsigned long long) total) : 0ull;
        }
        __syncthreads();
        if (emit) {
            const unsigned long long index = sh.outBase + sh.warpCounts[warp] + __popc(ballot & ((1u << lane) - 1u));
            if (index < args.outCapacity) {
                VoxelRecord rec;
                rec.x = ox;
                rec.y = oy;
                rec.z = oz;
                rec.argb = quantizeArgb(result.r, result.g, result.b);
                *reinterpret_cast<int4 *>(args.out + index) = *reinterpret_cast<const int4 *>(&rec);
                storeFloatRecord(args, index, result.w, result.r, result.g, result.b);
            }
            else {
                atomicAdd(&args.counters->outputOverflow, 1ull);
            }
        }
    }

    // per-block statistics
    for (int o = 16; o > 0; o >>= 1) {
        clipCalls += __shfl_xor_sync(0xffffffffu, clipCalls, o);
        contributions += __shfl_xor_sync(0xffffffffu, contributions, o);
    }
    if (lane == 0) {
        if (clipCalls != 0) {
            atomicAdd(&args.counters->clipCalls, clipCalls);
        }
        if (contributions != 0) {
            atomicAdd(&args.counters->contributions, contributions);
        }
    }
}

inline int gridFor(unsigned long long n, int threads, int cap)
{
    unsigned long long blocks = (n + threads - 1) / threads;
    if (blocks < 1) {
        blocks = 1;
    }
    if (blocks > (unsigned long long) cap) {
        blocks = cap;
    }
    return (int) blocks;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// launch wrappers

void launchBounds(const MeshView &mesh, RunCounters *counters, cudaStream_t stream)
{
    const unsigned long long vertices = mesh.count * 3;
    boundsKernel<<<gridFor(vertices, 256, 148 * 8), 256, 0, stream>>>(mesh.verts, vertices, counters);
}

/// The counters reach the host through a store from the SM into mapped pinned memory, not through a copy engine: a
/// 300-byte read-back must not queue behind a multi-hundred-megabyte download that another stream has in flight.
__global__ void publishCountersKernel(const RunCounters *__restrict__ counters, RunCounters *hostMapped)
{
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(counters);
    volatile unsigned long long *dst = reinterpret_cast<volatile unsigned long long *>(hostMapped);
    for (unsigned k = threadIdx.x; k < sizeof(RunCounters) / sizeof(unsigned long long); k += blockDim.x) {
        dst[k] = src[k];
    }
    __threadfence_system();
}

void launchPublishCounters(const RunCounters *counters, RunCounters *hostMapped, cudaStream_t stream)
{
    static_assert(sizeof(RunCounters) % sizeof(unsigned long long) == 0, "copied as 64-bit words");
    publishCountersKernel<<<1, 64, 0, stream>>>(counters, hostMapped);
}

/// Order-independent checksum of a record set: sum mod 2^64 of the splitmix64 finaliser of
/// x + (y << 21) + (z << 42) xor argb * 0x9E3779B97F4A7C15 (obj2voxel_b200/meshes.py record_hash is the host twin).
__global__ void __launch_bounds__(256)
recordHashKernel(const VoxelRecord *__restrict__ records, unsigned long long count, unsigned long long *sum)
{
    unsigned long long local = 0;
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const int4 r = __ldcs(reinterpret_cast<const int4 *>(records + i));
        unsigned long long k = (unsigned long long) (uint32_t) r.x + ((unsigned long long) (uint32_t) r.y << 21) +
                               ((unsigned long long) (uint32_t) r.z << 42);
        k ^= (unsigned long long) (uint32_t) r.w * 0x9E3779B97F4A7C15ull;
        k = (k ^ (k >> 30)) * 0xBF58476D1CE4E5B9ull;
        k = (k ^ (k >> 27)) * 0x94D049BB133111EBull;
        local += k ^ (k >> 31);
    }
    for (int o = 16; o > 0; o >>= 1) {
        local += __shfl_xor_sync(0xffffffffu, local, o);
    }
    if ((threadIdx.x & 31) == 0 && local != 0) {
        atomicAdd(sum, local);
    }
}

void launchRecordHash(const VoxelRecord *records, unsigned long long count, unsigned long long *sum, int smCount,
                      cudaStream_t stream)
{
    unsigned long long blocks = (count + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > (unsigned long long) smCount * 8 ? (unsigned long long) smCount * 8 : blocks);
    recordHashKernel<<<(unsigned) blocks, 256, 0, stream>>>(records, count, sum);
}

void launchFinishBounds(RunCounters *counters, cudaStream_t stream)
{
    finishBoundsKernel<<<1, 32, 0, stream>>>(counters);
}

void launchCountLeaves(const MeshView &mesh, const GridView &grid, uint32_t *leafCount, uint32_t *tileCount,
                       uint32_t *tileCandidates, RunCounters *counters, const HugeWork &work, cudaStream_t stream)
{
    const int blocks = gridFor(mesh.count, kSetupThreads, 148 * 64);
    if (mesh.uvs != nullptr) {
        countLeavesKernel<true><<<blocks, kSetupThreads, 0, stream>>>(mesh, grid, leafCount, tileCount, tileCandidates,
                                                                       counters, work);
    }
    else {
        countLeavesKernel<false><<<blocks, kSetupThreads, 0, stream>>>(mesh, grid, leafCount, tileCount, tileCandidates,
                                                                        counters, work);
    }
}

static int hugeBlocks(unsigned long long expected)
{
    const unsigned long long blocks = (expected * kHugeSubtrees + kHugeThreads - 1) / kHugeThreads;
    return (int) std::min<unsigned long long>(std::max<unsigned long long>(blocks, 1), 148ull * 32);
}

void launchHugeSubtreeScan(const MeshView &mesh, const GridView &grid, const HugeWork &work, RunCounters *counters,
                           uint32_t *perTriangle, bool occupancy, unsigned long long expected, cudaStream_t stream)
{
    hugeSubtreeCountKernel<<<hugeBlocks(expected), kHugeThreads, 0, stream>>>(mesh, grid, work, counters);
    const int blocks = (int) std::min<unsigned long long>(std::max<unsigned long long>(expected, 1), 148ull * 8);
    hugeScanKernel<<<blocks, kHugeSubtrees, 0, stream>>>(work, counters, perTriangle, occupancy);
}

void launchHugeCountTiles(const MeshView &mesh, const GridView &grid, const HugeWork &work, uint32_t *tileCount,
                          uint32_t *tileCandidates, RunCounters *counters, unsigned long long expected, cudaStream_t stream)
{
    hugeCountTilesKernel<<<hugeBlocks(expected), kHugeThreads, 0, stream>>>(mesh, grid, work, tileCount, tileCandidates,
                                                                            counters);
}

size_t scanScratchElems(size_t n)
{
    const size_t blocks = (n + kScanBlock - 1) / kScanBlock;
    return (blocks + 1) * 2;  // u64 block sums stored in a u32 scratch array
}

void launchExclusiveScan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *scratch, unsigned long long *total,
                         cudaStream_t stream)
{
    const size_t blocks = (n + kScanBlock - 1) / kScanBlock;
    auto *blockSums = reinterpret_cast<unsigned long long *>(scratch);
    if (blocks == 0) {
        cudaMemsetAsync(total, 0, sizeof(unsigned long long), stream);
        return;
    }
    scanReduceKernel<<<(unsigned) blocks, kScanThreads, 0, stream>>>(in, n, blockSums);
    scanTopKernel<<<1, 1024, 0, stream>>>(blockSums, blocks, total);
    scanApplyKernel<<<(unsigned) blocks, kScanThreads, 0, stream>>>(in, out, n, blockSums);
}

void launchCompactActiveTiles(const uint32_t *tileCount, const uint32_t *tileCandidates, const uint32_t *tileStart,
                              uint32_t tileTotal, uint32_t *allTiles, uint32_t *longTiles, uint32_t *heavyTiles,
                              LightTile *lightTiles, LightTile *bigLightTiles, RunCounters *counters,
                              cudaStream_t stream)
{
    if (tileTotal == 0) {
        return;
    }
    compactActiveTilesKernel<<<(tileTotal + 255) / 256, 256, 0, stream>>>(tileCount, tileCandidates, tileStart,
                                                                          tileTotal, allTiles, longTiles, heavyTiles,
                                                                          lightTiles, bigLightTiles, counters);
}

void launchEmitLeaves(const MeshView &mesh, const GridView &grid, const uint32_t *leafOffset, const uint32_t *tileStart,
                      uint32_t *tileFill, LeafRecord *leaves, LeafUv *leafUvs, uint32_t *tileList, uint32_t *pairTile,
                      RunCounters *counters, const HugeWork &work, unsigned long long hugeExpected, cudaStream_t stream)
{
    const int blocks = gridFor(mesh.count, kSetupThreads, 148 * 64);
    const EmitTargets out{leafOffset, tileStart, tileFill, leaves, leafUvs, tileList, pairTile};
    if (mesh.uvs != nullptr) {
        emitLeavesKernel<true><<<blocks, kSetupThreads, 0, stream>>>(mesh, grid, leafOffset, tileStart, tileFill, leaves,
                                                                      leafUvs, tileList, pairTile);
        if (work.capacity != 0) {
            hugeEmitLeavesKernel<true><<<hugeBlocks(hugeExpected), kHugeThreads, 0, stream>>>(mesh, grid, work, out,
                                                                                              counters);
        }
    }
    else {
        emitLeavesKernel<false><<<blocks, kSetupThreads, 0, stream>>>(mesh, grid, leafOffset, tileStart, tileFill,
                                                                       leaves, leafUvs, tileList, pairTile);
        if (work.capacity != 0) {
            hugeEmitLeavesKernel<false><<<hugeBlocks(hugeExpected), kHugeThreads, 0, stream>>>(mesh, grid, work, out,
                                                                                               counters);
        }
    }
}

void launchSortTileLists(const TileWork &work, uint32_t *tileList, cudaStream_t stream)
{
    if (work.allCount == 0) {
        return;
    }
    const int warpsPerBlock = 8;
    const int smallBlocks = gridFor(work.allCount, warpsPerBlock, 148 * 16);
    sortSmallListsKernel<<<smallBlocks, warpsPerBlock * 32, 0, stream>>>(work, tileList);
    if (work.longCount != 0) {
        const int largeBlocks = gridFor(work.longCount, 1, 148 * 4);
        sortLargeListsKernel<<<largeBlocks, 512, 0, stream>>>(work, tileList);
    }
}

void launchVoxelizeTiles(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    if (args.work.activeCount == 0) {
        return;
    }
    const size_t smem = sizeof(TileShared);
    unsigned blocks = (unsigned) smCount;
    if (blocks > args.work.activeCount) {
        blocks = args.work.activeCount;
    }
    if (args.mesh.uvs != nullptr) {
        cudaFuncSetAttribute(voxelizeTilesKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        voxelizeTilesKernel<true><<<blocks, kTileThreads, smem, stream>>>(args);
    }
    else {
        cudaFuncSetAttribute(voxelizeTilesKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        voxelizeTilesKernel<false><<<blocks, kTileThreads, smem, stream>>>(args);
    }
}

}  // namespace o2v
