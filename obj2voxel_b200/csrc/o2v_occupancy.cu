// Occupancy-only path (see OccupancyView in o2v_kernels.cuh).
//
// A mesh whose every triangle is MATERIALLESS (any STL, any OBJ without materials, obj2voxel_set_triangle_basic /
// _colored — SURVEY fact 8) voxelizes to 0xFFFFFFFF wherever a voxel is occupied: colorAt_f returns white
// (src/triangle.hpp:186), BLEND of equal colours is (w1 + w2) / (w1 + w2) = 1 exactly, MAX keeps one of two whites, the
// 2x downscale combines whites.  The weights — the only thing the fold order and the piece count feed — never reach the
// output, so the result is the set { voxel : some leaf's exact clip has >= 1 piece }, an order-independent OR:
//
//   classify  thread = candidate voxel     three-way SAT (o2v_sat.cuh): `certain` -> atomicOr into the tile's 512-bit
//                                          bitmap, `uncertain` -> queue entry (unless the bitmap already decides it)
//   clip      persistent lanes             the bit-exact six-plane clip (WarpClipper) for queued voxels only
//   expand    thread = tile                bitmap (OR-reduced 2x2x2 when supersampling) -> compacted Voxel32 records
//
// Exactness: `miss` and `certain` are proofs about the reference's result (o2v_sat.cuh header; fuzzed by
// tests/test_sat_classifier.py), everything else runs the reference arithmetic.  Tiles above kLightMaxCandidates keep
// the block-per-tile kernel.  prefilter = 0 sends every candidate through the exact clip (validation).
#include "o2v_device.cuh"

namespace o2v {

namespace {

constexpr int kOccPairThreads = 128;
constexpr int kOccClipThreads = 128;
constexpr int kOccExpandThreads = 128;
constexpr int kOccRefillThreshold = 8;

/// Bits of `m` (layout x + 8 y) smeared over their 2x2 xy blocks: the footprint of the parents that already have a child.
__device__ __forceinline__ unsigned long long smear2x2(unsigned long long m)
{
    const unsigned long long evenX = (m | (m >> 1)) & 0x5555555555555555ull;
    const unsigned long long pairX = evenX | (evenX << 1);
    const unsigned long long evenY = (pairX | (pairX >> 8)) & 0x00ff00ff00ff00ffull;
    return evenY | (evenY << 8);
}

/// Per-pair SAT constants staged in shared memory.  43 words: an odd stride, so the staging threads (one pair each) write
/// without bank conflicts; in the flat phase the lanes of a warp read the same one or two pairs (broadcast).
struct PairSat {
    float plane[4];
    float planeLimit, planeSure;
    float edge[27];
    float lo[3], hi[3];
    uint32_t flags;
    uint32_t box;               // tile-local AABB as in LeafStage::box; 0 = pair not on this path
    uint32_t magicX, magicXY;   // n / d == (n * magic) >> 16 for n < 512, d = dx resp. dx * dy (<= 64)
};

constexpr int kOccAccWords = 33;  // per pair: sure[8][2] | maybe[8][2] 32-bit halves of the layer masks, +1 pad word

/// One block = kOccPairThreads consecutive (leaf, tile) pairs.
///   stage    thread = pair        leaf -> SAT constants in shared memory; block-wide scan of the pairs' candidate counts
///   classify thread = candidate   the batch's candidate voxels as one flat index space (perfectly balanced: a pair has
///                                 1 .. 512 candidates), verdict bits OR-ed into the pair's layer masks in shared memory
///   flush    thread = pair        `certain` masks -> global tile bitmap (one 64-bit atomicOr per layer), `uncertain` ones
///                                 filtered by what the bitmap already shows, then appended to the queue (one global
///                                 atomic per block)
__global__ void __launch_bounds__(kOccPairThreads)
occupancyClassifyKernel(const VoxelizeArgs args)
{
    __shared__ PairSat sat[kOccPairThreads];
    __shared__ uint32_t acc[kOccPairThreads * kOccAccWords];
    __shared__ uint32_t prefix[kOccPairThreads + 1];
    __shared__ uint32_t warpSums[kOccPairThreads / 32];
    __shared__ unsigned long long queueBase;

    const SparseView &sp = args.sparse;
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t pair = blockIdx.x * kOccPairThreads + tid;

    // ---- stage ----
    uint32_t volume = 0, tile = 0;
    {
        PairSat &mine = sat[tid];
        mine.box = 0;
        bool active = pair < sp.pairCount;
        if (active) {
            tile = sp.pairTile[pair];
            active = sp.tileCandidates[tile] <= kLightMaxCandidates;  // heavy tiles: block-per-tile kernel
        }
        if (active) {
            uint32_t origin[3];
            tileOriginOf(args.grid, tile, origin);
            LeafStage s;
            stageLeaf<false>(s, args, args.work.tileList[pair], origin);
            const float originF[3] = {(float) origin[0], (float) origin[1], (float) origin[2]};
            LeafCertain c;
            buildCertain(c, s, originF);
            const uint32_t dx = ((s.box >> 12) & 15u) - (s.box & 15u), dy = ((s.box >> 16) & 15u) - ((s.box >> 4) & 15u),
                           dz = ((s.box >> 20) & 15u) - ((s.box >> 8) & 15u);
            volume = dx * dy * dz;
            if (volume != 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mine.plane[k] = s.plane[k];
                }
                mine.planeLimit = s.planeLimit;
                mine.planeSure = c.planeSure;
#pragma unroll
                for (int k = 0; k < 27; ++k) {
                    mine.edge[k] = s.edge[k];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    mine.lo[k] = c.lo[k];
                    mine.hi[k] = c.hi[k];
                }
                mine.flags = s.flags;
                mine.box = s.box;
                mine.magicX = (65536u + dx - 1u) / dx;
                mine.magicXY = (65536u + dx * dy - 1u) / (dx * dy);
            }
        }
#pragma unroll
        for (int k = 0; k < kOccAccWords - 1; ++k) {
            acc[tid * kOccAccWords + k] = 0;
        }
    }
    uint32_t inclusive = volume;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(full, inclusive, o);
        inclusive += lane >= (uint32_t) o ? up : 0u;
    }
    if (lane == 31) {
        warpSums[warp] = inclusive;
    }
    __syncthreads();
    uint32_t warpBase = 0;
    for (uint32_t w = 0; w < warp; ++w) {
        warpBase += warpSums[w];
    }
    prefix[tid + 1] = warpBase + inclusive;
    if (tid == 0) {
        prefix[0] = 0;
    }
    __syncthreads();
    const uint32_t total = prefix[kOccPairThreads];

    // ---- classify ----
    for (uint32_t i = tid; i < total; i += kOccPairThreads) {
        uint32_t p = 0;  // last pair with prefix[p] <= i
#pragma unroll
        for (uint32_t step = kOccPairThreads / 2; step > 0; step >>= 1) {
            p += prefix[p + step] <= i ? step : 0u;
        }
        const PairSat &s = sat[p];
        const uint32_t local = i - prefix[p];
        const uint32_t x0 = s.box & 15u, y0 = (s.box >> 4) & 15u, z0 = (s.box >> 8) & 15u;
        const uint32_t dx = ((s.box >> 12) & 15u) - x0, dy = ((s.box >> 16) & 15u) - y0;
        const uint32_t zi = (local * s.magicXY) >> 16;
        const uint32_t inLayer = local - zi * dx * dy;
        const uint32_t yi = (inLayer * s.magicX) >> 16;
        const uint32_t x = x0 + (inLayer - yi * dx), y = y0 + yi, z = z0 + zi;
        const int verdict = args.prefilter ? classifyVoxel(s, s, (float) x, (float) y, (float) z) : kSatUncertain;
        if (verdict != kSatMiss) {
            // word = layer z, half y / 4; bit = x + 8 (y % 4)
            atomicOr(&acc[p * kOccAccWords + (verdict == kSatCertain ? 0u : 16u) + z * 2u + (y >> 2)],
                     1u << (x + 8u * (y & 3u)));
        }
    }
    __syncthreads();

    // ---- flush ----
    uint32_t count = 0;
    if (volume != 0) {
        unsigned long long *bits = occ.tileBits + (size_t) occ.tileSlot[tile] * kTileEdge;
        uint32_t *mine = acc + tid * kOccAccWords;
        const uint32_t z0 = (sat[tid].box >> 8) & 15u, z1 = (sat[tid].box >> 20) & 15u;
        for (uint32_t z = z0; z < z1; ++z) {
            const unsigned long long sure = mine[z * 2] | ((unsigned long long) mine[z * 2 + 1] << 32);
            unsigned long long maybe = mine[16 + z * 2] | ((unsigned long long) mine[16 + z * 2 + 1] << 32);
            unsigned long long known = sure;
            if (sure != 0) {
                known |= atomicOr(bits + z, sure);
            }
            else if (maybe != 0) {
                known = __ldcg(bits + z);
            }
            if (downscale && maybe != 0) {
                // a parent that already has a child needs no further children
                known = smear2x2(known | __ldcg(bits + (z ^ 1u)));
            }
            maybe &= ~known;  // already decided by this or another leaf (a stale read only costs a redundant clip)
            mine[16 + z * 2] = (uint32_t) maybe;
            mine[16 + z * 2 + 1] = (uint32_t) (maybe >> 32);
            count += (uint32_t) __popcll(maybe);
        }
    }
    inclusive = count;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(full, inclusive, o);
        inclusive += lane >= (uint32_t) o ? up : 0u;
    }
    __syncthreads();  // warpSums is reused
    if (lane == 31) {
        warpSums[warp] = inclusive;
    }
    __syncthreads();
    uint32_t blockTotal = 0;
    warpBase = 0;
    for (uint32_t w = 0; w < kOccPairThreads / 32; ++w) {
        warpBase += w < warp ? warpSums[w] : 0u;
        blockTotal += warpSums[w];
    }
    if (blockTotal == 0) {
        return;  // block-uniform
    }
    if (tid == 0) {
        queueBase = atomicAdd(&args.counters->survivors, (unsigned long long) blockTotal);
    }
    __syncthreads();
    if (count != 0) {
        unsigned long long index = queueBase + warpBase + (inclusive - count);
        const uint32_t *mine = acc + tid * kOccAccWords;
        const uint32_t z0 = (sat[tid].box >> 8) & 15u, z1 = (sat[tid].box >> 20) & 15u;
        for (uint32_t z = z0; z < z1; ++z) {
            unsigned long long maybe = mine[16 + z * 2] | ((unsigned long long) mine[16 + z * 2 + 1] << 32);
            while (maybe != 0) {
                const uint32_t b = (uint32_t) __ffsll((long long) maybe) - 1u;
                maybe &= maybe - 1ull;
                if (index < occ.queueCapacity) {  // beyond: counted only; the engine grows the queue and reruns
                    occ.queue[index] = make_uint2(pair, (z << 6) | b);
                }
                ++index;
            }
        }
    }
}

/// Same persistent-lane scheme as sparseClipKernel (o2v_sparse.cu), fed from the queue; a surviving piece sets the bit.
__global__ void __launch_bounds__(kOccClipThreads)
occupancyClipKernel(const VoxelizeArgs args)
{
    const SparseView &sp = args.sparse;
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
                    const uint2 entry = occ.queue[e];
                    const uint32_t pair = entry.x;
                    const uint32_t tile = __ldg(sp.pairTile + pair);
                    const uint32_t z = entry.y >> 6, xy = entry.y & 63u;
                    unsigned long long *layer = occ.tileBits + (size_t) __ldg(occ.tileSlot + tile) * kTileEdge + z;
                    const unsigned long long mine = 1ull << xy;
                    unsigned long long known = __ldcg(layer);
                    if (downscale) {
                        known = smear2x2(known | __ldcg(occ.tileBits + (size_t) __ldg(occ.tileSlot + tile) * kTileEdge +
                                                        (z ^ 1u)));
                    }
                    if ((known & mine) == 0) {
                        const uint32_t leafIndex = __ldg(args.work.tileList + pair);
                        uint32_t origin[3];
                        tileOriginOf(args.grid, tile, origin);
                        const float4 *src = reinterpret_cast<const float4 *>(args.leaves + leafIndex);
                        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
                        Tri<false> leaf;
                        leaf.v[0] = a.x; leaf.v[1] = a.y; leaf.v[2] = a.z; leaf.v[3] = a.w;
                        leaf.v[4] = b.x; leaf.v[5] = b.y; leaf.v[6] = b.z; leaf.v[7] = b.w;
                        leaf.v[8] = c.x;
                        clipper.begin(leaf, origin[0] + (xy & 7u), origin[1] + (xy >> 3), origin[2] + z, c.z);
                        if ((__float_as_uint(c.w) & kLeafNeedsCull) != 0 &&
                            planeDistanceCulled(leaf.v, clipper.px, clipper.py, clipper.pz)) {
                            clipper.done = true;  // voxelization.cpp:451-458 (slivers only)
                        }
                        hasEntry = true;
                        word = layer;
                        bit = mine;
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

__global__ void __launch_bounds__(kOccExpandThreads)
occupancyExpandKernel(const VoxelizeArgs args)
{
    const OccupancyView &occ = args.occ;
    const unsigned int full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t stride = gridDim.x * blockDim.x;

    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x - lane); base < args.work.allCount; base += stride) {
        const uint32_t slot = base + lane;
        unsigned long long m[kTileEdge];
#pragma unroll
        for (int z = 0; z < (int) kTileEdge; ++z) {
            m[z] = 0;
        }
        uint32_t origin[3] = {0, 0, 0};
        if (slot < args.work.allCount) {
            const uint32_t tile = args.work.allTiles[slot];
            if (args.sparse.tileCandidates[tile] <= kLightMaxCandidates) {
                tileOriginOf(args.grid, tile, origin);
                const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(occ.tileBits + (size_t) slot * kTileEdge);
#pragma unroll
                for (int k = 0; k < (int) kTileEdge / 2; ++k) {
                    const ulonglong2 w = src[k];
                    m[2 * k] = w.x;
                    m[2 * k + 1] = w.y;
                }
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
            origin[0] >>= 1;
            origin[1] >>= 1;
            origin[2] >>= 1;
        }
        uint32_t count = 0;
#pragma unroll
        for (int z = 0; z < (int) kTileEdge; ++z) {
            count += (uint32_t) __popcll(m[z]);
        }
        uint32_t inclusive = count;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(full, inclusive, o);
            inclusive += lane >= (uint32_t) o ? up : 0u;
        }
        const uint32_t warpCount = __shfl_sync(full, inclusive, 31);
        if (warpCount == 0) {
            continue;
        }
        unsigned long long index = 0;
        if (lane == 31) {
            index = atomicAdd(&args.counters->voxels, (unsigned long long) warpCount);
        }
        index = __shfl_sync(full, index, 31) + (inclusive - count);
        unsigned long long overflow = 0;
        const uint32_t shift = downscale ? 1u : 0u;
#pragma unroll
        for (int z = 0; z < (int) kTileEdge; ++z) {
            unsigned long long w = m[z];
            while (w != 0) {
                const uint32_t b = (uint32_t) __ffsll((long long) w) - 1u;
                w &= w - 1ull;
                if (index < args.outCapacity) {
                    VoxelRecord rec;
                    rec.x = (int32_t) (origin[0] + ((b & 7u) >> shift));
                    rec.y = (int32_t) (origin[1] + ((b >> 3) >> shift));
                    rec.z = (int32_t) (origin[2] + (uint32_t) z);
                    rec.argb = 0xFFFFFFFFu;  // quantizeArgb(1, 1, 1)
                    *reinterpret_cast<int4 *>(args.out + index) = *reinterpret_cast<const int4 *>(&rec);
                }
                else {
                    ++overflow;
                }
                ++index;
            }
        }
        if (overflow != 0) {
            atomicAdd(&args.counters->outputOverflow, overflow);
        }
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

}  // namespace

void launchOccupancyClassify(const VoxelizeArgs &args, cudaStream_t stream)
{
    if (args.sparse.pairCount == 0) {
        return;
    }
    const unsigned blocks = (args.sparse.pairCount + kOccPairThreads - 1) / kOccPairThreads;
    occupancyClassifyKernel<<<blocks, kOccPairThreads, 0, stream>>>(args);
}

void launchOccupancyClip(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    // the queue length lives on the device (RunCounters::survivors)
    occupancyClipKernel<<<occupancyPersistentBlocks(occupancyClipKernel, kOccClipThreads, smCount), kOccClipThreads, 0,
                          stream>>>(args);
}

void launchOccupancyExpand(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    if (args.work.allCount == 0) {
        return;
    }
    unsigned blocks = occupancyPersistentBlocks(occupancyExpandKernel, kOccExpandThreads, smCount);
    const unsigned needed = (args.work.allCount + kOccExpandThreads - 1) / kOccExpandThreads;
    blocks = blocks < needed ? blocks : needed;
    occupancyExpandKernel<<<blocks, kOccExpandThreads, 0, stream>>>(args);
}

}  // namespace o2v
