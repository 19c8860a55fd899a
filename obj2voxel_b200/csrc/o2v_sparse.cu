// Staged sparse path: the voxelization of "light" tiles (few candidate voxels — the regime of every BASELINE config, where
// triangles are a few voxels across).  A tile-at-a-time kernel leaves most lanes idle there (a tile holds ~2 leaves, ~25
// candidate voxels, ~10 clips), so the work is re-batched between stages with stream compaction in HBM; every stage is
// dense over its own unit of work:
//
//   survivors (count, then write)  thread = (leaf, tile) pair   conservative SAT over the pair's candidate voxels
//   clip                           thread = surviving voxel     bit-exact six-plane clip (o2v_exact.cuh)
//   fold                           warp   = tile                sort contributions by (voxel, list position), replay the
//                                                               reference's fold order, write Voxel32 records
//
// Per-voxel order = ascending leaf index = (triangle index, DFS order) because tile lists are sorted before stage 1 and a
// tile's survivors are laid out contiguously in list order (pairOffset is an exclusive scan in tileList order).
// Reference semantics reproduced: src/voxelization.cpp:383-472 (clip, uv buffer), :513-526 (merge), util.hpp:160-172.
#include <stdlib.h>

#include "o2v_device.cuh"

namespace o2v {

namespace {

constexpr int kPairThreads = 128;
constexpr int kClipThreads = 128;
constexpr int kFoldWarpsPerBlock = 4;
constexpr int kBlockFoldThreads = 128;   // above that (up to kLightMaxCandidates) one block folds the tile
constexpr uint32_t kTinyFoldMax = 24;  // survivors per tile up to which one thread folds the whole tile
constexpr int kTinyFoldThreads = 128;

// ---------------------------------------------------------------------------------------------------------------------
// stages 1 + 2: thread per (leaf, tile) pair

/// ceil(65536 / d) for d = 1 .. 8: n / d == (n * kReciprocal16[d]) >> 16 for n < 64 (box positions of a masked pair).
__constant__ uint32_t kReciprocal16[9] = {0u, 65536u, 32768u, 21846u, 16384u, 13108u, 10923u, 9363u, 8192u};

/// Pass 1 (WRITE = false) evaluates the SAT for every candidate voxel of the pair, counts the survivors and, when the pair
/// has at most 64 candidates (always, for triangles up to 4 voxels across), keeps them as a bit mask; pass 2 (WRITE = true)
/// runs after the scan and only expands the mask into survivor entries (pairs with more candidates redo the SAT).
/// Both walk the box as ONE loop (position index, coordinates carried along): the lanes of a warp hold boxes of different
/// shapes, and nested z / y / x loops make the warp pay for the union of the shapes (measured: 8.7 of 32 lanes active).
template <bool WRITE>
__global__ void __launch_bounds__(kPairThreads)
sparseSurvivorsKernel(const VoxelizeArgs args)
{
    const SparseView &sp = args.sparse;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t pair = blockIdx.x * blockDim.x + threadIdx.x; pair < sp.pairCount; pair += stride) {
        const uint32_t tile = sp.pairTile[pair];
        uint32_t count = 0;
        if (sp.tileCandidates[tile] <= kLightMaxCandidates) {
            uint32_t box;
            LeafStage s;
            bool masked = false;
            unsigned long long mask = 0;
            if (WRITE) {
                box = sp.pairBox[pair];
                masked = (box >> 31) != 0;
                if (masked) {
                    mask = sp.pairMask[pair];
                }
            }
            if (!WRITE || !masked) {
                uint32_t origin[3];
                tileOriginOf(args.grid, tile, origin);
                stageLeaf<false>(s, args, args.work.tileList[pair], origin);
                box = s.box;
            }
            const uint32_t x0 = box & 15u, y0 = (box >> 4) & 15u, z0 = (box >> 8) & 15u;
            const uint32_t x1 = (box >> 12) & 15u, y1 = (box >> 16) & 15u, z1 = (box >> 20) & 15u;
            const uint32_t dx = x1 - x0, dy = y1 - y0;
            const uint32_t total = dx * dy * (z1 - z0);
            if (WRITE && masked) {
                uint32_t offset = sp.pairOffset[pair];
                const uint32_t rx = kReciprocal16[dx], ry = kReciprocal16[dy];
                while (mask != 0) {
                    const uint32_t bit = (uint32_t) __ffsll((long long) mask) - 1u;
                    mask &= mask - 1ull;
                    const uint32_t row = (bit * rx) >> 16, z = (row * ry) >> 16;
                    const uint32_t x = x0 + bit - row * dx, y = y0 + row - z * dy;
                    sp.entries[offset] = make_uint4(pair, x | (y << 3) | ((z0 + z) << 6), 0u, 0u);
                    ++offset;
                }
            }
            else {
                const bool small = total <= 64u;
                uint32_t offset = WRITE ? sp.pairOffset[pair] : 0u;
                uint32_t x = x0, y = y0, z = z0;
                for (uint32_t bit = 0; bit < total; ++bit) {
                    const bool pass = !args.prefilter || prefilterPass(s, (float) x, (float) y, (float) z);
                    if (pass) {
                        if (WRITE) {
                            sp.entries[offset] = make_uint4(pair, x | (y << 3) | (z << 6), 0u, 0u);
                            ++offset;
                        }
                        else if (small) {
                            mask |= 1ull << bit;
                        }
                        ++count;
                    }
                    if (++x == x1) {
                        x = x0;
                        if (++y == y1) {
                            y = y0;
                            ++z;
                        }
                    }
                }
                if (!WRITE) {
                    sp.pairBox[pair] = (box & 0x00ffffffu) | (small ? 0x80000000u : 0u);
                    if (small) {
                        sp.pairMask[pair] = mask;
                    }
                }
            }
        }
        if (!WRITE) {
            sp.pairSurvivors[pair] = count;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 3: thread per surviving candidate voxel -> exact clip

constexpr uint32_t kRefillThreshold = 8;  // idle lanes needed before the warp pays the load latency of a refill

/// Persistent lanes: each warp owns a contiguous range of survivors; a lane that finishes its clip stores the result and
/// fetches the next entry while the other lanes keep going (dynamic refill between the rounds of the warp-synchronous
/// clipper), so the long tail of one voxel's clip tree does not idle the other 31 lanes.
template <bool UV>
__global__ void __launch_bounds__(kClipThreads)
sparseClipKernel(const VoxelizeArgs args)
{
    const SparseView &sp = args.sparse;
    const unsigned int full = 0xffffffffu;
    const unsigned long long total = args.counters->survivors;
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
    const int refillThreshold = (int) kRefillThreshold;
    WarpClipper<UV> clipper;
    ClipStack<UV> stack;
    clipper.idle();
    clipper.r.pieces = 0;
    clipper.r.weight = clipper.r.u = clipper.r.v = 0.0f;
    bool hasEntry = false;
    unsigned long long current = 0;
    uint32_t currentTri = 0;

    for (;;) {
        const unsigned int idle = __ballot_sync(full, clipper.done);
        const bool moreWork = cursor < end;
        if (idle == full && !moreWork) {
            break;
        }
        if (moreWork && (__popc(idle) >= refillThreshold || idle == full)) {
            if (clipper.done) {
                if (hasEntry) {
                    // weight 0 = "no contribution" (area > 0, so a real contribution is never 0)
                    reinterpret_cast<uint2 *>(sp.entries + current)[1] =
                        make_uint2(__float_as_uint(clipper.r.pieces != 0 ? clipper.r.weight : 0.0f), currentTri);
                    if (UV) {
                        sp.uvs[current] = make_float2(clipper.r.u, clipper.r.v);
                    }
                    hasEntry = false;
                }
                const unsigned long long e = cursor + __popc(idle & below);
                if (e < end) {
                    const uint2 entry = reinterpret_cast<const uint2 *>(sp.entries + e)[0];
                    const uint32_t pair = entry.x;
                    const uint32_t leafIndex = __ldg(args.work.tileList + pair);
                    uint32_t origin[3];
                    tileOriginOf(args.grid, __ldg(sp.pairTile + pair), origin);
                    const float4 *src = reinterpret_cast<const float4 *>(args.leaves + leafIndex);
                    const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
                    Tri<UV> leaf;
                    leaf.v[0] = a.x; leaf.v[1] = a.y; leaf.v[2] = a.z; leaf.v[3] = a.w;
                    leaf.v[4] = b.x; leaf.v[5] = b.y; leaf.v[6] = b.z; leaf.v[7] = b.w;
                    leaf.v[8] = c.x;
                    if (UV) {
                        const float4 *uv = reinterpret_cast<const float4 *>(args.leafUvs + leafIndex);
                        const float4 u0 = __ldg(uv), u1 = __ldg(uv + 1);
                        leaf.t[0] = u0.x; leaf.t[1] = u0.y; leaf.t[2] = u0.z; leaf.t[3] = u0.w;
                        leaf.t[4] = u1.x; leaf.t[5] = u1.y;
                    }
                    clipper.begin(leaf, origin[0] + (entry.y & 7u), origin[1] + ((entry.y >> 3) & 7u),
                                  origin[2] + ((entry.y >> 6) & 7u), c.z);
                    if ((__float_as_uint(c.w) & kLeafNeedsCull) != 0 &&
                        planeDistanceCulled(leaf.v, clipper.px, clipper.py, clipper.pz)) {
                        clipper.done = true;  // voxelization.cpp:451-458 (slivers only): skipped, no contribution
                    }
                    hasEntry = true;
                    current = e;
                    currentTri = __float_as_uint(c.y);
                }
            }
            cursor += __popc(idle);
        }
        clipper.round(stack, caseTable);
    }
    if (hasEntry) {
        reinterpret_cast<uint2 *>(sp.entries + current)[1] =
            make_uint2(__float_as_uint(clipper.r.pieces != 0 ? clipper.r.weight : 0.0f), currentTri);
        if (UV) {
            sp.uvs[current] = make_float2(clipper.r.u, clipper.r.v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 4: warp per tile -> ordered fold + output

/// Per-warp scratch, carved out of dynamic shared memory: tri | sortKey | cW [| cU | cV], kWarpFoldMax entries each, then
/// the positions of the first contribution of every output voxel (16 bits each).
/// Sort key = (voxel key << 18) | (list slot << 9) | contribution slot — 9 bits each.
template <bool UV>
struct FoldWarpLayout {
    static constexpr uint32_t kArrays = UV ? 5u : 3u;
    static constexpr size_t kBytesPerWarp = (size_t) kArrays * kWarpFoldMax * 4u + (size_t) kWarpFoldMax * 2u;
};

template <bool UV>
__global__ void __launch_bounds__(kFoldWarpsPerBlock * 32)
sparseFoldKernel(const VoxelizeArgs args)
{
    extern __shared__ __align__(16) unsigned char foldSmem[];
    struct {
        uint32_t *tri, *sortKey;
        float *cW, *cU, *cV;
        uint16_t *heads;
    } sh;
    {
        uint32_t *warpBase = reinterpret_cast<uint32_t *>(foldSmem + (threadIdx.x >> 5) * FoldWarpLayout<UV>::kBytesPerWarp);
        sh.tri = warpBase;
        sh.sortKey = warpBase + kWarpFoldMax;
        sh.cW = reinterpret_cast<float *>(warpBase + 2 * kWarpFoldMax);
        sh.cU = UV ? sh.cW + kWarpFoldMax : sh.cW;
        sh.cV = UV ? sh.cW + 2 * kWarpFoldMax : sh.cW;
        sh.heads = reinterpret_cast<uint16_t *>(warpBase + FoldWarpLayout<UV>::kArrays * kWarpFoldMax);
    }
    const SparseView &sp = args.sparse;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t full = 0xffffffffu;
    const uint32_t warpsTotal = gridDim.x * kFoldWarpsPerBlock;
    const bool blend = args.grid.strategy == kBlend;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t groupShift = downscale ? 21u : 18u;  // group = parent voxel when downscaling, else the voxel
    unsigned long long contributions = 0;

    for (uint32_t t = blockIdx.x * kFoldWarpsPerBlock + (threadIdx.x >> 5); t < args.lightCount; t += warpsTotal) {
        const LightTile d = args.lightTiles[t];
        const uint32_t begin = sp.pairOffset[d.listStart];
        const uint32_t count = sp.pairOffset[d.listStart + d.leafCount] - begin;  // <= d.candidates <= 512
        if (count <= kTinyFoldMax) {
            continue;  // folded by sparseTinyFoldKernel (thread per tile)
        }
        uint32_t origin[3];
        tileOriginOf(args.grid, d.tile, origin);


        // ---- gather the tile's contributions (weight != 0), compacted in list order ----
        uint32_t kept = 0;
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t e = base + lane;
            uint4 entry = make_uint4(0u, 0u, 0u, 0u);
            if (e < count) {
                entry = sp.entries[begin + e];
            }
            const float w = __uint_as_float(entry.z);
            const bool keep = w != 0.0f;
            const uint32_t ballot = __ballot_sync(full, keep);
            if (keep) {
                const uint32_t pos = kept + __popc(ballot & below);
                const uint32_t x = entry.y & 7u, y = (entry.y >> 3) & 7u, z = (entry.y >> 6) & 7u;
                sh.sortKey[pos] = (voxelKey(x, y, z) << 18) | ((entry.x - d.listStart) << 9) | pos;
                sh.cW[pos] = w;
                sh.tri[pos] = entry.w;
                if (UV) {
                    const float2 uv = sp.uvs[begin + e];
                    sh.cU[pos] = uv.x;
                    sh.cV[pos] = uv.y;
                }
            }
            kept += __popc(ballot);
        }
        __syncwarp();
        if (kept == 0) {
            continue;
        }

        // ---- sort by (voxel key, list slot) ----
        if (kept <= 32) {
            uint32_t key = lane < kept ? sh.sortKey[lane] : 0xffffffffu;
            for (uint32_t k = 2; k <= 32; k <<= 1) {
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    const uint32_t other = __shfl_xor_sync(full, key, j);
                    const bool ascending = (lane & k) == 0;
                    const bool lower = (lane & j) == 0;
                    key = (lower == ascending) ? min(key, other) : max(key, other);
                }
            }
            sh.sortKey[lane] = key;
        }
        else if (kept <= 96) {
            // rank sort: a lane counts, for each of its (at most three) keys, the keys below it — one broadcast read per
            // key of the tile, no exchange steps (a padded bitonic network of 64 / 128 keys costs twice the instructions
            // and a warp barrier per stage); keys are unique (slot in the low bits), so the ranks are a permutation
            const uint32_t k0 = lane < kept ? sh.sortKey[lane] : 0xffffffffu;
            const uint32_t k1 = lane + 32 < kept ? sh.sortKey[lane + 32] : 0xffffffffu;
            const uint32_t k2 = lane + 64 < kept ? sh.sortKey[lane + 64] : 0xffffffffu;
            uint32_t r0 = 0, r1 = 0, r2 = 0;
            for (uint32_t j = 0; j < kept; ++j) {
                const uint32_t key = sh.sortKey[j];
                r0 += key < k0 ? 1u : 0u;
                r1 += key < k1 ? 1u : 0u;
                r2 += key < k2 ? 1u : 0u;
            }
            __syncwarp();
            if (lane < kept) {
                sh.sortKey[r0] = k0;
            }
            if (lane + 32 < kept) {
                sh.sortKey[r1] = k1;
            }
            if (lane + 64 < kept) {
                sh.sortKey[r2] = k2;
            }
        }
        else {
            uint32_t padded = 128u;
            while (padded < kept) {
                padded <<= 1;
            }
            for (uint32_t i = kept + lane; i < padded; i += 32) {
                sh.sortKey[i] = 0xffffffffu;
            }
            __syncwarp();
            for (uint32_t k = 2; k <= padded; k <<= 1) {  // ascending-only bitonic network
                for (uint32_t i = lane; i < padded; i += 32) {
                    const uint32_t l = i ^ (k - 1);
                    if (l > i) {
                        const uint32_t a = sh.sortKey[i], b = sh.sortKey[l];
                        if (a > b) {
                            sh.sortKey[i] = b;
                            sh.sortKey[l] = a;
                        }
                    }
                }
                __syncwarp();
                for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                    for (uint32_t i = lane; i < padded; i += 32) {
                        const uint32_t l = i ^ j;
                        if (l > i) {
                            const uint32_t a = sh.sortKey[i], b = sh.sortKey[l];
                            if (a > b) {
                                sh.sortKey[i] = b;
                                sh.sortKey[l] = a;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();

        // ---- one lane per output voxel replays the fold in order; one atomic per tile reserves the output range.  The
        // first contributions of the voxels are listed first, so that every lane of the fold loop has a voxel ----
        uint32_t runs = 0;
        for (uint32_t base = 0; base < kept; base += 32) {
            const uint32_t p = base + lane;
            const bool start = p < kept && (p == 0 || (sh.sortKey[p] >> groupShift) != (sh.sortKey[p - 1] >> groupShift));
            const uint32_t ballot = __ballot_sync(full, start);
            if (start) {
                sh.heads[runs + __popc(ballot & below)] = (uint16_t) p;
            }
            runs += __popc(ballot);
        }
        unsigned long long outBase = 0;
        if (lane == 0) {
            outBase = atomicAdd(&args.counters->voxels, (unsigned long long) runs);
        }
        outBase = __shfl_sync(full, outBase, 0);
        __syncwarp();
        for (uint32_t h = lane; h < runs; h += 32) {
            const uint32_t p = sh.heads[h];
            {
                const uint32_t group = sh.sortKey[p] >> groupShift;
                uint32_t currentVoxel = (sh.sortKey[p] >> 18) & 511u;
                VoxelAccumulator child;
                resetAccumulator(child);
                WeightedColor parent;
                parent.w = parent.r = parent.g = parent.b = 0.0f;
                bool hasParent = false;
                for (uint32_t q = p; q < kept; ++q) {
                    const uint32_t key = sh.sortKey[q];
                    if ((key >> groupShift) != group) {
                        break;
                    }
                    const uint32_t vk = (key >> 18) & 511u, slot = key & 511u;  // (bits 9 .. 17: the list slot, only a sort key)
                    if (vk != currentVoxel) {  // next child of the same parent (downscale only), ascending Morton order
                        flushPartial(child, args);
                        contributions += child.contributions;
                        if (!hasParent) {
                            hasParent = true;
                            parent = child.voxel;
                        }
                        else {
                            combineColorInto(parent, child.voxel.w, child.voxel.r, child.voxel.g, child.voxel.b, blend);
                        }
                        resetAccumulator(child);
                        currentVoxel = vk;
                    }
                    const uint32_t tri = sh.tri[slot];
                    if (child.hasPartial && child.partialTri != tri) {
                        flushPartial(child, args);  // the previous triangle's uv-buffer entry is complete
                    }
                    addContribution(child, tri, sh.cW[slot], UV ? sh.cU[slot] : 0.0f, UV ? sh.cV[slot] : 0.0f);
                }
                flushPartial(child, args);
                contributions += child.contributions;
                WeightedColor result = child.voxel;
                const uint32_t pk = currentVoxel >> 3, ck = currentVoxel & 7u;
                int32_t ox, oy, oz;
                if (downscale) {
                    if (hasParent) {
                        combineColorInto(parent, child.voxel.w, child.voxel.r, child.voxel.g, child.voxel.b, blend);
                        result = parent;
                    }
                    ox = (int32_t) (origin[0] / 2 + (pk & 3u));
                    oy = (int32_t) (origin[1] / 2 + ((pk >> 2) & 3u));
                    oz = (int32_t) (origin[2] / 2 + ((pk >> 4) & 3u));
                }
                else {
                    ox = (int32_t) (origin[0] + (((pk & 3u) << 1) | ((ck >> 2) & 1u)));
                    oy = (int32_t) (origin[1] + ((((pk >> 2) & 3u) << 1) | ((ck >> 1) & 1u)));
                    oz = (int32_t) (origin[2] + ((((pk >> 4) & 3u) << 1) | (ck & 1u)));
                }
                const unsigned long long index = outBase + h;
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
        __syncwarp();
    }

    for (int o = 16; o > 0; o >>= 1) {
        contributions += __shfl_xor_sync(full, contributions, o);
    }
    if (lane == 0 && contributions != 0) {
        atomicAdd(&args.counters->contributions, contributions);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 4, tiny tiles (<= kTinyFoldMax survivors, ~10 on BASELINE-sized triangles): a warp folds 32 tiles at a time.
// A warp per tile leaves most lanes idle, and a thread per tile (round 1) ran at 6.8 of 32 lanes: the tiles of a warp
// hold 1 .. 24 survivors and every lane ran its own load / insertion sort / fold loops.  Here the unit of work changes
// from stage to stage so that the lanes stay level: the survivors of the 32 tiles are loaded as one flat list (lane =
// entry), sorted inside their tile by rank (lane = entry: count the keys of the same tile that are smaller), and folded
// per output voxel (lane = voxel: ~2 contributions each); the warp reserves its output range with one atomic.

constexpr int kTinyWarps = kTinyFoldThreads / 32;
constexpr uint32_t kTinyBatchEntries = 32u * kTinyFoldMax;
static_assert(kTinyFoldMax <= 32u, "the entry of a tile takes 5 bits of the sort key");

template <bool UV>
__global__ void __launch_bounds__(kTinyFoldThreads)
sparseTinyFoldKernel(const VoxelizeArgs args)
{
    // key = (tile of the batch << 15) | (voxel key, 512 = no contribution << 5) | entry of the tile: ascending keys list a
    // tile's contributions by voxel (children of one parent adjacent, ascending Morton order) and, per voxel, in list order
    // (the entries of a tile lie in list order: the entry index orders like the list slot)
    __shared__ uint32_t keysShared[kTinyWarps][kTinyBatchEntries];
    __shared__ uint32_t sortedShared[kTinyWarps][kTinyBatchEntries];
    __shared__ uint32_t baseShared[kTinyWarps][33];
    __shared__ uint32_t beginShared[kTinyWarps][32];
    __shared__ uint32_t originShared[kTinyWarps][32][3];
    const SparseView &sp = args.sparse;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu;
    const uint32_t below = (1u << lane) - 1u;
    const bool blend = args.grid.strategy == kBlend;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t groupShift = downscale ? 8u : 5u;  // group = parent voxel when downscaling, else the voxel
    uint32_t *keys = keysShared[warp], *sorted = sortedShared[warp], *base = baseShared[warp];
    const uint32_t warpsTotal = gridDim.x * kTinyWarps;
    const uint32_t batches = (args.lightCount + 31u) / 32u;
    unsigned long long contributions = 0;

    for (uint32_t batch = blockIdx.x * kTinyWarps + warp; batch < batches; batch += warpsTotal) {
        // ---- lane = tile: where its survivors are ----
        const uint32_t t = batch * 32u + lane;
        uint32_t count = 0;
        if (t < args.lightCount) {
            const LightTile d = args.lightTiles[t];
            const uint32_t begin = sp.pairOffset[d.listStart];
            count = sp.pairOffset[d.listStart + d.leafCount] - begin;
            count = count <= kTinyFoldMax ? count : 0u;  // the others: sparseFoldKernel
            uint32_t origin[3];
            tileOriginOf(args.grid, d.tile, origin);
            beginShared[warp][lane] = begin;
            originShared[warp][lane][0] = origin[0];
            originShared[warp][lane][1] = origin[1];
            originShared[warp][lane][2] = origin[2];
        }
        uint32_t inclusive = count;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(full, inclusive, o);
            inclusive += lane >= (uint32_t) o ? up : 0u;
        }
        const uint32_t total = __shfl_sync(full, inclusive, 31);
        if (total == 0) {
            continue;  // warp-uniform
        }
        base[lane] = inclusive - count;
        if (lane == 31) {
            base[32] = total;
        }
        __syncwarp(full);

        // ---- lane = entry: load, build the key ----
        for (uint32_t i = lane; i < total; i += 32) {
            uint32_t tileOfBatch = 0;
            for (uint32_t step = 16; step > 0; step >>= 1) {
                tileOfBatch += base[tileOfBatch + step] <= i ? step : 0u;  // base[32] = total > i: never read past it
            }
            const uint32_t e = i - base[tileOfBatch];
            const uint4 entry = sp.entries[beginShared[warp][tileOfBatch] + e];
            const uint32_t x = entry.y & 7u, y = (entry.y >> 3) & 7u, z = (entry.y >> 6) & 7u;
            const uint32_t vk = __uint_as_float(entry.z) != 0.0f ? voxelKey(x, y, z) : 512u;
            keys[i] = (tileOfBatch << 15) | (vk << 5) | e;
        }
        __syncwarp(full);

        // ---- lane = entry: rank inside the tile ----
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t key = keys[i];
            const uint32_t b = base[key >> 15], n = base[(key >> 15) + 1] - b;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; ++j) {
                rank += keys[b + j] < key ? 1u : 0u;
            }
            sorted[b + rank] = key;
        }
        __syncwarp(full);

        // ---- the first contribution of every output voxel, as a dense list (the keys are no longer needed) ----
        uint32_t runs = 0;
        for (uint32_t p0 = 0; p0 < total; p0 += 32) {
            const uint32_t p = p0 + lane;
            const uint32_t key = p < total ? sorted[p] : 0xffffffffu;
            const bool start = p < total && ((key >> 5) & 1023u) < 512u &&
                               (p == 0 || (sorted[p - 1] >> groupShift) != (key >> groupShift));
            const uint32_t ballot = __ballot_sync(full, start);
            if (start) {
                keys[runs + __popc(ballot & below)] = p;
            }
            runs += __popc(ballot);
        }
        unsigned long long outBase = 0;
        if (lane == 0 && runs != 0) {
            outBase = atomicAdd(&args.counters->voxels, (unsigned long long) runs);
        }
        outBase = __shfl_sync(full, outBase, 0);
        __syncwarp(full);

        // ---- lane = output voxel: replay the fold in order ----
        for (uint32_t h = lane; h < runs; h += 32) {
            const uint32_t p = keys[h];
            const uint32_t first = sorted[p];
            {
                const uint32_t group = first >> groupShift;
                const uint32_t tileOfBatch = first >> 15;
                const uint32_t begin = beginShared[warp][tileOfBatch];
                const uint32_t tileEnd = base[tileOfBatch + 1];
                uint32_t currentVoxel = (first >> 5) & 511u;
                VoxelAccumulator child;
                resetAccumulator(child);
                WeightedColor parent;
                parent.w = parent.r = parent.g = parent.b = 0.0f;
                bool hasParent = false;
                for (uint32_t q = p; q < tileEnd; ++q) {
                    const uint32_t key = sorted[q];
                    if ((key >> groupShift) != group) {
                        break;
                    }
                    const uint32_t vk = (key >> 5) & 511u, e = key & 31u;
                    if (vk != currentVoxel) {  // next child of the same parent (downscale only), ascending Morton order
                        flushPartial(child, args);
                        contributions += child.contributions;
                        if (!hasParent) {
                            hasParent = true;
                            parent = child.voxel;
                        }
                        else {
                            combineColorInto(parent, child.voxel.w, child.voxel.r, child.voxel.g, child.voxel.b, blend);
                        }
                        resetAccumulator(child);
                        currentVoxel = vk;
                    }
                    const uint2 clipped = reinterpret_cast<const uint2 *>(sp.entries + begin + e)[1];  // {weight, triangle}
                    const uint32_t tri = clipped.y;
                    if (child.hasPartial && child.partialTri != tri) {
                        flushPartial(child, args);  // the previous triangle's uv-buffer entry is complete
                    }
                    float u = 0.0f, v = 0.0f;
                    if (UV) {
                        const float2 uv = sp.uvs[begin + e];
                        u = uv.x;
                        v = uv.y;
                    }
                    addContribution(child, tri, __uint_as_float(clipped.x), u, v);
                }
                flushPartial(child, args);
                contributions += child.contributions;
                WeightedColor result = child.voxel;
                const uint32_t pk = currentVoxel >> 3, ck = currentVoxel & 7u;
                const uint32_t originX = originShared[warp][tileOfBatch][0], originY = originShared[warp][tileOfBatch][1],
                               originZ = originShared[warp][tileOfBatch][2];
                int32_t ox, oy, oz;
                if (downscale) {
                    if (hasParent) {
                        combineColorInto(parent, child.voxel.w, child.voxel.r, child.voxel.g, child.voxel.b, blend);
                        result = parent;
                    }
                    ox = (int32_t) (originX / 2 + (pk & 3u));
                    oy = (int32_t) (originY / 2 + ((pk >> 2) & 3u));
                    oz = (int32_t) (originZ / 2 + ((pk >> 4) & 3u));
                }
                else {
                    ox = (int32_t) (originX + (((pk & 3u) << 1) | ((ck >> 2) & 1u)));
                    oy = (int32_t) (originY + ((((pk >> 2) & 3u) << 1) | ((ck >> 1) & 1u)));
                    oz = (int32_t) (originZ + ((((pk >> 4) & 3u) << 1) | (ck & 1u)));
                }
                const unsigned long long index = outBase + h;
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
        __syncwarp(full);
    }

    for (int o = 16; o > 0; o >>= 1) {
        contributions += __shfl_xor_sync(full, contributions, o);
    }
    if (lane == 0 && contributions != 0) {
        atomicAdd(&args.counters->contributions, contributions);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 4, large tiles (kWarpFoldMax < survivors <= kLightMaxCandidates): block per tile.  Same algorithm with 64-bit sort
// keys in shared memory ((voxel key << 40) | (list slot << 20) | entry index); payloads are read back from HBM/L2 by index.

template <bool UV>
__global__ void __launch_bounds__(kBlockFoldThreads)
sparseBlockFoldKernel(const VoxelizeArgs args)
{
    __shared__ unsigned long long keys[kLightMaxCandidates];
    __shared__ uint32_t keptShared, runsShared, emittedShared;
    __shared__ unsigned long long outBaseShared;
    const SparseView &sp = args.sparse;
    const bool blend = args.grid.strategy == kBlend;
    const bool downscale = args.grid.supersampling == 2;
    const uint32_t groupShift = downscale ? 43u : 40u;
    unsigned long long contributions = 0;

    for (uint32_t t = blockIdx.x; t < args.bigLightCount; t += gridDim.x) {
        const LightTile d = args.bigLightTiles[t];
        const uint32_t begin = sp.pairOffset[d.listStart];
        const uint32_t count = sp.pairOffset[d.listStart + d.leafCount] - begin;  // <= d.candidates <= 4096
        if (count == 0) {
            continue;  // block-uniform
        }
        uint32_t origin[3];
        tileOriginOf(args.grid, d.tile, origin);
        __syncthreads();
        if (threadIdx.x == 0) {
            keptShared = 0;
            runsShared = 0;
            emittedShared = 0;
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < count; e += blockDim.x) {
            const uint4 entry = sp.entries[begin + e];
            if (__uint_as_float(entry.z) != 0.0f) {
                const uint32_t x = entry.y & 7u, y = (entry.y >> 3) & 7u, z = (entry.y >> 6) & 7u;
                const uint32_t pos = atomicAdd(&keptShared, 1u);  // order is irrelevant before the sort
                keys[pos] = ((unsigned long long) voxelKey(x, y, z) << 40) |
                            ((unsigned long long) (entry.x - d.listStart) << 20) | e;
            }
        }
        __syncthreads();
        const uint32_t kept = keptShared;
        if (kept == 0) {
            continue;
        }
        uint32_t padded = 1;
        while (padded < kept) {
            padded <<= 1;
        }
        for (uint32_t i = kept + threadIdx.x; i < padded; i += blockDim.x) {
            keys[i] = ~0ull;
        }
        __syncthreads();
        for (uint32_t k = 2; k <= padded; k <<= 1) {  // ascending-only bitonic network
            for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x) {
                const uint32_t l = i ^ (k - 1);
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    if (a > b) {
                        keys[i] = b;
                        keys[l] = a;
                    }
                }
            }
            __syncthreads();
            for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x) {
                    const uint32_t l = i ^ j;
                    if (l > i) {
                        const unsigned long long a = keys[i], b = keys[l];
                        if (a > b) {
                            keys[i] = b;
                            keys[l] = a;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // reserve the output range: one atomic per tile
        uint32_t myRuns = 0;
        for (uint32_t p = threadIdx.x; p < kept; p += blockDim.x) {
            myRuns += (p == 0 || (keys[p] >> groupShift) != (keys[p - 1] >> groupShift)) ? 1u : 0u;
        }
        if (myRuns != 0) {
            atomicAdd(&runsShared, myRuns);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            outBaseShared = atomicAdd(&args.counters->voxels, (unsigned long long) runsShared);
        }
        __syncthreads();
        for (uint32_t p = threadIdx.x; p < kept; p += blockDim.x) {
            if (!(p == 0 || (keys[p] >> groupShift) != (keys[p - 1] >> groupShift))) {
                continue;
            }
            const unsigned long long group = keys[p] >> groupShift;
            uint32_t currentVoxel = (uint32_t) (keys[p] >> 40) & 511u;
            VoxelAccumulator child;
            resetAccumulator(child);
            WeightedColor parent;
            parent.w = parent.r = parent.g = parent.b = 0.0f;
            bool hasParent = false;
            for (uint32_t q = p; q < kept; ++q) {
                const unsigned long long key = keys[q];
                if ((key >> groupShift) != group) {
                    break;
                }
                const uint32_t vk = (uint32_t) (key >> 40) & 511u, e = (uint32_t) key & 0xfffffu;  // (bits 20 .. 39: list slot)
                if (vk != currentVoxel) {  // next child of the same parent (downscale only), ascending Morton order
                    flushPartial(child, args);
                    contributions += child.contributions;
                    if (!hasParent) {
                        hasParent = true;
                        parent = child.voxel;
                    }
                    else {
                        combineColorInto(parent, child.voxel.w, child.voxel.r, child.voxel.g, child.voxel.b, blend);
                    }
                    resetAccumulator(child);
                    currentVoxel = vk;
                }
                const uint2 clipped = reinterpret_cast<const uint2 *>(sp.entries + begin + e)[1];  // {weight, triangle}
                const uint32_t tri = clipped.y;
                if (child.hasPartial && child.partialTri != tri) {
                    flushPartial(child, args);
                }
                float u = 0.0f, v = 0.0f;
                if (UV) {
                    const float2 uv = sp.uvs[begin + e];
                    u = uv.x;
                    v = uv.y;
                }
                addContribution(child, tri, __uint_as_float(clipped.x), u, v);
            }
            flushPartial(child, args);
            contributions += child.contributions;
            WeightedColor result = child.voxel;
            const uint32_t pk = currentVoxel >> 3, ck = currentVoxel & 7u;
            int32_t ox, oy, oz;
            if (downscale) {
                if (hasParent) {
                    combineColorInto(parent, child.voxel.w, child.voxel.r, child.voxel.g, child.voxel.b, blend);
                    result = parent;
                }
                ox = (int32_t) (origin[0] / 2 + (pk & 3u));
                oy = (int32_t) (origin[1] / 2 + ((pk >> 2) & 3u));
                oz = (int32_t) (origin[2] / 2 + ((pk >> 4) & 3u));
            }
            else {
                ox = (int32_t) (origin[0] + (((pk & 3u) << 1) | ((ck >> 2) & 1u)));
                oy = (int32_t) (origin[1] + ((((pk >> 2) & 3u) << 1) | ((ck >> 1) & 1u)));
                oz = (int32_t) (origin[2] + ((((pk >> 4) & 3u) << 1) | (ck & 1u)));
            }
            const unsigned long long index = outBaseShared + atomicAdd(&emittedShared, 1u);
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

    for (int o = 16; o > 0; o >>= 1) {
        contributions += __shfl_xor_sync(0xffffffffu, contributions, o);
    }
    if ((threadIdx.x & 31) == 0 && contributions != 0) {
        atomicAdd(&args.counters->contributions, contributions);
    }
}

template <typename Kernel>
unsigned persistentBlocks(Kernel kernel, int threads, int smCount, unsigned long long needed, size_t dynamicSmem = 0)
{
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, dynamicSmem);
    perSm = perSm < 1 ? 1 : perSm;
    unsigned long long blocks = (unsigned long long) smCount * perSm;  // a multiple of the SM count
    blocks = blocks < needed ? blocks : needed;
    return (unsigned) (blocks < 1 ? 1 : blocks);
}

}  // namespace

void launchSparseSurvivors(const VoxelizeArgs &args, bool write, cudaStream_t stream)
{
    if (args.sparse.pairCount == 0) {
        return;
    }
    const unsigned blocks = (args.sparse.pairCount + kPairThreads - 1) / kPairThreads;
    if (write) {
        sparseSurvivorsKernel<true><<<blocks, kPairThreads, 0, stream>>>(args);
    }
    else {
        sparseSurvivorsKernel<false><<<blocks, kPairThreads, 0, stream>>>(args);
    }
}

void launchSparseClip(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    // the survivor count lives on the device (RunCounters::survivors): persistent grid-stride kernel
    if (args.mesh.uvs != nullptr) {
        const unsigned blocks = persistentBlocks(sparseClipKernel<true>, kClipThreads, smCount, ~0ull);
        sparseClipKernel<true><<<blocks, kClipThreads, 0, stream>>>(args);
    }
    else {
        const unsigned blocks = persistentBlocks(sparseClipKernel<false>, kClipThreads, smCount, ~0ull);
        sparseClipKernel<false><<<blocks, kClipThreads, 0, stream>>>(args);
    }
}

void launchSparseFold(const VoxelizeArgs &args, int smCount, cudaStream_t stream)
{
    {
        const unsigned long long tinyBlocks = (args.lightCount + kTinyFoldThreads - 1) / kTinyFoldThreads;
        if (args.mesh.uvs != nullptr) {
            sparseTinyFoldKernel<true><<<persistentBlocks(sparseTinyFoldKernel<true>, kTinyFoldThreads, smCount, tinyBlocks),
                                         kTinyFoldThreads, 0, stream>>>(args);
        }
        else {
            sparseTinyFoldKernel<false><<<persistentBlocks(sparseTinyFoldKernel<false>, kTinyFoldThreads, smCount,
                                                           tinyBlocks),
                                          kTinyFoldThreads, 0, stream>>>(args);
        }
    }
    if (args.bigLightCount != 0) {
        if (args.mesh.uvs != nullptr) {
            sparseBlockFoldKernel<true><<<persistentBlocks(sparseBlockFoldKernel<true>, kBlockFoldThreads, smCount,
                                                           args.bigLightCount),
                                          kBlockFoldThreads, 0, stream>>>(args);
        }
        else {
            sparseBlockFoldKernel<false><<<persistentBlocks(sparseBlockFoldKernel<false>, kBlockFoldThreads, smCount,
                                                            args.bigLightCount),
                                           kBlockFoldThreads, 0, stream>>>(args);
        }
    }
    if (args.lightCount == 0) {
        return;
    }
    const int threads = kFoldWarpsPerBlock * 32;
    const unsigned long long needed = (args.lightCount + kFoldWarpsPerBlock - 1) / kFoldWarpsPerBlock;
    if (args.mesh.uvs != nullptr) {
        const size_t smem = FoldWarpLayout<true>::kBytesPerWarp * kFoldWarpsPerBlock;
        cudaFuncSetAttribute(sparseFoldKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        sparseFoldKernel<true><<<persistentBlocks(sparseFoldKernel<true>, threads, smCount, needed, smem), threads, smem,
                                 stream>>>(args);
    }
    else {
        const size_t smem = FoldWarpLayout<false>::kBytesPerWarp * kFoldWarpsPerBlock;
        cudaFuncSetAttribute(sparseFoldKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        sparseFoldKernel<false><<<persistentBlocks(sparseFoldKernel<false>, threads, smCount, needed, smem), threads,
                                  smem, stream>>>(args);
    }
}

}  // namespace o2v
