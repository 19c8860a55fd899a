// Device-side helpers shared by the tile kernels (o2v_kernels.cu: block-per-tile; o2v_sparse.cu: staged sparse path).
#ifndef O2V_DEVICE_CUH
#define O2V_DEVICE_CUH

#include "o2v_kernels.cuh"
#include "o2v_sat.cuh"

namespace o2v {
namespace {

/// Adds a per-thread tally to a global counter with one atomic per warp (all 32 lanes must call it).
__device__ __forceinline__ void warpTally(unsigned long long *counter, unsigned long long value)
{
    for (int o = 16; o > 0; o >>= 1) {
        value += __shfl_xor_sync(0xffffffffu, value, o);
    }
    if ((threadIdx.x & 31) == 0 && value != 0) {
        atomicAdd(counter, value);
    }
}

__device__ __forceinline__ void tileOriginOf(const GridView &grid, uint32_t tile, uint32_t origin[3])
{
    const uint32_t T = grid.tilesPerAxis;
    if (grid.tileShift < 32u) {  // uniform; the three divisions below are ~60 instructions
        origin[0] = (tile & (T - 1u)) * kTileEdge;
        origin[1] = ((tile >> grid.tileShift) & (T - 1u)) * kTileEdge;
        origin[2] = ((tile >> (2u * grid.tileShift)) + grid.slabTileZ0) * kTileEdge;
        return;
    }
    origin[0] = (tile % T) * kTileEdge;
    origin[1] = ((tile / T) % T) * kTileEdge;
    origin[2] = (tile / (T * T) + grid.slabTileZ0) * kTileEdge;
}

// ---------------------------------------------------------------------------------------------------------------------
// triangle setup shared by the count and emit passes

/// Transforms the model-space triangle `in` (applyMeshTransform, src/obj2voxel.cpp:202-209) and computes its area.
/// false: the triangle contributes nothing (zero area, non-finite, or a negative voxel-space coordinate).
template <bool UV>
__device__ __forceinline__ bool setupTriangle(const GridView &grid, const float in[9], Tri<UV> &t, float &area)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        affineApply(grid.xf, in + k * 3, t.v + k * 3);
    }
    area = triArea(t.v);
    // A negative voxel-space coordinate wraps the reference's float -> u32 cast to a huge chunkMin (triangle.hpp:91-95,
    // obj2voxel.cpp:211-219; formally UB), so the triangle lands in no chunk: dropped as a whole.
    bool negative = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        negative |= floorf(min3(t.v[a], t.v[3 + a], t.v[6 + a])) < 0.0f;
    }
    // weight 0 never reaches the voxel map (voxelization.cpp:466); non-finite input is a contract violation
    return area > 0.0f && area < INFINITY && !negative;
}

template <bool UV>
__device__ __forceinline__ bool loadTriangle(const MeshView &mesh, const GridView &grid, unsigned long long i,
                                             Tri<UV> &t, float &area)
{
    const float *src = mesh.verts + i * 9;
    float in[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        in[k] = __ldg(src + k);
    }
    if (UV) {
        const float *uv = mesh.uvs + i * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            t.t[k] = __ldg(uv + k);
        }
    }
    return setupTriangle<UV>(grid, in, t, area);
}

/// floor of the least and the greatest voxel-space z of the model-space triangle `in`: the three transformed z
/// coordinates alone, with the arithmetic of affineApply (so that the answer is traverseLeaves' root test).
__device__ __forceinline__ void triangleZRange(const GridView &grid, const float in[9], float &zlo, float &zhi)
{
    const float *m = grid.xf;
    float z[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        z[k] = xadd(dot3(m[6], m[7], m[8], in[k * 3], in[k * 3 + 1], in[k * 3 + 2]), m[11]);
    }
    zlo = floorf(min3(z[0], z[1], z[2]));
    zhi = floorf(max3(z[0], z[1], z[2]));
}

/// true if the voxel z range of the model-space triangle `in` provably misses this rank's slab.
/// Triangles with a negative z are kept: the count pass drops and counts them.
__device__ __forceinline__ bool triangleMissesSlab(const GridView &grid, const float in[9])
{
    float zlo, zhi;
    triangleZRange(grid, in, zlo, zhi);
    return zlo >= 0.0f && (toU32(zhi) + 1u <= grid.slabZ0 || toU32(zlo) >= grid.slabZ1);
}

// ---------------------------------------------------------------------------------------------------------------------
// triangle batches through shared memory: bulk-async copies (the TMA engine, no tensor map needed for a 1-D run)

__device__ __forceinline__ uint32_t sharedAddress(const void *p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbarrierInit(unsigned long long *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sharedAddress(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

/// Arms `bar` with the byte count and starts the copy global -> shared; src, dst and bytes are multiples of 16.
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses of dst come first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sharedAddress(bar)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sharedAddress(dst)),
                 "l"(src), "r"(bytes), "r"(sharedAddress(bar))
                 : "memory");
}

__device__ __forceinline__ void mbarrierWait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred done;\n"
        "WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
        "@!done bra WAIT;\n"
        "}\n" ::"r"(sharedAddress(bar)),
        "r"(parity)
        : "memory");
}

template <int kThreads>
struct TriangleBatch {
    alignas(128) float v[kThreads * 9];
    alignas(8) unsigned long long bar;
};

/// Calls visit(i, in, valid) from every thread of the block, once per batch of kThreads triangles of the contiguous array
/// verts[9 * count] (grid-stride over batches): `in` = the nine model-space floats of triangle i, valid = false for the
/// threads beyond the end of the ragged last batch (so that visit may use block-wide barriers).  Thread 0 starts the bulk copy of the
/// block's next batch as soon as the current one sits in registers, so the copy runs under the visit.  Batches the bulk
/// copy cannot take (array not 16-byte aligned; the ragged last batch) are loaded by the block.
template <int kThreads, typename Visit>
__device__ __forceinline__ void streamTriangles(const float *verts, unsigned long long count,
                                                TriangleBatch<kThreads> &sh, Visit &&visit)
{
    const uint32_t tid = threadIdx.x;
    const unsigned long long batches = (count + kThreads - 1) / kThreads;
    const bool aligned = (reinterpret_cast<unsigned long long>(verts) & 15ull) == 0;
    constexpr uint32_t kBatchBytes = kThreads * 9 * sizeof(float);
    static_assert(kBatchBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (tid == 0) {
        mbarrierInit(&sh.bar, 1);
    }
    __syncthreads();
    uint32_t parity = 0;
    unsigned long long b = blockIdx.x;
    if (tid == 0 && b < batches && aligned && (b + 1) * kThreads <= count) {
        bulkLoad(sh.v, verts + b * kThreads * 9, kBatchBytes, &sh.bar);
    }
    for (; b < batches; b += gridDim.x) {
        const unsigned long long first = b * kThreads;
        const uint32_t n = (uint32_t) (count - first < (unsigned long long) kThreads ? count - first : kThreads);
        if (aligned && n == (uint32_t) kThreads) {
            mbarrierWait(&sh.bar, parity);
            parity ^= 1u;
        }
        else {
            for (uint32_t k = tid; k < n * 9u; k += kThreads) {
                sh.v[k] = __ldg(verts + first * 9 + k);
            }
            __syncthreads();
        }
        float in[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            in[k] = sh.v[tid * 9 + k];  // stride 9 words: conflict-free
        }
        __syncthreads();  // the batch sits in registers: its buffer is free again
        const unsigned long long next = b + gridDim.x;
        if (tid == 0 && next < batches && aligned && (next + 1) * kThreads <= count) {
            bulkLoad(sh.v, verts + next * kThreads * 9, kBatchBytes, &sh.bar);
        }
        visit(first + tid, in, tid < n);
    }
}

/// visit(leaf, lo, hi) with the leaf's voxel AABB clamped to the chunk grid and this rank's slab, if anything is left.
template <bool UV, typename Visit>
__device__ __forceinline__ void visitClamped(const Tri<UV> &leaf, const GridView &grid, Visit &&visit)
{
    uint32_t lo[3], hi[3];
    triVoxelBounds(leaf.v, lo, hi);
    // voxels beyond the chunk grid belong to chunks the reference never dispatches (obj2voxel.cpp:503-505)
    hi[0] = min(hi[0], grid.gridExtent);
    hi[1] = min(hi[1], grid.gridExtent);
    lo[2] = max(lo[2], grid.slabZ0);
    hi[2] = min(hi[2], min(grid.slabZ1, grid.gridExtent));
    if (lo[0] >= hi[0] || lo[1] >= hi[1] || lo[2] >= hi[2]) {
        return;
    }
    visit(leaf, lo, hi);
}

// Huge triangles.  The subdivision of a triangle is walked by the thread that owns the triangle; a non-aligned triangle
// that spans thousands of voxels has 10^4 .. 10^5 leaves (a low-poly model at a high resolution), and one thread walking
// them takes tens of milliseconds while the rest of the device is done in one.  From kHugeRootVolume on the owner only
// lists the triangle (traverseLeaves: `huge`; listHugeTriangle), and separate kernels spread the listed triangles over
// the device as (triangle, subtree) items: the 4^kHugeSplitDepth = 256 subtrees four splits below the root, numbered in
// forEachLeaf's order (forEachLeafOfSubtree, o2v_exact.cuh).  A count-only walk of every item and a scan per triangle
// give each subtree its place in the triangle's leaf sequence, so that the leaves keep the reference's order wherever
// it matters.
#ifndef O2V_HUGE_ROOT_VOLUME
#define O2V_HUGE_ROOT_VOLUME (1ull << 21)  // voxels of the root AABB (128^3): >= ~10^3 leaves; A/B builds raise it to "never"
#endif
constexpr unsigned long long kHugeRootVolume = O2V_HUGE_ROOT_VOLUME;
constexpr int kHugeSplitDepth = 4;
constexpr uint32_t kHugeSubtrees = 1u << (2 * kHugeSplitDepth);
static_assert(kHugeSubtrees == kHugeSubtreesPerTriangle, "the engine sizes HugeWork::subtree by it");

/// Calls visit(leaf, lo, hi) for every leaf whose voxel AABB intersects this rank's slab, in the reference's order.
/// With `huge` given, a huge triangle is not walked: *huge = true and the caller lists it (listHugeTriangle).
template <bool UV, typename Visit>
__device__ __forceinline__ bool traverseLeaves(const Tri<UV> &root, const GridView &grid, Visit &&visit,
                                               bool *huge = nullptr)
{
    uint32_t rlo[3], rhi[3];
    triVoxelBounds(root.v, rlo, rhi);
    if (rhi[2] <= grid.slabZ0 || rlo[2] >= grid.slabZ1) {
        return true;  // midpoints stay inside the parent's AABB, so no leaf can reach the slab
    }
    if ((rhi[0] - rlo[0]) * (rhi[1] - rlo[1]) * (rhi[2] - rlo[2]) < kSubdivisionVolumeLimit) {
        // The triangle is its own single leaf whether or not it is axis-aligned (aligned: voxelized whole,
        // voxelization.cpp:498-501; otherwise the subdivision pops it at once, :349-379): no normal, no square root, and
        // the bounds are computed once.  Almost every triangle of a mesh that is fine relative to the grid ends here.
        rhi[0] = min(rhi[0], grid.gridExtent);
        rhi[1] = min(rhi[1], grid.gridExtent);
        rlo[2] = max(rlo[2], grid.slabZ0);
        rhi[2] = min(rhi[2], min(grid.slabZ1, grid.gridExtent));
        if (rlo[0] < rhi[0] && rlo[1] < rhi[1] && rlo[2] < rhi[2]) {
            visit(root, rlo, rhi);
        }
        return true;
    }
    if (huge != nullptr &&
        (unsigned long long) (rhi[0] - rlo[0]) * (rhi[1] - rlo[1]) * (rhi[2] - rlo[2]) >= kHugeRootVolume &&
        !triRoughlyAxisAligned(root.v)) {
        *huge = true;
        return true;
    }
    return forEachLeaf<UV>(root, [&](const Tri<UV> &leaf) { visitClamped<UV>(leaf, grid, visit); });
}

__device__ __forceinline__ void listHugeTriangle(const HugeWork &work, RunCounters *counters, unsigned long long index)
{
    const unsigned long long slot = atomicAdd(&counters->hugeTriangles, 1ull);
    if (slot < work.capacity) {
        work.list[slot] = static_cast<uint32_t>(index);
    }
}

/// visit(tri, area, leaf, lo, hi, seq) for every leaf of every listed (triangle, subtree) item, the items spread over the
/// grid (a thread walks whole subtrees).  seq = the leaf's position in its triangle's leaf sequence when the subtree
/// offsets are in place (hugeScanKernel), else its position in the subtree.  Returns the walks that hit
/// kMaxSubdivisionDepth.
template <bool UV, typename Visit>
__device__ __forceinline__ unsigned long long forEachHugeLeaf(const MeshView &mesh, const GridView &grid,
                                                              const HugeWork &work, const RunCounters *counters,
                                                              bool offsetsInPlace, Visit &&visit)
{
    const unsigned long long seen = counters->hugeTriangles;
    const unsigned long long items = (seen < work.capacity ? seen : work.capacity) * kHugeSubtrees;
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    unsigned long long overflow = 0;
    for (unsigned long long item = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; item < items;
         item += stride) {
        const uint32_t tri = work.list[item / kHugeSubtrees];
        Tri<UV> root;
        float area = 0.0f;
        loadTriangle<UV>(mesh, grid, tri, root, area);
        uint32_t seq = offsetsInPlace ? work.subtree[item] : 0u;
        const bool ok = forEachLeafOfSubtree<UV>(root, kHugeSplitDepth, (uint32_t) (item % kHugeSubtrees),
                                                 [&](const Tri<UV> &leaf) {
            visitClamped<UV>(leaf, grid, [&](const Tri<UV> &l, const uint32_t *lo, const uint32_t *hi) {
                visit(tri, area, l, lo, hi, seq);
                ++seq;
            });
        });
        overflow += ok ? 0 : 1;
    }
    return overflow;
}

/// Debug / parity record: the voxel's float WeightedColor (weight, r, g, b) as the fold left it, before the ARGB8
/// truncation — what obj2voxel::Voxelizer::voxels() holds in the reference (src/voxelization.hpp:55-108).
__device__ __forceinline__ void storeFloatRecord(const VoxelizeArgs &args, unsigned long long index, float w, float r,
                                                 float g, float b)
{
    if (args.floatOut != nullptr) {
        args.floatOut[index] = make_float4(w, r, g, b);
    }
}

struct VoxelAccumulator {
    // per-triangle uv buffer entry (voxelization.cpp:426-472) and the voxel itself (voxelization.cpp:513-526)
    WeightedUv partial;
    WeightedColor voxel;
    uint32_t partialTri;
    bool hasPartial;
    bool hasVoxel;
    uint32_t contributions;
};

__device__ __forceinline__ void flushPartial(VoxelAccumulator &acc, const VoxelizeArgs &args)
{
    if (!acc.hasPartial) {
        return;
    }
    acc.hasPartial = false;
    const uint32_t tri = acc.partialTri;
    const MeshView &mesh = args.mesh;
    uint8_t type;
    if (mesh.types != nullptr) {
        type = mesh.types[tri];
    }
    else {
        type = (mesh.uvs != nullptr && args.textureCount != 0) ? kTextured : kMaterialless;
    }
    float rgb[3] = {1.0f, 1.0f, 1.0f};  // MATERIALLESS: triangle.hpp:186
    if (type == kUntextured && mesh.colors != nullptr) {
        rgb[0] = mesh.colors[(size_t) tri * 3];
        rgb[1] = mesh.colors[(size_t) tri * 3 + 1];
        rgb[2] = mesh.colors[(size_t) tri * 3 + 2];
    }
    else if (type == kTextured && args.textureCount != 0) {
        uint32_t id = mesh.textureIds != nullptr ? mesh.textureIds[tri] : 0u;
        id = id < args.textureCount ? id : 0u;
        textureLookup(args.textures[id], acc.partial.u, acc.partial.v, rgb);
    }
    ++acc.contributions;
    if (!acc.hasVoxel) {
        acc.hasVoxel = true;
        acc.voxel.w = acc.partial.w;
        acc.voxel.r = rgb[0];
        acc.voxel.g = rgb[1];
        acc.voxel.b = rgb[2];
    }
    else {
        combineColorInto(acc.voxel, acc.partial.w, rgb[0], rgb[1], rgb[2], args.grid.strategy == kBlend);
    }
}

__device__ __forceinline__ void addContribution(VoxelAccumulator &acc, uint32_t tri, float w, float u, float v)
{
    if (acc.hasPartial) {
        blendUvInto(acc.partial, w, u, v);  // same triangle, later leaf: insertWeighted<BLEND>
    }
    else {
        acc.hasPartial = true;
        acc.partialTri = tri;
        acc.partial.w = w;
        acc.partial.u = u;
        acc.partial.v = v;
    }
}

template <bool UV>
__device__ __forceinline__ void stageLeaf(LeafStage &s, const VoxelizeArgs &args, uint32_t leafIndex,
                                          const uint32_t tileOrigin[3])
{
    const float4 *src = reinterpret_cast<const float4 *>(args.leaves + leafIndex);
    const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w;
    s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
    s.v[8] = c.x;
    s.tri = __float_as_uint(c.y);
    s.area = c.z;
    s.flags = __float_as_uint(c.w);
    if (UV) {
        const float4 *uv = reinterpret_cast<const float4 *>(args.leafUvs + leafIndex);
        const float4 u0 = __ldg(uv), u1 = __ldg(uv + 1);
        s.t[0] = u0.x; s.t[1] = u0.y; s.t[2] = u0.z; s.t[3] = u0.w;
        s.t[4] = u1.x; s.t[5] = u1.y;
    }
    uint32_t lo[3], hi[3];
    triVoxelBounds(s.v, lo, hi);
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint32_t l = lo[i] > tileOrigin[i] ? min(lo[i] - tileOrigin[i], kTileEdge) : 0u;
        const uint32_t h = hi[i] > tileOrigin[i] ? min(hi[i] - tileOrigin[i], kTileEdge) : 0u;
        packed |= l << (4 * i);
        packed |= h << (12 + 4 * i);
    }
    s.box = packed;
    const float origin[3] = {(float) tileOrigin[0], (float) tileOrigin[1], (float) tileOrigin[2]};
    buildPrefilter(s, origin);
}

/// Warp-synchronous form of clipLeafInVoxel (o2v_exact.cuh) — identical arithmetic and piece order, but all 32 lanes of the
/// warp drive it together.  The per-thread version leaves reconvergence to the compiler, which serialises the lanes of a
/// warp through the data-dependent loops (measured: 1.9 active lanes per instruction); here every lane steps through the
/// same two phases under explicit warp votes: a cheap classify/advance step repeated until no lane can advance, then one
/// shared split step.  The state is a struct so that a persistent kernel can refill finished lanes between rounds.
/// planeFlags (o2v_exact.cuh) from sign bits: c < p  <=>  (c - p) is negative, and |c - p| < eps  <=>  (|c - p| - eps) is
/// negative — IEEE subtraction never flips the sign of a non-zero exact difference and yields +0 for equal operands, so the
/// six flags are the sign bits of six correctly-rounded differences (no compare / select chain).
__device__ __forceinline__ uint32_t planeFlagsFast(float c0, float c1, float c2, float planePos)
{
    const float d0 = __fsub_rn(c0, planePos), d1 = __fsub_rn(c1, planePos), d2 = __fsub_rn(c2, planePos);
    const float e0 = __fsub_rn(fabsf(d0), kEpsilon), e1 = __fsub_rn(fabsf(d1), kEpsilon),
                e2 = __fsub_rn(fabsf(d2), kEpsilon);
    const uint32_t L = (__float_as_uint(d0) >> 31) | ((__float_as_uint(d1) >> 31) << 1) |
                       ((__float_as_uint(d2) >> 31) << 2);
    const uint32_t P = (__float_as_uint(e0) >> 31) | ((__float_as_uint(e1) >> 31) << 1) |
                       ((__float_as_uint(e2) >> 31) << 2);
    return (P << 3) | L;
}

/// Fills the 64-entry case table (index = (planar flags << 3) | lo flags) cooperatively; the caller synchronises.
__device__ __forceinline__ void fillClipCaseTable(uint8_t *table)
{
    for (uint32_t i = threadIdx.x; i < 64u; i += blockDim.x) {
        table[i] = static_cast<uint8_t>(clipCaseOf(i >> 3, i & 7u));
    }
}

template <bool UV>
struct ClipStack {  // pieces waiting for their remaining planes; kept apart from WarpClipper so that only this array is
    Tri<UV> piece[6];  // dynamically indexed (local memory) while the clipper's scalars stay in registers
    uint8_t plane[6];
};

template <bool UV>
struct WarpClipper {
    Tri<UV> cur;
    ClipResult r;
    float wholeArea;
    uint32_t px, py, pz;
    int sp;
    int plane;
    bool done;  // no piece left (or the lane never had work)

    __device__ __forceinline__ void idle()
    {
        done = true;
        sp = 0;
        plane = 0;
    }

    __device__ __forceinline__ void begin(const Tri<UV> &leaf, uint32_t x, uint32_t y, uint32_t z, float area)
    {
        cur = leaf;
        px = x;
        py = y;
        pz = z;
        wholeArea = area;
        r.pieces = 0;
        r.weight = 0.0f;
        r.u = 0.0f;
        r.v = 0.0f;
        sp = 0;
        plane = 0;
        done = false;
    }

    /// One round for the whole warp: every lane that still has a piece advances it to its next real split (phase A) and
    /// performs that split (phase B).  Must be called by all 32 lanes.
    __device__ __forceinline__ void popOrFinish(ClipStack<UV> &stack)
    {
        if (sp == 0) {
            done = true;
        }
        else {
            --sp;
            cur = stack.piece[sp];
            plane = stack.plane[sp];
        }
    }

    /// One round for the whole warp = one node of every lane's clip tree: classify the current piece against all six
    /// voxel planes in one straight-line block (no per-lane axis selects, no data-dependent inner loop; the 16-way case
    /// switch is a 64-entry shared-memory table, clipCaseOf), then either count it as a survivor, drop it, or split it at
    /// the first plane >= `plane` that cuts it.  Must be called by all 32 lanes.
    __device__ __forceinline__ void round(ClipStack<UV> &stack, const uint8_t *caseTable)
    {
        const unsigned int full = 0xffffffffu;
        bool needSplit = false;
        uint32_t code = 0;
        if (!done) {
            uint32_t nonKeep = 0, dropMask = 0, codes = 0;
#pragma unroll
            for (int axis = 0; axis < 3; ++axis) {
                const float c0 = cur.v[axis], c1 = cur.v[3 + axis], c2 = cur.v[6 + axis];
                const uint32_t base = axis == 0 ? px : (axis == 1 ? py : pz);
                // plane `axis` keeps the hi side (DISCARD_LO), plane `3 + axis` = base + 1 keeps the lo side
                const uint32_t a = caseTable[planeFlagsFast(c0, c1, c2, static_cast<float>(base))];
                const uint32_t b = caseTable[planeFlagsFast(c0, c1, c2, static_cast<float>(base + 1u))];
                nonKeep |= ((a & 7u) == 0u ? 0u : 1u) << axis;          // kept whole iff unsplit and not lo
                nonKeep |= ((b & 7u) == 4u ? 0u : 1u) << (3 + axis);    // kept whole iff unsplit and lo
                dropMask |= ((a & 7u) == 4u ? 1u : 0u) << axis;
                dropMask |= ((b & 7u) == 0u ? 1u : 0u) << (3 + axis);
                codes |= (a << (5 * axis)) | (b << (5 * (3 + axis)));
            }
            const uint32_t cutting = nonKeep & 63u & ~((1u << plane) - 1u);  // planes still ahead of this piece
            if (cutting == 0) {
                // survived every remaining plane: result = mix(result, {area, textureCenter}) (util.hpp:160-165)
                const float weightSum = xadd(r.weight, wholeArea);
                if (UV) {
                    const float cu = xdiv(xadd(xadd(cur.t[0], cur.t[2]), cur.t[4]), 3.0f);
                    const float cv = xdiv(xadd(xadd(cur.t[1], cur.t[3]), cur.t[5]), 3.0f);
                    r.u = xdiv(xadd(xmul(r.weight, r.u), xmul(wholeArea, cu)), weightSum);
                    r.v = xdiv(xadd(xmul(r.weight, r.v), xmul(wholeArea, cv)), weightSum);
                }
                r.weight = weightSum;
                ++r.pieces;
                popOrFinish(stack);
            }
            else {
                const int first = __ffs(cutting) - 1;
                if ((dropMask >> first) & 1u) {
                    popOrFinish(stack);
                }
                else {
                    needSplit = true;
                    plane = first;
                    code = (codes >> (5 * first)) & 31u;
                }
            }
        }
        const ClipAction action = (code & 3u) == 1u ? kClipSplitRegular : kClipSplitOnePlanar;
        const int pivot = static_cast<int>(code >> 3);
        const bool sideLo = (code & 4u) != 0;

        // ---- phase B: one split for every lane that needs one ----
        if (needSplit) {
            const int axis = plane < 3 ? plane : plane - 3;
            const uint32_t base = axis == 0 ? px : (axis == 1 ? py : pz);
            const float planePos = static_cast<float>(base + (plane < 3 ? 0u : 1u));
            const bool keepHi = plane < 3;
            rotateToPivot<UV>(cur, pivot);
            if (action == kClipSplitRegular) {
                // splitTriangle_regularCase, voxelization.cpp:279-331: pivot = isolated vertex, X0/X1 on its two edges
                const float s0 = intersectAxisPlane(cur.v, cur.v + 3, axis, planePos);
                const float s1 = intersectAxisPlane(cur.v, cur.v + 6, axis, planePos);
                float g0[3], g1[3], x0[2] = {0.0f, 0.0f}, x1[2] = {0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    g0[i] = mix1(cur.v[i], cur.v[3 + i], s0);
                    g1[i] = mix1(cur.v[i], cur.v[6 + i], s1);
                }
                if (UV) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        x0[i] = mix1(cur.t[i], cur.t[2 + i], s0);
                        x1[i] = mix1(cur.t[i], cur.t[4 + i], s1);
                    }
                }
                if (sideLo != keepHi) {
                    setVertex<UV>(cur, 1, g0, x0);  // the isolated corner (iso, X0, X1) is kept
                    setVertex<UV>(cur, 2, g1, x1);
                }
                else {
                    Tri<UV> second;  // the quad is kept: (X0, a, b) now, (X0, X1, b) pending for the next plane
                    setVertex<UV>(second, 0, g0, x0);
                    setVertex<UV>(second, 1, g1, x1);
                    setVertex<UV>(second, 2, cur.v + 6, cur.t + (UV ? 4 : 0));
                    stack.piece[sp] = second;
                    stack.plane[sp] = static_cast<uint8_t>(plane + 1);
                    ++sp;
                    setVertex<UV>(cur, 0, g0, x0);
                }
            }
            else {
                // splitTriangle_onePlanarCase, voxelization.cpp:240-277: pivot = planar vertex, X on the opposite edge
                const float s = intersectAxisPlane(cur.v + 3, cur.v + 6, axis, planePos);
                float geo[3], tex[2] = {0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    geo[i] = mix1(cur.v[3 + i], cur.v[6 + i], s);
                }
                if (UV) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        tex[i] = mix1(cur.t[2 + i], cur.t[4 + i], s);
                    }
                }
                if (sideLo != keepHi) {
                    setVertex<UV>(cur, 2, geo, tex);  // keep (p, a, X)
                }
                else {
                    setVertex<UV>(cur, 1, geo, tex);  // keep (p, X, b)
                }
            }
            ++plane;
        }
        __syncwarp(full);
    }
};

/// Clips one leaf per lane to completion (all 32 lanes call it together; lanes without work pass valid = false).
template <bool UV>
__device__ __forceinline__ ClipResult clipLeafInVoxelWarp(bool valid, const Tri<UV> &leaf, uint32_t px, uint32_t py,
                                                          uint32_t pz, float wholeArea, const uint8_t *caseTable)
{
    WarpClipper<UV> clipper;
    ClipStack<UV> stack;
    clipper.begin(leaf, px, py, pz, wholeArea);
    if (!valid) {
        clipper.idle();
    }
    while (!__all_sync(0xffffffffu, clipper.done)) {
        clipper.round(stack, caseTable);
    }
    return clipper.r;
}

/// Voxel key: parent (2x2x2 block) index in the high 6 bits, child Morton code (x most significant, ileave.hpp:243-246) in
/// the low 3 — ascending keys visit the children of one parent in ascending Morton order, which is the downscale order.
__device__ __forceinline__ uint32_t voxelKey(uint32_t x, uint32_t y, uint32_t z)
{
    const uint32_t parent = (x >> 1) | ((y >> 1) << 2) | ((z >> 1) << 4);
    const uint32_t child = ((x & 1u) << 2) | ((y & 1u) << 1) | (z & 1u);
    return (parent << 3) | child;
}

__device__ __forceinline__ void resetAccumulator(VoxelAccumulator &acc)
{
    acc.hasPartial = false;
    acc.hasVoxel = false;
    acc.partialTri = 0;
    acc.contributions = 0;
    acc.partial.w = acc.partial.u = acc.partial.v = 0.0f;
    acc.voxel.w = acc.voxel.r = acc.voxel.g = acc.voxel.b = 0.0f;
}


}  // namespace
}  // namespace o2v

#endif  // O2V_DEVICE_CUH
