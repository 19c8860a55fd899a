// o2v::Engine — buffer management and kernel sequencing for one GPU (see o2v_engine.h).
#include "o2v_engine.h"
#include "o2v_sat.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace o2v {

#define O2V_CUDA(expr)                                                                                  \
    do {                                                                                                \
        const cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                                     \
            return fail(kErrCuda, std::string(#expr) + ": " + cudaGetErrorString(err__));               \
        }                                                                                               \
    } while (0)

DeviceBuffer::~DeviceBuffer()
{
    if (ptr_ != nullptr) {
        cudaFree(ptr_);
    }
}

bool DeviceBuffer::ensure(size_t bytes)
{
    if (bytes <= size_) {
        return true;
    }
    // grow with head-room so repeated runs of slightly different sizes do not reallocate every time
    const size_t wanted = bytes + bytes / 8 + 256;
    void *fresh = nullptr;
    if (cudaMalloc(&fresh, wanted) != cudaSuccess) {
        cudaGetLastError();
        if (cudaMalloc(&fresh, bytes) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        size_ = bytes;
    }
    else {
        size_ = wanted;
    }
    if (ptr_ != nullptr) {
        cudaFree(ptr_);
    }
    ptr_ = fresh;
    return true;
}

Engine *Engine::create(int device, std::string *error)
{
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        if (error != nullptr) {
            *error = std::string("no CUDA device available (") + cudaGetErrorString(err) +
                     "); obj2voxel_b200 has no CPU fallback";
        }
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= count) {
        if (error != nullptr) {
            *error = "invalid CUDA device index";
        }
        return nullptr;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    err = cudaGetDeviceProperties(&prop, device);
    if (err != cudaSuccess) {
        if (error != nullptr) {
            *error = cudaGetErrorString(err);
        }
        return nullptr;
    }
    Engine *e = new Engine();
    e->device_ = device;
    e->smCount_ = prop.multiProcessorCount;
    e->totalMemory_ = prop.totalGlobalMem;
    bool ok = cudaHostAlloc(&e->hostCounters_, sizeof(RunCounters), cudaHostAllocMapped) == cudaSuccess &&
              cudaHostGetDevicePointer(&e->hostCountersDevice_, e->hostCounters_, 0) == cudaSuccess;
    ok = ok && cudaMallocHost(&e->hostCountersInit_, sizeof(RunCounters)) == cudaSuccess;
    ok = ok && cudaEventCreate(&e->evStart_) == cudaSuccess && cudaEventCreate(&e->evSetup_) == cudaSuccess;
    ok = ok && cudaEventCreate(&e->evVoxStart_) == cudaSuccess && cudaEventCreate(&e->evVoxEnd_) == cudaSuccess;
    ok = ok && cudaEventCreate(&e->evClipStart_) == cudaSuccess && cudaEventCreate(&e->evClipEnd_) == cudaSuccess;
    ok = ok && cudaEventCreate(&e->evClassifyStart_) == cudaSuccess;
    ok = ok && cudaEventCreate(&e->evFilterStart_) == cudaSuccess;
    ok = ok && e->counters_.ensure(sizeof(RunCounters));
    if (!ok) {
        if (error != nullptr) {
            *error = std::string("engine setup failed: ") + cudaGetErrorString(cudaGetLastError());
        }
        delete e;
        return nullptr;
    }
    memset(e->hostCountersInit_, 0, sizeof(RunCounters));
    for (int i = 0; i < 3; ++i) {
        e->hostCountersInit_->boundsMinBits[i] = 0xffffffffu;
        e->hostCountersInit_->boundsMaxBits[i] = 0u;
    }
    return e;
}

Engine::~Engine()
{
    cudaSetDevice(device_);
    if (hostCounters_ != nullptr) {
        cudaFreeHost(hostCounters_);
    }
    if (hostCountersInit_ != nullptr) {
        cudaFreeHost(hostCountersInit_);
    }
    for (void *p : staging_) {
        if (p != nullptr) {
            cudaFreeHost(p);
        }
    }
    for (cudaEvent_t ev : {evStart_, evSetup_, evVoxStart_, evVoxEnd_, evClipStart_, evClipEnd_, evClassifyStart_, evFilterStart_}) {
        if (ev != nullptr) {
            cudaEventDestroy(ev);
        }
    }
}

int Engine::fail(int code, const std::string &message)
{
    error_ = message;
    cudaGetLastError();
    return code;
}

/// Bounds (given or computed on the device), the exact mesh transform and this run's slab: everything of GridView.
int Engine::setupGrid(const MeshView &mesh, const EngineParams &params, cudaStream_t stream, RunStats &st, GridView &grid,
                      bool *emptySlab)
{
    RunCounters *dCounters = counters_.as<RunCounters>();
    const uint32_t S = params.resolution * params.supersampling;
    // ---- bounds + transform (src/obj2voxel.cpp:475-482) ----
    float meshMin[3], meshMax[3];
    if (params.boundsKnown) {
        memcpy(meshMin, params.bounds, sizeof meshMin);
        memcpy(meshMax, params.bounds + 3, sizeof meshMax);
    }
    else {
        launchBounds(mesh, dCounters, stream);
        launchFinishBounds(dCounters, stream);
        st.kernelLaunches += 2;
        launchPublishCounters(dCounters, hostCountersDevice_, stream);
        ++st.kernelLaunches;
        O2V_CUDA(cudaStreamSynchronize(stream));
        memcpy(meshMin, hostCounters_->boundsMin, sizeof meshMin);
        memcpy(meshMax, hostCounters_->boundsMax, sizeof meshMax);
    }

    computeMeshTransform(meshMin, meshMax, S, params.unitTransform, grid.xf);
    memcpy(st.transform, grid.xf, sizeof st.transform);
    grid.sampleRes = S;
    grid.gridExtent = (S + 63u) / 64u * 64u;
    grid.tilesPerAxis = grid.gridExtent / kTileEdge;
    grid.tileShift = 32;
    for (uint32_t shift = 0; shift < 32; ++shift) {
        if (grid.tilesPerAxis == (1u << shift)) {
            grid.tileShift = shift;
        }
    }
    const uint32_t gridZ = (S + 63u) / 64u * 64u;
    uint32_t z0 = params.slabZ0, z1 = params.slabZ1;
    if (z0 == 0 && z1 == 0) {
        z1 = gridZ;
    }
    z1 = std::min(z1, gridZ);
    if (z0 % kTileEdge != 0 || z1 % kTileEdge != 0 || z0 > z1) {
        return fail(kErrBadParams, "slab bounds must be multiples of 8 with z0 <= z1");
    }
    grid.slabZ0 = z0;
    grid.slabZ1 = z1;
    grid.slabTileZ0 = z0 / kTileEdge;
    grid.slabTileZCount = (z1 - z0) / kTileEdge;
    grid.supersampling = params.supersampling;
    grid.strategy = params.strategy;
    *emptySlab = grid.slabTileZCount == 0;
    return kErrOk;
}

int Engine::voxelize(const MeshView &meshIn, const TextureView *textures, uint32_t textureCount,
                     const EngineParams &params, cudaStream_t stream, RunStats *stats)
{
    int rc = kRetryHugeWalk;
    for (int attempt = 0; attempt < 3 && rc == kRetryHugeWalk; ++attempt) {
        rc = voxelizeOnce(meshIn, textures, textureCount, params, stream, stats);
    }
    return rc == kRetryHugeWalk ? fail(kErrTooLarge, "the number of huge triangles keeps changing between attempts") : rc;
}

size_t Engine::accumulateBytes(const EngineParams &params)
{
    const unsigned long long S = (unsigned long long) params.resolution * params.supersampling;
    if (S == 0 || S > 8192ull) {
        return 0;
    }
    const uint32_t shift = params.supersampling == 2 ? 1u : 0u;
    const uint32_t gridExtent = (uint32_t) ((S + 63u) / 64u * 64u);
    const uint32_t z0 = params.slabZ0, z1 = (params.slabZ0 == 0 && params.slabZ1 == 0) ? gridExtent
                                                                                          : std::min(params.slabZ1, gridExtent);
    const unsigned long long perAxis = ((gridExtent >> shift) + kChunkEdge - 1) / kChunkEdge;
    const unsigned long long rows = ((z1 >> shift) + kChunkEdge - 1) / kChunkEdge - (z0 >> shift) / kChunkEdge;
    const unsigned long long bytes = perAxis * perAxis * rows * kChunkWords * 8ull;
    return bytes <= (2ull << 30) ? (size_t) (2 * bytes) : 0;
}

bool Engine::hugeWork(HugeWork &work)
{
    work = HugeWork{nullptr, nullptr, 0};
    if (!walkHuge_) {
        return true;
    }
    if (hugeExpected_ >= (1ull << 24) || !hugeList_.ensure((size_t) hugeExpected_ * 4) ||
        !hugeSubtree_.ensure((size_t) hugeExpected_ * kHugeSubtreesPerTriangle * 4)) {
        return false;
    }
    work = HugeWork{hugeList_.as<uint32_t>(), hugeSubtree_.as<uint32_t>(), (uint32_t) hugeExpected_};
    return true;
}

int Engine::voxelizeOnce(const MeshView &meshIn, const TextureView *textures, uint32_t textureCount,
                         const EngineParams &params, cudaStream_t stream, RunStats *stats)
{
    // Every triangle MATERIALLESS (no per-triangle types, no usable texture): the output is white wherever a voxel is
    // occupied, whatever the weights — the occupancy-only path decides just that (o2v_occupancy.cu).
    bool occupancy = params.occupancyPath != 0 && !params.floatRecords && meshIn.types == nullptr &&
                     !(meshIn.uvs != nullptr && textureCount != 0);
    MeshView mesh = meshIn;
    if (occupancy) {
        mesh.uvs = nullptr;     // uvs without a texture never reach a colour (flushPartial)
        mesh.colors = nullptr;  // colours without UNTEXTURED types neither
    }
    error_.clear();
    voxelCount_ = 0;
    bitmapValid_ = false;
    packedBits_ = 0;
    floatValid_ = false;
    RunStats local;
    RunStats &st = stats != nullptr ? *stats : local;
    st = RunStats();

    if (params.resolution == 0 || params.supersampling == 0 || params.supersampling > 2 || params.strategy > 1) {
        return fail(kErrBadParams, "resolution must be > 0, supersampling 1 or 2, strategy 0 or 1");
    }
    const unsigned long long sampleRes64 = (unsigned long long) params.resolution * params.supersampling;
    if (sampleRes64 > 8192ull) {
        return fail(kErrTooLarge, "sample resolution above 8192 is not supported");
    }
    if (mesh.count >= (1ull << 32)) {
        return fail(kErrTooLarge, "more than 2^32-1 triangles");
    }
    O2V_CUDA(cudaSetDevice(device_));

    RunCounters *dCounters = counters_.as<RunCounters>();
    O2V_CUDA(cudaMemcpyAsync(dCounters, hostCountersInit_, sizeof(RunCounters), cudaMemcpyHostToDevice, stream));
    O2V_CUDA(cudaEventRecord(evStart_, stream));
    memset(&st.counters, 0, sizeof st.counters);
    if (mesh.count == 0) {
        return kErrOk;  // src/obj2voxel.cpp:590-594: empty model, empty output
    }

    GridView grid;
    bool emptySlab = false;
    if (const int rc = setupGrid(mesh, params, stream, st, grid, &emptySlab)) {
        return rc;
    }
    if (emptySlab) {
        return kErrOk;
    }
    if (occupancy) {
        const int rc = voxelizeOccupancy(mesh, params, grid, stream, st);
        if (rc != kOccupancyFallback) {
            return rc;
        }
        // the chunk bitmaps do not fit: fold weights like for any other mesh (same result)
        st = RunStats();
        memcpy(st.transform, grid.xf, sizeof st.transform);
        occupancy = false;
        O2V_CUDA(cudaMemcpyAsync(dCounters, hostCountersInit_, sizeof(RunCounters), cudaMemcpyHostToDevice, stream));
    }

    const unsigned long long tileTotal64 =
        (unsigned long long) grid.tilesPerAxis * grid.tilesPerAxis * grid.slabTileZCount;
    if (tileTotal64 >= (1ull << 31)) {
        return fail(kErrTooLarge, "tile grid too large");
    }
    const uint32_t tileTotal = (uint32_t) tileTotal64;
    const size_t n = (size_t) mesh.count;

    const size_t scratchElems = std::max(scanScratchElems(n), scanScratchElems(tileTotal));
    if (!leafCount_.ensure(n * 4) || !leafOffset_.ensure(n * 4) || !tileCount_.ensure((size_t) tileTotal * 4) ||
        !tileStart_.ensure((size_t) tileTotal * 4) || !tileFill_.ensure((size_t) tileTotal * 4) ||
        !activeTiles_.ensure((size_t) tileTotal * 4) || !allTiles_.ensure((size_t) tileTotal * 4) ||
        !longTiles_.ensure((size_t) tileTotal * 4) ||
        !tileCand_.ensure((size_t) tileTotal * 4) ||
        !lightTiles_.ensure((size_t) tileTotal * sizeof(LightTile)) ||
        !bigLightTiles_.ensure((size_t) tileTotal * sizeof(LightTile)) || !scratch_.ensure(scratchElems * 4)) {
        return fail(kErrOutOfMemory, "device allocation failed (binning buffers)");
    }
    O2V_CUDA(cudaMemsetAsync(tileCount_.as<void>(), 0, (size_t) tileTotal * 4, stream));
    O2V_CUDA(cudaMemsetAsync(tileFill_.as<void>(), 0, (size_t) tileTotal * 4, stream));
    O2V_CUDA(cudaMemsetAsync(tileCand_.as<void>(), 0, (size_t) tileTotal * 4, stream));

    HugeWork huge;
    if (!hugeWork(huge)) {
        return fail(kErrOutOfMemory, "device allocation failed (huge triangles)");
    }
    launchCountLeaves(mesh, grid, leafCount_.as<uint32_t>(), tileCount_.as<uint32_t>(), tileCand_.as<uint32_t>(),
                      dCounters, huge, stream);
    if (huge.capacity != 0) {
        launchHugeSubtreeScan(mesh, grid, huge, dCounters, leafCount_.as<uint32_t>(), false, hugeExpected_, stream);
        launchHugeCountTiles(mesh, grid, huge, tileCount_.as<uint32_t>(), tileCand_.as<uint32_t>(), dCounters,
                             hugeExpected_, stream);
        st.kernelLaunches += 3;
    }
    launchExclusiveScan(leafCount_.as<uint32_t>(), leafOffset_.as<uint32_t>(), n, scratch_.as<uint32_t>(),
                        &dCounters->leaves, stream);
    launchExclusiveScan(tileCount_.as<uint32_t>(), tileStart_.as<uint32_t>(), tileTotal, scratch_.as<uint32_t>(),
                        &dCounters->pairs, stream);
    launchCompactActiveTiles(tileCount_.as<uint32_t>(), tileCand_.as<uint32_t>(), tileStart_.as<uint32_t>(), tileTotal,
                             allTiles_.as<uint32_t>(), longTiles_.as<uint32_t>(), activeTiles_.as<uint32_t>(),
                             lightTiles_.as<LightTile>(), bigLightTiles_.as<LightTile>(), dCounters, stream);
    st.kernelLaunches += 8;
    launchPublishCounters(dCounters, hostCountersDevice_, stream);
    ++st.kernelLaunches;
    O2V_CUDA(cudaStreamSynchronize(stream));
    O2V_CUDA(cudaGetLastError());

    if (hostCounters_->hugeTriangles > huge.capacity) {
        walkHuge_ = true;  // their leaves are not counted yet: once more, with room to list them
        hugeExpected_ = hostCounters_->hugeTriangles;
        return kRetryHugeWalk;
    }
    walkHuge_ = hostCounters_->hugeTriangles != 0;
    hugeExpected_ = hostCounters_->hugeTriangles;
    const unsigned long long leafTotal = hostCounters_->leaves;
    const unsigned long long pairTotal = hostCounters_->pairs;
    const unsigned long long activeTotal = hostCounters_->activeTiles;
    if (leafTotal >= (1ull << 32) || pairTotal >= (1ull << 32)) {
        return fail(kErrTooLarge, "more than 2^32-1 leaves or (leaf, tile) pairs in this slab");
    }

    // Output capacity: every voxel needs at least one candidate, and a tile emits at most 512 (64 when downscaled).
    const unsigned long long perTile = params.supersampling == 2 ? kTileVoxels / 8 : kTileVoxels;
    unsigned long long capacity = std::min(hostCounters_->candidateVoxels, activeTotal * perTile);
    capacity = std::max<unsigned long long>(capacity, 1);

    const bool hasUv = mesh.uvs != nullptr;
    if (!leaves_.ensure((size_t) std::max<unsigned long long>(leafTotal, 1) * sizeof(LeafRecord)) ||
        (hasUv && !leafUvs_.ensure((size_t) std::max<unsigned long long>(leafTotal, 1) * sizeof(LeafUv))) ||
        !tileList_.ensure((size_t) std::max<unsigned long long>(pairTotal, 1) * 4) ||
        !pairTile_.ensure((size_t) std::max<unsigned long long>(pairTotal, 1) * 4) ||
        !pairSurvivors_.ensure((size_t) (pairTotal + 1) * 4) || !pairOffset_.ensure((size_t) (pairTotal + 1) * 4) ||
        !pairMask_.ensure((size_t) (pairTotal + 1) * 8) || !pairBox_.ensure((size_t) (pairTotal + 1) * 4) ||
        !scratch_.ensure(std::max(scratchElems, scanScratchElems((size_t) pairTotal + 1)) * 4)) {
        return fail(kErrOutOfMemory, "device allocation failed (leaf buffers)");
    }
    if (capacity * sizeof(VoxelRecord) > out_.size()) {  // only when the buffer has to grow: bound it by free memory
        size_t freeBytes = 0, totalBytes = 0;
        O2V_CUDA(cudaMemGetInfo(&freeBytes, &totalBytes));
        const unsigned long long affordable = (freeBytes + out_.size()) / sizeof(VoxelRecord) * 9 / 10;
        capacity = std::min(capacity, std::max<unsigned long long>(affordable, 1));
    }
    if (!out_.ensure((size_t) capacity * sizeof(VoxelRecord)) ||
        (params.floatRecords && !floatOut_.ensure((size_t) capacity * sizeof(float4)))) {
        return fail(kErrOutOfMemory, "device allocation failed (voxel output)");
    }
    if (textureCount != 0) {
        if (!textures_.ensure(textureCount * sizeof(TextureView))) {
            return fail(kErrOutOfMemory, "device allocation failed (texture table)");
        }
        O2V_CUDA(cudaMemcpyAsync(textures_.as<void>(), textures, textureCount * sizeof(TextureView),
                                 cudaMemcpyHostToDevice, stream));
    }

    launchEmitLeaves(mesh, grid, leafOffset_.as<uint32_t>(), tileStart_.as<uint32_t>(), tileFill_.as<uint32_t>(),
                     leaves_.as<LeafRecord>(), hasUv ? leafUvs_.as<LeafUv>() : nullptr, tileList_.as<uint32_t>(),
                     pairTile_.as<uint32_t>(), dCounters, huge, hugeExpected_, stream);
    TileWork work;
    work.allTiles = allTiles_.as<uint32_t>();
    work.allCount = (uint32_t) activeTotal;
    work.longTiles = longTiles_.as<uint32_t>();
    work.longCount = (uint32_t) hostCounters_->longTiles;
    work.activeTiles = activeTiles_.as<uint32_t>();
    work.tileStart = tileStart_.as<uint32_t>();
    work.tileCount = tileCount_.as<uint32_t>();
    work.tileList = tileList_.as<uint32_t>();
    work.activeCount = (uint32_t) hostCounters_->heavyTiles;
    launchSortTileLists(work, tileList_.as<uint32_t>(), stream);
    st.kernelLaunches += 3;

    VoxelizeArgs args;
    args.grid = grid;
    args.work = work;
    args.leaves = leaves_.as<LeafRecord>();
    args.leafUvs = hasUv ? leafUvs_.as<LeafUv>() : nullptr;
    args.mesh = mesh;
    args.textures = textureCount != 0 ? textures_.as<TextureView>() : nullptr;
    args.textureCount = textureCount;
    args.out = out_.as<VoxelRecord>();
    args.floatOut = params.floatRecords ? floatOut_.as<float4>() : nullptr;
    args.outCapacity = capacity;
    args.counters = dCounters;
    args.lightTiles = lightTiles_.as<LightTile>();
    args.lightCount = (uint32_t) hostCounters_->lightTiles;
    args.bigLightTiles = bigLightTiles_.as<LightTile>();
    args.bigLightCount = (uint32_t) hostCounters_->bigLightTiles;
    args.variant = params.variant < 0 ? 0 : params.variant;
    args.prefilter = params.prefilter;

    // ---- staged sparse path, stages 1-2: SAT survivors of every (leaf, light tile) pair, compacted in list order ----
    SparseView &sparse = args.sparse;
    sparse.pairTile = pairTile_.as<uint32_t>();
    sparse.tileCandidates = tileCand_.as<uint32_t>();
    sparse.pairCount = (uint32_t) pairTotal;
    sparse.pairSurvivors = pairSurvivors_.as<uint32_t>();
    sparse.pairOffset = pairOffset_.as<uint32_t>();
    sparse.pairMask = pairMask_.as<unsigned long long>();
    sparse.pairBox = pairBox_.as<uint32_t>();
    sparse.entries = nullptr;
    sparse.uvs = nullptr;
    // Survivors <= candidate voxels (known from the first read-back).  When that bound is affordable the queue is sized
    // by it and the exact count stays on the device (no host round trip between the stages).
    const unsigned long long candidateBound = hostCounters_->candidateVoxels;
    const bool boundAffordable = candidateBound < (1ull << 32) && candidateBound * 16ull <= (8ull << 30);
    bool sparseActive = args.lightCount != 0 || args.bigLightCount != 0;
    if (sparseActive) {
        O2V_CUDA(cudaMemsetAsync(pairSurvivors_.as<uint32_t>() + pairTotal, 0, 4, stream));
        launchSparseSurvivors(args, false, stream);
        launchExclusiveScan(pairSurvivors_.as<uint32_t>(), pairOffset_.as<uint32_t>(), (size_t) pairTotal + 1,
                            scratch_.as<uint32_t>(), &dCounters->survivors, stream);
        st.kernelLaunches += 4;
        unsigned long long entryCapacity = candidateBound;
        if (!boundAffordable) {
            launchPublishCounters(dCounters, hostCountersDevice_, stream);
            ++st.kernelLaunches;
            O2V_CUDA(cudaStreamSynchronize(stream));
            O2V_CUDA(cudaGetLastError());
            entryCapacity = hostCounters_->survivors;
            if (entryCapacity >= (1ull << 32)) {
                return fail(kErrTooLarge, "more than 2^32-1 candidate voxels on the sparse path of this slab");
            }
        }
        entryCapacity = std::max<unsigned long long>(entryCapacity, 1);
        if (!entries_.ensure((size_t) entryCapacity * sizeof(uint4)) ||
            (hasUv && !contribUvs_.ensure((size_t) entryCapacity * sizeof(float2)))) {
            return fail(kErrOutOfMemory, "device allocation failed (sparse path buffers)");
        }
        sparse.entries = entries_.as<uint4>();
        sparse.uvs = hasUv ? contribUvs_.as<float2>() : nullptr;
        launchSparseSurvivors(args, true, stream);
        ++st.kernelLaunches;
    }
    O2V_CUDA(cudaEventRecord(evSetup_, stream));

    for (int attempt = 0; attempt < 2; ++attempt) {
        O2V_CUDA(cudaEventRecord(evVoxStart_, stream));
        if (sparseActive) {
            O2V_CUDA(cudaEventRecord(evClipStart_, stream));
            launchSparseClip(args, smCount_, stream);
            O2V_CUDA(cudaEventRecord(evClipEnd_, stream));
            launchSparseFold(args, smCount_, stream);
        }
        launchVoxelizeTiles(args, smCount_, stream);
        O2V_CUDA(cudaEventRecord(evVoxEnd_, stream));
        const int launched = (sparseActive ? 4 : 0) + (args.work.activeCount != 0 ? 1 : 0);
        st.voxelizeLaunches += launched;
        st.kernelLaunches += launched;
        launchPublishCounters(dCounters, hostCountersDevice_, stream);
        ++st.kernelLaunches;
        O2V_CUDA(cudaStreamSynchronize(stream));
        O2V_CUDA(cudaGetLastError());
        if (hostCounters_->outputOverflow == 0) {
            break;
        }
        if (attempt == 1) {
            return fail(kErrOutOfMemory, "voxel output does not fit device memory");
        }
        // the exact count is now known: grow once and redo the tile pass (setup results are still valid)
        capacity = hostCounters_->voxels;
        if (!out_.ensure((size_t) capacity * sizeof(VoxelRecord)) ||
            (params.floatRecords && !floatOut_.ensure((size_t) capacity * sizeof(float4)))) {
            return fail(kErrOutOfMemory, "device allocation failed (voxel output, exact size)");
        }
        args.out = out_.as<VoxelRecord>();
        args.floatOut = params.floatRecords ? floatOut_.as<float4>() : nullptr;
        args.outCapacity = capacity;
        RunCounters reset = *hostCounters_;
        reset.voxels = 0;
        reset.contributions = 0;
        reset.clipCalls = 0;
        reset.outputOverflow = 0;
        reset.tileCursor = 0;
        *hostCountersInit_ = reset;
        O2V_CUDA(cudaMemcpyAsync(dCounters, hostCountersInit_, sizeof(RunCounters), cudaMemcpyHostToDevice, stream));
        O2V_CUDA(cudaStreamSynchronize(stream));
        memset(hostCountersInit_, 0, sizeof(RunCounters));
        for (int i = 0; i < 3; ++i) {
            hostCountersInit_->boundsMinBits[i] = 0xffffffffu;
        }
    }

    hostCounters_->clipCalls += hostCounters_->survivors;  // every sparse-path survivor is one exact clip
    floatValid_ = params.floatRecords;
    st.counters = *hostCounters_;
    st.outCapacity = capacity;
    voxelCount_ = hostCounters_->voxels;
    cudaEventElapsedTime(&st.msTotal, evStart_, evVoxEnd_);
    cudaEventElapsedTime(&st.msSetup, evStart_, evSetup_);
    cudaEventElapsedTime(&st.msVoxelize, evVoxStart_, evVoxEnd_);
    if (sparseActive) {
        cudaEventElapsedTime(&st.msClip, evClipStart_, evClipEnd_);
    }
    return kErrOk;
}

int Engine::voxelizeOccupancy(const MeshView &meshIn, const EngineParams &params, const GridView &grid,
                              cudaStream_t stream, RunStats &st)
{
    RunCounters *dCounters = counters_.as<RunCounters>();
    MeshView mesh = meshIn;
    // slabFiltered: the caller's array already is what filterSlab() keeps for this slab (multi-GPU ingest distributes the
    // triangles by z range once; a step then never touches the rest of the mesh)
    const bool partOfGrid = (grid.slabZ0 != 0 || grid.slabZ1 < grid.gridExtent) && !params.slabFiltered;
    if (partOfGrid) {
        // One rank of several: keep only the triangles whose z range can reach the slab (one streaming pass over the
        // mesh); everything after works on that share.  How many were kept stays on the device until the count pass is
        // through as well (mesh.count is their upper bound until then).
        if (!slabVerts_.ensure((size_t) meshIn.count * 9 * sizeof(float))) {
            return fail(kErrOutOfMemory, "device allocation failed (occupancy path, slab triangles)");
        }
        launchOccupancySlabFilter(meshIn, grid, slabVerts_.as<float>(), dCounters, smCount_, stream);
        ++st.kernelLaunches;
        mesh.verts = slabVerts_.as<float>();
    }
    const size_t nBound = (size_t) mesh.count;  // = meshIn.count

    OccupancyView occ{};
    // bitmaps in OUTPUT space: chunks of 64^3 output voxels (128^3 samples when downscaling)
    occ.shift = params.supersampling == 2 ? 1u : 0u;
    occ.chunksPerAxis = ((grid.gridExtent >> occ.shift) + kChunkEdge - 1) / kChunkEdge;
    occ.chunkZ0 = (grid.slabZ0 >> occ.shift) / kChunkEdge;
    const uint32_t chunkRows = ((grid.slabZ1 >> occ.shift) + kChunkEdge - 1) / kChunkEdge - occ.chunkZ0;
    occ.chunkTotal = occ.chunksPerAxis * occ.chunksPerAxis * chunkRows;  // <= 128^3
    if (!leafCount_.ensure(nBound * 4) || !leafOffset_.ensure(nBound * 4) ||
        !scratch_.ensure(scanScratchElems(nBound) * 4) || !chunkFlag_.ensure(((size_t) occ.chunkTotal + 31) / 32 * 4) ||
        !chunkSlot_.ensure((size_t) occ.chunkTotal * 4) || !chunkList_.ensure((size_t) occ.chunkTotal * 4) ||
        !leaves_.ensure(nBound * sizeof(LeafRecord))) {
        return fail(kErrOutOfMemory, "device allocation failed (occupancy path, setup buffers)");
    }
    occ.chunkFlag = chunkFlag_.as<uint32_t>();
    occ.chunkSlot = chunkSlot_.as<uint32_t>();
    occ.chunkList = chunkList_.as<uint32_t>();
    O2V_CUDA(cudaMemsetAsync(occ.chunkFlag, 0, ((size_t) occ.chunkTotal + 31) / 32 * 4, stream));

    // the one pass over the triangles: statistics, chunk marks, and the first leaf of triangle i into leaf slot i
    HugeWork huge;
    if (!hugeWork(huge)) {
        return fail(kErrOutOfMemory, "device allocation failed (huge triangles)");
    }
    launchOccupancyCount(mesh, grid, occ, leafCount_.as<uint32_t>(), leaves_.as<LeafRecord>(), dCounters, partOfGrid,
                         huge, hugeExpected_, smCount_, stream);
    st.kernelLaunches += huge.capacity != 0 ? 3 : 0;
    if (params.accumulate == 0) {
        launchOccupancyAssignChunks(occ, dCounters, stream);
    }
    else if (params.accumulate == 1) {
        launchOccupancyAllChunks(occ, stream);  // (later pieces of the job find the two tables as this one leaves them)
    }
    st.kernelLaunches += 2;
    launchPublishCounters(dCounters, hostCountersDevice_, stream);
    ++st.kernelLaunches;
    O2V_CUDA(cudaStreamSynchronize(stream));
    O2V_CUDA(cudaGetLastError());

    if (hostCounters_->hugeTriangles > huge.capacity) {
        walkHuge_ = true;  // their leaves are not counted yet: once more, with room to list them
        hugeExpected_ = hostCounters_->hugeTriangles;
        return kRetryHugeWalk;
    }
    walkHuge_ = hostCounters_->hugeTriangles != 0;
    hugeExpected_ = hostCounters_->hugeTriangles;
    if (partOfGrid) {
        mesh.count = hostCounters_->slabTriangles;
    }
    st.slabTriangles = mesh.count;
    const size_t n = (size_t) mesh.count;  // triangles the passes work on = first-leaf slots
    occ.firstLeaves = (uint32_t) n;
    const unsigned long long extraLeaves = hostCounters_->extraLeaves;
    const unsigned long long leafTotal = hostCounters_->leaves == 0 ? 0 : n + extraLeaves;  // leaf slots
    const unsigned long long candidateBound = hostCounters_->candidateVoxels;
    const unsigned long long bigLeaves = hostCounters_->bigLeaves, bigBoxes = hostCounters_->bigBoxes;
    occ.activeChunks = params.accumulate != 0 ? occ.chunkTotal : (uint32_t) hostCounters_->activeTiles;
    if (leafTotal >= (1ull << 32) || bigBoxes >= (1ull << 32) || bigLeaves >= (1ull << 24)) {
        return fail(kErrTooLarge, "more than 2^32-1 leaves or boxes in this slab");
    }
    st.occupancyPath = true;
    const size_t accumulateBitmapBytes = params.accumulate != 0 ? (size_t) occ.chunkTotal * kChunkWords * 8 : 0;
    if (params.accumulate != 0) {
        // the job checked Engine::accumulateBytes: both bitmap sets exist for all of its pieces, cleared by the first
        if (!tileBits_.ensure(accumulateBitmapBytes) || !emittedBits_.ensure(accumulateBitmapBytes)) {
            return fail(kErrOutOfMemory, "device allocation failed (bitmaps of a job in pieces)");
        }
        if (params.accumulate == 1) {
            O2V_CUDA(cudaMemsetAsync(tileBits_.as<void>(), 0, accumulateBitmapBytes, stream));
            O2V_CUDA(cudaMemsetAsync(emittedBits_.as<void>(), 0, accumulateBitmapBytes, stream));
        }
    }
    if (leafTotal == 0) {
        st.counters = *hostCounters_;
        return kErrOk;
    }

    // Output capacity: every voxel needs a candidate, and a chunk emits at most 64^3 (8 times fewer when downscaled).
    const unsigned long long perChunk = (unsigned long long) kChunkEdge * kChunkEdge * kChunkEdge;
    unsigned long long capacity = std::min(candidateBound, occ.activeChunks * perChunk);
    capacity = std::max<unsigned long long>(capacity, 1);
    // Queue of SAT-undecided voxels: a fraction of the candidates in practice (~5 %); sized at a quarter of the bound and
    // grown to the need (one rerun) in the rare case that is not enough.
    unsigned long long queueCapacity =
        std::max<unsigned long long>(std::min<unsigned long long>(candidateBound, 1ull << 20), candidateBound / 4);

    unsigned long long rangeCapacity = queueCapacity;  // one entry per row with undecided voxels
    const size_t bitmapBytes = (size_t) occ.activeChunks * kChunkWords * 8;
    if (const char *env = getenv("O2V_B200_OCCUPANCY_MAX_BYTES")) {  // test hook for the fallback below
        if (bitmapBytes > strtoull(env, nullptr, 10) && params.accumulate == 0) {
            return kOccupancyFallback;
        }
    }
    if (bitmapBytes > tileBits_.size()) {  // (never with accumulate: the bitmaps exist already)
        size_t freeBytes = 0, totalBytes = 0;
        O2V_CUDA(cudaMemGetInfo(&freeBytes, &totalBytes));
        if (bitmapBytes > (freeBytes + tileBits_.size()) / 2) {
            return kOccupancyFallback;  // e.g. a dense 8192^3 job: the caller takes the weighted path
        }
    }
    if (params.bitmapResult && !chunkCounts_.ensure((size_t) std::max(occ.activeChunks, 1u) * 4)) {
        return fail(kErrOutOfMemory, "device allocation failed (chunk counts)");
    }
    if (!tileBits_.ensure(bitmapBytes) ||
        !extraLeaves_.ensure((size_t) std::max<unsigned long long>(extraLeaves, 1) * sizeof(LeafRecord)) ||
        !occQueue_.ensure((size_t) queueCapacity * sizeof(uint4)) ||
        !occRanges_.ensure((size_t) rangeCapacity * sizeof(uint4)) ||
        !bigLeaves_.ensure((size_t) std::max<unsigned long long>(bigLeaves, 1) * sizeof(uint2))) {
        return fail(kErrOutOfMemory, "device allocation failed (occupancy path buffers)");
    }
    if (params.bitmapResult) {
        capacity = ~0ull;  // no records are written
    }
    else {
        if (capacity * sizeof(VoxelRecord) > out_.size()) {  // only when the buffer has to grow: bound it by free memory
            size_t freeBytes = 0, totalBytes = 0;
            O2V_CUDA(cudaMemGetInfo(&freeBytes, &totalBytes));
            const unsigned long long affordable = (freeBytes + out_.size()) / sizeof(VoxelRecord) * 9 / 10;
            capacity = std::min(capacity, std::max<unsigned long long>(affordable, 1));
        }
        if (!out_.ensure((size_t) capacity * sizeof(VoxelRecord))) {
            return fail(kErrOutOfMemory, "device allocation failed (voxel output)");
        }
    }
    occ.bits = tileBits_.as<unsigned long long>();
    occ.emitted = params.accumulate != 0 ? emittedBits_.as<unsigned long long>() : nullptr;
    occ.queue = occQueue_.as<uint4>();
    occ.queueCapacity = queueCapacity;
    occ.ranges = occRanges_.as<uint4>();
    occ.rangeCapacity = rangeCapacity;
    occ.bigLeaves = bigLeaves_.as<uint2>();
    occ.bigCapacity = (uint32_t) bigLeaves;

    occ.extraLeaves = extraLeaves_.as<LeafRecord>();
    if (extraLeaves != 0 || bigLeaves != 0) {
        // second pass, only when some triangle subdivides or some leaf is big (axis-aligned triangles)
        launchExclusiveScan(leafCount_.as<uint32_t>(), leafOffset_.as<uint32_t>(), n, scratch_.as<uint32_t>(),
                            &dCounters->scanTotal, stream);
        launchOccupancyEmit(mesh, grid, occ, leafOffset_.as<uint32_t>(), extraLeaves_.as<LeafRecord>(), dCounters,
                            huge, hugeExpected_, smCount_, stream);
        st.kernelLaunches += 4;
    }
    O2V_CUDA(cudaEventRecord(evSetup_, stream));

    VoxelizeArgs args{};
    args.grid = grid;
    args.leaves = leaves_.as<LeafRecord>();
    args.mesh = mesh;
    args.out = out_.as<VoxelRecord>();
    args.outCapacity = capacity;
    args.counters = dCounters;
    args.occ = occ;
    args.variant = params.variant < 0 ? 0 : params.variant;
    args.prefilter = params.prefilter;
    args.certainMargin = certainMarginFor(grid.sampleRes);
    // packed positions: 10 bits per axis while the OUTPUT chunk grid allows it (voxels up to the chunk grid survive)
    args.packedBits = !params.packedResult || params.bitmapResult ? 0 : ((grid.gridExtent >> occ.shift) <= 1024u ? 32 : 64);

    for (int attempt = 0; attempt < 4; ++attempt) {
        O2V_CUDA(cudaEventRecord(evVoxStart_, stream));
        if (params.accumulate == 0) {
            O2V_CUDA(cudaMemsetAsync(occ.bits, 0, bitmapBytes, stream));
        }
        O2V_CUDA(cudaEventRecord(evClassifyStart_, stream));
        // variant 1 / 2 force the block-per-batch / the thread-per-leaf classifier (tests: both must agree)
        const bool microLeaves = args.variant == 2 ||
                                 (args.variant != 1 && candidateBound <= kOccDirectCandidates * hostCounters_->leaves);
        launchOccupancyClassify(args, leafTotal, microLeaves, (uint32_t) bigLeaves, bigBoxes, smCount_, stream);
        O2V_CUDA(cudaEventRecord(evFilterStart_, stream));
        launchOccupancyFilterQueue(args, smCount_, stream);
        O2V_CUDA(cudaEventRecord(evClipStart_, stream));
        launchOccupancyClip(args, smCount_, stream);
        O2V_CUDA(cudaEventRecord(evClipEnd_, stream));
        if (params.bitmapResult) {
            launchOccupancyChunkCount(args.occ, chunkCounts_.as<uint32_t>(), dCounters, smCount_, stream);
        }
        else {
            launchOccupancyExpand(args, smCount_, stream);
        }
        O2V_CUDA(cudaEventRecord(evVoxEnd_, stream));
        const int launched = 4 + (bigLeaves != 0 ? 1 : 0);
        st.voxelizeLaunches += launched;
        st.kernelLaunches += launched;
        launchPublishCounters(dCounters, hostCountersDevice_, stream);
        ++st.kernelLaunches;
        O2V_CUDA(cudaStreamSynchronize(stream));
        O2V_CUDA(cudaGetLastError());
        const bool queueOverflow = hostCounters_->survivors > queueCapacity;
        const bool rangeOverflow = hostCounters_->ranges > rangeCapacity;
        if (hostCounters_->outputOverflow == 0 && !queueOverflow && !rangeOverflow) {
            break;
        }
        if (attempt == 3) {
            return fail(kErrOutOfMemory, "voxel output does not fit device memory");
        }
        if (rangeOverflow) {
            rangeCapacity = hostCounters_->ranges;  // exact: the list does not depend on the order of execution
            if (!occRanges_.ensure((size_t) rangeCapacity * sizeof(uint4))) {
                return fail(kErrOutOfMemory, "device allocation failed (occupancy range list, grown)");
            }
            args.occ.ranges = occRanges_.as<uint4>();
            args.occ.rangeCapacity = rangeCapacity;
        }
        else if (queueOverflow) {
            // what is queued depends on which bits were already visible: leave head-room, capped by the true bound
            queueCapacity = std::min(candidateBound, hostCounters_->survivors * 2 + (1ull << 20));
            if (!occQueue_.ensure((size_t) queueCapacity * sizeof(uint4))) {
                return fail(kErrOutOfMemory, "device allocation failed (occupancy queue, grown)");
            }
            args.occ.queue = occQueue_.as<uint4>();
            args.occ.queueCapacity = queueCapacity;
        }
        else {  // the exact voxel count is known now
            capacity = hostCounters_->voxels;
            if (!out_.ensure((size_t) capacity * sizeof(VoxelRecord))) {
                return fail(kErrOutOfMemory, "device allocation failed (voxel output, exact size)");
            }
            args.out = out_.as<VoxelRecord>();
            args.outCapacity = capacity;
        }
        RunCounters reset = *hostCounters_;
        reset.survivors = 0;
        reset.ranges = 0;
        reset.voxels = 0;
        reset.outputOverflow = 0;
        *hostCountersInit_ = reset;
        O2V_CUDA(cudaMemcpyAsync(dCounters, hostCountersInit_, sizeof(RunCounters), cudaMemcpyHostToDevice, stream));
        O2V_CUDA(cudaStreamSynchronize(stream));
        memset(hostCountersInit_, 0, sizeof(RunCounters));
        for (int i = 0; i < 3; ++i) {
            hostCountersInit_->boundsMinBits[i] = 0xffffffffu;
        }
    }

    if (params.accumulate != 0) {
        // what this piece delivered is what every later piece leaves out (the expand kernel read `emitted`, nobody else
        // does: the copy follows it on the stream)
        O2V_CUDA(cudaMemcpyAsync(emittedBits_.as<void>(), occ.bits, accumulateBitmapBytes, cudaMemcpyDeviceToDevice, stream));
    }
    hostCounters_->clipCalls = hostCounters_->survivors;  // every queued voxel is (at most) one exact clip
    st.counters = *hostCounters_;
    st.outCapacity = params.bitmapResult ? 0 : capacity;
    voxelCount_ = hostCounters_->voxels;
    packedBits_ = args.packedBits;
    if (params.bitmapResult) {
        bitmapValid_ = true;
        bitmap_.bits = args.occ.bits;
        bitmap_.chunkIds = args.occ.chunkList;
        bitmap_.chunkCounts = chunkCounts_.as<uint32_t>();
        bitmap_.chunks = args.occ.activeChunks;
        bitmap_.chunksPerAxis = args.occ.chunksPerAxis;
        bitmap_.chunkZ0 = args.occ.chunkZ0;
    }
    cudaEventElapsedTime(&st.msTotal, evStart_, evVoxEnd_);
    cudaEventElapsedTime(&st.msSetup, evStart_, evSetup_);
    cudaEventElapsedTime(&st.msVoxelize, evVoxStart_, evVoxEnd_);
    cudaEventElapsedTime(&st.msClassify, evClassifyStart_, evFilterStart_);
    cudaEventElapsedTime(&st.msFilter, evFilterStart_, evClipStart_);
    cudaEventElapsedTime(&st.msClip, evClipStart_, evClipEnd_);
    cudaEventElapsedTime(&st.msExpand, evClipEnd_, evVoxEnd_);
    return kErrOk;
}

int Engine::meshBounds(const MeshView &mesh, cudaStream_t stream, float outMin[3], float outMax[3])
{
    error_.clear();
    O2V_CUDA(cudaSetDevice(device_));
    RunCounters *dCounters = counters_.as<RunCounters>();
    O2V_CUDA(cudaMemcpyAsync(dCounters, hostCountersInit_, sizeof(RunCounters), cudaMemcpyHostToDevice, stream));
    launchBounds(mesh, dCounters, stream);
    launchFinishBounds(dCounters, stream);
    launchPublishCounters(dCounters, hostCountersDevice_, stream);
    O2V_CUDA(cudaStreamSynchronize(stream));
    O2V_CUDA(cudaGetLastError());
    memcpy(outMin, hostCounters_->boundsMin, 3 * sizeof(float));
    memcpy(outMax, hostCounters_->boundsMax, 3 * sizeof(float));
    return kErrOk;
}

int Engine::zRowHistogram(const MeshView &mesh, const EngineParams &params, uint32_t unit, uint32_t rows,
                          cudaStream_t stream, unsigned long long *outHistogram)
{
    error_.clear();
    O2V_CUDA(cudaSetDevice(device_));
    for (uint32_t r = 0; r < rows; ++r) {
        outHistogram[r] = 0;
    }
    if (mesh.count == 0 || rows == 0) {
        return kErrOk;
    }
    if (!params.boundsKnown || rows > 128 || unit == 0) {
        return fail(kErrBadParams, "zRowHistogram needs the mesh bounds and at most 128 rows");
    }
    RunStats st;
    GridView grid;
    bool emptySlab = false;
    EngineParams whole = params;
    whole.slabZ0 = whole.slabZ1 = 0;
    if (const int rc = setupGrid(mesh, whole, stream, st, grid, &emptySlab)) {
        return rc;
    }
    if (!scatterCounts_.ensure(128 * sizeof(unsigned long long))) {
        return fail(kErrOutOfMemory, "device allocation failed (z histogram)");
    }
    O2V_CUDA(cudaMemsetAsync(scatterCounts_.as<void>(), 0, 128 * sizeof(unsigned long long), stream));
    launchOccupancyZHistogram(mesh, grid, unit, rows, scatterCounts_.as<unsigned long long>(), smCount_, stream);
    O2V_CUDA(cudaMemcpyAsync(outHistogram, scatterCounts_.as<void>(), rows * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, stream));
    O2V_CUDA(cudaStreamSynchronize(stream));
    O2V_CUDA(cudaGetLastError());
    return kErrOk;
}

float *Engine::receiveRegion(uint32_t source, uint32_t sources, unsigned long long capacity)
{
    cudaSetDevice(device_);
    if (!received_.ensure((size_t) sources * capacity * 9 * sizeof(float))) {
        return nullptr;
    }
    return received_.as<float>() + (size_t) source * capacity * 9;
}

const float *Engine::packReceived(const unsigned long long *counts, uint32_t sources, unsigned long long capacity,
                                  cudaStream_t stream, unsigned long long *total)
{
    cudaSetDevice(device_);
    unsigned long long sum = 0;
    for (uint32_t r = 0; r < sources; ++r) {
        sum += counts[r];
    }
    *total = sum;
    // a dense copy in a buffer of its own: packing in place would overlap source and destination as soon as a slab
    // receives more than its even share (boundary triangles go to two slabs)
    if (!receivedPacked_.ensure((size_t) std::max<unsigned long long>(sum, 1) * 9 * sizeof(float))) {
        return nullptr;
    }
    unsigned long long filled = 0;
    for (uint32_t r = 0; r < sources; ++r) {
        if (counts[r] != 0) {
            cudaMemcpyAsync(receivedPacked_.as<float>() + filled * 9, received_.as<float>() + (size_t) r * capacity * 9,
                            (size_t) counts[r] * 9 * sizeof(float), cudaMemcpyDeviceToDevice, stream);
        }
        filled += counts[r];
    }
    return receivedPacked_.as<float>();
}

int Engine::scatterToSlabs(const MeshView &mesh, const EngineParams &params, SlabScatter scatter, cudaStream_t stream,
                           unsigned long long *sentCounts)
{
    error_.clear();
    O2V_CUDA(cudaSetDevice(device_));
    for (uint32_t s = 0; s < scatter.slabs; ++s) {
        sentCounts[s] = 0;
    }
    if (mesh.count == 0) {
        return kErrOk;
    }
    if (!params.boundsKnown) {
        return fail(kErrBadParams, "scatterToSlabs needs the mesh bounds");
    }
    RunStats st;
    GridView grid;
    bool emptySlab = false;
    EngineParams whole = params;
    whole.slabZ0 = whole.slabZ1 = 0;
    if (const int rc = setupGrid(mesh, whole, stream, st, grid, &emptySlab)) {
        return rc;
    }
    if (!scatterCounts_.ensure(128 * sizeof(unsigned long long))) {
        return fail(kErrOutOfMemory, "device allocation failed (scatter counters)");
    }
    O2V_CUDA(cudaMemsetAsync(scatterCounts_.as<void>(), 0, kMaxSlabs * sizeof(unsigned long long), stream));
    scatter.count = scatterCounts_.as<unsigned long long>();
    launchOccupancySlabScatter(mesh, grid, scatter, smCount_, stream);
    O2V_CUDA(cudaMemcpyAsync(sentCounts, scatterCounts_.as<void>(), scatter.slabs * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, stream));
    O2V_CUDA(cudaStreamSynchronize(stream));
    O2V_CUDA(cudaGetLastError());
    return kErrOk;
}

int Engine::filterSlab(const MeshView &mesh, const EngineParams &params, cudaStream_t stream, const float **kept,
                       unsigned long long *keptCount)
{
    error_.clear();
    *kept = nullptr;
    *keptCount = 0;
    if (params.resolution == 0 || params.supersampling == 0 || params.supersampling > 2 ||
        (unsigned long long) params.resolution * params.supersampling > 8192ull) {
        return fail(kErrBadParams, "resolution must be > 0 (sample resolution <= 8192), supersampling 1 or 2");
    }
    O2V_CUDA(cudaSetDevice(device_));
    RunCounters *dCounters = counters_.as<RunCounters>();
    O2V_CUDA(cudaMemcpyAsync(dCounters, hostCountersInit_, sizeof(RunCounters), cudaMemcpyHostToDevice, stream));
    if (mesh.count == 0) {
        return kErrOk;
    }
    RunStats st;
    GridView grid;
    bool emptySlab = false;
    if (const int rc = setupGrid(mesh, params, stream, st, grid, &emptySlab)) {
        return rc;
    }
    if (emptySlab) {
        return kErrOk;
    }
    if (!slabKept_.ensure((size_t) mesh.count * 9 * sizeof(float))) {
        return fail(kErrOutOfMemory, "device allocation failed (slab triangles)");
    }
    launchOccupancySlabFilter(mesh, grid, slabKept_.as<float>(), dCounters, smCount_, stream);
    launchPublishCounters(dCounters, hostCountersDevice_, stream);
    O2V_CUDA(cudaStreamSynchronize(stream));
    O2V_CUDA(cudaGetLastError());
    *kept = slabKept_.as<float>();
    *keptCount = hostCounters_->slabTriangles;
    return kErrOk;
}

int Engine::resultHash(cudaStream_t stream, unsigned long long *out)
{
    *out = 0;
    O2V_CUDA(cudaSetDevice(device_));
    if (!hash_.ensure(sizeof(unsigned long long))) {
        return fail(kErrOutOfMemory, "device allocation failed (record hash)");
    }
    O2V_CUDA(cudaMemsetAsync(hash_.as<void>(), 0, sizeof(unsigned long long), stream));
    if (voxelCount_ != 0) {
        launchRecordHash(out_.as<VoxelRecord>(), voxelCount_, hash_.as<unsigned long long>(), smCount_, stream);
    }
    O2V_CUDA(cudaMemcpyAsync(out, hash_.as<void>(), sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    O2V_CUDA(cudaStreamSynchronize(stream));
    return kErrOk;
}

void *Engine::pinnedStaging(int slot, size_t bytes)
{
    if (slot < 0 || slot > 1) {
        return nullptr;
    }
    if (bytes > stagingBytes_[slot]) {
        if (staging_[slot] != nullptr) {
            cudaFreeHost(staging_[slot]);
            staging_[slot] = nullptr;
            stagingBytes_[slot] = 0;
        }
        if (cudaMallocHost(&staging_[slot], bytes) != cudaSuccess) {
            cudaGetLastError();
            staging_[slot] = nullptr;
            return nullptr;
        }
        stagingBytes_[slot] = bytes;
    }
    return staging_[slot];
}

int Engine::download(void *hostDst, cudaStream_t stream)
{
    if (voxelCount_ == 0) {
        return kErrOk;
    }
    O2V_CUDA(cudaSetDevice(device_));
    O2V_CUDA(cudaMemcpyAsync(hostDst, out_.as<void>(), (size_t) voxelCount_ * sizeof(VoxelRecord),
                             cudaMemcpyDeviceToHost, stream));
    O2V_CUDA(cudaStreamSynchronize(stream));
    return kErrOk;
}

}  // namespace o2v
