// C-ABI of libobj2voxel_b200.so: the 35 reference entry points of include/obj2voxel.h (reference implementation:
// src/obj2voxel.cpp:645-1003) plus the additive bulk/device API of include/obj2voxel_b200.h, both on top of
// o2v::Engine.  Host logic only — every voxel is produced by the CUDA kernels; without a CUDA device
// obj2voxel_voxelize() logs an error and returns OBJ2VOXEL_ERR_DEVICE (there is deliberately no CPU path).
#include "obj2voxel_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "o2v_engine.h"
#include "o2v_io.h"

using namespace o2v;

// ---------------------------------------------------------------------------------------------------------------------
// logging: process-global level + optional callback (reference: voxelio log + src/obj2voxel.cpp:639-682)

namespace {

std::mutex gLogMutex;
obj2voxel_enum_t gLogLevel = OBJ2VOXEL_LOG_LEVEL_INFO;  // RELEASE_LOG_LEVEL, src/constants.hpp:21
obj2voxel_log_callback *gLogCallback = nullptr;
void *gLogCallbackData = nullptr;
thread_local std::string gLastError;

const char *levelName(obj2voxel_enum_t level)
{
    switch (level) {
    case OBJ2VOXEL_LOG_LEVEL_ERROR: return "ERROR";
    case OBJ2VOXEL_LOG_LEVEL_WARNING: return "WARNING";
    case OBJ2VOXEL_LOG_LEVEL_INFO: return "INFO";
    default: return "DEBUG";
    }
}

}  // namespace

namespace o2v {

void logMessage(unsigned char level, const std::string &message)
{
    std::lock_guard<std::mutex> lock{gLogMutex};
    if (level > gLogLevel) {
        return;
    }
    if (gLogCallback != nullptr && gLogCallback(gLogCallbackData, message.c_str(), level)) {
        return;
    }
    fprintf(stdout, "[obj2voxel_b200] [%s] %s\n", levelName(level), message.c_str());
    fflush(stdout);
}

}  // namespace o2v

namespace {

[[noreturn]] void contractViolation(const char *what)
{
    // the reference turns these into VXIO_ASSERT failures -> std::terminate (voxelio/src/assert.cpp:11-14)
    fprintf(stderr, "[obj2voxel_b200] [FAILURE] contract violation: %s\n", what);
    fflush(stderr);
    std::terminate();
}

#define O2V_REQUIRE(cond, what) \
    do {                        \
        if (!(cond)) {          \
            contractViolation(what); \
        }                       \
    } while (0)

std::string withThousands(unsigned long long n)
{
    std::string digits = std::to_string(n), out;
    for (size_t i = 0; i < digits.size(); ++i) {
        if (i != 0 && (digits.size() - i) % 3 == 0) {
            out += ',';
        }
        out += digits[i];
    }
    return out;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// API types

struct obj2voxel_texture {
    std::vector<uint8_t> pixels;
    size_t width = 0, height = 0, channels = 0;
    uint8_t wrap = OBJ2VOXEL_UV_WRAP;  // voxelio Image default WrapMode::REPEAT (image.hpp:120)
    bool loaded = false;
};

struct obj2voxel_triangle {
    float v[9];
    float t[6];
    uint8_t type;
    float color[3];
    const obj2voxel_texture *texture;
};

namespace {

enum class IoKind { MISSING, CALLBACK, FILE, MEMORY, BULK };

struct EngineDeleter {
    void operator()(Engine *e) const { delete e; }
};

std::mutex gEngineMutex;
std::unordered_map<int, std::unique_ptr<Engine, EngineDeleter>> gEngines;

/// One engine per device per process, created on first use (obj2voxel instances are throwaway objects).
Engine *sharedEngine(std::string *error)
{
    int device = 0;
    if (const char *env = getenv("O2V_B200_DEVICE")) {
        device = atoi(env);
    }
    std::lock_guard<std::mutex> lock{gEngineMutex};
    auto found = gEngines.find(device);
    if (found != gEngines.end()) {
        return found->second.get();
    }
    Engine *engine = Engine::create(device, error);
    if (engine != nullptr) {
        gEngines[device].reset(engine);
    }
    return engine;
}

void statsToC(const RunStats &in, o2v_b200_stats *out)
{
    memset(out, 0, sizeof *out);
    out->voxels = in.counters.voxels;
    out->leaves = in.counters.leaves;
    out->pairs = in.counters.pairs;
    out->active_tiles = in.counters.activeTiles;
    out->candidate_voxels = in.counters.candidateVoxels;
    out->clip_calls = in.counters.clipCalls;
    out->contributions = in.counters.contributions;
    out->dropped_triangles = in.counters.droppedTriangles;
    out->depth_overflow = in.counters.depthOverflow;
    out->out_capacity = in.outCapacity;
    out->ms_total = in.msTotal;
    out->ms_setup = in.msSetup;
    out->ms_voxelize = in.msVoxelize;
    memcpy(out->transform, in.transform, sizeof out->transform);
    out->kernel_launches = in.kernelLaunches;
    out->voxelize_launches = in.voxelizeLaunches;
    out->light_tiles = in.counters.lightTiles + in.counters.bigLightTiles;
    out->heavy_tiles = in.counters.heavyTiles;
    out->survivors = in.counters.survivors;
    out->ms_clip = in.msClip;
    out->occupancy_path = in.occupancyPath ? 1 : 0;
    out->ms_classify = in.msClassify;
    out->reserved = 0.0f;
    out->slab_triangles = in.slabTriangles;
    out->undecided_ranges = in.counters.ranges;
    out->ms_filter = in.msFilter;
    out->ms_expand = in.msExpand;
}

constexpr uint32_t kMaxJobParts = 128;  // a part is at least one 64-voxel chunk row; sample resolution <= 8192

/// How obj2voxel_voxelize() cuts a job into z parts (bounds[0 .. parts], ascending, inner bounds multiples of 64; the
/// parts tile [z0, z1) clipped to the chunk grid).  requested > 0 forces the number of parts (at most one per chunk row),
/// otherwise jobs of at least 2^20 triangles run in up to four parts and smaller ones in one.
uint32_t planJobParts(uint32_t sampleRes, uint32_t slabZ0, uint32_t slabZ1, unsigned long long triangles, int requested,
                      uint32_t *bounds)
{
    const uint32_t gridExtent = (sampleRes + 63u) / 64u * 64u;
    uint32_t jobZ0 = slabZ0, jobZ1 = slabZ1;
    if (jobZ0 == 0 && jobZ1 == 0) {
        jobZ1 = gridExtent;
    }
    jobZ1 = std::min(jobZ1, gridExtent);
    jobZ0 = std::min(jobZ0, jobZ1);
    const uint32_t row0 = jobZ0 / 64u, row1 = std::max((jobZ1 + 63u) / 64u, row0 + 1u);
    const uint32_t rows = row1 - row0;
    uint32_t parts = (triangles >= (1ull << 20) && rows > 1u) ? std::min(4u, rows) : 1u;
    if (requested > 0) {
        parts = std::min(std::min((uint32_t) requested, rows), kMaxJobParts);
    }
    for (uint32_t k = 0; k <= parts; ++k) {
        const uint32_t z = (row0 + (uint32_t) ((unsigned long long) rows * k / parts)) * 64u;
        bounds[k] = std::min(std::max(z, jobZ0), jobZ1);
    }
    return parts;
}

/// Totals of a job that ran as several z parts (every voxel, leaf and clip belongs to exactly one part; a dropped
/// triangle may be seen by several parts: the largest count is reported).
void accumulateStats(RunStats &total, const RunStats &part)
{
    RunCounters &t = total.counters;
    const RunCounters &p = part.counters;
    t.voxels += p.voxels;
    t.leaves += p.leaves;
    t.pairs += p.pairs;
    t.activeTiles += p.activeTiles;
    t.candidateVoxels += p.candidateVoxels;
    t.clipCalls += p.clipCalls;
    t.contributions += p.contributions;
    t.droppedTriangles = std::max(t.droppedTriangles, p.droppedTriangles);
    t.depthOverflow = std::max(t.depthOverflow, p.depthOverflow);
    t.lightTiles += p.lightTiles;
    t.bigLightTiles += p.bigLightTiles;
    t.heavyTiles += p.heavyTiles;
    t.survivors += p.survivors;
    total.outCapacity = std::max(total.outCapacity, part.outCapacity);
    total.msTotal += part.msTotal;
    total.msSetup += part.msSetup;
    total.msVoxelize += part.msVoxelize;
    total.msClip += part.msClip;
    total.msClassify += part.msClassify;
    total.msExpand += part.msExpand;
    total.msFilter += part.msFilter;
    t.ranges += p.ranges;
    total.kernelLaunches += part.kernelLaunches;
    total.voxelizeLaunches += part.voxelizeLaunches;
    total.occupancyPath = total.occupancyPath && part.occupancyPath;
}

EngineParams paramsFromC(const o2v_b200_params &p)
{
    EngineParams e;
    e.resolution = p.resolution;
    e.supersampling = p.supersampling;
    e.strategy = static_cast<uint8_t>(p.strategy);
    e.boundsKnown = p.bounds_known != 0;
    memcpy(e.bounds, p.bounds, sizeof e.bounds);
    for (int i = 0; i < 9; ++i) {
        e.unitTransform[i] = p.unit_transform[i];
    }
    e.slabZ0 = p.slab_z0;
    e.slabZ1 = p.slab_z1;
    e.variant = p.variant;
    e.prefilter = p.prefilter;
    e.occupancyPath = p.occupancy_path;
    e.slabFiltered = p.slab_filtered != 0;
    return e;
}

/// Device copies of a host mesh + textures for one run.
struct UploadedMesh {
    DeviceBuffer verts, uvs, types, colors, textureIds;
    std::vector<std::unique_ptr<DeviceBuffer>> texturePixels;
    std::vector<TextureView> textureViews;
    MeshView view{};

    bool upload(const o2v_b200_mesh &mesh, const o2v_b200_texture *textures, uint32_t textureCount, cudaStream_t stream,
                std::string *error)
    {
        const size_t n = static_cast<size_t>(mesh.count);
        auto copy = [&](DeviceBuffer &dst, const void *src, size_t bytes) -> bool {
            if (src == nullptr || bytes == 0) {
                return true;
            }
            if (!dst.ensure(bytes)) {
                *error = "device allocation failed (mesh upload)";
                return false;
            }
            if (cudaMemcpyAsync(dst.as<void>(), src, bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) {
                *error = std::string("mesh upload failed: ") + cudaGetErrorString(cudaGetLastError());
                return false;
            }
            return true;
        };
        if (!copy(verts, mesh.verts, n * 9 * sizeof(float)) || !copy(uvs, mesh.uvs, n * 6 * sizeof(float)) ||
            !copy(types, mesh.types, n) || !copy(colors, mesh.colors, n * 3 * sizeof(float)) ||
            !copy(textureIds, mesh.texture_ids, n * sizeof(uint32_t))) {
            return false;
        }
        view.verts = mesh.verts != nullptr ? verts.as<float>() : nullptr;
        view.uvs = mesh.uvs != nullptr ? uvs.as<float>() : nullptr;
        view.types = mesh.types != nullptr ? types.as<uint8_t>() : nullptr;
        view.colors = mesh.colors != nullptr ? colors.as<float>() : nullptr;
        view.textureIds = mesh.texture_ids != nullptr ? textureIds.as<uint32_t>() : nullptr;
        view.count = mesh.count;
        for (uint32_t i = 0; i < textureCount; ++i) {
            texturePixels.emplace_back(new DeviceBuffer());
            const size_t bytes = (size_t) textures[i].width * textures[i].height * textures[i].channels;
            if (!copy(*texturePixels.back(), textures[i].pixels, bytes)) {
                return false;
            }
            textureViews.push_back(TextureView{texturePixels.back()->as<uint8_t>(), textures[i].width,
                                               textures[i].height, textures[i].channels, textures[i].wrap});
        }
        return true;
    }
};

}  // namespace

struct obj2voxel_instance {
    // configuration (mirrors reference src/obj2voxel.cpp:142-173)
    IoKind inputKind = IoKind::MISSING;
    IoKind outputKind = IoKind::MISSING;
    obj2voxel_triangle_callback *inputCallback = nullptr;
    void *inputCallbackData = nullptr;
    obj2voxel_voxel_callback *outputCallback = nullptr;
    void *outputCallbackData = nullptr;
    const char *inputFile = nullptr;  // not copied, like the reference (src/obj2voxel.cpp:719)
    const char *outputFile = nullptr;
    FileFormat inputFormat = FileFormat::UNKNOWN;
    FileFormat outputFormat = FileFormat::UNKNOWN;
    obj2voxel_texture *defaultTexture = nullptr;

    const float *bulkVerts = nullptr;
    const float *bulkUvs = nullptr;
    size_t bulkCount = 0;
    obj2voxel_texture *bulkTexture = nullptr;

    float bounds[6] = {0, 0, 0, 0, 0, 0};
    bool boundsKnown = false;
    uint8_t strategy = OBJ2VOXEL_MAX_STRATEGY;
    uint32_t outputResolution = 0;
    uint32_t supersampling = 1;
    bool parallel = false;
    int unitTransform[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    uint32_t slabZ0 = 0, slabZ1 = 0;

    // run state
    bool done = false;
    std::unique_ptr<VoxelSink> sink;
    o2v_b200_stats stats{};

    // worker bookkeeping (reference src/obj2voxel.cpp:957-1003): workers only block; the GPU does the work
    std::mutex workerMutex;
    std::condition_variable workerWake;
    uint32_t workerCount = 0;
    bool workersStopped = false;
};

// ---------------------------------------------------------------------------------------------------------------------
// the job: reference src/obj2voxel.cpp:578-637

namespace {

struct HostMesh {
    std::vector<float> verts, uvs, colors;
    std::vector<uint8_t> types;
    std::vector<uint32_t> textureIds;
    std::vector<const obj2voxel_texture *> textures;
    bool anyTextured = false, anyColored = false;

    uint32_t textureIndex(const obj2voxel_texture *texture)
    {
        for (size_t i = 0; i < textures.size(); ++i) {
            if (textures[i] == texture) {
                return static_cast<uint32_t>(i);
            }
        }
        textures.push_back(texture);
        return static_cast<uint32_t>(textures.size() - 1);
    }

    void push(const obj2voxel_triangle &t)
    {
        verts.insert(verts.end(), t.v, t.v + 9);
        uvs.insert(uvs.end(), t.t, t.t + 6);
        colors.insert(colors.end(), t.color, t.color + 3);
        uint8_t type = t.type;
        uint32_t id = 0;
        if (type == kTextured) {
            if (t.texture != nullptr && t.texture->loaded) {
                id = textureIndex(t.texture);
                anyTextured = true;
            }
            else {
                type = kMaterialless;
            }
        }
        anyColored |= type == kUntextured;
        types.push_back(type);
        textureIds.push_back(id);
    }
};

obj2voxel_error_t runJob(obj2voxel_instance &inst)
{
    // ---- gather the triangle stream into flat arrays (reference: "Caching triangles", obj2voxel.cpp:583-588) ----
    HostMesh host;
    o2v_b200_mesh mesh{};
    std::vector<o2v_b200_texture> textures;

    if (inst.inputKind == IoKind::BULK) {
        mesh.verts = inst.bulkVerts;
        mesh.count = inst.bulkCount;
        if (inst.bulkUvs != nullptr && inst.bulkTexture != nullptr && inst.bulkTexture->loaded) {
            mesh.uvs = inst.bulkUvs;
            host.textures.push_back(inst.bulkTexture);
        }
    }
    else {
        if (inst.inputKind == IoKind::CALLBACK) {
            obj2voxel_triangle triangle{};
            triangle.type = kMaterialless;
            while (inst.inputCallback(inst.inputCallbackData, &triangle)) {
                host.push(triangle);
            }
        }
        else {
            std::string error;
            TriangleAppender appender = [&](const float v[9], const float uv[6], uint8_t type, const float color[3],
                                            const obj2voxel_texture *texture) {
                obj2voxel_triangle t{};
                memcpy(t.v, v, sizeof t.v);
                if (uv != nullptr) {
                    memcpy(t.t, uv, sizeof t.t);
                }
                t.type = type;
                if (color != nullptr) {
                    memcpy(t.color, color, sizeof t.color);
                }
                t.texture = texture;
                host.push(t);
            };
            if (!readTriangleFile(inst.inputFile, inst.inputFormat, inst.defaultTexture, appender, &error)) {
                logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Failed to open input: " + error);
                return OBJ2VOXEL_ERR_IO_ERROR_ON_OPEN_INPUT_FILE;
            }
        }
        mesh.verts = host.verts.data();
        mesh.count = host.types.size();
        if (host.anyTextured) {
            mesh.uvs = host.uvs.data();
            mesh.texture_ids = host.textureIds.data();
        }
        if (host.anyTextured || host.anyColored) {
            mesh.types = host.types.data();
        }
        if (host.anyColored) {
            mesh.colors = host.colors.data();
        }
    }
    for (const obj2voxel_texture *t : host.textures) {
        textures.push_back(o2v_b200_texture{t->pixels.data(), (uint32_t) t->width, (uint32_t) t->height,
                                            (uint32_t) t->channels, t->wrap});
    }

    if (mesh.count == 0) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_WARNING, "Model has no triangles, aborting and writing empty voxel model");
        inst.sink->finalize();
        return inst.sink->good() ? OBJ2VOXEL_ERR_OK : OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE;
    }
    logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "Cached model with " + withThousands(mesh.count) + " triangles");

    // ---- device run ----
    std::string error;
    Engine *engine = sharedEngine(&error);
    if (engine == nullptr) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Cannot voxelize: " + error);
        return OBJ2VOXEL_ERR_DEVICE;
    }
    cudaSetDevice(engine->device());
    cudaStream_t stream = nullptr;
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, std::string("cudaStreamCreate failed: ") +
                                                  cudaGetErrorString(cudaGetLastError()));
        return OBJ2VOXEL_ERR_DEVICE;
    }
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } guard{stream};

    EngineParams params;
    params.resolution = inst.outputResolution;
    params.supersampling = inst.supersampling;
    params.strategy = inst.strategy;
    params.boundsKnown = inst.boundsKnown;
    memcpy(params.bounds, inst.bounds, sizeof params.bounds);
    memcpy(params.unitTransform, inst.unitTransform, sizeof params.unitTransform);
    params.slabZ0 = inst.slabZ0;
    params.slabZ1 = inst.slabZ1;
    if (const char *env = getenv("O2V_B200_VARIANT")) {
        params.variant = atoi(env);
    }
    if (const char *env = getenv("O2V_B200_PREFILTER")) {
        params.prefilter = atoi(env);
    }
    if (const char *env = getenv("O2V_B200_OCCUPANCY_PATH")) {
        params.occupancyPath = atoi(env);
    }

    if (inst.supersampling > 1) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "Chunks will be downscaled from " +
                                                 withThousands(inst.outputResolution * inst.supersampling) +
                                                 " to output resolution " + withThousands(inst.outputResolution) +
                                                 " ...");
    }

    RunStats stats;
    const auto tJob = std::chrono::steady_clock::now();
    auto msSince = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    {
        std::lock_guard<std::mutex> lock{gEngineMutex};  // one job at a time per process-wide engine
        // the mesh staging buffers are kept per device across jobs (grow-only) like the engine's own buffers
        static std::unordered_map<int, std::unique_ptr<UploadedMesh>> uploads;
        std::unique_ptr<UploadedMesh> &uploadSlot = uploads[engine->device()];
        if (uploadSlot == nullptr) {
            uploadSlot.reset(new UploadedMesh());
        }
        UploadedMesh &uploaded = *uploadSlot;
        uploaded.texturePixels.clear();
        uploaded.textureViews.clear();
        if (!uploaded.upload(mesh, textures.data(), (uint32_t) textures.size(), stream, &error)) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, error);
            return OBJ2VOXEL_ERR_DEVICE;
        }
        cudaStreamSynchronize(stream);
        const double msUpload = msSince(tJob);
        const auto tRun = std::chrono::steady_clock::now();

        // ---- plan: a big job runs as up to four z sub-slabs of whole 64-voxel chunk rows, so that the download of one
        // part (PCIe, the longest leg of a host-to-host job) runs under the kernels of the next.  Every voxel belongs to
        // exactly one part: same records, part by part. ----
        uint32_t partBounds[kMaxJobParts + 1];
        const char *partsEnv = getenv("O2V_B200_PIPELINE_PARTS");
        const uint32_t parts = planJobParts(inst.outputResolution * inst.supersampling, inst.slabZ0, inst.slabZ1, mesh.count,
                                            partsEnv != nullptr ? atoi(partsEnv) : 0, partBounds);
        auto partBound = [&](uint32_t k) { return partBounds[k]; };

        cudaStream_t copyStream = nullptr;
        cudaEvent_t copied[2] = {nullptr, nullptr};
        if (cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&copied[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&copied[1], cudaEventDisableTiming) != cudaSuccess) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, std::string("cudaStreamCreate failed: ") +
                                                      cudaGetErrorString(cudaGetLastError()));
            return OBJ2VOXEL_ERR_DEVICE;
        }
        struct CopyGuard {
            cudaStream_t s;
            cudaEvent_t *e;
            ~CopyGuard()
            {
                cudaStreamSynchronize(s);
                cudaStreamDestroy(s);
                cudaEventDestroy(e[0]);
                cudaEventDestroy(e[1]);
            }
        } copyGuard{copyStream, copied};

        const size_t batch = 1u << 21;  // records per sink call (32 MiB)
        bool sinkOk = true, deviceOk = true;
        double msKernels = 0;
        bool firstPart = true;

        // What is still on its way to the sink: a part whose records are being copied into pinned buffer `slot`.
        struct Pending {
            bool active = false;
            int slot = 0;
            uint32_t *records = nullptr;
            unsigned long long count = 0;
        } pending;
        auto deliver = [&](Pending &p) {  // waits for the copy, then hands the records to the sink batch by batch
            if (!p.active) {
                return;
            }
            p.active = false;
            deviceOk = deviceOk && cudaEventSynchronize(copied[p.slot]) == cudaSuccess;
            for (unsigned long long done = 0; done < p.count && sinkOk && deviceOk; done += batch) {
                sinkOk = inst.sink->write(p.records + done * 4, (size_t) std::min<unsigned long long>(batch, p.count - done));
            }
        };
        // Fallback for a part whose records do not fit a pinned buffer: two 32 MiB staging buffers, the copy of batch
        // k+1 under the sink call of batch k (also the whole story of a job that runs as one part).
        auto streamOut = [&](const unsigned char *deviceRecords, unsigned long long total) {
            uint32_t *staging[2] = {static_cast<uint32_t *>(engine->pinnedStaging(0, batch * 16)),
                                    static_cast<uint32_t *>(engine->pinnedStaging(1, batch * 16))};
            std::vector<uint32_t> pageable;
            if (staging[0] == nullptr || staging[1] == nullptr) {  // pinned memory exhausted: plain host memory still works
                pageable.resize(batch * 4 * 2);
                staging[0] = pageable.data();
                staging[1] = pageable.data() + batch * 4;
            }
            auto startCopy = [&](unsigned long long done, int slot) -> bool {
                const size_t count = (size_t) std::min<unsigned long long>(batch, total - done);
                return cudaMemcpyAsync(staging[slot], deviceRecords + done * 16, count * 16, cudaMemcpyDeviceToHost,
                                       copyStream) == cudaSuccess &&
                       cudaEventRecord(copied[slot], copyStream) == cudaSuccess;
            };
            int slot = 0;
            deviceOk = deviceOk && startCopy(0, 0);
            for (unsigned long long done = 0; done < total && sinkOk && deviceOk; done += batch, slot ^= 1) {
                const size_t count = (size_t) std::min<unsigned long long>(batch, total - done);
                if (done + batch < total) {
                    deviceOk = startCopy(done + batch, slot ^ 1);
                }
                deviceOk = deviceOk && cudaEventSynchronize(copied[slot]) == cudaSuccess;
                if (deviceOk) {
                    sinkOk = inst.sink->write(staging[slot], count);
                }
            }
            cudaStreamSynchronize(copyStream);
        };

        for (uint32_t k = 0; k < parts && sinkOk && deviceOk; ++k) {
            EngineParams partParams = params;
            if (parts > 1) {
                partParams.slabZ0 = partBound(k);
                partParams.slabZ1 = partBound(k + 1);
                if (partParams.slabZ0 >= partParams.slabZ1) {
                    continue;
                }
            }
            RunStats partStats;
            const int rc = engine->voxelize(uploaded.view, uploaded.textureViews.data(),
                                            (uint32_t) uploaded.textureViews.size(), partParams, stream, &partStats);
            if (rc != 0) {
                logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Voxelization failed on the device: " + engine->lastError());
                return OBJ2VOXEL_ERR_DEVICE;
            }
            msKernels += partStats.msTotal;
            if (firstPart) {
                stats = partStats;
                firstPart = false;
            }
            else {
                accumulateStats(stats, partStats);
            }
            const unsigned long long total = engine->voxelCount();
            const auto *deviceRecords = reinterpret_cast<const unsigned char *>(engine->deviceVoxels());
            if (parts == 1) {
                if (total != 0) {
                    streamOut(deviceRecords, total);
                }
                break;
            }
            // this part's records start their way to the host before the previous part is handed to the sink, and stay
            // in their device buffer while the next part writes the other one
            const int slot = (int) (k & 1u);
            uint32_t *host = total != 0 && total * 16 <= (1ull << 30)
                                 ? static_cast<uint32_t *>(engine->pinnedStaging(slot, (size_t) total * 16))
                                 : nullptr;
            Pending mine;
            if (host != nullptr) {
                deviceOk = deviceOk &&
                           cudaMemcpyAsync(host, deviceRecords, (size_t) total * 16, cudaMemcpyDeviceToHost, copyStream) ==
                               cudaSuccess &&
                           cudaEventRecord(copied[slot], copyStream) == cudaSuccess;
                mine.active = true;
                mine.slot = slot;
                mine.records = host;
                mine.count = total;
            }
            deliver(pending);
            if (host == nullptr && total != 0) {
                streamOut(deviceRecords, total);  // too big to pin in one piece (or pinning failed)
            }
            pending = mine;
            engine->swapOutputBuffers();
        }
        deliver(pending);
        cudaStreamSynchronize(copyStream);
        const double msRun = msSince(tRun);
        statsToC(stats, &inst.stats);
        if (!deviceOk) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR,
                       std::string("voxel download failed: ") + cudaGetErrorString(cudaGetLastError()));
            return OBJ2VOXEL_ERR_DEVICE;
        }
        if (!sinkOk || !inst.sink->good()) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Voxelization failed because of IO error");
            return OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE;
        }
        char timing[160];
        snprintf(timing, sizeof timing, "timing: upload %.2f ms, %u part(s): kernels %.2f ms, run + download + sink %.2f ms",
                 msUpload, parts, msKernels, msRun);
        logMessage(OBJ2VOXEL_LOG_LEVEL_DEBUG, timing);
    }

    logMessage(OBJ2VOXEL_LOG_LEVEL_INFO,
               "Voxelized " + withThousands(mesh.count) + " triangles, writing any buffered voxels ...");
    inst.sink->finalize();
    if (!inst.sink->good()) {
        return OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE;
    }
    logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "All " + withThousands(inst.sink->voxelsWritten()) + " voxels written");
    return OBJ2VOXEL_ERR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// extern "C": reference API

extern "C" {

obj2voxel_instance *obj2voxel_alloc(void)
{
    return new obj2voxel_instance;
}

void obj2voxel_free(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "obj2voxel_free(NULL)");
    delete instance;
}

void obj2voxel_set_log_level(obj2voxel_enum_t level)
{
    std::lock_guard<std::mutex> lock{gLogMutex};
    gLogLevel = level > OBJ2VOXEL_LOG_LEVEL_DEBUG ? OBJ2VOXEL_LOG_LEVEL_DEBUG : level;
}

obj2voxel_enum_t obj2voxel_get_log_level(void)
{
    std::lock_guard<std::mutex> lock{gLogMutex};
    return gLogLevel;
}

void obj2voxel_set_log_callback(obj2voxel_log_callback *callback, void *callback_data)
{
    // NULL resets to stdout (the documented behaviour; the reference's implementation would call a null pointer,
    // SURVEY Appendix B8)
    std::lock_guard<std::mutex> lock{gLogMutex};
    gLogCallback = callback;
    gLogCallbackData = callback_data;
}

void obj2voxel_set_resolution(obj2voxel_instance *instance, uint32_t resolution)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    O2V_REQUIRE(resolution != 0, "resolution must not be 0");
    instance->outputResolution = resolution;
}

void obj2voxel_set_supersampling(obj2voxel_instance *instance, uint32_t level)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    O2V_REQUIRE(level != 0, "supersampling level must not be 0");
    O2V_REQUIRE(level < 3, "supersampling level must be 1 or 2");  // src/obj2voxel.cpp:275
    instance->supersampling = level;
}

void obj2voxel_set_color_strategy(obj2voxel_instance *instance, obj2voxel_enum_t strategy)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    O2V_REQUIRE(strategy < 2, "unknown color strategy");
    instance->strategy = strategy;
}

void obj2voxel_set_texture(obj2voxel_instance *instance, obj2voxel_texture *texture)
{
    O2V_REQUIRE(instance != nullptr && texture != nullptr, "instance or texture is NULL");
    instance->defaultTexture = texture;
}

void obj2voxel_set_input_file(obj2voxel_instance *instance, const char *file, const char *type)
{
    O2V_REQUIRE(instance != nullptr && file != nullptr, "instance or file is NULL");
    const FileFormat format = detectFormat(file, type);
    O2V_REQUIRE(format != FileFormat::UNKNOWN, "input file has no recognizable extension");
    instance->inputKind = IoKind::FILE;
    instance->inputFile = file;
    instance->inputFormat = format;
}

void obj2voxel_set_input_callback(obj2voxel_instance *instance, obj2voxel_triangle_callback *callback,
                                  void *callback_data)
{
    O2V_REQUIRE(instance != nullptr && callback != nullptr, "instance or callback is NULL");
    instance->inputKind = IoKind::CALLBACK;
    instance->inputCallback = callback;
    instance->inputCallbackData = callback_data;
}

void obj2voxel_set_output_file(obj2voxel_instance *instance, const char *file, const char *type)
{
    O2V_REQUIRE(instance != nullptr && file != nullptr, "instance or file is NULL");
    const FileFormat format = detectFormat(file, type);
    O2V_REQUIRE(format != FileFormat::UNKNOWN, "output file has no recognizable extension");
    instance->outputKind = IoKind::FILE;
    instance->outputFile = file;
    instance->outputFormat = format;
}

void obj2voxel_set_output_memory(obj2voxel_instance *instance, const char *type)
{
    O2V_REQUIRE(instance != nullptr && type != nullptr, "instance or type is NULL");
    const FileFormat format = detectFormat(nullptr, type);
    O2V_REQUIRE(format != FileFormat::UNKNOWN, "not a recognized file extension");
    instance->outputKind = IoKind::MEMORY;
    instance->outputFile = nullptr;
    instance->outputFormat = format;
}

void obj2voxel_set_output_callback(obj2voxel_instance *instance, obj2voxel_voxel_callback *callback,
                                   void *callback_data)
{
    O2V_REQUIRE(instance != nullptr && callback != nullptr, "instance or callback is NULL");
    instance->outputKind = IoKind::CALLBACK;
    instance->outputCallback = callback;
    instance->outputCallbackData = callback_data;
}

void obj2voxel_set_parallel(obj2voxel_instance *instance, bool enabled)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    instance->parallel = enabled;
}

void obj2voxel_set_unit_transform(obj2voxel_instance *instance, const int transform[9])
{
    O2V_REQUIRE(instance != nullptr && transform != nullptr, "instance or transform is NULL");
    memcpy(instance->unitTransform, transform, sizeof instance->unitTransform);
}

void obj2voxel_set_mesh_boundaries(obj2voxel_instance *instance, const float bounds[6])
{
    O2V_REQUIRE(instance != nullptr && bounds != nullptr, "instance or bounds is NULL");
    for (int i = 0; i < 6; ++i) {
        O2V_REQUIRE(isfinite(bounds[i]), "infinite mesh boundaries provided");
    }
    for (int i = 0; i < 3; ++i) {
        O2V_REQUIRE(bounds[i] <= bounds[i + 3], "lower mesh bound must be <= the maximum on each axis");
    }
    memcpy(instance->bounds, bounds, sizeof instance->bounds);
    instance->boundsKnown = true;
}

uint32_t obj2voxel_get_resolution(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    return instance->outputResolution;
}

uint32_t obj2voxel_get_chunk_size(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    return 64;  // src/constants.hpp:10; Z-slabs handed to GPUs are multiples of it
}

const obj2voxel_byte_t *obj2voxel_get_output_memory(obj2voxel_instance *instance, size_t *out_size)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    O2V_REQUIRE(instance->sink != nullptr, "accessing output memory before voxelization");
    if (instance->outputKind != IoKind::MEMORY) {
        return nullptr;
    }
    const std::vector<uint8_t> *bytes = instance->sink->memory();
    O2V_REQUIRE(bytes != nullptr, "memory sink without a byte array");
    *out_size = bytes->size();
    return bytes->data();
}

void obj2voxel_set_triangle_basic(obj2voxel_triangle *triangle, const float vertices[9])
{
    triangle->type = kMaterialless;
    memcpy(triangle->v, vertices, sizeof triangle->v);
}

void obj2voxel_set_triangle_colored(obj2voxel_triangle *triangle, const float vertices[9], const float color[3])
{
    triangle->type = kMaterialless;  // sic: reference src/obj2voxel.cpp:832
    memcpy(triangle->v, vertices, sizeof triangle->v);
    memcpy(triangle->color, color, sizeof triangle->color);
}

void obj2voxel_set_triangle_textured(obj2voxel_triangle *triangle, const float vertices[9], const float textures[6],
                                     obj2voxel_texture *texture)
{
    triangle->type = kTextured;
    memcpy(triangle->v, vertices, sizeof triangle->v);
    memcpy(triangle->t, textures, sizeof triangle->t);
    triangle->texture = texture;
}

obj2voxel_texture *obj2voxel_texture_alloc(void)
{
    return new obj2voxel_texture;
}

void obj2voxel_texture_free(obj2voxel_texture *texture)
{
    O2V_REQUIRE(texture != nullptr, "texture is NULL");
    delete texture;
}

bool obj2voxel_texture_load_from_file(obj2voxel_texture *texture, const char *file, const char *type)
{
    O2V_REQUIRE(texture != nullptr && file != nullptr, "texture or file is NULL");
    if (detectFormat(file, type) != FileFormat::PNG) {
        return false;
    }
    std::vector<uint8_t> bytes;
    if (!readWholeFile(file, &bytes)) {
        return false;
    }
    return obj2voxel_texture_load_from_memory(texture, bytes.data(), bytes.size(), "png");
}

bool obj2voxel_texture_load_from_memory(obj2voxel_texture *texture, const obj2voxel_byte_t *data, size_t size,
                                        const char *type)
{
    O2V_REQUIRE(texture != nullptr && data != nullptr, "texture or data is NULL");
    if (detectFormat(nullptr, type) != FileFormat::PNG) {
        return false;
    }
    // the reference decodes to four channels (png::decode(.., 4, ..), src/obj2voxel.cpp:873-905)
    std::vector<uint8_t> rgba;
    size_t w = 0, h = 0;
    std::string error;
    if (!decodePng(data, size, &rgba, &w, &h, &error)) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_WARNING, "PNG decode failed: " + error);
        return false;
    }
    texture->pixels = std::move(rgba);
    texture->width = w;
    texture->height = h;
    texture->channels = 5;  // RGBA32 order (r,g,b,a): distinguished from load_pixels' 4 = "ARGB32"
    texture->loaded = true;
    return true;
}

bool obj2voxel_texture_load_pixels(obj2voxel_texture *texture, const obj2voxel_byte_t *pixels, size_t width,
                                   size_t height, size_t channels)
{
    O2V_REQUIRE(texture != nullptr && pixels != nullptr, "texture or pixels is NULL");
    O2V_REQUIRE(channels == 3 || channels == 4, "channels must be 3 or 4");  // colorFormatOfChannelCount, :349-356
    texture->pixels.assign(pixels, pixels + width * height * channels);
    texture->width = width;
    texture->height = height;
    texture->channels = channels;
    texture->loaded = true;
    return true;
}

void obj2voxel_teture_set_uv_mode(obj2voxel_texture *texture, obj2voxel_enum_t mode)
{
    O2V_REQUIRE(texture != nullptr && texture->loaded, "can't set UV mode of empty texture");
    texture->wrap = mode == OBJ2VOXEL_UV_CLAMP ? OBJ2VOXEL_UV_CLAMP : OBJ2VOXEL_UV_WRAP;
}

void obj2voxel_texture_get_meta(obj2voxel_texture *texture, size_t *out_width, size_t *out_height,
                                size_t *out_channels)
{
    O2V_REQUIRE(texture != nullptr && texture->loaded, "can't get metadata of empty image");
    *out_width = texture->width;
    *out_height = texture->height;
    *out_channels = texture->channels == 5 ? 4 : texture->channels;
}

void obj2voxel_texture_get_pixels(obj2voxel_texture *texture, obj2voxel_byte_t *out_pixels)
{
    O2V_REQUIRE(texture != nullptr && out_pixels != nullptr && texture->loaded, "can't get pixels of empty image");
    memcpy(out_pixels, texture->pixels.data(), texture->pixels.size());
}

obj2voxel_error_t obj2voxel_voxelize(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    obj2voxel_instance &inst = *instance;
    // argument checks in the reference's order: src/obj2voxel.cpp:604-618
    if (inst.done) {
        return OBJ2VOXEL_ERR_DOUBLE_VOXELIZATION;
    }
    if (inst.inputKind == IoKind::MISSING) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "No input was specified");
        return OBJ2VOXEL_ERR_NO_INPUT;
    }
    if (inst.outputKind == IoKind::MISSING) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "No output was specified");
        return OBJ2VOXEL_ERR_NO_OUTPUT;
    }
    if (inst.outputResolution == 0) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "No resolution was specified");
        return OBJ2VOXEL_ERR_NO_RESOLUTION;
    }
    if (inst.inputKind == IoKind::FILE && !canReadTriangles(inst.inputFormat)) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Unsupported input file type");
        return OBJ2VOXEL_ERR_IO_ERROR_ON_OPEN_INPUT_FILE;
    }

    std::string error;
    if (inst.outputKind == IoKind::CALLBACK) {
        inst.sink = makeCallbackSink(inst.outputCallback, inst.outputCallbackData);
    }
    else {
        inst.sink = makeFormatSink(inst.outputFormat, inst.outputKind == IoKind::MEMORY ? nullptr : inst.outputFile,
                                   inst.outputResolution, &error);
    }
    if (inst.sink == nullptr) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Failed to open output: " + error);
        return OBJ2VOXEL_ERR_IO_ERROR_ON_OPEN_OUTPUT_FILE;
    }

    const obj2voxel_error_t result = runJob(inst);
    if (inst.outputKind != IoKind::MEMORY) {
        inst.sink.reset();  // src/obj2voxel.cpp:631-633: only memory sinks outlive the job
    }
    inst.done = true;
    return result;
}

void obj2voxel_run_worker(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    std::unique_lock<std::mutex> lock{instance->workerMutex};
    if (instance->workersStopped) {
        return;
    }
    ++instance->workerCount;
    // The reference's workers pull 64^3-chunk commands from a ring buffer (src/obj2voxel.cpp:970-985); here the chunks
    // are GPU tiles, so a worker only has to honour the blocking contract until obj2voxel_stop_workers().
    instance->workerWake.wait(lock, [instance] { return instance->workersStopped; });
}

void obj2voxel_stop_workers(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    {
        std::lock_guard<std::mutex> lock{instance->workerMutex};
        instance->workersStopped = true;
        instance->workerCount = 0;
    }
    instance->workerWake.notify_all();
}

uint32_t obj2voxel_get_worker_count(obj2voxel_instance *instance)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    std::lock_guard<std::mutex> lock{instance->workerMutex};
    return instance->workerCount;
}

// ---------------------------------------------------------------------------------------------------------------------
// extern "C": additive API (include/obj2voxel_b200.h)

struct o2v_b200_engine {
    Engine *engine;
};

o2v_b200_engine *o2v_b200_engine_create(int device)
{
    std::string error;
    Engine *engine = Engine::create(device, &error);
    if (engine == nullptr) {
        gLastError = error;
        return nullptr;
    }
    return new o2v_b200_engine{engine};
}

void o2v_b200_engine_destroy(o2v_b200_engine *engine)
{
    if (engine != nullptr) {
        delete engine->engine;
        delete engine;
    }
}

const char *o2v_b200_last_error(void)
{
    return gLastError.c_str();
}

int o2v_b200_sm_count(const o2v_b200_engine *engine)
{
    return engine->engine->smCount();
}

void o2v_b200_default_params(o2v_b200_params *params)
{
    memset(params, 0, sizeof *params);
    params->supersampling = 1;
    params->unit_transform[0] = params->unit_transform[4] = params->unit_transform[8] = 1;
    params->variant = -1;
    params->prefilter = 1;
    params->occupancy_path = 1;
}

int o2v_b200_voxelize_device(o2v_b200_engine *engine, const o2v_b200_params *params, const o2v_b200_mesh *mesh,
                             const o2v_b200_texture *textures, uint32_t texture_count, void *cuda_stream,
                             o2v_b200_stats *out_stats)
{
    MeshView view;
    view.verts = mesh->verts;
    view.uvs = mesh->uvs;
    view.types = mesh->types;
    view.colors = mesh->colors;
    view.textureIds = mesh->texture_ids;
    view.count = mesh->count;
    std::vector<TextureView> views;
    for (uint32_t i = 0; i < texture_count; ++i) {
        views.push_back(TextureView{textures[i].pixels, textures[i].width, textures[i].height, textures[i].channels,
                                    textures[i].wrap});
    }
    RunStats stats;
    const int rc = engine->engine->voxelize(view, views.data(), texture_count, paramsFromC(*params),
                                            static_cast<cudaStream_t>(cuda_stream), &stats);
    if (rc != 0) {
        gLastError = engine->engine->lastError();
    }
    if (out_stats != nullptr) {
        statsToC(stats, out_stats);
    }
    return rc;
}

const void *o2v_b200_result_device(const o2v_b200_engine *engine)
{
    return engine->engine->deviceVoxels();
}

uint64_t o2v_b200_result_count(const o2v_b200_engine *engine)
{
    return engine->engine->voxelCount();
}

uint32_t o2v_b200_plan_parts(uint32_t sample_resolution, uint32_t slab_z0, uint32_t slab_z1, uint64_t triangles,
                             int32_t requested_parts, uint32_t *out_bounds, uint32_t bounds_capacity)
{
    uint32_t bounds[kMaxJobParts + 1];
    const uint32_t parts = planJobParts(sample_resolution, slab_z0, slab_z1, triangles, requested_parts, bounds);
    for (uint32_t k = 0; k <= parts && k < bounds_capacity; ++k) {
        out_bounds[k] = bounds[k];
    }
    return parts;
}

int o2v_b200_result_download(o2v_b200_engine *engine, void *host_dst, void *cuda_stream)
{
    const int rc = engine->engine->download(host_dst, static_cast<cudaStream_t>(cuda_stream));
    if (rc != 0) {
        gLastError = engine->engine->lastError();
    }
    return rc;
}

int o2v_b200_filter_slab(o2v_b200_engine *engine, const o2v_b200_params *params, const o2v_b200_mesh *mesh,
                         void *cuda_stream, const float **out_kept, uint64_t *out_count)
{
    MeshView view{};
    view.verts = mesh->verts;
    view.count = mesh->count;
    unsigned long long count = 0;
    const int rc = engine->engine->filterSlab(view, paramsFromC(*params), static_cast<cudaStream_t>(cuda_stream),
                                              out_kept, &count);
    if (rc != 0) {
        gLastError = engine->engine->lastError();
    }
    *out_count = count;
    return rc;
}

int o2v_b200_result_hash(o2v_b200_engine *engine, void *cuda_stream, uint64_t *out_hash)
{
    unsigned long long hash = 0;
    const int rc = engine->engine->resultHash(static_cast<cudaStream_t>(cuda_stream), &hash);
    if (rc != 0) {
        gLastError = engine->engine->lastError();
    }
    *out_hash = hash;
    return rc;
}

int o2v_b200_voxelize_host(o2v_b200_engine *engine, const o2v_b200_params *params, const o2v_b200_mesh *mesh,
                           const o2v_b200_texture *textures, uint32_t texture_count, uint32_t *out_voxels,
                           uint64_t out_capacity, uint64_t *out_count, o2v_b200_stats *out_stats)
{
    cudaSetDevice(engine->engine->device());
    UploadedMesh uploaded;
    std::string error;
    if (!uploaded.upload(*mesh, textures, texture_count, nullptr, &error)) {
        gLastError = error;
        return kErrCuda;
    }
    RunStats stats;
    const int rc = engine->engine->voxelize(uploaded.view, uploaded.textureViews.data(), texture_count,
                                            paramsFromC(*params), nullptr, &stats);
    if (out_stats != nullptr) {
        statsToC(stats, out_stats);
    }
    if (rc != 0) {
        gLastError = engine->engine->lastError();
        return rc;
    }
    const uint64_t count = engine->engine->voxelCount();
    if (out_count != nullptr) {
        *out_count = count;
    }
    if (count > out_capacity) {
        gLastError = "output buffer too small";
        return -5;
    }
    const int drc = engine->engine->download(out_voxels, nullptr);
    if (drc != 0) {
        gLastError = engine->engine->lastError();
    }
    return drc;
}

void obj2voxel_b200_set_input_triangles(obj2voxel_instance *instance, const float *vertices, const float *uvs,
                                        size_t count, obj2voxel_texture *texture)
{
    O2V_REQUIRE(instance != nullptr && (vertices != nullptr || count == 0), "instance or vertices is NULL");
    instance->inputKind = IoKind::BULK;
    instance->bulkVerts = vertices;
    instance->bulkUvs = uvs;
    instance->bulkCount = count;
    instance->bulkTexture = texture;
}

void obj2voxel_b200_set_slab(obj2voxel_instance *instance, uint32_t z0, uint32_t z1)
{
    O2V_REQUIRE(instance != nullptr, "instance is NULL");
    instance->slabZ0 = z0;
    instance->slabZ1 = z1;
}

void obj2voxel_b200_get_stats(obj2voxel_instance *instance, o2v_b200_stats *out_stats)
{
    O2V_REQUIRE(instance != nullptr && out_stats != nullptr, "instance or out_stats is NULL");
    *out_stats = instance->stats;
}

}  // extern "C"
