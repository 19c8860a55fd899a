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

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "o2v_engine.h"
#include "o2v_io.h"
#include "o2v_job.h"

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
    obj2voxel_log_callback *callback = nullptr;
    void *callbackData = nullptr;
    {
        std::lock_guard<std::mutex> lock{gLogMutex};
        if (level > gLogLevel) {
            return;
        }
        callback = gLogCallback;
        callbackData = gLogCallbackData;
    }
    // the callback runs outside the lock: it may call obj2voxel_set_log_level / _get_log_level itself
    if (callback != nullptr && callback(callbackData, message.c_str(), level)) {
        return;
    }
    static std::mutex printMutex;
    std::lock_guard<std::mutex> lock{printMutex};
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
    out->download_bytes = in.downloadBytes;
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
    e.floatRecords = p.float_records != 0;
    e.accumulate = p.accumulate;
    return e;
}

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
    std::vector<int> devices;  // obj2voxel_b200_set_devices; empty = the process default

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
    // positions and types for every triangle; uvs / texture ids / colours only from the first triangle that has them
    // (zero-filled before): a MATERIALLESS stream — any STL, obj2voxel_set_triangle_basic — costs 37 bytes per triangle
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
        const size_t index = types.size();
        verts.insert(verts.end(), t.v, t.v + 9);
        uint8_t type = t.type;
        if (type == kTextured) {
            if (t.texture != nullptr && t.texture->loaded) {
                if (!anyTextured) {
                    anyTextured = true;
                    uvs.assign(index * 6, 0.0f);
                    textureIds.assign(index, 0u);
                }
            }
            else {
                type = kMaterialless;
            }
        }
        if (type == kUntextured && !anyColored) {
            anyColored = true;
            colors.assign(index * 3, 0.0f);
        }
        if (anyTextured) {
            const bool textured = type == kTextured;
            static const float zeros[6] = {0, 0, 0, 0, 0, 0};
            uvs.insert(uvs.end(), textured ? t.t : zeros, (textured ? t.t : zeros) + 6);
            textureIds.push_back(textured ? textureIndex(t.texture) : 0u);
        }
        if (anyColored) {
            colors.insert(colors.end(), t.color, t.color + 3);
        }
        types.push_back(type);
    }
};

/// openSink: creates inst.sink — called once the input has been read, so that a missing or unreadable input leaves an
/// existing output file alone (the reference opens its input first as well: src/obj2voxel.cpp:617-625).
obj2voxel_error_t runJob(obj2voxel_instance &inst, const std::function<obj2voxel_error_t()> &openSink)
{
    // ---- gather the triangle stream into flat arrays (reference: "Caching triangles", obj2voxel.cpp:583-588) ----
    HostMesh host;
    o2v_b200_mesh mesh{};
    std::vector<o2v_b200_texture> textures;
    struct LoadedTextures {  // what an OBJ's material library made the reader load: freed when the job is over
        std::vector<obj2voxel_texture *> all;
        ~LoadedTextures()
        {
            for (obj2voxel_texture *t : all) {
                obj2voxel_texture_free(t);
            }
        }
    } loaded;

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
            if (!readTriangleFile(inst.inputFile, inst.inputFormat, inst.defaultTexture, appender, &loaded.all, &error)) {
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

    if (const obj2voxel_error_t opened = openSink()) {
        return opened;
    }
    if (mesh.count == 0) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_WARNING, "Model has no triangles, aborting and writing empty voxel model");
        inst.sink->finalize();
        return inst.sink->good() ? OBJ2VOXEL_ERR_OK : OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE;
    }
    logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "Cached model with " + withThousands(mesh.count) + " triangles");

    // ---- device run (o2v_job.cpp): upload, kernels per z part, download under the next part's kernels, sink ----
    JobOptions options;
    EngineParams &params = options.params;
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
    if (const char *env = getenv("O2V_B200_PIPELINE_PARTS")) {
        options.parts = atoi(env);
    }
    options.devices = inst.devices;  // empty: O2V_B200_DEVICES / O2V_B200_DEVICE / device 0

    if (inst.supersampling > 1) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "Chunks will be downscaled from " +
                                                 withThousands(inst.outputResolution * inst.supersampling) +
                                                 " to output resolution " + withThousands(inst.outputResolution) +
                                                 " ...");
    }

    RunStats stats;
    JobTimings timings;
    const obj2voxel_error_t rc = runDeviceJob(mesh, textures, options, *inst.sink, &stats, &timings);
    statsToC(stats, &inst.stats);
    if (rc != OBJ2VOXEL_ERR_OK) {
        return rc;
    }
    if (!inst.sink->good()) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Voxelization failed because of IO error");
        return OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE;
    }
    char timing[512];
    snprintf(timing, sizeof timing,
             "timing: %u device(s), upload %.2f ms%s, slab exchange %.2f ms%s, %u part(s) each: kernels %.2f ms, run + "
             "download%s + sink %.2f ms (device 0: voxelize calls %.2f, waiting for copies %.2f, host expansion %.2f, sink "
             "%.2f ms)",
             timings.devices, timings.msUpload,
             timings.streamedUpload ? " (the upload runs under the parts: a part = a piece of the triangle array, voxelized as it arrives)"
             : timings.stagedUpload ? " (pageable input staged by host threads)"
                                    : "",
             timings.msExchange, timings.peerExchange ? " (peer stores)" : "", timings.parts, timings.msKernels,
             timings.bitmapDownload ? " of bitmaps + host expansion" : "", timings.msRun, timings.msVoxelizeCalls,
             timings.msWaitCopy, timings.msExpandHost, timings.msSink);
    logMessage(OBJ2VOXEL_LOG_LEVEL_DEBUG, timing);
    if (timings.devices > 1) {
        std::string bounds = "slab bounds (sample-space z, balanced by triangles per chunk row):";
        for (uint32_t z : timings.slabBounds) {
            bounds += " " + std::to_string(z);
        }
        logMessage(OBJ2VOXEL_LOG_LEVEL_DEBUG, bounds);
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

    auto openSink = [&inst]() -> obj2voxel_error_t {
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
        return OBJ2VOXEL_ERR_OK;
    };

    const obj2voxel_error_t result = runJob(inst, openSink);
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

bool obj2voxel_b200_array_source_next(void *source, obj2voxel_triangle *out_triangle)
{
    o2v_b200_array_source *s = static_cast<o2v_b200_array_source *>(source);
    if (s->next >= s->count) {
        return false;
    }
    obj2voxel_set_triangle_basic(out_triangle, s->vertices + 9 * s->next++);
    return true;
}

bool obj2voxel_b200_counting_sink_write(void *sink, uint32_t *, size_t voxel_count)
{
    o2v_b200_counting_sink *s = static_cast<o2v_b200_counting_sink *>(sink);
    s->voxels += voxel_count;
    ++s->calls;
    return true;
}

void obj2voxel_b200_set_devices(obj2voxel_instance *instance, const int32_t *devices, uint32_t count)
{
    O2V_REQUIRE(instance != nullptr && (devices != nullptr || count == 0), "instance or devices is NULL");
    instance->devices.assign(devices, devices + count);
}

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

const float *o2v_b200_result_floats_device(const o2v_b200_engine *engine)
{
    return engine->engine->floatRecords();
}

uint64_t o2v_b200_result_count(const o2v_b200_engine *engine)
{
    return engine->engine->voxelCount();
}

uint64_t o2v_b200_expand_bitmaps(const uint64_t *bits, const uint32_t *chunk_ids, const uint32_t *chunk_counts,
                                 uint32_t chunks, uint32_t chunks_per_axis, uint32_t chunk_z0, uint32_t *out_quads)
{
    static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "bitmap words");
    return expandBitmapsOnHost(reinterpret_cast<const unsigned long long *>(bits), chunk_ids, chunk_counts, chunks,
                               chunks_per_axis, chunk_z0, out_quads);
}

uint64_t o2v_b200_scan_chunk_bitmap(const uint64_t *words, uint32_t cx, uint32_t cy, uint32_t cz, uint32_t buffer_quads,
                                    uint32_t *out_quads, uint64_t out_capacity)
{
    std::vector<uint32_t> buffer((size_t) std::max(buffer_quads, 64u) * 4 + 16);
    uint64_t written = 0;
    const unsigned long long n = scanChunkBitmap(
        reinterpret_cast<const unsigned long long *>(words), cx, cy, cz, buffer.data(), std::max(buffer_quads, 64u),
        [&](uint32_t *quads, size_t count) {
            if (written + count > out_capacity) {
                return false;
            }
            memcpy(out_quads + written * 4, quads, count * 16);
            written += count;
            return true;
        });
    return n == ~0ull ? UINT64_MAX : written;
}

void o2v_b200_expand_packed(const void *packed, int32_t bits, uint64_t count, uint32_t *out_quads)
{
    expandPackedOnHost(packed, bits, count, out_quads);
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

void o2v_b200_plan_slabs(uint32_t sample_resolution, uint32_t supersampling, uint32_t slab_z0, uint32_t slab_z1,
                         uint32_t devices, const uint64_t *row_histogram, uint32_t rows, uint32_t *out_bounds)
{
    if (devices == 0 || devices > kMaxSlabs || supersampling == 0) {
        return;
    }
    static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "histogram element");
    planJobSlabs(sample_resolution, supersampling, slab_z0, slab_z1, devices,
                 reinterpret_cast<const unsigned long long *>(row_histogram), rows, out_bounds);
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
    // plain: synchronous copies, one run, one download (obj2voxel_voxelize() is the pipelined host path)
    cudaSetDevice(engine->engine->device());
    struct Buffers {
        std::vector<void *> all;
        ~Buffers()
        {
            for (void *p : all) {
                cudaFree(p);
            }
        }
        void *upload(const void *src, size_t bytes)
        {
            if (src == nullptr || bytes == 0) {
                return nullptr;
            }
            void *dst = nullptr;
            if (cudaMalloc(&dst, bytes) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(dst);
                failed = true;
                return nullptr;
            }
            all.push_back(dst);
            return dst;
        }
        bool failed = false;
    } buffers;
    const size_t n = (size_t) mesh->count;
    MeshView view{};
    view.verts = static_cast<const float *>(buffers.upload(mesh->verts, n * 9 * sizeof(float)));
    view.uvs = static_cast<const float *>(buffers.upload(mesh->uvs, n * 6 * sizeof(float)));
    view.types = static_cast<const uint8_t *>(buffers.upload(mesh->types, n));
    view.colors = static_cast<const float *>(buffers.upload(mesh->colors, n * 3 * sizeof(float)));
    view.textureIds = static_cast<const uint32_t *>(buffers.upload(mesh->texture_ids, n * sizeof(uint32_t)));
    view.count = mesh->count;
    std::vector<TextureView> views;
    for (uint32_t i = 0; i < texture_count; ++i) {
        const size_t bytes = (size_t) textures[i].width * textures[i].height * textures[i].channels;
        views.push_back(TextureView{static_cast<const uint8_t *>(buffers.upload(textures[i].pixels, bytes)),
                                    textures[i].width, textures[i].height, textures[i].channels, textures[i].wrap});
    }
    if (buffers.failed) {
        gLastError = "mesh upload failed (device allocation or copy)";
        return kErrCuda;
    }
    RunStats stats;
    const int rc = engine->engine->voxelize(view, views.data(), texture_count, paramsFromC(*params), nullptr, &stats);
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
