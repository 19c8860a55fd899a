// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Harness around the UNMODIFIED reference (Eisenwave/obj2voxel @ 9fb8ae2) compiled from the
// sources where they lie under /root/reference (see oracle/Makefile).  It exists to
//   (1) pin oracle/o2v_oracle.c (the C restatement) against the real implementation,
//   (2) generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   (3) act as the "reference" CPU arm of bench.py (--impl reference, cpu_baseline.kind="reference").
//
// This translation unit textually includes the reference's src/obj2voxel.cpp so that the functions in
// its anonymous namespace (computeMeshTransform :370, applyMeshTransform :202, sortTriangleIntoChunks :226,
// computeChunkBounds :245, findMeshBounds :180) can be driven directly and float WeightedColor values read
// out of obj2voxel::Voxelizer::voxels() before the ARGB8 quantisation of voxelizeChunk (:283-296).
// No reference source is copied into this repository.

#include "obj2voxel.cpp"  // resolved through -I$(REF)/src by oracle/Makefile

#include <chrono>
#include <cstdlib>
#include <thread>

#include "voxelio/format/qef.hpp"
#include "voxelio/format/vl32.hpp"
#include "voxelio/format/vox.hpp"

namespace {

struct ArrayInput {
    const float *verts;   // 9 per triangle
    const float *uvs;     // 6 per triangle or null
    size_t count;
    size_t next = 0;
    obj2voxel_texture *texture;
};

bool arrayInputCallback(void *data, obj2voxel_triangle *out)
{
    auto *in = static_cast<ArrayInput *>(data);
    if (in->next >= in->count) {
        return false;
    }
    const float *v = in->verts + in->next * 9;
    if (in->uvs != nullptr && in->texture != nullptr) {
        obj2voxel_set_triangle_textured(out, v, in->uvs + in->next * 6, in->texture);
    }
    else {
        obj2voxel_set_triangle_basic(out, v);
    }
    ++in->next;
    return true;
}

struct CollectOutput {
    std::vector<uint32_t> voxels;
    size_t calls = 0;
};

bool collectOutputCallback(void *data, uint32_t *voxels, size_t count)
{
    auto *out = static_cast<CollectOutput *>(data);
    out->voxels.insert(out->voxels.end(), voxels, voxels + count * 4);
    ++out->calls;
    return true;
}

struct CountOutput {
    size_t voxels = 0;
};

bool countOutputCallback(void *data, uint32_t *, size_t count)
{
    static_cast<CountOutput *>(data)->voxels += count;
    return true;
}

obj2voxel_texture *makeTexture(const uint8_t *pixels, size_t w, size_t h, size_t channels, int wrapMode)
{
    if (pixels == nullptr) {
        return nullptr;
    }
    obj2voxel_texture *tex = obj2voxel_texture_alloc();
    obj2voxel_texture_load_pixels(tex, pixels, w, h, channels);
    obj2voxel_teture_set_uv_mode(tex, static_cast<obj2voxel_enum_t>(wrapMode));
    return tex;
}

}  // namespace

extern "C" {

/// Runs the reference through its public C API (include/obj2voxel.h) with `workers` worker threads
/// (0 = single-threaded, set_parallel(false)).  All setters precede worker start.
/// collect != 0: *outVoxels receives a malloc'ed array of 4*n u32 (x,y,z,argb) in the reference's own output order.
/// Returns the voxel count, or -(error code) on failure.
long long o2vref_run_api(const float *verts, const float *uvs, size_t triangleCount,
                         const uint8_t *texPixels, size_t texW, size_t texH, size_t texChannels, int texWrap,
                         uint32_t resolution, uint32_t supersampling, int strategy,
                         const float *bounds6, const int *unit9, int workers, int collect,
                         uint32_t **outVoxels, double *outSeconds, size_t *outSinkCalls)
{
    obj2voxel_set_log_level(OBJ2VOXEL_LOG_LEVEL_ERROR);

    obj2voxel_texture *tex = makeTexture(texPixels, texW, texH, texChannels, texWrap);
    ArrayInput input{verts, uvs, triangleCount, 0, tex};
    CollectOutput collected;
    CountOutput counted;

    obj2voxel_instance *instance = obj2voxel_alloc();
    obj2voxel_set_input_callback(instance, &arrayInputCallback, &input);
    if (collect) {
        obj2voxel_set_output_callback(instance, &collectOutputCallback, &collected);
    }
    else {
        obj2voxel_set_output_callback(instance, &countOutputCallback, &counted);
    }
    obj2voxel_set_resolution(instance, resolution);
    obj2voxel_set_supersampling(instance, supersampling);
    obj2voxel_set_color_strategy(instance, static_cast<obj2voxel_enum_t>(strategy));
    if (bounds6 != nullptr) {
        obj2voxel_set_mesh_boundaries(instance, bounds6);
    }
    if (unit9 != nullptr) {
        obj2voxel_set_unit_transform(instance, unit9);
    }
    obj2voxel_set_parallel(instance, workers > 0);

    std::vector<std::thread> threads;
    for (int i = 0; i < workers; ++i) {
        threads.emplace_back(&obj2voxel_run_worker, instance);
    }
    if (workers > 0) {
        while (obj2voxel_get_worker_count(instance) != static_cast<uint32_t>(workers)) {
            std::this_thread::yield();
        }
    }

    const auto t0 = std::chrono::steady_clock::now();
    const obj2voxel_error_t error = obj2voxel_voxelize(instance);
    const auto t1 = std::chrono::steady_clock::now();

    obj2voxel_stop_workers(instance);
    for (std::thread &t : threads) {
        t.join();
    }
    obj2voxel_free(instance);
    if (tex != nullptr) {
        obj2voxel_texture_free(tex);
    }

    if (outSeconds != nullptr) {
        *outSeconds = std::chrono::duration<double>(t1 - t0).count();
    }
    if (outSinkCalls != nullptr) {
        *outSinkCalls = collected.calls;
    }
    if (error != OBJ2VOXEL_ERR_OK) {
        return -static_cast<long long>(error);
    }
    if (!collect) {
        return static_cast<long long>(counted.voxels);
    }
    const size_t n = collected.voxels.size() / 4;
    if (outVoxels != nullptr) {
        auto *buffer = static_cast<uint32_t *>(std::malloc(std::max<size_t>(1, n * 4) * sizeof(uint32_t)));
        std::memcpy(buffer, collected.voxels.data(), n * 4 * sizeof(uint32_t));
        *outVoxels = buffer;
    }
    return static_cast<long long>(n);
}

/// Drives the reference's *internal* pipeline single-threaded (the same call sequence as
/// voxelize_specialized<false>, src/obj2voxel.cpp:467-520) but reads the float WeightedColor map of each chunk before
/// it is quantised.  types: 1 MATERIALLESS, 2 UNTEXTURED (colors, 3 per triangle), 3 TEXTURED; null = all MATERIALLESS
/// (or TEXTURED when uvs and a texture are given).
/// applyDownscale != 0 calls Voxelizer::downscale() per chunk exactly like voxelizeChunk does for supersampling > 1.
/// Output: xyz (3 u32 per voxel, sample-space or downscaled position) and wrgb (weight,r,g,b floats), malloc'ed.
/// outTransform receives the 12 floats of the mesh transform (row-major 3x3, then translation).
long long o2vref_run_internal(const float *verts, const float *uvs, const uint8_t *types, const float *colors,
                              size_t triangleCount,
                              const uint8_t *texPixels, size_t texW, size_t texH, size_t texChannels, int texWrap,
                              uint32_t resolution, uint32_t supersampling, int strategy,
                              const float *bounds6, const int *unit9, int applyDownscale,
                              uint32_t **outXyz, float **outWrgb, float *outTransform)
{
    obj2voxel_set_log_level(OBJ2VOXEL_LOG_LEVEL_ERROR);
    obj2voxel_texture *tex = makeTexture(texPixels, texW, texH, texChannels, texWrap);

    obj2voxel_instance *instance = obj2voxel_alloc();
    obj2voxel_set_resolution(instance, resolution);
    obj2voxel_set_supersampling(instance, supersampling);
    obj2voxel_set_color_strategy(instance, static_cast<obj2voxel_enum_t>(strategy));
    if (bounds6 != nullptr) {
        obj2voxel_set_mesh_boundaries(instance, bounds6);
    }
    if (unit9 != nullptr) {
        obj2voxel_set_unit_transform(instance, unit9);
    }

    const u32 chunkCountCbrt = divCeil(instance->sampleResolution, CHUNK_SIZE);
    instance->chunkCount = chunkCountCbrt * chunkCountCbrt * chunkCountCbrt;

    for (size_t i = 0; i < triangleCount; ++i) {
        CachedTriangle triangle{};
        const uint8_t type = types != nullptr ? types[i] : (uvs != nullptr && tex != nullptr ? 3 : 1);
        if (type == 3) {
            obj2voxel_set_triangle_textured(&triangle, verts + i * 9, uvs + i * 6, tex);
        }
        else {
            obj2voxel_set_triangle_basic(&triangle, verts + i * 9);
            if (type == 2) {
                triangle.type = TriangleType::UNTEXTURED;
                triangle.color = Vec3f{colors + i * 3};
            }
        }
        instance->triangles.push_back(triangle);
    }

    const u32 n = static_cast<u32>(triangleCount);
    if (not instance->boundsKnown) {
        for (u32 i = 0; i < n; i += BATCH_SIZE) {
            obj2voxel::findMeshBounds(*instance, i);
        }
    }
    instance->meshTransform = computeMeshTransform(*instance);
    if (outTransform != nullptr) {
        for (usize i = 0; i < 3; ++i) {
            for (usize j = 0; j < 3; ++j) {
                outTransform[i * 3 + j] = instance->meshTransform.matrix[i][j];
            }
            outTransform[9 + i] = instance->meshTransform.translation[i];
        }
    }
    for (u32 i = 0; i < n; i += BATCH_SIZE) {
        obj2voxel::applyMeshTransform(*instance, i);
    }
    for (u32 i = 0; i < n; ++i) {
        sortTriangleIntoChunks(*instance, i);
    }

    std::vector<uint32_t> xyz;
    std::vector<float> wrgb;
    Voxelizer voxelizer{instance->colorStrategy};
    for (u32 chunkIndex = 0; chunkIndex < instance->chunkCount; ++chunkIndex) {
        auto found = instance->chunks.find(chunkIndex);
        if (found == instance->chunks.end()) {
            continue;
        }
        Vec3u32 chunkMin, chunkMax;
        computeChunkBounds(chunkIndex, chunkMin, chunkMax);
        for (u32 triangle : found->second) {
            voxelizer.voxelize(instance->triangles[triangle], chunkMin, chunkMax);
        }
        if (applyDownscale) {
            voxelizer.downscale();
        }
        for (auto &[index, color] : voxelizer.voxels()) {
            const Vec3u32 pos = VoxelMap<WeightedColor>::posOf(index);
            xyz.insert(xyz.end(), {pos[0], pos[1], pos[2]});
            wrgb.insert(wrgb.end(), {color.weight, color.value[0], color.value[1], color.value[2]});
        }
        voxelizer.voxels().clear();
    }

    obj2voxel_free(instance);
    if (tex != nullptr) {
        obj2voxel_texture_free(tex);
    }

    const size_t count = wrgb.size() / 4;
    auto *xyzBuffer = static_cast<uint32_t *>(std::malloc(std::max<size_t>(1, count * 3) * sizeof(uint32_t)));
    auto *wrgbBuffer = static_cast<float *>(std::malloc(std::max<size_t>(1, count * 4) * sizeof(float)));
    std::memcpy(xyzBuffer, xyz.data(), count * 3 * sizeof(uint32_t));
    std::memcpy(wrgbBuffer, wrgb.data(), count * 4 * sizeof(float));
    *outXyz = xyzBuffer;
    *outWrgb = wrgbBuffer;
    return static_cast<long long>(count);
}

/// Reads a voxel file written by anyone with the reference's own voxelio reader for `type` ("qef", "vox", "vl32"):
/// *outVoxels receives a malloc'ed array of 4*n u32 (x, y, z, argb) in file order.  Returns n, or -1 if the file cannot
/// be opened, -2 if the type has no reader, -3 if the reader reports an error.
long long o2vref_read_voxel_file(const char *path, const char *type, uint32_t **outVoxels)
{
    std::optional<voxelio::FileInputStream> stream = voxelio::FileInputStream::open(path);
    if (not stream.has_value()) {
        return -1;
    }
    std::unique_ptr<voxelio::AbstractReader> reader;
    const std::string t = type;
    if (t == "qef") {
        reader.reset(new voxelio::qef::Reader{*stream});
    }
    else if (t == "vox") {
        reader.reset(new voxelio::vox::Reader{*stream});
    }
    else if (t == "vl32") {
        reader.reset(new voxelio::vl32::Reader{*stream});
    }
    else {
        return -2;
    }
    std::vector<uint32_t> voxels;
    std::vector<voxelio::Voxel64> buffer(8192);
    for (;;) {
        const voxelio::ReadResult result = reader->read(buffer.data(), buffer.size());
        if (result.isBad()) {
            return -3;
        }
        for (uint64_t i = 0; i < result.voxelsRead; ++i) {
            const voxelio::Voxel64 &v = buffer[i];
            voxels.insert(voxels.end(), {static_cast<uint32_t>(v.pos[0]), static_cast<uint32_t>(v.pos[1]),
                                         static_cast<uint32_t>(v.pos[2]), v.argb});
        }
        if (result.isEnd()) {
            break;
        }
    }
    const size_t n = voxels.size() / 4;
    auto *out = static_cast<uint32_t *>(std::malloc(std::max<size_t>(1, n * 4) * sizeof(uint32_t)));
    std::memcpy(out, voxels.data(), n * 4 * sizeof(uint32_t));
    *outVoxels = out;
    return static_cast<long long>(n);
}

/// Runs the reference from an input FILE (its own OBJ / STL readers: tinyobjloader, src/io.cpp) to either a voxel
/// callback (outputPath == nullptr: *outVoxels receives 4*n u32) or an output FILE written by voxelio's writers.
/// Returns the voxel count (callback) or 0 (file), or -(error code).
long long o2vref_run_file(const char *inputPath, const char *outputPath, uint32_t resolution, uint32_t supersampling,
                          int strategy, int workers, uint32_t **outVoxels)
{
    obj2voxel_set_log_level(OBJ2VOXEL_LOG_LEVEL_ERROR);
    CollectOutput collected;
    obj2voxel_instance *instance = obj2voxel_alloc();
    obj2voxel_set_input_file(instance, inputPath, nullptr);
    if (outputPath != nullptr) {
        obj2voxel_set_output_file(instance, outputPath, nullptr);
    }
    else {
        obj2voxel_set_output_callback(instance, &collectOutputCallback, &collected);
    }
    obj2voxel_set_resolution(instance, resolution);
    obj2voxel_set_supersampling(instance, supersampling);
    obj2voxel_set_color_strategy(instance, static_cast<obj2voxel_enum_t>(strategy));
    obj2voxel_set_parallel(instance, workers > 0);
    std::vector<std::thread> threads;
    for (int i = 0; i < workers; ++i) {
        threads.emplace_back(&obj2voxel_run_worker, instance);
    }
    if (workers > 0) {
        while (obj2voxel_get_worker_count(instance) != static_cast<uint32_t>(workers)) {
            std::this_thread::yield();
        }
    }
    const obj2voxel_error_t error = obj2voxel_voxelize(instance);
    obj2voxel_stop_workers(instance);
    for (std::thread &t : threads) {
        t.join();
    }
    obj2voxel_free(instance);
    if (error != OBJ2VOXEL_ERR_OK) {
        return -static_cast<long long>(error);
    }
    if (outputPath != nullptr) {
        return 0;
    }
    const size_t n = collected.voxels.size() / 4;
    auto *buffer = static_cast<uint32_t *>(std::malloc(std::max<size_t>(1, n * 4) * sizeof(uint32_t)));
    std::memcpy(buffer, collected.voxels.data(), n * 4 * sizeof(uint32_t));
    *outVoxels = buffer;
    return static_cast<long long>(n);
}

void o2vref_free(void *pointer)
{
    std::free(pointer);
}

unsigned o2vref_hardware_threads(void)
{
    return std::thread::hardware_concurrency();
}

}  // extern "C"
