// Host-side I/O adapters around the GPU hot path: voxel sinks (callback, VL32, PLY, XYZRGB; file or memory) and triangle
// file readers (binary STL, Wavefront OBJ).  Counterpart of the reference's src/io.{hpp,cpp}; formats follow
// voxelio/src/format/{vl32,ply,xyzrgb}.cpp.
#ifndef O2V_IO_H
#define O2V_IO_H

#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "obj2voxel.h"

namespace o2v {

void logMessage(unsigned char level, const std::string &message);

enum class FileFormat { UNKNOWN, OBJ, STL, VL32, PLY, XYZRGB, QEF, VOX, PNG };

/// Extension lookup like voxelio::fileTypeOfExtension / detectFileType (src/obj2voxel.cpp:316-335): `type` wins over the
/// path's suffix.
FileFormat detectFormat(const char *path, const char *type);
bool canReadTriangles(FileFormat format);

/// Receives batches of (x, y, z, argb) host-endian u32 quads; counterpart of IVoxelSink (src/io.hpp:69-92).
class VoxelSink {
public:
    virtual ~VoxelSink() = default;
    virtual bool write(uint32_t *quads, size_t count) = 0;
    virtual void finalize() = 0;
    virtual bool good() const = 0;
    virtual size_t voxelsWritten() const = 0;
    virtual const std::vector<uint8_t> *memory() const { return nullptr; }
};

std::unique_ptr<VoxelSink> makeCallbackSink(obj2voxel_voxel_callback *callback, void *data);
/// path == nullptr: in-memory byte array (obj2voxel_set_output_memory).  Returns nullptr + *error on failure.
std::unique_ptr<VoxelSink> makeFormatSink(FileFormat format, const char *path, uint32_t resolution, std::string *error);

using TriangleAppender = std::function<void(const float v[9], const float uv[6], uint8_t type, const float color[3],
                                            const obj2voxel_texture *texture)>;

/// Streams every triangle of an STL / OBJ file into `append`.  Textures an OBJ's material library names are loaded
/// through the C API and appended to *loadedTextures: the triangles point at them, the caller frees them
/// (obj2voxel_texture_free) once the job has uploaded them.
bool readTriangleFile(const char *path, FileFormat format, const obj2voxel_texture *defaultTexture,
                      const TriangleAppender &append, std::vector<obj2voxel_texture *> *loadedTextures, std::string *error);

bool readWholeFile(const char *path, std::vector<uint8_t> *out);

/// Minimal PNG decoder (8-bit, non-interlaced, colour types 0/2/3/4/6) to RGBA8; zlib does the inflate.
bool decodePng(const uint8_t *data, size_t size, std::vector<uint8_t> *rgba, size_t *width, size_t *height,
               std::string *error);

}  // namespace o2v

#endif  // O2V_IO_H
