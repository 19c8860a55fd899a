// Host-side I/O adapters (see o2v_io.h).  None of this is on the GPU hot path; it brackets it.
#include "o2v_io.h"

#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>

namespace o2v {

// ---------------------------------------------------------------------------------------------------------------------
// file types

static std::string lowerSuffix(const char *path)
{
    std::string s{path};
    const size_t dot = s.find_last_of('.');
    const size_t slash = s.find_last_of("/\\");
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) {
        return "";
    }
    std::string ext = s.substr(dot + 1);
    std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char) tolower(c); });
    return ext;
}

FileFormat detectFormat(const char *path, const char *type)
{
    std::string ext;
    if (type != nullptr) {
        ext = type;
        std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char) tolower(c); });
    }
    else if (path != nullptr) {
        ext = lowerSuffix(path);
    }
    if (ext == "obj") return FileFormat::OBJ;
    if (ext == "stl") return FileFormat::STL;
    if (ext == "vl32") return FileFormat::VL32;
    if (ext == "ply") return FileFormat::PLY;
    if (ext == "xyzrgb") return FileFormat::XYZRGB;
    if (ext == "qef") return FileFormat::QEF;
    if (ext == "vox") return FileFormat::VOX;
    if (ext == "png") return FileFormat::PNG;
    return FileFormat::UNKNOWN;
}

bool canReadTriangles(FileFormat format)
{
    return format == FileFormat::OBJ || format == FileFormat::STL;  // src/obj2voxel.cpp:533-537
}

bool readWholeFile(const char *path, std::vector<uint8_t> *out)
{
    FILE *f = fopen(path, "rb");
    if (f == nullptr) {
        return false;
    }
    uint8_t chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) != 0) {
        out->insert(out->end(), chunk, chunk + got);
    }
    const bool ok = !ferror(f);
    fclose(f);
    return ok;
}

// ---------------------------------------------------------------------------------------------------------------------
// sinks

namespace {

class CallbackSink final : public VoxelSink {
public:
    CallbackSink(obj2voxel_voxel_callback *callback, void *data) : callback_{callback}, data_{data} {}

    bool write(uint32_t *quads, size_t count) override
    {
        // layout = voxelio Voxel32 {i32 pos[3]; u32 argb} reinterpreted as four u32 (src/io.cpp:638-653)
        ok_ = ok_ && callback_(data_, quads, count);
        written_ += ok_ ? count : 0;
        return ok_;
    }
    void finalize() override {}
    bool good() const override { return ok_; }
    size_t voxelsWritten() const override { return written_; }

private:
    obj2voxel_voxel_callback *callback_;
    void *data_;
    bool ok_ = true;
    size_t written_ = 0;
};

/// Byte stream to a file or to memory with absolute seek (PLY patches its vertex count on finalize).
class ByteStream {
public:
    explicit ByteStream(FILE *file) : file_{file} {}
    ~ByteStream()
    {
        if (file_ != nullptr) {
            fclose(file_);
        }
    }
    void write(const void *data, size_t size)
    {
        if (file_ != nullptr) {
            ok_ = ok_ && fwrite(data, 1, size, file_) == size;
            return;
        }
        if (position_ + size > bytes_.size()) {
            bytes_.resize(position_ + size);
        }
        memcpy(bytes_.data() + position_, data, size);
        position_ += size;
    }
    size_t position()
    {
        return file_ != nullptr ? (size_t) ftell(file_) : position_;
    }
    void seek(size_t absolute)
    {
        if (file_ != nullptr) {
            ok_ = ok_ && fseek(file_, (long) absolute, SEEK_SET) == 0;
        }
        else {
            position_ = absolute;
        }
    }
    void flush()
    {
        if (file_ != nullptr) {
            ok_ = ok_ && fflush(file_) == 0;
        }
    }
    bool good() const { return ok_; }
    const std::vector<uint8_t> *memory() const { return file_ == nullptr ? &bytes_ : nullptr; }

private:
    FILE *file_;
    std::vector<uint8_t> bytes_;
    size_t position_ = 0;
    bool ok_ = true;
};

inline uint32_t toBigEndian(uint32_t v)
{
    return __builtin_bswap32(v);
}

/// VL32 body: big-endian i32 x, y, z then big-endian argb, 16 bytes per voxel (voxelio/src/format/vl32.cpp:83-88).
/// PLY = a fixed-size binary_big_endian header in front of the same body (voxelio/src/format/ply.cpp:18-35,63-77).
/// XYZRGB = text "x y z r g b\n" (voxelio/src/format/xyzrgb.cpp:31-43).
class FormatSink final : public VoxelSink {
public:
    FormatSink(FileFormat format, FILE *file) : format_{format}, stream_{file}
    {
        if (format_ == FileFormat::PLY) {
            writeText("ply\r\n");
            writeText("format binary_big_endian 1.0\r\n");
            writeText("comment voxel list written by obj2voxel_b200; body is byte-identical to the VL32 format\r\n");
            writeText("element vertex ");
            countOffset_ = stream_.position();
            writeText("....;....;....;....;....;...\r\n");  // patched with the count + a trailing comment on finalize
            writeText("property int x\r\nproperty int y\r\nproperty int z\r\n");
            writeText("property uchar alpha\r\nproperty uchar red\r\nproperty uchar green\r\nproperty uchar blue\r\n");
            writeText("end_header\r\n");
        }
    }

    bool write(uint32_t *quads, size_t count) override
    {
        if (format_ == FileFormat::XYZRGB) {
            std::string text;
            for (size_t i = 0; i < count; ++i) {
                const uint32_t *q = quads + i * 4;
                char line[96];
                const int n = snprintf(line, sizeof line, "%d %d %d %u %u %u\n", (int32_t) q[0], (int32_t) q[1],
                                       (int32_t) q[2], (q[3] >> 16) & 255u, (q[3] >> 8) & 255u, q[3] & 255u);
                text.append(line, (size_t) n);
            }
            stream_.write(text.data(), text.size());
        }
        else {
            scratch_.resize(count * 4);
            for (size_t i = 0; i < count * 4; ++i) {
                scratch_[i] = toBigEndian(quads[i]);
            }
            stream_.write(scratch_.data(), count * 16);
        }
        written_ += count;
        return stream_.good();
    }

    void finalize() override
    {
        if (finalized_) {
            return;
        }
        finalized_ = true;
        if (format_ == FileFormat::PLY) {
            const size_t end = stream_.position();
            stream_.seek(countOffset_);
            const std::string patch = std::to_string(written_) + "\r\ncomment ";
            stream_.write(patch.data(), patch.size());
            stream_.seek(end);
        }
        stream_.flush();
    }

    bool good() const override { return stream_.good(); }
    size_t voxelsWritten() const override { return written_; }
    const std::vector<uint8_t> *memory() const override { return stream_.memory(); }

private:
    void writeText(const char *s) { stream_.write(s, strlen(s)); }

    FileFormat format_;
    ByteStream stream_;
    std::vector<uint32_t> scratch_;
    size_t written_ = 0;
    size_t countOffset_ = 0;
    bool finalized_ = false;
};

}  // namespace

std::unique_ptr<VoxelSink> makeCallbackSink(obj2voxel_voxel_callback *callback, void *data)
{
    return std::unique_ptr<VoxelSink>(new CallbackSink{callback, data});
}

std::unique_ptr<VoxelSink> makeFormatSink(FileFormat format, const char *path, uint32_t, std::string *error)
{
    if (format != FileFormat::VL32 && format != FileFormat::PLY && format != FileFormat::XYZRGB) {
        *error = "only the streamable voxel formats vl32, ply and xyzrgb are implemented (palette formats qef/vox are not)";
        return nullptr;
    }
    FILE *file = nullptr;
    if (path != nullptr) {
        file = fopen(path, "wb");
        if (file == nullptr) {
            *error = std::string("cannot open \"") + path + "\" for writing";
            return nullptr;
        }
    }
    return std::unique_ptr<VoxelSink>(new FormatSink{format, file});
}

// ---------------------------------------------------------------------------------------------------------------------
// triangle files

namespace {

/// Binary STL: 80-byte header, u32 count, 50-byte records (normal, 3 vertices, attribute) — src/io.cpp:395-435.
bool readStl(const char *path, const TriangleAppender &append, std::string *error)
{
    FILE *f = fopen(path, "rb");
    if (f == nullptr) {
        *error = std::string("Failed to open STL file: \"") + path + "\"";
        return false;
    }
    unsigned char header[80];
    if (fread(header, 1, 80, f) != 80) {
        fclose(f);
        *error = "Binary STL file must start with a header of 80 characters";
        return false;
    }
    if (memcmp(header, "solid", 5) == 0) {
        fclose(f);
        *error = "The given file is an ASCII STL file which is not supported";
        return false;
    }
    uint32_t count = 0;
    if (fread(&count, 4, 1, f) != 1) {
        fclose(f);
        *error = "Couldn't read STL triangle count";
        return false;
    }
    for (uint32_t i = 0; i < count; ++i) {
        unsigned char record[50];
        if (fread(record, 1, 50, f) != 50) {
            fclose(f);
            *error = "Unexpected EOF or error when reading triangle";
            return false;
        }
        float v[9];
        memcpy(v, record + 12, sizeof v);  // little-endian host assumed (x86-64 / aarch64)
        append(v, nullptr, /*MATERIALLESS*/ 1, nullptr, nullptr);
    }
    fclose(f);
    return true;
}

struct ObjMaterial {
    std::string name;
    float kd[3] = {1, 1, 1};
    bool hasKd = false;
    std::string mapKd;
    std::shared_ptr<void> texture;  // obj2voxel_texture owned here (allocated through the C API)
};

std::string directoryOf(const std::string &path)
{
    const size_t slash = path.find_last_of("/\\");
    return slash == std::string::npos ? std::string{} : path.substr(0, slash + 1);
}

/// Wavefront OBJ: v / vt / f (fan triangulation, negative indices), mtllib / usemtl with Kd and map_Kd (PNG).
/// Material mapping follows src/io.cpp:194-312: no material -> MATERIALLESS (or the default texture when the face has
/// uvs), material with a diffuse texture -> TEXTURED, otherwise UNTEXTURED with the diffuse colour.
bool readObj(const char *path, const obj2voxel_texture *defaultTexture, const TriangleAppender &append,
             std::string *error)
{
    FILE *f = fopen(path, "r");
    if (f == nullptr) {
        *error = std::string("Failed to open OBJ file: \"") + path + "\"";
        return false;
    }
    const std::string dir = directoryOf(path);
    std::vector<float> positions, texcoords;
    std::vector<ObjMaterial> materials;
    int current = -1;

    auto loadMtl = [&](const std::string &name) {
        FILE *m = fopen((dir + name).c_str(), "r");
        if (m == nullptr) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_WARNING, "Failed to open material library \"" + dir + name + "\"");
            return;
        }
        char line[1024];
        while (fgets(line, sizeof line, m) != nullptr) {
            char key[64];
            if (sscanf(line, "%63s", key) != 1) {
                continue;
            }
            const char *rest = strstr(line, key) + strlen(key);
            if (strcmp(key, "newmtl") == 0) {
                char name2[512];
                if (sscanf(rest, "%511s", name2) == 1) {
                    materials.emplace_back();
                    materials.back().name = name2;
                }
            }
            else if (!materials.empty() && strcmp(key, "Kd") == 0) {
                ObjMaterial &mat = materials.back();
                mat.hasKd = sscanf(rest, "%f %f %f", &mat.kd[0], &mat.kd[1], &mat.kd[2]) == 3;
            }
            else if (!materials.empty() && strcmp(key, "map_Kd") == 0) {
                char name2[512];
                if (sscanf(rest, "%511s", name2) == 1) {
                    materials.back().mapKd = name2;
                }
            }
        }
        fclose(m);
        for (ObjMaterial &mat : materials) {
            if (mat.mapKd.empty() || mat.texture) {
                continue;
            }
            std::string file = dir + mat.mapKd;
            std::replace(file.begin(), file.end(), '\\', '/');
            obj2voxel_texture *tex = obj2voxel_texture_alloc();
            if (obj2voxel_texture_load_from_file(tex, file.c_str(), "png")) {
                mat.texture = std::shared_ptr<void>(tex, [](void *p) {
                    // textures must outlive the job (the engine uploads them before the stream ends); keep them alive
                    // for the process lifetime like a caller-owned texture would be
                    (void) p;
                });
                logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "Loaded texture \"" + file + "\"");
            }
            else {
                obj2voxel_texture_free(tex);
                logMessage(OBJ2VOXEL_LOG_LEVEL_WARNING,
                           "Failed to load texture \"" + file + "\" of material \"" + mat.name + "\"");
            }
        }
    };

    char line[4096];
    while (fgets(line, sizeof line, f) != nullptr) {
        if (line[0] == 'v' && line[1] == ' ') {
            float x, y, z;
            if (sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3) {
                positions.insert(positions.end(), {x, y, z});
            }
        }
        else if (line[0] == 'v' && line[1] == 't') {
            float u = 0, v = 0;
            if (sscanf(line + 3, "%f %f", &u, &v) >= 1) {
                texcoords.insert(texcoords.end(), {u, v});
            }
        }
        else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            int vi[64], ti[64];
            int n = 0;
            const char *p = line + 1;
            while (n < 64) {
                while (*p == ' ' || *p == '\t') {
                    ++p;
                }
                if (*p == '\0' || *p == '\n' || *p == '\r') {
                    break;
                }
                int v = 0, t = 0, consumed = 0;
                if (sscanf(p, "%d%n", &v, &consumed) != 1) {
                    break;
                }
                p += consumed;
                if (*p == '/') {
                    ++p;
                    if (*p != '/' && sscanf(p, "%d%n", &t, &consumed) == 1) {
                        p += consumed;
                    }
                    if (*p == '/') {
                        ++p;
                        int nrm;
                        if (sscanf(p, "%d%n", &nrm, &consumed) == 1) {
                            p += consumed;
                        }
                    }
                }
                const int vcount = (int) (positions.size() / 3), tcount = (int) (texcoords.size() / 2);
                vi[n] = v > 0 ? v - 1 : vcount + v;
                ti[n] = t > 0 ? t - 1 : (t < 0 ? tcount + t : -1);
                ++n;
            }
            for (int k = 1; k + 1 < n; ++k) {
                const int idx[3] = {0, k, k + 1};
                float v[9], uv[6] = {0, 0, 0, 0, 0, 0};
                bool hasUv = true, valid = true;
                for (int c = 0; c < 3; ++c) {
                    const int a = vi[idx[c]];
                    valid = valid && a >= 0 && (size_t) a * 3 + 2 < positions.size();
                    if (!valid) {
                        break;
                    }
                    memcpy(v + c * 3, positions.data() + (size_t) a * 3, sizeof(float) * 3);
                    const int t = ti[idx[c]];
                    if (t >= 0 && (size_t) t * 2 + 1 < texcoords.size()) {
                        uv[c * 2] = texcoords[(size_t) t * 2];
                        uv[c * 2 + 1] = texcoords[(size_t) t * 2 + 1];
                    }
                    else {
                        hasUv = false;
                    }
                }
                if (!valid) {
                    continue;
                }
                if (current < 0) {
                    if (hasUv && defaultTexture != nullptr) {
                        append(v, uv, 3, nullptr, defaultTexture);
                    }
                    else {
                        append(v, nullptr, 1, nullptr, nullptr);
                    }
                }
                else {
                    const ObjMaterial &mat = materials[(size_t) current];
                    if (mat.texture && hasUv) {
                        append(v, uv, 3, nullptr, static_cast<const obj2voxel_texture *>(mat.texture.get()));
                    }
                    else {
                        append(v, nullptr, 2, mat.kd, nullptr);
                    }
                }
            }
        }
        else if (strncmp(line, "mtllib", 6) == 0) {
            char name[512];
            if (sscanf(line + 6, "%511s", name) == 1) {
                loadMtl(name);
            }
        }
        else if (strncmp(line, "usemtl", 6) == 0) {
            char name[512];
            current = -1;
            if (sscanf(line + 6, "%511s", name) == 1) {
                for (size_t i = 0; i < materials.size(); ++i) {
                    if (materials[i].name == name) {
                        current = (int) i;
                    }
                }
            }
        }
    }
    fclose(f);
    return true;
}

}  // namespace

bool readTriangleFile(const char *path, FileFormat format, const obj2voxel_texture *defaultTexture,
                      const TriangleAppender &append, std::string *error)
{
    if (format == FileFormat::STL) {
        return readStl(path, append, error);
    }
    if (format == FileFormat::OBJ) {
        return readObj(path, defaultTexture, append, error);
    }
    *error = "unsupported triangle file type";
    return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// PNG (8-bit, non-interlaced) -> RGBA8

namespace {

uint32_t readBe32(const uint8_t *p)
{
    return (uint32_t) p[0] << 24 | (uint32_t) p[1] << 16 | (uint32_t) p[2] << 8 | p[3];
}

uint8_t paeth(int a, int b, int c)
{
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (uint8_t) (pa <= pb && pa <= pc ? a : (pb <= pc ? b : c));
}

}  // namespace

bool decodePng(const uint8_t *data, size_t size, std::vector<uint8_t> *rgba, size_t *width, size_t *height,
               std::string *error)
{
    static const uint8_t signature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (size < 8 || memcmp(data, signature, 8) != 0) {
        *error = "not a PNG file";
        return false;
    }
    uint32_t w = 0, h = 0;
    int depth = 0, colorType = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, paletteAlpha;
    size_t pos = 8;
    while (pos + 12 <= size) {
        const uint32_t length = readBe32(data + pos);
        const uint8_t *type = data + pos + 4;
        const uint8_t *body = data + pos + 8;
        if (pos + 12 + (size_t) length > size) {
            break;
        }
        if (memcmp(type, "IHDR", 4) == 0 && length >= 13) {
            w = readBe32(body);
            h = readBe32(body + 4);
            depth = body[8];
            colorType = body[9];
            interlace = body[12];
        }
        else if (memcmp(type, "PLTE", 4) == 0) {
            palette.assign(body, body + length);
        }
        else if (memcmp(type, "tRNS", 4) == 0) {
            paletteAlpha.assign(body, body + length);
        }
        else if (memcmp(type, "IDAT", 4) == 0) {
            idat.insert(idat.end(), body, body + length);
        }
        else if (memcmp(type, "IEND", 4) == 0) {
            break;
        }
        pos += 12 + (size_t) length;
    }
    if (w == 0 || h == 0 || depth != 8 || interlace != 0) {
        *error = "unsupported PNG (only 8-bit non-interlaced images)";
        return false;
    }
    int channels;
    switch (colorType) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: *error = "unsupported PNG colour type"; return false;
    }
    const size_t stride = (size_t) w * channels;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawSize = (uLongf) raw.size();
    if (uncompress(raw.data(), &rawSize, idat.data(), (uLong) idat.size()) != Z_OK || rawSize != raw.size()) {
        *error = "PNG inflate failed";
        return false;
    }
    std::vector<uint8_t> image(stride * h);
    for (size_t y = 0; y < h; ++y) {
        const uint8_t filter = raw[y * (stride + 1)];
        const uint8_t *in = raw.data() + y * (stride + 1) + 1;
        uint8_t *out = image.data() + y * stride;
        const uint8_t *up = y != 0 ? out - stride : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t) channels ? out[x - channels] : 0;
            const int b = up != nullptr ? up[x] : 0;
            const int c = (up != nullptr && x >= (size_t) channels) ? up[x - channels] : 0;
            uint8_t v = in[x];
            switch (filter) {
            case 1: v = (uint8_t) (v + a); break;
            case 2: v = (uint8_t) (v + b); break;
            case 3: v = (uint8_t) (v + ((a + b) >> 1)); break;
            case 4: v = (uint8_t) (v + paeth(a, b, c)); break;
            default: break;
            }
            out[x] = v;
        }
    }
    rgba->resize((size_t) w * h * 4);
    for (size_t i = 0; i < (size_t) w * h; ++i) {
        const uint8_t *p = image.data() + i * channels;
        uint8_t *o = rgba->data() + i * 4;
        switch (colorType) {
        case 0: o[0] = o[1] = o[2] = p[0]; o[3] = 255; break;
        case 2: o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = 255; break;
        case 3: {
            const size_t k = p[0];
            o[0] = k * 3 + 2 < palette.size() ? palette[k * 3] : 0;
            o[1] = k * 3 + 2 < palette.size() ? palette[k * 3 + 1] : 0;
            o[2] = k * 3 + 2 < palette.size() ? palette[k * 3 + 2] : 0;
            o[3] = k < paletteAlpha.size() ? paletteAlpha[k] : 255;
            break;
        }
        case 4: o[0] = o[1] = o[2] = p[0]; o[3] = p[1]; break;
        default: o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3]; break;
        }
    }
    *width = w;
    *height = h;
    return true;
}

}  // namespace o2v
