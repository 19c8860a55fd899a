// Host-side I/O adapters (see o2v_io.h).  None of this is on the GPU hot path; it brackets it.
#include "o2v_io.h"

#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <array>
#include <map>
#include <unordered_map>

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
            // the reference's own comment line, so that the 299-byte header is byte-identical to what voxelio writes
            // (voxelio/src/format/ply.cpp:24): a reader keyed to the fixed header length of obj2voxel's PLY files keeps working
            writeText("comment generated by voxel-io: a C++ library by Jan \"Eisenwave\" Schultke\r\n");
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

// ---- palette formats: QEF (text) and VOX (MagicaVoxel) -------------------------------------------------------------
//
// Like the reference's VoxelioVoxelSink (src/io.cpp:524-636) these formats need the whole palette before the first byte
// can be written: every voxel is buffered, colours are indexed in first-seen order (Palette32::insert,
// voxelio/include/voxelio/palette.hpp:74-80), and finalize() writes the file.  File layouts follow
// voxelio/src/format/qef.cpp:212-295 and voxelio/src/format/vox.cpp:828-1080 (version 150, 256^3 models, one nTRN + nSHP
// per model under a single nGRP, no LAYR chunk, RGBA chunk last).  Deviation, stated: when a VOX needs more than 255
// colours the reference reduces its palette with a randomly seeded k-means (voxelio/src/palette.cpp); here the reduction
// is a deterministic median cut (each colour maps to the mean of its box) — the same contract, other representatives.

struct PaletteBuilder {
    std::vector<uint32_t> colors;  // index -> argb
    std::unordered_map<uint32_t, uint32_t> indexOf;

    uint32_t insert(uint32_t argb)
    {
        auto found = indexOf.find(argb);
        if (found != indexOf.end()) {
            return found->second;
        }
        const uint32_t index = (uint32_t) colors.size();
        indexOf.emplace(argb, index);
        colors.push_back(argb);
        return index;
    }
};

/// Median cut of `colors` (argb) into at most `limit` boxes; returns per input colour the index of its box and fills
/// `representatives` with the channel-wise mean of each box.
std::vector<uint32_t> medianCut(const std::vector<uint32_t> &colors, size_t limit, std::vector<uint32_t> *representatives)
{
    struct Box {
        std::vector<uint32_t> members;  // indices into colors
    };
    auto channel = [&](uint32_t index, int c) { return (int) ((colors[index] >> (8 * c)) & 255u); };
    auto widest = [&](const Box &box, int *axis) {
        int best = -1;
        for (int c = 0; c < 4; ++c) {
            int lo = 255, hi = 0;
            for (uint32_t m : box.members) {
                lo = std::min(lo, channel(m, c));
                hi = std::max(hi, channel(m, c));
            }
            if (hi - lo > best) {
                best = hi - lo;
                *axis = c;
            }
        }
        return best;
    };
    std::vector<Box> boxes(1);
    boxes[0].members.resize(colors.size());
    for (uint32_t i = 0; i < colors.size(); ++i) {
        boxes[0].members[i] = i;
    }
    while (boxes.size() < limit) {
        size_t pick = boxes.size();
        int pickRange = 0, pickAxis = 0;
        for (size_t b = 0; b < boxes.size(); ++b) {
            int axis = 0;
            const int range = boxes[b].members.size() > 1 ? widest(boxes[b], &axis) : 0;
            if (range > pickRange) {
                pick = b;
                pickRange = range;
                pickAxis = axis;
            }
        }
        if (pick == boxes.size()) {
            break;  // every box holds one colour (or identical colours)
        }
        std::vector<uint32_t> &m = boxes[pick].members;
        std::sort(m.begin(), m.end(), [&](uint32_t a, uint32_t b) {
            const int ca = channel(a, pickAxis), cb = channel(b, pickAxis);
            return ca != cb ? ca < cb : a < b;
        });
        Box upper;
        upper.members.assign(m.begin() + (long) (m.size() / 2), m.end());
        m.resize(m.size() / 2);
        boxes.push_back(std::move(upper));
    }
    std::vector<uint32_t> mapping(colors.size(), 0);
    representatives->clear();
    for (size_t b = 0; b < boxes.size(); ++b) {
        uint64_t sum[4] = {0, 0, 0, 0};
        for (uint32_t m : boxes[b].members) {
            for (int c = 0; c < 4; ++c) {
                sum[c] += (uint64_t) channel(m, c);
            }
            mapping[m] = (uint32_t) b;
        }
        uint32_t mean = 0;
        const uint64_t n = std::max<uint64_t>(boxes[b].members.size(), 1);
        for (int c = 0; c < 4; ++c) {
            mean |= (uint32_t) ((sum[c] + n / 2) / n) << (8 * c);
        }
        representatives->push_back(mean);
    }
    return mapping;
}

/// voxelio stringifyFractionRpad (src/stringify.cpp:9-38): truncated decimal expansion, right-padded with zeros.
std::string fractionRpad(uint32_t num, uint32_t den, unsigned precision)
{
    std::string result = std::to_string(num / den);
    num %= den;
    result.push_back('.');
    for (unsigned i = 0; i < precision; ++i) {
        num *= 10;
        result.push_back((char) ('0' + num / den));
        num %= den;
    }
    return result;
}

class PaletteSink final : public VoxelSink {
public:
    PaletteSink(FileFormat format, FILE *file, uint32_t resolution)
        : format_{format}, stream_{file}, resolution_{resolution}
    {
    }

    bool write(uint32_t *quads, size_t count) override
    {
        voxels_.insert(voxels_.end(), quads, quads + count * 4);
        for (size_t i = 0; i < count; ++i) {
            palette_.insert(quads[i * 4 + 3]);
        }
        return stream_.good();
    }

    void finalize() override
    {
        if (finalized_) {
            return;
        }
        finalized_ = true;
        if (format_ == FileFormat::QEF) {
            writeQef();
        }
        else {
            writeVox();
        }
        stream_.flush();
        voxels_.clear();
        voxels_.shrink_to_fit();
    }

    bool good() const override { return stream_.good() && !failed_; }
    size_t voxelsWritten() const override { return finalized_ ? written_ : voxels_.size() / 4; }
    const std::vector<uint8_t> *memory() const override { return stream_.memory(); }

private:
    void text(const std::string &s) { stream_.write(s.data(), s.size()); }
    void u32le(uint32_t v) { stream_.write(&v, 4); }  // little-endian host (x86-64 / aarch64)
    void fourcc(const char *id) { stream_.write(id, 4); }
    void vstring(const std::string &s)
    {
        u32le((uint32_t) s.size());
        text(s);
    }
    void chunkHeader(const char *id, uint32_t selfSize, uint32_t childSize = 0)
    {
        fourcc(id);
        u32le(selfSize);
        u32le(childSize);
    }

    void writeQef()
    {
        const size_t count = voxels_.size() / 4;
        text("Qubicle Exchange Format\nVersion 0.2\nwww.minddesk.com\n");
        text(std::to_string(resolution_) + ' ' + std::to_string(resolution_) + ' ' + std::to_string(resolution_) + '\n');
        text(std::to_string(palette_.colors.size()) + '\n');
        for (uint32_t argb : palette_.colors) {
            text(fractionRpad((argb >> 16) & 255u, 255u, 4) + ' ' + fractionRpad((argb >> 8) & 255u, 255u, 4) + ' ' +
                 fractionRpad(argb & 255u, 255u, 4) + '\n');
        }
        std::string block;
        for (size_t i = 0; i < count; ++i) {
            const uint32_t *q = voxels_.data() + i * 4;
            // the reference rejects positions outside [0, resolution] (qef.cpp:297-312): WRITE_ERROR_POSITION_OUT_OF_BOUNDS
            if ((int32_t) q[0] < 0 || (int32_t) q[1] < 0 || (int32_t) q[2] < 0 || q[0] > resolution_ ||
                q[1] > resolution_ || q[2] > resolution_) {
                failed_ = true;
                break;
            }
            block += std::to_string((int32_t) q[0]) + ' ' + std::to_string((int32_t) q[1]) + ' ' +
                     std::to_string((int32_t) q[2]) + ' ' + std::to_string(palette_.indexOf[q[3]]) + '\n';
            if (block.size() > (1u << 20)) {
                text(block);
                block.clear();
            }
            ++written_;
        }
        text(block);
    }

    void nodeTransform(uint32_t id, uint32_t child, const int32_t t[3])
    {
        const std::string translation = std::to_string(t[0]) + ' ' + std::to_string(t[1]) + ' ' + std::to_string(t[2]);
        chunkHeader("nTRN", 11 * 4 + 2 + 1 + 2 + (uint32_t) translation.size());
        u32le(id);
        u32le(0);            // node attributes: empty dict
        u32le(child);
        u32le(0xffffffffu);  // reserved id, must be -1
        u32le(0);            // layer
        u32le(1);            // frames, must be 1
        u32le(2);            // frame dict
        vstring("_r");
        vstring("4");        // identity rotation
        vstring("_t");
        vstring(translation);
    }

    void writeVox()
    {
        const size_t count = voxels_.size() / 4;
        // palette: at most 255 usable entries (index 0 is reserved)
        std::vector<uint32_t> reduced = palette_.colors;
        std::vector<uint32_t> mapping(palette_.colors.size());
        for (uint32_t i = 0; i < mapping.size(); ++i) {
            mapping[i] = i;
        }
        if (palette_.colors.size() > 255) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_DEBUG, "Reducing palette from " + std::to_string(palette_.colors.size()) +
                                                      " to 255 colors ...");
            mapping = medianCut(palette_.colors, 255, &reduced);
        }
        // models: 256^3 chunks in ascending (x, y, z) order
        std::map<std::array<int32_t, 3>, std::vector<uint32_t>> models;
        for (size_t i = 0; i < count; ++i) {
            const uint32_t *q = voxels_.data() + i * 4;
            const int32_t p[3] = {(int32_t) q[0], (int32_t) q[1], (int32_t) q[2]};
            std::array<int32_t, 3> key;
            uint32_t xyzi = 0;
            for (int a = 0; a < 3; ++a) {
                key[(size_t) a] = p[a] >= 0 ? p[a] / 256 : -((255 - p[a]) / 256);  // floor division
                xyzi |= (uint32_t) (p[a] - key[(size_t) a] * 256) << (8 * a);      // bytes x, y, z, i in file order
            }
            xyzi |= (mapping[palette_.indexOf[q[3]]] + 1u) << 24;
            models[key].push_back(xyzi);
            ++written_;
        }

        text("VOX ");
        u32le(150);
        chunkHeader("MAIN", 0, 0);  // child size patched below
        for (const auto &model : models) {
            chunkHeader("SIZE", 12);
            u32le(256);
            u32le(256);
            u32le(256);
            chunkHeader("XYZI", (uint32_t) (model.second.size() + 1) * 4);
            u32le((uint32_t) model.second.size());
            stream_.write(model.second.data(), model.second.size() * 4);
        }
        const int32_t zero[3] = {0, 0, 0};
        nodeTransform(0, 1, zero);
        chunkHeader("nGRP", (3 + (uint32_t) models.size()) * 4);
        u32le(1);
        u32le(0);
        u32le((uint32_t) models.size());
        for (uint32_t k = 0; k < models.size(); ++k) {
            u32le(2 + 2 * k);
        }
        uint32_t k = 0;
        for (const auto &model : models) {
            const int32_t t[3] = {model.first[0] * 256 + 128, model.first[1] * 256 + 128, model.first[2] * 256 + 128};
            nodeTransform(2 + 2 * k, 3 + 2 * k, t);
            chunkHeader("nSHP", 5 * 4);
            u32le(3 + 2 * k);
            u32le(0);
            u32le(1);
            u32le(k);
            u32le(0);
            ++k;
        }
        chunkHeader("RGBA", 1024);
        uint8_t rgba[1024];
        memset(rgba, 0, sizeof rgba);
        for (size_t i = 0; i < reduced.size() && i < 255; ++i) {  // RGBA entry i is palette index i + 1
            rgba[i * 4 + 0] = (uint8_t) (reduced[i] >> 16);
            rgba[i * 4 + 1] = (uint8_t) (reduced[i] >> 8);
            rgba[i * 4 + 2] = (uint8_t) reduced[i];
            rgba[i * 4 + 3] = (uint8_t) (reduced[i] >> 24);
        }
        stream_.write(rgba, sizeof rgba);
        const size_t end = stream_.position();
        stream_.seek(16);
        u32le((uint32_t) (end - 20));
        stream_.seek(end);
    }

    FileFormat format_;
    ByteStream stream_;
    uint32_t resolution_;
    std::vector<uint32_t> voxels_;
    PaletteBuilder palette_;
    size_t written_ = 0;
    bool finalized_ = false;
    bool failed_ = false;
};

}  // namespace

std::unique_ptr<VoxelSink> makeCallbackSink(obj2voxel_voxel_callback *callback, void *data)
{
    return std::unique_ptr<VoxelSink>(new CallbackSink{callback, data});
}

std::unique_ptr<VoxelSink> makeFormatSink(FileFormat format, const char *path, uint32_t resolution, std::string *error)
{
    const bool palette = format == FileFormat::QEF || format == FileFormat::VOX;
    if (!palette && format != FileFormat::VL32 && format != FileFormat::PLY && format != FileFormat::XYZRGB) {
        *error = "not a voxel output format (supported: vl32, ply, xyzrgb, qef, vox)";
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
    if (palette) {
        return std::unique_ptr<VoxelSink>(new PaletteSink{format, file, resolution});
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
    obj2voxel_texture *texture = nullptr;  // allocated through the C API; the job that called the reader frees it
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
             std::vector<obj2voxel_texture *> *loadedTextures, std::string *error)
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
        char *line = nullptr;  // getline: lines of any length
        size_t lineCapacity = 0;
        while (getline(&line, &lineCapacity, m) >= 0) {
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
        free(line);
        fclose(m);
        for (ObjMaterial &mat : materials) {
            if (mat.mapKd.empty() || mat.texture != nullptr) {
                continue;
            }
            std::string file = dir + mat.mapKd;
            std::replace(file.begin(), file.end(), '\\', '/');
            obj2voxel_texture *tex = obj2voxel_texture_alloc();
            if (obj2voxel_texture_load_from_file(tex, file.c_str(), "png")) {
                mat.texture = tex;  // lives until the job has uploaded it: handed to the caller, who frees it
                loadedTextures->push_back(tex);
                logMessage(OBJ2VOXEL_LOG_LEVEL_INFO, "Loaded texture \"" + file + "\"");
            }
            else {
                obj2voxel_texture_free(tex);
                logMessage(OBJ2VOXEL_LOG_LEVEL_WARNING,
                           "Failed to load texture \"" + file + "\" of material \"" + mat.name + "\"");
            }
        }
    };

    char *line = nullptr;  // getline: an n-gon's `f` line may be longer than any fixed buffer
    size_t lineCapacity = 0;
    std::vector<int> vi, ti;
    while (getline(&line, &lineCapacity, f) >= 0) {
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
            vi.clear();
            ti.clear();
            const char *p = line + 1;
            for (;;) {
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
                vi.push_back(v > 0 ? v - 1 : vcount + v);
                ti.push_back(t > 0 ? t - 1 : (t < 0 ? tcount + t : -1));
            }
            const int n = (int) vi.size();
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
                    if (mat.texture != nullptr && hasUv) {
                        append(v, uv, 3, nullptr, mat.texture);
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
    free(line);
    fclose(f);
    return true;
}

}  // namespace

bool readTriangleFile(const char *path, FileFormat format, const obj2voxel_texture *defaultTexture,
                      const TriangleAppender &append, std::vector<obj2voxel_texture *> *loadedTextures, std::string *error)
{
    if (format == FileFormat::STL) {
        return readStl(path, append, error);
    }
    if (format == FileFormat::OBJ) {
        return readObj(path, defaultTexture, append, loadedTextures, error);
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
