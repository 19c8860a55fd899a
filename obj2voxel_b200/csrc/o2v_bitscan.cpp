// Host side of the bitmap download (o2v_job.cpp): one chunk's occupancy bitmap (4096 64-bit words: word = tile * 8 +
// layer, bit = x + 8 y inside the 8^3 tile, OUTPUT space) -> Voxel32 quads {x, y, z, 0xFFFFFFFF}.
//
// Two forms, chosen once at run time: a portable one (count trailing zeros, bit by bit) and one for CPUs with AVX-512
// VBMI2, where VPCOMPRESSB turns a word into the list of its set bit positions in one instruction and the quads are built
// sixteen at a time.  Reference counterpart: none — the reference's sink receives the quads of a chunk from the CPU
// voxelizer itself (src/obj2voxel.cpp:283-312); here they come out of a bitmap the device produced.
#include "o2v_job.h"

#include <stdlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace o2v {

namespace {

struct ChunkCursor {
    uint32_t *out;        // where the next quad goes
    uint32_t *const end;  // one past the buffer
};

/// Portable: appends the quads of `words[first .. last)` (tile-layer words of the chunk at output origin (cx, cy, cz));
/// stops and returns the index of the first word that did not fit entirely (the caller flushes and resumes there).
uint32_t scanWordsPortable(const unsigned long long *words, uint32_t first, uint32_t last, uint32_t cx, uint32_t cy,
                           uint32_t cz, ChunkCursor &cursor)
{
    for (uint32_t i = first; i < last; ++i) {
        unsigned long long w = words[i];
        if (w == 0) {
            continue;
        }
        if (cursor.out + 4 * (size_t) __builtin_popcountll(w) > cursor.end) {
            return i;
        }
        const uint32_t tile = i >> 3, layer = i & 7u;
        const uint32_t ox = cx + (tile & 7u) * kTileEdge, oy = cy + ((tile >> 3) & 7u) * kTileEdge;
        const uint32_t oz = cz + (tile >> 6) * kTileEdge + layer;
        do {
            const uint32_t b = (uint32_t) __builtin_ctzll(w);
            w &= w - 1;
            cursor.out[0] = ox + (b & 7u);
            cursor.out[1] = oy + (b >> 3);
            cursor.out[2] = oz;
            cursor.out[3] = 0xFFFFFFFFu;
            cursor.out += 4;
        } while (w != 0);
    }
    return last;
}

#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx512vbmi,avx512vbmi2")))
uint32_t scanWordsAvx512(const unsigned long long *words, uint32_t first, uint32_t last, uint32_t cx, uint32_t cy,
                         uint32_t cz, ChunkCursor &cursor)
{
    const __m512i iota = _mm512_set_epi8(63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43,
                                         42, 41, 40, 39, 38, 37, 36, 35, 34, 33, 32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22,
                                         21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    // lanes of a 4-quad vector: [x0 y0 z a | x1 y1 z a | x2 y2 z a | x3 y3 z a]; x from X (0..15), y from Y (16..31)
    const __m512i pick0 = _mm512_set_epi32(0, 0, 19, 3, 0, 0, 18, 2, 0, 0, 17, 1, 0, 0, 16, 0);
    const __m512i step = _mm512_set_epi32(0, 0, 4, 4, 0, 0, 4, 4, 0, 0, 4, 4, 0, 0, 4, 4);
    const __m512i seven = _mm512_set1_epi32(7);
    for (uint32_t i = first; i < last; ++i) {
        const unsigned long long w = words[i];
        if (w == 0) {
            continue;
        }
        uint32_t n = (uint32_t) __builtin_popcountll(w);
        if (cursor.out + 4 * (size_t) n > cursor.end) {
            return i;
        }
        const uint32_t tile = i >> 3, layer = i & 7u;
        const __m512i ox = _mm512_set1_epi32((int) (cx + (tile & 7u) * kTileEdge));
        const __m512i oy = _mm512_set1_epi32((int) (cy + ((tile >> 3) & 7u) * kTileEdge));
        const int oz = (int) (cz + (tile >> 6) * kTileEdge + layer);
        const __m512i za = _mm512_set_epi32(-1, oz, 0, 0, -1, oz, 0, 0, -1, oz, 0, 0, -1, oz, 0, 0);
        const __m512i positions = _mm512_maskz_compress_epi8((__mmask64) w, iota);  // the set bits' positions, packed
        alignas(64) unsigned char bytes[64];
        _mm512_store_si512(bytes, positions);
        for (uint32_t done = 0; done < n; done += 16) {  // sixteen quads per round (one round for all but dense words)
            const __m512i b = _mm512_cvtepu8_epi32(_mm_load_si128(reinterpret_cast<const __m128i *>(bytes + done)));
            const __m512i X = _mm512_add_epi32(_mm512_and_si512(b, seven), ox);
            const __m512i Y = _mm512_add_epi32(_mm512_srli_epi32(b, 3), oy);
            __m512i pick = pick0;
            const uint32_t left = n - done < 16u ? n - done : 16u;
            for (uint32_t q = 0; q < left; q += 4) {
                const __m512i xy = _mm512_permutex2var_epi32(X, pick, Y);
                const __m512i quads = _mm512_mask_blend_epi32(0xCCCC, xy, za);
                const uint32_t remaining = left - q;
                const __mmask16 keep = remaining >= 4u ? (__mmask16) 0xFFFF : (__mmask16) ((1u << (4u * remaining)) - 1u);
                _mm512_mask_storeu_epi32(cursor.out, keep, quads);
                cursor.out += 4 * (remaining >= 4u ? 4u : remaining);
                pick = _mm512_add_epi32(pick, step);
            }
        }
    }
    return last;
}
#endif

using ScanFn = uint32_t (*)(const unsigned long long *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, ChunkCursor &);

ScanFn chooseScan()
{
    if (const char *env = getenv("O2V_B200_PORTABLE_SCAN")) {  // measurement / tests
        if (atoi(env) != 0) {
            return scanWordsPortable;
        }
    }
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
        __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512vbmi2")) {
        return scanWordsAvx512;
    }
#endif
    return scanWordsPortable;
}

}  // namespace

bool hostHasFastBitScan()
{
    static const bool fast = chooseScan() != scanWordsPortable;
    return fast;
}

unsigned long long scanChunkBitmap(const unsigned long long *words, uint32_t cx, uint32_t cy, uint32_t cz, uint32_t *buffer,
                                   uint32_t bufferQuads, const std::function<bool(uint32_t *, size_t)> &flush)
{
    static const ScanFn scan = chooseScan();
    unsigned long long total = 0;
    ChunkCursor cursor{buffer, buffer + (size_t) bufferQuads * 4};
    uint32_t at = 0;
    while (at < kChunkWords) {
        at = scan(words, at, kChunkWords, cx, cy, cz, cursor);
        if (at < kChunkWords || cursor.out != buffer) {
            // the buffer is full (or the chunk is done): hand it on and start over
            const size_t quads = (size_t) (cursor.out - buffer) / 4;
            total += quads;
            if (quads != 0 && !flush(buffer, quads)) {
                return ~0ull;
            }
            cursor.out = buffer;
        }
    }
    return total;
}

}  // namespace o2v
