// TEST SHIM (not a product path): compiles the exact-arithmetic header for the host so that tests/test_hostmath.py can
// compare the very functions the kernels run against the oracle on a machine without a GPU.
#include <string.h>

#include "o2v_exact.cuh"
#include "o2v_sat.cuh"

using namespace o2v;

extern "C" {

int o2vt_clip_voxel(const float tri15[15], const uint32_t pos[3], float wholeArea, int textured, float outWuv[3])
{
    ClipResult r;
    if (textured) {
        Tri<true> t;
        memcpy(t.v, tri15, sizeof t.v);
        memcpy(t.t, tri15 + 9, sizeof t.t);
        r = clipLeafInVoxel<true>(t, pos[0], pos[1], pos[2], wholeArea);
    }
    else {
        Tri<false> t;
        memcpy(t.v, tri15, sizeof t.v);
        r = clipLeafInVoxel<false>(t, pos[0], pos[1], pos[2], wholeArea);
    }
    outWuv[0] = r.weight;
    outWuv[1] = r.u;
    outWuv[2] = r.v;
    return r.pieces;
}

size_t o2vt_subdivide(const float tri15[15], float *outLeaves, size_t cap)
{
    Tri<true> t;
    memcpy(t.v, tri15, sizeof t.v);
    memcpy(t.t, tri15 + 9, sizeof t.t);
    size_t count = 0;
    forEachLeaf<true>(t, [&](const Tri<true> &leaf) {
        if (count < cap) {
            memcpy(outLeaves + count * 15, leaf.v, sizeof leaf.v);
            memcpy(outLeaves + count * 15 + 9, leaf.t, sizeof leaf.t);
        }
        ++count;
    });
    return count;
}

/// The leaves of forEachLeaf, produced the way the huge-triangle kernels produce them: subtree after subtree
/// (forEachLeafOfSubtree, 4^depth subtrees).  Must equal o2vt_subdivide for a triangle that is not axis-aligned.
size_t o2vt_subdivide_by_subtrees(const float tri15[15], int depth, float *outLeaves, size_t cap)
{
    Tri<true> t;
    memcpy(t.v, tri15, sizeof t.v);
    memcpy(t.t, tri15 + 9, sizeof t.t);
    size_t count = 0;
    for (uint32_t subtree = 0; subtree < (1u << (2 * depth)); ++subtree) {
        forEachLeafOfSubtree<true>(t, depth, subtree, [&](const Tri<true> &leaf) {
            if (count < cap) {
                memcpy(outLeaves + count * 15, leaf.v, sizeof leaf.v);
                memcpy(outLeaves + count * 15 + 9, leaf.t, sizeof leaf.t);
            }
            ++count;
        });
    }
    return count;
}

float o2vt_area(const float v[9])
{
    return triArea(v);
}

int o2vt_aligned(const float v[9])
{
    return triRoughlyAxisAligned(v) ? 1 : 0;
}

void o2vt_transform(const float m[12], const float in[3], float out[3])
{
    affineApply(m, in, out);
}

void o2vt_texture_lookup(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t channels, uint32_t wrap, float u,
                         float v, float rgb[3])
{
    const TextureView tex{pixels, w, h, channels, wrap};
    textureLookup(tex, u, v, rgb);
}

uint32_t o2vt_quantize(const float rgb[3])
{
    return quantizeArgb(rgb[0], rgb[1], rgb[2]);
}

void o2vt_combine(float acc[4], const float incoming[4], int blend)
{
    WeightedColor c{acc[0], acc[1], acc[2], acc[3]};
    combineColorInto(c, incoming[0], incoming[1], incoming[2], incoming[3], blend != 0);
    acc[0] = c.w;
    acc[1] = c.r;
    acc[2] = c.g;
    acc[3] = c.b;
}

/// Sweeps every voxel of every leaf's AABB (leaves: n x 9 floats, voxel space) through the SAT users of the kernels
///   * prefilterPass with tile-relative constants (weighted path: o2v_sparse.cu / o2v_kernels.cu),
///   * buildRowSat / rowSpanMisses / classifyInRow (and classifyVoxel) with constants relative to the leaf's box — or to
///     its 16^3 sub-boxes when the box holds more than 4096 voxels — exactly as the thread-per-leaf classifier of
///     o2v_occupancy.cu stages them,
///   * classifySpan (the row-interval form of the block classifier) with the same constants,
/// and through the reference semantics (plane-distance cull + exact clip, o2v_exact.cuh), and counts disagreements.
/// certainMargin: the `certain` shrink (certainMarginFor(S) in the kernels).
/// out: [0] pairs, [1] miss, [2] uncertain, [3] certain, [4] reference hits, [5] `miss` verdicts (either per-voxel user)
/// the reference hits (must be 0), [6] `certain` verdicts the reference does not hit (must be 0), [7] leaves skipped,
/// [8] span: miss, [9] span: uncertain, [10] span: certain, [11] span `miss` the reference hits (must be 0),
/// [12] span `certain` the reference does not hit (must be 0), [13] voxels where span and per-voxel verdicts differ.
void o2vt_classify_fuzz(const float *leaves, size_t n, unsigned long long maxVolume, float certainMargin,
                        unsigned long long out[16])
{
    for (int i = 0; i < 16; ++i) {
        out[i] = 0;
    }
    for (size_t l = 0; l < n; ++l) {
        const float *v = leaves + l * 9;
        uint32_t lo[3], hi[3];
        triVoxelBounds(v, lo, hi);
        const unsigned long long volume =
            (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (unsigned long long) (hi[2] - lo[2]);
        if (!(triArea(v) > 0.0f) || volume > maxVolume) {
            ++out[7];
            continue;
        }
        Tri<false> tri;
        memcpy(tri.v, v, sizeof tri.v);
        const uint32_t flags = leafFlagsOf(v);
        const uint32_t boxEdge = volume > 4096 ? 16u : 0xffffffffu;  // o2v_occupancy.cu: kOccBigVolume, kOccBoxEdge
        for (uint32_t z = lo[2]; z < hi[2]; ++z) {
            for (uint32_t y = lo[1]; y < hi[1]; ++y) {
                // span form: once per row and box
                int spanVerdict[4096 + 16];
                for (uint32_t x = lo[0]; x < hi[0]; ++x) {
                    // weighted path: constants relative to the voxel's 8^3 tile
                    const float tileOrigin[3] = {(float) (x & ~7u), (float) (y & ~7u), (float) (z & ~7u)};
                    LeafStage ts;
                    memcpy(ts.v, v, sizeof ts.v);
                    ts.flags = flags;
                    buildPrefilter(ts, tileOrigin);
                    const bool pass = prefilterPass(ts, (float) (x & 7u), (float) (y & 7u), (float) (z & 7u));
                    // occupancy path: constants relative to the leaf's box or 16^3 sub-box
                    uint32_t bo[3] = {lo[0], lo[1], lo[2]};
                    if (boxEdge != 0xffffffffu) {
                        bo[0] += (x - lo[0]) / boxEdge * boxEdge;
                        bo[1] += (y - lo[1]) / boxEdge * boxEdge;
                        bo[2] += (z - lo[2]) / boxEdge * boxEdge;
                    }
                    const float boxOrigin[3] = {(float) bo[0], (float) bo[1], (float) bo[2]};
                    LeafStage bs;
                    memcpy(bs.v, v, sizeof bs.v);
                    bs.flags = flags;
                    buildPrefilter(bs, boxOrigin);
                    PairSat sat;
                    buildPairSat(sat, bs, boxOrigin, certainMargin);
                    // as the thread-per-leaf classifier does it: the row's 8-aligned x segment is tested as a whole first
                    const uint32_t boxHiX = boxEdge != 0xffffffffu && bo[0] + boxEdge < hi[0] ? bo[0] + boxEdge : hi[0];
                    const uint32_t segFirst = (x & ~7u) > bo[0] ? (x & ~7u) : bo[0];
                    const uint32_t segLast = ((x & ~7u) + 8u < boxHiX ? (x & ~7u) + 8u : boxHiX) - 1u;
                    RowSat row;
                    buildRowSat(sat, (float) (y - bo[1]), (float) (z - bo[2]), row);
                    int verdict = kSatUncertain, span = kSatUncertain;
                    if ((flags & kLeafNoPrefilter) == 0) {
                        const float spanFirst = (float) (segFirst - bo[0]), spanLast = (float) (segLast - bo[0]);
                        verdict = (rowPlaneSpanMisses(sat, row, spanFirst, spanLast) ||
                                   rowSpanMisses(sat, row, spanFirst, spanLast))
                                      ? (int) kSatMiss
                                      : classifyInRow(sat, row, (float) (x - bo[0]));
                        if (verdict != classifyVoxel(sat, (float) (x - bo[0]), (float) (y - bo[1]),
                                                     (float) (z - bo[2]))) {
                            ++out[5];  // the segment test must never reject what the voxel's own test accepts
                        }
                        if (x == bo[0]) {  // first voxel of this box's row: solve the row
                            SpanSat ss;
                            buildSpanSat(ss, sat);
                            int i0, i1, j0, j1;
                            classifySpan(ss, (float) (y - bo[1]), (float) (z - bo[2]), (float) (boxHiX - 1u - bo[0]), i0, i1,
                                         j0, j1);
                            for (uint32_t xx = bo[0]; xx < boxHiX; ++xx) {
                                const int lx = (int) (xx - bo[0]);
                                spanVerdict[xx - lo[0]] = (lx < i0 || lx > i1) ? (int) kSatMiss
                                                          : (lx >= j0 && lx <= j1) ? (int) kSatCertain
                                                                                   : (int) kSatUncertain;
                            }
                        }
                        span = spanVerdict[x - lo[0]];
                    }
                    const bool hit =
                        !planeDistanceCulled(v, x, y, z) && clipLeafInVoxel<false>(tri, x, y, z, 1.0f).pieces != 0;
                    ++out[0];
                    ++out[1 + verdict];
                    out[4] += hit ? 1 : 0;
                    out[5] += ((verdict == kSatMiss || !pass) && hit) ? 1 : 0;
                    out[6] += (verdict == kSatCertain && !hit) ? 1 : 0;
                    ++out[8 + span];
                    out[11] += (span == kSatMiss && hit) ? 1 : 0;
                    out[12] += (span == kSatCertain && !hit) ? 1 : 0;
                    out[13] += span != verdict ? 1 : 0;
                }
            }
        }
    }
}


// ---- study of two leads for the classify kernel, CPU only: nothing below is used by a kernel ----

namespace {

/// Every edge function divided by its `certain` threshold: miss <=> value < 0, sure <=> value >= 1, so that the six (plus
/// three) values of a voxel reduce to one minimum.  A degenerate edge (a = b = 0: value = c everywhere) becomes the
/// constant 1 (c >= 0), -1 (c < 0) or 0.5 (NaN: neither verdict).
struct ScaledSat {
    float a[9], b[9], c[9];
};

void buildScaledSat(ScaledSat &out, const PairSat &s)
{
    for (int k = 0; k < 9; ++k) {
        const SatEdge e = s.edge[k];
        if (e.k > 0.0f) {
            out.a[k] = e.a / e.k;
            out.b[k] = e.b / e.k;
            out.c[k] = e.c / e.k;
        }
        else {
            out.a[k] = out.b[k] = 0.0f;
            out.c[k] = e.c < 0.0f ? -1.0f : (e.c >= 0.0f ? 1.0f : 0.5f);
        }
    }
}

int classifyScaled(const PairSat &s, const ScaledSat &t, float lx, float ly, float lz, float certainMargin)
{
    const float dist = fabsf(s.plane[0] * lx + s.plane[1] * ly + s.plane[2] * lz + s.plane[3]);
    const float q[3] = {lx, ly, lz};
    float least = 2.0f;
    for (int proj = 0; proj < 3; ++proj) {
        const float qa = q[proj], qb = q[(proj + 1) % 3];
        for (int i = 0; i < 3; ++i) {
            const int k = proj * 3 + i;
            least = fminf(least, t.a[k] * qa + t.b[k] * qb + t.c[k]);
        }
    }
    if (dist > s.planeLimit || least < 0.0f) {
        return kSatMiss;
    }
    bool sure = least >= 1.0f && dist <= s.planeSure;
    for (int a = 0; a < 3; ++a) {
        sure = sure && q[a] <= s.hi[a] && q[a] >= s.lo[a];
    }
    return sure ? kSatCertain : kSatUncertain;
}

/// buildPairSat with the `certain` shrink as a parameter (the header's is the constant kCertainMargin).
void buildPairSatWith(PairSat &out, const LeafStage &s, const float origin[3], float certainMargin)
{
    buildPairSat(out, s, origin, certainMargin);
}

}  // namespace

/// The sweep of o2vt_classify_fuzz with the `certain` shrink given at run time and, if scaled != 0, the scaled-minimum
/// form of the edge tests.  out: [0] pairs, [1] miss, [2] uncertain, [3] certain, [4] reference hits, [5] `miss` verdicts
/// the reference hits, [6] `certain` verdicts the reference does not hit, [7] leaves skipped.
void o2vt_classify_study(const float *leaves, size_t n, unsigned long long maxVolume, float certainMargin, int scaled,
                         unsigned long long out[8])
{
    for (int i = 0; i < 8; ++i) {
        out[i] = 0;
    }
    for (size_t l = 0; l < n; ++l) {
        const float *v = leaves + l * 9;
        uint32_t lo[3], hi[3];
        triVoxelBounds(v, lo, hi);
        const unsigned long long volume =
            (unsigned long long) (hi[0] - lo[0]) * (hi[1] - lo[1]) * (unsigned long long) (hi[2] - lo[2]);
        if (!(triArea(v) > 0.0f) || volume > maxVolume) {
            ++out[7];
            continue;
        }
        Tri<false> tri;
        memcpy(tri.v, v, sizeof tri.v);
        const uint32_t flags = leafFlagsOf(v);
        const uint32_t boxEdge = volume > 4096 ? 16u : 0xffffffffu;
        for (uint32_t z = lo[2]; z < hi[2]; ++z) {
            for (uint32_t y = lo[1]; y < hi[1]; ++y) {
                for (uint32_t x = lo[0]; x < hi[0]; ++x) {
                    uint32_t bo[3] = {lo[0], lo[1], lo[2]};
                    if (boxEdge != 0xffffffffu) {
                        bo[0] += (x - lo[0]) / boxEdge * boxEdge;
                        bo[1] += (y - lo[1]) / boxEdge * boxEdge;
                        bo[2] += (z - lo[2]) / boxEdge * boxEdge;
                    }
                    const float boxOrigin[3] = {(float) bo[0], (float) bo[1], (float) bo[2]};
                    LeafStage bs;
                    memcpy(bs.v, v, sizeof bs.v);
                    bs.flags = flags;
                    buildPrefilter(bs, boxOrigin);
                    PairSat sat;
                    buildPairSatWith(sat, bs, boxOrigin, certainMargin);
                    int verdict = kSatUncertain;
                    if ((flags & kLeafNoPrefilter) == 0) {
                        const float lx = (float) (x - bo[0]), ly = (float) (y - bo[1]), lz = (float) (z - bo[2]);
                        if (scaled != 0) {
                            ScaledSat t;
                            buildScaledSat(t, sat);
                            verdict = classifyScaled(sat, t, lx, ly, lz, certainMargin);
                        }
                        else {  // the header's form with the run-time margin in the box-normal tests
                            verdict = classifyVoxel(sat, lx, ly, lz);
                            if (verdict != kSatMiss) {
                                const float q[3] = {lx, ly, lz};
                                bool sure = fabsf(sat.plane[0] * lx + sat.plane[1] * ly + sat.plane[2] * lz + sat.plane[3]) <=
                                            sat.planeSure;
                                for (int k = 0; k < 9; ++k) {
                                    const int proj = k / 3;
                                    const float value = sat.edge[k].a * q[proj] + sat.edge[k].b * q[(proj + 1) % 3] +
                                                        sat.edge[k].c;
                                    sure = sure && value >= sat.edge[k].k;
                                }
                                for (int a = 0; a < 3; ++a) {
                                    sure = sure && q[a] <= sat.hi[a] && q[a] >= sat.lo[a];
                                }
                                verdict = sure ? kSatCertain : kSatUncertain;
                            }
                        }
                    }
                    const bool hit =
                        !planeDistanceCulled(v, x, y, z) && clipLeafInVoxel<false>(tri, x, y, z, 1.0f).pieces != 0;
                    ++out[0];
                    ++out[1 + verdict];
                    out[4] += hit ? 1 : 0;
                    out[5] += (verdict == kSatMiss && hit) ? 1 : 0;
                    out[6] += (verdict == kSatCertain && !hit) ? 1 : 0;
                }
            }
        }
    }
}

}  // extern "C"
