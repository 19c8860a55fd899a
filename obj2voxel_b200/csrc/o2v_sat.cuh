// Triangle / unit-voxel separating-axis tests around the exact clip (host + device).  None of this is reference
// arithmetic: it only decides which (leaf, voxel) pairs need the exact clip of o2v_exact.cuh, and every decision it
// takes alone is provable from margins that dwarf the rounding of both sides:
//
//   miss      the leaf provably misses the voxel inflated by kPrefilterMargin      -> the exact clip would be empty
//   certain   the leaf provably reaches the voxel shrunk by kCertainMargin          -> the exact clip has >= 1 piece
//   uncertain everything in between                                                 -> run the exact clip
//
// SAT (Akenine-Moeller / Schwarz-Seidel form): box normals (the leaf's voxel AABB), the triangle normal (plane test)
// and the nine edge x axis cross products (three 2-D edge-function tests per projection).  All 13 axes are evaluated
// against the inflated box for `miss` and against the shrunk box for `certain`; one evaluation serves both.
//
// Why `certain` implies a surviving piece in the reference (src/voxelization.cpp:383-424): let q be a point of the leaf
// inside the voxel shrunk by s.  Invariant over the six sequential half-space clips: some piece holds a point within
// k * e of q, e = displacement of a computed intersection point (<= 1.5 ulp of the coordinate + 9e-8 * edge length:
// <= 1.7e-3 at S = 8192 for a grid-sized triangle, <= 3e-4 at S <= 2048).  That piece has a vertex deeper than
// s - k * e > 2^-16 on the kept side, so the case switch (voxelization.cpp:192-234) can only keep it whole or split it,
// and the kept sub-triangles have their vertices within e of the exact cut polygon (same convex combination => a point
// within e).  After six planes a piece is left, so weight = pieces * area > 0.  The plane-distance cull
// (voxelization.cpp:451-458) cannot fire either: `certain` is only granted to leaves whose normal is robust
// (leafFlagsOf).  Budget: s = kCertainMargin - (plane test error <= margin / 2 = 1/128, by kLeafNoPrefilter) - (edge
// function error <= 1e-3) >= 0.022 against 6 e <= 0.0102.  tests/test_sat_classifier.py fuzzes exactly this claim.
#ifndef O2V_SAT_CUH
#define O2V_SAT_CUH

#include "o2v_exact.cuh"

namespace o2v {

// Inflation of the voxel box in the conservative SAT, in voxels.  It must exceed every slack of the exact clip: the
// planarity epsilon (2^-16) plus the rounding of intersection points (two roundings of a coordinate < 8192: <= 2e-3).
constexpr float kPrefilterMargin = 0.015625f;
// Shrink of the voxel box for the `certain` verdict (see the budget above).
constexpr float kCertainMargin = 0.03125f;

struct LeafStage {
    float v[9];
    float t[6];
    float area;
    uint32_t tri;
    uint32_t box;      // tile-local AABB: 4 bits each lo.x lo.y lo.z hi.x hi.y hi.z (hi exclusive, <= 8)
    float plane[4];    // n . p + d for the tile-local voxel min corner p
    float planeLimit;  // (0.5 + margin) * (|nx| + |ny| + |nz|)
    float edge[27];    // 3 projections (xy, yz, zx) x 3 edges x (A, B, C): A*p.a + B*p.b + C >= 0 inside
    uint32_t flags;    // LeafRecord::flags.  51 words: odd stride, so lanes reading the same field of different leaves
                       // hit distinct banks
};

/// Conservative separating-axis coefficients for leaf vs. unit voxels of the tile at `origin` (Schwarz-Seidel edge
/// functions on a box inflated by kPrefilterMargin).  Not exact arithmetic: FMA contraction is welcome here.
O2V_HD void buildPrefilter(LeafStage &s, const float origin[3])
{
    float p[9];
O2V_UNROLL
    for (int k = 0; k < 9; ++k) {
        p[k] = s.v[k] - origin[k % 3];
    }
    const float e0[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
    const float e1[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    const float n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    s.plane[0] = n[0];
    s.plane[1] = n[1];
    s.plane[2] = n[2];
    s.plane[3] = n[0] * (0.5f - p[0]) + n[1] * (0.5f - p[1]) + n[2] * (0.5f - p[2]);
    s.planeLimit = (0.5f + kPrefilterMargin) * (fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]));
    const float grow = 1.0f + kPrefilterMargin;
O2V_UNROLL
    for (int proj = 0; proj < 3; ++proj) {
        const int a = proj, b = (proj + 1) % 3, c = (proj + 2) % 3;  // xy (n.z), yz (n.x), zx (n.y)
        const float sign = n[c] >= 0.0f ? 1.0f : -1.0f;
O2V_UNROLL
        for (int i = 0; i < 3; ++i) {
            const int j = (i + 1) % 3;
            const float ea = p[j * 3 + a] - p[i * 3 + a];
            const float eb = p[j * 3 + b] - p[i * 3 + b];
            const float A = -eb * sign, B = ea * sign;
            float C = -(A * p[i * 3 + a] + B * p[i * 3 + b]);
            C += A > 0.0f ? A * grow : -A * kPrefilterMargin;
            C += B > 0.0f ? B * grow : -B * kPrefilterMargin;
            s.edge[(proj * 3 + i) * 3 + 0] = A;
            s.edge[(proj * 3 + i) * 3 + 1] = B;
            s.edge[(proj * 3 + i) * 3 + 2] = C;
        }
    }
}

/// false only if the triangle provably misses the (inflated) voxel.  NaNs compare false => pass.
O2V_HD bool prefilterPass(const LeafStage &s, float lx, float ly, float lz)
{
    if ((s.flags & kLeafNoPrefilter) != 0) {
        return true;  // sliver: the computed normal is too noisy for the plane test (o2v_exact.cuh, leafFlagsOf)
    }
    const float dist = s.plane[0] * lx + s.plane[1] * ly + s.plane[2] * lz + s.plane[3];
    if (fabsf(dist) > s.planeLimit) {
        return false;
    }
    const float q[3] = {lx, ly, lz};
O2V_UNROLL
    for (int proj = 0; proj < 3; ++proj) {
        const float qa = q[proj], qb = q[(proj + 1) % 3];
O2V_UNROLL
        for (int i = 0; i < 3; ++i) {
            const float *e = s.edge + (proj * 3 + i) * 3;
            if (e[0] * qa + e[1] * qb + e[2] < 0.0f) {
                return false;
            }
        }
    }
    return true;
}

enum SatVerdict : int { kSatMiss = 0, kSatUncertain = 1, kSatCertain = 2 };

/// One edge function of the SAT with both thresholds: inside the inflated box iff a * qa + b * qb + c >= 0, inside the
/// shrunk box iff that value >= k, k = (|a| + |b|) * (kPrefilterMargin + kCertainMargin) (critical corner moved from the
/// inflated to the shrunk box).
struct alignas(16) SatEdge {
    float a, b, c, k;
};

/// Everything the three-way classification of one leaf against the unit voxels of a box needs, relative to the box's
/// min corner (`origin`): 48 floats, 16-byte aligned (128-bit shared-memory loads).
struct alignas(16) PairSat {
    float plane[4];        // n . q + d at the voxel centre, q = voxel min corner - origin
    float planeLimit;      // (0.5 + kPrefilterMargin) * |n|_1
    float planeSure;       // (0.5 - kCertainMargin) * |n|_1
    float lo[3], hi[3];    // origin-relative float AABB of the leaf (box-normal axes of the SAT), shrunk by the
                           // `certain` margin: voxel l is inside on axis a iff lo[a] <= l[a] <= hi[a]
    SatEdge edge[9];
};

/// Shrink of the voxel box for `certain` at sample resolution S: the displacement e of a computed intersection point in
/// the six sequential clips is <= 3e-4 for S <= 2048 (header), so s = 1/64 - 1/128 - 1e-3 = 0.0068 still exceeds 6 e =
/// 0.0018 almost fourfold; above that the header's 1/32 stands.  A third fewer `uncertain` voxels (exact clips) on the
/// resolutions BASELINE.json names.  tests/test_sat_classifier.py fuzzes both.
O2V_HD float certainMarginFor(uint32_t sampleResolution)
{
    return sampleResolution <= 2048u ? 0.015625f : kCertainMargin;
}

/// Builds the constants from a leaf staged with buildPrefilter(s, origin).
O2V_HD void buildPairSat(PairSat &out, const LeafStage &s, const float origin[3], float certainMargin = kCertainMargin)
{
    const float norm1 = fabsf(s.plane[0]) + fabsf(s.plane[1]) + fabsf(s.plane[2]);
O2V_UNROLL
    for (int k = 0; k < 4; ++k) {
        out.plane[k] = s.plane[k];
    }
    out.planeLimit = s.planeLimit;
    out.planeSure = (0.5f - certainMargin) * norm1;
O2V_UNROLL
    for (int a = 0; a < 3; ++a) {
        // l + margin <= max  and  l + 1 - margin >= min
        out.lo[a] = fminf(fminf(s.v[a], s.v[3 + a]), s.v[6 + a]) - origin[a] - 1.0f + certainMargin;
        out.hi[a] = fmaxf(fmaxf(s.v[a], s.v[3 + a]), s.v[6 + a]) - origin[a] - certainMargin;
    }
    const float shift = kPrefilterMargin + certainMargin;
O2V_UNROLL
    for (int k = 0; k < 9; ++k) {
        out.edge[k].a = s.edge[k * 3];
        out.edge[k].b = s.edge[k * 3 + 1];
        out.edge[k].c = s.edge[k * 3 + 2];
        out.edge[k].k = (fabsf(s.edge[k * 3]) + fabsf(s.edge[k * 3 + 1])) * shift;
    }
}

/// The SAT along one row of voxels (fixed y and z, x running): every axis is linear in x, so the y/z part of each of
/// the thirteen axes is evaluated once per row and a voxel costs one multiply-add per axis.
struct RowSat {
    float planeRow;    // n.y * ly + n.z * lz + d:       signed plane value = n.x * lx + planeRow
    float xyBase[3];   // e.b * ly + e.c (xy projection): value = e.a * lx + xyBase
    float zxBase[3];   // e.a * lz + e.c (zx projection): value = e.b * lx + zxBase
    bool miss;         // a yz edge function is negative: no voxel of the row can be hit
    bool sure;         // the yz edge functions and the y / z box normals allow `certain`
};

O2V_HD void buildRowSat(const PairSat &s, float ly, float lz, RowSat &r)
{
    r.planeRow = s.plane[1] * ly + s.plane[2] * lz + s.plane[3];
    r.miss = false;
    r.sure = ly <= s.hi[1] && ly >= s.lo[1] && lz <= s.hi[2] && lz >= s.lo[2];
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        r.xyBase[i] = s.edge[i].b * ly + s.edge[i].c;
        const SatEdge e = s.edge[3 + i];
        const float value = e.a * ly + e.b * lz + e.c;
        r.miss = r.miss || value < 0.0f;
        r.sure = r.sure && value >= e.k;
        r.zxBase[i] = s.edge[6 + i].a * lz + s.edge[6 + i].c;
    }
}

/// true only if every voxel lx in [lxFirst, lxLast] of the row is a `miss`: the row's yz edges say so, or one of the
/// seven axes that vary with x is beyond its threshold at both ends on the same side (each is a monotone function of
/// lx, in floating point too).  A NaN compares false: not a miss.
O2V_HD bool rowSpanMisses(const PairSat &s, const RowSat &r, float lxFirst, float lxLast)
{
    const float d0 = s.plane[0] * lxFirst + r.planeRow, d1 = s.plane[0] * lxLast + r.planeRow;
    bool miss = r.miss || (d0 > s.planeLimit && d1 > s.planeLimit) || (d0 < -s.planeLimit && d1 < -s.planeLimit);
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        const float a = s.edge[i].a, b = s.edge[6 + i].b;
        miss = miss || (a * lxFirst + r.xyBase[i] < 0.0f && a * lxLast + r.xyBase[i] < 0.0f);
        miss = miss || (b * lxFirst + r.zxBase[i] < 0.0f && b * lxLast + r.zxBase[i] < 0.0f);
    }
    return miss;
}

/// The cheap part of rowSpanMisses: the row's yz edges and the plane at both ends of the span.
O2V_HD bool rowPlaneSpanMisses(const PairSat &s, const RowSat &r, float lxFirst, float lxLast)
{
    const float d0 = s.plane[0] * lxFirst + r.planeRow, d1 = s.plane[0] * lxLast + r.planeRow;
    return r.miss || (d0 > s.planeLimit && d1 > s.planeLimit) || (d0 < -s.planeLimit && d1 < -s.planeLimit);
}

/// Three-way verdict for voxel lx of a row that is not RowSat::miss.
O2V_HD int classifyInRow(const PairSat &s, const RowSat &r, float lx)
{
    const float dist = fabsf(s.plane[0] * lx + r.planeRow);
    bool miss = dist > s.planeLimit;
    bool sure = r.sure && dist <= s.planeSure && lx <= s.hi[0] && lx >= s.lo[0];
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        const float xy = s.edge[i].a * lx + r.xyBase[i];
        miss = miss || xy < 0.0f;
        sure = sure && xy >= s.edge[i].k;
        const float zx = s.edge[6 + i].b * lx + r.zxBase[i];
        miss = miss || zx < 0.0f;
        sure = sure && zx >= s.edge[6 + i].k;
    }
    return miss ? kSatMiss : (sure ? kSatCertain : kSatUncertain);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same SAT solved per row instead of evaluated per voxel.  Along a row (fixed ly, lz) each of the seven axes that
// vary with x is a linear function g * lx + base, so "not a miss" and "certain" are each an interval of lx: the
// intersection of one half-line per edge function and one band for the plane.  A row costs the same whatever its length,
// every lane does the same work (no per-lane voxel loops), and the `certain` voxels come out as one run of bits.
// Against the per-voxel form the interval ends carry the rounding of one multiplication by a reciprocal (relative 2^-22):
// a voxel can only change verdict if its value lies within ~1e-6 of a threshold, i.e. the effective margins move by
// ~1e-6 of 1/64 — inside every slack of the header.  tests/test_sat_classifier.py fuzzes this form against the exact clip.

constexpr float kSpanBig = 1.0e30f;       // "no bound from this edge"
constexpr float kSpanInvLimit = 1.0e18f;  // reciprocal of a coefficient that is (nearly) zero: the function is row-constant

/// An x-varying edge function of a row as bounds on lx: with base = coef * (ly or lz) + c and t = base * inv = -base / g,
/// not-a-miss <=> t + dLo <= lx <= t + dHi and certain <=> t + dLoSure <= lx <= t + dHiSure; the side the edge does not
/// bound carries -/+ kSpanBig.  g == 0 is the limit g -> +0: t = -base * kSpanInvLimit puts the bound far below the row
/// when base passes and far above when it fails.
struct alignas(16) SpanEdge {
    float coef, c, inv, dLo;
    float dHi, dLoSure, dHiSure, pad;
};

struct alignas(16) SpanSat {
    float ny, nz, d, pn;                      // plane value = nx * lx + (ny * ly + nz * lz + d); pn = -1 / nx
    float wMiss, wSure, loSure, hiSure;       // half widths planeLimit / |nx|, planeSure / |nx|; x box normal of `certain`
    float ySureLo, ySureHi, zSureLo, zSureHi; // rows whose y / z box normals allow `certain`
    SatEdge yz[3];                            // the yz projection: constant along a row
    SpanEdge xy[3];                           // base = coef * ly + c
    SpanEdge zx[3];                           // base = coef * lz + c
};

/// 1 / x for x in [1e-18, 1e18]; one ulp more or less does not matter to the interval ends (see above).
O2V_HD float spanReciprocal(float x)
{
#if defined(__CUDA_ARCH__)
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}

O2V_HD void makeSpanEdge(SpanEdge &e, float g, float coef, float c, float k)
{
    const float ag = fabsf(g);
    const float invAbs = ag > 1.0f / kSpanInvLimit ? spanReciprocal(ag) : kSpanInvLimit;
    const bool lower = !(g < 0.0f);
    const float ks = k * invAbs;
    e.coef = coef;
    e.c = c;
    e.inv = lower ? -invAbs : invAbs;
    e.dLo = lower ? 0.0f : -kSpanBig;
    e.dHi = lower ? kSpanBig : 0.0f;
    e.dLoSure = lower ? ks : -kSpanBig;
    e.dHiSure = lower ? kSpanBig : -ks;
    e.pad = 0.0f;
}

O2V_HD void buildSpanSat(SpanSat &out, const PairSat &s)
{
    const float anx = fabsf(s.plane[0]);
    const float invAbs = anx > 1.0f / kSpanInvLimit ? spanReciprocal(anx) : kSpanInvLimit;
    out.ny = s.plane[1];
    out.nz = s.plane[2];
    out.d = s.plane[3];
    out.pn = s.plane[0] < 0.0f ? invAbs : -invAbs;
    out.wMiss = s.planeLimit * invAbs;
    out.wSure = s.planeSure * invAbs;
    out.loSure = s.lo[0];
    out.hiSure = s.hi[0];
    out.ySureLo = s.lo[1];
    out.ySureHi = s.hi[1];
    out.zSureLo = s.lo[2];
    out.zSureHi = s.hi[2];
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        out.yz[i] = s.edge[3 + i];
        makeSpanEdge(out.xy[i], s.edge[i].a, s.edge[i].b, s.edge[i].c, s.edge[i].k);          // a * lx + b * ly + c
        makeSpanEdge(out.zx[i], s.edge[6 + i].b, s.edge[6 + i].a, s.edge[6 + i].c, s.edge[6 + i].k);  // a * lz + b * lx + c
    }
}

/// The verdicts of row (ly, lz) of a box whose voxels are lx = 0 .. lastX: voxels in [i0, i1] are not a `miss`, those in
/// [j0, j1] (a sub-range, empty when j0 > j1) are `certain`, the rest of [i0, i1] is `uncertain`.  i0 > i1: the whole
/// row misses.
O2V_HD void classifySpan(const SpanSat &s, float ly, float lz, float lastX, int &i0, int &i1, int &j0, int &j1)
{
    bool miss = false;
    bool sure = ly >= s.ySureLo && ly <= s.ySureHi && lz >= s.zSureLo && lz <= s.zSureHi;
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        const float value = s.yz[i].a * ly + s.yz[i].b * lz + s.yz[i].c;
        miss = miss || value < 0.0f;
        sure = sure && value >= s.yz[i].k;
    }
    const float centre = (s.ny * ly + s.nz * lz + s.d) * s.pn;
    float loM = fmaxf(0.0f, centre - s.wMiss), hiM = fminf(lastX, centre + s.wMiss);
    float loS = fmaxf(s.loSure, centre - s.wSure), hiS = fminf(s.hiSure, centre + s.wSure);
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        const float t = (s.xy[i].coef * ly + s.xy[i].c) * s.xy[i].inv;
        loM = fmaxf(loM, t + s.xy[i].dLo);
        hiM = fminf(hiM, t + s.xy[i].dHi);
        loS = fmaxf(loS, t + s.xy[i].dLoSure);
        hiS = fminf(hiS, t + s.xy[i].dHiSure);
        const float u = (s.zx[i].coef * lz + s.zx[i].c) * s.zx[i].inv;
        loM = fmaxf(loM, u + s.zx[i].dLo);
        hiM = fminf(hiM, u + s.zx[i].dHi);
        loS = fmaxf(loS, u + s.zx[i].dLoSure);
        hiS = fminf(hiS, u + s.zx[i].dHiSure);
    }
    // into [-1, lastX + 1] before the conversions (the bounds may be +-1e30)
    loM = fminf(loM, lastX + 1.0f);
    hiM = fmaxf(hiM, -1.0f);
    loS = fminf(fmaxf(loS, loM), lastX + 1.0f);
    hiS = fmaxf(fminf(hiS, hiM), -1.0f);
#if defined(__CUDA_ARCH__)
    i0 = __float2int_ru(loM);
    i1 = miss ? i0 - 1 : __float2int_rd(hiM);
    j0 = __float2int_ru(loS);
    j1 = sure ? __float2int_rd(hiS) : j0 - 1;
#else
    i0 = (int) ceilf(loM);
    i1 = miss ? i0 - 1 : (int) floorf(hiM);
    j0 = (int) ceilf(loS);
    j1 = sure ? (int) floorf(hiS) : j0 - 1;
#endif
}

/// Three-way verdict for the voxel whose min corner is origin + (lx, ly, lz).  Leaves flagged kLeafNoPrefilter must not
/// be classified (they are `uncertain` throughout: their normal is too noisy for the plane test).  Any NaN makes the
/// comparisons fail towards `uncertain` or `miss` only where `miss` is proven by a comparison that held.
O2V_HD int classifyVoxel(const PairSat &s, float lx, float ly, float lz)
{
    RowSat r;
    buildRowSat(s, ly, lz, r);
    if (r.miss || rowSpanMisses(s, r, lx, lx)) {
        return kSatMiss;
    }
    return classifyInRow(s, r, lx);
}

}  // namespace o2v

#endif  // O2V_SAT_CUH
