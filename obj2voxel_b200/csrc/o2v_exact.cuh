// Exact (reference-identical) binary32 arithmetic of the obj2voxel hot path, written for sm_100a device code and
// compilable for the host (unit tests of the same functions run on CPU, see o2v_hostmath_test.cpp).
//
// Parity contract (SURVEY.md facts 2,4,5; Appendix A): IEEE binary32, round-to-nearest-even, NO fused multiply-add, the
// reference's operation order.  On the device every multiply goes through __fmul_rn and every add/sub that could be
// contracted through __fadd_rn/__fsub_rn, which ptxas never fuses; division and sqrt are the IEEE-rounded intrinsics.
// On the host this header must be compiled with -ffp-contract=off.
//
// Reference locations (relative to /root/reference) are cited per function.
#ifndef O2V_EXACT_CUH
#define O2V_EXACT_CUH

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define O2V_HD __host__ __device__ __forceinline__
#define O2V_UNROLL _Pragma("unroll")
#else
#define O2V_HD inline
#define O2V_UNROLL
#endif

namespace o2v {

#if defined(__CUDA_ARCH__)
O2V_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
O2V_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
O2V_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
O2V_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
O2V_HD float xsqrt(float a) { return __fsqrt_rn(a); }
#else
O2V_HD float xmul(float a, float b) { return a * b; }
O2V_HD float xadd(float a, float b) { return a + b; }
O2V_HD float xsub(float a, float b) { return a - b; }
O2V_HD float xdiv(float a, float b) { return a / b; }
O2V_HD float xsqrt(float a) { return sqrtf(a); }
#endif

// src/constants.hpp:13-15, src/voxelization.cpp:15,337
constexpr uint32_t kSubdivisionVolumeLimit = 512u;
constexpr float kEpsilon = 1.0f / 65536.0f;
constexpr float kSqrtThird = 0.5773502691896257645091487805019574556476017512701268760186023264f;
constexpr float kDiagonalityLimit = 0.5f;
constexpr int kMaxSubdivisionDepth = 24;

enum TriangleType : uint8_t { kMaterialless = 1, kUntextured = 2, kTextured = 3 };  // src/triangle.hpp:21-30
enum ColorStrategy : uint8_t { kMax = 0, kBlend = 1 };                               // include/obj2voxel.h:43-46

/// A triangle in voxel space.  v[k*3+axis]; t[k*2+c] only meaningful when UV is true.
template <bool UV>
struct Tri {
    float v[9];
    float t[UV ? 6 : 1];
};

// voxelio vec.hpp:378-385: dot is a left fold that starts at 0
O2V_HD float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    float r = xadd(0.0f, xmul(ax, bx));
    r = xadd(r, xmul(ay, by));
    r = xadd(r, xmul(az, bz));
    return r;
}

// src/util.hpp:143-146: (1 - t) * a + t * b
O2V_HD float mix1(float a, float b, float t)
{
    return xadd(xmul(xsub(1.0f, t), a), xmul(t, b));
}

/// Unnormalised normal cross(v1 - v0, v2 - v0): src/triangle.hpp:59-62, vec.hpp:389-399.
O2V_HD void triNormal(const float *v, float n[3])
{
    const float ax = xsub(v[3], v[0]), ay = xsub(v[4], v[1]), az = xsub(v[5], v[2]);
    const float bx = xsub(v[6], v[0]), by = xsub(v[7], v[1]), bz = xsub(v[8], v[2]);
    n[0] = xsub(xmul(ay, bz), xmul(az, by));
    n[1] = xsub(xmul(az, bx), xmul(ax, bz));
    n[2] = xsub(xmul(ax, by), xmul(ay, bx));
}

/// src/triangle.hpp:103-106: length(normal()) / 2
O2V_HD float triArea(const float *v)
{
    float n[3];
    triNormal(v, n);
    return xdiv(xsqrt(dot3(n[0], n[1], n[2], n[0], n[1], n[2])), 2.0f);
}

/// src/voxelization.cpp:335-347.  NaN (zero-area) compares false => subdivision path.
O2V_HD bool triRoughlyAxisAligned(const float *v)
{
    float n[3];
    triNormal(v, n);
    const float ax = fabsf(n[0]), ay = fabsf(n[1]), az = fabsf(n[2]);
    const float len = xsqrt(dot3(ax, ay, az, ax, ay, az));
    const float ux = xdiv(ax, len), uy = xdiv(ay, len), uz = xdiv(az, len);
    const float diagonality = dot3(ux, uy, uz, kSqrtThird, kSqrtThird, kSqrtThird);
    const float diagonality01 = xdiv(xsub(diagonality, kSqrtThird), xsub(1.0f, kSqrtThird));
    return diagonality01 < kDiagonalityLimit;
}

/// The plane-distance cull of voxelizeSubTriangle (src/voxelization.cpp:435-458) with the reference's arithmetic:
/// planeNormal = normal() / length(normal()) (util.hpp:127-139), center = Vec3(pos) + 0.5,
/// |dot(planeNormal, center - v0)| > 2 skips the voxel.  NaN normals (zero area) never cull.
O2V_HD bool planeDistanceCulled(const float *v, uint32_t px, uint32_t py, uint32_t pz)
{
    float n[3];
    triNormal(v, n);
    const float len = xsqrt(dot3(n[0], n[1], n[2], n[0], n[1], n[2]));
    const float ux = xdiv(n[0], len), uy = xdiv(n[1], len), uz = xdiv(n[2], len);
    const float rx = xsub(xadd(static_cast<float>(px), 0.5f), v[0]);
    const float ry = xsub(xadd(static_cast<float>(py), 0.5f), v[1]);
    const float rz = xsub(xadd(static_cast<float>(pz), 0.5f), v[2]);
    return fabsf(dot3(ux, uy, uz, rx, ry, rz)) > 2.0f;
}

/// true when the cull above provably cannot fire for any voxel the leaf's exact clip would contribute to, so it may be
/// skipped: a contributing voxel's centre lies within sqrt(3)/2 + 0.02 of the leaf's plane, the limit is 2, and the
/// computed unit normal is off by at most |dn| / |n| with |dn| <= 9 * 2^-24 * |e01| * |e02| (three roundings per
/// component of the cross product), which is multiplied by |centre - v0| <= max(|e01|, |e02|) + 2.  Needed slack: 1.1;
/// the test below grants itself another factor 4:  |n| > 4e-6 * |e01| * |e02| * (max(|e01|, |e02|) + 2), squared
/// (with (m + 2)^2 <= 2 m^2 + 8).  Slivers — and NaN/inf inputs, which compare false — fail it and take the exact cull.
/// Not part of the reference arithmetic (any float evaluation will do; the constants carry the slack).
///
/// The same quantity decides whether the conservative SAT prefilter (o2v_device.cuh) may be trusted: its plane test
/// |n . (c - p0)| <= (0.5 + margin) |n|_1 stays conservative as long as |dn| (D + 0.87) <= margin |n|, margin = 1/64,
/// i.e. |n| >= 3.5e-5 |e01| |e02| (D + 1); `slack` selects the threshold (kCullSlack / kPrefilterSlack, squared).
O2V_HD bool planeNormalIsRobust(const float *v, float slackSquared)
{
    const float ax = v[3] - v[0], ay = v[4] - v[1], az = v[5] - v[2];
    const float bx = v[6] - v[0], by = v[7] - v[1], bz = v[8] - v[2];
    const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const float n2 = nx * nx + ny * ny + nz * nz;
    const float a2 = ax * ax + ay * ay + az * az, b2 = bx * bx + by * by + bz * bz;
    const float m2 = a2 < b2 ? b2 : a2;
    return n2 > slackSquared * a2 * b2 * (2.0f * m2 + 8.0f);
}

constexpr float kCullSlackSquared = 1.6e-11f;       // (4e-6)^2
constexpr float kPrefilterSlackSquared = 1.6e-9f;   // (4e-5)^2
constexpr uint32_t kLeafNeedsCull = 1u;    // LeafRecord::flags: the exact distance cull must be evaluated (slivers)
constexpr uint32_t kLeafNoPrefilter = 2u;  // the SAT prefilter's plane test is not provably conservative: skip it

O2V_HD uint32_t leafFlagsOf(const float *v)
{
    return (planeNormalIsRobust(v, kCullSlackSquared) ? 0u : kLeafNeedsCull) |
           (planeNormalIsRobust(v, kPrefilterSlackSquared) ? 0u : kLeafNoPrefilter);
}

// util.hpp:80-98 (std::min/std::max nesting)
O2V_HD float min3(float a, float b, float c)
{
    const float bc = c < b ? c : b;
    return bc < a ? bc : a;
}
O2V_HD float max3(float a, float b, float c)
{
    const float bc = b < c ? c : b;
    return a < bc ? bc : a;
}

/// float -> u32 of src/triangle.hpp:91-100; negative / huge input is UB in the reference and is clamped here.
O2V_HD uint32_t toU32(float x)
{
    if (!(x > 0.0f)) {
        return 0u;
    }
    if (x >= 4294967040.0f) {
        return 4294967040u;
    }
    return static_cast<uint32_t>(x);
}

/// Inclusive min / exclusive max voxel bounds: src/triangle.hpp:91-100.
O2V_HD void triVoxelBounds(const float *v, uint32_t vmin[3], uint32_t vmax[3])
{
    for (int i = 0; i < 3; ++i) {
        vmin[i] = toU32(floorf(min3(v[i], v[3 + i], v[6 + i])));
        vmax[i] = toU32(floorf(max3(v[i], v[3 + i], v[6 + i]))) + 1u;
    }
}

/// v' = M v + t with the row-wise dot of src/util.hpp:262-268.  m = 3x3 row-major followed by the translation.
O2V_HD void affineApply(const float *m, const float *in, float *out)
{
    const float x = dot3(m[0], m[1], m[2], in[0], in[1], in[2]);
    const float y = dot3(m[3], m[4], m[5], in[0], in[1], in[2]);
    const float z = dot3(m[6], m[7], m[8], in[0], in[1], in[2]);
    out[0] = xadd(x, m[9]);
    out[1] = xadd(y, m[10]);
    out[2] = xadd(z, m[11]);
}

// ---------------------------------------------------------------------------------------------------------------------
// subdivision: src/voxelization.cpp:349-379, src/triangle.hpp:134-143

/// Child `c` of the 4-way split: 0 = centre (g0,g1,g2), 1 = (v0,g0,g2), 2 = (v1,g1,g0), 3 = (v2,g2,g1).
template <bool UV>
O2V_HD void triChild(const Tri<UV> &p, int c, Tri<UV> &out)
{
    float g[9];
    float x[6];
    for (int e = 0; e < 3; ++e) {
        const int a = e, b = (e + 1) % 3;
        for (int i = 0; i < 3; ++i) {
            g[e * 3 + i] = mix1(p.v[a * 3 + i], p.v[b * 3 + i], 0.5f);
        }
        if (UV) {
            for (int i = 0; i < 2; ++i) {
                x[e * 2 + i] = mix1(p.t[a * 2 + i], p.t[b * 2 + i], 0.5f);
            }
        }
    }
    // corner k = (v_k, g_k, g_{k+2}); centre = (g0, g1, g2)
    const int k = c - 1;
    const int ga = c == 0 ? 0 : k, gb = c == 0 ? 1 : (k + 2) % 3;
    for (int i = 0; i < 3; ++i) {
        out.v[i] = c == 0 ? g[i] : p.v[k * 3 + i];
        out.v[3 + i] = c == 0 ? g[3 + i] : g[ga * 3 + i];
        out.v[6 + i] = c == 0 ? g[6 + i] : g[gb * 3 + i];
    }
    if (UV) {
        for (int i = 0; i < 2; ++i) {
            out.t[i] = c == 0 ? x[i] : p.t[k * 2 + i];
            out.t[2 + i] = c == 0 ? x[2 + i] : x[ga * 2 + i];
            out.t[4 + i] = c == 0 ? x[4 + i] : x[gb * 2 + i];
        }
    }
    (void) ga;
    (void) gb;
}

/// true when the leaf test of forEachSubdividedTriangle pops the triangle: u32 AABB volume < 512 (wraps like the
/// reference).
O2V_HD bool triIsLeaf(const float *v)
{
    uint32_t lo[3], hi[3];
    triVoxelBounds(v, lo, hi);
    const uint32_t volume = (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    return volume < kSubdivisionVolumeLimit;
}

/// Visits the leaves of one input triangle in the reference's emission order: an aligned triangle is its own single
/// leaf; otherwise depth-first with children visited 3, 2, 1, centre (LIFO stack, centre replaces the top).
/// fn(const Tri<UV>&) is called per leaf.  Returns false if kMaxSubdivisionDepth was exceeded (leaf emitted as is).
template <bool UV, typename LeafFn>
O2V_HD bool forEachLeafBelow(const Tri<UV> &start, int startLevel, LeafFn &&fn)
{
    // `start` sits `startLevel` splits below its input triangle (0: it is the triangle): the depth limit counts from there
    Tri<UV> parents[kMaxSubdivisionDepth];
    uint8_t next[kMaxSubdivisionDepth];
    int level = 0;
    bool ok = true;
    Tri<UV> node = start;
    for (;;) {
        while (!triIsLeaf(node.v)) {
            if (level + startLevel >= kMaxSubdivisionDepth) {
                ok = false;
                break;
            }
            parents[level] = node;
            next[level] = 0;
            Tri<UV> child;
            triChild<UV>(node, 3, child);
            node = child;
            ++level;
        }
        fn(node);
        bool advanced = false;
        while (level > 0) {
            const int n = ++next[level - 1];
            if (n < 4) {
                triChild<UV>(parents[level - 1], 3 - n, node);
                advanced = true;
                break;
            }
            --level;
        }
        if (!advanced) {
            break;
        }
    }
    return ok;
}

template <bool UV, typename LeafFn>
O2V_HD bool forEachLeaf(const Tri<UV> &root, LeafFn &&fn)
{
    if (triRoughlyAxisAligned(root.v)) {
        fn(root);
        return true;
    }
    return forEachLeafBelow<UV>(root, 0, fn);
}

/// The leaves of one of the 4^depth subtrees `depth` splits below a (non-aligned) triangle, subtrees numbered in the
/// order forEachLeaf reaches them (two bits per level, most significant first; position p = child 3 - p): walking the
/// subtrees 0, 1, 2, ... one after the other visits the leaves in forEachLeaf's order.  A node that is a leaf above the
/// split depth belongs to the first subtree below it.
template <bool UV, typename LeafFn>
O2V_HD bool forEachLeafOfSubtree(const Tri<UV> &root, int depth, uint32_t subtree, LeafFn &&fn)
{
    Tri<UV> node = root;
    for (int level = 0; level < depth; ++level) {
        const int below = 2 * (depth - 1 - level);  // bits of the levels under this one
        if (triIsLeaf(node.v)) {
            if ((subtree & ((1u << (below + 2)) - 1u)) == 0) {
                fn(node);
            }
            return true;
        }
        Tri<UV> child;
        triChild<UV>(node, 3 - (int) ((subtree >> below) & 3u), child);
        node = child;
    }
    return forEachLeafBelow<UV>(node, depth, fn);
}

// ---------------------------------------------------------------------------------------------------------------------
// triangle splitting: src/voxelization.cpp:17-31,110-331

O2V_HD bool isZero(float x) { return fabsf(x) < kEpsilon; }

O2V_HD float axisOf(const float *p, int axis) { return axis == 0 ? p[0] : (axis == 1 ? p[1] : p[2]); }

/// src/voxelization.cpp:27-31 with dir = to - org
O2V_HD float intersectAxisPlane(const float *org, const float *to, int axis, float plane)
{
    const float d = -xsub(axisOf(to, axis), axisOf(org, axis));
    return isZero(d) ? 0.0f : xdiv(xsub(axisOf(org, axis), plane), d);
}

template <bool UV>
O2V_HD void setVertex(Tri<UV> &dst, int k, const float *p, const float *t)
{
    dst.v[k * 3] = p[0];
    dst.v[k * 3 + 1] = p[1];
    dst.v[k * 3 + 2] = p[2];
    if (UV) {
        dst.t[k * 2] = t[0];
        dst.t[k * 2 + 1] = t[1];
    }
}

/// Result of clipping one leaf against one voxel: src/voxelization.cpp:383-424.
struct ClipResult {
    int pieces;    // number of surviving pieces (0 => no contribution)
    float weight;  // pieces-fold repeated sum of the whole-triangle area
    float u, v;    // sequential weighted mean of the piece UV centres (only when UV)
};

/// What one half-space does to a triangle (the switch of src/voxelization.cpp:192-234 without the geometry).
enum ClipAction : int { kClipKeep = 0, kClipDrop = 1, kClipSplitRegular = 2, kClipSplitOnePlanar = 3 };

/// The reference's 16-way case switch (src/voxelization.cpp:192-234) as a function of the three "planar" flags P and the
/// three "lo" flags L (bit k = vertex k).  Packed result: bits 1:0 = kind (0 unsplit, 1 regular split, 2 one-planar
/// split), bit 2 = the unsplit triangle is lo / the piece that starts at the pivot is lo, bits 4:3 = pivot vertex
/// (isolated vertex of the regular case :289-293, planar vertex of the one-planar case :245-246).
O2V_HD uint32_t clipCaseOf(uint32_t P, uint32_t L)
{
    const bool p0 = (P & 1u) != 0, p1 = (P & 2u) != 0, p2 = (P & 4u) != 0;
    const bool l0 = (L & 1u) != 0, l1 = (L & 2u) != 0, l2 = (L & 4u) != 0;
    const int loSum = int(l0) + int(l1) + int(l2);
    const int planarSum = int(p0) + int(p1) + int(p2);
    uint32_t kind = 0, pivot = 0;
    bool flag;
    if (loSum == 0) {
        flag = false;
    }
    else if (loSum == 3) {
        flag = true;
    }
    else if (planarSum == 3) {
        flag = false;  // IS_LO_BIASED == false
    }
    else if (planarSum == 2) {
        flag = !p0 ? l0 : (!p1 ? l1 : l2);  // loVertices[firstNonplanar()]
    }
    else if (planarSum == 1) {
        const uint32_t p = p0 ? 0u : (p1 ? 1u : 2u);
        const bool la = p == 0 ? l1 : (p == 1 ? l2 : l0);  // vertex (p + 1) % 3
        const bool lb = p == 0 ? l2 : (p == 1 ? l0 : l1);  // vertex (p + 2) % 3
        flag = la;
        if (la != lb) {
            kind = 2;
            pivot = p;
        }
    }
    else {
        const bool isoLo = loSum == 1;
        kind = 1;
        pivot = isoLo ? (l0 ? 0u : (l1 ? 1u : 2u)) : (!l0 ? 0u : (!l1 ? 1u : 2u));
        flag = isoLo;
    }
    return kind | (flag ? 4u : 0u) | (pivot << 3);
}

/// Planar / lo flags of the three vertices against one axis plane (SplittingValues, src/voxelization.cpp:110-153),
/// packed as (P << 3) | L for clipCaseOf / the case table.
O2V_HD uint32_t planeFlags(float c0, float c1, float c2, float planePos)
{
    const uint32_t P = (isZero(xsub(c0, planePos)) ? 1u : 0u) | (isZero(xsub(c1, planePos)) ? 2u : 0u) |
                       (isZero(xsub(c2, planePos)) ? 4u : 0u);
    const uint32_t L = (c0 < planePos ? 1u : 0u) | (c1 < planePos ? 2u : 0u) | (c2 < planePos ? 4u : 0u);
    return (P << 3) | L;
}

/// Classifies triangle `v` against plane `planePos` on `axis`.  For the split actions `pivot` is the isolated vertex
/// (regular case) or the planar vertex (one-planar case) and `sideLo` tells whether the piece at the pivot
/// (regular: the isolated corner; one-planar: (p, a, X)) lies on the lo side.
O2V_HD ClipAction classifyAgainstPlane(const float *v, int axis, float planePos, bool keepHi, int &pivot, bool &sideLo)
{
    const uint32_t flags = planeFlags(axisOf(v, axis), axisOf(v + 3, axis), axisOf(v + 6, axis), planePos);
    const uint32_t code = clipCaseOf(flags >> 3, flags & 7u);
    const uint32_t kind = code & 3u;
    const bool flag = (code & 4u) != 0;
    if (kind == 0) {
        return flag != keepHi ? kClipKeep : kClipDrop;
    }
    pivot = int(code >> 3);
    sideLo = flag;
    return kind == 1 ? kClipSplitRegular : kClipSplitOnePlanar;
}

/// Rotates the triangle so that vertex `pivot` comes first (keeps the cyclic order, i.e. (pivot+1)%3 and (pivot+2)%3
/// become vertices 1 and 2 exactly like otherIndices / nonPlanarIndices in the reference).
template <bool UV>
O2V_HD void rotateToPivot(Tri<UV> &t, int pivot)
{
    if (pivot == 0) {
        return;
    }
    Tri<UV> r;
O2V_UNROLL
    for (int i = 0; i < 3; ++i) {
        r.v[i] = pivot == 1 ? t.v[3 + i] : t.v[6 + i];
        r.v[3 + i] = pivot == 1 ? t.v[6 + i] : t.v[i];
        r.v[6 + i] = pivot == 1 ? t.v[i] : t.v[3 + i];
    }
    if (UV) {
O2V_UNROLL
        for (int i = 0; i < 2; ++i) {
            r.t[i] = pivot == 1 ? t.t[2 + i] : t.t[4 + i];
            r.t[2 + i] = pivot == 1 ? t.t[4 + i] : t.t[i];
            r.t[4 + i] = pivot == 1 ? t.t[i] : t.t[2 + i];
        }
    }
    t = r;
}

/// Six sequential half-space clips of `leaf` against voxel (px,py,pz), visiting surviving pieces in the reference's list
/// order (a depth-first walk over the clip tree yields the same left-to-right leaf order as its breadth-first ping-pong
/// buffers), then the weighted fold of voxelization.cpp:414-420 / util.hpp:160-165.
///
/// SIMT shape: every lane runs the same two-phase loop — a cheap classify/advance loop (no geometry) until its current
/// piece needs a real split, then ONE shared split step — so the lanes of a warp execute the expensive split code together
/// instead of serialising on the reference's 16-way case switch.
template <bool UV>
O2V_HD ClipResult clipLeafInVoxel(const Tri<UV> &leaf, uint32_t px, uint32_t py, uint32_t pz, float wholeArea)
{
    ClipResult r;
    r.pieces = 0;
    r.weight = 0.0f;
    r.u = 0.0f;
    r.v = 0.0f;

    Tri<UV> pending[6];
    uint8_t pendingPlane[6];
    int sp = 0;

    Tri<UV> cur = leaf;
    int plane = 0;
    for (;;) {
        // ---- phase A: advance through planes that keep or drop the piece whole; pop finished / dead pieces ----
        ClipAction action = kClipKeep;
        int pivot = 0;
        bool sideLo = false;
        bool finished = false;
        for (;;) {
            if (plane == 6) {
                // a surviving piece: result = mix(result, {area, textureCenter}) (util.hpp:160-165, triangle.hpp:127-131)
                const float weightSum = xadd(r.weight, wholeArea);
                if (UV) {
                    const float cu = xdiv(xadd(xadd(cur.t[0], cur.t[2]), cur.t[4]), 3.0f);
                    const float cv = xdiv(xadd(xadd(cur.t[1], cur.t[3]), cur.t[5]), 3.0f);
                    r.u = xdiv(xadd(xmul(r.weight, r.u), xmul(wholeArea, cu)), weightSum);
                    r.v = xdiv(xadd(xmul(r.weight, r.v), xmul(wholeArea, cv)), weightSum);
                }
                r.weight = weightSum;
                ++r.pieces;
                action = kClipDrop;  // done with this piece: take the next pending one
            }
            else {
                const int axis = plane < 3 ? plane : plane - 3;
                const uint32_t base = axis == 0 ? px : (axis == 1 ? py : pz);
                const float planePos = static_cast<float>(base + (plane < 3 ? 0u : 1u));
                action = classifyAgainstPlane(cur.v, axis, planePos, plane < 3, pivot, sideLo);
            }
            if (action == kClipKeep) {
                ++plane;
                continue;
            }
            if (action == kClipDrop) {
                if (sp == 0) {
                    finished = true;
                    break;
                }
                --sp;
                cur = pending[sp];
                plane = pendingPlane[sp];
                continue;
            }
            break;  // a split is needed
        }
        if (finished) {
            break;
        }

        // ---- phase B: one split (the expensive geometry), shared by all lanes that reached it ----
        const int axis = plane < 3 ? plane : plane - 3;
        const uint32_t base = axis == 0 ? px : (axis == 1 ? py : pz);
        const float planePos = static_cast<float>(base + (plane < 3 ? 0u : 1u));
        const bool keepHi = plane < 3;
        rotateToPivot<UV>(cur, pivot);
        if (action == kClipSplitRegular) {
            // splitTriangle_regularCase, voxelization.cpp:279-331: pivot = isolated vertex, X0/X1 on its two edges
            const float s0 = intersectAxisPlane(cur.v, cur.v + 3, axis, planePos);
            const float s1 = intersectAxisPlane(cur.v, cur.v + 6, axis, planePos);
            float g0[3], g1[3], x0[2] = {0.0f, 0.0f}, x1[2] = {0.0f, 0.0f};
O2V_UNROLL
            for (int i = 0; i < 3; ++i) {
                g0[i] = mix1(cur.v[i], cur.v[3 + i], s0);
                g1[i] = mix1(cur.v[i], cur.v[6 + i], s1);
            }
            if (UV) {
O2V_UNROLL
                for (int i = 0; i < 2; ++i) {
                    x0[i] = mix1(cur.t[i], cur.t[2 + i], s0);
                    x1[i] = mix1(cur.t[i], cur.t[4 + i], s1);
                }
            }
            if (sideLo != keepHi) {
                // the isolated corner (iso, X0, X1) is kept
                setVertex<UV>(cur, 1, g0, x0);
                setVertex<UV>(cur, 2, g1, x1);
            }
            else {
                // the quad is kept: (X0, a, b) first, (X0, X1, b) pending for the next plane
                Tri<UV> second;
                setVertex<UV>(second, 0, g0, x0);
                setVertex<UV>(second, 1, g1, x1);
                setVertex<UV>(second, 2, cur.v + 6, cur.t + (UV ? 4 : 0));
                pending[sp] = second;
                pendingPlane[sp] = static_cast<uint8_t>(plane + 1);
                ++sp;
                setVertex<UV>(cur, 0, g0, x0);
            }
        }
        else {
            // splitTriangle_onePlanarCase, voxelization.cpp:240-277: pivot = planar vertex, X on the opposite edge a -> b;
            // (p, a, X) lies on a's side, (p, X, b) on the other
            const float s = intersectAxisPlane(cur.v + 3, cur.v + 6, axis, planePos);
            float geo[3], tex[2] = {0.0f, 0.0f};
O2V_UNROLL
            for (int i = 0; i < 3; ++i) {
                geo[i] = mix1(cur.v[3 + i], cur.v[6 + i], s);
            }
            if (UV) {
O2V_UNROLL
                for (int i = 0; i < 2; ++i) {
                    tex[i] = mix1(cur.t[2 + i], cur.t[4 + i], s);
                }
            }
            if (sideLo != keepHi) {
                setVertex<UV>(cur, 2, geo, tex);  // keep (p, a, X)
            }
            else {
                setVertex<UV>(cur, 1, geo, tex);  // keep (p, X, b)
            }
        }
        ++plane;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// weighted combination: src/util.hpp:150-175 (always called as combine(new, existing))

struct WeightedUv {
    float w, u, v;
};

struct WeightedColor {
    float w, r, g, b;
};

/// insertWeighted<BLEND> on the per-triangle uv buffer: src/voxelization.cpp:56-63
O2V_HD void blendUvInto(WeightedUv &existing, float w, float u, float v)
{
    const float weightSum = xadd(w, existing.w);
    existing.u = xdiv(xadd(xmul(w, u), xmul(existing.w, existing.u)), weightSum);
    existing.v = xdiv(xadd(xmul(w, v), xmul(existing.w, existing.v)), weightSum);
    existing.w = weightSum;
}

O2V_HD void combineColorInto(WeightedColor &existing, float w, float r, float g, float b, bool blend)
{
    if (blend) {
        const float weightSum = xadd(w, existing.w);
        existing.r = xdiv(xadd(xmul(w, r), xmul(existing.w, existing.r)), weightSum);
        existing.g = xdiv(xadd(xmul(w, g), xmul(existing.w, existing.g)), weightSum);
        existing.b = xdiv(xadd(xmul(w, b), xmul(existing.w, existing.b)), weightSum);
        existing.w = weightSum;
    }
    else if (w > existing.w) {  // ties keep the earlier (lower triangle index) value, util.hpp:169-172
        existing.w = w;
        existing.r = r;
        existing.g = g;
        existing.b = b;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// colour lookup and quantisation

/// voxelio image.hpp:87-95 (REPEAT; exact integers map to 1.0)
O2V_HD float wrapRepeat(float x)
{
    float integral;
    float fraction = modff(x, &integral);
    fraction = xadd(fraction, fraction < 0.0f ? 1.0f : 0.0f);
    fraction = xadd(fraction, fraction == 0.0f ? 1.0f : 0.0f);
    return fraction;
}

/// voxelio color.hpp:152-156
O2V_HD float clamp01(float x)
{
    const float lo = x < 0.0f ? 0.0f : x;
    return 1.0f < lo ? 1.0f : lo;
}

struct TextureView {
    const uint8_t *pixels;
    uint32_t width, height;
    uint32_t channels;  // 3 = RGB24, 4 = ARGB32 as decoded by voxelio/src/image.cpp:95-98 (load_pixels), 5 = RGBA32 (PNG)
    uint32_t wrap;      // 0 clamp, 1 repeat (include/obj2voxel.h:48-51)
};

/// src/triangle.hpp:181-194 (TEXTURED branch: lookup at (u, 1 - v)), image.hpp:159-194, color.hpp:38-52
O2V_HD void textureLookup(const TextureView &tex, float u, float v, float rgb[3])
{
    const float fv = xsub(1.0f, v);
    const float wu = tex.wrap == 0u ? clamp01(u) : wrapRepeat(u);
    const float wv = tex.wrap == 0u ? clamp01(fv) : wrapRepeat(fv);
    const size_t x = static_cast<size_t>(xmul(wu, static_cast<float>(tex.width - 1u)));
    const size_t y = static_cast<size_t>(xmul(wv, static_cast<float>(tex.height - 1u)));
    const uint32_t bytesPerPixel = tex.channels == 3u ? 3u : 4u;
    const uint8_t *in = tex.pixels + (y * tex.width + x) * bytesPerPixel;
    uint8_t r, g, b;
    if (tex.channels == 4u) {  // decodeArgb32 as written: Color32{in[3], in[0], in[1], in[2]} read as (r, g, b, a)
        r = in[3];
        g = in[0];
        b = in[1];
    }
    else {  // decodeRgb24 / decodeRgba32
        r = in[0];
        g = in[1];
        b = in[2];
    }
    rgb[0] = xdiv(static_cast<float>(r), 255.0f);
    rgb[1] = xdiv(static_cast<float>(g), 255.0f);
    rgb[2] = xdiv(static_cast<float>(b), 255.0f);
}

/// src/obj2voxel.cpp:283-296 + voxelio color.hpp:165-173,95: truncating float -> u8, alpha 0xFF
O2V_HD uint32_t quantizeArgb(float r, float g, float b)
{
    const uint32_t qr = static_cast<uint8_t>(xmul(clamp01(r), 255.0f));
    const uint32_t qg = static_cast<uint8_t>(xmul(clamp01(g), 255.0f));
    const uint32_t qb = static_cast<uint8_t>(xmul(clamp01(b), 255.0f));
    return 0xFF000000u | (qr << 16) | (qg << 8) | qb;
}

}  // namespace o2v

#endif  // O2V_EXACT_CUH
