/* TEST INFRASTRUCTURE ONLY — see o2v_oracle.h.  CPU restatement (plain C99, binary32, no FMA contraction: build with
 * -ffp-contract=off) of the obj2voxel hot path.  Written from the semantics of the reference, not from its text; each
 * function cites the reference file:line (relative to /root/reference) it follows.
 *
 * Parity status: PINNED against oracle/_ref (the unmodified reference) and tests/golden/ — tests/test_oracle.py.
 */
#include "o2v_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------------------------ */
/* constants: src/constants.hpp:10-15, src/voxelization.cpp:15,337,435 */

#define CHUNK_SIZE 64u
#define SUBDIVISION_VOLUME_LIMIT 512u
static const float EPSILON = 1.0f / 65536.0f;
static const float SQRT_THIRD = 0.5773502691896257645091487805019574556476017512701268760186023264f;
static const float DIAGONALITY_LIMIT = 0.5f;
static const float DISTANCE_LIMIT = 2.0f;

typedef struct {
    float v[3][3];
    float t[3][2];
} tri_t;

/* ------------------------------------------------------------------------------------------------------------------ */
/* vector arithmetic: voxelio vec.hpp:378-399 (dot is a left fold starting at 0), util.hpp:127-146 */

static float dot3(const float a[3], const float b[3])
{
    float r = 0;
    r += a[0] * b[0];
    r += a[1] * b[1];
    r += a[2] * b[2];
    return r;
}

static void cross3(const float a[3], const float b[3], float out[3])
{
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

static void sub3(const float a[3], const float b[3], float out[3])
{
    for (int i = 0; i < 3; ++i) {
        out[i] = a[i] - b[i];
    }
}

/* util.hpp:143-146: (1 - t) * a + t * b */
static float mixf(float a, float b, float t)
{
    return (1 - t) * a + t * b;
}

/* triangle.hpp:59-62 */
static void tri_normal(const tri_t *t, float out[3])
{
    float e01[3], e02[3];
    sub3(t->v[1], t->v[0], e01);
    sub3(t->v[2], t->v[0], e02);
    cross3(e01, e02, out);
}

/* triangle.hpp:103-106: length(normal()) / 2 */
static float tri_area(const tri_t *t)
{
    float n[3];
    tri_normal(t, n);
    return sqrtf(dot3(n, n)) / 2;
}

static float min3f(float a, float b, float c)
{
    /* util.hpp:80-89: std::min(a, std::min(b, c)) */
    float bc = c < b ? c : b;
    return bc < a ? bc : a;
}

static float max3f(float a, float b, float c)
{
    float bc = b < c ? c : b;
    return a < bc ? bc : a;
}

/* float -> u32 as the reference's cast<u32>() (triangle.hpp:91-100); out-of-range input is UB there,
 * clamped here. */
static uint32_t to_u32(float x)
{
    if (!(x > 0)) {
        return 0;
    }
    if (x >= 4294967040.0f) {
        return 4294967040u;
    }
    return (uint32_t) x;
}

/* triangle.hpp:91-100 */
static void tri_voxel_bounds(const tri_t *t, uint32_t vmin[3], uint32_t vmax[3])
{
    for (int i = 0; i < 3; ++i) {
        vmin[i] = to_u32(floorf(min3f(t->v[0][i], t->v[1][i], t->v[2][i])));
        vmax[i] = to_u32(floorf(max3f(t->v[0][i], t->v[1][i], t->v[2][i]))) + 1u;
    }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Morton keys: voxelio ileave.hpp:243-246 — x occupies the most significant bit of each triple */

static uint64_t spread3(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

static uint32_t compact3(uint64_t x)
{
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t) x;
}

uint64_t o2v_oracle_ileave3(uint32_t x, uint32_t y, uint32_t z)
{
    return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}

void o2v_oracle_dileave3(uint64_t n, uint32_t out[3])
{
    out[0] = compact3(n >> 2);
    out[1] = compact3(n >> 1);
    out[2] = compact3(n);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* mesh -> voxel transform: src/obj2voxel.cpp:370-402, src/util.hpp:212-281 */

typedef struct {
    float m[3][3];
    float t[3];
} affine_t;

static affine_t affine_scale(float scale, float tx, float ty, float tz)
{
    affine_t a;
    memset(&a, 0, sizeof a);
    a.m[0][0] = a.m[1][1] = a.m[2][2] = scale;
    a.t[0] = tx;
    a.t[1] = ty;
    a.t[2] = tz;
    return a;
}

/* util.hpp:270-281: matrix[i][j] = dot(lhs.row(i), rhs.col(j)); translation = lhs.matrix * rhs.translation + lhs.translation */
static affine_t affine_compose(const affine_t *lhs, const affine_t *rhs)
{
    affine_t r;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            const float col[3] = {rhs->m[0][j], rhs->m[1][j], rhs->m[2][j]};
            r.m[i][j] = dot3(lhs->m[i], col);
        }
        r.t[i] = dot3(lhs->m[i], rhs->t);
    }
    for (int i = 0; i < 3; ++i) {
        r.t[i] += lhs->t[i];
    }
    return r;
}

/* util.hpp:262-268 */
static void affine_apply(const affine_t *a, const float v[3], float out[3])
{
    const float x = dot3(a->m[0], v);
    const float y = dot3(a->m[1], v);
    const float z = dot3(a->m[2], v);
    out[0] = x + a->t[0];
    out[1] = y + a->t[1];
    out[2] = z + a->t[2];
}

static affine_t compute_mesh_transform(const float mesh_min[3], const float mesh_max[3], uint32_t sample_resolution,
                                       const int unit[9])
{
    const float ANTI_BLEED = 0.5f;
    float size[3];
    sub3(mesh_max, mesh_min, size);
    const float max_axis = max3f(size[0], size[1], size[2]);
    const float sample_scale = (float) sample_resolution - ANTI_BLEED;

    affine_t result = affine_scale(1, -mesh_min[0], -mesh_min[1], -mesh_min[2]);
    affine_t step = affine_scale(2.0f / max_axis, -1.0f, -1.0f, -1.0f);
    result = affine_compose(&step, &result);
    affine_t u;
    for (int i = 0; i < 9; ++i) {
        u.m[i / 3][i % 3] = (float) unit[i];
    }
    u.t[0] = u.t[1] = u.t[2] = 1.0f;
    result = affine_compose(&u, &result);
    step = affine_scale(sample_scale / 2, ANTI_BLEED / 2, ANTI_BLEED / 2, ANTI_BLEED / 2);
    result = affine_compose(&step, &result);
    return result;
}

void o2v_oracle_mesh_transform(const float mesh_min[3], const float mesh_max[3], uint32_t sample_resolution,
                               const int unit_transform[9], float out[12])
{
    const affine_t a = compute_mesh_transform(mesh_min, mesh_max, sample_resolution, unit_transform);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            out[i * 3 + j] = a.m[i][j];
        }
        out[9 + i] = a.t[i];
    }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* triangle splitting: src/voxelization.cpp:17-31,110-331 */

static int is_zero(float x)
{
    return fabsf(x) < EPSILON;
}

/* voxelization.cpp:27-31 */
static float intersect_ray_axis_plane(const float org[3], const float dir[3], uint32_t axis, uint32_t plane)
{
    const float d = -dir[axis];
    return is_zero(d) ? 0 : (org[axis] - (float) plane) / d;
}

typedef struct {
    tri_t *items;
    int count;
    int keep_hi; /* DISCARD_LO pass keeps the hi side (voxelization.cpp:388-390) */
} piece_sink;

/* LoHiPusher, voxelization.cpp:85-106 */
static void push_piece(piece_sink *sink, const tri_t *t, int is_lo)
{
    if ((is_lo != 0) != (sink->keep_hi != 0)) {
        sink->items[sink->count++] = *t;
    }
}

static void make_piece(tri_t *out, const float *p0, const float *p1, const float *p2, const float *t0, const float *t1,
                       const float *t2)
{
    memcpy(out->v[0], p0, sizeof(float) * 3);
    memcpy(out->v[1], p1, sizeof(float) * 3);
    memcpy(out->v[2], p2, sizeof(float) * 3);
    memcpy(out->t[0], t0, sizeof(float) * 2);
    memcpy(out->t[1], t1, sizeof(float) * 2);
    memcpy(out->t[2], t2, sizeof(float) * 2);
}

static void split_triangle(uint32_t axis, uint32_t plane, const tri_t *t, piece_sink *sink)
{
    /* SplittingValues, voxelization.cpp:110-153 */
    int lo[3], planar[3];
    int lo_sum = 0, planar_sum = 0;
    for (int i = 0; i < 3; ++i) {
        const float c = t->v[i][axis];
        planar[i] = is_zero(c - (float) plane);
        lo[i] = c < (float) plane;
        planar_sum += planar[i];
        lo_sum += lo[i];
    }

    /* switch of voxelization.cpp:192-234 */
    if (lo_sum == 0) {
        push_piece(sink, t, 0);
        return;
    }
    if (lo_sum == 3) {
        push_piece(sink, t, 1);
        return;
    }
    if (planar_sum == 3) {
        push_piece(sink, t, 0); /* IS_LO_BIASED == false */
        return;
    }
    if (planar_sum == 2) {
        const int first_nonplanar = !planar[0] ? 0 : !planar[1] ? 1 : 2;
        push_piece(sink, t, lo[first_nonplanar]);
        return;
    }
    if (planar_sum == 1) {
        /* splitTriangle_onePlanarCase, voxelization.cpp:240-277 */
        const int p = planar[0] ? 0 : planar[1] ? 1 : 2;
        const int a = (p + 1) % 3, b = (p + 2) % 3;
        const int nonplanar_lo = lo[a] + lo[b];
        if (nonplanar_lo != 1) {
            push_piece(sink, t, nonplanar_lo == 2);
            return;
        }
        float edge[3];
        sub3(t->v[b], t->v[a], edge);
        const float s = intersect_ray_axis_plane(t->v[a], edge, axis, plane);
        float geo[3], tex[2];
        for (int i = 0; i < 3; ++i) {
            geo[i] = mixf(t->v[a][i], t->v[b][i], s);
        }
        for (int i = 0; i < 2; ++i) {
            tex[i] = mixf(t->t[a][i], t->t[b][i], s);
        }
        tri_t first, second;
        make_piece(&first, t->v[p], t->v[a], geo, t->t[p], t->t[a], tex);
        make_piece(&second, t->v[p], geo, t->v[b], t->t[p], tex, t->t[b]);
        push_piece(sink, &first, lo[a]);
        push_piece(sink, &second, !lo[a]);
        return;
    }

    /* splitTriangle_regularCase, voxelization.cpp:279-331 */
    const int iso_lo = lo_sum == 1;
    const int iso = iso_lo ? (lo[0] ? 0 : lo[1] ? 1 : 2) : (!lo[0] ? 0 : !lo[1] ? 1 : 2);
    const int o0 = (iso + 1) % 3, o1 = (iso + 2) % 3;
    float e0[3], e1[3];
    sub3(t->v[o0], t->v[iso], e0);
    sub3(t->v[o1], t->v[iso], e1);
    const float s0 = intersect_ray_axis_plane(t->v[iso], e0, axis, plane);
    const float s1 = intersect_ray_axis_plane(t->v[iso], e1, axis, plane);
    float g0[3], g1[3], x0[2], x1[2];
    for (int i = 0; i < 3; ++i) {
        g0[i] = mixf(t->v[iso][i], t->v[o0][i], s0);
        g1[i] = mixf(t->v[iso][i], t->v[o1][i], s1);
    }
    for (int i = 0; i < 2; ++i) {
        x0[i] = mixf(t->t[iso][i], t->t[o0][i], s0);
        x1[i] = mixf(t->t[iso][i], t->t[o1][i], s1);
    }
    tri_t isolated, other0, other1;
    make_piece(&isolated, t->v[iso], g0, g1, t->t[iso], x0, x1);
    make_piece(&other0, g0, t->v[o0], t->v[o1], x0, t->t[o0], t->t[o1]);
    make_piece(&other1, g0, g1, t->v[o1], x0, x1, t->t[o1]);
    push_piece(sink, &isolated, iso_lo);
    push_piece(sink, &other0, !iso_lo);
    push_piece(sink, &other1, !iso_lo);
}

static void tri_from15(tri_t *t, const float f[15])
{
    memcpy(t->v, f, sizeof(float) * 9);
    memcpy(t->t, f + 9, sizeof(float) * 6);
}

static void tri_to15(const tri_t *t, float f[15])
{
    memcpy(f, t->v, sizeof(float) * 9);
    memcpy(f + 9, t->t, sizeof(float) * 6);
}

int o2v_oracle_split(uint32_t axis, uint32_t plane, const float tri15[15], int keep_hi, float out15[45])
{
    tri_t t, out[3];
    tri_from15(&t, tri15);
    piece_sink sink = {out, 0, keep_hi};
    split_triangle(axis, plane, &t, &sink);
    for (int i = 0; i < sink.count; ++i) {
        tri_to15(&out[i], out15 + 15 * i);
    }
    return sink.count;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* per-voxel clip: src/voxelization.cpp:383-424.  Two ping-pong lists; capacity follows ArrayVector<.,64>
 * (src/voxelization.hpp:57) with head-room: one plane turns a piece into at most two kept pieces => <= 64 after six. */

#define PIECE_CAP 192

typedef struct {
    float weight;
    float uv[2];
} weighted_uv;

static int clip_in_voxel(const tri_t *sub, const uint32_t pos[3], float whole_area, weighted_uv *result)
{
    tri_t buffer_a[PIECE_CAP], buffer_b[PIECE_CAP];
    tri_t *pre = buffer_a, *post = buffer_b;
    int pre_count = 1;
    pre[0] = *sub;

    for (uint32_t hi = 0; hi < 2; ++hi) {
        for (uint32_t axis = 0; axis < 3; ++axis) {
            const uint32_t plane = pos[axis] + hi;
            piece_sink sink = {post, 0, hi == 0};
            for (int i = 0; i < pre_count; ++i) {
                split_triangle(axis, plane, &pre[i], &sink);
            }
            if (sink.count == 0) {
                result->weight = 0;
                result->uv[0] = result->uv[1] = 0;
                return 0;
            }
            tri_t *swap = pre;
            pre = post;
            post = swap;
            pre_count = sink.count;
        }
    }

    /* fold of voxelization.cpp:414-420 with mix(Weighted) util.hpp:160-165: result = mix(result, {area, uvCentre}) */
    weighted_uv r = {0, {0, 0}};
    for (int i = 0; i < pre_count; ++i) {
        const tri_t *p = &pre[i];
        float centre[2];
        for (int k = 0; k < 2; ++k) {
            centre[k] = ((p->t[0][k] + p->t[1][k]) + p->t[2][k]) / 3; /* triangle.hpp:127-131 */
        }
        const float weight_sum = r.weight + whole_area;
        for (int k = 0; k < 2; ++k) {
            r.uv[k] = (r.weight * r.uv[k] + whole_area * centre[k]) / weight_sum;
        }
        r.weight = weight_sum;
    }
    *result = r;
    return pre_count;
}

int o2v_oracle_clip_voxel(const float tri15[15], const uint32_t pos[3], float whole_area, float out_wuv[3])
{
    tri_t t;
    tri_from15(&t, tri15);
    weighted_uv r;
    const int n = clip_in_voxel(&t, pos, whole_area, &r);
    out_wuv[0] = r.weight;
    out_wuv[1] = r.uv[0];
    out_wuv[2] = r.uv[1];
    return n;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* subdivision: src/voxelization.cpp:335-379, src/triangle.hpp:134-143 */

static int is_roughly_axis_aligned(const tri_t *t)
{
    float n[3];
    tri_normal(t, n);
    for (int i = 0; i < 3; ++i) {
        n[i] = fabsf(n[i]);
    }
    const float len = sqrtf(dot3(n, n));
    float unit[3];
    for (int i = 0; i < 3; ++i) {
        unit[i] = n[i] / len;
    }
    const float diag[3] = {SQRT_THIRD, SQRT_THIRD, SQRT_THIRD};
    const float diagonality = dot3(unit, diag);
    const float diagonality01 = (diagonality - SQRT_THIRD) / (1 - SQRT_THIRD);
    return diagonality01 < DIAGONALITY_LIMIT; /* NaN => false => subdivision path */
}

typedef struct {
    tri_t *items;
    size_t count, cap;
} tri_stack;

static void stack_push(tri_stack *s, const tri_t *t)
{
    if (s->count == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 64;
        s->items = (tri_t *) realloc(s->items, s->cap * sizeof(tri_t));
    }
    s->items[s->count++] = *t;
}

typedef void (*leaf_fn)(const tri_t *leaf, void *user);

/* forEachSubdividedTriangle: LIFO; the centre piece replaces the top, corners 1..3 are pushed, so corner 3 is visited
 * first and the centre subtree last. */
static void for_each_leaf(const tri_t *input, tri_stack *stack, leaf_fn fn, void *user)
{
    if (is_roughly_axis_aligned(input)) {
        fn(input, user);
        return;
    }
    stack->count = 0;
    stack_push(stack, input);
    while (stack->count != 0) {
        const tri_t top = stack->items[stack->count - 1];
        uint32_t vmin[3], vmax[3];
        tri_voxel_bounds(&top, vmin, vmax);
        const uint32_t volume = (vmax[0] - vmin[0]) * (vmax[1] - vmin[1]) * (vmax[2] - vmin[2]); /* u32 wrap: B6 */
        if (volume < SUBDIVISION_VOLUME_LIMIT) {
            --stack->count;
            fn(&top, user);
            continue;
        }
        float g[3][3], x[3][2];
        for (int e = 0; e < 3; ++e) {
            const int a = e, b = (e + 1) % 3;
            for (int i = 0; i < 3; ++i) {
                g[e][i] = mixf(top.v[a][i], top.v[b][i], 0.5f);
            }
            for (int i = 0; i < 2; ++i) {
                x[e][i] = mixf(top.t[a][i], top.t[b][i], 0.5f);
            }
        }
        tri_t centre, c1, c2, c3;
        make_piece(&centre, g[0], g[1], g[2], x[0], x[1], x[2]);
        make_piece(&c1, top.v[0], g[0], g[2], top.t[0], x[0], x[2]);
        make_piece(&c2, top.v[1], g[1], g[0], top.t[1], x[1], x[0]);
        make_piece(&c3, top.v[2], g[2], g[1], top.t[2], x[2], x[1]);
        stack->items[stack->count - 1] = centre;
        stack_push(stack, &c1);
        stack_push(stack, &c2);
        stack_push(stack, &c3);
    }
}

typedef struct {
    float *out;
    size_t cap, count;
} leaf_collect;

static void collect_leaf(const tri_t *leaf, void *user)
{
    leaf_collect *c = (leaf_collect *) user;
    if (c->count < c->cap) {
        tri_to15(leaf, c->out + 15 * c->count);
    }
    ++c->count;
}

size_t o2v_oracle_subdivide(const float tri15[15], float *out_leaves, size_t cap)
{
    tri_t t;
    tri_from15(&t, tri15);
    tri_stack stack = {0, 0, 0};
    leaf_collect c = {out_leaves, cap, 0};
    for_each_leaf(&t, &stack, collect_leaf, &c);
    free(stack.items);
    return c.count;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* open-addressed Morton-keyed maps standing in for VoxelMap<T> = std::unordered_map<u64, T> (util.hpp:179-208).
 * Iteration order never matters for the results (keys within one triangle are distinct, voxelization.cpp:513-526). */

typedef struct {
    uint64_t key_plus_one; /* 0 = empty */
    float w;
    float v[3];
} map_slot;

typedef struct {
    map_slot *slots;
    size_t cap, count;
    uint32_t *used; /* indices of occupied slots, for fast clear/iteration */
    size_t used_cap;
} voxel_map;

static uint64_t hash64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

static void map_init(voxel_map *m, size_t cap)
{
    m->cap = cap;
    m->count = 0;
    m->slots = (map_slot *) calloc(cap, sizeof(map_slot));
    m->used_cap = cap / 2 + 1;
    m->used = (uint32_t *) malloc(m->used_cap * sizeof(uint32_t));
}

static void map_free(voxel_map *m)
{
    free(m->slots);
    free(m->used);
    memset(m, 0, sizeof *m);
}

static map_slot *map_find_or_insert(voxel_map *m, uint64_t key, int *inserted);

static void map_grow(voxel_map *m)
{
    voxel_map bigger;
    map_init(&bigger, m->cap * 2);
    for (size_t i = 0; i < m->count; ++i) {
        const map_slot *s = &m->slots[m->used[i]];
        int inserted;
        map_slot *d = map_find_or_insert(&bigger, s->key_plus_one - 1, &inserted);
        d->w = s->w;
        memcpy(d->v, s->v, sizeof d->v);
    }
    map_free(m);
    *m = bigger;
}

static map_slot *map_find_or_insert(voxel_map *m, uint64_t key, int *inserted)
{
    if ((m->count + 1) * 2 > m->cap) {
        map_grow(m);
    }
    size_t i = hash64(key) & (m->cap - 1);
    for (;;) {
        map_slot *s = &m->slots[i];
        if (s->key_plus_one == key + 1) {
            *inserted = 0;
            return s;
        }
        if (s->key_plus_one == 0) {
            s->key_plus_one = key + 1;
            m->used[m->count++] = (uint32_t) i;
            *inserted = 1;
            return s;
        }
        i = (i + 1) & (m->cap - 1);
    }
}

static void map_clear(voxel_map *m)
{
    for (size_t i = 0; i < m->count; ++i) {
        m->slots[m->used[i]].key_plus_one = 0;
    }
    m->count = 0;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* colour lookup: src/triangle.hpp:181-194, voxelio image.hpp:87-95,159-194, image.cpp:85-98, color.hpp:38-52 */

static float wrap_repeat(float x)
{
    float integral;
    float fraction = modff(x, &integral);
    fraction += fraction < 0;
    fraction += fraction == 0;
    return fraction;
}

static float clamp01(float x)
{
    /* color.hpp:152-156: std::min(std::max(x, 0), 1) */
    const float lo = x < 0 ? 0.0f : x;
    return 1.0f < lo ? 1.0f : lo;
}

void o2v_oracle_texture_lookup(const o2v_oracle_texture *tex, const float uv[2], float out_rgb[3])
{
    const float u = uv[0];
    const float v = 1 - uv[1]; /* triangle.hpp:190 */
    const float wu = tex->wrap == O2V_ORACLE_UV_CLAMP ? clamp01(u) : wrap_repeat(u);
    const float wv = tex->wrap == O2V_ORACLE_UV_CLAMP ? clamp01(v) : wrap_repeat(v);
    const size_t x = (size_t) (wu * (float) (tex->width - 1));
    const size_t y = (size_t) (wv * (float) (tex->height - 1));
    const uint8_t *in = tex->pixels + (y * tex->width + x) * (size_t) tex->channels;
    uint8_t r, g, b;
    if (tex->channels == 3) { /* decodeRgb24 */
        r = in[0];
        g = in[1];
        b = in[2];
    }
    else { /* decodeArgb32 as written in voxelio/src/image.cpp:95-98: Color32{in[3], in[0], in[1], in[2]} = (r,g,b,a) */
        r = in[3];
        g = in[0];
        b = in[1];
    }
    out_rgb[0] = r / 255.f;
    out_rgb[1] = g / 255.f;
    out_rgb[2] = b / 255.f;
}

uint32_t o2v_oracle_quantize_argb(const float rgb[3])
{
    const uint32_t r = (uint8_t) (clamp01(rgb[0]) * 0xFF);
    const uint32_t g = (uint8_t) (clamp01(rgb[1]) * 0xFF);
    const uint32_t b = (uint8_t) (clamp01(rgb[2]) * 0xFF);
    return 0xFF000000u | (r << 16) | (g << 8) | b;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Voxelizer::voxelize for one (triangle, clip box): src/voxelization.cpp:426-526 */

typedef struct {
    tri_t tri;      /* voxel space */
    uint8_t type;
    float color[3];
    float area;     /* whole-triangle area (voxelization.cpp:416) */
    uint32_t vmin[3], vmax[3];
} prepared_tri;

typedef struct {
    const prepared_tri *input;
    uint32_t box_min[3], box_max[3];
    voxel_map *uv_buffer;
    uint64_t subtriangles;
} sub_context;

/* voxelizeSubTriangle, voxelization.cpp:426-472 */
static void voxelize_sub_triangle(const tri_t *sub, void *user)
{
    sub_context *ctx = (sub_context *) user;
    ++ctx->subtriangles;

    float n[3];
    tri_normal(sub, n);
    const float len = sqrtf(dot3(n, n));
    float unit[3];
    for (int i = 0; i < 3; ++i) {
        unit[i] = n[i] / len;
    }
    uint32_t lo[3], hi[3];
    tri_voxel_bounds(sub, lo, hi);
    for (int i = 0; i < 3; ++i) {
        lo[i] = lo[i] > ctx->box_min[i] ? lo[i] : ctx->box_min[i];
        hi[i] = hi[i] < ctx->box_max[i] ? hi[i] : ctx->box_max[i];
    }

    for (uint32_t z = lo[2]; z < hi[2]; ++z) {
        for (uint32_t y = lo[1]; y < hi[1]; ++y) {
            for (uint32_t x = lo[0]; x < hi[0]; ++x) {
                const uint32_t pos[3] = {x, y, z};
                const float centre[3] = {(float) x + 0.5f, (float) y + 0.5f, (float) z + 0.5f};
                float rel[3];
                sub3(centre, sub->v[0], rel);
                const float signed_distance = dot3(unit, rel);
                if (fabsf(signed_distance) > DISTANCE_LIMIT) {
                    continue;
                }
                weighted_uv uv;
                clip_in_voxel(sub, pos, ctx->input->area, &uv);
                if (uv.weight != 0.f) {
                    /* insertWeighted<BLEND>(uvBuffer, pos, uv): mix(new, existing), voxelization.cpp:56-63 */
                    int inserted;
                    map_slot *s = map_find_or_insert(ctx->uv_buffer, o2v_oracle_ileave3(x, y, z), &inserted);
                    if (inserted) {
                        s->w = uv.weight;
                        s->v[0] = uv.uv[0];
                        s->v[1] = uv.uv[1];
                    }
                    else {
                        const float weight_sum = uv.weight + s->w;
                        s->v[0] = (uv.weight * uv.uv[0] + s->w * s->v[0]) / weight_sum;
                        s->v[1] = (uv.weight * uv.uv[1] + s->w * s->v[1]) / weight_sum;
                        s->w = weight_sum;
                    }
                }
            }
        }
    }
}

/* combine(new, existing): util.hpp:160-172 */
static void combine_into(map_slot *existing, float w, const float c[3], int strategy)
{
    if (strategy == O2V_ORACLE_BLEND) {
        const float weight_sum = w + existing->w;
        for (int i = 0; i < 3; ++i) {
            existing->v[i] = (w * c[i] + existing->w * existing->v[i]) / weight_sum;
        }
        existing->w = weight_sum;
    }
    else if (w > existing->w) {
        existing->w = w;
        memcpy(existing->v, c, sizeof(float) * 3);
    }
}

typedef struct {
    voxel_map voxels;
    voxel_map uv_buffer;
    tri_stack stack;
    uint64_t contributions;
    uint64_t subtriangles;
} worker_state;

static void voxelize_triangle_in_box(worker_state *w, const prepared_tri *p, const uint32_t box_min[3],
                                     const uint32_t box_max[3], const o2v_oracle_texture *texture, int strategy)
{
    sub_context ctx;
    ctx.input = p;
    memcpy(ctx.box_min, box_min, sizeof ctx.box_min);
    memcpy(ctx.box_max, box_max, sizeof ctx.box_max);
    ctx.uv_buffer = &w->uv_buffer;
    ctx.subtriangles = 0;
    for_each_leaf(&p->tri, &w->stack, voxelize_sub_triangle, &ctx);
    w->subtriangles += ctx.subtriangles;

    /* moveUvBufferIntoVoxels, voxelization.cpp:513-526 */
    for (size_t i = 0; i < w->uv_buffer.count; ++i) {
        const map_slot *u = &w->uv_buffer.slots[w->uv_buffer.used[i]];
        float color[3];
        if (p->type == O2V_ORACLE_TEXTURED) {
            o2v_oracle_texture_lookup(texture, u->v, color);
        }
        else if (p->type == O2V_ORACLE_UNTEXTURED) {
            memcpy(color, p->color, sizeof color);
        }
        else {
            color[0] = color[1] = color[2] = 1.0f;
        }
        int inserted;
        map_slot *s = map_find_or_insert(&w->voxels, u->key_plus_one - 1, &inserted);
        ++w->contributions;
        if (inserted) {
            s->w = u->w;
            memcpy(s->v, color, sizeof color);
        }
        else {
            combine_into(s, u->w, color, strategy);
        }
    }
    map_clear(&w->uv_buffer);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* driver: src/obj2voxel.cpp:180-252,467-520 (bounds, transform, 64^3 chunks) */

typedef struct {
    uint64_t key;
    float w;
    float v[3];
} out_voxel;

static int compare_out_voxel(const void *a, const void *b)
{
    const uint64_t ka = ((const out_voxel *) a)->key, kb = ((const out_voxel *) b)->key;
    return ka < kb ? -1 : ka > kb ? 1 : 0;
}

typedef struct {
    const prepared_tri *prepared;
    size_t triangle_count;
    const o2v_oracle_texture *texture;
    int strategy;
    uint32_t chunks_per_axis;
    size_t chunk_count;
    size_t next_chunk; /* guarded by lock */
    pthread_mutex_t lock;
    out_voxel *all;
    size_t all_count, all_cap;
    uint64_t contributions, subtriangles;
} chunk_job;

/* One worker: pulls 64^3 chunks, voxelizes every overlapping triangle in ascending index order — the reference's
 * per-chunk list order (obj2voxel.cpp:226-243,270-272) — and appends the chunk's voxels to the shared output. */
static void *chunk_worker(void *user)
{
    chunk_job *job = (chunk_job *) user;
    worker_state w;
    memset(&w, 0, sizeof w);
    map_init(&w.voxels, 1u << 16);
    map_init(&w.uv_buffer, 1u << 12);

    for (;;) {
        pthread_mutex_lock(&job->lock);
        const size_t chunk = job->next_chunk++;
        pthread_mutex_unlock(&job->lock);
        if (chunk >= job->chunk_count) {
            break;
        }
        const uint32_t per_axis = job->chunks_per_axis;
        const uint32_t cx = (uint32_t) (chunk % per_axis);
        const uint32_t cy = (uint32_t) ((chunk / per_axis) % per_axis);
        const uint32_t cz = (uint32_t) (chunk / ((size_t) per_axis * per_axis));
        const uint32_t box_min[3] = {cx * CHUNK_SIZE, cy * CHUNK_SIZE, cz * CHUNK_SIZE};
        const uint32_t box_max[3] = {box_min[0] + CHUNK_SIZE, box_min[1] + CHUNK_SIZE, box_min[2] + CHUNK_SIZE};

        for (size_t t = 0; t < job->triangle_count; ++t) {
            const prepared_tri *p = &job->prepared[t];
            int overlaps = 1;
            for (int i = 0; i < 3; ++i) {
                overlaps &= p->vmin[i] < box_max[i] && p->vmax[i] > box_min[i];
            }
            if (overlaps) {
                voxelize_triangle_in_box(&w, p, box_min, box_max, job->texture, job->strategy);
            }
        }
        if (w.voxels.count != 0) {
            pthread_mutex_lock(&job->lock);
            if (job->all_count + w.voxels.count > job->all_cap) {
                job->all_cap = (job->all_count + w.voxels.count) * 2;
                job->all = (out_voxel *) realloc(job->all, job->all_cap * sizeof(out_voxel));
            }
            for (size_t i = 0; i < w.voxels.count; ++i) {
                const map_slot *s = &w.voxels.slots[w.voxels.used[i]];
                out_voxel *o = &job->all[job->all_count++];
                o->key = s->key_plus_one - 1;
                o->w = s->w;
                memcpy(o->v, s->v, sizeof o->v);
            }
            pthread_mutex_unlock(&job->lock);
            map_clear(&w.voxels);
        }
    }
    pthread_mutex_lock(&job->lock);
    job->contributions += w.contributions;
    job->subtriangles += w.subtriangles;
    pthread_mutex_unlock(&job->lock);
    map_free(&w.voxels);
    map_free(&w.uv_buffer);
    free(w.stack.items);
    return NULL;
}

void o2v_oracle_default_params(o2v_oracle_params *p)
{
    memset(p, 0, sizeof *p);
    p->supersampling = 1;
    p->strategy = O2V_ORACLE_MAX;
    p->unit_transform[0] = p->unit_transform[4] = p->unit_transform[8] = 1;
    p->downscale = 1;
}

int o2v_oracle_voxelize(const o2v_oracle_params *params, size_t n, const float *verts, const float *uvs,
                        const uint8_t *types, const float *colors, const o2v_oracle_texture *texture,
                        o2v_oracle_result *out)
{
    memset(out, 0, sizeof *out);
    if (params->resolution == 0 || params->supersampling == 0 || params->supersampling > 2) {
        return 1;
    }
    const uint32_t S = params->resolution * params->supersampling;

    /* findMeshBounds, obj2voxel.cpp:180-200 */
    float mesh_min[3], mesh_max[3];
    if (params->bounds_known) {
        memcpy(mesh_min, params->bounds, sizeof mesh_min);
        memcpy(mesh_max, params->bounds + 3, sizeof mesh_max);
    }
    else {
        for (int i = 0; i < 3; ++i) {
            mesh_min[i] = INFINITY;
            mesh_max[i] = -INFINITY;
        }
        for (size_t t = 0; t < n; ++t) {
            for (int k = 0; k < 3; ++k) {
                for (int i = 0; i < 3; ++i) {
                    const float c = verts[t * 9 + k * 3 + i];
                    mesh_min[i] = c < mesh_min[i] ? c : mesh_min[i];
                    mesh_max[i] = mesh_max[i] < c ? c : mesh_max[i];
                }
            }
        }
    }
    if (n == 0) {
        return 0; /* obj2voxel.cpp:590-594: empty model, empty output */
    }

    const affine_t transform = compute_mesh_transform(mesh_min, mesh_max, S, params->unit_transform);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            out->transform[i * 3 + j] = transform.m[i][j];
        }
        out->transform[9 + i] = transform.t[i];
    }

    /* applyMeshTransform, obj2voxel.cpp:202-224 */
    prepared_tri *prepared = (prepared_tri *) malloc(n * sizeof(prepared_tri));
    for (size_t t = 0; t < n; ++t) {
        prepared_tri *p = &prepared[t];
        memset(p, 0, sizeof *p);
        for (int k = 0; k < 3; ++k) {
            affine_apply(&transform, verts + t * 9 + k * 3, p->tri.v[k]);
            if (uvs != NULL) {
                p->tri.t[k][0] = uvs[t * 6 + k * 2];
                p->tri.t[k][1] = uvs[t * 6 + k * 2 + 1];
            }
        }
        p->type = types != NULL ? types[t]
                                : (uvs != NULL && texture != NULL ? O2V_ORACLE_TEXTURED : O2V_ORACLE_MATERIALLESS);
        if (p->type == O2V_ORACLE_UNTEXTURED && colors != NULL) {
            memcpy(p->color, colors + t * 3, sizeof p->color);
        }
        p->area = tri_area(&p->tri);
        tri_voxel_bounds(&p->tri, p->vmin, p->vmax);
        /* A negative voxel-space coordinate makes the reference's float -> u32 cast wrap to a huge chunkMin
         * (triangle.hpp:91-95, obj2voxel.cpp:211-219; formally UB): the triangle lands in no chunk and is
         * dropped as a whole.  Reproduced here as the observed x86-64 behaviour. */
        for (int i = 0; i < 3; ++i) {
            if (floorf(min3f(p->tri.v[0][i], p->tri.v[1][i], p->tri.v[2][i])) < 0) {
                p->vmin[i] = p->vmax[i] = 0; /* overlaps no chunk */
            }
        }
    }

    /* chunk grid: obj2voxel.cpp:245-252,580-581 (full coverage; the reference's lost chunks for non-power-of-two grids,
     * SURVEY B2, are not reproduced) */
    const uint32_t chunks_per_axis = (S + CHUNK_SIZE - 1) / CHUNK_SIZE;
    const size_t chunk_count = (size_t) chunks_per_axis * chunks_per_axis * chunks_per_axis;

    int threads = params->threads;
    if (threads <= 0) {
        const long online = sysconf(_SC_NPROCESSORS_ONLN);
        threads = online > 0 ? (int) online : 1;
    }
    if ((size_t) threads > chunk_count) {
        threads = (int) chunk_count;
    }

    chunk_job job;
    memset(&job, 0, sizeof job);
    job.prepared = prepared;
    job.triangle_count = n;
    job.texture = texture;
    job.strategy = params->strategy;
    job.chunks_per_axis = chunks_per_axis;
    job.chunk_count = chunk_count;
    pthread_mutex_init(&job.lock, NULL);

    if (threads <= 1) {
        chunk_worker(&job);
    }
    else {
        pthread_t *handles = (pthread_t *) malloc((size_t) threads * sizeof(pthread_t));
        for (int i = 0; i < threads; ++i) {
            pthread_create(&handles[i], NULL, chunk_worker, &job);
        }
        for (int i = 0; i < threads; ++i) {
            pthread_join(handles[i], NULL);
        }
        free(handles);
    }
    pthread_mutex_destroy(&job.lock);

    out_voxel *all = job.all;
    size_t all_count = job.all_count;
    const uint64_t contributions = job.contributions, subtriangles = job.subtriangles;
    free(prepared);

    /* supersampling: intended semantics (README.adoc:153-163); the unmodified Voxelizer::downscale
     * (voxelization.cpp:538-554) returns an empty map.  Children folded in ascending Morton order with combine(child, acc). */
    if (params->supersampling == 2 && params->downscale && all_count != 0) {
        qsort(all, all_count, sizeof(out_voxel), compare_out_voxel);
        size_t write = 0;
        for (size_t i = 0; i < all_count; ++i) {
            const uint64_t parent = all[i].key >> 3;
            if (write != 0 && all[write - 1].key == parent) {
                map_slot acc;
                acc.w = all[write - 1].w;
                memcpy(acc.v, all[write - 1].v, sizeof acc.v);
                combine_into(&acc, all[i].w, all[i].v, params->strategy);
                all[write - 1].w = acc.w;
                memcpy(all[write - 1].v, acc.v, sizeof acc.v);
            }
            else {
                out_voxel o = all[i];
                o.key = parent;
                all[write++] = o;
            }
        }
        all_count = write;
    }

    out->count = all_count;
    out->contributions = contributions;
    out->subtriangles = subtriangles;
    out->xyz = (uint32_t *) malloc((all_count ? all_count : 1) * 3 * sizeof(uint32_t));
    out->argb = (uint32_t *) malloc((all_count ? all_count : 1) * sizeof(uint32_t));
    out->wrgb = (float *) malloc((all_count ? all_count : 1) * 4 * sizeof(float));
    for (size_t i = 0; i < all_count; ++i) {
        o2v_oracle_dileave3(all[i].key, out->xyz + 3 * i);
        out->wrgb[4 * i] = all[i].w;
        memcpy(out->wrgb + 4 * i + 1, all[i].v, sizeof(float) * 3);
        out->argb[i] = o2v_oracle_quantize_argb(all[i].v);
    }
    free(all);
    return 0;
}

void o2v_oracle_free_result(o2v_oracle_result *r)
{
    free(r->xyz);
    free(r->argb);
    free(r->wrgb);
    memset(r, 0, sizeof *r);
}
