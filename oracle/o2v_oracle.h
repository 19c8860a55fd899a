/* TEST INFRASTRUCTURE ONLY — CPU restatement of the obj2voxel per-triangle voxelization hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.  The product
 * (obj2voxel_b200/csrc, libobj2voxel_b200.so) never links, calls or falls back to it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit-for-bit (positions, ARGB8, float weight and
 * float RGB) against the unmodified reference built into oracle/_ref/ (when present) and against the committed golden
 * fixtures under tests/golden/ that were generated from that build (tests/golden/make_golden.py).
 *
 * Every function cites the reference location (relative to /root/reference) whose semantics it restates.
 */
#ifndef O2V_ORACLE_H
#define O2V_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* obj2voxel::TriangleType, src/triangle.hpp:21-30 */
enum { O2V_ORACLE_MATERIALLESS = 1, O2V_ORACLE_UNTEXTURED = 2, O2V_ORACLE_TEXTURED = 3 };
/* include/obj2voxel.h:43-46 */
enum { O2V_ORACLE_MAX = 0, O2V_ORACLE_BLEND = 1 };
/* include/obj2voxel.h:48-51 */
enum { O2V_ORACLE_UV_CLAMP = 0, O2V_ORACLE_UV_WRAP = 1 };

typedef struct {
    const uint8_t *pixels; /* row-major, channels bytes per pixel */
    size_t width;
    size_t height;
    int channels; /* 3 = RGB24, 4 = ARGB32 as decoded by voxelio (src/obj2voxel.cpp:349-356, voxelio/src/image.cpp:85-98) */
    int wrap;     /* O2V_ORACLE_UV_* */
} o2v_oracle_texture;

typedef struct {
    uint32_t resolution;    /* output resolution R */
    uint32_t supersampling; /* 1 or 2; sample resolution S = R * supersampling (src/obj2voxel.cpp:689,697) */
    int strategy;           /* O2V_ORACLE_MAX / O2V_ORACLE_BLEND */
    int bounds_known;       /* src/obj2voxel.cpp:804-816 */
    float bounds[6];
    int unit_transform[9];  /* row-major, src/obj2voxel.cpp:797-802 */
    int downscale;          /* 0: emit at S (what the unmodified reference computes before its broken downscale);
                               1: intended-semantics downscale to R, children folded in ascending Morton order */
    int threads;            /* worker threads over 64^3 chunks; <=0 = all online cores */
} o2v_oracle_params;

typedef struct {
    size_t count;
    uint32_t *xyz;          /* 3 per voxel */
    uint32_t *argb;         /* 1 per voxel, src/obj2voxel.cpp:283-296 + voxelio color.hpp:165-173 */
    float *wrgb;            /* weight, r, g, b per voxel (pre-quantisation WeightedColor) */
    float transform[12];    /* mesh->voxel affine: 3x3 row-major then translation */
    uint64_t contributions; /* N_contrib: emplace attempts of voxelization.cpp:520 */
    uint64_t subtriangles;  /* leaves emitted by forEachSubdividedTriangle (or 1 per aligned triangle) */
} o2v_oracle_result;

void o2v_oracle_default_params(o2v_oracle_params *params);

/* verts: 9 floats per triangle (model space); uvs: 6 per triangle or NULL; types: 1 per triangle or NULL
 * (NULL = TEXTURED if uvs && texture, else MATERIALLESS); colors: 3 per triangle or NULL (UNTEXTURED only).
 * Returns 0 on success.  Output is in no particular order; callers sort. */
int o2v_oracle_voxelize(const o2v_oracle_params *params, size_t triangle_count, const float *verts, const float *uvs,
                        const uint8_t *types, const float *colors, const o2v_oracle_texture *texture,
                        o2v_oracle_result *out);

void o2v_oracle_free_result(o2v_oracle_result *result);

/* ---- building blocks, exported for unit tests ---- */

/* voxelio ileave.hpp:243-246 / :270-275 */
uint64_t o2v_oracle_ileave3(uint32_t x, uint32_t y, uint32_t z);
void o2v_oracle_dileave3(uint64_t n, uint32_t out[3]);

/* src/obj2voxel.cpp:370-402 with util.hpp:262-281; out = 3x3 row-major + translation */
void o2v_oracle_mesh_transform(const float mesh_min[3], const float mesh_max[3], uint32_t sample_resolution,
                               const int unit_transform[9], float out[12]);

/* src/voxelization.cpp:175-331.  tri15 = 9 position floats then 6 uv floats.  Returns number of kept pieces written to
 * out15 (0..3). */
int o2v_oracle_split(uint32_t axis, uint32_t plane, const float tri15[15], int keep_hi, float out15[45]);

/* src/voxelization.cpp:383-424: six-plane clip of one sub-triangle in voxel `pos`; whole_area = area of the whole input
 * triangle.  Returns the piece count; out_wuv = {weight, u, v}. */
int o2v_oracle_clip_voxel(const float tri15[15], const uint32_t pos[3], float whole_area, float out_wuv[3]);

/* src/voxelization.cpp:335-379: writes up to cap leaves (15 floats each) in emission order; returns the leaf count. */
size_t o2v_oracle_subdivide(const float tri15[15], float *out_leaves, size_t cap);

/* voxelio color.hpp:165-173, :95 */
uint32_t o2v_oracle_quantize_argb(const float rgb[3]);

/* triangle.hpp:181-194 + image.hpp:87-95,159-194 (TEXTURED lookup at uv) */
void o2v_oracle_texture_lookup(const o2v_oracle_texture *texture, const float uv[2], float out_rgb[3]);

#ifdef __cplusplus
}
#endif
#endif
