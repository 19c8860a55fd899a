/* Additive C-ABI of libobj2voxel_b200.so: bulk / device-resident entry points around the same kernels that serve
 * obj2voxel_voxelize().  Plain pointers and sizes only (no torch / C++ types).
 *
 * What each entry point stands in for in the reference:
 *   o2v_b200_voxelize_device   the hot path itself — obj2voxel::Voxelizer::voxelize() per (triangle, chunk)
 *                              (src/voxelization.cpp:480-526) together with the driver steps that feed it
 *                              (findMeshBounds / computeMeshTransform / applyMeshTransform / chunk sort / voxelizeChunk,
 *                              src/obj2voxel.cpp:180-314,370-402) — with inputs and outputs in HBM
 *   o2v_b200_voxelize_host     the same with host buffers (H2D + kernels + D2H), i.e. obj2voxel_voxelize() without the
 *                              one-indirect-call-per-triangle ITriangleStream (src/obj2voxel.cpp:585-588)
 *   obj2voxel_b200_set_input_triangles   bulk replacement for obj2voxel_set_input_callback (include/obj2voxel.h:177)
 *   slab_z0 / slab_z1          the unit of multi-GPU partitioning: a Z-slab of 64^3-chunk rows
 *                              (src/obj2voxel.cpp:245-252 computeChunkBounds gives the reference's clip boxes)
 */
#ifndef OBJ2VOXEL_B200_HEADER
#define OBJ2VOXEL_B200_HEADER

#include <stddef.h>
#include <stdint.h>

#include "obj2voxel.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct o2v_b200_engine o2v_b200_engine;

/* obj2voxel::TriangleType, reference src/triangle.hpp:21-30 */
#define O2V_B200_TRI_MATERIALLESS 1
#define O2V_B200_TRI_UNTEXTURED 2
#define O2V_B200_TRI_TEXTURED 3

typedef struct o2v_b200_params {
    uint32_t resolution;        /* output resolution R */
    uint32_t supersampling;     /* 1 or 2 (sample resolution S = R * supersampling) */
    uint32_t strategy;          /* OBJ2VOXEL_MAX_STRATEGY / OBJ2VOXEL_BLEND_STRATEGY */
    uint32_t bounds_known;      /* 0: compute mesh bounds on the device */
    float bounds[6];            /* min xyz, max xyz */
    int32_t unit_transform[9];  /* row-major */
    uint32_t slab_z0, slab_z1;  /* owned sample-space z range [z0, z1), multiples of 8; 0,0 = the whole grid */
    int32_t variant;            /* kernel A/B switch, -1 = default (occupancy-only path: 1 / 2 force the block-per-batch /
                                 * thread-per-leaf classifier; both give the same records) */
    int32_t prefilter;          /* 1 = conservative SAT prefilter on (default); 0 = off (validation only) */
    int32_t float_records;      /* parity / debug: 1 = also keep every voxel's float (weight, r, g, b) as the fold left it,
                                 * before the ARGB8 truncation — what obj2voxel::Voxelizer::voxels() holds in the
                                 * reference (src/voxelization.hpp:55-108); forces the weighted pipeline; read them with
                                 * o2v_b200_result_floats_device() */
    int32_t slab_filtered;      /* 1 = mesh is the output of o2v_b200_filter_slab() for this very slab: the step skips
                                 * its own filter pass (multi-GPU ingest distributes triangles by z range once) */
    int32_t occupancy_path;     /* 1 (default) = meshes whose every triangle is MATERIALLESS (output colour is white
                                 * whatever the weights, reference src/triangle.hpp:186) take the occupancy-only path;
                                 * 0 = always fold weights and colours (validation / measurement) */
    int32_t accumulate;         /* occupancy-only path, a mesh voxelized piece by piece (what obj2voxel_voxelize() does while
                                 * the triangles are still crossing PCIe): 0 = an ordinary run; 1 = the first piece, 2 = a
                                 * later piece of the same grid / slab on the same engine.  A piece's result is the voxels
                                 * no earlier piece has produced, so the pieces' results are disjoint and their union is
                                 * the whole mesh's (the output is an OR).  A piece runs once: o2v_b200_voxelize_host needs
                                 * room for all of its records on the first call (a repeated call finds them delivered). */
} o2v_b200_params;

typedef struct o2v_b200_mesh {
    const float *verts;          /* 9 floats per triangle, model space */
    const float *uvs;            /* 6 floats per triangle or NULL */
    const uint8_t *types;        /* O2V_B200_TRI_* per triangle or NULL (textured if uvs and textures, else materialless) */
    const float *colors;         /* 3 floats per triangle or NULL (O2V_B200_TRI_UNTEXTURED) */
    const uint32_t *texture_ids; /* per-triangle index into the texture array or NULL (0) */
    uint64_t count;
} o2v_b200_mesh;

typedef struct o2v_b200_texture {
    const uint8_t *pixels; /* row-major, `channels` bytes per pixel */
    uint32_t width, height;
    uint32_t channels; /* 3 or 4 */
    uint32_t wrap;     /* OBJ2VOXEL_UV_CLAMP / OBJ2VOXEL_UV_WRAP */
} o2v_b200_texture;

typedef struct o2v_b200_stats {
    uint64_t voxels;            /* emitted voxels */
    uint64_t leaves;            /* sub-triangles after the reference's subdivision */
    uint64_t pairs;             /* (leaf, tile) pairs */
    uint64_t active_tiles;
    uint64_t candidate_voxels;  /* sum of leaf AABB volumes */
    uint64_t clip_calls;        /* exact six-plane clips executed */
    uint64_t contributions;     /* (triangle, voxel) merges = N_contrib */
    uint64_t dropped_triangles; /* zero-area / non-finite */
    uint64_t depth_overflow;
    uint64_t out_capacity;
    float ms_total, ms_setup, ms_voxelize; /* device time, CUDA events on the run stream */
    float transform[12];
    int32_t kernel_launches;
    int32_t voxelize_launches;
    uint64_t light_tiles;       /* tiles voxelized warp-per-tile */
    uint64_t heavy_tiles;       /* tiles voxelized block-per-tile */
    uint64_t survivors;         /* sparse path: SAT survivors = exact clips */
    float ms_clip;              /* duration of the exact-clip kernel, CUDA events */
    int32_t occupancy_path;     /* 1 = this run took the occupancy-only path (survivors = voxels the SAT left undecided) */
    float ms_classify;          /* occupancy-only path: duration of the SAT classification kernel, CUDA events */
    float reserved;
    uint64_t slab_triangles;    /* occupancy-only path: triangles the step worked on (all of them, or what the slab filter
                                 * kept on a rank / job part that owns a part of the grid) */
    uint64_t undecided_ranges;  /* occupancy-only path: rows with voxels the SAT left undecided (their x runs are checked
                                 * against the bitmap; `survivors` of them reach the exact clip) */
    float ms_filter;            /* occupancy-only path: duration of that check (occupancyFilterQueueKernel) */
    float ms_expand;            /* occupancy-only path: bitmap -> records (occupancyExpandKernel) */
    uint64_t download_bytes;    /* obj2voxel_voxelize(): bytes copied device -> host (records, or bitmaps + chunk lists) */
} o2v_b200_stats;

/* NULL when no CUDA device is usable (no CPU fallback); see o2v_b200_last_error(). */
o2v_b200_engine *o2v_b200_engine_create(int device);
void o2v_b200_engine_destroy(o2v_b200_engine *engine);
const char *o2v_b200_last_error(void);
int o2v_b200_sm_count(const o2v_b200_engine *engine);
void o2v_b200_default_params(o2v_b200_params *params);

/* All pointers inside mesh and textures[].pixels are DEVICE pointers; textures itself is a host array.
 * cuda_stream is a cudaStream_t (NULL = default stream).  Returns 0 or a negative error code. */
int o2v_b200_voxelize_device(o2v_b200_engine *engine, const o2v_b200_params *params, const o2v_b200_mesh *mesh,
                             const o2v_b200_texture *textures, uint32_t texture_count, void *cuda_stream,
                             o2v_b200_stats *out_stats);

/* Result of the last run: `count` records of {int32 x, y, z; uint32 argb} in device memory owned by the engine. */
const void *o2v_b200_result_device(const o2v_b200_engine *engine);
uint64_t o2v_b200_result_count(const o2v_b200_engine *engine);
int o2v_b200_result_download(o2v_b200_engine *engine, void *host_dst, void *cuda_stream);
/* `count` x 4 floats (weight, r, g, b), index-aligned with o2v_b200_result_device(); NULL unless the last run had
 * float_records = 1. */
const float *o2v_b200_result_floats_device(const o2v_b200_engine *engine);

/* Multi-GPU ingest on the occupancy-only path (all-MATERIALLESS meshes): copies the triangles of the DEVICE mesh whose z
 * range can reach the slab [slab_z0, slab_z1) of `params` into an engine-owned dense device array (*out_kept: 9 floats
 * per triangle, arbitrary order, valid until the next call).  Voxelizing that array with slab_filtered = 1 gives the
 * slab's records without touching the rest of the mesh again — the "triangles distributed once" of the Z-slab scheme
 * (the reference re-tests every triangle against every chunk it might touch: src/obj2voxel.cpp:211-243). */
int o2v_b200_filter_slab(o2v_b200_engine *engine, const o2v_b200_params *params, const o2v_b200_mesh *mesh,
                         void *cuda_stream, const float **out_kept, uint64_t *out_count);

/* Order-independent 64-bit checksum of the last run's records, computed on the device: the sum mod 2^64 over the records
 * of splitmix64_finalise((x + (y << 21) + (z << 42)) ^ (argb * 0x9E3779B97F4A7C15)).  Z-slabs / job parts / ranks add their
 * sums; the total equals the same sum over the reference's output (voxelio Voxel32 list, src/io.cpp:638-653) whatever the
 * order, so a multi-GPU job is verified without gathering records. */
int o2v_b200_result_hash(o2v_b200_engine *engine, void *cuda_stream, uint64_t *out_hash);

/* Host-buffer convenience: uploads the mesh (and textures), runs, downloads up to out_capacity records into out_voxels
 * (4 u32 each).  *out_count receives the number of voxels produced (may exceed out_capacity: then nothing past the
 * capacity is written and -5 is returned). */
int o2v_b200_voxelize_host(o2v_b200_engine *engine, const o2v_b200_params *params, const o2v_b200_mesh *mesh,
                           const o2v_b200_texture *textures, uint32_t texture_count, uint32_t *out_voxels,
                           uint64_t out_capacity, uint64_t *out_count, o2v_b200_stats *out_stats);

/* Host side of the bitmap download that obj2voxel_voxelize() uses on one GPU (16 bytes per voxel over one PCIe link is
 * the longest leg of a host-to-host job; an all-white result travels as 1 bit per OUTPUT voxel of every touched 64^3
 * chunk instead and the host's threads write the Voxel32 quads the voxel callback receives, src/io.cpp:638-653).
 * bits: `chunks` bitmaps of 4096 64-bit words — word = tile (x | y << 3 | z << 6 in units of 8 voxels) * 8 + layer z,
 * bit = x + 8 y inside the tile; chunk_ids[c] = cx + chunks_per_axis * (cy + chunks_per_axis * (cz - chunk_z0));
 * chunk_counts[c] = set bits of bitmap c.  Writes sum(chunk_counts) quads {x, y, z, 0xFFFFFFFF} to out_quads and returns
 * that number, or UINT64_MAX if a bitmap disagrees with its count.  Pure host code: no device needed. */
uint64_t o2v_b200_expand_bitmaps(const uint64_t *bits, const uint32_t *chunk_ids, const uint32_t *chunk_counts,
                                 uint32_t chunks, uint32_t chunks_per_axis, uint32_t chunk_z0, uint32_t *out_quads);

/* The chunk scan behind the bitmap download as obj2voxel_voxelize() runs it (a host thread per chunk, quads handed on in
 * batches of at most buffer_quads >= 64 from a buffer that stays in the thread's cache; VPCOMPRESSB where the CPU has
 * AVX-512 VBMI2): writes the quads of one chunk bitmap (4096 words, chunk at output origin (cx, cy, cz)) to out_quads (at
 * most out_capacity) and returns their number.  Pure host code. */
uint64_t o2v_b200_scan_chunk_bitmap(const uint64_t *words, uint32_t cx, uint32_t cy, uint32_t cz, uint32_t buffer_quads,
                                    uint32_t *out_quads, uint64_t out_capacity);

/* The positions of an all-white result cross PCIe packed into `bits` = 32 (x | y << 10 | z << 20,
 * output grids up to 1024^3) or 64 (x | y << 21 | z << 42) bits per voxel and the host's threads write the Voxel32 quads
 * {x, y, z, 0xFFFFFFFF} the voxel callback receives.  Pure host code: no device needed. */
void o2v_b200_expand_packed(const void *packed, int32_t bits, uint64_t count, uint32_t *out_quads);

/* How obj2voxel_voxelize() would cut a job into z parts (see obj2voxel_b200_get_stats below): returns the number of parts
 * and writes parts + 1 ascending bounds (at most bounds_capacity) — part k covers sample-space z in
 * [out_bounds[k], out_bounds[k + 1]); inner bounds are multiples of 64 (the reference's chunk rows,
 * src/obj2voxel.cpp:245-252).  requested_parts > 0 forces the number (as O2V_B200_PIPELINE_PARTS does), <= 0 applies the
 * default rule.  Pure host arithmetic: no device needed. */
uint32_t o2v_b200_plan_parts(uint32_t sample_resolution, uint32_t slab_z0, uint32_t slab_z1, uint64_t triangles,
                             int32_t requested_parts, uint32_t *out_bounds, uint32_t bounds_capacity);

/* How obj2voxel_voxelize() cuts a job into Z-slabs for `devices` devices (at most 16): devices + 1 ascending sample-space
 * bounds, inner ones multiples of 64 * supersampling (a row of output chunks has one owner).  row_histogram = NULL: as
 * many chunk rows each as the count allows; else row_histogram[r] = triangles whose z range reaches row r (rows of
 * 64 * supersampling sample layers from z = 0) and the bounds are cut so that every device gets about the same number —
 * what the job runner does with the histogram its devices take of their shares.  Pure host arithmetic: no device needed. */
void o2v_b200_plan_slabs(uint32_t sample_resolution, uint32_t supersampling, uint32_t slab_z0, uint32_t slab_z1,
                         uint32_t devices, const uint64_t *row_histogram, uint32_t rows, uint32_t *out_bounds);

/* ---- additive setters on the reference-compatible instance ------------------------------------------------------- */

/* Bulk input: count triangles as 9 floats each (+ 6 uv floats each and a texture when textured).  The arrays are not
 * copied and must stay valid until obj2voxel_voxelize() returns. */
void obj2voxel_b200_set_input_triangles(obj2voxel_instance *instance, const float *vertices, const float *uvs,
                                        size_t count, obj2voxel_texture *texture);
/* A ready-made obj2voxel_triangle_callback (include/obj2voxel.h:177) over a flat array, for callers that want the
 * reference's one-call-per-triangle ingestion (src/obj2voxel.cpp:585-588) without writing the callback:
 * obj2voxel_set_input_callback(instance, obj2voxel_b200_array_source_next, &source).  Triangles are MATERIALLESS. */
typedef struct o2v_b200_array_source {
    const float *vertices; /* 9 floats per triangle */
    size_t count;
    size_t next; /* start at 0 */
} o2v_b200_array_source;
bool obj2voxel_b200_array_source_next(void *source, obj2voxel_triangle *out_triangle);

/* A ready-made obj2voxel_voxel_callback (include/obj2voxel.h:188) that only counts what it receives:
 * obj2voxel_set_output_callback(instance, obj2voxel_b200_counting_sink_write, &sink).  The callback of one job is never
 * called concurrently (calls come from several host threads, one at a time). */
typedef struct o2v_b200_counting_sink {
    uint64_t voxels; /* start at 0 */
    uint64_t calls;
} o2v_b200_counting_sink;
bool obj2voxel_b200_counting_sink_write(void *sink, uint32_t *voxel_data, size_t voxel_count);

/* The CUDA devices obj2voxel_voxelize() spreads this job over, one Z-slab of whole chunk rows each (default: the
 * environment's O2V_B200_DEVICES = "all" | count | "0,2,5", else O2V_B200_DEVICE, else device 0).  The reference spreads a
 * job over worker threads that pull 64^3 chunks (src/obj2voxel.cpp:957-996, CLI -j, src/main.cpp:155-158); here a device
 * owns a slab of chunk rows.  On the occupancy-only path each device uploads 1/N of the triangles and the devices
 * exchange them by slab over peer memory; the sink receives every device's records, one writer at a time. */
void obj2voxel_b200_set_devices(obj2voxel_instance *instance, const int32_t *devices, uint32_t count);
/* Restrict the job to a Z-slab [z0, z1) of the sample grid (multiples of 8).  (0, 0) restores the whole grid, any other
 * z0 == z1 is an empty slab (no voxels): a rank that owns no rows skips the job instead of passing (0, 0). */
void obj2voxel_b200_set_slab(obj2voxel_instance *instance, uint32_t z0, uint32_t z1);
/* Statistics of the last obj2voxel_voxelize() on this instance.  A big job runs as z parts (one per 2^20 triangles, at most four)
 * so that the download of one part overlaps the kernels of the next (O2V_B200_PIPELINE_PARTS overrides the number): the
 * counts and the device times are then sums over the parts (dropped_triangles: the largest of the parts), and the sink /
 * voxel callback receives the parts one after the other. */
void obj2voxel_b200_get_stats(obj2voxel_instance *instance, o2v_b200_stats *out_stats);

#ifdef __cplusplus
}
#endif

#endif /* OBJ2VOXEL_B200_HEADER */
