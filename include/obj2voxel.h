/* obj2voxel C API, B200 edition.
 *
 * This header declares, symbol for symbol and type for type, the 35 functions of the reference's
 * include/obj2voxel.h (Eisenwave/obj2voxel @ 9fb8ae2); a program compiled against the reference header links and runs
 * against libobj2voxel_b200.so unchanged.  Behind obj2voxel_voxelize() the per-triangle voxelization
 * (reference src/voxelization.cpp, src/triangle.hpp, the driver loop of src/obj2voxel.cpp:467-520) runs as hand-written
 * sm_100a CUDA kernels; there is no CPU implementation in this library.
 *
 * Each declaration names the reference header line it replaces ("ref :NNN").  Additive entry points (bulk/device input,
 * Z-slab selection, statistics) live in obj2voxel_b200.h and never change the behaviour of the functions below.
 */
#ifndef OBJ2VOXEL_HEADER
#define OBJ2VOXEL_HEADER

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar typedefs and opaque handles (ref :14-28) ---------------------------------------------------------- */

typedef unsigned char obj2voxel_enum_t;
typedef unsigned char obj2voxel_byte_t;
typedef unsigned char obj2voxel_error_t;

typedef struct obj2voxel_instance obj2voxel_instance; /* one voxelization job; single use */
typedef struct obj2voxel_texture obj2voxel_texture;   /* caller-owned image, must outlive the job */
typedef struct obj2voxel_triangle obj2voxel_triangle; /* only ever filled through obj2voxel_set_triangle_*() */

/* ---- callbacks (ref :32-38) ------------------------------------------------------------------------------------ */

/* Produces the next input triangle into out_triangle; returns false at the end of the stream. */
typedef bool(obj2voxel_triangle_callback)(void *callback_data, obj2voxel_triangle *out_triangle);
/* Receives voxel_count records of four host-endian u32 (x, y, z, argb); the buffer is only valid during the call and
 * the batching/order of calls is unspecified.  Returning false aborts the job with an I/O error. */
typedef bool(obj2voxel_voxel_callback)(void *callback_data, uint32_t *voxel_data, size_t voxel_count);
/* Receives a log line; returning false asks for the default stdout logging of that line. */
typedef bool(obj2voxel_log_callback)(void *callback_data, const char *msg, obj2voxel_enum_t level);

/* ---- enum constants (ref :43-79) ------------------------------------------------------------------------------- */

static const obj2voxel_enum_t OBJ2VOXEL_MAX_STRATEGY = 0;   /* colour of the heaviest triangle wins */
static const obj2voxel_enum_t OBJ2VOXEL_BLEND_STRATEGY = 1; /* weighted average of all triangle colours */

static const obj2voxel_enum_t OBJ2VOXEL_UV_CLAMP = 0;
static const obj2voxel_enum_t OBJ2VOXEL_UV_WRAP = 1;

static const obj2voxel_enum_t OBJ2VOXEL_LOG_LEVEL_SILENT = 0;
static const obj2voxel_enum_t OBJ2VOXEL_LOG_LEVEL_ERROR = 1;
static const obj2voxel_enum_t OBJ2VOXEL_LOG_LEVEL_WARNING = 2;
static const obj2voxel_enum_t OBJ2VOXEL_LOG_LEVEL_INFO = 3;
static const obj2voxel_enum_t OBJ2VOXEL_LOG_LEVEL_DEBUG = 4;

static const obj2voxel_error_t OBJ2VOXEL_ERR_OK = 0;
static const obj2voxel_error_t OBJ2VOXEL_ERR_NO_INPUT = 1;
static const obj2voxel_error_t OBJ2VOXEL_ERR_NO_OUTPUT = 2;
static const obj2voxel_error_t OBJ2VOXEL_ERR_NO_RESOLUTION = 3;
static const obj2voxel_error_t OBJ2VOXEL_ERR_IO_ERROR_ON_OPEN_INPUT_FILE = 4;
static const obj2voxel_error_t OBJ2VOXEL_ERR_IO_ERROR_ON_OPEN_OUTPUT_FILE = 5;
static const obj2voxel_error_t OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE = 6;
static const obj2voxel_error_t OBJ2VOXEL_ERR_DOUBLE_VOXELIZATION = 7;
/* Additive (not in the reference): no usable CUDA device / device failure.  This library never falls back to the CPU. */
static const obj2voxel_error_t OBJ2VOXEL_ERR_DEVICE = 8;

/* ---- instance life cycle and the job itself -------------------------------------------------------------------- */

obj2voxel_instance *obj2voxel_alloc(void);          /* ref :89 */
void obj2voxel_free(obj2voxel_instance *instance);  /* ref :95 */

/* Runs the job on the calling thread (kernels on the GPU selected by O2V_B200_DEVICE, default 0) and returns one of
 * OBJ2VOXEL_ERR_*.  A second call on the same instance returns OBJ2VOXEL_ERR_DOUBLE_VOXELIZATION.  ref :406 */
obj2voxel_error_t obj2voxel_voxelize(obj2voxel_instance *instance);

/* ---- logging: process-global, as in the reference (ref :105-120) ------------------------------------------------ */

void obj2voxel_set_log_level(obj2voxel_enum_t level);
void obj2voxel_set_log_callback(obj2voxel_log_callback *callback, void *callback_data); /* NULL restores stdout */
obj2voxel_enum_t obj2voxel_get_log_level(void);

/* ---- job configuration (ref :130-263) --------------------------------------------------------------------------- */

void obj2voxel_set_resolution(obj2voxel_instance *instance, uint32_t resolution);          /* ref :130, > 0 */
void obj2voxel_set_supersampling(obj2voxel_instance *instance, uint32_t level);            /* ref :138, 1 or 2 */
void obj2voxel_set_color_strategy(obj2voxel_instance *instance, obj2voxel_enum_t strategy); /* ref :146 */
void obj2voxel_set_texture(obj2voxel_instance *instance, obj2voxel_texture *texture);      /* ref :157, default texture */

/* `type` is a file extension without dot, or NULL to detect it from the path.  The path string is not copied. */
void obj2voxel_set_input_file(obj2voxel_instance *instance, const char *file, const char *type);  /* ref :167 */
void obj2voxel_set_input_callback(obj2voxel_instance *instance,                                   /* ref :177 */
                                  obj2voxel_triangle_callback *callback,
                                  void *callback_data);

void obj2voxel_set_output_file(obj2voxel_instance *instance, const char *file, const char *type);  /* ref :189 */
void obj2voxel_set_output_memory(obj2voxel_instance *instance, const char *type);                  /* ref :198 */
void obj2voxel_set_output_callback(obj2voxel_instance *instance,                                   /* ref :207 */
                                   obj2voxel_voxel_callback *callback,
                                   void *callback_data);

/* Accepted for compatibility: the GPU path does not need host workers, but the worker entry points below keep the
 * reference's blocking / counting behaviour.  ref :219 */
void obj2voxel_set_parallel(obj2voxel_instance *instance, bool enabled);

/* Row-major 3x3 integer axis permutation / mirroring applied in the [-1,1] cube.  ref :228 */
void obj2voxel_set_unit_transform(obj2voxel_instance *instance, const int transform[9]);
/* {min x,y,z, max x,y,z}; skips the bounds pass.  All values finite, min <= max.  ref :237 */
void obj2voxel_set_mesh_boundaries(obj2voxel_instance *instance, const float bounds[6]);

uint32_t obj2voxel_get_resolution(obj2voxel_instance *instance); /* ref :245 */
uint32_t obj2voxel_get_chunk_size(obj2voxel_instance *instance); /* ref :254, always 64 */
/* Bytes produced by set_output_memory(); owned by the instance, valid until obj2voxel_free().  ref :263 */
const obj2voxel_byte_t *obj2voxel_get_output_memory(obj2voxel_instance *instance, size_t *out_size);

/* ---- triangle construction inside a triangle callback (ref :272-292) --------------------------------------------- */

void obj2voxel_set_triangle_basic(obj2voxel_triangle *triangle, const float vertices[9]);
/* Bug-compatible with the reference (src/obj2voxel.cpp:828-837): the colour is stored but the triangle voxelizes white. */
void obj2voxel_set_triangle_colored(obj2voxel_triangle *triangle, const float vertices[9], const float color[3]);
void obj2voxel_set_triangle_textured(obj2voxel_triangle *triangle,
                                     const float vertices[9],
                                     const float textures[6],
                                     obj2voxel_texture *texture);

/* ---- textures (ref :300-370) ---------------------------------------------------------------------------------- */

obj2voxel_texture *obj2voxel_texture_alloc(void);
void obj2voxel_texture_free(obj2voxel_texture *texture);
bool obj2voxel_texture_load_from_file(obj2voxel_texture *texture, const char *file, const char *type);
bool obj2voxel_texture_load_from_memory(obj2voxel_texture *texture,
                                        const obj2voxel_byte_t *data,
                                        size_t size,
                                        const char *type);
/* Copies width*height*channels bytes; channels 3 (RGB) or 4 (decoded like the reference's ARGB32).  ref :341 */
bool obj2voxel_texture_load_pixels(
    obj2voxel_texture *texture, const obj2voxel_byte_t *pixels, size_t width, size_t height, size_t channels);
/* The misspelling is part of the reference ABI (ref :350). */
void obj2voxel_teture_set_uv_mode(obj2voxel_texture *texture, obj2voxel_enum_t mode);
void obj2voxel_texture_get_meta(obj2voxel_texture *texture,
                                size_t *out_width,
                                size_t *out_height,
                                size_t *out_channels);
void obj2voxel_texture_get_pixels(obj2voxel_texture *texture, obj2voxel_byte_t *out_pixels);

/* ---- worker threads (ref :380-396) ------------------------------------------------------------------------------ */

/* Registers the calling thread as a worker and blocks until obj2voxel_stop_workers(). */
void obj2voxel_run_worker(obj2voxel_instance *instance);
void obj2voxel_stop_workers(obj2voxel_instance *instance);
uint32_t obj2voxel_get_worker_count(obj2voxel_instance *instance);

#ifdef __cplusplus
}
#endif

#endif /* OBJ2VOXEL_HEADER */
