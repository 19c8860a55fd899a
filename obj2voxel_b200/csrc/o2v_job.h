// One voxelization job from host buffers to a voxel sink, on one or several GPUs of one process: what
// obj2voxel_voxelize() does once the triangle stream has been gathered (reference: voxelize_specialized,
// src/obj2voxel.cpp:467-520 — sort triangles into chunks, voxelize the chunks on the workers, write each chunk's voxels to
// the sink under a mutex, :296-312).
//
//   host triangles ──H2D──▶ device share ──(several devices: bin by Z-slab, peer stores over NVLink)──▶ slab triangles
//     ──kernels per z part──▶ records | occupancy bitmaps ──D2H (under the next part's kernels)──▶ pinned host memory
//     ──(bitmaps: host threads expand them into Voxel32 quads)──▶ sink / voxel callback
#ifndef O2V_JOB_H
#define O2V_JOB_H

#include <functional>
#include <string>
#include <vector>

#include "o2v_engine.h"
#include "o2v_io.h"
#include "obj2voxel_b200.h"

namespace o2v {

enum class DownloadMode { AUTO, RECORDS, BITMAP, PACKED };

struct JobOptions {
    EngineParams params;            // slabZ0 / slabZ1 restrict the whole job
    std::vector<int> devices;       // CUDA device ordinals; one Z-slab each
    int parts = 0;                  // z parts per device, 0 = default rule (planJobParts)
    DownloadMode download = DownloadMode::AUTO;
};

struct JobTimings {
    double msUpload = 0, msExchange = 0, msRun = 0, msKernels = 0;
    double msWaitCopy = 0, msExpandHost = 0, msSink = 0, msVoxelizeCalls = 0;  // device 0's host side, summed over its parts
    std::vector<uint32_t> slabBounds;  // devices + 1 sample-space z bounds of the devices' slabs
    uint32_t parts = 0, devices = 0;
    bool bitmapDownload = false, peerExchange = false, stagedUpload = false;
    bool streamedUpload = false;  // the mesh was voxelized piece by piece while it crossed PCIe (msUpload is not separate)
};

/// Process-wide engine of `device`, created on first use; nullptr + *error without a usable CUDA device.
Engine *sharedEngine(int device, std::string *error);
/// Devices a job uses by default: O2V_B200_DEVICES = "all" | count | comma-separated ordinals, else O2V_B200_DEVICE, else 0.
std::vector<int> defaultJobDevices();

/// How a job's z range is cut into parts (see o2v_b200_plan_parts).
constexpr uint32_t kMaxJobParts = 128;
uint32_t planJobParts(uint32_t sampleRes, uint32_t slabZ0, uint32_t slabZ1, unsigned long long triangles, int requested,
                      uint32_t *bounds);
void accumulateStats(RunStats &total, const RunStats &part);
/// The Z-slabs of a job over `devices` devices (see o2v_b200_plan_slabs): equal chunk rows, then — with a histogram of
/// triangles per row of 64 * supersampling sample layers — cut so that every device gets about the same number.
void planJobSlabs(uint32_t sampleRes, uint32_t supersampling, uint32_t jobZ0, uint32_t jobZ1, uint32_t devices,
                  const unsigned long long *histogram, uint32_t rows, uint32_t *bounds);

/// Runs the job.  mesh / textures: HOST pointers (o2v_b200_mesh layout).  Returns OBJ2VOXEL_ERR_OK,
/// OBJ2VOXEL_ERR_DEVICE (message logged) or OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE.
obj2voxel_error_t runDeviceJob(const o2v_b200_mesh &mesh, const std::vector<o2v_b200_texture> &textures,
                               const JobOptions &options, VoxelSink &sink, RunStats *stats, JobTimings *timings);

/// Expands downloaded occupancy bitmaps (BitmapResult layout, host copies) into Voxel32 quads on the host threads:
/// records[offset of chunk ...] receive {x, y, z, 0xFFFFFFFF}; returns the number written (= sum of chunkCounts).
unsigned long long expandBitmapsOnHost(const unsigned long long *bits, const uint32_t *chunkIds,
                                       const uint32_t *chunkCounts, uint32_t chunks, uint32_t chunksPerAxis,
                                       uint32_t chunkZ0, uint32_t *records);

/// One chunk's bitmap (kChunkWords words, chunk at OUTPUT origin (cx, cy, cz)) into quads, through `buffer` (room for
/// bufferQuads >= 64 quads): flush(buffer, n) is called whenever the buffer is full and once at the end.  Returns the
/// number of quads, ~0 if flush failed.  Uses VPCOMPRESSB where the CPU has AVX-512 VBMI2 (hostHasFastBitScan).
unsigned long long scanChunkBitmap(const unsigned long long *words, uint32_t cx, uint32_t cy, uint32_t cz, uint32_t *buffer,
                                   uint32_t bufferQuads, const std::function<bool(uint32_t *, size_t)> &flush);
bool hostHasFastBitScan();

/// Expands `count` packed positions (Engine::packedBits: 32 = x | y << 10 | z << 20, 64 = x | y << 21 | z << 42) into
/// Voxel32 quads {x, y, z, 0xFFFFFFFF} on the host threads.
void expandPackedOnHost(const void *packed, int bits, unsigned long long count, uint32_t *records);

}  // namespace o2v

#endif  // O2V_JOB_H
