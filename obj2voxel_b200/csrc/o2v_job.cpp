// See o2v_job.h.  Host orchestration only: every voxel is decided by the CUDA kernels behind o2v::Engine.
#include "o2v_job.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "o2v_pool.h"

namespace o2v {

namespace {

struct EngineDeleter {
    void operator()(Engine *e) const { delete e; }
};

std::mutex gEngineMutex;
std::unordered_map<int, std::unique_ptr<Engine, EngineDeleter>> gEngines;
std::mutex gJobMutex;  // one job at a time per process: the engines and their buffers are shared

double msSince(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

/// Host threads of the job runner: O2V_B200_HOST_THREADS or the hardware concurrency (the calling thread is one of them).
HostPool &hostPool()
{
    static HostPool *pool = [] {
        unsigned threads = std::thread::hardware_concurrency();
        if (const char *env = getenv("O2V_B200_HOST_THREADS")) {
            threads = (unsigned) std::max(1, atoi(env));
        }
        threads = std::min(std::max(threads, 1u), 64u);
        return new HostPool(threads - 1);  // never destroyed: worker threads must not outlive a static's destructor
    }();
    return *pool;
}

bool isPinnedHost(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

constexpr size_t kStageChunk = 4u << 20;

/// Host -> device copy of a caller's array.  Pinned memory goes to the copy engine as it is.  Pageable memory would make
/// the driver stage it through its own small buffer on one thread (~10 GB/s): instead the host threads copy 4 MiB
/// pieces into their own pinned buffers and send each piece on, so that the link — not a memcpy — sets the pace.
bool uploadArray(int device, void *dst, const void *src, size_t bytes, cudaStream_t stream, bool *staged)
{
    if (bytes == 0) {
        return true;
    }
    if (isPinnedHost(src) || bytes < 2 * kStageChunk) {
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream) == cudaSuccess;
    }
    *staged = true;
    std::atomic<bool> ok{true};
    const size_t chunks = (bytes + kStageChunk - 1) / kStageChunk;
    hostPool().parallelFor(chunks, [&](size_t i) {
        thread_local void *pinned = nullptr;
        if (pinned == nullptr && cudaHostAlloc(&pinned, kStageChunk, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            pinned = nullptr;
        }
        const size_t offset = i * kStageChunk, n = std::min(kStageChunk, bytes - offset);
        const char *from = static_cast<const char *>(src) + offset;
        char *to = static_cast<char *>(dst) + offset;
        bool good = cudaSetDevice(device) == cudaSuccess;
        if (pinned != nullptr) {
            memcpy(pinned, from, n);
            good = good && cudaMemcpyAsync(to, pinned, n, cudaMemcpyHostToDevice, cudaStreamPerThread) == cudaSuccess &&
                   cudaStreamSynchronize(cudaStreamPerThread) == cudaSuccess;
        }
        else {  // no pinned memory left: the driver's own staging still works
            good = good && cudaMemcpy(to, from, n, cudaMemcpyHostToDevice) == cudaSuccess;
        }
        if (!good) {
            ok = false;
        }
    });
    return ok;
}

/// Device copies of (a share of) a host mesh + textures, kept per device across jobs (grow-only).
struct UploadedMesh {
    DeviceBuffer verts, uvs, types, colors, textureIds;
    std::vector<std::unique_ptr<DeviceBuffer>> texturePixels;
    std::vector<TextureView> textureViews;
    MeshView view{};

    bool upload(int device, const o2v_b200_mesh &mesh, size_t first, size_t count,
                const std::vector<o2v_b200_texture> &textures, cudaStream_t stream, bool *staged, std::string *error)
    {
        auto copy = [&](DeviceBuffer &dst, const void *src, size_t stride) -> bool {
            if (src == nullptr || count == 0) {
                return true;
            }
            if (!dst.ensure(count * stride)) {
                *error = "device allocation failed (mesh upload)";
                return false;
            }
            if (!uploadArray(device, dst.as<void>(), static_cast<const char *>(src) + first * stride, count * stride,
                             stream, staged)) {
                *error = std::string("mesh upload failed: ") + cudaGetErrorString(cudaGetLastError());
                return false;
            }
            return true;
        };
        if (!copy(verts, mesh.verts, 9 * sizeof(float)) || !copy(uvs, mesh.uvs, 6 * sizeof(float)) ||
            !copy(types, mesh.types, 1) || !copy(colors, mesh.colors, 3 * sizeof(float)) ||
            !copy(textureIds, mesh.texture_ids, sizeof(uint32_t))) {
            return false;
        }
        view.verts = mesh.verts != nullptr ? verts.as<float>() : nullptr;
        view.uvs = mesh.uvs != nullptr ? uvs.as<float>() : nullptr;
        view.types = mesh.types != nullptr ? types.as<uint8_t>() : nullptr;
        view.colors = mesh.colors != nullptr ? colors.as<float>() : nullptr;
        view.textureIds = mesh.texture_ids != nullptr ? textureIds.as<uint32_t>() : nullptr;
        view.count = count;
        texturePixels.clear();
        textureViews.clear();
        for (const o2v_b200_texture &t : textures) {
            texturePixels.emplace_back(new DeviceBuffer());
            const size_t bytes = (size_t) t.width * t.height * t.channels;
            if (!texturePixels.back()->ensure(std::max<size_t>(bytes, 1)) ||
                cudaMemcpyAsync(texturePixels.back()->as<void>(), t.pixels, bytes, cudaMemcpyHostToDevice, stream) !=
                    cudaSuccess) {
                *error = "texture upload failed";
                return false;
            }
            textureViews.push_back(TextureView{texturePixels.back()->as<uint8_t>(), t.width, t.height, t.channels, t.wrap});
        }
        return true;
    }
};

/// What a device keeps between jobs besides its engine.
struct DeviceState {
    UploadedMesh upload;
    // a mesh that goes up in pieces (runDeviceJob): the stream and the events are made once — creating and destroying
    // them per job, and asking for the free memory, showed up as occasional 40 ms jobs
    cudaStream_t uploadStream = nullptr;
    std::vector<cudaEvent_t> pieceEvents;
    size_t approvedBytes = 0;  // largest bitmaps + mesh footprint the free-memory check has passed

    uint32_t *records[2] = {nullptr, nullptr};  // host buffers the bitmaps are expanded into (plain memory: CPUs write it)
    size_t recordBytes[2] = {0, 0};

    uint32_t *recordBuffer(int slot, size_t bytes)
    {
        if (bytes > recordBytes[slot]) {
            free(records[slot]);
            const size_t wanted = bytes + bytes / 8;
            records[slot] = static_cast<uint32_t *>(aligned_alloc(64, (wanted + 63) / 64 * 64));
            recordBytes[slot] = records[slot] != nullptr ? wanted : 0;
        }
        return records[slot];
    }
};

DeviceState &deviceState(int device)
{
    static std::mutex mutex;
    static std::unordered_map<int, std::unique_ptr<DeviceState>> states;
    std::lock_guard<std::mutex> lock{mutex};
    std::unique_ptr<DeviceState> &slot = states[device];
    if (slot == nullptr) {
        slot.reset(new DeviceState());
    }
    return *slot;
}

/// Peer access between every pair of the job's devices (once per process and pair); false if some pair cannot.
bool enablePeerAccess(const std::vector<int> &devices)
{
    static std::mutex mutex;
    static std::unordered_map<long long, bool> enabled;
    std::lock_guard<std::mutex> lock{mutex};
    bool all = true;
    for (int a : devices) {
        for (int b : devices) {
            if (a == b) {
                continue;
            }
            const long long key = ((long long) a << 32) | (unsigned) b;
            auto found = enabled.find(key);
            if (found == enabled.end()) {
                int can = 0;
                bool ok = cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can != 0 && cudaSetDevice(a) == cudaSuccess;
                if (ok) {
                    const cudaError_t err = cudaDeviceEnablePeerAccess(b, 0);
                    ok = err == cudaSuccess || err == cudaErrorPeerAccessAlreadyEnabled;
                }
                cudaGetLastError();
                found = enabled.emplace(key, ok).first;
            }
            all = all && found->second;
        }
    }
    return all;
}

template <bool kStreaming>
void unpackRange(const void *packed, int bits, unsigned long long first, unsigned long long last, uint32_t *out);

/// O2V_B200_EXPAND_BATCH=0: expand a whole part into a DRAM-sized buffer with streaming stores, then call the sink (the
/// alternative to SharedSink::writePacked, kept for measurements).
bool wholePartExpansion()
{
    const char *env = getenv("O2V_B200_EXPAND_BATCH");
    return env != nullptr && atoi(env) == 0;
}

/// Sink shared by the device threads of a job: writes are serialised (the reference writes each chunk's voxels under a
/// mutex from whichever worker voxelized it, src/obj2voxel.cpp:296-312).
struct SharedSink {
    VoxelSink &sink;
    std::mutex mutex;
    std::atomic<bool> failed{false};

    /// Packed positions -> quads -> sink.  Every host thread unpacks 64 Ki voxels at a time into a buffer of its own
    /// that never leaves its cache (1 MiB, written again and again by the same core) and hands it to the sink under the
    /// lock: the quads are 16 bytes per voxel that nobody needs in DRAM.  The sink is called from whichever thread has a
    /// batch ready, one at a time — as the reference calls it from whichever worker finished a chunk
    /// (src/obj2voxel.cpp:296-312).
    void writePacked(const void *packed, int bits, unsigned long long count)
    {
        constexpr unsigned long long kBatch = 1ull << 16;
        hostPool().parallelFor((size_t) ((count + kBatch - 1) / kBatch), [&](size_t task) {
            if (failed) {
                return;
            }
            thread_local uint32_t *mine = nullptr;
            if (mine == nullptr) {
                mine = static_cast<uint32_t *>(aligned_alloc(4096, kBatch * 16));  // (kept for the thread's lifetime)
                if (mine == nullptr) {
                    failed = true;
                    return;
                }
            }
            const unsigned long long first = task * kBatch, last = std::min(count, first + kBatch);
            unpackRange<false>(packed, bits, first, last, mine);
            std::lock_guard<std::mutex> lock{mutex};
            if (!failed && !sink.write(mine, (size_t) (last - first))) {
                failed = true;
            }
        });
    }

    /// Occupancy bitmaps -> quads -> sink, the same way: a host thread takes a chunk, scans its 4096 words into its own
    /// cache-resident buffer (scanChunkBitmap: VPCOMPRESSB on CPUs that have it) and hands the buffer to the sink whenever
    /// it is full.  Returns the number of voxels delivered.
    unsigned long long writeBitmaps(const unsigned long long *bits, const uint32_t *chunkIds, uint32_t chunks,
                                    uint32_t chunksPerAxis, uint32_t chunkZ0)
    {
        constexpr uint32_t kBatch = 1u << 16;
        std::atomic<unsigned long long> delivered{0};
        hostPool().parallelFor(chunks, [&](size_t c) {
            if (failed) {
                return;
            }
            thread_local uint32_t *mine = nullptr;
            if (mine == nullptr) {
                mine = static_cast<uint32_t *>(aligned_alloc(4096, (size_t) kBatch * 16));
                if (mine == nullptr) {
                    failed = true;
                    return;
                }
            }
            const uint32_t chunk = chunkIds[c], C = chunksPerAxis;
            const unsigned long long n = scanChunkBitmap(
                bits + c * kChunkWords, (chunk % C) * kChunkEdge, ((chunk / C) % C) * kChunkEdge,
                (chunk / (C * C) + chunkZ0) * kChunkEdge, mine, kBatch, [&](uint32_t *quads, size_t count) {
                    std::lock_guard<std::mutex> lock{mutex};
                    return !failed && sink.write(quads, count);
                });
            if (n == ~0ull) {
                failed = true;
            }
            else {
                delivered += n;
            }
        });
        return delivered;
    }

    void write(uint32_t *quads, unsigned long long count)
    {
        const size_t batch = 1u << 21;  // records per sink call (32 MiB)
        std::lock_guard<std::mutex> lock{mutex};
        for (unsigned long long done = 0; done < count && !failed; done += batch) {
            if (!sink.write(quads + done * 4, (size_t) std::min<unsigned long long>(batch, count - done))) {
                failed = true;
            }
        }
    }
};

/// The part pipeline of one device: the slab [params.slabZ0, params.slabZ1) of `mesh` (device-resident) in z parts, the
/// download of one part under the kernels of the next, records to the sink.
struct SlabRun {
    Engine *engine = nullptr;
    DeviceState *state = nullptr;
    MeshView mesh{};
    const TextureView *textures = nullptr;
    uint32_t textureCount = 0;
    EngineParams params;
    int requestedParts = 0;
    bool wantBitmap = false, wantPacked = false;
    cudaStream_t stream = nullptr;
    SharedSink *sink = nullptr;
    // A mesh that is still crossing PCIe (occupancy-only path): the parts are pieces of the triangle array instead of z
    // ranges — piece k = triangles [pieceFirst[k], pieceFirst[k + 1]), on the device once pieceReady[k] has happened —
    // and every piece delivers the voxels no earlier piece has (EngineParams::accumulate).
    std::vector<size_t> pieceFirst;
    std::vector<cudaEvent_t> pieceReady;
    // results
    RunStats stats;
    bool any = false;
    double msKernels = 0;
    uint32_t parts = 0;
    bool usedBitmap = false;
    unsigned long long downloadBytes = 0;
    double msWaitCopy = 0, msExpandHost = 0, msSink = 0, msVoxelizeCalls = 0;  // where the host side spent its time
    std::string error;
    std::atomic<bool> deviceFailed{false};
    std::mutex errorMutex;

    bool run();
    void failWith(const std::string &message)
    {
        std::lock_guard<std::mutex> lock{errorMutex};
        if (error.empty()) {
            error = message;
        }
        deviceFailed = true;
    }
};

bool SlabRun::run()
{
    const int device = engine->device();
    if (cudaSetDevice(device) != cudaSuccess) {
        error = "cudaSetDevice failed";
        return false;
    }
    cudaStream_t copyStream = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr};
    if (cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&copied[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&copied[1], cudaEventDisableTiming) != cudaSuccess) {
        error = std::string("cudaStreamCreate failed: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    struct Guard {
        cudaStream_t s;
        cudaEvent_t *e;
        ~Guard()
        {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
            cudaEventDestroy(e[0]);
            cudaEventDestroy(e[1]);
        }
    } guard{copyStream, copied};

    uint32_t partBounds[kMaxJobParts + 1];
    const bool pieces = !pieceFirst.empty();
    parts = pieces ? (uint32_t) pieceFirst.size() - 1
                   : planJobParts(params.resolution * params.supersampling, params.slabZ0, params.slabZ1, mesh.count,
                                  requestedParts, partBounds);
    const size_t batch = 1u << 21;

    // What is on its way to the sink: a part whose download sits in pinned buffer `slot` — records as they are, or
    // bitmaps that the host threads expand into the record buffer of the same slot first.  A delivery thread takes the
    // parts in order, so that this thread can launch the next part's kernels while the host threads expand the last
    // one's bitmaps; at most two parts are in flight (two device buffers, two pinned buffers).
    struct Pending {
        int slot = 0;
        unsigned long long count = 0;
        bool bitmap = false;
        int packedBits = 0;
        uint32_t chunks = 0;
        BitmapResult geometry;
        unsigned char *host = nullptr;
    };
    auto deliver = [&](const Pending &p) {
        auto t0 = std::chrono::steady_clock::now();
        if (cudaSetDevice(device) != cudaSuccess || cudaEventSynchronize(copied[p.slot]) != cudaSuccess) {
            failWith(std::string("voxel download failed: ") + cudaGetErrorString(cudaGetLastError()));
            return;
        }
        msWaitCopy += msSince(t0);
        if (deviceFailed || sink->failed) {
            return;
        }
        if (!p.bitmap && p.packedBits == 0) {
            t0 = std::chrono::steady_clock::now();
            sink->write(reinterpret_cast<uint32_t *>(p.host), p.count);
            msSink += msSince(t0);
            return;
        }
        t0 = std::chrono::steady_clock::now();
        if (p.packedBits != 0 && !wholePartExpansion()) {
            sink->writePacked(p.host, p.packedBits, p.count);
            msExpandHost += msSince(t0);
            return;
        }
        if (p.packedBits != 0) {
            uint32_t *unpacked = state->recordBuffer(p.slot, (size_t) p.count * 16);
            if (unpacked == nullptr) {
                failWith("host allocation failed (voxel records)");
                return;
            }
            expandPackedOnHost(p.host, p.packedBits, p.count, unpacked);
            msExpandHost += msSince(t0);
            t0 = std::chrono::steady_clock::now();
            sink->write(unpacked, p.count);
            msSink += msSince(t0);
            return;
        }
        const size_t bitsBytes = (size_t) p.chunks * kChunkWords * 8;
        const uint32_t *ids = reinterpret_cast<const uint32_t *>(p.host + bitsBytes);
        unsigned long long written = 0;
        if (!wholePartExpansion()) {
            written = sink->writeBitmaps(reinterpret_cast<const unsigned long long *>(p.host), ids, p.chunks,
                                         p.geometry.chunksPerAxis, p.geometry.chunkZ0);
            msExpandHost += msSince(t0);
        }
        else {
            uint32_t *records = state->recordBuffer(p.slot, (size_t) p.count * 16);
            if (records == nullptr) {
                failWith("host allocation failed (voxel records)");
                return;
            }
            written = expandBitmapsOnHost(reinterpret_cast<const unsigned long long *>(p.host), ids, ids + p.chunks,
                                          p.chunks, p.geometry.chunksPerAxis, p.geometry.chunkZ0, records);
            msExpandHost += msSince(t0);
            if (written == p.count) {
                t0 = std::chrono::steady_clock::now();
                sink->write(records, p.count);
                msSink += msSince(t0);
            }
        }
        if (written != p.count && !sink->failed) {
            failWith("bitmap expansion produced a different voxel count than the device");
        }
    };
    struct Delivery {
        std::mutex mutex;
        std::condition_variable wake;
        std::vector<Pending> queue;
        unsigned inFlight = 0;
        bool stop = false;
    } delivery;
    std::thread deliveryThread([&] {
        for (;;) {
            Pending p;
            {
                std::unique_lock<std::mutex> lock{delivery.mutex};
                delivery.wake.wait(lock, [&] { return delivery.stop || !delivery.queue.empty(); });
                if (delivery.queue.empty()) {
                    return;
                }
                p = delivery.queue.front();
                delivery.queue.erase(delivery.queue.begin());
            }
            deliver(p);
            {
                std::lock_guard<std::mutex> lock{delivery.mutex};
                --delivery.inFlight;
            }
            delivery.wake.notify_all();
        }
    });
    auto waitInFlightBelow = [&](unsigned limit) {  // limit 1: everything delivered
        std::unique_lock<std::mutex> lock{delivery.mutex};
        delivery.wake.wait(lock, [&] { return delivery.inFlight < limit; });
    };
    auto enqueue = [&](const Pending &p) {
        {
            std::lock_guard<std::mutex> lock{delivery.mutex};
            delivery.queue.push_back(p);
            ++delivery.inFlight;
        }
        delivery.wake.notify_all();
    };
    struct DeliveryGuard {
        Delivery &d;
        std::thread &t;
        ~DeliveryGuard()
        {
            {
                std::lock_guard<std::mutex> lock{d.mutex};
                d.stop = true;
            }
            d.wake.notify_all();
            t.join();
        }
    } deliveryGuard{delivery, deliveryThread};
    // Fallback for records that do not fit a pinned buffer (or a job in one part): two 32 MiB staging buffers, the copy of
    // batch k + 1 under the sink call of batch k.
    auto streamOut = [&](const unsigned char *deviceRecords, unsigned long long total) {
        uint32_t *staging[2] = {static_cast<uint32_t *>(engine->pinnedStaging(0, batch * 16)),
                                static_cast<uint32_t *>(engine->pinnedStaging(1, batch * 16))};
        std::vector<uint32_t> pageable;
        if (staging[0] == nullptr || staging[1] == nullptr) {  // pinned memory exhausted: plain host memory still works
            pageable.resize(batch * 4 * 2);
            staging[0] = pageable.data();
            staging[1] = pageable.data() + batch * 4;
        }
        auto startCopy = [&](unsigned long long done, int slot) -> bool {
            const size_t count = (size_t) std::min<unsigned long long>(batch, total - done);
            return cudaMemcpyAsync(staging[slot], deviceRecords + done * 16, count * 16, cudaMemcpyDeviceToHost,
                                   copyStream) == cudaSuccess &&
                   cudaEventRecord(copied[slot], copyStream) == cudaSuccess;
        };
        int slot = 0;
        bool ok = startCopy(0, 0);
        for (unsigned long long done = 0; done < total && ok && !sink->failed; done += batch, slot ^= 1) {
            const size_t count = (size_t) std::min<unsigned long long>(batch, total - done);
            if (done + batch < total) {
                ok = startCopy(done + batch, slot ^ 1);
            }
            ok = ok && cudaEventSynchronize(copied[slot]) == cudaSuccess;
            if (ok) {
                sink->write(staging[slot], count);
            }
        }
        cudaStreamSynchronize(copyStream);
        if (!ok) {
            failWith(std::string("voxel download failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
    };

    uint32_t launched = 0;  // parts handed to the delivery thread so far (slot = launched & 1)
    for (uint32_t k = 0; k < parts && !sink->failed && !deviceFailed; ++k) {
        EngineParams partParams = params;
        MeshView partMesh = mesh;
        if (pieces) {
            partMesh.verts = mesh.verts + pieceFirst[k] * 9;
            partMesh.count = pieceFirst[k + 1] - pieceFirst[k];
            partParams.accumulate = k == 0 ? 1 : 2;
            if (cudaStreamWaitEvent(stream, pieceReady[k], 0) != cudaSuccess) {
                failWith(std::string("waiting for the upload failed: ") + cudaGetErrorString(cudaGetLastError()));
                break;
            }
        }
        else {
            partParams.slabZ0 = partBounds[k];
            partParams.slabZ1 = partBounds[k + 1];
            if (partParams.slabZ0 >= partParams.slabZ1) {
                continue;
            }
            if (parts > 1) {
                partParams.slabFiltered = false;  // a part filters its own triangles out of the slab's
            }
        }
        partParams.bitmapResult = wantBitmap;
        partParams.packedResult = wantPacked;
        // the part before the last one has left its device buffers and its pinned buffer: this part takes them over
        waitInFlightBelow(2);
        if (deviceFailed || sink->failed) {
            break;
        }
        RunStats partStats;
        const auto tCall = std::chrono::steady_clock::now();
        const int rc = engine->voxelize(partMesh, textures, textureCount, partParams, stream, &partStats);
        msVoxelizeCalls += msSince(tCall);
        if (rc != 0) {
            failWith("Voxelization failed on the device: " + engine->lastError());
            break;
        }
        msKernels += partStats.msTotal;
        if (!any) {
            stats = partStats;
            any = true;
        }
        else {
            accumulateStats(stats, partStats);
        }
        const unsigned long long total = engine->voxelCount();
        if (total == 0) {
            continue;
        }
        const int slot = (int) (launched & 1u);
        Pending mine;
        mine.slot = slot;
        mine.count = total;
        if (engine->hasBitmapResult()) {
            usedBitmap = true;
            const BitmapResult b = engine->bitmapResult();
            const size_t bitsBytes = (size_t) b.chunks * kChunkWords * 8, listBytes = (size_t) b.chunks * 4;
            unsigned char *host = static_cast<unsigned char *>(engine->pinnedStaging(slot, bitsBytes + 2 * listBytes));
            if (host == nullptr) {
                failWith("pinned host allocation failed (bitmap download)");
                break;
            }
            const bool ok =
                cudaMemcpyAsync(host, b.bits, bitsBytes, cudaMemcpyDeviceToHost, copyStream) == cudaSuccess &&
                cudaMemcpyAsync(host + bitsBytes, b.chunkIds, listBytes, cudaMemcpyDeviceToHost, copyStream) ==
                    cudaSuccess &&
                cudaMemcpyAsync(host + bitsBytes + listBytes, b.chunkCounts, listBytes, cudaMemcpyDeviceToHost,
                                copyStream) == cudaSuccess &&
                cudaEventRecord(copied[slot], copyStream) == cudaSuccess;
            if (!ok) {
                failWith(std::string("bitmap download failed: ") + cudaGetErrorString(cudaGetLastError()));
                break;
            }
            downloadBytes += bitsBytes + 2 * listBytes;
            mine.bitmap = true;
            mine.chunks = b.chunks;
            mine.geometry = b;
            mine.host = host;
            enqueue(mine);
            ++launched;
            engine->swapBitmapBuffers();
            continue;
        }
        const auto *deviceRecords = reinterpret_cast<const unsigned char *>(engine->deviceVoxels());
        const int packedBits = engine->packedBits();
        const size_t voxelBytes = packedBits == 0 ? 16u : (size_t) packedBits / 8u;
        if (packedBits != 0) {
            usedBitmap = true;  // (the timing line: "host expansion")
        }
        unsigned char *host = (parts > 1 || packedBits != 0) && total * voxelBytes <= (1ull << 30)
                                  ? static_cast<unsigned char *>(engine->pinnedStaging(slot, (size_t) total * voxelBytes))
                                  : nullptr;
        if (host == nullptr && packedBits != 0) {
            failWith("pinned host allocation failed (packed download)");
            break;
        }
        if (host == nullptr) {
            // one part, or too big to pin in one piece (or pinning failed): batches through the two small buffers, once
            // nothing else uses them
            waitInFlightBelow(1);
            streamOut(deviceRecords, total);
            downloadBytes += total * 16;
            continue;
        }
        const bool ok = cudaMemcpyAsync(host, deviceRecords, (size_t) total * voxelBytes, cudaMemcpyDeviceToHost,
                                        copyStream) == cudaSuccess &&
                        cudaEventRecord(copied[slot], copyStream) == cudaSuccess;
        if (!ok) {
            failWith(std::string("voxel download failed: ") + cudaGetErrorString(cudaGetLastError()));
            break;
        }
        downloadBytes += total * voxelBytes;
        mine.packedBits = packedBits;
        mine.host = host;
        enqueue(mine);
        ++launched;
        engine->swapOutputBuffers();
    }
    waitInFlightBelow(1);
    cudaStreamSynchronize(copyStream);
    return !deviceFailed;
}

/// Rendezvous of the job's device threads.
class Barrier {
public:
    explicit Barrier(unsigned count) : count_(count) {}
    void arriveAndWait()
    {
        std::unique_lock<std::mutex> lock{mutex_};
        const unsigned generation = generation_;
        if (++arrived_ == count_) {
            arrived_ = 0;
            ++generation_;
            wake_.notify_all();
            return;
        }
        wake_.wait(lock, [&] { return generation_ != generation; });
    }

private:
    const unsigned count_;
    unsigned arrived_ = 0, generation_ = 0;
    std::mutex mutex_;
    std::condition_variable wake_;
};

/// Slab bounds of a job over `devices` devices: whole 64-sample chunk rows (128 when downscaling, so that an output
/// chunk row has one owner), as even as the row count allows; devices beyond the row count get an empty slab.
void planDeviceSlabs(uint32_t sampleRes, uint32_t supersampling, uint32_t jobZ0, uint32_t jobZ1, uint32_t devices,
                     uint32_t *bounds)
{
    const uint32_t gridExtent = (sampleRes + 63u) / 64u * 64u;
    if (jobZ0 == 0 && jobZ1 == 0) {
        jobZ1 = gridExtent;
    }
    jobZ1 = std::min(jobZ1, gridExtent);
    jobZ0 = std::min(jobZ0, jobZ1);
    const uint32_t unit = 64u * supersampling;
    const uint32_t row0 = jobZ0 / unit, row1 = std::max((jobZ1 + unit - 1) / unit, row0);
    const uint32_t rows = row1 - row0;
    for (uint32_t d = 0; d <= devices; ++d) {
        const uint32_t z = (row0 + (uint32_t) ((unsigned long long) rows * d / devices)) * unit;
        bounds[d] = std::min(std::max(z, jobZ0), jobZ1);
    }
    bounds[0] = jobZ0;
    bounds[devices] = jobZ1;
}

/// Slab bounds that give every device about the same number of (triangle, row) pairs: `histogram` = triangles per row
/// of `unit` sample layers from z = 0; the job covers [jobZ0, jobZ1) (as planDeviceSlabs left bounds[0] and
/// bounds[devices]).  Inner bounds stay multiples of `unit`; a device may end up with an empty slab (it sits the job out).
void balanceDeviceSlabs(const std::vector<unsigned long long> &histogram, uint32_t unit, uint32_t devices, uint32_t *bounds)
{
    const uint32_t jobZ0 = bounds[0], jobZ1 = bounds[devices];
    const uint32_t row0 = jobZ0 / unit, row1 = std::min<uint32_t>((jobZ1 + unit - 1) / unit, (uint32_t) histogram.size());
    unsigned long long total = 0;
    for (uint32_t r = row0; r < row1; ++r) {
        total += histogram[r];
    }
    if (total == 0 || row1 <= row0) {
        return;  // nothing to go by: the equal rows stand
    }
    unsigned long long running = 0;
    uint32_t row = row0;
    for (uint32_t d = 1; d < devices; ++d) {
        const unsigned long long target = total * d / devices;
        while (row < row1 && running + histogram[row] / 2 < target) {  // the row goes to the side its middle falls on
            running += histogram[row];
            ++row;
        }
        bounds[d] = std::min(std::max(row * unit, jobZ0), jobZ1);
    }
}

}  // namespace

Engine *sharedEngine(int device, std::string *error)
{
    std::lock_guard<std::mutex> lock{gEngineMutex};
    auto found = gEngines.find(device);
    if (found != gEngines.end()) {
        return found->second.get();
    }
    Engine *engine = Engine::create(device, error);
    if (engine != nullptr) {
        gEngines[device].reset(engine);
    }
    return engine;
}

std::vector<int> defaultJobDevices()
{
    std::vector<int> devices;
    if (const char *env = getenv("O2V_B200_DEVICES")) {
        int available = 0;
        if (cudaGetDeviceCount(&available) != cudaSuccess) {
            cudaGetLastError();
            available = 0;
        }
        const std::string spec = env;
        if (spec == "all") {
            for (int d = 0; d < available; ++d) {
                devices.push_back(d);
            }
        }
        else if (spec.find(',') != std::string::npos) {
            size_t at = 0;
            while (at < spec.size()) {
                const size_t comma = spec.find(',', at);
                const std::string item = spec.substr(at, comma == std::string::npos ? std::string::npos : comma - at);
                if (!item.empty()) {
                    devices.push_back(atoi(item.c_str()));
                }
                if (comma == std::string::npos) {
                    break;
                }
                at = comma + 1;
            }
        }
        else {
            const int count = std::min(std::max(atoi(env), 1), std::max(available, 1));
            for (int d = 0; d < count; ++d) {
                devices.push_back(d);
            }
        }
    }
    if (devices.empty()) {
        const char *env = getenv("O2V_B200_DEVICE");
        devices.push_back(env != nullptr ? atoi(env) : 0);
    }
    if (devices.size() > kMaxSlabs) {
        devices.resize(kMaxSlabs);
    }
    return devices;
}

void planJobSlabs(uint32_t sampleRes, uint32_t supersampling, uint32_t jobZ0, uint32_t jobZ1, uint32_t devices,
                  const unsigned long long *histogram, uint32_t rows, uint32_t *bounds)
{
    planDeviceSlabs(sampleRes, supersampling, jobZ0, jobZ1, devices, bounds);
    if (histogram != nullptr && rows != 0) {
        balanceDeviceSlabs(std::vector<unsigned long long>(histogram, histogram + rows), 64u * supersampling, devices, bounds);
    }
}

uint32_t planJobParts(uint32_t sampleRes, uint32_t slabZ0, uint32_t slabZ1, unsigned long long triangles, int requested,
                      uint32_t *bounds)
{
    const uint32_t gridExtent = (sampleRes + 63u) / 64u * 64u;
    uint32_t jobZ0 = slabZ0, jobZ1 = slabZ1;
    if (jobZ0 == 0 && jobZ1 == 0) {
        jobZ1 = gridExtent;
    }
    jobZ1 = std::min(jobZ1, gridExtent);
    jobZ0 = std::min(jobZ0, jobZ1);
    const uint32_t row0 = jobZ0 / 64u, row1 = std::max((jobZ1 + 63u) / 64u, row0 + 1u);
    const uint32_t rows = row1 - row0;
    // one part per 2^20 triangles, at most four: the download of a part (and the host threads' work on it) runs under the
    // kernels of the next, but every part pays a filter pass over the slab's triangles and two host round trips
    uint32_t parts = std::min<uint32_t>({4u, rows, (uint32_t) std::max<unsigned long long>(triangles >> 20, 1ull)});
    if (requested > 0) {
        parts = std::min(std::min((uint32_t) requested, rows), kMaxJobParts);
    }
    for (uint32_t k = 0; k <= parts; ++k) {
        const uint32_t z = (row0 + (uint32_t) ((unsigned long long) rows * k / parts)) * 64u;
        bounds[k] = std::min(std::max(z, jobZ0), jobZ1);
    }
    return parts;
}

void accumulateStats(RunStats &total, const RunStats &part)
{
    RunCounters &t = total.counters;
    const RunCounters &p = part.counters;
    t.voxels += p.voxels;
    t.leaves += p.leaves;
    t.pairs += p.pairs;
    t.activeTiles += p.activeTiles;
    t.candidateVoxels += p.candidateVoxels;
    t.clipCalls += p.clipCalls;
    t.contributions += p.contributions;
    t.droppedTriangles = std::max(t.droppedTriangles, p.droppedTriangles);
    t.depthOverflow = std::max(t.depthOverflow, p.depthOverflow);
    t.lightTiles += p.lightTiles;
    t.bigLightTiles += p.bigLightTiles;
    t.heavyTiles += p.heavyTiles;
    t.survivors += p.survivors;
    t.ranges += p.ranges;
    total.outCapacity = std::max(total.outCapacity, part.outCapacity);
    total.msTotal += part.msTotal;
    total.msSetup += part.msSetup;
    total.msVoxelize += part.msVoxelize;
    total.msClip += part.msClip;
    total.msClassify += part.msClassify;
    total.msFilter += part.msFilter;
    total.msExpand += part.msExpand;
    total.kernelLaunches += part.kernelLaunches;
    total.voxelizeLaunches += part.voxelizeLaunches;
    total.occupancyPath = total.occupancyPath && part.occupancyPath;
    total.slabTriangles += part.slabTriangles;
    total.downloadBytes += part.downloadBytes;
}

unsigned long long expandBitmapsOnHost(const unsigned long long *bits, const uint32_t *chunkIds,
                                       const uint32_t *chunkCounts, uint32_t chunks, uint32_t chunksPerAxis,
                                       uint32_t chunkZ0, uint32_t *records)
{
    std::vector<unsigned long long> offset((size_t) chunks + 1, 0);
    for (uint32_t c = 0; c < chunks; ++c) {
        offset[c + 1] = offset[c] + chunkCounts[c];
    }
    std::atomic<bool> consistent{true};
#if defined(__SSE2__)
    static const struct BitOffsets {
        __m128i v[64];
        BitOffsets()
        {
            for (int b = 0; b < 64; ++b) {
                v[b] = _mm_set_epi32(0, 0, b >> 3, b & 7);
            }
        }
    } table;
    const __m128i *bitOffsets = table.v;
#endif
    hostPool().parallelFor(chunks, [&](size_t c) {
        if (chunkCounts[c] == 0) {
            return;
        }
        const uint32_t chunk = chunkIds[c], C = chunksPerAxis;
        const uint32_t cx = (chunk % C) * kChunkEdge, cy = ((chunk / C) % C) * kChunkEdge;
        const uint32_t cz = (chunk / (C * C) + chunkZ0) * kChunkEdge;
        const unsigned long long *words = bits + c * kChunkWords;
        uint32_t *out = records + offset[c] * 4;
        uint32_t *const end = records + offset[c + 1] * 4;
        for (uint32_t tile = 0; tile < kChunkWords / kTileEdge; ++tile) {
            const uint32_t ox = cx + (tile & 7u) * kTileEdge, oy = cy + ((tile >> 3) & 7u) * kTileEdge;
            const uint32_t oz = cz + (tile >> 6) * kTileEdge;
            for (uint32_t layer = 0; layer < kTileEdge; ++layer) {
                unsigned long long w = words[tile * kTileEdge + layer];
                if (w == 0) {
                    continue;
                }
                if (out + 4 * (size_t) __builtin_popcountll(w) > end) {  // more bits than the device counted: never write
                    consistent = false;                                  // past the chunk's share
                    return;
                }
#if defined(__SSE2__)
                // records are written once and read by someone else later: keep them out of this core's cache
                const __m128i base = _mm_set_epi32(-1, (int) (oz + layer), (int) oy, (int) ox);
                do {
                    const uint32_t b = (uint32_t) __builtin_ctzll(w);
                    w &= w - 1;
                    _mm_stream_si128(reinterpret_cast<__m128i *>(out), _mm_add_epi32(base, bitOffsets[b]));
                    out += 4;
                } while (w != 0);
#else
                do {
                    const uint32_t b = (uint32_t) __builtin_ctzll(w);
                    w &= w - 1;
                    out[0] = ox + (b & 7u);
                    out[1] = oy + (b >> 3);
                    out[2] = oz + layer;
                    out[3] = 0xFFFFFFFFu;
                    out += 4;
                } while (w != 0);
#endif
            }
        }
        if (out != end) {
            consistent = false;
        }
    });
#if defined(__SSE2__)
    _mm_sfence();
#endif
    return consistent ? offset[chunks] : ~0ull;
}

namespace {

/// Voxels [first, last) of a packed array into records[0 ...] — one host thread.  streaming: store past the caches (the
/// records are written once and read by someone else later); otherwise ordinary stores (a buffer that is reused while
/// it is still cached).
template <bool kStreaming>
void unpackRange(const void *packed, int bits, unsigned long long first, unsigned long long last, uint32_t *out)
{
    unsigned long long i = first;
#if defined(__SSE2__)
    auto store = [](uint32_t *to, __m128i value) {
        if (kStreaming) {
            _mm_stream_si128(reinterpret_cast<__m128i *>(to), value);
        }
        else {
            _mm_store_si128(reinterpret_cast<__m128i *>(to), value);
        }
    };
#endif
    if (bits == 32) {
        const uint32_t *in = static_cast<const uint32_t *>(packed);
#if defined(__SSE2__)
        // four voxels per round: unpack the three fields, transpose into four records
        const __m128i mask = _mm_set1_epi32(1023), white = _mm_set1_epi32(-1);
        for (; i + 4 <= last; i += 4, out += 16) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(in + i));
            const __m128i x = _mm_and_si128(v, mask), y = _mm_and_si128(_mm_srli_epi32(v, 10), mask);
            const __m128i z = _mm_srli_epi32(v, 20);
            const __m128i xyLo = _mm_unpacklo_epi32(x, y), xyHi = _mm_unpackhi_epi32(x, y);
            const __m128i zaLo = _mm_unpacklo_epi32(z, white), zaHi = _mm_unpackhi_epi32(z, white);
            store(out, _mm_unpacklo_epi64(xyLo, zaLo));
            store(out + 4, _mm_unpackhi_epi64(xyLo, zaLo));
            store(out + 8, _mm_unpacklo_epi64(xyHi, zaHi));
            store(out + 12, _mm_unpackhi_epi64(xyHi, zaHi));
        }
#endif
        for (; i < last; ++i, out += 4) {
            const uint32_t v = in[i];
            out[0] = v & 1023u;
            out[1] = (v >> 10) & 1023u;
            out[2] = v >> 20;
            out[3] = 0xFFFFFFFFu;
        }
    }
    else {
        const unsigned long long *in = static_cast<const unsigned long long *>(packed);
        for (; i < last; ++i, out += 4) {
            const unsigned long long v = in[i];
#if defined(__SSE2__)
            store(out, _mm_set_epi32(-1, (int) (v >> 42), (int) ((v >> 21) & 0x1fffffu), (int) (v & 0x1fffffu)));
#else
            out[0] = (uint32_t) (v & 0x1fffffu);
            out[1] = (uint32_t) ((v >> 21) & 0x1fffffu);
            out[2] = (uint32_t) (v >> 42);
            out[3] = 0xFFFFFFFFu;
#endif
        }
    }
}

}  // namespace

void expandPackedOnHost(const void *packed, int bits, unsigned long long count, uint32_t *records)
{
    constexpr unsigned long long kBlock = 1ull << 15;  // voxels per task: 128 KiB in, 512 KiB out
    hostPool().parallelFor((size_t) ((count + kBlock - 1) / kBlock), [&](size_t task) {
        const unsigned long long first = task * kBlock, last = std::min(count, first + kBlock);
        unpackRange<true>(packed, bits, first, last, records + first * 4);
    });
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

obj2voxel_error_t runDeviceJob(const o2v_b200_mesh &mesh, const std::vector<o2v_b200_texture> &textures,
                               const JobOptions &options, VoxelSink &sinkTarget, RunStats *statsOut, JobTimings *timingsOut)
{
    std::lock_guard<std::mutex> jobLock{gJobMutex};
    const auto tJob = std::chrono::steady_clock::now();
    JobTimings timings;
    std::vector<int> devices = options.devices.empty() ? defaultJobDevices() : options.devices;
    if (devices.size() > kMaxSlabs) {
        devices.resize(kMaxSlabs);
    }

    // the occupancy-only path never reads attributes (Engine::voxelize takes the same decision)
    const bool occupancy = options.params.occupancyPath != 0 && mesh.types == nullptr &&
                           !(mesh.uvs != nullptr && !textures.empty());
    const uint32_t S = options.params.resolution * options.params.supersampling;
    uint32_t slabBounds[kMaxSlabs + 1];
    planDeviceSlabs(S, options.params.supersampling, options.params.slabZ0, options.params.slabZ1,
                    (uint32_t) devices.size(), slabBounds);
    // devices whose slab is empty (more devices than chunk rows) sit the job out
    std::vector<int> active;
    std::vector<uint32_t> activeBounds;
    for (size_t d = 0; d < devices.size(); ++d) {
        if (slabBounds[d] < slabBounds[d + 1]) {
            if (active.empty()) {
                activeBounds.push_back(slabBounds[d]);
            }
            active.push_back(devices[d]);
            activeBounds.push_back(slabBounds[d + 1]);
        }
    }
    if (active.empty()) {
        active.push_back(devices[0]);
        activeBounds = {options.params.slabZ0, options.params.slabZ1};
    }
    const uint32_t D = (uint32_t) active.size();

    std::vector<Engine *> engines(D, nullptr);
    for (uint32_t d = 0; d < D; ++d) {
        std::string error;
        engines[d] = sharedEngine(active[d], &error);
        if (engines[d] == nullptr) {
            logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Cannot voxelize: " + error);
            return OBJ2VOXEL_ERR_DEVICE;
        }
    }
    // Several devices on the occupancy-only path exchange triangles by Z-slab over peer memory; without peer access (or
    // with attributes to keep in step: the weighted path folds in triangle order) every device takes the whole mesh.
    const bool exchange = D > 1 && occupancy && enablePeerAccess(active);
    // Bitmaps instead of records over PCIe when one device would have to push 16 bytes per voxel through one link; with
    // several links the copy engines write the records faster than the host's threads could.
    DownloadMode download = options.download;
    if (const char *env = getenv("O2V_B200_DOWNLOAD")) {
        download = strcmp(env, "bitmap") == 0    ? DownloadMode::BITMAP
                   : strcmp(env, "records") == 0 ? DownloadMode::RECORDS
                   : strcmp(env, "packed") == 0  ? DownloadMode::PACKED
                                                 : download;
    }
    if (download == DownloadMode::AUTO) {
        // An all-white result crosses PCIe as positions packed into 4 (or 8) bytes per voxel, and the host's threads write
        // the quads batch by batch into buffers that stay in their caches.  Measured against it (profiles/r02_history.md):
        // 16-byte records through the copy engine (4 times the traffic), and the occupancy bitmaps themselves (a third
        // less traffic again, but scanning bits costs the host 4 times what unpacking positions does — with VPCOMPRESSB
        // as without).
        download = DownloadMode::PACKED;
    }
    const bool wantBitmap = download == DownloadMode::BITMAP && occupancy;
    const bool wantPacked = download == DownloadMode::PACKED && occupancy;
    timings.devices = D;
    timings.peerExchange = exchange;

    SharedSink sink{sinkTarget, {}, {}};
    Barrier barrier(D);
    std::mutex resultMutex;
    RunStats total;
    bool anyStats = false;
    std::atomic<bool> failed{false};
    std::string firstError;
    std::vector<float> shareMin((size_t) D * 3, 0.0f), shareMax((size_t) D * 3, 0.0f);
    std::vector<unsigned long long> sent((size_t) D * D, 0);  // [source][slab]
    const uint32_t slabUnit = 64u * options.params.supersampling;  // a slab bound: a whole row of output chunks
    const uint32_t histogramRows = std::min<uint32_t>(((S + 63u) / 64u * 64u + slabUnit - 1) / slabUnit, 128u);
    std::vector<unsigned long long> rowHistogram((size_t) D * 128, 0);  // [device][row]
    std::vector<uint32_t> slabs(activeBounds);  // the plan every device thread ends up with (equal rows, then balanced)
    static const bool balanceEnabled = getenv("O2V_B200_BALANCE") == nullptr || atoi(getenv("O2V_B200_BALANCE")) != 0;
    std::atomic<bool> staged{false}, usedBitmap{false}, streamedUpload{false};
    double msUpload = 0, msExchange = 0, msKernels = 0;
    uint32_t partsUsed = 0;

    auto fail = [&](const std::string &message) {
        std::lock_guard<std::mutex> lock{resultMutex};
        if (firstError.empty()) {
            firstError = message;
        }
        failed = true;
    };

    auto deviceThread = [&](uint32_t d) {
        const auto tThread = std::chrono::steady_clock::now();
        double msMyUpload = 0, msMyExchange = 0;
        Engine *engine = engines[d];
        DeviceState &state = deviceState(active[d]);
        bool ok = cudaSetDevice(active[d]) == cudaSuccess;
        cudaStream_t stream = nullptr;
        ok = ok && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess;
        if (!ok) {
            fail(std::string("cudaStreamCreate failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
        // ---- upload: this device's share when the slabs exchange triangles, the whole mesh otherwise ----
        const size_t n = (size_t) mesh.count;
        const size_t first = exchange ? n * d / D : 0, last = exchange ? n * (d + 1) / D : n;
        std::string error;
        bool wasStaged = false;
        // One device, an all-white mesh with known bounds in pinned memory: the array goes up in pieces and every piece is
        // voxelized while the next one crosses PCIe (the result is an OR: a piece delivers what no earlier piece has).
        std::vector<size_t> pieceFirst;
        std::vector<cudaEvent_t> pieceReady;
        cudaStream_t uploadStream = nullptr;
        const char *streamEnv = getenv("O2V_B200_STREAM_UPLOAD");  // (read per job: tests compare both ways)
        const bool streamEnabled = streamEnv == nullptr || atoi(streamEnv) != 0;
        const char *streamMinEnv = getenv("O2V_B200_STREAM_MIN");  // triangles from which it pays (tests lower it)
        const size_t streamMin = streamMinEnv != nullptr ? (size_t) strtoull(streamMinEnv, nullptr, 10) : (size_t) 2 << 20;
        bool streamed = false;
        const bool wholeGrid = options.params.slabZ0 == 0 && options.params.slabZ1 == 0;  // (a slab job filters per part)
        if (ok && !failed && D == 1 && occupancy && wantPacked && streamEnabled && options.params.boundsKnown && wholeGrid &&
            mesh.verts != nullptr && n >= streamMin && isPinnedHost(mesh.verts)) {
            const size_t bitmapBytes = Engine::accumulateBytes(options.params);
            // one piece per 1.5 M triangles, two to eight (cfg4, 10 M triangles: 4 pieces 10.1 ms, 6: 9.2, 8: 9.4, 12: 10.5
            // — every piece pays two host round trips and an expand pass over all bitmaps)
            const uint32_t pieces = options.parts > 0
                                        ? std::min((uint32_t) options.parts, kMaxJobParts)
                                        : (uint32_t) std::min<size_t>(std::max<size_t>(n / 1500000, 2), 8);
            bool fits = bitmapBytes != 0 && pieces > 1;
            if (fits && bitmapBytes + n * 36 > state.approvedBytes) {
                size_t freeBytes = 0, totalBytes = 0;
                fits = cudaMemGetInfo(&freeBytes, &totalBytes) == cudaSuccess && bitmapBytes + n * 36 < freeBytes / 2;
                if (fits) {
                    state.approvedBytes = bitmapBytes + n * 36;
                }
            }
            if (fits && state.uploadStream == nullptr &&
                cudaStreamCreateWithFlags(&state.uploadStream, cudaStreamNonBlocking) != cudaSuccess) {
                state.uploadStream = nullptr;
                fits = false;
            }
            while (fits && state.pieceEvents.size() < pieces) {
                cudaEvent_t made = nullptr;
                fits = cudaEventCreateWithFlags(&made, cudaEventDisableTiming) == cudaSuccess;
                if (fits) {
                    state.pieceEvents.push_back(made);
                }
            }
            uploadStream = state.uploadStream;
            if (fits && state.upload.verts.ensure(n * 9 * sizeof(float))) {
                streamed = true;
                state.upload.view = MeshView{};
                state.upload.view.verts = state.upload.verts.as<float>();
                state.upload.view.count = n;
                state.upload.textureViews.clear();
                for (uint32_t k = 0; k <= pieces; ++k) {
                    pieceFirst.push_back(n * k / pieces);
                }
                for (uint32_t k = 0; k < pieces && streamed; ++k) {
                    const size_t offset = pieceFirst[k] * 9, floats = (pieceFirst[k + 1] - pieceFirst[k]) * 9;
                    if (cudaMemcpyAsync(state.upload.verts.as<float>() + offset, mesh.verts + offset, floats * sizeof(float),
                                        cudaMemcpyHostToDevice, uploadStream) != cudaSuccess ||
                        cudaEventRecord(state.pieceEvents[k], uploadStream) != cudaSuccess) {
                        fail(std::string("mesh upload failed: ") + cudaGetErrorString(cudaGetLastError()));
                        streamed = false;
                    }
                    pieceReady.push_back(state.pieceEvents[k]);
                }
            }
        }
        if (!streamed && D == 1 && occupancy && n >= streamMin) {
            char why[256];
            snprintf(why, sizeof why,
                     "upload first, then the parts (streaming needs: packed download %d, enabled %d, known bounds %d, pinned "
                     "input %d, bitmaps of the whole grid %zu bytes, more than one piece)",
                     (int) wantPacked, (int) streamEnabled, (int) options.params.boundsKnown,
                     (int) (mesh.verts != nullptr && isPinnedHost(mesh.verts)), Engine::accumulateBytes(options.params));
            logMessage(OBJ2VOXEL_LOG_LEVEL_DEBUG, why);
        }
        if (!streamed && ok && !failed &&
            !state.upload.upload(active[d], mesh, first, last - first, textures, stream, &wasStaged, &error)) {
            fail(error);
        }
        if (ok && !streamed) {
            cudaStreamSynchronize(stream);
        }
        if (wasStaged) {
            staged = true;
        }
        msMyUpload = msSince(tThread);
        if (d == 0) {
            msUpload = msSince(tJob);
        }
        const auto tExchange = std::chrono::steady_clock::now();

        EngineParams params = options.params;
        MeshView view = state.upload.view;
        if (D > 1) {
            params.slabZ0 = activeBounds[d];
            params.slabZ1 = activeBounds[d + 1];
        }
        if (exchange) {
            // ---- bounds of the whole mesh from the shares, then every triangle to the device(s) of its slab(s) ----
            if (!options.params.boundsKnown) {
                if (ok && !failed && engine->meshBounds(view, stream, &shareMin[d * 3], &shareMax[d * 3]) != 0) {
                    fail("mesh bounds failed: " + engine->lastError());
                }
                barrier.arriveAndWait();
                params.boundsKnown = true;
                for (int a = 0; a < 3; ++a) {
                    float lo = shareMin[a], hi = shareMax[a];
                    for (uint32_t r = 1; r < D; ++r) {
                        lo = std::min(lo, shareMin[r * 3 + a]);
                        hi = std::max(hi, shareMax[r * 3 + a]);
                    }
                    params.bounds[a] = lo;
                    params.bounds[3 + a] = hi;
                }
            }
            // ---- slabs of equal work: triangles per chunk row, summed over the shares ----
            std::vector<uint32_t> mine(activeBounds);
            if (balanceEnabled) {
                if (ok && !failed && engine->zRowHistogram(view, params, slabUnit, histogramRows, stream,
                                                           &rowHistogram[(size_t) d * 128]) != 0) {
                    fail("z histogram failed: " + engine->lastError());
                }
                barrier.arriveAndWait();
                std::vector<unsigned long long> total(histogramRows, 0);
                for (uint32_t r = 0; r < D; ++r) {
                    for (uint32_t row = 0; row < histogramRows; ++row) {
                        total[row] += rowHistogram[(size_t) r * 128 + row];
                    }
                }
                balanceDeviceSlabs(total, slabUnit, D, mine.data());  // (every thread computes the same bounds)
                params.slabZ0 = mine[d];
                params.slabZ1 = mine[d + 1];
                if (d == 0) {
                    slabs = mine;
                }
            }
            const unsigned long long capacity = (unsigned long long) ((n + D - 1) / D);
            SlabScatter scatter{};
            scatter.slabs = D;
            for (uint32_t s = 0; s <= D; ++s) {
                scatter.bound[s] = mine[s];
            }
            scatter.capacity = capacity;
            if (ok && !failed && engine->receiveRegion(0, D, capacity) == nullptr) {  // this device's own buffer
                fail("device allocation failed (receive buffer)");
            }
            barrier.arriveAndWait();  // every receive buffer exists before anyone asks for its regions
            for (uint32_t s = 0; s < D && ok && !failed; ++s) {
                scatter.dest[s] = engines[s]->receiveRegion(d, D, capacity);  // (big enough already: a pointer, no allocation)
            }
            cudaSetDevice(active[d]);
            if (ok && !failed && engine->scatterToSlabs(view, params, scatter, stream, &sent[(size_t) d * D]) != 0) {
                fail("slab exchange failed: " + engine->lastError());
            }
            barrier.arriveAndWait();  // every source has delivered (its stream is synchronised)
            if (ok && !failed) {
                unsigned long long counts[kMaxSlabs], received = 0;
                for (uint32_t r = 0; r < D; ++r) {
                    counts[r] = sent[(size_t) r * D + d];
                }
                view = MeshView{};
                view.verts = engine->packReceived(counts, D, capacity, stream, &received);
                view.count = received;
                params.slabFiltered = true;
                if (view.verts == nullptr) {
                    fail("device allocation failed (received triangles)");
                }
            }
            msMyExchange = msSince(tExchange);
            if (d == 0) {
                msExchange = msMyExchange;
            }
        }

        // ---- this device's slab in parts ----
        SlabRun run;
        run.engine = engine;
        run.state = &state;
        run.mesh = view;
        run.textures = state.upload.textureViews.data();
        run.textureCount = (uint32_t) state.upload.textureViews.size();
        run.params = params;
        run.requestedParts = options.parts;
        run.wantBitmap = wantBitmap;
        run.wantPacked = wantPacked;
        run.stream = stream;
        run.sink = &sink;
        if (streamed) {
            run.pieceFirst = pieceFirst;
            run.pieceReady = pieceReady;
            run.textureCount = 0;
            streamedUpload = true;
        }
        const bool emptySlab = D > 1 && params.slabZ0 >= params.slabZ1;  // (never hand (0, 0) on: it means "the whole grid")
        if (ok && !failed && !emptySlab) {
            if (!run.run()) {
                fail(run.error);
            }
            std::lock_guard<std::mutex> lock{resultMutex};
            run.stats.downloadBytes = run.downloadBytes;
            if (run.any) {
                if (!anyStats) {
                    total = run.stats;
                    anyStats = true;
                }
                else {
                    accumulateStats(total, run.stats);
                }
            }
            msKernels = std::max(msKernels, run.msKernels);
            if (d == 0) {
                timings.msWaitCopy = run.msWaitCopy;
                timings.msExpandHost = run.msExpandHost;
                timings.msSink = run.msSink;
                timings.msVoxelizeCalls = run.msVoxelizeCalls;
            }
            partsUsed = std::max(partsUsed, run.parts);
            if (run.usedBitmap) {
                usedBitmap = true;
            }
        }
        if (streamed) {
            cudaStreamSynchronize(uploadStream);  // (the stream and the events stay with the device state)
        }
        if (stream != nullptr) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
        if (D > 1) {
            char line[320];
            snprintf(line, sizeof line,
                     "device %d: slab [%u, %u), %llu triangles; upload %.2f ms, exchange %.2f ms, %u part(s): voxelize calls "
                     "%.2f ms (device time %.2f), waiting for copies %.2f, host expansion %.2f, sink %.2f; thread %.2f ms",
                     active[d], params.slabZ0, params.slabZ1, (unsigned long long) view.count, msMyUpload, msMyExchange,
                     run.parts, run.msVoxelizeCalls, run.msKernels, run.msWaitCopy, run.msExpandHost, run.msSink,
                     msSince(tThread));
            logMessage(OBJ2VOXEL_LOG_LEVEL_DEBUG, line);
        }
    };

    if (D == 1) {
        deviceThread(0);
    }
    else {
        std::vector<std::thread> threads;
        for (uint32_t d = 0; d < D; ++d) {
            threads.emplace_back(deviceThread, d);
        }
        for (std::thread &t : threads) {
            t.join();
        }
    }

    timings.slabBounds = slabs;
    timings.msUpload = msUpload;
    timings.msExchange = msExchange;
    timings.msKernels = msKernels;
    timings.msRun = msSince(tJob) - msUpload;
    timings.parts = partsUsed;
    timings.bitmapDownload = usedBitmap;
    timings.stagedUpload = staged;
    timings.streamedUpload = streamedUpload;
    if (statsOut != nullptr) {
        *statsOut = total;
    }
    if (timingsOut != nullptr) {
        *timingsOut = timings;
    }
    if (failed) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, firstError.empty() ? "Voxelization failed on the device" : firstError);
        return OBJ2VOXEL_ERR_DEVICE;
    }
    if (sink.failed) {
        logMessage(OBJ2VOXEL_LOG_LEVEL_ERROR, "Voxelization failed because of IO error");
        return OBJ2VOXEL_ERR_IO_ERROR_DURING_VOXEL_WRITE;
    }
    return OBJ2VOXEL_ERR_OK;
}

}  // namespace o2v
