// Ingest: file / host memory -> device-resident stack, with the pass-1
// accumulation overlapped.  Replaces the reference's buffered np.fromfile loop
// (video_reader.py:94-123: 25 frames per read, consumed frame by frame on the
// host) with the payload MEMORY-MAPPED and copied by reader threads into a ring
// of pinned slots, H2D copies on a private copy stream and shg_accumulate on a
// private compute stream: slot i+1 is being filled while slot i crosses PCIe and
// slot i-1 is summed.  (Copying out of the mapping costs one user-space memcpy
// per byte; pread -- the first version, still the fall-back when the file cannot
// be mapped, SHG_INGEST_PREAD=1 forces it -- pays a kernel crossing and a
// per-page copy_to_user on top and fed the ring at 2.7 GB/s per thread.)
// A source that is already pinned is copied straight from where it lies.
// Bound: PCIe H2D (frame bytes cross once), then page-cache read rate.
#include <fcntl.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct shg_ingest {
    int device = 0;
    int64_t slot_bytes = 0;
    int n_slots = 0;
    int n_threads = 1;
    std::vector<void*> slots;
    std::vector<cudaEvent_t> slot_free;      // H2D out of the slot finished
    std::vector<bool> slot_used;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t compute_stream = nullptr;
    cudaEvent_t chunk_copied = nullptr;
};

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Fill dst with frames [f0, f0+nf) using n_threads readers.  Exactly one of fd / src is used.
int fill_slot(int fd, const unsigned char* src, int64_t payload_offset, int64_t frame_bytes, int64_t stride,
              int64_t f0, int64_t nf, unsigned char* dst, int n_threads, std::string* err) {
    std::atomic<int> failed{0};
    auto work = [&](int64_t a, int64_t b) {
        if (stride == frame_bytes) {
            int64_t off = payload_offset + a * stride, left = (b - a) * frame_bytes;
            unsigned char* d = dst + (a - f0) * frame_bytes;
            if (src) { memcpy(d, src + off, (size_t)left); return; }
            while (left > 0) {
                const ssize_t got = pread(fd, d, (size_t)std::min<int64_t>(left, 1 << 30), off);
                if (got <= 0) { failed = 1; return; }
                off += got; d += got; left -= got;
            }
        } else {
            for (int64_t f = a; f < b; ++f) {
                int64_t off = payload_offset + f * stride, left = frame_bytes;
                unsigned char* d = dst + (f - f0) * frame_bytes;
                if (src) { memcpy(d, src + off, (size_t)left); continue; }
                while (left > 0) {
                    const ssize_t got = pread(fd, d, (size_t)left, off);
                    if (got <= 0) { failed = 1; return; }
                    off += got; d += got; left -= got;
                }
            }
        }
    };
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, nf));
    if (nt == 1) {
        work(f0, f0 + nf);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const int64_t a = f0 + nf * t / nt, b = f0 + nf * (t + 1) / nt;
            th.emplace_back(work, a, b);
        }
        for (auto& t : th) t.join();
    }
    if (failed) { *err = "short read (file smaller than its header claims?)"; return 1; }
    return 0;
}

int run_ingest(shg_ingest* ing, int fd, const unsigned char* src, int64_t payload_offset, int64_t frame_bytes,
               int64_t stride, int64_t frame0, int64_t n_frames, void* d_stack, int bytes_per_px, uint64_t* d_sum,
               uint32_t* d_max, double* h_stats4) {
    SHG_REQUIRE(ing, "ingest: null handle");
    SHG_REQUIRE(frame_bytes > 0 && stride >= frame_bytes, "ingest: bad frame size / stride");
    SHG_REQUIRE(frame_bytes <= ing->slot_bytes, "ingest: a frame (%lld B) does not fit a slot (%lld B)",
                (long long)frame_bytes, (long long)ing->slot_bytes);
    SHG_CHECK(cudaSetDevice(ing->device));
    const double t0 = now_s();
    double t_read = 0.0;
    int64_t chunks = 0;
    const int64_t frame_px = frame_bytes / bytes_per_px;
    unsigned char* dst = static_cast<unsigned char*>(d_stack);

    // already-pinned contiguous source: copy from where it lies
    bool direct = false;
    if (src && stride == frame_bytes) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src + payload_offset + frame0 * stride) == cudaSuccess &&
            at.type == cudaMemoryTypeHost)
            direct = true;
        (void)cudaGetLastError();
    }
    const int64_t per_chunk = std::max<int64_t>(1, ing->slot_bytes / frame_bytes);
    for (int64_t f = 0; f < n_frames; f += per_chunk, ++chunks) {
        const int64_t nf = std::min(per_chunk, n_frames - f);
        const unsigned char* from;
        const int slot = (int)(chunks % ing->n_slots);
        if (direct) {
            from = src + payload_offset + (frame0 + f) * stride;
        } else {
            if (ing->slot_used[slot]) SHG_CHECK(cudaEventSynchronize(ing->slot_free[slot]));
            const double r0 = now_s();
            std::string err;
            if (fill_slot(fd, src, payload_offset, frame_bytes, stride, frame0 + f, nf,
                          static_cast<unsigned char*>(ing->slots[slot]), ing->n_threads, &err)) {
                shg_set_error("ingest: %s", err.c_str());
                return 3;
            }
            t_read += now_s() - r0;
            from = static_cast<unsigned char*>(ing->slots[slot]);
        }
        SHG_CHECK(cudaMemcpyAsync(dst + f * frame_bytes, from, (size_t)(nf * frame_bytes), cudaMemcpyHostToDevice,
                                  ing->copy_stream));
        if (!direct) {
            SHG_CHECK(cudaEventRecord(ing->slot_free[slot], ing->copy_stream));
            ing->slot_used[slot] = true;
        }
        if (d_sum && d_max) {
            SHG_CHECK(cudaEventRecord(ing->chunk_copied, ing->copy_stream));
            SHG_CHECK(cudaStreamWaitEvent(ing->compute_stream, ing->chunk_copied, 0));
            if (int rc = shg_accumulate(dst + f * frame_bytes, bytes_per_px, nf, frame_px, d_sum, d_max,
                                        ing->compute_stream))
                return rc;
        }
    }
    SHG_CHECK(cudaStreamSynchronize(ing->copy_stream));
    SHG_CHECK(cudaStreamSynchronize(ing->compute_stream));
    if (h_stats4) {
        h_stats4[0] = now_s() - t0;
        h_stats4[1] = t_read;
        h_stats4[2] = (double)(n_frames * frame_bytes);
        h_stats4[3] = (double)chunks;
    }
    return 0;
}

}  // namespace

extern "C" int shg_ingest_create(int device, int64_t slot_bytes, int n_slots, int n_threads, shg_ingest** out) {
    SHG_REQUIRE(out && slot_bytes > 0 && n_slots >= 2 && n_slots <= 64 && n_threads >= 1 && n_threads <= 256,
                "shg_ingest_create: bad arguments");
    SHG_CHECK(cudaSetDevice(device));
    shg_ingest* ing = new shg_ingest();
    ing->device = device;
    ing->slot_bytes = slot_bytes;
    ing->n_slots = n_slots;
    ing->n_threads = n_threads;
    ing->slot_used.assign(n_slots, false);
    for (int i = 0; i < n_slots; ++i) {
        void* p = nullptr;
        cudaEvent_t e;
        if (cudaHostAlloc(&p, (size_t)slot_bytes, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            shg_set_error("shg_ingest_create: cannot allocate pinned slot %d of %lld bytes: %s", i,
                          (long long)slot_bytes, cudaGetErrorString(cudaGetLastError()));
            shg_ingest_destroy(ing);
            return 1;
        }
        ing->slots.push_back(p);
        ing->slot_free.push_back(e);
    }
    SHG_CHECK(cudaStreamCreateWithFlags(&ing->copy_stream, cudaStreamNonBlocking));
    SHG_CHECK(cudaStreamCreateWithFlags(&ing->compute_stream, cudaStreamNonBlocking));
    SHG_CHECK(cudaEventCreateWithFlags(&ing->chunk_copied, cudaEventDisableTiming));
    *out = ing;
    return 0;
}

extern "C" int shg_ingest_destroy(shg_ingest* ing) {
    if (!ing) return 0;
    cudaSetDevice(ing->device);
    for (void* p : ing->slots) cudaFreeHost(p);
    for (cudaEvent_t e : ing->slot_free) cudaEventDestroy(e);
    if (ing->chunk_copied) cudaEventDestroy(ing->chunk_copied);
    if (ing->copy_stream) cudaStreamDestroy(ing->copy_stream);
    if (ing->compute_stream) cudaStreamDestroy(ing->compute_stream);
    delete ing;
    return 0;
}

extern "C" int shg_ingest_file(shg_ingest* ing, const char* path, int64_t payload_offset, int64_t frame_bytes,
                               int64_t frame_stride_bytes, int64_t frame0, int64_t n_frames, void* d_stack,
                               int bytes_per_px, uint64_t* d_sum, uint32_t* d_max, double* h_stats4) {
    SHG_REQUIRE(path, "shg_ingest_file: null path");
    const int fd = open(path, O_RDONLY);
    SHG_REQUIRE(fd >= 0, "shg_ingest_file: cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 ||
        st.st_size < payload_offset + (frame0 + n_frames - 1) * frame_stride_bytes + frame_bytes) {
        close(fd);
        shg_set_error("shg_ingest_file: %s is smaller than %lld frames of %lld bytes", path,
                      (long long)(frame0 + n_frames), (long long)frame_bytes);
        return 2;
    }
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    // map the bytes this call needs (page-aligned window) and let the reader threads memcpy out of the mapping
    const char* force = getenv("SHG_INGEST_PREAD");
    void* map = MAP_FAILED;
    int64_t map_off = 0, map_len = 0;
    if (!(force && atoi(force) != 0)) {
        const int64_t page = sysconf(_SC_PAGESIZE);
        const int64_t first = payload_offset + frame0 * frame_stride_bytes;
        const int64_t last = payload_offset + (frame0 + n_frames - 1) * frame_stride_bytes + frame_bytes;
        map_off = first / page * page;
        map_len = last - map_off;
        map = mmap(nullptr, (size_t)map_len, PROT_READ, MAP_SHARED, fd, (off_t)map_off);
        if (map != MAP_FAILED) {
            madvise(map, (size_t)map_len, MADV_SEQUENTIAL);
            madvise(map, (size_t)map_len, MADV_WILLNEED);
        }
    }
    int rc;
    if (map != MAP_FAILED) {
        // run_ingest addresses frame f at src + payload_offset + f*stride: rebase so that the mapping's first byte is
        // file offset map_off
        rc = run_ingest(ing, -1, static_cast<const unsigned char*>(map) - map_off, payload_offset, frame_bytes,
                        frame_stride_bytes, frame0, n_frames, d_stack, bytes_per_px, d_sum, d_max, h_stats4);
        munmap(map, (size_t)map_len);
    } else {
        rc = run_ingest(ing, fd, nullptr, payload_offset, frame_bytes, frame_stride_bytes, frame0, n_frames, d_stack,
                        bytes_per_px, d_sum, d_max, h_stats4);
    }
    close(fd);
    return rc;
}

extern "C" int shg_ingest_memory(shg_ingest* ing, const void* h_payload, int64_t frame_bytes,
                                 int64_t frame_stride_bytes, int64_t n_frames, void* d_stack, int bytes_per_px,
                                 uint64_t* d_sum, uint32_t* d_max, double* h_stats4) {
    SHG_REQUIRE(h_payload, "shg_ingest_memory: null payload");
    return run_ingest(ing, -1, static_cast<const unsigned char*>(h_payload), 0, frame_bytes, frame_stride_bytes, 0,
                      n_frames, d_stack, bytes_per_px, d_sum, d_max, h_stats4);
}

extern "C" int shg_host_alloc(int64_t bytes, void** out) {
    SHG_REQUIRE(out && bytes > 0, "shg_host_alloc: bad arguments");
    SHG_CHECK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return 0;
}

extern "C" int shg_host_free(void* p) {
    if (p) SHG_CHECK(cudaFreeHost(p));
    return 0;
}

extern "C" int shg_memcpy_async(void* dst, const void* src, int64_t bytes, int kind, void* stream) {
    SHG_REQUIRE(kind >= 1 && kind <= 3, "shg_memcpy_async: kind must be 1 (H2D), 2 (D2H) or 3 (D2D)");
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost
                                                                            : cudaMemcpyDeviceToDevice;
    SHG_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, k, as_stream(stream)));
    return 0;
}
