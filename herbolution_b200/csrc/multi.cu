// multi.cu -- one process, 1..8 GPUs behind a single handle (hb_create(device_ids, n_dev)).
//
// The reference's host is one process whose chunk work fans out over a thread pool (server/src/lib.rs:42-60,
// lib/src/task.rs:6-25).  hb_multi mirrors that shape: one hb_ctx per device, and a region call partitions the
// region's chunk columns into contiguous x-slabs, one per device (chunks are independent: no exchange, the slab's
// one-column apron of heights is recomputed from the noise function).  Every device is driven by its own host
// thread, which owns that ctx's stream pair, pinned ring and expander threads for the duration of the call; the
// caller's sink is serialised by one mutex, so it never runs concurrently with itself.
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "host_common.cuh"

using namespace hb;

struct hb_multi {
    std::vector<hb_ctx*> ctx;
    std::vector<int> devices;
    std::mutex sink_mu;
};

namespace {

// contiguous, balanced split of ix in [0, ncx) over n parts (herbolution_b200/sharding.py: slab_bounds)
void slab_bounds(int32_t ncx, int n, int i, int32_t* lo, int32_t* hi) {
    const int32_t base = ncx / n, rem = ncx % n;
    *lo = i * base + (i < rem ? i : rem);
    *hi = *lo + base + (i < rem ? 1 : 0);
}

struct SinkShim {
    hb_multi* m;
    hb_region_sink sink;
    void* user;
};

void locked_sink(void* u, uint64_t first, uint64_t n, const uint64_t* off, const uint32_t* cnt, const hb_instance3d* inst,
                 uint64_t n_inst, const uint16_t* ids) {
    SinkShim* s = static_cast<SinkShim*>(u);
    std::lock_guard<std::mutex> lk(s->m->sink_mu);
    s->sink(s->user, first, n, off, cnt, inst, n_inst, ids);
}

template <class F>
int for_each_ctx(hb_multi* m, F f) {
    if (!m) return fail(HB_ERR_INVALID, "multi handle is null");
    for (hb_ctx* c : m->ctx) {
        const int rc = f(c);
        if (rc != HB_OK) return rc;
    }
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_multi_create(const int32_t* devices, int32_t n_devices, hb_multi** out) {
    if (!out) return fail(HB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (n_devices < 0 || n_devices > HB_MAX_DEVICES) return fail(HB_ERR_INVALID, "n_devices must be in 0..%d", HB_MAX_DEVICES);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(HB_ERR_CUDA, "no CUDA device available (%s); libherbgpu has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> devs;
    if (n_devices == 0 || !devices) {  // all visible devices (at most HB_MAX_DEVICES)
        if (n_devices != 0) return fail(HB_ERR_INVALID, "devices is null");
        for (int i = 0; i < ndev && i < HB_MAX_DEVICES; ++i) devs.push_back(i);
    } else {
        for (int i = 0; i < n_devices; ++i) {
            for (int j = 0; j < i; ++j)
                if (devices[j] == devices[i]) return fail(HB_ERR_INVALID, "device %d listed twice", devices[i]);
            devs.push_back(devices[i]);
        }
    }
    hb_multi* m = new (std::nothrow) hb_multi();
    if (!m) return fail(HB_ERR_NOMEM, "out of host memory");
    for (int d : devs) {
        hb_ctx* c = nullptr;
        const int rc = hb_create(d, &c);
        if (rc != HB_OK) {
            hb_multi_destroy(m);
            return rc;  // hb_create recorded the message
        }
        m->ctx.push_back(c);
        m->devices.push_back(d);
    }
    // the expander threads of all devices share the host: split them evenly unless the caller says otherwise
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 8;
    const int per = (int)std::max<size_t>(1, std::min<size_t>(32, hw) / m->ctx.size());
    for (hb_ctx* c : m->ctx) hb_set_host_threads(c, per);
    *out = m;
    return HB_OK;
}

void hb_multi_destroy(hb_multi* m) {
    if (!m) return;
    for (hb_ctx* c : m->ctx) hb_destroy(c);
    delete m;
}

int32_t hb_multi_device_count(const hb_multi* m) { return m ? (int32_t)m->ctx.size() : 0; }

hb_ctx* hb_multi_ctx(hb_multi* m, int32_t i) {
    if (!m || i < 0 || (size_t)i >= m->ctx.size()) return nullptr;
    return m->ctx[i];
}

int hb_multi_set_world(hb_multi* m, int64_t seed, float freq, int32_t octaves, float lacunarity, float gain) {
    return for_each_ctx(m, [&](hb_ctx* c) { return hb_set_world(c, seed, freq, octaves, lacunarity, gain); });
}
int hb_multi_set_noise_kind(hb_multi* m, int32_t kind) {
    return for_each_ctx(m, [&](hb_ctx* c) { return hb_set_noise_kind(c, kind); });
}
int hb_multi_set_fma_mode(hb_multi* m, int32_t fused) {
    return for_each_ctx(m, [&](hb_ctx* c) { return hb_set_fma_mode(c, fused); });
}
int hb_multi_set_materials(hb_multi* m, int32_t n_materials, const int32_t* n_colors, const float* rgba) {
    return for_each_ctx(m, [&](hb_ctx* c) { return hb_set_materials(c, n_materials, n_colors, rgba); });
}
int hb_multi_set_region_transport(hb_multi* m, int32_t transport) {
    return for_each_ctx(m, [&](hb_ctx* c) { return hb_set_region_transport(c, transport); });
}
int hb_multi_set_host_threads(hb_multi* m, int32_t n_threads_per_device) {
    return for_each_ctx(m, [&](hb_ctx* c) { return hb_set_host_threads(c, n_threads_per_device); });
}

int hb_multi_slab(const hb_multi* m, const hb_region* region, int32_t i, int32_t* ix_begin, int32_t* ix_end) {
    if (!m || !region || !ix_begin || !ix_end) return fail(HB_ERR_INVALID, "null argument");
    if (i < 0 || (size_t)i >= m->ctx.size()) return fail(HB_ERR_INVALID, "device index %d out of range", i);
    if (region->ncx < 0) return fail(HB_ERR_INVALID, "bad region");
    slab_bounds(region->ncx, (int)m->ctx.size(), i, ix_begin, ix_end);
    return HB_OK;
}

int hb_multi_generate_mesh_region(hb_multi* m, const hb_region* region, int want_ids, hb_region_sink sink, void* user,
                                  uint64_t* total_chunks, uint64_t* total_quads) {
    if (!m || !region) return fail(HB_ERR_INVALID, "null argument");
    if (m->ctx.empty()) return fail(HB_ERR_INVALID, "multi handle owns no device");
    const int n = (int)m->ctx.size();
    std::vector<int> rc(n, HB_OK);
    std::vector<std::string> msg(n);
    std::vector<uint64_t> chunks(n, 0), quads(n, 0);
    SinkShim shim{m, sink, user};
    auto work = [&](int i) {
        int32_t lo, hi;
        slab_bounds(region->ncx, n, i, &lo, &hi);
        if (lo >= hi) return;  // more devices than columns
        rc[i] = hb_generate_mesh_region_slab(m->ctx[i], region, lo, hi, want_ids, sink ? locked_sink : nullptr, &shim, &chunks[i],
                                             &quads[i]);
        if (rc[i] != HB_OK) msg[i] = hb_last_error();  // thread-local: carry it to the caller's thread
    };
    // device 0 runs on the calling thread (a single-device handle spawns nothing)
    std::vector<std::thread> th;
    std::vector<int> inline_later;
    for (int i = 1; i < n; ++i) {
        try {
            th.emplace_back(work, i);
        } catch (...) {  // no thread to be had: this device's slab runs on the calling thread after device 0's
            inline_later.push_back(i);
        }
    }
    work(0);
    for (int i : inline_later) work(i);
    for (auto& t : th) t.join();
    uint64_t tc = 0, tq = 0;
    for (int i = 0; i < n; ++i) {
        if (rc[i] != HB_OK) return fail(rc[i], "device %d: %s", m->devices[i], msg[i].c_str());
        tc += chunks[i];
        tq += quads[i];
    }
    if (total_chunks) *total_chunks = tc;
    if (total_quads) *total_quads = tq;
    return HB_OK;
}

}  // extern "C"
