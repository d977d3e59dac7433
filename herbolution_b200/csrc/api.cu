// api.cu -- the C ABI of libherbgpu.so (include/herbgpu.h): context, host-pointer entry points,
// device-resident region plans.  Everything here is plumbing around the kernels in gen.cu/mesh.cu;
// there is deliberately no CPU implementation of any stage.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <emmintrin.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include "host_common.cuh"

using namespace hb;

namespace {

thread_local char g_err[512] = "";

}  // namespace

namespace hb {
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int copy_to_host(hb_ctx* c, cudaStream_t s, void* dst, const void* d_src, size_t bytes) {
    if (bytes == 0) return HB_OK;
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();  // an unregistered pointer is not an error here
    if (pinned) {
        HB_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
        HB_CUDA(cudaStreamSynchronize(s));
        return HB_OK;
    }
    const size_t piece = (size_t)16 << 20;
    for (int b = 0; b < 2; ++b) {
        HB_CUDA(c->h_bounce[b].reserve(piece));
        if (!c->ev_bounce[b]) HB_CUDA(cudaEventCreateWithFlags(&c->ev_bounce[b], cudaEventDisableTiming));
    }
    const size_t np = (bytes + piece - 1) / piece;
    auto issue = [&](size_t k) -> cudaError_t {
        const size_t o = k * piece, len = std::min(piece, bytes - o);
        cudaError_t e = cudaMemcpyAsync(c->h_bounce[k & 1].p, static_cast<const char*>(d_src) + o, len, cudaMemcpyDeviceToHost, s);
        return e != cudaSuccess ? e : cudaEventRecord(c->ev_bounce[k & 1], s);
    };
    HB_CUDA(issue(0));
    for (size_t k = 0; k < np; ++k) {
        if (k + 1 < np) HB_CUDA(issue(k + 1));
        HB_CUDA(cudaEventSynchronize(c->ev_bounce[k & 1]));
        const size_t o = k * piece, len = std::min(piece, bytes - o);
        memcpy(static_cast<char*>(dst) + o, c->h_bounce[k & 1].p, len);
    }
    return HB_OK;
}

}  // namespace hb

namespace {

// Instance3d model_0..2 per face: CubeFace::rotation (lib/src/spatial/face.rs:60-69) ->
// Quat::from_euler (lib/src/rotation.rs:77-88) -> to_axes (:109-116) -> * Mat3::from(Vec3::ONE)
// (lib/src/matrix/mat3.rs:63-90), columns extended with 0.0 (client/src/video/world/vertex.rs:53-62).
// Evaluated on the host with libm sinf/cosf, which is what f32::sin_cos resolves to on linux.
// volatile keeps the compiler from contracting or pre-folding the f32 expression tree.
void face_matrix(int face, float out12[12]) {
    const float FRAC_PI_2 = 1.57079632679489661923132169163975144f;
    const float PI = 3.14159265358979323846264338327950288f;
    float yaw = 0.0f, pitch = 0.0f, roll = 0.0f;
    switch (face) {
        case 0: pitch = FRAC_PI_2; break;   // East
        case 1: pitch = -FRAC_PI_2; break;  // West
        case 2: yaw = -FRAC_PI_2; break;    // Up
        case 3: yaw = FRAC_PI_2; break;     // Down
        case 4: break;                      // North
        default: pitch = PI; break;         // South
    }
    volatile float sx = sinf(yaw / 2.0f), cx = cosf(yaw / 2.0f);
    volatile float sy = sinf(pitch / 2.0f), cy = cosf(pitch / 2.0f);
    volatile float sz = sinf(roll / 2.0f), cz = cosf(roll / 2.0f);
    auto mul = [](float a, float b) { volatile float r = a * b; return (float)r; };
    auto add = [](float a, float b) { volatile float r = a + b; return (float)r; };
    auto sub = [](float a, float b) { volatile float r = a - b; return (float)r; };
    const float x = sub(mul(mul(sx, cy), cz), mul(mul(cx, sy), sz));
    const float y = add(mul(mul(cx, sy), cz), mul(mul(sx, cy), sz));
    const float z = sub(mul(mul(cx, cy), sz), mul(mul(sx, sy), cz));
    const float w = add(mul(mul(cx, cy), cz), mul(mul(sx, sy), sz));
    const float ax[3][3] = {
        {sub(1.f, mul(2.f, add(mul(y, y), mul(z, z)))), mul(2.f, add(mul(x, y), mul(z, w))), mul(2.f, sub(mul(x, z), mul(y, w)))},
        {mul(2.f, sub(mul(x, y), mul(z, w))), sub(1.f, mul(2.f, add(mul(x, x), mul(z, z)))), mul(2.f, add(mul(y, z), mul(x, w)))},
        {mul(2.f, add(mul(x, z), mul(y, w))), mul(2.f, sub(mul(y, z), mul(x, w))), sub(1.f, mul(2.f, add(mul(x, x), mul(y, y))))},
    };
    static const float S[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            out12[c * 4 + r] = add(add(mul(ax[0][r], S[c][0]), mul(ax[1][r], S[c][1])), mul(ax[2][r], S[c][2]));
    out12[3] = out12[7] = out12[11] = 0.0f;
}

// client/src/world/chunk.rs:271-275: `Vec4<u8>.cast() / 3.0` resolves to f64, then cast::<f32>(),
// `-a + 1.0`, max_each(0.15)
float ao_value(int k) {
    volatile double amount = (double)k / 3.0;
    volatile float a = (float)amount;
    volatile float factor = -a + 1.0f;
    return factor > 0.15f ? (float)factor : 0.15f;
}

void default_materials(MaterialsDev* m) {
    // Material::stone/dirt/grass, server/src/chunk/material.rs:27-61
    static const float c[3][3][4] = {
        {{0.5f, 0.5f, 0.5f, 1.0f}, {0.6f, 0.6f, 0.6f, 1.0f}, {0.7f, 0.7f, 0.7f, 1.0f}},
        {{0.4f, 0.3f, 0.2f, 1.0f}, {0.5f, 0.4f, 0.3f, 1.0f}, {0.6f, 0.5f, 0.4f, 1.0f}},
        {{0.1f, 0.8f, 0.1f, 1.0f}, {0.2f, 0.9f, 0.2f, 1.0f}, {0.3f, 1.0f, 0.3f, 1.0f}},
    };
    memset(m, 0, sizeof(*m));
    m->n_materials = 3;
    for (int i = 0; i < 3; ++i) {
        m->n_colors[i] = 3;
        memcpy(m->rgba[i], c[i], sizeof(c[i]));
    }
}

}  // namespace

struct hb_region_plan {
    hb_ctx* ctx = nullptr;
    hb_region region{};
    RegionDev dev{};
    int want_ids = 0;
    size_t max_chunks = 0, max_hcols = 0;
    uint64_t capacity = 0;
    Buf d_heights, d_counts, d_offsets, d_scan, d_total, d_out, d_ids, d_packed, d_status;
    Buf h_total;
    hb_region_plan() { h_total.pinned = true; }
    size_t n_chunks() const { return (size_t)(dev.ix_end - dev.ix_begin) * dev.ncz * dev.ncy; }
};

namespace {

void plan_target(hb_region_plan* p, int32_t ix_begin, int32_t ix_end) {
    const hb_region& r = p->region;
    RegionDev d;
    d.cx0 = r.cx0; d.cz0 = r.cz0; d.ncx = r.ncx; d.ncz = r.ncz; d.cy0 = r.cy0; d.ncy = r.ncy;
    d.ix_begin = ix_begin; d.ix_end = ix_end;
    d.hx_begin = std::max(0, ix_begin - 1);
    d.hx_end = std::min(r.ncx, ix_end + 1);
    p->dev = d;
}

int region_check(const hb_region* r) {
    if (!r) return fail(HB_ERR_INVALID, "region is null");
    if (r->ncx <= 0 || r->ncz <= 0 || r->ncy <= 0) return fail(HB_ERR_INVALID, "region extents must be positive");
    const int64_t x1 = (int64_t)r->cx0 + r->ncx - 1, z1 = (int64_t)r->cz0 + r->ncz - 1, y1 = (int64_t)r->cy0 + r->ncy - 1;
    if (!coord_ok(r->cx0) || !coord_ok(r->cz0) || !coord_ok(r->cy0) || x1 > HB_MAX_CHUNK_COORD ||
        z1 > HB_MAX_CHUNK_COORD || y1 > HB_MAX_CHUNK_COORD)
        return fail(HB_ERR_RANGE, "region exceeds +-%d chunks", HB_MAX_CHUNK_COORD);
    return HB_OK;
}

int plan_alloc(hb_region_plan* p, int32_t max_width) {
    const hb_region& r = p->region;
    p->max_chunks = (size_t)max_width * r.ncz * r.ncy;
    p->max_hcols = (size_t)std::min(r.ncx, max_width + 2) * r.ncz;
    HB_CUDA(p->d_heights.reserve(p->max_hcols * HB_CHUNK_AREA * sizeof(int32_t)));
    HB_CUDA(p->d_counts.reserve(p->max_chunks * sizeof(uint32_t)));
    HB_CUDA(p->d_offsets.reserve(p->max_chunks * sizeof(uint64_t)));
    HB_CUDA(p->d_scan.reserve(scan_workspace_bytes(p->max_chunks)));
    HB_CUDA(p->d_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(p->h_total.reserve(4 * sizeof(uint64_t)));
    if (p->want_ids) HB_CUDA(p->d_ids.reserve(p->max_chunks * HB_CHUNK_VOLUME * sizeof(uint16_t)));
    return HB_OK;
}

int plan_enqueue(hb_region_plan* p, int stages, cudaStream_t s) {
    hb_ctx* c = p->ctx;
    const size_t hcols = (size_t)(p->dev.hx_end - p->dev.hx_begin) * p->dev.ncz;
    if (stages & HB_STAGE_FUSED) {
        // heights + count + scan + emit as one launch; regions the fused kernel does not cover take the staged kernels
        stages &= ~HB_STAGE_FUSED;
        if (region_fused_supported(c->world, p->dev)) {
            if (!p->d_out.p && p->capacity) return fail(HB_ERR_INVALID, "plan has no instance buffer");
            HB_CUDA(p->d_status.reserve(region_fused_status_bytes(p->dev)));
            HB_CUDA(launch_region_fused(s, c->tables, c->d_mats, c->world, p->dev, p->d_heights.as<int32_t>(),
                                        p->d_counts.as<uint32_t>(), p->d_offsets.as<uint64_t>(), p->d_out.as<hb_instance3d>(),
                                        p->capacity, p->d_total.as<uint64_t>(), p->d_status.as<uint64_t>(),
                                        noise_fast_index_ok(c->world)));
            stages &= ~HB_STAGE_ALL_MESH;
        } else {
            stages |= HB_STAGE_ALL_MESH;
        }
    }
    if (stages & HB_STAGE_HEIGHTS) HB_CUDA(launch_heights_region(s, c->world, p->dev, p->d_heights.as<int32_t>(), 0, hcols));
    if (stages & HB_STAGE_COUNT) {
        HB_CUDA(cudaMemsetAsync(p->d_total.p, 0, 4 * sizeof(uint64_t), s));
        HB_CUDA(launch_mesh_region(s, false, c->tables, c->d_mats, p->dev, p->d_heights.as<int32_t>(),
                                   p->d_counts.as<uint32_t>(), nullptr, nullptr, 0, p->d_total.as<uint64_t>()));
    }
    if (stages & HB_STAGE_SCAN)
        HB_CUDA(launch_scan(s, p->d_counts.as<uint32_t>(), p->n_chunks(), p->d_offsets.as<uint64_t>(), p->d_scan.p,
                            p->d_total.as<uint64_t>(), false));
    if (stages & HB_STAGE_EMIT) {
        if (!p->d_out.p && p->capacity) return fail(HB_ERR_INVALID, "plan has no instance buffer");
        HB_CUDA(launch_mesh_region(s, true, c->tables, c->d_mats, p->dev, p->d_heights.as<int32_t>(),
                                   p->d_counts.as<uint32_t>(), p->d_offsets.as<uint64_t>(),
                                   p->d_out.as<hb_instance3d>(), p->capacity, p->d_total.as<uint64_t>()));
    }
    if (stages & HB_STAGE_EMIT_PACKED) {
        HB_CUDA(p->d_packed.reserve(std::max<uint64_t>(p->capacity, 4) * sizeof(hb_packed_quad)));
        HB_CUDA(launch_mesh_region(s, true, c->tables, c->d_mats, p->dev, p->d_heights.as<int32_t>(),
                                   p->d_counts.as<uint32_t>(), p->d_offsets.as<uint64_t>(), p->d_packed.p, p->capacity,
                                   p->d_total.as<uint64_t>(), 0, true));
    }
    if (stages & HB_STAGE_FILL) {
        if (!p->want_ids) return fail(HB_ERR_INVALID, "plan was created without want_ids");
        HB_CUDA(launch_fill_region(s, p->dev, p->d_heights.as<int32_t>(), p->d_ids.as<uint16_t>()));
    }
    return HB_OK;
}

void ensure_expander(hb_ctx* c);  // defined with the region streamer below

int plan_totals(hb_region_plan* p, cudaStream_t s, uint64_t tot[4]) {
    HB_CUDA(cudaMemcpyAsync(p->h_total.p, p->d_total.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    memcpy(tot, p->h_total.p, 4 * sizeof(uint64_t));
    return HB_OK;
}

}  // namespace

extern "C" {

const char* hb_last_error(void) { return g_err; }
const char* hb_version(void) { return "herbgpu 0.1 (sm_100a)"; }

int hb_create(int device, hb_ctx** out) {
    if (!out) return fail(HB_ERR_INVALID, "out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(HB_ERR_CUDA, "no CUDA device available (%s); libherbgpu has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(HB_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    HB_CUDA(cudaSetDevice(device));
    hb_ctx* c = new (std::nothrow) hb_ctx();
    if (!c) return fail(HB_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->seed = 0;
    c->world = WorldDev{0, 0.001f, 6, 2.0f, 0.6f, HB_NOISE_FBM, 1};
    for (int f = 0; f < 6; ++f) face_matrix(f, c->tables.face_mat[f]);
    for (int k = 0; k < 4; ++k) c->tables.ao_lut[k] = ao_value(k);
    default_materials(&c->mats);
    cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_mats, sizeof(MaterialsDev));
    if (ce == cudaSuccess) ce = cudaMemcpy(c->d_mats, &c->mats, sizeof(MaterialsDev), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        hb_destroy(c);  // releases whatever was created; nothing leaks on a failed create
        return fail(HB_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(ce));
    }
    *out = c;
    return HB_OK;
}

void hb_destroy(hb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    Buf* bufs[] = {&c->d_pts, &c->d_cols, &c->d_colidx, &c->d_heights, &c->d_noise, &c->d_ids, &c->d_cubes, &c->d_nbr,
                   &c->d_bitmaps, &c->d_summary, &c->d_counts, &c->d_offsets, &c->d_scan, &c->d_total, &c->d_out,
                   &c->h_total, &c->h_inst[0], &c->h_inst[1], &c->h_off[0], &c->h_off[1], &c->h_cnt[0], &c->h_cnt[1],
                   &c->h_ids[0], &c->h_ids[1], &c->h_heights[0], &c->h_heights[1], &c->h_bounce[0], &c->h_bounce[1], &c->h_packed[0], &c->h_packed[1]};
    for (Buf* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        if (c->stream_plan[i]) hb_region_plan_destroy(c->stream_plan[i]);
        if (c->ev_kernels[i]) cudaEventDestroy(c->ev_kernels[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_bounce[i]) cudaEventDestroy(c->ev_bounce[i]);
    }
    c->pool.resize(0);
    free(c->h_expanded);
    free(c->h_ids_plain);
    if (c->d_mats) cudaFree(c->d_mats);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy) cudaStreamDestroy(c->copy);
    delete c;
}

int hb_set_world(hb_ctx* c, int64_t seed, float freq, int32_t octaves, float lacunarity, float gain) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    if (octaves < 1 || octaves > 255) return fail(HB_ERR_INVALID, "octaves must be in 1..255 (u8 in the reference)");
    std::lock_guard<std::mutex> lk(c->mu);
    c->seed = seed;
    c->world = WorldDev{(int32_t)seed, freq, octaves, lacunarity, gain, c->world.kind, c->world.fma};
    return HB_OK;
}

int hb_set_noise_kind(hb_ctx* c, int32_t kind) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    if (kind < HB_NOISE_FBM || kind > HB_NOISE_GRADIENT) return fail(HB_ERR_INVALID, "unknown noise kind %d", kind);
    std::lock_guard<std::mutex> lk(c->mu);
    c->world.kind = kind;
    return HB_OK;
}

int hb_set_fma_mode(hb_ctx* c, int32_t fused) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    if (fused != 0 && fused != 1) return fail(HB_ERR_INVALID, "fma mode must be 0 (unfused) or 1 (fused)");
    std::lock_guard<std::mutex> lk(c->mu);
    c->world.fma = fused;
    return HB_OK;
}

int hb_set_materials(hb_ctx* c, int32_t n_materials, const int32_t* n_colors, const float* rgba) {
    if (!c || !n_colors || !rgba) return fail(HB_ERR_INVALID, "null argument");
    if (n_materials < 1 || n_materials > HB_MAX_MATERIALS)
        return fail(HB_ERR_INVALID, "n_materials must be in 1..%d", HB_MAX_MATERIALS);
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    MaterialsDev m;
    memset(&m, 0, sizeof(m));
    m.n_materials = n_materials;
    for (int i = 0; i < n_materials; ++i) {
        if (n_colors[i] < 1 || n_colors[i] > HB_MAX_COLORS)
            return fail(HB_ERR_INVALID, "material %d: n_colors must be in 1..%d", i + 1, HB_MAX_COLORS);
        m.n_colors[i] = n_colors[i];
        memcpy(m.rgba[i], rgba + (size_t)i * HB_MAX_COLORS * 4, sizeof(float) * HB_MAX_COLORS * 4);
    }
    c->mats = m;
    c->xt_valid = false;
    HB_CUDA(cudaMemcpyAsync(c->d_mats, &c->mats, sizeof(MaterialsDev), cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return HB_OK;
}

int hb_host_alloc(hb_ctx* c, size_t bytes, void** out) {
    if (!c || !out) return fail(HB_ERR_INVALID, "null argument");
    HB_CUDA(cudaSetDevice(c->device));
    HB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return HB_OK;
}

int hb_host_free(hb_ctx* c, void* p) {
    (void)c;  // pinned memory is not tied to the ctx: it may be freed after hb_destroy (ctx may be NULL)
    if (p) HB_CUDA(cudaFreeHost(p));
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// generation
// ------------------------------------------------------------------------------------------------

int hb_noise_tiles(hb_ctx* c, const int32_t* cols_xz, size_t n, float* out_noise, int32_t* out_heights) {
    if (!c || (n && !cols_xz)) return fail(HB_ERR_INVALID, "null argument");
    if (n == 0) return HB_OK;
    for (size_t i = 0; i < 2 * n; ++i)
        if (!coord_ok(cols_xz[i])) return fail(HB_ERR_RANGE, "column %zu outside +-%d", i / 2, HB_MAX_CHUNK_COORD);
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    const size_t batch = 16384;
    for (size_t b0 = 0; b0 < n; b0 += batch) {
        const size_t nb = std::min(batch, n - b0);
        HB_CUDA(c->d_cols.reserve(nb * 2 * sizeof(int32_t)));
        HB_CUDA(c->d_heights.reserve(nb * HB_CHUNK_AREA * sizeof(int32_t)));
        HB_CUDA(c->d_noise.reserve(nb * HB_CHUNK_AREA * sizeof(float)));
        HB_CUDA(cudaMemcpyAsync(c->d_cols.p, cols_xz + 2 * b0, nb * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(launch_heights(c->stream, c->world, c->d_cols.as<int32_t>(), nb, c->d_heights.as<int32_t>(),
                               out_noise ? c->d_noise.as<float>() : nullptr));
        if (out_noise)
            HB_CUDA(cudaMemcpyAsync(out_noise + b0 * HB_CHUNK_AREA, c->d_noise.p, nb * HB_CHUNK_AREA * sizeof(float),
                                    cudaMemcpyDeviceToHost, c->stream));
        if (out_heights)
            HB_CUDA(cudaMemcpyAsync(out_heights + b0 * HB_CHUNK_AREA, c->d_heights.p, nb * HB_CHUNK_AREA * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return HB_OK;
}

int hb_noise_fbm3d(hb_ctx* c, float sx, float sy, float sz, int32_t nx, int32_t ny, int32_t nz, const float* freq3,
                   float* out) {
    if (!c || !freq3 || !out) return fail(HB_ERR_INVALID, "null argument");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(HB_ERR_INVALID, "extents must be positive");
    // the reference advances coordinates by repeated f32 additions (noise/f32.rs:152-190); they are exact,
    // and equal to start + index, only for integer-valued starts below 2^24
    const float lim = 16777216.0f;
    const float s[3] = {sx, sy, sz};
    const int ext[3] = {nx, ny, nz};
    for (int a = 0; a < 3; ++a)
        if (!(s[a] == floorf(s[a])) || fabsf(s[a]) + (float)ext[a] > lim)
            return fail(HB_ERR_RANGE, "3-D tile start must be integer-valued with |start|+extent <= 2^24");
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    const size_t total = (size_t)nx * ny * nz;
    HB_CUDA(c->d_noise.reserve(total * sizeof(float)));
    HB_CUDA(launch_noise3d(c->stream, c->world, sx, sy, sz, nx, ny, nz, freq3[0], freq3[1], freq3[2], c->d_noise.as<float>()));
    HB_CUDA(cudaMemcpyAsync(out, c->d_noise.p, total * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return HB_OK;
}

int hb_generate(hb_ctx* c, const hb_chunk_pt* pts, size_t n, uint16_t* out_ids, hb_palette_cube* out_cubes,
                float* out_noise) {
    if (!c || (n && !pts)) return fail(HB_ERR_INVALID, "null argument");
    if (n == 0) return HB_OK;
    for (size_t i = 0; i < n; ++i)
        if (!coord_ok(pts[i].x) || !coord_ok(pts[i].y) || !coord_ok(pts[i].z))
            return fail(HB_ERR_RANGE, "chunk %zu outside +-%d", i, HB_MAX_CHUNK_COORD);
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    const size_t batch = out_cubes ? 2048 : 8192;
    // Block ids follow the region transport (hb_set_region_transport): with the packed one (default) the heights tiles
    // come back (4 KiB per distinct column) and the host threads write the ids straight into the caller's array
    // (fill_chunk_ids: the ids are a function of the column's heights); with HB_TRANSPORT_INSTANCES k_fill writes them
    // on the device and 64 KiB per chunk are copied.
    const bool host_ids = out_ids && c->transport == HB_TRANSPORT_PACKED && ((uintptr_t)out_ids & 15u) == 0;
    if (host_ids) ensure_expander(c);
    std::vector<int32_t> cols, colidx;
    std::unordered_map<KeyXZ, int32_t, KeyXZHash> map;
    for (size_t b0 = 0; b0 < n; b0 += batch) {
        const size_t nb = std::min(batch, n - b0);
        // every chunk of a column shares one noise tile: evaluate each (cx, cz) once per batch
        cols.clear(); colidx.resize(nb); map.clear();
        for (size_t i = 0; i < nb; ++i) {
            const KeyXZ k{pts[b0 + i].x, pts[b0 + i].z};
            auto it = map.find(k);
            if (it == map.end()) {
                it = map.emplace(k, (int32_t)(cols.size() / 2)).first;
                cols.push_back(k.x); cols.push_back(k.z);
            }
            colidx[i] = it->second;
        }
        const size_t ncol = cols.size() / 2;
        HB_CUDA(c->d_cols.reserve(ncol * 2 * sizeof(int32_t)));
        HB_CUDA(c->d_colidx.reserve(nb * sizeof(int32_t)));
        HB_CUDA(c->d_pts.reserve(nb * sizeof(hb_chunk_pt)));
        HB_CUDA(c->d_heights.reserve(ncol * HB_CHUNK_AREA * sizeof(int32_t)));
        if (out_noise) HB_CUDA(c->d_noise.reserve(ncol * HB_CHUNK_AREA * sizeof(float)));
        if (out_ids && !host_ids) HB_CUDA(c->d_ids.reserve(nb * HB_CHUNK_VOLUME * sizeof(uint16_t)));
        if (host_ids) HB_CUDA(c->h_heights[0].reserve(ncol * HB_CHUNK_AREA * sizeof(int32_t)));
        if (out_cubes) HB_CUDA(c->d_cubes.reserve(nb * HB_CHUNK_VOLUME * sizeof(hb_palette_cube)));
        HB_CUDA(cudaMemcpyAsync(c->d_cols.p, cols.data(), ncol * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(cudaMemcpyAsync(c->d_colidx.p, colidx.data(), nb * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(cudaMemcpyAsync(c->d_pts.p, pts + b0, nb * sizeof(hb_chunk_pt), cudaMemcpyHostToDevice, c->stream));
        int rc = hb_dev_generate(c, c->stream, c->d_cols.as<int32_t>(), ncol, c->d_pts.as<hb_chunk_pt>(),
                                 c->d_colidx.as<int32_t>(), nb, c->d_heights.as<int32_t>(),
                                 out_noise ? c->d_noise.as<float>() : nullptr,
                                 out_ids && !host_ids ? c->d_ids.as<uint16_t>() : nullptr,
                                 out_cubes ? c->d_cubes.as<hb_palette_cube>() : nullptr);
        if (rc != HB_OK) return rc;
        if (host_ids)
            HB_CUDA(cudaMemcpyAsync(c->h_heights[0].p, c->d_heights.p, ncol * HB_CHUNK_AREA * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                    c->stream));
        else if (out_ids)
            HB_CUDA(cudaMemcpyAsync(out_ids + b0 * HB_CHUNK_VOLUME, c->d_ids.p, nb * HB_CHUNK_VOLUME * sizeof(uint16_t),
                                    cudaMemcpyDeviceToHost, c->stream));
        if (out_cubes)
            HB_CUDA(cudaMemcpyAsync(out_cubes + b0 * HB_CHUNK_VOLUME, c->d_cubes.p,
                                    nb * HB_CHUNK_VOLUME * sizeof(hb_palette_cube), cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
        if (host_ids) {
            const int32_t* hts = c->h_heights[0].as<int32_t>();
            c->pool.run(nb, 4, [&](size_t i) {
                fill_chunk_ids(hts + (size_t)colidx[i] * HB_CHUNK_AREA, pts[b0 + i].y, out_ids + (b0 + i) * HB_CHUNK_VOLUME);
            });
        }
        if (out_noise) {
            // per-chunk tiles in the caller's order (a column's tile is repeated for each of its chunks)
            std::vector<float> tiles(ncol * HB_CHUNK_AREA);
            HB_CUDA(cudaMemcpy(tiles.data(), c->d_noise.p, tiles.size() * sizeof(float), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nb; ++i)
                memcpy(out_noise + (b0 + i) * HB_CHUNK_AREA, tiles.data() + (size_t)colidx[i] * HB_CHUNK_AREA,
                       HB_CHUNK_AREA * sizeof(float));
        }
    }
    return HB_OK;
}

int hb_dev_generate(hb_ctx* c, void* stream, const int32_t* d_cols_xz, size_t ncol, const hb_chunk_pt* d_pts,
                    const int32_t* d_col_of_chunk, size_t n, int32_t* d_heights, float* d_noise, uint16_t* d_ids,
                    hb_palette_cube* d_cubes) {
    if (!c || !d_cols_xz || !d_heights) return fail(HB_ERR_INVALID, "null argument");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    HB_CUDA(launch_heights(s, c->world, d_cols_xz, ncol, d_heights, d_noise));
    if (n && (d_ids || d_cubes)) {
        if (!d_pts || !d_col_of_chunk) return fail(HB_ERR_INVALID, "null argument");
        HB_CUDA(launch_fill(s, d_pts, d_col_of_chunk, n, d_heights, d_ids, d_cubes));
    }
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// meshing
// ------------------------------------------------------------------------------------------------

int hb_build_neighbor_index(const hb_chunk_pt* pts, size_t n, int32_t* out) {
    if ((n && !pts) || (n && !out)) return fail(HB_ERR_INVALID, "null argument");
    if (n > (size_t)INT32_MAX) return fail(HB_ERR_INVALID, "too many chunks");
    std::unordered_map<KeyXYZ, int32_t, KeyXYZHash> map;
    map.reserve(n * 2);
    for (size_t i = 0; i < n; ++i) map[KeyXYZ{pts[i].x, pts[i].y, pts[i].z}] = (int32_t)i;
    for (size_t i = 0; i < n; ++i) {
        int k = 0;
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz, ++k) {
                    auto it = map.find(KeyXYZ{pts[i].x + dx, pts[i].y + dy, pts[i].z + dz});
                    out[i * 27 + k] = it == map.end() ? -1 : it->second;
                }
    }
    return HB_OK;
}

size_t hb_dev_mesh_workspace_bytes(size_t n) {
    const size_t a = n * HB_CHUNK_AREA * sizeof(uint32_t);  // bitmaps
    const size_t b = (n * sizeof(uint32_t) + 255) & ~(size_t)255;  // summary
    const size_t d = n * 6 * 32 * sizeof(uint32_t) + (((n + 1) * sizeof(uint32_t) + 255) & ~(size_t)255);  // border records, work list
    return a + b + scan_workspace_bytes(n) + 256 + d;
}

}  // extern "C"

namespace {

// workspace layout shared by the ids and the PaletteCube (flags) paths: bitmaps | summary | scan | borders | [face planes]
struct MeshWs {
    uint32_t* bitmaps;
    uint32_t* summary;
    void* scan;
    uint32_t* borders;  // ids path: six 32-word planes per chunk (k_pack)
    uint32_t* planes;  // only in flags mode
};
size_t mesh_ws_bytes(size_t n, bool flags) {
    return hb_dev_mesh_workspace_bytes(n) + (flags ? n * 6 * HB_CHUNK_AREA * sizeof(uint32_t) : 0);
}
MeshWs mesh_ws(void* d_workspace, size_t n, bool flags) {
    uint8_t* ws = static_cast<uint8_t*>(d_workspace);
    MeshWs w;
    w.bitmaps = reinterpret_cast<uint32_t*>(ws);
    ws += n * HB_CHUNK_AREA * sizeof(uint32_t);
    w.summary = reinterpret_cast<uint32_t*>(ws);
    ws += (n * sizeof(uint32_t) + 255) & ~(size_t)255;
    w.scan = ws;
    ws += (scan_workspace_bytes(n) + 255) & ~(size_t)255;
    w.borders = reinterpret_cast<uint32_t*>(ws);
    ws += n * 6 * 32 * sizeof(uint32_t) + (((n + 1) * sizeof(uint32_t) + 255) & ~(size_t)255);  // + the emit work list
    w.planes = flags ? reinterpret_cast<uint32_t*>(ws) : nullptr;
    return w;
}

// pack + count + scan (+ emit when d_out): `blocks` are u16 ids, or hb_palette_cube when flags is set
int dev_mesh_impl(hb_ctx* c, cudaStream_t s, bool flags, const hb_chunk_pt* d_pts, size_t n, const void* d_blocks,
                  const int32_t* d_nbr, void* d_workspace, hb_instance3d* d_out, uint64_t out_capacity, uint64_t* d_offsets,
                  uint32_t* d_counts, uint64_t* d_total) {
    HB_CUDA(cudaMemsetAsync(d_total, 0, 4 * sizeof(uint64_t), s));
    if (n == 0) return HB_OK;
    const MeshWs w = mesh_ws(d_workspace, n, flags);
    if (flags)
        HB_CUDA(launch_pack_cubes(s, static_cast<const hb_palette_cube*>(d_blocks), n, c->mats.n_materials, w.bitmaps, w.summary,
                                  w.planes, d_total));
    else
        HB_CUDA(launch_pack(s, static_cast<const uint16_t*>(d_blocks), n, c->mats.n_materials, w.bitmaps, w.summary, w.borders,
                            d_counts, d_total));
    HB_CUDA(launch_mesh_ids(s, false, c->tables, c->d_mats, d_pts, n, d_blocks, d_nbr, w.bitmaps, w.summary, w.planes, d_counts,
                            nullptr, nullptr, 0, d_total, nullptr, w.borders));
    HB_CUDA(launch_scan(s, d_counts, n, d_offsets, w.scan, d_total, false));
    if (d_out)
        HB_CUDA(launch_mesh_ids(s, true, c->tables, c->d_mats, d_pts, n, d_blocks, d_nbr, w.bitmaps, w.summary, w.planes,
                                d_counts, d_offsets, d_out, out_capacity, d_total, nullptr, w.borders));
    return HB_OK;
}

// host-pointer meshing shared by hb_mesh (ids) and hb_mesh_cubes (PaletteCube with flags)
int mesh_host_impl(hb_ctx* c, bool flags, const hb_chunk_pt* pts, size_t n, const void* blocks, const int32_t* neighbor_index,
                   hb_instance3d* out, uint64_t out_capacity, uint64_t* out_offsets, uint32_t* out_counts, uint64_t* out_total) {
    if (!c || !out_total) return fail(HB_ERR_INVALID, "null argument");
    *out_total = 0;
    if (n == 0) return HB_OK;
    if (!pts || !blocks || !neighbor_index || !out_offsets || !out_counts) return fail(HB_ERR_INVALID, "null argument");
    for (size_t i = 0; i < n; ++i)
        if (!coord_ok(pts[i].x) || !coord_ok(pts[i].y) || !coord_ok(pts[i].z))
            return fail(HB_ERR_RANGE, "chunk %zu outside +-%d", i, HB_MAX_CHUNK_COORD);
    for (size_t i = 0; i < n * 27; ++i)
        if (neighbor_index[i] < -1 || neighbor_index[i] >= (int64_t)n)
            return fail(HB_ERR_INVALID, "neighbor_index[%zu] = %d out of range", i, neighbor_index[i]);
    const size_t block_bytes = n * HB_CHUNK_VOLUME * (flags ? sizeof(hb_palette_cube) : sizeof(uint16_t));
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    HB_CUDA(c->d_pts.reserve(n * sizeof(hb_chunk_pt)));
    HB_CUDA(c->d_ids.reserve(block_bytes));
    HB_CUDA(c->d_nbr.reserve(n * 27 * sizeof(int32_t)));
    HB_CUDA(c->d_bitmaps.reserve(mesh_ws_bytes(n, flags)));
    HB_CUDA(c->d_counts.reserve(n * sizeof(uint32_t)));
    HB_CUDA(c->d_offsets.reserve(n * sizeof(uint64_t)));
    HB_CUDA(c->d_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(c->h_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(cudaMemcpyAsync(c->d_pts.p, pts, n * sizeof(hb_chunk_pt), cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemcpyAsync(c->d_ids.p, blocks, block_bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemcpyAsync(c->d_nbr.p, neighbor_index, n * 27 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    // pass 1: pack + count + scan (no instance buffer yet)
    int rc = dev_mesh_impl(c, s, flags, c->d_pts.as<hb_chunk_pt>(), n, c->d_ids.p, c->d_nbr.as<int32_t>(), c->d_bitmaps.p,
                           nullptr, 0, c->d_offsets.as<uint64_t>(), c->d_counts.as<uint32_t>(), c->d_total.as<uint64_t>());
    if (rc != HB_OK) return rc;
    HB_CUDA(cudaMemcpyAsync(c->h_total.p, c->d_total.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    uint64_t tot[4];
    memcpy(tot, c->h_total.p, sizeof(tot));
    if (tot[2] & 2) return fail(HB_ERR_INVALID, "block array contains a material id above the palette size %d", c->mats.n_materials);
    *out_total = tot[0];
    HB_CUDA(cudaMemcpyAsync(out_offsets, c->d_offsets.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaMemcpyAsync(out_counts, c->d_counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (tot[0] > out_capacity || (tot[0] && !out)) {
        HB_CUDA(cudaStreamSynchronize(s));
        return fail(HB_ERR_CAPACITY, "instance buffer holds %llu, %llu required", (unsigned long long)out_capacity,
                    (unsigned long long)tot[0]);
    }
    if (tot[0]) {
        HB_CUDA(c->d_out.reserve(tot[0] * sizeof(hb_instance3d)));
        const MeshWs w = mesh_ws(c->d_bitmaps.p, n, flags);
        HB_CUDA(launch_mesh_ids(s, true, c->tables, c->d_mats, c->d_pts.as<hb_chunk_pt>(), n, c->d_ids.p, c->d_nbr.as<int32_t>(),
                                w.bitmaps, w.summary, w.planes, c->d_counts.as<uint32_t>(), c->d_offsets.as<uint64_t>(),
                                c->d_out.as<hb_instance3d>(), tot[0], c->d_total.as<uint64_t>(), nullptr, w.borders));
        HB_CUDA(cudaMemcpyAsync(out, c->d_out.p, tot[0] * sizeof(hb_instance3d), cudaMemcpyDeviceToHost, s));
    }
    HB_CUDA(cudaStreamSynchronize(s));
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_dev_mesh(hb_ctx* c, void* stream, const hb_chunk_pt* d_pts, size_t n, const uint16_t* d_ids,
                const int32_t* d_neighbor_index, void* d_workspace, hb_instance3d* d_out, uint64_t out_capacity,
                uint64_t* d_offsets, uint32_t* d_counts, uint64_t* d_total) {
    if (!c || !d_total) return fail(HB_ERR_INVALID, "null argument");
    if (n && (!d_pts || !d_ids || !d_neighbor_index || !d_workspace || !d_offsets || !d_counts))
        return fail(HB_ERR_INVALID, "null argument");
    return dev_mesh_impl(c, stream ? (cudaStream_t)stream : c->stream, false, d_pts, n, d_ids, d_neighbor_index, d_workspace,
                         d_out, out_capacity, d_offsets, d_counts, d_total);
}

int hb_mesh(hb_ctx* c, const hb_chunk_pt* pts, size_t n, const uint16_t* ids, const int32_t* neighbor_index,
            hb_instance3d* out, uint64_t out_capacity, uint64_t* out_offsets, uint32_t* out_counts, uint64_t* out_total) {
    return mesh_host_impl(c, false, pts, n, ids, neighbor_index, out, out_capacity, out_offsets, out_counts, out_total);
}

int hb_mesh_cubes(hb_ctx* c, const hb_chunk_pt* pts, size_t n, const hb_palette_cube* cubes, const int32_t* neighbor_index,
                  hb_instance3d* out, uint64_t out_capacity, uint64_t* out_offsets, uint32_t* out_counts, uint64_t* out_total) {
    return mesh_host_impl(c, true, pts, n, cubes, neighbor_index, out, out_capacity, out_offsets, out_counts, out_total);
}

// ------------------------------------------------------------------------------------------------
// fused region
// ------------------------------------------------------------------------------------------------

int hb_region_plan_create(hb_ctx* c, const hb_region* region, int32_t ix_begin, int32_t ix_end, uint64_t quad_capacity,
                          int want_ids, hb_region_plan** out) {
    if (!c || !out) return fail(HB_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = region_check(region);
    if (rc != HB_OK) return rc;
    if (ix_begin < 0 || ix_end > region->ncx || ix_begin >= ix_end) return fail(HB_ERR_INVALID, "bad slab [%d,%d)", ix_begin, ix_end);
    if ((uint64_t)(ix_end - ix_begin) * region->ncz * region->ncy > (uint64_t)INT32_MAX)
        return fail(HB_ERR_INVALID, "slab has too many chunks for one plan");
    HB_CUDA(cudaSetDevice(c->device));
    hb_region_plan* p = new (std::nothrow) hb_region_plan();
    if (!p) return fail(HB_ERR_NOMEM, "out of host memory");
    p->ctx = c;
    p->region = *region;
    p->want_ids = want_ids;
    plan_target(p, ix_begin, ix_end);
    rc = plan_alloc(p, ix_end - ix_begin);
    if (rc == HB_OK && quad_capacity == 0) {
        std::lock_guard<std::mutex> lk(c->mu);
        rc = plan_enqueue(p, HB_STAGE_HEIGHTS | HB_STAGE_COUNT | HB_STAGE_SCAN, c->stream);
        uint64_t tot[4] = {0, 0, 0, 0};
        if (rc == HB_OK) rc = plan_totals(p, c->stream, tot);
        quad_capacity = tot[0] ? tot[0] : 4;
    }
    if (rc == HB_OK) {
        p->capacity = quad_capacity;
        cudaError_t e = p->d_out.reserve(quad_capacity * sizeof(hb_instance3d));
        if (e != cudaSuccess) rc = fail(HB_ERR_CUDA, "instance buffer allocation failed: %s", cudaGetErrorString(e));
    }
    if (rc != HB_OK) { hb_region_plan_destroy(p); return rc; }
    *out = p;
    return HB_OK;
}

void hb_region_plan_destroy(hb_region_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    Buf* bufs[] = {&p->d_heights, &p->d_counts, &p->d_offsets, &p->d_scan, &p->d_total, &p->d_out, &p->d_ids, &p->d_packed,
                   &p->d_status, &p->h_total};
    for (Buf* b : bufs) b->release();
    delete p;
}

int hb_region_plan_enqueue(hb_region_plan* p, int stages, void* stream) {
    if (!p) return fail(HB_ERR_INVALID, "plan is null");
    std::lock_guard<std::mutex> lk(p->ctx->mu);  // reads ctx->world / tables: serialised with hb_set_world etc.
    HB_CUDA(cudaSetDevice(p->ctx->device));
    return plan_enqueue(p, stages, stream ? (cudaStream_t)stream : p->ctx->stream);
}

int hb_region_plan_launches(const hb_region_plan* p, int stages) {
    if (!p) return 0;
    int n = 0;
    if (stages & HB_STAGE_FUSED) {
        stages &= ~HB_STAGE_FUSED;
        if (region_fused_supported(p->ctx->world, p->dev)) { n += 1; stages &= ~HB_STAGE_ALL_MESH; }
        else stages |= HB_STAGE_ALL_MESH;
    }
    if (stages & HB_STAGE_HEIGHTS) n += 1;
    if (stages & HB_STAGE_COUNT) n += 1;
    if (stages & HB_STAGE_SCAN) n += scan_launch_count(p->n_chunks());
    if (stages & HB_STAGE_EMIT) n += 1;
    if (stages & HB_STAGE_EMIT_PACKED) n += 1;
    if (stages & HB_STAGE_FILL) n += 1;
    return n;
}

int hb_region_plan_result(hb_region_plan* p, void* stream, uint64_t* n_chunks, uint64_t* total_quads, uint64_t* total_span,
                          const hb_instance3d** d_instances, const uint64_t** d_offsets, const uint32_t** d_counts,
                          const uint16_t** d_ids, const int32_t** d_heights) {
    if (!p) return fail(HB_ERR_INVALID, "plan is null");
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    HB_CUDA(cudaSetDevice(p->ctx->device));
    uint64_t tot[4];
    int rc = plan_totals(p, stream ? (cudaStream_t)stream : p->ctx->stream, tot);
    if (rc != HB_OK) return rc;
    if (n_chunks) *n_chunks = p->n_chunks();
    if (total_span) *total_span = tot[0];
    if (total_quads) *total_quads = tot[1];
    if (d_instances) *d_instances = p->d_out.as<hb_instance3d>();
    if (d_offsets) *d_offsets = p->d_offsets.as<uint64_t>();
    if (d_counts) *d_counts = p->d_counts.as<uint32_t>();
    if (d_ids) *d_ids = p->d_ids.as<uint16_t>();
    if (d_heights) *d_heights = p->d_heights.as<int32_t>();
    if (tot[2] & 1) return fail(HB_ERR_CAPACITY, "instance buffer holds %llu, %llu required", (unsigned long long)p->capacity,
                                (unsigned long long)tot[0]);
    return HB_OK;
}

int hb_dev_download(hb_ctx* c, void* host_dst, const void* dev_src, size_t bytes) {
    if (!c || (bytes && (!host_dst || !dev_src))) return fail(HB_ERR_INVALID, "null argument");
    HB_CUDA(cudaSetDevice(c->device));
    HB_CUDA(cudaDeviceSynchronize());
    if (bytes) HB_CUDA(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return HB_OK;
}

}  // extern "C"

namespace {

enum StreamMode { kRawInstances = 0, kPackedExpand = 1, kPackedSink = 2 };

// plain host memory for the expanded slab: grow-only, 2 MiB aligned and huge-page advised so that the expander's
// stores do not pay a TLB miss every 4 KiB; the pages are touched once (first call of a given shape)
int reserve_plain(void** buf, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return HB_OK;
    free(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = ((bytes + bytes / 8) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    void* p = aligned_alloc((size_t)2 << 20, want);
    if (!p) return fail(HB_ERR_NOMEM, "out of host memory for a slab buffer (%zu bytes)", want);
    madvise(p, want, MADV_HUGEPAGE);
    *buf = p;
    *cap = want;
    return HB_OK;
}
int reserve_expanded(hb_ctx* c, size_t bytes) { return reserve_plain(&c->h_expanded, &c->h_expanded_cap, bytes); }

void ensure_expander(hb_ctx* c) {
    if (!c->xt_valid) {
        c->xt.build(c->tables, c->mats);
        c->xt_valid = true;
    }
    if (c->host_threads == 0) {
        c->host_threads = default_host_threads();
        c->pool.resize(c->host_threads);
    }
}

int stream_region(hb_ctx* c, const hb_region* region, int32_t ix_begin, int32_t ix_end, int want_ids, StreamMode mode,
                  hb_region_sink sink, hb_region_packed_sink psink, void* user, uint64_t* total_chunks, uint64_t* total_quads) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    int rc = region_check(region);
    if (rc != HB_OK) return rc;
    if (ix_begin < 0 || ix_end > region->ncx || ix_begin > ix_end) return fail(HB_ERR_INVALID, "bad slab [%d,%d)", ix_begin, ix_end);
    if (total_chunks) *total_chunks = 0;
    if (total_quads) *total_quads = 0;
    if (ix_begin == ix_end) return HB_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    const bool packed = mode != kRawInstances;
    // packed transports: the block ids are rebuilt on the host from the slab's heights tiles (fill_chunk_ids), so only
    // 4 KiB per chunk COLUMN cross the link instead of 64 KiB per chunk, and the device skips the FILL stage
    const bool host_ids = want_ids && packed;
    const int dev_ids = want_ids && !packed;
    if (mode == kPackedExpand || host_ids) ensure_expander(c);
    // slab width: aim at ~384 MiB of instances per slab (about 300 KB per column on this terrain),
    // bounded so that the optional id image (64 KiB/chunk) stays below ~1 GiB per slab.
    // Packed sink: only 4 B/quad leave the device, so the same budget is counted in packed bytes (8x wider slabs):
    // every slab costs a host synchronisation and three copies, and with 52 slabs of 14 MB that overhead kept the link
    // at 0.66 of its rate (measured on config 4: 192 / 384 / 768 / 1536 / 3072 / 6144 MiB -> 26.7 / 19.3 / 16.7 / 15.5 /
    // 14.9 / 14.7 ms per region; the expanding transport is host-write bound and flat between 384 and 768 MiB).
    const size_t per_col = 300 * 1024;
    const size_t slab_mib = mode == kPackedSink ? 3072 : 384;
    int32_t width = (int32_t)std::max<size_t>(1, (slab_mib << 20) / (per_col * (size_t)region->ncz));
    if (want_ids) {
        const size_t ids_per_ix = (size_t)region->ncz * region->ncy * HB_CHUNK_VOLUME * sizeof(uint16_t);
        width = (int32_t)std::max<size_t>(1, std::min<size_t>(width, ((size_t)1 << 30) / ids_per_ix));
    }
    width = std::min(width, ix_end - ix_begin);

    // the two plans (device buffers grow-only) and the events live in the ctx: after the first call of a given
    // shape nothing is allocated or freed here, which matters with several processes on one box (allocation
    // calls serialise in the driver and cudaFree synchronises the device)
    hb_region_plan* plan[2];
    for (int b = 0; b < 2; ++b) {
        if (!c->stream_plan[b]) {
            c->stream_plan[b] = new (std::nothrow) hb_region_plan();
            if (!c->stream_plan[b]) return fail(HB_ERR_NOMEM, "out of host memory");
            c->stream_plan[b]->ctx = c;
            HB_CUDA(cudaEventCreateWithFlags(&c->ev_kernels[b], cudaEventDisableTiming));
            HB_CUDA(cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
        }
        plan[b] = c->stream_plan[b];
        plan[b]->region = *region;
        plan[b]->want_ids = dev_ids;
        plan_target(plan[b], ix_begin, ix_begin + width);
        rc = plan_alloc(plan[b], width);
        if (rc != HB_OK) return rc;
    }
    cudaEvent_t* ev_kernels = c->ev_kernels;
    cudaEvent_t* ev_copied = c->ev_copied;
    struct Pending { bool live; uint64_t first, n, span; } pend[2] = {{false, 0, 0, 0}, {false, 0, 0, 0}};
    uint64_t sum_chunks = 0, sum_quads = 0;
    // hand slab `b` to the sink: wait for its copies; packed+expand writes the Instance3d records on the host threads
    // Handing slab `b` to the sink has two halves so that the host threads rebuild its records (and block ids) WHILE the
    // calling thread enqueues and sizes the next slab on the device: drain_begin waits for the slab's copies and starts
    // the workers, drain_end waits for them and calls the sink.
    uint16_t* host_ids_out = nullptr;
    auto drain_begin = [&](int b) -> int {
        if (!pend[b].live) return HB_OK;
        HB_CUDA(cudaEventSynchronize(ev_copied[b]));
        const bool want_expand = mode == kPackedExpand && sink;
        const bool want_fill = host_ids && (sink || psink);
        if (!want_expand && !want_fill) return HB_OK;
        if (want_fill) {
            const int rc2 = reserve_plain(&c->h_ids_plain, &c->h_ids_plain_cap, pend[b].n * HB_CHUNK_VOLUME * sizeof(uint16_t));
            if (rc2 != HB_OK) return rc2;
        }
        if (want_expand) {
            const int rc2 = reserve_expanded(c, std::max<uint64_t>(pend[b].span, 4) * sizeof(hb_instance3d));
            if (rc2 != HB_OK) return rc2;
        }
        host_ids_out = want_fill ? static_cast<uint16_t*>(c->h_ids_plain) : nullptr;
        const uint64_t* off = c->h_off[b].as<uint64_t>();
        const uint32_t* cnt = c->h_cnt[b].as<uint32_t>();
        const uint32_t* packed_words = c->h_packed[b].as<uint32_t>();
        hb_instance3d* inst = static_cast<hb_instance3d*>(c->h_expanded);
        uint16_t* out_ids = host_ids_out;
        const int32_t* hts = c->h_heights[b].as<int32_t>();
        const hb_region r = *region;
        const uint64_t first = pend[b].first, per_ix = (uint64_t)r.ncz * r.ncy;
        const ExpandTables* T = &c->xt;
        c->pool.start(pend[b].n, 8, [=](size_t i) {  // slab-relative chunk i = column i / ncy, layer i % ncy
            if (want_expand && cnt[i]) {
                const uint64_t g = first + i;
                const hb_chunk_pt pt{r.cx0 + (int32_t)(g / per_ix), r.cy0 + (int32_t)(g % r.ncy), r.cz0 + (int32_t)((g / r.ncy) % r.ncz)};
                expand_chunk(*T, pt, packed_words + off[i], cnt[i], inst + off[i]);
            }
            if (want_fill) fill_chunk_ids(hts + (i / r.ncy) * HB_CHUNK_AREA, r.cy0 + (int32_t)(i % r.ncy), out_ids + i * HB_CHUNK_VOLUME);
        });
        return HB_OK;
    };
    auto drain_end = [&](int b) -> int {
        if (!pend[b].live) return HB_OK;
        pend[b].live = false;
        c->pool.wait();
        const uint64_t* off = c->h_off[b].as<uint64_t>();
        const uint32_t* cnt = c->h_cnt[b].as<uint32_t>();
        const uint16_t* ids = dev_ids ? c->h_ids[b].as<uint16_t>() : (host_ids && (sink || psink) ? host_ids_out : nullptr);
        if (mode == kPackedSink) {
            if (psink) psink(user, pend[b].first, pend[b].n, off, cnt, c->h_packed[b].as<hb_packed_quad>(), pend[b].span, ids);
        } else if (mode == kPackedExpand) {
            if (sink) sink(user, pend[b].first, pend[b].n, off, cnt, static_cast<hb_instance3d*>(c->h_expanded), pend[b].span, ids);
        } else if (sink) {
            sink(user, pend[b].first, pend[b].n, off, cnt, c->h_inst[b].as<hb_instance3d>(), pend[b].span, ids);
        }
        return HB_OK;
    };
    auto drain = [&](int b) -> int {
        const int rc2 = drain_begin(b);
        return rc2 != HB_OK ? rc2 : drain_end(b);
    };
    int slab = 0;
    for (int32_t ix0 = ix_begin; ix0 < ix_end && rc == HB_OK; ix0 += width, ++slab) {
        const int b = slab & 1;
        hb_region_plan* p = plan[b];
        rc = drain(b);  // buffers of parity b are free again (device out + pinned)
        if (rc != HB_OK) break;
        rc = drain_begin(b ^ 1);  // the previous slab's records are rebuilt by the workers during the device calls below
        if (rc != HB_OK) break;
        plan_target(p, ix0, std::min(ix_end, ix0 + width));
        const size_t nch = p->n_chunks();
        rc = plan_enqueue(p, HB_STAGE_HEIGHTS | HB_STAGE_COUNT | HB_STAGE_SCAN, c->stream);
        if (rc != HB_OK) break;
        uint64_t tot[4];
        rc = plan_totals(p, c->stream, tot);
        if (rc != HB_OK) break;
        p->capacity = tot[0];
        const uint64_t span4 = std::max<uint64_t>(tot[0], 4);
        cudaError_t e = cudaSuccess;
        if (packed) e = c->h_packed[b].reserve(span4 * sizeof(hb_packed_quad));
        else {
            e = p->d_out.reserve(span4 * sizeof(hb_instance3d));
            if (e == cudaSuccess) e = c->h_inst[b].reserve(span4 * sizeof(hb_instance3d));
        }
        if (e == cudaSuccess) e = c->h_off[b].reserve(nch * sizeof(uint64_t));
        if (e == cudaSuccess) e = c->h_cnt[b].reserve(nch * sizeof(uint32_t));
        if (e == cudaSuccess && dev_ids) e = c->h_ids[b].reserve(nch * HB_CHUNK_VOLUME * sizeof(uint16_t));
        const size_t slab_cols = (size_t)(p->dev.ix_end - p->dev.ix_begin) * region->ncz;
        if (e == cudaSuccess && host_ids) e = c->h_heights[b].reserve(slab_cols * HB_CHUNK_AREA * sizeof(int32_t));
        if (e != cudaSuccess) { rc = fail(HB_ERR_CUDA, "slab buffer allocation failed: %s", cudaGetErrorString(e)); break; }
        rc = plan_enqueue(p, (packed ? HB_STAGE_EMIT_PACKED : HB_STAGE_EMIT) | (dev_ids ? HB_STAGE_FILL : 0), c->stream);
        if (rc != HB_OK) break;
        cudaEventRecord(ev_kernels[b], c->stream);
        cudaStreamWaitEvent(c->copy, ev_kernels[b], 0);
        cudaMemcpyAsync(c->h_off[b].p, p->d_offsets.p, nch * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->copy);
        cudaMemcpyAsync(c->h_cnt[b].p, p->d_counts.p, nch * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy);
        if (tot[0]) {
            if (packed)
                cudaMemcpyAsync(c->h_packed[b].p, p->d_packed.p, tot[0] * sizeof(hb_packed_quad), cudaMemcpyDeviceToHost, c->copy);
            else
                cudaMemcpyAsync(c->h_inst[b].p, p->d_out.p, tot[0] * sizeof(hb_instance3d), cudaMemcpyDeviceToHost, c->copy);
        }
        if (dev_ids)
            cudaMemcpyAsync(c->h_ids[b].p, p->d_ids.p, nch * HB_CHUNK_VOLUME * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->copy);
        if (host_ids)  // the slab's own columns inside the heights buffer (which also holds the one-column apron)
            cudaMemcpyAsync(c->h_heights[b].p,
                            p->d_heights.as<int32_t>() + (size_t)(p->dev.ix_begin - p->dev.hx_begin) * region->ncz * HB_CHUNK_AREA,
                            slab_cols * HB_CHUNK_AREA * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy);
        cudaError_t ce = cudaEventRecord(ev_copied[b], c->copy);
        if (ce != cudaSuccess) { rc = fail(HB_ERR_CUDA, "slab copy failed: %s", cudaGetErrorString(ce)); break; }
        pend[b] = Pending{true, (uint64_t)ix0 * region->ncz * region->ncy, nch, tot[0]};
        sum_chunks += nch;
        sum_quads += tot[1];
        // hand the previous slab to the sink while this one is computed and copied
        rc = drain_end(b ^ 1);
    }
    c->pool.wait();  // an error path may have left the loop between drain_begin and drain_end
    if (rc == HB_OK) rc = drain((slab & 1));
    if (rc == HB_OK) rc = drain((slab & 1) ^ 1);
    cudaStreamSynchronize(c->copy);
    cudaStreamSynchronize(c->stream);
    if (rc == HB_OK) {
        if (total_chunks) *total_chunks = sum_chunks;
        if (total_quads) *total_quads = sum_quads;
    }
    return rc;
}

}  // namespace

extern "C" {

int hb_set_region_transport(hb_ctx* c, int32_t transport) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    if (transport != HB_TRANSPORT_PACKED && transport != HB_TRANSPORT_INSTANCES) return fail(HB_ERR_INVALID, "unknown transport %d", transport);
    std::lock_guard<std::mutex> lk(c->mu);
    c->transport = transport;
    return HB_OK;
}

int hb_set_host_threads(hb_ctx* c, int32_t n) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    if (n < 1 || n > 256) return fail(HB_ERR_INVALID, "host threads must be in 1..256");
    std::lock_guard<std::mutex> lk(c->mu);
    c->host_threads = n;
    c->pool.resize(n);
    return HB_OK;
}

int hb_generate_mesh_region(hb_ctx* c, const hb_region* region, int want_ids, hb_region_sink sink, void* user,
                            uint64_t* total_chunks, uint64_t* total_quads) {
    if (!region) return fail(HB_ERR_INVALID, "region is null");
    return hb_generate_mesh_region_slab(c, region, 0, region->ncx, want_ids, sink, user, total_chunks, total_quads);
}

int hb_generate_mesh_region_slab(hb_ctx* c, const hb_region* region, int32_t ix_begin, int32_t ix_end, int want_ids,
                                 hb_region_sink sink, void* user, uint64_t* total_chunks, uint64_t* total_quads) {
    if (!c) return fail(HB_ERR_INVALID, "ctx is null");
    const StreamMode mode = c->transport == HB_TRANSPORT_INSTANCES ? kRawInstances : kPackedExpand;
    return stream_region(c, region, ix_begin, ix_end, want_ids, mode, sink, nullptr, user, total_chunks, total_quads);
}

int hb_generate_mesh_region_packed(hb_ctx* c, const hb_region* region, int32_t ix_begin, int32_t ix_end, int want_ids,
                                   hb_region_packed_sink sink, void* user, uint64_t* total_chunks, uint64_t* total_quads) {
    return stream_region(c, region, ix_begin, ix_end, want_ids, kPackedSink, nullptr, sink, user, total_chunks, total_quads);
}

int hb_expand_quads(hb_ctx* c, hb_chunk_pt pt, const hb_packed_quad* quads, uint32_t n, hb_instance3d* out) {
    if (!c || (n && (!quads || !out))) return fail(HB_ERR_INVALID, "null argument");
    if ((uintptr_t)out & 15u) return fail(HB_ERR_INVALID, "out must be 16-byte aligned");
    if (!c->xt_valid) {  // tables are rebuilt under the lock; afterwards concurrent callers only read them
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->xt_valid) {
            c->xt.build(c->tables, c->mats);
            c->xt_valid = true;
        }
    }
    expand_chunk(c->xt, pt, quads, n, out);
    return HB_OK;
}

// ---- in-run ceilings for the end-to-end numbers (measurement helpers; they move bytes, nothing else) ----

int hb_probe_d2h(hb_ctx* c, size_t bytes, int32_t reps, double* gbs) {
    if (!c || !gbs || bytes == 0 || reps < 1) return fail(HB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    HB_CUDA(c->d_out.reserve(bytes));
    HB_CUDA(c->h_inst[0].reserve(bytes));
    cudaEvent_t e0, e1;
    HB_CUDA(cudaEventCreate(&e0));
    HB_CUDA(cudaEventCreate(&e1));
    cudaMemcpyAsync(c->h_inst[0].p, c->d_out.p, bytes, cudaMemcpyDeviceToHost, c->copy);  // warm-up
    cudaEventRecord(e0, c->copy);
    for (int r = 0; r < reps; ++r) cudaMemcpyAsync(c->h_inst[0].p, c->d_out.p, bytes, cudaMemcpyDeviceToHost, c->copy);
    cudaEventRecord(e1, c->copy);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return fail(HB_ERR_CUDA, "D2H probe failed: %s", cudaGetErrorString(e));
    *gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
    return HB_OK;
}

int hb_probe_host_write(hb_ctx* c, size_t bytes, int32_t reps, double* gbs) {
    if (!c || !gbs || bytes < ((size_t)1 << 20) || reps < 1) return fail(HB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    ensure_expander(c);
    int rc = reserve_expanded(c, bytes);
    if (rc != HB_OK) return rc;
    char* base = static_cast<char*>(c->h_expanded);
    const size_t piece = (size_t)1 << 20, np = bytes / piece;
    auto fill = [&](size_t i) {
        const __m128i v = _mm_set1_epi32((int)i);
        char* p = base + i * piece;
        for (size_t o = 0; o < piece; o += 64) {
            _mm_stream_si128((__m128i*)(p + o), v);
            _mm_stream_si128((__m128i*)(p + o + 16), v);
            _mm_stream_si128((__m128i*)(p + o + 32), v);
            _mm_stream_si128((__m128i*)(p + o + 48), v);
        }
        _mm_sfence();
    };
    c->pool.run(np, 4, fill);  // touches the pages
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) c->pool.run(np, 4, fill);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *gbs = (double)np * piece * reps / dt / 1e9;
    return HB_OK;
}

// ---- the host half of the packed transports without a context (no CUDA call is made: usable, and tested, on a
// machine without a GPU) ----
int hb_host_expand_quads(hb_chunk_pt pt, const hb_packed_quad* quads, uint32_t n, int32_t n_materials, const int32_t* n_colors,
                         const float* rgba, hb_instance3d* out) {
    if (n && (!quads || !out)) return fail(HB_ERR_INVALID, "null argument");
    if ((uintptr_t)out & 15u) return fail(HB_ERR_INVALID, "out must be 16-byte aligned");
    MaterialsDev m;
    if (!n_colors || !rgba) {
        default_materials(&m);
    } else {
        if (n_materials < 1 || n_materials > HB_MAX_MATERIALS) return fail(HB_ERR_INVALID, "n_materials must be in 1..%d", HB_MAX_MATERIALS);
        memset(&m, 0, sizeof(m));
        m.n_materials = n_materials;
        for (int i = 0; i < n_materials; ++i) {
            if (n_colors[i] < 1 || n_colors[i] > HB_MAX_COLORS) return fail(HB_ERR_INVALID, "material %d: n_colors must be in 1..%d", i + 1, HB_MAX_COLORS);
            m.n_colors[i] = n_colors[i];
            memcpy(m.rgba[i], rgba + (size_t)i * HB_MAX_COLORS * 4, sizeof(float) * HB_MAX_COLORS * 4);
        }
    }
    MeshTables tb;
    for (int f = 0; f < 6; ++f) face_matrix(f, tb.face_mat[f]);
    for (int k = 0; k < 4; ++k) tb.ao_lut[k] = ao_value(k);
    ExpandTables T;
    T.build(tb, m);
    expand_chunk(T, pt, quads, n, out);
    return HB_OK;
}

int hb_host_fill_ids(const int32_t* heights, int32_t cy, uint16_t* out_ids) {
    if (!heights || !out_ids) return fail(HB_ERR_INVALID, "null argument");
    if ((uintptr_t)out_ids & 15u) return fail(HB_ERR_INVALID, "out_ids must be 16-byte aligned");
    fill_chunk_ids(heights, cy, out_ids);
    return HB_OK;
}

int hb_encode_palette(hb_ctx* c, const hb_material_desc* descs, uint8_t* out, size_t capacity, size_t* n_bytes) {
    if (!c || !n_bytes) return fail(HB_ERR_INVALID, "null argument");
    *n_bytes = 0;
    static const hb_material_desc kGame[3] = {{"herbolution", "stone", 0x3F, 1, 10.0f},
                                              {"herbolution", "dirt", 0x3F, 1, 0.95f},
                                              {"herbolution", "grass", 0x3F, 1, 1.05f}};
    std::lock_guard<std::mutex> lk(c->mu);
    const MaterialsDev& m = c->mats;
    if (!descs) {
        if (m.n_materials != 3) return fail(HB_ERR_INVALID, "descs is null but the ctx holds %d materials (the default palette has 3)", m.n_materials);
        descs = kGame;
    }
    std::vector<uint8_t> buf;
    auto put = [&](const void* p, size_t n) { buf.insert(buf.end(), static_cast<const uint8_t*>(p), static_cast<const uint8_t*>(p) + n); };
    for (int i = 0; i < m.n_materials; ++i) {
        const hb_material_desc& d = descs[i];
        if (!d.group || !d.key) return fail(HB_ERR_INVALID, "material %d: group/key is null", i + 1);
        const size_t gl = strlen(d.group), kl = strlen(d.key);
        if (gl > 255 || kl > 255) return fail(HB_ERR_INVALID, "material %d: group/key longer than 255 bytes", i + 1);
        const uint16_t idx = (uint16_t)i;  // little-endian hosts only (as the rest of the library)
        put(&idx, 2);
        const uint8_t lens[2] = {(uint8_t)gl, (uint8_t)kl};
        put(lens, 2);
        put(d.group, gl);
        put(d.key, kl);
        const uint8_t e0 = (uint8_t)((uint8_t)(d.cullable_faces << 6) | (d.has_collider ? 1u : 0u));  // u8 shift, material.rs:79
        const uint8_t tex = 0;
        put(&e0, 1);
        put(&tex, 1);
        const uint16_t nc = (uint16_t)m.n_colors[i];
        put(&nc, 2);
        put(m.rgba[i], (size_t)m.n_colors[i] * 4 * sizeof(float));
        put(&d.toughness, sizeof(float));
    }
    *n_bytes = buf.size();
    if (buf.size() > capacity || (!out && !buf.empty()))
        return fail(HB_ERR_CAPACITY, "palette needs %zu bytes, %zu given", buf.size(), capacity);
    if (!buf.empty()) memcpy(out, buf.data(), buf.size());
    return HB_OK;
}

int hb_region_plan_packed(hb_region_plan* p, const hb_packed_quad** d_packed) {
    if (!p || !d_packed) return fail(HB_ERR_INVALID, "null argument");
    *d_packed = p->d_packed.as<hb_packed_quad>();
    return HB_OK;
}

}  // extern "C"
