// world.cu -- device-resident ChunkMap (include/herbgpu.h, hb_world_*): the server's chunk store with its
// PaletteCube arrays in HBM, bulk load (generate + cull_shared_faces), the incremental dig/place path
// (ChunkMap::set_cube), the CubeUpdate wire format and the client's remesh-what-changed loop.
//
// Per slot in HBM: cubes[32768] (reference layout, 8 B), two 32768-bit sets over the voxel index (word = column
// x*32+z, bit = y -- the same layout as the mesher's occupancy bitmap): `touched` = CubeMesh::updated_positions as a
// set, `pending` = what the last sync drained; plus the mesher's metadata (occupancy bitmap, six face planes,
// summary), the chunk position and the 27-entry neighbour row.  The position -> slot map lives on the host, like
// the reference's HashMap<ChunkPt, Chunk>; kernels only ever see slots.
#include <string.h>

#include <algorithm>
#include <new>
#include <unordered_map>
#include <vector>

#include "host_common.cuh"

using namespace hb;

namespace {

constexpr int kWT = 256;

// CubeFlags (server/src/chunk/cube.rs:35-49,69-72): faces() is empty unless bit 0 is set; insert/remove go through
// set_opaque, which always sets bit 0.
__device__ __forceinline__ uint32_t flags_faces(uint32_t v) { return (v & 1u) ? ((v >> 1) & 63u) : 0u; }
__device__ __forceinline__ uint32_t flags_remove(uint32_t v, int f) { return ((flags_faces(v) & ~(1u << f)) << 1) | 1u; }
__device__ __forceinline__ uint32_t flags_insert(uint32_t v, int f) { return ((flags_faces(v) | (1u << f)) << 1) | 1u; }

// several culls of one load can hit the same edge voxel from different CTAs
__device__ __forceinline__ void atomic_remove_face(uint32_t* p, int f) {
    uint32_t old = *p;
    for (;;) {
        const uint32_t nv = flags_remove(old, f);
        if (nv == old) return;
        const uint32_t seen = atomicCAS(p, old, nv);
        if (seen == old) return;
        old = seen;
    }
}

struct CullPair {
    int32_t a, b, face, _pad;  // chunk b is a's neighbour in direction `face`
};

// CubeMesh::cull_shared_faces (server/src/chunk/mesh.rs:76-123): over the 32x32 pairs of facing border voxels,
// if both have a material the shared face is removed on both and both positions are pushed to updated_positions.
__global__ void __launch_bounds__(kWT) k_world_cull(hb_palette_cube* __restrict__ cubes, uint32_t* __restrict__ touched,
                                                    const CullPair* __restrict__ pairs) {
    const CullPair p = pairs[blockIdx.x];
    const int axis = p.face >> 1;
    const int fa = (p.face & 1) ? 0 : 31, fb = 31 - fa;  // mesh.rs:254-264 boundary()
    hb_palette_cube* ca = cubes + (size_t)p.a * HB_CHUNK_VOLUME;
    hb_palette_cube* cb = cubes + (size_t)p.b * HB_CHUNK_VOLUME;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = k * kWT + threadIdx.x, u = i >> 5, v = i & 31;
        int xa, ya, za, xb, yb, zb;
        if (axis == 0) { xa = fa; xb = fb; ya = yb = v; za = zb = u; }       // v = y keeps a warp on one 256-byte run
        else if (axis == 1) { ya = fa; yb = fb; xa = xb = u; za = zb = v; }
        else { za = fa; zb = fb; xa = xb = u; ya = yb = v; }
        const int La = xa * 1024 + za * 32 + ya, Lb = xb * 1024 + zb * 32 + yb;
        if (ca[La].material != 0 && cb[Lb].material != 0) {
            atomic_remove_face(&ca[La].flags, p.face);
            atomic_remove_face(&cb[Lb].flags, p.face ^ 1);
            atomicOr(touched + (size_t)p.a * HB_CHUNK_AREA + xa * 32 + za, 1u << ya);
            atomicOr(touched + (size_t)p.b * HB_CHUNK_AREA + xb * 32 + zb, 1u << yb);
        }
    }
}

struct EditRec {
    int32_t slot;     // own chunk, -1 = not loaded
    int32_t nbr[6];   // adjacent chunk across -x, +x, -y, +y, -z, +z; filled only where the voxel is on that border
    uint16_t x, y, z, material;
};

// ChunkMap::set_cube (server/src/chunk/map.rs:81-151) + CubeMesh::set (server/src/chunk/mesh.rs:125-186).  One warp
// applies the edits of one GROUP one after the other, in the caller's order: lane e < 6 owns direction e, the six
// neighbour voxels are distinct words, and the centre voxel's face mask is combined with a ballot.  An edit reads and
// writes only its own voxel and the six voxels next to it (in the chunk or on the adjacent chunk's border), so edits
// whose 7-voxel stencils share no voxel commute; the host partitions a batch into the connected components of
// "stencils intersect" (hb_world_set_cubes) and every component is a group, i.e. groups run concurrently on
// different warps while the reference's sequential semantics are preserved inside each.
constexpr int kEditWarps = 4;
__global__ void __launch_bounds__(32 * kEditWarps) k_world_edit(hb_palette_cube* __restrict__ cubes, uint32_t* __restrict__ touched,
                                                                const EditRec* __restrict__ recs,
                                                                const uint32_t* __restrict__ group_start, uint32_t n_groups) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x * kEditWarps + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const size_t i_begin = group_start[g], i_end = group_start[g + 1];
    for (size_t i = i_begin; i < i_end; ++i) {
        __syncwarp();  // the previous edit's stores (any lane) are visible to every lane
        const EditRec r = recs[i];
        const int l[3] = {r.x, r.y, r.z};
        // map.rs:85-140: the neighbouring chunk's border voxel gets the face pointing at us, unconditionally.
        // Edge order (-x, EAST) (+x, WEST) (-y, UP) (+y, DOWN) (-z, NORTH) (+z, SOUTH): direction e inserts face e.
        if (lane < 6) {
            const int axis = lane >> 1, neg = !(lane & 1);
            const bool on_border = neg ? l[axis] == 0 : l[axis] == 31;
            if (on_border && r.nbr[lane] >= 0) {
                int nl[3] = {l[0], l[1], l[2]};
                nl[axis] = neg ? 31 : 0;
                hb_palette_cube* c = cubes + (size_t)r.nbr[lane] * HB_CHUNK_VOLUME + nl[0] * 1024 + nl[2] * 32 + nl[1];
                c->flags = flags_insert(c->flags, lane);
                atomicOr(touched + (size_t)r.nbr[lane] * HB_CHUNK_AREA + nl[0] * 32 + nl[2], 1u << nl[1]);
            }
        }
        if (r.slot < 0) continue;  // map.rs:142: own chunk not loaded
        hb_palette_cube* base = cubes + (size_t)r.slot * HB_CHUNK_VOLUME;
        hb_palette_cube* me = base + l[0] * 1024 + l[2] * 32 + l[1];
        uint32_t* tch = touched + (size_t)r.slot * HB_CHUNK_AREA;
        const uint32_t old_mat = me->material, my_flags = me->flags;
        __syncwarp();
        if (old_mat == r.material) continue;  // mesh.rs:129-131
        // in-chunk neighbour of lane's direction (FaceIter order E, W, U, D, N, S = +x, -x, +y, -y, +z, -z)
        int nx = l[0], ny = l[1], nz = l[2];
        bool inside = false;
        if (lane < 6) {
            const int d = (lane & 1) ? -1 : 1;
            if ((lane >> 1) == 0) nx += d; else if ((lane >> 1) == 1) ny += d; else nz += d;
            inside = nx >= 0 && ny >= 0 && nz >= 0 && nx < 32 && ny < 32 && nz < 32;
        }
        hb_palette_cube* nb = base + nx * 1024 + nz * 32 + ny;
        uint32_t new_flags = my_flags;
        if (old_mat == 0 && r.material != 0) {
            // mesh.rs:135-145 + remove_neighboring_faces (:209-232)
            const uint32_t missing = ~flags_faces(my_flags) & 63u;
            bool nb_solid = false;
            if (inside && ((missing >> lane) & 1u)) {
                atomicOr(tch + nx * 32 + nz, 1u << ny);
                nb_solid = nb->material != 0;
                nb->flags = flags_remove(nb->flags, lane ^ 1);
            }
            const uint32_t removed = __ballot_sync(0xffffffffu, nb_solid) & 63u;
            new_flags = ((63u & ~removed) << 1) | 1u;
        } else if (old_mat != 0 && r.material == 0) {
            // mesh.rs:147-157 + add_neighboring_faces (:234-251), which walks faces.not() as written
            const uint32_t present = flags_faces(my_flags);
            if (inside && !((present >> lane) & 1u)) {
                atomicOr(tch + nx * 32 + nz, 1u << ny);
                nb->flags = flags_insert(nb->flags, lane ^ 1);
            }
            new_flags = 1u;  // set_opaque(none)
        }
        if (lane == 0) {
            me->material = r.material;
            me->flags = new_flags;
            atomicOr(tch + l[0] * 32 + l[2], 1u << l[1]);  // mesh.rs:185
        }
    }
}

// Chunk::sync_with_client, drain half (server/src/chunk/mod.rs:35-67): pending <- touched, touched <- {}, and the
// number of positions drained per chunk.
__global__ void __launch_bounds__(kWT) k_world_drain(uint32_t* __restrict__ touched, uint32_t* __restrict__ pending,
                                                     const int32_t* __restrict__ list, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_part[kWT / 32];
    const size_t slot = (size_t)list[blockIdx.x];
    uint4* t = reinterpret_cast<uint4*>(touched + slot * HB_CHUNK_AREA) + threadIdx.x;
    const uint4 v = *t;
    *t = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(pending + slot * HB_CHUNK_AREA)[threadIdx.x] = v;
    uint32_t c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < kWT / 32; ++i) tot += s_part[i];
        counts[blockIdx.x] = tot;
    }
}

// CubeUpdate { overwrites: [ChunkCube] } (server/src/chunk/handle.rs:19-32): one 12-byte entry per drained position,
// ascending voxel index, carrying the cube as it is now (the reference reads mesh.data at drain time too).
__global__ void __launch_bounds__(kWT) k_world_updates(const hb_palette_cube* __restrict__ cubes, const uint32_t* __restrict__ pending,
                                                       const int32_t* __restrict__ list, const uint64_t* __restrict__ offsets,
                                                       uint32_t* __restrict__ out_words) {
    __shared__ uint32_t s_warp[kWT / 32];
    const size_t slot = (size_t)list[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 v = reinterpret_cast<const uint4*>(pending + slot * HB_CHUNK_AREA)[tid];  // columns 4*tid .. 4*tid+3
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t mine = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = incl - mine;
    for (int i = 0; i < warp; ++i) before += s_warp[i];
    const hb_palette_cube* src = cubes + slot * HB_CHUNK_VOLUME;
    uint32_t* dst = out_words + (offsets[blockIdx.x] + before) * 3ULL;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = tid * 4 + k, x = c >> 5, z = c & 31;
        uint32_t m = w[k];
        while (m) {
            const int y = __ffs(m) - 1;
            m &= m - 1;
            const uint2 cube = *reinterpret_cast<const uint2*>(src + c * 32 + y);
            dst[0] = 1u | ((uint32_t)x << 11) | ((uint32_t)y << 6) | ((uint32_t)z << 1);  // vec3u5, upper half = pad
            dst[1] = cube.x;
            dst[2] = cube.y;
            dst += 3;
        }
    }
}

// (dx+1)*9 + (dy+1)*3 + (dz+1): create_chunk_shell order (client/src/world/chunk.rs:158-174)
inline int shell_index(int dx, int dy, int dz) { return (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1); }

const int kNormal[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};  // face.rs:47-58

// Position -> slot map of the loaded chunks (the reference's HashMap<ChunkPt, Chunk>): open addressing over packed
// 3 x 21-bit keys.  A loader update asks for ~120 000 lookups (27 neighbours per new chunk), which is the host side
// of hb_world_load; std::unordered_map made that 4 ms for 4 235 chunks.
class SlotTable {
   public:
    void init(size_t capacity) {
        size_t sz = 64;
        while (sz < capacity * 4) sz <<= 1;
        keys_.assign(sz, kEmpty);
        vals_.assign(sz, -1);
        mask_ = sz - 1;
        count_ = filled_ = 0;
    }
    size_t size() const { return count_; }
    int32_t find(int32_t x, int32_t y, int32_t z) const {
        if (!in_range(x) || !in_range(y) || !in_range(z)) return -1;  // cannot be loaded; and must not alias a packed key
        const uint64_t k = pack(x, y, z);
        for (size_t i = hash(k) & mask_;; i = (i + 1) & mask_) {
            if (keys_[i] == k) return vals_[i];
            if (keys_[i] == kEmpty) return -1;
        }
    }
    void insert(int32_t x, int32_t y, int32_t z, int32_t v) {  // key must be absent and in range
        if ((filled_ + 1) * 2 > keys_.size()) rehash();
        const uint64_t k = pack(x, y, z);
        size_t i = hash(k) & mask_;
        while (keys_[i] != kEmpty && keys_[i] != kTomb) i = (i + 1) & mask_;
        if (keys_[i] == kEmpty) ++filled_;
        keys_[i] = k;
        vals_[i] = v;
        ++count_;
    }
    void erase(int32_t x, int32_t y, int32_t z) {
        if (!in_range(x) || !in_range(y) || !in_range(z)) return;
        const uint64_t k = pack(x, y, z);
        for (size_t i = hash(k) & mask_;; i = (i + 1) & mask_) {
            if (keys_[i] == k) { keys_[i] = kTomb; vals_[i] = -1; --count_; return; }
            if (keys_[i] == kEmpty) return;
        }
    }

   private:
    static constexpr uint64_t kEmpty = ~0ULL, kTomb = ~0ULL - 1;
    static bool in_range(int32_t c) { return c >= -(1 << 20) && c < (1 << 20); }
    static uint64_t pack(int32_t x, int32_t y, int32_t z) {
        return ((uint64_t)((uint32_t)x & 0x1FFFFFu)) | ((uint64_t)((uint32_t)y & 0x1FFFFFu) << 21) | ((uint64_t)((uint32_t)z & 0x1FFFFFu) << 42);
    }
    static size_t hash(uint64_t k) {
        k ^= k >> 31; k *= 0x9E3779B97F4A7C15ULL; k ^= k >> 29; k *= 0xBF58476D1CE4E5B9ULL; k ^= k >> 32;
        return (size_t)k;
    }
    void rehash() {  // drops the tombstones; the table never grows beyond what init() sized for `capacity` live keys
        std::vector<uint64_t> ok;
        std::vector<int32_t> ov;
        ok.swap(keys_);
        ov.swap(vals_);
        keys_.assign(ok.size(), kEmpty);
        vals_.assign(ok.size(), -1);
        count_ = filled_ = 0;
        for (size_t i = 0; i < ok.size(); ++i)
            if (ok[i] != kEmpty && ok[i] != kTomb) {
                size_t j = hash(ok[i]) & mask_;
                while (keys_[j] != kEmpty) j = (j + 1) & mask_;
                keys_[j] = ok[i];
                vals_[j] = ov[i];
                ++count_;
                ++filled_;
            }
    }
    std::vector<uint64_t> keys_;
    std::vector<int32_t> vals_;
    size_t mask_ = 0, count_ = 0, filled_ = 0;
};

}  // namespace

struct hb_world {
    hb_ctx* ctx = nullptr;
    size_t capacity = 0;
    // per-slot device state
    Buf d_cubes, d_touched, d_pending, d_bitmaps, d_summary, d_planes, d_pts, d_nbr;
    // scratch
    Buf d_a, d_b, d_c, d_d, d_heights, d_counts, d_offsets, d_scan, d_total, d_out;
    Buf h_pin;
    // host mirror
    SlotTable map;
    std::vector<int32_t> free_slots;
    std::vector<hb_chunk_pt> slot_pt;
    std::vector<uint8_t> used;
    std::vector<uint32_t> epoch;   // bumped when a slot is freed: invalidates sync records
    std::vector<int32_t> h_nbr;    // capacity x 27
    bool tables_dirty = true;      // d_pts / d_nbr behind the host mirror
    std::vector<uint8_t> stale;    // mesher metadata (bitmaps/planes/summary) behind the cubes
    std::vector<uint8_t> touched;  // updated_positions may be non-empty
    std::vector<uint8_t> remesh;   // received a CubeUpdate since the last remesh
    struct SyncRec { int32_t slot; uint32_t epoch; uint32_t count; };
    std::vector<SyncRec> last_sync;
    uint32_t last_edit_groups = 0;  // conflict groups of the last hb_world_set_cubes batch (diagnostics)
    hb_world() { h_pin.pinned = true; }
    int find(int32_t x, int32_t y, int32_t z) const { return map.find(x, y, z); }
};

namespace {

int upload(cudaStream_t s, Buf& b, const void* src, size_t bytes) {
    HB_CUDA(b.reserve(bytes ? bytes : 16));
    if (bytes) HB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    return HB_OK;
}

int push_tables(hb_world* w, cudaStream_t s) {
    if (!w->tables_dirty) return HB_OK;
    HB_CUDA(cudaMemcpyAsync(w->d_pts.p, w->slot_pt.data(), w->capacity * sizeof(hb_chunk_pt), cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemcpyAsync(w->d_nbr.p, w->h_nbr.data(), w->capacity * 27 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    w->tables_dirty = false;
    return HB_OK;
}

// set (or clear, slot = -1) `slot` in the neighbour rows of the chunks around p, and fill slot's own row
void wire_neighbours(hb_world* w, const hb_chunk_pt& p, int32_t slot, bool present) {
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                const int o = (dx || dy || dz) ? w->find(p.x + dx, p.y + dy, p.z + dz) : slot;
                if (present) w->h_nbr[(size_t)slot * 27 + shell_index(dx, dy, dz)] = o;
                if (o >= 0 && o != slot) w->h_nbr[(size_t)o * 27 + shell_index(-dx, -dy, -dz)] = present ? slot : -1;
            }
    if (present) w->h_nbr[(size_t)slot * 27 + 13] = slot;
    w->tables_dirty = true;
}

std::vector<int32_t> slots_where(const hb_world* w, const std::vector<uint8_t>& flag) {
    std::vector<int32_t> v;
    for (size_t s = 0; s < w->capacity; ++s)
        if (flag[s] && w->used[s]) v.push_back((int32_t)s);
    return v;
}

// client ChunkMap::update: re-pack what changed, count, scan, emit.  Leaves the instances in w->d_out.
int remesh_impl(hb_world* w, hb_chunk_pt* out_pts, size_t pts_capacity, uint64_t out_capacity, bool host_out,
                uint64_t* out_offsets, uint32_t* out_counts, size_t* n_chunks, uint64_t* out_total,
                std::vector<int32_t>* list_out) {
    hb_ctx* c = w->ctx;
    cudaStream_t s = c->stream;
    const std::vector<int32_t> list = slots_where(w, w->remesh);
    const size_t n = list.size();
    *n_chunks = n;
    *out_total = 0;
    if (n == 0) {
        std::fill(w->remesh.begin(), w->remesh.end(), 0);
        return HB_OK;
    }
    int rc = push_tables(w, s);
    if (rc != HB_OK) return rc;
    HB_CUDA(w->d_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(w->h_pin.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(cudaMemsetAsync(w->d_total.p, 0, 4 * sizeof(uint64_t), s));
    const std::vector<int32_t> st = slots_where(w, w->stale);
    if (!st.empty()) {
        if ((rc = upload(s, w->d_a, st.data(), st.size() * sizeof(int32_t))) != HB_OK) return rc;
        HB_CUDA(launch_pack_cubes(s, w->d_cubes.as<hb_palette_cube>(), st.size(), c->mats.n_materials, w->d_bitmaps.as<uint32_t>(),
                                  w->d_summary.as<uint32_t>(), w->d_planes.as<uint32_t>(), w->d_total.as<uint64_t>(),
                                  w->d_a.as<int32_t>()));
        std::fill(w->stale.begin(), w->stale.end(), 0);
    }
    if ((rc = upload(s, w->d_b, list.data(), n * sizeof(int32_t))) != HB_OK) return rc;
    HB_CUDA(w->d_counts.reserve(n * sizeof(uint32_t)));
    HB_CUDA(w->d_offsets.reserve(n * sizeof(uint64_t)));
    HB_CUDA(w->d_scan.reserve(scan_workspace_bytes(n)));
    auto mesh = [&](bool emit, uint64_t cap) {
        return launch_mesh_ids(s, emit, c->tables, c->d_mats, w->d_pts.as<hb_chunk_pt>(), n, w->d_cubes.p, w->d_nbr.as<int32_t>(),
                               w->d_bitmaps.as<uint32_t>(), w->d_summary.as<uint32_t>(), w->d_planes.as<uint32_t>(),
                               w->d_counts.as<uint32_t>(), w->d_offsets.as<uint64_t>(), w->d_out.as<hb_instance3d>(), cap,
                               w->d_total.as<uint64_t>(), w->d_b.as<int32_t>());
    };
    HB_CUDA(mesh(false, 0));
    HB_CUDA(launch_scan(s, w->d_counts.as<uint32_t>(), n, w->d_offsets.as<uint64_t>(), w->d_scan.p, w->d_total.as<uint64_t>(), false));
    HB_CUDA(cudaMemcpyAsync(w->h_pin.p, w->d_total.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    uint64_t tot[4];
    memcpy(tot, w->h_pin.p, sizeof(tot));
    *out_total = tot[0];
    if (tot[2] & 2) return fail(HB_ERR_INVALID, "a loaded chunk holds a material id above the palette size %d", c->mats.n_materials);
    if (n > pts_capacity || (host_out && tot[0] > out_capacity))
        return fail(HB_ERR_CAPACITY, "remesh needs room for %zu chunks and %llu instances", n, (unsigned long long)tot[0]);
    HB_CUDA(cudaMemcpyAsync(out_offsets, w->d_offsets.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaMemcpyAsync(out_counts, w->d_counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    for (size_t i = 0; i < n; ++i) out_pts[i] = w->slot_pt[list[i]];
    if (tot[0]) {
        HB_CUDA(w->d_out.reserve(tot[0] * sizeof(hb_instance3d)));
        HB_CUDA(mesh(true, tot[0]));
    }
    HB_CUDA(cudaStreamSynchronize(s));
    std::fill(w->remesh.begin(), w->remesh.end(), 0);
    if (list_out) *list_out = list;
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_world_create(hb_ctx* c, size_t capacity, hb_world** out) {
    if (!c || !out) return fail(HB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (capacity == 0 || capacity > (size_t)INT32_MAX / 32) return fail(HB_ERR_INVALID, "bad world capacity %zu", capacity);
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    hb_world* w = new (std::nothrow) hb_world();
    if (!w) return fail(HB_ERR_NOMEM, "out of host memory");
    w->ctx = c;
    w->capacity = capacity;
    struct { Buf* b; size_t bytes; bool zero; } alloc[] = {
        {&w->d_cubes, capacity * HB_CHUNK_VOLUME * sizeof(hb_palette_cube), false},
        {&w->d_touched, capacity * HB_CHUNK_AREA * sizeof(uint32_t), true},
        {&w->d_pending, capacity * HB_CHUNK_AREA * sizeof(uint32_t), true},
        {&w->d_bitmaps, capacity * HB_CHUNK_AREA * sizeof(uint32_t), true},
        {&w->d_summary, capacity * sizeof(uint32_t), true},
        {&w->d_planes, capacity * 6 * HB_CHUNK_AREA * sizeof(uint32_t), true},
        {&w->d_pts, capacity * sizeof(hb_chunk_pt), true},
        {&w->d_nbr, capacity * 27 * sizeof(int32_t), true},
    };
    for (auto& a : alloc) {
        cudaError_t e = a.b->reserve(a.bytes);
        if (e == cudaSuccess && a.zero) e = cudaMemsetAsync(a.b->p, 0, a.bytes, c->stream);
        if (e != cudaSuccess) {
            const int rc = fail(HB_ERR_CUDA, "world allocation failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            for (auto& b : alloc) b.b->release();
            delete w;
            return rc;
        }
    }
    try {
        w->slot_pt.assign(capacity, hb_chunk_pt{0, 0, 0});
        w->used.assign(capacity, 0);
        w->epoch.assign(capacity, 0);
        w->h_nbr.assign(capacity * 27, -1);
        w->stale.assign(capacity, 0);
        w->touched.assign(capacity, 0);
        w->remesh.assign(capacity, 0);
        w->map.init(capacity);
        w->free_slots.resize(capacity);
        for (size_t i = 0; i < capacity; ++i) w->free_slots[i] = (int32_t)(capacity - 1 - i);  // pop_back hands out slot 0 first
    } catch (const std::bad_alloc&) {
        for (auto& b : alloc) b.b->release();
        delete w;
        return fail(HB_ERR_NOMEM, "out of host memory");
    }
    *out = w;
    return HB_OK;
}

void hb_world_destroy(hb_world* w) {
    if (!w) return;
    std::lock_guard<std::mutex> lk(w->ctx->mu);
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    Buf* bufs[] = {&w->d_cubes, &w->d_touched, &w->d_pending, &w->d_bitmaps, &w->d_summary, &w->d_planes, &w->d_pts, &w->d_nbr,
                   &w->d_a, &w->d_b, &w->d_c, &w->d_d, &w->d_heights, &w->d_counts, &w->d_offsets, &w->d_scan, &w->d_total,
                   &w->d_out, &w->h_pin};
    for (Buf* b : bufs) b->release();
    delete w;
}

size_t hb_world_len(const hb_world* w) { return w ? w->map.size() : 0; }

int hb_world_load(hb_world* w, const hb_chunk_pt* pts, size_t n) {
    if (!w || (n && !pts)) return fail(HB_ERR_INVALID, "null argument");
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // queue_load (map.rs:69-75) ignores loaded chunks; the provider generates each requested position once.
    // Positions are entered into the table as they are accepted (which also rejects repeats inside the batch) and
    // taken out again if the batch does not fit.
    for (size_t i = 0; i < n; ++i)
        if (!coord_ok(pts[i].x) || !coord_ok(pts[i].y) || !coord_ok(pts[i].z))
            return fail(HB_ERR_RANGE, "chunk %zu outside +-%d", i, HB_MAX_CHUNK_COORD);
    std::vector<hb_chunk_pt> fresh;
    std::vector<int32_t> slots;
    bool full = false;
    for (size_t i = 0; i < n && !full; ++i) {
        if (w->find(pts[i].x, pts[i].y, pts[i].z) >= 0) continue;
        if (w->free_slots.empty()) { full = true; break; }
        const int32_t slot = w->free_slots.back();
        w->free_slots.pop_back();
        w->map.insert(pts[i].x, pts[i].y, pts[i].z, slot);
        fresh.push_back(pts[i]);
        slots.push_back(slot);
    }
    if (full) {
        size_t wanted = fresh.size();
        for (size_t i = 0; i < n; ++i) wanted += w->find(pts[i].x, pts[i].y, pts[i].z) < 0;  // upper bound incl. repeats
        for (size_t i = fresh.size(); i-- > 0;) {
            w->map.erase(fresh[i].x, fresh[i].y, fresh[i].z);
            w->free_slots.push_back(slots[i]);
        }
        return fail(HB_ERR_CAPACITY, "world holds %zu of %zu chunks, about %zu more requested", w->map.size(), w->capacity, wanted);
    }
    const size_t m = fresh.size();
    if (m == 0) return HB_OK;
    // columns: all cy layers of one (cx, cz) share a heights tile
    std::vector<int32_t> cols, colidx(m);
    {
        SlotTable colmap;
        colmap.init(m);
        for (size_t i = 0; i < m; ++i) {
            int32_t ci = colmap.find(fresh[i].x, 0, fresh[i].z);
            if (ci < 0) {
                ci = (int32_t)(cols.size() / 2);
                colmap.insert(fresh[i].x, 0, fresh[i].z, ci);
                cols.push_back(fresh[i].x);
                cols.push_back(fresh[i].z);
            }
            colidx[i] = ci;
        }
    }
    const size_t ncol = cols.size() / 2;
    std::vector<uint8_t> is_new(w->capacity, 0);
    for (size_t i = 0; i < m; ++i) {
        is_new[slots[i]] = 1;
        w->slot_pt[slots[i]] = fresh[i];
        w->used[slots[i]] = 1;
    }
    // Device side in two halves with the host's neighbour bookkeeping in between, so that the generator kernels
    // (heights + fill: they need only the new chunks' positions and slots) run while the host wires the 27-shells and
    // collects the cull pairs.  If anything fails the host mirror is taken back to where it was, so no chunk is left
    // "loaded" with cubes that were never written (neighbours a failed cull may have touched keep valid, if stale,
    // flags; a CUDA error at this point is sticky for the context anyway).  Un-wiring a chunk that was never wired
    // only rewrites -1 entries, so the rollback below is the same wherever the failure happened.
    auto generate = [&]() -> int {
        int rc;
        if ((rc = upload(s, w->d_a, cols.data(), cols.size() * sizeof(int32_t))) != HB_OK) return rc;
        if ((rc = upload(s, w->d_b, colidx.data(), m * sizeof(int32_t))) != HB_OK) return rc;
        if ((rc = upload(s, w->d_c, fresh.data(), m * sizeof(hb_chunk_pt))) != HB_OK) return rc;
        if ((rc = upload(s, w->d_d, slots.data(), m * sizeof(int32_t))) != HB_OK) return rc;
        HB_CUDA(w->d_heights.reserve(ncol * HB_CHUNK_AREA * sizeof(int32_t)));
        HB_CUDA(launch_heights(s, c->world, w->d_a.as<int32_t>(), ncol, w->d_heights.as<int32_t>(), nullptr));
        HB_CUDA(launch_fill_world(s, w->d_c.as<hb_chunk_pt>(), w->d_b.as<int32_t>(), m, w->d_heights.as<int32_t>(),
                                  w->d_cubes.as<hb_palette_cube>(), w->d_d.as<int32_t>(), w->d_touched.as<uint32_t>()));
        return HB_OK;
    };
    int rc = generate();
    // Neighbour rows and cull pairs, one new chunk per work item on the context's host threads.  Every new chunk is
    // already in the table, so the entry two new neighbours have for each other comes from each one's own lookup;
    // back-pointers are written only into the rows of chunks that were loaded before, at the entry that faces this
    // chunk -- no two items write the same word.  load_provided (map.rs:203-228): one cull per (new chunk, loaded
    // face neighbour); a pair of two new chunks once.
    std::vector<CullPair> pair_buf(m * 6);
    std::vector<uint8_t> pair_cnt(m, 0);
    auto wire_one = [&](size_t i) {
        const hb_chunk_pt p = fresh[i];
        const int32_t slot = slots[i];
        int32_t* row = &w->h_nbr[(size_t)slot * 27];
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    const int o = (dx || dy || dz) ? w->find(p.x + dx, p.y + dy, p.z + dz) : slot;
                    row[shell_index(dx, dy, dz)] = o;
                    if (o >= 0 && o != slot && !is_new[o]) w->h_nbr[(size_t)o * 27 + shell_index(-dx, -dy, -dz)] = slot;
                }
        int k = 0;
        for (int f = 0; f < 6; ++f) {
            const int o = row[shell_index(kNormal[f][0], kNormal[f][1], kNormal[f][2])];
            if (o < 0 || (is_new[o] && o < slot)) continue;
            pair_buf[i * 6 + k++] = CullPair{slot, o, f, 0};
        }
        pair_cnt[i] = (uint8_t)k;
    };
    if (m >= 1024) {
        if (c->host_threads == 0) {
            c->host_threads = default_host_threads();
            c->pool.resize(c->host_threads);
        }
        c->pool.run(m, 64, wire_one);
    } else {
        for (size_t i = 0; i < m; ++i) wire_one(i);
    }
    w->tables_dirty = true;
    std::vector<CullPair> pairs;
    pairs.reserve(m * 3);
    for (size_t i = 0; i < m; ++i) {
        w->stale[slots[i]] = w->touched[slots[i]] = 1;
        for (int k = 0; k < pair_cnt[i]; ++k) {
            const CullPair& cp = pair_buf[i * 6 + k];
            pairs.push_back(cp);
            w->stale[cp.b] = w->touched[cp.b] = 1;
        }
    }
    auto cull = [&]() -> int {
        if (!pairs.empty()) {
            const int rc2 = upload(s, w->d_a, pairs.data(), pairs.size() * sizeof(CullPair));  // stream-ordered after the heights kernel's reads of d_a
            if (rc2 != HB_OK) return rc2;
            k_world_cull<<<(unsigned)pairs.size(), kWT, 0, s>>>(w->d_cubes.as<hb_palette_cube>(), w->d_touched.as<uint32_t>(),
                                                                w->d_a.as<CullPair>());
            HB_CUDA(cudaGetLastError());
        }
        HB_CUDA(cudaStreamSynchronize(s));
        return HB_OK;
    };
    if (rc == HB_OK) rc = cull();
    else cudaStreamSynchronize(s);
    if (rc != HB_OK) {
        for (size_t i = m; i-- > 0;) {
            w->map.erase(fresh[i].x, fresh[i].y, fresh[i].z);
            wire_neighbours(w, fresh[i], slots[i], false);
            for (int k = 0; k < 27; ++k) w->h_nbr[(size_t)slots[i] * 27 + k] = -1;
            w->used[slots[i]] = 0;
            w->epoch[slots[i]]++;
            w->stale[slots[i]] = w->touched[slots[i]] = w->remesh[slots[i]] = 0;
            w->free_slots.push_back(slots[i]);
        }
    }
    return rc;
}

int hb_world_unload(hb_world* w, const hb_chunk_pt* pts, size_t n) {
    if (!w || (n && !pts)) return fail(HB_ERR_INVALID, "null argument");
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    for (size_t i = 0; i < n; ++i) {
        const int32_t slot = w->find(pts[i].x, pts[i].y, pts[i].z);
        if (slot < 0) continue;
        w->map.erase(pts[i].x, pts[i].y, pts[i].z);
        wire_neighbours(w, pts[i], slot, false);
        for (int k = 0; k < 27; ++k) w->h_nbr[(size_t)slot * 27 + k] = -1;
        w->used[slot] = 0;
        w->epoch[slot]++;
        w->stale[slot] = w->touched[slot] = w->remesh[slot] = 0;
        // the slot's update sets need no clearing: a freed slot is never a sync candidate, the next load overwrites
        // `touched` completely (k_fill_cubes) and the sync after it overwrites `pending`
        w->free_slots.push_back(slot);
    }
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return HB_OK;
}

namespace {
// Conflict groups of an edit batch: the connected components of "the two edits' 7-voxel stencils (the voxel and its six
// neighbours) share a voxel".  group_of[i] = group of edit i; groups are numbered by their first edit; group_start is the
// exclusive prefix of the group sizes (n_groups + 1 entries).  Pure host code (hb_host_group_edits exposes it to tests).
void group_edits(const int32_t* xyz, size_t n, std::vector<uint32_t>& group_of, std::vector<uint32_t>& group_start) {
    // ---- conflict groups: union-find over the edits, joined wherever two 7-voxel stencils share a voxel.  The stencil
    // voxels go into an open-addressing table (key = world voxel, value = the first edit that touched it).
    std::vector<uint32_t> parent(n);
    for (size_t i = 0; i < n; ++i) parent[i] = (uint32_t)i;
    auto root = [&](uint32_t a) {
        while (parent[a] != a) a = parent[a] = parent[parent[a]];
        return a;
    };
    {
        struct Slot { int32_t x, y, z; uint32_t edit; };
        size_t cap = 16;
        while (cap < n * 14) cap <<= 1;  // 7 stencil voxels per edit, load factor <= 1/2
        std::vector<Slot> table(cap, Slot{0, 0, 0, UINT32_MAX});
        static const int kStencil[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        for (size_t i = 0; i < n; ++i) {
            for (int k = 0; k < 7; ++k) {
                const int32_t vx = xyz[3 * i] + kStencil[k][0], vy = xyz[3 * i + 1] + kStencil[k][1], vz = xyz[3 * i + 2] + kStencil[k][2];
                uint64_t h = (uint64_t)(uint32_t)vx * 0x9E3779B97F4A7C15ULL ^ (uint64_t)(uint32_t)vy * 0xC2B2AE3D27D4EB4FULL ^
                             (uint64_t)(uint32_t)vz * 0x165667B19E3779F9ULL;
                h ^= h >> 29;
                for (size_t q = (size_t)h & (cap - 1);; q = (q + 1) & (cap - 1)) {
                    Slot& t = table[q];
                    if (t.edit == UINT32_MAX) {
                        t = Slot{vx, vy, vz, (uint32_t)i};
                        break;
                    }
                    if (t.x == vx && t.y == vy && t.z == vz) {
                        const uint32_t a = root(t.edit), b = root((uint32_t)i);
                        if (a != b) parent[a > b ? a : b] = a > b ? b : a;  // the root is the component's first edit
                        break;
                    }
                }
            }
        }
    }
    // stable grouping: groups ordered by their first edit, edits inside a group in the caller's order
    group_of.assign(n, 0);
    group_start.clear();
    {
        std::vector<uint32_t> gid(n, UINT32_MAX), count;
        for (size_t i = 0; i < n; ++i) {
            const uint32_t r = root((uint32_t)i);
            if (gid[r] == UINT32_MAX) { gid[r] = (uint32_t)count.size(); count.push_back(0); }
            group_of[i] = gid[r];
            ++count[gid[r]];
        }
        group_start.resize(count.size() + 1);
        group_start[0] = 0;
        for (size_t g = 0; g < count.size(); ++g) group_start[g + 1] = group_start[g] + count[g];
    }
}

// one sub-batch of hb_world_set_cubes (c->mu held, device current): group, upload, launch, wait
int apply_edits(hb_world* w, const int32_t* xyz, const uint16_t* materials, size_t n, uint32_t* groups_out) {
    hb_ctx* c = w->ctx;
    std::vector<uint32_t> group_of, group_start;
    group_edits(xyz, n, group_of, group_start);
    const uint32_t n_groups = (uint32_t)group_start.size() - 1;
    std::vector<uint32_t> fill(group_start.begin(), group_start.end() - 1);
    std::vector<EditRec> recs(n);
    for (size_t i = 0; i < n; ++i) {
        const int32_t p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const int32_t ch[3] = {p[0] >> 5, p[1] >> 5, p[2] >> 5};  // lib/src/point.rs:25-32
        const int l[3] = {p[0] & 31, p[1] & 31, p[2] & 31};
        EditRec& r = recs[fill[group_of[i]]++];
        r.slot = w->find(ch[0], ch[1], ch[2]);
        r.x = (uint16_t)l[0]; r.y = (uint16_t)l[1]; r.z = (uint16_t)l[2];
        r.material = materials[i];
        if (r.slot >= 0) w->stale[r.slot] = w->touched[r.slot] = 1;
        for (int e = 0; e < 6; ++e) {
            const int axis = e >> 1, neg = !(e & 1);
            r.nbr[e] = -1;
            if (neg ? l[axis] != 0 : l[axis] != 31) continue;
            int32_t q[3] = {ch[0], ch[1], ch[2]};
            q[axis] += neg ? -1 : 1;
            r.nbr[e] = w->find(q[0], q[1], q[2]);
            if (r.nbr[e] >= 0) w->stale[r.nbr[e]] = w->touched[r.nbr[e]] = 1;
        }
    }
    int rc = upload(c->stream, w->d_a, recs.data(), n * sizeof(EditRec));
    if (rc != HB_OK) return rc;
    if ((rc = upload(c->stream, w->d_b, group_start.data(), group_start.size() * sizeof(uint32_t))) != HB_OK) return rc;
    k_world_edit<<<(n_groups + kEditWarps - 1) / kEditWarps, 32 * kEditWarps, 0, c->stream>>>(
        w->d_cubes.as<hb_palette_cube>(), w->d_touched.as<uint32_t>(), w->d_a.as<EditRec>(), w->d_b.as<uint32_t>(), n_groups);
    HB_CUDA(cudaGetLastError());
    *groups_out = n_groups;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return HB_OK;
}
}  // namespace

int hb_world_set_cubes(hb_world* w, const int32_t* xyz, const uint16_t* materials, size_t n) {
    if (!w || (n && (!xyz || !materials))) return fail(HB_ERR_INVALID, "null argument");
    if (n == 0) return HB_OK;
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    for (size_t i = 0; i < n; ++i)
        if (materials[i] > (uint32_t)c->mats.n_materials)
            return fail(HB_ERR_INVALID, "edit %zu: material id %u above the palette size %d", i, materials[i], c->mats.n_materials);
    // sub-batches bound the host-side grouping tables (224 bytes per edit); they run one after the other, so the order
    // between them is kept like the order inside a group
    const size_t kSub = 65536;
    w->last_edit_groups = 0;
    for (size_t first = 0; first < n; first += kSub) {
        uint32_t groups = 0;
        const int rc = apply_edits(w, xyz + 3 * first, materials + first, std::min(kSub, n - first), &groups);
        if (rc != HB_OK) return rc;
        w->last_edit_groups += groups;
    }
    return HB_OK;
}

uint32_t hb_world_last_edit_groups(const hb_world* w) { return w ? w->last_edit_groups : 0; }

int hb_host_group_edits(const int32_t* cubes_xyz, size_t n, uint32_t* group_of, uint32_t* n_groups) {
    if ((n && (!cubes_xyz || !group_of)) || !n_groups) return fail(HB_ERR_INVALID, "null argument");
    std::vector<uint32_t> g, start;
    if (n) group_edits(cubes_xyz, n, g, start);
    for (size_t i = 0; i < n; ++i) group_of[i] = g[i];
    *n_groups = n ? (uint32_t)start.size() - 1 : 0;
    return HB_OK;
}

int hb_world_sync(hb_world* w, size_t* n_dirty, uint64_t* n_entries) {
    if (!w) return fail(HB_ERR_INVALID, "null argument");
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    if (n_dirty) *n_dirty = 0;
    if (n_entries) *n_entries = 0;
    w->last_sync.clear();
    const std::vector<int32_t> cand = slots_where(w, w->touched);
    if (cand.empty()) return HB_OK;
    int rc = upload(s, w->d_a, cand.data(), cand.size() * sizeof(int32_t));
    if (rc != HB_OK) return rc;
    HB_CUDA(w->d_counts.reserve(cand.size() * sizeof(uint32_t)));
    k_world_drain<<<(unsigned)cand.size(), kWT, 0, s>>>(w->d_touched.as<uint32_t>(), w->d_pending.as<uint32_t>(), w->d_a.as<int32_t>(),
                                                        w->d_counts.as<uint32_t>());
    HB_CUDA(cudaGetLastError());
    std::vector<uint32_t> counts(cand.size());
    HB_CUDA(cudaMemcpyAsync(counts.data(), w->d_counts.p, cand.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    std::fill(w->touched.begin(), w->touched.end(), 0);
    uint64_t total = 0;
    for (size_t i = 0; i < cand.size(); ++i) {
        if (counts[i] == 0) continue;
        w->last_sync.push_back(hb_world::SyncRec{cand[i], w->epoch[cand[i]], counts[i]});
        w->remesh[cand[i]] = 1;
        total += counts[i];
    }
    if (n_dirty) *n_dirty = w->last_sync.size();
    if (n_entries) *n_entries = total;
    return HB_OK;
}

int hb_world_fetch_updates(hb_world* w, hb_chunk_pt* out_pts, size_t pts_capacity, uint64_t* out_offsets, uint32_t* out_counts,
                           hb_chunk_cube* out_entries, uint64_t entries_capacity, size_t* n_runs, uint64_t* n_entries) {
    if (!w || !n_runs || !n_entries) return fail(HB_ERR_INVALID, "null argument");
    *n_runs = 0;
    *n_entries = 0;
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    std::vector<int32_t> list;
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> counts;
    uint64_t total = 0;
    for (const auto& r : w->last_sync) {
        if (!w->used[r.slot] || w->epoch[r.slot] != r.epoch) continue;  // unloaded since the sync
        list.push_back(r.slot);
        offsets.push_back(total);
        counts.push_back(r.count);
        total += r.count;
    }
    const size_t n = list.size();
    *n_runs = n;
    *n_entries = total;
    if (n == 0) return HB_OK;
    if (n > pts_capacity || total > entries_capacity)
        return fail(HB_ERR_CAPACITY, "updates need room for %zu chunks and %llu entries", n, (unsigned long long)total);
    if (!out_pts || !out_offsets || !out_counts || !out_entries) return fail(HB_ERR_INVALID, "null argument");
    int rc;
    if ((rc = upload(s, w->d_a, list.data(), n * sizeof(int32_t))) != HB_OK) return rc;
    if ((rc = upload(s, w->d_b, offsets.data(), n * sizeof(uint64_t))) != HB_OK) return rc;
    HB_CUDA(w->d_out.reserve(total * sizeof(hb_chunk_cube)));
    k_world_updates<<<(unsigned)n, kWT, 0, s>>>(w->d_cubes.as<hb_palette_cube>(), w->d_pending.as<uint32_t>(), w->d_a.as<int32_t>(),
                                                w->d_b.as<uint64_t>(), w->d_out.as<uint32_t>());
    HB_CUDA(cudaGetLastError());
    if ((rc = copy_to_host(c, s, out_entries, w->d_out.p, total * sizeof(hb_chunk_cube))) != HB_OK) return rc;
    for (size_t i = 0; i < n; ++i) {
        out_pts[i] = w->slot_pt[list[i]];
        out_offsets[i] = offsets[i];
        out_counts[i] = counts[i];
    }
    return HB_OK;
}

int hb_world_remesh(hb_world* w, hb_chunk_pt* out_pts, size_t pts_capacity, hb_instance3d* out, uint64_t out_capacity,
                    uint64_t* out_offsets, uint32_t* out_counts, size_t* n_chunks, uint64_t* out_total) {
    if (!w || !n_chunks || !out_total) return fail(HB_ERR_INVALID, "null argument");
    if (pts_capacity && (!out_pts || !out_offsets || !out_counts)) return fail(HB_ERR_INVALID, "null argument");
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    int rc = remesh_impl(w, out_pts, pts_capacity, out ? out_capacity : 0, true, out_offsets, out_counts, n_chunks, out_total, nullptr);
    if (rc != HB_OK) return rc;
    // pinned destination: one DMA at link speed; pageable: streamed through the pinned bounce buffers
    if (*out_total) return copy_to_host(c, c->stream, out, w->d_out.p, *out_total * sizeof(hb_instance3d));
    return HB_OK;
}

int hb_world_remesh_dev(hb_world* w, hb_chunk_pt* out_pts, size_t pts_capacity, uint64_t* out_offsets, uint32_t* out_counts,
                        size_t* n_chunks, uint64_t* out_total, const hb_instance3d** d_instances) {
    if (!w || !n_chunks || !out_total) return fail(HB_ERR_INVALID, "null argument");
    if (pts_capacity && (!out_pts || !out_offsets || !out_counts)) return fail(HB_ERR_INVALID, "null argument");
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    int rc = remesh_impl(w, out_pts, pts_capacity, 0, false, out_offsets, out_counts, n_chunks, out_total, nullptr);
    if (rc != HB_OK) return rc;
    if (d_instances) *d_instances = *out_total ? w->d_out.as<hb_instance3d>() : nullptr;
    return HB_OK;
}

int hb_world_read_chunk(hb_world* w, hb_chunk_pt pt, hb_palette_cube* out_cubes) {
    if (!w || !out_cubes) return fail(HB_ERR_INVALID, "null argument");
    hb_ctx* c = w->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    const int slot = w->find(pt.x, pt.y, pt.z);
    if (slot < 0) return fail(HB_ERR_INVALID, "chunk (%d, %d, %d) is not loaded", pt.x, pt.y, pt.z);
    HB_CUDA(cudaStreamSynchronize(c->stream));
    HB_CUDA(cudaMemcpy(out_cubes, w->d_cubes.as<hb_palette_cube>() + (size_t)slot * HB_CHUNK_VOLUME,
                       HB_CHUNK_VOLUME * sizeof(hb_palette_cube), cudaMemcpyDeviceToHost));
    return HB_OK;
}

}  // extern "C"
