// herbolution_host.hpp -- C++17 host-side mirror of the reference's generator / mesher interface over the
// C ABI of libherbgpu (include/herbgpu.h).  Header-only; link with -lherbgpu.
//
// The reference is Rust; its toolchain is not available here, so this header plays the role of the thin
// `herbgpu-sys` + glue a maintainer would write, with the reference's names, argument
// meaning and error behaviour:
//
//   server/src/generator/mod.rs      ChunkGenerator::{new, request, dequeue}, GenerationParams::{new, generate, get_noise}
//   server/src/chunk/mesh.rs         CubeMesh {position, data}, CubeMesh::get
//   client/src/world/chunk.rs        ChunkData, create_chunk_shell, generate_mesh, ChunkMap::update
//   server/src/chunk/map.rs          ChunkMap::{queue_load, queue_unload, set_cube, update} -> ServerChunkMap
//   server/src/chunk/mod.rs          Chunk::sync_with_client -> ServerChunkMap::sync_with_client (CubeUpdate lists)
//
// Errors: the reference panics (unwrap) on internal failures; here every failing hb_* call throws
// herbolution::Error carrying the hb_status and hb_last_error().  There is no CPU fallback.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/herbgpu.h"

namespace herbolution {

constexpr int CHUNK_LENGTH = HB_CHUNK_LENGTH;  // lib/src/world.rs:5
constexpr int CHUNK_AREA = HB_CHUNK_AREA;
constexpr int CHUNK_VOLUME = HB_CHUNK_VOLUME;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("herbgpu error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
    if (rc != HB_OK) throw Error(rc, hb_last_error());
}

// lib::point::ChunkPt (lib/src/point.rs:34-35)
struct ChunkPt {
    int32_t x = 0, y = 0, z = 0;
    bool operator==(const ChunkPt& o) const { return x == o.x && y == o.y && z == o.z; }
    ChunkPt operator+(const ChunkPt& o) const { return {x + o.x, y + o.y, z + o.z}; }
};
static_assert(sizeof(ChunkPt) == sizeof(hb_chunk_pt), "ChunkPt layout");
struct ChunkPtHash {
    size_t operator()(const ChunkPt& p) const {
        uint64_t h = (uint64_t)(uint32_t)p.x * 0x9E3779B97F4A7C15ULL;
        h ^= ((uint64_t)(uint32_t)p.y + 0x7F4A7C15ULL) * 0xC2B2AE3D27D4EB4FULL;
        h = (h << 31) | (h >> 33);
        h ^= (uint64_t)(uint32_t)p.z * 0x165667B19E3779F9ULL;
        return (size_t)(h ^ (h >> 29));
    }
};

using PaletteCube = hb_palette_cube;  // server/src/chunk/material.rs:144
using Instance3d = hb_instance3d;     // client/src/video/world/vertex.rs:41-51

// lib/src/vector.rs:1274-1276
inline int linearize(int x, int y, int z) { return x * CHUNK_LENGTH * CHUNK_LENGTH + z * CHUNK_LENGTH + y; }

// One GPU.  Shared by generator and mesher objects (the reference shares Arc<GenerationParams> etc.).
class Context {
   public:
    explicit Context(int device = 0) { check(hb_create(device, &h_)); }
    ~Context() { hb_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    hb_ctx* raw() const { return h_; }
    // server/src/generator/mod.rs:103-108 are the defaults
    void set_world(int64_t seed, float freq = 0.001f, int octaves = 6, float lacunarity = 2.0f, float gain = 0.6f) {
        check(hb_set_world(h_, seed, freq, octaves, lacunarity, gain));
    }

   private:
    hb_ctx* h_ = nullptr;
};

// server/src/chunk/mesh.rs:12-19.  `ids` are the bare materials (0 = None), `data` the PaletteCube image
// when requested.
struct CubeMesh {
    ChunkPt position;
    std::vector<uint16_t> ids;
    std::vector<PaletteCube> data;
    explicit CubeMesh(ChunkPt p = {}) : position(p) {}  // CubeMesh::new (mesh.rs:22-30)
    uint16_t get(int x, int y, int z) const { return ids.empty() ? 0 : ids[(size_t)linearize(x, y, z)]; }  // mesh.rs:32-34
};

// server/src/generator/mod.rs:51-111
class GenerationParams {
   public:
    GenerationParams(int64_t seed, std::shared_ptr<Context> ctx, bool want_cubes = false)
        : seed_(seed), ctx_(std::move(ctx)), want_cubes_(want_cubes) {
        ctx_->set_world(seed);
    }
    int64_t seed() const { return seed_; }
    Context& context() const { return *ctx_; }

    // get_noise (:98-111): the 32x32 tile of chunk column (cx, cz), index x + z*32
    std::array<float, CHUNK_AREA> get_noise(int32_t cx, int32_t cz) const {
        std::array<float, CHUNK_AREA> out{};
        const int32_t col[2] = {cx, cz};
        check(hb_noise_tiles(ctx_->raw(), col, 1, out.data(), nullptr));
        return out;
    }

    // generate(&self, chunk: &mut CubeMesh) (:63-95)
    void generate(CubeMesh& chunk) const {
        std::vector<CubeMesh*> one{&chunk};
        generate_into(one);
    }

    // the batched form ChunkGenerator::dequeue uses
    std::vector<CubeMesh> generate_batch(const std::vector<ChunkPt>& positions) const {
        std::vector<CubeMesh> out;
        out.reserve(positions.size());
        for (const ChunkPt& p : positions) out.emplace_back(p);
        std::vector<CubeMesh*> ptrs;
        for (CubeMesh& m : out) ptrs.push_back(&m);
        generate_into(ptrs);
        return out;
    }

   private:
    void generate_into(std::vector<CubeMesh*>& chunks) const {
        const size_t n = chunks.size();
        if (!n) return;
        std::vector<hb_chunk_pt> pts(n);
        for (size_t i = 0; i < n; ++i) pts[i] = {chunks[i]->position.x, chunks[i]->position.y, chunks[i]->position.z};
        std::vector<uint16_t> ids(n * CHUNK_VOLUME);
        std::vector<PaletteCube> cubes(want_cubes_ ? n * CHUNK_VOLUME : 0);
        check(hb_generate(ctx_->raw(), pts.data(), n, ids.data(), want_cubes_ ? cubes.data() : nullptr, nullptr));
        for (size_t i = 0; i < n; ++i) {
            chunks[i]->ids.assign(ids.begin() + (ptrdiff_t)(i * CHUNK_VOLUME), ids.begin() + (ptrdiff_t)((i + 1) * CHUNK_VOLUME));
            if (want_cubes_)
                chunks[i]->data.assign(cubes.begin() + (ptrdiff_t)(i * CHUNK_VOLUME), cubes.begin() + (ptrdiff_t)((i + 1) * CHUNK_VOLUME));
        }
    }
    int64_t seed_;
    std::shared_ptr<Context> ctx_;
    bool want_cubes_;
};

// server/src/generator/mod.rs:14-49.  request() may be called from any thread (the reference takes &self);
// the batch accumulated since the last dequeue() is generated by ONE hb_generate call.
class ChunkGenerator {
   public:
    explicit ChunkGenerator(std::shared_ptr<GenerationParams> generator) : params_(std::move(generator)) {}
    void request(ChunkPt position) {
        std::lock_guard<std::mutex> lk(mu_);
        pending_.push_back(position);
    }
    std::vector<CubeMesh> dequeue() {
        std::vector<ChunkPt> batch;
        {
            std::lock_guard<std::mutex> lk(mu_);
            batch.swap(pending_);
        }
        return params_->generate_batch(batch);
    }

   private:
    std::shared_ptr<GenerationParams> params_;
    std::mutex mu_;
    std::vector<ChunkPt> pending_;
};

// client/src/world/chunk.rs:97-102.  `cubes` = material ids (faces recomputed on the GPU); if `data` is non-empty
// it holds the PaletteCubes as received from the server and the faces are read from cube.flags (literal
// generate_mesh: hb_mesh_cubes), which is what the incremental dig/place path needs.
struct ChunkData {
    ChunkPt position;
    std::vector<uint16_t> cubes;    // CHUNK_VOLUME entries, index linearize(x, y, z)
    std::vector<PaletteCube> data;  // optional, CHUNK_VOLUME entries
};
using ChunkShell = std::array<std::shared_ptr<const ChunkData>, 27>;
using ChunkTable = std::unordered_map<ChunkPt, std::shared_ptr<ChunkData>, ChunkPtHash>;

// create_chunk_shell (chunk.rs:158-174): x outer, y, z inner; centre = index 13
inline ChunkShell create_chunk_shell(const ChunkTable& map, ChunkPt position) {
    ChunkShell shell{};
    int i = 0;
    for (int x = -1; x <= 1; ++x)
        for (int y = -1; y <= 1; ++y)
            for (int z = -1; z <= 1; ++z, ++i) {
                auto it = map.find(position + ChunkPt{x, y, z});
                if (it != map.end()) shell[(size_t)i] = it->second;
            }
    return shell;
}

// One batched hb_mesh call: instances of chunk i are result.instances[offsets[i] .. offsets[i] + counts[i]).
struct MeshBatch {
    std::vector<Instance3d> instances;
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> counts;
};
inline MeshBatch mesh_chunks(Context& ctx, const std::vector<const ChunkData*>& chunks) {
    const size_t n = chunks.size();
    MeshBatch out;
    out.offsets.assign(n, 0);
    out.counts.assign(n, 0);
    if (!n) return out;
    std::vector<hb_chunk_pt> pts(n);
    const bool flags = !chunks[0]->data.empty();
    std::vector<uint16_t> ids(flags ? 0 : n * CHUNK_VOLUME);
    std::vector<PaletteCube> cubes(flags ? n * CHUNK_VOLUME : 0);
    for (size_t i = 0; i < n; ++i) {
        pts[i] = {chunks[i]->position.x, chunks[i]->position.y, chunks[i]->position.z};
        if (flags) {
            if (chunks[i]->data.size() != (size_t)CHUNK_VOLUME) throw Error(HB_ERR_INVALID, "ChunkData::data must hold 32768 cubes");
            std::copy(chunks[i]->data.begin(), chunks[i]->data.end(), cubes.begin() + (ptrdiff_t)(i * CHUNK_VOLUME));
        } else {
            if (chunks[i]->cubes.size() != (size_t)CHUNK_VOLUME) throw Error(HB_ERR_INVALID, "ChunkData::cubes must hold 32768 ids");
            std::copy(chunks[i]->cubes.begin(), chunks[i]->cubes.end(), ids.begin() + (ptrdiff_t)(i * CHUNK_VOLUME));
        }
    }
    std::vector<int32_t> nbr(n * 27);
    check(hb_build_neighbor_index(pts.data(), n, nbr.data()));
    uint64_t total = 0;
    auto run = [&](Instance3d* dst, uint64_t cap) {
        return flags ? hb_mesh_cubes(ctx.raw(), pts.data(), n, cubes.data(), nbr.data(), dst, cap, out.offsets.data(),
                                     out.counts.data(), &total)
                     : hb_mesh(ctx.raw(), pts.data(), n, ids.data(), nbr.data(), dst, cap, out.offsets.data(),
                               out.counts.data(), &total);
    };
    int rc = run(nullptr, 0);
    if (rc != HB_OK && rc != HB_ERR_CAPACITY) check(rc);
    out.instances.resize(total);
    if (total) check(run(out.instances.data(), total));
    return out;
}

// generate_mesh(shell, instances) (chunk.rs:176-223): appends the centre chunk's instances.
inline void generate_mesh(Context& ctx, const ChunkShell& shell, std::vector<Instance3d>& instances) {
    if (!shell[13]) throw Error(HB_ERR_INVALID, "shell[13] (the centre chunk) must be present");  // reference: unwrap()
    std::vector<const ChunkData*> present;
    size_t centre = 0;
    for (size_t i = 0; i < 27; ++i)
        if (shell[i]) {
            if (i == 13) centre = present.size();
            present.push_back(shell[i].get());
        }
    MeshBatch b = mesh_chunks(ctx, present);
    instances.insert(instances.end(), b.instances.begin() + (ptrdiff_t)b.offsets[centre],
                     b.instances.begin() + (ptrdiff_t)(b.offsets[centre] + b.counts[centre]));
}

// client ChunkMap (chunk.rs:32-83): holds ChunkData; update() remeshes every dirty chunk in one batch.
class ChunkMap {
   public:
    explicit ChunkMap(std::shared_ptr<Context> ctx) : ctx_(std::move(ctx)) {}
    void insert(ChunkPt position, std::vector<uint16_t> cubes) {
        auto d = std::make_shared<ChunkData>();
        d->position = position;
        d->cubes = std::move(cubes);
        map[position] = std::move(d);
        dirty_.push_back(position);
    }
    size_t update() {
        if (dirty_.empty()) return 0;
        std::vector<const ChunkData*> all;
        std::vector<ChunkPt> order;
        for (auto& kv : map) {
            all.push_back(kv.second.get());
            order.push_back(kv.first);
        }
        MeshBatch b = mesh_chunks(*ctx_, all);
        std::unordered_map<ChunkPt, size_t, ChunkPtHash> index;
        for (size_t i = 0; i < order.size(); ++i) index[order[i]] = i;
        size_t done = 0;
        for (const ChunkPt& p : dirty_) {
            auto it = index.find(p);
            if (it == index.end()) continue;
            const size_t i = it->second;
            meshes[p].assign(b.instances.begin() + (ptrdiff_t)b.offsets[i],
                             b.instances.begin() + (ptrdiff_t)(b.offsets[i] + b.counts[i]));
            ++done;
        }
        dirty_.clear();
        return done;
    }
    ChunkTable map;
    std::unordered_map<ChunkPt, std::vector<Instance3d>, ChunkPtHash> meshes;

   private:
    std::shared_ptr<Context> ctx_;
    std::vector<ChunkPt> dirty_;
};

// lib::point::CubePt (lib/src/point.rs:7-12): world cube coordinates
struct CubePt {
    int32_t x = 0, y = 0, z = 0;
};
using ChunkCube = hb_chunk_cube;  // server/src/chunk/handle.rs:28-32
// server/src/chunk/handle.rs:19-22, with the chunk it is addressed to (the reference routes it through that
// chunk's own channel)
struct CubeUpdate {
    ChunkPt chunk;
    std::vector<ChunkCube> overwrites;
};

// Server ChunkMap (server/src/chunk/map.rs) over a device-resident store, plus the two steps that follow it each
// tick: Chunk::sync_with_client (server/src/chunk/mod.rs:35-67) and the client's remesh of every chunk that got a
// CubeUpdate (client/src/world/chunk.rs:47-76).  queue_load / queue_unload / set_cube / update keep the
// reference's names and meaning; update() flushes the queues like load_provided + unload_requested (map.rs:236-245).
class ServerChunkMap {
   public:
    ServerChunkMap(std::shared_ptr<Context> ctx, size_t capacity) : ctx_(std::move(ctx)) {
        check(hb_world_create(ctx_->raw(), capacity, &w_));
    }
    ~ServerChunkMap() { hb_world_destroy(w_); }
    ServerChunkMap(const ServerChunkMap&) = delete;
    ServerChunkMap& operator=(const ServerChunkMap&) = delete;

    void queue_load(ChunkPt p) { load_.push_back(p); }      // map.rs:69-75
    void queue_unload(ChunkPt p) { unload_.push_back(p); }  // map.rs:65-67
    // map.rs:81-151; material 0 = None
    void set_cube(CubePt p, uint16_t material) {
        const int32_t xyz[3] = {p.x, p.y, p.z};
        check(hb_world_set_cubes(w_, xyz, &material, 1));
    }
    void set_cubes(const std::vector<CubePt>& ps, const std::vector<uint16_t>& materials) {
        if (ps.size() != materials.size()) throw Error(HB_ERR_INVALID, "one material per edit");
        check(hb_world_set_cubes(w_, reinterpret_cast<const int32_t*>(ps.data()), materials.data(), ps.size()));
    }
    // map.rs:236-245: load_provided, then unload_requested
    void update() {
        check(hb_world_load(w_, reinterpret_cast<const hb_chunk_pt*>(load_.data()), load_.size()));
        load_.clear();
        check(hb_world_unload(w_, reinterpret_cast<const hb_chunk_pt*>(unload_.data()), unload_.size()));
        unload_.clear();
    }
    size_t len() const { return hb_world_len(w_); }
    // Chunk::sync_with_client for every chunk; want_wire = false skips building the CubeUpdate lists
    std::vector<CubeUpdate> sync_with_client(bool want_wire = true) {
        size_t n = 0;
        uint64_t e = 0;
        check(hb_world_sync(w_, &n, &e));
        std::vector<CubeUpdate> out;
        if (!want_wire || n == 0) return out;
        std::vector<hb_chunk_pt> pts(n);
        std::vector<uint64_t> off(n);
        std::vector<uint32_t> cnt(n);
        std::vector<ChunkCube> entries(e);
        check(hb_world_fetch_updates(w_, pts.data(), n, off.data(), cnt.data(), entries.data(), e, &n, &e));
        out.resize(n);  // chunks unloaded since the sync are skipped: n may have shrunk
        for (size_t i = 0; i < n; ++i) {
            out[i].chunk = ChunkPt{pts[i].x, pts[i].y, pts[i].z};
            out[i].overwrites.assign(entries.begin() + (ptrdiff_t)off[i], entries.begin() + (ptrdiff_t)(off[i] + cnt[i]));
        }
        return out;
    }
    // client ChunkMap::update: (position, instances) of every chunk that received a CubeUpdate since the last call
    std::vector<std::pair<ChunkPt, std::vector<Instance3d>>> remesh() {
        size_t n = 0;
        uint64_t total = 0;
        int rc = hb_world_remesh(w_, nullptr, 0, nullptr, 0, nullptr, nullptr, &n, &total);
        if (rc != HB_OK && rc != HB_ERR_CAPACITY) check(rc);
        std::vector<std::pair<ChunkPt, std::vector<Instance3d>>> out;
        if (rc == HB_OK) return out;  // nothing to do
        std::vector<hb_chunk_pt> pts(n);
        std::vector<uint64_t> off(n);
        std::vector<uint32_t> cnt(n);
        std::vector<Instance3d> inst(total);
        check(hb_world_remesh(w_, pts.data(), n, inst.data(), total, off.data(), cnt.data(), &n, &total));
        out.resize(n);
        for (size_t i = 0; i < n; ++i) {
            out[i].first = ChunkPt{pts[i].x, pts[i].y, pts[i].z};
            out[i].second.assign(inst.begin() + (ptrdiff_t)off[i], inst.begin() + (ptrdiff_t)(off[i] + cnt[i]));
        }
        return out;
    }
    std::vector<PaletteCube> read(ChunkPt p) const {
        std::vector<PaletteCube> out((size_t)CHUNK_VOLUME);
        check(hb_world_read_chunk(w_, hb_chunk_pt{p.x, p.y, p.z}, out.data()));
        return out;
    }

   private:
    std::shared_ptr<Context> ctx_;
    hb_world* w_ = nullptr;
    std::vector<ChunkPt> load_, unload_;
};

}  // namespace herbolution
