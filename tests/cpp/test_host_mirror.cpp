// C++ parity test of the host mirror (herbolution_b200/host/herbolution_host.hpp) against the CPU oracle.
// Reads like a test of the reference would: request chunks from a ChunkGenerator, dequeue them, put them
// into a client ChunkMap, remesh, and compare with the oracle's literal pipeline.
//   usage: test_host_mirror [--cpu-only]
#include <cstdio>
#include <cstring>
#include <thread>

#include "../../herbolution_b200/host/herbolution_host.hpp"
extern "C" {
#include "../../oracle/hb_oracle.h"
}

using namespace herbolution;

static int failures = 0;
#define EXPECT(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) {                                     \
            ++failures;                                    \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                      \
            std::printf("\n");                             \
        }                                                  \
    } while (0)

static void cpu_only_checks() {
    // create_chunk_shell order == client/src/world/chunk.rs:158-174
    ChunkTable map;
    for (int x = 0; x < 2; ++x)
        for (int z = 0; z < 2; ++z) {
            auto d = std::make_shared<ChunkData>();
            d->position = {x, 0, z};
            map[d->position] = d;
        }
    ChunkShell s = create_chunk_shell(map, {0, 0, 0});
    int present = 0;
    for (auto& e : s) present += e != nullptr;
    EXPECT(present == 4 && s[13] && s[13]->position == (ChunkPt{0, 0, 0}), "shell centre");
    EXPECT(s[14] && s[14]->position == (ChunkPt{0, 0, 1}), "shell +z = index 14");
    EXPECT(s[22] && s[22]->position == (ChunkPt{1, 0, 0}), "shell +x = index 22");
    EXPECT(linearize(1, 2, 3) == 1024 + 96 + 2, "linearize");
    static_assert(sizeof(Instance3d) == 100 && sizeof(PaletteCube) == 8, "layouts");
}

int main(int argc, char** argv) {
    cpu_only_checks();
    if (argc > 1 && !std::strcmp(argv[1], "--cpu-only")) {
        try {
            Context c(0);
            std::printf("note: a GPU is present\n");
        } catch (const Error& e) {
            EXPECT(e.code == HB_ERR_CUDA, "expected HB_ERR_CUDA without a GPU, got %d", e.code);
            std::printf("no GPU: %s\n", e.what());
        }
        std::printf(failures ? "FAILED\n" : "OK (cpu-only)\n");
        return failures ? 1 : 0;
    }

    const int64_t seed = 5;
    const int32_t reg[6] = {-2, 1, 3, 2, -2, 3};  // cx0, cz0, ncx, ncz, cy0, ncy
    const int n = reg[2] * reg[3] * reg[5];
    auto ctx = std::make_shared<Context>(0);
    auto params = std::make_shared<GenerationParams>(seed, ctx, /*want_cubes=*/true);
    ChunkGenerator gen(params);

    // requests arrive from several threads, like rayon workers calling request(&self)
    std::vector<ChunkPt> pts;
    for (int ix = 0; ix < reg[2]; ++ix)
        for (int iz = 0; iz < reg[3]; ++iz)
            for (int iy = 0; iy < reg[5]; ++iy) pts.push_back({reg[0] + ix, reg[4] + iy, reg[1] + iz});
    std::thread t1([&] { for (int i = 0; i < n / 2; ++i) gen.request(pts[(size_t)i]); });
    t1.join();
    std::thread t2([&] { for (int i = n / 2; i < n; ++i) gen.request(pts[(size_t)i]); });
    t2.join();
    std::vector<CubeMesh> meshes = gen.dequeue();
    EXPECT((int)meshes.size() == n, "dequeue returned %zu meshes", meshes.size());
    EXPECT(gen.dequeue().empty(), "second dequeue must be empty");

    // oracle: literal CubeMesh::set path per chunk
    hbo_world w{seed, 0.001f, 6, 2.0f, 0.6f, 1};
    hbo_palette pal;
    hbo_default_palette(&pal);
    std::vector<uint32_t> counts((size_t)n);
    std::vector<uint16_t> ref_ids((size_t)n * CHUNK_VOLUME);
    double secs[3];
    const uint64_t total = hbo_region_pipeline(&w, &pal, reg[0], reg[1], reg[2], reg[3], reg[4], reg[5], 1, 4, counts.data(),
                                               nullptr, 0, ref_ids.data(), secs);
    std::vector<hbo_instance> ref_inst(total);
    hbo_region_pipeline(&w, &pal, reg[0], reg[1], reg[2], reg[3], reg[4], reg[5], 1, 4, counts.data(), ref_inst.data(), total,
                        nullptr, secs);

    auto* one = new hbo_cube_mesh;
    for (int i = 0; i < n; ++i) {
        const CubeMesh& m = meshes[(size_t)i];
        EXPECT(m.position == pts[(size_t)i], "mesh %d position", i);
        EXPECT(!std::memcmp(m.ids.data(), &ref_ids[(size_t)i * CHUNK_VOLUME], CHUNK_VOLUME * 2), "chunk %d ids differ", i);
        hbo_mesh_init(one, m.position.x, m.position.y, m.position.z);
        hbo_generate_literal(&w, one);
        EXPECT(!std::memcmp(m.data.data(), one->data, sizeof(one->data)), "chunk %d PaletteCube image differs", i);
        EXPECT(m.get(3, 0, 7) == one->data[linearize(3, 0, 7)].material, "CubeMesh::get");
    }
    delete one;

    // client side: insert, update (one batched remesh), compare per chunk
    ChunkMap cm(ctx);
    for (const CubeMesh& m : meshes) cm.insert(m.position, m.ids);
    EXPECT((int)cm.update() == n, "ChunkMap::update remeshed count");
    EXPECT(cm.update() == 0, "nothing dirty after update");
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        const auto& got = cm.meshes[pts[(size_t)i]];
        EXPECT(got.size() == counts[(size_t)i], "chunk %d: %zu quads, oracle %u", i, got.size(), counts[(size_t)i]);
        if (got.size() == counts[(size_t)i])
            EXPECT(!std::memcmp(got.data(), &ref_inst[off], got.size() * sizeof(Instance3d)), "chunk %d: quads differ", i);
        off += counts[(size_t)i];
    }

    // single-chunk drop-in signature
    std::vector<Instance3d> inst;
    generate_mesh(*ctx, create_chunk_shell(cm.map, pts[4]), inst);
    EXPECT(inst.size() == cm.meshes[pts[4]].size() &&
               !std::memcmp(inst.data(), cm.meshes[pts[4]].data(), inst.size() * sizeof(Instance3d)),
           "generate_mesh(shell) == batched mesh");

    // flags-driven (literal) mesher: PaletteCubes as generate() leaves them, before any cull_shared_faces
    {
        ChunkTable table;
        for (const CubeMesh& m : meshes) {
            auto d = std::make_shared<ChunkData>();
            d->position = m.position;
            d->data = m.data;
            table[m.position] = d;
        }
        const ChunkPt centre = pts[7];
        ChunkShell shell = create_chunk_shell(table, centre);
        std::vector<Instance3d> got;
        generate_mesh(*ctx, shell, got);
        const hbo_cube* oshell[27];
        for (int i = 0; i < 27; ++i) oshell[i] = shell[(size_t)i] ? reinterpret_cast<const hbo_cube*>(shell[(size_t)i]->data.data()) : nullptr;
        const int32_t cpos[3] = {centre.x, centre.y, centre.z};
        const size_t nref = hbo_generate_mesh(oshell, cpos, &pal, nullptr, 0);
        std::vector<hbo_instance> ref(nref);
        hbo_generate_mesh(oshell, cpos, &pal, ref.data(), nref);
        EXPECT(got.size() == nref, "flags mode: %zu quads, oracle %zu", got.size(), nref);
        if (got.size() == nref) EXPECT(!std::memcmp(got.data(), ref.data(), nref * sizeof(Instance3d)), "flags mode: quads differ");
    }

    // server ChunkMap on the device-resident store: load, dig/place in order, CubeUpdate wire, client remesh
    {
        ctx->set_world(seed);
        ServerChunkMap server(ctx, 16);
        hbo_map* om = hbo_map_create(&w);
        const ChunkPt lp[4] = {{0, -1, 0}, {1, -1, 0}, {0, 0, 0}, {1, 0, 0}};
        for (const ChunkPt& p : lp) {
            server.queue_load(p);
            hbo_map_load(om, p.x, p.y, p.z);
        }
        server.update();
        EXPECT(server.len() == 4, "ServerChunkMap::len %zu", server.len());
        const CubePt ed[5] = {{31, -3, 4}, {32, -3, 4}, {31, -3, 4}, {5, 40, 5}, {0, -1, 31}};
        const uint16_t em[5] = {0, 0, 2, 3, 0};
        for (int i = 0; i < 5; ++i) {
            server.set_cube(ed[i], em[i]);
            hbo_map_set_cube(om, ed[i].x, ed[i].y, ed[i].z, em[i]);
        }
        std::unordered_map<ChunkPt, std::vector<hbo_cube>, ChunkPtHash> client;
        for (const CubeUpdate& u : server.sync_with_client()) {
            std::vector<hbo_cube>& cc = client[u.chunk];
            cc.assign((size_t)CHUNK_VOLUME, hbo_cube{0, 0, 0});
            hbo_apply_updates(cc.data(), reinterpret_cast<const hbo_chunk_cube*>(u.overwrites.data()), u.overwrites.size());
        }
        EXPECT(client.size() == 4, "CubeUpdates for %zu chunks", client.size());
        for (const ChunkPt& p : lp) {
            const hbo_cube* ref = hbo_map_cubes(om, hbo_map_find(om, p.x, p.y, p.z));
            const std::vector<PaletteCube> got = server.read(p);
            EXPECT(!std::memcmp(got.data(), ref, sizeof(hbo_cube) * CHUNK_VOLUME), "server cubes (%d,%d,%d)", p.x, p.y, p.z);
            EXPECT(client.count(p) && !std::memcmp(client[p].data(), ref, sizeof(hbo_cube) * CHUNK_VOLUME),
                   "client copy after CubeUpdate (%d,%d,%d)", p.x, p.y, p.z);
        }
        auto remeshed = server.remesh();
        EXPECT(remeshed.size() == 4, "remeshed %zu chunks", remeshed.size());
        for (const auto& pr : remeshed) {
            const ChunkPt c = pr.first;
            const hbo_cube* oshell[27];
            int k = 0;
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz, ++k) {
                        const int h = hbo_map_find(om, c.x + dx, c.y + dy, c.z + dz);
                        oshell[k] = h >= 0 ? hbo_map_cubes(om, h) : nullptr;
                    }
            const int32_t cpos[3] = {c.x, c.y, c.z};
            const size_t nref = hbo_generate_mesh(oshell, cpos, &pal, nullptr, 0);
            std::vector<hbo_instance> ref(nref);
            hbo_generate_mesh(oshell, cpos, &pal, ref.data(), nref);
            EXPECT(pr.second.size() == nref, "remesh (%d,%d,%d): %zu quads, oracle %zu", c.x, c.y, c.z, pr.second.size(), nref);
            if (pr.second.size() == nref)
                EXPECT(!std::memcmp(pr.second.data(), ref.data(), nref * sizeof(Instance3d)), "remesh (%d,%d,%d) differs", c.x, c.y, c.z);
        }
        EXPECT(server.remesh().empty(), "nothing left to remesh");
        hbo_map_destroy(om);
    }

    // get_noise and error behaviour
    auto tile = params->get_noise(3, -4);
    float ref_tile[CHUNK_AREA];
    hbo_noise_tile2(&w, 3, -4, ref_tile);
    EXPECT(!std::memcmp(tile.data(), ref_tile, sizeof(ref_tile)), "get_noise tile differs");
    try {
        CubeMesh far(ChunkPt{1 << 20, 0, 0});
        params->generate(far);
        EXPECT(false, "out-of-range chunk must throw");
    } catch (const Error& e) {
        EXPECT(e.code == HB_ERR_RANGE, "expected HB_ERR_RANGE, got %d", e.code);
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK: %d chunks, %llu quads byte-exact vs oracle\n", failures ? failures : n,
                (unsigned long long)total);
    return failures ? 1 : 0;
}
