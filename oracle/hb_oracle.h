/*
 * hb_oracle.h -- CPU restatement ("oracle") of herbolution's chunk pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (herbolution_b200/,
 * include/, libherbgpu.so) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the CPU arm.
 *
 * PARITY PINNING STATUS
 *   The reference (Rust nightly, un-vendored deps) cannot be compiled here and
 *   ships no tests, golden vectors or fixtures for this path.
 *   => "parity unpinned" by the reference itself.  What IS pinned:
 *     - SipHash-1-3 (std DefaultHasher) against CPython's siphash13 with a
 *       zero key (tests/test_oracle_hash.py);
 *     - the literal CubeMesh::set / cull_shared_faces transcription against
 *       the closed-form visibility rule (tests/test_oracle_gen.py);
 *     - noise against the survey's independent float32 numpy restatement
 *       values and structural invariants.
 *   wyrand (fastrand 2.3.0, un-vendored, Cargo.lock:712-715) is restated from
 *   its published algorithm and is unpinned: the colour channel is therefore
 *   compared/reported separately from geometry everywhere.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * it follows.
 */
#ifndef HB_ORACLE_H
#define HB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HBO_CHUNK_LENGTH 32      /* lib/src/world.rs:5 */
#define HBO_CHUNK_AREA 1024      /* lib/src/world.rs:6 */
#define HBO_CHUNK_VOLUME 32768   /* lib/src/world.rs:7 */

/* Face order == CubeFace discriminants, lib/src/spatial/face.rs:13-25 */
enum { HBO_EAST = 0, HBO_WEST, HBO_UP, HBO_DOWN, HBO_NORTH, HBO_SOUTH };

/* server/src/generator/mod.rs:98-111 (hard-coded there; parameters here) */
typedef struct hbo_world {
    int64_t seed;
    float freq;        /* 0.001 */
    int32_t octaves;   /* 6 */
    float lacunarity;  /* 2.0 */
    float gain;        /* 0.6 */
    int32_t fused;     /* 1: AVX2/NEON column (true FMA at neg_mul_add),
                          0: SSE/Scalar column (unfused)  simd/ops/f32.rs:141-162 */
    int32_t kind;      /* HBO_NOISE_*: which Sample32 impl the tile walk calls (noise/{fbm,ridge,turbulence,
                          gradient}.rs); the game uses FBM (generator/mod.rs:103) */
} hbo_world;
#define HBO_NOISE_FBM 0
#define HBO_NOISE_RIDGE 1
#define HBO_NOISE_TURBULENCE 2
#define HBO_NOISE_GRADIENT 3

/* == Cube<Option<PaletteMaterialId>>, server/src/chunk/cube.rs:5-10 (repr(C)) */
typedef struct hbo_cube {
    uint16_t material; /* 0 = None, else NonZeroU16(index+1) material.rs:215-226 */
    uint16_t pad;
    uint32_t flags;    /* CubeFlags, cube.rs:27-89 */
} hbo_cube;

/* == Instance3d, client/src/video/world/vertex.rs:41-51 (100 bytes) */
typedef struct hbo_instance {
    float model[16];
    float rgba[4];
    uint32_t light;
    float ao[4];
} hbo_instance;

/* server/src/chunk/material.rs:17-71: colours per material; index 0 unused (air) */
#define HBO_MAX_COLORS 8
typedef struct hbo_material {
    int32_t n_colors;
    float rgba[HBO_MAX_COLORS][4];
} hbo_material;

typedef struct hbo_palette {
    int32_t n_materials;           /* ids 1..n */
    hbo_material mat[16];          /* mat[id-1] */
} hbo_palette;

/* == CubeMesh, server/src/chunk/mesh.rs:12-19 (palette implicit) */
typedef struct hbo_cube_mesh {
    int32_t pos[3];
    hbo_cube data[HBO_CHUNK_VOLUME];
    uint64_t n_updated;            /* updated_positions.len() */
    uint16_t *upd;                 /* optional log of updated_positions (vec3u5 bits); NULL = count only */
    uint64_t upd_cap;              /* entries `upd` can hold; pushes beyond it are counted, not stored */
} hbo_cube_mesh;

/* ---- noise (crates/simd-noise) ---- */
float hbo_simplex2(float x, float y, int64_t seed, int fused);
float hbo_fbm2(float x, float y, const hbo_world *w);
float hbo_ridge2(float x, float y, const hbo_world *w);        /* functions/ridge_32.rs:18-31 */
float hbo_turbulence2(float x, float y, const hbo_world *w);   /* functions/turbulence_32.rs:18-32 */
float hbo_sample2(float x, float y, const hbo_world *w);       /* Sample32::sample_2d of w->kind */
float hbo_sample3(float x, float y, float z, const hbo_world *w);
void  hbo_noise_tile2(const hbo_world *w, int32_t cx, int32_t cz, float out[HBO_CHUNK_AREA]);
float hbo_simplex3(float x, float y, float z, int64_t seed);
float hbo_fbm3(float x, float y, float z, const hbo_world *w);
/* get_3d_noise_helper_f32: out index x + y*X + z*X*Y */
void  hbo_noise_tile3(const hbo_world *w, float sx, float sy, float sz,
                      int nx, int ny, int nz, float fx, float fy, float fz, float *out);

/* ---- generation ---- */
int32_t hbo_height(float noise);                                /* generator/mod.rs:79 */
void hbo_default_palette(hbo_palette *p);                       /* material.rs:27-61 */
void hbo_mesh_init(hbo_cube_mesh *m, int32_t cx, int32_t cy, int32_t cz); /* mesh.rs:22-30 */
void hbo_mesh_set(hbo_cube_mesh *m, int x, int y, int z, uint16_t material); /* mesh.rs:125-186 */
void hbo_generate_literal(const hbo_world *w, hbo_cube_mesh *m); /* generator/mod.rs:63-95 via set() */
/* closed form: ids only (index L = x*1024+z*32+y) */
void hbo_generate_ids(const hbo_world *w, int32_t cx, int32_t cy, int32_t cz, uint16_t *ids);
/* closed form of the PaletteCube image after generate() (in-chunk flags only) */
void hbo_cubes_from_ids(const uint16_t *ids, hbo_cube *out);
void hbo_cull_shared_faces(hbo_cube_mesh *a, hbo_cube_mesh *b); /* mesh.rs:76-123 */

/* ---- meshing ---- */
uint64_t hbo_siphash13_chunk(int32_t x, int32_t y, int32_t z);  /* client/src/world/chunk.rs:182-184 */
uint64_t hbo_wyrand_next(uint64_t *state);                      /* fastrand 2.3.0 Rng::gen_u64 */
float hbo_wyrand_f32(uint64_t *state);                          /* fastrand 2.3.0 Rng::f32 */
void hbo_face_matrix(int face, float out12[12]);                /* face.rs:60-69 + rotation.rs + vertex.rs:53-58 */
float hbo_ao_value(int k);                                      /* chunk.rs:271-275 */
/* literal generate_mesh: shell[27] of cube arrays (NULL = absent); flags are READ from cubes.
 * Returns number of instances; writes at most cap. client/src/world/chunk.rs:176-223 */
size_t hbo_generate_mesh(const hbo_cube *const shell[27], const int32_t center_pos[3],
                         const hbo_palette *pal, hbo_instance *out, size_t cap);
/* closed form: shell of u16 id arrays; visibility recomputed from ids */
size_t hbo_mesh_ids(const uint16_t *const shell[27], const int32_t center_pos[3],
                    const hbo_palette *pal, hbo_instance *out, size_t cap);

/* ---- region pipeline (CPU baseline driver; one chunk per task like the reference) ----
 * chunk index = ((ix*ncz)+iz)*ncy+iy.  literal!=0: set()-path + cull_shared_faces + flag mesher.
 * counts[n] receives per-chunk quad counts; if out!=NULL instances are concatenated in
 * chunk-index order (cap instances).  Returns total quads.  secs[3] = gen, cull, mesh wall s. */
uint64_t hbo_region_pipeline(const hbo_world *w, const hbo_palette *pal,
                             int32_t cx0, int32_t cz0, int32_t ncx, int32_t ncz,
                             int32_t cy0, int32_t ncy, int literal, int threads,
                             uint32_t *counts, hbo_instance *out, size_t cap,
                             uint16_t *ids_out, double secs[3]);

/* ---- server ChunkMap + wire format + client remesh ----
 * hbo_map is the server's ChunkMap (server/src/chunk/map.rs) holding CubeMeshes with their
 * updated_positions logs.  Chunk handles are dense indices in load order. */
typedef struct hbo_map hbo_map;
/* == ChunkCube (server/src/chunk/handle.rs:28-32): vec3u5 position + PaletteCube.  The Rust struct
 * is repr(Rust); this 12-byte C layout is ours. */
typedef struct hbo_chunk_cube {
    uint16_t pos;                  /* vec3u5 bits: 1 | x<<11 | y<<6 | z<<1 (lib/src/vector.rs:1181-1193) */
    uint16_t pad;
    hbo_cube cube;
} hbo_chunk_cube;

hbo_map *hbo_map_create(const hbo_world *w);
void hbo_map_destroy(hbo_map *m);
/* generator (literal set() path) + ChunkMap::load_provided (map.rs:203-228): cull_shared_faces
 * against every already-loaded face neighbour, then insert.  Returns the chunk's handle, or the
 * existing one if it is already loaded (queue_load ignores those, map.rs:69-75). */
int hbo_map_load(hbo_map *m, int32_t cx, int32_t cy, int32_t cz);
int hbo_map_find(const hbo_map *m, int32_t cx, int32_t cy, int32_t cz); /* -1 = not loaded */
void hbo_map_unload(hbo_map *m, int32_t cx, int32_t cy, int32_t cz);   /* map.rs:230-235 */
int hbo_map_len(const hbo_map *m);                                     /* handles issued so far */
/* ChunkMap::set_cube (map.rs:81-151), world cube coordinates, material 0 = None */
void hbo_map_set_cube(hbo_map *m, int32_t wx, int32_t wy, int32_t wz, uint16_t material);
const hbo_cube *hbo_map_cubes(const hbo_map *m, int handle);           /* NULL if unloaded */
void hbo_map_position(const hbo_map *m, int handle, int32_t pos[3]);
uint64_t hbo_map_pending(const hbo_map *m, int handle);                /* updated_positions.len() */
/* Chunk::sync_with_client (server/src/chunk/mod.rs:35-67): drains updated_positions (duplicates
 * and order kept) into ChunkCube entries carrying the cube as it is now.  Returns the number of
 * entries; writes at most cap. */
uint64_t hbo_map_drain(hbo_map *m, int handle, hbo_chunk_cube *out, uint64_t cap);
/* client apply_updates_from_server (client/src/world/chunk.rs:132-149) */
void hbo_apply_updates(hbo_cube *client_cubes, const hbo_chunk_cube *entries, uint64_t n);

/* ---- on-disk cube codec (server/src/chunk/codec.rs) ----
 * raw ids: 0 = None, k = palette index k-1 (the NonZeroU16 niche the reference stores). */
size_t hbo_rle_encode(const uint16_t *raw_ids, size_t n, uint8_t *out, size_t cap); /* codec.rs:73-101 */
int hbo_rle_decode(const uint8_t *bytes, size_t nbytes, uint16_t *raw_ids);        /* codec.rs:52-66 */
size_t hbo_encode_palette(const hbo_palette *pal, const char *const *groups, const char *const *keys, const uint8_t *cullable,
                          const uint8_t *collider, const float *toughness, uint8_t *out, size_t cap); /* codec.rs:70-75, material.rs:73-99 */

#ifdef __cplusplus
}
#endif
#endif
