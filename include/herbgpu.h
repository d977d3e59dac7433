/*
 * herbgpu.h -- C ABI of libherbgpu.so: herbolution's chunk pipeline (terrain generation +
 * chunk meshing) as hand-written sm_100a CUDA.  This is the drop-in boundary: plain pointers
 * and sizes, no C++/torch types.  A Rust host binds it with a ~100-line `extern "C"` block
 * (INTEGRATION.md); the Python mirror in herbolution_b200/ binds it with ctypes.
 *
 * Reference citations are relative to the herbolution source tree.
 *
 * Conventions
 *   - chunk = 32^3 voxels (lib/src/world.rs:4-7); voxel index L = x*1024 + z*32 + y
 *     (lib/src/vector.rs:1274-1276).
 *   - block id: u16, 0 = air (Option<PaletteMaterialId>::None), k = palette index k-1
 *     (server/src/chunk/material.rs:215-226).  Generated terrain uses 1 = stone, 2 = dirt,
 *     3 = grass (server/src/generator/mod.rs:64-72).
 *   - face order / bit: East(+x)=0, West(-x)=1, Up(+y)=2, Down(-y)=3, North(+z)=4, South(-z)=5
 *     (lib/src/spatial/face.rs:13-25,118-123).
 *   - every call returns HB_OK (0) or a negative hb_status; nothing unwinds across the ABI;
 *     hb_last_error() gives a thread-local message.
 *   - one hb_ctx per GPU (one process or one host thread per GPU); calls on one ctx are
 *     serialised by an internal mutex, so rayon workers may share it.
 *   - there is NO CPU fallback: without a CUDA device hb_create fails with HB_ERR_CUDA.
 */
#ifndef HERBGPU_H
#define HERBGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Only the entry points below are exported from libherbgpu.so (it is built with -fvisibility=hidden). */
#if defined(__GNUC__)
#define HB_API __attribute__((visibility("default")))
#else
#define HB_API
#endif

#define HB_CHUNK_LENGTH 32
#define HB_CHUNK_AREA 1024
#define HB_CHUNK_VOLUME 32768
#define HB_MAX_MATERIALS 16
#define HB_MAX_COLORS 8
/* Chunk coordinates must satisfy |c| <= HB_MAX_CHUNK_COORD so that world block coordinates are
 * exact in f32, as the reference's tile walk assumes (crates/simd-noise/src/noise/f32.rs:87-112). */
#define HB_MAX_CHUNK_COORD 262143

typedef enum hb_status {
    HB_OK = 0,
    HB_ERR_INVALID = -1,   /* bad argument (null pointer, unknown material id, ...) */
    HB_ERR_CUDA = -2,      /* CUDA runtime error, or no device */
    HB_ERR_CAPACITY = -3,  /* output buffer too small; *out_total holds the required size */
    HB_ERR_RANGE = -4,     /* chunk coordinate outside +-HB_MAX_CHUNK_COORD */
    HB_ERR_NOMEM = -5
} hb_status;

typedef struct hb_ctx hb_ctx;

/* == lib::point::ChunkPt (repr of Vec3<i32>), lib/src/point.rs:34-35 */
typedef struct hb_chunk_pt { int32_t x, y, z; } hb_chunk_pt;

/* == client Instance3d, client/src/video/world/vertex.rs:41-51: 100 bytes, 4-byte aligned */
typedef struct hb_instance3d {
    float model[16];  /* model_0..model_3 (columns) */
    float rgba[4];
    uint32_t light;
    float ao[4];
} hb_instance3d;

/* == PaletteCube = Cube<Option<PaletteMaterialId>>, server/src/chunk/cube.rs:5-10 (repr(C)), 8 bytes */
typedef struct hb_palette_cube {
    uint16_t material;
    uint16_t _pad;
    uint32_t flags;   /* CubeFlags: bit0 opaque tag, bits1..6 faces (cube.rs:27-72) */
} hb_palette_cube;

/* The information content of one Instance3d, as the device sends it to the host (4 bytes instead of 100):
 *   bits  0-14  voxel index L = x*1024 + z*32 + y            (lib/src/vector.rs:1274-1276)
 *   bits 15-17  face (East, West, Up, Down, North, South)    (lib/src/spatial/face.rs:13-25)
 *   bits 18-25  the four vertex_ao values, 2 bits each, in Instance3d::ao order (client/src/world/chunk.rs:255-276)
 *   bits 26-29  material id - 1
 * Everything else in the 100-byte record (face matrix, world position, colour draw, light, ao floats) is a pure
 * function of these bits and the chunk position: hb_expand_quads rebuilds it bit for bit. */
typedef uint32_t hb_packed_quad;

/* A box of chunks: cx in [cx0,cx0+ncx), cz in [cz0,cz0+ncz), cy in [cy0,cy0+ncy).
 * Chunk index inside a region = ((ix*ncz)+iz)*ncy+iy.  Chunks outside the region are ABSENT:
 * border faces toward them are emitted and AO probes into them read "empty", exactly like
 * unloaded chunks in the reference (client/src/world/chunk.rs:158-174,225-236). */
typedef struct hb_region { int32_t cx0, cz0, ncx, ncz, cy0, ncy; } hb_region;

/* ---- lifetime ---- */
HB_API int hb_create(int device, hb_ctx **out);
HB_API void hb_destroy(hb_ctx *ctx);
HB_API const char *hb_last_error(void);
HB_API const char *hb_version(void);

/* World parameters; defaults are the reference's hard-coded values
 * (server/src/generator/mod.rs:103-108): freq 0.001, 6 octaves, lacunarity 2.0, gain 0.6.
 * seed is GenerationParams::seed (i64). */
HB_API int hb_set_world(hb_ctx *ctx, int64_t seed, float freq, int32_t octaves, float lacunarity, float gain);
/* Which of the crate's noise types the tile walks sample (crates/simd-noise/src/noise/{fbm,ridge,turbulence,
 * gradient}.rs over functions/{fbm,ridge,turbulence}_32.rs); applies to hb_noise_tiles, hb_noise_fbm3d and to the
 * heights every generation path derives.  The game uses HB_NOISE_FBM (server/src/generator/mod.rs:103), the default.
 * Gradient = one simplex evaluation (octaves, lacunarity and gain unused). */
#define HB_NOISE_FBM 0
#define HB_NOISE_RIDGE 1
#define HB_NOISE_TURBULENCE 2
#define HB_NOISE_GRADIENT 3
HB_API int hb_set_noise_kind(hb_ctx *ctx, int32_t kind);
/* Which arithmetic column of the reference's runtime SIMD dispatch (crates/simd-noise/src/simd/invoking.rs:169-199)
 * the 2-D/3-D noise reproduces bit for bit.  neg_mul_add (simd/ops/f32.rs:141-162; six sites in simplex_2d, one in
 * ridge) is ONE fused operation on AVX2+FMA and NEON hosts -- fused = 1, the default, what any x86-64 machine since 2013
 * selects -- and a rounded product followed by a subtraction on SSE2 / SSE4.1 / scalar / wasm hosts -- fused = 0.  The
 * two columns differ by a few ulp (inside north_star's 1e-5 relative tolerance). */
HB_API int hb_set_fma_mode(hb_ctx *ctx, int32_t fused);

/* Palette colours, Material::texture (server/src/chunk/material.rs:27-71): material id m (1-based)
 * has n_colors[m-1] colours, rgba[(m-1)*HB_MAX_COLORS + c][4].  Default = stone, dirt, grass. */
HB_API int hb_set_materials(hb_ctx *ctx, int32_t n_materials, const int32_t *n_colors, const float *rgba);

/* Pinned host memory for full-speed async copies (optional; any host pointer is accepted). */
HB_API int hb_host_alloc(hb_ctx *ctx, size_t bytes, void **out);
HB_API int hb_host_free(hb_ctx *ctx, void *p);   /* ctx may be NULL (the memory outlives the ctx) */

/* ---- generation ---- */

/* Replaces GenerationParams::get_noise (server/src/generator/mod.rs:98-111) for n chunk columns.
 * cols_xz: n x {cx, cz}.  out_noise: n x 1024 f32, index x + z*32 (the crate's tile order,
 * crates/simd-noise/src/noise/f32.rs:69-129).  out_heights (nullable): n x 1024 i32,
 * h = trunc(noise*32) (generator/mod.rs:79), index x*32 + z. */
HB_API int hb_noise_tiles(hb_ctx *ctx, const int32_t *cols_xz, size_t n, float *out_noise, int32_t *out_heights);

/* get_3d_noise (crates/simd-noise/src/noise/f32.rs:131-198) with FbmNoise: nx*ny*nz samples starting
 * at (sx,sy,sz), step 1, multiplied by freq[3]; out index x + y*nx + z*nx*ny.  Uses the ctx's seed,
 * octaves, lacunarity, gain. */
HB_API int hb_noise_fbm3d(hb_ctx *ctx, float sx, float sy, float sz, int32_t nx, int32_t ny, int32_t nz,
                   const float *freq3, float *out);

/* Replaces ChunkGenerator::request/dequeue + GenerationParams::generate
 * (server/src/generator/mod.rs:32-48,63-95) for a batch of n chunks.
 * out_ids: n x 32768 u16.  out_cubes (nullable): n x 32768 hb_palette_cube = CubeMesh::data right
 * after generate() (in-chunk face flags as CubeMesh::set leaves them, server/src/chunk/mesh.rs:125-232;
 * cross-chunk culling not yet applied).  out_noise (nullable): n x 1024 f32 as hb_noise_tiles.
 * out_ids follow the region transport (hb_set_region_transport): with the packed one (default) the heights tiles come
 * back and the host threads write the ids into out_ids (which should then be 16-byte aligned; otherwise, and with
 * HB_TRANSPORT_INSTANCES, the device writes them and 64 KiB per chunk are copied).  Same ids either way. */
HB_API int hb_generate(hb_ctx *ctx, const hb_chunk_pt *pts, size_t n, uint16_t *out_ids,
                hb_palette_cube *out_cubes, float *out_noise);

/* ---- meshing ---- */

/* Host helper (no GPU work): out[i*27 + (dx+1)*9 + (dy+1)*3 + (dz+1)] = index of chunk pts[i]+(dx,dy,dz)
 * in pts, or -1.  Same shell order as create_chunk_shell (client/src/world/chunk.rs:158-174). */
HB_API int hb_build_neighbor_index(const hb_chunk_pt *pts, size_t n, int32_t *out);

/* Replaces the remesh loop of client ChunkMap::update + generate_mesh
 * (client/src/world/chunk.rs:66-75,176-276) together with the server-side face culling it relies on
 * (CubeMesh::set in-chunk flags, cull_shared_faces across borders: server/src/chunk/mesh.rs:76-232).
 * ids: n x 32768 u16.  neighbor_index: n x 27 as above (-1 = absent chunk).
 * Output: chunk i's instances are out[out_offsets[i] .. out_offsets[i]+out_counts[i]) in the
 * reference's emission order (ascending L, then face order).  out_offsets[i] is a multiple of 4;
 * the gap after each chunk is zero-filled.  *out_total = instances spanned (incl. gaps).
 * If out_capacity (in instances) is too small nothing is written to `out`, *out_total holds the
 * requirement and HB_ERR_CAPACITY is returned (out may be NULL with capacity 0 to query). */
HB_API int hb_mesh(hb_ctx *ctx, const hb_chunk_pt *pts, size_t n, const uint16_t *ids,
            const int32_t *neighbor_index, hb_instance3d *out, uint64_t out_capacity,
            uint64_t *out_offsets, uint32_t *out_counts, uint64_t *out_total);

/* Literal drop-in of client generate_mesh (client/src/world/chunk.rs:176-223): the visible faces are READ from
 * cube.flags -- CubeFlags::faces(), server/src/chunk/cube.rs:35-41 -- exactly as the server left them (CubeMesh::set,
 * cull_shared_faces, ChunkMap::set_cube incl. its border quirks, server/src/chunk/map.rs:81-151), instead of being
 * recomputed from the materials; AO probes read material.is_some() of the 27-shell.  Use this for the incremental
 * dig/place path, where the reference's flags are not the closed-form visibility.
 * cubes: n x 32768 hb_palette_cube.  Everything else as hb_mesh. */
HB_API int hb_mesh_cubes(hb_ctx *ctx, const hb_chunk_pt *pts, size_t n, const hb_palette_cube *cubes,
                         const int32_t *neighbor_index, hb_instance3d *out, uint64_t out_capacity,
                         uint64_t *out_offsets, uint32_t *out_counts, uint64_t *out_total);

/* ---- fused generate + mesh of a region ---- */

/* Sink for hb_generate_mesh_region: called once per slab (a contiguous range of region chunk
 * indices [first_chunk, first_chunk+n_chunks)), in order, from the calling thread.  Pointers are
 * host buffers owned by the library and valid only during the call.  offsets are relative
 * to `instances`.  ids is NULL unless want_ids.
 * The sink runs while the ctx is locked: it must NOT call back into the library on the same ctx (or on a world of
 * that ctx) -- that would deadlock; hand the data to another thread or use a second ctx instead. */
typedef void (*hb_region_sink)(void *user, uint64_t first_chunk, uint64_t n_chunks,
                               const uint64_t *offsets, const uint32_t *counts,
                               const hb_instance3d *instances, uint64_t n_instances,
                               const uint16_t *ids);

/* Generate and mesh every chunk of `region` without a block-array round trip; results stream back
 * through pinned double-buffered async copies.  Equivalent to hb_generate + hb_mesh over the same
 * chunk set.  total_* (nullable) receive the sums. */
HB_API int hb_generate_mesh_region(hb_ctx *ctx, const hb_region *region, int want_ids,
                            hb_region_sink sink, void *user,
                            uint64_t *total_chunks, uint64_t *total_quads);

/* How hb_generate_mesh_region[_slab] moves a slab's quads from the device to the sink's Instance3d buffer:
 *   HB_TRANSPORT_PACKED (default): the device writes hb_packed_quad words (4 B/quad), they cross the link, and the
 *     library's host threads (hb_set_host_threads; default min(hardware threads, 32) or $HB_HOST_THREADS) rebuild the
 *     100-byte records with non-temporal stores into a library-owned host buffer -- 25x less link traffic;
 *   HB_TRANSPORT_INSTANCES: the device writes the 100-byte records and they are copied as they are (pinned buffers).
 * The sink sees byte-identical data either way.
 * want_ids follows the transport: with the packed one the slab's heights tiles cross the link (4 KiB per chunk COLUMN)
 * and the host threads write the u16 block ids from them -- the generator's ids are a function of the column's heights
 * alone (server/src/generator/mod.rs:85-91) --; with HB_TRANSPORT_INSTANCES the device's FILL stage writes them and
 * 64 KiB per chunk are copied.  Same ids either way (hb_generate_mesh_region_packed always rebuilds them on the host). */
#define HB_TRANSPORT_PACKED 0
#define HB_TRANSPORT_INSTANCES 1
HB_API int hb_set_region_transport(hb_ctx *ctx, int32_t transport);
HB_API int hb_set_host_threads(hb_ctx *ctx, int32_t n_threads);

/* Same for the x-slab ix in [ix_begin, ix_end) of `region` only: how a region is sharded across GPUs (one ctx per
 * GPU, herbolution_b200/sharding.py).  The slab's border columns see their neighbours inside the region (the
 * one-column apron of heights is recomputed from the noise function, no exchange); first_chunk passed to the sink
 * is still the index within the whole region. */
HB_API int hb_generate_mesh_region_slab(hb_ctx *ctx, const hb_region *region, int32_t ix_begin, int32_t ix_end,
                                        int want_ids, hb_region_sink sink, void *user,
                                        uint64_t *total_chunks, uint64_t *total_quads);

/* The packed transport without the expansion: the sink receives the hb_packed_quad words themselves (chunk i of the
 * slab at quads[offsets[i] .. + counts[i]), offsets multiples of 4, gaps zero) and expands what it needs, when it
 * needs it, on its own threads with hb_expand_quads -- e.g. straight into a mapped vertex buffer.  Sink rules as above. */
typedef void (*hb_region_packed_sink)(void *user, uint64_t first_chunk, uint64_t n_chunks,
                                      const uint64_t *offsets, const uint32_t *counts,
                                      const hb_packed_quad *quads, uint64_t n_quads,
                                      const uint16_t *ids);
HB_API int hb_generate_mesh_region_packed(hb_ctx *ctx, const hb_region *region, int32_t ix_begin, int32_t ix_end,
                                          int want_ids, hb_region_packed_sink sink, void *user,
                                          uint64_t *total_chunks, uint64_t *total_quads);
/* n packed quads of chunk `pt` -> out[0 .. n) (zero-filled up to the next multiple of 4 records, so `out` needs room
 * for (n + 3) & ~3 records and must be 16-byte aligned).  No GPU work, no lock: any number of host threads may call
 * it concurrently on one ctx (as long as nobody changes the materials meanwhile).  Uses the ctx's palette colours. */
HB_API int hb_expand_quads(hb_ctx *ctx, hb_chunk_pt pt, const hb_packed_quad *quads, uint32_t n, hb_instance3d *out);

/* ---- one process, several GPUs (a handle owning 1..8 devices) ----
 * The reference's host is ONE process whose chunk work fans out over a thread pool (server/src/lib.rs:42-60,
 * lib/src/task.rs:6-25); hb_multi gives that host the same shape on a multi-GPU box: one hb_ctx per device behind one
 * handle.  hb_multi_generate_mesh_region splits the region's chunk columns into contiguous, balanced x-slabs, one per
 * device in the order the devices were listed (hb_multi_slab reports the split; chunks are independent, so there is
 * no exchange: each slab recomputes its one-column apron of heights), and drives every device from its own host
 * thread with that ctx's stream pair, pinned ring and expander threads.
 * Sink: called once per streamed piece exactly as for hb_generate_mesh_region, from the per-device threads, but
 * SERIALISED by the handle (never two calls at once).  Pieces of one device arrive in region order; pieces of
 * different devices interleave.  first_chunk is the index within the whole region, so a consumer places each piece
 * without knowing which device produced it.  The sink must not call back into the handle.
 * devices = NULL with n_devices = 0 takes every visible device (at most HB_MAX_DEVICES).  The setters broadcast to all
 * devices; hb_multi_ctx exposes a device's ctx for anything else (per-device worlds, probes).  Host expander threads
 * default to min(hardware threads, 32) / n_devices per device. */
#define HB_MAX_DEVICES 8
typedef struct hb_multi hb_multi;
HB_API int hb_multi_create(const int32_t *devices, int32_t n_devices, hb_multi **out);
HB_API void hb_multi_destroy(hb_multi *multi);
HB_API int32_t hb_multi_device_count(const hb_multi *multi);
HB_API hb_ctx *hb_multi_ctx(hb_multi *multi, int32_t i);   /* owned by the handle: never pass it to hb_destroy */
HB_API int hb_multi_set_world(hb_multi *multi, int64_t seed, float freq, int32_t octaves, float lacunarity, float gain);
HB_API int hb_multi_set_noise_kind(hb_multi *multi, int32_t kind);
HB_API int hb_multi_set_fma_mode(hb_multi *multi, int32_t fused);
HB_API int hb_multi_set_materials(hb_multi *multi, int32_t n_materials, const int32_t *n_colors, const float *rgba);
HB_API int hb_multi_set_region_transport(hb_multi *multi, int32_t transport);
HB_API int hb_multi_set_host_threads(hb_multi *multi, int32_t n_threads_per_device);
HB_API int hb_multi_slab(const hb_multi *multi, const hb_region *region, int32_t i, int32_t *ix_begin, int32_t *ix_end);
HB_API int hb_multi_generate_mesh_region(hb_multi *multi, const hb_region *region, int want_ids,
                                         hb_region_sink sink, void *user,
                                         uint64_t *total_chunks, uint64_t *total_quads);

/* The host half of the packed transports without a context: no CUDA call is made, so a consumer on another machine
 * (or a test on a machine without a GPU) can rebuild the records and the block ids.
 * hb_host_expand_quads = hb_expand_quads with the palette passed in (n_colors / rgba as for hb_set_materials; NULL = the
 * game's stone, dirt, grass); the tables are rebuilt per call, so pass whole chunks.
 * hb_host_fill_ids: the 32 768 block ids of the chunk at layer cy of a column from the column's heights tile
 * (int32[1024], index x*32 + z, as hb_noise_tiles returns it): GenerationParams::generate's threshold rule
 * (server/src/generator/mod.rs:85-91).  Both outputs 16-byte aligned. */
HB_API int hb_host_expand_quads(hb_chunk_pt pt, const hb_packed_quad *quads, uint32_t n, int32_t n_materials,
                                const int32_t *n_colors, const float *rgba, hb_instance3d *out);
HB_API int hb_host_fill_ids(const int32_t *heights, int32_t cy, uint16_t *out_ids);

/* Measurement helpers: the two ceilings of the end-to-end path, measured with the ctx's own buffers and threads so
 * that several ctxs (GPUs) can be probed at the same moment.  hb_probe_d2h: plain cudaMemcpyAsync device -> pinned
 * host of `bytes`, `reps` times back to back (GB/s by CUDA events).  hb_probe_host_write: the expander's worker pool
 * filling `bytes` of the expanded-slab buffer with non-temporal stores (GB/s by wall clock). */
HB_API int hb_probe_d2h(hb_ctx *ctx, size_t bytes, int32_t reps, double *gbs);
HB_API int hb_probe_host_write(hb_ctx *ctx, size_t bytes, int32_t reps, double *gbs);

/* ---- device-resident pipeline (kernel-only measurement, chaining with other CUDA work) ----
 * A plan owns the device buffers for region chunks with ix in [ix_begin, ix_end).  Stages only
 * ENQUEUE kernels on `stream` (a cudaStream_t passed as void*; NULL = the ctx stream). */
typedef struct hb_region_plan hb_region_plan;
enum {
    HB_STAGE_HEIGHTS = 1,  /* noise -> heights tiles */
    HB_STAGE_COUNT = 2,    /* visible faces per chunk */
    HB_STAGE_SCAN = 4,     /* exclusive scan -> offsets */
    HB_STAGE_EMIT = 8,     /* write hb_instance3d */
    HB_STAGE_FILL = 16,    /* write u16 block ids (plan created with want_ids) */
    HB_STAGE_EMIT_PACKED = 32, /* write hb_packed_quad words at the same offsets (hb_region_plan_packed) */
    HB_STAGE_FUSED = 64,   /* HEIGHTS + COUNT + SCAN + EMIT as ONE kernel launch (noise evaluated per column incl. its
                              apron, offsets by a chained scan over the columns); same results, bit for bit.  The
                              heights buffer then holds the slab's own columns only (enough for FILL).  Regions
                              with more than 32 layers or a non-FBM noise kind run the four staged kernels instead. */
    HB_STAGE_ALL_MESH = 15
};
/* quad_capacity 0 = size the instance buffer by running HEIGHTS+COUNT+SCAN once (synchronous). */
HB_API int hb_region_plan_create(hb_ctx *ctx, const hb_region *region, int32_t ix_begin, int32_t ix_end,
                          uint64_t quad_capacity, int want_ids, hb_region_plan **out);
HB_API void hb_region_plan_destroy(hb_region_plan *plan);
HB_API int hb_region_plan_enqueue(hb_region_plan *plan, int stages, void *stream);
/* Synchronises `stream`, then reports totals and device pointers (any out pointer may be NULL). */
HB_API int hb_region_plan_result(hb_region_plan *plan, void *stream, uint64_t *n_chunks, uint64_t *total_quads,
                          uint64_t *total_span, const hb_instance3d **d_instances,
                          const uint64_t **d_offsets, const uint32_t **d_counts,
                          const uint16_t **d_ids, const int32_t **d_heights);
/* Device pointer of the packed words HB_STAGE_EMIT_PACKED wrote (NULL before the first such enqueue). */
HB_API int hb_region_plan_packed(hb_region_plan *plan, const hb_packed_quad **d_packed);
/* Number of kernel launches a given stage mask enqueues (for launch accounting). */
HB_API int hb_region_plan_launches(const hb_region_plan *plan, int stages);
/* Blocking device->host copy of plan / hb_dev_* results (synchronises the ctx device). */
HB_API int hb_dev_download(hb_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes);

/* Device-pointer forms of hb_generate / hb_mesh: all pointers are device memory, kernels are
 * enqueued on `stream`, nothing is synchronised.
 * hb_dev_generate: d_cols_xz (ncol x 2), d_col_of_chunk (n), d_heights (ncol x 1024 i32 scratch/out). */
HB_API int hb_dev_generate(hb_ctx *ctx, void *stream, const int32_t *d_cols_xz, size_t ncol,
                    const hb_chunk_pt *d_pts, const int32_t *d_col_of_chunk, size_t n,
                    int32_t *d_heights, float *d_noise, uint16_t *d_ids, hb_palette_cube *d_cubes);
/* hb_dev_mesh: d_workspace of hb_dev_mesh_workspace_bytes(n) bytes (about 5.1 KiB per chunk: occupancy bitmaps, border
 * records, scan scratch and the emit work list; contents are scratch between calls).  d_total[0] = span (instances),
 * d_total[1] = quads, d_total[2] = error flags (1 = capacity exceeded, 2 = unknown material id), d_total[3] = scratch
 * (the emit kernel's work-ticket counter); d_total is 4 x u64. */
HB_API size_t hb_dev_mesh_workspace_bytes(size_t n);
HB_API int hb_dev_mesh(hb_ctx *ctx, void *stream, const hb_chunk_pt *d_pts, size_t n, const uint16_t *d_ids,
                const int32_t *d_neighbor_index, void *d_workspace, hb_instance3d *d_out,
                uint64_t out_capacity, uint64_t *d_offsets, uint32_t *d_counts, uint64_t *d_total);

/* ------------------------------------------------------------------------------------------------
 * Device-resident world: the server's ChunkMap (server/src/chunk/map.rs) with its PaletteCube arrays kept
 * in HBM, the incremental dig/place path, the server->client CubeUpdate wire format and the client's
 * remesh-what-changed loop.  One hb_world serves both sides: the client's
 * ChunkData is a copy of the server's cubes kept current by CubeUpdates (client/src/world/chunk.rs:132-149),
 * so meshing from the resident cubes after a sync equals meshing the client's copy.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hb_world hb_world;

/* == ChunkCube (server/src/chunk/handle.rs:28-32): vec3u5 position + PaletteCube.  The Rust struct is
 * repr(Rust) and never leaves the process, so this 12-byte C layout is the binding's own. */
typedef struct hb_chunk_cube {
    uint16_t pos;   /* vec3u5 bits: 1 | x<<11 | y<<6 | z<<1 (lib/src/vector.rs:1181-1193) */
    uint16_t _pad;
    hb_palette_cube cube;
} hb_chunk_cube;

/* A store of `capacity` chunk slots (292 KiB of HBM each: cubes, update sets, mesher bitmaps).  A world borrows its
 * ctx (stream, tables, mutex): destroy every world before hb_destroy(ctx).  Calls on one world are serialised. */
HB_API int hb_world_create(hb_ctx *ctx, size_t capacity, hb_world **out);
HB_API void hb_world_destroy(hb_world *world);
HB_API size_t hb_world_len(const hb_world *world);          /* chunks currently loaded */

/* ChunkMap::queue_load + ChunkGenerator + ChunkMap::load_provided (server/src/chunk/map.rs:69-75,203-228,
 * server/src/generator/mod.rs:32-95): generates every not-yet-loaded chunk of pts[0..n) straight into its slot
 * (PaletteCube image as CubeMesh::set leaves it), then runs cull_shared_faces (server/src/chunk/mesh.rs:76-123)
 * for every (new chunk, loaded face neighbour) pair.  Already-loaded and repeated positions are ignored, as
 * queue_load does.  HB_ERR_CAPACITY (nothing changed) if the free slots do not suffice.  Batches of 1024 or more
 * new chunks wire their neighbour rows on the context's host threads (hb_set_host_threads). */
HB_API int hb_world_load(hb_world *world, const hb_chunk_pt *pts, size_t n);
/* ChunkMap::unload_requested (map.rs:230-235): the slot is freed; neighbours keep whatever faces the earlier
 * cull removed, exactly as the reference leaves them. */
HB_API int hb_world_unload(hb_world *world, const hb_chunk_pt *pts, size_t n);

/* ChunkMap::set_cube (map.rs:81-151) applied to n edits IN ORDER: cubes_xyz = n x 3 world cube coordinates
 * (chunk = c >> 5, local = c & 31, lib/src/point.rs:25-32), materials[i] = 0 (None: dig) or a palette id.
 * Includes the reference's border behaviour: the face of the adjacent chunk's border voxel is inserted
 * unconditionally, and edits to a position whose own chunk is not loaded still touch loaded neighbours. */
HB_API int hb_world_set_cubes(hb_world *world, const int32_t *cubes_xyz, const uint16_t *materials, size_t n);
/* How many conflict groups the last hb_world_set_cubes batch was split into (edits whose 7-voxel stencils -- the voxel
 * and its six neighbours -- share no voxel commute and are applied concurrently, one warp per group; inside a group
 * the caller's order is kept, so the result equals the reference's sequential ChunkMap::set_cube calls). */
HB_API uint32_t hb_world_last_edit_groups(const hb_world *world);
/* The grouping itself, without a world or a GPU (host code; what the CPU tests check against a brute-force graph):
 * group_of[i] = conflict group of edit i of the batch (one sub-batch: hb_world_set_cubes groups every 65 536 edits
 * separately), groups numbered by their first edit. */
HB_API int hb_host_group_edits(const int32_t *cubes_xyz, size_t n, uint32_t *group_of, uint32_t *n_groups);

/* Chunk::sync_with_client (server/src/chunk/mod.rs:35-67) for every chunk: drains each chunk's
 * updated_positions.  Chunks with a non-empty set join the remesh set (they are the chunks whose client copy
 * receives a CubeUpdate, client/src/world/chunk.rs:52-63).  Reports how many chunks were dirty and how many
 * hb_chunk_cube entries hb_world_fetch_updates would deliver.  The reference keeps updated_positions as a Vec with
 * duplicates; here it is a set (one bit per voxel), which the client cannot tell apart: apply_updates_from_server
 * overwrites cubes[pos] with the same final cube for every duplicate. */
HB_API int hb_world_sync(hb_world *world, size_t *n_dirty, uint64_t *n_entries);
/* The CubeUpdates of the last hb_world_sync, one run per dirty chunk: out_pts[i] = chunk, its entries are
 * out_entries[out_offsets[i] .. +out_counts[i]) in ascending voxel index, each carrying the cube as it is now.
 * Optional (a fused server+client can skip the wire format and go straight to hb_world_remesh).
 * *n_runs / *n_entries (required) receive how many runs / entries there are: chunks unloaded since the sync are
 * skipped, so this can be fewer than hb_world_sync reported -- only out_*[0 .. *n_runs) are written.
 * HB_ERR_CAPACITY (nothing written, *n_runs / *n_entries hold the requirement) if pts_capacity / entries_capacity
 * are too small.  out_entries from hb_host_alloc (pinned) is one DMA at link speed; pageable memory is accepted
 * and streamed through pinned bounce buffers. */
HB_API int hb_world_fetch_updates(hb_world *world, hb_chunk_pt *out_pts, size_t pts_capacity, uint64_t *out_offsets,
                           uint32_t *out_counts, hb_chunk_cube *out_entries, uint64_t entries_capacity,
                           size_t *n_runs, uint64_t *n_entries);

/* Client ChunkMap::update (client/src/world/chunk.rs:47-76): generate_mesh (faces read from cube.flags, AO and
 * occupancy through the 27-chunk shell of loaded chunks) for every chunk of the remesh set that is still loaded.
 * Two-call pattern like hb_mesh: with too small a pts_capacity / out_capacity the call reports *n_chunks and
 * *out_total, returns HB_ERR_CAPACITY and keeps the remesh set; on HB_OK the set is cleared.  Chunk i's
 * instances are out[out_offsets[i] .. +out_counts[i]) (offsets are multiples of 4, gaps zero-filled). */
HB_API int hb_world_remesh(hb_world *world, hb_chunk_pt *out_pts, size_t pts_capacity, hb_instance3d *out,
                    uint64_t out_capacity, uint64_t *out_offsets, uint32_t *out_counts, size_t *n_chunks,
                    uint64_t *out_total);
/* Same, leaving the instances in device memory (valid until the next hb_world_* call on this world). */
HB_API int hb_world_remesh_dev(hb_world *world, hb_chunk_pt *out_pts, size_t pts_capacity, uint64_t *out_offsets,
                        uint32_t *out_counts, size_t *n_chunks, uint64_t *out_total,
                        const hb_instance3d **d_instances);
/* Copy one loaded chunk's PaletteCube array (32768 x 8 B) to the host; HB_ERR_INVALID if it is not loaded. */
HB_API int hb_world_read_chunk(hb_world *world, hb_chunk_pt pt, hb_palette_cube *out_cubes);

/* ------------------------------------------------------------------------------------------------
 * On-disk cube codec (server/src/chunk/codec.rs): the run-length records of
 * encode_cubes (codec.rs:73-101) and the record loop of CubeGrid::decode (codec.rs:52-66), bit for bit.
 * A record is [u8 count][u16 LE id]; a run is cut every 255 voxels; the encoder starts in state (None, 0), so a
 * chunk whose first voxel has a material begins with an empty record.  `ids` are the raw in-memory ids used
 * everywhere in this library (0 = None, k = palette index k-1).  The reference writes PaletteMaterialId::to_u16
 * (the 0-based index) and 0 for None, and reads back NonZeroU16::new(v + 1): None and the first material share wire
 * value 0 and come back as the first material -- the codec is lossy by construction (and unused: no writer is ever
 * called, server/src/chunk/provider.rs:48-58); it is reproduced here as it is, not repaired.  The palette half of
 * CubeGrid::encode is hb_encode_palette below (CubeGrid::decode does not skip it -- another reason the reference's
 * codec cannot round-trip; the bytes are reproduced, the defect is not repaired).
 * ---------------------------------------------------------------------------------------------- */
/* encode_palette (codec.rs:70-75) + Material::encode (server/src/chunk/material.rs:73-99): for palette entry i
 * (0-based, the ctx's materials in id order)
 *   [u16 LE i][u8 len(group)][u8 len(key)][group bytes][key bytes][u8 (cullable_faces << 6) | has_collider]
 *   [u8 0 = Texture::Colors][u16 LE n_colors][n_colors x (r, g, b, a) f32 LE][f32 LE toughness]
 * (the shift is done in u8 as in the reference, so only the low two face bits survive).  The colours are the ctx's
 * (hb_set_materials); everything else comes from `descs`, one per material, or -- descs = NULL, only valid while the
 * ctx holds the default three materials -- from Material::stone / dirt / grass (material.rs:27-61).
 * CubeGrid::encode(chunk) = these bytes followed by the chunk's hb_rle_encode stream.  Pure host code.
 * Two-call pattern: *n_bytes is always set; HB_ERR_CAPACITY if capacity is too small. */
typedef struct hb_material_desc {
    const char *group, *key;   /* GroupKeyBuf, e.g. "herbolution", "stone" (each at most 255 bytes) */
    uint8_t cullable_faces;    /* CubeFaces bits, 0x3F = all */
    uint8_t has_collider;
    float toughness;
} hb_material_desc;
HB_API int hb_encode_palette(hb_ctx *ctx, const hb_material_desc *descs, uint8_t *out, size_t capacity, size_t *n_bytes);
/* ids: n x 32768.  Chunk i's stream is out[out_offsets[i] .. + out_sizes[i]) (bytes; starts are multiples of 12,
 * gaps zero-filled); *out_total = bytes spanned.  Two-call pattern: HB_ERR_CAPACITY reports offsets, sizes, total. */
HB_API int hb_rle_encode(hb_ctx *ctx, const uint16_t *ids, size_t n, uint8_t *out, uint64_t out_capacity,
                  uint64_t *out_offsets, uint32_t *out_sizes, uint64_t *out_total);
/* Stream i = bytes[offsets[i] .. + sizes[i]) -> out_ids[i][32768].  A partial trailing record is ignored, voxels
 * the stream does not reach are None, wire 65535 wraps to None (release-build arithmetic); a stream that decodes to
 * more than 32768 voxels is HB_ERR_INVALID (the reference panics). */
HB_API int hb_rle_decode(hb_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, const uint64_t *offsets,
                  const uint32_t *sizes, size_t n, uint16_t *out_ids);
/* Device-pointer form of the encoder: counts -> offsets (in records; x3 for bytes) -> records if d_out.
 * d_scan_workspace as for hb_dev_mesh; d_total (4 x u64): [0] = span in records, [2] & 1 = capacity exceeded, [3] = scratch (work tickets).  d_ids must be
 * 16-byte aligned and d_out 4-byte aligned (any cudaMalloc pointer is); d_out should be zeroed first if the gaps matter. */
HB_API int hb_dev_rle_encode(hb_ctx *ctx, void *stream, const uint16_t *d_ids, size_t n, void *d_scan_workspace,
                      uint8_t *d_out, uint64_t capacity_records, uint64_t *d_offsets, uint32_t *d_counts,
                      uint64_t *d_total);

#ifdef __cplusplus
}
#endif
#endif /* HERBGPU_H */
