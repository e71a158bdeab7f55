/*
 * voxplat_b200.h -- C ABI of the B200-native chunk-rebuild path.
 *
 * This is the drop-in boundary for ONE path of the Voxplat engine (reference: kosshi-net/voxplat):
 * RLE decode/encode -> exposed-voxel cull (with cross-chunk halos) -> 5-level LOD reduction -> splat-list
 * and near-field quad-mesh buffers.  Every entry point below names the reference interface it replaces
 * (file:line in the reference tree).  Plain pointers and sizes only; no CUDA, torch or C++ types.
 *
 * All byte formats are the reference's:
 *   voxels   uint8, chunk-local index (z<<2b | y<<b | x), b = root_bitw          chunkset.h:171-188
 *   chunk id (cz<<by | cy)<<bx | cx over the chunk grid                          chunkset.c:124-126
 *   RLE      uint32 LE words: run (24 bit) | value<<24, terminated by a 0 word   rle.c:17-26
 *   splats   int16 x,y,z,colour|shadow<<6 ; per chunk [L0|L1|L2|L3|L4]           mesher.c:497-536, chunkset.c:380-458
 *   mesh     int16 x,y,z,data per vertex (4 per quad), uint32 indices (6 per quad) mesher.c:321-349
 *   shadow   uint16 height map, (X+Y) entries per z row                          shadow.h:26-51
 *
 * Threading: a vp_ctx may be used from any ONE host thread at a time (the reference's mesher thread,
 * game.c:77-89); it owns its device buffers, streams and pinned staging.  No GL context is needed.
 * Errors: every function returns 0 on success or a negative vp_status; vp_last_error() gives text.
 * There is NO CPU fallback: without a CUDA device vp_ctx_create fails with VP_ERR_NO_DEVICE.
 */
#ifndef VOXPLAT_B200_H
#define VOXPLAT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VP_API __attribute__((visibility("default")))
#else
#define VP_API
#endif

#define VP_MAX_LOD_LEVEL 5              /* chunkset.h:11 MAX_LOD_LEVEL */

typedef enum {
	VP_OK             =  0,
	VP_ERR_ARG        = -1,             /* bad argument / unsupported geometry */
	VP_ERR_NO_DEVICE  = -2,             /* no CUDA device: the product never falls back to the CPU */
	VP_ERR_CUDA       = -3,             /* a CUDA runtime call failed */
	VP_ERR_ARENA_FULL = -4,             /* output arena too small; counts are valid, grow and retry */
	VP_ERR_RLE        = -5,             /* malformed RLE stream (length != chunk volume) */
	VP_ERR_NOT_RESIDENT = -6,           /* chunk id outside this context's slab */
	VP_ERR_IO         = -7              /* world file missing, truncated or not a VOXPLAT file */
} vp_status;

/* Rebuild flags: which of the dispatcher's two branches to run (chunkset.c:337 `if (c->make_mesh)`). */
#define VP_REBUILD_SPLAT 1u             /* chunk_make_mask + 5x chunk_make_splatlist + 4x chunk_mask_downsample */
#define VP_REBUILD_MESH  2u             /* chunk_make_mesh */

typedef struct vp_ctx vp_ctx;

/* World geometry = the fields of struct ChunkSet the path reads (chunkset.h:29-54) plus the slab of
 * chunk rows this context owns when the world is sharded over several GPUs (z is the slowest chunk
 * index, so a slab is a contiguous chunk-id range). */
typedef struct {
	int32_t  device;                /* CUDA device ordinal */
	int32_t  root_bitw;             /* ChunkSet.root_bitw: 4..7 (chunk edge 16..128) */
	int32_t  max_bitw[3];           /* ChunkSet.max_bitw: chunk grid of the WHOLE world */
	int32_t  slab_z0, slab_z1;      /* owned chunk rows [z0,z1); 0,0 = whole world */
	uint64_t splat_arena_bytes;     /* device arena for splat lists; 0 = auto */
	uint64_t mesh_arena_bytes;      /* device arena for VBO+IBO; 0 = auto */
	uint64_t rle_arena_bytes;       /* device arena for RLE streams in/out; 0 = auto */
} vp_config;

/* Per-chunk result of one rebuild = what chunkset_manage publishes into struct ChunkMD
 * (chunkset.c:347-366 mesh branch, :463-501 splat branch).  Offsets are bytes into the arena the
 * call wrote (host staging for vp_rebuild_batch, device arena for vp_rebuild_device). */
typedef struct {
	uint64_t svl_offset;                      /* ChunkMD.svl                       */
	uint32_t svl_items[VP_MAX_LOD_LEVEL];     /* ChunkMD.svl_items[5] (int16 units) */
	uint32_t svl_items_total;                 /* ChunkMD.svl_items_total           */
	uint64_t vbo_offset;                      /* ChunkMD.mesh_vbo                  */
	uint64_t ibo_offset;                      /* ChunkMD.mesh_ibo                  */
	uint32_t vbo_items;                       /* ChunkMD.mesh_vbo_items (int16 units)  */
	uint32_t ibo_items;                       /* ChunkMD.mesh_ibo_items (uint32 units) */
} vp_chunk_result;

/* ---- context ------------------------------------------------------------------------------- */

/* Replaces chunkset_create + chunkset_clear + shadow_init (chunkset.c:31-121, shadow.h:26-43) for the
 * device-resident copy of the world: all chunks start as the null (all-air) chunk, the shadow map is
 * zero-filled and padded (SURVEY 8a' u2/u3). */
VP_API int  vp_ctx_create(const vp_config *cfg, vp_ctx **out);
VP_API void vp_ctx_destroy(vp_ctx *ctx);
VP_API const char *vp_last_error(const vp_ctx *ctx);          /* ctx may be NULL for creation errors */
VP_API const char *vp_version(void);

/* Re-allocate the output arenas (0 = keep).  Use after VP_ERR_ARENA_FULL: the byte counts reported by
 * vp_rebuild_device_results are the sizes the batch needs. */
VP_API int  vp_ctx_resize_arenas(vp_ctx *ctx, uint64_t splat_bytes, uint64_t mesh_bytes);

/* Run all work of this context on an externally owned CUDA stream (a cudaStream_t passed as void*),
 * e.g. the caller's torch stream, so that the caller's events bracket the kernels.  NULL restores the
 * context's own stream. */
VP_API int  vp_ctx_set_stream(vp_ctx *ctx, void *cuda_stream);
VP_API int  vp_ctx_synchronize(vp_ctx *ctx);
/* Counters for the bench: kernels launched by this library since the last reset. */
VP_API uint64_t vp_kernel_launches(vp_ctx *ctx, int reset);

/* ---- world residency (host -> device) -------------------------------------------------------- */

/* Dense upload of n chunks, host_dense = n * R^3 bytes.  Replaces chunk_open_rw + write + chunk_close_rw
 * (chunkset.c:167-204) for the device copy.  A chunk whose bytes are all zero becomes the null chunk,
 * like chunk_compress's all-air test (chunkset.c:225-228). */
VP_API int  vp_upload_chunks_dense(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n, const uint8_t *host_dense);

/* RLE upload + device decode: rle_decompress (rle.c:90-116) for n chunks at once.  words = the n streams
 * back to back (each with its 0 terminator); word_offsets[n+1] = start of each stream in `words`.
 * A stream equal to the all-air stream {R^3, 0} makes the chunk null (chunkset.c:144-145). */
VP_API int  vp_upload_chunks_rle(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n,
                          const uint32_t *words, const uint64_t *word_offsets);

/* resident[i] = 1 when the chunk holds voxels on the device, 0 when it is the null chunk (the reference's test is
 * `c->rle == set->null_chunk->rle`, chunkset.c:144-145).  Host-side table lookup, no device work. */
VP_API int  vp_chunks_resident(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n, uint8_t *resident);

/* Make chunks the null chunk (chunkset_clear, chunkset.c:116-117). */
VP_API int  vp_set_chunks_null(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n);

/* Read chunks back as dense bytes (n * R^3); null chunks read as zeros. */
VP_API int  vp_download_chunks_dense(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n, uint8_t *host_dense);

/* rle_compress (rle.c:44-87) of n resident chunks on the device.  On return word_offsets[n+1] holds the
 * start of every stream inside `words`; total words = word_offsets[n].  VP_ERR_ARENA_FULL if cap_words
 * is too small (word_offsets is still filled so the caller can size the buffer). */
VP_API int  vp_encode_chunks_rle(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n,
                          uint32_t *words, uint64_t cap_words, uint64_t *word_offsets);

/* Shadow map rows [z0,z1) in world voxel rows, (X+Y) uint16 each (shadow.h:26-51).  Rows outside the
 * context's slab (+17 rows of reach) are ignored. */
VP_API int  vp_upload_shadow_rows(vp_ctx *ctx, uint32_t z0, uint32_t z1, const uint16_t *rows);
/* The same without waiting for the copy: the rows travel on the context stream in front of whatever is enqueued next
 * (e.g. vp_rebuild_from_rle).  `rows` must be page-locked memory and stay valid until the next synchronous call. */
VP_API int  vp_upload_shadow_rows_async(vp_ctx *ctx, uint32_t z0, uint32_t z1, const uint16_t *rows);

/* ---- flat RLE codec: drop-in bodies for rle_compress / rle_decompress (rle.h:7-8) ------------- */

/* Encode `length` bytes (any length >= 1, like rle.h:7; a run stops at 0xFFFFFF like rle.c:62); writes at most
 * cap_words words (incl. terminator); *n_words = words needed. */
VP_API int  vp_rle_compress(vp_ctx *ctx, const uint8_t *data, uint32_t length,
                     uint32_t *out_words, uint32_t cap_words, uint32_t *n_words);
/* Decode a 0-terminated stream of n_words words (incl. terminator); *n_bytes = bytes produced. */
VP_API int  vp_rle_decompress(vp_ctx *ctx, const uint32_t *words, uint32_t n_words,
                       uint8_t *out, uint32_t cap_bytes, uint32_t *n_bytes);

/* ---- rebuild: the loop body of chunkset_manage (chunkset.c:318-505) for a batch of chunks ------- */

/* Synchronous host-facing call.  For each chunk: VP_REBUILD_SPLAT fills svl_* (five LOD segments,
 * byte-identical to chunk_make_mask/downsample/splatlist), VP_REBUILD_MESH fills vbo_/ibo_*
 * (byte-identical to chunk_make_mesh).  `flags` applies to all chunks; if per_chunk_flags != NULL it
 * overrides per chunk (the reference picks by ChunkMD.make_mesh).  Output bytes are in pinned host
 * staging owned by ctx, valid until the next rebuild call: *splat_base + svl_offset, *mesh_base +
 * vbo_offset / ibo_offset. */
VP_API int  vp_rebuild_batch(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n, uint32_t flags,
                      const uint8_t *per_chunk_flags, vp_chunk_result *results,
                      const void **splat_base, const void **mesh_base);

/* The whole end-to-end step in one call: host RLE streams in (as vp_upload_chunks_rle), rebuild (as
 * vp_rebuild_batch), host buffers out -- internally pipelined over `n_blocks` blocks of ascending chunk ids on three
 * streams (upload+decode | kernels | download) so PCIe runs in both directions at once.  n_blocks = 1, or ids not
 * ascending, degenerates to the sequential order.  Results are identical to the two separate calls. */
VP_API int  vp_rebuild_from_rle(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n, const uint32_t *words,
                         const uint64_t *word_offsets, uint32_t flags, const uint8_t *per_chunk_flags, uint32_t n_blocks,
                         vp_chunk_result *results, const void **splat_base, const void **mesh_base);

/* Asynchronous device-resident variant (what the bench times as `value`): chunk ids are taken from
 * the host array once (vp_batch_prepare), kernels are enqueued on the context stream, outputs stay
 * in the device arenas.  vp_rebuild_device_results copies the per-chunk records back. */
/* vp_batch_prepare must be repeated after anything that turns a chunk null or resident (uploads, vp_set_chunks_null,
 * vp_edit_sphere, vp_generate_world, vp_world_load): the prepared lists leave out chunks that had nothing to show;
 * vp_rebuild_device fails with VP_ERR_ARG otherwise. */
VP_API int  vp_batch_prepare(vp_ctx *ctx, const uint32_t *chunk_ids, uint32_t n, const uint8_t *per_chunk_flags,
                      uint32_t flags);
VP_API int  vp_rebuild_device(vp_ctx *ctx);
/* The same in two parts, for slab contexts that exchange border planes every step: part 0 launches the chunks that do
 * not read a ghost row (all but the slab's last chunk row for splats, plus the first row for meshes), part 1 the
 * others -- call it after vp_halo_unpack, so the exchange overlaps part 0.  vp_rebuild_device == part 0 + part 1. */
VP_API int  vp_rebuild_device_part(vp_ctx *ctx, int part);
VP_API int  vp_rebuild_device_results(vp_ctx *ctx, vp_chunk_result *results, uint64_t *splat_bytes, uint64_t *mesh_bytes);
/* Device time of the kernels of the last n (<= 256) vp_rebuild_device calls, oldest first (CUDA events on the
 * context stream): splat_ms[k] = cull+LOD+splat kernel, mesh_ms[k] = mesh kernel (0 if not launched).
 * Synchronises the stream. */
VP_API int  vp_kernel_ms_history(vp_ctx *ctx, uint32_t n, float *splat_ms, float *mesh_ms);
/* Device pointers of the arenas (for zero-copy consumers / tests). */
VP_API void *vp_splat_arena_device(vp_ctx *ctx);
VP_API void *vp_mesh_arena_device(vp_ctx *ctx);
/* Copy [0,bytes) of an arena to host memory (which: 0 splat, 1 mesh). */
VP_API int  vp_arena_download(vp_ctx *ctx, int which, void *host_dst, uint64_t bytes);

/* ---- single-chunk wrappers with the reference's own signatures' meaning (mesher.h:7-37) -------- */

/* chunk_make_mask + chunk_make_splatlist(level 0..4) for one chunk: geometry receives the five
 * segments, items[5] the int16 counts.  Returns total int16 items or a negative vp_status. */
VP_API int64_t vp_chunk_make_splatlists(vp_ctx *ctx, uint32_t chunk_id, int16_t *geometry, uint64_t cap_items,
                                 uint32_t items[VP_MAX_LOD_LEVEL]);
/* chunk_make_mesh for one chunk. */
VP_API int  vp_chunk_make_mesh(vp_ctx *ctx, uint32_t chunk_id, int16_t *geometry, uint64_t cap_geometry_items,
                        uint32_t *geometry_items, uint32_t *index, uint64_t cap_index_items, uint32_t *index_items);

/* ---- device-side edits ------------------------------------------------------------ */

/* chunkset_edit_sphere (chunkset/edit.c:179-244) on the device copy: writes `voxel` into every cell closer than
 * `radius` to (x,y,z) (cells with y < 2 are protected, edit.c:151), applies shadow_place_update (shadow.h:77-89) in
 * the reference's order when voxel != 0, refreshes the x-face planes.  dirty_ids receives the chunks the reference
 * marks dirty (all chunks of the box [c-r-1, c+r+1], in its list order); *n_dirty their number.  The caller
 * rebuilds them with vp_rebuild_batch.  No voxel data crosses PCIe. */
VP_API int  vp_edit_sphere(vp_ctx *ctx, int32_t x, int32_t y, int32_t z, uint32_t radius, uint8_t voxel,
                    uint32_t *dirty_ids, uint32_t cap, uint32_t *n_dirty);
/* chunkset_edit_raycast_until_solid (chunkset/edit.c:248-314; the pick ray of game.c:212) for n rays at once on the
 * device copy: origins / vectors are n x 3 floats; coords (n x 3) receives the cell where each walk ended, voxels (n) the
 * voxel hit (0 = nothing within 4095 steps), normals (n x 3, in/out) gets +1 / -1 on the axis of the last step of a
 * hit and is left alone otherwise, exactly like the reference's output arguments. */
VP_API int  vp_raycast(vp_ctx *ctx, uint32_t n, const float *origins, const float *vectors, uint32_t *coords, int8_t *normals,
                uint8_t *voxels);
/* Read height-map rows [z0,z1) back (the device copy is authoritative after vp_edit_sphere). */
VP_API int  vp_download_shadow_rows(vp_ctx *ctx, uint32_t z0, uint32_t z1, uint16_t *rows);

/* ---- LOD-node aggregation: the consumer right after the path ------------------------ */

/* One octree node of LOD level `lod` = what gfx_update_svl builds into set->gsvl[lod][node] (gfx/vsplat.c:209-323):
 * the level-`lod` splat segments of all member chunks, concatenated x outer / y / z inner. */
typedef struct {
	uint64_t offset;                /* bytes into the node buffer returned in *base */
	uint32_t items;                 /* GeometrySVL.vbo_items (int16 units) */
	uint32_t members;               /* chunks under the node */
} vp_node_result;

/* Gather every node of level `lod` on the device from the splat lists of the LAST rebuild (which must have covered
 * all chunks of the world in id order).  nodes[flatten3(chunk offset >> lod, max_bitw - min(lod, max_bitw))];
 * *n_nodes = number of nodes of the level.  base == NULL keeps the node buffers on the device only;
 * kernel_ms (optional) receives the device time of the gather. */
VP_API int  vp_build_lod_nodes(vp_ctx *ctx, uint32_t lod, vp_node_result *nodes, uint32_t cap_nodes, uint32_t *n_nodes,
                        const void **base, float *kernel_ms);

/* ---- world generation on the device ------------------------------------------------- */

/* Fill the context's slab with the deterministic synthetic world of seed `seed` without any host -> device voxel
 * traffic: terrain + trees of the integer generator (csrc/vp_worldgen_core.h: the structure of chunkset/gen.c:89-184,
 * 305-323 with a pinned integer noise; gen.c's own arithmetic lives in the un-vendored FastNoise), all-air chunks become
 * null chunks (chunkset.c:225-228), and the height map is built on the device by the shadow_place_update rule
 * (shadow.h:77-89) in a fixed order.  Same resident state as vp_upload_chunks_dense + vp_upload_shadow_rows of the host
 * generator (libvpworldgen.so), byte for byte.  Worlds taller than 1024 voxels: VP_ERR_ARG. */
VP_API int  vp_generate_world(vp_ctx *ctx, uint32_t seed);

/* ---- world file: checkpoint / resume ------------------------------------------------ */

/* The layout of the reference's exporter command_export (src/deadcode.c:320-350): byte 0x89 + "VOXPLAT", root_bitw
 * (1 byte), max_bitw (3 bytes), every chunk's RLE stream in chunk-id order (rle_compress words with their 0
 * terminator, rle.c:44-87; null chunks as the all-air stream {R^3, 0}, chunkset.c:98), then the whole shadow map,
 * (X+Y)*Z uint16 (shadow.h:26-51).  Streams are encoded / decoded on the device.  The context must hold the whole
 * world (slab = all chunk rows).  vp_world_load requires a context created with the file's root_bitw / max_bitw
 * (vp_world_file_info reads them without a device). */
VP_API int  vp_world_save(vp_ctx *ctx, const char *path, uint64_t *bytes_written);
VP_API int  vp_world_load(vp_ctx *ctx, const char *path);
VP_API int  vp_world_file_info(const char *path, int32_t *root_bitw, int32_t max_bitw[3], uint64_t *file_bytes);

/* ---- multi-GPU slab borders (new: the reference is single-process) ----------------------------- */

/* A border plane = for every chunk column (cx,cy) of one chunk row, one R*R-byte z-slice, packed
 * (cy*nx + cx)*R*R, nx*ny*R*R bytes in total.  which: 0 = z-slice 0 of the context's FIRST owned chunk
 * row (send to the rank below: its +z halo), 1 = z-slice R-1 of the LAST owned row (send to the rank
 * above: its -z halo, needed by mesh AO only).  `device_buf` is a device pointer (e.g. a torch tensor's
 * data_ptr) that NCCL sends/receives; pack/unpack run on the context stream. */
VP_API uint64_t vp_halo_plane_bytes(vp_ctx *ctx);
/* The border work of a slab context -- vp_halo_pack, vp_halo_unpack and vp_rebuild_device_part(ctx, 1) -- runs on the
 * context's border stream (a cudaStream_t returned as void*), beside the interior chunks on the context stream.  The
 * caller's transport (NCCL send / receive of the packed planes) has to be ordered against THIS stream: after the pack,
 * before the unpack.  vp_rebuild_device_part(ctx, 1) joins the border stream back into the context stream. */
VP_API void *vp_ctx_border_stream(vp_ctx *ctx);
VP_API int  vp_halo_pack(vp_ctx *ctx, int which, void *device_buf);
/* which: 0 = plane received from the rank ABOVE (becomes z-slice 0 of ghost row slab_z1),
 *        1 = plane received from the rank BELOW (becomes z-slice R-1 of ghost row slab_z0-1). */
VP_API int  vp_halo_unpack(vp_ctx *ctx, int which, const void *device_buf);

/* ---- several GPUs behind one handle, one host thread (new: the reference is single-device) ------------ */

/* The world is cut into z-slabs of chunk rows, one vp_ctx per device (devices[i], or 0..ndev-1 when devices is NULL;
 * ndev must divide the chunk rows).  Border planes are written by one kernel per plane straight into the neighbour's
 * ghost chunks over NVLink (peer access), no NCCL.  This is what the C drop-in for chunkset_manage
 * (chunkset.c:246-507; voxplat_b200/host/vp_chunkset_manage.c) talks to: every call below is the vp_* call of the same
 * name routed to the device that owns each chunk. */
typedef struct vp_multi vp_multi;
VP_API int32_t vp_device_count(void);                             /* visible CUDA devices (0 when there is none) */
VP_API int  vp_multi_create(const vp_config *base, const int32_t *devices, int32_t ndev, vp_multi **out);
VP_API void vp_multi_destroy(vp_multi *m);
VP_API const char *vp_multi_last_error(const vp_multi *m);       /* m may be NULL for creation errors */
VP_API int32_t vp_multi_devices(const vp_multi *m);
VP_API vp_ctx *vp_multi_ctx(vp_multi *m, int32_t i);              /* the i-th slab's context (owned by m) */
VP_API int32_t vp_multi_owner(const vp_multi *m, uint32_t chunk_id);
VP_API int  vp_multi_upload_chunks_dense(vp_multi *m, const uint32_t *chunk_ids, uint32_t n, const uint8_t *host_dense);
VP_API int  vp_multi_upload_chunks_rle(vp_multi *m, const uint32_t *chunk_ids, uint32_t n, const uint32_t *words, const uint64_t *word_offsets);
VP_API int  vp_multi_set_chunks_null(vp_multi *m, const uint32_t *chunk_ids, uint32_t n);
VP_API int  vp_multi_upload_shadow_rows(vp_multi *m, uint32_t z0, uint32_t z1, const uint16_t *rows);
/* Push every slab's border planes into its neighbours' ghost chunks (mesh != 0: also the planes only mesh AO reads). */
VP_API int  vp_multi_exchange_halos(vp_multi *m, int32_t mesh);
VP_API int  vp_multi_batch_prepare(vp_multi *m, const uint32_t *chunk_ids, uint32_t n, const uint8_t *per_chunk_flags, uint32_t flags);
/* One device-resident step on all devices: interior chunks | border planes peer to peer | border chunks.  Asynchronous. */
VP_API int  vp_multi_rebuild_device(vp_multi *m, int32_t mesh);
VP_API int  vp_multi_synchronize(vp_multi *m);
/* vp_rebuild_batch over all devices.  results[i] is relative to splat_bases[owner[i]] / mesh_bases[owner[i]] (arrays of
 * vp_multi_devices() entries, pinned staging of the owning context, valid until the next rebuild call). */
VP_API int  vp_multi_rebuild_batch(vp_multi *m, const uint32_t *chunk_ids, uint32_t n, uint32_t flags, const uint8_t *per_chunk_flags,
                            vp_chunk_result *results, uint8_t *owner, const void **splat_bases, const void **mesh_bases);

#ifdef __cplusplus
}
#endif
#endif /* VOXPLAT_B200_H */
