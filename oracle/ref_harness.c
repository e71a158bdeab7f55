/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Thin C driver around the UNMODIFIED reference sources, which are compiled where they lie under
 * /root/reference/src by oracle/Makefile into oracle/_ref/libvoxref.so (git-ignored).  Nothing in
 * here re-implements the reference: every vr_* function only sequences reference calls the way the
 * reference's own dispatcher does (chunkset.c:318-458 for the splat path, :339-343 for the mesh
 * path), with the oracle decisions of SURVEY.md section 8(a') applied:
 *   u2/u3  shadow map is calloc'ed with 17*SH_MAX_X+64 trailing zero entries (LOD-l splats sample at
 *          +(1<<l) on every axis, mesher.c:526-531: up to 16 rows past the end for tiny worlds),
 *   u5     one rle_compress happens before any decompress (chunkset_clear does it),
 *   u6     no chunk_compress runs concurrently with meshing,
 *   u7     scratch sized as the reference does (geom N*10, work M*10, mask M) unless told larger.
 *
 * Used by tests/ (parity checker) and by bench.py's cpu_baseline / --impl reference legs.
 */
#define __FILENAME__ "oracle/ref_harness.c"
#include "chunkset.h"
#include "chunkset/mesher.h"
#include "chunkset/rle.h"
#include "chunkset/edit.h"
#include "mem.h"
#include "event.h"
#include "ctx.h"

#include <string.h>
#include <stdio.h>
#include <omp.h>

#define VR_EXPORT __attribute__((visibility("default")))

static int vr_inited = 0;

VR_EXPORT int vr_init(uint64_t heap_bytes)
{
	if (vr_inited) return 0;
	mem_init((size_t)heap_bytes);      /* main.c:85 */
	log_init();                        /* main.c:87 */
	vr_inited = 1;
	return 0;
}

VR_EXPORT uint32_t vr_shadow_pad(struct ChunkSet *set) { return 17 * set->shadow_map_size[0] + 64; }

/* chunkset_manage keeps per-thread scratch in a global that it sizes ONCE, for the first ChunkSet it sees
 * (chunkset.c:233-271): a process that creates worlds of several chunk sizes (the test suite) has to drop it in between,
 * or the second world overflows the first one's buffers.  Same layout as the anonymous struct at chunkset.c:233-242. */
extern struct { void *geom[2]; void *mask[2]; void *work[2]; uint32_t work_size, geom_size, mask_size; uint8_t lock[2]; } mesher[4];
static void vr_reset_dispatcher_scratch(void)
{
	for (int t = 0; t < 4; t++)
		for (int b = 0; b < 2; b++) {
			if (mesher[t].geom[b]) { mem_free(mesher[t].geom[b]); mem_free(mesher[t].work[b]); mem_free(mesher[t].mask[b]); }
			mesher[t].geom[b] = mesher[t].work[b] = mesher[t].mask[b] = NULL;
			mesher[t].lock[b] = 0;
		}
}

VR_EXPORT struct ChunkSet *vr_world_create(int root_bitw, int bx, int by, int bz)
{
	vr_reset_dispatcher_scratch();
	uint8_t mb[3] = { (uint8_t)bx, (uint8_t)by, (uint8_t)bz };
	struct ChunkSet *set = chunkset_create((uint8_t)root_bitw, mb);   /* game.c:263 */
	chunkset_clear(set);                                              /* game.c:265 */
	shadow_init(set);                                                 /* game.c:267 */
	/* u2/u3: replace the uninitialised, unpadded map by a zeroed, padded one */
	mem_free(set->shadow_map);
	set->shadow_map = calloc((size_t)set->shadow_map_length + vr_shadow_pad(set), sizeof(uint16_t));
	return set;
}

VR_EXPORT uint32_t  vr_chunk_count(struct ChunkSet *set) { return set->count; }
VR_EXPORT uint16_t *vr_shadow_ptr(struct ChunkSet *set)  { return set->shadow_map; }
VR_EXPORT uint32_t  vr_shadow_len(struct ChunkSet *set)  { return set->shadow_map_length; }

VR_EXPORT void vr_world_set_shadow(struct ChunkSet *set, const uint16_t *map, uint32_t n)
{
	if (n > set->shadow_map_length + vr_shadow_pad(set)) n = set->shadow_map_length + vr_shadow_pad(set);
	memcpy(set->shadow_map, map, (size_t)n * sizeof(uint16_t));
}

/* Store dense voxels into a chunk through the reference's own rw-open path (chunkset.c:167-204). */
VR_EXPORT void vr_world_set_chunk(struct ChunkSet *set, uint32_t id, const uint8_t *dense)
{
	struct ChunkMD *c = &set->chunks[id];
	chunk_open_rw(set, c);
	memcpy(c->voxels, dense, c->count);
	c->dirty = 1;
	chunk_close_rw(set, c);
}

VR_EXPORT void vr_world_compress_all(struct ChunkSet *set)
{
	for (uint32_t i = 0; i < set->count; i++) {
		struct ChunkMD *c = &set->chunks[i];
		chunk_lock(set, c);
		chunk_compress(set, c);            /* chunkset.c:213-230, incl. the first-word all-air test */
		chunk_unlock(set, c);
	}
}

VR_EXPORT void vr_world_decode_all(struct ChunkSet *set)
{
	for (uint32_t i = 0; i < set->count; i++) {
		chunk_open_ro(set, &set->chunks[i]);
		chunk_close_ro(set, &set->chunks[i]);
	}
}

/* state: bit0 voxels!=NULL, bit1 rle!=NULL, bit2 rle is the shared null rle, bit3 voxels alias null */
VR_EXPORT int vr_chunk_state(struct ChunkSet *set, uint32_t id)
{
	struct ChunkMD *c = &set->chunks[id];
	return (c->voxels != NULL) | ((c->rle != NULL) << 1) | ((c->rle == set->null_chunk->rle) << 2)
	     | ((c->voxels == set->null_chunk->voxels) << 3);
}

/* RLE words of a chunk (incl. terminator); returns word count, 0 if the chunk holds no rle. */
VR_EXPORT uint32_t vr_chunk_rle(struct ChunkSet *set, uint32_t id, const uint32_t **words)
{
	struct ChunkMD *c = &set->chunks[id];
	if (!c->rle) { *words = NULL; return 0; }
	const uint32_t *w = (const uint32_t *)c->rle;
	uint32_t n = 0;
	do { n++; } while (w[n]);          /* same termination rule as rle.c:100-108 */
	*words = w;
	return n + 1;
}

VR_EXPORT const uint8_t *vr_chunk_voxels(struct ChunkSet *set, uint32_t id)
{
	struct ChunkMD *c = &set->chunks[id];
	chunk_open_ro(set, c);
	chunk_close_ro(set, c);
	return c->voxels;
}

VR_EXPORT void vr_shadow_place(struct ChunkSet *set, uint32_t x, uint32_t y, uint32_t z)
{
	uint32_t ws[3] = { x, y, z };
	shadow_place_update(set, ws);
}

VR_EXPORT int vr_edit_read(struct ChunkSet *set, uint32_t x, uint32_t y, uint32_t z)
{
	uint32_t ws[3] = { x, y, z };
	return chunkset_edit_read(set, ws);
}

VR_EXPORT void vr_edit_write(struct ChunkSet *set, uint32_t x, uint32_t y, uint32_t z, int v)
{
	uint32_t ws[3] = { x, y, z };
	chunkset_edit_write(set, ws, (Voxel)v);
}

VR_EXPORT void vr_edit_sphere(struct ChunkSet *set, int32_t x, int32_t y, int32_t z, uint32_t radius, int v)
{
	int32_t ws[3] = { x, y, z };
	chunkset_edit_sphere(set, ws, radius, (Voxel)v);
}

/* ---- the consumer right after the path: gfx_update_svl (gfx/vsplat.c:197-338), run unmodified against the GL capture
 * shim (gfx_shim/gl_capture.c).  For a chunk whose splat list was published it rebuilds, for every LOD level, the node
 * buffer that chunk belongs to. ---- */
void gfx_update_svl(struct ChunkSet *set, uint32_t index);
const void *vr_gl_buffer(unsigned int id, size_t *size);
VR_EXPORT void vr_gfx_update_svl(struct ChunkSet *set, uint32_t id) { gfx_update_svl(set, id); }
VR_EXPORT int vr_chunk_svl_dirty(struct ChunkSet *set, uint32_t id) { return set->chunks[id].svl_dirty; }
/* The node buffer of (lod, node index) as the reference left it in "GPU memory": returns GeometrySVL.vbo_items. */
VR_EXPORT uint32_t vr_gsvl_node(struct ChunkSet *set, int lod, uint32_t node, const void **data, uint64_t *bytes)
{
	struct GeometrySVL *g = &set->gsvl[lod][node];
	size_t size = 0;
	*data = vr_gl_buffer(g->vbo, &size);
	*bytes = size;
	return g->vbo ? g->vbo_items : 0;
}

/* The pick ray of game.c:212, as is. */
VR_EXPORT int vr_raycast(struct ChunkSet *set, float *origin, float *vector, uint32_t *coord, int8_t *normal)
{
	return (int)chunkset_edit_raycast_until_solid(set, origin, vector, coord, normal);
}

VR_EXPORT int vr_chunk_dirty(struct ChunkSet *set, uint32_t id, int clear)
{
	int d = set->chunks[id].dirty;
	if (clear) set->chunks[id].dirty = 0;
	return d;
}

/* ---- per-thread scratch, sized as chunkset.c:263-265 (times `scale` for adversarial inputs) ---- */
struct vr_scratch { uint8_t *geom, *work, *mask; size_t geom_size, work_size, mask_size; };
static __thread struct vr_scratch tls;

static int vr_scale = 1;
/* scale=1: the reference's own scratch sizes.  Random (non-terrain) test data can need up to 96 B of
 * mesh per voxel (u7), so tests raise the scale to 10 instead of letting the reference overflow. */
VR_EXPORT void vr_set_scratch_scale(int scale) { vr_scale = scale < 1 ? 1 : scale; }

static struct vr_scratch *scratch_get(struct ChunkSet *set)
{
	size_t R = set->root, N = R * R * R, M = (R + 1) * (R + 1) * (R + 1);
	size_t g = N * 10 * vr_scale, w = M * 10 * vr_scale;      /* chunkset.c:263-264 */
	if (tls.geom_size != g || tls.mask_size != M) {
		free(tls.geom); free(tls.work); free(tls.mask);
		tls.geom = malloc(g); tls.work = malloc(w); tls.mask = malloc(M);
		tls.geom_size = g; tls.work_size = w; tls.mask_size = M;
	}
	return &tls;
}

/* Splat path of one chunk = chunkset.c:318-458 verbatim in call order; the scratch is cleared only
 * where the reference clears it (only the M bytes the functions can touch; the rest is never read). */
VR_EXPORT int vr_chunk_splat(struct ChunkSet *set, uint32_t id, int16_t *out, uint32_t cap_items, uint32_t items[5])
{
	struct ChunkMD *c = &set->chunks[id];
	struct vr_scratch *s = scratch_get(set);
	size_t M = s->mask_size;
	chunk_open_ro(set, c);
	memset(s->work, 0, M);
	memset(s->mask, 0, M);
	int16_t *geom = (int16_t *)s->geom;

	uint32_t n = 0, m;
	chunk_make_mask(set, c, s->mask);
	chunk_make_splatlist(set, c, 0, s->mask, geom, &n);
	items[0] = n;

	chunk_mask_downsample(set, 1, s->mask, s->work);
	m = 0; chunk_make_splatlist(set, c, 1, s->work, geom + n, &m); items[1] = m; n += m;

	memset(s->mask, 0, M);
	chunk_mask_downsample(set, 2, s->work, s->mask);
	m = 0; chunk_make_splatlist(set, c, 2, s->mask, geom + n, &m); items[2] = m; n += m;

	memset(s->work, 0, M);
	chunk_mask_downsample(set, 3, s->mask, s->work);
	m = 0; chunk_make_splatlist(set, c, 3, s->work, geom + n, &m); items[3] = m; n += m;

	memset(s->mask, 0, M);
	chunk_mask_downsample(set, 4, s->work, s->mask);
	m = 0; chunk_make_splatlist(set, c, 4, s->mask, geom + n, &m); items[4] = m; n += m;

	chunk_close_ro(set, c);
	if (out && n <= cap_items) memcpy(out, geom, (size_t)n * sizeof(int16_t));
	return (int)n;
}

/* Level-0 mask only (debug aid for differential tests). */
VR_EXPORT void vr_chunk_mask(struct ChunkSet *set, uint32_t id, uint8_t *mask_out)
{
	struct ChunkMD *c = &set->chunks[id];
	size_t R = set->root, M = (R + 1) * (R + 1) * (R + 1);
	chunk_open_ro(set, c);
	memset(mask_out, 0, M);
	chunk_make_mask(set, c, mask_out);
	chunk_close_ro(set, c);
}

/* Mesh path of one chunk = chunkset.c:339-343. */
VR_EXPORT int vr_chunk_mesh(struct ChunkSet *set, uint32_t id, int16_t *vbo, uint32_t cap_v,
                            uint32_t *ibo, uint32_t cap_i, uint32_t *nv_out, uint32_t *ni_out)
{
	struct ChunkMD *c = &set->chunks[id];
	struct vr_scratch *s = scratch_get(set);
	uint32_t nv = 0, ni = 0;
	chunk_open_ro(set, c);
	chunk_make_mesh(set, c, (int16_t *)s->geom, &nv, (uint32_t *)s->work, &ni);
	chunk_close_ro(set, c);
	*nv_out = nv; *ni_out = ni;
	if (vbo && nv <= cap_v) memcpy(vbo, s->geom, (size_t)nv * sizeof(int16_t));
	if (ibo && ni <= cap_i) memcpy(ibo, s->work, (size_t)ni * sizeof(uint32_t));
	return 0;
}

VR_EXPORT uint32_t vr_rle_compress(const uint8_t *dense, uint32_t n, uint32_t *out, uint32_t cap_words)
{
	uint32_t *w = (uint32_t *)rle_compress((Voxel *)dense, n);
	uint32_t k = 0;
	do { k++; } while (w[k]);
	k++;
	if (out && k <= cap_words) memcpy(out, w, (size_t)k * 4);
	mem_free(w);
	return k;
}

/* NB: rle_decompress sizes its scratch from the global set by the last rle_compress (u5). */
VR_EXPORT uint32_t vr_rle_decompress(const uint32_t *words, uint8_t *out, uint32_t n_expected)
{
	Voxel *v = rle_decompress((void *)words);
	memcpy(out, v, n_expected);
	mem_free(v);
	return n_expected;
}

static uint64_t fnv1a(const void *p, size_t n, uint64_t h)
{
	const uint8_t *b = p;
	for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}
VR_EXPORT uint64_t vr_fnv1a(const void *p, uint64_t n) { return fnv1a(p, (size_t)n, 1469598103934665603ull); }

/*
 * Whole-world rebuild on the host cores: OpenMP dynamic schedule over `ids` (NULL = all chunks),
 * the throttle-free best case of chunkset_manage's loop.
 *   mode 0 = splat path, mode 1 = mesh path.
 * Outputs per chunk: hash (FNV-1a 64 of the output bytes; for mesh VBO then IBO), counts[8]
 * (splat: items[0..4]; mesh: [5]=vbo items, [6]=ibo items).  Returns seconds of wall time.  With hashes == NULL
 * nothing is hashed (timing runs: only the reference's own work is inside the clock).
 */
VR_EXPORT double vr_world_rebuild(struct ChunkSet *set, const uint32_t *ids, uint32_t n_ids, int mode,
                                  int nthreads, uint64_t *hashes, uint32_t *counts)
{
	if (!ids) n_ids = set->count;
	if (nthreads <= 0) nthreads = omp_get_max_threads();
	double t0 = ctx_time();
	#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
	for (uint32_t k = 0; k < n_ids; k++) {
		uint32_t id = ids ? ids[k] : k;
		struct vr_scratch *s = scratch_get(set);
		uint32_t it[8] = {0};
		uint64_t h = 0;                    /* hashes == NULL: a timing run, the outputs are not read back */
		if (mode == 0) {
			int n = vr_chunk_splat(set, id, NULL, 0, it);
			if (hashes) h = fnv1a(s->geom, (size_t)n * 2, 1469598103934665603ull);
		} else {
			vr_chunk_mesh(set, id, NULL, 0, NULL, 0, &it[5], &it[6]);
			if (hashes) {
				h = fnv1a(s->geom, (size_t)it[5] * 2, 1469598103934665603ull);
				h = fnv1a(s->work, (size_t)it[6] * 4, h);
			}
		}
		if (hashes) hashes[k] = h;
		if (counts) memcpy(counts + (size_t)k * 8, it, sizeof(it));
	}
	return ctx_time() - t0;
}

/* ---- the reference's own dispatcher, for drop-in comparisons ---- */
VR_EXPORT void vr_manage(struct ChunkSet *set) { chunkset_manage(set); }
#ifdef VR_WITH_GPU
/* libvoxref_gpu.so only: chunkset_manage is the CUDA drop-in, the reference's own loop is chunkset_manage_cpu */
void chunkset_manage_cpu(struct ChunkSet *set);
VR_EXPORT void vr_manage_cpu(struct ChunkSet *set) { chunkset_manage_cpu(set); }
#endif
VR_EXPORT int vr_chunk_pending(struct ChunkSet *set, uint32_t id)
{
	struct ChunkMD *c = &set->chunks[id];
	return c->dirty | c->remesh | c->changing;
}

VR_EXPORT void vr_chunk_set_make_mesh(struct ChunkSet *set, uint32_t id, int v)
{
	struct ChunkMD *c = &set->chunks[id];
	if (c->make_mesh != v) { c->make_mesh = v; c->remesh = 1; }      /* game.c:621-624 */
}

/* Read back what chunkset_manage published (chunkset.c:347-366, :463-501) and acknowledge it the way
 * gfx_update_svl / gfx_update_mesh do (svl_dirty=0 vsplat.c:334, mesh_dirty=0 vmesh.c:232). */
VR_EXPORT int vr_chunk_published(struct ChunkSet *set, uint32_t id, const uint16_t **svl, uint32_t items[5],
                                 uint32_t *total, const void **vbo, uint32_t *nv, const void **ibo, uint32_t *ni,
                                 int ack)
{
	struct ChunkMD *c = &set->chunks[id];
	int flags = c->svl_dirty | (c->mesh_dirty << 1) | (c->no_geometry << 2) | (c->dirty << 3) | (c->remesh << 4);
	*svl = c->svl; memcpy(items, c->svl_items, 5 * sizeof(uint32_t)); *total = c->svl_items_total;
	*vbo = c->mesh_vbo; *nv = c->mesh_vbo_items; *ibo = c->mesh_ibo; *ni = c->mesh_ibo_items;
	if (ack) { c->svl_dirty = 0; c->mesh_dirty = 0; }
	return flags;
}
