/*
 * vp_chunkset_manage.c -- host-side drop-in for the reference's rebuild dispatcher.
 *
 * Keeps the engine's entry point `void chunkset_manage(struct ChunkSet*)` (chunkset.c:246, called every
 * 5 ms from mesher_loop, game.c:77-89) and its observable contract -- which chunks are picked, what is
 * published into struct ChunkMD and in which order the flags flip -- but replaces the per-chunk CPU work
 * (chunk_open_ro -> chunk_make_mask / splatlist / downsample or chunk_make_mesh, chunkset.c:318-458) by ONE
 * batched call into the CUDA library (include/voxplat_b200.h).
 *
 * This file is compiled AGAINST THE REFERENCE'S OWN HEADERS (chunkset.h, mem.h, ctx.h: -I<reference>/src);
 * nothing of the reference is copied.  To adopt it, drop the file into src/, remove the body of
 * chunkset_manage from chunkset.c and link libvoxplat_b200.so.
 *
 * Differences that are deliberate:
 *   - the "at most ~9 chunks per pass" throttle (chunkset.c:255) existed to bound CPU time per pass; a batch
 *     costs the GPU microseconds per chunk, so every eligible chunk is rebuilt in the same pass;
 *   - scratch buffers (chunkset.c:233-271) are gone: results arrive in pinned staging and are copied once
 *     into mem_alloc'd blocks, exactly sized (the reference guesses R^3*10 bytes, SURVEY 8a' u7).
 */
#define __FILENAME__ "vp_chunkset_manage.c"
#include "chunkset.h"
#include "mem.h"
#include "event.h"
#include "ctx.h"

#include <string.h>
#include <stdlib.h>

#include "voxplat_b200.h"

static struct {
	struct ChunkSet *set;
	vp_ctx *ctx;
	uint8_t *stale;            /* per chunk: the device copy is older than the host voxels */
	uint32_t *ids;
	uint8_t *flags;
	uint8_t *was_dirty;
	vp_chunk_result *res;
	uint32_t *up_ids;
} G;

static void vp_die(const char *what)
{
	logf_error("voxplat_b200: %s: %s", what, vp_last_error(G.ctx));
	panic();                                   /* the reference's error convention (event.c:158-161) */
}

/* (Re)attach to a ChunkSet: device copy of the world geometry, everything marked stale. */
static void vp_attach(struct ChunkSet *set)
{
	if (G.ctx) { vp_ctx_destroy(G.ctx); free(G.stale); free(G.ids); free(G.flags); free(G.was_dirty); free(G.res); free(G.up_ids); }
	memset(&G, 0, sizeof G);
	vp_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.device = 0;
	cfg.root_bitw = set->root_bitw;
	for (int i = 0; i < 3; i++) cfg.max_bitw[i] = set->max_bitw[i];
	if (vp_ctx_create(&cfg, &G.ctx) != VP_OK) vp_die("vp_ctx_create");
	G.set = set;
	G.stale = malloc(set->count); memset(G.stale, 1, set->count);
	G.ids = malloc(sizeof(uint32_t) * set->count);
	G.up_ids = malloc(sizeof(uint32_t) * set->count);
	G.flags = malloc(set->count);
	G.was_dirty = malloc(set->count);
	G.res = malloc(sizeof(vp_chunk_result) * set->count);
}

/* Bring the device copy of chunk i up to date from whatever form the host holds (state machine of
 * chunkset.h:149-169): null chunk, dense voxels, or RLE only (decoded on the device, no host decode). */
static void vp_push_chunk(struct ChunkSet *set, uint32_t i)
{
	struct ChunkMD *c = &set->chunks[i];
	int rc;
	pthread_mutex_lock(&c->mutex_read);              /* same lock order as chunk_open_ro (chunkset.c:135) */
	if ((!c->voxels && c->rle == set->null_chunk->rle) || c->voxels == set->null_chunk->voxels) {
		rc = vp_set_chunks_null(G.ctx, &i, 1);
	} else if (c->voxels) {
		rc = vp_upload_chunks_dense(G.ctx, &i, 1, c->voxels);
	} else {
		const uint32_t *w = (const uint32_t *)c->rle;
		uint64_t off[2] = { 0, 0 };
		uint32_t n = 0;
		do { n++; } while (w[n]);                    /* rle.c:98-108 termination rule */
		off[1] = n + 1;
		rc = vp_upload_chunks_rle(G.ctx, &i, 1, w, off);
	}
	pthread_mutex_unlock(&c->mutex_read);
	if (rc != VP_OK) vp_die("chunk upload");
	G.stale[i] = 0;
}

void chunkset_manage(struct ChunkSet *set)
{
	if (G.set != set) vp_attach(set);

	/* ---- 1. residency: every chunk whose voxels changed since the last pass goes to the device first, so
	 * that the halos the kernels read are current even for chunks that are throttled below ---- */
	int any_upload = 0;
	for (uint32_t i = 0; i < set->count; i++) {
		if (set->chunks[i].dirty) G.stale[i] = 1;
		if (G.stale[i]) { vp_push_chunk(set, i); any_upload = 1; }
	}
	if (any_upload) {            /* edits move the height map too (shadow_place_update, shadow.h:77-89) */
		uint32_t rows = set->shadow_map_size[1];
		if (vp_upload_shadow_rows(G.ctx, 0, rows, set->shadow_map) != VP_OK) vp_die("vp_upload_shadow_rows");
	}

	/* ---- 2. selection with the reference's predicates (chunkset.c:284-316) ---- */
	uint32_t n = 0;
	for (uint32_t i = 0; i < set->count; i++) {
		struct ChunkMD *c = &set->chunks[i];
		if (c->svl_dirty || c->mesh_dirty) continue;             /* previous geometry not uploaded yet (:284) */
		c->remesh = c->remesh | c->dirty;
		if (c->remesh == 0) {
			if (!c->mesh_dirty && c->mesh_vbo) {                 /* uploaded: drop the CPU copy (:290-295) */
				c->mesh_vbo = mem_free(c->mesh_vbo);
				c->mesh_ibo = mem_free(c->mesh_ibo);
			}
			if (c->voxels && c->last_access + 1.0 < ctx_time()) {  /* idle: compress (:299-305) */
				chunk_lock(set, c);
				chunk_compress(set, c);
				chunk_unlock(set, c);
			}
			continue;
		}
		if (c->last_meshing + 0.100 > ctx_time()) continue;      /* :309 */
		G.was_dirty[n] = c->dirty;
		c->changing = c->dirty;
		c->dirty = 0;
		c->remesh = 0;
		c->last_meshing = ctx_time();
		G.ids[n] = i;
		G.flags[n] = c->make_mesh ? VP_REBUILD_MESH : VP_REBUILD_SPLAT;     /* :337 */
		n++;
	}
	if (!n) return;

	/* ---- 3. one batched rebuild on the GPU ---- */
	const void *splat_base = NULL, *mesh_base = NULL;
	if (vp_rebuild_batch(G.ctx, G.ids, n, 0, G.flags, G.res, &splat_base, &mesh_base) != VP_OK) vp_die("vp_rebuild_batch");

	/* ---- 4. publication, in the reference's order: buffers -> counts -> *_dirty (chunkset.c:347-366, :463-501) ---- */
	for (uint32_t k = 0; k < n; k++) {
		struct ChunkMD *c = &set->chunks[G.ids[k]];
		const vp_chunk_result *r = &G.res[k];
		uint8_t clear_svl = 0;
		if (G.flags[k] & VP_REBUILD_MESH) {
			if (c->mesh_vbo) mem_free(c->mesh_vbo);
			c->mesh_vbo_items = r->vbo_items;
			c->mesh_vbo = mem_alloc(c->mesh_vbo_items * sizeof(uint16_t));
			memcpy(c->mesh_vbo, (const uint8_t *)mesh_base + r->vbo_offset, c->mesh_vbo_items * sizeof(uint16_t));
			if (c->mesh_ibo) mem_free(c->mesh_ibo);
			c->mesh_ibo_items = r->ibo_items;
			c->mesh_ibo = mem_alloc(c->mesh_ibo_items * sizeof(uint32_t));
			memcpy(c->mesh_ibo, (const uint8_t *)mesh_base + r->ibo_offset, c->mesh_ibo_items * sizeof(uint32_t));
			clear_svl = c->no_geometry = !r->vbo_items;
			c->mesh_dirty = 1;
		} else {
			pthread_mutex_lock(&c->mutex_svl);
			c->no_geometry = !r->svl_items_total;
			clear_svl = c->no_geometry;
			if (!clear_svl) {
				for (int l = 0; l < MAX_LOD_LEVEL; l++) c->svl_items[l] = r->svl_items[l];
				if (c->svl) c->svl = mem_free(c->svl);
				c->svl = mem_alloc(r->svl_items_total * sizeof(uint16_t));
				memcpy(c->svl, (const uint8_t *)splat_base + r->svl_offset, r->svl_items_total * sizeof(uint16_t));
				c->svl_items_total = r->svl_items_total;
				c->svl_dirty = 1;
			}
			pthread_mutex_unlock(&c->mutex_svl);
		}
		if (clear_svl) {
			pthread_mutex_lock(&c->mutex_svl);
			if (c->svl) c->svl = mem_free(c->svl);
			memset(c->svl_items, 0, sizeof c->svl_items);
			c->svl_items_total = 0;
			c->svl_dirty = 1;
			pthread_mutex_unlock(&c->mutex_svl);
		}
		c->changing = 0;
	}
}

/* rle.h:7-8 drop-ins on top of the flat device codec (same allocation contract: mem_alloc'd result). */
Voxel *vp_rle_compress_dropin(Voxel *data, uint32_t length)
{
	uint32_t n = 0, cap = length + 1;
	uint32_t *tmp = malloc(sizeof(uint32_t) * cap);
	if (!G.ctx || vp_rle_compress(G.ctx, data, length, tmp, cap, &n) != VP_OK) vp_die("vp_rle_compress");
	Voxel *out = mem_alloc(n * sizeof(uint32_t));
	memcpy(out, tmp, n * sizeof(uint32_t));
	free(tmp);
	return out;
}

Voxel *vp_rle_decompress_dropin(void *vdata, uint32_t expected_bytes)
{
	const uint32_t *w = vdata;
	uint32_t n = 0, nb = 0;
	do { n++; } while (w[n]);
	Voxel *out = mem_alloc(expected_bytes);
	if (!G.ctx || vp_rle_decompress(G.ctx, w, n + 1, out, expected_bytes, &nb) != VP_OK) vp_die("vp_rle_decompress");
	return out;
}
