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
 *   - uploads are batched (one call per kind of chunk per pass) and only the height-map rows of edited chunk rows travel;
 *   - the world is sharded by chunk rows (z-slabs) over every visible GPU through the vp_multi layer (one host thread,
 *     border planes pushed peer to peer): VP_DEVICES=0,1,2,3 picks the devices, the default is all of them (the largest
 *     power of two that leaves every slab at least two chunk rows); the reference's dispatcher is 4 OpenMP threads on one
 *     host (chunkset.c:251);
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
	vp_multi *m;               /* one vp_ctx per device, slabs of chunk rows */
	int ndev;
	const void **splat_bases, **mesh_bases;
	uint8_t *owner;            /* per entry of the batch: index of the device that rebuilt it */
	uint8_t *stale;            /* per chunk: the device copy is older than the host voxels */
	uint8_t *pending;          /* per chunk: its dirty flag was consumed (cleared) by a residency pass and not rebuilt yet */
	uint32_t *ids;
	uint8_t *flags;
	uint8_t *was_dirty;
	vp_chunk_result *res;
	/* upload staging of one residency pass */
	uint32_t *null_ids, *dense_ids, *rle_ids;
	uint8_t *dense_stage;      /* VP_PUSH_BATCH chunks of R^3 bytes */
	uint32_t *rle_words; uint64_t rle_cap; uint64_t *rle_offs;
	int first_pass;
} G;

#define VP_PUSH_BATCH 256u     /* dense chunks per upload call (64 MB of staging at 64^3) */

static void vp_die(const char *what)
{
	logf_error("voxplat_b200: %s: %s", what, vp_multi_last_error(G.m));
	panic();                                   /* the reference's error convention (event.c:158-161) */
}

/* (Re)attach to a ChunkSet: device copy of the world geometry, everything marked stale. */
static void vp_attach(struct ChunkSet *set)
{
	if (G.m) {
		vp_multi_destroy(G.m);
		free(G.splat_bases); free(G.mesh_bases); free(G.owner);
		free(G.stale); free(G.pending); free(G.ids); free(G.flags); free(G.was_dirty); free(G.res);
		free(G.null_ids); free(G.dense_ids); free(G.rle_ids); free(G.dense_stage); free(G.rle_words); free(G.rle_offs);
	}
	memset(&G, 0, sizeof G);
	vp_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.root_bitw = set->root_bitw;
	for (int i = 0; i < 3; i++) cfg.max_bitw[i] = set->max_bitw[i];
	/* devices: VP_DEVICES=a,b,... (a device may repeat: several slabs on one GPU), else every visible GPU */
	int32_t devs[64];
	int ndev = 0;
	const int nz = 1 << set->max_bitw[2];
	const char *e = getenv("VP_DEVICES");
	if (e && *e) {
		while (*e && ndev < 64) {
			devs[ndev++] = (int32_t)strtol(e, (char **)&e, 10);
			while (*e == ',' || *e == ' ') e++;
		}
	} else {
		const int vis = vp_device_count();
		ndev = 1;
		while (ndev * 2 <= vis && ndev * 4 <= nz) ndev *= 2;
		for (int i = 0; i < ndev; i++) devs[i] = i;
	}
	if (ndev < 1 || nz % ndev) { logf_error("voxplat_b200: %i devices do not divide %i chunk rows", ndev, nz); panic(); }
	if (vp_multi_create(&cfg, devs, ndev, &G.m) != VP_OK) vp_die("vp_multi_create");
	G.ndev = ndev;
	G.splat_bases = calloc(ndev, sizeof(void *));
	G.mesh_bases = calloc(ndev, sizeof(void *));
	G.owner = malloc(set->count);
	const size_t N = (size_t)1 << (3 * set->root_bitw);
	G.set = set;
	G.stale = malloc(set->count); memset(G.stale, 1, set->count);
	G.pending = calloc(set->count, 1);
	G.ids = malloc(sizeof(uint32_t) * set->count);
	G.flags = malloc(set->count);
	G.was_dirty = malloc(set->count);
	G.res = malloc(sizeof(vp_chunk_result) * set->count);
	G.null_ids = malloc(sizeof(uint32_t) * set->count);
	G.dense_ids = malloc(sizeof(uint32_t) * VP_PUSH_BATCH);
	G.rle_ids = malloc(sizeof(uint32_t) * set->count);
	G.rle_offs = malloc(sizeof(uint64_t) * ((size_t)set->count + 1));
	G.dense_stage = malloc(N * VP_PUSH_BATCH);
	G.first_pass = 1;
}

static void vp_flush_dense(uint32_t *n_dense)
{
	if (*n_dense && vp_multi_upload_chunks_dense(G.m, G.dense_ids, *n_dense, G.dense_stage) != VP_OK) vp_die("vp_upload_chunks_dense");
	*n_dense = 0;
}

/* Residency pass: bring the device copy of every chunk whose voxels changed up to date, in three batched calls (null
 * chunks, dense chunks through a staging buffer, RLE-only chunks decoded on the device) instead of one call per chunk.
 *
 * Ordering against the engine's edit threads (chunk_close_rw sets c->dirty = 1 after writing): a chunk's dirty flag is
 * consumed -- cleared, with c->changing raised in its place so that `dirty || changing` (game.c:643) stays true --
 * BEFORE its voxels are read under the read lock.  An edit that lands later sets dirty again and is picked up by the
 * next pass; an edit between the clear and the read is simply included twice.  (Reading first and clearing later, as
 * round 1 did, could lose an edit for good.) */
static void vp_residency_pass(struct ChunkSet *set)
{
	const size_t N = (size_t)1 << (3 * set->root_bitw);
	const uint32_t R = 1u << set->root_bitw;
	const int zshift = set->max_bitw[0] + set->max_bitw[1];
	uint32_t n_null = 0, n_dense = 0, n_rle = 0, zmin = 0xFFFFFFFFu, zmax = 0;
	uint64_t n_words = 0;
	for (uint32_t i = 0; i < set->count; i++) {
		struct ChunkMD *c = &set->chunks[i];
		if (c->dirty) {
			c->changing = 1;
			c->dirty = 0;
			G.pending[i] = 1;
			G.stale[i] = 1;
			const uint32_t cz = i >> zshift;
			if (cz < zmin) zmin = cz;
			if (cz > zmax) zmax = cz;
		}
		if (!G.stale[i]) continue;
		pthread_mutex_lock(&c->mutex_read);              /* same lock order as chunk_open_ro (chunkset.c:135) */
		if ((!c->voxels && c->rle == set->null_chunk->rle) || c->voxels == set->null_chunk->voxels) {
			G.null_ids[n_null++] = i;
		} else if (c->voxels) {
			memcpy(G.dense_stage + (size_t)n_dense * N, c->voxels, N);
			G.dense_ids[n_dense++] = i;
		} else {
			const uint32_t *w = (const uint32_t *)c->rle;
			uint32_t nw = 0;
			do { nw++; } while (w[nw]);                  /* rle.c:98-108 termination rule */
			nw++;                                        /* + the terminator */
			if (n_words + nw > G.rle_cap) {
				G.rle_cap = (n_words + nw) * 2 + 1024;
				G.rle_words = realloc(G.rle_words, G.rle_cap * sizeof(uint32_t));
			}
			memcpy(G.rle_words + n_words, w, (size_t)nw * sizeof(uint32_t));
			G.rle_offs[n_rle] = n_words;
			G.rle_ids[n_rle++] = i;
			n_words += nw;
		}
		pthread_mutex_unlock(&c->mutex_read);
		G.stale[i] = 0;
		if (n_dense == VP_PUSH_BATCH) vp_flush_dense(&n_dense);
	}
	vp_flush_dense(&n_dense);
	if (n_null && vp_multi_set_chunks_null(G.m, G.null_ids, n_null) != VP_OK) vp_die("vp_set_chunks_null");
	if (n_rle) {
		G.rle_offs[n_rle] = n_words;
		if (vp_multi_upload_chunks_rle(G.m, G.rle_ids, n_rle, G.rle_words, G.rle_offs) != VP_OK) vp_die("vp_upload_chunks_rle");
	}
	/* edits move the height map too (shadow_place_update, shadow.h:77-89), but only rows of the chunk rows they touched */
	const uint32_t shw = set->shadow_map_size[0], rows = set->shadow_map_size[1];
	if (G.first_pass) {
		if (vp_multi_upload_shadow_rows(G.m, 0, rows, set->shadow_map) != VP_OK) vp_die("vp_upload_shadow_rows");
		G.first_pass = 0;
	} else if (zmin <= zmax) {
		const uint32_t z0 = zmin * R, z1 = (zmax + 1) * R < rows ? (zmax + 1) * R : rows;
		if (vp_multi_upload_shadow_rows(G.m, z0, z1, set->shadow_map + (size_t)z0 * shw) != VP_OK) vp_die("vp_upload_shadow_rows");
	}
}

/* Test seam (NULL in production): called between the residency pass and the selection pass -- the window in which an edit
 * could be lost before the dirty flag was consumed ahead of the voxel read (tests/test_dropin_mock.py edits a chunk here). */
void (*vp_chunkset_manage_between_passes)(struct ChunkSet *set) = NULL;

void chunkset_manage(struct ChunkSet *set)
{
	if (G.set != set) vp_attach(set);

	/* ---- 1. residency: every chunk whose voxels changed since the last pass goes to the device first, so
	 * that the halos the kernels read are current even for chunks that are throttled below ---- */
	vp_residency_pass(set);
	if (vp_chunkset_manage_between_passes) vp_chunkset_manage_between_passes(set);

	/* ---- 2. selection with the reference's predicates (chunkset.c:284-316); the dirty flag of the reference is the
	 * `pending` flag here (consumed in step 1) ---- */
	uint32_t n = 0;
	for (uint32_t i = 0; i < set->count; i++) {
		struct ChunkMD *c = &set->chunks[i];
		if (c->svl_dirty || c->mesh_dirty) continue;             /* previous geometry not uploaded yet (:284) */
		c->remesh = c->remesh | G.pending[i];
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
		G.was_dirty[n] = G.pending[i];
		c->changing = G.pending[i];                              /* :311 (already 1 when the chunk was dirty) */
		G.pending[i] = 0;
		c->remesh = 0;
		c->last_meshing = ctx_time();
		G.ids[n] = i;
		G.flags[n] = c->make_mesh ? VP_REBUILD_MESH : VP_REBUILD_SPLAT;     /* :337 */
		n++;
	}
	if (!n) return;

	/* ---- 3. one batched rebuild on the GPU ---- */
	if (vp_multi_rebuild_batch(G.m, G.ids, n, 0, G.flags, G.res, G.owner, G.splat_bases, G.mesh_bases) != VP_OK) vp_die("vp_multi_rebuild_batch");

	/* ---- 4. publication, in the reference's order: buffers -> counts -> *_dirty (chunkset.c:347-366, :463-501) ---- */
	for (uint32_t k = 0; k < n; k++) {
		struct ChunkMD *c = &set->chunks[G.ids[k]];
		const vp_chunk_result *r = &G.res[k];
		const void *splat_base = G.splat_bases[G.owner[k]], *mesh_base = G.mesh_bases[G.owner[k]];
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
		/* the reference drops `changing` here (:503); a chunk edited again meanwhile keeps it through its pending flag */
		c->changing = G.pending[G.ids[k]];
	}
}

/* rle.h:7-8 drop-ins on top of the flat device codec (same allocation contract: mem_alloc'd result). */
Voxel *vp_rle_compress_dropin(Voxel *data, uint32_t length)
{
	uint32_t n = 0, cap = length + 1;
	uint32_t *tmp = malloc(sizeof(uint32_t) * cap);
	if (!G.m || vp_rle_compress(vp_multi_ctx(G.m, 0), data, length, tmp, cap, &n) != VP_OK) vp_die("vp_rle_compress");
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
	if (!G.m || vp_rle_decompress(vp_multi_ctx(G.m, 0), w, n + 1, out, expected_bytes, &nb) != VP_OK) vp_die("vp_rle_decompress");
	return out;
}
