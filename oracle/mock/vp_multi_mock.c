/* oracle/mock/vp_multi_mock.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A host stand-in for the vp_multi_* entry points of include/voxplat_b200.h, so that the C drop-in dispatcher
 * (voxplat_b200/host/vp_chunkset_manage.c) can be exercised WITHOUT a GPU: its selection predicates, the dirty / pending /
 * stale bookkeeping, the batched residency pass, the publication order -- and, above all, its behaviour while another thread
 * edits the world, which the single-threaded GPU drop-in test cannot show.
 *
 * Like the device, the mock keeps ITS OWN COPY of the world (dense chunks + height map) that only changes through the
 * upload calls, and rebuilds from that copy with the oracle restatement (vox_oracle.c, linked into the same library):
 * a dispatcher that forgets to re-upload an edited chunk publishes stale geometry here exactly as it would on the GPUs.
 * Never linked into the product. */
#include "voxplat_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
	int32_t rb;
	int32_t bits[3];
	const uint8_t *const *chunks;
	const uint16_t *shadow;
} vo_world;
uint32_t vo_chunk_splat(const vo_world *w, uint32_t id, int16_t *out, uint32_t items[5]);
uint32_t vo_chunk_mesh_faces(const vo_world *w, uint32_t id);
void vo_chunk_mesh(const vo_world *w, uint32_t id, int16_t *vbo, uint32_t *ibo, uint32_t *nv, uint32_t *ni);
uint32_t vo_rle_decode(const uint32_t *words, uint8_t *out, uint32_t cap);

struct vp_multi {
	int32_t rb, bits[3];
	uint32_t n_chunks;
	size_t N, shw, rows;
	uint8_t **chunks;
	uint16_t *shadow;
	uint8_t *splat_stage, *mesh_stage;
	size_t splat_cap, mesh_cap;
	uint64_t uploads, rebuilt;
	char err[128];
};

int32_t vp_device_count(void) { return 1; }
const char *vp_multi_last_error(const vp_multi *m) { return m ? m->err : "mock"; }
int32_t vp_multi_devices(const vp_multi *m) { return m ? 1 : 0; }
vp_ctx *vp_multi_ctx(vp_multi *m, int32_t i) { (void)m; (void)i; return NULL; }
int32_t vp_multi_owner(const vp_multi *m, uint32_t id) { return m && id < m->n_chunks ? 0 : -1; }

int vp_multi_create(const vp_config *cfg, const int32_t *devices, int32_t ndev, vp_multi **out)
{
	(void)devices; (void)ndev;
	vp_multi *m = calloc(1, sizeof *m);
	m->rb = cfg->root_bitw;
	for (int i = 0; i < 3; i++) m->bits[i] = cfg->max_bitw[i];
	m->n_chunks = 1u << (m->bits[0] + m->bits[1] + m->bits[2]);
	m->N = (size_t)1 << (3 * m->rb);
	m->shw = (((size_t)1 << m->bits[0]) + ((size_t)1 << m->bits[1])) << m->rb;
	m->rows = (size_t)1 << (m->bits[2] + m->rb);
	m->chunks = calloc(m->n_chunks, sizeof *m->chunks);
	m->shadow = calloc(m->shw * m->rows + 17 * m->shw + 64, sizeof(uint16_t));      /* zero-filled and padded like the device copy */
	*out = m;
	return VP_OK;
}

void vp_multi_destroy(vp_multi *m)
{
	if (!m) return;
	for (uint32_t i = 0; i < m->n_chunks; i++) free(m->chunks[i]);
	free(m->chunks); free(m->shadow); free(m->splat_stage); free(m->mesh_stage); free(m);
}

static int all_zero(const uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) if (p[i]) return 0; return 1; }

static void store(vp_multi *m, uint32_t id, const uint8_t *dense)
{
	m->uploads++;
	if (all_zero(dense, m->N)) { free(m->chunks[id]); m->chunks[id] = NULL; return; }      /* chunkset.c:225-228 */
	if (!m->chunks[id]) m->chunks[id] = malloc(m->N);
	memcpy(m->chunks[id], dense, m->N);
}

int vp_multi_upload_chunks_dense(vp_multi *m, const uint32_t *ids, uint32_t n, const uint8_t *host)
{
	for (uint32_t i = 0; i < n; i++) { if (ids[i] >= m->n_chunks) return VP_ERR_NOT_RESIDENT; store(m, ids[i], host + (size_t)i * m->N); }
	return VP_OK;
}

int vp_multi_set_chunks_null(vp_multi *m, const uint32_t *ids, uint32_t n)
{
	for (uint32_t i = 0; i < n; i++) { if (ids[i] >= m->n_chunks) return VP_ERR_NOT_RESIDENT; free(m->chunks[ids[i]]); m->chunks[ids[i]] = NULL; m->uploads++; }
	return VP_OK;
}

int vp_multi_upload_chunks_rle(vp_multi *m, const uint32_t *ids, uint32_t n, const uint32_t *words, const uint64_t *offs)
{
	uint8_t *tmp = malloc(m->N);
	for (uint32_t i = 0; i < n; i++) {
		if (ids[i] >= m->n_chunks) { free(tmp); return VP_ERR_NOT_RESIDENT; }
		if (vo_rle_decode(words + offs[i], tmp, (uint32_t)m->N) != m->N) { free(tmp); snprintf(m->err, sizeof m->err, "malformed stream"); return VP_ERR_RLE; }
		store(m, ids[i], tmp);
	}
	free(tmp);
	return VP_OK;
}

int vp_multi_upload_shadow_rows(vp_multi *m, uint32_t z0, uint32_t z1, const uint16_t *rows)
{
	if (z1 > m->rows || z0 > z1) return VP_ERR_ARG;
	memcpy(m->shadow + (size_t)z0 * m->shw, rows, (size_t)(z1 - z0) * m->shw * sizeof(uint16_t));
	return VP_OK;
}

static uint8_t *grow(uint8_t **buf, size_t *cap, size_t need)
{
	if (need > *cap) { *cap = need * 2 + 4096; *buf = realloc(*buf, *cap); }
	return *buf;
}

int vp_multi_rebuild_batch(vp_multi *m, const uint32_t *ids, uint32_t n, uint32_t flags, const uint8_t *per_chunk_flags,
                           vp_chunk_result *results, uint8_t *owner, const void **splat_bases, const void **mesh_bases)
{
	vo_world w = { m->rb, { m->bits[0], m->bits[1], m->bits[2] }, (const uint8_t *const *)m->chunks, m->shadow };
	const size_t R1 = ((size_t)1 << m->rb) + 1;
	int16_t *geom = malloc(R1 * R1 * R1 * 5 * sizeof(int16_t));
	size_t so = 0, mo = 0;
	for (uint32_t k = 0; k < n; k++) {
		const uint32_t f = per_chunk_flags ? per_chunk_flags[k] : flags;
		vp_chunk_result *r = &results[k];
		memset(r, 0, sizeof *r);
		if (owner) owner[k] = 0;
		if (ids[k] >= m->n_chunks) { free(geom); return VP_ERR_NOT_RESIDENT; }
		m->rebuilt++;
		if (f & VP_REBUILD_SPLAT) {
			const uint32_t items = vo_chunk_splat(&w, ids[k], geom, r->svl_items);
			r->svl_items_total = items;
			r->svl_offset = so;
			memcpy(grow(&m->splat_stage, &m->splat_cap, so + (size_t)items * 2) + so, geom, (size_t)items * 2);
			so += (size_t)items * 2;
		}
		if (f & VP_REBUILD_MESH) {
			const uint32_t faces = vo_chunk_mesh_faces(&w, ids[k]);
			int16_t *vbo = malloc(((size_t)faces + 1) * 16 * sizeof(int16_t));
			uint32_t *ibo = malloc(((size_t)faces + 1) * 6 * sizeof(uint32_t));
			vo_chunk_mesh(&w, ids[k], vbo, ibo, &r->vbo_items, &r->ibo_items);
			grow(&m->mesh_stage, &m->mesh_cap, mo + (size_t)r->vbo_items * 2 + (size_t)r->ibo_items * 4);
			r->vbo_offset = mo; memcpy(m->mesh_stage + mo, vbo, (size_t)r->vbo_items * 2); mo += (size_t)r->vbo_items * 2;
			r->ibo_offset = mo; memcpy(m->mesh_stage + mo, ibo, (size_t)r->ibo_items * 4); mo += (size_t)r->ibo_items * 4;
			free(vbo); free(ibo);
		}
	}
	free(geom);
	if (splat_bases) splat_bases[0] = m->splat_stage;
	if (mesh_bases) mesh_bases[0] = m->mesh_stage;
	return VP_OK;
}

/* the flat codec is not part of what this mock is for */
int vp_rle_compress(vp_ctx *c, const uint8_t *d, uint32_t l, uint32_t *o, uint32_t cap, uint32_t *n) { (void)c; (void)d; (void)l; (void)o; (void)cap; (void)n; return VP_ERR_NO_DEVICE; }
int vp_rle_decompress(vp_ctx *c, const uint32_t *w, uint32_t nw, uint8_t *o, uint32_t cap, uint32_t *n) { (void)c; (void)w; (void)nw; (void)o; (void)cap; (void)n; return VP_ERR_NO_DEVICE; }
const char *vp_last_error(const vp_ctx *c) { (void)c; return "mock"; }
