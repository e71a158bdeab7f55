// vp_multi.cu -- several GPUs behind one C handle, one host thread, no NCCL: the world is cut into z-slabs of chunk rows
// (one vp_ctx per device), and the only data that crosses devices -- the 1-voxel border planes of SURVEY 8(e) -- is
// written by ONE kernel per plane straight into the neighbour's ghost chunks over NVLink (peer-mapped pools), together
// with the x-face rows the ghost chunks need.  No pack buffer, no transport call, no unpack pass.
//
// New relative to the reference, which is single-process / single-device (chunkset.c:246-253, game.c:77-89): this is
// the layer the C drop-in dispatcher (host/vp_chunkset_manage.c) talks to, so the engine's own entry point can use
// every GPU of the box.  The torchrun / NCCL form of the same exchange (one process per GPU) is voxplat_b200/slab.py.
//
// Ordering (all on streams, the host never waits inside a step):
//   part 0 of every device (chunks that read no ghost row, context stream)
//   pushes: device i's border stream waits for (a) its own context stream (voxels ready) and (b) the neighbour's
//           "previous step finished" event (nobody still reads the ghost slice), then runs k_halo_push, then records
//           pushed[i][dir];
//   part 1 of every device: its border stream waits for the pushed events of both neighbours, rebuilds the border
//           chunks beside the interior ones and joins the context stream (vp_rebuild_device_part).
#include "vp_internal.h"
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace {

// One CTA per chunk column of a chunk row: z-slice `zslice` of the source chunk -> the same slice of the neighbour's ghost
// chunk (zeros when the source is the null chunk), plus row `zslice` of the ghost chunk's two x-face planes.
__global__ void __launch_bounds__(256)
k_halo_push(int rb, const uint8_t *__restrict__ vox_src, const int32_t *__restrict__ src_row_slots, uint8_t *__restrict__ vox_dst,
            uint8_t *__restrict__ xlo_dst, uint8_t *__restrict__ xhi_dst, const int32_t *__restrict__ dst_row_slots, int zslice)
{
	const int R = 1 << rb, RR = R * R;
	const int ss = src_row_slots[blockIdx.x], ds = dst_row_slots[blockIdx.x];
	if (ds < 0) return;                                  // cannot happen: ghost rows own permanent slots
	const uint8_t *src = ss >= 0 ? vox_src + ((size_t)ss << (3 * rb)) + (size_t)zslice * RR : nullptr;
	uint4 *dst = reinterpret_cast<uint4 *>(vox_dst + ((size_t)ds << (3 * rb)) + (size_t)zslice * RR);
	const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
	for (int i = threadIdx.x; i < RR / 16; i += blockDim.x) dst[i] = src ? s4[i] : make_uint4(0, 0, 0, 0);
	for (int y = threadIdx.x; y < R; y += blockDim.x) {
		xlo_dst[(size_t)ds * RR + (size_t)zslice * R + y] = src ? src[(size_t)y * R] : (uint8_t)0;
		xhi_dst[(size_t)ds * RR + (size_t)zslice * R + y] = src ? src[(size_t)y * R + R - 1] : (uint8_t)0;
	}
}

} // namespace

struct vp_multi {
	int n = 0;
	std::vector<vp_ctx *> ctx;
	int rows_per = 0, per_row = 0;
	std::vector<cudaEvent_t> pushed[2];      // [dir][i]: device i's push towards below (0) / above (1) is done
	std::vector<cudaEvent_t> done;           // device i finished the step (its ghost slices may be overwritten)
	// batch bookkeeping
	std::vector<std::vector<uint32_t>> ids, pos;         // per device: chunk ids of the batch, their positions in the caller's list
	std::vector<std::vector<uint8_t>> flags;
	std::vector<uint8_t> owner_of;
	uint32_t batch_n = 0;
	std::string err;
};

static thread_local std::string g_multi_err;
static int mfail(vp_multi *m, int code, const std::string &what) { if (m) m->err = what; else g_multi_err = what; return code; }

extern "C" const char *vp_multi_last_error(const vp_multi *m) { return m ? m->err.c_str() : g_multi_err.c_str(); }

extern "C" void vp_multi_destroy(vp_multi *m)
{
	if (!m) return;
	for (int i = 0; i < m->n; i++) {
		if (!m->ctx[i]) continue;
		cudaSetDevice(m->ctx[i]->cfg.device);
		cudaDeviceSynchronize();
	}
	for (int i = 0; i < (int)m->done.size(); i++) {
		cudaSetDevice(m->ctx[i]->cfg.device);
		if (m->pushed[0][i]) cudaEventDestroy(m->pushed[0][i]);
		if (m->pushed[1][i]) cudaEventDestroy(m->pushed[1][i]);
		if (m->done[i]) cudaEventDestroy(m->done[i]);
	}
	for (int i = 0; i < m->n; i++) vp_ctx_destroy(m->ctx[i]);
	delete m;
}

extern "C" int vp_multi_create(const vp_config *base, const int32_t *devices, int32_t ndev, vp_multi **out)
{
	if (!base || !out || ndev < 1 || ndev > 64) return mfail(nullptr, VP_ERR_ARG, "vp_multi_create: bad argument");
	*out = nullptr;
	const int nz = 1 << base->max_bitw[2];
	if (nz % ndev) return mfail(nullptr, VP_ERR_ARG, "vp_multi_create: the device count must divide the chunk rows of the world");
	vp_multi *m = new (std::nothrow) vp_multi();
	if (!m) return mfail(nullptr, VP_ERR_ARG, "out of host memory");
	m->n = ndev; m->ctx.assign(ndev, nullptr);
	m->rows_per = nz / ndev;
	m->per_row = 1 << (base->max_bitw[0] + base->max_bitw[1]);
	for (int i = 0; i < ndev; i++) {
		vp_config cfg = *base;
		cfg.device = devices ? devices[i] : i;
		cfg.slab_z0 = i * m->rows_per; cfg.slab_z1 = (i + 1) * m->rows_per;
		const int rc = vp_ctx_create(&cfg, &m->ctx[i]);
		if (rc) { const std::string e = vp_last_error(nullptr); vp_multi_destroy(m); return mfail(nullptr, rc, "vp_multi_create: " + e); }
	}
	// slab neighbours write into each other's pools
	for (int i = 0; i < ndev; i++)
		for (int j : {i - 1, i + 1}) {
			if (j < 0 || j >= ndev) continue;
			const int a = m->ctx[i]->cfg.device, b = m->ctx[j]->cfg.device;
			if (a == b) continue;                          // several slabs on one device (tests): plain device memory
			int can = 0;
			cudaDeviceCanAccessPeer(&can, a, b);
			if (!can) { vp_multi_destroy(m); return mfail(nullptr, VP_ERR_CUDA, "vp_multi_create: no peer access between slab neighbours"); }
			cudaSetDevice(a);
			const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { vp_multi_destroy(m); return mfail(nullptr, VP_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
			cudaGetLastError();
		}
	m->pushed[0].assign(ndev, nullptr); m->pushed[1].assign(ndev, nullptr); m->done.assign(ndev, nullptr);
	for (int i = 0; i < ndev; i++) {
		cudaSetDevice(m->ctx[i]->cfg.device);
		cudaEventCreateWithFlags(&m->pushed[0][i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&m->pushed[1][i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&m->done[i], cudaEventDisableTiming);
		cudaEventRecord(m->done[i], m->ctx[i]->stream);
	}
	m->ids.resize(ndev); m->pos.resize(ndev); m->flags.resize(ndev);
	*out = m;
	return VP_OK;
}

extern "C" int32_t vp_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" int32_t vp_multi_devices(const vp_multi *m) { return m ? m->n : 0; }
extern "C" vp_ctx *vp_multi_ctx(vp_multi *m, int32_t i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
extern "C" int32_t vp_multi_owner(const vp_multi *m, uint32_t chunk_id)
{
	if (!m) return -1;
	const uint32_t cz = chunk_id / (uint32_t)m->per_row;
	const int o = (int)(cz / (uint32_t)m->rows_per);
	return o < m->n ? o : -1;
}

#define M_CUDA(m, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return mfail(m, VP_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(e__)); } while (0)
#define M_CTX(m, i, call) do { int rc__ = (call); if (rc__) return mfail(m, rc__, std::string(#call ": ") + vp_last_error((m)->ctx[i])); } while (0)

// Border planes of every slab, written straight into the neighbours' ghost chunks.  mesh != 0 also sends the planes that
// only mesh AO needs (z-slice R-1 of a slab's last row, to the slab above).
extern "C" int vp_multi_exchange_halos(vp_multi *m, int32_t mesh)
{
	if (!m) return VP_ERR_ARG;
	for (int i = 0; i < m->n; i++) {
		vp_ctx *c = m->ctx[i];
		M_CUDA(m, cudaSetDevice(c->cfg.device));
		M_CUDA(m, cudaEventRecord(c->ev_bready, c->stream));
		M_CUDA(m, cudaStreamWaitEvent(c->border_stream, c->ev_bready, 0));
		for (int dir = 0; dir < 2; dir++) {
			const int j = dir == 0 ? i - 1 : i + 1;
			if (j < 0 || j >= m->n || (dir == 1 && !mesh)) continue;
			vp_ctx *p = m->ctx[j];
			// dir 0: z-slice 0 of my first row -> ghost row slab_z1 of the slab below (its +z halo);
			// dir 1: z-slice R-1 of my last row -> ghost row slab_z0 - 1 of the slab above (its -z halo)
			const int src_row = dir == 0 ? c->cfg.slab_z0 : c->cfg.slab_z1 - 1;
			const int dst_row = dir == 0 ? p->cfg.slab_z1 : p->cfg.slab_z0 - 1;
			const int32_t *src_slots = c->d_slot + (size_t)(src_row - c->ez0) * m->per_row;
			const int32_t *dst_slots = p->d_slot + (size_t)(dst_row - p->ez0) * m->per_row;
			M_CUDA(m, cudaStreamWaitEvent(c->border_stream, m->done[j], 0));      // the neighbour no longer reads the old plane
			k_halo_push<<<m->per_row, 256, 0, c->border_stream>>>(c->rb, c->vox_pool, src_slots, p->vox_pool, p->xlo_pool, p->xhi_pool, dst_slots,
			                                                       dir == 0 ? 0 : c->R - 1);
			M_CUDA(m, cudaGetLastError());
			c->launches++;
			M_CUDA(m, cudaEventRecord(m->pushed[dir][i], c->border_stream));
		}
	}
	// receivers: the border stream of slab j waits for the planes pushed into it
	for (int j = 0; j < m->n; j++) {
		vp_ctx *p = m->ctx[j];
		M_CUDA(m, cudaSetDevice(p->cfg.device));
		if (j + 1 < m->n) M_CUDA(m, cudaStreamWaitEvent(p->border_stream, m->pushed[0][j + 1], 0));
		if (j > 0 && mesh) M_CUDA(m, cudaStreamWaitEvent(p->border_stream, m->pushed[1][j - 1], 0));
		// host-facing calls that follow on the context stream (vp_rebuild_batch of border chunks) see the planes too
		M_CUDA(m, cudaEventRecord(p->ev_bjoin, p->border_stream));
		M_CUDA(m, cudaStreamWaitEvent(p->stream, p->ev_bjoin, 0));
	}
	return VP_OK;
}

// ---- residency, routed to the owning device -------------------------------------------------------------------------

static int split_by_owner(vp_multi *m, const uint32_t *ids, uint32_t n, std::vector<std::vector<uint32_t>> &idx)
{
	idx.assign(m->n, {});
	for (uint32_t i = 0; i < n; i++) {
		const int o = vp_multi_owner(m, ids[i]);
		if (o < 0) return mfail(m, VP_ERR_NOT_RESIDENT, "chunk id outside the world");
		idx[o].push_back(i);
	}
	return VP_OK;
}

extern "C" int vp_multi_upload_chunks_dense(vp_multi *m, const uint32_t *ids, uint32_t n, const uint8_t *host)
{
	if (!m || (n && (!ids || !host))) return mfail(m, VP_ERR_ARG, "vp_multi_upload_chunks_dense: null argument");
	std::vector<std::vector<uint32_t>> idx;
	int rc = split_by_owner(m, ids, n, idx);
	if (rc) return rc;
	const size_t N = (size_t)1 << (3 * m->ctx[0]->rb);
	for (int d = 0; d < m->n; d++) {
		// runs of consecutive list positions go up in one call without a copy
		size_t k = 0;
		while (k < idx[d].size()) {
			size_t e = k + 1;
			while (e < idx[d].size() && idx[d][e] == idx[d][e - 1] + 1) e++;
			M_CTX(m, d, vp_upload_chunks_dense(m->ctx[d], ids + idx[d][k], (uint32_t)(e - k), host + (size_t)idx[d][k] * N));
			k = e;
		}
	}
	return VP_OK;
}

extern "C" int vp_multi_set_chunks_null(vp_multi *m, const uint32_t *ids, uint32_t n)
{
	if (!m || (n && !ids)) return mfail(m, VP_ERR_ARG, "vp_multi_set_chunks_null: null argument");
	std::vector<std::vector<uint32_t>> idx;
	int rc = split_by_owner(m, ids, n, idx);
	if (rc) return rc;
	for (int d = 0; d < m->n; d++) {
		std::vector<uint32_t> sub;
		for (uint32_t i : idx[d]) sub.push_back(ids[i]);
		if (!sub.empty()) M_CTX(m, d, vp_set_chunks_null(m->ctx[d], sub.data(), (uint32_t)sub.size()));
	}
	return VP_OK;
}

extern "C" int vp_multi_upload_chunks_rle(vp_multi *m, const uint32_t *ids, uint32_t n, const uint32_t *words, const uint64_t *word_offsets)
{
	if (!m || (n && (!ids || !words || !word_offsets))) return mfail(m, VP_ERR_ARG, "vp_multi_upload_chunks_rle: null argument");
	std::vector<std::vector<uint32_t>> idx;
	int rc = split_by_owner(m, ids, n, idx);
	if (rc) return rc;
	for (int d = 0; d < m->n; d++) {
		if (idx[d].empty()) continue;
		std::vector<uint32_t> sub, w;
		std::vector<uint64_t> off(1, 0);
		for (uint32_t i : idx[d]) {
			sub.push_back(ids[i]);
			w.insert(w.end(), words + word_offsets[i], words + word_offsets[i + 1]);
			off.push_back(w.size());
		}
		M_CTX(m, d, vp_upload_chunks_rle(m->ctx[d], sub.data(), (uint32_t)sub.size(), w.data(), off.data()));
	}
	return VP_OK;
}

// every device keeps the rows of its slab (+17 rows of reach); rows outside are ignored by vp_upload_shadow_rows
extern "C" int vp_multi_upload_shadow_rows(vp_multi *m, uint32_t z0, uint32_t z1, const uint16_t *rows)
{
	if (!m) return VP_ERR_ARG;
	for (int d = 0; d < m->n; d++) M_CTX(m, d, vp_upload_shadow_rows(m->ctx[d], z0, z1, rows));
	return VP_OK;
}

// ---- rebuild ------------------------------------------------------------------------------------------------------------

extern "C" int vp_multi_batch_prepare(vp_multi *m, const uint32_t *ids, uint32_t n, const uint8_t *per_chunk_flags, uint32_t flags)
{
	if (!m || (n && !ids)) return mfail(m, VP_ERR_ARG, "vp_multi_batch_prepare: null argument");
	std::vector<std::vector<uint32_t>> idx;
	int rc = split_by_owner(m, ids, n, idx);
	if (rc) return rc;
	m->owner_of.assign(n, 0);
	for (int d = 0; d < m->n; d++) {
		m->ids[d].clear(); m->pos[d] = idx[d]; m->flags[d].clear();
		for (uint32_t i : idx[d]) {
			m->ids[d].push_back(ids[i]);
			m->flags[d].push_back((uint8_t)(per_chunk_flags ? per_chunk_flags[i] : flags));
			m->owner_of[i] = (uint8_t)d;
		}
		M_CTX(m, d, vp_batch_prepare(m->ctx[d], m->ids[d].data(), (uint32_t)m->ids[d].size(), m->flags[d].data(), 0));
	}
	m->batch_n = n;
	return VP_OK;
}

// One device-resident step on all devices: interior chunks, border planes pushed peer to peer, border chunks.
extern "C" int vp_multi_rebuild_device(vp_multi *m, int32_t mesh)
{
	if (!m) return VP_ERR_ARG;
	for (int d = 0; d < m->n; d++) M_CTX(m, d, vp_rebuild_device_part(m->ctx[d], 0));
	int rc = vp_multi_exchange_halos(m, mesh);
	if (rc) return rc;
	for (int d = 0; d < m->n; d++) {
		M_CTX(m, d, vp_rebuild_device_part(m->ctx[d], 1));
		M_CUDA(m, cudaSetDevice(m->ctx[d]->cfg.device));
		M_CUDA(m, cudaEventRecord(m->done[d], m->ctx[d]->stream));
	}
	return VP_OK;
}

extern "C" int vp_multi_synchronize(vp_multi *m)
{
	if (!m) return VP_ERR_ARG;
	for (int d = 0; d < m->n; d++) M_CTX(m, d, vp_ctx_synchronize(m->ctx[d]));
	return VP_OK;
}

// vp_rebuild_batch over all devices: results[i] is relative to splat_bases[owner[i]] / mesh_bases[owner[i]] (pinned
// staging of the owning device's context, valid until the next rebuild call).
extern "C" int vp_multi_rebuild_batch(vp_multi *m, const uint32_t *ids, uint32_t n, uint32_t flags, const uint8_t *per_chunk_flags,
                                      vp_chunk_result *results, uint8_t *owner, const void **splat_bases, const void **mesh_bases)
{
	if (!m || !results || (n && !ids)) return mfail(m, VP_ERR_ARG, "vp_multi_rebuild_batch: null argument");
	int rc = vp_multi_batch_prepare(m, ids, n, per_chunk_flags, flags);
	if (rc) return rc;
	bool mesh = false;
	for (uint32_t i = 0; i < n && !mesh; i++) mesh = ((per_chunk_flags ? per_chunk_flags[i] : flags) & VP_REBUILD_MESH) != 0;
	if ((rc = vp_multi_rebuild_device(m, mesh ? 1 : 0))) return rc;
	// all devices are running; collect device by device (records, then the arenas into the pinned staging)
	std::vector<std::vector<vp_chunk_result>> r(m->n);
	for (int d = 0; d < m->n; d++) {
		vp_ctx *c = m->ctx[d];
		r[d].resize(m->ids[d].size());
		uint64_t sb = 0, mb = 0;
		M_CTX(m, d, vp_rebuild_device_results(c, r[d].data(), &sb, &mb));
		const void *sbase = nullptr, *mbase = nullptr;
		M_CTX(m, d, vp_stage_arenas(c, sb, mb, &sbase, &mbase));
		if (splat_bases) splat_bases[d] = sbase;
		if (mesh_bases) mesh_bases[d] = mbase;
		for (size_t k = 0; k < r[d].size(); k++) results[m->pos[d][k]] = r[d][k];
	}
	if (owner) memcpy(owner, m->owner_of.data(), n);
	return VP_OK;
}
