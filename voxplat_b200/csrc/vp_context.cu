// vp_context.cu -- host side of the C ABI (include/voxplat_b200.h): context, resident world,
// uploads, batch rebuild orchestration, staging.  No CPU compute path exists here: every data
// transformation is a CUDA kernel (vp_splat.cu, vp_mesh.cu, vp_rle.cu, the small kernels below).
#include "vp_internal.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <new>

static thread_local std::string g_create_err;

int vp_fail(vp_ctx *c, int code, const char *what, cudaError_t e)
{
	char buf[512];
	if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	else snprintf(buf, sizeof buf, "%s", what);
	if (c) c->err = buf; else g_create_err = buf;
	return code;
}

extern "C" const char *vp_last_error(const vp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" const char *vp_version(void) { return "voxplat_b200 0.1 (sm_100a)"; }

VpWorldDev vp_world_dev(const vp_ctx *c)
{
	VpWorldDev w;
	w.rb = c->rb;
	for (int i = 0; i < 3; i++) w.bits[i] = c->cfg.max_bitw[i];
	w.ez0 = c->ez0; w.ez1 = c->ez1;
	w.slot = c->d_slot;
	w.vox_pool = c->vox_pool; w.xlo_pool = c->xlo_pool; w.xhi_pool = c->xhi_pool;
	w.shadow = c->d_shadow; w.sh_z0 = c->sh_z0; w.sh_z1 = c->sh_z1;
	w.sh_w = (uint32_t)((c->nx + c->ny) << c->rb);
	return w;
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------

// Copies of the x = 0 and x = R-1 planes of freshly written chunks (the contiguous +x / -x halo sources).
// One CTA per chunk, one thread per (z,y) row: two byte loads at stride R, two coalesced byte stores.
__global__ void k_extract_xfaces(int rb, const uint8_t *__restrict__ vox_pool, uint8_t *__restrict__ xlo_pool,
                                 uint8_t *__restrict__ xhi_pool, const int32_t *__restrict__ slots)
{
	const int R = 1 << rb, RR = R * R;
	const int slot = slots[blockIdx.x];
	if (slot < 0) return;
	const uint8_t *v = vox_pool + ((size_t)slot << (3 * rb));
	uint8_t *lo = xlo_pool + (size_t)slot * RR, *hi = xhi_pool + (size_t)slot * RR;
	for (int r = threadIdx.x; r < RR; r += blockDim.x) {
		lo[r] = v[(size_t)r << rb];
		hi[r] = v[((size_t)r << rb) + R - 1];
	}
}

cudaError_t vp_launch_extract_xfaces(int rb, const uint8_t *vox_pool, uint8_t *xlo_pool, uint8_t *xhi_pool,
                                     const int32_t *d_slots, uint32_t n, cudaStream_t s)
{
	if (!n) return cudaSuccess;
	k_extract_xfaces<<<n, 256, 0, s>>>(rb, vox_pool, xlo_pool, xhi_pool, d_slots);
	return cudaGetLastError();
}

// Border plane pack / unpack: one R*R slice per chunk column of a chunk row <-> contiguous plane.
__global__ void k_plane_copy(int rb, uint8_t *__restrict__ vox_pool, const int32_t *__restrict__ row_slots,
                             uint8_t *__restrict__ plane, int zslice, int unpack)
{
	const int R = 1 << rb, RR = R * R;
	const int slot = row_slots[blockIdx.x];
	uint4 *pl = reinterpret_cast<uint4 *>(plane + (size_t)blockIdx.x * RR);
	if (slot < 0) {
		if (!unpack) for (int i = threadIdx.x; i < RR / 16; i += blockDim.x) pl[i] = make_uint4(0, 0, 0, 0);
		return;
	}
	uint4 *sl = reinterpret_cast<uint4 *>(vox_pool + ((size_t)slot << (3 * rb)) + (size_t)zslice * RR);
	for (int i = threadIdx.x; i < RR / 16; i += blockDim.x) { if (unpack) sl[i] = pl[i]; else pl[i] = sl[i]; }
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------

static void ctx_free(vp_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->cfg.device);
	cudaFree(c->vox_pool); cudaFree(c->xlo_pool); cudaFree(c->xhi_pool); cudaFree(c->d_slot); cudaFree(c->d_shadow);
	cudaFree(c->d_ids); cudaFree(c->d_flags); cudaFree(c->d_splat_ids); cudaFree(c->d_mesh_ids); cudaFree(c->d_results);
	cudaFree(c->d_splat_pos); cudaFree(c->d_mesh_pos);
	cudaFree(c->d_splat_arena); cudaFree(c->d_mesh_arena); cudaFree(c->d_rle_arena); cudaFree(c->d_arena_state);
	cudaFree(c->d_splat_scratch); cudaFree(c->d_mesh_scratch); cudaFree(c->d_splat_scratch_b); cudaFree(c->d_mesh_scratch_b);
	cudaFree(c->d_tmp_slots); cudaFree(c->d_io); cudaFree(c->d_node_arena); cudaFree(c->d_nodes); cudaFreeHost(c->h_node_stage);
	cudaFreeHost(c->h_results); cudaFreeHost(c->h_arena_state); cudaFreeHost(c->h_splat_stage); cudaFreeHost(c->h_mesh_stage);
	cudaFreeHost(c->h_io_stage);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->mesh_stream) cudaStreamDestroy(c->mesh_stream);
	if (c->border_stream) cudaStreamDestroy(c->border_stream);
	if (c->ev_reset) cudaEventDestroy(c->ev_reset);
	if (c->ev_bjoin) cudaEventDestroy(c->ev_bjoin);
	if (c->ev_bready) cudaEventDestroy(c->ev_bready);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->ev_join) cudaEventDestroy(c->ev_join);
	if (c->down_stream) cudaStreamDestroy(c->down_stream);
	for (int k = 0; k < 2; k++) for (int i = 0; i < 64; i++) if (c->ev_pipe[k][i]) cudaEventDestroy(c->ev_pipe[k][i]);
	cudaFreeHost(c->h_steps);
	if (c->ev_a) cudaEventDestroy(c->ev_a);
	if (c->ev_b) cudaEventDestroy(c->ev_b);
	for (int h = 0; h < vp_ctx::kHist; h++) for (int i = 0; i < 4; i++) if (c->ev_k[h][i]) cudaEventDestroy(c->ev_k[h][i]);
	delete c;
}

static int ghost_row_init(vp_ctx *c, int row);

extern "C" int vp_ctx_create(const vp_config *cfg, vp_ctx **out)
{
	if (!cfg || !out) return vp_fail(nullptr, VP_ERR_ARG, "vp_ctx_create: null argument");
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return vp_fail(nullptr, VP_ERR_NO_DEVICE, "vp_ctx_create: no CUDA device (this library has no CPU path)", e);
	if (cfg->device < 0 || cfg->device >= ndev) return vp_fail(nullptr, VP_ERR_ARG, "vp_ctx_create: bad device ordinal");
	if (cfg->root_bitw < 4 || cfg->root_bitw > 7) return vp_fail(nullptr, VP_ERR_ARG, "vp_ctx_create: root_bitw must be 4..7 (chunk edge 16..128)");
	for (int i = 0; i < 3; i++)
		if (cfg->max_bitw[i] < 0 || cfg->max_bitw[i] + cfg->root_bitw > 15)
			return vp_fail(nullptr, VP_ERR_ARG, "vp_ctx_create: world dimension must be <= 32768 voxels");
	vp_ctx *c = new (std::nothrow) vp_ctx();
	if (!c) return vp_fail(nullptr, VP_ERR_ARG, "out of host memory");
	c->cfg = *cfg;
	c->rb = cfg->root_bitw; c->R = 1 << c->rb;
	c->nx = 1 << cfg->max_bitw[0]; c->ny = 1 << cfg->max_bitw[1]; c->nz = 1 << cfg->max_bitw[2];
	if (c->cfg.slab_z1 <= c->cfg.slab_z0) { c->cfg.slab_z0 = 0; c->cfg.slab_z1 = c->nz; }
	if (c->cfg.slab_z0 < 0 || c->cfg.slab_z1 > c->nz) { delete c; return vp_fail(nullptr, VP_ERR_ARG, "vp_ctx_create: slab outside the world"); }
	c->ez0 = std::max(0, c->cfg.slab_z0 - 1);
	c->ez1 = std::min(c->nz, c->cfg.slab_z1 + 1);
	c->n_ext = (uint32_t)(c->ez1 - c->ez0) * c->nx * c->ny;
	c->n_slots = c->n_ext;
	const size_t N = (size_t)1 << (3 * c->rb), RR = (size_t)1 << (2 * c->rb);
	const size_t owned_vox = (size_t)(c->cfg.slab_z1 - c->cfg.slab_z0) * c->nx * c->ny * N;
	if (!c->cfg.splat_arena_bytes) c->cfg.splat_arena_bytes = std::max<size_t>(owned_vox * 2, 16u << 20);
	if (!c->cfg.mesh_arena_bytes) c->cfg.mesh_arena_bytes = std::max<size_t>(owned_vox, 16u << 20);
	if (!c->cfg.rle_arena_bytes) c->cfg.rle_arena_bytes = std::max<size_t>(owned_vox / 2, 16u << 20);

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { int rc = vp_fail(nullptr, VP_ERR_CUDA, #call, e__); ctx_free(c); return rc; } } while (0)
	CK(cudaSetDevice(cfg->device));
	CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	{
		// Scheduling priorities.  The context stream (the bulk of a rebuild) runs at the default = lowest priority.  The border
		// stream of a slab context gets the highest: its few chunks wait for the neighbour's plane and must not queue behind
		// the thousands of interior CTAs launched before them, or the step ends one kernel later than it has to.
		// VP_MESH_PRIO / VP_BORDER_PRIO = hi | lo override (measurements: DESIGN.md section 6).
		int lo = 0, hi = 0;
		CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		const char *e = getenv("VP_MESH_PRIO");
		// the mesh kernels (a few hundred near-field chunks) run beside the splat kernels; at high priority they finish early
		// instead of trailing the step: 0.441 vs 0.474 ms per C2 step
		CK(cudaStreamCreateWithPriority(&c->mesh_stream, cudaStreamNonBlocking, (e && e[0] == 'l') ? lo : hi));
		const char *b = getenv("VP_BORDER_PRIO");
		CK(cudaStreamCreateWithPriority(&c->border_stream, cudaStreamNonBlocking, (b && b[0] == 'l') ? lo : hi));
	}
	CK(cudaEventCreateWithFlags(&c->ev_reset, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&c->ev_bjoin, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&c->ev_bready, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
	CK(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
	for (int k = 0; k < 2; k++) for (int i = 0; i < 64; i++) CK(cudaEventCreateWithFlags(&c->ev_pipe[k][i], cudaEventDisableTiming));
	CK(cudaHostAlloc(&c->h_steps, 64 * 2 * sizeof(VpArenaDev) + 64 * sizeof(uint32_t), cudaHostAllocDefault));
	memset(c->h_steps, 0, 64 * 2 * sizeof(VpArenaDev) + 64 * sizeof(uint32_t));
	CK(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
	for (int h = 0; h < vp_ctx::kHist; h++) for (int i = 0; i < 4; i++) CK(cudaEventCreate(&c->ev_k[h][i]));
	c->stream = c->own_stream;
	CK(cudaMalloc(&c->vox_pool, (size_t)c->n_slots * N));
	CK(cudaMalloc(&c->xlo_pool, (size_t)c->n_slots * RR));
	CK(cudaMalloc(&c->xhi_pool, (size_t)c->n_slots * RR));
	CK(cudaMalloc(&c->d_slot, (size_t)c->n_ext * sizeof(int32_t)));
	CK(cudaMalloc(&c->d_tmp_slots, (size_t)c->n_ext * sizeof(int32_t)));
	c->h_slot.assign(c->n_ext, -1);
	c->free_slots.resize(c->n_slots);
	for (uint32_t i = 0; i < c->n_slots; i++) c->free_slots[i] = c->n_slots - 1 - i;      // pop_back hands out 0,1,2,...
	CK(cudaMemsetAsync(c->d_slot, 0xFF, (size_t)c->n_ext * sizeof(int32_t), c->stream));
	// shadow rows: owned voxel rows + 17 rows of reach (LOD-4 samples at z+16, mesh diamond at z+1), plus the
	// out-of-bounds slack of SURVEY 8a' u3
	const uint32_t shw = (uint32_t)((c->nx + c->ny) << c->rb), Zw = (uint32_t)c->nz << c->rb;
	c->sh_z0 = (uint32_t)c->cfg.slab_z0 << c->rb;
	c->sh_z1 = std::min<uint32_t>(Zw, ((uint32_t)c->cfg.slab_z1 << c->rb) + 17);
	c->shadow_entries = (size_t)(c->sh_z1 - c->sh_z0) * shw + (size_t)17 * shw + 64;
	CK(cudaMalloc(&c->d_shadow, c->shadow_entries * sizeof(uint16_t)));
	CK(cudaMemsetAsync(c->d_shadow, 0, c->shadow_entries * sizeof(uint16_t), c->stream));
	CK(cudaMalloc(&c->d_splat_arena, c->cfg.splat_arena_bytes));
	CK(cudaMalloc(&c->d_mesh_arena, c->cfg.mesh_arena_bytes));
	CK(cudaMalloc(&c->d_rle_arena, c->cfg.rle_arena_bytes));
	CK(cudaMalloc(&c->d_arena_state, 3 * sizeof(VpArenaDev)));
	CK(cudaHostAlloc(&c->h_arena_state, 6 * sizeof(VpArenaDev), cudaHostAllocDefault));     // [0..2] readback, [3..5] reset template
	memset(c->h_arena_state, 0, 6 * sizeof(VpArenaDev));
	CK(cudaStreamSynchronize(c->stream));
#undef CK
	// Ghost rows (the border slices of the slab neighbours) own permanent zero-filled slots from the start, so that a
	// border exchange never changes which chunks are resident.
	for (int side = 0; side < 2; side++) {
		const int row = side ? c->cfg.slab_z1 : c->cfg.slab_z0 - 1;
		if (row < c->ez0 || row >= c->ez1 || (row >= c->cfg.slab_z0 && row < c->cfg.slab_z1)) continue;
		const int rc = ghost_row_init(c, row);
		if (rc) { g_create_err = c->err; ctx_free(c); return rc; }
	}
	*out = c;
	return VP_OK;
}

extern "C" void vp_ctx_destroy(vp_ctx *ctx) { if (ctx) { cudaSetDevice(ctx->cfg.device); cudaDeviceSynchronize(); ctx_free(ctx); } }

extern "C" int vp_ctx_set_stream(vp_ctx *c, void *cuda_stream)
{
	if (!c) return VP_ERR_ARG;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
	return VP_OK;
}

extern "C" int vp_ctx_synchronize(vp_ctx *c)
{
	if (!c) return VP_ERR_ARG;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

extern "C" uint64_t vp_kernel_launches(vp_ctx *c, int reset)
{
	uint64_t v = c->launches;
	if (reset) c->launches = 0;
	return v;
}

// pinned staging that only grows
static int stage_reserve(vp_ctx *c, uint8_t **buf, size_t *cap, size_t need)
{
	if (need <= *cap) return VP_OK;
	if (*buf) { VP_CUDA(c, cudaStreamSynchronize(c->stream)); VP_CUDA(c, cudaFreeHost(*buf)); *buf = nullptr; *cap = 0; }
	size_t want = std::max(need, (size_t)1 << 20);
	VP_CUDA(c, cudaHostAlloc((void **)buf, want, cudaHostAllocDefault));
	*cap = want;
	return VP_OK;
}

// extended-slab index of a world chunk id, or -1 when the row is not held here
static inline int64_t ext_index(const vp_ctx *c, uint32_t id)
{
	const uint32_t per_row = (uint32_t)c->nx * c->ny;
	const uint32_t cz = id / per_row;
	if (id >= per_row * (uint32_t)c->nz || (int)cz < c->ez0 || (int)cz >= c->ez1) return -1;
	return (int64_t)id - (int64_t)c->ez0 * per_row;
}

static inline bool all_zero(const uint8_t *p, size_t n)
{
	const uint64_t *q = reinterpret_cast<const uint64_t *>(p);      // chunk volumes are multiples of 8
	uint64_t acc = 0;
	for (size_t i = 0; i < n / 8; i++) { acc |= q[i]; if ((i & 511) == 511 && acc) return false; }
	return acc == 0;
}

// assign / release slots for a list of chunks; want[i] != 0 means the chunk needs storage.  All ids and the pool capacity
// are checked BEFORE any state changes, so a failing call leaves the host and device slot tables as they were.
static int assign_slots(vp_ctx *c, const uint32_t *ids, uint32_t n, const uint8_t *want, std::vector<int32_t> &slots)
{
	slots.resize(n);
	size_t need = 0, freed = 0;
	for (uint32_t i = 0; i < n; i++) {
		const int64_t e = ext_index(c, ids[i]);
		if (e < 0) return vp_fail(c, VP_ERR_NOT_RESIDENT, "chunk id outside this context's slab");
		const int32_t s = c->h_slot[(size_t)e];
		if (want[i]) need += s < 0; else freed += s >= 0;
	}
	// (a list that names a chunk twice may over-count `need`; that only makes the check conservative)
	if (need > c->free_slots.size() + freed) return vp_fail(c, VP_ERR_ARENA_FULL, "chunk pool exhausted");
	for (int pass = 0; pass < 2; pass++)               // releases first: their slots may be handed out again in the same call
		for (uint32_t i = 0; i < n; i++) {
			if ((want[i] != 0) != (pass == 1)) continue;
			const int64_t e = ext_index(c, ids[i]);
			int32_t s = c->h_slot[(size_t)e];
			if (want[i]) {
				if (s < 0) {
					if (c->free_slots.empty()) return vp_fail(c, VP_ERR_ARENA_FULL, "chunk pool exhausted");      // unreachable: checked above
					s = (int32_t)c->free_slots.back(); c->free_slots.pop_back();
					c->h_slot[(size_t)e] = s;
					c->residency_epoch++;
				}
			} else if (s >= 0) {
				c->free_slots.push_back((uint32_t)s);
				c->h_slot[(size_t)e] = -1; s = -1;
				c->residency_epoch++;
			}
			slots[i] = s;
		}
	return VP_OK;
}

// push the slot-table entries of the listed chunks: one copy of the table range they span (the host table mirrors the
// device table entry for entry, so the untouched entries in between are rewritten with what they already hold; a copy
// per run of consecutive ids cost 0.4 ms of driver calls for the 2550 non-null chunks of the 2048x256x2048 world)
static int push_slot_table(vp_ctx *c, const uint32_t *ids, uint32_t n)
{
	if (!n) return VP_OK;
	int64_t lo = ext_index(c, ids[0]), hi = lo;
	for (uint32_t i = 1; i < n; i++) { const int64_t e = ext_index(c, ids[i]); lo = std::min(lo, e); hi = std::max(hi, e); }
	VP_CUDA(c, cudaMemcpyAsync(c->d_slot + lo, c->h_slot.data() + lo, (size_t)(hi - lo + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	return VP_OK;
}

extern "C" int vp_upload_chunks_dense(vp_ctx *c, const uint32_t *ids, uint32_t n, const uint8_t *host)
{
	if (!c || (n && (!ids || !host))) return vp_fail(c, VP_ERR_ARG, "vp_upload_chunks_dense: null argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const size_t N = (size_t)1 << (3 * c->rb);
	std::vector<uint8_t> want(n);
	for (uint32_t i = 0; i < n; i++) want[i] = !all_zero(host + (size_t)i * N, N);      // chunkset.c:225-228
	std::vector<int32_t> slots;
	int rc = assign_slots(c, ids, n, want.data(), slots);
	if (rc) return rc;
	// coalesce runs where both the source chunks and the destination slots are consecutive
	uint32_t i = 0;
	while (i < n) {
		if (slots[i] < 0) { i++; continue; }
		uint32_t j = i + 1;
		while (j < n && slots[j] == slots[j - 1] + 1) j++;
		VP_CUDA(c, cudaMemcpyAsync(c->vox_pool + (size_t)slots[i] * N, host + (size_t)i * N, (size_t)(j - i) * N, cudaMemcpyHostToDevice, c->stream));
		i = j;
	}
	rc = push_slot_table(c, ids, n);
	if (rc) return rc;
	VP_CUDA(c, cudaMemcpyAsync(c->d_tmp_slots, slots.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, vp_launch_extract_xfaces(c->rb, c->vox_pool, c->xlo_pool, c->xhi_pool, c->d_tmp_slots, n, c->stream));
	c->launches++;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));      // `slots` / caller memory may go away
	return VP_OK;
}

// World generation on the device: the owned chunk rows are generated batch by batch into a staging
// buffer (k_gen_chunks), all-air chunks become null chunks exactly like an upload would make them (chunkset.c:225-228),
// the others are copied to their pool slots; then the height map rows of the slab (+17 rows of reach into the next chunk
// row, generated transiently) are built from the voxels (k_shadow_rows).  Same resident state as
// vp_upload_chunks_dense + vp_upload_shadow_rows of the host generator's world (csrc/vp_worldgen.c).
namespace {
struct DevBuf {
	void *p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
	template <class T> T *as() const { return static_cast<T *>(p); }
};
}

extern "C" int vp_generate_world(vp_ctx *c, uint32_t seed)
{
	if (!c) return VP_ERR_ARG;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const size_t N = (size_t)1 << (3 * c->rb);
	const uint32_t per_row = (uint32_t)c->nx * c->ny;
	const uint32_t first = (uint32_t)c->cfg.slab_z0 * per_row, count = (uint32_t)(c->cfg.slab_z1 - c->cfg.slab_z0) * per_row;
	const uint32_t Zw = (uint32_t)c->nz << c->rb, z_own_end = (uint32_t)c->cfg.slab_z1 << c->rb;
	if (((uint32_t)c->ny << c->rb) > 1024u) return vp_fail(c, VP_ERR_ARG, "vp_generate_world: worlds taller than 1024 voxels are generated on the host");
	const int bits[3] = {c->cfg.max_bitw[0], c->cfg.max_bitw[1], c->cfg.max_bitw[2]};
	const uint32_t B = std::max<uint32_t>(per_row, std::min<uint32_t>(count, 1024));       // a whole chunk row fits (reach rows below)
	const bool trace = getenv("VP_TRACE") != nullptr;
	auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t0 = now();
	auto lap = [&](const char *what) { if (trace) { cudaStreamSynchronize(c->stream); double t = now(); fprintf(stderr, "[vp_generate_world] %-24s %8.2f ms\n", what, t - t0); t0 = t; } };
	DevBuf staging, dids, dsolid, dslots, dtable;
	VP_CUDA(c, cudaMalloc(&staging.p, (size_t)B * N));
	VP_CUDA(c, cudaMalloc(&dids.p, (size_t)B * 4));
	VP_CUDA(c, cudaMalloc(&dsolid.p, (size_t)B * 4));
	VP_CUDA(c, cudaMalloc(&dslots.p, (size_t)B * 4));
	std::vector<uint32_t> ids(B), solid(B);
	std::vector<uint8_t> want(B);
	std::vector<int32_t> slots;
	lap("allocations");
	for (uint32_t done = 0; done < count; done += B) {
		const uint32_t n = std::min(B, count - done);
		for (uint32_t i = 0; i < n; i++) ids[i] = first + done + i;
		VP_CUDA(c, cudaMemcpyAsync(dids.p, ids.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, vp_launch_gen_chunks(seed, c->rb, bits, dids.as<uint32_t>(), n, staging.as<uint8_t>(), dsolid.as<uint32_t>(), c->stream));
		VP_CUDA(c, cudaMemcpyAsync(solid.data(), dsolid.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
		lap("k_gen_chunks + flags");
		for (uint32_t i = 0; i < n; i++) want[i] = solid[i] != 0;
		int rc = assign_slots(c, ids.data(), n, want.data(), slots);
		if (rc) return rc;
		if ((rc = push_slot_table(c, ids.data(), n))) return rc;
		VP_CUDA(c, cudaMemcpyAsync(dslots.p, slots.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, vp_launch_scatter_chunks(c->rb, staging.as<uint8_t>(), dslots.as<int32_t>(), n, c->vox_pool, c->stream));
		VP_CUDA(c, vp_launch_extract_xfaces(c->rb, c->vox_pool, c->xlo_pool, c->xhi_pool, dslots.as<int32_t>(), n, c->stream));
		c->launches += 3;
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
		lap("slots + scatter + xfaces");
	}
	// height map: rows of the owned chunk rows from the pool ...
	const uint32_t shw = (uint32_t)((c->nx + c->ny) << c->rb);
	std::vector<const uint8_t *> table((size_t)std::max(count, per_row));
	for (uint32_t i = 0; i < count; i++) {
		const int32_t s = c->h_slot[(size_t)ext_index(c, first + i)];
		table[i] = s >= 0 ? c->vox_pool + (size_t)s * N : nullptr;
	}
	VP_CUDA(c, cudaMalloc(&dtable.p, table.size() * sizeof(void *)));
	VP_CUDA(c, cudaMemcpyAsync(dtable.p, table.data(), (size_t)count * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
	cudaError_t e = vp_launch_shadow_rows(c->rb, dtable.as<const uint8_t *>(), c->nx, c->ny, (uint32_t)c->cfg.slab_z0, c->sh_z0, std::min(c->sh_z1, z_own_end), c->d_shadow, c->stream);
	if (e != cudaSuccess) return vp_fail(c, VP_ERR_CUDA, "vp_generate_world: shadow rows", e);
	c->launches++;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	lap("k_shadow_rows");
	// ... and the rows of reach past the slab (LOD splats sample up to 16 rows ahead) from the next chunk row, generated transiently
	if (c->sh_z1 > z_own_end && z_own_end < Zw) {
		for (uint32_t i = 0; i < per_row; i++) ids[i] = (uint32_t)c->cfg.slab_z1 * per_row + i;
		VP_CUDA(c, cudaMemcpyAsync(dids.p, ids.data(), (size_t)per_row * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, vp_launch_gen_chunks(seed, c->rb, bits, dids.as<uint32_t>(), per_row, staging.as<uint8_t>(), dsolid.as<uint32_t>(), c->stream));
		VP_CUDA(c, cudaMemcpyAsync(solid.data(), dsolid.p, (size_t)per_row * 4, cudaMemcpyDeviceToHost, c->stream));
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
		for (uint32_t i = 0; i < per_row; i++) table[i] = solid[i] ? staging.as<uint8_t>() + (size_t)i * N : nullptr;
		VP_CUDA(c, cudaMemcpyAsync(dtable.p, table.data(), (size_t)per_row * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
		e = vp_launch_shadow_rows(c->rb, dtable.as<const uint8_t *>(), c->nx, c->ny, (uint32_t)c->cfg.slab_z1, z_own_end, c->sh_z1,
		                          c->d_shadow + (size_t)(z_own_end - c->sh_z0) * shw, c->stream);
		if (e != cudaSuccess) return vp_fail(c, VP_ERR_CUDA, "vp_generate_world: shadow rows of reach", e);
		c->launches += 2;
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	return VP_OK;
}

// resident[i] = 1 when chunk ids[i] holds voxels on the device, 0 for the null chunk (chunkset.c:116-117, :225-228).
extern "C" int vp_chunks_resident(vp_ctx *c, const uint32_t *ids, uint32_t n, uint8_t *resident)
{
	if (!c || (n && (!ids || !resident))) return vp_fail(c, VP_ERR_ARG, "vp_chunks_resident: null argument");
	for (uint32_t i = 0; i < n; i++) {
		const int64_t e = ext_index(c, ids[i]);
		if (e < 0) return vp_fail(c, VP_ERR_NOT_RESIDENT, "chunk id outside this context's slab");
		resident[i] = c->h_slot[(size_t)e] >= 0;
	}
	return VP_OK;
}

extern "C" int vp_set_chunks_null(vp_ctx *c, const uint32_t *ids, uint32_t n)
{
	if (!c || (n && !ids)) return vp_fail(c, VP_ERR_ARG, "vp_set_chunks_null: null argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	std::vector<uint8_t> want(n, 0);
	std::vector<int32_t> slots;
	int rc = assign_slots(c, ids, n, want.data(), slots);
	if (rc) return rc;
	rc = push_slot_table(c, ids, n);
	if (rc) return rc;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

extern "C" int vp_download_chunks_dense(vp_ctx *c, const uint32_t *ids, uint32_t n, uint8_t *host)
{
	if (!c || (n && (!ids || !host))) return vp_fail(c, VP_ERR_ARG, "vp_download_chunks_dense: null argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const size_t N = (size_t)1 << (3 * c->rb);
	for (uint32_t i = 0; i < n; i++) {
		int64_t e = ext_index(c, ids[i]);
		if (e < 0) return vp_fail(c, VP_ERR_NOT_RESIDENT, "chunk id outside this context's slab");
		int32_t s = c->h_slot[(size_t)e];
		if (s < 0) memset(host + (size_t)i * N, 0, N);
		else VP_CUDA(c, cudaMemcpyAsync(host + (size_t)i * N, c->vox_pool + (size_t)s * N, N, cudaMemcpyDeviceToHost, c->stream));
	}
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

static int upload_shadow_rows(vp_ctx *c, uint32_t z0, uint32_t z1, const uint16_t *rows, bool wait)
{
	if (!c || !rows || z1 < z0) return vp_fail(c, VP_ERR_ARG, "vp_upload_shadow_rows: bad argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const uint32_t shw = (uint32_t)((c->nx + c->ny) << c->rb);
	uint32_t a = std::max(z0, c->sh_z0), b = std::min(z1, c->sh_z1);
	if (a < b)
		VP_CUDA(c, cudaMemcpyAsync(c->d_shadow + (size_t)(a - c->sh_z0) * shw, rows + (size_t)(a - z0) * shw,
		                           (size_t)(b - a) * shw * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
	if (wait) VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

extern "C" int vp_upload_shadow_rows(vp_ctx *c, uint32_t z0, uint32_t z1, const uint16_t *rows) { return upload_shadow_rows(c, z0, z1, rows, true); }
extern "C" int vp_upload_shadow_rows_async(vp_ctx *c, uint32_t z0, uint32_t z1, const uint16_t *rows) { return upload_shadow_rows(c, z0, z1, rows, false); }

// ------------------------------------------------------------------------------------------------
// rebuild
// ------------------------------------------------------------------------------------------------

// Pipeline step done: the arena states go straight into pinned host memory (zero-copy stores) followed by the step's
// ticket.  A cudaMemcpyAsync would queue behind the multi-megabyte downloads on the device-to-host copy engine and make
// every step look as long as a download; the host polls the ticket instead of sleeping on an event.
__global__ void k_publish_step(const VpArenaDev *__restrict__ d_state, VpArenaDev *h_state, volatile uint32_t *h_ticket, uint32_t ticket)
{
	if (threadIdx.x < 2) h_state[threadIdx.x] = d_state[threadIdx.x];
	__threadfence_system();
	__syncwarp();
	if (threadIdx.x == 0) *h_ticket = ticket;
}

static int batch_reserve(vp_ctx *c, uint32_t n)
{
	if (n <= c->batch_cap) return VP_OK;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(c->d_ids); cudaFree(c->d_flags); cudaFree(c->d_splat_ids); cudaFree(c->d_mesh_ids); cudaFree(c->d_results);
	cudaFree(c->d_splat_pos); cudaFree(c->d_mesh_pos);
	cudaFreeHost(c->h_results);
	c->d_ids = c->d_splat_ids = c->d_mesh_ids = c->d_splat_pos = c->d_mesh_pos = nullptr; c->d_flags = nullptr; c->d_results = nullptr; c->h_results = nullptr;
	c->batch_cap = 0;
	uint32_t cap = std::max<uint32_t>(n, 1024);
	VP_CUDA(c, cudaMalloc(&c->d_ids, (size_t)cap * 4));
	VP_CUDA(c, cudaMalloc(&c->d_flags, (size_t)cap));
	VP_CUDA(c, cudaMalloc(&c->d_splat_ids, (size_t)cap * 4));
	VP_CUDA(c, cudaMalloc(&c->d_mesh_ids, (size_t)cap * 4));
	VP_CUDA(c, cudaMalloc(&c->d_splat_pos, (size_t)cap * 4));
	VP_CUDA(c, cudaMalloc(&c->d_mesh_pos, (size_t)cap * 4));
	VP_CUDA(c, cudaMalloc(&c->d_results, (size_t)cap * sizeof(VpResultDev)));
	VP_CUDA(c, cudaHostAlloc(&c->h_results, (size_t)cap * sizeof(VpResultDev), cudaHostAllocDefault));
	c->batch_cap = cap;
	return VP_OK;
}

// Scratch of a splat / mesh rebuild of up to n chunks per launch (per-chunk arrival counters first: they start at zero
// and the kernels leave them at zero).
static int scratch_reserve(vp_ctx *c, uint8_t **buf, uint32_t *cap_chunks, uint32_t n, size_t (*bytes_for)(int, uint32_t))
{
	if (n <= *cap_chunks) return VP_OK;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->mesh_stream));
	VP_CUDA(c, cudaStreamSynchronize(c->border_stream));
	cudaFree(*buf); *buf = nullptr; *cap_chunks = 0;
	const uint32_t cap = std::max<uint32_t>(n + n / 8, 64);
	VP_CUDA(c, cudaMalloc(buf, bytes_for(c->rb, cap)));
	VP_CUDA(c, cudaMemsetAsync(*buf, 0, (size_t)cap * 4 + 256, c->stream));
	*cap_chunks = cap;
	return VP_OK;
}
static int splat_scratch_reserve(vp_ctx *c, uint32_t n) { return scratch_reserve(c, &c->d_splat_scratch, &c->splat_scratch_chunks, n, vp_splat_scratch_bytes); }
static int mesh_scratch_reserve(vp_ctx *c, uint32_t n) { return scratch_reserve(c, &c->d_mesh_scratch, &c->mesh_scratch_chunks, n, vp_mesh_scratch_bytes); }
static int splat_scratch_b_reserve(vp_ctx *c, uint32_t n) { return scratch_reserve(c, &c->d_splat_scratch_b, &c->splat_scratch_b_chunks, n, vp_splat_scratch_bytes); }
static int mesh_scratch_b_reserve(vp_ctx *c, uint32_t n) { return scratch_reserve(c, &c->d_mesh_scratch_b, &c->mesh_scratch_b_chunks, n, vp_mesh_scratch_bytes); }

extern "C" int vp_batch_prepare(vp_ctx *c, const uint32_t *ids, uint32_t n, const uint8_t *per_chunk_flags, uint32_t flags)
{
	if (!c || (n && !ids)) return vp_fail(c, VP_ERR_ARG, "vp_batch_prepare: null argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	int rc = batch_reserve(c, n);
	if (rc) return rc;
	std::vector<uint32_t> sid, spos, mid, mpos;
	for (uint32_t i = 0; i < n; i++) {
		const uint32_t per_row = (uint32_t)c->nx * c->ny, cz = ids[i] / per_row;
		if (ids[i] >= per_row * (uint32_t)c->nz || (int)cz < c->cfg.slab_z0 || (int)cz >= c->cfg.slab_z1)
			return vp_fail(c, VP_ERR_NOT_RESIDENT, "vp_batch_prepare: chunk id outside the owned slab");
		uint32_t f = per_chunk_flags ? per_chunk_flags[i] : flags;
		if (f & VP_REBUILD_SPLAT) {
			// mesher.c:404-409: a null chunk whose +x,+y,+z neighbours are null too has nothing visible; its
			// (zeroed) result record is final, so it is not launched at all
			const uint32_t cx = ids[i] % (uint32_t)c->nx, cy = (ids[i] / (uint32_t)c->nx) % (uint32_t)c->ny;
			auto slot_of = [&](uint32_t x, uint32_t y, uint32_t z) -> int32_t {
				if (x >= (uint32_t)c->nx || y >= (uint32_t)c->ny || z >= (uint32_t)c->nz || (int)z < c->ez0 || (int)z >= c->ez1) return -1;
				return c->h_slot[(size_t)(z - (uint32_t)c->ez0) * per_row + (size_t)y * c->nx + x];
			};
			if (slot_of(cx, cy, cz) >= 0 || slot_of(cx + 1, cy, cz) >= 0 || slot_of(cx, cy + 1, cz) >= 0 || slot_of(cx, cy, cz + 1) >= 0) {
				sid.push_back(ids[i]); spos.push_back(i);
			}
		}
		if (f & VP_REBUILD_MESH) { mid.push_back(ids[i]); mpos.push_back(i); }
	}
	// Chunks whose rebuild reads a ghost row (the border plane a slab neighbour sends every step) go to the END of the
	// lists: vp_rebuild_device_part(0) launches the others while the exchange is still in flight, part 1 the rest.
	// splat: only the +z neighbour row matters (mesher.c:391-432); mesh AO looks both ways (mesher.c:119-171).
	{
		const uint32_t per_row = (uint32_t)c->nx * c->ny;
		auto needs_ghost = [&](uint32_t id, bool mesh) {
			const int cz = (int)(id / per_row);
			if (cz + 1 == c->cfg.slab_z1 && c->cfg.slab_z1 < c->nz) return true;
			return mesh && cz == c->cfg.slab_z0 && c->cfg.slab_z0 > 0;
		};
		auto interior_first = [&](std::vector<uint32_t> &ids_, std::vector<uint32_t> &pos_, bool mesh) -> uint32_t {
			std::vector<uint32_t> a, ap, b, bp;
			for (size_t k = 0; k < ids_.size(); k++) {
				if (needs_ghost(ids_[k], mesh)) { b.push_back(ids_[k]); bp.push_back(pos_[k]); }
				else { a.push_back(ids_[k]); ap.push_back(pos_[k]); }
			}
			const uint32_t n_int = (uint32_t)a.size();
			a.insert(a.end(), b.begin(), b.end()); ap.insert(ap.end(), bp.begin(), bp.end());
			ids_.swap(a); pos_.swap(ap);
			return n_int;
		};
		c->n_splat_int = interior_first(sid, spos, false);
		c->n_mesh_int = interior_first(mid, mpos, true);
	}
	c->batch_n = n; c->n_splat = (uint32_t)sid.size(); c->n_mesh = (uint32_t)mid.size();
	c->batch_epoch = c->residency_epoch;
	if ((rc = splat_scratch_reserve(c, c->n_splat_int))) return rc;
	if ((rc = mesh_scratch_reserve(c, c->n_mesh_int))) return rc;
	if (c->n_splat > c->n_splat_int && (rc = splat_scratch_b_reserve(c, c->n_splat - c->n_splat_int))) return rc;
	if (c->n_mesh > c->n_mesh_int && (rc = mesh_scratch_b_reserve(c, c->n_mesh - c->n_mesh_int))) return rc;
	if (c->n_splat) {
		VP_CUDA(c, cudaMemcpyAsync(c->d_splat_ids, sid.data(), sid.size() * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, cudaMemcpyAsync(c->d_splat_pos, spos.data(), spos.size() * 4, cudaMemcpyHostToDevice, c->stream));
	}
	if (c->n_mesh) {
		VP_CUDA(c, cudaMemcpyAsync(c->d_mesh_ids, mid.data(), mid.size() * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, cudaMemcpyAsync(c->d_mesh_pos, mpos.data(), mpos.size() * 4, cudaMemcpyHostToDevice, c->stream));
	}
	// reset template for the arena states
	for (int a = 0; a < 3; a++) { c->h_arena_state[3 + a].cursor = 0; c->h_arena_state[3 + a].overflow = 0; c->h_arena_state[3 + a].pad = 0; }
	c->h_arena_state[3].capacity = c->cfg.splat_arena_bytes;
	c->h_arena_state[4].capacity = c->cfg.mesh_arena_bytes;
	c->h_arena_state[5].capacity = c->cfg.rle_arena_bytes;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

// part 0: reset the arenas / records, launch the chunks that do not read a ghost row; part 1: the chunks that do (after
// the caller unpacked the received border planes on the context stream).  Within a part the mesh kernel (few chunks,
// latency bound) runs on its own stream beside the splat kernels: the two write disjoint fields of the result records
// and separate arenas.
extern "C" int vp_rebuild_device_part(vp_ctx *c, int part)
{
	if (!c || part < 0 || part > 1) return VP_ERR_ARG;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	VpWorldDev w = vp_world_dev(c);
	// the prepared lists leave out chunks that had nothing to show when they were made (null chunk with null +x,+y,+z
	// neighbours): a chunk that changed between null and resident since then needs a new vp_batch_prepare
	if (c->batch_epoch != c->residency_epoch)
		return vp_fail(c, VP_ERR_ARG, "vp_rebuild_device: chunk residency changed since vp_batch_prepare -- prepare the batch again");
	if (part == 0) {
		VP_CUDA(c, cudaMemcpyAsync(c->d_arena_state, c->h_arena_state + 3, 2 * sizeof(VpArenaDev), cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, cudaMemsetAsync(c->d_results, 0, (size_t)c->batch_n * sizeof(VpResultDev), c->stream));
		VP_CUDA(c, cudaEventRecord(c->ev_reset, c->stream));
		c->ev_k_valid[c->rebuilds % vp_ctx::kHist] = 0;
		c->rebuilds++;
	}
	if (!c->rebuilds) return vp_fail(c, VP_ERR_ARG, "vp_rebuild_device_part: part 1 before part 0");
	cudaEvent_t *ev = c->ev_k[(c->rebuilds - 1) % vp_ctx::kHist];
	uint8_t &valid = c->ev_k_valid[(c->rebuilds - 1) % vp_ctx::kHist];
	const bool fork = c->n_splat && c->n_mesh;
	cudaStream_t ms = fork ? c->mesh_stream : c->stream;
	if (part == 0) {
		// interior: the chunks that do not read a ghost row.  The mesh kernels (few chunks) run beside the splat kernels.
		if (fork) {
			VP_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
			VP_CUDA(c, cudaStreamWaitEvent(ms, c->ev_fork, 0));
		}
		if (c->n_mesh) {
			VP_CUDA(c, cudaEventRecord(ev[2], ms));
			if (c->n_mesh_int) {
				VP_CUDA(c, vp_launch_mesh(w, c->d_mesh_ids, c->n_mesh_int, c->d_results, c->d_mesh_pos, c->d_mesh_arena, c->d_arena_state + 1,
				                          c->d_mesh_scratch, c->mesh_scratch_chunks, ms));
				c->launches += kMeshLaunches;
			}
		}
		if (c->n_splat) {
			VP_CUDA(c, cudaEventRecord(ev[0], c->stream));
			if (c->n_splat_int) {
				VP_CUDA(c, vp_launch_splat(w, c->d_splat_ids, c->n_splat_int, c->d_results, c->d_splat_pos, c->d_splat_arena, c->d_arena_state + 0,
				                           c->d_splat_scratch, c->splat_scratch_chunks, c->stream));
				c->launches += kSplatLaunches;
			}
		}
		return VP_OK;
	}
	// part 1: the slab's border chunks on the border stream -- behind the unpack of the received planes (vp_halo_unpack
	// runs there too) and BESIDE the interior kernels, with their own scratch; both parts bump the same arena cursors.
	const uint32_t sb = c->n_splat - c->n_splat_int, mb = c->n_mesh - c->n_mesh_int;
	if (sb || mb) {
		cudaStream_t bs = c->border_stream;
		VP_CUDA(c, cudaStreamWaitEvent(bs, c->ev_reset, 0));            // arena cursors and result records are reset
		if (sb) {
			VP_CUDA(c, vp_launch_splat(w, c->d_splat_ids + c->n_splat_int, sb, c->d_results, c->d_splat_pos + c->n_splat_int, c->d_splat_arena,
			                           c->d_arena_state + 0, c->d_splat_scratch_b, c->splat_scratch_b_chunks, bs));
			c->launches += kSplatLaunches;
		}
		if (mb) {
			VP_CUDA(c, vp_launch_mesh(w, c->d_mesh_ids + c->n_mesh_int, mb, c->d_results, c->d_mesh_pos + c->n_mesh_int, c->d_mesh_arena,
			                          c->d_arena_state + 1, c->d_mesh_scratch_b, c->mesh_scratch_b_chunks, bs));
			c->launches += kMeshLaunches;
		}
		VP_CUDA(c, cudaEventRecord(c->ev_bjoin, bs));
		VP_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_bjoin, 0));
		if (fork) VP_CUDA(c, cudaStreamWaitEvent(ms, c->ev_bjoin, 0));
	}
	if (c->n_mesh) { VP_CUDA(c, cudaEventRecord(ev[3], ms)); valid |= 2; }
	if (c->n_splat) { VP_CUDA(c, cudaEventRecord(ev[1], c->stream)); valid |= 1; }
	if (fork) {
		VP_CUDA(c, cudaEventRecord(c->ev_join, ms));
		VP_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
	}
	return VP_OK;
}

extern "C" int vp_rebuild_device(vp_ctx *c)
{
	int rc = vp_rebuild_device_part(c, 0);
	return rc ? rc : vp_rebuild_device_part(c, 1);
}

extern "C" int vp_rebuild_device_results(vp_ctx *c, vp_chunk_result *results, uint64_t *splat_bytes, uint64_t *mesh_bytes)
{
	if (!c) return VP_ERR_ARG;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	if (c->batch_n) VP_CUDA(c, cudaMemcpyAsync(c->h_results, c->d_results, (size_t)c->batch_n * sizeof(VpResultDev), cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(c->h_arena_state, c->d_arena_state, 2 * sizeof(VpArenaDev), cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (results && c->batch_n) memcpy(results, c->h_results, (size_t)c->batch_n * sizeof(VpResultDev));
	if (splat_bytes) *splat_bytes = c->h_arena_state[0].cursor;
	if (mesh_bytes) *mesh_bytes = c->h_arena_state[1].cursor;
	if (c->h_arena_state[0].overflow || c->h_arena_state[1].overflow)
		return vp_fail(c, VP_ERR_ARENA_FULL, "output arena too small (cursor values give the required bytes)");
	return VP_OK;
}

extern "C" int vp_kernel_ms_history(vp_ctx *c, uint32_t n, float *splat_ms, float *mesh_ms)
{
	if (!c || n > (uint32_t)vp_ctx::kHist || n > c->rebuilds) return vp_fail(c, VP_ERR_ARG, "vp_kernel_ms_history: n exceeds the recorded history");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	for (uint32_t k = 0; k < n; k++) {
		const uint64_t step = c->rebuilds - n + k;
		cudaEvent_t *ev = c->ev_k[step % vp_ctx::kHist];
		const uint8_t valid = c->ev_k_valid[step % vp_ctx::kHist];
		float a = 0.0f, b = 0.0f;
		if (valid & 1) VP_CUDA(c, cudaEventElapsedTime(&a, ev[0], ev[1]));
		if (valid & 2) VP_CUDA(c, cudaEventElapsedTime(&b, ev[2], ev[3]));
		if (splat_ms) splat_ms[k] = a;
		if (mesh_ms) mesh_ms[k] = b;
	}
	return VP_OK;
}

extern "C" void *vp_splat_arena_device(vp_ctx *c) { return c ? c->d_splat_arena : nullptr; }
extern "C" void *vp_mesh_arena_device(vp_ctx *c) { return c ? c->d_mesh_arena : nullptr; }

extern "C" int vp_arena_download(vp_ctx *c, int which, void *host_dst, uint64_t bytes)
{
	if (!c || !host_dst || which < 0 || which > 1) return vp_fail(c, VP_ERR_ARG, "vp_arena_download: bad argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	VP_CUDA(c, cudaMemcpyAsync(host_dst, which ? c->d_mesh_arena : c->d_splat_arena, bytes, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

extern "C" int vp_ctx_resize_arenas(vp_ctx *c, uint64_t splat_bytes, uint64_t mesh_bytes)
{
	if (!c) return VP_ERR_ARG;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (splat_bytes && splat_bytes != c->cfg.splat_arena_bytes) {
		cudaFree(c->d_splat_arena); c->d_splat_arena = nullptr;
		VP_CUDA(c, cudaMalloc(&c->d_splat_arena, splat_bytes));
		c->cfg.splat_arena_bytes = splat_bytes;
	}
	if (mesh_bytes && mesh_bytes != c->cfg.mesh_arena_bytes) {
		cudaFree(c->d_mesh_arena); c->d_mesh_arena = nullptr;
		VP_CUDA(c, cudaMalloc(&c->d_mesh_arena, mesh_bytes));
		c->cfg.mesh_arena_bytes = mesh_bytes;
	}
	return VP_OK;
}

// The first sb / mb bytes of the splat / mesh arenas into the context's pinned staging (grown on demand).
int vp_stage_arenas(vp_ctx *c, uint64_t sb, uint64_t mb, const void **splat_base, const void **mesh_base)
{
	int rc;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	if ((rc = stage_reserve(c, &c->h_splat_stage, &c->splat_stage_cap, sb))) return rc;
	if ((rc = stage_reserve(c, &c->h_mesh_stage, &c->mesh_stage_cap, mb))) return rc;
	if (sb) VP_CUDA(c, cudaMemcpyAsync(c->h_splat_stage, c->d_splat_arena, sb, cudaMemcpyDeviceToHost, c->stream));
	if (mb) VP_CUDA(c, cudaMemcpyAsync(c->h_mesh_stage, c->d_mesh_arena, mb, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (splat_base) *splat_base = c->h_splat_stage;
	if (mesh_base) *mesh_base = c->h_mesh_stage;
	return VP_OK;
}

extern "C" int vp_rebuild_batch(vp_ctx *c, const uint32_t *ids, uint32_t n, uint32_t flags, const uint8_t *per_chunk_flags,
                                vp_chunk_result *results, const void **splat_base, const void **mesh_base)
{
	int rc = vp_batch_prepare(c, ids, n, per_chunk_flags, flags);
	if (rc) return rc;
	rc = vp_rebuild_device(c);
	if (rc) return rc;
	uint64_t sb = 0, mb = 0;
	rc = vp_rebuild_device_results(c, results, &sb, &mb);
	if (rc) return rc;
	return vp_stage_arenas(c, sb, mb, splat_base, mesh_base);
}

// ---- single-chunk wrappers (mesher.h:7-37 semantics) ------------------------------------------------

extern "C" int64_t vp_chunk_make_splatlists(vp_ctx *c, uint32_t chunk_id, int16_t *geometry, uint64_t cap_items, uint32_t items[VP_MAX_LOD_LEVEL])
{
	vp_chunk_result r;
	const void *sb = nullptr;
	int rc = vp_rebuild_batch(c, &chunk_id, 1, VP_REBUILD_SPLAT, nullptr, &r, &sb, nullptr);
	if (rc) return rc;
	for (int l = 0; l < VP_MAX_LOD_LEVEL; l++) items[l] = r.svl_items[l];
	if (r.svl_items_total > cap_items) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_chunk_make_splatlists: geometry buffer too small");
	if (r.svl_items_total) memcpy(geometry, (const uint8_t *)sb + r.svl_offset, (size_t)r.svl_items_total * 2);
	return (int64_t)r.svl_items_total;
}

extern "C" int vp_chunk_make_mesh(vp_ctx *c, uint32_t chunk_id, int16_t *geometry, uint64_t cap_geometry_items, uint32_t *geometry_items,
                                  uint32_t *index, uint64_t cap_index_items, uint32_t *index_items)
{
	vp_chunk_result r;
	const void *mb = nullptr;
	int rc = vp_rebuild_batch(c, &chunk_id, 1, VP_REBUILD_MESH, nullptr, &r, nullptr, &mb);
	if (rc) return rc;
	*geometry_items = r.vbo_items; *index_items = r.ibo_items;
	if (r.vbo_items > cap_geometry_items || r.ibo_items > cap_index_items)
		return vp_fail(c, VP_ERR_ARENA_FULL, "vp_chunk_make_mesh: output buffer too small");
	if (r.vbo_items) memcpy(geometry, (const uint8_t *)mb + r.vbo_offset, (size_t)r.vbo_items * 2);
	if (r.ibo_items) memcpy(index, (const uint8_t *)mb + r.ibo_offset, (size_t)r.ibo_items * 4);
	return VP_OK;
}

// ---- multi-GPU slab borders ---------------------------------------------------------------------------

// Give every chunk of ghost row `row` a zero-filled slot (once).  Ghost chunks only ever hold the border slice.
static int ghost_row_init(vp_ctx *c, int row)
{
	const uint32_t per_row = (uint32_t)c->nx * c->ny;
	const size_t N = (size_t)1 << (3 * c->rb);
	bool all = true;
	for (uint32_t i = 0; i < per_row && all; i++) all = c->h_slot[(size_t)(row - c->ez0) * per_row + i] >= 0;
	if (all) return VP_OK;
	std::vector<uint32_t> ids(per_row);
	std::vector<uint8_t> want(per_row, 1);
	for (uint32_t i = 0; i < per_row; i++) ids[i] = (uint32_t)row * per_row + i;
	std::vector<int32_t> slots;
	int rc = assign_slots(c, ids.data(), per_row, want.data(), slots);
	if (rc) return rc;
	for (uint32_t i = 0; i < per_row; i++) VP_CUDA(c, cudaMemsetAsync(c->vox_pool + (size_t)slots[i] * N, 0, N, c->stream));
	if ((rc = push_slot_table(c, ids.data(), per_row))) return rc;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

extern "C" uint64_t vp_halo_plane_bytes(vp_ctx *c) { return c ? (uint64_t)c->nx * c->ny << (2 * c->rb) : 0; }

extern "C" int vp_halo_pack(vp_ctx *c, int which, void *device_buf)
{
	if (!c || !device_buf || which < 0 || which > 1) return vp_fail(c, VP_ERR_ARG, "vp_halo_pack: bad argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const uint32_t per_row = (uint32_t)c->nx * c->ny;
	const int row = which == 0 ? c->cfg.slab_z0 : c->cfg.slab_z1 - 1;
	const int32_t *row_slots = c->d_slot + (size_t)(row - c->ez0) * per_row;
	// on the border stream, behind whatever the context stream has done to the voxels so far
	VP_CUDA(c, cudaEventRecord(c->ev_bready, c->stream));
	VP_CUDA(c, cudaStreamWaitEvent(c->border_stream, c->ev_bready, 0));
	k_plane_copy<<<per_row, 256, 0, c->border_stream>>>(c->rb, c->vox_pool, row_slots, (uint8_t *)device_buf, which == 0 ? 0 : c->R - 1, 0);
	VP_CUDA(c, cudaGetLastError());
	c->launches++;
	return VP_OK;
}

extern "C" int vp_halo_unpack(vp_ctx *c, int which, const void *device_buf)
{
	if (!c || !device_buf || which < 0 || which > 1) return vp_fail(c, VP_ERR_ARG, "vp_halo_unpack: bad argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const uint32_t per_row = (uint32_t)c->nx * c->ny;
	const int row = which == 0 ? c->cfg.slab_z1 : c->cfg.slab_z0 - 1;
	if (row < c->ez0 || row >= c->ez1) return vp_fail(c, VP_ERR_ARG, "vp_halo_unpack: no ghost row on that side");
	// Ghost rows own permanent, zero-filled slots (allocated on first use), so an exchange is two kernels and
	// no host round trip: copy the plane into the border slice, refresh the x-face planes.
	int rc = ghost_row_init(c, row);
	if (rc) return rc;
	const int32_t *row_slots = c->d_slot + (size_t)(row - c->ez0) * per_row;
	k_plane_copy<<<per_row, 256, 0, c->border_stream>>>(c->rb, c->vox_pool, row_slots, (uint8_t *)const_cast<void *>(device_buf), which == 0 ? 0 : c->R - 1, 1);
	VP_CUDA(c, cudaGetLastError());
	VP_CUDA(c, vp_launch_extract_xfaces(c->rb, c->vox_pool, c->xlo_pool, c->xhi_pool, row_slots, per_row, c->border_stream));
	c->launches += 2;
	// anything the caller enqueues on the context stream next (e.g. vp_rebuild_batch of border chunks) sees the planes
	VP_CUDA(c, cudaEventRecord(c->ev_bjoin, c->border_stream));
	VP_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_bjoin, 0));
	return VP_OK;
}

extern "C" void *vp_ctx_border_stream(vp_ctx *c) { return c ? (void *)c->border_stream : nullptr; }

// ------------------------------------------------------------------------------------------------
// RLE: upload + decode, encode of resident chunks, flat codec
// ------------------------------------------------------------------------------------------------

static int dio_reserve(vp_ctx *c, size_t need)
{
	if (need <= c->d_io_cap) return VP_OK;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(c->d_io); c->d_io = nullptr; c->d_io_cap = 0;
	size_t want = std::max(need, (size_t)4 << 20);
	VP_CUDA(c, cudaMalloc(&c->d_io, want));
	c->d_io_cap = want;
	return VP_OK;
}

extern "C" int vp_upload_chunks_rle(vp_ctx *c, const uint32_t *ids, uint32_t n, const uint32_t *words, const uint64_t *word_offsets)
{
	if (!c || (n && (!ids || !words || !word_offsets))) return vp_fail(c, VP_ERR_ARG, "vp_upload_chunks_rle: null argument");
	if (!n) return VP_OK;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const uint32_t N = 1u << (3 * c->rb);
	const uint64_t total_words = word_offsets[n] - word_offsets[0];
	if (total_words * 4 > c->cfg.rle_arena_bytes) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_upload_chunks_rle: rle arena too small for this batch");
	std::vector<uint8_t> want(n);
	for (uint32_t i = 0; i < n; i++) {
		if (word_offsets[i + 1] < word_offsets[i] + 2) return vp_fail(c, VP_ERR_RLE, "vp_upload_chunks_rle: stream shorter than 2 words");
		want[i] = words[word_offsets[i]] != N;             // first word == {run N, value 0}: the null chunk (chunkset.c:225-228)
	}
	std::vector<int32_t> slots;
	int rc = assign_slots(c, ids, n, want.data(), slots);
	if (rc) return rc;
	// device copies: words -> rle arena, offsets (relative to the first word) + slots + status -> d_io
	std::vector<unsigned long long> rel(n + 1);
	for (uint32_t i = 0; i <= n; i++) rel[i] = word_offsets[i] - word_offsets[0];
	const size_t off_bytes = (size_t)(n + 1) * 8, slot_bytes = (size_t)n * 4;
	if ((rc = dio_reserve(c, off_bytes + slot_bytes + 16))) return rc;
	unsigned long long *d_off = reinterpret_cast<unsigned long long *>(c->d_io);
	int32_t *d_slots = reinterpret_cast<int32_t *>(c->d_io + off_bytes);
	uint32_t *d_status = reinterpret_cast<uint32_t *>(c->d_io + off_bytes + ((slot_bytes + 7) & ~(size_t)7));
	VP_CUDA(c, cudaMemcpyAsync(c->d_rle_arena, words + word_offsets[0], total_words * 4, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_off, rel.data(), off_bytes, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_slots, slots.data(), slot_bytes, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemsetAsync(d_status, 0, 4, c->stream));
	VP_CUDA(c, vp_launch_rle_decode(reinterpret_cast<const uint32_t *>(c->d_rle_arena), d_off, d_slots, n, c->vox_pool, N, d_status, c->stream));
	VP_CUDA(c, vp_launch_extract_xfaces(c->rb, c->vox_pool, c->xlo_pool, c->xhi_pool, d_slots, n, c->stream));
	c->launches += 2;
	if ((rc = push_slot_table(c, ids, n))) return rc;
	uint32_t status = 0;
	VP_CUDA(c, cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (status) return vp_fail(c, VP_ERR_RLE, "vp_upload_chunks_rle: a stream does not expand to exactly one chunk volume");
	return VP_OK;
}

extern "C" int vp_encode_chunks_rle(vp_ctx *c, const uint32_t *ids, uint32_t n, uint32_t *words, uint64_t cap_words, uint64_t *word_offsets)
{
	if (!c || (n && (!ids || !word_offsets))) return vp_fail(c, VP_ERR_ARG, "vp_encode_chunks_rle: null argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const uint32_t N = 1u << (3 * c->rb);
	std::vector<int32_t> slots(n);
	for (uint32_t i = 0; i < n; i++) {
		int64_t e = ext_index(c, ids[i]);
		if (e < 0) return vp_fail(c, VP_ERR_NOT_RESIDENT, "chunk id outside this context's slab");
		slots[i] = c->h_slot[(size_t)e];
	}
	const size_t off_bytes = (size_t)n * 8, cnt_bytes = (size_t)n * 4, slot_bytes = (size_t)n * 4;
	int rc = dio_reserve(c, off_bytes + cnt_bytes + slot_bytes + 16);
	if (rc) return rc;
	unsigned long long *d_off = reinterpret_cast<unsigned long long *>(c->d_io);
	uint32_t *d_cnt = reinterpret_cast<uint32_t *>(c->d_io + off_bytes);
	int32_t *d_slots = reinterpret_cast<int32_t *>(c->d_io + off_bytes + cnt_bytes);
	c->h_arena_state[5].cursor = 0; c->h_arena_state[5].capacity = c->cfg.rle_arena_bytes; c->h_arena_state[5].overflow = 0; c->h_arena_state[5].pad = 0;
	VP_CUDA(c, cudaMemcpyAsync(c->d_arena_state + 2, c->h_arena_state + 5, sizeof(VpArenaDev), cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_slots, slots.data(), slot_bytes, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, vp_launch_rle_encode(c->vox_pool, d_slots, n, N, reinterpret_cast<uint32_t *>(c->d_rle_arena), c->d_arena_state + 2, d_off, d_cnt, c->stream));
	c->launches++;
	std::vector<unsigned long long> off(n);
	std::vector<uint32_t> cnt(n);
	VP_CUDA(c, cudaMemcpyAsync(off.data(), d_off, off_bytes, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(cnt.data(), d_cnt, cnt_bytes, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(c->h_arena_state + 2, c->d_arena_state + 2, sizeof(VpArenaDev), cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->h_arena_state[2].overflow) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_encode_chunks_rle: rle arena too small");
	uint64_t acc = 0;
	for (uint32_t i = 0; i < n; i++) { word_offsets[i] = acc; acc += slots[i] < 0 ? 2 : cnt[i]; }
	word_offsets[n] = acc;
	if (acc > cap_words || !words) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_encode_chunks_rle: output buffer too small (word_offsets[n] words needed)");
	const uint64_t used = c->h_arena_state[2].cursor;
	if ((rc = stage_reserve(c, &c->h_io_stage, &c->io_stage_cap, used))) return rc;
	if (used) VP_CUDA(c, cudaMemcpyAsync(c->h_io_stage, c->d_rle_arena, used, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	for (uint32_t i = 0; i < n; i++) {
		uint32_t *dst = words + word_offsets[i];
		if (slots[i] < 0) { dst[0] = N; dst[1] = 0; }                    // the shared null stream (chunkset.c:98,117)
		else memcpy(dst, c->h_io_stage + off[i] * 4, (size_t)cnt[i] * 4);
	}
	return VP_OK;
}

extern "C" int vp_rle_compress(vp_ctx *c, const uint8_t *data, uint32_t length, uint32_t *out_words, uint32_t cap_words, uint32_t *n_words)
{
	if (!c || !data || !n_words || length == 0) return vp_fail(c, VP_ERR_ARG, "vp_rle_compress: null argument or empty input");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	// rle.h:7 takes any length.  The kernel works on 16-byte groups and packs a run into 24 bits, so the input is padded to a
	// multiple of 16 with copies of its last byte (they extend the last run; taken off again below) and encoded in segments
	// of at most 0xFFFFF0 bytes (no run of a segment reaches 24 bits); the host then joins runs across segment borders and
	// applies the reference's split rule: a run stops when its count reaches 0xFFFFFF (rle.c:62).
	constexpr uint32_t kSeg = 0xFFFFF0u;
	const uint64_t padded = ((uint64_t)length + 15u) & ~15ull;
	const uint32_t pad = (uint32_t)(padded - length), nseg = (uint32_t)((padded + kSeg - 1) / kSeg);
	if ((padded + 4ull * nseg) * 4 > c->cfg.rle_arena_bytes) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rle_compress: rle arena too small");
	int rc = dio_reserve(c, (size_t)padded + 64 + (size_t)nseg * 16);
	if (rc) return rc;
	unsigned long long *d_off = reinterpret_cast<unsigned long long *>(c->d_io);
	uint32_t *d_cnt = reinterpret_cast<uint32_t *>(c->d_io + (size_t)nseg * 8);
	uint8_t *d_src = c->d_io + (((size_t)nseg * 12 + 63) & ~(size_t)63);
	c->h_arena_state[5].cursor = 0; c->h_arena_state[5].capacity = c->cfg.rle_arena_bytes; c->h_arena_state[5].overflow = 0; c->h_arena_state[5].pad = 0;
	VP_CUDA(c, cudaMemcpyAsync(c->d_arena_state + 2, c->h_arena_state + 5, sizeof(VpArenaDev), cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_src, data, length, cudaMemcpyHostToDevice, c->stream));
	if (pad) VP_CUDA(c, cudaMemsetAsync(d_src + length, data[length - 1], pad, c->stream));
	for (uint32_t k = 0; k < nseg; k++) {
		const uint32_t seg_len = (uint32_t)std::min<uint64_t>(kSeg, padded - (uint64_t)k * kSeg);
		VP_CUDA(c, vp_launch_rle_encode(d_src + (size_t)k * kSeg, nullptr, 1, seg_len, reinterpret_cast<uint32_t *>(c->d_rle_arena), c->d_arena_state + 2,
		                                d_off + k, d_cnt + k, c->stream));
		c->launches++;
	}
	std::vector<unsigned long long> off(nseg);
	std::vector<uint32_t> cnt(nseg);
	VP_CUDA(c, cudaMemcpyAsync(off.data(), d_off, (size_t)nseg * 8, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)nseg * 4, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	uint64_t raw = 0;
	for (uint32_t k = 0; k < nseg; k++) { if (off[k] == ~0ull) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rle_compress: rle arena too small"); raw += cnt[k]; }
	if (nseg == 1 && pad == 0) {               // the common case (a chunk volume): the device stream is the answer
		*n_words = cnt[0];
		if (cnt[0] > cap_words || !out_words) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rle_compress: output buffer too small");
		VP_CUDA(c, cudaMemcpyAsync(out_words, reinterpret_cast<uint32_t *>(c->d_rle_arena) + off[0], (size_t)cnt[0] * 4, cudaMemcpyDeviceToHost, c->stream));
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
		return VP_OK;
	}
	std::vector<uint32_t> seg(raw), out;
	uint64_t at = 0;
	for (uint32_t k = 0; k < nseg; k++) {
		VP_CUDA(c, cudaMemcpyAsync(seg.data() + at, reinterpret_cast<uint32_t *>(c->d_rle_arena) + off[k], (size_t)cnt[k] * 4, cudaMemcpyDeviceToHost, c->stream));
		at += cnt[k];
	}
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	uint64_t acc = 0; uint32_t val = 0;
	auto flush = [&] {
		while (acc > 0xFFFFFFull) { out.push_back(0xFFFFFFu | (val << 24)); acc -= 0xFFFFFFull; }
		if (acc) out.push_back((uint32_t)acc | (val << 24));
		acc = 0;
	};
	for (uint64_t i = 0; i < raw; i++) {
		const uint32_t wd = seg[i];
		if (!wd) continue;                                   // a segment's terminator
		if (acc && (wd >> 24) == val) acc += wd & 0xFFFFFFu;
		else { flush(); val = wd >> 24; acc = wd & 0xFFFFFFu; }
	}
	acc -= pad;                                              // the padding copies the last byte: it sits in the last run
	flush();
	out.push_back(0u);
	*n_words = (uint32_t)out.size();
	if (out.size() > cap_words || !out_words) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rle_compress: output buffer too small");
	memcpy(out_words, out.data(), out.size() * 4);
	return VP_OK;
}

extern "C" int vp_rle_decompress(vp_ctx *c, const uint32_t *words, uint32_t n_words, uint8_t *out, uint32_t cap_bytes, uint32_t *n_bytes)
{
	if (!c || !words || !out || !n_bytes || n_words < 2) return vp_fail(c, VP_ERR_ARG, "vp_rle_decompress: bad argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	uint64_t total = 0;
	for (uint32_t i = 0; i + 1 < n_words; i++) total += words[i] & 0xFFFFFFu;        // output size (metadata only)
	*n_bytes = (uint32_t)std::min<uint64_t>(total, 0xFFFFFFFFu);
	if (total > cap_bytes) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rle_decompress: output buffer too small");
	if (total == 0 || total > 0xFFFFFFF0ull) return vp_fail(c, VP_ERR_ARG, "vp_rle_decompress: decoded length must be 1 .. 2^32 - 16 bytes");
	// the kernel writes 16-byte groups: a stream of any other length gets one run of padding in front of its terminator
	const uint32_t pad = (uint32_t)((16u - (total & 15u)) & 15u);
	const uint32_t dev_words = n_words + (pad ? 1u : 0u);
	if ((size_t)dev_words * 4 > c->cfg.rle_arena_bytes) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rle_decompress: rle arena too small");
	int rc = dio_reserve(c, (size_t)total + pad + 64);
	if (rc) return rc;
	unsigned long long h_off[2] = {0, dev_words};
	const uint32_t tail[2] = {pad, 0u};
	unsigned long long *d_off = reinterpret_cast<unsigned long long *>(c->d_io);
	uint32_t *d_status = reinterpret_cast<uint32_t *>(c->d_io + 16);
	uint8_t *d_dst = c->d_io + 64;
	VP_CUDA(c, cudaMemcpyAsync(c->d_rle_arena, words, (size_t)(n_words - 1) * 4, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(c->d_rle_arena + (size_t)(n_words - 1) * 4, pad ? tail : tail + 1, pad ? 8 : 4, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_off, h_off, 16, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemsetAsync(d_status, 0, 4, c->stream));
	VP_CUDA(c, vp_launch_rle_decode(reinterpret_cast<const uint32_t *>(c->d_rle_arena), d_off, nullptr, 1, d_dst, (uint32_t)(total + pad), d_status, c->stream));
	c->launches++;
	uint32_t status = 0;
	VP_CUDA(c, cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(out, d_dst, total, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (status) return vp_fail(c, VP_ERR_RLE, "vp_rle_decompress: malformed stream");
	return VP_OK;
}

// ------------------------------------------------------------------------------------------------
// Pipelined end-to-end rebuild: host RLE streams in, host buffers out, in ONE call.
// Same result as vp_upload_chunks_rle + vp_rebuild_batch, but the batch is cut into blocks of ascending ids
// that flow through three streams -- upload+decode | kernels | download -- so that the PCIe transfers of
// different blocks run in both directions at once and hide the kernels.  Blocks are processed from the
// HIGHEST ids down: the cull of a chunk reads its +x/+y/+z neighbours (larger ids), which are then already
// decoded; the mesh of a chunk also reads neighbours with smaller ids, so it is launched in the step in
// which the lowest of them (id - nx*ny - nx - 1) has been decoded.
// ------------------------------------------------------------------------------------------------
extern "C" int vp_rebuild_from_rle(vp_ctx *c, const uint32_t *ids, uint32_t n, const uint32_t *words, const uint64_t *word_offsets,
                                   uint32_t flags, const uint8_t *per_chunk_flags, uint32_t n_blocks,
                                   vp_chunk_result *results, const void **splat_base, const void **mesh_base)
{
	const double t_entry = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	const bool trace = getenv("VP_TRACE") != nullptr;
	auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	auto mark = [&](const char *what) { if (trace) fprintf(stderr, "[vp_rebuild_from_rle] %-28s at %.3f ms\n", what, now_ms() - t_entry); };
	if (!c || !n || !ids || !words || !word_offsets || !results) return vp_fail(c, VP_ERR_ARG, "vp_rebuild_from_rle: null argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	bool ascending = true;
	for (uint32_t i = 1; i < n && ascending; i++) ascending = ids[i] > ids[i - 1];
	if (!ascending || n_blocks < 1) n_blocks = 1;                    // pipelining needs ascending ids
	n_blocks = std::min<uint32_t>(std::min<uint32_t>(n_blocks, 64u), n);
	const uint32_t N = 1u << (3 * c->rb), per_row = (uint32_t)c->nx * c->ny;
	const uint64_t total_words = word_offsets[n] - word_offsets[0];
	if (total_words * 4 > c->cfg.rle_arena_bytes) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_rebuild_from_rle: rle arena too small");
	int rc = batch_reserve(c, n);
	if (rc) return rc;

	// ---- host bookkeeping for the whole batch ----
	std::vector<uint8_t> want(n);
	for (uint32_t i = 0; i < n; i++) {
		if (word_offsets[i + 1] < word_offsets[i] + 2) return vp_fail(c, VP_ERR_RLE, "vp_rebuild_from_rle: stream shorter than 2 words");
		want[i] = words[word_offsets[i]] != N;
		const uint32_t cz = ids[i] / per_row;
		if (ids[i] >= per_row * (uint32_t)c->nz || (int)cz < c->cfg.slab_z0 || (int)cz >= c->cfg.slab_z1)
			return vp_fail(c, VP_ERR_NOT_RESIDENT, "vp_rebuild_from_rle: chunk id outside the owned slab");
	}
	std::vector<int32_t> slots;
	mark("checked");
	if ((rc = assign_slots(c, ids, n, want.data(), slots))) return rc;
	mark("slots assigned");
	if ((rc = push_slot_table(c, ids, n))) return rc;
	mark("slot table pushed");
	std::vector<unsigned long long> rel(n + 1);
	for (uint32_t i = 0; i <= n; i++) rel[i] = word_offsets[i] - word_offsets[0];
	const size_t off_bytes = (size_t)(n + 1) * 8, slot_bytes = (size_t)n * 4;
	if ((rc = dio_reserve(c, off_bytes + slot_bytes + 16))) return rc;
	unsigned long long *d_off = reinterpret_cast<unsigned long long *>(c->d_io);
	int32_t *d_slots = reinterpret_cast<int32_t *>(c->d_io + off_bytes);
	uint32_t *d_status = reinterpret_cast<uint32_t *>(c->d_io + off_bytes + ((slot_bytes + 7) & ~(size_t)7));
	VP_CUDA(c, cudaMemcpyAsync(d_off, rel.data(), off_bytes, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_slots, slots.data(), slot_bytes, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemsetAsync(d_status, 0, 4, c->stream));

	// block b covers batch positions [bstart[b], bstart[b+1]); step t = 0.. processes block n_blocks-1-t
	std::vector<uint32_t> bstart(n_blocks + 1);
	for (uint32_t b = 0; b <= n_blocks; b++) bstart[b] = (uint32_t)((uint64_t)n * b / n_blocks);
	auto slot_of = [&](uint32_t x, uint32_t y, uint32_t z) -> int32_t {
		if (x >= (uint32_t)c->nx || y >= (uint32_t)c->ny || z >= (uint32_t)c->nz || (int)z < c->ez0 || (int)z >= c->ez1) return -1;
		return c->h_slot[(size_t)(z - (uint32_t)c->ez0) * per_row + (size_t)y * c->nx + x];
	};
	std::vector<std::vector<uint32_t>> s_ids(n_blocks), s_pos(n_blocks), m_ids(n_blocks), m_pos(n_blocks);
	for (uint32_t i = 0; i < n; i++) {
		const uint32_t f = per_chunk_flags ? per_chunk_flags[i] : flags;
		uint32_t b = (uint32_t)(std::upper_bound(bstart.begin(), bstart.end(), i) - bstart.begin()) - 1;
		const uint32_t step = n_blocks - 1 - b;
		if (f & VP_REBUILD_SPLAT) {
			const uint32_t cx = ids[i] % (uint32_t)c->nx, cy = (ids[i] / (uint32_t)c->nx) % (uint32_t)c->ny, cz = ids[i] / per_row;
			if (slot_of(cx, cy, cz) >= 0 || slot_of(cx + 1, cy, cz) >= 0 || slot_of(cx, cy + 1, cz) >= 0 || slot_of(cx, cy, cz + 1) >= 0) {
				s_ids[step].push_back(ids[i]); s_pos[step].push_back(i);
			}
		}
		if (f & VP_REBUILD_MESH) {
			// the lowest neighbour id; the mesh may run once the block that holds it (or the first block) is decoded
			const int64_t low = (int64_t)ids[i] - per_row - c->nx - 1;
			uint32_t mb = 0;
			if (low > (int64_t)ids[0]) {
				const uint32_t pos = (uint32_t)(std::lower_bound(ids, ids + n, (uint32_t)low) - ids);
				mb = (uint32_t)(std::upper_bound(bstart.begin(), bstart.end(), std::min(pos, n - 1)) - bstart.begin()) - 1;
			}
			const uint32_t mstep = n_blocks - 1 - std::min(mb, b);
			m_ids[mstep].push_back(ids[i]); m_pos[mstep].push_back(i);
		}
	}
	std::vector<uint32_t> sid, spos, mid, mpos, s_first(n_blocks + 1, 0), m_first(n_blocks + 1, 0);
	for (uint32_t t = 0; t < n_blocks; t++) {
		sid.insert(sid.end(), s_ids[t].begin(), s_ids[t].end()); spos.insert(spos.end(), s_pos[t].begin(), s_pos[t].end());
		mid.insert(mid.end(), m_ids[t].begin(), m_ids[t].end()); mpos.insert(mpos.end(), m_pos[t].begin(), m_pos[t].end());
		s_first[t + 1] = (uint32_t)sid.size(); m_first[t + 1] = (uint32_t)mid.size();
	}
	{
		uint32_t largest = 0, largest_m = 0;
		for (uint32_t t = 0; t < n_blocks; t++) { largest = std::max(largest, s_first[t + 1] - s_first[t]); largest_m = std::max(largest_m, m_first[t + 1] - m_first[t]); }
		if ((rc = splat_scratch_reserve(c, largest))) return rc;
		if ((rc = mesh_scratch_reserve(c, largest_m))) return rc;
	}
	if (!sid.empty()) {
		VP_CUDA(c, cudaMemcpyAsync(c->d_splat_ids, sid.data(), sid.size() * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, cudaMemcpyAsync(c->d_splat_pos, spos.data(), spos.size() * 4, cudaMemcpyHostToDevice, c->stream));
	}
	if (!mid.empty()) {
		VP_CUDA(c, cudaMemcpyAsync(c->d_mesh_ids, mid.data(), mid.size() * 4, cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, cudaMemcpyAsync(c->d_mesh_pos, mpos.data(), mpos.size() * 4, cudaMemcpyHostToDevice, c->stream));
	}
	c->batch_n = n;
	for (int a = 0; a < 3; a++) { c->h_arena_state[3 + a].cursor = 0; c->h_arena_state[3 + a].overflow = 0; c->h_arena_state[3 + a].pad = 0; }
	c->h_arena_state[3].capacity = c->cfg.splat_arena_bytes;
	c->h_arena_state[4].capacity = c->cfg.mesh_arena_bytes;
	VP_CUDA(c, cudaMemcpyAsync(c->d_arena_state, c->h_arena_state + 3, 2 * sizeof(VpArenaDev), cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemsetAsync(c->d_results, 0, (size_t)n * sizeof(VpResultDev), c->stream));
	// staging sized from the previous call (grown afterwards if this batch turns out larger)
	if ((rc = stage_reserve(c, &c->h_splat_stage, &c->splat_stage_cap, std::max<uint64_t>(c->last_splat_bytes + c->last_splat_bytes / 4, 32u << 20)))) return rc;
	if ((rc = stage_reserve(c, &c->h_mesh_stage, &c->mesh_stage_cap, std::max<uint64_t>(c->last_mesh_bytes + c->last_mesh_bytes / 4, 32u << 20)))) return rc;
	VP_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
	VP_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_a, 0));

	// ---- download every step's new output range as soon as its kernels are done ----
	// Steps publish their arena cursors + a ticket in pinned memory (k_publish_step); drain() turns every finished step
	// into two asynchronous downloads on down_stream, either opportunistically (while later steps are still being
	// enqueued) or to the end (polling the tickets: the wake-up latency of an event wait would leave the copy engine idle
	// between steps).
	volatile uint32_t *tickets = reinterpret_cast<volatile uint32_t *>(c->h_steps + 64 * 2);
	const uint32_t ticket = ++c->pipe_ticket ? c->pipe_ticket : ++c->pipe_ticket;          // never 0
	uint64_t done_s = 0, done_m = 0;
	uint32_t next_dl = 0, enqueued_steps = 0;
	bool staged = true;
	auto drain = [&](bool to_the_end) -> int {
		while (next_dl < (to_the_end ? n_blocks : enqueued_steps)) {
			const uint32_t t = next_dl;
			if (tickets[t] != ticket) {
				if (!to_the_end) return VP_OK;
				const cudaError_t q = cudaEventQuery(c->ev_pipe[1][t]);                          // surfaces launch failures instead of spinning forever
				if (q != cudaSuccess && q != cudaErrorNotReady) return vp_fail(c, VP_ERR_CUDA, "vp_rebuild_from_rle: pipeline step", q);
				if (q == cudaSuccess && tickets[t] != ticket) { std::atomic_thread_fence(std::memory_order_seq_cst); }
				continue;
			}
			std::atomic_thread_fence(std::memory_order_acquire);
			next_dl++;
			if (trace) fprintf(stderr, "[vp_rebuild_from_rle] step %u kernels done at %.3f ms\n", t, now_ms() - t_entry);
			const uint64_t cs = c->h_steps[2 * t].cursor, cm = c->h_steps[2 * t + 1].cursor;
			if (c->h_steps[2 * t].overflow || c->h_steps[2 * t + 1].overflow) { staged = false; next_dl = n_blocks; return VP_OK; }
			if (cs > c->splat_stage_cap || cm > c->mesh_stage_cap) { staged = false; continue; }      // staging too small: bulk copy below
			if (!staged) continue;
			VP_CUDA(c, cudaStreamWaitEvent(c->down_stream, c->ev_pipe[1][t], 0));
			if (cs > done_s) VP_CUDA(c, cudaMemcpyAsync(c->h_splat_stage + done_s, c->d_splat_arena + done_s, cs - done_s, cudaMemcpyDeviceToHost, c->down_stream));
			if (cm > done_m) VP_CUDA(c, cudaMemcpyAsync(c->h_mesh_stage + done_m, c->d_mesh_arena + done_m, cm - done_m, cudaMemcpyDeviceToHost, c->down_stream));
			done_s = cs; done_m = cm;
		}
		return VP_OK;
	};

	// ---- enqueue: upload+decode on copy_stream, kernels on the main stream ----
	mark("lists built, first launch");
	VpWorldDev w = vp_world_dev(c);
	for (uint32_t t = 0; t < n_blocks; t++) {
		const uint32_t b = n_blocks - 1 - t, i0 = bstart[b], i1 = bstart[b + 1];
		const uint64_t w0 = rel[i0], w1 = rel[i1];
		VP_CUDA(c, cudaMemcpyAsync(c->d_rle_arena + w0 * 4, words + word_offsets[0] + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, c->copy_stream));
		VP_CUDA(c, vp_launch_rle_decode(reinterpret_cast<const uint32_t *>(c->d_rle_arena), d_off + i0, d_slots + i0, i1 - i0, c->vox_pool, N, d_status, c->copy_stream));
		VP_CUDA(c, vp_launch_extract_xfaces(c->rb, c->vox_pool, c->xlo_pool, c->xhi_pool, d_slots + i0, i1 - i0, c->copy_stream));
		VP_CUDA(c, cudaEventRecord(c->ev_pipe[0][t], c->copy_stream));
		c->launches += 2;
		VP_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_pipe[0][t], 0));
		enqueued_steps = t + 1;
		if (s_first[t + 1] > s_first[t]) {
			VP_CUDA(c, vp_launch_splat(w, c->d_splat_ids + s_first[t], s_first[t + 1] - s_first[t], c->d_results, c->d_splat_pos + s_first[t], c->d_splat_arena, c->d_arena_state + 0,
			                           c->d_splat_scratch, c->splat_scratch_chunks, c->stream));
			c->launches += kSplatLaunches;
		}
		if (m_first[t + 1] > m_first[t]) {
			VP_CUDA(c, vp_launch_mesh(w, c->d_mesh_ids + m_first[t], m_first[t + 1] - m_first[t], c->d_results, c->d_mesh_pos + m_first[t], c->d_mesh_arena, c->d_arena_state + 1,
			                          c->d_mesh_scratch, c->mesh_scratch_chunks, c->stream));
			c->launches += kMeshLaunches;
		}
		k_publish_step<<<1, 32, 0, c->stream>>>(c->d_arena_state, c->h_steps + 2 * t, tickets + t, ticket);
		VP_CUDA(c, cudaGetLastError());
		c->launches++;
		VP_CUDA(c, cudaEventRecord(c->ev_pipe[1][t], c->stream));
		if ((rc = drain(false))) return rc;           // start the downloads of finished steps while the later ones are enqueued
	}
	VP_CUDA(c, cudaMemcpyAsync(c->h_results, c->d_results, (size_t)n * sizeof(VpResultDev), cudaMemcpyDeviceToHost, c->stream));
	if (trace) fprintf(stderr, "[vp_rebuild_from_rle] enqueued after %.3f ms\n", now_ms() - t_entry);
	if ((rc = drain(true))) return rc;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->down_stream));
	if (trace) fprintf(stderr, "[vp_rebuild_from_rle] downloads done at %.3f ms\n", now_ms() - t_entry);
	uint32_t status = 0;
	VP_CUDA(c, cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
	VP_CUDA(c, cudaMemcpy(c->h_arena_state, c->d_arena_state, 2 * sizeof(VpArenaDev), cudaMemcpyDeviceToHost));
	if (status) return vp_fail(c, VP_ERR_RLE, "vp_rebuild_from_rle: a stream does not expand to exactly one chunk volume");
	if (c->h_arena_state[0].overflow || c->h_arena_state[1].overflow)
		return vp_fail(c, VP_ERR_ARENA_FULL, "output arena too small (cursor values give the required bytes)");
	const uint64_t sb = c->h_arena_state[0].cursor, mb = c->h_arena_state[1].cursor;
	c->last_splat_bytes = sb; c->last_mesh_bytes = mb;
	if (!staged) {           // first call / grown world: size the staging now and copy in bulk
		if ((rc = stage_reserve(c, &c->h_splat_stage, &c->splat_stage_cap, sb))) return rc;
		if ((rc = stage_reserve(c, &c->h_mesh_stage, &c->mesh_stage_cap, mb))) return rc;
		if (sb) VP_CUDA(c, cudaMemcpyAsync(c->h_splat_stage, c->d_splat_arena, sb, cudaMemcpyDeviceToHost, c->stream));
		if (mb) VP_CUDA(c, cudaMemcpyAsync(c->h_mesh_stage, c->d_mesh_arena, mb, cudaMemcpyDeviceToHost, c->stream));
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	memcpy(results, c->h_results, (size_t)n * sizeof(VpResultDev));
	if (splat_base) *splat_base = c->h_splat_stage;
	if (mesh_base) *mesh_base = c->h_mesh_stage;
	return VP_OK;
}

// ------------------------------------------------------------------------------------------------
// LOD-node aggregation (gfx_update_svl's gather, vsplat.c:209-323) over the results of the last rebuild
// ------------------------------------------------------------------------------------------------
extern "C" int vp_build_lod_nodes(vp_ctx *c, uint32_t lod, vp_node_result *nodes, uint32_t cap_nodes, uint32_t *n_nodes, const void **base, float *kernel_ms)
{
	if (!c || lod >= VP_MAX_LOD_LEVEL || !n_nodes) return vp_fail(c, VP_ERR_ARG, "vp_build_lod_nodes: bad argument");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	if (c->cfg.slab_z0 != 0 || c->cfg.slab_z1 != c->nz) return vp_fail(c, VP_ERR_ARG, "vp_build_lod_nodes: needs a context that owns the whole world");
	if (c->batch_n != (uint32_t)c->nx * c->ny * c->nz) return vp_fail(c, VP_ERR_ARG, "vp_build_lod_nodes: the last rebuild must have covered every chunk in id order");
	const int bits[3] = { c->cfg.max_bitw[0], c->cfg.max_bitw[1], c->cfg.max_bitw[2] };
	uint32_t nn = 1;
	for (int i = 0; i < 3; i++) nn <<= bits[i] - std::min<int>((int)lod, bits[i]);
	*n_nodes = nn;
	if (!nodes || cap_nodes < nn) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_build_lod_nodes: node table too small (*n_nodes entries needed)");
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (nn > c->nodes_cap) { cudaFree(c->d_nodes); c->d_nodes = nullptr; VP_CUDA(c, cudaMalloc(&c->d_nodes, (size_t)nn * sizeof(VpNodeDev))); c->nodes_cap = nn; }
	// the level's segments can never exceed what the rebuild wrote
	VP_CUDA(c, cudaMemcpy(c->h_arena_state, c->d_arena_state, sizeof(VpArenaDev), cudaMemcpyDeviceToHost));
	const size_t need = std::max<size_t>(c->h_arena_state[0].cursor, 1u << 20);
	if (need > c->node_arena_cap) { cudaFree(c->d_node_arena); c->d_node_arena = nullptr; VP_CUDA(c, cudaMalloc(&c->d_node_arena, need)); c->node_arena_cap = need; }
	c->h_arena_state[5].cursor = 0; c->h_arena_state[5].capacity = c->node_arena_cap; c->h_arena_state[5].overflow = 0; c->h_arena_state[5].pad = 0;
	VP_CUDA(c, cudaMemcpyAsync(c->d_arena_state + 2, c->h_arena_state + 5, sizeof(VpArenaDev), cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaEventRecord(c->ev_k[0][0], c->stream));
	{ int rc2 = dio_reserve(c, (size_t)c->batch_n * 8); if (rc2) return rc2; }
	VP_CUDA(c, vp_launch_lod_nodes((int)lod, bits, nn, c->d_results, c->d_splat_arena, c->d_node_arena, c->d_arena_state + 2, c->d_nodes,
	                               reinterpret_cast<unsigned long long *>(c->d_io), c->stream));
	VP_CUDA(c, cudaEventRecord(c->ev_k[0][1], c->stream));
	c->launches += 2;
	static_assert(sizeof(vp_node_result) == sizeof(VpNodeDev), "node layout must match the C ABI");
	VP_CUDA(c, cudaMemcpyAsync(nodes, c->d_nodes, (size_t)nn * sizeof(VpNodeDev), cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(c->h_arena_state + 2, c->d_arena_state + 2, sizeof(VpArenaDev), cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	if (kernel_ms) VP_CUDA(c, cudaEventElapsedTime(kernel_ms, c->ev_k[0][0], c->ev_k[0][1]));
	if (c->h_arena_state[2].overflow) return vp_fail(c, VP_ERR_ARENA_FULL, "vp_build_lod_nodes: node arena overflow");
	if (base) {
		const uint64_t used = c->h_arena_state[2].cursor;
		int rc = stage_reserve(c, &c->h_node_stage, &c->node_stage_cap, used);
		if (rc) return rc;
		if (used) VP_CUDA(c, cudaMemcpyAsync(c->h_node_stage, c->d_node_arena, used, cudaMemcpyDeviceToHost, c->stream));
		VP_CUDA(c, cudaStreamSynchronize(c->stream));
		*base = c->h_node_stage;
	}
	return VP_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-side brush edit (chunkset_edit_sphere, edit.c:179-244) and height-map read-back
// ------------------------------------------------------------------------------------------------
extern "C" int vp_edit_sphere(vp_ctx *c, int32_t x, int32_t y, int32_t z, uint32_t radius, uint8_t voxel,
                              uint32_t *dirty_ids, uint32_t cap, uint32_t *n_dirty)
{
	if (!c || !n_dirty || radius > 64) return vp_fail(c, VP_ERR_ARG, "vp_edit_sphere: bad argument (radius <= 64)");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const int rb = c->rb;
	const size_t N = (size_t)1 << (3 * rb);
	// chunks in the box (chunkset_get_chunks_in_aabb, edit.c:89-126: inclusive chunk range of [c-r-1, c+r+1], x,y,z loops)
	const int r = (int)radius;
	const int lo[3] = { (x - r - 1) >> rb, (y - r - 1) >> rb, (z - r - 1) >> rb }, hi[3] = { (x + r + 1) >> rb, (y + r + 1) >> rb, (z + r + 1) >> rb };
	std::vector<uint32_t> ids;
	for (int gx = lo[0]; gx <= hi[0]; gx++) for (int gy = lo[1]; gy <= hi[1]; gy++) for (int gz = lo[2]; gz <= hi[2]; gz++) {
		if (gx < 0 || gy < 0 || gz < 0 || gx >= c->nx || gy >= c->ny || gz >= c->nz) continue;
		ids.push_back(((uint32_t)gz * c->ny + (uint32_t)gy) * c->nx + (uint32_t)gx);
	}
	*n_dirty = (uint32_t)ids.size();
	if (ids.empty()) return VP_OK;                                  // edit.c:204-205
	if (dirty_ids && cap >= ids.size()) memcpy(dirty_ids, ids.data(), ids.size() * 4);
	// chunk_open_rw on a null chunk makes a fresh all-air copy (chunkset.c:174-180): give owned box chunks a zeroed slot
	std::vector<uint32_t> own;
	for (uint32_t id : ids) { const int cz = (int)(id / ((uint32_t)c->nx * c->ny)); if (cz >= c->cfg.slab_z0 && cz < c->cfg.slab_z1) own.push_back(id); }
	std::vector<int32_t> before(own.size()), slots;
	for (size_t i = 0; i < own.size(); i++) before[i] = c->h_slot[(size_t)ext_index(c, own[i])];
	std::vector<uint8_t> want(own.size(), 1);
	int rc = assign_slots(c, own.data(), (uint32_t)own.size(), want.data(), slots);
	if (rc) return rc;
	for (size_t i = 0; i < own.size(); i++) if (before[i] < 0) VP_CUDA(c, cudaMemsetAsync(c->vox_pool + (size_t)slots[i] * N, 0, N, c->stream));
	if ((rc = push_slot_table(c, own.data(), (uint32_t)own.size()))) return rc;
	VpWorldDev w = vp_world_dev(c);
	VP_CUDA(c, vp_launch_edit_sphere(w, c->vox_pool, c->d_shadow, x, y, z, r, voxel, c->cfg.slab_z0, c->cfg.slab_z1, c->stream));
	if (!slots.empty()) {                 // (a slab context may own none of the box's chunks: only its height-map rows change)
		VP_CUDA(c, cudaMemcpyAsync(c->d_tmp_slots, slots.data(), slots.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
		VP_CUDA(c, vp_launch_extract_xfaces(c->rb, c->vox_pool, c->xlo_pool, c->xhi_pool, c->d_tmp_slots, (uint32_t)slots.size(), c->stream));
		c->launches++;
	}
	c->launches++;
	VP_CUDA(c, cudaStreamSynchronize(c->stream));                   // `slots` goes out of scope
	return VP_OK;
}

extern "C" int vp_raycast(vp_ctx *c, uint32_t n, const float *origins, const float *vectors, uint32_t *coords, int8_t *normals, uint8_t *voxels)
{
	if (!c || (n && (!origins || !vectors || !coords || !normals || !voxels))) return vp_fail(c, VP_ERR_ARG, "vp_raycast: null argument");
	if (!n) return VP_OK;
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	// device scratch: origins | vectors | coords (n x 12 bytes each), normals (n x 3), voxels (n)
	const size_t f3 = (size_t)n * 12, total = 3 * f3 + (size_t)n * 4 + 64;
	int rc = dio_reserve(c, total);
	if (rc) return rc;
	float *d_o = reinterpret_cast<float *>(c->d_io), *d_v = reinterpret_cast<float *>(c->d_io + f3);
	uint32_t *d_c = reinterpret_cast<uint32_t *>(c->d_io + 2 * f3);
	int8_t *d_n = reinterpret_cast<int8_t *>(c->d_io + 3 * f3);
	uint8_t *d_x = reinterpret_cast<uint8_t *>(c->d_io + 3 * f3 + (size_t)n * 3);
	VP_CUDA(c, cudaMemcpyAsync(d_o, origins, f3, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_v, vectors, f3, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(d_n, normals, (size_t)n * 3, cudaMemcpyHostToDevice, c->stream));
	VP_CUDA(c, vp_launch_raycast(vp_world_dev(c), c->vox_pool, n, d_o, d_v, d_c, d_n, d_x, c->stream));
	c->launches++;
	VP_CUDA(c, cudaMemcpyAsync(coords, d_c, f3, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(normals, d_n, (size_t)n * 3, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaMemcpyAsync(voxels, d_x, n, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}

extern "C" int vp_download_shadow_rows(vp_ctx *c, uint32_t z0, uint32_t z1, uint16_t *rows)
{
	if (!c || !rows || z1 < z0 || z0 < c->sh_z0 || z1 > c->sh_z1) return vp_fail(c, VP_ERR_ARG, "vp_download_shadow_rows: rows outside the slab's range");
	VP_CUDA(c, cudaSetDevice(c->cfg.device));
	const uint32_t shw = (uint32_t)((c->nx + c->ny) << c->rb);
	if (z1 > z0) VP_CUDA(c, cudaMemcpyAsync(rows, c->d_shadow + (size_t)(z0 - c->sh_z0) * shw, (size_t)(z1 - z0) * shw * 2, cudaMemcpyDeviceToHost, c->stream));
	VP_CUDA(c, cudaStreamSynchronize(c->stream));
	return VP_OK;
}
