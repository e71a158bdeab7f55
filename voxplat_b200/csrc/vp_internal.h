// vp_internal.h -- shared declarations of the CUDA implementation behind include/voxplat_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/voxplat_b200.h"

// ---------------------------------------------------------------------------------------------
// Device view of the resident world.  HBM layout:
//   vox_pool  : slots * R^3 bytes, one slot per non-null chunk, chunk-local index (z<<2b|y<<b|x)
//   xlo_pool  : slots * R^2 bytes, copy of the x = 0   plane of each slot, [z][y]   (the +x halo source)
//   xhi_pool  : slots * R^2 bytes, copy of the x = R-1 plane of each slot, [z][y]   (mesh AO, -x halo)
//   slot      : int32 per chunk of the EXTENDED slab (owned rows plus one ghost row on each side),
//               -1 = null chunk (all air, chunkset.c:116-117)
//   shadow    : uint16 rows [sh_z0, sh_z1) of (X+Y) entries + zero padding (SURVEY 8a' u3)
// ---------------------------------------------------------------------------------------------
struct VpWorldDev {
	int32_t rb;                // root_bitw
	int32_t bits[3];           // chunk grid bit widths of the whole world
	int32_t ez0, ez1;          // chunk rows covered by `slot` (extended slab)
	const int32_t *slot;       // [(ez1-ez0) << (bits[0]+bits[1])]
	const uint8_t *vox_pool;
	const uint8_t *xlo_pool;
	const uint8_t *xhi_pool;
	const uint16_t *shadow;    // addresses row sh_z0
	uint32_t sh_z0, sh_z1;     // shadow rows held: [sh_z0, sh_z1)
	uint32_t sh_w;             // X + Y
};

// Result record written by the kernels (device mirror of vp_chunk_result).
struct VpResultDev {
	unsigned long long svl_offset;
	uint32_t svl_items[5];
	uint32_t svl_items_total;
	unsigned long long vbo_offset;
	unsigned long long ibo_offset;
	uint32_t vbo_items;
	uint32_t ibo_items;
};
static_assert(sizeof(VpResultDev) == sizeof(vp_chunk_result), "result layout must match the C ABI");

// Per-node record of the LOD aggregation (device mirror of vp_node_result).
struct VpNodeDev {
	unsigned long long offset;     // bytes into the node arena (~0 = not reserved)
	uint32_t items;                // int16 items = GeometrySVL.vbo_items
	uint32_t members;              // chunks under the node
};

// Arena bump allocator state in device memory.
struct VpArenaDev {
	unsigned long long cursor;     // bytes used
	unsigned long long capacity;   // bytes available
	unsigned int overflow;         // set to 1 by a kernel that could not reserve
	unsigned int pad;
};

struct vp_ctx {
	vp_config cfg{};
	int R, rb;
	int nx, ny, nz;               // chunk grid of the world
	int ez0, ez1;                 // extended slab rows held on this device
	uint32_t n_ext;               // chunks in the extended slab
	cudaStream_t own_stream, stream;
	cudaStream_t copy_stream, down_stream;
	cudaStream_t mesh_stream;     // the mesh kernel of a rebuild runs beside the splat kernels
	cudaStream_t border_stream;   // slab contexts: border pack / unpack and the rebuild of the chunks that read a ghost row
	cudaEvent_t ev_reset, ev_bjoin, ev_bready;
	cudaEvent_t ev_fork, ev_join;
	cudaEvent_t ev_pipe[2][64];   // [0] decode done, [1] kernels done, per pipeline step
	VpArenaDev *h_steps;          // pinned: arena states after every pipeline step [64][2], then 64 uint32 step tickets
	uint32_t pipe_ticket;         // ticket value of the running / last vp_rebuild_from_rle call
	uint64_t last_splat_bytes, last_mesh_bytes;
	cudaEvent_t ev_a, ev_b;
	static constexpr int kHist = 256;
	cudaEvent_t ev_k[kHist][4];   // timing events around the splat [0,1] and mesh [2,3] kernels of the last kHist rebuilds
	uint8_t ev_k_valid[kHist];    // bit0 splat, bit1 mesh recorded
	uint64_t rebuilds;            // vp_rebuild_device calls so far
	// world
	uint8_t *vox_pool, *xlo_pool, *xhi_pool;
	uint32_t n_slots;             // capacity of the pools in chunks
	std::vector<int32_t> h_slot;  // host mirror of slot table
	std::vector<uint32_t> free_slots;
	int32_t *d_slot;
	uint16_t *d_shadow; uint32_t sh_z0, sh_z1; size_t shadow_entries;
	// batch state
	uint32_t *d_ids; uint8_t *d_flags; uint32_t batch_n, batch_cap; uint32_t batch_flags;
	uint64_t residency_epoch, batch_epoch;               // bumped by every null <-> resident change / value at the last vp_batch_prepare
	uint32_t *d_splat_ids, *d_mesh_ids; uint32_t n_splat, n_mesh;
	uint32_t n_splat_int, n_mesh_int;                    // list entries that do not read a ghost row come first
	uint32_t *d_splat_pos, *d_mesh_pos;                  // position in the batch of each list entry
	int32_t *d_tmp_slots;                                // scratch list of slots for helper kernels
	VpResultDev *d_results; VpResultDev *h_results;      // h_results pinned
	// arenas
	uint8_t *d_splat_arena, *d_mesh_arena, *d_rle_arena;
	VpArenaDev *d_arena_state;    // [0] splat, [1] mesh, [2] rle / nodes
	VpArenaDev *h_arena_state;    // pinned: [0..2] read-back, [3..5] reset template
	uint8_t *h_splat_stage, *h_mesh_stage; size_t splat_stage_cap, mesh_stage_cap;
	uint8_t *h_io_stage; size_t io_stage_cap;            // pinned staging for uploads / rle
	uint8_t *d_node_arena; size_t node_arena_cap; VpNodeDev *d_nodes; uint32_t nodes_cap;   // LOD-node aggregation (vp_nodes.cu)
	uint8_t *h_node_stage; size_t node_stage_cap;
	uint8_t *d_splat_scratch; uint32_t splat_scratch_chunks;  // arrival counters + slab records of the splat kernel
	uint8_t *d_mesh_scratch; uint32_t mesh_scratch_chunks;    // occupancy tiles / counts between the mesh kernels
	uint8_t *d_splat_scratch_b; uint32_t splat_scratch_b_chunks;   // the same for the border part of a slab step (runs beside the interior part)
	uint8_t *d_mesh_scratch_b; uint32_t mesh_scratch_b_chunks;
	uint8_t *d_io; size_t d_io_cap;                      // device scratch for the flat RLE codec / stream offsets
	uint64_t launches;
	std::string err;
};

VpWorldDev vp_world_dev(const vp_ctx *c);
int vp_stage_arenas(vp_ctx *c, uint64_t sb, uint64_t mb, const void **splat_base, const void **mesh_base);
int vp_fail(vp_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess);
#define VP_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, #call, e__); } while (0)

// kernel launchers (each returns a cudaError_t from the launch)
// splat rebuild = 2 launches (count + reserve, emit); scratch was allocated (and zero-filled once) for launches of up to
// scratch_chunks chunks: vp_splat_scratch_bytes(rb, scratch_chunks) bytes
cudaError_t vp_launch_splat(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                            uint8_t *arena, VpArenaDev *state, uint8_t *scratch, uint32_t scratch_chunks, cudaStream_t s);
size_t vp_splat_scratch_bytes(int rb, uint32_t n);
constexpr int kSplatLaunches = 2;
// mesh rebuild = 2 launches (count + reserve, emit) with the occupancy tiles of the slabs carried through `scratch`
cudaError_t vp_launch_mesh(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                           uint8_t *arena, VpArenaDev *state, uint8_t *scratch, uint32_t scratch_chunks, cudaStream_t s);
size_t vp_mesh_scratch_bytes(int rb, uint32_t n);
constexpr int kMeshLaunches = 2;
cudaError_t vp_launch_extract_xfaces(int rb, const uint8_t *vox_pool, uint8_t *xlo_pool, uint8_t *xhi_pool,
                                     const int32_t *d_slots, uint32_t n, cudaStream_t s);
cudaError_t vp_launch_rle_decode(const uint32_t *d_words, const unsigned long long *d_offsets, const int32_t *d_slots,
                                 uint32_t n, uint8_t *dst_base, uint32_t N, uint32_t *d_status, cudaStream_t s);
cudaError_t vp_launch_rle_encode(const uint8_t *src_base, const int32_t *d_slots, uint32_t n, uint32_t N, uint32_t *d_arena_words,
                                 VpArenaDev *state, unsigned long long *d_offsets, uint32_t *d_counts, cudaStream_t s);
cudaError_t vp_launch_lod_nodes(int lod, const int bits[3], uint32_t n_nodes, const VpResultDev *d_chunk_res, const uint8_t *d_splat_arena,
                                uint8_t *d_node_arena, VpArenaDev *state, VpNodeDev *d_nodes, unsigned long long *d_chunk_dst, cudaStream_t s);
cudaError_t vp_launch_raycast(const VpWorldDev &w, const uint8_t *vox_pool, uint32_t n, const float *origins, const float *vectors,
                              uint32_t *coords, int8_t *normals, uint8_t *voxels, cudaStream_t s);
cudaError_t vp_launch_edit_sphere(const VpWorldDev &w, uint8_t *vox_pool, uint16_t *shadow, int cx, int cy, int cz, int radius, int voxel,
                                  int own_z0, int own_z1, cudaStream_t s);

// device world generator (vp_worldgen_dev.cu)
cudaError_t vp_launch_gen_chunks(uint32_t seed, int rb, const int bits[3], const uint32_t *d_ids, uint32_t n, uint8_t *d_out, uint32_t *d_solid, cudaStream_t s);
cudaError_t vp_launch_scatter_chunks(int rb, const uint8_t *staging, const int32_t *d_slots, uint32_t n, uint8_t *pool, cudaStream_t s);
cudaError_t vp_launch_shadow_rows(int rb, const uint8_t *const *table, int nx, int ny, uint32_t row0, uint32_t z0, uint32_t z1, uint16_t *rows_out, cudaStream_t s);
