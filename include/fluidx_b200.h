/* fluidx_b200.h — C ABI of the B200-native smoke-solver step (CSAdvect + CSProject2D/3D).
 *
 * This is the drop-in boundary for the simulation half of the reference's `Fluid` class
 * (FluidX12/Content/Fluid.h:20-33).  The reference has no FFI layer — `Fluid` is a C++ class on
 * D3D12/XUSG — so each entry point below names the reference member it replaces.  Signatures use
 * plain pointers and sizes only.  All functions return 0 on success or a negative fxb_status;
 * no exception crosses the boundary; fxb_last_error() describes the last failure on the calling
 * thread.  A handle is not re-entrant: drive it from one host thread (the reference runs
 * everything on its UI thread, Common/Win32Application.cpp:205-211).
 *
 * There is NO CPU fallback: fxb_create fails with FXB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FLUIDX_B200_H
#define FLUIDX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXB_ABI_VERSION 2

typedef struct fxb_sim fxb_sim;

typedef enum fxb_status {
    FXB_OK = 0,
    FXB_ERR_INVALID = -1,      /* bad argument / configuration (e.g. nx != ny, Fluid.cpp:201) */
    FXB_ERR_CUDA = -2,         /* CUDA runtime/driver failure, or no sm_100 device */
    FXB_ERR_NCCL = -3,         /* NCCL failure or NCCL not loadable when nranks > 1 */
    FXB_ERR_SIZE = -4,         /* host buffer size does not match the field */
    FXB_ERR_HALO_OVERFLOW = -5,/* a back-trace left the exchanged z-halo (multi-GPU only; sticky) */
    FXB_ERR_IO = -6            /* a volume file could not be written / read or is not a valid one */
} fxb_status;

/* Sampler addressing of the advection fetches: Fluid uses LINEAR_MIRROR (Fluid.cpp:452),
 * FluidEZ uses LINEAR_CLAMP (FluidEZ.cpp:406). */
typedef enum fxb_address_mode { FXB_ADDRESS_MIRROR = 0, FXB_ADDRESS_CLAMP = 1 } fxb_address_mode;

/* Fields addressable through fxb_get_field / fxb_set_field.  Host layout is the texture's logical
 * layout: dense, x fastest, [z][y][x][4] IEEE half for velocity/colour (R16G16B16A16_FLOAT,
 * Fluid.cpp:207,213) and [z][y][x] float for pressure (R32_FLOAT, Fluid.cpp:219).  With nranks > 1
 * a rank reads/writes only its own z-slab (fxb_get_slab). */
typedef enum fxb_field {
    FXB_FIELD_VELOCITY = 0,          /* m_velocities[0]: projected velocity (.w is a don't-care) */
    FXB_FIELD_COLOR = 1,             /* m_colors[m_frameParity]: what Fluid::Render samples */
    FXB_FIELD_PRESSURE = 2,          /* m_incompress: persists across frames (warm start) */
    FXB_FIELD_VELOCITY_ADVECTED = 3, /* m_velocities[1]: output of CSAdvect */
    FXB_FIELD_COLOR_PREV = 4         /* m_colors[!m_frameParity] */
} fxb_field;

/* How z-slab face halos travel between neighbouring ranks (nranks > 1). */
typedef enum fxb_halo_backend {
    FXB_HALO_PEER = 0,  /* one kernel per exchange stores the face planes straight into the neighbours' memory (CUDA IPC over NVLink) */
    FXB_HALO_NCCL = 1,  /* grouped ncclSend / ncclRecv pairs */
    FXB_HALO_FUSED = 2  /* no exchange at all: every kernel of the step stores the planes next to a slab face into the
                           neighbour's halo as well, ordered by one event counter per rank (default) */
} fxb_halo_backend;

typedef struct fxb_config {
    uint32_t struct_size;   /* = sizeof(fxb_config); set by fxb_config_default */
    uint32_t nx, ny, nz;    /* GLOBAL grid (gridSize of Fluid::Init, Fluid.cpp:189-201); nz == 1 selects the 2D path */
    int32_t address_mode;   /* fxb_address_mode; default MIRROR */
    int32_t early_exit;     /* 1 (default): per-cell `|x-x0| < 0.001` freeze of CSPoisson.hlsli:22-24; 0: all sweeps */
    int32_t jacobi_iters;   /* ITER, default 64 (CSProject3D.hlsl:13) */
    int32_t fuse_t;         /* Jacobi sweeps fused per HBM pass; 0 = library default */
    int32_t device;         /* CUDA device ordinal */
    int32_t rank, nranks;   /* z-slab decomposition: this rank owns planes [rank*nz/nranks, (rank+1)*nz/nranks) */
    int32_t h_adv;          /* advection z-halo in planes (multi-GPU); 0 = default 8 */
    int32_t use_graph;      /* 1 (default): the step is a captured CUDA graph; 0: plain stream launches */
    int32_t kernel_path;    /* 0 = tuned kernels (default); 1 = one simple kernel per logical pass (cross-check path) */
    int32_t phase_timing;   /* 1: a one-thread kernel after every phase accumulates the phase's device time (fxb_get_phase_times) */
    int32_t halo_backend;   /* fxb_halo_backend (multi-GPU); default FXB_HALO_FUSED */
    int32_t jacobi_group;   /* multi-GPU: fused passes per pressure-halo exchange; 0 = library default (1) */
    const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks; required iff nranks > 1 */
} fxb_config;

typedef struct fxb_stats {
    int32_t s_exec;            /* Jacobi sweeps in which >= 1 cell was still active, last step (global) */
    int32_t jacobi_passes;     /* fused HBM passes actually executed in the last step */
    int32_t fuse_t;            /* sweeps fused per pass (before pass tail_from) */
    int32_t halo_overflow;     /* sticky: 1 if an advection back-trace left the z-halo */
    int32_t frame_parity;      /* m_frameParity */
    int32_t kernels_per_step;  /* CUDA kernels launched by one fxb_simulate */
    uint64_t steps;            /* fxb_simulate calls so far */
    uint64_t active_after_first_sweep; /* cells still active after sweep 1, last step (this rank) */
    uint64_t total_sweeps;     /* cumulative s_exec over all steps */
    uint64_t total_passes;     /* cumulative jacobi_passes over all steps */
    uint64_t bricks_processed; /* cumulative brick passes: bricks relaxed by a fused pass (frozen bricks are skipped) */
    uint64_t bricks_copied;    /* cumulative frozen bricks copied once into the other pressure buffer */
    uint64_t brick_cells;      /* output cells per brick (tile minus halo, x bz planes) */
    uint64_t bricks_per_pass;  /* bricks in the grid of one pass */
    int32_t jacobi_fused;      /* 1 when the fused Jacobi kernel is in use, 0 for one sweep per launch */
    int32_t tail_from;         /* first pass of a frame that fuses FOUR sweeps (the tail schedule of the tuned T = 2 solve:
                                  passes 0 .. tail_from-1 fuse fuse_t sweeps each); 0 = every pass fuses fuse_t sweeps */
} fxb_stats;

/* Fills cfg with the defaults above (grid 128^3 as FluidX12.cpp:44). */
int fxb_config_default(fxb_config* cfg);

/* Replaces Fluid::Init (Fluid.cpp:189-270, simulation part): allocates 2x velocity + 2x colour +
 * pressure on the device, zero-initialised like freshly created D3D12 committed resources, builds
 * the emitter table and captures the step graph. */
int fxb_create(const fxb_config* cfg, fxb_sim** out);
void fxb_destroy(fxb_sim* sim);

/* Replaces Fluid::UpdateFrame's simulation part (Fluid.cpp:288-290, 344-345): stores the time
 * step for the next fxb_simulate and flips the frame parity iff dt > 0. */
int fxb_update_frame(fxb_sim* sim, float dt);

/* Replaces Fluid::Simulate (Fluid.cpp:348-410): enqueues advect + project on `cuda_stream`
 * (a cudaStream_t, NULL = legacy default stream) and returns without waiting. */
int fxb_simulate(fxb_sim* sim, void* cuda_stream);

/* Waits for everything enqueued by the handle and reports sticky asynchronous errors. */
int fxb_sync(fxb_sim* sim);

/* dt rule of FluidX::OnUpdate (FluidX12.cpp:266-267): (nz > 1 ? 2 : 1) / ny. */
int fxb_dt_for_grid(uint32_t nx, uint32_t ny, uint32_t nz, float* dt);

/* z-range [z0, z0+count) of this rank's slab (whole grid when nranks == 1). */
int fxb_get_slab(const fxb_sim* sim, uint32_t* z0, uint32_t* count);

/* Synchronous host<->device copies of one field of this rank's slab; `bytes` must equal
 * nx*ny*count*(8 or 4).  These double as checkpoint/restore (the reference has none). */
int fxb_get_field(fxb_sim* sim, int field, void* host, size_t bytes);
int fxb_set_field(fxb_sim* sim, int field, const void* host, size_t bytes);

/* Asynchronous variant used by the end-to-end path: enqueues the device->host copy of a field
 * into (preferably pinned) host memory on `cuda_stream`. */
int fxb_get_field_async(fxb_sim* sim, int field, void* host, size_t bytes, void* cuda_stream);

/* Reads back the device-side counters of the last step (synchronises the handle's stream). */
int fxb_get_stats(fxb_sim* sim, fxb_stats* out);

/* The same record without draining the pipeline: fxb_post_stats enqueues, behind everything the handle has enqueued so
 * far, a copy of the counters into pinned slot `slot` (0..3) and an event; fxb_wait_stats blocks until that copy has
 * landed and returns the record of the step that was last enqueued when it was posted.  A frame loop posts after every
 * fxb_simulate and waits for the PREVIOUS frame's slot, so the device never idles — the reference keeps FrameCount = 3
 * frames in flight the same way (Fluid.h:35, FluidX12.cpp:605-622). */
int fxb_post_stats(fxb_sim* sim, int slot);
int fxb_wait_stats(fxb_sim* sim, int slot, fxb_stats* out);

/* Diagnostic, needs no GPU: the plane arithmetic of the peer-memory halo exchange (FXB_HALO_PEER) for `rank` of `nranks`
 * z-slabs with `halo` allocated planes and an exchange `depth` planes deep.  out4 = {first local plane sent to rank-1,
 * first local plane of rank-1's array it lands in, first local plane sent to rank+1, first local plane of rank+1's
 * array it lands in} (entries for a missing neighbour are meaningless). */
int fxb_p2p_plan(int32_t nz, int32_t nranks, int32_t rank, int32_t halo, int32_t depth, int64_t* out4);

/* Diagnostic, needs no GPU: the pass schedule of the fused pressure solve for `iters` sweeps when a pass fuses `fuse_t`
 * sweeps and, from pass `tail_from` on (0 = never), four (fxb_stats.tail_from).  *npass = passes of a frame;
 * s0[k] (k < min(n, *npass)) = sweeps completed before pass k.  Replaces nothing in the reference: its loop runs the 64
 * sweeps one by one (CSPoisson.hlsli:8-26). */
int fxb_jacobi_schedule(int32_t iters, int32_t fuse_t, int32_t tail_from, int32_t* npass, int32_t* s0, int32_t n);

/* Diagnostic, needs no GPU: the order in which a fused-halo kernel hands out its z chunks (`chunk` planes each over `n`
 * owned planes; the chunks within `reach` planes of an interior face wait for the neighbour rank and come last).
 * out[b] (b < nchunks = ceil(n / chunk)) = the chunk CTA row b works on. */
int fxb_face_last_order(int32_t n, int32_t chunk, int32_t reach, int32_t has_lo, int32_t has_hi, int32_t* out, int32_t nchunks);

/* Diagnostic, needs no GPU: the voxel box {x0,y0,z0,x1,y1,z1} (half-open) outside of which the advection kernel
 * skips the emitter (CSAdvect.hlsl:57-68) because the Gaussian basis there is below exp(-4). */
int fxb_emitter_box(uint32_t nx, uint32_t ny, uint32_t nz, int32_t* out6);

/* Freeze histogram of the last step: out[k] = cells still active after sweep k+1 (this rank; global after the
 * all-reduce when nranks > 1), k < n <= 128.  The oracle reports the same numbers. */
int fxb_get_freeze_histogram(fxb_sim* sim, uint64_t* out, int n);

/* Runs one step un-graphed with CUDA events between phases.  ms[0]=advect, ms[1]=divergence,
 * ms[2]=all Jacobi passes, ms[3]=gradient-subtract, ms[4]=halo exchange (0 when nranks == 1),
 * ms[5]=whole step.  The caller must have called fxb_update_frame. */
int fxb_profile_step(fxb_sim* sim, float* ms, int n);

/* Checksums of the current state, one 64-bit word per field (out3[0] velocity.xyz, [1] colour, [2] pressure): the
 * wrap-around sum over this rank's own voxels of a mix of the voxel's global index and its bits.  The sums of all ranks
 * of a z-slab run add up (mod 2^64) to the value a single GPU reports for the same state.  Synchronises the device. */
int fxb_state_checksum(fxb_sim* sim, uint64_t* out3);

/* Device time per phase accumulated by the phase marks since creation (or the last reset): ms[0]=advect (with its halo
 * exchange), ms[1]=divergence, ms[2]=all Jacobi passes, ms[3]=gradient-subtract.  Needs fxb_config.phase_timing = 1;
 * works for graph-launched steps, so the times belong to the very steps the caller timed.  Synchronises the device. */
int fxb_get_phase_times(fxb_sim* sim, double* ms, int n, int reset);

/* ---- Light-map pass ---------------------------------------------------------------------------
 * The pass that follows Fluid::Simulate in the reference's default render mode: Fluid::rayMarchL (Fluid.cpp:857-878)
 * runs CSRayMarchL.hlsl:15-80 over m_colors[m_frameParity] and writes one R11G11B10_FLOAT texel per voxel into
 * m_lightMap (Fluid.cpp:223-227): transmittance towards the light x light colour + ambient, or with light probes
 * ambient-occlusion x SH irradiance.  fxb_light_params holds the constants that shader reads, register for register. */
typedef struct fxb_light_params {
    float light_pt[3];         /* cbPerFrame g_lightPt = m_lightPt (Fluid.cpp:304; default (75, 75, -75), :171) */
    float light_color[4];      /* g_lightColor: rgb, intensity (Fluid.cpp:305; default (1, .7, .3, 3 pi), :172) */
    float ambient[4];          /* g_ambient (Fluid.cpp:306; default (1, 1, 1, 1.5 pi), :173); used iff !has_light_probes */
    float world_i[12];         /* cbPerObject g_worldI: three float4 registers as XMStoreFloat3x4 writes them (Fluid.cpp:318) */
    float world[12];           /* g_world (Fluid.cpp:319; default scaling by 10, :182) */
    uint32_t num_samples;      /* cbSampleRes g_numSamples = m_maxLightSamples (Fluid.cpp:872; default 64, :175) */
    uint32_t has_light_probes; /* m_coeffSH ? 1 : 0 (Fluid.cpp:873) */
    float sh[9][3];            /* g_roSHCoeffs: nine float3 order-3 SH coefficients (Fluid::SetSH, Fluid.cpp:278-281, 874) */
} fxb_light_params;

/* Replaces Fluid::rayMarchL: enqueues the pass on `cuda_stream` over the current colour field (so after an
 * fxb_simulate on the same stream it sees that step's result).  3D grids.  With nranks > 1 every rank calls it: a light
 * ray crosses every z-slab, so the ranks first exchange the density channel (2 bytes per voxel of the grid per rank,
 * ncclSend/ncclRecv) and each then writes the light map of its own planes; needs nx * ny to be a multiple of 4.  Light map and density scratch are allocated on first use. */
int fxb_light_map(fxb_sim* sim, const fxb_light_params* params, void* cuda_stream);
/* Synchronous copy of the light map of this rank's planes to the host: [z][y][x] uint32,
 * DXGI_FORMAT_R11G11B10_FLOAT packing (R in bits 0-10, G 11-21, B 22-31); bytes = nx*ny*count*4 (fxb_get_slab).
 * Fails if fxb_light_map has not run. */
int fxb_get_light_map(fxb_sim* sim, void* host, size_t bytes);

/* ---- Cube-map ray march with the separate light pass ------------------------------------------
 * Fluid::rayMarchV (Fluid.cpp:880-908): CSRayMarchV (= CSRayMarch.hlsl:98-196 with _LIGHT_PASS_) marches one view ray
 * per texel of the six faces of mip m_cubeMapLOD of the R8G8B8A8_UNORM cube map (Fluid.cpp:229-232) through
 * m_colors[m_frameParity], lit by the light map of fxb_light_map.  The shipped shader culls faces by a host-computed
 * mask (_CPU_CUBE_FACE_CULL_ == 1, Fluid.cpp:51-63, 900). */
typedef struct fxb_view_params {
    float eye_pt[3];          /* cbPerFrame g_eyePt (Fluid.cpp:302) */
    float world_i[12];        /* cbPerObject g_worldI, three float4 registers (Fluid.cpp:318) */
    uint32_t num_samples;     /* cbSampleRes g_numSamples = m_raySampleCount (Fluid.cpp:324-327, 898; at most 192, :174) */
    uint32_t visibility_mask; /* bit i: face i (+X, -X, +Y, -Y, +Z, -Z) is marched; fxb_cube_visibility_mask */
    uint32_t cube_size;       /* edge of the cube-map mip written: m_gridSize.x >> m_cubeMapLOD (Fluid.cpp:906) */
} fxb_view_params;

/* GenVisibilityMask (Fluid.cpp:51-63), no GPU needed: a face is marched iff the eye, in volume space, is on the inner
 * side of its plane (the cube map holds what is seen THROUGH the volume on the far faces). */
int fxb_cube_visibility_mask(const float world_i[12], const float eye_pt[3], uint32_t* mask);
/* EstimateCubeMapLOD (Fluid.cpp:141-166, with ProjectToViewport :86-106 and EstimateCubeEdgePixelSize :108-139), no GPU
 * needed: from the row-major world-view-projection matrix (row vectors, as DirectXMath holds it before the transpose
 * of Fluid.cpp:316) and the viewport size, the longest projected edge of the volume's cube gives the ideal ray sample
 * count (clamped to max_ray_samples = m_maxRaySamples) and the cube-map mip to march (0 .. num_mips - 1; the reference
 * uses 5 mips of a cube map of edge gridSize.x, Fluid.cpp:229-232).  *ray_samples is fxb_view_params.num_samples,
 * cube_size0 >> *lod is fxb_view_params.cube_size. */
int fxb_estimate_cube_lod(const float world_view_proj[16], float viewport_w, float viewport_h, uint32_t max_ray_samples,
                          uint32_t num_mips, uint32_t cube_size0, uint32_t* ray_samples, uint32_t* lod);
/* Replaces Fluid::rayMarchV: enqueues the march on `cuda_stream`.  fxb_light_map must have run (the light map is an
 * input).  Texels of culled faces and of rays that miss the volume keep their previous contents, as in the reference
 * (zero after allocation or after a change of cube_size).  1 <= cube_size <= 4096.  3D grids.  With nranks > 1 every
 * rank calls it: a view ray crosses every z-slab, so the ranks first gather the colour field and the light map of the
 * whole grid (8 + 4 bytes per voxel per rank, ncclSend/ncclRecv) and every rank then holds the complete cube map. */
int fxb_ray_march_v(fxb_sim* sim, const fxb_view_params* params, void* cuda_stream);
/* Replaces Fluid::rayMarch (Fluid.cpp:825-855), the mode without the separate light pass (Fluid::RAY_MARCH_CUBEMAP
 * alone): CSRayMarch casts the light ray — and with light probes the occlusion ray plus the SH irradiance — at every
 * view sample instead of reading the light map.  `light->num_samples` is g_numLightSamples (m_maxLightSamples,
 * Fluid.cpp:844).  Needs no fxb_light_map; same cube map, same rules as fxb_ray_march_v. */
int fxb_ray_march(fxb_sim* sim, const fxb_view_params* view, const fxb_light_params* light, void* cuda_stream);
/* Synchronous copy of the cube map written by the last fxb_ray_march_v / fxb_ray_march: [6][S][S][4] bytes (R, G, B, A UNORM8). */
int fxb_get_cube_map(fxb_sim* sim, void* host, size_t bytes);

/* ---- Volume files: the hand-off format of a field to a renderer ------------------------------
 * The reference never leaves the GPU: Fluid::Render binds m_colors[m_frameParity] as a Texture3D SRV
 * (Fluid.cpp:760-770, 841/870/897) and its ray marchers sample it as premultiplied RGBA with the density in .w
 * (RayMarch.hlsli:62-68 GetSample; CSRayMarch.hlsl:157; _PRE_MULTIPLIED_, Common.hlsli:5) at
 * uvw = pos * 0.5 + 0.5 (RayMarch.hlsli LocalToTex3DSpace), texel (x, y, z) centred at ((x, y, z) + 0.5) / N.
 * A volume file is that texture's logical contents: a 64-byte little-endian header followed by the dense payload,
 * x fastest, [z][y][x][4] IEEE half (format 1, R16G16B16A16_FLOAT) or [z][y][x] float (format 2, R32_FLOAT) — the
 * byte layout of fxb_get_field.  A rank of a multi-GPU run writes its own z-slab (z0, nz_local); a reader assembles
 * the slabs by z0. */
#define FXB_VOLUME_MAGIC "FXBV"
#define FXB_VOLUME_VERSION 1u
#define FXB_VOLUME_FLAG_PREMULTIPLIED 1u /* colour fields: rgb already multiplied by the density in .w */
typedef struct fxb_volume_header {
    char magic[4];          /* "FXBV" */
    uint32_t version;       /* FXB_VOLUME_VERSION */
    uint32_t nx, ny, nz;    /* global grid */
    uint32_t z0, nz_local;  /* planes [z0, z0 + nz_local) are stored in this file */
    uint32_t field;         /* fxb_field */
    uint32_t format;        /* 1 = half x 4 (8 bytes/voxel), 2 = float (4 bytes/voxel) */
    uint32_t flags;         /* FXB_VOLUME_FLAG_* */
    uint64_t frame;         /* fxb_simulate calls that produced the contents */
    float dt;               /* time step of the last frame (0 = paused) */
    uint32_t frame_parity;  /* m_frameParity when written */
    uint64_t payload_bytes; /* = nx * ny * nz_local * (8 or 4) */
} fxb_volume_header;        /* 64 bytes */

/* Host-only (no GPU needed).  fxb_volume_write validates the header (magic and version are filled in), writes
 * header + payload to `path`.tmp and renames it over `path`.  fxb_volume_read_header reads and validates a header.
 * fxb_volume_read also reads the payload into `data` (`capacity` >= payload_bytes, else FXB_ERR_SIZE) and fails
 * with FXB_ERR_IO on a truncated file. */
int fxb_volume_write(const char* path, const fxb_volume_header* hdr, const void* data);
int fxb_volume_read_header(const char* path, fxb_volume_header* out);
int fxb_volume_read(const char* path, fxb_volume_header* out, void* data, size_t capacity);

/* Copies `field` of this rank's slab to the host (synchronously) and writes it as a volume file. */
int fxb_export_field(fxb_sim* sim, int field, const char* path);

/* Writes a fresh 128-byte ncclUniqueId (rank 0 calls this, then ships it to the other ranks). */
int fxb_nccl_unique_id(void* out128);

const char* fxb_last_error(void);
int fxb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDX_B200_H */
