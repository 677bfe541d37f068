// fxb_render_api.cu — the C ABI entry points of the rows past the simulation step: the light-map
// pass (Fluid::rayMarchL), the cube-map ray march (Fluid::rayMarchV) and volume files (the renderer hand-off format).
// Kernels: lightmap.cu, raymarch.cu.  The step itself (Init / UpdateFrame / Simulate, field I/O, statistics) is in
// fxb_api.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "fxb_internal.h"

namespace {
int fail(int code, const std::string& msg) { return fxb::api_fail(code, msg); }
using fxb::field_device_ptr;
}  // namespace

extern "C" {

// ---- light-map pass (Fluid::rayMarchL, Fluid.cpp:857-878; kernels in lightmap.cu) ---------------------------------
static_assert(sizeof(fxb_light_params) == 4 * (3 + 4 + 4 + 12 + 12 + 2 + 27), "fxb_light_params is passed to the kernel as is");

int fxb_light_map(fxb_sim* s, const fxb_light_params* params, void* cuda_stream) {
    if (!s || !params) return fail(FXB_ERR_INVALID, "fxb_light_map: null argument");
    if (s->cfg.nz <= 1) return fail(FXB_ERR_INVALID, "fxb_light_map: 3D grids only (the reference renders none other, Fluid.cpp:296)");
    if (s->multi() && s->plane_voxels() % 4 != 0)  // 16-byte loads / 8-byte stores of the extraction start at a plane
        return fail(FXB_ERR_INVALID, "fxb_light_map: with nranks > 1 nx * ny must be a multiple of 4");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    if (!s->light_map || !s->light_density) {
        // the light map covers the rank's own planes; the density scratch covers the WHOLE grid, because a light ray
        // crosses every z-slab (2 bytes per voxel; the other ranks' planes arrive over NCCL inside the pass)
        if (!s->light_map) FXB_CUDA(cudaMalloc((void**)&s->light_map, s->own_voxels() * sizeof(unsigned)));
        FXB_CUDA(cudaMalloc((void**)&s->light_density, (s->plane_voxels() * s->cfg.nz + 4) * sizeof(unsigned short)));
    }
    // what Fluid::Render binds: m_colors[m_frameParity] (SRV_TABLE_RAY_MARCH + !m_frameParity, Fluid.cpp:760-770, 870)
    const char* colour_own = static_cast<const char*>(s->col[s->parity]) + s->own_offset() * 8;
    // (the launcher has already consumed the sticky error: report the code it returned)
    const cudaError_t le = fxb::launch_light_map(s->dom, colour_own, s->light_density, s->light_map, params, &s->comm,
                                                 (cudaStream_t)cuda_stream);
    if (le != cudaSuccess)
        return fail(s->multi() ? FXB_ERR_NCCL : FXB_ERR_CUDA, "fxb_light_map: launch failed: " +
                    (s->multi() ? fxb::halo_last_error() + " / " : std::string()) + cudaGetErrorString(le));
    s->last_stream = (cudaStream_t)cuda_stream;
    return FXB_OK;
}

int fxb_get_light_map(fxb_sim* s, void* host, size_t bytes) {
    if (!s || !host) return fail(FXB_ERR_INVALID, "fxb_get_light_map: null argument");
    if (!s->light_map) return fail(FXB_ERR_INVALID, "fxb_get_light_map: fxb_light_map has not run");
    if (bytes != s->own_voxels() * sizeof(unsigned)) return fail(FXB_ERR_SIZE, "fxb_get_light_map: size mismatch");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    FXB_CUDA(cudaMemcpy(host, s->light_map, bytes, cudaMemcpyDeviceToHost));
    return FXB_OK;
}

// ---- cube-map ray march (Fluid::rayMarchV, Fluid.cpp:880-908; kernel in raymarch.cu) --------------------------------
static_assert(sizeof(fxb_view_params) == 4 * (3 + 12 + 3), "fxb_view_params is passed to the kernel as is");

int fxb_cube_visibility_mask(const float world_i[12], const float eye_pt[3], uint32_t* mask) {
    if (!world_i || !eye_pt || !mask) return fail(FXB_ERR_INVALID, "fxb_cube_visibility_mask: null argument");
    uint32_t m = 0;
    for (int face = 0; face < 6; ++face) {
        const float* w = world_i + 4 * (face >> 1);
        const float v = ((eye_pt[0] * w[0] + eye_pt[1] * w[1]) + eye_pt[2] * w[2]) + w[3];
        m |= (uint32_t)((face & 1) ? v > -1.0f : v < 1.0f) << face;  // IsCubeFaceVisible, Fluid.cpp:41-46
    }
    *mask = m;
    return FXB_OK;
}

namespace {
// (Re)allocates the cube-map mip for edge `size`, zero-filled like a new committed resource.
int ensure_cube_map(fxb_sim* s, uint32_t size) {
    if (size < 1 || size > 4096) return fail(FXB_ERR_INVALID, "cube_size out of range (1..4096)");
    if (s->cube_size == size) return FXB_OK;
    FXB_CUDA(cudaDeviceSynchronize());
    cudaFree(s->cube_map);
    s->cube_map = nullptr;
    s->cube_size = 0;
    const size_t bytes = (size_t)6 * size * size * sizeof(unsigned);
    FXB_CUDA(cudaMalloc((void**)&s->cube_map, bytes));
    FXB_CUDA(cudaMemset(s->cube_map, 0, bytes));
    s->cube_size = size;
    return FXB_OK;
}
// The colour field (and, if asked, the light map) of the whole grid as the view-ray march needs them.  One GPU: the
// fields themselves.  z-slabs: every rank copies its planes into a whole-grid array and the ranks exchange their slabs
// (HaloComm::all_gather_slabs; 8 + 4 bytes per voxel of the grid per rank), after which every rank marches all 6 S^2
// rays — a few hundred thousand, far cheaper than a second exchange of the cube map.
int whole_fields(fxb_sim* s, bool need_light_map, cudaStream_t st, const void** colour, const unsigned** light_map) {
    *colour = s->col[s->parity];
    *light_map = s->light_map;
    if (!s->multi()) return FXB_OK;
    const size_t plane = s->plane_voxels(), whole = plane * s->cfg.nz, z0 = (size_t)s->dom.z_own0;
    if (!s->whole_colour) FXB_CUDA(cudaMalloc(&s->whole_colour, whole * 8));
    FXB_CUDA(cudaMemcpyAsync(static_cast<char*>(s->whole_colour) + plane * z0 * 8,
                             static_cast<const char*>(s->col[s->parity]) + s->own_offset() * 8, s->own_voxels() * 8,
                             cudaMemcpyDeviceToDevice, st));
    if (!s->comm.all_gather_slabs(s->whole_colour, plane * 8, (int)s->cfg.nz, st))
        return fail(FXB_ERR_NCCL, "colour gather failed: " + fxb::halo_last_error());
    *colour = s->whole_colour;
    if (need_light_map) {
        if (!s->whole_light_map) FXB_CUDA(cudaMalloc((void**)&s->whole_light_map, whole * 4));
        FXB_CUDA(cudaMemcpyAsync(s->whole_light_map + plane * z0, s->light_map, s->own_voxels() * 4,
                                 cudaMemcpyDeviceToDevice, st));
        if (!s->comm.all_gather_slabs(s->whole_light_map, plane * 4, (int)s->cfg.nz, st))
            return fail(FXB_ERR_NCCL, "light-map gather failed: " + fxb::halo_last_error());
        *light_map = s->whole_light_map;
    }
    return FXB_OK;
}
}  // namespace

int fxb_estimate_cube_lod(const float m[16], float viewport_w, float viewport_h, uint32_t max_ray_samples,
                          uint32_t num_mips, uint32_t cube_size0, uint32_t* ray_samples, uint32_t* lod) {
    if (!m || !ray_samples || !lod || num_mips < 1 || cube_size0 < 1)
        return fail(FXB_ERR_INVALID, "fxb_estimate_cube_lod: bad argument");
    // ProjectToViewport (Fluid.cpp:86-106): the eight corners of [-1, 1]^3 through XMVector3TransformCoord (row vector
    // times matrix, divided by w), to viewport pixels
    static const float corner[8][3] = {{1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {-1, 1, -1}, {1, 1, -1}, {-1, -1, -1}, {1, -1, -1}};
    float px[8], py[8];
    for (int i = 0; i < 8; ++i) {
        const float x = corner[i][0], y = corner[i][1], z = corner[i][2];
        const float rx = x * m[0] + y * m[4] + z * m[8] + m[12], ry = x * m[1] + y * m[5] + z * m[9] + m[13];
        const float rw = x * m[3] + y * m[7] + z * m[11] + m[15];
        px[i] = (rx / rw * 0.5f + 0.5f) * viewport_w;
        py[i] = (ry / rw * -0.5f + 0.5f) * viewport_h;
    }
    // EstimateCubeEdgePixelSize (Fluid.cpp:108-139): the longest of the twelve projected edges
    static const unsigned char edge[12][2] = {{0, 1}, {3, 2}, {1, 3}, {2, 0}, {4, 5}, {7, 6}, {5, 7}, {6, 4}, {1, 4}, {6, 3}, {5, 0}, {2, 7}};
    float longest = 0.0f;
    for (int i = 0; i < 12; ++i) {
        const float ex = px[edge[i][1]] - px[edge[i][0]], ey = py[edge[i][1]] - py[edge[i][0]];
        longest = std::max(std::sqrt(ex * ex + ey * ey), longest);
    }
    // EstimateCubeMapLOD (Fluid.cpp:141-166), upscale = 2, raySampleCountScale = 2
    float s = longest / 2.0f;
    float amount = 2.0f * s / std::sqrt(3.0f);
    const uint32_t count = (uint32_t)std::ceil(amount);
    *ray_samples = std::min(count, max_ray_samples);
    amount = std::min(amount, (float)*ray_samples);
    s = amount / 2.0f * std::sqrt(3.0f);
    const float level = std::min(std::max(std::log2((float)cube_size0 / s), 0.0f), 255.0f);  // the reference casts to uint8_t
    *lod = std::min<uint32_t>((uint32_t)(unsigned char)level, num_mips - 1);
    return FXB_OK;
}

int fxb_ray_march_v(fxb_sim* s, const fxb_view_params* params, void* cuda_stream) {
    if (!s || !params) return fail(FXB_ERR_INVALID, "fxb_ray_march_v: null argument");
    if (s->cfg.nz <= 1) return fail(FXB_ERR_INVALID, "fxb_ray_march_v: 3D grids only");
    if (!s->light_map) return fail(FXB_ERR_INVALID, "fxb_ray_march_v: fxb_light_map has not run (the light map is an input)");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    if (const int rc = ensure_cube_map(s, params->cube_size)) return rc;
    const void* colour;
    const unsigned* light_map;
    if (const int rc = whole_fields(s, true, (cudaStream_t)cuda_stream, &colour, &light_map)) return rc;
    FXB_CUDA(fxb::launch_ray_march_v(s->dom, colour, light_map, s->cube_map, params, (cudaStream_t)cuda_stream));
    s->last_stream = (cudaStream_t)cuda_stream;
    return FXB_OK;
}

int fxb_ray_march(fxb_sim* s, const fxb_view_params* view, const fxb_light_params* light, void* cuda_stream) {
    if (!s || !view || !light) return fail(FXB_ERR_INVALID, "fxb_ray_march: null argument");
    if (s->cfg.nz <= 1) return fail(FXB_ERR_INVALID, "fxb_ray_march: 3D grids only");
    if (s->multi() && s->plane_voxels() % 4 != 0)
        return fail(FXB_ERR_INVALID, "fxb_ray_march: with nranks > 1 nx * ny must be a multiple of 4");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    if (const int rc = ensure_cube_map(s, view->cube_size)) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const void* colour;
    const unsigned* unused;
    if (const int rc = whole_fields(s, false, st, &colour, &unused)) return rc;
    if (!s->light_density)
        FXB_CUDA(cudaMalloc((void**)&s->light_density, (s->plane_voxels() * s->cfg.nz + 4) * sizeof(unsigned short)));
    // the density channel of the whole grid: extracted from the (gathered) colour field in one go
    FXB_CUDA(fxb::launch_extract_density(colour, s->light_density, s->plane_voxels() * s->cfg.nz, st));
    FXB_CUDA(fxb::launch_ray_march(s->dom, colour, s->light_density, s->cube_map, view, light, st));
    s->last_stream = (cudaStream_t)cuda_stream;
    return FXB_OK;
}

int fxb_get_cube_map(fxb_sim* s, void* host, size_t bytes) {
    if (!s || !host) return fail(FXB_ERR_INVALID, "fxb_get_cube_map: null argument");
    if (!s->cube_map) return fail(FXB_ERR_INVALID, "fxb_get_cube_map: fxb_ray_march_v has not run");
    if (bytes != (size_t)6 * s->cube_size * s->cube_size * 4) return fail(FXB_ERR_SIZE, "fxb_get_cube_map: size mismatch");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    FXB_CUDA(cudaMemcpy(host, s->cube_map, bytes, cudaMemcpyDeviceToHost));
    return FXB_OK;
}

// ---- volume files (include/fluidx_b200.h: the renderer hand-off format) --------------------------
namespace {
static_assert(sizeof(fxb_volume_header) == 64, "fxb_volume_header is a 64-byte wire structure");

int check_volume_header(const fxb_volume_header& h, const char* who) {
    const std::string w(who);
    if (memcmp(h.magic, FXB_VOLUME_MAGIC, 4) != 0) return fail(FXB_ERR_IO, w + ": not a volume file (magic)");
    if (h.version != FXB_VOLUME_VERSION) return fail(FXB_ERR_IO, w + ": unsupported volume file version");
    if (h.format != 1 && h.format != 2) return fail(FXB_ERR_IO, w + ": unknown element format");
    if (h.nx == 0 || h.ny == 0 || h.nz == 0 || h.nz_local == 0 || (uint64_t)h.z0 + h.nz_local > h.nz)
        return fail(FXB_ERR_IO, w + ": bad grid / slab extent");
    if (h.field > FXB_FIELD_COLOR_PREV || (h.format == 2) != (h.field == FXB_FIELD_PRESSURE))
        return fail(FXB_ERR_IO, w + ": field and element format do not match");
    if (h.payload_bytes != (uint64_t)h.nx * h.ny * h.nz_local * (h.format == 1 ? 8u : 4u))
        return fail(FXB_ERR_IO, w + ": payload size does not match the extent");
    return FXB_OK;
}
}  // namespace

int fxb_volume_write(const char* path, const fxb_volume_header* hdr, const void* data) {
    if (!path || !*path || !hdr || !data) return fail(FXB_ERR_INVALID, "fxb_volume_write: null argument");
    fxb_volume_header h = *hdr;
    memcpy(h.magic, FXB_VOLUME_MAGIC, 4);
    h.version = FXB_VOLUME_VERSION;
    if (const int rc = check_volume_header(h, "fxb_volume_write")) return rc == FXB_ERR_IO ? FXB_ERR_INVALID : rc;
    const std::string tmp = std::string(path) + ".tmp";
    FILE* fp = fopen(tmp.c_str(), "wb");
    if (!fp) return fail(FXB_ERR_IO, "fxb_volume_write: cannot create " + tmp);
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1 && fwrite(data, 1, h.payload_bytes, fp) == h.payload_bytes;
    ok = (fclose(fp) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return fail(FXB_ERR_IO, std::string("fxb_volume_write: writing ") + path + " failed");
    }
    return FXB_OK;
}

int fxb_volume_read_header(const char* path, fxb_volume_header* out) {
    if (!path || !out) return fail(FXB_ERR_INVALID, "fxb_volume_read_header: null argument");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(FXB_ERR_IO, std::string("fxb_volume_read_header: cannot open ") + path);
    const bool ok = fread(out, sizeof *out, 1, fp) == 1;
    fclose(fp);
    if (!ok) return fail(FXB_ERR_IO, "fxb_volume_read_header: file shorter than a header");
    return check_volume_header(*out, "fxb_volume_read_header");
}

int fxb_volume_read(const char* path, fxb_volume_header* out, void* data, size_t capacity) {
    if (!data) return fail(FXB_ERR_INVALID, "fxb_volume_read: null argument");
    if (const int rc = fxb_volume_read_header(path, out)) return rc;
    if (capacity < out->payload_bytes) return fail(FXB_ERR_SIZE, "fxb_volume_read: buffer smaller than the payload");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(FXB_ERR_IO, std::string("fxb_volume_read: cannot open ") + path);
    bool ok = fseek(fp, (long)sizeof *out, SEEK_SET) == 0 && fread(data, 1, out->payload_bytes, fp) == out->payload_bytes;
    ok = ok && fgetc(fp) == EOF;  // nothing may follow the payload
    fclose(fp);
    if (!ok) return fail(FXB_ERR_IO, "fxb_volume_read: payload truncated or followed by extra bytes");
    return FXB_OK;
}

int fxb_export_field(fxb_sim* s, int field, const char* path) {
    if (!s || !path) return fail(FXB_ERR_INVALID, "fxb_export_field: null argument");
    size_t eb; int err;
    if (!field_device_ptr(s, field, &eb, &err)) return fail(err, "fxb_export_field: bad field");
    fxb_volume_header h = {};
    h.nx = s->cfg.nx; h.ny = s->cfg.ny; h.nz = s->cfg.nz;
    h.z0 = (uint32_t)s->dom.z_own0;
    h.nz_local = (uint32_t)(s->dom.z_own1 - s->dom.z_own0);
    h.field = (uint32_t)field;
    h.format = eb == 8 ? 1u : 2u;
    h.flags = (field == FXB_FIELD_COLOR || field == FXB_FIELD_COLOR_PREV) ? FXB_VOLUME_FLAG_PREMULTIPLIED : 0u;
    h.frame = s->steps;
    h.dt = s->dt;
    h.frame_parity = (uint32_t)s->parity;
    h.payload_bytes = (uint64_t)s->own_voxels() * eb;
    std::vector<char> host;
    try {
        host.resize(h.payload_bytes);
    } catch (const std::bad_alloc&) {
        return fail(FXB_ERR_IO, "fxb_export_field: no host memory for the staging buffer");
    }
    if (const int rc = fxb_get_field(s, field, host.data(), host.size())) return rc;
    return fxb_volume_write(path, &h, host.data());
}

}  // extern "C"
