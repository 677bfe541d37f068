// fluid.hpp — header-only C++ mirror of the reference's `Fluid` simulation surface over the C ABI.
//
// Reference: FluidX12/Content/Fluid.h:20-33 (class Fluid: Init / UpdateFrame / Simulate), callers
// FluidX12/FluidX12.cpp:197-201 (Init), :282 (UpdateFrame), :536 (Simulate).  Same member names, same argument
// order; the D3D12-only arguments (command list, descriptor-table library, uploaders, render formats, view and
// projection matrices, eye point) have no meaning on the CUDA path and are dropped — `Simulate` takes the CUDA
// stream where the reference takes the command list.  Like the reference, Init returns false on failure
// (XUSG_N_RETURN, XUSG/Core/XUSG.h:12-15) and UpdateFrame / Simulate cannot fail loudly (they are `void` there);
// here they return the fxb_status so a caller may check it.
#pragma once

#include <cstdint>
#include <string>

#include "fluidx_b200.h"

namespace fluidx_b200 {

struct UInt3 {  // stands in for DirectX::XMUINT3 (gridSize of Fluid::Init)
    uint32_t x, y, z;
};

class Fluid {
public:
    static const uint8_t FrameCount = 3;  // Fluid.h:35 — the reference's constant-buffer ring; kept for call-site parity

    Fluid() = default;
    Fluid(const Fluid&) = delete;
    Fluid& operator=(const Fluid&) = delete;
    virtual ~Fluid() { fxb_destroy(m_sim); }

    // Fluid::Init (Fluid.cpp:189-270).  `cfg` carries what the reference hard-codes or takes from D3D12:
    // sampler addressing (Fluid.cpp:452 MIRROR / FluidEZ.cpp:406 CLAMP), ITER, device, z-slab rank.
    bool Init(const UInt3& gridSize, const fxb_config* cfg = nullptr) {
        fxb_config c;
        if (cfg) c = *cfg; else fxb_config_default(&c);
        c.struct_size = sizeof(fxb_config);
        c.nx = gridSize.x; c.ny = gridSize.y; c.nz = gridSize.z;
        fxb_destroy(m_sim);
        m_sim = nullptr;
        if (fxb_create(&c, &m_sim) != FXB_OK) {
            m_error = fxb_last_error();
            return false;
        }
        m_gridSize = gridSize;
        m_frameParity = 0;
        return true;
    }

    // Fluid::UpdateFrame, simulation part (Fluid.cpp:283-291, 344-345): store dt, flip the parity iff dt > 0.
    int UpdateFrame(float timeStep, uint8_t frameIndex = 0) {
        (void)frameIndex;  // the reference indexes its 3-slot constant-buffer ring with it (Fluid.cpp:288)
        m_timeStep = timeStep;
        if (timeStep > 0.0f) m_frameParity = !m_frameParity;
        return fxb_update_frame(m_sim, timeStep);
    }

    // Fluid::Simulate (Fluid.cpp:348-410): enqueue advect + project; `stream` replaces the command list.
    int Simulate(void* stream = nullptr, uint8_t frameIndex = 0) {
        (void)frameIndex;
        return fxb_simulate(m_sim, stream);
    }

    // dt rule of FluidX::OnUpdate (FluidX12.cpp:266-267)
    float TimeStepForGrid() const {
        float dt = 0.0f;
        fxb_dt_for_grid(m_gridSize.x, m_gridSize.y, m_gridSize.z, &dt);
        return dt;
    }

    // Fluid::rayMarchL (Fluid.cpp:857-878), the light-map pass of Fluid::Render: CSRayMarchL over m_colors[m_frameParity].
    int RayMarchL(const fxb_light_params& params, void* stream = nullptr) { return fxb_light_map(m_sim, &params, stream); }

    // Fluid::rayMarchV (Fluid.cpp:880-908): CSRayMarchV into one mip of the cube map, lit by RayMarchL's light map.
    int RayMarchV(const fxb_view_params& params, void* stream = nullptr) { return fxb_ray_march_v(m_sim, &params, stream); }

    // Fluid::rayMarch (Fluid.cpp:825-855): CSRayMarch, the march that computes the light at every view sample.
    int RayMarch(const fxb_view_params& view, const fxb_light_params& light, void* stream = nullptr) {
        return fxb_ray_march(m_sim, &view, &light, stream);
    }

    // No counterpart in the reference, whose renderer binds m_colors[m_frameParity] in place (Fluid.cpp:760-770, 841):
    // writes the field as a volume file for a renderer outside the process (fluidx_b200.h, fxb_volume_header).
    bool Export(const char* path, int field = FXB_FIELD_COLOR) {
        if (fxb_export_field(m_sim, field, path) == FXB_OK) return true;
        m_error = fxb_last_error();
        return false;
    }

    fxb_sim* handle() const { return m_sim; }
    const std::string& last_error() const { return m_error; }
    uint8_t frameParity() const { return m_frameParity; }

protected:
    fxb_sim* m_sim = nullptr;
    UInt3 m_gridSize{0, 0, 0};  // Fluid.h:111
    float m_timeStep = 0.0f;     // Fluid.h:126
    uint8_t m_frameParity = 0;   // Fluid.h:124
    std::string m_error;
};

// FluidEZ (Content/FluidEZ.h:20-32; the app's default runtime path, FluidX12.cpp:40): the same two dispatches
// (FluidEZ.cpp:373-449) with the advection sampler LINEAR_CLAMP (FluidEZ.cpp:406) instead of LINEAR_MIRROR.
class FluidEZ : public Fluid {
public:
    bool Init(const UInt3& gridSize, const fxb_config* cfg = nullptr) {
        fxb_config c;
        if (cfg) c = *cfg; else { fxb_config_default(&c); c.address_mode = FXB_ADDRESS_CLAMP; }
        return Fluid::Init(gridSize, &c);
    }
};

}  // namespace fluidx_b200
