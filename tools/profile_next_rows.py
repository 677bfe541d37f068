"""Drives the light-map pass and the ray marches on a developed plume so that ncu can capture them (no torch).

    ncu --set full --clock-control none --import-source on -k regex:"light_map_kernel|ray_march" -c 3 \
        -o gpurun_out/next_rows -f python tools/profile_next_rows.py 256
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fluidx12_b200 as fx
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    f = fx.Fluid()
    if not f.Init(gridSize=(n, n, n)):
        raise SystemExit(f.last_error)
    dt = fx.dt_for_grid(n, n, n)
    for _ in range(100):
        f.step(dt)
    lp = fx.FxbLightParams.reference_defaults()
    f.RayMarchL(lp)                       # extract_density_kernel + light_map_kernel
    v = fx.FxbViewParams()
    v.eye_pt[:] = [4.0, 16.0, -40.0]
    v.world_i[:] = lp.world_i[:]
    v.num_samples, v.cube_size = 192, n
    mask = C.c_uint32()
    fx.lib().fxb_cube_visibility_mask(v.world_i, v.eye_pt, C.byref(mask))
    v.visibility_mask = mask.value
    f.RayMarchV(v)                        # ray_march_v_kernel
    v.cube_size = max(n // 4, 8)
    f.RayMarch(v, lp)                     # ray_march_kernel (light computed per view sample: keep the cube small)
    f.sync()
    print("light map words:", len(set(f.get_light_map()[::8, ::8, ::8].reshape(-1).tolist())), "cube texels with smoke:",
          int((f.get_cube_map()[..., 3] > 0).sum()))
    f.close()


if __name__ == "__main__":
    main()
