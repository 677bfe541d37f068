"""One short GPU session for the dynamic pressure-solve schedule (FXB_TAIL=1): parity first, then timing.

Written for a GPU budget of well under a minute: no torch import, results are appended to gpurun_out/shot.jsonl
stage by stage (flushed), so whatever finished before a time limit is kept.  Timing is host-side around
fxb_simulate ... fxb_sync over many steps (the graph launch is asynchronous; the sync is outside the loop)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.environ.get("FXB_SHOT_OUT") or os.path.join(ROOT, "gpurun_out", "shot.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
T0 = time.time()


def emit(**kw):
    kw["t"] = round(time.time() - T0, 2)
    with open(OUT, "a") as fh:
        fh.write(json.dumps(kw) + "\n")
        fh.flush()
        os.fsync(fh.fileno())
    print(kw, flush=True)


def make(fx, n, tail, **env):
    keys = {"FXB_TAIL": "1" if tail else "0"}
    keys.update({k: str(v) for k, v in env.items()})
    old = {k: os.environ.get(k) for k in keys}
    os.environ.update(keys)
    try:
        f = fx.Fluid()
        ok = f.Init(gridSize=n)
        if not ok:
            raise RuntimeError(f.last_error)
        return f
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def same_fields(fx, a, b):
    bad = {}
    for name, fld in (("velocity", fx.FIELD_VELOCITY), ("colour", fx.FIELD_COLOR), ("pressure", fx.FIELD_PRESSURE)):
        x, y = a.get_field(fld), b.get_field(fld)
        if name == "velocity":
            x, y = x[..., :3], y[..., :3]
        bad[name] = int((x != y).sum())
    return bad


def time_steps(f, dt, steps):
    f.sync()
    t = time.perf_counter()
    for _ in range(steps):
        f.step(dt)
    f.sync()
    return (time.perf_counter() - t) / steps * 1e3


def oracle_check(fx, label, env, n=(64, 64, 64), steps=6):
    try:
        import oracle
        dt = fx.dt_for_grid(*n)
        f, o = make(fx, n, env is not None, **(env or {})), oracle.FluidOracle(*n)
        for _ in range(steps):
            f.step(dt)
            o.step(dt)
        f.sync()
        bad = {}
        for name, gf, of in (("velocity", fx.FIELD_VELOCITY, oracle.FIELD_VEL), ("colour", fx.FIELD_COLOR, oracle.FIELD_COLOR),
                             ("pressure", fx.FIELD_PRESSURE, oracle.FIELD_PRESSURE)):
            x, y = f.get_field(gf), o.get_field(of)
            if name == "velocity":
                x, y = x[..., :3], y[..., :3]
            bad[name] = int((x != y).sum())
        emit(stage="oracle", variant=label, grid=n, mismatches=bad, s_exec=[f.stats().s_exec, o.s_exec], tail=f.tail_stats())
        f.close()
    except Exception as e:
        emit(stage="oracle", variant=label, error=repr(e))


def timing(fx, n, spin, steps, variants, all_fields=False):
    """Spin up on the default schedule, copy the state into each variant, time `steps` graph-launched steps."""
    dt = fx.dt_for_grid(*n)
    base = make(fx, n, False)
    for _ in range(spin):
        base.step(dt)
    base.sync()
    flds = (fx.FIELD_VELOCITY, fx.FIELD_COLOR, fx.FIELD_PRESSURE)
    state = {fld: base.get_field(fld) for fld in flds}
    res = {"default": round(time_steps(base, dt, steps), 4)}
    base.UpdateFrame(dt)
    res["default_phases"] = {k: round(v, 4) for k, v in base.profile_step().items()}
    # "pc": pressure and colour (the colour field is advected by the velocity, so the pair pins all three at 2/3 of
    # the host traffic — used at 512^3 inside bench.py's time budget)
    cmp_flds = (fx.FIELD_PRESSURE, fx.FIELD_COLOR) if all_fields == "pc" else flds if all_fields else (fx.FIELD_PRESSURE,)
    ref = {fld: base.get_field(fld) for fld in cmp_flds}
    emit(stage="timing", grid=n, **res)
    base.close()
    for label, env in variants:
        try:
            f = make(fx, n, True, **env)
            for fld, arr in state.items():
                f.set_field(fld, arr)
            ms = time_steps(f, dt, steps)
            f.UpdateFrame(dt)
            ph = f.profile_step()
            bad = {str(fld): int((f.get_field(fld) != ref[fld]).sum()) for fld in cmp_flds}
            emit(stage="timing", grid=n, variant=label, ms=round(ms, 4), phases={k: round(v, 4) for k, v in ph.items()},
                 tail=f.tail_stats(), passes=f.stats().jacobi_passes, s_exec=f.stats().s_exec, mismatch_vs_default=bad)
            f.close()
        except Exception as e:  # keep going: the other variants are still informative
            emit(stage="timing", grid=n, variant=label, error=repr(e))


def light_map_timing(fx, n, spin=100, reps=5):
    """The light-map pass (fxb_light_map) on a developed plume: host-timed launches, reference default
    constants, without and with light probes (seeded SH coefficients)."""
    try:
        f = make(fx, n, False)
        dt = fx.dt_for_grid(*n)
        for _ in range(spin):
            f.step(dt)
        f.sync()
        lit = int((f.get_field(fx.FIELD_COLOR)[..., 3].astype(np.float32) >= 0.01).sum())
        for probes in (0, 1):
            p = fx.FxbLightParams.reference_defaults()
            p.has_light_probes = probes
            sh = np.random.default_rng(1).standard_normal((9, 3)) * 0.3
            sh[0] = np.abs(sh[0]) + 0.8
            for i in range(9):
                p.sh[i][:] = sh[i].tolist()
            f.RayMarchL(p)
            f.sync()
            t = time.perf_counter()
            for _ in range(reps):
                f.RayMarchL(p)
            f.sync()
            ms = (time.perf_counter() - t) / reps * 1e3
            vox = n[0] * n[1] * n[2]
            emit(stage="light_map", grid=n, probes=probes, ms=round(ms, 4), voxels_with_smoke=lit,
                 gvoxels_per_s=round(vox / ms / 1e6, 2), algorithmic_gbs=round(12.0 * vox / ms / 1e6, 1),
                 words_distinct=int(len(np.unique(f.get_light_map()[::4, ::4, ::4]))))
        # the view-ray march into the cube map (fxb_ray_march_v), lit by the light map just written
        import ctypes as C
        v = fx.FxbViewParams()
        v.eye_pt[:] = [4.0, 16.0, -40.0]
        v.world_i[:] = [0.1, 0, 0, 0, 0, 0.1, 0, 0, 0, 0, 0.1, 0]
        v.num_samples, v.cube_size = 192, n[0]
        mask = C.c_uint32()
        fx.lib().fxb_cube_visibility_mask(v.world_i, v.eye_pt, C.byref(mask))
        v.visibility_mask = mask.value
        f.RayMarchV(v)
        f.sync()
        t = time.perf_counter()
        for _ in range(reps):
            f.RayMarchV(v)
        f.sync()
        ms = (time.perf_counter() - t) / reps * 1e3
        cube = f.get_cube_map()
        emit(stage="light_map", grid=n, pass_="ray_march_v", cube_size=n[0], ray_samples=192, ms=round(ms, 4),
             mrays_per_s=round(6 * n[0] * n[0] / ms / 1e3, 2), texels_with_smoke=int((cube[..., 3] > 0).sum()))
        f.close()
    except Exception as e:
        emit(stage="light_map", grid=n, error=repr(e))


def bench_mode(fx):
    """The short list bench.py runs (in a child process, after its own measurement) so that every default bench run
    also times the opt-in variants: most informative first, nothing that can spin on the device (no TMA-staged
    window), every variant compared bit for bit with the default schedule's fields (the oracle is not used here)."""
    timing(fx, (256, 256, 256), 100, 40, [("tail", {}), ("advect2_only", {"FXB_TAIL": 0, "FXB_ADVECT": 2}),
                                          ("tail_advect2", {"FXB_ADVECT": 2}),
                                          ("tail_dense2_only", {"FXB_TAIL_DENSE": 2, "FXB_TAIL_SPARSE_CAP": 0}),
                                          ("tail_cap2048_dense2", {"FXB_TAIL_SPARSE_CAP": 2048, "FXB_TAIL_DENSE": 2}),
                                          ("tail_dense_only", {"FXB_TAIL_SPARSE_CAP": 0}),
                                          ("tail_cpasync", {"FXB_TAIL_CPASYNC": 1}),
                                          ("tail_pass0", {"FXB_PASS0": 2}),
                                          ("tail_thr256", {"FXB_TAIL_THRESHOLD": 256, "FXB_TAIL_MAINS": 12}),
                                          ("tail_mains1", {"FXB_TAIL_MAINS": 1})], all_fields=True)
    light_map_timing(fx, (256, 256, 256))
    timing(fx, (512, 512, 512), 100, 20, [("tail", {}), ("advect2_only", {"FXB_TAIL": 0, "FXB_ADVECT": 2}),
                                          ("tail_advect2", {"FXB_ADVECT": 2}),
                                          ("tail_dense2_only", {"FXB_TAIL_DENSE": 2, "FXB_TAIL_SPARSE_CAP": 0}),
                                          ("tail_pass0", {"FXB_PASS0": 2})],
           all_fields="pc")
    light_map_timing(fx, (512, 512, 512))
    emit(stage="done")


def main():
    import fluidx12_b200 as fx
    emit(stage="import")
    if "--bench" in sys.argv:
        return bench_mode(fx)
    oracle_check(fx, "tail", {})
    oracle_check(fx, "tail_dense_only", {"FXB_TAIL_SPARSE_CAP": 0}, n=(64, 64, 40), steps=4)
    timing(fx, (256, 256, 256), 100, 40, [("tail", {}), ("tail_dense_only", {"FXB_TAIL_SPARSE_CAP": 0}),
                                          ("tail_cap1024", {"FXB_TAIL_SPARSE_CAP": 1024}),
                                          ("tail_cap2048", {"FXB_TAIL_SPARSE_CAP": 2048}),
                                          ("tail_cap2048_dense2", {"FXB_TAIL_SPARSE_CAP": 2048, "FXB_TAIL_DENSE": 2}),
                                          ("tail_grid592", {"FXB_TAIL_GRID": 592}),
                                          ("tail_mains1", {"FXB_TAIL_MAINS": 1}),
                                          ("tail_thr8192_m5", {"FXB_TAIL_THRESHOLD": 8192, "FXB_TAIL_MAINS": 5}),
                                          ("tail_cpasync", {"FXB_TAIL_CPASYNC": 1}),
                                          ("tail_tma", {"FXB_TAIL_CPASYNC": 2}),
                                          ("tail_dense2", {"FXB_TAIL_DENSE": 2}),
                                          ("tail_dense2_only", {"FXB_TAIL_DENSE": 2, "FXB_TAIL_SPARSE_CAP": 0}),
                                          ("tail_dense2_only_tma", {"FXB_TAIL_DENSE": 2, "FXB_TAIL_SPARSE_CAP": 0, "FXB_TAIL_CPASYNC": 2}),
                                          ("tail_pass0", {"FXB_PASS0": 2}),
                                          ("tail_advect2", {"FXB_ADVECT": 2}),
                                          ("advect2_only", {"FXB_TAIL": 0, "FXB_ADVECT": 2}),
                                          ("tail_thr256", {"FXB_TAIL_THRESHOLD": 256, "FXB_TAIL_MAINS": 12}),
                                          ("tail_thr1024", {"FXB_TAIL_THRESHOLD": 1024, "FXB_TAIL_MAINS": 12})],
           all_fields=True)
    timing(fx, (512, 512, 512), 100, 20, [("tail", {}), ("advect2_only", {"FXB_TAIL": 0, "FXB_ADVECT": 2}), ("tail_pass0", {"FXB_PASS0": 2}), ("tail_thr2048", {"FXB_TAIL_THRESHOLD": 2048, "FXB_TAIL_MAINS": 12})])
    oracle_check(fx, "default", None)
    emit(stage="done")


if __name__ == "__main__":
    try:
        main()
    except Exception as e:
        emit(stage="fatal", error=repr(e))
        raise
