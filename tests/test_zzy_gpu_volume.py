"""Export of a live field as a volume file through the C ABI.  In a late-sorting file: a surprise in
a "next" row must not stop the hot path's own GPU tests under -x."""
import numpy as np
import pytest

import fluidx12_b200 as fx
from fluidx12_b200 import volume


@pytest.mark.gpu
def test_export_of_a_live_field_is_what_get_field_returns(tmp_path):
    n = (64, 64, 24)
    f = fx.Fluid()
    assert f.Init(gridSize=n), f.last_error
    dt = fx.dt_for_grid(*n)
    for _ in range(7):
        f.step(dt)
    f.step(0.0)  # a paused frame: dt = 0 is recorded, the parity does not flip
    for fld, name in ((fx.FIELD_COLOR, "c"), (fx.FIELD_VELOCITY, "v"), (fx.FIELD_PRESSURE, "p")):
        p = str(tmp_path / (name + ".fxbv"))
        f.export(p, fld)
        a, h = volume.read_numpy(p)
        want = f.get_field(fld)
        assert a.dtype == want.dtype and a.tobytes() == want.tobytes()
        assert (h["nx"], h["ny"], h["nz"], h["z0"], h["nz_local"]) == (64, 64, 24, 0, 24)
        assert h["frame"] == 8 and h["dt"] == 0.0 and h["frame_parity"] == f.stats().frame_parity == 1
        assert h["flags"] == (volume.FLAG_PREMULTIPLIED if fld == fx.FIELD_COLOR else 0)
    assert np.abs(f.get_field(fx.FIELD_COLOR).astype(np.float32)).max() > 0
    with pytest.raises(fx.FluidError):
        f.export(str(tmp_path / "bad.fxbv"), 99)
    f.close()
