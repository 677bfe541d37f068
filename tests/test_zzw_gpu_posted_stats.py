"""fxb_post_stats / fxb_wait_stats through the C ABI (late-sorting file: added after the round's last GPU run)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fx():
    import fluidx12_b200
    return fluidx12_b200


def test_posted_stats_are_the_synchronous_ones_one_frame_late(fx):
    """fxb_post_stats / fxb_wait_stats: the record of step k read while step k + 1 is already enqueued equals what
    fxb_get_stats returned for step k on a second, synchronously driven handle."""
    n = (64, 64, 64)
    dt = fx.dt_for_grid(*n)
    a, b = fx.Fluid(), fx.Fluid()
    assert a.Init(gridSize=n) and b.Init(gridSize=n)
    want = []
    for k in range(12):
        b.step(0.0 if k == 5 else dt)
        st = b.stats()
        want.append((st.s_exec, st.jacobi_passes, st.steps, st.frame_parity, st.total_sweeps, st.bricks_processed))
    got = []
    for k in range(12):
        a.step(0.0 if k == 5 else dt)
        a.post_stats(k % 4)
        if k > 0:
            st = a.wait_stats((k - 1) % 4)
            got.append((st.s_exec, st.jacobi_passes, st.steps, st.frame_parity, st.total_sweeps, st.bricks_processed))
    st = a.wait_stats(11 % 4)
    got.append((st.s_exec, st.jacobi_passes, st.steps, st.frame_parity, st.total_sweeps, st.bricks_processed))
    assert got == want
    with pytest.raises(fx.FluidError):
        a.post_stats(4)
    c = fx.Fluid()
    assert c.Init(gridSize=n)
    with pytest.raises(fx.FluidError):
        c.wait_stats(0)          # nothing posted
    for f in (a, b, c):
        f.close()
