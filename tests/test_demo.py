"""The reference's smoke test (src/quadrotor_ilqr_test.py:6-8: main() runs without raising), on the GPU path,
plus the assertions that test never made."""
import pytest


def test_demo_module_imports_without_gpu():
    import quadrotorilqr_b200.demo as demo

    assert callable(demo.main) and callable(demo.extract_traj_array)


@pytest.mark.gpu
def test_demo_main_runs():
    from quadrotorilqr_b200 import demo

    traj_dict, costs = demo.main(plot_iters=True, verbose=False)
    assert len(costs) == 76 and abs(costs[-1] - 22556.502591980552) < 1e-6
    assert set(["desired", "optimized", "iter 0", "iter 75"]) <= set(traj_dict)
    arr = demo.extract_traj_array(traj_dict["optimized"])
    assert arr.shape == (40, 18)
