"""GPU: batched IMU preintegration kernel against the oracle (float32 state, tolerance 1e-5 relative as stated in
oracle/imu_oracle.cc)."""
import numpy as np
import pytest

import imu_synth
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_batched_preintegration_matches_oracle(pkg):
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    s, ft = imu_synth.make_stream(n_frames=40, seed=2)
    rng = np.random.default_rng(0)
    bias = (rng.normal(size=(39, 6)) * [0.02, 0.02, 0.02, 0.002, 0.002, 0.002]).astype(np.float32)
    out = ctx.imu_preintegrate(s, ft[:-1], ft[1:], bias, imu_synth.NOISE)
    for j in range(39):
        r = ol.imu_preintegrate(s, ft[j], ft[j + 1], bias[j], imu_synth.NOISE)
        assert out[j]["n_steps"] == r["n_steps"] and out[j]["n_consumed"] == r["n_consumed"]
        for f in ("dT", "dR", "dV", "dP", "JRg", "JVg", "JVa", "JPg", "JPa", "avgA", "avgW"):
            a, b = np.asarray(out[j][f], np.float64), np.asarray(r[f], np.float64)
            assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1.0), (j, f)
        a, b = out[j]["C"].astype(np.float64), r["C"].astype(np.float64)
        assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max()
    ctx.close()


def test_queue_edge_cases(pkg):
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    s, ft = imu_synth.make_stream(n_frames=3, seed=1)
    out = ctx.imu_preintegrate(s[:1], ft[0], ft[1], np.zeros(6, np.float32), imu_synth.NOISE)
    assert out[0]["n_steps"] == 0 and np.array_equal(out[0]["dR"].reshape(3, 3), np.eye(3, dtype=np.float32))
    ctx.close()
