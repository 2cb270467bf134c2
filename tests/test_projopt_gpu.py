"""GPU: vido_pose_opt_proj (PoseOptimizationNew / PoseOptimizationObjMot) against the oracle, several problems per launch."""
import numpy as np
import pytest

import oracle_lib as ol
import proj_synth

pytestmark = pytest.mark.gpu


def test_projection_only_optimisers_match_oracle(pkg):
    cases = [proj_synth.camera_case(seed=1)[0], proj_synth.object_case(seed=2)[0], proj_synth.camera_case(n=900, seed=5, outliers=0.3)[0],
             proj_synth.object_case(n=60, seed=7, outliers=0.0)[0], proj_synth.camera_case(n=2, seed=3)[0]]
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    got = ctx.pose_opt_proj(cases)
    for case, (T, inl, st) in zip(cases, got):
        d = dict(case)
        kind = d.pop("kind")
        T0, inl0, st0 = ol.pose_opt_proj(kind, d["obs_xy"], d["pts3d"], d["T_init"], K=d.get("K"), P=d.get("P"))
        assert st.iterations == st0.iterations and st.total_trials == st0.total_trials, (kind, st.iterations, st0.iterations)
        for (c1, l1, t1), (c2, l2, t2) in zip(st.records(), st0.records()):
            assert t1 == t2 and abs(c1 - c2) <= 1e-6 * max(abs(c2), 1e-9)
        assert np.abs(T - T0).max() <= 1e-4 * max(np.abs(T0).max(), 1.0)
        assert np.array_equal(inl, inl0)
    ctx.close()
