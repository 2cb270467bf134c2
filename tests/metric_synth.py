"""Ground truth of the synthetic scene in the conventions of Tracking::GetMetricError (src/Tracking.cc:3531-3674): camera poses
relative to the first frame (Map::vmCameraPose_GT), and per estimated object motion the object's previous pose (vmObjPosePre) and
its body-frame motion (vmRigidMotion_GT)."""
import numpy as np


def ground_truth(sc, frames, objects_of_frame):
    """objects_of_frame(f) -> semantic labels of the estimated object motions of frame f >= 1, in Map order"""
    T0i = np.linalg.inv(frames[0]["Twc"].numpy())
    cam_gt = np.stack([T0i @ f["Twc"].numpy() for f in frames]).astype(np.float32)
    pre, mgt = [], []
    for f in range(1, len(frames)):
        for s in objects_of_frame(f):
            j = int(s) - 1
            L0 = T0i @ sc.object_pose(j, f - 1).numpy()
            L1 = T0i @ sc.object_pose(j, f).numpy()
            pre.append(L0)
            mgt.append(np.linalg.inv(L0) @ L1)
    if not pre:
        return cam_gt, None, None
    return cam_gt, np.stack(pre).astype(np.float32), np.stack(mgt).astype(np.float32)
