"""Golden fixtures (see make_golden.py for how each file was produced)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def path(name: str) -> str:
    return os.path.join(HERE, name)


def load_cam0(type: str = None, inner: int = None):
    """The bundled cam0 data set + shipped config.cfg as a Problem (optionally another
    projection type / Inner_Constraints value)."""
    import feba_b200 as fb
    z = np.load(path("cam0_problem.npz"), allow_pickle=False)
    s = fb.Settings(**{k[2:]: (z[k].item() if z[k].dtype.kind != "U" else str(z[k]))
                       for k in z.files if k.startswith("s_")})
    if s.Meas_std_y is not None and np.isnan(s.Meas_std_y):
        s.Meas_std_y = None
    for k in ("Iteration_Cap", "Inner_Constraints", "Estimate_Xc", "Estimate_Yc", "Estimate_Zc",
              "Estimate_w", "Estimate_p", "Estimate_k", "Estimate_xp", "Estimate_yp", "Estimate_c",
              "Estimate_radial", "Num_Radial_Distortions", "Estimate_decent", "Estimate_tie",
              "Estimate_AllGCP", "Check_Points"):
        setattr(s, k, int(getattr(s, k)))
    if type is not None:
        s.type = type
    if inner is not None:
        s.Inner_Constraints = int(inner)
    return fb.Problem(settings=s, obs_x=z["obs_x"], obs_y=z["obs_y"], obs_img=z["obs_img"],
                      obs_pt=z["obs_pt"], img_cam=z["img_cam"], eop0=z["eop0"], iop0=z["iop0"],
                      cam_box=z["cam_box"], xyz0=z["xyz0"], pt_tie=z["pt_tie"], tie_pt=z["tie_pt"],
                      point_ids=[str(v) for v in z["point_ids"]],
                      image_ids=[str(v) for v in z["image_ids"]],
                      camera_ids=[str(v) for v in z["camera_ids"]])
