import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from ppg_slam_b200 import cameras, capi
from tests.parity_util import diff_records, oracle_post
cam = cameras.EUROC
H, W = cam.height, cam.width
rs = np.random.RandomState(0)
desc = rs.normal(size=(256, H // 8, W // 8)).astype(np.float32)
heat = np.zeros((H, W), np.float32)
for name, prob in (("uniform", np.full((H, W), 0.5, np.float32)),
                   ("ramp", np.linspace(0.1, 0.9, H * W).astype(np.float32).reshape(H, W)),
                   ("noise", (rs.rand(H, W) * 0.9 + 0.05).astype(np.float32))):
    e = capi.Extractor(cam, max_batch=1)
    t0 = time.time()
    got = e.run_from_maps(prob[None], heat[None], desc[None], allow_capacity=True)[0]
    t1 = time.time()
    ref = oracle_post(cam, prob, heat, desc)
    print(name, "gpu s", round(t1 - t0, 3), "rounds", got["nms_rounds"], "status", got["status"], "n_kp", got["n_kp"],
          "diff", diff_records(got, ref)[:2], flush=True)
    e.close()
