"""ppg_slam_b200 — B200-native (sm_100a) front-end hot path of PPG-SLAM.

Only what the path needs lives here:
  csrc/      CUDA kernels + the C-ABI (include/ppg_b200.h) -> libppg_b200.so
  capi.py    ctypes binding of that C-ABI (tests / bench drive the product through it)
  weights/   flat fp32 export of the reference's net/*.pt (tools/export_weights.py)
  synth.py   seeded synthetic frames / association inputs of SURVEY.md §8(d)
  cameras.py the four shipped calibrations (config/*.yaml of the reference)

There is no CPU fallback: importing capi without the built library raises.
"""
__all__ = ["capi", "synth", "cameras", "weights_io"]
