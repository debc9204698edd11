"""The calibrations the reference ships (config/*.yaml), as the extractor sees them.

Reference: system/src/System.cpp:45-70 builds `vCalibration`; for KannalaBrandt8 it reads
Camera.k0..k3 while every shipped YAML defines k1..k4, so k0 reads as 0 and D = (0,k1,k2,k3)
(SURVEY.md §3.1).  That quirk is reproduced here because it decides whether the heat-map remap
runs (feature/src/PPGExtractor.cpp:261).  Values are rounded to float32 exactly like
`std::vector<float> vCalibration`.
"""
import numpy as np


def _f32(*v):
    return [float(np.float32(x)) for x in v]


class Camera:
    def __init__(self, name, width, height, fx, fy, cx, cy, d, fisheye):
        self.name, self.width, self.height, self.fisheye = name, width, height, bool(fisheye)
        self.K = _f32(fx, 0, cx, 0, fy, cy, 0, 0, 1)
        self.D = _f32(*d)

    def __repr__(self):
        return "Camera(%s %dx%d %s)" % (self.name, self.width, self.height, "KB8" if self.fisheye else "pinhole")


# config/EuRoC.yaml:11-25
EUROC = Camera("EuRoC", 752, 480, 458.654, 457.296, 367.215, 248.375,
               (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05), False)
# config/TUM-VI.yaml:11-25  (k0 missing -> 0; k4 never read)
TUMVI = Camera("TUM-VI", 512, 512, 190.978477, 190.973307, 254.931706, 256.897442,
               (0.0, 0.003482389402, 0.000715034845, -0.002053236141), True)
# config/TUM-VI-1024.yaml:11-25
TUMVI1024 = Camera("TUM-VI-1024", 1024, 1024, 380.81042871360756, 380.81194179427075,
                   510.29465304840727, 514.3304630538506,
                   (0.0, 0.010171079892421483, -0.010816440029919381, 0.005942781769412756), True)
# config/UMA.yaml:11-25
UMA = Camera("UMA-VI", 1024, 768, 545.7402000594014, 546.4624873938807, 516.7898455908171,
             399.68834148863493,
             (0.0, -0.06983837053126551, 0.030679193251357234, -0.029318268716673087), True)

ALL = {c.name: c for c in (EUROC, TUMVI, TUMVI1024, UMA)}
