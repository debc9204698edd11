"""CPU oracle for the PPG-SLAM front-end hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package.  The product (ppg_slam_b200/, include/) never does.

Parity status: the reference ships no tests or golden vectors for this path, so
parity is UNPINNED by the reference's own tests.  What pins this oracle instead:
  * L0 (networks): net_ref.py is checked here against torch.jit running the reference's own
    net/*.pt (tests/golden/make_golden.py writes the fixtures from the TorchScript modules).
  * the four OpenCV routines the extractor calls (initUndistortRectifyMap, undistortPoints,
    fisheye::undistortPoints, remap) are restated in ppg_oracle.c and checked bit-for-bit against
    cv2 4.13 in tests/test_oracle_cv.py.
  * everything else is a line-by-line restatement of feature/src/PPGExtractor.cpp and
    matching/src/Matcher.cpp with the file:line each function follows.
"""
