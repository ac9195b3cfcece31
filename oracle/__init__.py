"""
CPU oracle for the SerStacker stacking hot path (register -> warp -> accumulate).

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it, and only as the checker or
as the timed CPU baseline.  Nothing under serstacker_b200/ imports it.

It is a restatement in Python of the reference's C++ control flow; every pixel primitive the
reference delegates to OpenCV (cv::remap, cv::pyrDown, cv::sepFilter2D, cv::meanStdDev, cv::invert,
cv::solve, cv::erode, cv::morphologyEx, cv::resize, cv::GaussianBlur ...) is delegated here to the
same OpenCV through cv2 (4.13.0 in this image).  OpenCV itself is a third-party dependency of the
reference that is neither vendored nor version-pinned (CMakeLists.txt:47-59 requires >= 4.2).

PARITY UNPINNED: the reference has no tests, golden vectors or fixtures of any kind for this path
(SURVEY.md section 4), and its sources cannot be compiled in this image (no OpenCV C++ headers, TBB,
libconfig).  The oracle is pinned only by construction (same OpenCV primitives, same call order) and
by the committed fixtures under tests/golden/ that were generated from it.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
