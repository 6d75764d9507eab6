"""TEST INFRASTRUCTURE: builds oracle/liboracle.so (+ oracle/_ref when /root/reference exists)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libiou3d_ref.so")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("voxel_oracle.c", "iou3d_oracle.c", "Makefile")]
    if force or _stale(LIB, srcs):
        subprocess.run(["make", "-C", HERE, "liboracle.so"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    ref_src = "/root/reference/det3d/ops/iou3d/src/iou3d_kernel.cu"
    if os.path.exists(ref_src) and (force or not os.path.exists(REF_LIB)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
