"""ORACLE build recipe — test infrastructure only.

* ``build_oracle()``: gcc -> oracle/liboracle.so (voxelize.c + nms.c, fp contraction
  off, no fast-math).
* ``build_ref()``: when /root/reference is present (the build container only),
  compile the reference's rotated-NMS CUDA source *where it lies*
  (mmdet/ops/iou3d/src/iou3d_kernel.cu, unmodified, sm_100a) into
  oracle/_ref/libiou3d_ref.so.  No reference source is copied into the repo;
  oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot.
  The other reference native code on this path (spconv v1.0) is third-party and
  absent; the numba voxelizer is Python and cannot travel (its outputs are
  committed as tests/golden/*.npz instead).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/mmdet/ops/iou3d/src/iou3d_kernel.cu"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_oracle(force=False):
    out = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, "voxelize.c"), os.path.join(HERE, "nms.c")]
    if force or _newer(out, srcs):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
               "-o", out] + srcs + ["-lm"]
        subprocess.check_call(cmd)
    return out


def ref_path():
    return os.path.join(HERE, "_ref", "libiou3d_ref.so")


def build_ref(force=False):
    """Returns the path of the reference NMS library, or None if it can neither be
    built (no /root/reference) nor found prebuilt."""
    out = ref_path()
    if os.path.isfile(REF_SRC) and (force or _newer(out, [REF_SRC])):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cmd = ["nvcc", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
               "-Xcompiler", "-fPIC", "-o", out, REF_SRC]
        subprocess.check_call(cmd)
    return out if os.path.isfile(out) else None


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_ref(force=True))
