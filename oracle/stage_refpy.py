"""TEST INFRASTRUCTURE ONLY (called by oracle/build_ref.sh): stages the reference's Python op wrappers as ONE archive
next to the extensions they bind -- oracle/_ref/refpy.zip, importable through zipimport.

tests/test_gpu_dropin.py runs these files UNMODIFIED twice, over this repo's drop-in modules and over the reference
extensions built by build_ref.sh, and compares the results.  Like the .so files the archive
is a build output: git-ignored, shipped to the GPU box by gpurun, never part of the repository's history.

    python oracle/stage_refpy.py /root/reference oracle/_ref/refpy.zip
"""
import sys
import zipfile

FILES = ["pointnet2_lib/pointnet2/pointnet2_utils.py", "pointnet2_lib/pointnet2/pointnet2_modules.py",
         "pointnet2_lib/pointnet2/pytorch_utils.py", "lib/utils/iou3d/iou3d_utils.py",
         "lib/utils/roipool3d/roipool3d_utils.py", "lib/utils/kitti_utils.py", "lib/utils/object3d.py"]


def main(ref, out):
    with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED) as z:
        dirs = set()
        for f in FILES:
            z.write(ref + "/" + f, f)
            parts = f.split("/")[:-1]
            for k in range(1, len(parts) + 1):
                dirs.add("/".join(parts[:k]))
        for d in sorted(dirs):
            z.writestr(d + "/__init__.py", "")   # generated: the reference tree relies on namespace packages
    print("build_ref: wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
