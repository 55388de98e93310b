"""Stages the reference's UNMODIFIED pure-Python layer (gstex_cuda/*.py and example.py) under the git-ignored
baseline/_ref/, so that the GPU box (which has no /root/reference) can run the reference's own Python over this repo's
backend: tests/test_gpu_reference_python.py rebinds `gstex_cuda.cuda` to `gstex_cuda_b200.cuda` and runs example.py's
trainer.  Test infrastructure only; nothing is copied into tracked files.  The compiled extension is NOT staged here -
oracle/build_ref.py builds it into oracle/_ref/."""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["__init__.py", "_torch_impl.py", "get_aabb_2d.py", "sh.py", "texture.py", "texture_edit.py", "texture_sample.py",
         "timer.py", "utils.py"]


def stage() -> str:
    if not os.path.isdir(os.path.join(SRC, "gstex_cuda")):
        raise RuntimeError(f"{SRC}/gstex_cuda not found (the reference exists in the build container only)")
    os.makedirs(os.path.join(DST, "gstex_cuda"), exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, "gstex_cuda", f), os.path.join(DST, "gstex_cuda", f))
    shutil.copyfile(os.path.join(SRC, "example.py"), os.path.join(DST, "example.py"))
    return DST


if __name__ == "__main__":
    print(stage())
