"""Recipe: compile the UNMODIFIED reference CUDA extension from where its sources lie.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this.

Sources are read in place from /root/reference/gstex_cuda/cuda/csrc (never copied into
this repo); the only outputs are the ninja build tree and ``gstex_ref_C.so`` under
``oracle/_ref/`` (git-ignored, but it travels to the GPU box with the gpurun snapshot).

Flags follow what upstream actually runs, i.e. the JIT flags of
``gstex_cuda/cuda/_backend.py:36-40`` (``-O3``, no ``--use_fast_math``; the fast-math
``setup.py`` build is never imported), with the arch pinned to sm_100.
This is a direct ninja/nvcc build through ``torch.utils.cpp_extension.load`` - the
reference's own build system (setup.py) is not run.

Usage:  python oracle/build_ref.py        (about 6 minutes, CPU only; no GPU needed)
"""
import glob
import os
import sys

REF_CSRC = "/root/reference/gstex_cuda/cuda/csrc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "gstex_ref_C"


def build(verbose: bool = True) -> str:
    if not os.path.isdir(REF_CSRC):
        raise FileNotFoundError(f"{REF_CSRC} not present (only exists in the build container)")
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    sources = sorted(glob.glob(os.path.join(REF_CSRC, "*.cu"))) + sorted(
        glob.glob(os.path.join(REF_CSRC, "*.cpp"))
    )
    load(
        name=NAME,
        sources=sources,
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3"],
        extra_include_paths=[os.path.join(REF_CSRC, "third_party/glm")],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=True,
    )
    so = os.path.join(OUT, NAME + ".so")
    assert os.path.exists(so), so
    return so


if __name__ == "__main__":
    print(build(verbose="-q" not in sys.argv))
