"""-m gpu: the reference's OWN, UNMODIFIED Python layer (gstex_cuda/texture.py, utils.py, get_aabb_2d.py, _torch_impl.py
and example.py, staged under the git-ignored baseline/_ref/ by oracle/stage_ref_python.py) running over THIS repo's
backend: `gstex_cuda.cuda` is rebound to `gstex_cuda_b200.cuda` - the one-line change INTEGRATION.md gives a maintainer
(route 2) - and example.py's trainer is run as upstream runs it.

  C1  `example.py --height 32 --width 32 --num_points 10 --iterations 10 --torch_compare True` (BASELINE config 1): the
      trainer itself asserts, every iteration, that our kernels' outputs and gradients equal the reference's pure-PyTorch
      rasteriser (torch.testing.assert_close defaults, example.py:270-275), through the reference's stateless
      texture_backward signature (texture.py:342).
  C2  the default image overfit for 30 iterations: the loss falls.
"""
import importlib
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def example():
    if not os.path.exists(os.path.join(REF, "gstex_cuda", "texture.py")):
        pytest.skip("reference Python layer not staged (python oracle/stage_ref_python.py in the build container)")
    import gstex_cuda_b200.cuda as backend

    for k in [k for k in sys.modules if k == "gstex_cuda" or k.startswith("gstex_cuda.") or k == "example"]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        pkg = importlib.import_module("gstex_cuda")
        assert os.path.realpath(os.path.dirname(pkg.__file__)) == os.path.realpath(os.path.join(REF, "gstex_cuda"))
        sys.modules["gstex_cuda.cuda"] = backend  # the rebinding: gstex_cuda/cuda/__init__.py -> gstex_cuda_b200.cuda
        pkg.cuda = backend
        mod = importlib.import_module("example")
        import gstex_cuda.texture as ref_texture
        assert ref_texture._C is backend
        yield mod
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "gstex_cuda" or k.startswith("gstex_cuda.") or k == "example"]:
            del sys.modules[k]


def _gt(height, width):
    gt = torch.ones((height, width, 3))
    gt[: height // 2, : width // 2, :] = torch.tensor([1.0, 0.0, 0.0])
    gt[height // 2:, width // 2:, :] = torch.tensor([0.0, 0.0, 1.0])
    return gt


def test_c1_torch_compare_through_the_reference_python(example, capsys):
    example.seed_everything(1)
    trainer = example.SimpleTrainer(gt_image=_gt(32, 32), num_points=10, num_texels=1000000)
    trainer.train(iterations=10, lr=1e-2, save_imgs=False, torch_compare=True)  # asserts CUDA == torch every iteration
    out = capsys.readouterr().out
    assert "Iteration 10/10" in out


def test_c2_default_overfit_through_the_reference_python(example, capsys):
    example.seed_everything(1)
    trainer = example.SimpleTrainer(gt_image=_gt(256, 256), num_points=100, num_texels=1000000)
    trainer.train(iterations=30, lr=1e-2, save_imgs=False, torch_compare=False)
    out = capsys.readouterr().out
    losses = [float(l.split("Loss:")[1]) for l in out.splitlines() if "Loss:" in l]
    assert len(losses) == 30 and losses[-1] < 0.8 * losses[0]
    print(f"  C2 through the reference's Python: loss {losses[0]:.5f} -> {losses[-1]:.5f}")
