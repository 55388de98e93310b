"""-m gpu: SH evaluation and texture sampling vs the reference's golden vectors and the oracle
(mirrors the reference's own tests/test_sh.py and tests/test_sample.py)."""
import os

import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200 import sh, texture_sample as TS, _torch_impl as _T
from gstex_cuda_b200 import cuda as _C
from gpu_util import DEV, to_np

pytestmark = pytest.mark.gpu


def test_sample_matches_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "torch_impl_sample.npz")))
    dims, tex, uvs = (torch.from_numpy(g[k]).to(DEV) for k in ("texture_dims", "texture", "uvs"))
    out = TS.texture_sample((5, 1, tex.shape[1]), dims, tex, uvs)
    torch.testing.assert_close(out.cpu(), torch.from_numpy(g["out"]))  # reference tests/test_sample.py:31-34
    torch.testing.assert_close(out, _T.sample_texture(dims, tex, uvs))


def test_sample_like_reference_test():
    """tests/test_sample.py:9-34 of the reference, seeded."""
    torch.manual_seed(0)
    sz, C, nq = 5, 10, 100
    dims = torch.stack([torch.randint(6, (sz,), dtype=torch.int32) + 2, torch.randint(7, (sz,), dtype=torch.int32) + 2,
                        torch.zeros((sz,), dtype=torch.int32)], dim=-1).to(DEV)
    hws = dims[:, 0] * dims[:, 1]
    dims[:, -1] = torch.cumsum(hws, 0) - hws
    tex = torch.rand((int(hws.sum()), C)).to(DEV)
    uvs = torch.rand((nq, 2)).to(DEV)
    qd = dims[torch.randint(sz, (nq,))].contiguous()
    fast = TS.texture_sample((sz, 1, C), qd, tex, uvs)
    torch.testing.assert_close(fast, _T.sample_texture(qd, tex, uvs))
    # scatter (transpose of the fetch) against the oracle and against torch autograd of the torch twin
    v = torch.randn(nq, C, device=DEV)
    vt = _C.texture_sample_backward((sz, 1, C), qd, uvs, tex, v)
    np.testing.assert_allclose(to_np(vt), oracle.texture_sample_backward(to_np(qd), to_np(uvs), to_np(tex), to_np(v)),
                               rtol=1e-5, atol=1e-6)
    t2 = tex.clone().requires_grad_(True)
    (_T.sample_texture(qd, t2, uvs) * v).sum().backward()
    torch.testing.assert_close(vt, t2.grad, rtol=1e-5, atol=1e-6)
    # the autograd default follows the reference: no gradient reaches the texture (texture_sample.py:71-78)
    t3 = tex.clone().requires_grad_(True)
    TS.texture_sample((sz, 1, C), qd, t3, uvs).sum().backward()
    assert t3.grad is None
    TS.texture_sample((sz, 1, C), qd, t3, uvs, texture_grad=True).sum().backward()
    assert t3.grad is not None


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_sh_matches_reference_golden(golden_dir, deg):
    g = dict(np.load(os.path.join(golden_dir, "torch_impl_sh.npz")))
    dirs = torch.from_numpy(g["viewdirs"]).to(DEV)
    coeffs = torch.from_numpy(g[f"coeffs{deg}"]).to(DEV).requires_grad_(True)
    colors = sh.spherical_harmonics(deg, dirs, coeffs)
    torch.testing.assert_close(colors.detach().cpu(), torch.from_numpy(g[f"colors{deg}"]))
    (colors * torch.from_numpy(g[f"v_colors{deg}"]).to(DEV)).sum().backward()
    torch.testing.assert_close(coeffs.grad.cpu(), torch.from_numpy(g[f"v_coeffs{deg}"]))


def test_sh_optimisation_like_reference_test():
    """tests/test_sh.py:9-46 of the reference (shortened to 200 Adam steps): CUDA colours and coefficient
    gradients equal torch autograd of the torch twin at every step."""
    torch.manual_seed(0)
    n, degree = 1, 4
    gt = torch.ones(n, 3, device=DEV) * 0.5
    dirs = torch.randn(n, 3, device=DEV)
    dirs /= torch.linalg.norm(dirs, dim=-1, keepdim=True)
    coeffs = torch.rand(n, sh.num_sh_bases(degree), 3, device=DEV, requires_grad=True)
    opt = torch.optim.Adam([coeffs], lr=1e-2)
    for _ in range(200):
        opt.zero_grad()
        check = _T.compute_sh_color(dirs, coeffs)
        torch.square(check - gt).mean().backward()
        check_grad = coeffs.grad.detach().clone()
        opt.zero_grad()
        colors = sh.spherical_harmonics(degree, dirs, coeffs)
        torch.square(colors - gt).mean().backward()
        torch.testing.assert_close(check_grad, coeffs.grad.detach())
        torch.testing.assert_close(check, colors)
        opt.step()


def test_sh_degrees_to_use_and_unnormalised_dirs():
    g = torch.Generator().manual_seed(4)
    n = 1000
    dirs = (torch.randn(n, 3, generator=g) * 3.0).to(DEV)  # the kernel normalises (sh.cuh:61-66)
    coeffs = torch.rand(n, 25, 3, generator=g).to(DEV)
    for use in range(5):
        got = _C.compute_sh_forward(n, 4, use, dirs, coeffs)
        np.testing.assert_allclose(to_np(got), oracle.sh_forward(4, use, to_np(dirs), to_np(coeffs)), rtol=2e-5, atol=2e-6)
        v = torch.randn(n, 3, generator=g).to(DEV)
        gb = _C.compute_sh_backward(n, 4, use, dirs, v)
        np.testing.assert_allclose(to_np(gb), oracle.sh_backward(4, use, to_np(dirs), to_np(v)), rtol=2e-5, atol=2e-6)
        assert float(gb[:, (use + 1) ** 2:].abs().max()) == 0.0 if use < 4 else True
    with pytest.raises(RuntimeError):
        _C.compute_sh_forward(n, 3, 3, dirs, coeffs)  # wrong number of bases (bindings.cu:27-30)
