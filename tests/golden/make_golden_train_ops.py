"""Golden vectors for the training-step glue, generated in the build container:

    python tests/golden/make_golden_train_ops.py        # needs /root/reference; writes train_ops.npz here

* preprocess: the statements of the reference's example.py:126-143 and :162-163, :171 restated with the SAME torch
  ops around the reference's own ``gstex_cuda._torch_impl.normalized_quat_to_rotmat`` (imported in place), forward
  values plus torch-autograd gradients of a random linear functional of the activated parameters.
* adam: ``torch.optim.Adam`` (the optimiser of example.py:223-225), three steps on a random parameter vector.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import gstex_cuda._torch_impl as _T  # noqa: E402  (the reference, imported in place)


def main():
    g = torch.Generator().manual_seed(21)
    N = 257
    raw_scales = (0.5 * math.log(1 / N) * torch.rand(N, 3, generator=g)).requires_grad_(True)
    raw_quats = (torch.randn(N, 4, generator=g) * 1.7).requires_grad_(True)  # deliberately NOT unit length
    raw_rgbs = torch.randn(N, 3, generator=g).requires_grad_(True)
    raw_opac = torch.randn(N, 1, generator=g).requires_grad_(True)
    mapping = torch.zeros(N, 1, 4)
    mapping[:, :, :2] = 0.5 + 0.1 * torch.randn(N, 1, 2, generator=g)
    mapping[:, :, 2] = math.log(0.3) + 0.2 * torch.randn(N, 1, generator=g)
    mapping[:, :, 3] = 2.0 * math.pi * torch.rand(N, 1, generator=g)
    mapping.requires_grad_(True)
    raw_texture = torch.randn(N * 6, 3, generator=g).requires_grad_(True)

    # example.py:126-137
    scales = torch.zeros_like(raw_scales)
    scales[:, :2] = torch.exp(raw_scales[:, :2])
    scales[:, -1] = 1e-5 * torch.mean(scales[:, :-1], dim=-1).detach()
    quats = raw_quats / raw_quats.norm(dim=-1, keepdim=True)
    Rs = _T.normalized_quat_to_rotmat(quats)
    uv0 = mapping[:, :, :2]
    uvscale = torch.exp(mapping[:, :, None, 2])
    theta = mapping[:, :, None, 3]
    ax1, ax2 = Rs[:, None, :, 0], Rs[:, None, :, 1]
    umap = uvscale * (ax1 * torch.cos(theta) + ax2 * torch.sin(theta))
    vmap = uvscale * (-ax1 * torch.sin(theta) + ax2 * torch.cos(theta))
    colors, opac, texture = torch.sigmoid(raw_rgbs), torch.sigmoid(raw_opac), torch.sigmoid(raw_texture)

    outs = dict(scales=scales, quats=quats, uv0=uv0, umap=umap, vmap=vmap, colors=colors, opacities=opac,
                texture=texture)
    up = {k: torch.randn(v.shape, generator=g) for k, v in outs.items()}
    up["scales"][:, 2] = 0.0  # the rasteriser never produces a thickness gradient
    loss = sum((outs[k] * up[k]).sum() for k in outs)
    loss.backward()
    d = dict(raw_scales=raw_scales, raw_quats=raw_quats, raw_rgbs=raw_rgbs, raw_opacities=raw_opac, mapping=mapping,
             raw_texture=raw_texture)
    npz = {k: v.detach().numpy() for k, v in d.items()}
    npz.update({k: v.detach().numpy() for k, v in outs.items()})
    npz.update({"v_" + k: v.numpy() for k, v in up.items()})
    npz.update({"g_" + k: v.grad.numpy() for k, v in d.items()})

    # torch.optim.Adam, example.py:223-225 (lr as given, default betas / eps)
    p = torch.randn(1003, generator=g).requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.01)
    npz["adam_p0"] = p.detach().numpy().copy()
    for t in range(3):
        grad = torch.randn(1003, generator=g) * (10.0 ** (t - 1))
        opt.zero_grad()
        p.grad = grad.clone()
        opt.step()
        npz[f"adam_g{t + 1}"] = grad.numpy()
        npz[f"adam_p{t + 1}"] = p.detach().numpy().copy()
    st = opt.state[p]
    npz["adam_m3"], npz["adam_v3"] = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
    np.savez_compressed(os.path.join(HERE, "train_ops.npz"), **npz)
    print("wrote train_ops.npz", {k: v.shape for k, v in npz.items() if k.startswith("g_")})


if __name__ == "__main__":
    main()
