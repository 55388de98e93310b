"""One GStex optimiser step from RAW parameters, fused.

The reference trainer (example.py:121-225, :278) goes, every iteration, through ~25 small torch kernels of
"preprocess" (exp / normalise / quat->R / uv maps / sigmoids), the rasteriser, the same number of autograd kernels
on the way back, and a 7-tensor ``torch.optim.Adam`` step.  Here the whole iteration is

    preprocess_forward (1 launch) -> FusedTrainStep over this rank's views (pipeline.py; the texel sigmoid and its
    VJP ride on the float4 padding passes) -> preprocess_backward (1 launch) -> [one NCCL all-reduce of the raw
    gradient arena] -> adam_step (1 launch over the whole parameter arena)

with no host synchronisation.  Raw parameters, their gradients and both Adam moments live in four contiguous
fp32 arenas with the same field layout, so the collective and the optimiser each touch one buffer.

Because nothing in the iteration depends on the host - the intersection count stays on the device, and so does Adam's
step counter (``gstex_adam_step_device``) - a whole optimiser step can be captured into ONE CUDA graph
(``capture()`` / ``replay()``): the ~25 launches of a small scene (BASELINE configs 1-3 are launch-bound) become one
graph launch.

Parameter names follow example.py:69-119: ``means``, ``scales`` (log), ``quats`` (un-normalised), ``opacities``
(pre-sigmoid), ``mapping`` (N,1,4) = (u0, v0, log uv-scale, theta), ``texture`` (pre-sigmoid, (X,3)), and for the
colours either ``rgbs`` (pre-sigmoid, example.py:162) or ``sh_coeffs`` (N,K,3) with colours = clamp(SH + 0.5, 0, 1)
(SURVEY 8d C4).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .pipeline import DataParallelTrainStep, FusedTrainStep


class GStexTrainStep:
    def __init__(self, raw: Dict[str, torch.Tensor], texture_dims: torch.Tensor, img_height: int, img_width: int, *,
                 intrins: Tuple[float, float, float, float], sh_degree: int = 3, lr: float = 0.01,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0,
                 train_mapping: bool = False, rank: int = 0, world_size: int = 1, group=None, check_every: int = 100,
                 visible_only: bool = False, **fused_kwargs):
        """``raw``: the leaf parameters (copied into this object's arena; ``self.raw`` are views of it).
        ``train_mapping``: example.py:118 freezes the uv mapping (its gradient is computed, the update skipped).
        ``check_every``: every that many steps (and in ``loss_value()``) the running maximum of the intersection count
        is read back and compared with the preallocated capacity - an overflow would otherwise drop intersections
        silently, also under CUDA-graph replay (0 disables the periodic check).
        ``visible_only``: the optimiser touches only Gaussians that hit at least one tile in at least one view of the
        step (on any rank) - parameters and both Adam moments of the others stay as they are ("sparse /
        visible-only update").  The per-step visibility counts ride in the gradient arena, so the one all-reduce sums
        them too."""
        self.lib = _lib.load()
        dev = raw["means"].device
        if dev.type != "cuda":
            raise RuntimeError("GStexTrainStep needs CUDA tensors (there is no CPU path)")
        self.dev = dev
        self.use_sh = "sh_coeffs" in raw
        n = raw["means"].shape[0]
        X, C = raw["texture"].shape
        if C != 3:
            raise RuntimeError("GStexTrainStep trains 3-channel textures (example.py:95)")
        K = (int(sh_degree) + 1) ** 2
        col = ("sh_coeffs", (n, K, 3)) if self.use_sh else ("rgbs", (n, 3))
        # the mapping sits after the trained fields so that a frozen mapping is simply left out of the Adam range; the last
        # slot is the loss accumulator (no parameter: only its gradient-arena twin is used), so that the data-parallel
        # all-reduce of the gradient arena sums the loss in the same collective
        self.fields: List[Tuple[str, Tuple[int, ...]]] = [
            ("means", (n, 3)), ("scales", (n, 3)), ("quats", (n, 4)), ("opacities", (n, 1)), col, ("texture", (X, 3)),
            ("mapping", (n, 1, 4)), ("loss", (1,)), ("visible", (n,))]
        pad = lambda sz: -(-sz // 64) * 64  # noqa: E731  (fields start on 256-byte boundaries: vector accesses)
        total = sum(pad(math.prod(shp)) for _, shp in self.fields)
        f32 = dict(dtype=torch.float32, device=dev)
        self.param_arena, self.grad_arena = torch.zeros(total, **f32), torch.zeros(total, **f32)
        self.exp_avg, self.exp_avg_sq = torch.zeros(total, **f32), torch.zeros(total, **f32)
        self.raw: Dict[str, torch.Tensor] = {}
        self.raw_grads: Dict[str, torch.Tensor] = {}
        off = 0
        for name, shp in self.fields:
            sz = math.prod(shp)
            self.raw_grads[name] = self.grad_arena[off:off + sz].view(*shp)
            if name not in ("loss", "visible"):
                if tuple(raw[name].shape) != shp:
                    raise RuntimeError(f"raw[{name!r}] must have shape {shp}, got {tuple(raw[name].shape)}")
                self.raw[name] = self.param_arena[off:off + sz].view(*shp)
                self.raw[name].copy_(raw[name])
            off += pad(sz)
        self.n_train = total - pad(n) - pad(1) - (0 if train_mapping else pad(n * 4))
        self.visible_only, self.train_mapping = bool(visible_only), bool(train_mapping)
        # texel -> Gaussian for the visible-only update: a uniform layout (every Gaussian owns th*tw consecutive texels,
        # example.py:139-143) needs no table
        self._tex_unit, self._tex_owner = 3, None
        if self.visible_only:
            d = texture_dims.to(torch.int64)
            per = d[:, 0] * d[:, 1]
            if n > 0 and bool((per == per[0]).all()) and bool((d[:, 2] == torch.arange(n, device=d.device) * per[0]).all()) \
                    and int(per[0]) * n == X:
                self._tex_unit = 3 * int(per[0])
            else:
                owner = torch.full((X,), 0, dtype=torch.int32, device=dev)
                idx = torch.repeat_interleave(torch.arange(n, device=dev), per)
                pos = torch.repeat_interleave(d[:, 2], per) + (torch.arange(int(per.sum()), device=dev)
                                                                - torch.repeat_interleave(torch.cumsum(per, 0) - per, per))
                owner[pos] = idx.to(torch.int32)
                self._tex_owner = owner
        self.check_every = int(check_every)
        self.n, self.X = n, X
        self.lr, self.betas, self.eps, self.grad_scale = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(grad_scale)
        self.rank, self.world_size, self.group = int(rank), int(world_size), group
        self.step_count = 0  # host mirror of the device-resident step counter (adam_state)
        self.adam_state = torch.zeros(self.lib.gstex_adam_state_bytes(), dtype=torch.uint8, device=dev)
        self._graph = None
        self._graph_loss = None

        # activated parameters (what the rasteriser consumes) and the gradients w.r.t. them
        self.act = dict(scales=torch.empty((n, 3), **f32), quats=torch.empty((n, 4), **f32),
                        opacities=torch.empty((n, 1), **f32), uv0=torch.empty((n, 1, 2), **f32),
                        umap=torch.empty((n, 1, 3), **f32), vmap=torch.empty((n, 1, 3), **f32))
        params = dict(self.act, means=self.raw["means"], texture=self.raw["texture"])
        grad_views = dict(v_means=self.raw_grads["means"], v_texture=self.raw_grads["texture"], loss=self.raw_grads["loss"])
        if self.visible_only:
            grad_views["visible"] = self.raw_grads["visible"]
        if self.use_sh:
            params["sh_coeffs"] = self.raw["sh_coeffs"]
            grad_views["v_sh_coeffs"] = self.raw_grads["sh_coeffs"]
        else:
            self.act["colors"] = torch.empty((n, 3), **f32)
            params["colors"] = self.act["colors"]
        self.fused = FusedTrainStep(params, texture_dims, img_height, img_width, intrins=intrins, sh_degree=sh_degree,
                                    grad_views=grad_views, texture_is_raw=True, **fused_kwargs)
        self.launches = 0

    def _s(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def preprocess(self) -> None:
        """raw -> activated parameters (example.py:126-143, :162-163); one launch."""
        P = lambda t: t.data_ptr()  # noqa: E731
        r, a = self.raw, self.act
        _lib.check(self.lib.gstex_preprocess_forward(
            self.n, P(r["scales"]), P(r["quats"]), P(r["mapping"]), 0 if self.use_sh else P(r["rgbs"]), P(r["opacities"]),
            P(a["scales"]), P(a["quats"]), P(a["uv0"]), P(a["umap"]), P(a["vmap"]), 0 if self.use_sh else P(a["colors"]),
            P(a["opacities"]), self._s()), "preprocess_forward")
        self.launches += 1

    def backward_preprocess(self) -> None:
        """gradients w.r.t. activated parameters (FusedTrainStep arena) -> raw gradient arena; one launch."""
        P = lambda t: t.data_ptr()  # noqa: E731
        r, g, rg = self.raw, self.fused.grads, self.raw_grads
        _lib.check(self.lib.gstex_preprocess_backward(
            self.n, P(r["scales"]), P(r["quats"]), P(r["mapping"]), 0 if self.use_sh else P(r["rgbs"]), P(r["opacities"]),
            P(g["v_scales"]), P(g["v_quats"]), P(g["v_uv0"]), P(g["v_umap"]), P(g["v_vmap"]),
            0 if self.use_sh else P(g["v_colors"]), P(g["v_opacity"]), P(rg["scales"]), P(rg["quats"]), P(rg["mapping"]),
            0 if self.use_sh else P(rg["rgbs"]), P(rg["opacities"]), self._s()), "preprocess_backward")
        self.launches += 1

    def forward_backward(self, cameras: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                         targets: Sequence[torch.Tensor]) -> torch.Tensor:
        """Loss (summed over this rank's views, device tensor) and raw gradients (``self.raw_grads``, summed over
        all ranks) of the batch; no parameter update."""
        mine = DataParallelTrainStep.shard(len(cameras), self.rank, self.world_size)
        self.preprocess()
        loss = self.fused.step([cameras[i] for i in mine], [targets[i] for i in mine])
        if mine:
            self.backward_preprocess()
        else:
            self.grad_arena.zero_()
        if self.world_size > 1:
            import torch.distributed as dist

            dist.all_reduce(self.grad_arena, op=dist.ReduceOp.SUM, group=self.group)  # the loss rides in its last slot
        return loss

    def optimizer_step(self) -> None:
        """torch.optim.Adam's update over the parameter arena (example.py:223-225, :278); the step counter and its bias
        corrections live on the device, so the call is the same every iteration (and CUDA-graph replayable)."""
        self.step_count += 1
        P = lambda t: t.data_ptr()  # noqa: E731
        if not self.visible_only:
            _lib.check(self.lib.gstex_adam_step_device(self.n_train, P(self.param_arena), P(self.grad_arena), P(self.exp_avg),
                                                       P(self.exp_avg_sq), self.lr, self.betas[0], self.betas[1], self.eps,
                                                       P(self.adam_state), self.grad_scale, self._s()), "adam_step")
            self.launches += 2
            return
        # visible-only: one counter advance, then one row-gated launch per field of the arena
        _lib.check(self.lib.gstex_adam_prepare_device(P(self.adam_state), self.lr, self.betas[0], self.betas[1], self.eps,
                                                      self.grad_scale, self._s()), "adam_prepare")
        vis = self.raw_grads["visible"]
        base = self.param_arena.data_ptr()
        for name, shp in self.fields:
            if name in ("loss", "visible") or (name == "mapping" and not self.train_mapping):
                continue
            t = self.raw[name]
            off = t.data_ptr() - base
            unit, owner = (self._tex_unit, self._tex_owner) if name == "texture" else (t.numel() // self.n, None)
            _lib.check(self.lib.gstex_adam_apply_rows_device(
                t.numel(), t.data_ptr(), self.grad_arena.data_ptr() + off, self.exp_avg.data_ptr() + off,
                self.exp_avg_sq.data_ptr() + off, P(self.adam_state), P(vis), unit, 0 if owner is None else P(owner),
                self._s()), f"adam_apply_rows[{name}]")
            self.launches += 1
        self.launches += 1

    def step(self, cameras, targets) -> torch.Tensor:
        loss = self.forward_backward(cameras, targets)
        self.optimizer_step()
        self._periodic_check()
        return loss

    def _periodic_check(self) -> None:
        if self.check_every > 0 and self.step_count % self.check_every == 0 and not torch.cuda.is_current_stream_capturing():
            self.fused.check_overflow()

    def loss_value(self) -> float:
        """Host value of the last step's loss (synchronises) after checking that no view overflowed the capacity."""
        self.fused.check_overflow()
        return float(self.raw_grads["loss"].item())

    # ---- CUDA graph of one whole optimiser step ------------------------------------------------------------------
    def capture(self, cameras: Sequence[Tuple[torch.Tensor, torch.Tensor]], targets: Sequence[torch.Tensor]) -> None:
        """Capture ``step(cameras, targets)`` into a CUDA graph.  The camera matrices and target images are captured BY
        ADDRESS: write new views into the same tensors (``copy_``) before a ``replay()``.  One eager warm-up step runs
        first (kernel loading, allocator) and its effect on the parameters and optimiser state is rolled back, so
        capture() itself does not train.  Single-process only (the NCCL all-reduce is left out of graphs here)."""
        if self.world_size > 1:
            raise RuntimeError("GStexTrainStep.capture(): graph capture is single-process; with world_size > 1 call step()")
        saved = [t.clone() for t in (self.param_arena, self.exp_avg, self.exp_avg_sq, self.adam_state)]
        count, launches, fused_launches = self.step_count, self.launches, self.fused.launches
        self.step(cameras, targets)
        self.fused.check_overflow()
        for dst, src in zip((self.param_arena, self.exp_avg, self.exp_avg_sq, self.adam_state), saved):
            dst.copy_(src)
        self.step_count = count
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self.step(cameras, targets)
        # capturing enqueues nothing: undo the bookkeeping of the captured call
        self.step_count, self.launches, self.fused.launches = count, launches, fused_launches
        self._graph, self._graph_loss = graph, loss

    def replay(self) -> torch.Tensor:
        """One optimiser step = one graph launch.  Returns the (device) loss tensor of the captured step."""
        if self._graph is None:
            raise RuntimeError("GStexTrainStep.replay() before capture()")
        self._graph.replay()
        self.step_count += 1
        self._periodic_check()
        return self._graph_loss

    @property
    def total_launches(self) -> int:
        return self.launches + self.fused.launches
