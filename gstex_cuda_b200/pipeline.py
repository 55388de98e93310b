"""Fused multi-view training step: SH colours -> project/AABB/count -> scan -> key emit -> radix sort ->
tile ranges -> pack -> rasterise forward -> image loss -> rasterise backward -> per-Gaussian epilogue ->
SH backward, for every view of the rank's share of a batch, with NO host synchronisation inside the step.

This is the host side of the hot path for BASELINE configs 4 and 5.  It replaces, per view,
the ~60 torch/extension launches and the blocking ``.item()`` of the reference trainer (example.py:121-209,
utils.py:58) with ~16 launches of libgstex_b200 kernels on one stream over preallocated buffers:

* binning is the fused bucket-by-tile + per-tile shared-memory sort of csrc/binning_tiles.cu (bit-identical ids and
  tile ranges, no cumulative sum, no global 64-bit sort); the intersection count never leaves the device, the id
  buffer has a fixed capacity (``max_intersects``) and what does not fit is dropped - ``check_overflow()`` reports
  it when the caller next synchronises anyway;
* the texture is padded to float4 once per step, texel gradients of all views accumulate in one padded
  buffer and are un-padded once per step;
* parameter gradients of all views accumulate in ONE contiguous fp32 arena (``grad_arena``), which is what
  the data-parallel wrapper all-reduces over NCCL (one collective per step).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

_GRAD_FIELDS = (("v_means", 3), ("v_scales", 3), ("v_quats", 4), ("v_opacity", 1), ("v_uv0", 2), ("v_umap", 3),
                ("v_vmap", 3))


class FusedTrainStep:
    """Owns the replicated scene parameters' gradient arena and all per-view scratch buffers."""

    def __init__(self, params: Dict[str, torch.Tensor], texture_dims: torch.Tensor, img_height: int, img_width: int,
                 *, intrins: Tuple[float, float, float, float], sh_degree: int = 3, block_width: int = 16,
                 settings: int = 1 << 8, glob_scale: float = 1.0, background: Optional[torch.Tensor] = None,
                 max_intersects: Optional[int] = None, grad_views: Optional[Dict[str, torch.Tensor]] = None,
                 texture_is_raw: bool = False):
        """``params`` holds the ACTIVATED parameters the rasteriser consumes.  Colours come from ``sh_coeffs``
        (N,K,3) through clamp(SH + 0.5, 0, 1) or, when ``params`` has ``colors`` (N,3) instead,
        are used as given (example.py:162).  ``grad_views`` places named gradients (e.g. ``v_means``,
        ``v_sh_coeffs``, ``v_texture``) in caller-owned tensors instead of this object's arena (trainer.py keeps
        them in its raw-parameter gradient arena).  ``texture_is_raw``: ``params["texture"]`` holds pre-sigmoid
        texels; the sigmoid and its VJP are fused into the padding / un-padding passes (example.py:171)."""
        self.lib = _lib.load()
        self.p = params
        dev = params["means"].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainStep needs CUDA tensors (there is no CPU path)")
        self.dev = dev
        self.use_sh = "sh_coeffs" in params
        self.texture_is_raw = bool(texture_is_raw)
        for k in ("means", "scales", "quats", "opacities", "sh_coeffs" if self.use_sh else "colors", "uv0", "umap",
                  "vmap", "texture"):
            t = params[k]
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError(f"{k} must be a contiguous float32 CUDA tensor")
        self.texture_dims = texture_dims.contiguous()
        self.n = params["means"].shape[0]
        self.X, self.C = params["texture"].shape
        self.H, self.W, self.bw = int(img_height), int(img_width), int(block_width)
        self.intr = tuple(float(v) for v in intrins)
        self.sh_degree = int(sh_degree)
        self.K = (self.sh_degree + 1) ** 2
        if self.use_sh:
            assert params["sh_coeffs"].shape == (self.n, self.K, 3)
        else:
            assert params["colors"].shape == (self.n, 3)
        if self.texture_is_raw and self.C != 3:
            raise RuntimeError("texture_is_raw needs a 3-channel texture (the sigmoid is fused into the float4 padding)")
        self.settings, self.glob_scale = int(settings), float(glob_scale)
        self.tiles_x, self.tiles_y = -(-self.W // self.bw), -(-self.H // self.bw)
        self.num_tiles = self.tiles_x * self.tiles_y
        self.end_bit = 32 + max(1, math.ceil(math.log2(max(2, self.num_tiles))))
        self.cap = int(max_intersects if max_intersects is not None else 8 * self.n)
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        n, X, C, H, W = self.n, self.X, self.C, self.H, self.W
        self.background = (background if background is not None else torch.zeros(3, **f32)).contiguous()

        # ---- gradient arena: [means 3 | scales 3 | quats 4 | opacity 1 | uv0 2 | umap 3 | vmap 3 | sh 3K | texture C*X/n]
        gv = dict(grad_views or {})
        col_name, col_shape = ("v_sh_coeffs", (n, self.K, 3)) if self.use_sh else ("v_colors", (n, 3))
        shapes = [(name, (n, w)) for name, w in _GRAD_FIELDS] + [(col_name, col_shape), ("v_texture", (X, C))]
        own = [(name, shp) for name, shp in shapes if name not in gv]
        # every field starts on a 256-byte boundary: the kernels use 8- and 16-byte vector accesses on some of them
        pad = lambda sz: -(-sz // 64) * 64  # noqa: E731
        self.grad_arena = torch.zeros(sum(pad(math.prod(shp)) for _, shp in own), **f32)
        self.grads: Dict[str, torch.Tensor] = {}
        off = 0
        for name, shp in own:
            sz = math.prod(shp)
            self.grads[name] = self.grad_arena[off:off + sz].view(*shp)
            off += pad(sz)
        for name, shp in shapes:
            if name in gv:
                t = gv[name]
                if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == math.prod(shp)):
                    raise RuntimeError(f"grad_views[{name!r}] must be a contiguous float32 CUDA tensor of {shp}")
                self.grads[name] = t.view(*shp)

        # ---- per-step buffers
        self.tex4 = torch.empty((X, 4), **f32) if C == 3 else None
        self.vtex4 = torch.empty((X, 4), **f32) if C == 3 else None
        self.loss = torch.zeros(1, **f32)
        # ---- per-view buffers
        if self.use_sh:  # per-view colours and their gradient (they feed this view's SH backward)
            self.colors, self.mask, self.v_colors = torch.empty((n, 3), **f32), torch.empty((n,), dtype=torch.uint8, device=dev), torch.empty((n, 3), **f32)
        else:
            self.colors, self.mask, self.v_colors = params["colors"], None, self.grads["v_colors"]
        self.centers, self.extents, self.depths = torch.empty((n, 2), **f32), torch.empty((n, 2), **f32), torch.empty((n,), **f32)
        self.nth = torch.empty((n,), **i32)
        self.ids_sorted = torch.empty((self.cap,), **i32)
        self.num_isect = torch.zeros((1,), **i32)
        self.masks = torch.empty((self.cap, 8), **i32)  # blend masks: forward -> backward (csrc/raster.cuh)
        self.bin_temp = torch.empty((self.lib.gstex_bin_tiles_temp_bytes(self.num_tiles, self.cap),), dtype=torch.uint8, device=dev)
        self.tile_bins = torch.empty((self.num_tiles, 2), **i32)
        self.recs, self.mean2d, self.acc = torch.empty((n, 32), **f32), torch.empty((n, 2), **f32), torch.empty((n, 32), **f32)
        self.out = dict(out_img=torch.empty((H, W, 3), **f32), out_depth=torch.empty((H, W), **f32),
                        out_reg=torch.empty((H, W), **f32), out_texture=torch.empty((H, W, C), **f32),
                        out_normal=torch.empty((H, W, 3), **f32), final_Ts=torch.empty((H, W), **f32),
                        final_idx=torch.empty((H, W), **i32), depth_idx=torch.empty((H, W), **i32),
                        out_reg_s=torch.empty((H, W, 3), **f32))
        self.vout = dict(v_out_img=torch.empty((H, W, 3), **f32), v_out_depth=torch.empty((H, W), **f32),
                         v_out_reg=torch.empty((H, W), **f32), v_out_alpha=torch.empty((H, W), **f32),
                         v_out_texture=torch.empty((H, W, C), **f32), v_out_normal=torch.empty((H, W, 3), **f32))
        self.max_count_seen = torch.zeros(1, **i32)
        # bookkeeping for bench.py: number of libgstex_b200 kernels launched, optional per-kernel CUDA events
        self.launches = 0
        self.time_kernels = False
        self.kernel_events: List[Tuple[str, torch.cuda.Event, torch.cuda.Event]] = []
        self._bin_launches = 6  # tile count, tile scan, scatter, three per-tile sort size classes
        # The rasterisers are issue-bound and leave most of the HBM bandwidth idle; the bandwidth-bound housekeeping that
        # does not depend on them (texture padding, zero-fills of the moment lines / texel-gradient buffer) runs on a
        # side stream underneath binning and the forward rasteriser and is joined with events where its result is needed.
        self.side = torch.cuda.Stream(device=dev)
        self._tex_ready = torch.cuda.Event()
        self._acc_ready = torch.cuda.Event()

    # ------------------------------------------------------------------------------------------
    def _s(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _ck(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed (code {rc}): {_lib.last_error()}")

    def _timed(self, name: str):
        """Context manager recording a CUDA event pair around one kernel launch (bench.py roofline)."""
        step = self

        class _T:
            def __enter__(self_inner):
                if step.time_kernels:
                    self_inner.a = torch.cuda.Event(enable_timing=True)
                    self_inner.a.record()

            def __exit__(self_inner, *exc):
                if step.time_kernels:
                    b = torch.cuda.Event(enable_timing=True)
                    b.record()
                    step.kernel_events.append((name, self_inner.a, b))
                return False

        return _T()

    def _cam(self, m: torch.Tensor, name: str) -> torch.Tensor:
        """Camera matrices are read through raw pointers as row-major 4x4 float32: reject anything else loudly
        (torch.linalg.inv, for one, returns column-major strides)."""
        if not (isinstance(m, torch.Tensor) and m.is_cuda and m.dtype == torch.float32 and tuple(m.shape) == (4, 4)):
            raise RuntimeError(f"{name} must be a float32 CUDA tensor of shape (4, 4)")
        if not m.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous (row-major); call .contiguous() on it")
        return m

    def begin_step(self) -> None:
        """Once per optimiser step: pad the texture, clear the texel-gradient buffer and the loss."""
        lib = self.lib
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)  # the parameters (and last step's consumers of tex4 / vtex4) are settled on `main`
        with torch.cuda.stream(self.side):
            if self.C == 3:
                pad = lib.gstex_sigmoid_pad_texture if self.texture_is_raw else lib.gstex_pad_texture
                self._ck(pad(self.X, self.p["texture"].data_ptr(), self.tex4.data_ptr(), self._s()), "pad_texture")
                self.launches += 1
                self.vtex4.zero_()
            else:
                self.grads["v_texture"].zero_()
            self._tex_ready.record(self.side)
        self.loss.zero_()
        self._first_view = True

    def view_forward(self, viewmat: torch.Tensor, c2w: torch.Tensor) -> Dict[str, torch.Tensor]:
        """SH colours + binning + rasterise forward of one camera.  Returns the output buffers (reused per view)."""
        lib, s, p = self.lib, self._s(), self.p
        n, H, W, bw = self.n, self.H, self.W, self.bw
        fx, fy, cx, cy = self.intr
        P = lambda t: t.data_ptr()  # noqa: E731
        viewmat, c2w = self._cam(viewmat, "viewmat"), self._cam(c2w, "c2w")
        # side stream: clear this view's moment lines (and per-view SH colour gradients) while the view is binned / rendered
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)  # the previous view's epilogue / SH backward have consumed them
        with torch.cuda.stream(self.side):
            self.acc.zero_()
            if self.use_sh and not self._first_view:
                self.v_colors.zero_()  # SH colours are per view (they feed this view's SH backward), never accumulated
            self._acc_ready.record(self.side)
        if self.use_sh:
            self._ck(lib.gstex_sh_colors_forward(n, self.sh_degree, self.sh_degree, P(p["means"]), P(c2w),
                                                 P(p["sh_coeffs"]), P(self.colors), P(self.mask), s), "sh_colors_forward")
        self._ck(lib.gstex_project_aabb_count(n, P(p["means"]), P(p["scales"]), self.glob_scale, P(p["quats"]), P(viewmat),
                                              fx, fy, cx, cy, H, W, bw, P(self.centers), P(self.extents), P(self.depths),
                                              P(self.nth), s), "project_aabb_count")
        # fused tile binning: bucket by tile + per-tile shared-memory sort; the intersection count stays on the device
        self._ck(lib.gstex_bin_tiles(n, P(self.centers), P(self.extents), P(self.depths), self.tiles_x, self.tiles_y, bw,
                                     self.cap, P(self.ids_sorted), 0, P(self.tile_bins), P(self.num_isect),
                                     P(self.bin_temp), self.bin_temp.numel(), s), "bin_tiles")
        torch.maximum(self.max_count_seen, self.num_isect, out=self.max_count_seen)
        self._ck(lib.gstex_pack_records(n, P(self.texture_dims), P(self.colors), P(p["opacities"]), P(p["means"]),
                                        P(p["scales"]), self.glob_scale, P(p["quats"]), P(p["uv0"]), P(p["umap"]),
                                        P(p["vmap"]), P(viewmat), P(c2w), fx, fy, cx, cy, P(self.recs), P(self.mean2d), s),
                 "pack_records")
        tex = self.tex4 if self.C == 3 else p["texture"]
        o = self.out
        if self._first_view:
            main.wait_event(self._tex_ready)  # padded texture (begin_step, side stream)
        with self._timed("raster_forward"):
            self._ck(lib.gstex_raster_forward(H, W, bw, self.C, self.settings, P(self.ids_sorted), P(self.tile_bins),
                                              P(self.recs), P(self.mean2d), P(tex), P(viewmat), P(c2w), fx, fy, cx, cy,
                                              P(self.background), P(o["out_img"]), P(o["out_depth"]), P(o["out_reg"]),
                                              P(o["out_texture"]), P(o["out_normal"]), P(o["final_Ts"]),
                                              P(o["final_idx"]), P(o["depth_idx"]), P(o["out_reg_s"]), P(self.masks), self.cap,
                                              P(self.num_isect), s), "raster_forward")
        self.launches += int(self.use_sh) + 1 + self._bin_launches + 1 + 2  # sh, project, binning, pack, mask zero-fill + raster
        return o

    def view_loss(self, target: torch.Tensor) -> None:
        """example.py:189-209 loss, accumulated into self.loss; fills the upstream-gradient buffers."""
        o, v = self.out, self.vout
        P = lambda t: t.data_ptr()  # noqa: E731
        self._ck(self.lib.gstex_image_loss(self.H, self.W, P(o["out_texture"]), P(o["out_reg"]), P(o["out_normal"]),
                                           P(target), P(self.loss), P(v["v_out_img"]), P(v["v_out_depth"]),
                                           P(v["v_out_reg"]), P(v["v_out_alpha"]), P(v["v_out_texture"]),
                                           P(v["v_out_normal"]), self._s()), "image_loss")
        self.launches += 1

    def view_backward(self, viewmat: torch.Tensor, c2w: torch.Tensor, vout: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """Rasterise backward + epilogue + SH backward of the view last rendered; accumulates into the arena."""
        lib, s, p = self.lib, self._s(), self.p
        n, H, W, bw = self.n, self.H, self.W, self.bw
        fx, fy, cx, cy = self.intr
        P = lambda t: t.data_ptr()  # noqa: E731
        v = vout if vout is not None else self.vout
        viewmat, c2w = self._cam(viewmat, "viewmat"), self._cam(c2w, "c2w")
        o, g = self.out, self.grads
        acc_flag = 0 if self._first_view else 1
        torch.cuda.current_stream(self.dev).wait_event(self._acc_ready)  # zero-fills of view_forward (side stream)
        tex = self.tex4 if self.C == 3 else p["texture"]
        vtex = self.vtex4 if self.C == 3 else g["v_texture"]
        with self._timed("raster_backward"):
            self._ck(lib.gstex_raster_backward(H, W, bw, self.C, self.settings, P(self.ids_sorted), P(self.tile_bins),
                                               P(self.recs), P(self.mean2d), P(tex), P(viewmat), P(c2w), fx, fy, cx, cy,
                                               P(self.background), P(o["final_Ts"]), P(o["final_idx"]),
                                               P(o["depth_idx"]), P(o["out_reg_s"]), P(v["v_out_img"]),
                                               P(v["v_out_depth"]), P(v["v_out_reg"]), P(v["v_out_alpha"]),
                                               P(v["v_out_texture"]), P(v["v_out_normal"]), P(self.masks), P(self.acc), P(vtex), s),
                     "raster_backward")
        self.launches += 2 + int(self.use_sh)  # raster backward, epilogue, SH backward
        self._ck(lib.gstex_raster_epilogue(n, P(p["means"]), P(p["scales"]), self.glob_scale, P(p["quats"]), P(p["umap"]),
                                           P(p["vmap"]), P(viewmat), P(c2w), fx, fy, cx, cy, P(self.acc), P(self.v_colors),
                                           P(g["v_opacity"]), P(g["v_means"]), P(g["v_scales"]), P(g["v_quats"]),
                                           P(g["v_uv0"]), P(g["v_umap"]), P(g["v_vmap"]), acc_flag, s), "raster_epilogue")
        if self.use_sh:
            self._ck(lib.gstex_sh_colors_backward(n, self.sh_degree, self.sh_degree, P(p["means"]), P(c2w),
                                                  P(self.v_colors), P(self.mask), P(g["v_sh_coeffs"]), acc_flag, s),
                     "sh_colors_backward")
        self._first_view = False

    def end_step(self) -> None:
        """Once per step: un-pad the texel gradients into the arena."""
        if self.C == 3:
            if self.texture_is_raw:  # gradient w.r.t. the pre-sigmoid texels: g * t (1 - t), fused into the un-padding
                self._ck(self.lib.gstex_unpad_texture_grad_sigmoid(self.X, self.vtex4.data_ptr(), self.tex4.data_ptr(),
                                                                   self.grads["v_texture"].data_ptr(), 0, self._s()),
                         "unpad_texture_grad_sigmoid")
            else:
                self._ck(self.lib.gstex_unpad_texture_grad(self.X, self.vtex4.data_ptr(),
                                                           self.grads["v_texture"].data_ptr(), 0, self._s()),
                         "unpad_texture_grad")
            self.launches += 1

    def step(self, cameras: Sequence[Tuple[torch.Tensor, torch.Tensor]], targets: Sequence[torch.Tensor]) -> torch.Tensor:
        """forward + loss + backward for every (viewmat, c2w) of this rank; returns the summed loss (device)."""
        self.begin_step()
        for (viewmat, c2w), target in zip(cameras, targets):
            self.view_forward(viewmat, c2w)
            self.view_loss(target)
            self.view_backward(viewmat, c2w)
        self.end_step()
        return self.loss

    def check_overflow(self) -> int:
        """Largest intersection count seen so far (synchronises); raises if it exceeded the buffers."""
        m = int(self.max_count_seen.item())
        if m > self.cap:
            raise RuntimeError(f"{m} tile intersections exceed the preallocated capacity {self.cap}; "
                               f"construct FusedTrainStep with a larger max_intersects")
        return m


class DataParallelTrainStep:
    """View-batch data parallelism: replicated parameters, views sharded round-robin over the
    ranks, ONE all-reduce of the contiguous gradient arena per step (NCCL on GPUs, gloo in the CPU tests of the
    sharding logic)."""

    def __init__(self, inner, rank: int, world_size: int, group=None):
        self.inner, self.rank, self.world_size, self.group = inner, int(rank), int(world_size), group

    @staticmethod
    def shard(num_views: int, rank: int, world_size: int) -> List[int]:
        return [v for v in range(num_views) if v % world_size == rank]

    def step(self, cameras, targets) -> torch.Tensor:
        mine = self.shard(len(cameras), self.rank, self.world_size)
        loss = self.inner.step([cameras[i] for i in mine], [targets[i] for i in mine])
        if self.world_size > 1:
            import torch.distributed as dist

            dist.all_reduce(self.inner.grad_arena, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss
