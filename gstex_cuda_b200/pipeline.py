"""Fused multi-view training step: SH colours -> project/AABB/count -> scan -> key emit -> radix sort ->
tile ranges -> pack -> rasterise forward -> image loss -> rasterise backward -> per-Gaussian epilogue ->
SH backward, for every view of the rank's share of a batch, with NO host synchronisation inside the step.

This is the host side of the hot path for BASELINE configs 4 and 5.  It replaces, per view,
the ~60 torch/extension launches and the blocking ``.item()`` of the reference trainer (example.py:121-209,
utils.py:58) with ~15 launches of libgstex_b200 kernels over preallocated buffers.  Every device operation of the step
goes through the C ABI - kernels are counted in ``launches``, the few stream-ordered memsets (texel-gradient buffer and
loss once per step, tile counters once per view) in ``fills``; no torch op runs inside ``step()``:

* binning is the fused bucket-by-tile + per-tile shared-memory sort of csrc/binning_tiles.cu (bit-identical ids and
  tile ranges, no cumulative sum, no global 64-bit sort); the intersection count never leaves the device, the id
  buffer has a fixed capacity (``max_intersects``; by default n * tiles when that is small - which cannot overflow -
  else 16 n) and what does not fit is dropped: the binning kernel keeps the running maximum of the count on the device
  and ``check_overflow()`` (called by the trainer every ``check_every`` steps and by ``loss_value()``) raises if it
  ever exceeded the capacity;
* texels are read as one aligned float4: a (X,3) texture is padded once per step and its gradients un-padded once per
  step; a texture stored as (X,4) (``texture_rgba=True``: channels r, g, b, unused) needs neither pass - the
  rasterisers read it and accumulate its gradient in place;
* parameter gradients of all views accumulate in ONE contiguous fp32 arena (``grad_arena``), which is what
  the data-parallel wrapper all-reduces over NCCL (one collective per step);
* the step is a two-stream software pipeline over the views: the two rasteriser kernels (instruction-issue bound, most
  of the HBM bandwidth idle) run back to back on the caller's stream, while everything bandwidth-bound - SH colours,
  projection, binning and record packing of the NEXT view, the per-Gaussian epilogue and SH backward of the PREVIOUS
  one, texture padding and zero-fills - runs underneath them on a side stream over double-buffered per-view state.
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
                 texture_is_raw: bool = False, texture_rgba: bool = False):
        """``params`` holds the ACTIVATED parameters the rasteriser consumes.  Colours come from ``sh_coeffs``
        (N,K,3) through clamp(SH + 0.5, 0, 1) or, when ``params`` has ``colors`` (N,3) instead,
        are used as given (example.py:162).  ``grad_views`` places named gradients (e.g. ``v_means``,
        ``v_sh_coeffs``, ``v_texture``) in caller-owned tensors instead of this object's arena (trainer.py keeps
        them in its raw-parameter gradient arena; ``loss`` places the loss accumulator).  ``texture_is_raw``:
        ``params["texture"]`` holds pre-sigmoid texels; the sigmoid and its VJP are fused into the padding / un-padding
        passes (example.py:171).  ``texture_rgba``: ``params["texture"]`` is (X,4) - three channels stored at a 16-byte
        pitch - and so is its gradient ``v_texture``; no padding passes run."""
        self.lib = _lib.load()
        self.p = params
        dev = params["means"].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainStep needs CUDA tensors (there is no CPU path)")
        self.dev = dev
        self.use_sh = "sh_coeffs" in params
        self.texture_is_raw = bool(texture_is_raw)
        self.texture_rgba = bool(texture_rgba)
        for k in ("means", "scales", "quats", "opacities", "sh_coeffs" if self.use_sh else "colors", "uv0", "umap",
                  "vmap", "texture"):
            t = params[k]
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError(f"{k} must be a contiguous float32 CUDA tensor")
        self.texture_dims = texture_dims.contiguous()
        self.n = params["means"].shape[0]
        self.X, self.C = params["texture"].shape
        if self.texture_rgba:
            if self.C != 4 or self.texture_is_raw:
                raise RuntimeError("texture_rgba needs an activated (X,4) texture")
            self.C = 3  # three channels at a float4 pitch
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
        # default capacity: the exact bound n * tiles when that is small (cannot overflow), else 16 n
        self.cap = int(max_intersects if max_intersects is not None
                       else min(self.n * self.num_tiles, max(16 * self.n, 1 << 20)))
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        n, X, C, H, W = self.n, self.X, self.C, self.H, self.W
        self.background = (background if background is not None else torch.zeros(3, **f32)).contiguous()

        # ---- gradient arena: [means 3 | scales 3 | quats 4 | opacity 1 | uv0 2 | umap 3 | vmap 3 | sh 3K | texture C*X/n]
        gv = dict(grad_views or {})
        col_name, col_shape = ("v_sh_coeffs", (n, self.K, 3)) if self.use_sh else ("v_colors", (n, 3))
        shapes = ([(name, (n, w)) for name, w in _GRAD_FIELDS] +
                  [(col_name, col_shape), ("v_texture", (X, 4 if self.texture_rgba else C)), ("loss", (1,))])
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
        if self.texture_rgba:  # read and accumulated in place
            self.tex4, self.vtex4 = params["texture"], self.grads["v_texture"]
        else:
            self.tex4 = torch.empty((X, 4), **f32) if C == 3 else None
            self.vtex4 = torch.empty((X, 4), **f32) if C == 3 else None
        # the loss accumulator is the last slot of the gradient arena: the data-parallel all-reduce sums it with the
        # gradients in ONE collective
        self.loss = self.grads["loss"]
        self.has_grad_views = bool(gv)
        # optional per-step visibility counts (n floats, caller-owned): += 1 per view in which the Gaussian hits a tile
        self.visible = gv.get("visible")
        if self.visible is not None and not (self.visible.is_cuda and self.visible.dtype == torch.float32
                                             and self.visible.numel() == n and self.visible.is_contiguous()):
            raise RuntimeError("grad_views['visible'] must be a contiguous float32 CUDA tensor of n elements")
        # ---- per-view buffers.  Everything the side stream produces for a view (and the moment lines it consumes)
        # exists twice, so that view k+1 can be prepared while view k is still being rasterised.
        self.sets = [self._make_view_set(f32, i32) for _ in range(2)]
        self.cur = self.sets[0]  # the set of the view rendered last
        self.masks = torch.empty((self.cap, 8), **i32)  # blend masks: forward -> backward (csrc/raster.cuh)
        self.out = dict(out_img=torch.empty((H, W, 3), **f32), out_depth=torch.empty((H, W), **f32),
                        out_reg=torch.empty((H, W), **f32), out_texture=torch.empty((H, W, C), **f32),
                        out_normal=torch.empty((H, W, 3), **f32), final_Ts=torch.empty((H, W), **f32),
                        final_idx=torch.empty((H, W), **i32), depth_idx=torch.empty((H, W), **i32),
                        out_reg_s=torch.empty((H, W, 3), **f32))
        self.vout = dict(v_out_img=torch.empty((H, W, 3), **f32), v_out_depth=torch.empty((H, W), **f32),
                         v_out_reg=torch.empty((H, W), **f32), v_out_alpha=torch.empty((H, W), **f32),
                         v_out_texture=torch.empty((H, W, C), **f32), v_out_normal=torch.empty((H, W, 3), **f32))
        # what view_loss() leaves for the backward: the gradients of the outputs its loss does not use are None (zeros)
        self.vout_loss = dict(self.vout, v_out_img=None, v_out_depth=None, v_out_alpha=None)
        self._cur_vout = self.vout
        self.max_count_seen = torch.zeros(1, **i32)
        # bookkeeping for bench.py: number of libgstex_b200 kernels launched, optional per-kernel CUDA events
        self.launches = 0
        self.time_kernels = False
        self.kernel_events: List[Tuple[str, torch.cuda.Event, torch.cuda.Event]] = []
        self._bin_launches = 5  # tile count, tile scan, scatter, two per-tile sort size classes (+ one memset)
        self.fills = 0  # stream-ordered memsets issued through the C ABI (not kernels: not part of `launches`)
        # The rasterisers are issue-bound and leave most of the HBM bandwidth idle; the bandwidth-bound housekeeping that
        # does not depend on them (texture padding, zero-fills of the moment lines / texel-gradient buffer) runs on a
        # side stream underneath binning and the forward rasteriser and is joined with events where its result is needed.
        # default priority: measured on C5 (8 views/step, 1 GPU) a high-priority side stream only moves time from its own
        # kernels into the rasterisers they displace (32.7 ms vs 32.4 ms per step)
        self.side = torch.cuda.Stream(device=dev)
        # the exposed head of a step's first view (_prepare): SH colours + record packing / the binning chain
        self.side2 = torch.cuda.Stream(device=dev)
        self.side_hi = torch.cuda.Stream(device=dev, priority=-1)

    def _make_view_set(self, f32, i32):
        n, dev = self.n, self.dev

        class _ViewSet:
            pass

        v = _ViewSet()
        if self.use_sh:  # per-view colours and their gradient (they feed this view's SH backward)
            v.colors, v.mask = torch.empty((n, 3), **f32), torch.empty((n,), dtype=torch.uint8, device=dev)
            v.v_colors = torch.empty((n, 3), **f32)
        else:
            v.colors, v.mask, v.v_colors = self.p["colors"], None, self.grads["v_colors"]
        v.centers, v.extents, v.depths = torch.empty((n, 2), **f32), torch.empty((n, 2), **f32), torch.empty((n,), **f32)
        v.nth = torch.empty((n,), **i32)
        v.ids_sorted = torch.empty((self.cap,), **i32)
        v.num_isect = torch.zeros((1,), **i32)
        v.bin_temp = torch.empty((self.lib.gstex_bin_tiles_temp_bytes(self.num_tiles, self.cap),), dtype=torch.uint8,
                                 device=dev)
        v.tile_bins = torch.empty((self.num_tiles, 2), **i32)
        v.recs, v.mean2d, v.acc = torch.empty((n, 32), **f32), torch.empty((n, 2), **f32), torch.empty((n, 32), **f32)
        v.prep_done, v.bwd_done = torch.cuda.Event(), torch.cuda.Event()
        return v

    # buffers of the view rendered last (bench.py, tools/ and the tests read them)
    num_isect = property(lambda self: self.cur.num_isect)
    ids_sorted = property(lambda self: self.cur.ids_sorted)
    tile_bins = property(lambda self: self.cur.tile_bins)
    recs = property(lambda self: self.cur.recs)
    acc = property(lambda self: self.cur.acc)
    colors = property(lambda self: self.cur.colors)

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

    def _zero(self, t: torch.Tensor) -> None:
        """Stream-ordered zero fill through the C ABI (a cudaMemsetAsync, counted in `fills`)."""
        self._ck(self.lib.gstex_fill_zero(t.data_ptr(), t.numel() * t.element_size(), self._s()), "fill_zero")
        self.fills += 1

    def begin_step(self) -> None:
        """Once per optimiser step: pad the texture (unless it is stored padded), clear the texel-gradient buffer and the
        loss."""
        lib = self.lib
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)  # the parameters (and last step's consumers of tex4 / vtex4) are settled on `main`
        with torch.cuda.stream(self.side):
            if self.C == 3 and not self.texture_rgba:
                pad = lib.gstex_sigmoid_pad_texture if self.texture_is_raw else lib.gstex_pad_texture
                self._ck(pad(self.X, self.p["texture"].data_ptr(), self.tex4.data_ptr(), self._s()), "pad_texture")
                self.launches += 1
            self._zero(self.vtex4 if self.C == 3 else self.grads["v_texture"])
        self._zero(self.loss)
        if self.visible is not None:
            self._zero(self.visible)
        self._first_view = True

    def _prepare(self, v, viewmat: torch.Tensor, c2w: torch.Tensor, first: bool) -> None:
        """Head of a view, ordered after the CURRENT stream's work: SH colours, record packing (which also clears the
        view's moment lines), project / AABB / count, fused tile binning.  Records ``v.prep_done`` on the current stream.

        ``first`` (the step's first view: nothing else is running) splits the head into two independent chains on helper
        streams - [SH colours -> packing] moves 0.8 GB at HBM speed, [project -> binning] is bound by L2 atomics, a one-CTA
        scan and the per-tile sorts.  The binning chain gets the higher priority so that its small kernels do not queue
        behind the thousands of CTAs of the two bandwidth kernels (C4: head 0.356 -> 0.329 ms).  The heads of later views
        already run underneath the previous view's rasterisers; they stay one in-order chain (two more streams competing
        with the rasterisers measured 0.8 % slower on 64 views)."""
        lib, p = self.lib, self.p
        n, H, W, bw = self.n, self.H, self.W, self.bw
        fx, fy, cx, cy = self.intr
        P = lambda t: t.data_ptr()  # noqa: E731
        cur = torch.cuda.current_stream(self.dev)

        def colours_and_records(s):
            if self.use_sh:
                self._ck(lib.gstex_sh_colors_forward(n, self.sh_degree, self.sh_degree, P(p["means"]), P(c2w),
                                                     P(p["sh_coeffs"]), P(v.colors), P(v.mask), s), "sh_colors_forward")
            self._ck(lib.gstex_pack_records(n, P(self.texture_dims), P(v.colors), P(p["opacities"]), P(p["means"]),
                                            P(p["scales"]), self.glob_scale, P(p["quats"]), P(p["uv0"]), P(p["umap"]),
                                            P(p["vmap"]), P(viewmat), P(c2w), fx, fy, cx, cy, P(v.recs), P(v.mean2d),
                                            P(v.acc), s), "pack_records")  # also clears the view's moment lines

        def project_and_bin(s):
            self._ck(lib.gstex_project_aabb_count(n, P(p["means"]), P(p["scales"]), self.glob_scale, P(p["quats"]),
                                                  P(viewmat), fx, fy, cx, cy, H, W, bw, P(v.centers), P(v.extents),
                                                  P(v.depths), P(v.nth), 0 if self.visible is None else P(self.visible), s),
                     "project_aabb_count")
            # fused tile binning: bucket by tile + per-tile sort; the intersection count stays on the device
            self._ck(lib.gstex_bin_tiles(n, P(v.centers), P(v.extents), P(v.depths), self.tiles_x, self.tiles_y, bw,
                                         self.cap, P(v.ids_sorted), 0, P(v.tile_bins), P(v.num_isect),
                                         P(self.max_count_seen), P(v.bin_temp), v.bin_temp.numel(), s), "bin_tiles")

        if first:
            self.side2.wait_stream(cur)
            self.side_hi.wait_stream(cur)
            with torch.cuda.stream(self.side2):
                colours_and_records(self._s())
            with torch.cuda.stream(self.side_hi):
                project_and_bin(self._s())
            cur.wait_stream(self.side2)
            cur.wait_stream(self.side_hi)
        else:
            colours_and_records(self._s())
            project_and_bin(self._s())
        v.prep_done.record(cur)
        self.launches += int(self.use_sh) + 1 + self._bin_launches + 1  # sh, project, binning, pack

    def _render(self, v, viewmat: torch.Tensor, c2w: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Rasterise forward of a prepared view on the current stream."""
        lib, s = self.lib, self._s()
        fx, fy, cx, cy = self.intr
        P = lambda t: t.data_ptr()  # noqa: E731
        tex = self.tex4 if self.C == 3 else self.p["texture"]
        o = self.out
        with self._timed("raster_forward"):
            self._ck(lib.gstex_raster_forward(self.H, self.W, self.bw, self.C, self.settings, P(v.ids_sorted),
                                              P(v.tile_bins), P(v.recs), P(v.mean2d), P(tex), P(viewmat), P(c2w), fx, fy,
                                              cx, cy, P(self.background), P(o["out_img"]), P(o["out_depth"]),
                                              P(o["out_reg"]), P(o["out_texture"]), P(o["out_normal"]), P(o["final_Ts"]),
                                              P(o["final_idx"]), P(o["depth_idx"]), P(o["out_reg_s"]), P(self.masks),
                                              self.cap, P(v.num_isect), s), "raster_forward")
        self.launches += 2  # mask zero-fill + raster
        self.cur = v
        self._cur_vout = self.vout  # a caller that fills self.vout itself fills all six; view_loss() narrows it
        return o

    def _raster_backward(self, v, viewmat: torch.Tensor, c2w: torch.Tensor, vout: Dict[str, torch.Tensor]) -> None:
        lib, s = self.lib, self._s()
        fx, fy, cx, cy = self.intr
        P = lambda t: 0 if t is None else t.data_ptr()  # noqa: E731  (None: an upstream gradient that is all zeros)
        o = self.out
        tex = self.tex4 if self.C == 3 else self.p["texture"]
        vtex = self.vtex4 if self.C == 3 else self.grads["v_texture"]
        with self._timed("raster_backward"):
            self._ck(lib.gstex_raster_backward(self.H, self.W, self.bw, self.C, self.settings, P(v.ids_sorted),
                                               P(v.tile_bins), P(v.recs), P(v.mean2d), P(tex), P(viewmat), P(c2w), fx, fy,
                                               cx, cy, P(self.background), P(o["final_Ts"]), P(o["final_idx"]),
                                               P(o["depth_idx"]), P(o["out_reg_s"]), P(vout["v_out_img"]),
                                               P(vout["v_out_depth"]), P(vout["v_out_reg"]), P(vout["v_out_alpha"]),
                                               P(vout["v_out_texture"]), P(vout["v_out_normal"]), P(self.masks), P(v.acc),
                                               P(vtex), s), "raster_backward")
        v.bwd_done.record(torch.cuda.current_stream(self.dev))
        self.launches += 1

    def _tail(self, v, viewmat: torch.Tensor, c2w: torch.Tensor, first: bool) -> None:
        """Bandwidth-bound tail of a view on the CURRENT stream: per-Gaussian epilogue (moments -> parameter gradients,
        accumulated into the arena) and SH backward."""
        lib, s, p, g = self.lib, self._s(), self.p, self.grads
        fx, fy, cx, cy = self.intr
        P = lambda t: t.data_ptr()  # noqa: E731
        # geometry gradients accumulate over the views (bit 0); the colour gradient too when colours are parameters, but a
        # view's SH colour gradient feeds only this view's SH backward and is overwritten (bit 1 clear)
        acc_flag = 0 if first else (1 if self.use_sh else 3)
        self._ck(lib.gstex_raster_epilogue(self.n, P(p["means"]), P(p["scales"]), self.glob_scale, P(p["quats"]),
                                           P(p["umap"]), P(p["vmap"]), P(viewmat), P(c2w), fx, fy, cx, cy, P(v.acc),
                                           P(v.recs), P(v.v_colors), P(g["v_opacity"]), P(g["v_means"]), P(g["v_scales"]),
                                           P(g["v_quats"]), P(g["v_uv0"]), P(g["v_umap"]), P(g["v_vmap"]), acc_flag, s),
                 "raster_epilogue")
        if self.use_sh:
            self._ck(lib.gstex_sh_colors_backward(self.n, self.sh_degree, self.sh_degree, P(p["means"]), P(c2w),
                                                  P(v.v_colors), P(v.mask), P(g["v_sh_coeffs"]), acc_flag & 1, s),
                     "sh_colors_backward")
        self.launches += 1 + int(self.use_sh)

    # ---- one view at a time (tests, tools, callers that inject their own loss) -------------------------------------
    def view_forward(self, viewmat: torch.Tensor, c2w: torch.Tensor) -> Dict[str, torch.Tensor]:
        """SH colours + binning + rasterise forward of one camera.  Returns the output buffers (reused per view)."""
        viewmat, c2w = self._cam(viewmat, "viewmat"), self._cam(c2w, "c2w")
        main = torch.cuda.current_stream(self.dev)
        v = self.sets[0]
        self.side.wait_stream(main)  # the previous view's backward / tail have consumed this set
        with torch.cuda.stream(self.side):
            self._prepare(v, viewmat, c2w, self._first_view)
        main.wait_event(v.prep_done)  # the side stream is in order: this also covers begin_step's padding / fills
        return self._render(v, viewmat, c2w)

    def view_loss(self, target: torch.Tensor) -> None:
        """example.py:189-209 loss, accumulated into self.loss; fills the upstream-gradient buffers."""
        o, v = self.out, self.vout
        P = lambda t: t.data_ptr()  # noqa: E731
        # the loss does not use out_img / out_depth / out_alpha: their zero gradients are neither written nor read (NULL)
        self._ck(self.lib.gstex_image_loss(self.H, self.W, P(o["out_texture"]), P(o["out_reg"]), P(o["out_normal"]),
                                           P(target), P(self.loss), 0, 0, P(v["v_out_reg"]), 0, P(v["v_out_texture"]),
                                           P(v["v_out_normal"]), self._s()), "image_loss")
        self._cur_vout = self.vout_loss
        self.launches += 1

    def view_backward(self, viewmat: torch.Tensor, c2w: torch.Tensor, vout: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """Rasterise backward + epilogue + SH backward of the view last rendered; accumulates into the arena."""
        viewmat, c2w = self._cam(viewmat, "viewmat"), self._cam(c2w, "c2w")
        v = self.cur
        self._raster_backward(v, viewmat, c2w, vout if vout is not None else self._cur_vout)
        self._tail(v, viewmat, c2w, self._first_view)
        self._first_view = False

    def end_step(self) -> None:
        """Once per step: un-pad the texel gradients into the arena (nothing to do for a (X,4) texture)."""
        if self.C == 3 and not self.texture_rgba:
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
        """forward + loss + backward for every (viewmat, c2w) of this rank; returns the summed loss (device).

        Two-stream pipeline: the caller's stream runs [rasterise forward, loss, rasterise backward] of view k while the
        side stream runs the head of view k+1 and the tail of view k-1 (see the module docstring).  Gradients reach the
        arena in view order (the tails are serialised on the side stream), exactly as in a serial loop."""
        cams = [(self._cam(a, "viewmat"), self._cam(b, "c2w")) for a, b in cameras]
        main, side = torch.cuda.current_stream(self.dev), self.side
        self.begin_step()
        nv = len(cams)
        if nv:
            with torch.cuda.stream(side):
                self._prepare(self.sets[0], *cams[0], True)
        for k in range(nv):
            v, (viewmat, c2w) = self.sets[k & 1], cams[k]
            if k + 1 < nv:  # head of the next view, underneath this view's rasterisers (its set was freed by tail k-1,
                with torch.cuda.stream(side):  # which precedes it on the side stream)
                    self._prepare(self.sets[(k + 1) & 1], *cams[k + 1], False)
            main.wait_event(v.prep_done)
            self._render(v, viewmat, c2w)
            self.view_loss(targets[k])
            self._raster_backward(v, viewmat, c2w, self._cur_vout)
            with torch.cuda.stream(side):  # tail of this view, underneath the next view's rasterisers
                side.wait_event(v.bwd_done)
                self._tail(v, viewmat, c2w, k == 0)
        main.wait_stream(side)
        if nv == 0:
            # a rank without views (fewer views than ranks, short last batch) contributes zeros: nothing above has
            # overwritten last step's gradients
            self._zero(self.grad_arena)
            for name, t in self.grads.items():
                if t.data_ptr() < self.grad_arena.data_ptr() or t.data_ptr() >= self.grad_arena.data_ptr() + 4 * max(1, self.grad_arena.numel()):
                    self._zero(t)
            return self.loss
        self._first_view = False
        self.end_step()
        return self.loss

    def loss_value(self) -> float:
        """Host value of the loss of the last step (synchronises); raises if any view so far overflowed the capacity."""
        val = float(self.loss.item())
        self.check_overflow()
        return val

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
        if getattr(inner, "has_grad_views", False):
            raise RuntimeError("DataParallelTrainStep all-reduces inner.grad_arena: gradients placed in caller-owned "
                               "grad_views would be left out (GStexTrainStep reduces its own arena)")
        self.inner, self.rank, self.world_size, self.group = inner, int(rank), int(world_size), group
        self.time_collective = False
        self.collective_events = []

    @staticmethod
    def shard(num_views: int, rank: int, world_size: int) -> List[int]:
        return [v for v in range(num_views) if v % world_size == rank]

    def step(self, cameras, targets) -> torch.Tensor:
        mine = self.shard(len(cameras), self.rank, self.world_size)
        loss = self.inner.step([cameras[i] for i in mine], [targets[i] for i in mine])
        if self.world_size > 1:
            import torch.distributed as dist

            # ONE collective per step: the loss accumulator is the last slot of the arena
            if self.time_collective:
                a = torch.cuda.Event(enable_timing=True)
                a.record()
            dist.all_reduce(self.inner.grad_arena, op=dist.ReduceOp.SUM, group=self.group)
            if self.time_collective:
                b = torch.cuda.Event(enable_timing=True)
                b.record()
                self.collective_events.append((a, b))
        return loss
