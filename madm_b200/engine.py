"""Python-side owner of one ``madm_ctx``: registers parameter pointers, keeps the packed-weight arena and the
workspace (torch tensors used purely as device memory), tracks parameter versions / the active LoRA adapter, and
calls ``madm_extract`` on torch's current stream.  No PyTorch compute on the path; no fallback."""
import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MadmExtractArgs, MadmTensor, STAGE_ALL, STAGE_ALL_S0, STAGE_DEC, STAGE_HEAD, STAGE_PROJ

TAP_SHAPES = ((512, 128), (320, 64), (640, 32), (1280, 16))  # enc tap, unet taps (C, HW side)
OUT_SIDES = (128, 64, 32, 16)                                  # s2..s5
# variant -> per-output (C, side) and the first "tap" (base: encoder tap; s0: the decoded image)
OUT_SHAPES = {"base": ((512, 128), (512, 64), (512, 32), (512, 16)), "s0": ((128, 512), (512, 64), (512, 32), (512, 16))}


class Engine:
    def __init__(self, device: torch.device, compute_dtype: str = "fp16", variant: str = "base"):
        if device.type != "cuda":
            raise _lib.MadmError("madm_b200 runs on sm_100a CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        _lib.check(self.lib.madm_create(C.byref(h), self.index), None, "madm_create")
        self.ctx = h
        if compute_dtype not in ("fp16", "bf16"):
            raise _lib.MadmError(f"compute_dtype must be 'fp16' or 'bf16', got {compute_dtype!r}")
        self.compute_dtype = compute_dtype
        _lib.check(self.lib.madm_set_compute_dtype(self.ctx, _lib.DTYPE_FP16 if compute_dtype == "fp16" else _lib.DTYPE_BF16),
                   self.ctx, "madm_set_compute_dtype")
        if variant not in OUT_SHAPES:
            raise _lib.MadmError(f"variant must be 'base' or 's0', got {variant!r}")
        self.variant = variant
        _lib.check(self.lib.madm_set_variant(self.ctx, _lib.VARIANT_S0 if variant == "s0" else _lib.VARIANT_BASE), self.ctx, "madm_set_variant")
        self.stage_all = STAGE_ALL_S0 if variant == "s0" else STAGE_ALL
        self._named: List[Tuple[str, torch.Tensor]] = []
        self._named_src = None  # the caller's list object behind the last bind (fast path: compare data pointers only)
        self._ptrs = None
        self._sig = None
        self._versions = None
        self._packed: Optional[torch.Tensor] = None
        self._packed_adapter: Optional[str] = "\0unset"
        self._ws: Optional[torch.Tensor] = None
        self._keep = []
        # image_im2col raises this device flag when a normalised image leaves [-1, 1] (the reference asserts that with a host sync,
        # ldm_diffusers.py:147).  It is copied to pinned host memory asynchronously after every call and looked at on the NEXT call
        # (or by check_input_range()), so steady-state inference keeps running without a synchronisation.
        self.range_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._range_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._range_event: Optional[torch.cuda.Event] = None
        # CUDA graphs: at small batch the ~550 launches of one forward are CPU-launch-bound, and even at B = 8 the graph saves the
        # inter-kernel launch gaps (24.9 -> 24.1 ms per step); one graph per call signature
        # replays them.  graph_max_batch: largest B that is graphed (0 disables).
        self.graph_max_batch = 32  # covers the sliding-window crop batches (9, 18, 21 crops per call)
        self._graphs: Dict[Tuple, Dict[str, object]] = {}

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.madm_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def bind(self, named: Sequence[Tuple[str, torch.Tensor]]):
        """(Re)register fp32 CUDA parameter tensors under their reference state_dict names."""
        if named is self._named_src:  # the caller's cached list: only the storage pointers can have moved
            ptrs = tuple(t.data_ptr() for _, t in named)
            if ptrs == self._ptrs:
                return False
        src = named
        named = [(n, t) for n, t in named]
        sig = tuple((n, t.data_ptr(), tuple(t.shape)) for n, t in named)
        if sig == self._sig:
            self._named_src, self._ptrs = src, tuple(p for _, p, _ in sig)
            return False
        for n, t in named:
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.MadmError(f"parameter {n} must be a contiguous fp32 tensor on {self.device} (got {t.dtype} on {t.device})")
        arr = (MadmTensor * len(named))()
        names = [n.encode() for n, _ in named]
        for i, (n, t) in enumerate(named):
            arr[i].name = names[i]
            arr[i].data = t.data_ptr()
            arr[i].ndim = t.dim()
            for k, s in enumerate(t.shape):
                arr[i].shape[k] = s
        _lib.check(self.lib.madm_set_tensors(self.ctx, arr, len(named)), self.ctx, "madm_set_tensors")
        self._graphs.clear()  # captured graphs hold the old parameter pointers (biases / norm affines are read in place)
        self._named, self._sig = named, sig
        self._named_src, self._ptrs = src, tuple(p for _, p, _ in sig)
        self._versions = None  # force a full repack
        return True

    def _version_vector(self):
        return tuple(t._version for _, t in self._named)

    def only_trainables_changed(self, old, new) -> bool:
        """True if every tensor whose version moved between two version vectors belongs to the LoRA training step's trainable set
        (LoRA factors, feature_projections / ema_feature_projections) and nothing else did — then a partial repack suffices."""
        if old is None or len(old) != len(new):
            return False
        for (n, _), a, b in zip(self._named, old, new):
            if a != b and not (".lora_A." in n or ".lora_B." in n or n.startswith("feature_projections.") or n.startswith("ema_feature_projections.")):
                return False
        return True

    def ensure_packed(self, adapter: Optional[str], scaling: float):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        need = self.lib.madm_packed_bytes(self.ctx)
        if need == 0:
            _lib.check(-1, self.ctx, "madm_packed_bytes")
        if self._packed is None or self._packed.numel() != need:
            self._packed = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._versions = None
        vers = self._version_vector()
        ad = (adapter or "").encode()
        if vers != self._versions:
            # a LoRA training step moves only LoRA factors and projection weights (optimizer step, EMA update): repack just those
            mode = 2 if self.only_trainables_changed(self._versions, vers) else 0
            _lib.check(self.lib.madm_pack_weights(self.ctx, C.c_void_p(self._packed.data_ptr()), ad, scaling, mode, st), self.ctx,
                       "madm_pack_weights")
            self._versions, self._packed_adapter = vers, adapter
        elif adapter != self._packed_adapter:  # adapter switch: re-fold only the 128 LoRA-targeted projections
            _lib.check(self.lib.madm_pack_weights(self.ctx, C.c_void_p(self._packed.data_ptr()), ad, scaling, 1, st), self.ctx,
                       "madm_pack_weights(lora_only)")
            self._packed_adapter = adapter

    def workspace(self, B: int, head_hw: Tuple[int, int] = (0, 0)) -> torch.Tensor:
        need = self.lib.madm_workspace_bytes_head(self.ctx, B, head_hw[0], head_hw[1])
        if need == 0:
            _lib.check(-1, self.ctx, "madm_workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._graphs.clear()  # captured graphs hold the old workspace pointer
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------ input range guard (ldm_diffusers.py:147)
    def _range_poll(self, wait: bool = False):
        ev = self._range_event
        if ev is None:
            return
        if wait:
            ev.synchronize()
        elif not ev.query():
            return
        self._range_event = None
        if int(self._range_host[0]) != 0:
            self._range_host.zero_()
            self.range_flag.zero_()
            raise _lib.MadmError("input image outside [0, 1] (normalised: outside [-1, 1]) in an earlier madm_extract call: the reference "
                                 "asserts `batched_inputs['img'].min() >= -1.0 and .max() <= 1.0` (ldm_diffusers.py:147)")

    def _range_publish(self):
        self._range_host.copy_(self.range_flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._range_event = ev

    def check_input_range(self):
        """Wait for the last call's range flag and raise if any image since the last check left the reference's input range."""
        self._range_poll(wait=True)

    def extract_graphed(self, img, cond_inputs, cond_emb, timesteps, shared_noise, *, ema=False, stages=STAGE_ALL, want_taps=False,
                        want_latents=False, want_final=False, out_dtype=torch.float32):
        """`extract` through a captured CUDA graph (static input / output buffers, one graph per call signature)."""
        B = img.shape[0]
        self._range_poll()
        key = (B, bool(ema), stages, bool(want_taps), bool(want_latents), bool(want_final), self._packed.data_ptr(), shared_noise.data_ptr(),
               out_dtype)
        g = self._graphs.get(key)
        if g is None:
            st = dict(img=torch.empty_like(img, dtype=torch.float32), cond_inputs=torch.empty(B, 77, 768, device=self.device),
                      cond_emb=torch.empty(B, 1280, device=self.device), timesteps=torch.zeros(B, dtype=torch.int64, device=self.device))
            for k, v in (("img", img), ("cond_inputs", cond_inputs), ("cond_emb", cond_emb), ("timesteps", timesteps)):
                st[k].copy_(v)
            kw = dict(ema=ema, stages=stages, want_taps=want_taps, want_latents=want_latents, want_final=want_final, out_dtype=out_dtype)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture: plan build, cudaFuncSetAttribute, allocations
                self.extract(st["img"], st["cond_inputs"], st["cond_emb"], st["timesteps"], shared_noise, **kw)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                res = self.extract(st["img"], st["cond_inputs"], st["cond_emb"], st["timesteps"], shared_noise, **kw)
            g = dict(graph=graph, static=st, res=res)
            self._graphs[key] = g
        st = g["static"]
        st["img"].copy_(img)
        st["cond_inputs"].copy_(cond_inputs)
        st["cond_emb"].copy_(cond_emb)
        st["timesteps"].copy_(timesteps)
        g["graph"].replay()
        if stages & _lib.STAGE_VAE:
            self._range_publish()
        out = {}
        for k, v in g["res"].items():  # fresh tensors, like the eager path: the static buffers are overwritten by the next replay
            out[k] = [t.clone() for t in v] if isinstance(v, list) else v.clone()
        return out

    # ------------------------------------------------------------------ DAFormer head
    def head(self, feats: Sequence[torch.Tensor], num_classes: int) -> torch.Tensor:
        """``MADM_STAGE_HEAD`` on the feature dict: s2..s5 fp32 NCHW [B,512,h/r,w/r] (r = 1, 2, 4, 8) -> logits [B,num_classes,h,w];
        h = w = 128 for a 512^2 crop, larger for the merged maps of sliding-window inference (s0 variant: s0 [B,128,h,w] with
        h = w = 512 for a crop, s3..s5 at 1/8, 1/16, 1/32 -> logits on the s0 grid)."""
        if self._packed is None:
            raise _lib.MadmError("Engine.head called before ensure_packed()")
        dev = self.device
        B = feats[0].shape[0]
        a = MadmExtractArgs()
        a.B, a.stages, a.ema = B, STAGE_HEAD, 0
        keep = []
        h0, w0 = int(feats[0].shape[2]), int(feats[0].shape[3])
        ratios = (1, 8, 16, 32) if self.variant == "s0" else (1, 2, 4, 8)
        for i, (ch, _) in enumerate(OUT_SHAPES[self.variant]):
            t = feats[i]
            want = (B, ch, h0 // ratios[i], w0 // ratios[i])
            if t.device != dev or tuple(t.shape) != want or h0 % ratios[3] or w0 % ratios[3]:
                raise _lib.MadmError(f"feature map {i} must be {list(want)} on {dev}, got {tuple(t.shape)} on {t.device}")
            t = t.to(torch.float32).contiguous()
            keep.append(t)
            a.out[i] = t.data_ptr()
        side0 = OUT_SHAPES[self.variant][0][1]
        head_hw = (0, 0) if (h0, w0) == (side0, side0) else (h0, w0)
        a.head_h, a.head_w = head_hw
        logits = torch.empty(B, num_classes, h0, w0, dtype=torch.float32, device=dev)
        a.logits = logits.data_ptr()
        ws = self.workspace(B, head_hw)
        a.packed = self._packed.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.madm_extract(self.ctx, C.byref(a), st), self.ctx, "madm_extract(head)")
        self._keep = keep
        return logits

    # ------------------------------------------------------------------ the hot path
    def extract(self, img: Optional[torch.Tensor], cond_inputs: torch.Tensor, cond_emb: torch.Tensor, timesteps: torch.Tensor,
                shared_noise: torch.Tensor, *, ema: bool = False, stages: int = STAGE_ALL, want_taps: bool = False,
                want_latents: bool = False, noisy_latents_in: Optional[torch.Tensor] = None, B: Optional[int] = None,
                out: Optional[Sequence[torch.Tensor]] = None, want_final: bool = False, img_normalised: bool = False,
                out_dtype=torch.float32) -> Dict[str, object]:
        if self._packed is None:
            raise _lib.MadmError("Engine.extract called before ensure_packed()")
        B = B if B is not None else (img.shape[0] if img is not None else noisy_latents_in.shape[0])
        dev = self.device
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._range_poll()

        def f32c(t, shape, name):
            if t is None:
                return None
            if t.device != dev:
                raise _lib.MadmError(f"{name} must be on {dev}")
            t = t.to(torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise _lib.MadmError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
            return t

        img = f32c(img, (B, 3, 512, 512), "img")
        cond_inputs = f32c(cond_inputs, (B, 77, 768), "cond_inputs")
        cond_emb = f32c(cond_emb, (B, 1280), "cond_emb")
        shared_noise = f32c(shared_noise, (1, 4, 64, 64), "shared_noise")
        noisy_latents_in = f32c(noisy_latents_in, (B, 4, 64, 64), "noisy_latents_in")
        timesteps = timesteps.to(device=dev, dtype=torch.int64).contiguous()
        ws = self.workspace(B)
        a = MadmExtractArgs()
        a.B, a.stages, a.ema = B, stages, 1 if ema else 0
        if out_dtype not in (torch.float32, torch.float16) or (out_dtype == torch.float16 and self.variant != "base"):
            raise _lib.MadmError("out_dtype must be torch.float32, or torch.float16 with the base variant")
        a.flags = (_lib.FLAG_IMG_NORMALISED if img_normalised else 0) | (_lib.FLAG_OUT_FP16 if out_dtype == torch.float16 else 0)
        a.img = img.data_ptr() if img is not None else None
        a.cond_inputs, a.cond_emb, a.timesteps = cond_inputs.data_ptr(), cond_emb.data_ptr(), timesteps.data_ptr()
        a.shared_noise = shared_noise.data_ptr() if shared_noise is not None else None
        a.noisy_latents_in = noisy_latents_in.data_ptr() if noisy_latents_in is not None else None
        res: Dict[str, object] = {}
        outs = []
        if stages & STAGE_PROJ:
            for i, (ch, side) in enumerate(OUT_SHAPES[self.variant]):
                t = out[i] if out is not None else torch.empty(B, ch, side, side, dtype=out_dtype, device=dev)
                if tuple(t.shape) != (B, ch, side, side) or t.dtype != out_dtype or not t.is_contiguous():
                    raise _lib.MadmError(f"out[{i}] must be a contiguous {out_dtype} [{B},{ch},{side},{side}] tensor")
                outs.append(t)
                a.out[i] = t.data_ptr()
            res["features"] = outs
        if want_taps:
            shapes = TAP_SHAPES if self.variant == "base" else ((3, 512),) + TAP_SHAPES[1:]
            taps = [torch.empty(B, c, s, s, dtype=torch.float32, device=dev) for c, s in shapes]
            for i, t in enumerate(taps):
                a.taps[i] = t.data_ptr()
            if self.variant == "s0":  # first feature = decoder_output (ldm_diffusers.py:199), written by the decoder stage
                a.taps[0] = None
                if stages & STAGE_DEC:
                    a.decoded_raw = taps[0].data_ptr()
            res["taps"] = taps
        if want_latents:
            res["latents"] = torch.empty(B, 4, 64, 64, dtype=torch.float32, device=dev)
            res["noisy_latents"] = torch.empty(B, 4, 64, 64, dtype=torch.float32, device=dev)
            a.latents, a.noisy_latents = res["latents"].data_ptr(), res["noisy_latents"].data_ptr()
        if want_final:  # return_unet_final_output (ldm_diffusers.py:211-215)
            if self.variant != "s0" or not (stages & STAGE_DEC):
                raise _lib.MadmError("want_final needs the s0 variant and MADM_STAGE_DEC")
            res["unet_sample"] = torch.empty(B, 4, 64, 64, dtype=torch.float32, device=dev)
            res["decoded"] = torch.empty(B, 3, 512, 512, dtype=torch.float32, device=dev)
            a.unet_sample, a.decoded = res["unet_sample"].data_ptr(), res["decoded"].data_ptr()
        a.packed = self._packed.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.range_flag = self.range_flag.data_ptr()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.madm_extract(self.ctx, C.byref(a), st), self.ctx, "madm_extract")
        if (stages & _lib.STAGE_VAE) and not capturing:
            self._range_publish()
        # inputs must outlive the asynchronous launches: park references until the next call
        self._keep = [img, cond_inputs, cond_emb, timesteps, shared_noise, noisy_latents_in]
        return res

    def set_profiling(self, on: bool):
        _lib.check(self.lib.madm_set_profiling(self.ctx, 1 if on else 0), self.ctx, "madm_set_profiling")

    def profile(self, stages: int = -1) -> Dict[str, Dict[str, float]]:
        """Per kernel family (optionally restricted to a MADM_STAGE_* mask): launches, device ms (CUDA events on the launch stream),
        algorithmic FLOPs / bytes and executed FLOPs."""
        p = _lib.MadmProfile()
        _lib.check(self.lib.madm_get_profile_stages(self.ctx, stages, C.byref(p)), self.ctx, "madm_get_profile_stages")
        return {k.name.decode(): dict(launches=k.launches, ms=k.ms, flops=k.flops, bytes=k.bytes, exec_flops=k.exec_flops) for k in p.kind}

    def launch_count(self, B: int, stages: int = STAGE_ALL) -> int:
        return int(self.lib.madm_launch_count(self.ctx, B, stages))
