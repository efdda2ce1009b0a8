"""TrainStep -- the reference's training step without Lightning, on the B200 engine.

Semantics restated from the reference:
    zero_grad -> forward (src/vit.py:83-92 -> src/models/specvit.py:68-94) -> backward
    -> [DDP mean all-reduce, src/hardware_utils.py:95] -> clip_grad_norm_(train.grad_clip = 0.5,
    src/basemodule.py:244) -> AdamW.step (src/opt/optimizer.py:108)
All kernels of a step are launched through the C ABI on one stream and, by default, captured into a single
CUDA graph (the configured shape is launch-bound).  `step()` takes device tensors; `step_host()` is the
end-to-end call: pinned-host inputs -> H2D -> step -> D2H of the loss.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import dp as _dp
from .engine import ROWS_SLOT
from .model import MyViT


class _PreTrainer:
    """A TRAINABLE LinearPreprocessor (ZCA / PCA matrix as Parameters, src/models/layers.py:51-60, builder.py:168: the
    default `freeze_epochs: 0`) inside the captured step.  Its matrix lives outside the flat arena, so the step gets four
    more launches around the encoder's three: forward GEMM in front; pixel gradient (vitb200_patch_embed_dgrad) -> weight /
    bias gradient (vitb200_linear_wgrad) -> their squared norm into state[5] (vitb200_sumsq_accum) behind the backward
    kernel, so that the optimizer tail clips by the norm over ALL parameters like torch's clip_grad_norm_; and
    vitb200_adamw on the matrix (+ its bf16 operand copy) and bias with the tail's clip coefficient and bias corrections."""

    def __init__(self, model: MyViT, eng) -> None:
        from . import _lib

        self.lib = _lib.load()
        self.eng, self.lin = eng, model.preprocessor.linear
        w, b = self.lin.weight, self.lin.bias
        if w.dtype != torch.float32 or not w.is_contiguous() or w.data_ptr() % 16 or (b is not None and b.data_ptr() % 16):
            raise ValueError("vit_b200: the trainable preprocessor matrix must be a contiguous, 16-byte aligned fp32 tensor")
        N, K = w.shape
        if N % 4 or K % 4 or N != eng.cfg.image_size:
            raise ValueError("vit_b200: preprocessor dimensions must be multiples of 4 and match the model's image_size")
        B, dev = eng.B, eng.device
        self.N, self.K, self.bf = N, K, eng.dt == _lib.BF16
        f32 = dict(dtype=torch.float32, device=dev)
        self.raw = model._raw_buffer(eng)
        self.dx = torch.zeros(B, N, **f32)
        self.gw, self.m_w, self.v_w = (torch.zeros(N, K, **f32) for _ in range(3))
        self.gb = self.m_b = self.v_b = None
        if b is not None:
            self.gb, self.m_b, self.v_b = (torch.zeros(N, **f32) for _ in range(3))
        self.x_op, self.dy_op, self.w16, self.y16 = self.raw, self.dx, None, None
        if self.bf:
            self.x_op = torch.zeros(B, K, dtype=torch.bfloat16, device=dev)
            self.dy_op = torch.zeros(B, N, dtype=torch.bfloat16, device=dev)
            self.w16 = torch.zeros(N, K, dtype=torch.bfloat16, device=dev)
            self.y16 = torch.zeros(B, N, dtype=torch.bfloat16, device=dev)
            self.refresh_operand()
        need = max(int(self.lib.vitb200_linear_wgrad_ws_bytes(B, N, K)) + 4096, int(self.lib.vitb200_grad_norm_ws_bytes(N * K)))
        self.ws = torch.zeros(need, dtype=torch.uint8, device=dev)
        self.ws_sq = torch.zeros(int(self.lib.vitb200_grad_norm_ws_bytes(N * K)), dtype=torch.uint8, device=dev)
        self.ws_fwd = None
        if self.bf and N % 8 == 0 and K % 8 == 0:
            self.ws_fwd = torch.zeros(int(self.lib.vitb200_tc_prelinear_ws_bytes(B, N, K)), dtype=torch.uint8, device=dev)

    def _st(self) -> int:
        return torch.cuda.current_stream(self.eng.device).cuda_stream

    def refresh_operand(self) -> None:
        """bf16 GEMM operand <- the fp32 matrix (after load_state_dict / a restore)."""
        if self.bf:
            from . import _lib
            _lib.check(self.lib.vitb200_cast_bf16(self.lin.weight.data_ptr(), self.w16.data_ptr(), self.N * self.K, self._st()), "cast_bf16")

    def forward(self) -> None:
        from . import _lib
        from ._lib import ACT_NONE, BF16, F32

        lib, e, st = self.lib, self.eng, self._st()
        B, N, K = e.B, self.N, self.K
        bptr = None if self.lin.bias is None else self.lin.bias.data_ptr()
        if not self.bf:
            _lib.check(lib.vitb200_linear_fwd(self.raw.data_ptr(), self.lin.weight.data_ptr(), bptr, e.x.data_ptr(), None, B, N, K,
                                              ACT_NONE, F32, st), "preprocessor linear_fwd")
            return
        _lib.check(lib.vitb200_cast_bf16(self.raw.data_ptr(), self.x_op.data_ptr(), B * K, st), "cast_bf16")
        if self.ws_fwd is not None:
            _lib.check(lib.vitb200_tc_prelinear_fwd(self.x_op.data_ptr(), self.w16.data_ptr(), bptr, e.x.data_ptr(), B, N, K,
                                                    self.ws_fwd.data_ptr(), st), "preprocessor prelinear_fwd")
        else:
            _lib.check(lib.vitb200_linear_fwd(self.x_op.data_ptr(), self.w16.data_ptr(), bptr, self.y16.data_ptr(), None, B, N, K,
                                              ACT_NONE, BF16, st), "preprocessor linear_fwd")
            _lib.check(lib.vitb200_cast_f32(self.y16.data_ptr(), e.x.data_ptr(), B * N, st), "cast_f32")

    def backward(self, train: bool) -> None:
        """Behind the encoder's backward: pixel gradient -> dW, db -> state[5] += |dW|^2 + |db|^2."""
        from . import _lib
        from ._lib import BF16, F32

        lib, e, st = self.lib, self.eng, self._st()
        B, N, K = e.B, self.N, self.K
        e.pixel_grad(train, out=self.dx)
        if self.bf:
            _lib.check(lib.vitb200_cast_bf16(self.dx.data_ptr(), self.dy_op.data_ptr(), B * N, st), "cast_bf16")
        _lib.check(lib.vitb200_linear_wgrad(self.dy_op.data_ptr(), self.x_op.data_ptr(), self.gw.data_ptr(),
                                            None if self.gb is None else self.gb.data_ptr(), B, N, K, 0, BF16 if self.bf else F32,
                                            self.ws.data_ptr(), st), "preprocessor linear_wgrad")
        acc = e.state.data_ptr() + 5 * 4
        _lib.check(lib.vitb200_sumsq_accum(self.gw.data_ptr(), N * K, acc, self.ws_sq.data_ptr(), st), "sumsq")
        if self.gb is not None:
            _lib.check(lib.vitb200_sumsq_accum(self.gb.data_ptr(), N, acc, self.ws_sq.data_ptr(), st), "sumsq")

    def update(self) -> None:
        """Behind the optimizer tail (which left clip coefficient and bias corrections in `state`)."""
        from . import _lib

        lib, e, st = self.lib, self.eng, self._st()
        _lib.check(lib.vitb200_adamw(self.lin.weight.data_ptr(), self.gw.data_ptr(), self.m_w.data_ptr(), self.v_w.data_ptr(),
                                     None if self.w16 is None else self.w16.data_ptr(), self.N * self.K, e.hyper.data_ptr(),
                                     e.state.data_ptr(), None, st), "preprocessor adamw")
        if self.gb is not None:
            _lib.check(lib.vitb200_adamw(self.lin.bias.data_ptr(), self.gb.data_ptr(), self.m_b.data_ptr(), self.v_b.data_ptr(),
                                         None, self.N, e.hyper.data_ptr(), e.state.data_ptr(), None, st), "preprocessor adamw")

    def tensors(self):
        return [t for t in (self.lin.weight.data, None if self.lin.bias is None else self.lin.bias.data, self.m_w, self.v_w,
                            self.m_b, self.v_b, self.w16) if t is not None]

    def optimizer_state(self, step: int, prefix: str = "preprocessor.linear.") -> dict:
        """torch.optim.AdamW per-parameter state of the matrix / bias, keyed by parameter name (checkpoint.py `extra`)."""
        out = {}
        if step > 0:
            for name, m, v in (("weight", self.m_w, self.v_w), ("bias", self.m_b, self.v_b)):
                if m is not None:
                    out[prefix + name] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().cpu().clone(),
                                          "exp_avg_sq": v.detach().cpu().clone()}
        return out

    def load_optimizer_state(self, extra: dict, prefix: str = "preprocessor.linear.") -> None:
        for name, m, v in (("weight", self.m_w, self.v_w), ("bias", self.m_b, self.v_b)):
            st = extra.get(prefix + name)
            if m is not None and st:
                m.copy_(st["exp_avg"].reshape(m.shape))
                v.copy_(st["exp_avg_sq"].reshape(v.shape))
        self.refresh_operand()


class TrainStep:
    def __init__(self, model: MyViT, batch_size: int, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, grad_clip: float = 0.5, use_graph: bool = True, process_group=None,
                 world_size: int = 1, noise_level: float = 0.0, train: bool = True, peer_allreduce: bool = True):
        self.model = model
        model._opt_hyper = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=grad_clip)
        model._engines.pop(batch_size, None)
        self.eng = model._engine(batch_size)
        self.B = batch_size
        self.train = train  # dropout on (model.train()) or off
        self.world = int(world_size)
        self.group = process_group
        if self.world > 1:
            self.eng.set_grad_scale(1.0 / self.world)
            # fused programs: the gradient all-reduce runs inside the optimizer kernel over NVLink peer memory
            if self.eng.fused_bwd and peer_allreduce:
                try:
                    self.eng.peer = _dp.PeerExchange(self.eng.arena.layout.n_opt, self.eng.device, group=process_group)
                except _dp.PeerUnavailable:
                    self.eng.peer = None    # collective decision (all ranks): NCCL all-reduce of the gradient arena
        self.noise_level = float(noise_level)
        self.use_graph = use_graph
        # a trainable (unfrozen) ZCA / PCA matrix in front of the encoder: trained inside the captured step (_PreTrainer)
        self._pre_tr: Optional[_PreTrainer] = None
        pre = model.preprocessor
        if pre is not None and not getattr(pre, "is_frozen", True):
            from .preprocessor import LinearPreprocessor
            if not isinstance(pre, LinearPreprocessor):
                raise NotImplementedError(
                    f"vit_b200: TrainStep trains a LinearPreprocessor (ZCA / PCA) in its graph; train a {type(pre).__name__} "
                    "through MyViT.forward / ViTLModule with a torch optimizer, or freeze it")
            if self.world > 1:
                raise NotImplementedError("vit_b200: a trainable preprocessor inside TrainStep is single-GPU (its gradient is "
                                          "not part of the in-kernel exchange); freeze it or use ViTLModule under DDP")
            self._pre_tr = _PreTrainer(model, self.eng)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_slot: dict = {}    # slot -> the same step on input slot 1 .. 3 (fit_host pipeline)
        self._graph_group: dict = {}   # (slots) -> one graph of consecutive steps on those slots (fit_host pipeline)
        self._graph_rows: dict = {}   # unroll -> graph of `unroll` steps in device-resident dataset mode (fit_device)
        self._rows_key = None
        self._rows_buf = None
        c = model.config
        self.h_x = torch.empty(batch_size, model.input_dim, dtype=torch.float32, pin_memory=True)
        self.h_y = torch.empty(self.eng.labels.shape, dtype=self.eng.labels.dtype, pin_memory=True)
        self.h_loss = torch.empty(1, dtype=torch.float32, pin_memory=True)
        self._segments = None

    def _check_pre(self) -> None:
        """The captured step either trains the preprocessor matrix or treats it as constant -- decided at construction."""
        pre = self.model.preprocessor
        if pre is None:
            return
        frozen = bool(getattr(pre, "is_frozen", True))
        if self._pre_tr is None and not frozen:
            raise NotImplementedError(
                "vit_b200: this TrainStep was built around a FROZEN preprocessor and its captured step does not train the "
                "matrix; build a new TrainStep after model.set_preprocessor_trainable(True)")
        if self._pre_tr is not None and frozen:
            raise NotImplementedError(
                "vit_b200: this TrainStep was built to train the preprocessor matrix, which has been frozen since; build a "
                "new TrainStep")

    # ---- one step's kernel sequence (also what gets captured) ---------------------------------
    def _launch(self, slot: int = 0) -> None:
        eng = self.eng
        eng.cls_only = True   # a training step reads the loss only: the last layer runs for the CLS row alone
        if self._pre_tr is not None:   # trainable preprocessor: 4 more launches around the encoder's (see _PreTrainer)
            fh = eng.can_fuse_head
            self._pre_tr.forward()
            eng.forward(train=self.train, with_labels=True, head_bwd=fh)
            eng.backward(train=self.train, skip_reduce=True, skip_head=fh)
            self._pre_tr.backward(self.train)
            eng.optimizer_step(fused_reduce=True)
            self._pre_tr.update()
            return
        # VITB200_STREAM=1: the optimizer kernel consumes each layer's gradient partials as soon as the backward kernel
        # signals them (bucketed overlap).  Off by default: measured neutral at 1 and 2 GPUs (the blocks that own the last
        # groups still pay the full reduce -> exchange -> barrier chain after the backward kernel, and every signal is a
        # MEMBAR.ALL.GPU inside the backward kernel), see DESIGN.md section 5.
        sg = bool(eng.mega_bwd and os.environ.get("VITB200_STREAM", "0") == "1")
        if slot != 0:
            eng.forward(train=self.train, with_labels=True, head_bwd=True, slot=slot)
            eng.backward(train=self.train, skip_reduce=True, skip_head=True, slot=slot, streamed=sg)
            eng.optimizer_step(fused_reduce=True, streamed=sg)
        elif self.world > 1 and getattr(eng, "peer", None) is None:
            eng.forward(train=self.train, with_labels=True)
            self._backward_overlapped()   # NCCL: gradients are summed across ranks before the optimizer kernel reads them
            eng.optimizer_step()
        else:
            # single GPU: head backward rides on forward's last launch, the gradient-partial reduction on the optimizer's
            fh = eng.can_fuse_head
            eng.forward(train=self.train, with_labels=True, head_bwd=fh)
            eng.backward(train=self.train, skip_reduce=True, skip_head=fh, streamed=sg)
            eng.optimizer_step(fused_reduce=True, streamed=sg)

    def _backward_overlapped(self) -> None:
        """Backward + gradient all-reduce (SUM; the 1/world mean is folded into the optimizer's grad_scale).

        Unfused programs: backward runs in bucket-sized segments (head -> layer L-1 -> ... -> embeddings) and each
        bucket's all-reduce is issued as soon as its kernels are enqueued, so NCCL overlaps the remaining kernels.
        Fused programs: layer gradients only exist after the final partial-sum reduction kernel, and the last bucket is
        on the critical path either way, so the whole optimised range goes out as ONE all-reduce (161 KB at the
        configured shape: latency-bound, ~one NCCL launch)."""
        eng = self.eng
        key = ("bwd", self.train, None, bool(eng.cls_only and eng.mega))
        if key not in eng._progs:
            eng._progs[key] = eng._build_backward(self.train, None)
            eng.launches[key] = len(eng._progs[key])
        prog = eng._progs[key]
        st = torch.cuda.current_stream(eng.device).cuda_stream
        from . import _lib

        def run(seg):
            for fn, args in seg:
                rc = fn(*args, st)
                if rc != 0:
                    _lib.check(rc, fn.__name__)

        if eng.fused_bwd:
            run(prog)
            w = _dp.allreduce_bucket(eng.arena.grad, 0, eng.arena.layout.n_opt, group=self.group, async_op=True)
            if w is not None:
                w.wait()
            return
        L = eng.cfg.num_hidden_layers
        # program layout: [head_bwd, final_ln_bwd] + 11 calls per layer (L-1 .. 0) + [embed_bwd]
        cuts = [2] + [2 + 11 * (i + 1) for i in range(L)] + [len(prog)]
        buckets = _dp.backward_bucket_order(eng.arena.layout.buckets)
        works, lo = [], 0
        for (name, s, e), hi in zip(buckets, cuts):
            run(prog[lo:hi])
            lo = hi
            works.append(_dp.allreduce_bucket(eng.arena.grad, s, e, group=self.group, async_op=True))
        for w in works:
            if w is not None:
                w.wait()

    def _snapshot(self):
        e = self.eng
        e._ensure_opt_state()
        return [t.clone() for t in (e.arena.data, e.exp_avg, e.exp_avg_sq, e.state, e.rng)] + \
               ([e.arena.shadow.clone()] if e.arena.shadow is not None else []) + \
               ([t.clone() for t in self._pre_tr.tensors()] if self._pre_tr is not None else [])

    def _restore(self, snap) -> None:
        e = self.eng
        dst = [e.arena.data, e.exp_avg, e.exp_avg_sq, e.state, e.rng] + \
              ([e.arena.shadow] if e.arena.shadow is not None else []) + \
              (self._pre_tr.tensors() if self._pre_tr is not None else [])
        for d, s in zip(dst, snap):
            d.copy_(s)
        e.arena.mark_shadow_fresh()

    def _capture(self, slot: int = 0, unroll: int = 1, slots=None) -> None:
        """Capture one step on `slot` (x `unroll` in device-resident dataset mode), or -- `slots` given -- one step per
        listed input slot, in order, as ONE graph (fit_host: consecutive steps chained by programmatic dependent launch)."""
        eng = self.eng
        eng.refresh_shadow()
        snap = self._snapshot()
        seq = list(slots) if slots is not None else [slot] * unroll
        side = torch.cuda.Stream(device=eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side):
            for sl in dict.fromkeys(seq):
                self._launch(sl)  # loads every kernel before capture; state is restored below
        torch.cuda.current_stream(eng.device).wait_stream(side)
        self._restore(snap)
        torch.cuda.synchronize(eng.device)
        g = torch.cuda.CUDAGraph()
        # thread_local: other threads (NCCL watchdog, data loaders) may touch CUDA while this thread captures
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for sl in seq:
                self._launch(sl)
        if slots is not None:
            self._graph_group[tuple(seq)] = g
        elif slot == 0:
            self.graph = g
        elif slot == ROWS_SLOT:
            self._graph_rows[unroll] = g
        else:
            self._graph_slot[slot] = g

    @property
    def two_slots(self) -> bool:
        """fit_host can upload straight into the engine's two input slots (whole-network programs, one graph per slot):
        no device-to-device staging copy between the upload and the step."""
        e = self.eng
        return bool(self.use_graph and e.mega and e.mega_bwd and self.model.preprocessor is None
                    and (self.world == 1 or getattr(e, "peer", None) is not None))

    def _run_slot(self, slot: int) -> None:
        """One step on the inputs sitting in the engine's input slot `slot`."""
        if slot == 0:
            return self._run_staged()
        if slot not in self._graph_slot:
            self._capture(slot)
        self._graph_slot[slot].replay()

    # ---- public -----------------------------------------------------------------------------
    def step(self, flux: torch.Tensor, labels: torch.Tensor, error: Optional[torch.Tensor] = None) -> torch.Tensor:
        """flux [B, L] and labels on the device (or pinned host): runs one full training step and returns the
        loss as a 0-dim device tensor (no host sync)."""
        eng = self.eng
        self._check_pre()
        if self.noise_level > 0 and error is not None:  # src/vit.py:86-88
            flux = flux.to(eng.device, non_blocking=True)
            flux = flux + torch.randn_like(flux) * error.to(eng.device, non_blocking=True) * self.noise_level
        self.model._stage_raw(eng, flux, labels, run_pre=self._pre_tr is None)
        self._run_staged()
        if self._pre_tr is not None:   # the matrix was rewritten through its raw pointer: drop cached operand copies
            self._pre_tr.lin.__dict__.pop("_w16", None)
        return eng.loss[0]

    def _run_staged(self) -> None:
        """One step on the inputs already sitting in the engine's buffers."""
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self.eng.refresh_shadow()
            self._launch()

    def _fit_device_rows(self, dataset, epochs, shuffle, seed, tail, start_epoch, max_steps) -> list:
        """fit_device for the whole-network kernels: they read the dataset rows of the epoch's permutation themselves
        (ViTEngine.bind_rows), so a step is the bare 3-launch graph -- no gather launch, no staging copy, no per-step loss
        copy (the forward kernel logs the loss of step i at loss_log[i]).  Steps are replayed `unroll` at a time (one graph
        of `unroll` consecutive steps, chained by programmatic dependent launch) plus single steps for the remainder."""
        from .data import epoch_indices

        eng = self.eng
        B = self.B
        rank = 0
        if self.world > 1:
            import torch.distributed as dist

            rank = dist.get_rank(self.group)
        unroll = max(1, int(os.environ.get("VITB200_UNROLL", "8")))
        lab = dataset.labels if dataset.labels.dtype == eng.labels.dtype else dataset.labels.to(eng.labels.dtype)
        out, left = [], max_steps
        for ep in range(start_epoch, start_epoch + epochs):
            order = epoch_indices(len(dataset), ep, seed=seed, shuffle=shuffle, rank=rank, world=self.world, batch=B,
                                  tail=tail)
            nb = order.numel() // B
            key = (dataset.flux.data_ptr(), lab.data_ptr(), len(dataset))
            if self._rows_key != key or self._rows_buf.numel() < order.numel():
                self._rows_buf = torch.zeros(order.numel(), dtype=torch.int64, device=eng.device)
                self._rows_log = torch.zeros(max(nb, 1), dtype=torch.float32, device=eng.device)
                self._rows_lab = lab
                eng.bind_rows(dataset.flux, lab, self._rows_buf, self._rows_log)
                self._rows_key, self._graph_rows = key, {}
            # the permutation goes through one of two page-locked staging buffers, so that the upload is a stream-ordered
            # DMA behind the previous epoch's graphs instead of a pageable copy that blocks the host until they finish
            pin = self.__dict__.setdefault("_rows_pin", [None, None, 0, None, None])
            i = pin[2]
            pin[2] ^= 1
            if pin[i] is None or pin[i].numel() < order.numel():
                pin[i] = torch.empty(order.numel(), dtype=torch.int64, pin_memory=True)
            if pin[3 + i] is not None:
                pin[3 + i].synchronize()        # the upload that last used this staging buffer is done
            pin[i][:order.numel()].copy_(order)
            self._rows_buf[:order.numel()].copy_(pin[i][:order.numel()], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(eng.device))
            pin[3 + i] = ev
            eng.start_rows()
            if left is not None:
                nb = min(nb, left)
                left -= nb
            for u in ((unroll, 1) if unroll > 1 and nb >= unroll else (1,)):
                if u not in self._graph_rows:
                    self._capture(ROWS_SLOT, u)
            q, r = (nb // unroll, nb % unroll) if unroll > 1 and nb >= unroll else (0, nb)
            gq, g1 = self._graph_rows.get(unroll), self._graph_rows.get(1)
            for _ in range(q):
                gq.replay()
            for _ in range(r):
                g1.replay()
            out.append(self._rows_log[:nb].clone())
            if left is not None and left <= 0:
                break
        return out

    def fit_device(self, dataset, epochs: int = 1, shuffle: bool = True, seed: int = 0, tail: str = "wrap",
                   start_epoch: int = 0, max_steps: Optional[int] = None) -> list:
        """The training loop over a DEVICE-resident dataset (`vit_b200.data.DeviceDataset`): per step
        one gather kernel assembles the batch (row gather by the epoch's permutation, noise injection when
        noise_level > 0, labels) directly in the engine's input buffers, then the step graph runs; the loss of every
        step is kept on the device.  No host<->device traffic and no host synchronisation inside an epoch.  Data-parallel
        runs take rank r's share of the permutation (DistributedSampler semantics, vit_b200.data.epoch_indices).
        Returns one device tensor of per-step losses per epoch."""
        from .data import epoch_indices

        eng, model = self.eng, self.model
        self._check_pre()
        if dataset.length != model.input_dim:
            raise ValueError(f"dataset spectra have {dataset.length} pixels, the model expects {model.input_dim}")
        if dataset.labels is None:
            raise ValueError("fit_device needs a dataset with labels")
        pre = model.preprocessor
        noisy = self.noise_level > 0 and getattr(dataset, "error", None) is not None
        if self.two_slots and not noisy and os.environ.get("VITB200_ROWS", "1") != "0":
            return self._fit_device_rows(dataset, epochs, shuffle, seed, tail, start_epoch, max_steps)
        x_dst = model._raw_buffer(eng)    # eng.x, or the staging buffer in front of a frozen preprocessor
        rank = 0
        if self.world > 1:
            import torch.distributed as dist

            rank = dist.get_rank(self.group)
        out = []
        B = self.B
        for ep in range(start_epoch, start_epoch + epochs):
            order = epoch_indices(len(dataset), ep, seed=seed, shuffle=shuffle, rank=rank, world=self.world, batch=B,
                                  tail=tail).to(eng.device, non_blocking=True)
            nb = order.numel() // B
            if max_steps is not None:
                nb = min(nb, max_steps)
                max_steps -= nb
            losses = torch.empty(nb, dtype=torch.float32, device=eng.device)
            for i in range(nb):
                dataset.gather(order[i * B:(i + 1) * B], x_dst, eng.labels, noise_level=self.noise_level, rng=eng.rng)
                if pre is not None and self._pre_tr is None:
                    pre.forward_into(x_dst, eng.x)
                self._run_staged()
                losses[i:i + 1].copy_(eng.loss, non_blocking=True)
            out.append(losses)
            if max_steps is not None and max_steps <= 0:
                break
        return out

    def _pinned(self, flux_host: torch.Tensor, labels_host: torch.Tensor, slot_x: torch.Tensor, slot_y: torch.Tensor):
        """Host batch -> page-locked memory (batches that already are pinned, e.g. from a DataLoader with
        pin_memory=True -- src/basemodule.py:76-85 -- are used in place)."""
        if flux_host.is_pinned() and flux_host.dtype == torch.float32 and flux_host.is_contiguous():
            hx = flux_host
        else:
            slot_x.copy_(flux_host)
            hx = slot_x
        ly = labels_host.reshape(slot_y.shape)
        if ly.is_pinned() and ly.dtype == slot_y.dtype and ly.is_contiguous():
            hy = ly
        else:
            slot_y.copy_(ly)
            hy = slot_y
        return hx, hy

    def step_host(self, flux_host: torch.Tensor, labels_host: torch.Tensor) -> float:
        """End-to-end step: host inputs -> pinned staging -> H2D -> step -> D2H loss (blocking)."""
        hx, hy = self._pinned(flux_host, labels_host, self.h_x, self.h_y)
        loss = self.step(hx, hy)
        self.h_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream(self.eng.device).synchronize()
        return float(self.h_loss[0])

    def fit_host(self, batches, on_loss=None) -> list:
        """The training loop over an iterable of HOST batches `(flux [B, L], labels)` -- what Lightning's fit loop does
        with the reference's in-RAM dataset (src/dataloader/base.py:219-245): every step copies its inputs host ->
        device and reads its loss back, but the copies are pipelined around the step instead of serialised with it:

          * batch i+1 goes pinned host -> device staging slot on a COPY stream while step i runs;
          * the loss of step i is copied to pinned memory behind the step and read by the host after step i+1 has
            been enqueued (one step late, like a logger), so the GPU never waits for the host round trip.

        Returns the list of per-step losses (floats); `on_loss(i, loss)` is called as each one becomes available."""
        eng = self.eng
        dev = eng.device
        main = torch.cuda.current_stream(dev)
        if self.two_slots and os.environ.get("VITB200_HOST_GROUP", "2") not in ("0", "1"):
            return self._fit_host_grouped(batches, on_loss)
        if getattr(self, "_pipe", None) is None:
            c = self.model.config
            self._pipe = dict(
                copy=torch.cuda.Stream(device=dev),
                d_x=[torch.empty(self.B, self.model.input_dim, dtype=torch.float32, device=dev) for _ in range(2)],
                d_y=[torch.empty_like(eng.labels) for _ in range(2)],
                h_x=[torch.empty(self.B, self.model.input_dim, dtype=torch.float32, pin_memory=True) for _ in range(2)],
                h_y=[torch.empty(eng.labels.shape, dtype=eng.labels.dtype, pin_memory=True) for _ in range(2)],
                h_loss=[torch.empty(1, dtype=torch.float32, pin_memory=True) for _ in range(2)],
                ev_in=[torch.cuda.Event() for _ in range(2)],     # staging slot filled (copy stream)
                ev_free=[torch.cuda.Event() for _ in range(2)],   # staging slot consumed (main stream)
                ev_loss=[torch.cuda.Event() for _ in range(2)],   # loss of the step landed in pinned memory
            )
        P = self._pipe
        copy = P["copy"]
        losses = []
        direct = self.two_slots   # uploads land in the engine's own input slots: the step reads them in place
        if direct:
            P["d_x"] = [eng.input_slot(0)[0], eng.input_slot(1)[0]]
            P["d_y"] = [eng.input_slot(0)[1], eng.input_slot(1)[1]]
            if self.graph is None:
                self._capture(0)
            if 1 not in self._graph_slot:
                self._capture(1)

        def upload(i, batch):
            s = i & 1
            if i >= 2:
                P["ev_free"][s].synchronize()   # the host slot / device slot of step i-2 are reusable
            hx, hy = self._pinned(batch[0], batch[1], P["h_x"][s], P["h_y"][s])
            with torch.cuda.stream(copy):
                if direct and i < 2:
                    copy.wait_stream(main)      # (the slots may still be read by steps enqueued before this loop)
                P["d_x"][s].copy_(hx, non_blocking=True)
                P["d_y"][s].copy_(hy.reshape(P["d_y"][s].shape), non_blocking=True)
                P["ev_in"][s].record(copy)

        def collect(i):
            s = i & 1
            P["ev_loss"][s].synchronize()
            v = float(P["h_loss"][s][0])
            losses.append(v)
            if on_loss is not None:
                on_loss(i, v)

        it = iter(batches)
        nxt = next(it, None)
        i = 0
        if nxt is not None:
            upload(0, nxt)
        while nxt is not None:
            s = i & 1
            main.wait_event(P["ev_in"][s])
            if direct:
                self._run_slot(s)
                loss = eng.loss[0]
            else:
                loss = self.step(P["d_x"][s], P["d_y"][s])
            P["ev_free"][s].record(main)
            if direct:
                P["h_loss"][s] = eng.loss_pinned[s:s + 1]   # written by the forward kernel itself (4 bytes over PCIe)
            else:
                P["h_loss"][s].copy_(loss.reshape(1), non_blocking=True)
            P["ev_loss"][s].record(main)
            nxt = next(it, None)
            if nxt is not None:
                upload(i + 1, nxt)       # overlaps step i
            if i >= 1:
                collect(i - 1)           # one step late: never stalls the GPU
            i += 1
        if i >= 1:
            collect(i - 1)
        return losses

    def _fit_host_grouped(self, batches, on_loss=None) -> list:
        """fit_host for the whole-network programs: the engine has eight input slots = two groups of G = 4.  The G batches of
        group k + 1 are uploaded (copy stream, straight into the engine's slots) while the G steps of group k run as ONE
        graph (3 G kernels chained by programmatic dependent launch -- no graph boundary, no event wait between the
        steps of a group); every step's loss is stored by the forward kernel itself into pinned host memory and read by the host
        after the next group has been enqueued.  A tail of fewer than G batches runs per-slot graphs."""
        from .engine import HOST_SLOTS

        eng = self.eng
        dev = eng.device
        main = torch.cuda.current_stream(dev)
        G = HOST_SLOTS // 2
        if getattr(self, "_gpipe", None) is None:
            self._gpipe = dict(
                copy=torch.cuda.Stream(device=dev),
                h_x=[torch.empty(self.B, self.model.input_dim, dtype=torch.float32, pin_memory=True) for _ in range(HOST_SLOTS)],
                h_y=[torch.empty(eng.labels.shape, dtype=eng.labels.dtype, pin_memory=True) for _ in range(HOST_SLOTS)],
                ev_in=[torch.cuda.Event() for _ in range(2)],      # the group's slots are filled (copy stream)
                ev_done=[torch.cuda.Event() for _ in range(2)],    # the group's steps are finished (main stream)
            )
        Q = self._gpipe
        copy = Q["copy"]
        d = [eng.input_slot(sl) for sl in range(HOST_SLOTS)]
        for p in range(2):
            key = tuple(range(p * G, p * G + G))
            if key not in self._graph_group:
                self._capture(slots=key)
        losses = []
        it = iter(batches)

        def take():
            out = []
            for _ in range(G):
                b = next(it, None)
                if b is None:
                    break
                out.append(b)
            return out

        def upload(k, group):
            p = k & 1
            if k >= 2:
                Q["ev_done"][p].synchronize()   # host staging + device slots of group k - 2 are reusable
            with torch.cuda.stream(copy):
                if k < 2:
                    copy.wait_stream(main)      # (the slots may still be read by steps enqueued before this loop)
                for j, batch in enumerate(group):
                    sl = p * G + j
                    hx, hy = self._pinned(batch[0], batch[1], Q["h_x"][sl], Q["h_y"][sl])
                    d[sl][0].copy_(hx, non_blocking=True)
                    d[sl][1].copy_(hy.reshape(d[sl][1].shape), non_blocking=True)
                Q["ev_in"][p].record(copy)

        def launch(k, n):
            p = k & 1
            main.wait_event(Q["ev_in"][p])
            if n == G:
                self._graph_group[tuple(range(p * G, p * G + G))].replay()
            else:
                for j in range(n):
                    self._run_slot(p * G + j)
            Q["ev_done"][p].record(main)

        def collect(k, n, first):
            p = k & 1
            Q["ev_done"][p].synchronize()
            for j in range(n):
                v = float(eng.loss_pinned[p * G + j])   # written by the forward kernel itself (4 bytes over PCIe)
                losses.append(v)
                if on_loss is not None:
                    on_loss(first + j, v)

        group = take()
        k, done, prev = 0, 0, None
        if group:
            upload(0, group)
        while group:
            n = len(group)
            launch(k, n)
            nxt = take()
            if nxt:
                upload(k + 1, nxt)          # overlaps the steps of group k
            if prev is not None:
                collect(*prev)              # one group late: never stalls the GPU
            prev = (k, n, done)
            done += n
            group = nxt
            k += 1
        if prev is not None:
            collect(*prev)
        return losses

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.h_x.numel() * 4 + self.h_y.numel() * self.h_y.element_size()

    @property
    def d2h_bytes_per_step(self) -> int:
        return 4

    def set_lr(self, lr: float) -> None:
        self.model._opt_hyper["lr"] = float(lr)
        self.eng.set_lr(lr)

    def close(self) -> None:
        """Release the captured CUDA graph (and the host pipeline).  Data-parallel runs MUST call this before
        torch.distributed.destroy_process_group(): NCCL (2.28) does not finish destroying a communicator while a live
        CUDA graph still holds collectives captured on it -- the call blocks forever on every rank."""
        import gc

        self.graph = None
        self._graph_slot, self._graph_group, self._gpipe = {}, {}, None
        self._graph_rows, self._rows_key = {}, None
        self._pipe = None
        gc.collect()
        torch.cuda.synchronize(self.eng.device)
        if getattr(self.eng, "peer", None) is not None:
            self.eng.peer.close()
            self.eng.peer = None

    def kernel_launches(self) -> int:
        """Our kernels per step (+2 device-to-device staging copies of the inputs done by torch in step())."""
        return self.eng.kernel_launches(self.train, fused_tail=self.world == 1 or getattr(self.eng, "peer", None) is not None)


class EvalStep:
    """scripts/test.py semantics: model.eval(), no_grad, forward only -- ONE forward per batch (the reference's
    validation/test steps run the model twice per batch, src/vit.py:127-150,194-215) with the metrics accumulated on the
    device (`evaluate`)."""

    def __init__(self, model: MyViT, batch_size: int, use_graph: bool = True):
        self.model = model
        self.eng = model._engine(batch_size)
        self.use_graph = use_graph
        self.graphs = {}
        c = model.config
        self.h_x = torch.empty(batch_size, model.input_dim, dtype=torch.float32, pin_memory=True)
        self.h_logits = torch.empty(batch_size, c.num_labels, dtype=torch.float32, pin_memory=True)

    @property
    def graph(self):
        return self.graphs.get(False)

    def _run(self, with_labels: bool) -> None:
        eng = self.eng
        eng.cls_only = True   # logits / loss only
        if not self.use_graph:
            eng.forward(train=False, with_labels=with_labels)
            return
        g = self.graphs.get(with_labels)
        if g is None:
            eng.refresh_shadow()
            side = torch.cuda.Stream(device=eng.device)
            side.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(side):
                eng.forward(train=False, with_labels=with_labels)
            torch.cuda.current_stream(eng.device).wait_stream(side)
            torch.cuda.synchronize(eng.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                eng.forward(train=False, with_labels=with_labels)
            self.graphs[with_labels] = g
        else:
            eng.refresh_shadow()
        g.replay()

    def forward(self, flux: torch.Tensor, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Logits [B, num_labels] (device, no host sync); with labels the loss is left in `eng.loss`."""
        self.model._stage_raw(self.eng, flux, labels)
        self._run(labels is not None)
        return self.eng.logits

    def forward_host(self, flux_host: torch.Tensor) -> torch.Tensor:
        self.h_x.copy_(flux_host)
        logits = self.forward(self.h_x)
        self.h_logits.copy_(logits, non_blocking=True)
        torch.cuda.current_stream(self.eng.device).synchronize()
        return self.h_logits

    def evaluate(self, dataset, return_preds: bool = True) -> dict:
        """One pass over a device-resident dataset in order: loss, MAE / MSE / R2 (regression) or accuracy
        (classification) accumulated on the device, predictions kept on the device; ONE host read at the end.  The last
        partial batch runs on a second engine of that size (no padding, so the metrics are exact)."""
        from .data import EvalMetrics

        model, eng = self.model, self.eng
        if dataset.labels is None:
            raise ValueError("evaluate needs a dataset with labels")
        if dataset.length != model.input_dim:
            raise ValueError(f"dataset spectra have {dataset.length} pixels, the model expects {model.input_dim}")
        N, B, C = len(dataset), eng.B, model.config.num_labels
        dev = eng.device
        is_cls = model.task_type == "cls"
        metrics = EvalMetrics(C, is_cls, dev)
        preds = torch.empty(N, C, dtype=torch.float32, device=dev) if return_preds else None
        idx = torch.arange(N, dtype=torch.int64, device=dev)
        pre = model.preprocessor

        def run(step: "EvalStep", lo: int, hi: int) -> None:
            e = step.eng
            dst = model._raw_buffer(e)
            dataset.gather(idx[lo:hi], dst, e.labels)
            if pre is not None:
                pre.forward_into(dst, e.x)
            step._run(True)
            metrics.update(e.logits, e.labels, e.loss)
            if preds is not None:
                preds[lo:hi].copy_(e.logits, non_blocking=True)

        nb = N // B
        for i in range(nb):
            run(self, i * B, (i + 1) * B)
        if N - nb * B:
            run(EvalStep(model, N - nb * B, use_graph=False), nb * B, N)
        out = metrics.compute()
        out["preds"] = preds
        out["labels"] = dataset.labels
        return out
