"""Training engine: the hot loop of examples/intermediate_downscaling.py:695-742 re-hosted for one process per GPU.

    training_step (:281-306)  forward -> clip_replace_constant -> crop target -> loss         (+ backward, :723-742)
    optimizer (:642-644)      AdamW(lr, betas=(beta_1, beta_2), weight_decay)
    data parallel (:618-621)  FSDP NO_SHARD == gradient all-reduce over the data-parallel group

Here the whole step is a fixed kernel schedule (``reslim_forward`` / ``o2_loss_fwd_bwd`` / ``reslim_backward``) with no
autograd graph around the large tensors.  Parameters live in ONE flat fp32 master buffer (the module's parameters are
views into it), gradients in one flat fp32 buffer, and the tcgen05 operands in one flat bf16 buffer that the fused AdamW
kernel refreshes while it updates the master copy -- so a step contains no cast kernels.  With world_size > 1 each
parameter group's gradient slice is all-reduced (NCCL, average) as soon as the backward schedule has finished it, on
NCCL's own stream, overlapping the rest of the backward.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops
from .dp import BucketReducer, ChainedMap, FlatLayout, FullShard, ShardedReducer, shard_size
from .losses import Metric, clip_spec
from .reslim import Res_Slim_ViT, reslim_backward, reslim_forward


class GradScaler:
    """Dynamic loss scaling with the semantics of the reference's bf16 branch (examples/intermediate_downscaling.py:493-495,
    733-742: ``ShardedGradScaler(init_scale=8192, growth_interval=100)`` + the ``min_scale = 128`` floor; torch defaults
    growth_factor 2, backoff_factor 0.5): gradients are produced pre-multiplied by ``scale``; a step whose gradients contain
    inf / NaN on ANY rank is skipped and halves the scale, ``growth_interval`` clean steps in a row double it."""

    def __init__(self, init_scale: float = 8192.0, growth_interval: int = 100, min_scale: float = 128.0,
                 growth_factor: float = 2.0, backoff_factor: float = 0.5):
        self.scale, self.growth_interval, self.min_scale = float(init_scale), int(growth_interval), float(min_scale)
        self.growth_factor, self.backoff_factor = float(growth_factor), float(backoff_factor)
        self.growth_tracker = 0
        self.skipped = 0

    def update(self, found_inf: bool):
        if found_inf:
            self.scale *= self.backoff_factor
            self.growth_tracker = 0
            self.skipped += 1
        else:
            self.growth_tracker += 1
            if self.growth_tracker == self.growth_interval:
                self.scale *= self.growth_factor
                self.growth_tracker = 0
        if self.scale < self.min_scale:                      # intermediate_downscaling.py:741-742
            self.scale = self.min_scale

    def state_dict(self):
        return {"scale": self.scale, "growth_tracker": self.growth_tracker}

    def load_state_dict(self, sd):
        self.scale, self.growth_tracker = float(sd["scale"]), int(sd["growth_tracker"])


class TrainEngine:
    def __init__(self, model: Res_Slim_ViT, loss: Metric, in_variables: Sequence[str], out_variables: Sequence[str],
                 var_weights: Optional[Dict[str, float]] = None, lr: float = 2e-4, betas=(0.9, 0.99),
                 weight_decay: float = 1e-5, eps: float = 1e-8, process_group=None, clip_constants: bool = True,
                 shard_optimizer: bool = False, shard_params: bool = False, grad_scaler: Optional[GradScaler] = None):
        self.model = model
        self.scaler = grad_scaler
        self.loss = loss
        self.in_variables, self.out_variables = list(in_variables), list(out_variables)
        self.var_weights = dict(var_weights or {})
        self.lr, self.betas, self.weight_decay, self.eps = lr, betas, weight_decay, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.clip = clip_spec(self.out_variables) if clip_constants else (-1, 0)
        self.step_count = 0
        self.act = model._act_dtype()
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("TrainEngine needs the model on a CUDA device (no CPU fallback)")
        self.device = dev

        # ---- FULL_SHARD mode (interm_1b / interm_10b; reference: FSDP FULL_SHARD with one unit per Block,
        # intermediate_downscaling.py:583-617): the GEMM weights of the kernel schedule -- 99.9 % of the parameters --
        # leave the flat buffers and live as per-unit 1/world shards (dp.FullShard: all-gather per Block in forward and
        # backward, reduce-scatter of the Block's gradient, sharded master / Adam state / update); everything else
        # (biases, LayerNorm, convolutions, embeddings, the var_agg q / kv weights the host-side tables are built from)
        # stays replicated below with an all-reduced gradient.  Works with world size 1 (gathers become copies).
        named = [(n, p) for n, p in model.named_parameters()]
        self.fs: Optional[FullShard] = None
        if shard_params:
            if shard_optimizer:
                raise ValueError("shard_params already shards the optimizer; pass only one of the two modes")
            kernel = set(model._names)
            big = {n: p for n, p in named if n in kernel and n.endswith(".weight") and p.dim() == 2 and p.requires_grad}
            root = [n for n in big if not n.startswith("blocks.")]
            units = [root] + [[n for n in big if n.startswith(f"blocks.{i}.")] for i in range(len(model.blocks))]
            self.fs = FullShard(units, big, self.act, process_group)
            for p in big.values():                      # the module keeps no full copy (engine.full_state_dict() gathers)
                p.data = torch.empty(0, device=dev, dtype=p.dtype)
                p.grad = None
            named = [(n, p) for n, p in named if n not in big]

        # ---- flat buffers; parameters become views of the fp32 master
        self.names = [n for n, _ in named]
        sizes = [p.numel() for _, p in named]
        self.layout = FlatLayout(self.names, sizes, align=8)
        offs = [self.layout.range[n][0] for n in self.names]
        total = self.total = self.layout.total
        # FSDP-style mode: Adam state and the update are sharded (rank r owns [r*S, (r+1)*S) of the flat buffers), gradients
        # are reduce-scattered bucket by bucket, updated fp32 / bf16 parameters are all-gathered (in place) after the step
        self.sharded = bool(shard_optimizer) and self.world > 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.shard = shard_size(total, self.world) if self.sharded else total
        if self.sharded:
            total = self.shard * self.world             # flat buffers padded so that every rank's shard has equal size
        self.own = (self.rank * self.shard, (self.rank + 1) * self.shard) if self.sharded else (0, total)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        n_state = self.shard if self.sharded else total
        self.flat_m = torch.zeros(n_state, device=dev, dtype=torch.float32)      # sharded mode: this rank's shard only
        self.flat_v = torch.zeros(n_state, device=dev, dtype=torch.float32)
        self.flat_b = torch.zeros(total, device=dev, dtype=torch.bfloat16) if self.act == torch.bfloat16 else None
        self.P: Dict[str, torch.Tensor] = {}
        self.G: Dict[str, torch.Tensor] = {}
        self.Wc: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for (n, p), o, s in zip(named, offs, sizes):
                view = self.flat_p[o:o + s].view(p.shape)
                view.copy_(p.detach().float())
                p.data = view
                self.P[n] = view
                self.G[n] = self.flat_g[o:o + s].view(p.shape)
                p.grad = self.G[n] if p.requires_grad else None
                if self.flat_b is not None and n.endswith(".weight") and p.dim() == 2:
                    self.Wc[n] = self.flat_b[o:o + s].view(p.shape)
        self.frozen = [n for (n, p) in named if not p.requires_grad]
        if self.flat_b is not None:
            ops.cast_bf16(self.flat_p, self.flat_b)
        self.reducer = (ShardedReducer if self.sharded else BucketReducer)(self.flat_g, self.layout, process_group)
        self._train_runs = None
        self._lat = None
        self._chw = None
        # The module's own forward (validation, tiled inference, the public-API training path) must see the operands
        # this engine maintains: the fused AdamW updates the flat fp32 master through raw pointers, so neither
        # ``_version`` nor ``data_ptr()`` of a parameter changes and a bf16 copy cached by the module would go stale
        # after the first optimizer step.  Replicated bf16: the flat bf16 buffer AdamW refreshes.  FULL_SHARD: lookups
        # gather the unit (``external_begin`` restarts the forward prefetch order before every module forward).
        if self.fs is not None:
            model.external_wc = ChainedMap(self.Wc if self.flat_b is not None else self.P, self.fs.params)
            model.external_begin = lambda: self.fs.begin(+1)
        elif self.flat_b is not None:
            model.external_wc = self.Wc
        # CUDA-graph mode (enable_graph): the whole step is captured once and replayed; see graph_step
        self._graph = None
        self._drop_word = None
        self._graph_warm = 0
        self._sc_dev = None

    # ------------------------------------------------------------------ pieces
    def _on_ready(self, names: List[str]):
        if self.fs is not None:
            self.fs.ready(names)
            names = [n for n in names if n not in self.fs.where]
        self.reducer.ready(names)

    def full_state_dict(self) -> Dict[str, torch.Tensor]:
        """The module's state dict with the sharded weights gathered (collective in FULL_SHARD mode)."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        if self.fs is not None:
            for n in self.fs.where:
                sd[n] = self.fs.full_tensor(n)
        return sd

    def forward_backward(self, x: torch.Tensor, y: torch.Tensor):
        """x [B,V,H,W] fp32, y [B,C,H',W'] fp32 on the device.  Returns the [C+1] loss vector (device, fp32);
        leaves the (all-reduced) gradients in the flat gradient buffer."""
        m = self.model
        self.flat_g.zero_()
        if x.dim() == 5:
            x = x.flatten(1, 2)
        g = m.geometry(x, self.in_variables, self.out_variables, self.act)
        # host prep on the small parameters (autograd-visible)
        tab_s, tab_v = m.frontend_tables(m.get_var_ids(self.in_variables))
        posres = m.pos_res_embed(g.gh, g.gw, self.act)
        with torch.no_grad():
            posres_act = posres if self.act == torch.float32 else ops.cast_bf16(posres)
            Wc = self.P if self.act == torch.float32 else self.Wc
            G = self.G
            if self.fs is not None:                 # lookups in these mappings drive the gathers / gradient slots
                Wc, G = self.fs.params, ChainedMap(self.G, self.fs.grads)
                self.fs.begin(+1)
            preds, S = reslim_forward(g, self.P, Wc, x, tab_s.detach(), tab_v.detach(), posres_act)
            if self._lat is None:
                self._lat = self.loss._lat(preds)
                self._chw = self.loss._ch_w(preds, self.out_variables, self.var_weights)
            vec, dpred = ops.loss_fwd_bwd(preds, y, self.loss.kind, lat_w=self._lat, ch_w=self._chw, clamp_ch=self.clip[0],
                                          const_mask=self.clip[1],
                                          grad_scale=self.scaler.scale if self.scaler is not None else 1.0)
            if self.fs is not None:
                self.fs.begin(-1)
            dts, dtv, dpos = reslim_backward(g, self.P, Wc, x, tab_s.detach(), tab_v.detach(), S, dpred, G,
                                             on_ready=self._on_ready)
        torch.autograd.backward([tab_s, tab_v, posres], [dts, dtv, dpos])
        for n in self.frozen:                       # e.g. pos_embed when learn_pos_emb=False
            self.G[n].zero_()
        if self.world > 1:
            self._on_ready([n for n in self.names if not self._is_kernel_param(n)])
            self.reducer.finish()
        if self.fs is not None:
            self.fs.finish()
        return vec

    def _is_kernel_param(self, n):
        return n in self._kernel_set

    @property
    def _kernel_set(self):
        s = getattr(self, "_ks", None)
        if s is None:
            s = self._ks = set(self.model._names)
        return s

    def _adam_scalars(self, grad_scale: float = 1.0):
        t = self.step_count
        return [self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, 1.0 - self.betas[0] ** t,
                (1.0 - self.betas[1] ** t) ** 0.5, grad_scale]

    def optimizer_step(self, grad_scale: float = 1.0):
        self.step_count += 1
        if self._train_runs is None:                # frozen parameters get neither an update nor weight decay
            self._train_runs = self.layout.runs([n for n in self.names if n not in self.frozen])
        o0, o1 = self.own
        for lo, hi in self._train_runs:
            lo, hi = max(lo, o0), min(hi, o1)           # sharded mode: only this rank's part of every run
            if lo >= hi:
                continue
            if self._sc_dev is not None:                # graph mode: step-dependent scalars come from device memory
                ops.adamw_dev(self.flat_p[lo:hi], self.flat_g[lo:hi], self.flat_m[lo - o0:hi - o0],
                              self.flat_v[lo - o0:hi - o0], self.flat_b[lo:hi] if self.flat_b is not None else None,
                              self._sc_dev)
                continue
            ops.adamw(self.flat_p[lo:hi], self.flat_g[lo:hi], self.flat_m[lo - o0:hi - o0], self.flat_v[lo - o0:hi - o0],
                      self.flat_b[lo:hi] if self.flat_b is not None else None, self.lr, self.betas[0], self.betas[1],
                      self.eps, self.weight_decay, self.step_count, grad_scale)
        if self.fs is not None:                         # FULL_SHARD units: each rank updates its slice of every unit
            for p32, g32, m, v, low in self.fs.shards():
                ops.adamw(p32, g32, m, v, low, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                          self.step_count, grad_scale)
            self.fs.after_step()
        if self.sharded:                                # in-place all-gather of the updated shards
            dist.all_gather_into_tensor(self.flat_p, self.flat_p[o0:o1], group=self.pg)
            if self.flat_b is not None:
                dist.all_gather_into_tensor(self.flat_b, self.flat_b[o0:o1], group=self.pg)

    def grads_nonfinite(self) -> bool:
        """inf / NaN anywhere in this step's (reduced) gradients, agreed on by all ranks (one streaming pass + one sync,
        like GradScaler.step's found_inf.item())."""
        flag = torch.zeros(1, device=self.device, dtype=torch.int32)
        o0, o1 = self.own
        ops.nonfinite(self.flat_g[o0:o1], flag)
        if self.fs is not None:
            for _, g32, _, _, _ in self.fs.shards():
                ops.nonfinite(g32, flag)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.pg)
        return bool(flag.item())

    def step(self, x, y):
        if self._graph_warm:
            return self.graph_step(x, y)
        vec = self.forward_backward(x, y)
        if self.scaler is None:
            self.optimizer_step()
            return vec
        found_inf = self.grads_nonfinite()                   # gradients carry the factor scaler.scale
        if not found_inf:
            self.optimizer_step(grad_scale=1.0 / self.scaler.scale)
        self.scaler.update(found_inf)
        return vec

    # ------------------------------------------------------------------ CUDA-graph mode
    def enable_graph(self, warm_steps: int = 3):
        """Replay the whole step (forward, loss, backward, AdamW) as ONE captured CUDA graph from step ``warm_steps`` + 1
        on.  For launch-bound configurations (interm_8m: ~180 kernels of a few microseconds each, paced by the host's
        issue rate when launched one by one).  Restrictions: one GPU (world size 1), replicated parameters, fixed batch
        shape, no dynamic grad scaler.  ``lr`` may change between steps: AdamW reads its scalars from device memory
        (o2_adamw_dev).  Dropout / drop-path work under replay: the seeds the kernels were captured with stay fixed, but
        every mask hash also mixes in a 64-bit step word read from device memory at run time (o2_dropout_seed_source) and
        the per-sample drop-path factors live in fixed device buffers; both are rewritten before each replay.  The loss
        vector returned by ``step`` is then ONE static tensor that every replay overwrites (read or clone it before the
        next step)."""
        if self.world > 1 or self.fs is not None or self.sharded:
            raise RuntimeError("enable_graph: single-GPU, replicated-parameter engines only")
        if self.scaler is not None:
            raise RuntimeError("enable_graph: the dynamic grad scaler decides on the host whether to step")
        self._graph_warm = max(1, int(warm_steps))

    def graph_step(self, x, y):
        if self._graph is None and self.step_count < self._graph_warm:       # lazy initialisation happens eagerly
            vec = self.forward_backward(x, y)
            self.optimizer_step()
            return vec
        if self._graph is None:
            self._gx, self._gy = x.clone(), y.clone()
            self._sc_dev = torch.zeros(8, device=self.device, dtype=torch.float32)
            m = self.model
            self._drop_word = None
            if m.training and (m.drop_rate > 0 or m.drop_path > 0):
                from .reslim import DropPlan
                dpr = [float(v) for v in torch.linspace(0, m.drop_path, m.depth)] if m.depth else []
                m.static_drop_plan = DropPlan(m.drop_rate, dpr, x.shape[0], int(torch.randint(0, 2 ** 62, (1,)).item()),
                                              self.device)
                self._drop_word = torch.zeros(1, device=self.device, dtype=torch.int64)
            g = torch.cuda.CUDAGraph()
            count = self.step_count
            ops.dropout_seed_source(self._drop_word)    # baked into the arguments of every captured dropout kernel
            try:
                with torch.cuda.graph(g):
                    self._gvec = self.forward_backward(self._gx, self._gy)
                    self.optimizer_step()
            finally:
                ops.dropout_seed_source(None)
            self.step_count = count                     # capture executes nothing
            self._graph = g
        if x.shape != self._gx.shape or y.shape != self._gy.shape:
            raise RuntimeError("graph_step: the batch shape changed after the graph was captured")
        self._gx.copy_(x, non_blocking=True)
        self._gy.copy_(y, non_blocking=True)
        self.step_count += 1
        # pageable source: the driver stages it before the call returns, so the next step cannot overwrite it in flight
        self._sc_dev.copy_(torch.tensor(self._adam_scalars(), dtype=torch.float32), non_blocking=True)
        if self._drop_word is not None:                 # new masks for this replay (seeds come from torch's CPU generator)
            word = torch.randint(-2 ** 62, 2 ** 62, (1,), dtype=torch.int64)
            self._drop_word.copy_(word, non_blocking=True)
            self.model.static_drop_plan.redraw_paths(int(word.item()) & 0x3FFFFFFFFFFFFFFF)
        self._graph.replay()
        return self._gvec
