"""Flat-buffer Adam for the drop-in models.

The reference trains with ``torch.optim.Adam(self.parameters(), lr, weight_decay)`` under ``StepLR``
(models/magnet_gnn.py:378-386, models/mpnn_2d.py:205-213): ~150 parameter tensors, each with its own chain of
element-wise launches per step.  ``FlatAdam`` keeps parameters, gradients and both moments in four flat fp32 buffers
(the parameters / ``.grad`` of the model become views into them) and updates everything with ONE CUDA launch
(``mgb_adam_step``); the flat gradient buffer is also what the training all-reduce sends (``allreduce_flat_gradient``), so
no per-step concatenation or copy-back is needed.  It is a ``torch.optim.Optimizer``: ``torch.optim.lr_scheduler.StepLR``
drives its ``lr`` exactly as it drives torch's Adam.
"""
from typing import Iterable

import torch

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FlatAdam got no trainable parameters")
        dev = params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in params):
            raise RuntimeError("FlatAdam needs fp32 CUDA parameters on one device (magnet_b200 has no CPU fallback)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        # every tensor starts on a 16-byte boundary of the flat buffers (float4 accesses in the kernel and in the layers)
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self._params, self._offs = params, offs
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                                   # the module's Parameter objects stay, their storage moves
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.steps = 0

    @torch.no_grad()
    def zero_grad(self, set_to_none: bool = False):           # the views must survive: always zero in place
        self.flat_grad.zero_()
        for p, o in zip(self._params, self._offs):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        for p, o in zip(self._params, self._offs):             # a gradient that autograd re-created (p.grad was None) is copied in
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                self.flat_grad[o:o + p.numel()].view_as(p).copy_(p.grad)
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.steps += 1
        with torch.cuda.device(self.flat_param.device):
            _lib.check(_lib.lib().mgb_adam_step(_lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                                _lib.ptr(self.exp_avg_sq), self.flat_param.numel(), float(g["lr"]),
                                                float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                                float(g["weight_decay"]), self.steps, float(grad_scale), _lib.stream()),
                       "adam_step")
        for p in self._params:                                  # the kernel wrote through raw pointers: tell the caches
            torch._C._increment_version(p)                      # (packed weight copies are keyed on the version counter)
        return loss

    # ---- checkpoints: the layout of torch.optim.Adam (per-parameter "step" / "exp_avg" / "exp_avg_sq"), so that a
    # Lightning checkpoint written with FlatAdam resumes under torch's Adam and vice versa.
    # Semantics note: parameters that receive no gradient are updated as if their gradient were zero (the flat gradient
    # buffer is zeroed, never None); torch's Adam skips such parameters entirely.  The two agree unless weight_decay != 0
    # or the parameter still carries momentum from earlier steps.  Every parameter of the drop-in models receives a
    # gradient in every training step.
    def state_dict(self):
        state = {}
        for i, (p, o) in enumerate(zip(self._params, self._offs)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.steps)),
                        "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        groups = []
        for g in self.param_groups:
            g2 = {k: v for k, v in g.items() if k != "params"}
            g2["params"] = list(range(len(self._params)))
            groups.append(g2)
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError("FlatAdam.load_state_dict: the checkpoint holds a different parameter list")
        for k, v in groups[0].items():
            if k != "params" and k in self.param_groups[0]:
                self.param_groups[0][k] = v
        state = state_dict["state"]
        steps = 0
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, (p, o) in enumerate(zip(self._params, self._offs)):
            st = state.get(i, state.get(str(i)))
            if st is None:                      # torch's Adam has no state for a parameter that never received a gradient
                continue
            n = p.numel()
            self.exp_avg[o:o + n].view_as(p).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + n].view_as(p).copy_(st["exp_avg_sq"])
            steps = max(steps, int(float(st["step"])))
        self.steps = steps
        # the flat views of parameters / gradients stay as they are (load_state_dict of the MODULE copies into them)


def allreduce_flat_gradient(opt: FlatAdam, world: int = None) -> float:
    """Sum the flat gradient buffer across ranks (one collective, no copies); returns the ``grad_scale`` to pass to
    ``opt.step`` so that the update uses the mean over ranks."""
    import torch.distributed as dist
    world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
    if world > 1:
        dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


class OverlappedFlatAllReduce:
    """The training all-reduce of SURVEY §8e / K12, overlapped with the tail of backward: FlatAdam's flat gradient buffer is cut
    into ``n_buckets`` contiguous parameter ranges; a post-accumulate hook per parameter counts a bucket down and, when the
    last gradient of the bucket has been written, launches its sum all-reduce on a side stream behind an event — so every
    bucket but the one autograd finishes last travels over NVLink while the remaining backward kernels run (what Lightning's
    DDP does with 25 MB buckets, models/magnet_gnn.py:378-386 + scripts/magnet_gnn/*.sh `--gpus`).  ``finish()`` launches what
    is left (parameters that received no gradient), makes the compute stream wait for the side stream and returns the
    ``grad_scale`` for ``FlatAdam.step``.  One collective per bucket on views of the buffer itself: no copies."""

    def __init__(self, opt: FlatAdam, world: int = None, n_buckets: int = 4, group=None):
        import torch.distributed as dist
        self.opt, self.group = opt, group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        params, offs = opt._params, opt._offs
        total = opt.flat_grad.numel()
        n_buckets = max(1, min(n_buckets, len(params)))
        # contiguous parameter ranges of about equal size, in buffer order
        bounds, target, acc, start = [], total / n_buckets, 0, 0
        for i, p in enumerate(params):
            acc += (p.numel() + 3) // 4 * 4
            if acc >= target * (len(bounds) + 1) and len(bounds) < n_buckets - 1:
                bounds.append((start, i + 1))
                start = i + 1
        bounds.append((start, len(params)))
        self.buckets = []            # (first param, last param + 1, flat lo, flat hi)
        for a, b in bounds:
            if a >= b:
                continue
            hi = offs[b] if b < len(params) else total
            self.buckets.append((a, b, offs[a], hi))
        self._bucket_of = {}
        for bi, (a, b, _, _) in enumerate(self.buckets):
            for i in range(a, b):
                self._bucket_of[id(params[i])] = bi
        self._cuda = opt.flat_grad.is_cuda
        self._side = torch.cuda.Stream(device=opt.flat_grad.device) if self._cuda else None
        self._left, self._launched, self._works = [], [], []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in params]
        self.reset()

    def reset(self):
        self._left = [b - a for a, b, _, _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []

    def _launch(self, bi):
        import torch.distributed as dist
        self._launched[bi] = True
        if self.world <= 1:
            return
        _, _, lo, hi = self.buckets[bi]
        view = self.opt.flat_grad[lo:hi]
        if self._cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self._side.wait_event(ev)
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _hook(self, p):
        bi = self._bucket_of.get(id(p))
        if bi is None or self._launched[bi]:
            return
        self._left[bi] -= 1
        if self._left[bi] == 0:
            # autograd may have re-created p.grad outside the flat buffer (first step after a set_to_none): FlatAdam.step
            # copies such gradients in, and a bucket holding one must be reduced after that copy -> leave it to finish()
            a, b, _, _ = self.buckets[bi]
            fg = self.opt.flat_grad
            ok = all(q.grad is not None and q.grad.data_ptr() == fg.data_ptr() + 4 * o
                     for q, o in zip(self.opt._params[a:b], self.opt._offs[a:b]))
            if ok:
                self._launch(bi)

    def finish(self) -> float:
        """Call after ``loss.backward()``: returns the grad_scale (1 / world) for ``FlatAdam.step``."""
        opt = self.opt
        with torch.no_grad():
            for p, o in zip(opt._params, opt._offs):          # stray gradients into the buffer before the late buckets go out
                if p.grad is not None and p.grad.data_ptr() != opt.flat_grad.data_ptr() + 4 * o:
                    opt.flat_grad[o:o + p.numel()].view_as(p).copy_(p.grad)
                    p.grad = opt.flat_grad[o:o + p.numel()].view_as(p)
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._launch(bi)
        if self._cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self._side)
        for w in self._works:
            w.wait()
        self.reset()
        return 1.0 / self.world

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
