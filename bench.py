#!/usr/bin/env python
"""bench.py — benchmarks of the B200-native MAgNet hot path, one JSON line per run.

    python bench.py --gpus N --steps K --warmup W                       # headline: mp_layer (+ the other metrics inside "metrics")
    python bench.py --metric {mp_layer,in_layer,inr_decode,rollout}     # one metric as the line itself
    python bench.py --metric inr_sweep                                  # BASELINE configs[4]: Q x k sweep -> profiles/r02_inr_sweep.json
    python bench.py --impl reference [--metric ...]                     # the reference's own CPU implementation (oracle/_ref)

BASELINE.json metric: "edges/s per MP layer fwd+bwd; query pts/s INR decode; rollout steps/s @1/2/4/8 B200".
  mp_layer    edges/s per message-passing layer, forward + backward — BASELINE configs[1]: the MP-PDE (mpnn_2d) processor, 5 GNN_Layers,
              hidden 128, time_window 10, synthetic 2-D irregular-uniform 64x64-point meshes, 32 samples per GPU, radius such that the
              reference's 32-neighbour index-order truncation is active.  A step = forward + backward of the 5 layers + gradient
              all-reduce (N > 1) + Adam update.
  in_layer    the same metric for MAgNet[GNN]'s InteractionNetwork (models/magnet_gnn.py:44-90) on the stage-3 graph of BASELINE
              configs[2] at test resolution 256 (65,536 nodes per sample, r = 0.08, 32-cap active); forward-only (fused kernel) next to it.
  inr_decode  query points/s of continuous_decoder + projector (kNN search, gather, proj_head, blend, 5-layer MLP), configs[4] shape.
  rollout     MAgNet[GNN] validation rollout steps/s (models/magnet_gnn.py:442-475), configs[2]: B = 32, L = Nq = 256 per GPU.
Every metric line carries `roofline` (algorithmic FLOPs / CUDA-event time / MEASURED_PEAKS.json), `e2e` (host buffers in, result out,
copies inside the timed region) and, at N = 1, `cpu_baseline` = the reference's own model files (byte-compiled into oracle/_ref by
oracle/stage_ref.py; restated third-party surface underneath) timed on the box's host cores on a bounded sample.
All metrics shard by independent samples with no data-path collective (weak scaling); values are summed over ranks, times are the max.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LAYERS = 5
TW = 10
NODES_PER_SAMPLE = 4096
SAMPLES_PER_GPU = 32
RADIUS = 0.09
FLOP_PER_EDGE_FWD = 101_632          # 2*(269*128 + 128*128), GNN_Layer message, reference formulation
FLOP_PER_NODE_FWD = 98_560
EXEC_FLOP_PER_EDGE_FWD = 2 * 128 * 128        # what the fused edge kernel executes per edge (factorised first Linear)
IN_FLOP_PER_EDGE_FWD = 229_376       # 2*(384*128 + 4*128^2), InteractionNetwork edge_fn
IN_FLOP_PER_NODE_FWD = 196_608
ENC_FLOP_PER_ROW = 2 * (13 * 128 + 4 * 128 * 128)      # Encoder MLPs (approx.: 12/13-wide first Linear)
INR_FLOP_PER_QUERY_T = lambda k: k * 2 * 132 * 128 + 131_328     # per (query, time step): k proj_head rows + projector


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained"), "measured"
    return 6650.0, 1590.0, None, "fallback"


def _ncu(name):
    """dram bytes per launch of a kernel from the committed ncu --set full summaries (profiles/*.json): measured once per round
    under ncu, NOT in this run (a number taken under a profiler is never a bench value; the traffic figure is a property of
    the kernel + problem size and is stated with its source)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        return d.get("dram_bytes_per_launch"), d.get("source_report")
    except Exception:  # noqa: BLE001
        return None, None


class HP(dict):
    __getattr__ = dict.__getitem__


def magnet_hparams(**over):
    hp = dict(time_slice=10, latent_dim=128, num_message_passing_steps=5, mlp_layers=4, mlp_hidden=128, radius=0.08, n_chan=128,
              teacher_forcing=True, codec_neighbors=4, noise=0, interpolation="area", factor=0.3, step_size=50, loss="l1",
              lr=1e-3, weight_decay=0)
    hp.update(over)
    return HP(hp)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The nvidia-smi process is started before the
    warm-up (its start-up takes longer than a short timed region, more so on an 8-GPU box); samples are time-stamped on
    arrival and only those between mark_start() and mark_end() are summarised."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0, self.t1 = None, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, timeout: float = 8.0):
        """Block until nvidia-smi has delivered its first sample (its start-up can outlast the warm-up)."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.02)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        rows = []
        if self.t0 is not None and self.t1 is not None and self.rows:
            # a sample describes the 50 ms before it arrived: keep those whose window overlaps the timed region; a region
            # shorter than one period keeps the sample that closes it
            rows = [r for t, r in self.rows if self.t0 <= t <= self.t1 + 0.06]
            if not rows:
                after = [(t, r) for t, r in self.rows if t >= self.t0]
                rows = [min(after, key=lambda tr: tr[0])[1]] if after else [self.rows[-1][1]]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class Env:
    """Rank / device / collective plumbing shared by the metrics."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = None
        self.L = None

    def init_gpu(self):
        import torch.distributed as dist
        from magnet_b200 import _lib
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.L = _lib.lib()
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, value: float, op: str) -> float:
        if self.world == 1:
            return float(value)
        import torch.distributed as dist
        t = torch.tensor([float(value)], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t)

    def prof(self, kid):
        t, c = ctypes.c_double(0), ctypes.c_int64(0)
        self.L.mgb_profile_collect(kid, ctypes.byref(t), ctypes.byref(c))
        return t.value, c.value

    def timed(self, fn, steps, warmup, clocks=None):
        """W untimed + K timed calls of fn between barrier + synchronize, CUDA events on the current stream; returns
        (max-over-ranks ms for the K steps, launches per step)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(max(warmup, 3)):
            fn()
        self.barrier()
        self.L.mgb_profile_enable(1)
        l0 = self.L.mgb_launch_count()
        if clocks is not None:
            clocks.wait_ready()
            self.barrier()
            clocks.mark_start()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        self.barrier()
        if clocks is not None:
            clocks.mark_end()
        self.L.mgb_profile_enable(0)
        launches = (self.L.mgb_launch_count() - l0) // max(steps, 1)
        return self.reduce(ev0.elapsed_time(ev1), "max"), int(launches)


def base_line(env, metric, unit, value, ms_per_step, steps, warmup, dtype, config, **more):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": env.world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": config}
    line.update(more)
    return line


# =============================================================================================================
# mp_layer — BASELINE configs[1]
# =============================================================================================================
def make_workload(samples: int, seed: int, device):
    """Layer inputs of the config-2 processor for `samples` samples of one shared 4096-node mesh."""
    from magnet_b200 import synthetic as S
    g = S._gen(seed)
    mesh = S.mesh("uniform", NODES_PER_SAMPLE, 2, g)
    n = samples * NODES_PER_SAMPLE
    pos_xy = mesh.repeat(samples, 1)
    x = torch.randn(n, 128, generator=g)
    u = torch.randn(n, TW, generator=g)
    pos = pos_xy[:, :1].repeat(1, 2).contiguous()       # quirk F6: both position inputs are the x coordinate
    var = torch.rand(samples, 1, generator=g).repeat_interleave(NODES_PER_SAMPLE, 0)
    gy = torch.randn(n, 128, generator=g)
    return dict(coords=pos_xy, x=x, u=u, pos=pos, var=var, gy=gy, n=n, samples=samples)


def layer_state_dicts():
    from magnet_b200 import synthetic as S
    shapes = {"message_net_1.0.weight": (128, 256 + TW + 2 + 1), "message_net_1.0.bias": (128,),
              "message_net_2.0.weight": (128, 128), "message_net_2.0.bias": (128,),
              "update_net_1.0.weight": (128, 257), "update_net_1.0.bias": (128,),
              "update_net_2.0.weight": (128, 128), "update_net_2.0.bias": (128,)}
    return [S.seeded_state_dict(shapes, 100 + l) for l in range(N_LAYERS)]


def mp_config(n_gpus, precision=None):
    c = {"workload": "mpnn_2d processor (5x GNN_Layer, hidden 128, tw 10) on synthetic 2-D irregular-uniform "
                     "64x64-point meshes, BASELINE configs[1]",
         "samples_per_gpu": SAMPLES_PER_GPU, "nodes_per_sample": NODES_PER_SAMPLE, "radius": RADIUS,
         "max_num_neighbors": 32, "layers": N_LAYERS, "parallelism": f"samples sharded over {n_gpus} GPU(s), no data-path collective"
                        + ("; one NCCL all-reduce of the flat gradient buffer per step" if n_gpus > 1 else ""),
         "l2": "working set per layer (~1 GB) exceeds the 126 MB L2; no explicit flush"}
    if precision:
        c["precision"] = precision
    return c


def ref_modules():
    """The reference's own model files: source tree in the build container, byte-compiled copies (oracle/_ref) on the GPU box."""
    from oracle import reference_loader as rl
    ns = rl.load()
    return ns, ("reference", f"unmodified reference model files ({ns.kind}: {'/root/reference' if ns.kind == 'source' else 'oracle/_ref'}) "
                             "behind the restated torch_geometric/torch_cluster surface (oracle/thirdparty)")


def cpu_mp_layer(samples: int, steps: int, warmup: int, threads: int):
    """Reference CPU path: models/mpnn_2d.py GNN_Layer x5, forward + backward, on a bounded sample."""
    from oracle import graph as OG
    torch.set_num_threads(threads)
    ns, (kind, how) = ref_modules()
    w = make_workload(samples, 0, "cpu")
    batch = torch.arange(samples).repeat_interleave(NODES_PER_SAMPLE)
    ei = OG.radius_graph(w["coords"], RADIUS, batch, loop=False, threads=threads)
    layers = []
    for sd in layer_state_dicts():
        m = ns.mpnn_2d.GNN_Layer(128, 128, 128, TW, 1)
        m.load_state_dict(sd, strict=True)
        layers.append(m)
    E = ei.shape[1]

    def step():
        out = w["x"].clone().requires_grad_()
        for m in layers:
            out = m(out, w["u"], w["pos"], w["var"], ei, batch)
        out.backward(w["gy"])
        for m in layers:
            m.zero_grad(set_to_none=True)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    sample = (f"{samples} of {SAMPLES_PER_GPU} samples per step ({samples * NODES_PER_SAMPLE} nodes, {E} edges), all {N_LAYERS} "
              f"layers fwd+bwd, {steps} timed steps after {warmup} warm-up; {how}")
    return {"value": E * N_LAYERS * steps / dt, "unit": "edges/s", "cores": threads, "kind": kind, "sample": sample,
            "ms_per_step": dt / steps * 1e3}


def bench_mp_layer(env, clocks):
    from magnet_b200 import graph as MG, functional as MF
    from magnet_b200.mpnn import GNN_Layer
    from magnet_b200.optim import FlatAdam, OverlappedFlatAllReduce
    args, dev, world, L = env.args, env.dev, env.world, env.L
    MF.set_precision(args.precision)
    w = make_workload(SAMPLES_PER_GPU, 1 + env.rank, dev)
    host = {k: v.pin_memory() for k, v in w.items() if torch.is_tensor(v) and k in ("x", "u", "pos", "var", "gy")}
    d = {k: v.to(dev) for k, v in w.items() if torch.is_tensor(v)}
    seg = MG.uniform_segments(SAMPLES_PER_GPU, NODES_PER_SAMPLE, dev)
    ei = MG.radius_graph(d["coords"], RADIUS, loop=False, ptr=seg.gptr)
    plan = MG.plan_for(ei, w["n"])
    batch = torch.arange(SAMPLES_PER_GPU, device=dev).repeat_interleave(NODES_PER_SAMPLE)
    E = ei.shape[1]
    layers = []
    for sd in layer_state_dicts():
        m = GNN_Layer(128, 128, 128, TW, 1, pos_dim=2).to(dev)
        m.load_state_dict(sd, strict=True)
        layers.append(m)
    params = [p for m in layers for p in m.parameters()]
    # the reference's optimizer (Adam + weight decay, models/mpnn_2d.py:205-213) as one flat-buffer launch; its flat
    # gradient buffer is what the all-reduce sends.  The step is part of the timed region: so is the re-packing of the
    # kernel-side weight copies that every update triggers.
    opt = FlatAdam(params, lr=1e-5, weight_decay=1e-8)
    # the training all-reduce leaves in one bucket per layer from the gradient hooks, on a side stream, while the backward of
    # the layers below is still running; finish() makes the compute stream wait for what is left
    ar = OverlappedFlatAllReduce(opt, world, n_buckets=N_LAYERS) if world > 1 else None

    def step(x, u, pos, var, gy):
        h = x.detach().requires_grad_()
        out = h
        for m in layers:
            out = m(out, u, pos, var, ei, batch, plan=plan, segments=seg)
        out.backward(gy)
        scale = ar.finish() if ar is not None else 1.0
        opt.step(grad_scale=scale)
        opt.zero_grad()
        return out

    total_ms, launches = env.timed(lambda: step(d["x"], d["u"], d["pos"], d["var"], d["gy"]), args.steps, args.warmup, clocks)
    if args.steps * 0.025 < 0.12:
        time.sleep(0.12)          # a very short timed region: let the sample that covers it arrive before nvidia-smi is stopped
    prof = {"edge_fwd": env.prof(0), "edge_bwd": env.prof(1), "node_gemm": env.prof(2), "wgrad": env.prof(3)}
    # ---- end to end: pinned host inputs -> H2D -> 5 layers fwd+bwd -> D2H of the result checksum ----
    # Every step copies its own inputs from pinned host memory (double-buffered on a side stream: the transfer of step k+1
    # overlaps the kernels of step k) and reads its result back (a training step's result is its loss-like scalar: 4 bytes).
    e2e_steps = max(1, args.steps)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    dbuf = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()} for _ in range(2)]
    results = torch.zeros(e2e_steps, dtype=torch.float32).pin_memory()
    used = [None, None]

    def upload(j):
        with torch.cuda.stream(copy_stream):
            if used[j] is not None:
                copy_stream.wait_event(used[j])
            for k, v in host.items():
                dbuf[j][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ee0.record()
    copy_stream.wait_stream(main_stream)
    nxt = upload(0)
    for i in range(e2e_steps):
        j = i & 1
        main_stream.wait_event(nxt)
        if i + 1 < e2e_steps:
            nxt = upload(j ^ 1)
        dx = dbuf[j]
        out = step(dx["x"], dx["u"], dx["pos"], dx["var"], dx["gy"])
        used[j] = torch.cuda.Event()
        used[j].record(main_stream)
        results[i:i + 1].copy_(out.sum().reshape(1), non_blocking=True)
    ee1.record()
    env.barrier()
    checksum = float(results.sum())
    assert checksum == checksum, "e2e produced NaN"
    e2e_ms = env.reduce(ee0.elapsed_time(ee1), "max")
    edges = env.reduce(float(E), "sum")
    del d, out, dbuf
    torch.cuda.empty_cache()
    if env.rank != 0:
        return None
    value = edges * N_LAYERS * args.steps / (total_ms * 1e-3)
    e2e_value = edges * N_LAYERS * e2e_steps / (e2e_ms * 1e-3)
    hbm, tf, tf_sus, which = _peaks()
    bt, bc = prof["edge_bwd"]
    ft, fc = prof["edge_fwd"]
    bwd_ms, fwd_ms = bt / max(bc, 1), ft / max(fc, 1)
    alg_flops = 2 * FLOP_PER_EDGE_FWD * E            # dgrad + wgrad of both message Linears, reference formulation
    achieved = alg_flops / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0
    traffic, tsrc = _ncu("dominant_kernel_ncu.json")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf,
                "traffic": traffic, "traffic_source": f"profiles/dominant_kernel_ncu.json ({tsrc}): ncu --set full capture of the same kernel and problem size, not measured in this run",
                "kernel": "gnn_edge_bwd_tc_kernel" if args.precision != "fp32" else "gnn_edge_bwd_kernel",
                "peak_source": which, "frac_of_sustained_peak": achieved / tf_sus if tf_sus else None,
                "kernel_ms": bwd_ms, "kernel_share_of_step": bt / total_ms if total_ms > 0 else None,
                "executed_tflops": 3 * EXEC_FLOP_PER_EDGE_FWD * E / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0,
                "edge_fwd_kernel_ms": fwd_ms,
                "edge_fwd_algorithmic_tflops": FLOP_PER_EDGE_FWD * E / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else 0.0,
                # node-level stages (row-wise Linears and weight gradients on [N,128] tensors): HBM-bound kernels
                "node_linear_ms_per_step": prof["node_gemm"][0] / max(args.steps, 1), "node_linear_launches_per_step": prof["node_gemm"][1] // max(args.steps, 1),
                "node_wgrad_ms_per_step": prof["wgrad"][0] / max(args.steps, 1), "node_wgrad_launches_per_step": prof["wgrad"][1] // max(args.steps, 1),
                "note": {"fp32": "fp32 FFMA path (1e-5 contract) measured against the bf16 tensor peak",
                         "fp32_tc": "tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate (1e-5 contract)",
                         "bf16": "tcgen05 bf16 operands, fp32 accumulate (1e-2 contract)"}[args.precision]}
    line = base_line(env, "edges/s per MP layer fwd+bwd", "edges/s", value, total_ms / args.steps, args.steps, args.warmup,
                     "bf16" if args.precision == "bf16" else "f32", mp_config(world, args.precision),
                     clocks=clocks.summary() if clocks else None,
                     e2e={"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                          "result": "a training step returns its scalar: the 4-byte read is the whole result (inference metrics below download fields)"},
                     gpu_launches=launches, roofline=roofline, edges_per_gpu=E)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_mp_layer(1, 2, 1, os.cpu_count() or 1)
    return line


# =============================================================================================================
# in_layer — InteractionNetwork on the stage-3 graph of BASELINE configs[2] at test resolution 256
# =============================================================================================================
IN_NODES_PER_SAMPLE, IN_SAMPLES = 65536, 2


def in_workload(samples, nodes, seed, dev):
    from magnet_b200 import synthetic as S
    g = S._gen(seed)
    pts = []
    for _ in range(samples):
        c = S.mesh("concentrated", nodes, 2, g)
        pts.append(2 * (c - c.min(0).values) / (c.max(0).values - c.min(0).values) - 1)
    pos = torch.cat(pts, 0)
    x = torch.randn(samples * nodes, 128, generator=g)
    return pos, x, g


def in_state_dict():
    from magnet_b200 import synthetic as S
    from magnet_b200.magnet_gnn import InteractionNetwork
    m = InteractionNetwork(128, 128, 128, 128, 4, 128)
    return S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5)


def in_config(env, samples, nodes, E=None):
    return {"workload": "MAgNet[GNN] InteractionNetwork (edge MLP 384->128x4->128 + LayerNorm, mean at the unsorted row, node MLP) on the "
                        "stage-3 radius graph (r = 0.08, loop, 32-cap active) of concentrated 2-D meshes, BASELINE configs[2] at test resolution 256",
            "samples_per_gpu": samples, "nodes_per_sample": nodes, "edges_per_gpu": E, "radius": 0.08,
            "parallelism": f"samples sharded over {env.world} GPU(s), no data-path collective",
            "l2": "e_features alone (E x 512 B) exceed the 126 MB L2; no explicit flush"}


def cpu_in_layer(threads, nodes=8192, steps=2):
    from oracle import graph as OG
    torch.set_num_threads(threads)
    ns, (kind, how) = ref_modules()
    pos, x, g = in_workload(1, nodes, 0, "cpu")
    e = OG.radius_graph(pos, 0.08 * (65536 / nodes) ** 0.5, torch.zeros(nodes, dtype=torch.int64), loop=True, threads=threads)
    ei = torch.stack([e[1], e[0]])
    E = ei.shape[1]
    ef = torch.randn(E, 128, generator=g)
    m = ns.magnet_gnn.InteractionNetwork(128, 128, 128, 128, 4, 128)
    m.load_state_dict(in_state_dict(), strict=True)
    gy = torch.randn(nodes, 128, generator=g)

    def step():
        xi, ei_ = x.clone().requires_grad_(), ef.clone().requires_grad_()
        y, _ = m(xi, ei, ei_)
        y.backward(gy)
        m.zero_grad(set_to_none=True)

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        m(x, ei, ef)
        t1 = time.perf_counter()
        for _ in range(steps):
            m(x, ei, ef)
        dtf = time.perf_counter() - t1
    return {"value": E * steps / dt, "unit": "edges/s", "cores": threads, "kind": kind, "forward_only_value": E * steps / dtf,
            "sample": f"one {nodes}-node sample (radius scaled to the same mean degree; {E} edges), InteractionNetwork fwd+bwd, "
                      f"{steps} timed steps after 1 warm-up; {how}"}


def bench_in_layer(env, clocks=None):
    from magnet_b200 import graph as MG, functional as MF, synthetic as S
    from magnet_b200.magnet_gnn import InteractionNetwork
    args, dev = env.args, env.dev
    MF.set_precision("fp32_tc" if args.precision == "fp32" else args.precision)
    pos, x, g = in_workload(IN_SAMPLES, IN_NODES_PER_SAMPLE, 900 + env.rank, dev)
    pos, x = pos.to(dev), x.to(dev)
    seg = MG.uniform_segments(IN_SAMPLES, IN_NODES_PER_SAMPLE, dev)
    ei = MG.radius_graph(pos, 0.08, loop=True, ptr=seg.gptr, swap_rows=True)
    N, E = x.shape[0], ei.shape[1]
    plan = MG.plan_for(ei, N)
    ef_host = torch.randn(E, 128, generator=g).pin_memory()
    x_host = x.cpu().pin_memory()
    ef = ef_host.to(dev)
    layer = InteractionNetwork(128, 128, 128, 128, 4, 128).to(dev)
    layer.load_state_dict(in_state_dict(), strict=True)
    gy = torch.randn(N, 128, generator=g).to(dev)
    steps = max(3, args.steps // 2)

    def fwd():
        with torch.no_grad():
            return layer(x, ei, ef, plan=plan, return_e=False)[0]

    def train():
        xi, ei_ = x.detach().requires_grad_(), ef.detach().requires_grad_()
        y, _ = layer(xi, ei, ei_, plan=plan, return_e=False)
        y.backward(gy)
        layer.zero_grad(set_to_none=True)

    f_ms, f_launch = env.timed(fwd, steps, args.warmup, clocks)
    k_t, k_c = env.prof(5)
    t_ms, t_launch = env.timed(train, steps, args.warmup)
    kf_t, kf_c = env.prof(5)            # the forward launches of the training steps
    kb_t, kb_c = env.prof(6)            # mgb_in_edge_bwd: both recompute passes, per call
    # e2e (inference): node + edge features from pinned host memory in, aggregated node update out
    out_host = torch.empty(N, 128).pin_memory()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ee0.record()
    for _ in range(steps):
        x.copy_(x_host, non_blocking=True)
        ef.copy_(ef_host, non_blocking=True)
        out_host.copy_(fwd(), non_blocking=True)
    ee1.record()
    env.barrier()
    e2e_ms = env.reduce(ee0.elapsed_time(ee1), "max")
    edges = env.reduce(float(E), "sum")
    del ef, x, gy
    torch.cuda.empty_cache()
    if env.rank != 0:
        return None
    hbm, tf, tf_sus, which = _peaks()
    k_ms = k_t / max(k_c, 1)
    kb_ms = kb_t / max(kb_c, 1)
    # dominant kernels of the training step: the two backward passes.  Algorithmic work = data + weight gradients of the four
    # 128x128 Linears they cover (2 x 4 x 2*128*128 FLOP per edge); the recompute (7 MMA groups of 15) is overhead by definition.
    bwd_flop = 2 * 4 * 2 * 128 * 128
    nterm = 3 if args.precision != "bf16" else 1
    achieved = bwd_flop * E / (kb_ms * 1e-3) / 1e12 if kb_ms > 0 else 0.0
    step_achieved = 3 * IN_FLOP_PER_EDGE_FWD * E / (t_ms / steps * 1e-3) / 1e12
    fwd_achieved = IN_FLOP_PER_EDGE_FWD * E / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    traffic, tsrc = _ncu("r02_in_edge_bwd_v1.json")
    line = base_line(env, "edges/s per MP layer fwd+bwd", "edges/s", edges * steps / (t_ms * 1e-3), t_ms / steps, steps, args.warmup,
                     "bf16" if args.precision == "bf16" else "f32", in_config(env, IN_SAMPLES, IN_NODES_PER_SAMPLE, E),
                     clocks=clocks.summary() if clocks else None,
                     forward_only={"value": edges * steps / (f_ms * 1e-3), "unit": "edges/s", "ms_per_step": f_ms / steps, "gpu_launches": f_launch,
                                   "note": "rollout / decode path: P|Q Linear + ONE fused edge launch (mgb_in_edge_fwd) + node MLP + LayerNorm"},
                     e2e={"value": edges * steps / (e2e_ms * 1e-3), "unit": "edges/s (forward)", "h2d_bytes_per_step": (N + E) * 512,
                          "d2h_bytes_per_step": N * 512, "steps": steps},
                     gpu_launches=t_launch,
                     roofline={"bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf,
                               "frac_of_sustained_peak": achieved / tf_sus if tf_sus else None, "peak_source": which,
                               "kernel": "in_edge_bwd_tc_kernel<0> + <1> (the two recompute passes of mgb_in_edge_bwd, one call)", "kernel_ms": kb_ms,
                               "kernel_share_of_step": kb_ms / (t_ms / steps) if t_ms > 0 else None,
                               "algorithmic_flop_per_edge": bwd_flop, "executed_flop_per_edge": 15 * 2 * 128 * 128 * nterm,
                               "executed_tflops": 15 * 2 * 128 * 128 * nterm * E / (kb_ms * 1e-3) / 1e12 if kb_ms > 0 else 0.0,
                               "compulsory_bytes_per_edge": 2 * 512 + 512 + 12,
                               "traffic": traffic, "traffic_source": f"profiles/r02_in_edge_bwd_v1.json ({tsrc}): sum of both passes' first captured launches is in the file; not measured in this run",
                               "forward_kernel": {"kernel": "in_edge_fwd_tc_kernel", "kernel_ms": k_ms, "achieved": fwd_achieved, "frac": fwd_achieved / tf,
                                                  "algorithmic_flop_per_edge": IN_FLOP_PER_EDGE_FWD, "executed_flop_per_edge": 5 * 2 * 128 * 128 * nterm,
                                                  "training_launch_ms": kf_t / max(kf_c, 1)},
                               "whole_step": {"achieved": step_achieved, "frac": step_achieved / tf,
                                              "algorithmic_flop_per_edge": 3 * IN_FLOP_PER_EDGE_FWD,
                                              "note": "reference formulation (384-wide first Linear per edge), forward + data + weight gradients, over the whole fwd+bwd step incl. the node MLP"},
                               "note": "training = fused forward launch + two fused recompute backward passes (weight gradients resident in tensor memory) "
                                       "+ by-source sum for dQ + one tensor-core Linear backward for d e_features / dWe; nothing of size [E,128] is saved by the forward"})
    if env.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_in_layer(os.cpu_count() or 1)
    return line


# =============================================================================================================
# inr_decode — BASELINE configs[4] shape
# =============================================================================================================
def decode_inputs(Lr, Q, T, seed, dev):
    from magnet_b200 import synthetic as S
    g = S._gen(seed)
    lr_coords = 2 * torch.rand(1, Lr, 2, generator=g) - 1
    hr_coords = 2 * torch.rand(1, Q, 2, generator=g) - 1
    enc = torch.randn(1, Lr, 128, generator=g)
    x_lr = torch.randn(1, T, 1, Lr, generator=g)
    t = torch.linspace(0, 1, 2 * T)[None]
    return [v.to(dev) for v in (lr_coords, hr_coords, enc, x_lr, t)]


def magnet_model(dev, ns=None, **over):
    from magnet_b200 import synthetic as S
    if ns is None:
        from magnet_b200.magnet_gnn import MAgNetGNN
        m = MAgNetGNN(magnet_hparams(**over))
    else:
        m = ns.magnet_gnn.MAgNetGNN(magnet_hparams(**over))
    m.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 7), strict=True)
    return m.to(dev).eval()


def cpu_inr_decode(threads, Lr=1 << 14, Q=1 << 14, k=4, T=10):
    torch.set_num_threads(threads)
    ns, (kind, how) = ref_modules()
    m = magnet_model("cpu", ns, codec_neighbors=k)
    lr_coords, hr_coords, enc, x_lr, t = decode_inputs(Lr, Q, T, 500, "cpu")
    with torch.no_grad():
        t0 = time.perf_counter()
        hr = m.projector(m.continuous_decoder(x_lr, enc, lr_coords, hr_coords, t))
        dt = time.perf_counter() - t0
    assert hr.shape[0] == Q
    return {"value": Q / dt, "unit": "query points/s", "cores": threads, "kind": kind,
            "sample": f"capped: {Q} queries over a {Lr}-node low-res mesh, k = {k}, T = {T}, one call (the reference's brute-force kNN and "
                      f"Python T x k loop make the full size infeasible); {how}"}


def bench_inr_decode(env, clocks=None, Lr=1 << 18, Q=1 << 20, k=4, T=10, cpu=True, flop_k=None):
    from magnet_b200 import functional as MF
    args, dev = env.args, env.dev
    MF.set_precision("fp32_tc" if args.precision == "fp32" else args.precision)
    m = magnet_model(dev, codec_neighbors=k)
    lr_coords, hr_coords, enc, x_lr, t = decode_inputs(Lr, Q, T, 500 + env.rank, dev)
    hr_host = hr_coords.cpu().pin_memory()
    out_host = torch.empty(Q, T).pin_memory()
    steps = max(3, args.steps // 2)

    def decode():
        with torch.no_grad():
            return m.decode_queries(x_lr, enc, lr_coords, hr_coords, t)

    ms, launches = env.timed(decode, steps, args.warmup, clocks)
    k_t, k_c = env.prof(7)
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ee0.record()
    for _ in range(steps):
        hr_coords.copy_(hr_host.reshape(hr_coords.shape), non_blocking=True)
        out_host.copy_(decode().reshape(Q, T), non_blocking=True)
    ee1.record()
    env.barrier()
    e2e_ms = env.reduce(ee0.elapsed_time(ee1), "max")
    queries = env.reduce(float(Q), "sum")
    if env.rank != 0:
        return None
    hbm, tf, tf_sus, which = _peaks()
    flop_k = k if flop_k is None else flop_k
    flops = INR_FLOP_PER_QUERY_T(flop_k) * T * Q
    achieved = flops / (ms / steps * 1e-3) / 1e12
    line = base_line(env, "query points/s, INR decode", "query points/s", queries * steps / (ms * 1e-3), ms / steps, steps, args.warmup,
                     "bf16" if args.precision == "bf16" else "f32",
                     {"workload": "MAgNetGNN.continuous_decoder + projector (kNN search, gather, proj_head, blend, 5-layer MLP), BASELINE configs[4] shape",
                      "queries_per_gpu": Q, "lowres_nodes": Lr, "k": k, "time_steps": T, "interpolation": "area",
                      "parallelism": f"queries sharded over {env.world} GPU(s), no collective",
                      "l2": "the query stream (Q x T rows of 512 B through the projector) exceeds the 126 MB L2; no explicit flush"},
                     clocks=clocks.summary() if clocks else None,
                     e2e={"value": queries * steps / (e2e_ms * 1e-3), "unit": "query points/s", "h2d_bytes_per_step": Q * 8, "d2h_bytes_per_step": Q * T * 4,
                          "steps": steps, "result": "hr_points [Q, T] fp32"},
                     gpu_launches=launches,
                     roofline={"bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf,
                               "frac_of_sustained_peak": achieved / tf_sus if tf_sus else None, "peak_source": which,
                               "kernel": "whole decode call (reference-formulation FLOPs over its CUDA-event time)",
                               "algorithmic_flop_per_query": INR_FLOP_PER_QUERY_T(flop_k) * T,
                               "decode_kernel_ms": k_t / max(k_c, 1), "traffic": None})
    if env.world == 1 and cpu and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_inr_decode(os.cpu_count() or 1, k=k, T=T)
    return line


def run_inr_sweep(env):
    """BASELINE configs[4]: Q in 1e4..1e8 x k in {4,8,16,32}; queries beyond 2^22 are decoded in chunks of 2^22 (the low-res mesh and its
    latents stay resident); the CPU reference beside it on a capped problem."""
    res = {"lowres_nodes": 1 << 18, "time_steps": 10, "rows": [],
           "note": "roofline_frac: FLOPs of the two blended neighbours + projector (k-independent useful work) over the measured bf16 "
                   "tensor peak; the fused decoder's work does not depend on k (it searches the two nearest nodes, which are the "
                   "first two of any k-nearest list), the reference's does"}
    cpu = {}
    for k in (4, 8, 16, 32):
        if env.rank == 0 and env.world == 1:
            cpu[k] = cpu_inr_decode(os.cpu_count() or 1, Lr=1 << 12, Q=1 << 12, k=k)
        for q_exp in (4, 5, 6, 7, 8):
            Q = 10 ** q_exp
            chunk = min(Q, 1 << 22)
            env.args.steps, env.args.warmup = 6, 3
            # FLOPs of the two neighbours the reference blends (F9); it evaluates proj_head for all k and discards k - 2 of them,
            # so the k-neighbour formulation would credit discarded work (fractions above 1 at k = 32)
            line = bench_inr_decode(env, None, Q=chunk, k=k, cpu=False, flop_k=2)
            if line is not None:
                per_gpu = line["value"] / env.world
                res["rows"].append({"queries": Q, "k": k, "chunk": chunk, "chunks": -(-Q // chunk), "n_gpus": env.world,
                                    "query_points_per_s": line["value"], "e2e_query_points_per_s": line["e2e"]["value"],
                                    "seconds_for_Q": Q / per_gpu / env.world, "roofline_frac": line["roofline"]["frac"],
                                    "flop_per_query": line["roofline"]["algorithmic_flop_per_query"],
                                    "cpu_reference_query_points_per_s": cpu.get(k, {}).get("value")})
                print(json.dumps(res["rows"][-1]), flush=True)
    if env.rank == 0:
        res["cpu_reference"] = cpu
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"r02_inr_sweep_n{env.world}.json"), "w"), indent=1)


# =============================================================================================================
# rollout — BASELINE configs[2]
# =============================================================================================================
def rollout_flops(B, L, Nq, E1, E3, T=10, k=4):
    proc = lambda E, N: 5 * (E * IN_FLOP_PER_EDGE_FWD + N * IN_FLOP_PER_NODE_FWD) + (E + N) * ENC_FLOP_PER_ROW
    return proc(E1, B * L) + proc(E3, B * (L + Nq)) + B * Nq * T * INR_FLOP_PER_QUERY_T(k) + B * (L + Nq) * ENC_FLOP_PER_ROW


def cpu_rollout(threads, B=32, L=256, Nq=256):
    from magnet_b200 import synthetic as S
    torch.set_num_threads(threads)
    ns, (kind, how) = ref_modules()
    m = magnet_model("cpu", ns)
    b = S.implicit_batch(B=B, L=L, Nq=Nq, nt=50, d=2, kind="concentrated", seed=600)
    with torch.no_grad():
        t0 = time.perf_counter()
        m.validation_step(b, 0)
        dt = time.perf_counter() - t0
    return {"value": 4 / dt, "unit": "rollout steps/s", "cores": threads, "kind": kind,
            "sample": f"one validation_step (4 autoregressive steps) at the full shape B = {B}, L = Nq = {L}, no warm-up; {how}"}


def bench_rollout(env, clocks=None, B=32, L=256, Nq=256, cpu=True):
    from magnet_b200 import synthetic as S, functional as MF
    args, dev = env.args, env.dev
    MF.set_precision("fp32_tc" if args.precision == "fp32" else args.precision)
    m = magnet_model(dev)
    bh = {k: v.pin_memory() for k, v in S.implicit_batch(B=B, L=L, Nq=Nq, nt=50, d=2, kind="concentrated", seed=600 + env.rank).items()}
    b = {k: v.to(dev) for k, v in bh.items()}
    steps = max(3, args.steps // 2)

    def run(batch=b):
        with torch.no_grad():
            return m.rollout(batch, teacher_forcing=False)[0]

    # our kernels per rollout, counted on an eager pass (the replayed CUDA graph holds the same kernels but launches as one)
    m.cuda_graph = False
    eager_ms, launches = env.timed(run, steps, args.warmup)
    m.cuda_graph = graphed = os.environ.get("MGB_CUDA_GRAPH", "1") != "0"
    ms, _ = env.timed(run, steps, args.warmup, clocks)
    pred = run()
    out_host = torch.empty(pred.shape).pin_memory()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ee0.record()
    for _ in range(steps):
        # host batch in (the coordinate tensors keep their identity so that the cached graphs stay valid: same meshes, new fields)
        for k in ("lr_frames", "hr_points", "t"):
            b[k].copy_(bh[k], non_blocking=True)
        out_host.copy_(run(), non_blocking=True)
    ee1.record()
    env.barrier()
    e2e_ms = env.reduce(ee0.elapsed_time(ee1), "max")
    total_steps = env.reduce(4.0 * steps, "sum")
    if env.rank != 0:
        return None
    hbm, tf, tf_sus, which = _peaks()
    with torch.no_grad():
        u = b["lr_frames"][:, :10].permute(0, 3, 1, 2).reshape(B, L, -1)
        E1 = m._build_graph(u, b["coords_lr"], b["t"][:, :10])[1].shape[1]
        allc = m._all_coords(b["coords_lr"], b["coords_hr"])
        E3 = m._build_graph(torch.cat([u, u[:, :Nq]], 1), allc, b["t"][:, :10])[1].shape[1]
    fl = rollout_flops(B, L, Nq, E1, E3)
    achieved = fl * 4 * steps / (ms * 1e-3) / 1e12
    line = base_line(env, "rollout steps/s", "rollout steps/s", total_steps / (ms * 1e-3), ms / (4 * steps), steps, args.warmup,
                     "bf16" if args.precision == "bf16" else "f32",
                     {"workload": "MAgNet[GNN] validation rollout (encode LR graph, INR decode, LR u HR graph, 10 InteractionNetwork layers; "
                                  "4 autoregressive steps per rollout), BASELINE configs[2]",
                      "batch_per_gpu": B, "lowres_nodes": L, "query_points": Nq, "time_slice": 10, "radius": 0.08, "edges_stage1": E1, "edges_stage3": E3,
                      "parallelism": f"one batch per GPU on {env.world} GPU(s), no collective", "step": "one MAgNetGNN.forward over the batch",
                      "l2": "the per-step working set fits the 126 MB L2 at this shape (the reference's training shape): launch-bound regime"},
                     clocks=clocks.summary() if clocks else None,
                     e2e={"value": total_steps / (e2e_ms * 1e-3), "unit": "rollout steps/s",
                          "h2d_bytes_per_step": sum(bh[k].numel() * 4 for k in ("lr_frames", "hr_points", "t")) // 4,
                          "d2h_bytes_per_step": pred.numel() * 4 // 4, "steps": steps, "result": "predicted fields [B, T_future, Nq+L, 1]"},
                     gpu_launches=launches // 4,
                     cuda_graph={"enabled": graphed, "eager_steps_per_s": 4 * steps / (eager_ms * 1e-3),
                                 "note": "gpu_launches = kernels of this library per step (counted on the eager pass); with the graph they "
                                         "are replayed by ONE cudaGraphLaunch per step plus the input / output copies"},
                     roofline={"bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf,
                               "peak_source": which, "kernel": "whole rollout step (reference-formulation FLOPs over its CUDA-event time)",
                               "algorithmic_flop_per_step": fl, "traffic": None,
                               "note": "16k-node problem: bounded by launch latency / occupancy, not by the tensor pipe"})
    if env.world == 1 and cpu and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_rollout(os.cpu_count() or 1, B, L, Nq)
    return line


# =============================================================================================================
# magnet_train — BASELINE configs[3] (C4): MAgNet[GNN] training, bf16, batch 64 = 8 samples per GPU, batch-sharded, flat
# gradient all-reduce overlapped with backward.  The mesh size per sample is capped by what 180 GB hold (see config.cap).
# =============================================================================================================
C4_SAMPLES_PER_GPU, C4_L, C4_NQ = 8, 16384, 16384


def cpu_magnet_train(threads, B=1, L=2048, Nq=2048):
    from magnet_b200 import synthetic as S
    torch.set_num_threads(threads)
    ns, (kind, how) = ref_modules()
    m = magnet_model("cpu", ns, teacher_forcing=False).train()
    b = S.implicit_batch(B=B, L=L, Nq=Nq, nt=20, d=2, kind="uniform", seed=700)
    t0 = time.perf_counter()
    loss = m.training_step(b, 0)
    loss.backward()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        u = b["lr_frames"][:, :10].permute(0, 3, 1, 2).reshape(B, L, -1)
        E1 = m._build_graph(u, b["coords_lr"], b["t"][:, :10])[1].shape[1]
        E3 = m._build_graph(torch.cat([u, u[:, :Nq]], 1), torch.cat([b["coords_lr"], b["coords_hr"]], 1), b["t"][:, :10])[1].shape[1]
    return {"value": 5 * (E1 + E3) / dt, "unit": "edges/s", "cores": threads, "kind": kind, "ms_per_step": dt * 1e3,
            "sample": f"capped: one training step (forward + backward, no optimizer) of the full model on ONE sample with L = Nq = {L} "
                      f"(radius 0.08: {E1} + {E3} edges), fp32; {how}"}


def bench_magnet_train(env, clocks=None):
    from magnet_b200 import synthetic as S, functional as MF
    from magnet_b200.optim import FlatAdam, OverlappedFlatAllReduce
    args, dev = env.args, env.dev
    MF.set_precision("fp32_tc" if args.precision == "fp32" else args.precision)
    B, L, Nq = C4_SAMPLES_PER_GPU, C4_L, C4_NQ
    m = magnet_model(dev, teacher_forcing=False).train()
    opt = FlatAdam(m.parameters(), lr=1e-5, weight_decay=1e-8)
    ar = OverlappedFlatAllReduce(opt, env.world, n_buckets=4)
    bh = {k: v.pin_memory() for k, v in S.implicit_batch(B=B, L=L, Nq=Nq, nt=20, d=2, kind="uniform", seed=700 + env.rank).items()}
    b = {k: v.to(dev) for k, v in bh.items()}
    steps = max(3, args.steps // 3)
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exposed = []

    def step(batch=b):
        loss = m.training_step(batch, 0)           # one autoregressive step (nt = 20): rollout, both L1 losses
        loss.backward()                            # bucketed all-reduces leave from the gradient hooks
        ev_a.record()
        scale = ar.finish()                        # what is left of the collective: the compute stream waits here
        ev_b.record()
        opt.step(grad_scale=scale)
        opt.zero_grad()
        exposed.append((ev_a, ev_b))
        return loss.detach()

    ms, launches = env.timed(step, steps, 2, clocks)
    torch.cuda.synchronize()
    exposed_ms = ev_a.elapsed_time(ev_b)          # last step's (events are re-recorded every step)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    kb_t, kb_c = env.prof(6)
    kf_t, kf_c = env.prof(5)
    loss_host = torch.zeros(1).pin_memory()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ee0.record()
    for _ in range(steps):
        for k in ("lr_frames", "hr_points", "t"):                      # same meshes (cached graphs), new fields from the host
            b[k].copy_(bh[k], non_blocking=True)
        loss_host.copy_(step().reshape(1), non_blocking=True)
    ee1.record()
    env.barrier()
    e2e_ms = env.reduce(ee0.elapsed_time(ee1), "max")
    with torch.no_grad():
        u = b["lr_frames"][:, :10].permute(0, 3, 1, 2).reshape(B, L, -1)
        E1 = m._build_graph(u, b["coords_lr"], b["t"][:, :10])[1].shape[1]
        E3 = m._build_graph(torch.cat([u, u[:, :Nq]], 1), m._all_coords(b["coords_lr"], b["coords_hr"]), b["t"][:, :10])[1].shape[1]
    edge_layers = env.reduce(5.0 * (E1 + E3), "sum")
    exposed_ms = env.reduce(exposed_ms, "max")
    peak_gb = env.reduce(peak_gb, "max")
    if env.rank != 0:
        return None
    hbm, tf, tf_sus, which = _peaks()
    fl = 3 * (rollout_flops(B, L, Nq, E1, E3))           # forward + data + weight gradients, reference formulation
    achieved = fl * steps / (ms * 1e-3) / 1e12
    n_param = sum(p.numel() for p in m.parameters())
    line = base_line(env, "edges/s per MP layer fwd+bwd", "edges/s", edge_layers * steps / (ms * 1e-3), ms / steps, steps, 3,
                     "bf16" if args.precision == "bf16" else "f32",
                     {"workload": "MAgNet[GNN] full-model training step (encoders, 2 x 5 InteractionNetwork layers, INR decoder, decoder; L1 losses; "
                                  "flat Adam), synthetic 2-D irregular-uniform meshes, BASELINE configs[3]",
                      "samples_per_gpu": B, "global_batch": B * env.world, "lowres_nodes": L, "query_points": Nq, "nodes_per_sample": L + Nq,
                      "edges_stage1": E1, "edges_stage3": E3, "radius": 0.08, "time_slice": 10, "rollout_steps_per_training_step": 1,
                      "cap": "configs[3] names 1 M-node meshes: at 32 in-edges per node e_features alone are 16 GB per sample and the edge encoder's "
                             "saved activations ~4x that; the largest power-of-two mesh that trains with 8 samples per GPU in 180 GB is used "
                             "(the InteractionNetwork layers themselves save 1 KB per node: the limit is the edge encoder, DESIGN.md §7.6)",
                      "parallelism": f"batch sharded over {env.world} GPU(s); one flat-buffer gradient all-reduce in 4 buckets launched from gradient hooks, overlapped with backward",
                      "l2": "per-layer working sets (GBs) exceed the 126 MB L2; no explicit flush"},
                     clocks=clocks.summary() if clocks else None,
                     train_steps_per_s=steps / (ms * 1e-3) , samples_per_s=env.world * B * steps / (ms * 1e-3),
                     e2e={"value": edge_layers * steps / (e2e_ms * 1e-3), "unit": "edges/s",
                          "h2d_bytes_per_step": sum(bh[k].numel() * 4 for k in ("lr_frames", "hr_points", "t")), "d2h_bytes_per_step": 4, "steps": steps,
                          "result": "the training loss (4 bytes) is the step's result"},
                     gpu_launches=launches,
                     allreduce={"bytes": n_param * 4, "buckets": len(ar.buckets), "exposed_ms_last_step": exposed_ms,
                                "note": "exposed = time the compute stream waits in finish() after backward (CUDA events); the rest of the collective ran under backward"},
                     peak_memory_gb=peak_gb,
                     roofline={"bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf, "peak_source": which,
                               "kernel": "whole training step (reference-formulation FLOPs, forward + data + weight gradients, over its CUDA-event time)",
                               "algorithmic_flop_per_step": fl, "in_edge_bwd_ms_per_call": kb_t / max(kb_c, 1), "in_edge_bwd_calls_per_step": kb_c / max(steps, 1),
                               "in_edge_fwd_ms_per_call": kf_t / max(kf_c, 1), "traffic": None})
    if env.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_magnet_train(os.cpu_count() or 1)
    return line


# =============================================================================================================
def run_reference(args, env):
    """--impl reference: the reference's own CPU implementation of the metric, all host threads, rank 0 only."""
    if env.rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.metric == "in_layer":
        cb = cpu_in_layer(threads)
        metric, unit, cfg = "edges/s per MP layer fwd+bwd", "edges/s", in_config(env, IN_SAMPLES, IN_NODES_PER_SAMPLE)
    elif args.metric == "inr_decode":
        cb = cpu_inr_decode(threads)
        metric, unit, cfg = "query points/s, INR decode", "query points/s", {"workload": "continuous_decoder + projector, BASELINE configs[4] shape (capped)"}
    elif args.metric == "magnet_train":
        cb = cpu_magnet_train(threads)
        metric, unit, cfg = "edges/s per MP layer fwd+bwd", "edges/s", {"workload": "MAgNet[GNN] full-model training step, BASELINE configs[3] (capped)"}
    elif args.metric == "rollout":
        cb = cpu_rollout(threads)
        metric, unit, cfg = "rollout steps/s", "rollout steps/s", {"workload": "MAgNet[GNN] validation rollout, BASELINE configs[2]"}
    else:
        cb = cpu_mp_layer(1, max(args.steps, 1), min(args.warmup, 1), threads)
        metric, unit, cfg = "edges/s per MP layer fwd+bwd", "edges/s", mp_config(args.gpus, args.precision)
    value = cb["value"]
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--metric", default="mp_layer", choices=["mp_layer", "in_layer", "inr_decode", "rollout", "inr_sweep", "rollout_res256", "magnet_train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline run without the other metrics")
    ap.add_argument("--precision", default="fp32_tc", choices=["fp32", "fp32_tc", "bf16"],
                    help="arithmetic of the 128-wide contractions: fp32 FFMA | tcgen05 hi/lo split (1e-5 contract) | tcgen05 bf16 (1e-2)")
    args = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    env = Env(args)
    if args.impl == "reference":
        run_reference(args, env)
        return
    env.init_gpu()
    import torch.distributed as dist
    if args.metric == "inr_sweep":
        run_inr_sweep(env)
    else:
        with ClockSampler(env.local_rank) as clocks:
            if args.metric == "mp_layer":
                line = bench_mp_layer(env, clocks)
                if not args.no_extra:
                    others = {}
                    for name, fn in (("in_layer", bench_in_layer), ("inr_decode", bench_inr_decode), ("rollout", bench_rollout)):
                        with ClockSampler(env.local_rank) as c2:
                            others[name] = fn(env, c2)
                    if line is not None:
                        line["metrics"] = others
            elif args.metric == "in_layer":
                line = bench_in_layer(env, clocks)
            elif args.metric == "inr_decode":
                line = bench_inr_decode(env, clocks)
            elif args.metric == "magnet_train":
                line = bench_magnet_train(env, clocks)
            elif args.metric == "rollout_res256":
                line = bench_rollout(env, clocks, B=1, L=32768, Nq=32768, cpu=False)
            else:
                line = bench_rollout(env, clocks)
        if env.rank == 0:
            print(json.dumps(line), flush=True)
    if env.world > 1 and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
