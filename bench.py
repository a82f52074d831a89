#!/usr/bin/env python
"""Benchmark of the DDiffPG hot path on B200 (contract: see the task prompt / DESIGN.md section 6).

Default line = BASELINE.json configs[1]: antmaze-v1 actor shapes (S=34, A=8, T=5, trunk 1024/512/256), synthetic
batch of 65,536 states per GPU, one fused T-step sampler launch per step; metric: denoised actions/sec (whole job).
The same line carries ``secondary``: the two other pieces of the hot path at the sizes of configs[2] / configs[3]
(1 M rows over 8 GPUs = 131,072 rows per GPU, weak scaling) -- the mode-conditioned Q-gradient action ascent (H2) and
the denoiser training step with its gradient all-reduce (H3), each with its own roofline fraction, end-to-end number
and, under torchrun, the cost of the collective and a replica-consistency self-check.
One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload sample|ascent|train] [--precision bf16|fp32] [--batch B] [--T T] [--width h]
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, A, O = 34, 8, 29
L2_FLUSH_BYTES = 256 << 20
CPU_SAMPLE_ROWS = {"sample": 4096, "ascent": 2048, "train": 4096, "critic": 4096}     # rows per step of the CPU arms (bounded samples)
SECONDARY_ROWS = 131072                                                 # configs[2] / [3]: 1 M rows over 8 GPUs
UNITS = {"sample": "actions/s", "ascent": "states/s", "train": "rows/s", "critic": "rows/s"}
Q_HID, Q_ATOMS = (512, 256, 128), 51           # DistributionalDoubleQ defaults (reference mlp.py:131-141)

_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, torchrun notices) goes to stderr; the JSON line is
    written to the original stdout by emit(), so stdout carries exactly one line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def algorithmic_flops(workload, T, h):
    """minimal FLOPs after removing batch- and loop-invariant terms."""
    if workload == "sample":
        return 2.0 * (T * (0.625 * h * h + 10 * h) + 34 * h)              # per action (6 725 632 at T=5, h=1024)
    if workload == "ascent":
        return 20 * 1395712.0 + 59392.0                                    # per state, 20 iterations
    if workload == "critic":
        # per row, two nets each: forward of the target critic and of the critic, dX through layers 4..2, dW of all four
        h1, h2, h3 = Q_HID
        fwd = (O + A) * h1 + h1 * h2 + h2 * h3 + h3 * Q_ATOMS
        return 2.0 * 2 * (fwd + fwd + (fwd - (O + A) * h1) + fwd)          # 2 953 216
    return 2.0 * (2 * (42 * h + 0.625 * h * h + 2 * h) + (0.625 * h * h + 2 * h))   # train: fwd + dW + dX (4 116 480)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def kernel_source_hash(files):
    """sha256 (16 hex) over the kernel sources an ncu traffic figure was captured from: profiles/traffic.json entries
    carry the hash of the day of the capture, and a figure whose sources have changed since is not reported."""
    h = hashlib.sha256()
    for name in files:
        path = os.path.join(ROOT, "ddiffpg_b200", "csrc", name)
        if not os.path.exists(path):
            return None
        with open(path, "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def metric_name(workload):
    return {"sample": "denoised actions/sec (T-step chain)", "ascent": "Q-ascent states/sec (20 Adam iterations)",
            "train": "denoiser train rows/sec (eps-loss fwd+bwd)",
            "critic": "critic update rows/sec (C51 target + BCE fwd+bwd)"}[workload]


def workload_config(args, workload, rows):
    cfg = {"workload": {"sample": "antmaze-v1 actor shapes, fused T-step sampler (BASELINE configs[1])",
                        "ascent": "mode-conditioned double-Q action ascent (BASELINE configs[2])",
                        "train": "denoiser eps-loss fwd+bwd + gradient all-reduce + clip/AdamW (BASELINE configs[3])",
                        "critic": "critic update of one mode: target heads, C51 projection, BCE, backward, gradient "
                                  "all-reduce, clip + AdamW on the flat vector, one CUDA graph (SURVEY 8f row N1)"}[workload],
           "rows_per_gpu": rows, "S": S, "A": A, "T": args.T, "trunk": [args.width, args.width // 2, args.width // 4],
           "precision_path": args.precision, "l2": "flushed between timed steps (256 MiB write, untimed)",
           "cpu_arm_rows_per_step": CPU_SAMPLE_ROWS[workload]}
    if workload == "ascent":
        cfg["modes"] = args.modes
    return cfg


# ------------------------------------------------------------------------------------------ CPU arms (oracle port)
def cpu_fn(workload, p, T, rows, seed=0):
    """One pass of the reference's CPU path (the oracle port of its modules: torch CPU fp32) over `rows` synthetic rows."""
    import torch
    from oracle import port        # the checker's CPU restatement, timed as the baseline only
    g = torch.Generator().manual_seed(seed)
    if workload == "sample":
        cs, cn = torch.randn(rows, S, generator=g), torch.randn(T, rows, A, generator=g)
        return lambda: port.actor_sample(p, cs, cn, T)
    if workload == "ascent":
        co, ca = torch.randn(rows, O, generator=g), torch.rand(rows, A, generator=g) * 2 - 1
        return lambda: port.q_action_ascent(p, co, ca.clone(), iters=20)
    if workload == "critic":
        co, cn = torch.randn(rows, O, generator=g), torch.randn(rows, O, generator=g)
        ca, cb = torch.rand(rows, A, generator=g) * 2 - 1, torch.rand(rows, A, generator=g) * 2 - 1
        cr, cd = torch.rand(rows, 1, generator=g), (torch.rand(rows, 1, generator=g) < 0.2).float()

        def critic_step():
            tq = port.critic_target_dist(p, cn, cb, cr, cd, 0.97).clamp_max(1.0)
            return port.critic_loss_and_grads(p, tq, co, ca)
        return critic_step
    c1, c2 = torch.randn(rows, S, generator=g), torch.rand(rows, A, generator=g) * 2 - 1
    c3, c4 = torch.randn(rows, A, generator=g), torch.randint(0, T, (rows,), generator=g)
    return lambda: port.adamw_train_step(p, c1, c2, c3, c4, T)


def cpu_params(workload, h):
    from oracle import port
    return port.init_critic_params(0) if workload in ("ascent", "critic") else port.init_actor_params(0, h=h)


def cpu_baseline(workload, p, T, budget_s, with_configs0):
    """Bounded sample of the CPU path on this box's host cores: >= 1 pass, at most `budget_s` seconds or 50 passes."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = CPU_SAMPLE_ROWS[workload]
    fn = cpu_fn(workload, p, T, rows)
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 50:
            break
    out = {"value": rows * n / dt, "unit": UNITS[workload], "cores": cores, "kind": "port",
           "sample": f"{n} passes of {rows} rows in {dt:.1f} s, torch CPU fp32, {cores} threads (oracle port of the reference modules)"}
    if with_configs0:
        # BASELINE configs[0]: the reference's own 256-row call (latency-bound on the CPU as well)
        f0 = cpu_fn(workload, p, T, 256, seed=1)
        f0(); f0()
        ts = []
        for _ in range(5):
            t1 = time.perf_counter(); f0(); ts.append(time.perf_counter() - t1)
        out["configs0"] = {"rows": 256, "ms_per_call": min(ts) * 1e3, "value": 256 / min(ts), "unit": UNITS[workload],
                           "how": "best of 5 after 2 warm-ups"}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port, torch CPU, all host threads), each
    step a bounded sample (`cpu_arm_rows_per_step` rows) of the workload `config` names.  Rank 0 only."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = args.workload
    rows = CPU_SAMPLE_ROWS[wl]
    fn = cpu_fn(wl, cpu_params(wl, args.width), args.T, rows)
    for _ in range(max(1, min(args.warmup, 2))):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = (f"{rows} rows per step x {args.steps} steps (a bounded sample of the {args.batch}-row workload), torch CPU fp32, "
              f"{cores} threads (oracle port of the reference modules)")
    line = {"impl": "reference", "metric": metric_name(wl), "value": value, "unit": UNITS[wl],
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "rows_per_step": rows, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, wl, args.batch),
            "cpu_baseline": {"value": value, "unit": UNITS[wl], "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNITS[wl], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------ CUDA arm
class Ctx:
    """Per-process state shared by the workload legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.dev)
        self.gen = torch.Generator().manual_seed(1234 + self.rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def timed(self, fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed (untimed) before each; returns the
        summed device time in ms (max over ranks taken by the caller)."""
        torch = self.torch
        total = 0.0
        for _ in range(steps):
            self.flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total


def setup_sample(cx, rows):
    torch, args, dev = cx.torch, cx.args, cx.dev
    from ddiffpg_b200 import DiffusionPolicy
    from ddiffpg_b200._lib import check, lib, ptr, stream_ptr
    T, h = args.T, args.width
    torch.manual_seed(0)                          # reference default init (nn.Linear), random weights
    pol = DiffusionPolicy(S, A, T, device="cuda", hidden=(h, h // 2, h // 4), precision=args.precision)
    params = {k: v.clone() for k, v in pol.state_dict().items()}
    pol.to(dev)
    state_h = torch.randn(rows, S, generator=cx.gen).pin_memory()
    state = state_h.to(dev)
    noise = torch.randn(T, rows, A, generator=cx.gen).to(dev)
    out_h = torch.empty(rows, A).pin_memory()
    # kernel-only leg: the same C-ABI call get_actions() makes, with the buffers resolved once, so that no Python
    # work sits between the start event and the launch (it showed up as +0.15 ms per step under torchrun)
    pol.get_actions(state[:256], noise=noise[:, :256].contiguous())     # builds the packed weights
    packed, shape, prec = pol._packed(args.precision, need=1)
    ws_bytes = lib().ddp_actor_sample_workspace_bytes(shape, rows, prec)
    ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
    out_d = torch.empty(rows, A, device=dev)
    sample_args = (shape, ptr(packed), ptr(state), ptr(noise), ptr(out_d), rows, prec, ptr(ws), ws_bytes)
    fn_sample = lib().ddp_actor_sample

    def step():
        check(fn_sample(*sample_args, stream_ptr()), "ddp_actor_sample")

    def e2e_step():
        # the call a user with host-resident observations makes: pinned obs up, noise drawn on the device,
        # actions back down (H2D / D2H of row chunks overlap the sampler launches)
        pol.get_actions_host(state_h, out_h)
    return {"step": step, "e2e": e2e_step, "launches": 1, "h2d": state_h.numel() * 4, "d2h": out_h.numel() * 4,
            "params": params, "keep": (pol, state, noise, ws, out_d)}


def setup_ascent(cx, rows):
    torch, args, dev = cx.torch, cx.args, cx.dev
    from ddiffpg_b200 import DistributionalDoubleQ, q_action_ascent_segments
    from ddiffpg_b200.models import _PackCache
    K = args.modes
    critics, params = [], None
    for m in range(K):
        torch.manual_seed(m)
        c = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=51, device="cuda")
        if m == 0:
            params = {k: v.clone() for k, v in c.state_dict().items()}
        critics.append(c.to(dev).requires_grad_(False))
    seg = [rows * m // K for m in range(K + 1)]
    obs_h = torch.randn(rows, O, generator=cx.gen).pin_memory()
    act_h = (torch.rand(rows, A, generator=cx.gen) * 2 - 1).pin_memory()
    obs, act0 = obs_h.to(dev), act_h.to(dev)
    work = act0.clone()
    out_h = torch.empty(rows, A).pin_memory()
    cache = _PackCache()

    def step():
        work.copy_(act0)
        q_action_ascent_segments(critics, obs, work, seg, iters=20, cache=cache, precision=args.precision)

    def e2e_step():
        o = obs_h.to(dev, non_blocking=True)
        w = act_h.to(dev, non_blocking=True)
        q_action_ascent_segments(critics, o, w, seg, iters=20, cache=cache, precision=args.precision)
        out_h.copy_(w, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    def step_global():
        # the same rows as a slice of a world x larger mode-sorted batch: global 1/B and the clip norm of the whole mode
        # batch through the K-float exchange step of every iteration (ddp_q_action_ascent_sharded)
        work.copy_(act0)
        return q_action_ascent_segments(critics, obs, work, seg, iters=20, cache=cache, precision=args.precision,
                                        mean_counts=[(seg[m + 1] - seg[m]) * cx.world for m in range(K)],
                                        process_group=cx.dist.group.WORLD, return_norms=True)
    return {"step": step, "e2e": e2e_step, "launches": 2 + 20 * 2 + 2, "h2d": (obs_h.numel() + act_h.numel()) * 4,
            "d2h": out_h.numel() * 4, "params": params, "keep": (critics, obs, work), "step_global": step_global}


def setup_train(cx, rows):
    torch, args, dev = cx.torch, cx.args, cx.dev
    from ddiffpg_b200 import DiffusionPolicy, FusedActorTrainer
    T, h = args.T, args.width
    torch.manual_seed(0)
    pol = DiffusionPolicy(S, A, T, device="cuda", hidden=(h, h // 2, h // 4))
    params = {k: v.clone() for k, v in pol.state_dict().items()}
    pol.to(dev)
    trainer = FusedActorTrainer(pol, precision=args.precision, graph=not args.no_graph, buckets=args.train_buckets)
    st_h = torch.randn(rows, S, generator=cx.gen).pin_memory()
    ac_h = (torch.rand(rows, A, generator=cx.gen) * 2 - 1).pin_memory()
    st, ac = st_h.to(dev), ac_h.to(dev)
    nz = torch.randn(rows, A, generator=cx.gen).to(dev)
    ts = torch.randint(0, T, (rows,), generator=cx.gen).to(dev)
    loss_h = torch.empty(2).pin_memory()

    def step():
        trainer.step(st, ac, noise=nz, timesteps=ts)

    def e2e_step():
        s = st_h.to(dev, non_blocking=True)
        a = ac_h.to(dev, non_blocking=True)
        loss, gn = trainer.step(s, a)
        loss_h[0:1].copy_(loss.reshape(1), non_blocking=True)
        loss_h[1:2].copy_(gn.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return {"step": step, "e2e": e2e_step, "launches": trainer.launches_per_step(), "h2d": (st_h.numel() + ac_h.numel()) * 4,
            "d2h": 8, "params": params, "trainer": trainer, "pol": pol, "inputs": (st, ac, nz, ts), "keep": (st_h, ac_h)}


def setup_critic(cx, rows):
    """N1: update_critic for one mode's critic as FusedCriticTrainer runs it -- fused target / loss / backward
    (ddp_q_critic_loss_fwd_bwd), the flat gradient averaged over the ranks (one all-reduce), clip + AdamW on the flat
    parameter vector (ddp_clip_adamw_step_dev), the whole update replayed as one CUDA graph.  `--no-graph`: the same with
    eager launches; `--torch-tail`: update_critic with torch's own clip_grad_norm_ + AdamW instead."""
    torch, args, dev = cx.torch, cx.args, cx.dev
    from ddiffpg_b200 import DistributionalDoubleQ, FusedCriticTrainer, update_critic
    torch.manual_seed(0)
    critic = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=Q_ATOMS, device="cuda").to(dev)
    params = {k: v.clone().cpu() for k, v in critic.state_dict().items()}
    torch.manual_seed(1)
    target = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=Q_ATOMS, device="cuda").to(dev).requires_grad_(False)
    critic.train_precision = args.precision
    host = [torch.randn(rows, O, generator=cx.gen), torch.rand(rows, A, generator=cx.gen) * 2 - 1,
            torch.rand(rows, 1, generator=cx.gen), torch.randn(rows, O, generator=cx.gen),
            torch.rand(rows, A, generator=cx.gen) * 2 - 1, (torch.rand(rows, 1, generator=cx.gen) < 0.2).float()]
    host = [t.pin_memory() for t in host]
    devt = [t.to(dev) for t in host]
    loss_h = torch.empty(2).pin_memory()
    group = cx.dist.group.WORLD if cx.world > 1 else None
    trainer = None
    if args.torch_tail:
        opt = torch.optim.AdamW(critic.parameters(), lr=5e-4)
        run = lambda batch: update_critic(critic, target, opt, *batch, gamma_n=0.97, max_grad_norm=1.0, process_group=group,
                                          sync=False)[1:]
    else:
        trainer = FusedCriticTrainer(critic, target, lr=5e-4, process_group=group if group is not None else False,
                                     graph=not args.no_graph)
        run = lambda batch: trainer.step(*batch, gamma_n=0.97)

    def step():
        run(devt)

    def e2e_step():
        up = [t.to(dev, non_blocking=True) for t in host]
        loss, gn = run(up)
        loss_h[0:1].copy_(loss.reshape(1), non_blocking=True)
        loss_h[1:2].copy_(gn.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    # launches of this repo's kernels per update.  Tensor path: 2 packs x (biases + 2 nets) = 6, target pass (prep, chain,
    # projection) = 3, prep + column map = 2, 8 forward GEMMs, BCE, 2 x (4 dW + 3 dX) = 14: 34 (+ norm, scalars, clip/AdamW
    # = 37 with the fused tail).  FMA path: 2 packs x 25, target tile kernel + projection, tile kernel, 8 dW: 61 (64).
    base = 34 if args.precision == "bf16" else 61
    leg = {"step": step, "e2e": e2e_step, "launches": base if args.torch_tail else base + 3,
           "h2d": sum(t.numel() for t in host) * 4, "d2h": 8, "params": params, "keep": (critic, target, devt)}
    if trainer is not None:
        leg["trainer"] = trainer
    return leg


SETUP = {"sample": setup_sample, "ascent": setup_ascent, "train": setup_train, "critic": setup_critic}


def train_collective_report(cx, leg, rows, steps):
    """H3 under torchrun: what the one collective of the path costs and whether the replicas agree.
    * `allreduce_us`: the gradient all-reduce of one step alone (flat fp32 gradient + loss, CUDA events, max over ranks);
    * `local_ms_per_step`: the same step on the same rows with the collective switched off (a trainer without a process
      group, timed on every rank at once) -- `exposed_us` = ms_per_step - local_ms_per_step is what the collective adds;
    * replica check: parameter checksum identical on all ranks after the timed steps;
    * gradient check (fp32 path, 512 rows per rank): all-reduced shard gradients == gradient of the gathered batch."""
    torch, dist, args, dev = cx.torch, cx.dist, cx.args, cx.dev
    from ddiffpg_b200 import DiffusionPolicy, FusedActorTrainer
    from ddiffpg_b200 import dist as ddist
    trainer = leg["trainer"]
    out = {}
    # -- the collective alone
    g = torch.zeros(trainer.flat.numel() + 1, device=dev)
    for _ in range(5):
        ddist.allreduce_sum_(g)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ddist.allreduce_sum_(g)
    e1.record(); e1.synchronize()
    out["allreduce_us"] = cx.max_over_ranks([e0.elapsed_time(e1) / n * 1e3])[0]
    out["allreduce_bytes"] = g.numel() * 4
    # -- the same step without the collective
    torch.manual_seed(0)
    h, T = args.width, args.T
    pol2 = DiffusionPolicy(S, A, T, device="cuda", hidden=(h, h // 2, h // 4)).to(dev)
    local = FusedActorTrainer(pol2, precision=args.precision, graph=not args.no_graph, process_group=False)
    st, ac, nz, ts = leg["inputs"]
    lstep = lambda: local.step(st, ac, noise=nz, timesteps=ts, global_batch=rows * cx.world)
    for _ in range(max(args.warmup, 3)):
        lstep()
    cx.barrier()
    ms_local = cx.max_over_ranks([cx.timed(lstep, steps)])[0] / steps
    out["local_ms_per_step"] = ms_local
    local.close()
    # -- replicas identical after the steps every rank has taken
    flat = trainer.flat.double()
    mine = torch.stack([flat.sum(), flat.abs().sum()])
    allc = [torch.zeros_like(mine) for _ in range(cx.world)]
    dist.all_gather(allc, mine)
    allc = torch.stack(allc)
    out["replica_checksum_spread"] = float((allc.max(0).values - allc.min(0).values).abs().max().item())
    out["replicas_identical"] = out["replica_checksum_spread"] == 0.0
    # -- reduced shard gradients == full-batch gradient (fp32 path: the 1e-5 statement is about the collective, not bf16)
    nb = 512
    gen = torch.Generator().manual_seed(99 + cx.rank)
    sh = [torch.randn(nb, S, generator=gen), torch.rand(nb, A, generator=gen) * 2 - 1, torch.randn(nb, A, generator=gen),
          torch.randint(0, T, (nb,), generator=gen).float()]
    sh = [x.to(dev) for x in sh]
    torch.manual_seed(0)
    pol3 = DiffusionPolicy(S, A, T, device="cuda", hidden=(h, h // 2, h // 4)).to(dev)
    inv = 1.0 / (nb * cx.world * A)
    l_sh, g_sh = pol3._loss_and_grads(sh[0], sh[1], sh[2], sh[3].long(), inv_count=inv, precision="fp32")
    g_sh, l_sh = g_sh.clone(), l_sh.clone()
    ddist.allreduce_sum_(g_sh, l_sh)
    full = []
    for x in sh:
        parts = [torch.zeros_like(x) for _ in range(cx.world)]
        dist.all_gather(parts, x)
        full.append(torch.cat(parts))
    l_full, g_full = pol3._loss_and_grads(full[0], full[1], full[2], full[3].long(), inv_count=inv, precision="fp32")
    rel = ((g_sh - g_full).norm() / g_full.norm()).item()
    out["reduced_vs_full_batch_grad_rel_l2"] = rel
    out["reduced_grad_ok"] = rel <= 1e-5 and abs(l_sh.item() - l_full.item()) <= 1e-5 * abs(l_full.item())
    return out


def ascent_exchange_report(cx, leg, steps, ms_local_per_step):
    """H2 under torchrun: the default is shard-local semantics (no collective, the timed line); this times the same rows
    with global-batch semantics -- 20 all-reduces of K floats per ascent, one between every gradient pass and Adam step --
    and checks that every rank saw the same global norms."""
    torch, dist = cx.torch, cx.dist
    step = leg["step_global"]
    for _ in range(3):
        step()
    cx.barrier()
    ms = cx.max_over_ranks([cx.timed(step, steps)])[0] / steps
    _, norms = step()
    allc = [torch.zeros_like(norms) for _ in range(cx.world)]
    dist.all_gather(allc, norms)
    return {"default_semantics": "shard-local (no collective)", "global_batch_ms_per_step": ms,
            "exchange_steps_per_ascent": 20, "exchange_floats": int(norms.shape[0]),
            "exposed_us_per_exchange": (ms - ms_local_per_step) * 1e3 / 20,
            "global_norms_identical_on_all_ranks": all(torch.equal(allc[0], c) for c in allc)}


def run_leg(cx, workload, rows, steps, want_cpu, with_configs0, cpu_budget_s):
    """Set one workload up, time its device-resident and end-to-end forms, and describe it as a dict."""
    torch, args = cx.torch, cx.args
    leg = SETUP[workload](cx, rows)
    step, e2e_step = leg["step"], leg["e2e"]
    for _ in range(max(args.warmup, 3)):
        step()
    # a fresh box ramps its clocks over the first milliseconds of load: a few more untimed iterations of exactly the timed
    # pattern (L2 flush + step) -- back-to-back launches without the flush would instead push the part into its power cap
    cx.timed(step, 20)
    cx.barrier()
    sampler = ClockSampler(cx.local)
    if cx.rank == 0:
        sampler.start()
    cx.barrier()
    ms = cx.timed(step, steps)
    # nvidia-smi polls every 100 ms and K steps can be shorter than one poll: keep the identical loop (flush + step)
    # running, untimed, until the sampler has seen >= 0.6 s of this load pattern, then stop it
    # (same step count on every rank: the training step contains a collective)
    ms_all = cx.max_over_ranks([ms])[0]
    per_step = max(ms_all / steps, 0.05) + 0.15          # + the untimed L2 flush
    cx.timed(step, min(2000, max(0, int((600.0 - ms_all) / per_step))))
    cx.barrier()
    clocks = sampler.stop() if cx.rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed steps + continuation of the same flush/step loop to 0.6 s (100 ms polls)"
    for _ in range(2):
        e2e_step()
    cx.barrier()
    ms_e2e = cx.timed(e2e_step, steps)
    cx.barrier()
    ms, ms_e2e = cx.max_over_ranks([ms, ms_e2e])
    extra = {}
    if workload == "train" and cx.world > 1:
        extra = train_collective_report(cx, leg, rows, steps)
        extra["exposed_us"] = (ms / steps - extra["local_ms_per_step"]) * 1e3
    if workload == "ascent" and cx.world > 1:
        extra = ascent_exchange_report(cx, leg, steps, ms / steps)
    if "trainer" in leg:
        leg["trainer"].close()            # graphs that captured the all-reduce go before the process group does
    if cx.rank != 0:
        return None
    T, h = args.T, args.width
    unit = UNITS[workload]
    peaks = measured_peaks()
    flops_unit = algorithmic_flops(workload, T, h)
    achieved = rows * flops_unit / (ms / steps * 1e-3) / 1e12            # per GPU, the step is the kernel sequence
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": None,
                "peak_source": f"{peaks['source']} bf16 burst (MEASURED_PEAKS.json)",
                "algorithmic_flops_per_unit": flops_unit,
                "hbm_bytes_per_unit": {"sample": 4 * (S + T * A + A), "ascent": 180, "train": 208,
                                       "critic": 4 * (2 * (O + A) + 2)}[workload]}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            ent = json.load(f).get(f"{workload}_{args.precision}")
        # one ncu --set full capture of the dominant kernel at exactly this workload size (never measured here: a number
        # taken under a profiler is not a bench value); reported only while the kernel sources still hash to what was profiled
        if (isinstance(ent, dict) and ent.get("rows") == rows and ent.get("T") == T and ent.get("width") == h
                and ent.get("source_hash") == kernel_source_hash(ent.get("source_files", []))):
            roofline["traffic"] = ent["bytes"]
            roofline["traffic_source"] = ent["source"]
    res = {"metric": metric_name(workload), "value": cx.world * rows * steps / (ms * 1e-3), "unit": unit,
           "ms_per_step": ms / steps, "config": workload_config(args, workload, rows), "clocks": clocks,
           "e2e": {"value": cx.world * rows * steps / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": leg["h2d"],
                   "d2h_bytes_per_step": leg["d2h"], "ms_per_step": ms_e2e / steps},
           "gpu_launches": leg["launches"] * steps, "roofline": roofline}
    if extra:
        res["collective"] = extra
    if want_cpu:
        res["cpu_baseline"] = cpu_baseline(workload, leg["params"], T, cpu_budget_s, with_configs0)
    return res


def run_cuda(args):
    cx = Ctx(args)
    steps = args.steps
    want_cpu = cx.world == 1 and not args.no_cpu_baseline
    main = run_leg(cx, args.workload, args.batch, steps, want_cpu, True, 10.0)
    secondary = {}
    if args.workload == "sample" and not args.no_secondary:
        for wl in ("ascent", "train", "critic"):
            secondary[wl] = run_leg(cx, wl, args.secondary_batch, max(3, min(steps, 10)), want_cpu, False, 4.0)
    if cx.rank == 0:
        line = {"metric": main["metric"], "value": main["value"], "unit": main["unit"], "n_gpus": cx.world, "steps": steps,
                "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic", "config": main["config"], "clocks": main["clocks"], "e2e": main["e2e"],
                "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
                "cpu_baseline": main.get("cpu_baseline")}
        if "collective" in main:
            line["collective"] = main["collective"]
        if secondary:
            line["secondary"] = secondary
        emit(line)
    finish(cx)


def finish(cx):
    """End of a rank's work: every CUDA graph that captured a collective has been released by now (run_leg closes the
    trainers), so the process group can be torn down the ordinary way."""
    sys.stdout.flush()
    sys.stderr.flush()
    if cx.world > 1:
        cx.torch.cuda.synchronize()
        cx.dist.barrier()
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="sample", choices=["sample", "ascent", "train", "critic"])
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"],
                    help="default: bf16 tensor-core paths (fp32 = warp-FMA parity paths)")
    ap.add_argument("--batch", type=int, default=65536, help="rows per GPU of the primary workload")
    ap.add_argument("--secondary-batch", type=int, default=SECONDARY_ROWS,
                    help="rows per GPU of the secondary legs (ascent, train) of the default line")
    ap.add_argument("--no-secondary", action="store_true", help="default line without the ascent / train legs")
    ap.add_argument("--T", type=int, default=5)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--modes", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-buckets", type=int, default=None, choices=[1, 2, 4],
                    help="train workload under torchrun: gradient all-reduces per step (default: by world size)")
    ap.add_argument("--torch-tail", action="store_true",
                    help="critic workload: update_critic with torch's own clip_grad_norm_ + AdamW instead of FusedCriticTrainer")
    ap.add_argument("--no-graph", action="store_true", help="train workload: eager launches instead of the CUDA graph")
    args = ap.parse_args()
    quiet_stdout()
    if args.precision is None:
        args.precision = os.environ.get("DDP_BENCH_PRECISION", "bf16")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
