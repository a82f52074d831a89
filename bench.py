#!/usr/bin/env python
"""Benchmark of the DDiffPG hot path on B200 (contract: see the task prompt / DESIGN.md section 6).

Default workload = BASELINE.json configs[1]: antmaze-v1 actor shapes (S=34, A=8, T=5, trunk
1024/512/256), synthetic batch of 65,536 states per GPU, one fused T-step sampler launch per step.
Metric: denoised actions/sec (whole job, all ranks).  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload sample|ascent|train] [--precision bf16|fp32] [--batch B] [--T T] [--width h]
"""
import argparse
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
    return 2.0 * (2 * (42 * h + 0.625 * h * h + 2 * h) + (0.625 * h * h + 2 * h))   # train: fwd + dW + dX (4 116 480)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


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


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port, torch CPU, all host threads), on a
    bounded sample of the same workload.  Rank 0 only."""
    import torch
    from oracle import port
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T, h = args.T, args.width
    sample_rows = {"sample": 4096, "ascent": 2048, "train": 4096}[args.workload]
    gen = torch.Generator().manual_seed(0)
    if args.workload == "sample":
        p = port.init_actor_params(0, h=h)
        state, noise = torch.randn(sample_rows, S, generator=gen), torch.randn(T, sample_rows, A, generator=gen)
        fn = lambda: port.actor_sample(p, state, noise, T)
    elif args.workload == "ascent":
        p = port.init_critic_params(0)
        obs, act = torch.randn(sample_rows, O, generator=gen), torch.rand(sample_rows, A, generator=gen) * 2 - 1
        fn = lambda: port.q_action_ascent(p, obs, act.clone(), iters=20)
    else:
        p = port.init_actor_params(0, h=h)
        st = torch.randn(sample_rows, S, generator=gen); ac = torch.rand(sample_rows, A, generator=gen) * 2 - 1
        nz = torch.randn(sample_rows, A, generator=gen); ts = torch.randint(0, T, (sample_rows,), generator=gen)
        fn = lambda: port.actor_loss_and_grads(p, st, ac, nz, ts, T)
    for _ in range(max(1, min(args.warmup, 2))):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = sample_rows * args.steps / dt
    unit = {"sample": "actions/s", "ascent": "states/s", "train": "rows/s"}[args.workload]
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch),
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                             "sample": f"{sample_rows} rows per step x {args.steps} steps, torch CPU fp32, "
                                       f"{cores} threads (oracle port of the reference modules)"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def metric_name(workload):
    return {"sample": "denoised actions/sec (T-step chain)", "ascent": "Q-ascent states/sec (20 Adam iterations)",
            "train": "denoiser train rows/sec (eps-loss fwd+bwd)"}[workload]


def workload_config(args, rows):
    return {"workload": {"sample": "antmaze-v1 actor shapes, fused T-step sampler (BASELINE configs[1])",
                         "ascent": "mode-conditioned double-Q action ascent (BASELINE configs[2])",
                         "train": "denoiser eps-loss fwd+bwd (BASELINE configs[3])"}[args.workload],
            "rows_per_gpu": rows, "S": S, "A": A, "T": args.T, "trunk": [args.width, args.width // 2, args.width // 4],
            "precision_path": args.precision, "l2": "flushed between timed steps (256 MiB write, untimed)"}


# ------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from ddiffpg_b200 import DiffusionPolicy, DistributionalDoubleQ, FusedActorTrainer, q_action_ascent_segments

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, h, B = args.T, args.width, args.batch
    hidden = (h, h // 2, h // 4)
    gen = torch.Generator().manual_seed(1234 + rank)
    unit = {"sample": "actions/s", "ascent": "states/s", "train": "rows/s"}[args.workload]
    launches = 0

    if args.workload == "sample":
        torch.manual_seed(0)                          # reference default init (nn.Linear), random weights
        pol = DiffusionPolicy(S, A, T, device="cuda", hidden=hidden, precision=args.precision)
        cpu_params = {k: v.clone() for k, v in pol.state_dict().items()}
        pol.to(dev)
        state_h = torch.randn(B, S, generator=gen).pin_memory()
        state = state_h.to(dev)
        noise = torch.randn(T, B, A, generator=gen).to(dev)
        out_h = torch.empty(B, A).pin_memory()
        # kernel-only leg: the same C-ABI call get_actions() makes, with the buffers resolved once, so that no Python
        # work sits between the start event and the launch (it showed up as +0.15 ms per step under torchrun)
        from ddiffpg_b200._lib import lib, check, ptr, stream_ptr
        pol.get_actions(state[:256], noise=noise[:, :256].contiguous())     # builds the packed weights
        packed, shape, prec = pol._packed(args.precision, need=1)
        ws_bytes = lib().ddp_actor_sample_workspace_bytes(shape, B, prec)
        ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        out_d = torch.empty(B, A, device=dev)
        sample_args = (shape, ptr(packed), ptr(state), ptr(noise), ptr(out_d), B, prec, ptr(ws), ws_bytes)
        fn_sample = lib().ddp_actor_sample

        def step():
            check(fn_sample(*sample_args, stream_ptr()), "ddp_actor_sample")
        launches_per_step = 1

        def e2e_step():
            # the call a user with host-resident observations makes: pinned obs up, noise drawn on the device,
            # actions back down (H2D / D2H of row chunks overlap the sampler launches)
            pol.get_actions_host(state_h, out_h)
        h2d, d2h = state_h.numel() * 4, out_h.numel() * 4
    elif args.workload == "ascent":
        K = args.modes
        critics = []
        for m in range(K):
            torch.manual_seed(m)
            c = DistributionalDoubleQ(O, A, v_min=0, v_max=5, num_atoms=51, device="cuda")
            if m == 0:
                cpu_params = {k: v.clone() for k, v in c.state_dict().items()}
            critics.append(c.to(dev).requires_grad_(False))
        seg = [B * m // K for m in range(K + 1)]
        obs_h = torch.randn(B, O, generator=gen).pin_memory()
        act_h = (torch.rand(B, A, generator=gen) * 2 - 1).pin_memory()
        obs, act0 = obs_h.to(dev), act_h.to(dev)
        work = act0.clone()
        out_h = torch.empty(B, A).pin_memory()
        from ddiffpg_b200.models import _PackCache
        cache = _PackCache()

        def step():
            work.copy_(act0)
            q_action_ascent_segments(critics, obs, work, seg, iters=20, cache=cache, precision=args.precision)
        launches_per_step = 2 + 20 * 2 + 2

        def e2e_step():
            o = obs_h.to(dev, non_blocking=True)
            w = act_h.to(dev, non_blocking=True)
            q_action_ascent_segments(critics, o, w, seg, iters=20, cache=cache, precision=args.precision)
            out_h.copy_(w, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = (obs_h.numel() + act_h.numel()) * 4, out_h.numel() * 4
    else:
        torch.manual_seed(0)
        pol = DiffusionPolicy(S, A, T, device="cuda", hidden=hidden)
        cpu_params = {k: v.clone() for k, v in pol.state_dict().items()}
        pol.to(dev)
        trainer = FusedActorTrainer(pol, precision=args.precision, graph=not args.no_graph)
        st_h = torch.randn(B, S, generator=gen).pin_memory()
        ac_h = (torch.rand(B, A, generator=gen) * 2 - 1).pin_memory()
        st, ac = st_h.to(dev), ac_h.to(dev)
        nz = torch.randn(B, A, generator=gen).to(dev)
        ts = torch.randint(0, T, (B,), generator=gen).to(dev)
        loss_h = torch.empty(2).pin_memory()
        step = lambda: trainer.step(st, ac, noise=nz, timesteps=ts)
        # own kernels per step (profiles/r01/launches_train_bf16_v3.txt): re-pack 12, prep 1, fused forward 1, backward
        # row GEMMs 3, dW GEMMs 4, time branch 6, norm / clip / AdamW 3, + 2 of the fp32 transposes
        launches_per_step = 32

        def e2e_step():
            s = st_h.to(dev, non_blocking=True)
            a = ac_h.to(dev, non_blocking=True)
            loss, gn = trainer.step(s, a)
            loss_h[0:1].copy_(loss.reshape(1), non_blocking=True)
            loss_h[1:2].copy_(gn.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = (st_h.numel() + ac_h.numel()) * 4, 8

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, device_events=True):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed (untimed) before
        each; returns the summed device time in ms (max over ranks taken by the caller)."""
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms = timed(step, args.steps)
    # nvidia-smi polls every 100 ms and K steps can be shorter than one poll: keep the identical loop (flush + step)
    # running, untimed, until the sampler has seen >= 0.6 s of this load pattern, then stop it
    # (same step count on every rank: the training step contains a collective)
    ms_all = ms
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_all = tmax.item()
    per_step = max(ms_all / args.steps, 0.05) + 0.15          # + the untimed L2 flush
    timed(step, min(2000, max(0, int((600.0 - ms_all) / per_step))))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed steps + continuation of the same flush/step loop to 0.6 s (100 ms polls)"
    for _ in range(2):
        e2e_step()
    barrier()
    ms_e2e = timed(e2e_step, args.steps)
    barrier()
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        finish(world)
        return

    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    peaks = measured_peaks()
    flops_unit = algorithmic_flops(args.workload, T, h)
    achieved = B * flops_unit / (ms / args.steps * 1e-3) / 1e12            # per GPU, the step is the kernel sequence
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": None,
                "peak_source": f"{peaks['source']} bf16 burst (MEASURED_PEAKS.json)",
                "algorithmic_flops_per_unit": flops_unit,
                "hbm_bytes_per_unit": 4 * (S + T * A + A) if args.workload == "sample" else None}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            ent = json.load(f).get(f"{args.workload}_{args.precision}")
        # one ncu --set full capture of the dominant kernel at exactly this workload size (never measured here:
        # a number taken under a profiler is not a bench value, and ncu is not run inside the bench)
        if isinstance(ent, dict) and ent.get("rows") == B and ent.get("T") == T and ent.get("width") == h:
            roofline["traffic"] = ent["bytes"]
            roofline["traffic_source"] = ent["source"]

    # CPU baseline: the oracle port on this box's host cores, bounded sample (N=1 runs only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import port        # the checker's CPU restatement, timed as the baseline only
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        rows = 4096
        cg = torch.Generator().manual_seed(0)
        p = cpu_params
        if args.workload == "sample":
            cs, cn = torch.randn(rows, S, generator=cg), torch.randn(T, rows, A, generator=cg)
            cfn = lambda: port.actor_sample(p, cs, cn, T)
        elif args.workload == "ascent":
            rows = 2048
            co, ca = torch.randn(rows, O, generator=cg), torch.rand(rows, A, generator=cg) * 2 - 1
            cfn = lambda: port.q_action_ascent(p, co, ca.clone(), iters=20)
        else:
            c1, c2 = torch.randn(rows, S, generator=cg), torch.rand(rows, A, generator=cg) * 2 - 1
            c3, c4 = torch.randn(rows, A, generator=cg), torch.randint(0, T, (rows,), generator=cg)
            cfn = lambda: port.actor_loss_and_grads(p, c1, c2, c3, c4, T)
        cfn()
        n, t0 = 0, time.perf_counter()
        while True:
            cfn(); n += 1
            dt = time.perf_counter() - t0
            if dt > 10.0 or n >= 50:
                break
        cpu = {"value": rows * n / dt, "unit": unit, "cores": cores, "kind": "port",
               "sample": f"{n} passes of {rows} rows in {dt:.1f} s, torch CPU fp32, {cores} threads"}

    line = {"metric": metric_name(args.workload), "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, B), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu}
    emit(line)
    finish(world)


def finish(world):
    """End of a rank's work.  Under torchrun the process leaves without tearing NCCL down: the training workload replays
    CUDA graphs that captured the gradient all-reduce, and destroy_process_group() with such graphs alive was seen to
    block for good (2 GPUs, after the result line had been printed).  Every collective of the run has completed by now
    (the last one is the all-reduce of the timings), so nothing is lost."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="sample", choices=["sample", "ascent", "train"])
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"],
                    help="default: bf16 tensor-core paths (fp32 = warp-FMA parity paths)")
    ap.add_argument("--batch", type=int, default=65536, help="rows per GPU")
    ap.add_argument("--T", type=int, default=5)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--modes", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
