#!/usr/bin/env python
"""Headline benchmark: VMM EM objects/sec, fwd+bwd (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32|bf16_fast]
    python bench.py --impl reference ...       # the reference's CPU path (oracle port) on the host cores

One "step" = one training pass over a batch of synthetic objects: 12 views x 4096-d features,
K = 3 latent views, 10 EM iterations -> head -> cross-entropy + NEM loss -> hand-written backward
-> (N > 1: one NCCL all-reduce of the flat gradient bucket) -> Adam.  Weak scaling: every rank
holds `--batch` objects (default 4096, configs[1]); `value` is the whole-job objects/s.

Timing: CUDA events on the launching stream around exactly K steps, bracketed by a barrier and
torch.cuda.synchronize(), max over ranks.  The step's working set (tens of GB of activations)
is far larger than the 126 MB L2, so no explicit L2 flush is needed ("l2": "inputs>L2").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_LATENT, M_VIEWS, D_FEAT, H_HID, T_ITERS, N_CLASSES = 3, 12, 4096, 1024, 10, 40
# Reference-algorithm GEMM FLOPs per object, fwd+bwd incl. head, counted with torch's flop counter on the
# reference's own code (SURVEY.md 8d / BASELINE.md 2).
ALG_FLOPS_PER_OBJECT = 39.681e9


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def build_models(precision, device):
    import torch
    from multi_view_sort_b200 import NEM, Multi_View_Net
    torch.manual_seed(10)
    nem = NEM(nclasses=N_CLASSES, view_num=M_VIEWS, k=K_LATENT, input_dim=D_FEAT, rnn_hidden_size=H_HID, iter_num=T_ITERS,
              init_type="bin_uniform", gamma_prior_loss=1, e_sigma=0.1, precision=precision)
    head = Multi_View_Net(None, "vggm", nclasses=N_CLASSES, k=K_LATENT, rnn_hidden_size=H_HID, precision=precision)
    return nem.to(device), head.to(device)


def cpu_oracle_rate(batch=8, reps=3, fwd_bwd=True):
    """The reference's CPU path (oracle port of train_nem.py / Trainer.py:160-219) on the host cores."""
    import torch
    from oracle import nem_oracle as O
    from oracle.make_golden import seeded_models
    cfg = dict(nclasses=N_CLASSES, M=M_VIEWS, K=K_LATENT, D=D_FEAT, H=H_HID, T=T_ITERS, stop_gamma_grad=0,
               if_gamma_prior=1, sigma=0.1, if_sort=1, if_pooling=0)
    torch.set_num_threads(os.cpu_count() or 1)
    nem, head = seeded_models(cfg)
    pn = {k: v for k, v in nem.named_parameters()}
    ph = {k: v for k, v in head.named_parameters()}
    x, y = O.make_inputs(batch, M_VIEWS, D_FEAT, N_CLASSES, seed=10)
    times = []
    for i in range(reps + 1):
        for p in list(pn.values()) + list(ph.values()):
            p.grad = None
        t0 = time.perf_counter()
        r = O.train_step(pn, ph, x, y, K_LATENT, T_ITERS, 0.1, 1, 0, 1, 0)
        if fwd_bwd:
            r["loss"].backward()
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
    best = min(times)
    return dict(value=batch / best, unit="objects/s", cores=torch.get_num_threads(), kind="port",
                sample=f"oracle (torch CPU restatement of the reference, bit-identical to it) fwd+bwd, batch {batch}, "
                       f"best of {reps} after 1 warm-up; CPU rate is flat in batch (SURVEY.md 6)"), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup, batch = args.steps, args.warmup, 8
    import torch
    from oracle import nem_oracle as O
    from oracle.make_golden import seeded_models
    cfg = dict(nclasses=N_CLASSES, M=M_VIEWS, K=K_LATENT, D=D_FEAT, H=H_HID, T=T_ITERS, stop_gamma_grad=0,
               if_gamma_prior=1, sigma=0.1, if_sort=1, if_pooling=0)
    torch.set_num_threads(os.cpu_count() or 1)
    nem, head = seeded_models(cfg)
    params = list(nem.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=5e-5, weight_decay=1e-4)
    pn = {k: v for k, v in nem.named_parameters()}
    ph = {k: v for k, v in head.named_parameters()}
    x, y = O.make_inputs(batch, M_VIEWS, D_FEAT, N_CLASSES, seed=10)

    def step():
        opt.zero_grad(set_to_none=True)
        r = O.train_step(pn, ph, x, y, K_LATENT, T_ITERS, 0.1, 1, 0, 1, 0)
        r["loss"].backward()
        opt.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = batch * steps / dt
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": "VMM EM objects/sec fwd+bwd (12x4096, K=3, 10 iters)", "value": val,
        "unit": "objects/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] shape (12 views x 4096-d, K=3, T=10, head + CE + Adam); bounded sample of 8 "
                               "objects per step on the host cores (the reference cannot hold batch 4096: 511 GB of autograd state)",
                   "batch_per_step": batch},
        "cpu_baseline": {"value": val, "unit": "objects/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {batch} objects, oracle port of the reference (the reference tree does not "
                                   "travel to the GPU box; the port is bit-identical to it on CPU)"},
        "e2e": {"value": val, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", help="bf16 (bf16 features, 1e-2 parity), fp32 (1e-4 parity), bf16_fast")
    ap.add_argument("--batch", type=int, default=4096, help="objects per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from multi_view_sort_b200._lib import lib
    from multi_view_sort_b200.dist import GradBucket
    from multi_view_sort_b200.functional import cross_entropy

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU path)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    xdtype = torch.float32 if args.precision.startswith("fp32") else torch.bfloat16
    nem, head = build_models(args.precision, dev)
    params = list(nem.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=5e-5, weight_decay=1e-4, fused=True)
    bucket = GradBucket(params)

    # synthetic features, generated once on the host (pinned) with a per-rank seed
    g = torch.Generator().manual_seed(10 + rank)
    x_host = torch.relu(torch.randn(B, M_VIEWS, D_FEAT, generator=g)).to(xdtype).pin_memory()
    y_host = torch.randint(0, N_CLASSES, (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    loss_host = torch.zeros(1).pin_memory()

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        h, gamma, nem_loss, _, _ = nem.run_em(x, T_ITERS)
        logits = head(h, gamma)
        loss = cross_entropy(logits, y) + nem_loss
        loss.backward()
        if world > 1:
            bucket.allreduce_mean()
        if not args.no_optimizer:
            opt.step()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return ms.item()

    L = lib()
    L.vmm_launch_count.restype = C.c_longlong
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)

    # ---- device-resident run (the `value`) --------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.vmm_launch_count()
    ms = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = L.vmm_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers (the `e2e`) ----------------------------
    # The timed region holds K host->device copies of the step's inputs (pinned memory) and K device->host reads of
    # the loss.  Inputs go through the package's DevicePrefetcher, as in ModelNetTrainer.train_nem_mvcnn: batch i+1
    # is copied on a side stream while step i computes; batch 0 of the region is not overlapped with anything.
    from multi_view_sort_b200.features import DevicePrefetcher

    feeder = DevicePrefetcher([], dev)                      # its two device buffer sets are allocated by the warm-up run

    def e2e_run(steps):
        feeder.loader = ((y_host, x_host) for _ in range(steps))
        for y, x in feeder:
            loss = step(x, y)
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()       # the caller reads the loss every step

    def timed_once(fn):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sync_all()
        return t.item()

    e2e_run(3)
    ms_e2e = timed_once(lambda: e2e_run(args.steps))
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- dominant kernel (tcgen05 contraction): summed launch time over one extra pass of K steps -------
    L.vmm_gemm_profile(1)
    timed(lambda: step(x_dev, y_dev), args.steps)
    tot_ms, ex_fl, nl = C.c_double(), C.c_double(), C.c_longlong()
    L.vmm_gemm_profile_read(C.byref(tot_ms), C.byref(ex_fl), C.byref(nl))
    L.vmm_gemm_profile(0)
    pk = peaks()
    traffic = None      # dram bytes of all contraction launches of one step, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            tr = json.load(f)
        if tr.get("precision") == args.precision and tr.get("batch") == B:
            traffic = tr["dram_bytes_per_step"]
    except Exception:
        pass
    gemm_ms_per_step = tot_ms.value / args.steps
    alg_flops_step = ALG_FLOPS_PER_OBJECT * B
    achieved = alg_flops_step / (gemm_ms_per_step * 1e-3) / 1e12
    executed = ex_fl.value / (tot_ms.value * 1e-3) / 1e12 if tot_ms.value > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        "metric": "VMM EM objects/sec fwd+bwd (12x4096, K=3, 10 iters)", "value": value, "unit": "objects/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision.startswith("bf16") else "f32 (split-bf16 tensor-core passes)",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 12 views x 4096-d features, K=3, T=10, batch 4096 objects per GPU, "
                               "training step = EM fwd + head + CE + NEM loss + backward"
                               + (" + NCCL grad all-reduce" if world > 1 else "") + ("" if args.no_optimizer else " + Adam"),
                   "precision": args.precision, "features": str(xdtype).replace("torch.", ""), "batch_per_gpu": B,
                   "global_batch": B * world, "l2": "inputs>L2 (tens of GB of activations per step)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "objects/s", "h2d_bytes_per_step": x_host.numel() * x_host.element_size()
                + y_host.numel() * 8, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_sustained"], "traffic": traffic,
                     "kernel": "gemm_tcgen05_kernel (all contraction launches of the step)",
                     "launches_per_step": nl.value / args.steps, "kernel_ms_per_step": gemm_ms_per_step,
                     "kernel_share_of_step": gemm_ms_per_step / (ms / args.steps),
                     "executed_tflops": executed, "executed_frac": executed / pk["bf16_sustained"],
                     "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
                     "note": "achieved = reference-algorithm FLOPs (39.681 GFLOP/object) / summed launch time; executed counts "
                             "every split-bf16 pass the precision mode issues",
                     "step_frac": (alg_flops_step / (ms / args.steps * 1e-3) / 1e12) / pk["bf16_sustained"],
                     "hbm_frac": ((110592 if xdtype == torch.bfloat16 else 208896) * B / (ms / args.steps * 1e-3) / 1e9)
                     / pk["hbm_gbs"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_oracle_rate()
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
