#!/usr/bin/env python
"""bench.py — UNet3d training-step throughput on B200 (BASELINE.json configs[1]).

Workload (cfg 2): UNet3d(1, 2, default_feature(2)) on one synthetic 160x192x160 T1-like volume,
batch 1 per GPU; one step = [visual_perception_augmentation of the sample on the GPU] -> one N=1 micro-batch
(forward, 5-level CE+Dice+MSE deep supervision, backward) -> [NCCL allreduce when N>1] -> grad/batch, clip 12,
Nesterov SGD.  `value` = steps/s with the sample already resident in HBM (CUDA events on the library's stream);
`e2e` = the same step through the public C-ABI with HOST (pinned) buffers, H2D/D2H inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
N>1 is launched by torchrun (one rank per GPU); rank 0 prints ONE JSON line.
--impl reference times the reference's own CPU implementation (oracle/_ref/unet_ref = /root/reference/unet.cpp
compiled unchanged against libtorch + restated step glue) on the host cores, on a bounded sample of the workload."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, D = 160, 192, 160
IN_C, OUT_C = 1, 2
METRIC = "train_steps_per_s"
UNIT = "steps/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "unet_ref")
CPU_SAMPLE_DIM = (96, 96, 96)   # bounded CPU sample: ~10-30 s of host work per step


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), which="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, which="fallback")


def synth_sample(seed):
    """Smooth ellipsoid 'head' + noise, max-normalised to [0,1] like tipl::normalize (train.cpp:30); label = mask."""
    import numpy as np
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(D, dtype=np.float32), np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    c = np.array([D, H, W]) * (0.5 + 0.03 * rng.uniform(-1, 1, 3))
    r = np.sqrt(((z - c[0]) / (0.42 * D)) ** 2 + ((y - c[1]) / (0.40 * H)) ** 2 + ((x - c[2]) / (0.38 * W)) ** 2)
    img = np.where(r < 1, 0.2 + 0.8 * np.clip(1 - r, 0, 1), 0).astype(np.float32)
    img += (rng.uniform(0, 0.05, r.shape).astype(np.float32)) * (r < 1)
    img /= img.max()
    lab = (r < 1).astype(np.float32)
    return img[None, None].copy(), lab[None].copy()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc, self.windows = gpu, [], None, []

    def start(self):
        """Started BEFORE the warm-up (nvidia-smi needs ~100 ms to deliver its first row); rows are time-stamped on arrival and
        only those inside a timed window (begin() .. end()) are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [t.strip() for t in line.split(",")]))

    def begin(self):
        self._t0 = time.perf_counter()

    def end(self):
        self.windows.append((self._t0, time.perf_counter()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, r in self.rows:
            if not any(b <= t <= e + 0.005 for b, e in self.windows):
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "rows_total": len(self.rows),
                "windows": "device-timed steps + end-to-end steps + inference windows"}


def workload_name(augment=True, simulate=False):
    return (f"cfg2: UNet3d({IN_C},{OUT_C},default_feature) single-template training step, human T1 skull-strip "
            f"{W}x{H}x{D}, batch 1 per GPU, {'with' if augment else 'WITHOUT'} "
            f"{'simulate_modality + ' if augment and simulate else ''}visual_perception_augmentation, "
            f"ce+dice+mse deep supervision, clip 12, Nesterov SGD")


def cpu_reference_step(steps, warmup, threads=None):
    """Times the reference (oracle/_ref) training step on the host on the bounded sample grid; returns dict."""
    if not os.path.exists(REF_BIN):
        return None
    threads = threads or os.cpu_count() or 1
    w, h, d = CPU_SAMPLE_DIM
    cmd = [REF_BIN, "time", "--mode", "step", "--in_c", str(IN_C), "--out_c", str(OUT_C), "--feature", "default", "--dim", str(w), str(h),
           str(d), "--steps", str(steps), "--warmup", str(warmup), "--threads", str(threads), "--batch", "1"]
    out = subprocess.check_output(cmd, text=True, timeout=1500)
    r = json.loads(out.strip().splitlines()[-1])
    scale = (W * H * D) / float(w * h * d)
    ms_full = r["ms_per_step"] * scale
    return {"value": 1000.0 / ms_full, "unit": UNIT, "cores": int(r["threads"]), "kind": "reference",
            "sample": f"oracle/_ref/unet_ref (reference unet.cpp + libtorch CPU): one full step (fwd+5-level loss+bwd+clip+SGD, no augmentation) at "
                      f"{w}x{h}x{d}, {r['ms_per_step']:.0f} ms, scaled x{scale:.2f} by voxel count to {W}x{H}x{D}",
            "ms_per_step_sample": r["ms_per_step"]}


def run_reference(args, rank):
    if rank != 0:
        return 0
    r = cpu_reference_step(max(1, min(args.steps, 2)), max(0, min(args.warmup, 1)))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/unet_ref not built (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / r["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(True, True), "grid": [W, H, D], "in_count": IN_C, "out_count": OUT_C,
                       "arm": "reference unet.cpp + libtorch CPU (oracle/_ref/unet_ref) on the host cores; simulate_modality and the augmentation "
                              "(TIPL) are not compilable here, so the CPU step excludes them -- they are in the GPU arm's timed region"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-augment", action="store_true")
    ap.add_argument("--no-simulate", action="store_true", help="skip simulate_modality (train.cpp:459) in front of the augmentation")
    ap.add_argument("--no-inference", action="store_true", help="skip the cfg1 inference leg (profiling runs under ncu)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)
    # libraries (NCCL's version banner, ...) may write to fd 1: keep stdout clean for the ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import numpy as np
    import torch
    from tests._pkg import load
    pkg = load()
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist
        import ctypes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = pkg.dist.bootstrap_nccl(pkg, dist, world, rank)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    net = pkg.UNet3d(IN_C, OUT_C, None, gpu=local_rank)
    net.init_params(0)          # identical on every rank (same seed) => replicas start in sync, no broadcast needed
    net.set_dim(W, H, D)
    net.train(True)
    lr0 = 1e-3
    net.create_optimizer(lr0)
    if comm is not None and not os.environ.get("U3D_NO_AR_OVERLAP"):
        net.attach_comm(comm, 1)     # one micro-batch per rank and step: overlap the gradient all-reduce with the backward pass
    img, lab = synth_sample(rank)                      # each rank trains on its own sample (data parallel)
    augment = not args.no_augment
    simulate = augment and not args.no_simulate
    if simulate:
        pkg.set_simulate_modality(net, 1)   # the template sample goes through the labelled overload first (train.cpp:459-460)
    x_host = torch.from_numpy(img).pin_memory()
    l_host = torch.from_numpy(lab).pin_memory()
    x_dev, l_dev = x_host.cuda(), l_host.cuda()
    total_steps = 1000
    step_no = [0]
    pending = [False]   # a prefetched sample is waiting in the handle

    def seed_of(s):
        return s * world + rank

    def one_step_device():
        # the sample is resident in HBM; its augmentation for step s+1 runs on the handle's prefetch stream while step s computes
        # (the reference augments in worker threads beside the trainer, train.cpp:446-485)
        s = step_no[0]
        if augment:
            if not pending[0]:
                pkg.prefetch_augmented(net, x_dev.data_ptr(), l_dev.data_ptr(), seed=seed_of(s), where=1)
            pkg.prefetch_augmented(net, x_dev.data_ptr(), l_dev.data_ptr(), seed=seed_of(s + 1), where=1)
            pending[0] = True
            loss = pkg.train_microbatch_prefetched(net)
        else:
            loss = net.device_train_microbatch(x_dev.data_ptr(), l_dev.data_ptr())
        net.step(world, pkg.poly_lr(lr0, s, total_steps), comm)
        step_no[0] += 1
        return loss

    def one_step_host():
        # end to end through the C-ABI with HOST buffers: the raw sample of step s+1 is uploaded from pinned memory and augmented on the
        # prefetch stream (train.cpp:446-485, 615-626) while step s trains; the losses come back to the host, then the update
        s = step_no[0]
        if augment:
            if not pending[0]:
                pkg.prefetch_augmented(net, x_host.numpy(), l_host.numpy(), seed=seed_of(s), where=0)
            pkg.prefetch_augmented(net, x_host.numpy(), l_host.numpy(), seed=seed_of(s + 1), where=0)
            pending[0] = True
            loss = pkg.train_microbatch_prefetched(net)
        else:
            loss = net.train_microbatch(x_host.numpy(), l_host.numpy())
        net.step(world, pkg.poly_lr(lr0, s, total_steps), comm)
        step_no[0] += 1
        return loss

    # ---------------- device-resident timing ----------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        one_step_device()
    launches0 = net.launch_count()
    barrier()
    clocks.begin()
    net.timer_start()
    loss = None
    for _ in range(args.steps):
        loss = one_step_device()
    ms = net.timer_stop()
    barrier()
    clocks.end()
    launches = net.launch_count() - launches0
    # ---------------- per-family attribution: the same steps again with a CUDA-event pair around every tensor-kernel launch.  The
    # library serialises the weight-gradient side stream while profiling, so these per-kernel times are not the concurrent ones ----
    net.profile(True)
    net.profile_read(reset=True)
    for _ in range(args.steps):
        one_step_device()
    prof = net.profile_read(reset=True)
    net.profile(False)
    # ---------------- end-to-end timing through the C-ABI with host buffers ----------------
    for _ in range(2):
        one_step_host()
    barrier()
    clocks.begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks.end()
    barrier()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_s * 1000.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1000.0
    # ---------------- inference (BASELINE configs[0]/[4] building block): forward()[0] of one 160x192x160 window per GPU ----------------
    train_loss_scale = net.loss_scale()
    del net
    inf_ms = inf_e2e_ms = float('nan')
    n_inf = 1
    if not args.no_inference:
        inf = pkg.UNet3d(IN_C, 1, None, gpu=local_rank)       # skull-strip 1 in / 1 out (cfg 1)
        inf.init_params(0)
        inf.set_dim(W, H, D)
        inf.prepare_for_inference()
        y_dev = torch.empty(1, 1, D, H, W, device="cuda")
        y_host = torch.empty(1, 1, D, H, W).pin_memory()
        for _ in range(3):
            inf.device_forward(x_dev.data_ptr(), [y_dev.data_ptr()])
        inf.sync()
        n_inf = max(args.steps, 5)
        barrier()
        clocks.begin()
        inf.timer_start()
        for _ in range(n_inf):
            inf.device_forward(x_dev.data_ptr(), [y_dev.data_ptr()])
        inf_ms = inf.timer_stop()
        clocks.end()
        barrier()
        y_host2 = torch.empty(1, 1, D, H, W).pin_memory()
        outs = [(y_host if i % 2 == 0 else y_host2).numpy() for i in range(n_inf)]
        inf.evaluate_windows([x_host.numpy()] * 2, outs[:2])     # warm the staging slots
        t0 = time.perf_counter()
        # evaluate.cpp:223-230 over n_inf windows from / to pinned host buffers: H2D, forward()[0], D2H per window, pipelined by the library
        inf.evaluate_windows([x_host.numpy()] * n_inf, outs)
        inf_e2e_ms = (time.perf_counter() - t0) * 1000.0
        barrier()
    clk = clocks.stop()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([inf_ms, inf_e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        inf_ms, inf_e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        import torch.distributed as dist
        dist.destroy_process_group()
        return 0
    pk = peaks()
    value = world * args.steps / (ms / 1000.0)
    e2e_value = world * args.steps / (e2e_ms / 1000.0)
    vox = W * H * D
    fams = {}
    for i, name in enumerate(("conv_igemm_kernel", "conv_wgrad_kernel", "conv_halo_s2_kernel", "conv_wgrad_band_kernel", "conv_tma_kernel", "conv_band_kernel")):
        ms_k, n_k, fl_k = prof[3 * i:3 * i + 3]
        fams[name] = {"ms_per_step": ms_k / args.steps, "launches_per_step": n_k / args.steps, "gflop_per_step": fl_k / args.steps / 1e9,
                      "achieved_tflops": (fl_k / 1e12) / (ms_k / 1e3) if ms_k > 0 else None}
    dom = max(fams, key=lambda k: fams[k]["ms_per_step"])
    achieved = fams[dom]["achieved_tflops"]
    ncu_summary = {}
    sp = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(sp):
        ncu_summary = json.load(open(sp))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": workload_name(augment, simulate),
                   "grid": [W, H, D], "in_count": IN_C, "out_count": OUT_C, "micro_batches_per_gpu_per_step": 1,
                   "global_batch": world, "parallelism": f"dp{world}", "augmentation": bool(augment), "simulate_modality": bool(simulate),
                   "l2": "no explicit flush: each step streams > 3 GB of activations, far above the 126 MB L2",
                   "arithmetic": "fp16 operands, fp32 accumulate (tcgen05), fp32 stats/loss/optimizer, loss scale %g" % train_loss_scale},
        "loss": [float(v) for v in loss],
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (IN_C + 1) * vox * 4,
                "d2h_bytes_per_step": 15 * 4 + 16,
                "api": "unet3d_prefetch_augmented(next host sample) + unet3d_train_microbatch_prefetched (losses out) + unet3d_step"},
        "gpu_launches": int(launches),
        "inference": {"workload": f"cfg1: UNet3d({IN_C},1,default) forward()[0] of one {W}x{H}x{D} window per GPU (windows sharded, no collective)",
                      "value": world * n_inf * (W * H * D) / 1e6 / (inf_ms / 1e3), "unit": "Mvoxel/s",
                      "ms_per_window": inf_ms / n_inf,
                      "e2e": {"value": world * n_inf * (W * H * D) / 1e6 / (inf_e2e_ms / 1e3), "unit": "Mvoxel/s", "api": "unet3d_evaluate_windows (host buffers)",
                              "h2d_bytes_per_window": IN_C * W * H * D * 4, "d2h_bytes_per_window": W * H * D * 4}},
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": dom + " (the tensor-core kernel family with the largest share of the step)",
                     "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": (achieved / pk["tf_sustained"]) if achieved else None, "peak_source": pk["which"] + " bf16/fp16 sustained",
                     "traffic": ncu_summary.get(dom + "_dram_bytes_per_launch"),
                     "how": "sum of algorithmic FLOPs (2*Cin*Cout*k^3*V_out per layer) / sum of CUDA-event durations of that family's launches inside the timed steps",
                     "families": fams},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_step(1, 0)
            if cb:
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is reported, never fatal
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
