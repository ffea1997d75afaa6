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
# one "step" = one batch-1 training step (one micro-batch: forward, 5-level loss, backward, + the update).  At N GPUs every rank runs
# one of them per optimizer update (weak scaling), so value = N x optimizer updates/s; at N = 1 the two are the same thing.
UNIT = "micro-batch steps/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "unet_ref")
CPU_BUDGET_S = 150.0            # wall-time bound of the --impl reference run (the reference's CPU path at the REAL grid)


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


def run_ref_bin(mode, out_c, dims, steps, warmup, batch=1, device="cpu", threads=None, budget_s=None, timeout=1700):
    """oracle/_ref/unet_ref time ... -> its JSON line (None when the binary is not there)."""
    if not os.path.exists(REF_BIN):
        return None
    threads = threads or os.cpu_count() or 1
    cmd = [REF_BIN, "time", "--mode", mode, "--in_c", str(IN_C), "--out_c", str(out_c), "--feature", "default", "--dim", *[str(v) for v in dims],
           "--steps", str(steps), "--warmup", str(warmup), "--threads", str(threads), "--batch", str(batch), "--device", device]
    if budget_s:
        cmd += ["--budget_s", str(budget_s)]
    out = subprocess.check_output(cmd, text=True, timeout=timeout, stderr=subprocess.DEVNULL)
    return json.loads(out.strip().splitlines()[-1])


def cpu_reference_step(steps, warmup, batch=1, budget_s=None):
    """The reference training step (unet.cpp unchanged + calc_losses + the restated update) on the host cores at the REAL cfg-2 grid.
    `steps` is an upper bound; the binary stops taking timed steps when budget_s is spent and reports how many it ran."""
    r = run_ref_bin("step", OUT_C, (W, H, D), steps, warmup, batch=batch, budget_s=budget_s)
    if r is None:
        return None
    ms = r["ms_per_step"]
    return {"value": batch * 1000.0 / ms, "unit": UNIT, "cores": int(r["threads"]), "kind": "reference",
            "sample": f"oracle/_ref/unet_ref = /root/reference/unet.cpp compiled unchanged + the reference's calc_losses, libtorch {r.get('torch', '?')} CPU "
                      f"(fp32, oneDNN), {r['threads']} threads: {r['steps']} timed optimizer step(s) of {batch} micro-batch(es) after {warmup} warm-up at the "
                      f"real {W}x{H}x{D} grid, {ms:.0f} ms per optimizer step (fwd + 5-level loss + bwd + clip + SGD).  simulate_modality and "
                      f"visual_perception_augmentation are NOT in this figure: they need TIPL (un-vendored) and do not compile here; in the "
                      f"reference they run in worker threads beside the trainer (train.cpp:446-485)",
            "steps_run": int(r["steps"]), "ms_per_optimizer_step": ms, "micro_batches_per_step": batch}


def cpu_augmentation_port_seconds():
    """The oracle's numpy port of simulate_modality + visual_perception_augmentation on ONE host core at the real grid (the stage the CPU
    reference figure leaves out).  Reported next to the baseline, never added to it."""
    try:
        from oracle import vpa_oracle as V, simulate_oracle as S
        img, lab = synth_sample(0)
        t0 = time.perf_counter()
        im = S.simulate_modality(img[0, 0].copy(), lab[0].copy(), OUT_C, 0)
        t1 = time.perf_counter()
        V.augment(dict(V.OPTION_DEFAULTS), im[None].copy(), lab[0].copy(), True, (W, H, D), 0)
        t2 = time.perf_counter()
        return {"simulate_modality_s": t1 - t0, "visual_perception_augmentation_s": t2 - t1, "kind": "port", "cores": 1,
                "what": "oracle/simulate_oracle.py + oracle/vpa_oracle.py (numpy restatements, parity unpinned), one sample at the real grid"}
    except Exception as ex:
        return {"failed": str(ex)}


def gpu_baseline(steps):
    """The reference on libtorch CUDA + cuDNN on THIS B200 ('the bar'): the same unet_ref binary with --device cuda,
    libtorch defaults (cuDNN may use TF32 for convolutions).  cfg-2 training step and cfg-1 inference window."""
    out = {}
    try:
        r = run_ref_bin("step", OUT_C, (W, H, D), max(3, min(steps, 20)), 3, device="cuda", timeout=600)
        out["train"] = {"value": 1000.0 / r["ms_per_step"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "ms_median": r["ms_median"],
                        "steps_run": r["steps"], "cudnn_tf32": bool(r["cudnn_tf32"]), "torch": r.get("torch"),
                        "what": "unet_ref time --mode step --device cuda: H2D of the sample, fwd, 5-level calc_losses, bwd, losses D2H, /batch, clip, SGD "
                                "(train.cpp:615-706,755-766); no augmentation"}
    except Exception as ex:
        out["train"] = {"failed": str(ex)[-300:]}
    try:
        r = run_ref_bin("fwd", 1, (W, H, D), max(3, min(steps, 20)), 3, device="cuda", timeout=600)
        vox = W * H * D / 1e6
        out["inference"] = {"value": vox / (r["ms_resident"] / 1e3), "unit": "Mvoxel/s", "ms_per_window_resident": r["ms_resident"],
                            "e2e": {"value": vox / (r["ms_per_step"] / 1e3), "unit": "Mvoxel/s", "ms_per_window": r["ms_per_step"],
                                    "what": "evaluate.cpp:226-229 per window: H2D, forward()[0], D2H from pageable host memory"},
                            "cudnn_tf32": bool(r["cudnn_tf32"])}
    except Exception as ex:
        out["inference"] = {"failed": str(ex)[-300:]}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    # N GPUs run N micro-batches per optimizer update (weak scaling); the reference's CPU path runs them one after the other
    est_s = 5.0 * world
    steps = max(1, min(args.steps, int(CPU_BUDGET_S / est_s)))
    r = cpu_reference_step(steps, 1, batch=world, budget_s=CPU_BUDGET_S)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/unet_ref not built (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps_run"],
            "steps_requested": args.steps, "warmup": 1, "ms_per_step": r["ms_per_optimizer_step"] / world,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(True, True), "grid": [W, H, D], "in_count": IN_C, "out_count": OUT_C,
                       "micro_batches_per_optimizer_step": world,
                       "arm": "reference unet.cpp + libtorch CPU (oracle/_ref/unet_ref) on the host cores at the real grid, one process, all host "
                              "threads; 'steps' is the number of optimizer steps actually timed inside the wall-time budget",
                       "excluded_stages": "simulate_modality + visual_perception_augmentation (TIPL, not compilable here; worker threads in the reference)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def extra_configs(pkg, torch, comm, world, rank, local_rank, barrier, clocks):
    """BASELINE.json configs 3, 4, 5 as extra keys (short runs; the headline stays cfg 2).  Every rank takes part."""
    import numpy as np

    def maxr(v):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return v

    def train_cfg(out_c, dims, mb_per_gpu, steps, flop_per_mb):
        w, h, d = dims
        net = pkg.UNet3d(IN_C, out_c, None, gpu=local_rank)
        net.init_params(0)
        net.set_dim(w, h, d)
        net.train(True)
        net.create_optimizer(1e-3)
        if comm is not None:
            net.attach_comm(comm, mb_per_gpu)
        pkg.set_simulate_modality(net, 1)
        rng = np.random.default_rng(100 + rank)
        z, y, x = np.meshgrid(np.arange(d, dtype=np.float32), np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
        r = np.sqrt(((z - d / 2) / (0.42 * d)) ** 2 + ((y - h / 2) / (0.40 * h)) ** 2 + ((x - w / 2) / (0.38 * w)) ** 2)
        img = (np.clip(1.1 - r, 0, 1) + 0.05 * rng.random(r.shape, dtype=np.float32)).astype(np.float32)
        img /= img.max()
        lab = np.zeros_like(r)
        for k in range(1, out_c):
            lab += (r < 1.0 - (k - 1) * (0.9 / max(out_c - 1, 1)))
        xh = torch.from_numpy(img[None, None].copy()).pin_memory()
        lh = torch.from_numpy(lab[None].astype(np.float32)).pin_memory()
        B = mb_per_gpu * world
        seq = [0]

        def one_step():
            # host-buffer path: sample b+1 is uploaded + simulated + augmented on the prefetch stream while micro-batch b trains
            for b in range(mb_per_gpu):
                if seq[0] == 0:
                    pkg.prefetch_augmented(net, xh.numpy(), lh.numpy(), seed=rank, where=0)
                seq[0] += 1
                pkg.prefetch_augmented(net, xh.numpy(), lh.numpy(), seed=seq[0] * world + rank, where=0)
                loss = pkg.train_microbatch_prefetched(net)
            net.step(B, 1e-3, comm)
            return loss

        one_step()
        barrier()
        clocks.begin()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = one_step()
        torch.cuda.synchronize()
        dt = maxr(time.perf_counter() - t0)
        clocks.end()
        barrier()
        skipped = net.last_step_skipped()
        del net
        return {"optimizer_steps_per_s": steps / dt, "micro_batches_per_s": steps * B / dt, "ms_per_optimizer_step": dt / steps * 1e3,
                "micro_batches_per_gpu_per_step": mb_per_gpu, "global_batch": B, "grid": [w, h, d], "out_count": out_c, "steps_timed": steps,
                "tflops_per_gpu": flop_per_mb * mb_per_gpu * steps / dt / 1e12, "loss": [float(v) for v in loss], "last_step_skipped": bool(skipped),
                "api": "host buffers: unet3d_prefetch_augmented + unet3d_train_microbatch_prefetched per micro-batch, unet3d_step per update"}

    out = {}
    try:
        out["cfg3"] = dict(train_cfg(6, (W, H, D), 8, 3, 1.72e12), workload="cfg3: UNet3d(1,6) multi-class tissue, 160x192x160, 8 micro-batches per GPU "
                           "per optimizer step (gradient accumulation), fp16 operands")
        out["cfg4"] = dict(train_cfg(2, (128, 160, 96), 8, 5, 3 * 229.51e9), workload="cfg4: UNet3d(1,2) rodent grid 128x160x96, 8 micro-batches per GPU per "
                           "optimizer step, one NCCL all-reduce per step when N > 1")
        # cfg5: the 8 windows (160x192x160 at stride 160,128,160) of one 320^3 volume, window i -> rank i % world, no collective
        net = pkg.UNet3d(IN_C, 6, None, gpu=local_rank)
        net.init_params(0)
        net.set_dim(W, H, D)
        net.prepare_for_inference()
        rng = np.random.default_rng(5)
        mine = pkg.dist.shard_windows(8, world, rank)
        wins = [torch.from_numpy(rng.random((1, IN_C, D, H, W), dtype=np.float32)).pin_memory() for _ in mine]
        outs = [torch.empty(1, 6, D, H, W).pin_memory() for _ in range(min(2, len(mine)))]
        ob = [outs[i % len(outs)].numpy() for i in range(len(mine))]
        net.evaluate_windows([w_.numpy() for w_ in wins], ob)
        barrier()
        clocks.begin()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            net.evaluate_windows([w_.numpy() for w_ in wins], ob)
        dt = maxr(time.perf_counter() - t0)
        clocks.end()
        barrier()
        out["cfg5"] = {"workload": "cfg5: UNet3d(1,6) 0.5 mm 320^3 volume as 8 windows of 160x192x160 sharded over the GPUs (window i -> rank i % N), "
                                   "host buffers in, fp32 logits of 6 classes out, no collective",
                       "value": reps * 8 * W * H * D / 1e6 / dt, "unit": "Mvoxel/s", "ms_per_volume": dt / reps * 1e3, "windows_per_gpu": len(mine),
                       "h2d_bytes_per_window": IN_C * W * H * D * 4, "d2h_bytes_per_window": 6 * W * H * D * 4,
                       "api": "unet3d_evaluate_windows (host buffers)"}
        if world == 1:
            # the same volume through unet3d_evaluate_volume: windows cut, forwarded, soft-maxed, re-assembled and arg-maxed on the GPU;
            # 1 byte per voxel comes back instead of 6 x 4
            vol = torch.from_numpy(rng.random((IN_C, 320, 320, 320), dtype=np.float32)).pin_memory()
            pkg.evaluate_volume(net, vol.numpy(), (160, 128, 160), 0.5, want_fg=False)
            t0 = time.perf_counter()
            for _ in range(reps):
                _, _, _, nwin = pkg.evaluate_volume(net, vol.numpy(), (160, 128, 160), 0.5, want_fg=False)
            dtl = time.perf_counter() - t0
            out["cfg5"]["label_map"] = {"value": reps * 320 ** 3 / 1e6 / dtl, "unit": "Mvoxel/s", "ms_per_volume": dtl / reps * 1e3, "windows": nwin,
                                        "h2d_bytes_per_volume": IN_C * 320 ** 3 * 4, "d2h_bytes_per_volume": 320 ** 3,
                                        "api": "unet3d_evaluate_volume (host volume in, uint8 label map out; softmax + create_mask + argmax on the GPU)"}
            del vol
            S = 320
            net.set_dim(S, S, S)
            xv = torch.rand(1, IN_C, S, S, S, device="cuda")
            yv = torch.empty(1, 6, S, S, S, device="cuda")
            net.device_forward(xv.data_ptr(), [yv.data_ptr()])
            net.sync()
            net.timer_start()
            for _ in range(3):
                net.device_forward(xv.data_ptr(), [yv.data_ptr()])
            ms1 = net.timer_stop() / 3
            out["cfg5"]["single_pass_320"] = {"ms": ms1, "value": S ** 3 / 1e6 / (ms1 / 1e3), "unit": "Mvoxel/s", "tflops": 3830.70e9 / (ms1 / 1e3) / 1e12,
                                             "what": "the whole 320^3 volume in one forward, device-resident"}
            del xv, yv
        del net
    except Exception as ex:   # extra keys never take the headline down
        out["extra_configs_error"] = repr(ex)[-400:]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-augment", action="store_true")
    ap.add_argument("--no-simulate", action="store_true", help="skip simulate_modality (train.cpp:459) in front of the augmentation")
    ap.add_argument("--no-inference", action="store_true", help="skip the cfg1 inference leg (profiling runs under ncu)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the libtorch/cuDNN baseline on this GPU")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the cfg 3 / 4 / 5 extra keys")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
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
    total_steps = 100000
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
    # (>= 5: the library captures a micro-batch as a CUDA graph at its SECOND occurrence per sample buffer, and the sample buffers
    # alternate between two slots -- the captures fall on steps 3 and 4 and must not be timed)
    n_warm = max(args.warmup, 5)
    for _ in range(n_warm):
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
    # ---------------- the same loop over a longer window (the contract times exactly --steps; a 20-step window is 0.15 s) ----------------
    n_sus = max(200, args.steps)
    barrier()
    clocks.begin()
    net.timer_start()
    for _ in range(n_sus):
        one_step_device()
    ms_sus = net.timer_stop()
    barrier()
    clocks.end()
    # ---------------- N > 1: the same steps WITHOUT the gradient all-reduce (replicas may drift apart: timing only) = what the
    # collective costs on top of the compute, i.e. its exposed (non-overlapped) time plus the SM share it takes from the backward pass
    ms_noar = None
    if world > 1:
        net.attach_comm(None, 1)
        real_comm, comm = comm, None
        for _ in range(5):   # (new graph variants without the communicator: captured on steps 3 and 4)
            one_step_device()
        barrier()
        net.timer_start()
        for _ in range(args.steps):
            one_step_device()
        ms_noar = net.timer_stop()
        barrier()
        comm = real_comm
        if not os.environ.get("U3D_NO_AR_OVERLAP"):
            net.attach_comm(comm, 1)
    # ---------------- per-family attribution: the same steps again with a CUDA-event pair around every tensor-kernel launch.  The
    # library serialises the weight-gradient side stream while profiling, so these per-kernel times are not the concurrent ones ----
    net.profile(True)
    net.profile_read(reset=True)
    for _ in range(args.steps):
        one_step_device()
    prof = net.profile_read(reset=True)
    net.profile(False)
    # ---------------- end-to-end timing through the C-ABI with host buffers ----------------
    for _ in range(5):   # the host path runs through the two prefetch slots: its graphs are captured on steps 3 and 4
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
        t = torch.tensor([ms, e2e_s * 1000.0, ms_sus, ms_noar], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, ms_sus, ms_noar = float(t[0]), float(t[1]), float(t[2]), float(t[3])
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
    extras = {}
    if not args.no_extra_configs:
        extras = extra_configs(pkg, torch, comm, world, rank, local_rank, barrier, clocks)
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
    for i, name in enumerate(("conv_igemm_kernel", "conv_wgrad_kernel", "conv_s2_kernel", "conv_wgrad_band_kernel", "conv_tma_kernel", "conv_band_kernel",
                              "conv_wgrad_quad_kernel")):
        ms_k, n_k, fl_k = prof[3 * i:3 * i + 3]
        fams[name] = {"ms_per_step": ms_k / args.steps, "launches_per_step": n_k / args.steps, "gflop_per_step": fl_k / args.steps / 1e9,
                      "achieved_tflops": (fl_k / 1e12) / (ms_k / 1e3) if ms_k > 0 else None}
    for f in fams.values():
        f["frac"] = (f["achieved_tflops"] / pk["tf_sustained"]) if f["achieved_tflops"] else None
    live = {k: f for k, f in fams.items() if f["ms_per_step"] > 0}
    dom = max(live, key=lambda k: live[k]["ms_per_step"])
    low = min(live, key=lambda k: live[k]["achieved_tflops"])
    achieved = fams[dom]["achieved_tflops"]
    tensor_gflop = sum(f["gflop_per_step"] for f in fams.values())
    ncu_summary = {}
    sp = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(sp):
        ncu_summary = json.load(open(sp))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "optimizer_updates_per_s": args.steps / (ms / 1000.0),
        "sustained": {"steps": n_sus, "ms_per_step": ms_sus / n_sus, "value": world * n_sus / (ms_sus / 1000.0), "unit": UNIT,
                      "what": "the same device-timed loop over a longer window, right after the contract's --steps window"},
        "config": {"workload": workload_name(augment, simulate),
                   "step_definition": "one batch-1 training step (micro-batch + update); N GPUs run N of them per optimizer update (weak scaling), "
                                      "so value = N x optimizer_updates_per_s",
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
                     "lowest": {"kernel": low, "achieved": fams[low]["achieved_tflops"], "frac": fams[low]["frac"],
                                "ms_per_step": fams[low]["ms_per_step"]},
                     "whole_step": {"tensor_gflop_per_step": tensor_gflop, "achieved": tensor_gflop / 1e3 / (ms / args.steps / 1e3),
                                    "frac": tensor_gflop / 1e3 / (ms / args.steps / 1e3) / pk["tf_sustained"],
                                    "what": "algorithmic conv FLOPs of one micro-batch / the device-timed step (everything included)"},
                     "families": fams},
    }
    line["inference"]["roofline"] = {"gflop_per_window": 573.56, "achieved": 573.56 / 1e3 / (inf_ms / n_inf / 1e3), "unit": "TFLOP/s",
                                     "frac": 573.56 / 1e3 / (inf_ms / n_inf / 1e3) / pk["tf_sustained"]}
    if ms_noar is not None:
        line["allreduce"] = {"ms_per_step_with": ms / args.steps, "ms_per_step_without": ms_noar / args.steps,
                             "cost_ms_per_step": (ms - ms_noar) / args.steps, "bytes_per_step": 15023818 * 4,
                             "what": "device-timed step with and without the NCCL gradient all-reduce (max over ranks): the difference is the "
                                     "collective's exposed time plus the SM share it takes from the overlapped backward pass"}
    line.update(extras)
    if world == 1 and not args.no_gpu_baseline:
        gb = gpu_baseline(args.steps)
        line["gpu_baseline"] = gb
        try:
            line["gpu_baseline"]["ratio_train_e2e"] = e2e_value / gb["train"]["value"]
            line["gpu_baseline"]["ratio_inference_resident"] = line["inference"]["value"] / gb["inference"]["value"]
            line["gpu_baseline"]["ratio_inference_e2e"] = line["inference"]["e2e"]["value"] / gb["inference"]["e2e"]["value"]
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_step(3, 1, budget_s=30.0)
            if cb:
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["augmentation_port"] = cpu_augmentation_port_seconds()
        except Exception as ex:  # the baseline is reported, never fatal
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
