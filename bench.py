#!/usr/bin/env python
"""bench.py -- stereo pairs/sec for GwcNet @256x512, D=192 (BASELINE.json metric), 1..8 x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one GwcNet inference forward (2D backbone -> cost volume -> 3D aggregation -> soft-argmin) over one batch
of B=8 synthetic SceneFlow-shaped pairs per GPU (BASELINE.json configs[1]); batches shard over the GPUs (weak scaling,
independent pairs) and ONE NCCL all_gather of the per-image EPE partial sums closes the timed region, as in the
reference's eval loop (stereo/modeling/trainer_template.py:313-329).

Reported in one JSON line (rank 0):
  value      pairs/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public model call with PINNED HOST inputs: H2D copy of the images + ground truth
             and D2H read-back of the per-image EPE inside the timed region, every step
  roofline   the cost-volume kernel named by the metric (HBM-bound), timed live with CUDA events around its launch
             inside the timed steps, against MEASURED_PEAKS.json hbm_gbs
  roofline_dominant  the kernel family that dominates the hot path (3x3x3 Conv3d / ConvTranspose3d on tcgen05, 3xTF32),
             same live timing; roofline_cuda_core = the fp32 layers still on CUDA cores
  cpu_baseline  the oracle port of the reference (same aten CPU kernels) on the host cores, bounded sample
--impl reference times that oracle port as the reference arm (the reference is pure Python/PyTorch and cannot travel;
oracle/__init__.py explains the provenance).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stereo_pairs_per_sec_gwcnet_256x512_d192"
CFG = {"MAX_DISP": 192, "USE_CONCAT_VOLUME": True, "CONCAT_CHANNELS": 12, "DOWNSAMPLE": 4, "NUM_GROUPS": 40}
H, W = 256, 512


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def synthetic_weights(model, seed=1):
    """Architecture-shaped random weights (no checkpoints ship with the reference; no network).  Variance-preserving
    normal conv weights and randomised BN statistics, so activations stay O(1) -- default init collapses the logits."""
    gen = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    for key in sorted(sd):
        t = sd[key]
        if key.endswith("num_batches_tracked") or "disp_regression" in key:
            continue
        if key.endswith("running_var"):
            v = torch.rand(t.shape, generator=gen) + 0.5
        elif key.endswith("running_mean") or (t.dim() == 1 and not key.endswith("weight")):
            v = torch.randn(t.shape, generator=gen) * 0.1
        elif t.dim() == 1:
            v = torch.rand(t.shape, generator=gen) * 0.5 + 0.5
        else:
            fan_in = t[0].numel()
            v = torch.randn(t.shape, generator=gen) * (1.0 / fan_in) ** 0.5
        sd[key] = v.to(t.dtype)
    sd["DispProcessor.classif3.2.weight"] = sd["DispProcessor.classif3.2.weight"] * 145.0
    model.load_state_dict(sd)
    return model


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc, self.path = None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, n_gpus):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8 or not f[0].isdigit() or int(f[0]) >= n_gpus:
                    continue
                try:
                    power = float(f[3])
                    clk = float(f[1])
                except ValueError:
                    continue
                mx = float(f[2])
                if power > 250.0:                                  # sample taken under load
                    clocks.append(clk)
                for name, val in zip(names, f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            out["sm_mhz"] = statistics.median(clocks)
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        out["samples_under_load"] = len(clocks)
        return out


def usable_cores():
    """Host threads this process may really use: min(affinity mask, cgroup CPU quota) -- os.cpu_count() alone
    oversubscribes a quota-limited container and makes the CPU arm look far slower than it is."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    elif n_gpus > 1:
        raise SystemExit("--gpus %d needs torchrun (one process per GPU); WORLD_SIZE is 1" % n_gpus)
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def oracle_model():
    from oracle import models as omodels                           # CPU oracle: cpu_baseline / reference arm only
    m = omodels.GwcNet(CFG["MAX_DISP"], CFG["USE_CONCAT_VOLUME"], CFG["CONCAT_CHANNELS"], CFG["DOWNSAMPLE"],
                       CFG["NUM_GROUPS"]).eval()
    return synthetic_weights(m)


def time_cpu(model, pairs_per_step, steps, warmup):
    gen = torch.Generator().manual_seed(0)
    x = {"left": torch.randn(pairs_per_step, 3, H, W, generator=gen), "right": torch.randn(pairs_per_step, 3, H, W, generator=gen)}
    with torch.no_grad():
        for _ in range(warmup):
            model(dict(x))
        t0 = time.perf_counter()
        for _ in range(steps):
            model(dict(x))
        dt = time.perf_counter() - t0
    return dt


def run_reference(args):
    """Reference arm: the oracle port of the reference's CPU path with every host thread, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    torch.set_num_threads(cores)
    model = oracle_model()
    pairs = 1                                                       # bounded sample: one pair of the B=8 batch per step
    dt = time_cpu(model, pairs, args.steps, min(args.warmup, 2))
    value = pairs * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "GwcNet cfgs/gwcnet_sceneflow 256x512 D=192 (configs[1]); CPU step = 1 pair sample",
                   "global_batch": pairs, "parallelism": "cpu-threads"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d forward(s) of 1 pair, oracle port of the reference (same aten CPU kernels)" % args.steps},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = False                         # fp32 end to end (cfg AMP: false; 1e-3 px EPE bar)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    from openstereo_b200 import _lib, host_models, ops

    B = args.batch
    model = synthetic_weights(host_models.GwcNet(CFG)).eval().to(dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    rot = args.rotate                                               # distinct input batches: rot * 12.6 MB > L2 (126 MB)
    host_left = [torch.randn(B, 3, H, W, generator=gen).pin_memory() for _ in range(rot)]
    host_right = [torch.randn(B, 3, H, W, generator=gen).pin_memory() for _ in range(rot)]
    host_gt = [(torch.rand(B, H, W, generator=gen) * 190 + 1).pin_memory() for _ in range(rot)]
    dev_left = [t.to(dev) for t in host_left]
    dev_right = [t.to(dev) for t in host_right]
    dev_gt = [t.to(dev) for t in host_gt]
    host_epe = torch.empty(B, 2).pin_memory()

    def step_resident(i):
        k = i % rot
        with torch.no_grad():
            disp = model({"left": dev_left[k], "right": dev_right[k]})["disp_pred"]
            return ops.epe_partial(disp, dev_gt[k], CFG["MAX_DISP"])

    def step_e2e(i):
        k = i % rot
        with torch.no_grad():
            left = host_left[k].to(dev, non_blocking=True)
            right = host_right[k].to(dev, non_blocking=True)
            gt = host_gt[k].to(dev, non_blocking=True)
            disp = model({"left": left, "right": right})["disp_pred"]
            part = ops.epe_partial(disp, gt, CFG["MAX_DISP"])
            host_epe.copy_(part, non_blocking=True)
            torch.cuda.current_stream().synchronize()               # the caller reads the metric every step
            return part

    from openstereo_b200.distributed import gather_epe_partials

    def gather(part):
        return gather_epe_partials(part)[0]                         # the single collective of the path (NCCL all_gather)

    def timed(step_fn, steps, profile):
        barrier(world)
        torch.cuda.synchronize()
        if profile:
            ops.profile_start()
        launches0 = _lib.launch_count()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        part = None
        for i in range(steps):
            part = step_fn(i)
        allparts = gather(part)
        stop.record()
        torch.cuda.synchronize()
        barrier(world)
        ms = start.elapsed_time(stop)
        prof = ops.profile_stop() if profile else None
        launches = _lib.launch_count() - launches0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms, float(launches)], device=dev, dtype=torch.float64)
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, launches = mx[0].item(), int(sm[1].item())
        return ms, launches, prof, allparts

    for i in range(max(args.warmup, 3)):
        step_resident(i)
        step_e2e(i)
    torch.cuda.synchronize()

    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    if os.environ.get("OSB_NCU_RANGE"):                             # `ncu --profile-from-start off`: capture only the timed steps
        torch.cuda.cudart().cudaProfilerStart()
    ms, launches, prof, parts = timed(step_resident, args.steps, profile=True)
    if os.environ.get("OSB_NCU_RANGE"):
        torch.cuda.cudart().cudaProfilerStop()
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, profile=False)
    clocks = sampler.stop(args.gpus) if rank == 0 else None
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    pairs = B * world * args.steps
    value = pairs / (ms / 1e3)
    e2e_value = pairs / (ms_e2e / 1e3)
    epe = (parts[:, 0] / parts[:, 1].clamp(min=1)).mean().item()
    hbm_peak, peak_src, sm_max = measured_peaks()

    def kernel_stats(name):
        ev = prof.get(name, [])
        t = [a.elapsed_time(b) for a, b in ev]
        return (sum(t) / len(t), len(t), sum(t)) if t else (None, 0, 0.0)

    step_ms = ms / args.steps
    vol_ms, vol_n, vol_total = kernel_stats("osb_gwc_concat_volume_fwd")
    vol_bytes = 4 * (2 * B * (320 + 12) * 64 * 128 + B * 64 * 48 * 64 * 128)
    roofline = None
    if vol_ms:
        ach = vol_bytes / vol_ms / 1e6
        roofline = {"kernel": "volume_kernel (gwc+concat fused, osb_gwc_concat_volume_fwd)", "bound": "hbm",
                    "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                    "frac_of_8TBs_nominal": round(ach / 8000.0, 4), "peak_source": peak_src,
                    "alg_bytes_per_launch": vol_bytes, "ms_per_launch": round(vol_ms, 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of this launch, one `ncu --set full` capture of the same
                    # command (profiles/r1_ncu_summary_final.md, r1_volume_final): 174.2 MB read (= the algorithmic input)
                    # + 745.8 MB written (the rest of the 805.3 MB output is still dirty in the 126 MB L2 at kernel end)
                    "traffic": 919946496 if B == 8 else None,
                    "share_of_step": round(vol_total / ms, 4)}
    # ---- 3D aggregation: MACs per pair of GwcNet-gc at D'=48, H'=64, W'=128
    vox = 48 * 64 * 128
    macs = {
        # stem dres0a (64->32) + dres0b, dres1a, dres1b, classif3a (32->32) at full resolution; 3 x (conv2, conv4)
        "osb_conv3d_k3_tc_fwd": vox * 27 * 32 * (64 + 4 * 32) + 3 * (vox // 8 * 27 * 64 * 64 + vox // 64 * 27 * 128 * 128),
        # 3 x (conv1 32->64 to 1/2, conv3 64->128 to 1/4): MACs counted at the OUTPUT voxels
        "osb_conv3d_k3_s2_tc_fwd": 3 * (vox // 8 * 27 * 32 * 64 + vox // 64 * 27 * 64 * 128),
        # 3 x (conv5 128->64, conv6 64->32): every INPUT voxel feeds 27 taps
        "osb_deconv3d_k3_tc_fwd": 3 * (vox // 64 * 27 * 128 * 64 + vox // 8 * 27 * 64 * 32),
        # redir1 (32->32, full) and redir2 (64->64, half) of the three hourglasses
        "osb_conv1x1_ndhwc_fwd": 3 * (vox * 32 * 32 + vox // 8 * 64 * 64),
        # classif3b 32->1 head
        "osb_conv3d_k3_c1_ndhwc_fwd": vox * 27 * 32,
        # 2D backbone residual blocks on the same kernels (two images per pair): 8 front convs 32->32 @128x256, 30 layer2 convs
        # 64->64, 4 layer3 + 6 dilated layer4 convs 128->128 @64x128 (gwcnet_backbone.py:38-60)
        # + lastconv's 320->128 3x3
        "osb_conv2d_k3_tc_fwd": 2 * 9 * (8 * 128 * 256 * 32 * 32 + 30 * 64 * 128 * 64 * 64 + 10 * 64 * 128 * 128 * 128
                                         + 64 * 128 * 320 * 128),
    }
    agg_macs = sum(v for k, v in macs.items() if k != "osb_conv2d_k3_tc_fwd")      # = 116.30 GMAC, SURVEY.md section 8a row a6
    tc_names = ["osb_conv3d_k3_tc_fwd", "osb_conv3d_k3_s2_tc_fwd", "osb_deconv3d_k3_tc_fwd", "osb_conv2d_k3_tc_fwd"]
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            bf16_peak = float(json.load(f).get("bf16_tflops_sustained"))
    except Exception:
        bf16_peak = 1400.0
    tf32_peak = bf16_peak / 2.0                                     # dense tf32 = half the bf16 rate
    roof_dom = None
    tc_total = sum(kernel_stats(n)[2] for n in tc_names)
    if tc_total > 0:
        per = {}
        for n in tc_names:
            _, cnt, tot = kernel_stats(n)
            if tot > 0:
                per[n] = {"launches_per_step": cnt // args.steps, "ms_per_step": round(tot / args.steps, 3),
                          "useful_tflops": round(2 * macs[n] * B * args.steps / (tot / 1e3) / 1e12, 1)}
        ran = [n for n in tc_names if kernel_stats(n)[2] > 0]
        useful = 2 * sum(macs[n] for n in ran) * B * args.steps / (tc_total / 1e3) / 1e12
        roof_dom = {"kernel": "tcgen05 conv family: conv3d_tc_kernel / conv3d_tcg_kernel (3x3x3 s1), conv3d_tcs2_kernel (s2), "
                              "conv3d_tcdc_kernel (transposed), incl. the backbone's 3x3 residual blocks as one-plane volumes "
                              "(osb_conv2d_k3_tc_fwd); kind::tf32 with the 3xTF32 split",
                    "bound": "tensor", "achieved": round(3 * useful, 1), "peak": round(tf32_peak, 1), "unit": "TFLOP/s",
                    "frac": round(3 * useful / tf32_peak, 4), "useful_fp32_equivalent_tflops": round(useful, 1),
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (dense tf32 rate); achieved counts the 3 MMAs "
                                   "issued per fp32-accurate product",
                    "alg_flops_per_step": 2 * sum(macs[n] for n in ran) * B, "share_of_step": round(tc_total / ms, 4),
                    "per_kernel": per,
                    # conv3d_tc_kernel<32> (32->32 stem layer), ncu --set full: 403.2 MB read + 362.8 MB written per launch =
                    # its algorithmic 402.7 + 402.7 MB (profiles/r1_ncu_summary_final.md, r1_conv3d_tc_final)
                    "traffic": 765931264 if B == 8 else None}
    cc_names = ["osb_conv3d_k3_bn_act_fwd", "osb_deconv3d_bn_act_fwd", "osb_conv3d_1x1_bn_act_fwd", "osb_conv1x1_ndhwc_fwd",
                "osb_conv3d_k3_c1_ndhwc_fwd"]
    cc_total = sum(kernel_stats(n)[2] for n in cc_names)
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12                 # derived: SMs x fp32 lanes x 2 x max clock
    roof_cc = None
    if cc_total > 0:
        cc_flops = 2 * (agg_macs - (sum(macs[n] for n in tc_names[:3]) if tc_total > 0 else 0)) * B
        ach = cc_flops * args.steps / (cc_total / 1e3) / 1e12
        roof_cc = {"kernel": "fp32 CUDA-core layers left in the aggregation (classif3b 32->1 head conv3d_k3_c1_ndhwc_kernel, channels-last "
                             "1x1 redir convs; everything else when the tensor-core variants do not cover a shape)",
                   "bound": "fp32_fma", "achieved": round(ach, 2), "peak": round(fp32_peak, 1), "unit": "TFLOP/s",
                   "frac": round(ach / fp32_peak, 4), "peak_source": "derived 148 SM x 128 lanes x 2 x %.0f MHz" % sm_max,
                   "alg_flops_per_step": cc_flops, "share_of_step": round(cc_total / ms, 4), "traffic": None}
    shares = {}
    for name, ev in prof.items():
        shares[name] = round(sum(a.elapsed_time(b) for a, b in ev) / ms, 4)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = usable_cores()
        torch.set_num_threads(cores)
        cm = oracle_model()
        n = 4
        dt = time_cpu(cm, 1, n, 1)
        cpu = {"value": round(n / dt, 4), "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": "%d forwards of 1 pair (of the B=%d batch) through the oracle port of the reference GwcNet "
                         "(same aten CPU kernels), %d threads" % (n, B, cores)}

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "GwcNet cfgs/gwcnet/gwcnet_sceneflow.yaml, batch %d/GPU @256x512 D=192 (BASELINE configs[1])" % B,
                   "global_batch": B * world, "parallelism": "dp%d batch-shard, 1 all_gather of per-image EPE" % world,
                   "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2); per-step activations ~6 GB" % (rot, rot * 2 * B * 3 * H * W * 4 / 1e6),
                   "weights": "synthetic seeded init (no checkpoints ship with the reference)"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 3), "unit": "pairs/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                "h2d_bytes_per_step": (2 * B * 3 * H * W + B * H * W) * 4, "d2h_bytes_per_step": B * 2 * 4},
        "gpu_launches": launches,
        "roofline": roofline, "roofline_dominant": roof_dom, "roofline_cuda_core": roof_cc, "kernel_share_of_step": shares,
        "cpu_baseline": cpu, "mean_epe_vs_synthetic_gt": round(epe, 3),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--rotate", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
