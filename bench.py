#!/usr/bin/env python
"""bench.py -- appearance-optimisation steps/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (BASELINE.json configs[1]): one train_clip step = render 512 rays x (64+64) samples through the 8x256
SDF + 4x256 colour MLPs (placement + fine pass), shading/canvas/losses on a 224x224 canvas, CLIP ViT-B/32 on the
texture and the shading canvas, full backward (second-order SDF terms included) and Adam.  Synthetic data:
seeded geometric-init weights, seeded random CLIP weights (the real ViT-B-32.pt is not on disk), disc-shaped
silhouettes (avatarclip_b200/workload.py).

N > 1: one camera view per rank per step (weak scaling), one NCCL all-reduce of the flat fp32 gradient per step.

Native arm JSON keys follow the driver contract; see DESIGN.md "Measurement" for how each number is taken.
"""
from __future__ import annotations

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
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "appearance-optim steps/sec (512 rays x 128 samples, CLIP loss)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu (profiles/r1_launches_tcgen05_engine.txt): mean over the
# 73 NT / 21 TN launches of one step (NT: 7.99 GB per step; the launches of the fine pass move 87-350 MB each)
TRAFFIC_PER_LAUNCH = {"avc::tc::gemm_tc_tn_kernel": 128.6e6, "avc::tc::gemm_tc_nt_kernel": 109.4e6}
N_RAYS, CANVAS = 512, 224
SDF_KW = dict(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
              geometric_init=True, weight_norm=True)
COL_KW = dict(d_feature=256, mode="no_view_dir", d_in=6, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
              multires_view=0, squeeze_out=True, extra_color=True)
REN_KW = dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0, extra_color=True)


def nt_designed_bytes_per_step():
    """HBM bytes the NT launches of one step move BY DESIGN (unfused layers: every linear reads its split operand and
    its stashed epilogue operands and writes its outputs once), B2 config, in units of one 256-wide fp32 row (1 KB):
    fine pass per point: value 8 layers (reads 7.16, writes sp' stash + split 16), features 2, gradient chain 7 x 3
    + 1.3, colour 4 x 2 + 1, colour dgrad 3 x 2.5 + 1 + 2, second-order sweep 8 x 5, value dgrad 8 x 4;
    placement passes per point: 8 layers, split in / split out (15.16)."""
    fine = 23.16 + 2 + 21 + 1.3 + 9 + 8.5 + 2 + 40 + 32
    place = 15.16
    p_fine = N_RAYS * (REN_KW["n_samples"] + REN_KW["n_importance"])
    p_place = N_RAYS * (REN_KW["n_samples"] + 3 * REN_KW["n_importance"] // 4)
    return 1024.0 * (fine * p_fine + place * p_place)


def algorithmic_flops_per_step():
    """(6.875 F_sdf + 3 F_col) per fine sample point, F = 2*MAC; + CLIP fwd+dgrad for two images.
    Returns (all MLP FLOP, the part executed by the NT tiles, the part executed by the TN (weight-gradient) tiles,
    CLIP FLOP).  NT: value fwd 1 + gradient chain 1 + placement passes 0.875 + value dgrad 1 + second-order sweep 1
    = 4.875 F_sdf, colour fwd + dgrad = 2 F_col.  TN: 2 F_sdf + 1 F_col."""
    mac_sdf = 39 * 256 + 6 * 256 * 256 + 256 * 217 + 256 * 257        # 9 linears of the 8x256 net
    mac_col = 262 * 256 + 3 * 256 * 256 + 2 * 3 * 256                 # 4x256 + two 3-wide heads
    pts = N_RAYS * (REN_KW["n_samples"] + REN_KW["n_importance"])
    mlp = (6.875 * 2 * mac_sdf + 3 * 2 * mac_col) * pts
    nt = (4.875 * 2 * mac_sdf + 2 * 2 * mac_col) * pts
    tn = (2.0 * 2 * mac_sdf + 1 * 2 * mac_col) * pts
    clip = 2 * 17.6e9
    return mlp, nt, tn, clip


def host_cores() -> int:
    """Usable host cores: the cgroup CPU quota when there is one (the GPU boxes expose 128 logical CPUs but
    grant a 16-CPU quota; 128 threads then run ~50x slower than 16), else the affinity mask."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1425.6), d.get("hbm_gbs", 6566.4), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_world(device, engine):
    import util_neus as U
    from oracle import clip_vit as cv          # weights generator only (seeded random ViT-B/32 state)
    from avatarclip_b200.clip_vit import ClipImageTower
    from avatarclip_b200.trainer import AppearanceTrainer
    sp, cp = U.synth_state(SDF_KW, COL_KW, seed=0)
    clip_sd = cv.random_vit_state(seed=0)
    text = torch.randn(2, 512, generator=torch.Generator().manual_seed(5))
    sdf, col, var, ren = U.build_product(SDF_KW, COL_KW, REN_KW, sp, cp, 0.3, device, engine=engine, chunk=4096)
    tower = ClipImageTower(clip_sd, device=device)
    return sp, cp, clip_sd, text, ren, tower


def kernel_key(name):
    """'void avc::tc::gemm_tc_nt_kernel<128, 3, ...>(...)' -> 'avc::tc::gemm_tc_nt_kernel';
    'void (anonymous namespace)::k_gemm16<EpiFc>(...)' -> 'k_gemm16'."""
    import re
    n = re.sub(r"^void\s+", "", name).replace("(anonymous namespace)::", "")
    m = re.match(r"[A-Za-z_][A-Za-z0-9_:]*", n)
    return m.group(0) if m else n


def count_my_launches(fn):
    """Kernels of libavc_b200.so launched by one call of fn (torch.profiler / CUPTI), by name."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        mine, total, table = 0, 0, {}
        for ev in prof.events():
            if ev.device_type is not None and str(ev.device_type).endswith("CUDA") and ev.name:
                total += 1
                n = ev.name
                if "avc::" in n or "k_gemm16" in n or "(anonymous namespace)::k_" in n or n.startswith("k_"):
                    mine += 1
                    key = kernel_key(n)
                    t = table.setdefault(key, [0, 0.0])
                    t[0] += 1
                    t[1] += ev.device_time if hasattr(ev, "device_time") else getattr(ev, "cuda_time", 0.0)
        return mine, total, table
    except Exception as e:  # pragma: no cover
        return None, None, {"error": repr(e)}


def run_native(args):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints "NCCL version ..." on STDOUT when NCCL_DEBUG=VERSION/INFO: keep stdout to the one JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("AVC_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    from avatarclip_b200.trainer import AppearanceTrainer, DeviceView
    from avatarclip_b200.workload import make_view
    sp, cp, clip_sd, text, ren, tower = build_world(device, args.engine)
    tr = AppearanceTrainer(ren, tower, text, lr=5e-4, process_group=pg, device=device)
    K, Wm = args.steps, args.warmup
    n_views = 8
    views = [make_view(rank + world * i, n_rays=N_RAYS, H=CANVAS, W=CANVAS, seed=0, bg_choice=3, pin=True)
             for i in range(n_views)]
    dv = DeviceView(views[0], device)
    resident = [DeviceView(v, device) for v in views]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) + per-phase events for the roofline
    for i in range(Wm):
        tr.step(resident[i % n_views])
    barrier()
    # one extra step on EVERY rank (it contains the all-reduce); only rank 0 records it with CUPTI
    launches_per_step, total_launches, table = (None, None, {})
    if rank == 0:
        launches_per_step, total_launches, table = count_my_launches(lambda: tr.step(resident[0]))
    else:
        tr.step(resident[0])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    use_graph = bool(args.graph)
    if use_graph:
        # the whole step as one CUDA graph on the static view `dv`; each step refreshes dv with ONE copy
        dv.copy_from(resident[0])
        tr.capture(dv)
        for i in range(2):
            dv.copy_from(resident[i % n_views])
            tr.replay(dv)
        barrier()

    def one_step(i):
        if use_graph:
            dv.copy_from(resident[i % n_views])      # D2D, inputs resident in HBM
            return tr.replay(dv)
        return tr.step(resident[i % n_views])

    ev[0].record()
    for i in range(K):
        one_step(i)
    ev[1].record()
    barrier()
    ms_value = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if rank == 0 else None

    # render-only share (events around the two render calls) for the roofline of the MLP contractions
    from avatarclip_b200 import renderer as R
    tot_render = 0.0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for i in range(min(K, 5)):
        d = resident[i % n_views]
        e[0].record()
        out, ws, chunk = R.render_forward_raw(ren, d.rays_o, d.rays_d, d.near, d.far, d.jitter, None, 0, 1.0, None,
                                              keep_ws=False, out=tr._out)
        e[1].record()
        cot = {k: torch.ones_like(out[k]) for k in ("color_fine", "extra_color_fine", "weights", "gradients")}
        e[2].record()
        R.render_backward_raw(ren, d.rays_o, d.rays_d, None, 0, 1.0, out, ws, chunk, cot, grad=tr.grad)
        e[3].record()
        torch.cuda.synchronize()
        tot_render += e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3])
    ms_render = tot_render / min(K, 5)

    # per-phase device times (CUDA events between the C-ABI calls), averaged over a few extra steps
    phases = {}
    for i in range(5):
        tr.phase_events = []
        tr.step(resident[i % n_views])
        torch.cuda.synchronize()
        evs = tr.phase_events
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            phases[n1] = phases.get(n1, 0.0) + e0.elapsed_time(e1) / 5.0
    tr.phase_events = None

    # ---------------- end-to-end timing: pinned-host inputs copied every step, loss read back every step
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    # The loss of step i is copied to pinned host memory on the stream right after step i and read by the host
    # while step i+1 runs (the reference logs its loss the same lagged way only every report_freq steps); every
    # step's loss is read inside the timed region, the last one before the closing synchronise.
    loss_host = torch.zeros(K, dtype=torch.float32).pin_memory()
    done = [torch.cuda.Event() for _ in range(K)]
    ev2[0].record()
    last = 0.0
    for i in range(K):
        dv.upload(views[i % n_views])                # one H2D copy of the packed, pinned view
        loss = tr.replay(dv) if use_graph else tr.step(dv)
        loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        done[i].record()
        if i > 0:
            done[i - 1].synchronize()
            last = float(loss_host[i - 1])
    done[K - 1].synchronize()
    last = float(loss_host[K - 1])
    ev2[1].record()
    barrier()
    ms_e2e = ev2[0].elapsed_time(ev2[1])

    t = torch.tensor([ms_value, ms_e2e], device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_value, ms_e2e = t.tolist()

    if rank == 0:
        mlp_flops, nt_flops, tn_flops, clip_flops = algorithmic_flops_per_step()
        peak_tf, peak_hbm, peak_src = peaks()
        achieved_step = mlp_flops / (ms_render * 1e-3) / 1e12
        # The CLIP tower's kernels are chained with programmatic dependent launch: each starts early and its CUPTI
        # duration includes the time it spends in griddepcontrol.wait for its predecessor, so those durations overlap
        # and do not add up (phases_ms.clip_* are the true times).  The dominant kernel is picked among the rest.
        plain = {k: v for k, v in table.items() if isinstance(v, list) and k.startswith("avc::")}
        dom = max(plain.items(), key=lambda kv: kv[1][1])[0] if plain else None
        # dominant kernel: algorithmic FLOP its launches execute per step / the sum of their device durations in the
        # profiled step (CUPTI kernel records taken live in this process, not under ncu)
        kern_us = {k: v for k, v in table.items() if isinstance(v, list)}
        dom_flops = {"avc::tc::gemm_tc_nt_kernel": nt_flops, "avc::tc::gemm_tc_tn_kernel": tn_flops,
                     "avc::gemm_nt_kernel": nt_flops, "avc::gemm_tn_kernel": tn_flops}.get(dom)
        if dom_flops is not None and dom in kern_us and kern_us[dom][1] > 0:
            n_launch, dom_us = kern_us[dom]
            achieved = dom_flops / (dom_us * 1e-6) / 1e12
            per_launch = {"launches_per_step": n_launch, "avg_launch_us": dom_us / n_launch,
                          "algorithmic_gflop_per_launch": dom_flops / n_launch / 1e9}
        else:
            achieved, per_launch = achieved_step, None
        line = {
            "metric": METRIC, "value": world * K / (ms_value * 1e-3), "unit": "steps/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.engine == 0 else "bf16x3(split)->f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 512 rays x (64+64) samples, 8x256 SDF + 4x256 colour, "
                                   "CLIP ViT-B/32 loss on 2 canvases 224x224, Adam; 1 view per GPU per step",
                       "views_per_step": world, "engine": "fp32 FFMA tiles" if args.engine == 0 else "tcgen05 split",
                       "l2": "per-step working set (activation stash ~2.4 GB) >> 126 MB L2; no flush needed",
                       "parallelism": f"view-sharded dp{world}" if world > 1 else "single",
                       "launch": ("one CUDA graph per step" + (" + all-reduce + Adam" if world > 1 else ""))
                                 if use_graph else "eager C-ABI calls"},
            "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": "steps/s",
                    "h2d_bytes_per_step": views[0].h2d_bytes(), "d2h_bytes_per_step": 4},
            "gpu_launches": None if launches_per_step is None else launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         "traffic": TRAFFIC_PER_LAUNCH.get(dom),
                         "per_launch": per_launch,
                         "scope": "algorithmic FLOP executed by the dominant kernel's launches in one step (SURVEY 8d "
                                  "split: NT tiles 4.875 F_sdf + 2 F_col per point, TN tiles 2 F_sdf + F_col) / sum of "
                                  "their device durations (CUPTI, live); peak = " + peak_src + "; the kernel runs 3 "
                                  "bf16 MMAs per product (two-term split), so its ceiling is peak/3",
                         "hbm_view": (lambda b, us: {"designed_bytes_per_step": b, "kernel_us_per_step": us,
                                                     "achieved_GBps": b / (us * 1e-6) / 1e9 if us else None,
                                                     "peak_GBps": peak_hbm,
                                                     "frac": (b / (us * 1e-6) / 1e9 / peak_hbm) if us else None,
                                                     "note": "the layers are not fused, so the NT launches are bound by "
                                                             "the activation traffic they move by design (operands in, "
                                                             "stash + split out), not by the tensor pipe"})(
                             nt_designed_bytes_per_step(), kern_us.get("avc::tc::gemm_tc_nt_kernel", [0, 0.0])[1]),
                         "step_level": {"achieved": achieved_step, "frac": achieved_step / peak_tf,
                                        "ms_render_fwd_bwd": ms_render,
                                        "note": "all MLP FLOP/step (0.577 T) / CUDA-event time of render fwd+bwd"},
                         "kernel_time_us_per_step": {k: round(v[1], 1) for k, v in
                                                     sorted(((k, v) for k, v in kern_us.items() if k.startswith("avc::")),
                                                            key=lambda kv: -kv[1][1])[:8]}},
            "last_loss": last,
            "phases_ms": {k: round(v, 4) for k, v in phases.items()},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sp, cp, clip_sd, text, views[0], sample_rays=args.cpu_sample_rays)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def oracle_trainer(sp, cp, clip_sd, text):
    import util_neus as U
    from oracle.train_step import OracleTrainer
    sconf, cconf, rconf = U.confs_from_kw(SDF_KW, COL_KW, REN_KW)
    return OracleTrainer(sconf, cconf, rconf, sp, cp, 0.3, clip_sd, text, lr=5e-4)


def subsample_view(hv, n):
    """First n rays of the view (a contiguous part of the disc), same canvas."""
    import copy
    import numpy as np
    v = copy.copy(hv)
    for name in ("rays_o", "rays_d", "near", "far", "jitter", "pix", "ray_background"):
        t = getattr(hv, name)
        if t is not None:
            setattr(v, name, t[:n].clone())
    m = torch.zeros_like(hv.in_mask)
    m[v.pix.long()] = 1
    v.in_mask = m
    v.flat = None          # the tensors above are no longer views of the packed buffer
    return v


def timed_oracle_step(orc, view, full_rays):
    """One oracle step on a ray sample; the ray-march part (render fwd + its backward) is linear in rays and is
    scaled to the full ray count, the CLIP / loss-stage part is run and counted in full."""
    R = view.rays_o.shape[0]
    t0 = time.perf_counter()
    total, aux = orc.loss(view)
    t1 = time.perf_counter()
    outs = aux["out"]
    # weight_sum is derived from `weights` inside render(), so its path is covered by the `weights` cotangent
    keys = ["color_fine", "extra_color_fine", "gradients", "weights", "gradient_error"]
    t3 = time.perf_counter()
    cots = torch.autograd.grad(total, [outs[k] for k in keys], retain_graph=True, allow_unused=True)
    t4 = time.perf_counter()
    params = [p for _, p in orc.named_params()]
    live = [(outs[k], c) for k, c in zip(keys, cots) if c is not None]
    grads = torch.autograd.grad([o for o, _ in live], params, [c for _, c in live], allow_unused=True)
    t5 = time.perf_counter()
    for p, g in zip(params, grads):
        p.grad = g
    orc.opt.step()
    t6 = time.perf_counter()
    return {"fwd_all": t1 - t0, "bwd_clip_stage": t4 - t3, "bwd_render": t5 - t4, "adam": t6 - t5, "rays": R,
            "loss": float(total)}


def cpu_baseline(sp, cp, clip_sd, text, view, sample_rays=128):
    """Oracle ('port') step timed on this box's host cores.  The render part runs on `sample_rays` of the 512 rays
    and is scaled linearly; CLIP + loss stage + Adam run in full."""
    import oracle.neus as on
    torch.set_num_threads(host_cores())
    orc = oracle_trainer(sp, cp, clip_sd, text)
    v = subsample_view(view, sample_rays)
    # time the render forward alone to split fwd_all
    t0 = time.perf_counter()
    with torch.no_grad():
        z = on.hierarchical_z(lambda x: on.sdf_value(orc.sp, orc.sconf, x), orc.rconf, v.rays_o, v.rays_d,
                              v.near.reshape(-1, 1), v.far.reshape(-1, 1), v.jitter.reshape(-1, 1))
    t_place = time.perf_counter() - t0
    timed_oracle_step(orc, v, N_RAYS)          # untimed warm-up (allocator, thread pool)
    tm = timed_oracle_step(orc, v, N_RAYS)
    # forward of render_core alone (with graph) ~ fwd_all - placement - clip/stage forward; measure clip fwd directly
    from oracle import clip_vit as cv
    canv = torch.rand(CANVAS, CANVAS, 3)
    t0 = time.perf_counter()
    with torch.no_grad():
        for b in range(2):
            cv.clip_cosine(orc.clip_state, canv, orc.text[b])
    t_clip_fwd = time.perf_counter() - t0
    scale = N_RAYS / float(sample_rays)
    t_render_fwd = max(tm["fwd_all"] - t_clip_fwd, 0.0)
    t_full = t_render_fwd * scale + t_clip_fwd + tm["bwd_clip_stage"] + tm["bwd_render"] * scale + tm["adam"]
    return {"value": 1.0 / t_full, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"ray-march on {sample_rays}/{N_RAYS} rays scaled x{scale:g} (linear in rays); CLIP x2, loss stage, "
                      f"Adam in full; fp32 torch CPU; s/step est. {t_full:.2f}",
            "detail": {k: round(v, 4) if isinstance(v, float) else v for k, v in tm.items()}}


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (the oracle port: /root/reference is a
    Python package that cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import util_neus as U
    from oracle import clip_vit as cv
    from avatarclip_b200.workload import make_view
    torch.set_num_threads(host_cores())
    sp, cp = U.synth_state(SDF_KW, COL_KW, seed=0)
    clip_sd = cv.random_vit_state(seed=0)
    text = torch.randn(2, 512, generator=torch.Generator().manual_seed(5))
    orc = oracle_trainer(sp, cp, clip_sd, text)
    sample = args.cpu_sample_rays
    views = [subsample_view(make_view(i, n_rays=N_RAYS, H=CANVAS, W=CANVAS, seed=0, bg_choice=3), sample) for i in range(4)]
    scale = N_RAYS / float(sample)
    canv = torch.rand(CANVAS, CANVAS, 3)
    with torch.no_grad():
        t0 = time.perf_counter()
        for b in range(2):
            cv.clip_cosine(orc.clip_state, canv, orc.text[b])
        t_clip_fwd = time.perf_counter() - t0
    for i in range(args.warmup):
        timed_oracle_step(orc, views[i % 4], N_RAYS)
    tot = 0.0
    for i in range(args.steps):
        tm = timed_oracle_step(orc, views[i % 4], N_RAYS)
        t_render_fwd = max(tm["fwd_all"] - t_clip_fwd, 0.0)
        tot += t_render_fwd * scale + t_clip_fwd + tm["bwd_clip_stage"] + tm["bwd_render"] * scale + tm["adam"]
    val = args.steps / tot
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] (same as the native arm), reference algorithm on host cores"},
            "cpu_baseline": {"value": val, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"per step: ray-march on {sample}/{N_RAYS} rays, time scaled x{scale:g}; CLIP x2, "
                                       "loss stage, Adam in full"},
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--engine", type=int, default=int(os.environ.get("AVC_ENGINE", "1")),
                    help="MLP contraction engine: 1 = tcgen05 split-bf16 tiles (default), 0 = fp32 FFMA tiles")
    ap.add_argument("--graph", type=int, default=int(os.environ.get("AVC_GRAPH", "1")),
                    help="1: replay the step as one captured CUDA graph (default); 0: eager C-ABI calls")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rays", type=int, default=64)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
        run_native(args)


if __name__ == "__main__":
    main()
