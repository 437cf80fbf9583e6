#!/usr/bin/env python
"""Benchmark of the Feature2Face generator hot path (BASELINE.json: 512x512 frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--mode parity|fast] [--impl reference]

A step = one pass of the generator over one batch of B synthetic 512x512 frames of the May.yaml ('large')
network (BASELINE.json configs[1]; frames of a clip are independent, so the clip is rendered B frames per call).
Under torchrun (N > 1) every rank renders its own block of the clip and the rendered frames are all-gathered
over NCCL (configs[3]), gather of step i overlapped with the rendering of step i+1; weak scaling.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
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
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "512x512 frames/sec (Feature2Face_G large / May.yaml)"
VARIANT, RECIPE, H, W = "large", "A", 512, 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--mode", default="parity", choices=["parity", "fast"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="large", choices=["large", "normal"],
                    help="large = May.yaml (BASELINE.json configs[1], default); normal = Obama1/Nadella/... (configs[2])")
    ap.add_argument("--no-extras", action="store_true", help="skip the fast-mode / single-frame side measurements")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1451.7), d.get("hbm_gbs", 6572.9), "MEASURED_PEAKS.json (sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.1] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for nm, val in zip(names, r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_net(mode: str):
    from livespeechportraits_b200.generator import Feature2Face_G
    from oracle import f2f_oracle as O
    opt = types.SimpleNamespace(isTrain=False, size=VARIANT, n_downsample_G=8, ngf=64, fp16=0)
    net = Feature2Face_G(opt, precision=mode)
    sd = O.make_state_dict(VARIANT, RECIPE)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), sd


def cpu_reference_fps(frames_per_step: int, steps: int, warmup: int):
    """The reference's own CPU implementation of the path: the unmodified ATen convs on the host cores, through the
    oracle port (the Python reference checkout does not travel to the GPU box)."""
    from oracle import f2f_oracle as O
    sd = O.make_state_dict(VARIANT, RECIPE)
    fm, cand = O.make_inputs(frames_per_step, H, W)
    x = torch.cat([fm, cand], 1)
    # "all the host threads it can use": calibrate the thread count on one frame each (oneDNN does not always
    # scale to every logical CPU of a big host, and the container may be pinned to fewer than os.cpu_count()).
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (torch.get_num_threads(), usable, usable // 2, 64, 32, 16, 8) if 1 <= c <= usable})
    best, best_t = None, None
    for c in cands:
        torch.set_num_threads(c)
        O.generator_forward(sd, x[:1], VARIANT)
        t0 = time.perf_counter()
        O.generator_forward(sd, x[:1], VARIANT)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    for _ in range(warmup):
        O.generator_forward(sd, x, VARIANT)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.generator_forward(sd, x, VARIANT)
    dt = time.perf_counter() - t0
    return frames_per_step * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args, rank: int):
    if rank != 0:
        return
    per_step = min(args.batch, 2)               # bounded sample of the step so K steps finish in minutes
    steps = min(args.steps, 40)
    fps, ms, cores = cpu_reference_fps(per_step, steps, min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded weights with the reference init distribution, seeded inputs)",
        "config": {"workload": f"{VARIANT} 512x512, {per_step} frames per step on the host CPU", "variant": VARIANT,
                   "batch": per_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {per_step} frames, fp32, torch ATen convs (oracle/f2f_oracle.py)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


_REAL_STDOUT = None


def emit(text: str) -> None:
    """The ONE line of the contract, written to the process's original stdout."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


def main():
    global VARIANT, METRIC, _REAL_STDOUT
    args = parse()
    # Libraries chat on stdout (NCCL prints its version line there): keep the original stdout for the JSON line only and
    # send everything else that writes to fd 1 to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    VARIANT = args.variant
    METRIC = f"512x512 frames/sec (Feature2Face_G {VARIANT} / {'May.yaml' if VARIANT == 'large' else 'Obama1.yaml'})"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path for the product arm (use --impl reference for the CPU arm)")
    import torch.distributed as dist
    from oracle import f2f_oracle as O
    from livespeechportraits_b200.pipeline import ClipRenderer

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    net, sd = make_net(args.mode)
    tflops_peak, hbm_peak, peak_src = measured_peaks()

    # ---- synthetic inputs: a pool of feature-map batches (> L2 together with weights and activations)
    pool = 4
    fm_pool, cand = [], None
    for i in range(pool):
        fm, cd = O.make_inputs(B, H, W, seed=100 + rank * 16 + i)
        fm_pool.append(fm.cuda())
        cand = cd[:1].cuda()
    flops_step = net.flops_per_frame(H, W) * B
    launches = net.launches_per_forward()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- output buffers; with N > 1 the tail kernel writes straight into this rank's slot of the gather buffer
    if world > 1:
        gbuf = [torch.empty((world, B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        comm = torch.cuda.Stream(dev)
        gdone = [None, None]
    else:
        obuf = [torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]

    def step(i):
        if world == 1:
            net.render(fm_pool[i % pool], cand, out=obuf[i & 1])
            return
        j = i & 1
        cur = torch.cuda.current_stream(dev)
        if gdone[j] is not None:
            cur.wait_event(gdone[j])
        net.render(fm_pool[i % pool], cand, out=gbuf[j][rank])
        ready = torch.cuda.Event()
        ready.record(cur)
        comm.wait_event(ready)
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(gbuf[j].view(world * B, 3, H, W), gbuf[j][rank])
            ev = torch.cuda.Event()
            ev.record(comm)
            gdone[j] = ev

    def finish():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm)

    # ---- parity spot check of what is being timed (frame 0 of the first pool batch against the oracle)
    parity_err = None
    if rank == 0:
        out0 = net.render(fm_pool[0][:1], cand)
        x0 = torch.cat([fm_pool[0][:1].cpu(), cand.cpu()], 1)
        parity_err = (out0.cpu() - O.generator_forward(sd, x0, VARIANT)).abs().max().item()

    for i in range(Wm):
        step(i)
    finish()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for i in range(K):
        step(i)
    finish()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    launches = net.launches_per_forward()       # pack + convs + split-K finishers of the plan that was just timed
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    # Per-launch breakdown: the same K steps again with a CUDA event after every kernel.  Recording events forces
    # plain stream launches (no CUDA-graph replay, no PDL overlap), so this pass is a little slower than the timed
    # region; it supplies the SHARE of each launch, the timed region supplies the time.
    net.profile_enable(True)
    for i in range(min(K, 200)):
        net.render(fm_pool[i % pool], cand, out=(obuf[i & 1] if world == 1 else gbuf[i & 1][rank]))
    torch.cuda.synchronize()
    prof_ms, prof_n = net.profile_read()
    net.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
    frames = B * K * world
    value = frames / (ms_total / 1e3)

    # ---- roofline of the dominant kernel (the tcgen05 conv family), from the per-launch events of the timed region
    rows = net.layer_table(H, W)
    conv_ms = prof_ms[1:]
    groups = {}
    for r, ms in zip(rows, conv_ms):
        key = f"k{r['kind']} {r['cin']}->{r['cout']} @{r['out_h']}"
        g = groups.setdefault(key, {"ms": 0.0, "flops": 0.0, "launches": 0})
        g["ms"] += ms
        g["flops"] += r["flops"] * B
        g["launches"] += 1
    tot_ms = sum(conv_ms)
    conv_share = tot_ms / (tot_ms + prof_ms[0]) if tot_ms > 0 else 1.0
    ms_step_local = e0.elapsed_time(e1) / K                         # this rank's timed-region time per step (graph replay)
    conv_ms_timed = ms_step_local * conv_share
    achieved = flops_step / (conv_ms_timed / 1e3) / 1e12 if conv_ms_timed > 0 else 0.0
    # tensor work actually issued: folded upsample does 4/9 of the MACs; parity mode issues 3 MMA-equivalents per K step
    exec_flops = 0.0
    for r in rows:
        f = r["flops"] * B
        if r["kind"] in (3, 4):
            f *= 4.0 / 9.0 if r["kind"] == 3 else 1.0
        exec_flops += f * (3.0 if args.mode == "parity" else 1.0)
    top = sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:6]
    # dram__bytes_read.sum + dram__bytes_write.sum of the top kernel, per launch, from the committed ncu --set full capture of
    # this workload (profiles/ncu_traffic.json names the report it was read from); ncu cannot run inside the timed region
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            cand_t = [t for t in json.load(open(tpath)) if t.get("batch") == B and t.get("mode") == args.mode and t.get("variant") == VARIANT]
            traffic = cand_t[0] if cand_t else None
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": "lspg::conv_pair_kernel / conv_umma_kernel / conv_patch_kernel (all conv launches of one step)",
        "achieved": achieved, "peak": tflops_peak, "unit": "TFLOP/s", "frac": achieved / tflops_peak, "peak_source": peak_src,
        "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
        "traffic_detail": traffic,
        "algorithmic_flops_per_step": flops_step, "kernel_ms_per_step": conv_ms_timed,
        "kernel_ms_per_step_event_pass": tot_ms, "pack_input_ms": prof_ms[0], "forwards_profiled": prof_n,
        "executed_tflops": exec_flops / (conv_ms_timed / 1e3) / 1e12 if conv_ms_timed > 0 else None,
        "executed_frac": (exec_flops / (conv_ms_timed / 1e3) / 1e12) / tflops_peak if conv_ms_timed > 0 else None,
        "note": "achieved = reference-conv FLOPs / time of the conv launches in the timed region; parity mode issues "
                "3 tcgen05 MMAs per K step (hi*hi, hi*lo, lo*hi), so executed_* is the tensor-pipe view",
        "top_groups": [{"group": k, "ms": round(v["ms"], 4), "launches": v["launches"],
                        "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["ms"] > 0 else None} for k, v in top],
    }

    # ---- end to end through the public batched API: pinned host feature maps in, pinned host frames out
    # clip length of the end-to-end leg: the 672 frames demo.py renders for 00083.wav (BASELINE.json configs[1])
    Ke = min(K, max(1, 672 // B))
    n_clip = B * Ke
    fm_host = torch.empty((n_clip, 1, H, W), dtype=torch.float32, pin_memory=True)
    for i in range(Ke):
        fm_host[i * B:(i + 1) * B].copy_(fm_pool[i % pool])
    out_host = torch.empty((n_clip, 3, H, W), dtype=torch.float32, pin_memory=True)
    clip = ClipRenderer(net, batch=B, device=dev)
    clip.render_clip(fm_host[: B * min(Wm, Ke)], cand, out_host[: B * min(Wm, Ke)])
    barrier()
    t0 = time.perf_counter()
    clip.render_clip(fm_host, cand, out_host)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    e2e = {"value": n_clip * world / dt, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 4,
           "d2h_bytes_per_step": B * 3 * H * W * 4,
           "frames": n_clip * world,
           "api": "livespeechportraits_b200.pipeline.ClipRenderer.render_clip (pinned host feature maps -> pinned host frames; "
                  "candidates resident on the device as in demo.py:95)"}

    extras = {}
    if rank == 0 and not args.no_extras:
        def timed(fn, n):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        other = "fast" if args.mode == "parity" else "parity"
        ms_o = timed(lambda: net.render(fm_pool[0], cand, out=(obuf[0] if world == 1 else gbuf[0][rank]), precision=other), 10)
        x0 = torch.cat([fm_pool[0][:1].cpu(), cand.cpu()], 1)
        err_o = (net.render(fm_pool[0][:1], cand, precision=other).cpu() - O.generator_forward(sd, x0, VARIANT)).abs().max().item()
        extras[f"{other}_mode"] = {"value": B / ms_o * 1e3, "unit": "frames/s", "max_abs_err_vs_oracle": err_o,
                                   "tflops_algorithmic": flops_step / (ms_o / 1e3) / 1e12}
        # N1: frames leave the GPU as uint8 HWC images (util.tensor2im fused into the tail kernel)
        img_host = torch.empty((n_clip, H, W, 3), dtype=torch.uint8, pin_memory=True)
        clip8 = ClipRenderer(net, batch=B, device=dev, uint8=True)
        clip8.render_clip(fm_host[: B * 2], cand, img_host[: B * 2])
        torch.cuda.synchronize()
        t8 = time.perf_counter()
        clip8.render_clip(fm_host, cand, img_host)
        torch.cuda.synchronize()
        dt8 = time.perf_counter() - t8
        extras["e2e_uint8_images"] = {"value": n_clip / dt8, "unit": "frames/s", "d2h_bytes_per_step": B * H * W * 3,
                                      "api": "ClipRenderer(uint8=True): lspg_forward_image (tensor2im fused, util/util.py:19-42)"}
        # N2: landmark tracks in, uint8 images out - the maps are drawn on the GPU (lspg_draw_feature_maps), so
        # 728 B per frame cross PCIe instead of 1 MB; the host-side cost this removes (88 cv2.line calls per frame,
        # datasets/face_dataset.py:312-323) is timed next to it on one host thread, as the reference runs it
        from oracle import raster_oracle as RO
        lm_np, sh_np = RO.make_landmarks(n_clip, (W, H), seed=3)
        lm_host, sh_host = torch.from_numpy(lm_np).pin_memory(), torch.from_numpy(sh_np).pin_memory()
        clip8.render_clip_from_landmarks(lm_host[: B * 2], sh_host[: B * 2], cand, img_host[: B * 2], (W, H))
        torch.cuda.synchronize()
        tl = time.perf_counter()
        clip8.render_clip_from_landmarks(lm_host, sh_host, cand, img_host, (W, H))
        torch.cuda.synchronize()
        dtl = time.perf_counter() - tl
        ms_r = timed(lambda: net.draw_feature_maps(lm_host[:B].cuda(), sh_host[:B].cuda(), (W, H)), 20)
        chk = net.draw_feature_maps(lm_host[:2].cuda(), sh_host[:2].cuda(), (W, H)).cpu().numpy()
        exact = all(np.array_equal(chk[i, 0] * 255, RO.draw_feature_map_cv2(lm_np[i], (W, H), sh_np[i])) for i in range(2))
        t_cv = time.perf_counter()
        for i in range(32):
            RO.draw_feature_map_cv2(lm_np[i % n_clip], (W, H), sh_np[i % n_clip])
        cv_ms = (time.perf_counter() - t_cv) / 32 * 1e3
        extras["e2e_from_landmarks_uint8"] = {
            "value": n_clip / dtl, "unit": "frames/s", "h2d_bytes_per_step": B * (73 + 18) * 2 * 4, "d2h_bytes_per_step": B * H * W * 3,
            "rasterise_ms_per_batch": ms_r, "bit_exact_vs_cv2": bool(exact), "cv2_host_ms_per_frame": cv_ms,
            "api": "ClipRenderer.render_clip_from_landmarks: lspg_draw_feature_maps + lspg_forward_image"}
        one = fm_pool[0][:1].contiguous()
        ms_1 = timed(lambda: net.render(one, cand), 30)
        extras["single_frame"] = {"value": 1e3 / ms_1, "unit": "frames/s", "ms_per_frame": ms_1, "mode": args.mode,
                                  "note": "batch 1 per call, as demo.py:266 calls inference()"}

    cpu = None
    if rank == 0 and world == 1:
        fps, ms, cores = cpu_reference_fps(1, 12, 2)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "12 frames (+2 warm-up), batch 1, fp32, all host threads, oracle/f2f_oracle.py (torch ATen convs = "
                         "what the reference executes on CPU)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.mode == "fast" else "bf16 hi+lo split operands (3 tcgen05 MMAs per K step), fp32 accumulate",
            "data": "synthetic (seeded weights with the reference's init distribution - no checkpoint ships; seeded inputs)",
            "config": {"workload": f"{'May.yaml (large)' if VARIANT == 'large' else 'Obama1.yaml (normal)'} 512x512, clip rendered in batches of {B} frames per step on one compute stream "
                                   "(BASELINE.json configs[1]/[3]; the one-frame-per-call rate of demo.py is reported as single_frame), "
                                   + ("single GPU" if world == 1 else f"frame-sharded over {world} GPUs + NCCL all-gather of the frames"),
                       "variant": VARIANT, "batch": B, "height": H, "width": W, "precision_mode": args.mode,
                       "parallelism": f"dp{world}",
                       "l2": "per-step working set (weights 0.24-0.49 GB + activations > 1 GB) exceeds the 126 MB L2; "
                             "input batches rotate over a pool"},
            "max_abs_err_vs_oracle": parity_err,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * K,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        line.update(extras)
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
