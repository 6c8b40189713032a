#!/usr/bin/env python
"""bench.py — MADM diffusion feature extraction on B200 (BASELINE.json metric: feature-extract images/s @512^2).

  python bench.py --gpus N --steps K --warmup W            # product arm (CUDA engine through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the fp32 oracle on the host cores

One "step" = one pass of the hot path (VAE encode -> q-sample -> UNet + taps -> GN-bottleneck projections, SURVEY §8
rows a-1..a-9) over one batch of 8 synthetic 512x512 images per GPU (BASELINE.json configs[1]; random-init weights,
seeded as SURVEY §8d).  For N > 1 the driver launches this file with torchrun: one rank per GPU, images batch-sharded,
no collective on the data path (weak scaling); timing = max over ranks of the CUDA-event time of exactly K steps,
bracketed by barrier + synchronize.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "UNet feature-extract images/s @512^2 (VAE-enc + q-sample + SD-1.4 UNet+LoRA taps + projections)"
UNIT = "images/s"
PER_GPU_BATCH = 8
# algorithmic FLOPs per 512^2 image (2*MAC), SURVEY §8d / BASELINE.md §2
GF_PER_IMG = {"vae_encoder": 1116.66, "unet_taps": 803.18, "projections": 14.35, "total": 1934.2}
# --variant s0 (not the default bench line): the vae_decoder_loss configuration of the shipped experiment files:
# + UNet conv_out, VAE decoder 2514.5, Bottleneck(3->128->128) at 512^2 instead of the s2 projection; SURVEY §8d total 4526.0
GF_PER_IMG_S0 = {"vae_encoder": 1116.66, "unet": 803.27, "vae_decoder": 2514.5, "projections": 91.6, "total": 4526.0}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


def cpu_oracle_throughput(timed: int = 3, warm: int = 1, variant: str = "base"):
    """The fp32 oracle (CPU restatement of the reference path) on the host cores: config 1, 1x3x512x512."""
    import torch
    from oracle import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ob = synthetic.build_backbone(with_ema=False, variant=variant)
    img = synthetic.synthetic_images(1)
    with torch.no_grad():
        for _ in range(warm):
            ob(img, input_modal="others")
        t0 = time.perf_counter()
        for _ in range(timed):
            ob(img, input_modal="others")
        dt = (time.perf_counter() - t0) / timed
    return dict(value=1.0 / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"BASELINE config 1: 1x3x512x512 fp32 oracle forward, {warm} warm-up + {timed} timed, {dt:.2f} s/img, "
                       f"{(GF_PER_IMG if variant == 'base' else GF_PER_IMG_S0)['total'] / dt:.0f} GFLOP/s implied"), dt


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path = the oracle port (the reference itself cannot be
    imported here: diffusers/peft/detectron2 absent).  Rank 0 alone runs it."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    base, dt = cpu_oracle_throughput(timed=steps, warm=1, variant=args.variant)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic (seeded, SURVEY §8d)",
        "config": {"workload": "MADM SD-1.4 UNet+LoRA backbone feature extraction, 8x3x512x512 per GPU in the product arm; the CPU "
                               "arm times a bounded sample of it: 1 image per step", "bounded_sample": "1x3x512x512 per step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_product(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from helpers import build_product_backbone, set_lora_adapter

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(1234 + rank)
    bb = build_product_backbone(dev, compute_dtype=args.dtype, variant=args.variant)  # random-init SD-1.4 weights + 2 LoRA adapters (r16)
    gf = GF_PER_IMG if args.variant == "base" else GF_PER_IMG_S0
    ldm = bb.feature_extractor.ldm_extractor
    with torch.no_grad():  # de-degenerate the zero-init pieces like the parity fixtures do
        g = torch.Generator(device=dev).manual_seed(99)
        for _, m in ldm.unet.lora_layers():
            for a in m.lora_B:
                m.lora_B[a].weight.copy_(torch.randn(m.lora_B[a].weight.shape, device=dev, generator=g) * 0.02)
    set_lora_adapter(ldm.unet, "Depth")
    gi = torch.Generator().manual_seed(rank)
    img_host = (torch.rand(B, 3, 512, 512, generator=gi)).pin_memory()
    img_dev = img_host.to(dev, non_blocking=True)
    from madm_b200.engine import OUT_SHAPES
    outs_host = [torch.empty(B, c, s, s, dtype=torch.float32).pin_memory() for c, s in OUT_SHAPES[args.variant]]
    outs_dev = [torch.empty(B, c, s, s, dtype=torch.float32, device=dev) for c, s in OUT_SHAPES[args.variant]]

    # B <= engine.graph_max_batch (8): the backbone replays the step's ~555 launches as one CUDA graph (static input buffer,
    # results cloned out of the static output buffers); larger batches launch on the stream into `outs_dev`.
    graphed = 0 < B <= ldm.engine().graph_max_batch

    def step_resident():
        return bb._extract(img_dev, "others", False, None, out=None if graphed else outs_dev)

    def step_e2e():
        x = img_host.to(dev, non_blocking=True)               # H2D of this step's inputs from pinned memory
        res = bb._extract(x, "others", False, None, out=None if graphed else outs_dev)
        for h, d in zip(outs_host, res["features"]):           # D2H of this step's result (the feature dict)
            h.copy_(d, non_blocking=True)
        return res

    # the same end-to-end work, with the copies of neighbouring steps overlapped with compute on a second stream
    from madm_b200.pipeline import HostPipeline
    pipe = HostPipeline(lambda x: bb._extract(x, "others", False, None), dev)

    def run_e2e_pipelined(steps):
        return pipe.run([img_host] * steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks; the data path itself has no collective
        barrier()
        return ms.item(), clocks

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step_resident()
        ms, clocks = timed(step_resident, args.steps, ClockSampler(local_rank) if rank == 0 else None)
        for _ in range(2):
            step_e2e()
        ms_e2e_serial, _ = timed(step_e2e, args.steps)
        run_e2e_pipelined(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e_pipelined(args.steps)  # K steps: every step's H2D, compute and D2H inside the timed region
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        ms_e2e = t.item()
        # roofline inputs: per-kernel-family CUDA-event timing of instrumented steps on the launch stream
        eng = ldm.engine()
        prof = None
        if rank == 0:
            eng.set_profiling(True)
            acc = None
            psteps = 2
            for _ in range(psteps):
                bb._extract(img_dev, "others", False, None, out=outs_dev)  # stream launches: the per-launch events live there
                p = eng.profile()
                if acc is None:
                    acc = p
                else:
                    for k in p:
                        for f in ("launches", "ms", "flops", "bytes"):
                            acc[k][f] += p[k][f]
            eng.set_profiling(False)
            prof = {k: {f: v[f] / psteps for f in v} for k, v in acc.items()}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = read_peaks()
    imgs = B * world * args.steps
    value = imgs / (ms / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)
    launches = eng.launch_count(B)
    gemm = prof["gemm_tc"]
    gemm_tflops = gemm["flops"] / (gemm["ms"] / 1e3) / 1e12
    total_ms = sum(v["ms"] for v in prof.values())
    gn = prof["groupnorm"]
    gemm_traffic = None
    try:  # measured once under ncu for this build; null if the summary is not in the tree
        with open(os.path.join(ROOT, "profiles", "r01_gemm_dram_traffic.json")) as f:
            gemm_traffic = json.load(f)["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM: all convs / linears)", "bound": "tensor",
        "achieved": gemm_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": gemm_tflops / peaks["tflops_sustained"],
        "peak_source": peaks["source"] + ", sustained cuBLAS bf16 figure (kernel timed inside a long step)",
        "traffic": gemm_traffic, "traffic_unit": "bytes per launch (ncu dram read + write, profiles/r01_gemm_dram_traffic.json)",
        "algorithmic_bytes_per_launch": gemm["bytes"] / max(1, gemm["launches"]),
        "launches_per_step": gemm["launches"], "avg_launch_ms": gemm["ms"] / max(1, gemm["launches"]),
        "algorithmic_gflop_per_launch": gemm["flops"] / max(1, gemm["launches"]) / 1e9,
        "share_of_step": gemm["ms"] / total_ms,
        "families": {k: {"ms_per_step": v["ms"], "launches": v["launches"], "share": v["ms"] / total_ms} for k, v in prof.items()},
        "groupnorm_hbm": {"achieved_gbs": gn["bytes"] / (gn["ms"] / 1e3) / 1e9 if gn["ms"] > 0 else None, "peak_gbs": peaks["hbm_gbs"]},
        "whole_path_tflops": value / world * gf["total"] / 1e3,
    }
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_oracle_throughput(timed=3, warm=1, variant=args.variant)
    ws_gb = eng._ws.numel() / 2 ** 30
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic (seeded rand images, random-init SD-1.4 weights + r16 LoRA, SURVEY §8d)",
        "config": {"workload": (f"BASELINE configs[1]: {B}x3x512x512 per GPU, feature extraction a-1..a-9 "
                                "(VAE-enc + q-sample t=0 + UNet taps + s2..s5 projections); sem_seg_head is SURVEY §8 f-2 (next), not timed")
                   if args.variant == "base" else
                   (f"{B}x3x512x512 per GPU, vae_decoder_loss / s0 variant of the shipped experiment configs (SURVEY §8 a-11): VAE-enc + "
                    "q-sample t=0 + UNet to its final output + VAE decoder + s0/s3/s4/s5 projections"),
                   "variant": args.variant,
                   "per_gpu_batch": B, "global_batch": B * world, "input_modal": "others", "adapter": "Depth_r16_a16 (folded)",
                   "l2": f"working set (packed weights 1.8 GB + workspace {ws_gb:.1f} GB) >> 126 MB L2; no explicit flush",
                   "accumulate": "fp32", "residual_stream": "fp32 (UNet, projections); fp16 in the VAE 512^2 / 256^2 stages with fp16 operands, like the reference's fp16 VAE", "gflop_per_image": gf},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": img_host.numel() * 4, "d2h_bytes_per_step": sum(t.numel() * 4 for t in outs_host),
                "api": "madm_b200.pipeline.HostPipeline around AttentionFeatureExtractorBackbone._extract -> madm_extract (C ABI): pinned "
                       "host buffers, H2D / D2H of neighbouring steps overlapped with compute on a copy stream",
                "serial_value": imgs / (ms_e2e_serial / 1e3), "serial_ms_per_step": ms_e2e_serial / args.steps,
                "cuda_graph": bool(graphed)},
        "gpu_launches": launches * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="madm_b200", choices=["madm_b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"], help="GEMM operand dtype (fp32 accumulate)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="base", choices=["base", "s0"],
                    help="base = BASELINE configs[1] (the bench line); s0 = vae_decoder_loss configuration of the shipped experiment files")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        # convenience: spawn torchrun ourselves when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_product(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
