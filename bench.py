#!/usr/bin/env python
"""Headline benchmark: frames/sec through sampler + ViT-g + Q-Former (+ projection) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic batch of BASELINE.json config 2 per rank
(32 videos x 64 candidate frames 224^2, 32-frame optical flow, sampler selects 32 frames per video -> 1024 frames
through EVA-ViT-g (39 layers), the BLIP-2 Q-Former and language_projection -> [32, 1024, 2048] LLM tokens):

  value  device-timed (CUDA events, max over ranks) with the inputs already resident in HBM
  e2e    the same step through the public API (`VideoTGBEncoder.encode_u8`) from PINNED HOST buffers: H2D of the u8
         frames, the flow and the ids inside the timed region, D2H of the resulting tokens
  roofline      the tcgen05 GEMM kernel family: algorithmic flops of its launches / their CUDA-event time
  cpu_baseline  the oracle port (reference algorithm, fp32 torch CPU, all host threads) on a bounded sample

Weak scaling: every rank processes its own 32 videos; the only exchange is the NCCL all-gather of the bf16 Q-Former
tokens (48 KiB/frame).  `--impl reference` times the reference algorithm (oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec sampler+ViT-g+Q-Former"
UNIT = "frames/s"

# algorithmic GFLOP per selected frame: ViT-g 520.72, Q-Former (BLIP-2) 12.75, projection 0.10
GFLOP_PER_FRAME = 533.57


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--videos", type=int, default=32)
    ap.add_argument("--nframe", type=int, default=32)
    ap.add_argument("--candidates", type=int, default=64)
    ap.add_argument("--lof", type=int, default=32)
    ap.add_argument("--tq", type=int, default=16)
    ap.add_argument("--vit-layers", type=int, default=39, help="debug only; the reported config uses 39")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: 1 untimed warm-up + 1 step of the resident path, nothing else (not a bench value)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU (oracle) arm
def cpu_reference(frames: int, steps: int, warmup: int, vit_layers: int = 39, lof: int = 32, tq: int = 16):
    """The reference algorithm (oracle port, fp32) on the host cores: 1 video, `2*frames` candidates, `frames` selected.
    Returns (frames_per_s, ms_per_step, cores)."""
    import torch
    from oracle import select_ref, torch_ref, transform_ref
    from videotgb_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vsd = syn.vision_state_dict(syn.VisionCfg(layers=vit_layers))
    qsd, ssd, hsd = syn.qformer_state_dict(), syn.sampler_state_dict(), syn.head_state_dict()
    n = 2 * frames
    u8 = syn.synthetic_frames_u8(n, seed=3)
    of = syn.synthetic_flow(1, lof, seed=5)
    of_mask = torch.ones(1, lof + 2, dtype=torch.long)
    g = torch.Generator().manual_seed(9)
    qids = torch.randint(1000, 30522, (1, tq), generator=g)
    qmask = torch.ones(1, tq, dtype=torch.long)

    def step():
        with torch.no_grad():
            px = torch.from_numpy(transform_ref.normalize_u8_thwc(u8.numpy())).view(1, n, 3, 224, 224)
            _, logits = torch_ref.sampler_forward(ssd, of, of_mask, qids, qmask, mode="fusion")
            sl, el = logits.split(1, dim=-1)
            cs, ce = [], []
            for _ in range(2):
                gum = -torch.empty(2, lof, 1).exponential_(generator=g).log()
                s, e = select_ref.gumbel_argmax(sl, el, gum)
                cs.append(s); ce.append(e)
            idx = select_ref.select_batch(cs, ce, [lof], n, frames)
            sampled = select_ref.gather_frames(px, idx)
            img, _ = torch_ref.vit_forward(vsd, sampled, layers=vit_layers)
            qt = hsd["query_tokens"].expand(frames, -1, -1)
            qo = torch_ref.qformer_forward(qsd, qt, img, None)
            return torch_ref.project_tokens(hsd, qo, 1)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return frames / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = args.cpu_frames if (args.steps + args.warmup) <= 10 else max(2, args.cpu_frames // 2)
    fps, ms, cores = cpu_reference(frames, args.steps, args.warmup, args.vit_layers, args.lof, args.tq)
    sample = f"1 video x {2 * frames} candidates -> {frames} frames per step, oracle port (torch fp32 CPU)"
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"BASELINE config 2: BLIP2-Flan-T5-xl path, {args.videos} videos x {args.nframe} sampled frames 224^2 "
                        f"(from {args.candidates} candidates, {args.lof}-frame flow) per GPU, bf16",
            "videos_per_gpu": args.videos, "nframe": args.nframe, "candidates": args.candidates, "flow_frames": args.lof,
            "vit_layers": args.vit_layers, "llm_hidden": 2048, "parallelism": "frame/video sharding, dp",
            "l2": "inputs_larger_than_l2"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], 0, set(), []
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                if float(p[2]) < 300:          # idle samples around the region
                    continue
                sm.append(int(float(p[0]))); mx = max(mx, int(float(p[1]))); pw.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def gemm_flops(shape):
    m, n, k, _ = shape
    return 2.0 * m * n * k


def run_ours(args):
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace as NS
    from videotgb_b200 import ops
    from videotgb_b200.distributed import TokenAllGather
    from videotgb_b200.pipeline import VideoTGBEncoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)

    with torch.device(dev):
        enc = VideoTGBEncoder(NS(num_hidden_layers=args.vit_layers), None, NS(fusion_layer=6, encoder_width=768), 2048)
        with torch.no_grad():
            enc.model.query_tokens.normal_(0.0, 0.02)
    enc.eval()
    if world > 1:
        enc.token_exchange = TokenAllGather()

    b, n, k, lof, tq = args.videos, args.candidates, args.nframe, args.lof, args.tq
    g = torch.Generator().manual_seed(1234 + rank)
    host = {
        "frames_u8": torch.randint(0, 256, (b, n, 224, 224, 3), generator=g, dtype=torch.uint8).pin_memory(),
        "of": (torch.randn(b, lof, 2, 224, 224, generator=g) * 0.25).pin_memory(),
        "sampler_question": torch.randint(1000, 30522, (b, tq), generator=g).pin_memory(),
        "sampler_question_attention_mask": torch.ones(b, tq, dtype=torch.long).pin_memory(),
    }
    of_lengths = torch.randint(8, lof + 1, (b,), generator=g).tolist()
    of_mask = torch.zeros(b, lof + 2, dtype=torch.long)
    for j, ln in enumerate(of_lengths):
        of_mask[j, : ln + 2] = 1
    host["of_mask"] = of_mask.pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    of_len_dev = torch.tensor(of_lengths, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    def to_batch(tensors):
        d = dict(tensors)
        d["of_lengths"] = of_len_dev
        d["nframe"] = k
        return d

    resident = to_batch({kk: v.to(dev) for kk, v in host.items()})
    out_host = torch.empty((b, k * 32, 2048), dtype=torch.float32).pin_memory()
    d2h_bytes = out_host.numel() * 4

    def step_resident():
        out = enc.encode_u8(resident, generator=gen)
        if enc.token_exchange is not None:
            enc.token_exchange.wait()
        return out

    # e2e: the same step through the public API from pinned host buffers.  Input upload (H2D), compute and token
    # download (D2H) run on three streams with double-buffered device inputs, so step k+1's upload and step k-1's
    # download overlap step k's compute -- every byte still crosses PCIe inside the timed region.
    h2d_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    dev_in = [{kk: torch.empty_like(v, device=dev) for kk, v in host.items()} for _ in range(2)]

    def run_e2e(steps):
        main = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = torch.cuda.Event()
        h2d_stream.wait_stream(main)
        d2h_stream.wait_stream(main)

        def upload(sidx):
            slot = sidx % 2
            with torch.cuda.stream(h2d_stream):
                if sidx >= 2:
                    h2d_stream.wait_event(ev_done[slot])          # the compute that last read this slot
                for kk, v in host.items():
                    dev_in[slot][kk].copy_(v, non_blocking=True)
                ev_in[slot].record(h2d_stream)

        upload(0)
        for sidx in range(steps):
            slot = sidx % 2
            if sidx + 1 < steps:
                upload(sidx + 1)
            main.wait_event(ev_in[slot])
            out = enc.encode_u8(to_batch(dev_in[slot]), generator=gen)
            if enc.token_exchange is not None:
                enc.token_exchange.wait()
            ev_done[slot].record(main)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev_done[slot])
                out_host.copy_(out, non_blocking=True)
                out.record_stream(d2h_stream)
                ev_out.record(d2h_stream)
        main.wait_event(ev_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, whole=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    if args.profile:
        step_resident()
        torch.cuda.synchronize()
        step_resident()
        torch.cuda.synchronize()
        print(json.dumps({"profile_only": True, "launches_per_step": ops.LAUNCHES // 2}))
        return
    for _ in range(max(3, args.warmup)):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ops.LAUNCHES
    ms = timed(step_resident, args.steps)
    launches = ops.LAUNCHES - launches0
    clock_info = clocks.stop() if rank == 0 else None

    run_e2e(2)                                   # warm the pinned paths
    ms_e2e = timed(run_e2e, args.steps, whole=True)

    # ---- roofline leg: CUDA-event time of every GEMM launch of one more step (same stream, same shapes)
    with ops.profile_launches() as prof:
        step_resident()
    gemm_ms = sum(t for (nm, shp, t) in prof.records if nm == "gemm")
    gemm_fl = sum(gemm_flops(shp) for (nm, shp, t) in prof.records if nm == "gemm")
    attn_ms = sum(t for (nm, shp, t) in prof.records if nm == "attention")
    ln_ms = sum(t for (nm, shp, t) in prof.records if nm == "layernorm")
    n_gemm = sum(1 for r in prof.records if r[0] == "gemm")
    by_shape = {}
    for nm, shp, t in prof.records:
        if nm == "gemm":
            e = by_shape.setdefault(shp, [0, 0.0])
            e[0] += 1
            e[1] += t
    # the dominant kernel = the GEMM shape with the largest share of the step (fc1 + GELU of the ViT MLP)
    dom_shape, (dom_n, dom_ms) = max(by_shape.items(), key=lambda kv: kv[1][1])
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved_tf = gemm_flops(dom_shape) / (dom_ms / dom_n * 1e-3) / 1e12
    all_gemm_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    try:                      # DRAM bytes of that launch from the committed `ncu --set full` capture (profiles/)
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f)
        if list(dom_shape) == tr.get("shape"):
            traffic = tr.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    epi_names = {0: "bias", 1: "bias+erf-GELU", 2: "bias (fp32 out)", 3: "bias+fp32 residual", 4: "row-bias (patch embed)",
                 5: "bias+ReLU", 6: "folded LayerNorm+bias", 7: "folded LayerNorm+bias+erf-GELU",
                 8: "bias+fp32 residual+bf16 copy+LayerNorm statistics"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames_total = b * k * world
    line = {
        "metric": METRIC, "value": frames_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": frames_total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "clocks": clock_info,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                     "kernel": f"gemm_bf16_tcgen05_kernel M={dom_shape[0]} N={dom_shape[1]} K={dom_shape[2]} "
                               f"epilogue={epi_names.get(dom_shape[3], dom_shape[3])}",
                     "launches": dom_n, "avg_launch_ms": dom_ms / dom_n, "share_of_step": dom_ms / ms,
                     "algorithmic_flops_per_launch": gemm_flops(dom_shape),
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                     "all_gemm": {"launches": n_gemm, "ms_per_step": gemm_ms, "tflops": all_gemm_tf,
                                  "frac": all_gemm_tf / peak_tf if peak_tf else None},
                     "gemm_shapes": [{"M": k[0], "N": k[1], "K": k[2], "epilogue": epi_names.get(k[3], k[3]), "launches": v[0],
                                      "ms": round(v[1], 3), "tflops": round(gemm_flops(k) * v[0] / (v[1] * 1e-3) / 1e12, 1)}
                                     for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:6]],
                     "attention_ms_per_step": attn_ms, "layernorm_ms_per_step": ln_ms},
        "path_tflops": frames_total / world * GFLOP_PER_FRAME / (ms * 1e-3) / 1e3,
        "path_frac_of_peak": (b * k * GFLOP_PER_FRAME / (ms * 1e-3) / 1e3) / peak_tf if peak_tf else None,
    }
    if world == 1 and not args.no_cpu_baseline:
        fps, cms, cores = cpu_reference(args.cpu_frames, 1, 1, args.vit_layers, args.lof, args.tq)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"1 video x {2 * args.cpu_frames} candidates -> {args.cpu_frames} frames, 1 timed step "
                                          f"after 1 warm-up, oracle port (torch fp32 CPU), {cms:.0f} ms"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
