#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 1080p B-frames/s (encode+decode) of the LHBDC hierarchical GOP-8 codec,
plus hot-path HBM GB/s vs peak.

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (b200vc kernels)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle), rank 0 only

A "step" = one GOP-8 of synthetic 1920x1080 video per GPU (7 B-frames; frames of one hierarchy level are one
batched Model call; anchors uncoded).  One encode+decode = one ``Model.forward(train=False)`` per B-frame
(SURVEY.md 8d).  N > 1 is launched by torchrun, one rank per GPU; GOPs are independent, so ranks share no
data-path collective (weak scaling) -- only the timing max and the per-frame records cross ranks.

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM; ``e2e``: same steps through the public
API from pinned host frames, H2D + D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "1080p B-frames/s encode+decode (LHBDC hierarchical GOP-8)"
UNIT = "B-frames/s"
HOT_KERNELS = ("gdn_f32", "warp_f32", "warp2_lhbdc_f32", "blend_residual_f32", "gauss_cond_f32",
               "entropy_bottleneck_f32", "spynet_level_f32", "spynet_pyramid_f32")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200vc", choices=["b200vc", "reference"])
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--gops-per-step", type=int, default=1)
    ap.add_argument("--conv-tf32", action="store_true",
                    help="let cuDNN use TF32 for the (out-of-scope) convolutions, as torch defaults do")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="wall-clock budget of the reference arm")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_product_model(device):
    import b200vc
    from b200vc import synthetic
    torch.manual_seed(0)
    model = b200vc.Model().eval()
    synthetic.calibrate_(model, 0)
    model.mv_compressor.update(force=True)
    model.residual_compressor.update(force=True)
    return model.to(device)


def build_oracle_model():
    """Same weights as the product arm (state dict copied), torch ops only, CPU."""
    import b200vc
    from b200vc import synthetic
    from oracle import lhbdc as o_lhbdc
    torch.manual_seed(0)
    prod = b200vc.Model().eval()
    synthetic.calibrate_(prod, 0)
    orc = o_lhbdc.Model().eval()
    orc.load_state_dict(prod.state_dict())
    orc.mv_compressor.update(force=True)
    orc.residual_compressor.update(force=True)
    return orc


def cpu_reference_frames_per_s(args, steps, warmup, budget_s):
    """The reference's CPU PyTorch path (oracle restatement: compressai is not installable offline), all host
    threads, on a bounded sample: the level-0 B-frame of a GOP-8 (anchors 0 and 8).  If the full frame would
    blow the time budget the sample shrinks to a top band of the frame and is scaled by its pixel share."""
    from b200vc import synthetic
    from b200vc.lhbdc import reflect_pad64
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = build_oracle_model()
    frames = reflect_pad64(synthetic.make_sequence(9, args.height, args.width, seed=1234))
    full_rows = frames.shape[2]
    n_total = steps + warmup

    def one(rows):
        t0 = time.perf_counter()
        with torch.no_grad():
            orc(frames[0:1, :, :rows], frames[4:5, :, :rows], frames[8:9, :, :rows], train=False)
        return time.perf_counter() - t0

    rows = full_rows
    first = one(rows)
    times = [] if warmup > 0 else [first]
    done = 1
    if first * n_total > budget_s:
        for cand in (full_rows // 2, full_rows // 4, full_rows // 8):
            cand -= cand % 64
            if cand >= 192:
                rows = cand
                if first * (cand / full_rows) * n_total <= budget_s:
                    break
        times, done = [], min(done, warmup)  # the full-size run only counts as a warm-up
    while done < n_total or not times:
        dt = one(rows)
        if done >= warmup:
            times.append(dt)
        done += 1
    share = rows / full_rows
    med = sorted(times)[len(times) // 2]
    value = share / med
    sample = (f"level-0 B-frame of a GOP-8, {rows}x{frames.shape[3]} of the padded {full_rows}x{frames.shape[3]} frame"
              f"{'' if rows == full_rows else f' (scaled by pixel share {share:.3f})'}; median of {len(times)} runs; "
              f"oracle torch path, fp32")
    return value, cores, sample, sum(times) / max(1, len(times)) * 1e3


# --------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, cores, sample, ms = cpu_reference_frames_per_s(args, steps, warmup, args.cpu_budget_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LHBDC hierarchical GOP-8 B-frame encode/decode, synthetic {args.width}x{args.height}, "
                               "random-init calibrated weights", "step": "one B-frame (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------- product arm
def run_product(args):
    import b200vc
    from b200vc import dist as bd
    from b200vc import gop, ops, synthetic
    from b200vc.lhbdc import reflect_pad64

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the b200vc arm needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU path")
    rank, local_rank, world = bd.init()
    if world != args.gpus:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    torch.backends.cudnn.allow_tf32 = bool(args.conv_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True  # as LHBDC/test/testing.py:31

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    G, sched = max(1, args.gops_per_step), gop.LHBDC_GOP8
    h, w = args.height, args.width
    model = build_product_model(device)
    coder = gop.GopCoder(model, sched)

    # two distinct GOP batches per rank, alternated across steps; one batch = G*9 frames = G*224 MB > L2
    n_sets = 2
    host = synthetic.make_sequence(n_sets * G * sched.gop + 1, h, w, seed=1234 + rank)
    idx = [[k * sched.gop + t for t in range(sched.gop + 1)] for k in range(n_sets * G)]
    host_sets = [torch.stack([host[idx[s * G + g]] for g in range(G)], 0).pin_memory() for s in range(n_sets)]
    dev_sets = [reflect_pad64(hs.to(device).flatten(0, 1)).unflatten(0, (G, sched.gop + 1)) for hs in host_sets]
    frames_per_step = G * len(sched.refs)

    def step_resident(i):
        return coder.code(dev_sets[i % n_sets], (h, w))

    def step_e2e(i):
        x = host_sets[i % n_sets].to(device, non_blocking=True)                       # H2D (pinned)
        x = reflect_pad64(x.flatten(0, 1)).unflatten(0, (G, sched.gop + 1))
        bits, sse = coder.code(x, (h, w))
        out = torch.stack([bits, sse]).to("cpu", non_blocking=False)                  # D2H of the step's result
        return out

    for i in range(warmup):
        step_resident(i)
    torch.cuda.synchronize()

    # ---- timed region 1: resident inputs, per-kernel CUDA events recorded live ------------------------
    launches0 = ops.launch_count()
    bd.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        ops.profile_begin()
        if args.ncu_range:
            torch.cuda.profiler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            bits, sse = step_resident(i)
        ev1.record()
        torch.cuda.synchronize()
        if args.ncu_range:
            torch.cuda.profiler.stop()
        prof = ops.profile_end()
        bd.barrier()
    ms_total = bd.max_over_ranks(ev0.elapsed_time(ev1), device)
    launches = ops.launch_count() - launches0
    clock_summary = clocks.summary()

    # ---- timed region 2: end to end from pinned host memory ------------------------------------------
    for i in range(2):
        step_e2e(i)
    bd.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        res = step_e2e(i)
    ev1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    bd.barrier()
    e2e_ms = bd.max_over_ranks(max(ev0.elapsed_time(ev1), wall_ms), device)
    h2d = host_sets[0].numel() * 4
    d2h = res.numel() * 8

    # ---- records: gather per-frame (gop, frame, bits, sse) over ranks; totals in global frame order ----
    rec = []
    bits_c, sse_c = bits.cpu(), sse.cpu()
    for g in range(G):
        for f in sched.order:
            rec.append([float(rank * G + g), float(f), bits_c[g, f].item(), sse_c[g, f].item()])
    table = bd.gather_records(torch.tensor(rec, dtype=torch.float64, device=device))
    tot = bd.totals(table)
    n_frames = table.shape[0]
    bpp = tot[0].item() / (n_frames * h * w)
    psnr = gop.psnr_from_sse(table[:, 3].cpu(), 3 * h * w).mean().item()

    if rank == 0:
        W_ = bd.world_size()
        value = frames_per_step * steps * W_ / (ms_total / 1e3)
        e2e_value = frames_per_step * steps * W_ / (e2e_ms / 1e3)
        peak, peak_src = measured_peak()
        kernels = {}
        for name, r in prof.items():
            gbs = r["bytes"] / (r["ms"] / 1e3) / 1e9 if r["ms"] > 0 else None
            kernels[name] = {"launches": r["launches"], "ms_per_step": r["ms"] / steps,
                             "algorithmic_MB_per_step": r["bytes"] / steps / 1e6, "GBps": gbs,
                             "frac_of_peak": gbs / peak if gbs else None}
            if "by_shape" in r:
                kernels[name]["by_shape"] = {
                    t: {"launches": v["launches"], "us_per_launch": 1e3 * v["ms"] / v["launches"],
                        "GBps": v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else None}
                    for t, v in r["by_shape"].items()}
        hot = {k: v for k, v in kernels.items() if k in HOT_KERNELS}
        dom = max(hot, key=lambda k: hot[k]["ms_per_step"]) if hot else None
        hot_ms = sum(v["ms_per_step"] for v in hot.values())
        hot_bytes = sum(v["algorithmic_MB_per_step"] for v in hot.values()) * 1e6
        roofline = None
        if dom:
            d = kernels[dom]
            roofline = {"bound": "hbm", "kernel": dom, "achieved": d["GBps"], "peak": peak, "unit": "GB/s",
                        "frac": d["frac_of_peak"], "traffic": load_traffic(dom), "traffic_note": load_traffic_note(dom),
                        "peak_source": peak_src,
                        "launches_per_step": d["launches"] / steps,
                        "hot_path_ms_per_step": hot_ms, "hot_path_GBps": hot_bytes / (hot_ms / 1e3) / 1e9 if hot_ms else None,
                        "hot_path_share_of_step": hot_ms / (ms_total / steps)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": W_, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"LHBDC hierarchical GOP-8 B-frame encode/decode, synthetic {w}x{h}, 1 B200 per rank "
                            "(BASELINE.json configs[1])",
                "step": f"{G} GOP-8 per GPU = {frames_per_step} B-frames, levels batched (1/2/4 frames per call)",
                "weights": "random-init (seed 0) + deterministic calibration (b200vc/synthetic.py)",
                "anchors": "uncoded source frames (I-frame codec is outside the B-frame hot path)",
                "conv_math": "cuDNN TF32 allowed (torch default, as the reference runs)" if args.conv_tf32
                             else "cuDNN fp32 (allow_tf32=False): the configuration parity is proven in",
                "l2": "inputs larger than L2: each step streams >= 224 MB of frames + GBs of activations; "
                      "two GOP sets alternate",
                "parallelism": f"gop-sharded x{W_}",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / steps},
            "gpu_launches": launches,
            "clocks": clock_summary,
            "roofline": roofline,
            "kernels": kernels,
            "quality": {"bpp": bpp, "psnr_db": psnr, "frames": n_frames, "total_bits": tot[0].item()},
        }
        if W_ == 1 and not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_reference_frames_per_s(args, 1, 0, 120.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    bd.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return 0


def load_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel from the committed
    `ncu --set full` capture (profiles/traffic.json, profiles/r01_ncu_summary.md), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            entry = json.load(f).get(kernel)
        return entry["bytes_per_launch"] if entry else None
    except Exception:
        return None


def load_traffic_note(kernel):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(kernel)
        return (f"ncu capture at {e['shape']}: {e['bytes_per_launch'] / 1e6:.1f} MB DRAM vs "
                f"{e['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic per launch") if e else None
    except Exception:
        return None


_RESULT_FD = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else any library prints (NCCL's version
    banner, cuDNN / torch warnings) was redirected to stderr by `claim_stdout`."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def claim_stdout():
    """Keep fd 1 for the result line only: duplicate it, then point fd 1 (C and Python level) at stderr."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    a = parse()
    claim_stdout()
    sys.exit(run_reference(a) if a.impl == "reference" else run_product(a))
