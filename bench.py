#!/usr/bin/env python
"""bench.py -- predict_instances throughput of the B200 path (and of the reference CPU path).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle)

A "step" is one StarDist2D.predict_instances() over one synthetic 1024x1024 image
(BASELINE.json configs[1]: default Config2D, n_rays=32, grid (1,1), random-init U-Net).
`value` = instances/s with the normalized, padded input already resident in HBM;
`e2e`   = the same metric through the public API with a HOST image (pinned H2D inside, D2H of the
          label map + polygons inside).  N>1: every rank processes its own images (weak scaling,
          no data-path collective); value = instances of all ranks / max-over-ranks time.
"""
import argparse, json, os, sys, time, threading, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (1024, 1024)
N_RAYS = 32
PROB_THRESH = 0.5          # the reference's default thresholds (base.py:241)
NMS_THRESH = 0.4
REF_CROP = 512             # CPU arms time a 512x512 crop of the same image per step (bounded sample)


class ClockSampler(threading.Thread):
    """samples nvidia-smi SM clocks / throttle reasons while the timed region runs"""
    def __init__(self, gpu_index=0, period=0.2):
        super().__init__(daemon=True)
        self.gpu, self.period, self.samples, self._halt = gpu_index, period, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set(); self.join(timeout=5)
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def conv_flops(config, shape):
    """algorithmic FLOPs of one forward pass (2*pixels*Cin*Cout*k^d per conv layer)"""
    from stardist_b200.models.weights import unet_layers
    sp = np.array(shape, dtype=np.float64)
    fl = 0.0
    for l in unet_layers(config):
        if l['kind'] in ('conv', 'head'):
            fl += 2.0 * np.prod(sp) * l['cin'] * l['cout'] * np.prod(l['k'])
        elif l['kind'] == 'pool':
            sp = sp / np.array(l['pool'])
        elif l['kind'] == 'up':
            sp = sp * np.array(l['pool'])
    return fl


def run_reference(args):
    """the reference's CPU path on the host cores: torch-CPU fp32 U-Net (stand-in for TF-CPU, which is
    not installable here) + the reference's own C++/OpenMP NMS (oracle/_ref) + numpy label painting."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")     # see cpu_baseline(): two OpenMP runtimes share the process
    import torch
    from stardist_b200.models.config import Config2D
    from oracle import pipeline2d
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    import bench_data
    cfg = Config2D(n_rays=N_RAYS)
    w = bench_data.bench_weights_2d(cfg)
    img, _ = bench_data.synthetic_image(SHAPE, seed=0)
    img = np.ascontiguousarray(img[:REF_CROP, :REF_CROP])      # bounded sample of the workload (one quarter of the image)
    pthr = PROB_THRESH
    def step():
        prob, dist, pads = pipeline2d.predict(cfg, w, img)
        pa, da, pts = pipeline2d.candidates(cfg, prob, dist, pads, img.shape, pthr)
        labels, res = pipeline2d.instances(cfg, img.shape, pa, da, pts, NMS_THRESH)
        return len(res['prob'])
    for _ in range(args.warmup): step()
    t0 = time.perf_counter(); n_inst = 0
    for _ in range(args.steps): n_inst += step()
    dt = time.perf_counter() - t0
    v = n_inst / dt
    print(json.dumps({
        "impl": "reference", "metric": "instances/sec (predict_instances end-to-end)", "value": v, "unit": "instances/s",
        "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "StarDist2D predict_instances, 1024x1024, n_rays=32, seeded synthetic U-Net weights (configs[1])",
                   "prob_thresh": PROB_THRESH, "weights": "seeded Glorot body + fitted heads (bench_data.py)", "nms_thresh": NMS_THRESH},
        "cpu_baseline": {"value": v, "unit": "instances/s", "cores": cores, "kind": "reference",
                         "sample": "%d steps of one %dx%d crop of the 1024^2 image; NMS = reference C++/OpenMP (oracle/_ref), U-Net = torch-CPU fp32 stand-in for TF-CPU, labels = numpy restatement" % (args.steps, REF_CROP, REF_CROP)},
        "e2e": {"value": v, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from stardist_b200 import Config2D, StarDist2D, _lib
    import bench_data
    cfg = Config2D(n_rays=N_RAYS)
    model = StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
    img, _ = bench_data.synthetic_image(SHAPE, seed=rank)
    pthr = PROB_THRESH
    x_dev = torch.from_numpy(img[None, ..., None]).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(max(3, args.warmup)):
        model.predict_instances_device(x_dev, SHAPE, prob_thresh=pthr, nms_thresh=NMS_THRESH)
    sampler = ClockSampler(local); sampler.start()
    _lib.launch_count(reset=True)
    _lib.profile_enable(True)
    barrier()
    n_inst = 0; dev_ms = 0.0; stage = {}
    for _ in range(args.steps):
        flush.zero_()
        model._events = []
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        labels, res = model.predict_instances_device(x_dev, SHAPE, prob_thresh=pthr, nms_thresh=NMS_THRESH)
        e1.record(); torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1); n_inst += len(res['prob'])
        ev = dict(model._events); model._events = None
        for a, b, k in (("net_begin", "net_end", "unet"), ("net_end", "cand_end", "threshold_sort_gather"),
                        ("cand_end", "nms_end", "nms"), ("nms_end", "label_end", "coord_label")):
            if a in ev and b in ev: stage[k] = stage.get(k, 0.0) + ev[a].elapsed_time(ev[b])
    launches = _lib.launch_count()
    prof_clip = _lib.profile_get("nms2d_clip"); prof_conv = _lib.profile_get("conv_tc")
    prof_fast = _lib.profile_get("nms2d_fast")
    prof_nms = {k: _lib.profile_get("nms2d_" + k) for k in ("frontier", "pairs", "fast", "clip")}
    _lib.profile_enable(False)
    barrier()
    t = torch.tensor([dev_ms, float(n_inst)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms_max, n_total = float(tmax[0]), float(tsum[1])
    else:
        dev_ms_max, n_total = dev_ms, float(n_inst)
    value = n_total / (dev_ms_max / 1000.0)

    # ---- end-to-end through the public API, host image in / host results out
    model._stats = {}
    for _ in range(2): model.predict_instances(img, prob_thresh=pthr, nms_thresh=NMS_THRESH)
    model._stats = {}
    barrier(); t0 = time.perf_counter(); n_e2e = 0
    for _ in range(args.steps):
        flush.zero_()
        labels, res = model.predict_instances(img, prob_thresh=pthr, nms_thresh=NMS_THRESH)
        n_e2e += len(res['prob'])
    torch.cuda.synchronize(); e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(n_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX); b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, n_e2e = float(a[0]), float(b[1])
    clocks = sampler.stop()
    h2d = model._stats.get('h2d_bytes', 0) // max(1, args.steps); d2h = model._stats.get('d2h_bytes', 0) // max(1, args.steps)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fl = conv_flops(cfg, SHAPE)
        unet_ms = stage.get("unet", 0.0) / args.steps
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_bw = peaks.get("hbm_gbs", 6650.0)
        src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        # Kernel rooflines, all timed live with CUDA events on the launching stream (sdb_profile_*):
        #  conv_tc   tcgen05 3x3 convolutions of the U-Net: tensor bound; algorithmic flop = 2*9*Cin*Cout*H*W per layer
        #            (each is issued as 3 fp16 MMAs -- hi*hi + lo*hi + hi*lo -- to carry fp32 accuracy)
        #  nms2d_fast  closed-form overlap integral, one warp per candidate pair: integer/fp64 ALU work on L1/L2-resident
        #            rows; algorithmic bytes per pair = two vertex rows (2*R*8) + two suffix rows (2*R*8) + 28
        #  nms2d_clip  exact Clipper-equivalent sweep for the pairs the filter leaves open: serial per-thread latency
        bytes_fast = 4 * N_RAYS * 8 + 28
        bytes_clip = 2 * N_RAYS * 8 + 28
        def rl(prof, bound, kernel, units_to_alg, peak, unit, extra=None):
            ms = prof["ms"]
            ach = units_to_alg * prof["units"] / (ms / 1e3) / (1e12 if unit == "TFLOP/s" else 1e9) if ms > 0 else 0.0
            d = {"bound": bound, "kernel": kernel, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak if peak else None,
                 "traffic": None, "ms_per_step": ms / args.steps, "launches_per_step": prof["launches"] / args.steps}
            if extra: d.update(extra)
            return d
        rls = {
            "conv_tc": rl(prof_conv, "tensor", "k_conv_tc (tcgen05 3x3 conv + 1x1 heads, all launches of the forward pass)", 1.0, peak_tf, "TFLOP/s",
                          {"algorithmic_flop_per_step": prof_conv["units"] / args.steps, "issued_flop_factor": 3, "unet_forward_ms": unet_ms,
                           "unet_algorithmic_tflops": fl / (unet_ms / 1e3) / 1e12 if unet_ms > 0 else None, "peak_source": src + " bf16_tflops_sustained"}),
            "nms2d_fast": rl(prof_fast, "hbm", "k_fast (NMS pair pre-filter, closed-form overlap integral; ALU bound on cache-resident rows)", bytes_fast, peak_bw, "GB/s",
                             {"pairs_per_step": prof_fast["units"] / args.steps, "algorithmic_bytes_per_pair": bytes_fast, "peak_source": src + " hbm_gbs"}),
            "nms2d_clip": rl(prof_clip, "hbm", "k_clip<32> (exact Clipper-equivalent sweep of the pairs the pre-filter leaves open; latency bound)", bytes_clip, peak_bw, "GB/s",
                             {"pairs_per_step": prof_clip["units"] / args.steps, "algorithmic_bytes_per_pair": bytes_clip, "peak_source": src + " hbm_gbs"}),
        }
        try:
            for k, v in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).items():
                if k in rls and rls[k]["launches_per_step"]:
                    rls[k]["traffic"] = v["dram_bytes_per_step"] / rls[k]["launches_per_step"]   # per launch, like achieved
                    rls[k]["traffic_source"] = v["source"]
        except Exception:
            pass
        dominant = max(rls, key=lambda k: rls[k]["ms_per_step"])
        out = {
            "metric": "instances/sec (predict_instances end-to-end)", "value": value, "unit": "instances/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "StarDist2D predict_instances, 1024x1024, n_rays=32, seeded synthetic U-Net weights (configs[1])",
                       "prob_thresh": PROB_THRESH, "weights": "seeded Glorot body + fitted heads (bench_data.py)", "nms_thresh": NMS_THRESH, "l2": "flushed (256 MiB write) between steps",
                       "instances_per_image": n_total / (args.steps * world),
                       "stages_ms": {k: v / args.steps for k, v in stage.items()},
                       "nms_kernels_ms": {k: v["ms"] / args.steps for k, v in prof_nms.items()}},
            "clocks": clocks,
            "e2e": {"value": n_e2e / e2e_s, "unit": "instances/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": dict(rls[dominant], name=dominant),
            "roofline_other": {k: v for k, v in rls.items() if k != dominant},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline(cfg, model.weights, img, pthr)
            except Exception as e:      # the baseline is a report, never a reason to lose the measurement
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cfg, weights, img, pthr):
    """bounded CPU sample of the same workload, timed in a fresh process (`--impl reference`) so that the
    OpenMP runtime of the reference extension starts with OMP_WAIT_POLICY=PASSIVE (torch and the reference
    C++ bring two OpenMP runtimes into one process; with the default spin-waiting they starve each other
    on many-core hosts)"""
    env = dict(os.environ); env.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"): env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)["cpu_baseline"]


if __name__ == "__main__":
    main()
