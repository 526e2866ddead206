#!/usr/bin/env python
"""Benchmark of the deformable-GAN training step (BASELINE.json metric: training images/sec at 256x256,
warp_skip=mask; warp-kernel HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

One "step" = one main.py iteration = dis_update + gen_update (training_ratio = 1, src_deformable/main.py:78-108)
on one synthetic batch (oracle/synth.py).  `value` = whole-job images/s with the batch resident in
HBM; `e2e.value` = the same through DeformablePose_GAN's public methods with the batch in pinned HOST memory
(H2D copies of input/target/warps/masks and the D2H loss read inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H = W = 256
P = 18
KPARTS = 10
PER_GPU_BATCH = 8
# SURVEY 8d work model (256^2, P=18): necessary conv FLOPs per image per iteration and warp bytes per image
CONV_GFLOP_PER_IMG = 582.3
WARP_FWD_BYTES_PER_IMG = 66.40e6


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (set) of a kernel, from the committed `ncu --set full`
    capture of the CURRENT kernel (profiles/traffic.json, written by tools/ncu_extract.py from the capture named there);
    None when no capture of the current kernel is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get(key)
    return e.get("bytes") if isinstance(e, dict) and e.get("batch") == PER_GPU_BATCH else None


def make_opt(N, content="block1_conv2", area=5, l1_w=0.01):
    return argparse.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4,
                              gen_type="baseline", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                              content_loss_layer=content, nn_loss_area_size=area, gan_penalty_weight=1.0,
                              l1_penalty_weight=l1_w)


def measured_peaks():
    """(HBM GB/s, bf16 TFLOP/s sustained, bf16 TFLOP/s burst, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------- reference arm (CPU)
class _RealReference:
    """The UNMODIFIED reference (src_deformable DeformablePose_GAN imported from baseline/_ref, or /root/reference in the
    build container) with its two file loads patched to seeded in-memory objects (oracle/ref_import.py).  device='cpu':
    its hard-coded .cuda() calls become the identity, i.e. the reference's own torch CPU path."""
    kind = "reference"

    def __init__(self, N, device="cpu"):
        import torchvision
        from oracle import ref_import, synth
        self.ri, self.N, self.device = ref_import, N, device
        self.opt = make_opt(N)
        vgg = torchvision.models.vgg19(weights=None)
        vw, vb = synth.vgg_conv1_1(0)
        with torch.no_grad():
            vgg.features[0].weight.copy_(vw)
            vgg.features[0].bias.copy_(vb)
        dsd = synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), 1)
        with self._ctx():
            self.model = ref_import.make_reference_gan(self.opt, dsd, vgg)
            self.model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), 0))
        self.batches = [synth.make_batch(N, H, W, P, seed=s) for s in (0, 1, 2)]
        if device != "cpu":
            self.batches = [{k: v.to(device) for k, v in b.items()} for b in self.batches]

    def _ctx(self):
        import contextlib
        return self.ri.cpu_only() if self.device == "cpu" else contextlib.nullcontext()

    def iteration(self):
        b, r, b2 = self.batches
        od = vars(self.opt)
        t0 = time.perf_counter()
        with self._ctx():
            self.model.dis_update(b["input"], b["target"], {"warps": b["warps"].float(), "masks": b["masks"]},
                                  r["input"], r["target"], od)
            self.model.gen_update(b2["input"], b2["target"], {"warps": b2["warps"].float(), "masks": b2["masks"]}, od)
        if self.device != "cpu":
            torch.cuda.synchronize()
        return time.perf_counter() - t0


class _PortReference:
    """Fallback when no copy of the reference is available: oracle/restate.py (the reference's torch CPU op sequence incl.
    its wasted generator backward in dis_update)."""
    kind = "port"

    def __init__(self, N, device="cpu"):
        from oracle import restate, synth
        vw, vb = synth.vgg_conv1_1(0)
        self.N = N
        self.model = restate.OracleGAN(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), 0),
                                       synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), 1), vw, vb, (H, W), P, N,
                                       faithful_waste=True)
        self.batches = [synth.make_batch(N, H, W, P, seed=s) for s in (0, 1, 2)]

    def iteration(self):
        from oracle import synth
        b, r, b2 = self.batches
        N = self.N
        t0 = time.perf_counter()
        self.model.dis_update(b["input"], b["target"], b["warps"], b["masks"], r["input"], r["target"], 1.0,
                              synth.dropout_masks(N, 512, 3, seed=0))
        self.model.gen_update(b2["input"], b2["target"], b2["warps"], b2["masks"], 1.0, 0.01, synth.dropout_masks(N, 512, 3, seed=1))
        return time.perf_counter() - t0


def make_cpu_reference(N):
    from oracle import fetch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    if fetch_ref.root("src_deformable") is not None:
        return _RealReference(N, "cpu")
    return _PortReference(N)


def cpu_reference_iteration(ref):
    return ref.iteration()


def reference_on_gpu(dev, N, iters=3):
    """Informative column: the unmodified reference on the SAME B200 through its stock ATen/cuDNN path
    (torch defaults: cudnn.allow_tf32 = True, i.e. TF32 convs like ours).  None when baseline/_ref is absent."""
    from oracle import fetch_ref
    if fetch_ref.root("src_deformable") is None:
        return None
    import contextlib
    import io
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ref = _RealReference(N, dev)
            ref.iteration()
            ref.iteration()
            torch.cuda.synchronize()
            t = sum(ref.iteration() for _ in range(iters)) / iters
        out = {"value": N / t, "unit": "img/s", "ms_per_step": 1e3 * t, "batch": N,
               "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
               "what": "unmodified reference DeformablePose_GAN.dis_update + gen_update on this GPU (ATen/cuDNN), inputs resident"}
        del ref
        torch.cuda.empty_cache()
        return out
    except Exception as e:      # the column is informative: never fail the bench on it
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: the UNMODIFIED reference modules from
    baseline/_ref (kind "reference"; oracle/restate.py, kind "port", only where no copy exists).  Each step is a bounded
    sample: one iteration at N=2 (the smallest batch the reference supports, models/networks.py:169) of the 256x256
    workload."""
    rank, world, _ = dist_info()
    if rank != 0:
        return
    N = 2
    ref = make_cpu_reference(N)
    # Bounded run: one CPU iteration takes seconds, so warm-up + timed iterations are capped to ~4 minutes of wall time
    # (at least one of each); the line reports how many timed iterations actually ran.
    budget_s = 240.0
    t_first = cpu_reference_iteration(ref)
    fit = max(int(budget_s / max(t_first, 1e-3)), 2)
    warm = max(min(args.warmup, fit // 4) - 1, 0)
    for _ in range(warm):
        cpu_reference_iteration(ref)
    steps = max(min(args.steps, fit - 1 - warm), 1)
    times = [cpu_reference_iteration(ref) for _ in range(steps)]
    total = sum(times)
    v = N * steps / total
    args.steps = steps
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "training images/sec at 256x256 warp_skip=mask", "value": v, "unit": "img/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "src_deformable warp_skip=mask, fasion 256x256, 18 kpts (BASELINE configs[1])",
                       "per_step_sample": "1 iteration (dis_update+gen_update) at batch 2"},
            "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": ref.kind,
                             "sample": "%d iterations at batch 2, 256x256, torch CPU fp32, %d threads" % (args.steps, cores)},
            "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- our arm (B200)
def run_ours(args):
    import torch.distributed as dist
    rank, world, local = dist_info()
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device -- there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Everything below until the result line may write to fd 1 from native code (NCCL prints its version banner to
    # stdout): park stdout on stderr so that the JSON line is the only thing on stdout.
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import _lib, kernels as K
    from pose_transfer_b200.models import pose_gan
    from oracle import synth

    N = args.batch
    opt = make_opt(N)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the constructor prints like the reference's; keep stdout = the JSON line
        model = pose_gan.DeformablePose_GAN(opt).cuda()
    model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), 0))
    model.disc.load_state_dict(synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), 1))
    vw, vb = synth.vgg_conv1_1(0)
    with torch.no_grad():
        model.content_model.features[0].weight.copy_(vw)
        model.content_model.features[0].bias.copy_(vb)
    od = vars(opt)

    # three distinct synthetic batches per step (D-fake, D-real, G), as in main.py:81-82,105; rank-dependent seeds
    host = [synth.make_batch(N, H, W, P, seed=100 * rank + s) for s in range(3)]
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    resident = [dict(b, warps=b["warps"].float()) for b in resident]
    h2d_bytes = sum(pinned[i][k].numel() * pinned[i][k].element_size() for i in (0, 2) for k in ("input", "target", "warps", "masks"))
    h2d_bytes += sum(pinned[1][k].numel() * pinned[1][k].element_size() for k in ("input", "target"))
    d2h_bytes = 2 * 4 * 4   # two 4-float loss buffers per step

    def step_resident():
        b, r, b2 = resident
        model.dis_update(b["input"], b["target"], {"warps": b["warps"], "masks": b["masks"]}, r["input"], r["target"], od)
        model.gen_update(b2["input"], b2["target"], {"warps": b2["warps"], "masks": b2["masks"]}, od)

    # End-to-end step: what main.py does per iteration (main.py:81-86,105-107) -- host batch -> device -> update ->
    # python floats -- with the loader-side copies issued on a side stream so that the H2D transfer of the NEXT
    # call's batch overlaps the current update (pinned memory + non_blocking, as a DataLoader(pin_memory=True) would).
    copy_stream = torch.cuda.Stream(device=dev)

    def upload(batch, keys):
        out = {}
        with torch.cuda.stream(copy_stream):
            for k in keys:
                t = batch[k].to(dev, non_blocking=True)
                out[k] = t.float() if k == "warps" else t
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return out, ev

    def ready(pair):
        d, ev = pair
        torch.cuda.current_stream().wait_event(ev)
        for t in d.values():
            t.record_stream(torch.cuda.current_stream())
        return d

    pending = {}

    def step_e2e():
        b, r, b2 = pinned
        if not pending:
            pending["b"] = upload(b, ("input", "target", "warps", "masks"))
            pending["r"] = upload(r, ("input", "target"))
        gb, gr = ready(pending.pop("b")), ready(pending.pop("r"))
        pending["b2"] = upload(b2, ("input", "target", "warps", "masks"))      # overlaps dis_update
        model.dis_update(gb["input"], gb["target"], {"warps": gb["warps"], "masks": gb["masks"]}, gr["input"], gr["target"], od)
        g2 = ready(pending.pop("b2"))
        pending["b"] = upload(b, ("input", "target", "warps", "masks"))        # next step's batches overlap gen_update
        pending["r"] = upload(r, ("input", "target"))
        model.gen_update(g2["input"], g2["target"], {"warps": g2["warps"], "masks": g2["masks"]}, od)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- end to end through the DEVICE data path: the batch crosses PCIe as images + key-points + 80
    # transform coefficients per sample; heat-maps and body-part masks are generated in HBM (csrc/pose_data.cu).  The
    # masks are then the reference's pose-derived masks (pose_transform.py:143-183) instead of SURVEY 8d's random
    # rectangles, so this number is reported beside `e2e`, not instead of it.
    from pose_transfer_b200.datasets.device_pipeline import DevicePoseBatcher
    batcher = DevicePoseBatcher((H, W), P, dev)
    kp_sets = [(synth.make_keypoints(N, H, W, P, seed=100 * rank + 2 * s), synth.make_keypoints(N, H, W, P, seed=100 * rank + 2 * s + 1))
               for s in range(3)]
    dev_host = []
    for (kf, kt), hb in zip(kp_sets, host):
        dev_host.append({"img_from": hb["input"][:, :3].contiguous().pin_memory(), "img_to": hb["target"].pin_memory(),
                         "kf": kf.to(torch.int32).pin_memory(), "kt": kt.to(torch.int32).pin_memory(),
                         "warps": torch.from_numpy(batcher.warps_on_host(kf.numpy(), kt.numpy(), P)).pin_memory()})
    dev_h2d_bytes = sum(t.numel() * t.element_size() for i, b in enumerate(dev_host) for k, t in b.items() if not (i == 1 and k == "warps"))

    def build(i, need_masks):
        with torch.cuda.stream(copy_stream):
            b = dev_host[i]
            out = batcher(b["img_from"], b["img_to"], b["kf"], b["kt"], warps=b["warps"], need_masks=need_masks, checked=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return out, ev

    pending_dev = {}

    def step_e2e_device():
        if not pending_dev:
            pending_dev["b"], pending_dev["r"] = build(0, True), build(1, False)
        gb, gr = ready(pending_dev.pop("b")), ready(pending_dev.pop("r"))
        pending_dev["b2"] = build(2, True)                                       # overlaps dis_update
        model.dis_update(gb["input"], gb["target"], {"warps": gb["warps"], "masks": gb["masks"]}, gr["input"], gr["target"], od)
        g2 = ready(pending_dev.pop("b2"))
        pending_dev["b"], pending_dev["r"] = build(0, True), build(1, False)     # next step's batches overlap gen_update
        model.gen_update(g2["input"], g2["target"], {"warps": g2["warps"], "masks": g2["masks"]}, od)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    if args.diag:
        # host-bound or GPU-bound?  Time the host spends blocked in the two loss read-backs of a step: ~0 => the GPU waits
        # for the host's launches; large => the host runs ahead and only the read-back latency is exposed.
        # (the read-back waits on an event recorded right behind the loss kernels: models/pose_gan.py::_loss_readback_end)
        wait = [0.0]
        orig = model._loss_readback_end

        def timed_readback(token, loss):
            t0 = time.perf_counter()
            r = orig(token, loss)
            wait[0] += time.perf_counter() - t0
            return r
        model._loss_readback_end = timed_readback
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            step_resident()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        del model._loss_readback_end
        print("diag: %.2f ms/step wall; host: %.2f ms/step blocked in loss read-backs, %.2f ms/step enqueueing, %.2f ms/step of GPU "
              "work still queued when the host finished" % (100 * total, 100 * wait[0], 100 * (t_enq - wait[0]), 100 * (total - t_enq)),
              file=sys.stderr)
    if args.ncu_step:
        # for `ncu --profile-from-start off`: exactly one resident step inside the profiler range, nothing else
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    step_e2e_device()
    step_e2e_device()
    ms_e2e_dev = timed(step_e2e_device, args.steps)

    # per-kernel-family device times (CUDA events on the launching stream) over a few more steps
    prof_steps = min(args.steps, 3)
    barrier()
    K.PROFILE_DETAIL = bool(args.layers)
    from pose_transfer_b200 import engine as _engine
    _engine.STREAMS = False           # per-kernel durations must be exclusive: no concurrent side-stream kernels
    step_resident()
    K.profile_start()
    for _ in range(prof_steps):
        step_resident()
    prof = K.profile_stop()
    _engine.STREAMS = True
    if args.layers and rank == 0:
        rows = []
        for k, (n, t) in prof.items():
            if "|" in k:
                kind, geo, flop = k.split("|")
                rows.append((t / n, kind, geo, float(flop), n // prof_steps))
        rows.sort(reverse=True)
        with open(args.layers, "w") as f:
            for t, kind, geo, flop, n in rows:
                f.write("%-13s %-40s x%d  %8.3f ms  %7.1f TFLOP/s\n" % (kind, geo, n, t, flop / t / 1e9))
        prof = {k: v for k, v in prof.items() if "|" not in k}
    hbm_peak, tf_peak, tf_burst, peak_src = measured_peaks()
    pbytes = dict(K.PROFILE_BYTES)
    wl, wms = prof.get("warp_forward", (0, 0.0))
    # 2 generator forwards per step (dis_update + gen_update), one launch set (the 4 warped levels) each
    warp_bytes_per_launch_set = WARP_FWD_BYTES_PER_IMG * N
    warp_sets = wl            # one fused launch per generator forward (2 per step)
    warp_gbs = warp_bytes_per_launch_set * warp_sets / (wms * 1e-3) / 1e9 if wms > 0 else 0.0

    def hbm_roofline(family, kernel):
        n, t = prof.get(family, (0, 0.0))
        b = pbytes.get(family, 0)
        gbs = b / (t * 1e-3) / 1e9 if t > 0 else 0.0
        return {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "algorithmic_bytes_per_step": b / prof_steps, "ms_per_step": t / prof_steps, "launches_per_step": n // prof_steps,
                "peak_source": peak_src}
    conv_ms = sum(prof.get(k, (0, 0.0))[1] for k in ("conv_forward", "conv_wgrad")) / prof_steps
    conv_tflops = CONV_GFLOP_PER_IMG * N / (conv_ms * 1e-3) / 1e3 if conv_ms > 0 else 0.0

    value = world * N * args.steps / (ms * 1e-3)
    e2e_value = world * N * args.steps / (ms_e2e * 1e-3)
    line = {"metric": "training images/sec at 256x256 warp_skip=mask", "value": value, "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 operands / f32 accumulate on the tensor-core convs (cuDNN's default for the reference too); f32 elsewhere",
            "data": "synthetic",
            "config": {"workload": "src_deformable warp_skip=mask, fasion 256x256, 18 kpts, batch %d/GPU (BASELINE configs[1])" % N,
                       "global_batch": N * world, "parallelism": "dp%d" % world, "step": "dis_update + gen_update",
                       "content_loss_layer": "block1_conv2", "nn_loss_area_size": 5,
                       "l2": "working set per step (>10 GB of activations) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "e2e_device_data_path": {"value": world * N * args.steps / (ms_e2e_dev * 1e-3), "unit": "img/s",
                                     "h2d_bytes_per_step": dev_h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                                     "ms_per_step": ms_e2e_dev / args.steps,
                                     "what": "same step fed by DevicePoseBatcher: images + key-points + affine coefficients over PCIe, "
                                             "pose heat-maps and body-part masks generated on the GPU (SURVEY 8f-2)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "warp_forward_tiles_kernel (ONE launch = the 4 warped skip levels of one generator forward)", "achieved": warp_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": warp_gbs / hbm_peak,
                         "traffic": measured_traffic("warp_forward_launch_set") if N == PER_GPU_BATCH else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch_set": warp_bytes_per_launch_set,
                         "ms_per_launch_set": wms / warp_sets if warp_sets else None},
            "conv_roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": tf_burst / 2, "unit": "TFLOP/s",
                              "frac": conv_tflops / (tf_burst / 2), "frac_of_sustained": conv_tflops / (tf_peak / 2),
                              "peak_source": peak_src + " bf16 burst / 2 (tf32); frac_of_sustained uses bf16 sustained / 2",
                              "algorithmic_gflop_per_img": CONV_GFLOP_PER_IMG, "conv_ms_per_step": conv_ms},
            "warp_backward_roofline": hbm_roofline("warp_backward", "warp_backward kernels (4 levels, gen_update only)"),
            "gn_roofline": hbm_roofline("gn", "gn_apply / gn_bwd_reduce / gn_bwd_apply / gn_stats (all launches of the step)"),
            "adam_roofline": hbm_roofline("adam", "adam_kernel (generator buckets + discriminator)"),
            "kernel_ms_per_step": {k: v[1] / prof_steps for k, v in sorted(prof.items())},
            "kernel_timing_note": "per-kernel CUDA-event times (roofline, conv_roofline, kernel_ms_per_step) come from extra steps run "
                                  "with the side-stream overlap switched off; value / e2e are measured with it on"}
    sys.stdout.flush()
    os.dup2(saved_stdout_fd, 1)
    os.close(saved_stdout_fd)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            torch.cuda.synchronize()
            line["reference_cudnn_on_this_gpu"] = reference_on_gpu(dev, N)
            cref = make_cpu_reference(2)
            t = cpu_reference_iteration(cref)
            line["cpu_baseline"] = {"value": 2 / t, "unit": "img/s", "cores": torch.get_num_threads(), "kind": cref.kind,
                                    "sample": "1 iteration (dis_update+gen_update) at batch 2, 256x256, torch CPU fp32"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (BASELINE configs[1]: 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", default="", help="write a per-conv-geometry timing table (CUDA events) to this file")
    ap.add_argument("--diag", action="store_true", help="print host-vs-GPU-bound diagnostics to stderr")
    ap.add_argument("--ncu-step", action="store_true", help="profiling aid: warm up, then run ONE step inside "
                    "cudaProfilerStart/Stop and exit (use with ncu --profile-from-start off); prints no result line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
