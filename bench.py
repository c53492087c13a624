#!/usr/bin/env python
"""bench.py — Gaussians->voxels throughput of the splat hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one forward pass of the splat op (pack + supertile lists + render; 3 kernel launches)
over one synthetic sample of BASELINE.json configs[1]: `nuscenes_gs25600_solid.py` shape —
25 600 Gaussians (+ the "empty" Gaussian) into the 200x200x16 grid, 18 classes, batch 1 per GPU.
Prints ONE JSON line (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".

* value        whole-job Gaussians/s with inputs resident in HBM, CUDA-event timed, max over ranks
* e2e          the same metric through the public module (`local_aggregate.LocalAggregator`) with
               HOST (pinned) inputs: H2D of every input + forward (logits + fused arg-max) + D2H of
               the occupancy prediction inside the timed region
* roofline     the tile render kernel timed alone (events recorded around it inside the C ABI),
               algorithmic bytes 112*G + 84*N   over the measured HBM copy peak
* cpu_baseline the oracle's C/OpenMP port of the reference algorithm on the host cores (rank 0)

`--impl reference` times that CPU port as the reference arm (the reference has no CPU
implementation of the splat; its CUDA op cannot run without a GPU build of torch extensions —
when oracle/_ref was shipped, its timing on this GPU is reported as `ref_cuda_op` in the main line).
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

WORKLOAD = "gs25600_solid"
METRIC = "gaussians_to_voxels_per_sec"
UNIT = "Gaussians/s"
G_COUNTED = 25600                 # learned Gaussians per sample (the empty one is overhead)
N_SETS = 6                        # rotating input/output sets so the working set exceeds L2


def _algorithmic_bytes(G, N, C=18):
    return (3 + 6 + 1 + C) * 4 * G + 12 * N + 4 * C * N      # 112*G + 84*N for C = 18


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz = None

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=10)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to the GPU, so that the pinned host
    buffers of the e2e leg are first-touched on the GPU's NUMA node (a remote node can cut the H2D
    rate several-fold).  Best effort: returns a short description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {index}"
    except Exception:
        pass
    return None


def _measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def _cpu_port_once(kw, inp):
    """One forward of the CPU port (oracle C code, all OpenMP threads).  Returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    a = {k: v[0].numpy() for k, v in inp.items()}
    dims = (kw["H"], kw["W"], kw["D"])
    t0 = time.perf_counter()
    pi, mi, radii = oracle.host_prep(a["pts"], a["means"], a["scales"], kw["pc_min"], kw["grid_size"],
                                     kw["scale_multiplier"])
    cov6 = oracle.cov6_from_3x3(a["cov"])
    out, pairs = oracle.splat_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims, "f32")
    sec = time.perf_counter() - t0
    _LAST_PORT["logits"], _LAST_PORT["pairs"] = out, int(pairs)
    return sec


_LAST_PORT = {}   # logits / (voxel, Gaussian) pair count of the most recent CPU-port forward (checker side figures)


def run_reference_arm(args):
    """Reference arm: the reference algorithm's CPU port on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from gaussianformer_b200.synthetic import make_splat_inputs
    kw, inp, _ = make_splat_inputs(WORKLOAD, seed=0, perturb=False)
    for _ in range(max(1, min(args.warmup, 2))):
        _cpu_port_once(kw, inp)
    steps = max(1, min(args.steps, 10))
    t = [_cpu_port_once(kw, inp) for _ in range(steps)]
    sec = sum(t) / len(t)
    value = G_COUNTED / sec
    cores = oracle.num_threads()
    sample = f"{steps} full forward passes of the {WORKLOAD} sample (N=640000 points, G=25601)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: 25600 Gaussians (+1 empty) -> 200x200x16x18, batch 1"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (bwd, prob, DAF, ref op)")
    ap.add_argument("--perturb", action="store_true", help="jitter every point inside its voxel (LoadOccupancySurroundOcc(perturb=True)); "
                    "the shipped configs and BASELINE.md use exact voxel centres")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from gaussianformer_b200 import _lib
    from gaussianformer_b200.splat import _make_desc, _ptr, LocalAggregator
    from gaussianformer_b200.synthetic import make_splat_inputs
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local_rank)
    numa = _bind_to_gpu_numa_node(local_rank)   # pinned staging buffers must live next to the GPU
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    W = max(args.warmup, 3)
    K = args.steps

    # ---- resident inputs: N_SETS different samples (seeded), rotated so no step re-reads L2 -------
    PERTURB = args.perturb
    kw, inp0, _ = make_splat_inputs(WORKLOAD, seed=rank * 100, perturb=PERTURB)
    sets = []
    for i in range(N_SETS):
        _, inp, _ = (kw, inp0, None) if i == 0 else make_splat_inputs(WORKLOAD, seed=rank * 100 + i, perturb=PERTURB)
        t = {k: v[0].to(dev).contiguous() for k, v in inp.items()}
        t["cov"] = t["cov"].reshape(-1, 9).contiguous()
        sets.append(t)
    G, N, C = sets[0]["means"].shape[0], sets[0]["pts"].shape[0], 18
    desc = _make_desc(G, N, C, kw["H"], kw["W"], kw["D"], _lib.GF_SPLAT_BASE, 1, 9, kw["pc_min"], kw["grid_size"],
                      float(kw["scale_multiplier"]), 0)
    ws_bytes = L.gf_splat_forward_workspace_bytes(ctypes.byref(desc))
    outs = [torch.empty((N, C), device=dev) for _ in range(N_SETS)]
    wss = [torch.empty(ws_bytes, dtype=torch.uint8, device=dev) for _ in range(N_SETS)]
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    calls = []
    for t, o, w in zip(sets, outs, wss):
        ins = _lib.SplatInputs(_ptr(t["pts"]), None, _ptr(t["means"]), None, _ptr(t["opa"]), _ptr(t["sem"]),
                               _ptr(t["cov"]), None, _ptr(t["scales"]))
        ou = _lib.SplatOutputs(_ptr(o), None, None, None, None)
        calls.append((ins, ou, w))
    loss_buf = torch.zeros(1, device=dev)
    pending = []

    def step(i):
        ins, ou, w = calls[i % N_SETS]
        _lib.check(L.gf_splat_forward(ctypes.byref(desc), ctypes.byref(ins), ctypes.byref(ou), _ptr(w), ws_bytes, sptr))
        if world > 1:
            # north star: NCCL only for the (scalar) loss all-reduce.  It is issued asynchronously on
            # NCCL's stream so that it overlaps the next sample's kernels; every handle is waited for
            # before the timed region closes.
            pending.append(dist.all_reduce(loss_buf, async_op=True))
            if len(pending) > 8:
                pending.pop(0).wait()

    def timed_region(nsteps, fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(nsteps):
            fn(i)
        while pending:
            pending.pop(0).wait()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    for i in range(W):
        step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms = timed_region(K, step)
    clocks = sampler.stop()
    ms_per_step = total_ms / K
    value = world * G_COUNTED / (ms_per_step * 1e-3)

    # ---- roofline: the render kernel alone, events recorded around it inside the C ABI -----------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(K, 50))]
    for i, (a, b) in enumerate(evs):
        a.record(stream); b.record(stream)          # materialise the underlying cudaEvent_t
        L.gf_debug_set_render_events(ctypes.c_void_p(a.cuda_event), ctypes.c_void_p(b.cuda_event))
        step(i)
    L.gf_debug_set_render_events(None, None)
    torch.cuda.synchronize(dev)
    render_ms = sorted(a.elapsed_time(b) for a, b in evs)
    render_ms = sum(render_ms) / len(render_ms)
    peak, peak_kind = _measured_peak()
    alg = _algorithmic_bytes(G, N, C)
    achieved = alg / (render_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "render_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "render_tc_kernel<18,false>" if os.environ.get("GF_B200_RENDER", "")[:1] == "t" else "render_tile_kernel<18,false>", "kernel_ms": render_ms,
                "algorithmic_bytes": alg, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)"}

    # ---- e2e: public module, HOST inputs: H2D + forward (logits + fused arg-max) + D2H every step ----
    # Each sample's six input tensors live in ONE pinned staging buffer (what a collate function
    # would produce), so a step is one H2D copy, the module call, and one D2H copy of the occupancy
    # prediction.  The host consumes prediction i-1 while step i is in flight (two result buffers),
    # i.e. copies and kernels of consecutive steps overlap CPU work but every step's copies are
    # inside the timed region.
    module = LocalAggregator(**kw).to(dev)
    module.validate = False     # the D2H read of the prediction is the step's synchronisation point
    n_host = 3
    layouts, host = [], []
    for i in range(n_host):
        inp = inp0 if i == 0 else make_splat_inputs(WORKLOAD, seed=rank * 100 + i, perturb=PERTURB)[1]
        offs, off = {}, 0
        for k, v in inp.items():
            offs[k] = (off, v.numel(), tuple(v.shape))
            off += (v.numel() + 63) // 64 * 64          # 256-byte aligned slices
        buf = torch.empty(off, dtype=torch.float32).pin_memory()
        for k, v in inp.items():
            o, nel, _ = offs[k]
            buf[o:o + nel].copy_(v.reshape(-1))
        layouts.append(offs)
        host.append(buf)
    h2d = host[0].numel() * 4
    pred_host = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    d2h = pred_host[0].numel()
    consumed = [0]

    copy_stream = torch.cuda.Stream(dev)
    h2d_done = [torch.cuda.Event() for _ in range(n_host)]

    def e2e_step(i):
        # the H2D copy of step i runs on a copy stream and overlaps the kernels of step i-1
        with torch.cuda.stream(copy_stream):
            dbuf = host[i % n_host].to(dev, non_blocking=True)
            h2d_done[i % n_host].record(copy_stream)
        stream.wait_event(h2d_done[i % n_host])
        dbuf.record_stream(stream)
        d = {k: dbuf[o:o + nel].view(shape) for k, (o, nel, shape) in layouts[i % n_host].items()}
        _logits, occ = module.forward_with_occupancy(d["pts"], d["means"], d["opa"], d["sem"], d["scales"], d["cov"])
        pred_host[i & 1].copy_(occ, non_blocking=True)
        done[i & 1].record(stream)
        if world > 1:
            pending.append(dist.all_reduce(loss_buf, async_op=True))
        if i > 0:                                  # consume the previous step's prediction on the host
            done[(i - 1) & 1].synchronize()
            consumed[0] += int(pred_host[(i - 1) & 1][0])

    for i in range(4):
        e2e_step(i)
    e2e_steps = max(5, min(K, 100))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e2e_ms_dev = timed_region(e2e_steps, e2e_step)
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(e2e_ms_dev, 0.0) / e2e_steps
    e2e = {"value": world * G_COUNTED / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall * 1e3 / e2e_steps,
           "host_numa_binding": numa,
           "api": "local_aggregate.LocalAggregator.forward_with_occupancy (validate=False): one pinned staging "
                  "buffer per sample -> H2D on a copy stream, logits + fused arg-max, D2H of the uint8 occupancy; "
                  "host reads prediction i-1 while step i runs"}

    extras = {}
    if rank == 0 and not args.no_extras:
        extras = side_measurements(dev, kw, inp0, sets[0], desc)

    cpu_baseline = None
    if rank == 0 and world == 1:
        import oracle
        _cpu_port_once(kw, inp0)
        reps = [_cpu_port_once(kw, inp0) for _ in range(3)]
        sec = sum(reps) / len(reps)
        cpu_baseline = {"value": G_COUNTED / sec, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                        "sample": f"3 full forward passes of the {WORKLOAD} sample, {sec:.3f} s each (C/OpenMP oracle)"}
        try:
            # secondary, work-normalised figure: in-box (voxel, Gaussian) pairs per second and the fp32
            # rate they stand for at ~56 flop + 1 exp per pair; and the headline sample checked against the CPU port
            pairs = _LAST_PORT["pairs"]
            extras["work"] = {"pairs": pairs, "pairs_per_s": pairs / (render_ms * 1e-3), "flop_per_pair": 56,
                              "fp32_tflops": 56 * pairs / (render_ms * 1e-3) / 1e12,
                              "what": "in-box (voxel, Gaussian) pairs of the sample over the render kernel's time"}
            got = outs[0].float().cpu().numpy()
            want = _LAST_PORT["logits"]
            import numpy as np
            err = np.abs(got - want)
            extras["parity_vs_cpu_port"] = {
                "max_abs_err": float(err.max()), "max_rel_err_above_1e-3": float((err / np.maximum(np.abs(want), 1e-3)).max()),
                "argmax_differences": int((got.argmax(1) != want.argmax(1)).sum()), "voxels": int(got.shape[0]),
                "what": "logits of the timed sample (resident set 0) vs the fp32 C/OpenMP oracle on the host"}
        except Exception as e:
            extras["parity_error"] = repr(e)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{WORKLOAD}: 25600 Gaussians (+1 empty) -> 200x200x16x18 voxels, "
                                       f"batch 1 per GPU, points = " + ("perturbed voxel centres" if PERTURB else "voxel centres (BASELINE.md §3)"),
                           "l2": f"{N_SETS} rotating input/output/workspace sets "
                                 f"({N_SETS * (alg + ws_bytes) / 1e6:.0f} MB) > 126 MB L2",
                           "parallelism": f"dp{world} (one sample per GPU, scalar loss all-reduce)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": 3 * K, "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def side_measurements(dev, kw, inp0, resident, desc):
    """Reported next to the headline (not part of it): backward, prob variant, DAF, and the
    reference CUDA op on the same GPU when oracle/_ref travelled with the repo."""
    import torch
    from gaussianformer_b200.splat import LocalAggregator, LocalAggregatorProb
    from gaussianformer_b200.synthetic import make_daf_inputs, make_splat_inputs
    from gaussianformer_b200.ops import DeformableAggregationFunction as DAF
    out = {}

    def timeit(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    try:
        m = LocalAggregator(**kw).to(dev)
        m.validate = False
        t = {k: v.to(dev) for k, v in inp0.items()}
        for k in ("means", "opa", "sem", "cov"):
            t[k].requires_grad_(True)
        logits = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
        g = torch.randn_like(logits)
        out["splat_bwd_ms"] = timeit(lambda: torch.autograd.grad(logits, [t["means"], t["opa"], t["sem"], t["cov"]], g,
                                                                 retain_graph=True), reps=10)
        kwp, inpp, _ = make_splat_inputs("prob_gs6400", seed=0, perturb=True)
        mp = LocalAggregatorProb(**kwp).to(dev)
        mp.validate = False
        tp = {k: v.to(dev) for k, v in inpp.items()}
        out["prob_gs6400_fwd_ms"] = timeit(lambda: mp(tp["pts"], tp["means"], tp["opa"], tp["sem"], tp["scales"], tp["cov"]), reps=10)
        for k in ("means", "opa", "sem", "cov"):
            tp[k].requires_grad_(True)
        lg, bl, de = mp(tp["pts"], tp["means"], tp["opa"], tp["sem"], tp["scales"], tp["cov"])
        gp = [torch.randn_like(lg), torch.randn_like(bl), torch.randn_like(de)]
        out["prob_gs6400_bwd_ms"] = timeit(lambda: torch.autograd.grad([lg, bl, de], [tp["means"], tp["opa"], tp["sem"], tp["cov"]],
                                                                       gp, retain_graph=True), reps=5)
        del lg, bl, de, gp, tp
        # BASELINE.json configs[2]: 144 000 Gaussians into the same grid (forward, voxel-centre points)
        kwl, inpl, _ = make_splat_inputs("gs144000", seed=0, perturb=False)
        ml = LocalAggregator(**kwl).to(dev)
        ml.validate = False
        tl = {k: v.to(dev) for k, v in inpl.items()}
        out["gs144000_fwd_ms"] = timeit(lambda: ml(tl["pts"], tl["means"], tl["opa"], tl["sem"], tl["scales"], tl["cov"]), reps=10)
        try:
            # north star: mIoU unchanged (SURVEY.md 8d recipe on the 144 000-Gaussian sample, whose arg-max spreads over
            # all 18 classes): fused arg-max of the CUDA path vs the CPU port's arg-max, scored with the reference's MeanIoU
            from gaussianformer_b200.metric import miou_parity, synthetic_labels
            _lg, occ = ml.forward_with_occupancy(tl["pts"], tl["means"], tl["opa"], tl["sem"], tl["scales"], tl["cov"])
            _cpu_port_once(kwl, inpl)
            want = _LAST_PORT["logits"]
            labels, mask = synthetic_labels(want, want.shape[1])
            r = miou_parity(occ.cpu(), want.argmax(1), labels, mask, want.shape[1])
            out["miou_parity_gs144000"] = {
                "miou_cuda": r["new"][0], "miou_cpu_port": r["ref"][0], "abs_diff": r["abs_diff"],
                "argmax_differences": int((occ.cpu().numpy().astype("int64") != want.argmax(1)).sum()),
                "voxels": int(want.shape[0])}
        except Exception as e:
            out["miou_parity_error"] = repr(e)
        del tl
        fms, loc, w = make_daf_inputs(seed=0)
        feat, shape, start = DAF.feature_maps_format([f.to(dev) for f in fms])
        feat = feat.contiguous()
        loc, w = loc.to(dev), w.to(dev)
        dev_maps = [f.to(dev) for f in fms]
        out["daf_format_ms"] = timeit(lambda: DAF.feature_maps_format(dev_maps)[0], reps=10)

        def torch_format():      # the reference's route: reshape + cat + permute, then the copy .contiguous() makes
            bs, cams, ch = dev_maps[0].shape[:3]
            return torch.cat([m.reshape(bs, cams, ch, -1) for m in dev_maps], dim=-1).permute(0, 1, 3, 2).contiguous()
        out["ref_format_torch_ms"] = timeit(torch_format, reps=10)
        del dev_maps
        out["daf_fwd_ms"] = timeit(lambda: DAF.apply(feat, shape, start, loc, w), reps=10)
        out["daf_fwd_alg_gbs"] = 4 * (feat.numel() + loc.numel() + w.numel() + loc.shape[1] * 128) / (out["daf_fwd_ms"] * 1e-3) / 1e9
    except Exception as e:  # side figures must never sink the headline
        out["extras_error"] = repr(e)
    try:
        # fused caller path: masked softmax + op + key-point sum, against the same three steps done
        # the reference's way (PyTorch softmax -> our drop-in op -> .sum) on the same inputs
        from gaussianformer_b200.ops import deformable_aggregation_fused
        from gaussianformer_b200.synthetic import make_daf_fused_inputs, reference_fused_composition
        fms, loc, lg, pm, _ = make_daf_fused_inputs(seed=0)
        feat, shape, start = DAF.feature_maps_format([f.to(dev) for f in fms])
        feat = feat.contiguous()
        loc, lg, pm = loc.to(dev), lg.to(dev), pm.to(dev)
        with torch.no_grad():
            out["daf_fused_fwd_ms"] = timeit(lambda: deformable_aggregation_fused(feat, shape, start, loc, lg, pm), reps=10)
            out["daf_unfused_fwd_ms"] = timeit(lambda: reference_fused_composition(DAF.apply, feat, shape, start, loc, lg, pm), reps=10)
        feat.requires_grad_(True); loc.requires_grad_(True); lg.requires_grad_(True)
        g = torch.randn(1, lg.shape[1], feat.shape[-1], device=dev)

        def fwd_bwd(fn):
            o = fn(feat, shape, start, loc, lg, pm)
            return torch.autograd.grad(o, [feat, loc, lg], g)
        out["daf_fused_fwd_bwd_ms"] = timeit(lambda: fwd_bwd(deformable_aggregation_fused), reps=5)
        out["daf_unfused_fwd_bwd_ms"] = timeit(lambda: fwd_bwd(lambda *a: reference_fused_composition(DAF.apply, *a)), reps=5)
        del feat, loc, lg, pm, g
    except Exception as e:
        out["daf_fused_error"] = repr(e)
    try:
        # the reference's own fallback for the sampling op (PyTorch grid_sample route, deformable_module.py:307-353,
        # restated in oracle.daf_torch_fallback) on the host cores, on a bounded sample of the same workload
        import oracle
        fms, loc, w = make_daf_inputs(seed=0)
        nsub = 23040                                        # one tenth of the 230 400 sampling points
        loc_s, w_s = loc[:, :nsub].contiguous(), w[:, :nsub].contiguous()
        with torch.no_grad():
            oracle.daf_torch_fallback(fms, loc_s, w_s, 4)
            t0 = time.perf_counter()
            oracle.daf_torch_fallback(fms, loc_s, w_s, 4)
            sec = time.perf_counter() - t0
        out["daf_cpu_fallback"] = {"ms_per_full_call_extrapolated": sec * 1e3 * loc.shape[1] / nsub, "sample_ms": sec * 1e3,
                                   "sample": f"{nsub} of {loc.shape[1]} sampling points, 6 cameras x 4 levels, C = 128",
                                   "cores": torch.get_num_threads(), "kind": "port",
                                   "what": "reference's PyTorch fallback of the op (grid_sample + weighted fusion) on the host"}
    except Exception as e:
        out["daf_cpu_fallback"] = {"error": repr(e)}
    try:
        from oracle import build_ref
        if build_ref.available("gf_ref_daf"):
            mod = build_ref.load_ref("gf_ref_daf")
            fms, loc, w = make_daf_inputs(seed=0)
            feat, shape, start = DAF.feature_maps_format([f.to(dev) for f in fms])
            feat = feat.contiguous()
            loc, w, shape_i, start_i = loc.to(dev), w.to(dev), shape.int(), start.int()
            ref_f = timeit(lambda: mod.deformable_aggregation_forward(feat, shape_i, start_i, loc, w), reps=10)
            g = torch.randn(1, loc.shape[1], 128, device=dev)
            gf, gl, gw = torch.zeros_like(feat), torch.zeros_like(loc), torch.zeros_like(w)
            ref_b = timeit(lambda: mod.deformable_aggregation_backward(feat, shape_i, start_i, loc, w, g, gf, gl, gw), reps=5)
            feat.requires_grad_(True); loc.requires_grad_(True); w.requires_grad_(True)
            o = DAF.apply(feat, shape, start, loc, w)
            out["daf_bwd_ms"] = timeit(lambda: torch.autograd.grad(o, [feat, loc, w], g, retain_graph=True), reps=5)
            out["ref_daf_op"] = {"fwd_ms": ref_f, "bwd_kernel_ms": ref_b, "what": "reference deformable_aggregation op "
                                 "compiled for sm_100a, same GPU, same inputs (backward without its three zero fills)"}
    except Exception as e:
        out["ref_daf_op"] = {"error": repr(e)}
    try:
        from oracle import build_ref
        if build_ref.available("gf_ref_localagg"):
            mod = build_ref.load_ref("gf_ref_localagg")
            t = {k: v[0].to(dev) for k, v in inp0.items()}
            pc_min = torch.tensor(kw["pc_min"], device=dev)[None]

            def ref_call():   # the reference's Python wrapper + native op (asserts included)
                pi = ((t["pts"] - pc_min) / kw["grid_size"]).to(torch.int)
                assert pi.min() >= 0
                mi = ((t["means"] - pc_min) / kw["grid_size"]).to(torch.int)
                assert mi.min() >= 0
                radii = torch.ceil(t["scales"].max(dim=-1)[0] * kw["scale_multiplier"] / kw["grid_size"]).to(torch.int)
                assert radii.min() >= 1
                cov6 = t["cov"].flatten(1)[:, [0, 4, 8, 1, 5, 2]]
                return mod.local_aggregate(t["pts"], pi, t["means"], mi, t["opa"], t["sem"], radii, cov6,
                                           kw["H"], kw["W"], kw["D"])
            ref_fwd = timeit(ref_call, reps=10)
            R, logits, geom, binning, img = ref_call()
            pi = ((t["pts"] - pc_min) / kw["grid_size"]).to(torch.int)
            cov6 = t["cov"].flatten(1)[:, [0, 4, 8, 1, 5, 2]].contiguous()
            g = torch.randn_like(logits)
            ref_bwd = timeit(lambda: mod.local_aggregate_backward(geom, binning, img, kw["H"], kw["W"], kw["D"], R,
                                                                  t["means"], t["pts"], pi, cov6, t["opa"], t["sem"], g),
                             reps=3, warm=1)
            out["ref_cuda_op"] = {"fwd_ms": ref_fwd, "bwd_ms": ref_bwd, "what": "reference localagg op "
                                  "(its Python prep + sort-based kernels; backward = native call only) compiled for "
                                  "sm_100a, same GPU, same sample"}
    except Exception as e:
        out["ref_cuda_op"] = {"error": repr(e)}
    return out


if __name__ == "__main__":
    main()
