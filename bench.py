#!/usr/bin/env python
"""Benchmark of the Any-Stereo iterative cost-volume hot path (BASELINE.json metric:
"pairs/s @384x1248, 32 iters, 1-8 B200; corr-lookup HBM GB/s vs peak").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own code on the host cores, rank 0 only

One "step" = one pass of the hot path over one batch of 8 synthetic KITTI-shaped pairs per GPU
(BASELINE.json configs[1]: coreContinuous_IGEV, 384x1248 -> 96x312 at 1/4, 32 iterations):
    build_gwc_volume -> Combined_Geo_Encoding_Volume(...) (all-pairs correlation + both pyramids)
    -> 32 x { geo/corr lookup -> BasicMultiUpdateBlock -> disp += delta }.
Backbones, the 3-D hourglass and the LIIF upsampler are outside the path and outside the
step; the GWC volume stands in for the aggregated geometry volume (same shape, same traffic).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H4, W4, FEAT_D, GEO_D, GROUPS, ITERS = 96, 312, 96, 48, 8, 32     # 384x1248 at 1/4 resolution
LOOKUP_BYTES_PER_PIXEL = 1372                                      # IGEV L=2, r=4, G=8
# fused lookup+convc1: the same 724 B of windows + 4 B disp read, 64 bf16 hi(+lo) channels written
FUSED_BYTES_PER_PIXEL = {"bf16x3": 728 + 256, "f16f8": 728 + 256, "bf16": 728 + 128, "fp16": 728 + 128}
PASSES = {"bf16x3": 3, "f16f8": 2}          # tensor-core pass-equivalents per MAC of the fp32-parity engines


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region.  Read in-process through NVML (the library
    nvidia-smi itself uses; `nvidia_ml_py`) so that no process is forked next to the kernel-launching thread; falls back
    to spawning `nvidia-smi --query-gpu=clocks.sm,...` when the module is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)) if not str(uuid).startswith("GPU-") else str(uuid))
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = (pynvml, h)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv, h = self._nvml
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        act = lambda bit: "Active" if (r & bit) else "Not Active"      # noqa: E731
        # bit values of nvmlClocksEventReasons: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        return [str(sm), str(mx), "%.1f" % pw, act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 7:
                        self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.1 if self._nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(3)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples),
                "source": "nvml (in-process)" if self._nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's OWN code (baseline/_ref) on the host cores; oracle port if absent
# --------------------------------------------------------------------------------------------------
WORKLOAD = ("coreContinuous_IGEV hot path (BASELINE configs[1]): 384x1248 -> 96x312 @1/4, batch %d pairs/GPU, 32 iters, "
            "corr_levels=2, radius=4")


class CpuReference:
    """One synthetic 384x1248 pair through the path on the CPU: build_gwc_volume -> Combined_Geo_Encoding_Volume(...) ->
    32 x {geo_fn(disp, coords) -> update_block(...) -> disp += delta}, ALL 32 iterations timed (no extrapolation).

    kind "reference": the unmodified reference modules (models/coreContinuous_IGEV/{submodule,geometry,update}.py from
    the pristine copy baseline/_ref, imported by oracle/ref_loader.py), called exactly as continuous_IGEVstereo.py:262,
    275-295 calls them.  kind "port": oracle/hotpath_oracle.py (the same torch calls restated), only when the reference
    tree is absent."""

    def __init__(self, threads=None):
        import torch
        from oracle import ref_loader
        self.torch = torch
        self.threads = threads or os.cpu_count()
        torch.set_num_threads(self.threads)
        g = torch.Generator().manual_seed(0)
        sizes = [(H4, W4), (H4 // 2, W4 // 2), (H4 // 4, W4 // 4)]
        self.f1 = torch.randn(1, FEAT_D, H4, W4, generator=g)
        self.f2 = torch.randn(1, FEAT_D, H4, W4, generator=g)
        self.net = [torch.tanh(torch.randn(1, 128, h, w, generator=g)) for h, w in sizes]
        self.inp = [[torch.relu(torch.randn(1, 128, h, w, generator=g)) for _ in range(3)] for h, w in sizes]
        self.disp = torch.rand(1, 1, H4, W4, generator=g) * 12.0
        self.coords = torch.arange(W4).float().reshape(1, 1, W4, 1).repeat(1, H4, 1, 1)
        if ref_loader.available():
            R = ref_loader.load()
            self.kind = "reference"
            self.R = R
            torch.manual_seed(0)
            self.block = R.IGEVUpdateBlock(ref_loader.update_block_args("igev"), hidden_dims=[128, 128, 128]).eval()
            self.where = os.path.relpath(ref_loader.REF_ROOT, ROOT) if ref_loader.REF_ROOT.startswith(ROOT) else ref_loader.REF_ROOT
        else:
            from oracle import hotpath_oracle as O
            self.kind = "port"
            self.O = O
            self.params = O.make_update_block_params(162, seed=0)
            self.where = "oracle/hotpath_oracle.py"

    def sample(self, iters=ITERS):
        """-> seconds for one pair (volume build + `iters` iterations)."""
        torch = self.torch
        with torch.no_grad():
            net = [t.clone() for t in self.net]
            disp = self.disp.clone()
            t0 = time.perf_counter()
            if self.kind == "reference":
                R = self.R
                geo = R.build_gwc_volume(self.f1, self.f2, GEO_D, GROUPS)              # continuous_IGEVstereo.py:262
                geo_fn = R.Combined_Geo_Encoding_Volume(self.f1.float(), self.f2.float(), geo.float(), radius=4,
                                                        num_levels=2)                    # :275-276
                for _ in range(iters):                                                   # :284-295
                    disp = disp.detach()
                    feat = geo_fn(disp, self.coords)
                    net, delta = self.block(net, self.inp, feat, disp, iter16=True, iter08=True)
                    disp = disp + delta
            else:
                O = self.O
                geo = O.gwc_volume(self.f1, self.f2, GEO_D, GROUPS)
                cp = O.corr_pyramid(O.all_pairs_corr(self.f1, self.f2), 2)
                gp = O.geo_pyramid(geo, 2)
                for _ in range(iters):
                    feat = O.geo_lookup(gp, cp, disp, self.coords, 4)
                    net, delta = O.update_block(self.params, net, self.inp, feat, disp)
                    disp = disp + delta
            return time.perf_counter() - t0

    def describe(self, iters=ITERS):
        return ("1 pair 384x1248 per step on %d host threads: build_gwc_volume + Combined_Geo_Encoding_Volume + all %d "
                "iterations of lookup + update block timed (no extrapolation); code = %s (%s)"
                % (self.threads, iters, self.where, "unmodified reference modules" if self.kind == "reference" else "oracle port"))


def cpu_reference_sample(threads=None):
    """cpu_baseline leg of our arm: one untimed warm-up of 2 iterations, then ONE full pair (~2-4 s)."""
    ref = CpuReference(threads)
    ref.sample(iters=2)
    s = ref.sample()
    return {"pairs_per_s": 1.0 / s, "s_per_pair": s, "cores": ref.threads, "kind": ref.kind, "sample": ref.describe()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    for _ in range(args.warmup):
        ref.sample(iters=2)                       # thread pools, allocator, oneDNN primitive caches
    times = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        times.append(ref.sample())
    wall = time.perf_counter() - t0
    s_step = sum(times) / len(times)
    v = 1.0 / s_step
    line = {
        "impl": "reference", "metric": "pairs/s @384x1248, 32 iters", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.pairs_per_gpu, "pairs_per_gpu": args.pairs_per_gpu, "iters": ITERS,
                   "reference_sample": "each step = 1 pair of that workload on the CPU (bounded sample)"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": ref.threads, "kind": ref.kind, "sample": ref.describe()},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall, "steps_done": len(times),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def make_inputs(torch, B, device, seed, pinned=False):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sizes = [(H4, W4), (H4 // 2, W4 // 2), (H4 // 4, W4 // 4)]

    def mk(*shape, fn=None):
        t = torch.randn(*shape, generator=g)
        if fn is not None:
            t = fn(t)
        if pinned:
            return t.pin_memory()
        return t.to(device)

    d = {
        "ml": mk(B, FEAT_D, H4, W4), "mr": mk(B, FEAT_D, H4, W4),
        "net": [mk(B, 128, h, w, fn=torch.tanh) for h, w in sizes],
        "inp": [[mk(B, 128, h, w, fn=torch.relu) for _ in range(3)] for h, w in sizes],
        "disp": mk(B, 1, H4, W4, fn=lambda t: t.abs() * 12.0),
    }
    return d


def host_bytes(d):
    n = d["ml"].numel() + d["mr"].numel() + d["disp"].numel()
    n += sum(t.numel() for t in d["net"]) + sum(t.numel() for l in d["inp"] for t in l)
    return 4 * n


def run_ours(args):
    # NCCL_DEBUG=VERSION makes NCCL print its banner on stdout, next to the one JSON line the contract asks for
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist

    import anystereo_b200 as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (printed on stdout at communicator creation when
        # NCCL_DEBUG >= VERSION) is sent to stderr by swapping the descriptor around init + the first collective
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    A.set_update_engine(args.engine)
    A.update_umma.set_encoder_overlap(not args.no_overlap)
    A.set_lookup_fusion(not args.no_fusion)
    A.set_corr_mode(args.corr_mode or {"fp32": "fp32", "bf16": "bf16"}.get(args.engine, "bf16x3"))
    B = args.pairs_per_gpu
    torch.manual_seed(0)
    uargs = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    block = A.BasicMultiUpdateBlock(uargs, hidden_dims=[128, 128, 128]).to(dev).eval()   # random init, seed 0
    dd = make_inputs(torch, B, dev, seed=rank)                    # device-resident inputs for `value`
    hh = make_inputs(torch, B, dev, seed=rank, pinned=True)      # pinned host inputs for `e2e`
    stage = make_inputs(torch, B, dev, seed=1234 + rank)         # device staging buffers the H2D copies fill
    host_out = torch.empty(B, 1, H4, W4).pin_memory()
    L = A._lib

    def step(d, events=None, uevents=None):
        geo = A.build_gwc_volume(d["ml"], d["mr"], GEO_D, GROUPS)
        disp, net = A.igev_iterations(block, d["ml"], d["mr"], geo, d["net"], d["inp"], d["disp"], ITERS,
                                      radius=4, num_levels=2, lookup_events=events, update_events=uevents)
        return disp

    # e2e: every step copies ITS inputs from pinned host memory and reads its result back.  Two device staging
    # sets + a copy stream let step i+1's H2D overlap step i's kernels (plain double buffering).
    stages = [stage, make_inputs(torch, B, dev, seed=4321 + rank)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]      # staging set filled
    freed = [torch.cuda.Event(), torch.cuda.Event()]      # staging set consumed
    e2e_state = {"i": 0, "primed": False}

    def h2d(dst):
        dst["ml"].copy_(hh["ml"], non_blocking=True)
        dst["mr"].copy_(hh["mr"], non_blocking=True)
        dst["disp"].copy_(hh["disp"], non_blocking=True)
        for d_, s_ in zip(dst["net"], hh["net"]):
            d_.copy_(s_, non_blocking=True)
        for dl, sl in zip(dst["inp"], hh["inp"]):
            for d_, s_ in zip(dl, sl):
                d_.copy_(s_, non_blocking=True)

    def e2e_step():
        cur = e2e_state["i"] & 1
        main = torch.cuda.current_stream()
        if not e2e_state["primed"]:                      # first call: copy this step's inputs synchronously
            with torch.cuda.stream(copy_stream):
                h2d(stages[cur])
                ready[cur].record()
            e2e_state["primed"] = True
        with torch.cuda.stream(copy_stream):             # prefetch the NEXT step's inputs
            copy_stream.wait_event(freed[cur ^ 1])
            h2d(stages[cur ^ 1])
            ready[cur ^ 1].record()
        main.wait_event(ready[cur])
        disp = step(stages[cur])
        freed[cur].record()
        host_out.copy_(disp, non_blocking=True)
        e2e_state["i"] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.no_grad():
        # untimed warm-up: at least W (>= 3) steps AND at least ~1.5 s, so that clocks / power management have settled
        # before the timed region (two outliers at 55-60 pairs/s were seen when the bench started right after another
        # process had left the GPU idle)
        t_w = time.perf_counter()
        n_w = 0
        while n_w < max(args.warmup, 3) or (time.perf_counter() - t_w < 1.5 and n_w < 40):
            step(dd)
            torch.cuda.synchronize()
            n_w += 1
        sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None)) if rank == 0 else None
        if sampler:
            sampler.start()
        # timed region: the public call (igev_iterations replays its captured step; bench.py --eager times launches one by one)
        A.set_graph_replay(not args.eager)
        step(dd)                                          # builds the graph outside the timed region
        torch.cuda.synchronize()
        l0 = L.launch_count
        ms = timed(lambda: step(dd), args.steps)
        launches = L.launch_count - l0
        clocks = sampler.stop() if sampler else None
        torch.cuda.synchronize()
        # per-kernel event timing needs eager launches: two more steps of the same loop with CUDA events around the lookup
        # launch and around the update block
        step(dd, [], [])                                  # untimed: the eager path warms its own caches
        torch.cuda.synchronize()
        events, uevents = [], []
        for _ in range(2):
            step(dd, events, uevents)
        torch.cuda.synchronize()
        look_us = [a.elapsed_time(b) * 1e3 for a, b in events]
        upd_us = [a.elapsed_time(b) * 1e3 for a, b in uevents]
        # The lookup kernel shares the GPU with the 1/8- and 1/16-scale GRU convolutions (motion encoder on a side
        # stream): its event-to-event time in the region above is not its own duration.  Two more steps of the SAME loop
        # with the side stream off give the kernel's exclusive launch durations, which the roofline line uses.
        look_us_overlapped = look_us
        if not args.no_overlap:
            A.update_umma.set_encoder_overlap(False)
            ev_x = []
            for _ in range(2):
                step(dd, ev_x, None)
            torch.cuda.synchronize()
            look_us = [a.elapsed_time(b) * 1e3 for a, b in ev_x]
            A.update_umma.set_encoder_overlap(True)
        for _ in range(4):
            e2e_step()
        torch.cuda.synchronize()
        ms_e2e = timed(e2e_step, args.steps)

    # the other BASELINE.json configs that run the same path (parity-test cases, reported for context).  Every rank runs
    # them on its own pairs (weak scaling, like the headline); times are the MAX over ranks, values the whole-job sum.
    other = {}

    def maxr(ms_):
        if world > 1:
            t_ = torch.tensor([ms_], device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item())
        return ms_

    def ev_ms(fn, reps):
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            fn()
        r1.record()
        torch.cuda.synchronize()
        return maxr(r0.elapsed_time(r1) / reps)

    if not args.no_other_configs:
        with torch.no_grad():
            rblock = A.BasicMultiUpdateBlockRAFT(types.SimpleNamespace(corr_levels=4, corr_radius=4, n_gru_layers=3),
                                                 hidden_dims=[128, 128, 128]).to(dev).eval()
            for name, (rb, rh, rw) in {"config1_raft_320x736_b1": (1, 80, 184), "config3_raft_1984x2880_b1": (1, 496, 720)}.items():
                g = torch.Generator(device="cpu").manual_seed(7 + rank)
                rs = [(rh, rw), (rh // 2, rw // 2), (rh // 4, rw // 4)]
                rf1 = (torch.randn(rb, 256, rh, rw, generator=g) / 4).to(dev)
                rf2 = (torch.randn(rb, 256, rh, rw, generator=g) / 4).to(dev)
                rnet = [torch.tanh(torch.randn(rb, 128, h_, w_, generator=g)).to(dev) for h_, w_ in rs]
                rinp = [[torch.relu(torch.randn(rb, 128, h_, w_, generator=g)).to(dev) for _ in range(3)] for h_, w_ in rs]
                A.set_graph_replay(False)                       # every kernel launched one by one
                for _ in range(2):
                    A.raft_iterations(rblock, rf1, rf2, rnet, rinp, ITERS)
                rms = ev_ms(lambda: A.raft_iterations(rblock, rf1, rf2, rnet, rinp, ITERS), 3)
                A.set_graph_replay(not args.eager)              # the public call as it is: replays its captured step
                for _ in range(2):
                    A.raft_iterations(rblock, rf1, rf2, rnet, rinp, ITERS)
                gms = ev_ms(lambda: A.raft_iterations(rblock, rf1, rf2, rnet, rinp, ITERS), 3)
                other[name] = {"ms_per_pair": gms / rb, "pairs_per_s": world * rb / (gms / 1e3), "iters": ITERS,
                               "corr_levels": 4, "pairs_per_gpu": rb, "eager_ms_per_pair": rms / rb,
                               "eager_pairs_per_s": world * rb / (rms / 1e3),
                               "note": "raft_iterations() as called by a user (CUDA-graph replay incl. the copies into its "
                                       "static buffers); eager_* = the same call with set_graph_replay(False)"}
                A.hotpath.graph_cache_clear()
                del rf1, rf2, rnet, rinp
            # the mixed-precision analogue (IEEE-half operands, single MMA): same step, reported for context only --
            # inside the 0.01 px EPE gate at the BASELINE shapes (tests/test_gpu_dropin.py, profiles/dropin_epe_r02.json)
            # but not inside the 1e-4 operator tolerance, so never the headline
            if args.engine in ("bf16x3", "f16f8"):
                A.set_update_engine("fp16")
                block.reset_caches()
                for _ in range(2):
                    step(dd)
                fms = ev_ms(lambda: step(dd), 3)
                other["engine_fp16_same_step"] = {"pairs_per_s": world * B / (fms / 1e3), "ms_per_step": fms}
                A.set_update_engine(args.engine)
                block.reset_caches()
            if args.engine == "f16f8":                       # both cross terms everywhere (the default: one pass in gru08 / gru16,
                A.set_lowres_single_pass(False)              # weight-residual term only in the 1/4-resolution gates)
                A.set_gate_weight_residual_only(False)
                block.reset_caches()
                for _ in range(2):
                    step(dd)
                lms = ev_ms(lambda: step(dd), 6)
                other["engine_f16f8_two_passes_everywhere_same_step"] = {
                    "pairs_per_s": world * B / (lms / 1e3), "ms_per_step": lms,
                    "what": "set_lowres_single_pass(False) + set_gate_weight_residual_only(False): every layer hi*hi + both cross "
                            "terms (final-disparity EPE 1.51e-4 px instead of 1.44e-4 px on the real IGEV graph, tests/test_gpu_dropin.py)"}
                A.set_lowres_single_pass(True)
                A.set_gate_weight_residual_only(None)
                block.reset_caches()
            if args.engine == "f16f8":                       # the 3-pass split of round 1, same step, for context
                A.set_update_engine("bf16x3")
                block.reset_caches()
                for _ in range(2):
                    step(dd)
                bms = ev_ms(lambda: step(dd), 3)
                other["engine_bf16x3_same_step"] = {"pairs_per_s": world * B / (bms / 1e3), "ms_per_step": bms}
                A.set_update_engine(args.engine)
                block.reset_caches()
            # the drop-in call pattern (what a reference user gets by rebinding the names only): geo_fn(disp, coords)
            # materialises the [B,162,h,w] tensor, update_block(...) consumes NCHW tensors -- no `deferred`, no fusion
            if args.engine != "fp32":
                def dropin_step():
                    geo = A.build_gwc_volume(dd["ml"], dd["mr"], GEO_D, GROUPS)
                    geo_fn = A.Combined_Geo_Encoding_Volume(dd["ml"].float(), dd["mr"].float(), geo.float(), radius=4, num_levels=2)
                    coords = A.hotpath.pixel_coords(B, H4, W4, dev)
                    net, disp = list(dd["net"]), dd["disp"]
                    for _ in range(ITERS):
                        feat = geo_fn(disp, coords)
                        net, delta = block(net, dd["inp"], feat, disp, iter16=True, iter08=True)
                        disp = disp + delta
                    return disp
                for _ in range(2):
                    dropin_step()
                dms = ev_ms(dropin_step, 6)
                # the same loop with install_into_reference(defer_lookup=True): geo_fn returns the lookup unevaluated and the
                # adopted update block runs it fused with convc1
                def dropin_step_deferred():
                    geo = A.build_gwc_volume(dd["ml"], dd["mr"], GEO_D, GROUPS)
                    geo_fn = A.geometry.Combined_Geo_Encoding_Volume_Deferred(dd["ml"].float(), dd["mr"].float(), geo.float(),
                                                                              radius=4, num_levels=2)
                    coords = A.hotpath.pixel_coords(B, H4, W4, dev)
                    net, disp = list(dd["net"]), dd["disp"]
                    for _ in range(ITERS):
                        feat = geo_fn(disp, coords)
                        net, delta = block(net, dd["inp"], feat, disp, iter16=True, iter08=True)
                        disp = disp + delta
                    return disp
                for _ in range(2):
                    dropin_step_deferred()
                dms2 = ev_ms(dropin_step_deferred, 6)
                other["dropin_call_pattern_deferred_lookup_same_step"] = {
                    "pairs_per_s": world * B / (dms2 / 1e3), "ms_per_step": dms2,
                    "what": "the same reference call pattern with install_into_reference(defer_lookup=True): the lookup is "
                            "fused with convc1 inside the adopted update block"}
                # one pair per forward (the reference's evaluation / demo batch size): the Python loop then runs at the host's pace;
                # adopt_update_block(..., replay=True) replays each update-block call from a CUDA graph (bit-identical results)
                one = {k: (v[:1].contiguous() if torch.is_tensor(v) else v) for k, v in dd.items()}
                one["net"] = [t[:1].contiguous() for t in dd["net"]]
                one["inp"] = [[t[:1].contiguous() for t in lst] for lst in dd["inp"]]

                def dropin_one_pair():
                    geo = A.build_gwc_volume(one["ml"], one["mr"], GEO_D, GROUPS)
                    geo_fn = A.geometry.Combined_Geo_Encoding_Volume_Deferred(one["ml"].float(), one["mr"].float(), geo.float(),
                                                                              radius=4, num_levels=2)
                    coords = A.hotpath.pixel_coords(1, H4, W4, dev)
                    net, disp = list(one["net"]), one["disp"]
                    for _ in range(ITERS):
                        feat = geo_fn(disp, coords)
                        net, delta = block(net, one["inp"], feat, disp, iter16=True, iter08=True)
                        disp = disp + delta
                    return disp
                one_ms = {}
                for tag, flag in (("eager", False), ("replayed", True)):
                    block.call_replay = flag
                    for _ in range(3):
                        d_one = dropin_one_pair()
                    one_ms[tag] = ev_ms(dropin_one_pair, 6)
                    one[tag] = d_one
                block.call_replay = None
                A.update_umma.call_replay_clear(block)
                other["dropin_call_pattern_one_pair"] = {
                    "eager_ms_per_pair": one_ms["eager"], "replayed_ms_per_pair": one_ms["replayed"],
                    "bit_identical": bool(torch.equal(one["eager"], one["replayed"])),
                    "what": "the reference call pattern at batch 1 (384x1248, 32 iterations, deferred lookup): launch by launch vs "
                            "adopt_update_block(..., replay=True) (each update-block call replayed from a CUDA graph)"}
                other["dropin_call_pattern_same_step"] = {
                    "pairs_per_s": world * B / (dms / 1e3), "ms_per_step": dms,
                    "what": "reference call pattern (continuous_IGEVstereo.py:275-295) on this library's operators: "
                            "materialised 162-channel lookup, NCHW tensors, torch add for disp += delta"}
                block.reset_caches()
            # config 4: arbitrary-scale disparity query after the loop, one 384x1248 pair per call
            aff = {"win_w": 3, "win_h": 3, "dilation": [1, 2, 4, 8]}
            liif = A.liif_out_multi_scale_Training(encoder_dim=208, mlphidden_list=[128, 64, 64], pos_dim=0,
                                                   unfold="with_v2ISU", affinity_settings=aff, number_input=2,
                                                   chanels=[176, 32]).to(dev).eval()
            g = torch.Generator(device="cpu").manual_seed(11 + rank)
            stem4 = torch.randn(1, 48, H4, W4, generator=g).to(dev)
            hid = torch.tanh(torch.randn(1, 128, H4, W4, generator=g)).to(dev)
            stem2 = torch.randn(1, 32, 2 * H4, 2 * W4, generator=g).to(dev)
            dlow = (torch.rand(1, 1, H4, W4, generator=g) * GEO_D).to(dev)
            loop_ms_per_pair = ms / args.steps / B
            for scale in (2.5, 3.7):
                Ho, Wo = int(H4 * 4 * scale), int(W4 * 4 * scale)
                ys = -1 + 1.0 / Ho + (2.0 / Ho) * torch.arange(Ho, device=dev).float()
                xs = -1 + 1.0 / Wo + (2.0 / Wo) * torch.arange(Wo, device=dev).float()
                hr = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), -1).reshape(1, -1, 2).contiguous()
                sc = torch.full((1,), scale, device=dev)
                for _ in range(2):
                    A.upsample_disp(liif, dlow, hid, stem4, stem2, None, hr_coord=hr, scale=sc)
                ums = ev_ms(lambda: A.upsample_disp(liif, dlow, hid, stem4, stem2, None, hr_coord=hr, scale=sc), 5)
                other["config4_igev_query_x%.1f" % scale] = {
                    "queries_per_pair": int(hr.shape[1]), "upsampler_ms_per_pair": ums,
                    "Mqueries_per_s": world * hr.shape[1] / ums / 1e3,
                    "loop_plus_upsampler_pairs_per_s": world * 1e3 / (loop_ms_per_pair + ums)}
                del hr
        # config 5: one training step of the hot path (forward + explicit adjoints + NCCL gradient all-reduce + clip +
        # AdamW), 8 pairs of 320x736 per GPU, 16 iterations; forward / data / weight gradients on tcgen05
        # (tools/train_step.py).  8 steps, the first two are warm-up, median of the rest; per-phase split.
        if args.engine != "fp32":
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import train_step
                with torch.enable_grad():
                    tr = train_step.run(args.engine, 8, 16, args.train_steps, 80, 184, dev, world, rank)
                other["config5_igev_train_step_320x736_b8"] = {
                    "ms_per_step": tr["ms_per_step_median"], "ms_per_step_all": tr["ms_per_step"],
                    "phase_ms_median": tr["phase_ms_median"], "allreduce": tr["allreduce"],
                    "pairs_per_s": tr["pairs_per_s"], "pairs_per_gpu": 8, "iters": 16,
                    "engine": args.engine, "loss": tr["loss"][-1], "grad_norm": tr["grad_norm_after_allreduce"][-1],
                    "peak_mem_GB": tr["peak_mem_GB"],
                    "note": "update block only is trained (backbones are synthetic leaf tensors that receive gradients)"}
            except Exception as e:                        # context only: never fail the bench line
                other["config5_igev_train_step_320x736_b8"] = {"error": repr(e)[:300]}
            finally:
                block.reset_caches()

    pairs = world * B * args.steps
    value = pairs / (ms / 1e3)
    e2e_value = pairs / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    n_pix = B * H4 * W4
    look_avg_us = sum(look_us) / max(len(look_us), 1)
    fused = (not args.no_fusion) and args.engine != "fp32"
    look_bpp = FUSED_BYTES_PER_PIXEL[args.engine] if fused else LOOKUP_BYTES_PER_PIXEL
    tap = fused and A.geometry._TAP_MAJOR
    look_kernel = (("geo_lookup_convc1_tap_kernel" if tap else "geo_lookup_convc1_kernel<2>") +
                   " (Combined_Geo_Encoding_Volume.__call__ fused with BasicMotionEncoder.convc1+ReLU)" if fused else
                   "geo_lookup_fwd_kernel<2> (Combined_Geo_Encoding_Volume.__call__)")
    achieved = look_bpp * n_pix / (look_avg_us * 1e-6) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get("geo_lookup_convc1_bytes_per_launch" if fused else "geo_lookup_fwd_bytes_per_launch")
    except Exception:
        pass
    # second roofline: the update block (tensor-bound).  FLOPs per SURVEY.md 8(d); in the bf16x3 mode every MAC is
    # issued 3 times, so "issued" is what the tensor pipe executes; peak = measured sustained cuBLAS bf16.
    upd_avg_us = sorted(upd_us)[len(upd_us) // 2] if upd_us else 0.0          # median: immune to a one-off stall
    upd_flops = 2.0 * (n_pix * (1847488 + 64 * 162) + n_pix / 4 * 1327104 + n_pix / 16 * 884736)
    issued = upd_flops * PASSES.get(args.engine, 1)
    if args.engine == "f16f8" and A.update_umma._LOWRES_1PASS["on"]:      # gru08 / gru16 run one pass
        issued -= 2.0 * (n_pix / 4 * 1327104 + n_pix / 16 * 884736)
    if args.engine == "f16f8" and A.update_umma._GATE_WL["on"] is not False and block.gate_weight_residual_only:
        issued -= 0.5 * n_pix * 2.0 * 9 * 384 * 256                      # gru04 z|r: half of the e5m2 pass dropped
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tpeak = float(json.load(f)["bf16_tflops_sustained"])
        tpeak_src = "measured sustained (MEASURED_PEAKS.json)"
    except Exception:
        tpeak, tpeak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
    cpu = cpu_reference_sample() if not args.no_cpu_baseline else None
    line = {
        "metric": "pairs/s @384x1248, 32 iters", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "warmup_steps_run": n_w, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "f32 (3x split-bf16 on tcgen05, fp32 accumulate)",
                  "f16f8": "f32 (2-pass split on tcgen05: IEEE-half hi*hi + one e5m2 pass for both cross terms, fp32 accumulate; "
                           "the 1/8- and 1/16-resolution GRUs in one IEEE-half pass, the 1/4-resolution gates with the weight-residual "
                           "cross term only)",
                  "bf16": "bf16",
                  "fp16": "f16 (IEEE half operands, fp32 accumulate; mixed-precision analogue)"}[args.engine],
        "data": "synthetic",
        "config": {"workload": WORKLOAD % B, "pairs_per_gpu": B, "iters": ITERS, "engine": args.engine, "corr_mode": A.get_corr_mode(),
                   "lookup_fused_with_convc1": fused, "cuda_graph_replay": not args.eager,
                   "parallelism": "pairs sharded across ranks, no data-path collective",
                   "l2": "inputs larger than L2: per step ~1 GB of pyramids + ~1.8 GB of activations per iteration stream through the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": host_bytes(hh),
                "d2h_bytes_per_step": 4 * B * H4 * W4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"kernel": look_kernel, "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": look_bpp * n_pix, "algorithmic_bytes_per_pixel": look_bpp,
                     "avg_launch_us": look_avg_us, "launches_timed": len(look_us),
                     "avg_launch_us_in_timed_region_sharing_the_gpu": sum(look_us_overlapped) / max(len(look_us_overlapped), 1),
                     "timing": "CUDA events on the launching stream inside the 32-iteration loop; exclusive durations from 2 "
                               "extra steps of the same loop without the side stream (in the timed region the kernel overlaps "
                               "the low-resolution GRU convolutions, so its event-to-event time there is not its own)",
                     "note": ("fused kernel: the three kernels it replaces (lookup, bf16 split, convc1) moved 3812 B/pixel; "
                              "second generation (16-byte gathers, shuffle interpolation, 32-byte stores; AS_LOOKUP_C1_TAP=0 "
                              "selects the first): gather and interpolation do not fully overlap (LDG issue blocks in the LSU "
                              "queue), see DESIGN.md section 5.2; --no-fusion reports the plain lookup kernel (1372 B/pixel, "
                              "frac 0.47-0.49)")
                     if fused else "plain lookup kernel (--no-fusion)"},
        "roofline_update_block": None if args.engine == "fp32" else {
            "kernels": "conv_umma_kernel (tcgen05) + small kernels per iteration" + (" + the fused lookup/convc1 kernel" if fused else ""), "bound": "tensor",
            "achieved": issued / (upd_avg_us * 1e-6) / 1e12, "peak": tpeak, "unit": "TFLOP/s", "frac": issued / (upd_avg_us * 1e-6) / 1e12 / tpeak,
            "logical_tflops": upd_flops / (upd_avg_us * 1e-6) / 1e12, "mma_issue_factor": issued / upd_flops,
            "avg_us_per_iteration": upd_avg_us, "peak_source": tpeak_src},
        "cpu_baseline": None if cpu is None else {"value": cpu["pairs_per_s"], "unit": "pairs/s", "cores": cpu["cores"],
                                                  "kind": cpu["kind"], "sample": cpu["sample"]},
        "clocks": clocks,
        "other_configs": other,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("ANYSTEREO_ENGINE", "f16f8"), choices=["fp32", "bf16x3", "f16f8", "bf16", "fp16"])
    ap.add_argument("--corr-mode", default=None, choices=[None, "fp32", "bf16x3", "bf16"])
    ap.add_argument("--pairs-per-gpu", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--train-steps", type=int, default=8)
    ap.add_argument("--no-fusion", action="store_true", help="materialise the 162-channel lookup tensor (A/B knob)")
    ap.add_argument("--no-overlap", action="store_true", help="keep the motion encoder on the main stream (A/B knob)")
    ap.add_argument("--eager", action="store_true", help="launch every kernel of the step instead of replaying the captured step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
