"""Per-kernel micro-benchmarks at BASELINE.json config-2 shapes (IGEV 384x1248 -> 96x312, B=8) and config-3
for the RAFT lookup: CUDA-event timing, >=20 reps after warm-up, L2 flushed between reps.  Algorithmic bytes
per SURVEY.md 8(d).  Usage: python tools/microbench.py [--B 8] [--json out.json]
`--only train` times the weight-gradient kernels of the training path at config-5 sizes, tensor cores vs CUDA cores
(added at the end of round 1 after the GPU budget was spent: first numbers are due in round 2)."""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import anystereo_b200 as A  # noqa: E402

PEAK = 6543.1
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, reps=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)   # us
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def fused_lookup_bench(blk, d, coords, rec, N):
    """8(f)-1: lookup + convc1 + ReLU in one kernel; algorithmic bytes 728 read + 256 written per pixel (two output planes)
    or + 128 (one plane).  One line per arithmetic mode of the kernel's own GEMM / output planes."""
    B, _, H, W = d.shape
    for engine, split in (("bf16x3", True), ("f16f8", True), ("bf16", False)):
        for tap in (False, True):                   # first kernel / tap-major kernel (as_geo_lookup_convc1_tap)
            A.set_update_engine(engine)
            bias = torch.randn(64, device="cuda")
            with A._lib.operand_format_scope(A.update_umma._sixteen_bit_format()):
                w_hi, w_lo = A.geometry.DeferredGeoLookup.pack_convc1_weight(torch.randn(64, 162, 1, 1, device="cuda") * 0.1,
                                                                             split, tap, bias)
            o_hi = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
            o_lo = torch.empty_like(o_hi) if split else None
            dl = blk.deferred(d, coords)
            med, best = timeit(lambda: dl.convc1_planes(w_hi, w_lo, bias, o_hi, o_lo, tap))
            rec("geo_lookup_convc1_fused_" + engine + ("_tap" if tap else ""), med, best, (728 + (256 if split else 128)) * N)
    A.set_update_engine("fp32")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--json", default=None)
    ap.add_argument("--update", action="store_true", help="also time the update block")
    ap.add_argument("--only", default=None, help="comma list of sections: lookup")
    ap.add_argument("--l2-fetch", type=int, default=0, help="experiment: cudaLimitMaxL2FetchGranularity (32/64/128)")
    a = ap.parse_args()
    if a.l2_fetch:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        torch.cuda.init()
        print("cudaDeviceSetLimit(L2 fetch granularity, %d) ->" % a.l2_fetch, rt.cudaDeviceSetLimit(5, ctypes.c_size_t(a.l2_fetch)))
    torch.manual_seed(0)
    dev = "cuda"
    B, D, H, W, Dg = a.B, 96, 96, 312, 48
    N = B * H * W
    res = {}

    def rec(name, med_us, best_us, bytes_):
        res[name] = dict(us_median=round(med_us, 2), us_best=round(best_us, 2), alg_MB=round(bytes_ / 1e6, 2),
                         GBps=round(bytes_ / med_us / 1e3, 1), frac_of_peak=round(bytes_ / med_us / 1e3 / PEAK, 3))
        print(name, res[name], flush=True)

    f1 = torch.randn(B, D, H, W, device=dev)
    f2 = torch.randn(B, D, H, W, device=dev)
    if a.only == "corr3":
        # all-pairs correlation GEMM at BASELINE config 3 (Middlebury-F 1984x2880 -> 496x720, D=256, 4 levels)
        Br, Dr, Hr, Wr = 1, 256, 496, 720
        r1 = torch.randn(Br, Dr, Hr, Wr, device=dev)
        r2 = torch.randn(Br, Dr, Hr, Wr, device=dev)
        flops = 2.0 * Br * Hr * Wr * Wr * Dr
        byts = 2 * 4 * Br * Dr * Hr * Wr + 4 * Br * Hr * Wr * Wr * (1 + 0.5 + 0.25 + 0.125)
        for mode in ("fp32", "bf16x3", "bf16"):
            A.set_corr_mode(mode)
            med, best = timeit(lambda: A.geometry._build_corr_levels(r1, r2, 4), reps=10)
            rec("corr_build_c3_" + mode, med, best, byts)
            res["corr_build_c3_" + mode]["TFLOPs_logical"] = round(flops / med / 1e6, 1)
            res["corr_build_c3_" + mode]["TFLOPs_issued"] = round(flops * (3 if mode == "bf16x3" else 1) / med / 1e6, 1)
            print("   ", res["corr_build_c3_" + mode])
        A.set_corr_mode("fp32")
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
        return
    if a.only == "train":
        # a13-vi kernels of the tensor-core training path at BASELINE config 5 (8 pairs of 320x736 -> 80x184)
        from anystereo_b200 import update_train as T
        Bt, Ht, Wt = a.B, 80, 184
        Nt = Bt * Ht * Wt
        A.set_update_engine("bf16x3")
        shapes = {"gru04.zr": ([128, 128, 128], 256, 3), "gru04.q": ([128, 128, 128], 128, 3), "dh1": ([128], 256, 3),
                  "conv": ([128], 127, 3), "convc2": ([64], 64, 3), "convc1": ([162], 64, 1), "dh2": ([256], 1, 3)}
        for name, (chans, Cout, k) in shapes.items():
            conv = torch.nn.Conv2d(sum(chans), Cout, k, padding=k // 2).to(dev)
            xs = [torch.randn(Bt, Ht, Wt, c, device=dev) for c in chans]
            pitch = (Cout + 63) // 64 * 64 if Cout >= 32 else Cout
            dy = torch.zeros(Bt, Ht, Wt, pitch, device=dev)
            dy[..., :Cout] = torch.randn(Bt, Ht, Wt, Cout, device=dev)
            c = T._Conv([conv])
            srcs = [T._src(x) for x in xs]
            flops = 2.0 * k * k * Cout * sum(chans) * Nt
            byts = 4 * Nt * (sum(chans) + Cout) + 4 * k * k * Cout * sum(chans)      # fp32 operands read once + dW
            for label, on in (("wgrad_tcgen05_incl_transposes", True), ("wgrad_cuda_cores", False)):
                T.set_wgrad_tensor_cores(on)
                T._KNOBS["small_tc"] = on
                med, best = timeit(lambda: c.wgrad(Bt, Ht, Wt, srcs, dy, pitch), reps=5 if not on else 10)
                rec("%s_%s" % (name, label), med, best, byts)
                res["%s_%s" % (name, label)]["TFLOPs_logical"] = round(flops / med / 1e6, 1)
            T.set_wgrad_tensor_cores(True)
            T._KNOBS["small_tc"] = True
        A.set_update_engine("fp32")
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
        return
    if a.only == "calib":
        # calibration of the box: what plain streaming achieves (same timing harness, L2 flushed)
        n = 1 << 28                                   # 1 GiB of fp32
        x = torch.empty(n, device=dev)
        y = torch.empty(n, device=dev)
        med, best = timeit(lambda: y.copy_(x))
        rec("CALIB_copy_1GiB(read+write)", med, best, 8 * n)
        med, best = timeit(lambda: y.zero_())
        rec("CALIB_memset_1GiB(write only)", med, best, 4 * n)
        med, best = timeit(lambda: torch.sum(x))
        rec("CALIB_sum_1GiB(read only)", med, best, 4 * n)
        n2 = 75 * (1 << 20)                           # 300 MB: the size of one correlation volume
        x2 = torch.empty(n2, device=dev); y2 = torch.empty(n2, device=dev)
        med, best = timeit(lambda: y2.copy_(x2))
        rec("CALIB_copy_300MB(read+write)", med, best, 8 * n2)
        med, best = timeit(lambda: y2.zero_())
        rec("CALIB_memset_300MB(write only)", med, best, 4 * n2)
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
        return
    gwc = A.build_gwc_volume(f1, f2, Dg, 8)
    if a.only == "mem":
        # the memory-bound volume kernels only, few repetitions (the command ncu wraps)
        blk = A.Combined_Geo_Encoding_Volume(f1, f2, gwc, num_levels=2, radius=4)
        coords = torch.arange(W, device=dev, dtype=torch.float32).reshape(1, 1, W, 1).repeat(B, H, 1, 1)
        d = (torch.rand(B, 1, H, W, device=dev) * Dg).contiguous()
        A.set_update_engine("f16f8")
        with A._lib.operand_format_scope(A.update_umma._sixteen_bit_format()):
            w_hi, w_lo = A.geometry.DeferredGeoLookup.pack_convc1_weight(torch.randn(64, 162, 1, 1, device="cuda") * 0.1, True)
        bias = torch.randn(64, device="cuda")
        o_hi = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
        o_lo = torch.empty_like(o_hi)
        for _ in range(3):
            A.build_gwc_volume(f1, f2, Dg, 8)
            A.geometry._build_geo_levels(gwc, 2)
            blk(d, coords)
            blk.deferred(d, coords).convc1_planes(w_hi, w_lo, bias, o_hi, o_lo)
        torch.cuda.synchronize()
        A.set_update_engine("fp32")
        return
    if a.only == "initdisp":
        # classifier Conv3d + softmax + disparity_regression fused; bytes = geo read once + disp written
        geo = torch.randn(B, 8, Dg, H, W, device=dev)
        wcl = torch.randn(1, 8, 3, 3, 3, device=dev) * 0.2
        med, best = timeit(lambda: A.init_disparity(geo, wcl))
        rec("init_disparity_fused", med, best, 4 * B * 8 * Dg * H * W + 4 * N)
        torch.backends.cudnn.allow_tf32 = False
        def ref():
            p = torch.softmax(torch.nn.functional.conv3d(geo, wcl, padding=1).squeeze(1), dim=1)
            return (p * torch.arange(Dg, device=dev).view(1, Dg, 1, 1)).sum(1, keepdim=True)
        med, best = timeit(ref)
        rec("TORCH_same_gpu_conv3d_softmax_regression", med, best, 4 * B * 8 * Dg * H * W + 4 * N)
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
        return
    if a.only == "instnorm":
        # InstanceNorm2d + ReLU (+ residual add + ReLU) on channels-last tensors at the sizes RAFT's fnet runs
        # them for one 384x1248 pair (both images in one batch).  Algorithmic bytes: x read twice (statistics, apply) + written
        # once (+ the residual read once).  Beside it: the ATen formulation the reference executes.
        from anystereo_b200 import extractor
        import torch.nn.functional as F
        for (Bn, Cn, Hn, Wn) in ((2, 64, 384, 1248), (2, 96, 192, 624), (2, 128, 96, 312)):
            x = torch.randn(Bn, Cn, Hn, Wn, device=dev).contiguous(memory_format=torch.channels_last)
            r = torch.randn(Bn, Cn, Hn, Wn, device=dev).contiguous(memory_format=torch.channels_last)
            nb = 4 * x.numel()
            tag = "%dx%dx%dx%d" % (Bn, Cn, Hn, Wn)
            med, best = timeit(lambda: extractor._instnorm_(x, 1e-5, relu=True))
            rec("instnorm_relu_" + tag, med, best, 3 * nb)
            med, best = timeit(lambda: extractor._instnorm_(x, 1e-5, relu=True, resid=r))
            rec("instnorm_relu_add_relu_" + tag, med, best, 4 * nb)
            xn = x.contiguous()
            rn = r.contiguous()
            med, best = timeit(lambda: F.relu(F.instance_norm(xn, eps=1e-5), inplace=True))
            rec("ATEN_instance_norm_relu_nchw_" + tag, med, best, 3 * nb)
            med, best = timeit(lambda: F.relu(rn + F.relu(F.instance_norm(xn, eps=1e-5), inplace=True)))
            rec("ATEN_instance_norm_relu_add_relu_nchw_" + tag, med, best, 4 * nb)
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
        return
    if a.only == "stem":
        # build_gwc_volume + corr_stem (Conv3d 8->8 + eval BatchNorm3d + LeakyReLU) + FeatureAtt multiply fused;
        # bytes = the two feature maps read once + the stem output written once (the GWC volume is never stored)
        fl = torch.randn(B, 96, H, W, device=dev) * 0.3
        fr = torch.randn(B, 96, H, W, device=dev) * 0.3
        wst = torch.randn(8, 8, 3, 3, 3, device=dev) * 0.2
        sc, sh = torch.rand(8, device=dev) + 0.5, torch.randn(8, device=dev) * 0.1
        att = torch.sigmoid(torch.randn(B, 8, H, W, device=dev))
        alg = 4 * (2 * B * 96 * H * W + B * 8 * Dg * H * W + B * 8 * H * W)
        med, best = timeit(lambda: A.gwc_corr_stem(fl, fr, Dg, 8, wst, sc, sh, 0.01, att))
        rec("gwc_corr_stem_fused", med, best, alg)
        def chain():
            v = A.build_gwc_volume(fl, fr, Dg, 8)
            v = torch.nn.functional.conv3d(v, wst, padding=1)
            v = torch.nn.functional.leaky_relu(v * sc.view(1, 8, 1, 1, 1) + sh.view(1, 8, 1, 1, 1), 0.01)
            return v * att.unsqueeze(2)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            med, best = timeit(chain)
            rec("UNFUSED_gwc_kernel_then_torch_conv3d_affine_lrelu_mul_%s" % ("tf32" if tf32 else "fp32"), med, best, alg)
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
        return
    if a.only == "lookup":
        blk = A.Combined_Geo_Encoding_Volume(f1, f2, gwc, num_levels=2, radius=4)
        coords = torch.arange(W, device=dev, dtype=torch.float32).reshape(1, 1, W, 1).repeat(B, H, 1, 1)
        d = (torch.rand(B, 1, H, W, device=dev) * Dg).contiguous()
        med, best = timeit(lambda: blk(d, coords))
        rec("geo_lookup_uniform", med, best, 1372 * N)
        fused_lookup_bench(blk, d, coords, rec, N)
        return
    # a7 GWC
    med, best = timeit(lambda: A.build_gwc_volume(f1, f2, Dg, 8))
    rec("gwc_build", med, best, 2 * 4 * B * D * H * W + 4 * B * 8 * Dg * H * W)
    # a1+a2 corr pyramid, per mode
    for mode in ("fp32", "bf16x3", "bf16"):
        try:
            A.set_corr_mode(mode)
            med, best = timeit(lambda: A.geometry._build_corr_levels(f1, f2, 2))
            rec("corr_build_" + mode, med, best, 2 * 4 * B * D * H * W + 4 * B * H * W * W * 1.5)
            res["corr_build_" + mode]["TFLOPs"] = round(2.0 * B * H * W * W * D / med / 1e6, 2)
        except RuntimeError as e:
            print("corr mode", mode, "unavailable:", str(e)[:80])
    A.set_corr_mode("fp32")
    # a2 geo pyramid
    med, best = timeit(lambda: A.geometry._build_geo_levels(gwc, 2))
    rec("geo_pyramid", med, best, 4 * B * 8 * Dg * H * W * 2.5)
    blk = A.Combined_Geo_Encoding_Volume(f1, f2, gwc, num_levels=2, radius=4)
    coords = torch.arange(W, device=dev, dtype=torch.float32).reshape(1, 1, W, 1).repeat(B, H, 1, 1)
    xs = torch.arange(W, device=dev).view(1, 1, 1, W)
    ys = torch.arange(H, device=dev).view(1, 1, H, 1)
    disps = {
        "uniform": torch.rand(B, 1, H, W, device=dev) * Dg,
        "smooth": 20 + 10 * torch.sin(2 * 3.14159265 * xs / W) * torch.cos(2 * 3.14159265 * ys / H)
                  + 0.5 * torch.randn(B, 1, H, W, device=dev),
    }
    for name, d in disps.items():
        d = d.float().contiguous()
        med, best = timeit(lambda: blk(d, coords))
        rec("geo_lookup_" + name, med, best, 1372 * N)
    fused_lookup_bench(blk, disps["uniform"].float().contiguous(), coords, rec, N)
    # a6 sampler on the level-0 volume
    vol = blk.init_corr_pyramid[0].reshape(B, H, W, W)
    x0 = (coords.reshape(B, 1, H, W) - disps["uniform"]).contiguous()
    med, best = timeit(lambda: A.corr_sampler.forward(vol, x0, 4))
    rec("sampler_fwd", med, best, 80 * N)
    g = torch.randn(B, 9, H, W, device=dev)
    med, best = timeit(lambda: A.corr_sampler.backward(vol, x0, g, 4))
    rec("sampler_bwd", med, best, (36 + 4) * N + 4 * N * W)
    try:   # the reference's own CUDA kernels recompiled for sm_100a: the "kernel to beat"
        from oracle import ref_sampler
        rs = ref_sampler.load()
        if rs is not None:
            x2 = torch.cat([x0, torch.zeros_like(x0)], 1).contiguous()
            med, best = timeit(lambda: rs.forward(vol, x2, 4))
            rec("REFERENCE_sampler_fwd", med, best, 80 * N)
            med, best = timeit(lambda: rs.backward(vol, x2, g, 4))
            rec("REFERENCE_sampler_bwd", med, best, (36 + 4) * N + 4 * N * W)
    except Exception as e:
        print("reference sampler unavailable:", e)
    # a6 at the size SURVEY 8(d) names: config-3 level-0 volume [1,496,720,720] (1.03 GB), 357,120 pixels x 80 B = 28.6 MB
    del vol
    Bs, Hs, Ws = 1, 496, 720
    vol3 = torch.randn(Bs, Hs, Ws, Ws, device=dev)
    c3 = torch.arange(Ws, device=dev, dtype=torch.float32).reshape(1, 1, 1, Ws).repeat(Bs, 1, Hs, 1)
    x3 = (c3 - torch.rand(Bs, 1, Hs, Ws, device=dev) * 64).contiguous()
    N3 = Bs * Hs * Ws
    med, best = timeit(lambda: A.corr_sampler.forward(vol3, x3, 4))
    rec("sampler_fwd_c3_level0", med, best, 80 * N3)
    g3 = torch.randn(Bs, 9, Hs, Ws, device=dev)
    med, best = timeit(lambda: A.corr_sampler.backward(vol3, x3, g3, 4))
    rec("sampler_bwd_c3_level0", med, best, (36 + 4) * N3 + 4 * N3 * Ws)
    try:
        from oracle import ref_sampler
        rs = ref_sampler.load()
        if rs is not None:
            x32 = torch.cat([x3, torch.zeros_like(x3)], 1).contiguous()
            med, best = timeit(lambda: rs.forward(vol3, x32, 4))
            rec("REFERENCE_sampler_fwd_c3_level0", med, best, 80 * N3)
            med, best = timeit(lambda: rs.backward(vol3, x32, g3, 4))
            rec("REFERENCE_sampler_bwd_c3_level0", med, best, (36 + 4) * N3 + 4 * N3 * Ws)
    except Exception as e:
        print("reference sampler unavailable:", e)
    del vol3, g3
    # a3 RAFT lookup at config 3 (496x720, L=4), D reduced to keep the build short: only the lookup is timed
    Br, Hr, Wr = 1, 496, 720
    r1 = torch.randn(Br, 64, Hr, Wr, device=dev)
    r2 = torch.randn(Br, 64, Hr, Wr, device=dev)
    rb = A.CorrBlock1D(r1, r2, num_levels=4, radius=4)
    rc = torch.arange(Wr, device=dev, dtype=torch.float32).reshape(1, 1, Wr, 1).repeat(Br, Hr, 1, 1)
    rd = (torch.rand(Br, 1, Hr, Wr, device=dev) * 64).contiguous()
    med, best = timeit(lambda: rb(rd, rc))
    rec("raft_lookup_c3", med, best, 308 * Br * Hr * Wr)
    del rb, r1, r2
    if a.update:
        args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
        for engine in ("fp32", "bf16x3", "bf16"):
            try:
                A.set_update_engine(engine)
                m = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128]).cuda().eval()
                sizes = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]
                net = [torch.tanh(torch.randn(B, 128, h, w, device=dev)) for h, w in sizes]
                inp = [[torch.randn(B, 128, h, w, device=dev) for _ in range(3)] for h, w in sizes]
                feat = blk(disps["uniform"].contiguous(), coords)
                dsp = disps["uniform"].contiguous()
                with torch.no_grad():
                    med, best = timeit(lambda: m(list(net), inp, feat, dsp), reps=5, warm=2, flush=False)
                flops = 2.0 * (N * (1847488 + 64 * 162) + N / 4 * 1327104 + N / 16 * 884736)
                res["update_block_" + engine] = dict(us_median=round(med, 1), TFLOPs=round(flops / med / 1e6, 2))
                print("update_block_" + engine, res["update_block_" + engine], flush=True)
            except (RuntimeError, NotImplementedError, ImportError) as e:
                print("update engine", engine, "unavailable:", str(e)[:100])
        A.set_update_engine("fp32")
    if a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
