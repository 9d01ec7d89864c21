"""BASELINE.json config 5 structure: one training step of the IGEV hot path on synthetic 320x736 crops
(80x184 at 1/4), `--batch` pairs per GPU, `--iters` unrolled iterations, sequence loss, gradient all-reduce
over NCCL (one process per GPU, launched with torchrun), AdamW on the update block.

Everything between the backbone outputs and the low-resolution disparities runs on this library's kernels
(forward and backward); backbones / hourglass / LIIF are outside the path, so their outputs are synthetic
leaf tensors that receive gradients.  Prints one JSON line with the step time (max over ranks)."""
import argparse
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import anystereo_b200 as A  # noqa: E402


def run(engine="fp32", batch=8, iters=16, steps=8, H=80, W=184, dev=None, world=1, rank=0, overlap=True):
    """`steps` training steps of the IGEV hot path (structure of train_continuous_IGEV.py:186-240); returns the
    per-step device times (ms) with a per-phase split, losses and gradient norms.  The first two steps are warm-up
    (weight packing, allocator growth: 2.2 s and 0.3 s were seen); the reported step time is the MEDIAN of the rest.
    The update engine / correlation mode are restored."""
    prev_engine, prev_corr = A.get_update_engine(), A.get_corr_mode()
    try:
        A.set_update_engine(engine)
        A.set_corr_mode("fp32")
        return _run(engine, batch, iters, steps, H, W, dev, world, rank, overlap)
    finally:
        A.set_update_engine(prev_engine)
        A.set_corr_mode(prev_corr)


def _median(v):
    v = sorted(v)
    return v[len(v) // 2] if len(v) % 2 else 0.5 * (v[len(v) // 2 - 1] + v[len(v) // 2])


def _run(engine, B, iters, steps, H, W, dev, world, rank, overlap):
    torch.manual_seed(0)                                    # identical replicas
    torch.cuda.empty_cache()                                # the caller's cached blocks have other sizes (bench.py)
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    block = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128]).to(dev).train()
    opt = torch.optim.AdamW(block.parameters(), lr=2e-4, weight_decay=1e-5, eps=1e-8)   # train_continuous_IGEV.py:127
    # bucket order = the order in which the first iteration's backward completes the gradients
    order = [p for name in ("disp_head", "gru04", "encoder", "gru08", "gru16") for p in getattr(block, name).parameters()]
    reducer = A.GradientAllReducer(order) if (world > 1 and overlap) else None
    # every PAIR has its own generator (seeded by its global index): rank r owns pairs [r*B, (r+1)*B), so N ranks x B pairs
    # and one rank x N*B pairs see the same global batch (the loss / gradient-norm equality check of SURVEY 4 item 5)
    gens = [torch.Generator(device="cpu").manual_seed(100 + rank * B + i) for i in range(B)]
    sizes = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]

    def per_pair(fn):
        return torch.cat([fn(g) for g in gens], 0).to(dev)

    def leaf(*shape, scale=1.0):
        return per_pair(lambda g: torch.randn(1, *shape, generator=g) * scale).requires_grad_(True)

    f1, f2 = leaf(96, H, W), leaf(96, H, W)
    geo = leaf(8, 48, H, W)
    net0 = [per_pair(lambda g: torch.tanh(torch.randn(1, 128, h, w, generator=g))).requires_grad_(True) for h, w in sizes]
    inp = [[leaf(128, h, w, scale=0.5) for _ in range(3)] for h, w in sizes]
    init_disp = per_pair(lambda g: torch.rand(1, 1, H, W, generator=g) * 40)
    gt = per_pair(lambda g: torch.rand(1, 1, H, W, generator=g) * 48)
    coords = torch.arange(W, device=dev, dtype=torch.float32).reshape(1, 1, W, 1).repeat(B, H, 1, 1)
    times, losses, gnorms, phases = [], [], [], []
    leaves = [f1, f2, geo] + net0 + [t for l in inp for t in l]
    for step in range(steps):
        for t in leaves:
            t.grad = None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        opt.zero_grad(set_to_none=True)
        fn = A.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
        net, disp, loss = list(net0), init_disp, 0.0
        for it in range(iters):
            disp = disp.detach()
            feat = fn(disp, coords)
            net, delta = block(net, inp, feat, disp)
            disp = disp + delta
            loss = loss + 0.9 ** (iters - 1 - it) * (disp - gt).abs().mean()
        ev[1].record()
        loss.backward()
        ev[2].record()
        if reducer is not None:
            reducer.finish()                               # buckets launched from the hooks during backward
        else:
            A.allreduce_gradients(list(block.parameters()))
        ev[3].record()
        gn = torch.nn.utils.clip_grad_norm_(block.parameters(), 1.0)       # train_continuous_IGEV.py:234
        opt.step()
        ev[4].record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev[0].elapsed_time(ev[4])] + [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = [float(v) for v in ms]
        times.append(ms[0])
        phases.append(ms[1:])
        lt = loss.detach().clone().reshape(1)
        if world > 1:
            dist.all_reduce(lt)                           # mean over ranks = the loss of the global batch
            lt /= world
        losses.append(float(lt))
        gnorms.append(float(gn))
    if reducer is not None:
        reducer.remove()
    steady = list(range(min(2, steps - 1), steps))
    med = _median([times[i] for i in steady])
    ph = [_median([phases[i][k] for i in steady]) for k in range(4)]
    return {"config": "IGEV hot-path training step, %dx%d (1/4: %dx%d), batch %d/GPU, %d iters, update engine %s"
                      % (4 * H, 4 * W, H, W, B, iters, engine),
            "n_gpus": world, "ms_per_step": times, "ms_per_step_median": med, "steady_steps": len(steady),
            "phase_ms_median": {"forward": ph[0], "backward": ph[1], "allreduce_wait": ph[2], "clip_and_adamw": ph[3]},
            "allreduce": ("bucketed NCCL all-reduce launched from post-accumulate hooks during backward, waited after it"
                          if reducer is not None else "bucketed NCCL all-reduce after backward") if world > 1 else "none (1 GPU)",
            "pairs_per_s": world * B / (med / 1e3),
            "loss": losses, "grad_norm_after_allreduce": gnorms,
            "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=16)       # train_iters, train_continuous_IGEV.py:297
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--engine", default="fp32", choices=["fp32", "bf16x3", "bf16", "fp16", "f16f8"],
                    help="update-block engine: fp32 = CUDA cores; others = forward, data and weight gradients on tcgen05")
    ap.add_argument("--h", type=int, default=80)
    ap.add_argument("--w", type=int, default=184)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run(a.engine, a.batch, a.iters, a.steps, a.h, a.w, dev, world, rank, overlap=not a.no_overlap)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
