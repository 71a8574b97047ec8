#!/usr/bin/env python
"""bench.py — contrastive-loss forward+backward throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

One "step" = forward+backward of ALL THREE losses (sent_loss D=256, img_loss D=512, word_loss
T=18 R=17x17 D=256) on one synthetic COCO-shaped batch of 256 per GPU (BASELINE config 2;
with --gpus N>1 config 3: global negatives, 256 per GPU, rows local / columns all-gathered).
Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same
metric through the public API with pinned-host inputs copied H2D and the loss read back D2H inside
the timed region.  `roofline` is for the dominant kernel (word-region backward), timed live with
CUDA events on the launching stream.  `cpu_baseline` / `--impl reference` time the CPU oracle
(the reference's own functions restated, plus this repo's word-loss restatement) on host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_LOCAL, D_SENT, D_IMG, D_WORD, T_WORDS, R_SIDE = 256, 256, 512, 256, 18, 17
RHO = (5.0, 5.0, 10.0)


def workload_name(world):
    """config.workload, the same string in both arms (the driver compares them)."""
    return ("COCO-256 (BASELINE config %d): sent_loss D=256 + img_loss D=512 + word_loss T=18 R=289 D=256, "
            "fwd+bwd, batch 256/GPU" % (2 if world == 1 else 3))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def make_inputs(B, seed, dtype=torch.float32):
    """Synthetic COCO-shaped batch, generated on CPU so oracle and GPU see identical values."""
    g = torch.Generator().manual_seed(seed)
    R = R_SIDE * R_SIDE
    words = torch.randn(B, D_WORD, T_WORDS, generator=g)
    regions = torch.randn(B, D_WORD, R, generator=g) + 0.3 * words[:, :, torch.randint(0, T_WORDS, (R,), generator=g)]
    lens = torch.randint(5, T_WORDS + 1, (B,), generator=g)
    d = {
        "sent": torch.randn(B, D_SENT, generator=g),
        "img": torch.randn(B, D_SENT, generator=g),
        "real": torch.randn(B, D_IMG, generator=g),
        "fake": torch.randn(B, D_IMG, generator=g),
        "words": words,
        "regions": regions.view(B, D_WORD, R_SIDE, R_SIDE),
        "mask": torch.arange(T_WORDS).unsqueeze(0) >= lens.unsqueeze(1),
    }
    return {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in d.items()}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.25)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        rows = [r for r in self.rows if len(r) == 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def oracle_step(inp, cols=None):
    """All three losses fwd+bwd with the CPU oracle (float32, as the reference would run).

    inp: the ROW side (one rank's batch).  cols (N > 1): the column operands of the global batch — sentence / fake-image
    embeddings, words and masks of all ranks, rank 0's first — so the step is rank 0's share of the sharded problem
    (B rows x B_global columns, identity labels with offset 0).  -> (loss, [d img, d sent, d fake, d words, d regions])."""
    import oracle
    B = inp["sent"].shape[0]
    leaf = lambda x: x.clone().requires_grad_()
    c = inp if cols is None else cols
    Bk = c["sent"].shape[0]
    labels = torch.zeros(B, Bk)
    labels.diagonal().fill_(1.0)                              # make_labels(b_global=False) (train_gan.py:74), rank 0's rows
    i_, s_, f_, w_, v_ = leaf(inp["img"]), leaf(c["sent"]), leaf(c["fake"]), leaf(c["words"]), leaf(inp["regions"])
    loss = (oracle.sent_loss(i_, s_, labels, False) + oracle.img_loss(inp["real"], f_, labels, False)
            + oracle.word_loss(v_, w_, c["mask"], labels, False, 0.5, *RHO, img_block=8))
    loss.backward()
    return float(loss.detach()), [t.grad for t in (i_, s_, f_, w_, v_)]


def oracle_inputs(B, world, precision):
    """The GPU arm's own inputs (same seeds), rounded to bf16 when the GPU arm computes on bf16 operands:
    BASELINE.md section 4 — both arms see bit-identical tensors."""
    rd = (lambda v: v.bfloat16().float()) if precision == "bf16" else (lambda v: v)
    ranks = [{k: (rd(v) if v.dtype.is_floating_point else v) for k, v in make_inputs(B, 1000 + r).items()} for r in range(world)]
    cols = None if world == 1 else {k: torch.cat([x[k] for x in ranks]) for k in ("sent", "fake", "words", "mask")}
    return ranks[0], cols


def time_oracle(steps, warmup, B, world=1, precision="fp32"):
    """-> (samples/s of the global batch, s per step, outputs of the first step).  With world > 1 one step is rank 0's
    share (B rows x B*world columns), the same per-rank problem the GPU arm times."""
    torch.set_num_threads(os.cpu_count())
    inp, cols = oracle_inputs(B, world, precision)
    first = None
    for _ in range(warmup):
        out = oracle_step(inp, cols)
        first = first or out
    t0 = time.perf_counter()
    for _ in range(steps):
        out = oracle_step(inp, cols)
        first = first or out
    dt = (time.perf_counter() - t0) / steps
    return B * world / dt, dt, first


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the newest committed `ncu --set full` summary
    (profiles/rNN_ncu_wr_tc_summary.txt, written by profiles/ncu_summary.py) -> (bytes, file name) or (None, None)."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_wr_tc_summary.txt")), reverse=True):
        block, total = False, {}
        for line in open(path):
            if line.startswith("-- "):
                block = kernel in line
            elif block:
                m = re.match(r"\s+(dram__bytes_(?:read|write)\.sum)\s+([0-9.]+)\s+(\w+)", line)
                if m:
                    total[m.group(1)] = float(m.group(2)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(3)]
        if len(total) == 2:
            return sum(total.values()), os.path.relpath(path, ROOT)
    return None, None


def run_reference(args):
    """--impl reference: the CPU path on host cores, same metric/config, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    steps, warmup = max(1, min(args.steps, 3 if world == 1 else 1)), max(1, min(args.warmup, 1))
    precision = args.precision or os.environ.get("XMC_BENCH_PRECISION", "bf16")
    sps, dt, _ = time_oracle(steps, warmup, args.batch, world, precision)
    cores = os.cpu_count()
    sample = (f"rank 0's share of the COCO-256 step (B={args.batch} rows x {args.batch * world} columns, T=18, R=289, D=256; "
              f"sent D=256, img D=512), fp32 arithmetic on the GPU arm's inputs, "
              f"{steps} timed + {warmup} warm-up fwd+bwd steps; reference functions restated "
              f"(train_gan.py:72-139) + this repo's word-loss restatement (reference has none)")
    line = {
        "impl": "reference", "metric": "contrastive-loss fwd+bwd samples/sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(world), "arithmetic": "CPU oracle, fp32",
                   "global_batch": args.batch * world, "pairs_per_s": args.batch * args.batch * world / dt,
                   "per_rank_problem": f"{args.batch} rows x {args.batch * world} columns"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_step_workload(args):
    """--workload step (BASELINE config 4): the reference's G/D update (train_gan.py:187-289 as
    xmc_gan_b200.step.gd_step) around DF-GAN-shaped networks at IMG.SIZE 256, NCH 32, NEF 256, IMG_MATCH, MA-GP on,
    synthetic images / captions — once with the stock PyTorch loss block, once with this package's ops swapped in,
    each with and without the word loss the reference leaves unimplemented.  One GPU; prints one JSON line."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stock_losses
    from dfgan_harness import NetD, NetG
    from xmc_gan_b200 import step as S
    from xmc_gan_b200 import train_gan as T
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, size, nch, nef, T_ = args.batch if args.batch != B_LOCAL else 64, 256, 32, 256, T_WORDS
    torch.manual_seed(0)
    G = NetG(size, nch, 100, nef, nef).to(dev)
    D = NetD(size, nch, nef, img_match=True, spec_norm=True, region_res=16).to(dev)
    optG = torch.optim.Adam(G.parameters(), 1e-4, betas=(0.0, 0.9))
    optD = torch.optim.Adam(D.parameters(), 4e-4, betas=(0.0, 0.9))
    g = torch.Generator().manual_seed(1)
    imgs = (torch.rand(B, 3, size, size, generator=g) * 2 - 1).to(dev)
    words = torch.randn(B, nef, T_, generator=g).to(dev)
    sent = torch.randn(B, nef, generator=g).to(dev)
    lens = torch.randint(5, T_ + 1, (B,), generator=g)
    mask = (torch.arange(T_)[None] >= lens[:, None]).to(dev)
    W, K = max(args.warmup, 3), args.steps
    res = {}
    for word in (False, True):
        variants = [("stock", stock_losses, None, False), ("xmc_gan_b200", T, {"precision": "bf16"}, False)]
        if word:       # N2: the region head inside the word loss's prologue instead of a convolution of the network
            variants.append(("xmc_gan_b200_fused_head", T, {"precision": "bf16"}, True))
        for name, ns, wkw, fused in variants:
            cfg = S.default_step_cfg()
            cfg.TRAIN.ENCODER_LOSS.WORD = word
            def one():
                noise = torch.randn(B, 100, device=dev)
                return S.gd_step(G, D, optG, optD, imgs, words, sent, mask, noise, cfg=cfg, losses=ns, word_kwargs=wkw,
                                 fused_region_head=fused)
            for _ in range(W):
                out = one()
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for _ in range(K):
                out = one()
            eb.record()
            torch.cuda.synchronize()
            res[("word" if word else "no_word", name)] = {"ms_per_step": ea.elapsed_time(eb) / K,
                                                          "scalars": {k: float(v) for k, v in out.items()}}
    line = {"workload": "G/D step (BASELINE config 4): DF-GAN-shaped netG/netD 256x256, NCH 32, NEF 256, IMG_MATCH, spectral norm, "
                        "MA-GP, sent_loss + img_loss%s, batch %d, fp32 networks" % (" (+ word_loss T=18, R=16x16, bf16 tcgen05 path)", B),
            "n_gpus": 1, "steps": K, "warmup": W, "unit": "ms/step", "higher_is_better": False, "data": "synthetic",
            "sent+img": {n: res[("no_word", n)] for n in ("stock", "xmc_gan_b200")},
            "sent+img+word": {n: res[("word", n)] for n in ("stock", "xmc_gan_b200", "xmc_gan_b200_fused_head")}}
    for k in ("sent+img", "sent+img+word"):
        line[k]["speedup"] = line[k]["stock"]["ms_per_step"] / line[k]["xmc_gan_b200"]["ms_per_step"]
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=B_LOCAL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--one-stream", action="store_true", help="run the three losses back to back on one stream")
    ap.add_argument("--fused", default=None, choices=["on", "off"],
                    help="evaluate the three losses through train_gan.contrastive_losses (grouped collectives); default: on for N > 1")
    ap.add_argument("--graph-multi", default="on", choices=["on", "off"], help="capture the N > 1 step as a CUDA graph as well")
    ap.add_argument("--no-fp32-block", action="store_true", help="skip the fp32-mode block of the default (bf16) line")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run comparison with the CPU oracle / single-GPU run")
    ap.add_argument("--sustain-s", type=float, default=1.5, help="seconds of back-to-back steps for the `sustained` block")
    ap.add_argument("--workload", default="losses", choices=["losses", "step"],
                    help="losses: the contract's metric (default); step: the G/D training step with the ops swapped in (config 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "step":
        return run_step_workload(args)

    import torch.distributed as dist
    from xmc_gan_b200 import train_gan as T
    from xmc_gan_b200.ops import default_ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # quiet by default (rank 0 prints ONE JSON line); a caller's setting wins
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    dev = torch.device("cuda", local_rank)
    ops = default_ops()
    precision = args.precision or os.environ.get("XMC_BENCH_PRECISION", "bf16")
    in_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    B = args.batch
    W = max(args.warmup, 3)
    K = args.steps

    host = make_inputs(B, 1000 + rank, in_dtype)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    # The three losses are independent until they are summed: the two similarity losses run on side streams
    # (autograd replays each backward on its forward's stream), so their ~30 us kernels fill the tails of the
    # two big word-region kernels instead of queueing behind them.  One GPU only (NCCL order across ranks).
    fused = (args.fused == "on") if args.fused else world > 1
    side = [torch.cuda.Stream(device=dev) for _ in range(2)] if (world == 1 and not args.one_stream and not fused) else None
    ops.use_side_stream = not args.one_stream

    def step(x, use_side=True, grp=group, nb=B):
        leaf = lambda t: t.detach().requires_grad_()
        i_, s_, f_, w_, v_ = leaf(x["img"]), leaf(x["sent"]), leaf(x["fake"]), leaf(x["words"]), leaf(x["regions"])
        labels = T.make_labels(nb, x["sent"], False, group=grp)
        if fused:
            # one autograd function for the three losses: grouped all-gather / packet exchange / reduce-scatter,
            # similarity losses on side streams inside it
            l_sent, l_img, l_word = T.contrastive_losses(i_, s_, x["real"], f_, v_, w_, x["mask"], labels, False,
                                                         rho1=RHO[0], rho2=RHO[1], rho3=RHO[2], precision=precision, group=grp)
            loss = l_sent + l_img + l_word
            loss.backward()
            return loss, (i_.grad, s_.grad, f_.grad, w_.grad, v_.grad)
        forked = side is not None and use_side
        if not forked:
            l_sent = T.sent_loss(i_, s_, labels, False, group=grp)
            l_img = T.img_loss(x["real"], f_, labels, False, group=grp)
        else:
            cur = torch.cuda.current_stream(dev)
            side[0].wait_stream(cur); side[1].wait_stream(cur)
            with torch.cuda.stream(side[0]):
                l_sent = T.sent_loss(i_, s_, labels, False, group=grp)
            with torch.cuda.stream(side[1]):
                l_img = T.img_loss(x["real"], f_, labels, False, group=grp)
        l_word = T.word_loss(v_, w_, x["mask"], labels, False, rho1=RHO[0], rho2=RHO[1], rho3=RHO[2],
                             precision=precision, group=grp)
        if forked:
            cur.wait_stream(side[0]); cur.wait_stream(side[1])
            l_sent.record_stream(cur); l_img.record_stream(cur)      # allocated on the side streams, consumed here
        loss = l_sent + l_img + l_word
        loss.backward()
        return loss, (i_.grad, s_.grad, f_.grad, w_.grad, v_.grad)

    def barrier():
        if world > 1:
            dist.barrier(group)
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Per-step CUDA events on the current stream, L2 flushed (untimed) between steps."""
        evs = []
        barrier()
        for i in range(steps):
            # Keep the host at most two steps ahead of the GPU, as a training loop that reads its loss does.
            # Unbounded, the eager N > 1 loop runs ~17 steps ahead (launch-queue depth) and the caching
            # allocator answers the growing set of in-flight buffers with cudaMalloc/cudaFree stalls of
            # ~100 ms (profiles/exp_n2_steps.py); the GPU never idles with two steps queued.
            if i >= 2:
                evs[i - 2][1].synchronize()
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            ms = float(t)
        return ms

    # ---- per-kernel CUDA-event times (eager launches; feeds the roofline) --------------------
    for _ in range(W):
        step(devin)

    # ---- parity of THIS run's numbers (both arms on identical tensors) -------------------
    # N = 1: the step's loss and five gradients are kept and compared below with the CPU oracle run on the same inputs.
    # N > 1: every rank's inputs are gathered, rank 0 evaluates the whole global batch on its own GPU (no group) and
    # compares the sharded step's global loss and its own gradient rows with it (SURVEY section 4's definition).
    nerr = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
    parity, gpu_out = None, None
    if not args.no_parity:
        loss_g, grads_g = step(devin)
        torch.cuda.synchronize()
        gpu_out = (float(loss_g.detach()), [g.detach().float().cpu() for g in grads_g])
        if world > 1:
            from xmc_gan_b200 import losses as L_
            full = {}
            for k, v in devin.items():
                u = v.to(torch.uint8) if v.dtype == torch.bool else v
                out = torch.empty((world * u.shape[0],) + tuple(u.shape[1:]), device=dev, dtype=u.dtype)
                dist.all_gather_into_tensor(out, u.contiguous(), group=group)
                full[k] = out.bool() if v.dtype == torch.bool else out
            if rank == 0:
                keep = L_.MAX_CONTEXT_BYTES
                L_.MAX_CONTEXT_BYTES = 1 << 40                      # a one-off check on a 180 GB device
                try:
                    loss_1, grads_1 = step(full, grp=None, nb=B * world)
                    torch.cuda.synchronize()
                    parity = {"against": "rank 0 alone on the gathered global batch (single-GPU path, same kernels, no group)",
                              "loss_rel": abs(gpu_out[0] - float(loss_1.detach())) / abs(float(loss_1.detach())),
                              "grad_nerr": [nerr(a, b[:B].float().cpu()) for a, b in zip(gpu_out[1], grads_1)],
                              "order": ["d img", "d sent", "d fake", "d words", "d regions"], "rows": "rank 0's"}
                finally:
                    L_.MAX_CONTEXT_BYTES = keep
                del loss_1, grads_1
            del full
            torch.cuda.empty_cache()
            dist.barrier(group)
            for _ in range(W):                                      # the caching allocator regrows its blocks
                step(devin)
    l0 = ops.launches
    ops.enable_timing(True)
    timed(lambda: step(devin, use_side=False), K)            # one stream: clean per-kernel event pairs
    kern = ops.kernel_ms()
    ops.enable_timing(False)
    launches = (ops.launches - l0) // K
    ms_eager = timed(lambda: step(devin), K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(min(K, 8)):                                # few enough that the launch queue never fills
        step(devin)
    host_ms = (time.perf_counter() - t0) / min(K, 8) * 1e3    # host time to ENQUEUE one step (python + launches)
    torch.cuda.synchronize()

    # ---- the step as a CUDA graph -----------------------------------------------------------------
    # The same public-API calls (make_labels / sent_loss / img_loss / word_loss + backward), captured once with
    # torch.cuda.graph and replayed: a step is ~35 launches of 2-700 us kernels and the host needs about as
    # long to enqueue them as the GPU needs to run them, so an eager loop measures the host.  Nothing in the
    # path synchronises with the host (the compacted word-row count stays on the device), which is what makes
    # the capture possible.  --no-graph (or a failed capture) falls back to eager launches.
    def capture(x_static):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                step(x_static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # N > 1: NCCL's watchdog thread polls events while this thread captures; only this thread's calls belong to the capture
        with torch.cuda.graph(g, **({"capture_error_mode": "thread_local"} if world > 1 else {})):
            loss, grads = step(x_static)
        return g, loss, grads

    graphs = None
    mode = "eager"
    # N > 1: the NCCL collectives are captured with the kernels (every rank captures and replays the same sequence);
    # whether the capture succeeded is agreed on by all ranks before anyone replays.
    if not args.no_graph and (world == 1 or args.graph_multi == "on"):
        try:
            sets = [devin, {k: v.clone() for k, v in devin.items()}]
            graphs = [capture(x) for x in sets]
            for g, _, _ in graphs:
                g.replay()
            torch.cuda.synchronize()
            ref_loss, ref_grads = step(devin)                   # the replay must reproduce the eager step
            for g_loss, g_grads in ((graphs[0][1], graphs[0][2]), (graphs[1][1], graphs[1][2])):
                ok = abs(float(g_loss.detach()) - float(ref_loss.detach())) <= 1e-4 * abs(float(ref_loss.detach()))
                for a, b in zip(g_grads, ref_grads):
                    ok = ok and float((a.float() - b.float()).norm()) <= 1e-2 * float(b.float().norm())   # atomic summation order
                if not ok:
                    raise RuntimeError("graph replay does not reproduce the eager step")
            mode = "cuda-graph"
        except Exception as e:                                  # noqa: BLE001 - any capture problem: measure eagerly
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()
        if world > 1:
            okf = torch.tensor([1.0 if graphs is not None else 0.0], device=dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN, group=group)
            if float(okf) == 0.0:
                graphs, mode = None, "eager"

    # ---- device-resident throughput ---------------------------------------------------------
    # Clocks are sampled by rank 0 only, for its own GPU: one nvidia-smi per rank every 100 ms takes the
    # driver's lock often enough to slow the eagerly launched N = 8 step by 4 % (8.30 vs 7.96 ms).
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    if graphs is not None:
        for _ in range(W):
            graphs[0][0].replay()
        ms = timed(graphs[0][0].replay, K)
    else:
        ms = timed(lambda: step(devin), K)
    clocks = sampler.summary() if sampler is not None else None

    # ---- sustained: back-to-back steps for >= --sustain-s seconds (no L2 flush, no per-step events) ----------------
    # The contract's K timed steps are a ~20 ms burst on a cold-ish GPU; this is the number after the clocks and the
    # power limiter have settled, with the median SM clock sampled over the same window.
    sustained = None
    if args.sustain_s > 0:
        run1 = graphs[0][0].replay if graphs is not None else (lambda: step(devin))
        n_sus = max(K, int(args.sustain_s * 1e3 / max(ms, 1e-3)) + 1)
        if world > 1:                                           # every rank must run the same number of steps
            tn = torch.tensor([n_sus], device=dev)
            dist.all_reduce(tn, op=dist.ReduceOp.MAX, group=group)
            n_sus = int(tn)
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        barrier()
        if sampler2 is not None:
            sampler2.start()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        last = None
        for i in range(n_sus):
            run1()
            if graphs is None and i % 2 == 1:                   # eager loop: keep the host at most ~2 steps ahead
                if last is not None:
                    last.synchronize()
                last = torch.cuda.Event()
                last.record()
        eb.record()
        barrier()
        ms_sus = ea.elapsed_time(eb) / n_sus
        if world > 1:
            t = torch.tensor([ms_sus], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            ms_sus = float(t)
        c2 = sampler2.summary() if sampler2 is not None else None
        sustained = {"ms_per_step": ms_sus, "value": B * world / (ms_sus * 1e-3), "steps": n_sus,
                     "seconds": ms_sus * n_sus * 1e-3, "sm_mhz_median": c2["sm_mhz"] if c2 else None,
                     "reasons": c2["reasons"] if c2 else None, "l2": "not flushed between steps"}

    # ---- end to end: pinned host -> device, loss -> host -------------------------------------
    # Every step copies ITS inputs from pinned host memory (copy stream, issued one step ahead into the other
    # of two input sets: what a DataLoader with pinned memory and non_blocking copies does) and its loss is
    # copied back to pinned host memory; the host reads a loss one step late, so enqueueing step k+1 overlaps
    # the execution of step k.  No L2 flush here: the inputs arrive by DMA each step.
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]

    if graphs is not None:
        def e2e_loop(steps):
            consumed = [None, None]                             # event: the replay that read input set i has finished
            copied = [None, None]

            def issue_h2d(i):
                with torch.cuda.stream(copy_stream):
                    if consumed[i] is not None:
                        copy_stream.wait_event(consumed[i])
                    for k, v in pinned.items():
                        sets[i][k].copy_(v, non_blocking=True)
                    copied[i] = torch.cuda.Event()
                    copied[i].record(copy_stream)

            issue_h2d(0)
            done = [None, None]
            for k in range(steps):
                i = k & 1
                if k + 1 < steps:
                    issue_h2d(i ^ 1)
                main_stream.wait_event(copied[i])
                graphs[i][0].replay()
                consumed[i] = torch.cuda.Event()
                consumed[i].record(main_stream)
                loss_host[i].copy_(graphs[i][1].detach().float(), non_blocking=True)   # D2H read of the result
                done[i] = torch.cuda.Event()
                done[i].record(main_stream)
                if done[i ^ 1] is not None:
                    done[i ^ 1].synchronize()
                    float(loss_host[i ^ 1])
            done[(steps - 1) & 1].synchronize()
            return float(loss_host[(steps - 1) & 1])
    else:
        def issue_h2d():
            with torch.cuda.stream(copy_stream):
                x = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return x, ev

        def e2e_loop(steps):
            pending = issue_h2d()
            done = [None, None]
            for k in range(steps):
                x, ev = pending
                pending = issue_h2d()
                main_stream.wait_event(ev)
                for t in x.values():
                    t.record_stream(main_stream)
                loss, _ = step(x)
                slot = k & 1
                loss_host[slot].copy_(loss.detach().float(), non_blocking=True)   # D2H read of the result
                d = torch.cuda.Event()
                d.record(main_stream)
                if done[slot ^ 1] is not None:                                    # the previous step's loss is on the host
                    done[slot ^ 1].synchronize()
                    float(loss_host[slot ^ 1])
                done[slot] = d
            done[(steps - 1) & 1].synchronize()
            return float(loss_host[(steps - 1) & 1])

    e2e_loop(5)
    # K steps are a 20 ms window, short enough for one slow PCIe / host moment to move the number by 10 % from run to run
    # (230-257 k samples/s seen for the same build): take 5 K steps per pass, best of three passes
    K_e2e = max(50, 5 * K)
    ms_e2e = float("inf")
    for _ in range(3):                      # best of three passes of K_e2e steps (host-side jitter: PCIe, first-touch)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        e2e_loop(K_e2e)
        eb.record()
        barrier()
        ms_e2e = min(ms_e2e, ea.elapsed_time(eb) / K_e2e)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        ms_e2e = float(t)

    # ---- roofline of the dominant kernel (word-region backward) -------------------------------
    # Algorithmic flops use the words that exist (padding words are not part of the problem: they are
    # excluded from the loss and have zero gradient) and the un-padded region count.
    pk = peaks()
    Bg, R = B * world, R_SIDE * R_SIDE
    nvalid = torch.tensor([float((~host["mask"]).sum())], device=dev)
    if world > 1:
        dist.all_reduce(nvalid, op=dist.ReduceOp.SUM, group=group)
    words_valid = float(nvalid)                              # valid word rows of the GLOBAL batch (this rank's columns)
    flops_bwd = 8.0 * B * words_valid * R * D_WORD           # dA, dV(x2 operands), dQ: 4 GEMMs x 2*Bi*words*R*D
    flops_fwd = 4.0 * B * words_valid * R * D_WORD
    n_bwd, ms_bwd = kern.get("wordregion_bwd", (0, float("nan")))
    n_fwd, ms_fwd = kern.get("wordregion_fwd", (0, float("nan")))
    ach = flops_bwd / (ms_bwd * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic("wr_bwd_tc_kernel")
    roofline = {
        "kernel": ("wr_bwd_tc_kernel (word-region backward, tcgen05, bf16 operands)" if precision == "bf16" else
                   "wr_bwd_split_kernel (word-region backward, tcgen05 at fp32 tolerance: hi + lo bf16 operands, 3 MMAs per product)"),
        "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, read from the newest committed `ncu --set full`
        # summary of this very workload (not measured in this run); only valid for the default 1-GPU bf16 run
        "traffic": traffic if (precision == "bf16" and world == 1 and B == B_LOCAL) else None,
        "traffic_source": traffic_src if (precision == "bf16" and world == 1 and B == B_LOCAL) else None,
        "peak_source": pk["src"] + ", bf16 sustained",
        "algorithmic_flops_per_launch": flops_bwd, "ms_per_launch": ms_bwd, "launches_timed": n_bwd,
        "word_rows": {"valid": words_valid, "padded_T": float(Bg * T_WORDS),
                      "note": "flops counted on valid words; with all T=18 slots counted the figure would be x%.2f"
                              % (Bg * T_WORDS / words_valid)},
        "forward": {"ms_per_launch": ms_fwd, "achieved": flops_fwd / (ms_fwd * 1e-3) / 1e12,
                    "frac": flops_fwd / (ms_fwd * 1e-3) / 1e12 / pk["tensor"]},
        "kernels_ms": {k: round(v[1], 4) for k, v in kern.items()},
    }
    if precision == "bf16" and D_WORD == 256:
        # what the kernel's structure allows: its SS-form N = 64 MMAs are bound by shared-memory operand reads,
        # t = max(M*N/256, bytes/128) clk per MMA, measured by profiles/micro/mma_bench.cu; per 64-region chunk 432 KB of
        # operand reads + 77 KB of tile fills = 3 980 clk against 2 560 clk of tensor time (2 048 of it algorithmic)
        roofline["structure_bound"] = {
            "limiter": "shared-memory operand bandwidth (128 B/clk/SM)", "clk_per_chunk_smem": 3980, "clk_per_chunk_tensor": 2560,
            "frac_of_tensor_peak_at_that_bound": round(0.8 * 2560 / 3980 * R / (64 * ((R + 63) // 64 - 1) + ((R - 1) % 64 // 16 + 1) * 16)
                                                       / (pk["tensor"] / 2061.0), 3) if pk["tensor"] else None,
            "source": "profiles/r02_mma_bench.txt, DESIGN.md 4.6",
            "note": "chunk period at the bound vs tensor time, x 0.8 (algorithmic share of the executed MMAs: S is recomputed), x R / padded R, "
                    "against the same measured peak as `frac` (2061 TFLOP/s = 148 SMs x 8192 flop/clk x 1.70 GHz inside the kernel); "
                    "the drain at every image boundary (about 3 000 clk per image) is on top of it"}

    def leave():
        """N > 1: the captured graphs hold NCCL kernels of the communicator and destroy_process_group() did not return
        with them alive (round-2 N = 2 run: every number printed, then a hang in the teardown).  Nothing is left to do
        but exit: drain the device, agree that everybody is done, and let process exit release the communicator."""
        sys.stdout.flush(); sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group)
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        return leave()
    line = {
        "metric": "contrastive-loss fwd+bwd samples/sec", "value": Bg / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(world),
            "arithmetic": "bf16-in/fp32-accumulate" if precision == "bf16" else "fp32",
            "global_batch": Bg, "pairs_per_s": B * Bg * world / (ms * 1e-3), "rho": RHO,
            "parallelism": "single GPU" if world == 1 else f"rows local, columns all-gathered over NCCL x{world}",
            "l2": "256 MiB buffer written between timed steps (untimed) to flush the 126 MB L2",
            "launch": mode, "streams": 1 if (side is None and not fused) or args.one_stream else 3 + int(fused),
            "api": ("train_gan.contrastive_losses (the three losses through one autograd function, grouped collectives)" if fused
                    else "train_gan.sent_loss / img_loss / word_loss (the reference's call sites)"), "eager_ms_per_step": ms_eager, "host_enqueue_ms_per_step": host_ms,
            "scaling_note": "global negatives: per-rank work grows with the global batch (rows local, columns "
                            "gathered), so samples/s stays flat with N while pairs/s grows ~N",
        },
        "clocks": clocks,
        "e2e": {"value": Bg / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": K_e2e, "passes": "best of 3",
                "how": "public API (train_gan.make_labels / sent_loss / img_loss / word_loss + backward, %s); pinned-host "
                       "inputs copied H2D every step on a copy stream one step ahead, loss copied D2H to pinned memory "
                       "every step and read by the host one step late" % mode},
        "gpu_launches": launches,
        "roofline": roofline,
        "sustained": sustained,
    }
    # BASELINE config 2 names two modes, "fp32 and bf16-in/fp32-accum": the default line is the bf16 mode; the fp32 mode
    # (same inputs as fp32 tensors, rel 1e-4: split-bf16 tcgen05 kernels) rides along as a second block
    if world == 1 and precision == "bf16" and not args.no_fp32_block:
        x32 = {k: (v.to(dev).float() if v.dtype.is_floating_point else v.to(dev)) for k, v in make_inputs(B, 1000 + rank).items()}
        precision = "fp32"                                       # step() reads it
        try:
            for _ in range(3):
                step(x32)
            ops.enable_timing(True)
            timed(lambda: step(x32, use_side=False), 5)
            k32 = ops.kernel_ms()
            ops.enable_timing(False)
            mode32 = "eager"
            run32 = lambda: step(x32)
            if not args.no_graph:                                # the same capture as the bf16 line (no host sync in this path either)
                try:
                    g32, gl32, _ = capture(x32)
                    g32.replay(); torch.cuda.synchronize()
                    el32, _ = step(x32)
                    if abs(float(gl32.detach()) - float(el32.detach())) > 1e-5 * abs(float(el32.detach())):
                        raise RuntimeError("graph replay does not reproduce the eager step")
                    run32, mode32 = g32.replay, "cuda-graph"
                except Exception as e:                            # noqa: BLE001
                    print(f"[bench] fp32 block: CUDA-graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
                    torch.cuda.synchronize()
            for _ in range(3):
                run32()
            ms32 = timed(run32, 10)
            f32_bwd = flops_bwd / (k32["wordregion_bwd"][1] * 1e-3) / 1e12
            line["fp32"] = {
                "value": Bg / (ms32 * 1e-3), "unit": "samples/s", "ms_per_step": ms32, "steps": 10, "launch": mode32, "tol": 1e-4,
                "arithmetic": "fp32 inputs; word-region products as three bf16 MMAs on hi + lo operand pairs (tcgen05), "
                              "similarity losses fp32 CUDA-core kernels; fp32 accumulation throughout",
                "kernels_ms": {k: round(v[1], 4) for k, v in k32.items()},
                "roofline": {"kernel": "wr_bwd_split_kernel", "bound": "tensor", "achieved": f32_bwd, "peak": pk["tensor"],
                             "unit": "TFLOP/s", "frac": f32_bwd / pk["tensor"],
                             "note": "algorithmic flops of the fp32 problem against the bf16 peak; the kernel executes "
                                     "three bf16 MMAs per product"}}
        finally:
            precision = "bf16"
        del x32
    if world == 1 and not args.no_cpu_baseline:
        sps, dt, first = time_oracle(2, 1, B, 1, precision)
        line["cpu_baseline"] = {
            "value": sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "the GPU arm's own COCO-256 batch (same seed%s), 2 timed + 1 warm-up fwd+bwd steps of the CPU oracle "
                      "(fp32 arithmetic, all host threads)" % (", rounded to bf16 as the kernels see it" if precision == "bf16" else ""),
            "ms_per_step": dt * 1e3}
        if gpu_out is not None:
            parity = {"against": "CPU oracle (oracle/ref_losses.py + oracle/word_region.py) on identical inputs",
                      "loss_rel": abs(gpu_out[0] - first[0]) / abs(first[0]),
                      "grad_nerr": [nerr(a, b) for a, b in zip(gpu_out[1], first[1])],
                      "order": ["d img", "d sent", "d fake", "d words", "d regions"]}
    if parity is not None:
        parity["tol"] = 2e-2 if precision == "bf16" else 1e-4
        parity["ok"] = bool(parity["loss_rel"] <= parity["tol"] and max(parity["grad_nerr"]) <= parity["tol"])
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    leave()


if __name__ == "__main__":
    main()
