#!/usr/bin/env python
"""Benchmark of the MuRCL pre-training hot path (BASELINE.json: "WSI bags/sec (MIL fwd+bwd + NT-Xent)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl murcl|reference] [--precision bf16|fp32]

Workload (BASELINE config 3, SURVEY.md 8d "cfg3"): one optimiser step of stage-3 pre-training per "step":
B=128 slides per GPU (ragged, N_i ~ U[500,15500] x 512-d, K=10 clusters), T=6 patch-steps x 2 views,
RL actor chooses the windows, select+gather+mixup -> ABMIL(512,512,128) -> Full_layer(512,1024,128) ->
NT-Xent over the global batch, backward, Adam (optim.ArenaAdam: torch.optim.Adam's update as one launch over the flat
parameter arena).  The whole step - head / loss chain on its side stream included - replays as one CUDA graph; the step's
random draws are issued up front (`--rng batched`; `--rng reference` keeps the reference's per-patch-step call order).
Weak scaling: every rank owns 128 slides; embeddings are all-gathered (NCCL) for the loss (from 4 ranks on each rank
reduces only its own rows of the global score matrix and the per-row statistics are all-gathered too) and gradients are
all-reduced once per step.  `roofline.yardstick` times the dominant GEMM shape alone next to cuBLAS on the same shape.

One JSON line on stdout (rank 0).  `value` = device-timed throughput with the slides already in HBM;
`e2e` = the same step driven from the PINNED HOST staging of a finite dataset shard (`--dataset-slides` per rank) through
the resident slide cache (csr.ResidentSlides): a slide crosses PCIe the first time a step touches it (prefetched on a copy
stream one step ahead), afterwards a step uploads only its slide-id list; the timed region spans `--e2e-epochs` epochs
INCLUDING the cold first one, and the loss is read back every step.  Secondary keys of the same line: `fp32_mode` (the
reference-precision step), `strong_scaling` (BASELINE config 3 verbatim: 128 bags per step over all ranks) and `cfg5`
(one 100k x 1024 bag, rows sharded over the ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "WSI bags/sec (MIL fwd+bwd + NT-Xent)"
UNIT = "bags/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="murcl", choices=["murcl", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MURCL_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--bags", type=int, default=128, help="slides per GPU per step")
    ap.add_argument("--T", type=int, default=6)
    ap.add_argument("--stage", type=int, default=3, choices=[1, 3],
                    help="train_stage of the step: 3 (default, the headline: the PPO actor chooses the windows) or 1 (random windows)")
    ap.add_argument("--feat-size", type=int, default=1024)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--clusters", type=int, default=10)
    ap.add_argument("--min-patches", type=int, default=500)
    ap.add_argument("--max-patches", type=int, default=15500)
    ap.add_argument("--dataset-slides", type=int, default=256, help="slides in a rank's dataset shard (resident cache)")
    ap.add_argument("--e2e-epochs", type=int, default=10, help="epochs of the dataset shard in the end-to-end region (first one cold)")
    ap.add_argument("--fp32-steps", type=int, default=3, help="timed steps of the secondary fp32-mode measurement (0 = skip)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the fp32-mode / strong-scaling / cfg5 measurements")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="exchange the head gradients under the aggregators' backward (measured slower: NCCL's CTAs take SMs from "
                         "the persistent one-CTA-per-SM GEMMs, whose last CTAs then run as a second wave)")
    ap.add_argument("--cpu-bags", type=int, default=8, help="slides in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fp32-store", action="store_true", help="keep the slide store in fp32 even in bf16 mode")
    ap.add_argument("--rng", default="batched", choices=["batched", "reference"],
                    help="random draws of the step: issued once up front (default) or per patch-step in the reference's call order")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_config(a, world):
    return {"workload": "cfg3: MuRCL stage-3 pre-train step (RL window selection + gather + mixup + ABMIL + Full_layer GRU "
                        "+ NT-Xent, fwd+bwd+Adam)",
            "bags_per_gpu": a.bags, "global_bags": a.bags * world, "T": a.T, "views": 2, "feat_size": a.feat_size,
            "feat_dim": a.dim, "clusters": a.clusters, "patches_per_bag": f"U[{a.min_patches},{a.max_patches}]",
            "arch": "ABMIL(512,512,128)+Full_layer(512,1024,128)", "parallelism": f"dp{world} (bags sharded per rank)",
            "l2_policy": "inputs larger than L2 (resident slide arena ~2 GB/GPU, a different batch of slides every step, activations 268 MB each)",
            "dataset_slides_per_gpu": a.dataset_slides,
            "slide_store_dtype": "bf16" if (a.precision == "bf16" and not a.fp32_store) else "f32",
            "bag_passes_per_step": a.bags * world * a.T * 2}


# ------------------------------------------------------------------------------------------------------
# synthetic slides
# ------------------------------------------------------------------------------------------------------
def make_host_batch(a, seed, pin=True, n_slides=None, precision=None):
    from murcl_b200 import synth
    from murcl_b200.csr import HostBags
    n_slides = n_slides or a.bags
    precision = precision or a.precision
    sizes = synth.camelyon_sizes(n_slides, a.min_patches, a.max_patches, seed=seed)
    feats, _clusters, labels = synth.make_bags(sizes, a.dim, a.clusters, seed=seed)
    # bf16 mode keeps the slide features as bf16 on the host too (what the first GEMM consumes): half the H2D bytes
    store_dtype = torch.bfloat16 if (precision == "bf16" and not a.fp32_store) else torch.float32
    return HostBags(feats, labels, a.clusters, pin=pin, dtype=store_dtype)


# ------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------------
class Clocks:
    """nvidia-smi sampler (25 ms period).  The process takes a few hundred ms to deliver its first line - longer than a short
    timed region - so it is started BEFORE the warm-up steps (`start` waits for the first sample) and only the samples that
    arrive between `mark_begin()` and `stop()` are reported; if the region was shorter than a sampling period the nearest
    samples around it are used and `samples_in_region` says so."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index
        self.t_begin = None

    def start(self, wait_s=5.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < wait_s:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.06)                                    # let the sample that was in flight at t_end arrive
        self.proc.terminate()
        t_begin = self.t_begin if self.t_begin is not None else 0.0
        inside = [l for t, l in self.lines if t_begin <= t <= t_end + 0.05]
        used = inside
        if not used:                                        # region shorter than a sampling period: the samples around it
            before = [l for t, l in self.lines if t < t_begin][-1:]
            after = [l for t, l in self.lines if t > t_end][:1]
            used = before + after
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in used:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_region": len(inside)}


# ------------------------------------------------------------------------------------------------------
# the job
# ------------------------------------------------------------------------------------------------------
class Job:
    def __init__(self, a, rank, world, device, bags=None, precision=None):
        from murcl_b200 import dist as mdist
        from murcl_b200 import synth
        from murcl_b200.dropin import abmil, cl, losses, rlmil
        self.a, self.rank, self.world, self.device = a, rank, world, device
        self.bags = bags or a.bags
        self.precision = precision or a.precision
        os.environ["MURCL_PRECISION"] = self.precision      # heads (Full_layer, actor, decoder) follow the same mode
        torch.manual_seed(985 + rank)
        enc = abmil.ABMIL(a.dim, L=512, D=128, dim_out=128, precision=self.precision)
        enc.load_state_dict(synth.abmil_state(a.dim, 512, 128, 128, seed=985, peak=2.0))
        self.model = cl.CL(enc.to(device), projection_dim=128, n_features=512)
        fc = rlmil.Full_layer(512, 1024, True, 128)
        fc.load_state_dict(synth.full_layer_state(512, 1024, 128, seed=986))
        self.fc = fc.to(device)
        self.ppo = rlmil.PPO(a.dim, 512, 512, False, action_std=0.5, lr=1e-5, gamma=0.1, K_epochs=3, action_size=a.clusters)
        self.ppo.policy_old.load_state_dict(synth.actor_state(512, 512, a.clusters, seed=987))
        self.memories = [rlmil.Memory(), rlmil.Memory()]
        for m in (self.fc, self.ppo.policy, self.ppo.policy_old):
            m.precision = self.precision
        if world > 1:
            self.crit = mdist.DistributedNTXent(self.bags, 1.0)
        else:
            self.crit = losses.NT_Xent(self.bags, 1.0)
        self.params = list(self.model.parameters()) + list(self.fc.parameters())
        # one flat fp32 parameter / gradient / bf16-shadow arena: gradients are summed inside the kernels, one memset
        # clears them, one cast refreshes the bf16 weights, one all-reduce exchanges them, Adam updates one tensor
        from murcl_b200.arena import ParamArena
        self.arena = ParamArena(self.params, shadow_dtype=torch.bfloat16 if self.precision == "bf16" else None)
        # Adam as upstream (train_MuRCL.py:154-171), as ONE fused launch over the arena that also refreshes the bf16 shadow
        # weights (murcl_adam_step); MURCL_TORCH_ADAM=1 runs torch's multi-tensor Adam + the separate cast instead
        if os.environ.get("MURCL_TORCH_ADAM", "0") == "1":
            self.opt = torch.optim.Adam(self.arena.optimizer_params(), lr=1e-4, weight_decay=1e-5, capturable=True)
        else:
            from murcl_b200.optim import ArenaAdam
            self.opt = ArenaAdam(self.arena, lr=1e-4, weight_decay=1e-5)
        self.head_range = self.arena.range_of(list(self.fc.parameters()))      # Full_layer: 80 % of the gradient bytes
        self.mdist = mdist
        self.graphs = {}
        self.launches_per_step = None

    def capture(self, store, slot_bag):
        """The whole optimiser step as ONE CUDA graph over the resident arena; the step's slides are chosen by the
        contents of the static ``slot_bag`` buffer.  Returns False (on every rank) if the capture fails on any rank."""
        import torch.distributed as dist
        from murcl_b200 import pretrain
        ok = True
        try:
            # NCCL's watchdog thread polls events while we capture: keep the capture thread-local under torchrun
            mode = "thread_local" if self.world > 1 else "global"
            g = pretrain.GraphedStep(lambda: self.step(store, slot_bag), warmup=2, capture_error_mode=mode)
            self.graphs[id(store)] = g
            self.launches_per_step = g.launches
        except Exception as e:                                  # noqa: BLE001 - report and fall back to eager launches
            import traceback
            sys.stderr.write(f"[bench] rank {self.rank}: CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}\n"
                             + "".join(traceback.format_exc().splitlines(keepends=True)[-14:]))
            ok = False
        torch.cuda.synchronize()
        if self.world > 1:
            flag = torch.tensor([1 if ok else 0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if not ok:
            self.graphs = {}
        return ok

    def run(self, store, slot_bag):
        g = self.graphs.get(id(store))
        return g() if g is not None else self.step(store, slot_bag)

    def step(self, store, slot_bag):
        from murcl_b200 import pretrain
        self.arena.zero_grad()
        pending = []
        lo, hi = self.head_range

        def exchange_heads():       # the head gradients are final here: their all-reduce overlaps the aggregators' backward
            pending.append(self.arena.allreduce(lo=lo, hi=hi, async_op=True))

        loss, _ = pretrain.pretrain_step(store, self.model, self.fc, self.crit, T=self.a.T, feat_size=self.a.feat_size,
                                         alpha=0.9, stage=self.a.stage, ppo=self.ppo, memories=self.memories,
                                         precision=self.precision, slot_bag=slot_bag, rng=self.a.rng,
                                         after_head_backward=exchange_heads if (self.world > 1 and self.a.overlap_allreduce) else None)
        if self.world > 1:
            if self.a.overlap_allreduce:
                self.arena.allreduce(lo=0, hi=lo)
                self.arena.allreduce(lo=hi)
                for w in pending:
                    if w is not None:
                        w.wait()
            else:
                self.arena.allreduce()      # ONE all-reduce of the flat gradient buffer (24 MB), after the backward
        self.opt.step()
        if isinstance(self.opt, torch.optim.Optimizer):
            self.arena.refresh()
        return loss


def timed(fn, steps, world, device):
    """EXACTLY `steps` calls bracketed by barrier + synchronize; CUDA events; max over ranks (ms)."""
    import torch.distributed as dist
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def roofline_probe(job, store, slot_bag, peaks):
    """Times every dense-layer launch of one extra step with CUDA events on the launch stream and reports the
    dominant kernel family (the encoder/attention GEMMs) against the measured tensor peak."""
    from murcl_b200 import ops
    log = []
    ops.set_profile(log)
    torch.cuda.nvtx.range_push("murcl_probe_step")      # ncu --nvtx --nvtx-include "murcl_probe_step/" profiles exactly this step
    try:
        job.step(store, slot_bag)
        torch.cuda.synchronize()
    finally:
        torch.cuda.nvtx.range_pop()
        ops.set_profile(None)
    by = {}
    for name, flops, e0, e1 in log:
        ms = e0.elapsed_time(e1)
        t = by.setdefault(name, [0.0, 0.0, 0])
        t[0] += flops; t[1] += ms; t[2] += 1
    if not by:
        return None
    # dominant kernel = the tcgen05 GEMM on the instance-level layers (encoder, attention projection and their
    # gradients: M = all instance rows of the step's bags); the batch-sized head layers are reported beside it
    big = {k: v for k, v in by.items() if k.startswith("linear_")}
    heads = {k: v for k, v in by.items() if k.startswith("head_")}
    tot_flops = sum(v[0] for v in big.values())
    tot_ms = sum(v[1] for v in big.values())
    n_big = sum(v[2] for v in big.values())
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    ach = tot_flops / (tot_ms * 1e-3) / 1e12
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
        except (ValueError, OSError):
            traffic = None
    hbm_fams = ("attnpool_fwd", "attnpool_bwd")
    fam = {k: {"launches": v[2], "ms": round(v[1], 3), "tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1)} for k, v in by.items()
           if k not in hbm_fams}
    hbm = peaks.get("hbm_gbs") or 6500.0
    for name in hbm_fams:        # the fused attention-pooling kernels are HBM-bound: their records carry algorithmic bytes
        if name in by:
            v = by[name]
            gbs = v[0] / (v[1] * 1e-3) / 1e9
            fam[name] = {"launches": v[2], "ms": round(v[1], 3), "bound": "hbm", "algorithmic_gbs": round(gbs, 1),
                         "frac_of_hbm_peak": round(gbs / hbm, 3)}
    return {"bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
            "traffic": traffic,
            "kernel": "gemm_tc_kernel (tcgen05) on the instance-level dense layers: murcl_linear_fwd / bwd_input / bwd_weight, "
                      "M = 262144 rows per launch",
            "launches": n_big, "ms_per_launch": round(tot_ms / max(n_big, 1), 4),
            "algorithmic_flop_per_launch": tot_flops / max(n_big, 1), "gemm_ms_per_step": round(tot_ms, 3),
            "algorithmic_flop_per_step": tot_flops, "families": fam,
            "head_layers_ms_per_step": round(sum(v[1] for v in heads.values()), 3),
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback",
            "timing": "CUDA events around every launch of one extra eager step on the launch stream"}


def step_roofline(a, ms_per_step, peaks):
    """Roofline of the WHOLE step from its algorithmic work: per kernel of a patch-step the
    compulsory HBM bytes (bf16 storage) and FLOPs, roof = sum over kernels of max(bytes / HBM peak, FLOPs / tensor peak).
    The batch-sized head layers, NT-Xent and the optimiser (< 3 % of the bytes, < 1 % of the FLOPs) are left out of the roof,
    which makes `frac` conservative."""
    s = 2 if a.precision == "bf16" else 4
    rows = 2.0 * a.bags * a.feat_size
    D, L, DA = float(a.dim), 512.0, 128.0
    hbm = (peaks.get("hbm_gbs") or 6500.0) * 1e9
    tc = (peaks.get("bf16_tflops_sustained") or 1400.0) * 1e12
    k = []                                                    # (name, bytes, flops) of one patch-step, both views
    k.append(("pack_gather", rows * (3 * D * s + 4), 0.0))
    for kin in (D, L, L):                                     # encoder forward: x -> h1 -> h2 -> h3 (+ 1 mask bit per output)
        k.append(("enc_fwd", rows * (kin * s + L * s + L / 8), 2 * rows * kin * L))
    k.append(("attnpool_fwd", rows * (L * s + DA * s + 8), 2 * rows * L * DA + 2 * rows * L))
    k.append(("attnpool_bwd", rows * (L * s + 2 * DA * s + 4), 2 * rows * L))
    k.append(("attn_wgrad", rows * (DA * s + L * s), 2 * rows * DA * L))
    k.append(("attn_dgrad", rows * (DA * s + L * s + L / 8), 2 * rows * DA * L))
    for kin in (L, L, D):                                     # encoder weight gradients (dZ and the layer's input)
        k.append(("enc_wgrad", rows * (L * s + kin * s), 2 * rows * L * kin))
    for _ in range(2):                                        # encoder input gradients (none for the first layer)
        k.append(("enc_dgrad", rows * (2 * L * s + L / 8), 2 * rows * L * L))
    roof = a.T * sum(max(b / hbm, f / tc) for _, b, f in k)
    tot_b, tot_f = a.T * sum(b for _, b, _ in k), a.T * sum(f for _, _, f in k)
    return {"roof_ms": round(roof * 1e3, 3), "frac": round(roof * 1e3 / ms_per_step, 4), "hbm_bytes": tot_b, "flop": tot_f,
            "hbm_only_ms": round(tot_b / hbm * 1e3, 3), "tensor_only_ms": round(tot_f / tc * 1e3, 3),
            "def": "sum over the step's instance-level kernels of max(algorithmic bytes / measured HBM peak, algorithmic FLOP / "
                   "measured sustained bf16 peak) / measured ms_per_step"}


def gemm_yardstick(rows, n=512, k=512, reps=20):
    """The dominant GEMM shape timed ALONE (burst clocks, back-to-back launches, inputs larger than L2): this repo's fused
    forward kernel (bias + ReLU + bit mask) beside cuBLAS through torch.matmul (no epilogue).  Context for `roofline.frac`:
    a [rows, 512] x [512, 512] product sits at the HBM / tensor ridge, where no kernel reaches the square-GEMM peak."""
    from murcl_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(rows, k, generator=g).to("cuda", torch.bfloat16)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to("cuda", torch.bfloat16)
    b = torch.zeros(n, device="cuda")
    wt = w.t().contiguous()

    def t_us(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    own = t_us(lambda: ops.linear_fwd(x, w, b, ops.ACT_RELU, relu_bits=True))
    lib = t_us(lambda: torch.matmul(x, wt))
    fl = 2.0 * rows * n * k
    return {"shape": [rows, n, k], "own_fused_fwd_us": round(own, 1), "own_fused_fwd_tflops": round(fl / own / 1e6, 1),
            "cublas_us": round(lib, 1), "cublas_tflops": round(fl / lib / 1e6, 1),
            "note": "timed alone at burst clocks (the step runs power-capped, see clocks); cuBLAS = torch.matmul, bf16, no epilogue"}


def cpu_baseline(a, threads=None, reps=2):
    """The oracle's restatement of the same step on the host cores, on a bounded sample (a.cpu_bags slides)."""
    from murcl_b200 import synth
    from oracle import murcl_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sizes = synth.camelyon_sizes(a.cpu_bags, a.min_patches, a.max_patches, seed=4242)
    feats, clusters, _ = synth.make_bags(sizes, a.dim, a.clusters, seed=4242)
    sd_m = synth.abmil_state(a.dim, 512, 128, 128, seed=985, peak=2.0)
    sd_f = synth.full_layer_state(512, 1024, 128, seed=986)
    g = synth.gen(1)
    O.pretrain_step(feats, clusters, sd_m, sd_f, T=1, feat_size=a.feat_size, generator=g)       # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.pretrain_step(feats, clusters, sd_m, sd_f, T=a.T, feat_size=a.feat_size, generator=g)
        times.append(time.perf_counter() - t0)
    best = min(times)
    out = {"value": round(a.cpu_bags / best, 3), "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{a.cpu_bags} slides (of {a.bags}), same size distribution, T={a.T} x 2 views, fwd+bwd, "
                     f"best of {reps}, torch CPU fp32 oracle, {best:.2f} s/step"}
    # the reference trainers pin torch to ONE thread (train_MuRCL.py:484, train_RLMIL.py:1162): the same sample that way,
    # one patch-step per view pair (T = 1) scaled to T - the oracle's cost is linear in T - to keep the bench short
    if threads > 1 and not getattr(a, "no_cpu_1thread", False):
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        O.pretrain_step(feats, clusters, sd_m, sd_f, T=1, feat_size=a.feat_size, generator=g)
        one = (time.perf_counter() - t0) * a.T
        torch.set_num_threads(threads)
        out["single_thread"] = {"value": round(a.cpu_bags / one, 3), "unit": UNIT, "cores": 1,
                                "sample": f"same {a.cpu_bags} slides, one patch-step timed and scaled by T={a.T}, {one:.2f} s/step"}
    return out, best


def run_reference(a):
    """`--impl reference`: the reference algorithm on the host cores (the oracle port: the reference is pure
    Python/PyTorch and /root/reference does not travel to the GPU box).  Each step = one bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from murcl_b200 import synth
    from oracle import murcl_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    sizes = synth.camelyon_sizes(a.cpu_bags, a.min_patches, a.max_patches, seed=4242)
    feats, clusters, _ = synth.make_bags(sizes, a.dim, a.clusters, seed=4242)
    sd_m = synth.abmil_state(a.dim, 512, 128, 128, seed=985, peak=2.0)
    sd_f = synth.full_layer_state(512, 1024, 128, seed=986)
    g = synth.gen(1)
    for _ in range(a.warmup):                     # warm-up steps are one patch-step each (thread pool, allocator, caches)
        O.pretrain_step(feats, clusters, sd_m, sd_f, T=1, feat_size=a.feat_size, generator=g)
    t0 = time.perf_counter()
    steps = max(1, a.steps)
    for _ in range(steps):
        O.pretrain_step(feats, clusters, sd_m, sd_f, T=a.T, feat_size=a.feat_size, generator=g)
    dt = (time.perf_counter() - t0) / steps
    v = round(a.cpu_bags / dt, 3)
    sample = (f"{a.cpu_bags} slides per step (bounded sample of the {a.bags}-slide step), T={a.T} x 2 views, fwd+bwd, "
              f"{steps} timed steps, torch CPU fp32 oracle port, {threads} threads")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
                      "warmup": a.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": workload_config(a, world),
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


_T0 = time.time()


def vlog(msg):
    if os.environ.get("MURCL_BENCH_VERBOSE"):
        sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - _T0:6.1f}s] {msg}\n")
        sys.stderr.flush()


def epoch_batches(n_slides, bags, n_epochs, seed):
    """Slide ids of every step of ``n_epochs`` epochs: a fresh shuffle per epoch, remainder dropped
    (train_MuRCL.py:211,229-233)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_epochs):
        perm = torch.randperm(n_slides, generator=g)
        for i in range(n_slides // bags):
            out.append(perm[i * bags:(i + 1) * bags].to(torch.int32))
    return out


def measure_fp32(a, rank, world, device):
    """Secondary measurement: the same optimiser step in fp32 mode (the reference's precision: fp32 slide store, fp32-exact
    GEMMs - bf16 tensor cores on 3-plane splits of the fp32 operands -, 1e-5 parity budget), eager launches."""
    from murcl_b200.csr import BagStore
    host = make_host_batch(a, seed=3000 + rank, pin=False, precision="fp32")
    store = BagStore.empty_like_host(host, device)
    store.copy_from_host(host)
    job = Job(a, rank, world, device, precision="fp32")
    slot_bag = torch.arange(a.bags, dtype=torch.int32, device=device).repeat(2)
    for _ in range(2):
        job.step(store, slot_bag)
    ms = timed(lambda i: job.step(store, slot_bag), a.fp32_steps, world, device)
    out = {"dtype": "f32", "value": round(a.bags * world * a.fp32_steps / (ms * 1e-3), 2), "unit": UNIT,
           "ms_per_step": round(ms / a.fp32_steps, 3), "steps": a.fp32_steps, "warmup": 2, "cuda_graph": False,
           "slide_store_dtype": "f32", "gemm": "tcgen05 split precision: 3 bf16 planes per fp32 operand, 6 exact products (MURCL_FP32_GEMM=split3)",
           "note": "same cfg3 step in the reference's precision: fp32 storage, fp32-exact arithmetic (1e-5 parity mode)"}
    del job, store, host
    torch.cuda.empty_cache()
    return out


def measure_strong(a, rank, world, device, steps):
    """Secondary measurement (world > 1): BASELINE config 3 verbatim - 128 bags per optimiser step over ALL ranks
    (strong scaling: 128 / world slides per rank)."""
    from murcl_b200.csr import BagStore
    if a.bags % world:
        return None
    local = a.bags // world
    host = make_host_batch(a, seed=5000 + rank, pin=False, n_slides=2 * local)
    store = BagStore.empty_like_host(host, device)
    store.copy_from_host(host)
    job = Job(a, rank, world, device, bags=local)
    slot_bag = torch.empty(2 * local, dtype=torch.int32, device=device)
    table = [torch.arange(k * local, (k + 1) * local, dtype=torch.int32, device=device).repeat(2) for k in range(2)]
    slot_bag.copy_(table[0])
    graphed = (not a.no_graph) and job.capture(store, slot_bag)

    def one(i):
        slot_bag.copy_(table[i % 2])
        job.run(store, slot_bag)

    for i in range(3):
        one(i)
    ms = timed(one, steps, world, device)
    return {"global_bags": a.bags, "bags_per_gpu": local, "value": round(a.bags * steps / (ms * 1e-3), 2), "unit": UNIT,
            "ms_per_step": round(ms / steps, 3), "steps": steps, "scaling": "strong", "cuda_graph": bool(graphed)}


def measure_cfg5(a, rank, world, device, iters=10):
    """Secondary measurement: BASELINE config 5 - ONE bag of 100 000 patches x 1024-d through CLAM_SB (gated attention,
    bf16 mode), its rows sharded over the ranks (`CLAM_SB.shard_bags`): local encoder + fused pooling, one all-gather
    of the pooling partials, backward without a collective, then the gradient all-reduce."""
    from murcl_b200 import dist as mdist
    from murcl_b200 import synth
    from murcl_b200.dropin import clam
    n_total, dim = 100000, 1024
    lo, hi = mdist.shard_range(n_total, rank, world)
    g = torch.Generator().manual_seed(7000 + rank)
    x = torch.clamp_min(0.5 * torch.randn(hi - lo, dim, generator=g) + 0.3, 0).to(device)
    m = clam.CLAM_SB(gate=True, size_arg="small", in_dim=dim, precision="bf16")
    m.load_state_dict(synth.clam_state(dim, "small", True, False, 2, seed=71, peak=3.0))
    m = m.to(device).eval()
    if world > 1:
        m.shard_bags(True)
    params = [p for p in m.parameters()]
    cot = torch.randn(1, 512, generator=torch.Generator().manual_seed(7)).to(device)

    def one(_i):
        for p in params:
            p.grad = None
        out, _ = m([x])
        (out * cot).sum().backward()
        if world > 1:
            mdist.allreduce_grads(params)

    for i in range(3):
        one(i)
    ms = timed(one, iters, world, device)
    per = ms / iters
    return {"workload": "cfg5: one bag of 100000 x 1024 patches, CLAM_SB small (gated), bf16 mode, fwd+bwd", "n_gpus": world,
            "rows_per_gpu": hi - lo, "ms_per_bag": round(per, 3), "bags_per_s": round(1e3 / per, 1),
            "patches_per_s": round(n_total / (per * 1e-3), 0), "iters": iters,
            "sharding": "rows of the bag split over the ranks; pooling partials merged by one all-gather" if world > 1 else "whole bag on one GPU"}


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MIL hot path has no CPU fallback; use --impl reference)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ["MURCL_PRECISION"] = a.precision          # heads (Full_layer, actor, decoder) follow the same mode
    vlog("init_process_group")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    vlog("process group up")
    from murcl_b200 import _lib
    from murcl_b200.csr import ResidentSlides
    _lib.load()
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())

    if a.dataset_slides < 2 * a.bags:
        raise SystemExit("--dataset-slides must hold at least two batches (different slides on consecutive steps)")
    host = make_host_batch(a, seed=1000 + 10 * rank, n_slides=a.dataset_slides)       # the rank's dataset shard, pinned
    slides = ResidentSlides(host, device)
    slides.ensure(range(a.dataset_slides))
    store = slides.store
    torch.cuda.synchronize()
    vlog("dataset shard staged (pinned) and resident")
    job = Job(a, rank, world, device)
    vlog("job built")
    steps_per_epoch = a.dataset_slides // a.bags
    batches = epoch_batches(a.dataset_slides, a.bags, max(2, -(-(a.steps + a.warmup) // steps_per_epoch)), seed=77 + rank)
    table = torch.stack([b.repeat(2) for b in batches]).to(device)              # [n_batches, 2B] slot -> slide
    slot_bag = torch.empty(2 * a.bags, dtype=torch.int32, device=device)        # static buffer the graph reads
    slot_bag.copy_(table[0])

    # ---- device-resident throughput --------------------------------------------------------------
    graphed = (not a.no_graph) and job.capture(store, slot_bag)
    vlog(f"graph capture: {graphed}")

    def resident_step(i):
        slot_bag.copy_(table[i % table.shape[0]])          # device-to-device: this step's slide ids
        job.run(store, slot_bag)

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()                                      # before the warm-up: the sampler needs a moment to come up
    for i in range(a.warmup):
        resident_step(i)
    torch.cuda.synchronize()
    vlog("warm-up done")
    if rank == 0:
        clocks.mark_begin()
    l0 = _lib.launch_count()
    ms = timed(lambda i: resident_step(a.warmup + i), a.steps, world, device)
    launches = job.launches_per_step * a.steps if graphed else _lib.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    value = a.bags * world * a.steps / (ms * 1e-3)
    vlog(f"timed region done: {ms / a.steps:.2f} ms/step")

    # ---- end to end: finite dataset shard on the host, resident slide cache ---------------------------
    e2e = None
    if not a.no_e2e:
        copy_stream = torch.cuda.Stream(device)
        plan = epoch_batches(a.dataset_slides, a.bags, a.e2e_epochs, seed=991 + rank)
        n_e2e = len(plan)
        ids_pinned = torch.stack([b.repeat(2) for b in plan]).pin_memory()
        ready = [torch.cuda.Event() for _ in range(n_e2e)]
        h2d = [0] * n_e2e
        losses = []
        marks = {}

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                h2d[i] += slides.ensure(plan[i].tolist())     # first touch: the slide's rows cross PCIe once
                ready[i].record(copy_stream)

        def e2e_step(i):
            if i == 0:
                prefetch(0)
            if i + 1 < n_e2e:
                prefetch(i + 1)                                # next batch's missing slides overlap this step's compute
            torch.cuda.current_stream().wait_event(ready[i])
            slot_bag.copy_(ids_pinned[i], non_blocking=True)   # H2D: the step's slide ids
            h2d[i] += ids_pinned[i].numel() * 4
            loss = job.run(store, slot_bag)
            losses.append(float(loss.item()))                  # D2H read of the step's result
            if i + 1 == steps_per_epoch or i + 1 == n_e2e:
                marks[i + 1] = time.perf_counter()

        slides.evict_all()                                     # cold start: nothing is resident
        torch.cuda.synchronize()
        t_start = time.perf_counter()
        ms_e2e = timed(e2e_step, n_e2e, world, device)
        cold_s = marks[steps_per_epoch] - t_start              # host clock: every step ends with loss.item()
        warm_s = marks[n_e2e] - marks[steps_per_epoch]
        warm_steps = n_e2e - steps_per_epoch
        e2e = {"value": round(a.bags * world * n_e2e / (ms_e2e * 1e-3), 2), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(h2d) / n_e2e), "d2h_bytes_per_step": 4, "steps": n_e2e,
               "ms_per_step": round(ms_e2e / n_e2e, 3), "epochs": a.e2e_epochs, "steps_per_epoch": steps_per_epoch,
               "h2d_bytes_total": int(sum(h2d)),
               "cold_epoch": {"steps": steps_per_epoch, "ms_per_step": round(cold_s * 1e3 / steps_per_epoch, 3),
                              "value": round(a.bags * world * steps_per_epoch / cold_s, 2),
                              "h2d_bytes_per_step": int(sum(h2d[:steps_per_epoch]) / steps_per_epoch)},
               "steady_state": ({"steps": warm_steps, "ms_per_step": round(warm_s * 1e3 / warm_steps, 3),
                                 "value": round(a.bags * world * warm_steps / warm_s, 2),
                                 "h2d_bytes_per_step": int(sum(h2d[steps_per_epoch:]) / warm_steps)} if warm_steps else None),
               "note": "finite dataset shard in pinned host memory; resident slide cache: a slide is uploaded (copy stream, one "
                       "step ahead) the first time a step touches it, later steps upload only their slide-id list; the timed "
                       "region covers all epochs including the cold first one; loss.item() every step; cold/steady splits are "
                       "rank-0 host-clock marks inside the device-timed region"}

    roof = roofline_probe(job, store, slot_bag, peaks)      # every rank runs it: the step contains collectives
    if roof is not None and rank == 0 and a.precision == "bf16":
        try:
            roof["yardstick"] = gemm_yardstick(2 * a.bags * a.feat_size)
        except Exception as e:                               # noqa: BLE001 - context only, never fatal
            roof["yardstick"] = {"error": f"{type(e).__name__}: {e}"}
    if roof is not None:
        roof["step"] = step_roofline(a, ms / a.steps, peaks)
    vlog("roofline probe done")

    secondary = {}
    if not a.no_secondary:
        del job
        torch.cuda.empty_cache()
        if world > 1:
            secondary["strong_scaling"] = measure_strong(a, rank, world, device, a.steps)
        else:
            secondary["strong_scaling"] = {"global_bags": a.bags, "bags_per_gpu": a.bags, "value": round(value, 2), "unit": UNIT,
                                           "ms_per_step": round(ms / a.steps, 3), "scaling": "strong",
                                           "note": "at 1 GPU the strong-scaling workload is the headline workload"}
        vlog("strong scaling done")
        secondary["cfg5"] = measure_cfg5(a, rank, world, device)
        vlog("cfg5 done")
        if a.fp32_steps > 0 and a.precision == "bf16":
            secondary["fp32_mode"] = measure_fp32(a, rank, world, device)
        vlog("fp32 mode done")
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu, _ = cpu_baseline(a)

    if rank == 0:
        out = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16" if a.precision == "bf16" else "f32", "data": "synthetic",
               "config": dict(workload_config(a, world), cuda_graph=bool(graphed)), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roof, "cpu_baseline": cpu,
               "bag_passes_per_s": round(value * a.T * 2, 1)}
        out.update(secondary)
        print(json.dumps(out), flush=True)
    if world > 1:
        # Tear down without ncclCommDestroy: destroying a communicator whose kernels live inside captured CUDA graphs
        # can block; every rank has finished its work and rank 0 has printed, so synchronise the device and exit.
        torch.cuda.synchronize()          # all collectives this rank took part in have completed
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
