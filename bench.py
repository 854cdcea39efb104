#!/usr/bin/env python
"""Benchmark of the MuRCL pre-training hot path (BASELINE.json: "WSI bags/sec (MIL fwd+bwd + NT-Xent)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl murcl|reference] [--precision bf16|fp32]

Workload (BASELINE config 3, SURVEY.md 8d "cfg3"): one optimiser step of stage-3 pre-training per "step":
B=128 slides per GPU (ragged, N_i ~ U[500,15500] x 512-d, K=10 clusters), T=6 patch-steps x 2 views,
RL actor chooses the windows, select+gather+mixup -> ABMIL(512,512,128) -> Full_layer(512,1024,128) ->
NT-Xent over the global batch, backward, Adam.  Weak scaling: every rank owns 128 slides; embeddings are
all-gathered (NCCL) for the loss and gradients all-reduced once per step.

One JSON line on stdout (rank 0).  `value` = device-timed throughput with the slides already in HBM;
`e2e` = the same step driven from PINNED HOST buffers (H2D of the step's slides inside the timed region,
double-buffered on a copy stream, loss read back every step).
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
    ap.add_argument("--feat-size", type=int, default=1024)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--clusters", type=int, default=10)
    ap.add_argument("--min-patches", type=int, default=500)
    ap.add_argument("--max-patches", type=int, default=15500)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-bags", type=int, default=8, help="slides in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fp32-store", action="store_true", help="keep the slide store in fp32 even in bf16 mode")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_config(a, world):
    return {"workload": "cfg3: MuRCL stage-3 pre-train step (RL window selection + gather + mixup + ABMIL + Full_layer GRU "
                        "+ NT-Xent, fwd+bwd+Adam)",
            "bags_per_gpu": a.bags, "global_bags": a.bags * world, "T": a.T, "views": 2, "feat_size": a.feat_size,
            "feat_dim": a.dim, "clusters": a.clusters, "patches_per_bag": f"U[{a.min_patches},{a.max_patches}]",
            "arch": "ABMIL(512,512,128)+Full_layer(512,1024,128)", "parallelism": f"dp{world} (bags sharded per rank)",
            "l2_policy": "inputs larger than L2 (CSR store >= 1 GB/GPU, activations 268 MB each)",
            "slide_store_dtype": "bf16" if (a.precision == "bf16" and not a.fp32_store) else "f32",
            "bag_passes_per_step": a.bags * world * a.T * 2}


# ------------------------------------------------------------------------------------------------------
# synthetic slides
# ------------------------------------------------------------------------------------------------------
def make_host_batch(a, seed, pin=True):
    from murcl_b200 import synth
    from murcl_b200.csr import HostBags
    sizes = synth.camelyon_sizes(a.bags, a.min_patches, a.max_patches, seed=seed)
    feats, _clusters, labels = synth.make_bags(sizes, a.dim, a.clusters, seed=seed)
    # bf16 mode keeps the slide features as bf16 on the host too (what the first GEMM consumes): half the H2D bytes
    store_dtype = torch.bfloat16 if (a.precision == "bf16" and not a.fp32_store) else torch.float32
    return HostBags(feats, labels, a.clusters, pin=pin, dtype=store_dtype)


# ------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
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
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# the job
# ------------------------------------------------------------------------------------------------------
class Job:
    def __init__(self, a, rank, world, device):
        from murcl_b200 import dist as mdist
        from murcl_b200 import synth
        from murcl_b200.dropin import abmil, cl, losses, rlmil
        self.a, self.rank, self.world, self.device = a, rank, world, device
        torch.manual_seed(985 + rank)
        enc = abmil.ABMIL(a.dim, L=512, D=128, dim_out=128, precision=a.precision)
        enc.load_state_dict(synth.abmil_state(a.dim, 512, 128, 128, seed=985, peak=2.0))
        self.model = cl.CL(enc.to(device), projection_dim=128, n_features=512)
        fc = rlmil.Full_layer(512, 1024, True, 128)
        fc.load_state_dict(synth.full_layer_state(512, 1024, 128, seed=986))
        self.fc = fc.to(device)
        self.ppo = rlmil.PPO(a.dim, 512, 512, False, action_std=0.5, lr=1e-5, gamma=0.1, K_epochs=3, action_size=a.clusters)
        self.ppo.policy_old.load_state_dict(synth.actor_state(512, 512, a.clusters, seed=987))
        self.memories = [rlmil.Memory(), rlmil.Memory()]
        if world > 1:
            self.crit = mdist.DistributedNTXent(a.bags, 1.0)
        else:
            self.crit = losses.NT_Xent(a.bags, 1.0)
        self.params = list(self.model.parameters()) + list(self.fc.parameters())
        self.opt = torch.optim.Adam(self.params, lr=1e-4, weight_decay=1e-5, capturable=True)
        self.mdist = mdist
        self.graphs = {}
        self.launches_per_step = None

    def capture(self, stores):
        """One CUDA graph per (double-buffered) store, sharing a memory pool.  Returns False (on every rank) if the
        capture fails on any rank."""
        import torch.distributed as dist
        from murcl_b200 import pretrain
        ok = True
        try:
            pool = None
            # NCCL's watchdog thread polls events while we capture: keep the capture thread-local under torchrun
            mode = "thread_local" if self.world > 1 else "global"
            for i, st in enumerate(stores):
                g = pretrain.GraphedStep(lambda st=st: self.step(st), warmup=2 if i == 0 else 0, pool=pool,
                                         capture_error_mode=mode)
                pool = g.pool()
                self.graphs[id(st)] = g
                self.launches_per_step = g.launches
        except Exception as e:                                  # noqa: BLE001 - report and fall back to eager launches
            sys.stderr.write(f"[bench] rank {self.rank}: CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}\n")
            ok = False
        torch.cuda.synchronize()
        if self.world > 1:
            flag = torch.tensor([1 if ok else 0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if not ok:
            self.graphs = {}
        return ok

    def run(self, store):
        g = self.graphs.get(id(store))
        return g() if g is not None else self.step(store)

    def step(self, store):
        from murcl_b200 import pretrain
        self.opt.zero_grad(set_to_none=True)
        loss, _ = pretrain.pretrain_step(store, self.model, self.fc, self.crit, T=self.a.T, feat_size=self.a.feat_size,
                                         alpha=0.9, stage=3, ppo=self.ppo, memories=self.memories,
                                         precision=self.a.precision)
        if self.world > 1:
            self.mdist.allreduce_grads(self.params)
        self.opt.step()
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


def roofline_probe(job, store, peaks):
    """Times every dense-layer launch of one extra step with CUDA events on the launch stream and reports the
    dominant kernel family (the encoder/attention GEMMs) against the measured tensor peak."""
    from murcl_b200 import ops
    log = []
    ops.set_profile(log)
    torch.cuda.nvtx.range_push("murcl_probe_step")      # ncu --nvtx --nvtx-include "murcl_probe_step/" profiles exactly this step
    try:
        job.step(store)
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
    fam = {k: {"launches": v[2], "ms": round(v[1], 3), "tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1)} for k, v in by.items()
           if k != "attnpool_fwd"}
    hbm = peaks.get("hbm_gbs") or 6500.0
    if "attnpool_fwd" in by:     # the fused attention-pooling forward is HBM-bound: its record carries algorithmic bytes
        v = by["attnpool_fwd"]
        gbs = v[0] / (v[1] * 1e-3) / 1e9
        fam["attnpool_fwd"] = {"launches": v[2], "ms": round(v[1], 3), "bound": "hbm", "algorithmic_gbs": round(gbs, 1),
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
    return {"value": round(a.cpu_bags / best, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{a.cpu_bags} slides (of {a.bags}), same size distribution, T={a.T} x 2 views, fwd+bwd, "
                      f"best of {reps}, torch CPU fp32 oracle, {best:.2f} s/step"}, best


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
    for _ in range(min(a.warmup, 1)):
        O.pretrain_step(feats, clusters, sd_m, sd_f, T=1, feat_size=a.feat_size, generator=g)
    t0 = time.perf_counter()
    steps = max(1, min(a.steps, 3))
    for _ in range(steps):
        O.pretrain_step(feats, clusters, sd_m, sd_f, T=a.T, feat_size=a.feat_size, generator=g)
    dt = (time.perf_counter() - t0) / steps
    v = round(a.cpu_bags / dt, 3)
    sample = (f"{a.cpu_bags} slides per step (bounded sample of the {a.bags}-slide step), T={a.T} x 2 views, fwd+bwd, "
              f"{steps} timed steps, torch CPU fp32 oracle port, {threads} threads")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
                      "warmup": min(a.warmup, 1), "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
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
    from murcl_b200.csr import BagStore
    _lib.load()
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())

    host = [make_host_batch(a, seed=1000 + 10 * rank + i) for i in range(2)]
    stores = [BagStore.empty_like_host(h, device) for h in host]
    for s, h in zip(stores, host):
        s.copy_from_host(h)
    torch.cuda.synchronize()
    vlog("host batches staged, stores on device")
    job = Job(a, rank, world, device)
    vlog("job built")

    # ---- device-resident throughput --------------------------------------------------------------
    graphed = (not a.no_graph) and job.capture(stores)
    vlog(f"graph capture: {graphed}")
    for i in range(a.warmup):
        job.run(stores[i % 2])
    torch.cuda.synchronize()
    vlog("warm-up done")
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms = timed(lambda i: job.run(stores[i % 2]), a.steps, world, device)
    launches = job.launches_per_step * a.steps if graphed else _lib.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    value = a.bags * world * a.steps / (ms * 1e-3)
    vlog(f"timed region done: {ms / a.steps:.2f} ms/step")

    # ---- end to end from pinned host buffers -------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        copy_stream = torch.cuda.Stream(device)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        h2d = host[0].nbytes
        losses = []

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i % 2])
                stores[i % 2].copy_from_host(host[i % 2])
                ready[i % 2].record(copy_stream)

        def e2e_step(i):
            if i == 0:
                prefetch(0)
            if i + 1 < a.e2e_steps:
                prefetch(i + 1)                      # next batch's H2D overlaps this step's compute
            torch.cuda.current_stream().wait_event(ready[i % 2])
            loss = job.run(stores[i % 2])
            freed[i % 2].record()
            losses.append(float(loss.item()))        # D2H read of the step's result

        for f in freed:
            f.record()
        ms_e2e = timed(e2e_step, a.e2e_steps, world, device)
        e2e = {"value": round(a.bags * world * a.e2e_steps / (ms_e2e * 1e-3), 2), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": a.e2e_steps,
               "ms_per_step": round(ms_e2e / a.e2e_steps, 3),
               "note": "per step: H2D of the step's slides (CSR features + cluster ids) from pinned memory on a copy stream, "
                       "double-buffered against compute; loss.item() each step"}

    roof = roofline_probe(job, stores[0], peaks)      # every rank runs it: the step contains collectives
    vlog("roofline probe done")
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
