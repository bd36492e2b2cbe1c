"""Benchmark of the hot path: joint network + transducer loss, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

Prints ONE JSON line (rank 0).  A step is one joint -> RNNTLoss -> backward pass over one synthetic
batch of the named workload (weak scaling: every GPU gets the full per-GPU batch).  See DESIGN.md
"Measurement" for the definition of every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the microbench the headline metric is quoted on (espnet joint dims)
    "cfg2": dict(B=32, T=400, U=40, V=4232, D=512, H=512, joint="espnet", ragged=False,
                 desc="joint+RNN-T loss microbench B=32 T=400 U=40 V=4232 D=512 H=512 fp32"),
    # same lattice, joint width 256 (leaves shared memory for a deeper TMA ring: pipeline experiments)
    "cfg2h256": dict(B=32, T=400, U=40, V=4232, D=512, H=256, joint="espnet", ragged=False,
                     desc="B=32 T=400 U=40 V=4232 D=512 H=256 fp32 (experiment)"),
    # BASELINE.json configs[0]: aishell.yaml joint (tt JointNet 1024 -> 1024 -> V), the reference's CPU-runnable case
    "cfg1": dict(B=4, T=200, U=30, V=4232, D=512, H=1024, joint="tt", ragged=False,
                 desc="aishell.yaml joint B=4 T=200 U=30 V=4232 D=512 H=1024 fp32"),
    # BASELINE.json configs[2]: joint_streaming.yaml joint dims (H=2048, V=6485) at the global batch of 64
    "cfg3": dict(B=64, T=410, U=42, V=6485, D=512, H=2048, joint="tt", ragged=False,
                 desc="joint_streaming.yaml joint B=64 T=410 U=42 V=6485 D=512 H=2048 fp32"),
    # BASELINE.json configs[0] shapes but with a fused-path joint width (parity-sized smoke workload)
    "small": dict(B=4, T=200, U=30, V=4232, D=512, H=512, joint="espnet", ragged=False,
                  desc="B=4 T=200 U=30 V=4232 D=512 H=512 fp32"),
    # BASELINE.json configs[3]: long-utterance stress (dense logits would be 54 GB)
    "cfg4": dict(B=16, T=1000, U=200, V=4232, D=512, H=512, joint="espnet", ragged=False,
                 desc="long-utterance stress B=16 T=1000 U=200 V=4232 D=512 H=512 fp32"),
    # BASELINE.json configs[4]: ragged batch, T in [50,1000], U in [5,200], bf16 joint inputs (espnet dims, V=4233)
    "cfg5": dict(B=32, T=1000, U=200, V=4233, D=512, H=512, joint="espnet", ragged=True, bf16=True,
                 desc="ragged batch B=32 T~U[50,1000] U~U[5,200] V=4233 D=512 H=512 bf16 joint inputs"),
}


def peaks():
    """Roofline denominators: the sustained cuBLAS figure is the one for a kernel timed inside a long step (the roofline
    'frac'); the burst figure is reported next to it ('frac_burst') -- the timed region here is short."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (frac: bf16_tflops_sustained, frac_burst: bf16_tflops)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback of B200_PROFILING.md")


def synth(w, seed, device=None, pin=False):
    g = torch.Generator().manual_seed(seed)
    B, T, U, V, D = w["B"], w["T"], w["U"], w["V"], w["D"]
    enc = torch.randn(B, T, D, generator=g)
    pred = torch.randn(B, U + 1, D, generator=g)
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
    act_lens = torch.full((B,), T, dtype=torch.int32)
    label_lens = torch.full((B,), U, dtype=torch.int32)
    if w.get("ragged"):
        act_lens = torch.randint(50, T + 1, (B,), generator=g, dtype=torch.int32)
        label_lens = torch.randint(5, U + 1, (B,), generator=g, dtype=torch.int32)
        act_lens[0], label_lens[0] = T, U          # one utterance at the maxima so the length checks pass
        for i in range(B):
            labels[i, int(label_lens[i]):] = -1      # tt/dataset.py:46-48 padding
    if w.get("bf16"):
        enc, pred = enc.bfloat16(), pred.bfloat16()
    out = [enc, pred, labels, act_lens, label_lens]
    if pin:
        out = [t.pin_memory() for t in out]
    if device is not None:
        out = [t.to(device) for t in out]
    return out


class ClockSampler:
    """SM clock, board power and throttle reasons DURING the timed region: NVML polled from a thread every 5 ms (the
    default timed region is ~150 ms, `nvidia-smi -lms 100` sees one or two samples of it); `nvidia-smi` when NVML's
    Python binding is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.proc = None
        self.thread = None
        self.rows = []                       # (sm MHz, power W, reason bits)
        self.maxc = None
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = nv.nvmlDeviceGetHandleByIndex(index)
            self.maxc = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                         nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
            self.halt = threading.Event()

            def poll():
                while not self.halt.is_set():
                    try:
                        self.rows.append((float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)),
                                          nv.nvmlDeviceGetPowerUsage(handle) / 1000.0,
                                          int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))))
                    except Exception:
                        pass
                    self.halt.wait(0.005)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def _summary(self, clocks, power, reasons, maxc, source):
        load = [c for c, p in zip(clocks, power) if p > 300.0] or clocks
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": maxc,
                "reasons": sorted(reasons), "samples": len(clocks), "power_w_max": max(power) if power else None,
                "source": source}

    def stop(self):
        if self.thread is not None:
            self.halt.set()
            self.thread.join(timeout=2)
            reasons = {n for _, _, r in self.rows for n, b in zip(self.NAMES, self.bits) if r & b}
            return self._summary([r[0] for r in self.rows], [r[1] for r in self.rows], reasons, self.maxc, "nvml, 5 ms")
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        clocks, maxc, reasons, power = [], None, set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clocks.append(float(f[0]))
                maxc = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return self._summary(clocks, power, reasons, maxc, "nvidia-smi -lms 100")


def _cpu_joint(w):
    """The reference's own joint module (unmodified, from the staged copy baseline/_ref/ -- /root/reference does not exist
    on the GPU box) or, without it, the oracle's restatement."""
    from oracle import joint_ref, ref_import
    tt = w.get("joint") == "tt"
    if ref_import.available():
        try:
            if tt:
                return ref_import.tt_model().JointNet(2 * w["D"], w["H"], w["V"]), "tt.model.JointNet (unmodified reference)"
            jn = ref_import.espnet_joint_module().JointNetwork
            return (jn(w["V"], w["D"], w["D"], w["H"], "tanh"),
                    "espnet...joint_network.JointNetwork (unmodified reference)")
        except Exception:                          # a stale or partial copy: the restatement is pinned to it by the tests
            pass
    if tt:
        return joint_ref.TTJointNet(2 * w["D"], w["H"], w["V"]), "oracle/joint_ref.TTJointNet"
    return joint_ref.EspnetJointNetwork(w["V"], w["D"], w["D"], w["H"], "tanh"), "oracle/joint_ref.EspnetJointNetwork"


def cpu_port_step(w, B_s, seed=1234, loss_impl="port"):
    """One step of the reference's CPU path on B_s utterances of workload w (reference joint + the oracle's port of the
    un-vendored warprnnt_pytorch loss, or torchaudio's independent CPU rnnt_loss); returns seconds."""
    from oracle import rnnt_oracle
    ws = dict(w, B=B_s, bf16=False)
    enc, pred, labels, act_lens, label_lens = synth(ws, seed)
    torch.manual_seed(seed)
    tt = w.get("joint") == "tt"
    joint, cpu_port_step.joint_name = _cpu_joint(w)
    crit = rnnt_oracle.RNNTLoss(blank=0)
    if loss_impl == "torchaudio":
        import torchaudio

        def crit(logits, labels, act_lens, label_lens):                     # noqa: F811
            return torchaudio.functional.rnnt_loss(logits, labels.clamp_min(0), act_lens, label_lens, blank=0,
                                                   reduction="mean")
    enc.requires_grad_()
    pred.requires_grad_()
    t0 = time.perf_counter()
    logits = joint(enc, pred) if tt else joint(enc[:, :, None], pred[:, None])
    loss = crit(logits, labels, act_lens, label_lens)
    loss.backward()
    loss.item()
    return time.perf_counter() - t0


def cpu_baseline(w, budget_s=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ["OMP_NUM_THREADS"] = str(cores)     # before the oracle's OpenMP library loads (torchrun sets it to 1)
    t1 = cpu_port_step(w, 1)                       # warm-up + calibration on one utterance
    B_s = int(max(1, min(w["B"], budget_s / 2 / max(t1, 1e-3))))
    ts = [cpu_port_step(w, B_s) for _ in range(2)]
    t = min(ts)
    out = {"value": B_s / t, "unit": "utt/s", "cores": cores, "kind": "port",
           "sample": "%d utterances of the workload per step (T=%d U=%d V=%d H=%d), best of 2 steps, %s + "
                     "oracle/rnnt_cpu.c (OpenMP port of warprnnt_pytorch)" %
                     (B_s, w["T"], w["U"], w["V"], w["H"], cpu_port_step.joint_name)}
    try:        # a second, independent CPU number: the same joint with torchaudio's CPU rnnt_loss
        cpu_port_step(w, 1, loss_impl="torchaudio")
        t2 = min(cpu_port_step(w, B_s, loss_impl="torchaudio") for _ in range(2))
        out["torchaudio_rnnt_loss"] = {"value": B_s / t2, "unit": "utt/s", "sample": "same joint and sample, "
                                       "torchaudio.functional.rnnt_loss (CPU) as the loss"}
    except Exception as e:                        # torchaudio missing or without its CPU op
        out["torchaudio_rnnt_loss"] = {"unavailable": repr(e)[:120]}
    return out


def run_reference(args, w):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python reference and
    its un-vendored warprnnt_pytorch dependency cannot travel to the GPU box), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ["OMP_NUM_THREADS"] = str(cores)     # before the oracle's OpenMP library loads (torchrun sets it to 1)
    t1 = cpu_port_step(w, 1)
    budget = 150.0 / max(1, args.steps + args.warmup)
    B_s = int(max(1, min(w["B"], budget / max(t1, 1e-3))))
    for _ in range(args.warmup):
        cpu_port_step(w, B_s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(w, B_s)
    dt = time.perf_counter() - t0
    val = B_s * args.steps / dt
    sample = "%d utterances per step of %s; %s + oracle/rnnt_cpu.c (OpenMP port of warprnnt_pytorch)" % (
        B_s, w["desc"], cpu_port_step.joint_name)
    print(json.dumps({
        "impl": "reference", "metric": "joint+RNN-T loss fwd+bwd utterances/s", "value": val, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, int(os.environ.get("WORLD_SIZE", "1")), args),      # the same dict as the GPU arm's
        "cpu_baseline": {"value": val, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(w, world, args):
    """The `config` object of a bench line: the workload only (both arms print the same dict for the same arguments)."""
    B = w["B"] // world if args.scaling == "strong" else w["B"]
    _, _, _, act_lens, label_lens = synth(dict(w, B=B), 1234)                # lengths of rank 0's shard
    M = int((act_lens.long() * (label_lens.long() + 1)).sum())
    return {"workload": w["desc"] + (" (global batch, sharded)" if args.scaling == "strong" else ""),
            "global_batch": world * B, "parallelism": "dp%d" % world, "logits": args.logits, "route": args.route,
            "l2": "per-step working set (A16 %.2f GB + dA %.2f GB) exceeds the 126 MB L2; no explicit flush" %
                  (M * w["H"] * 2 / 1e9, M * w["H"] * 4 / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU gets the workload's batch; strong: the workload's batch is sharded over the GPUs")
    ap.add_argument("--logits", default="init", choices=["init", "peaked"],
                    help="peaked: trained-like output layer (larger weights, boosted blank / label biases)")
    ap.add_argument("--breakdown", action="store_true", help="also print the per-kernel table to stderr")
    ap.add_argument("--no-overlap", action="store_true", help="A/B runs: every kernel on the launching stream")
    ap.add_argument("--route", default="default", choices=["default", "fused", "chunked"],
                    help="A/B runs: fused = the recomputing kernels at H = 512 (nothing V-wide in HBM), chunked = library GEMMs")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    args.warmup = max(args.warmup, 3)

    import transformer_transducer_b200 as ttb
    from transformer_transducer_b200 import functional as F
    if args.route != "default":
        F.ROUTE = args.route
    if args.no_overlap:
        F.WideJointRNNT.OVERLAP = False
        F.MERGE_PROJ_BACKWARD = False

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    w0 = w
    if args.scaling == "strong":
        if w["B"] % world:
            raise SystemExit("--scaling strong: the workload's batch %d is not divisible by %d GPUs" % (w["B"], world))
        w = dict(w, B=w["B"] // world, desc=w["desc"] + " (global batch, sharded)")
    torch.manual_seed(1234)
    tt = w.get("joint") == "tt"
    joint = (ttb.JointNet(2 * w["D"], w["H"], w["V"]) if tt else
             ttb.JointNetwork(w["V"], w["D"], w["D"], w["H"], "tanh")).to(dev)
    if args.logits == "peaked":
        # a trained model's output layer rather than nn.Linear's initialisation: logits several nats wide, the blank and a
        # few hundred frequent units boosted -- exercises the moving softmax reference of the forward kernels
        out_layer = joint.project_layer if tt else joint.lin_out
        with torch.no_grad():
            out_layer.weight.mul_(3.0)
            out_layer.bias[0] += 4.0
            out_layer.bias[torch.randperm(w["V"])[: w["V"] // 16].to(dev)] += 6.0
    if w.get("bf16"):
        joint = joint.bfloat16()

    def logits_of(e, p_):
        return model(e, p_) if tt else model(e[:, :, None], p_[:, None])
    model = joint
    if world > 1:
        # lin_out.weight (V x H fp32, 8.7 MB at configs[1]) is final first: its own bucket, so that its all-reduce runs
        # under the pre-projections' backward instead of after it
        model = torch.nn.parallel.DistributedDataParallel(joint, device_ids=[local], gradient_as_bucket_view=True,
                                                          bucket_cap_mb=8)
    crit = ttb.RNNTLoss(blank=0, reduction="mean")
    enc, pred, labels, act_lens, label_lens = synth(w, 1234 + rank, device=dev)
    enc.requires_grad_()
    pred.requires_grad_()
    host = synth(w, 1234 + rank, pin=True)
    h2d = sum(t.numel() * t.element_size() for t in host)

    def step_resident():
        for p_ in joint.parameters():
            p_.grad = None
        enc.grad = None
        pred.grad = None
        loss = crit(logits_of(enc, pred), labels, act_lens, label_lens)
        loss.backward()
        return loss

    # End to end like a training loop with a pinned-memory loader: every step's inputs are copied host -> device inside the
    # timed region, on a copy stream, while the previous step computes (the reference's DataLoader prefetches the same way);
    # the step then waits for its own copy, runs through the public API and reads its loss back.
    copy_stream = torch.cuda.Stream(dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            ts = [t.to(dev, non_blocking=True) for t in host]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ts, ev

    pending = []
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def step_e2e():
        for p_ in joint.parameters():
            p_.grad = None
        (e, p_, lab, al, ll), ev = pending.pop() if pending else fetch()
        torch.cuda.current_stream(dev).wait_event(ev)
        e.requires_grad_()
        p_.requires_grad_()
        loss = crit(logits_of(e, p_), lab, al, ll)
        # device -> host read of the step's result: the loss is final when the forward is, so its copy is queued here and
        # waited for once the backward has been launched -- the host then prepares the next step while the backward runs
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        read = torch.cuda.Event()
        read.record()
        pending.append(fetch())                                   # the next step's inputs, copied under this step's kernels
        loss.backward()
        read.synchronize()
        return float(loss_host[0])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    # Stand-alone kernel times: in the timed region the bandwidth-bound kernels run NEXT TO the products on a second
    # stream, so the events around them measure a stretched, overlapped duration.  Three untimed steps with everything on
    # one stream give each kernel's own time (used for the per-kernel rooflines below).
    alone = {}
    if getattr(F.WideJointRNNT, "OVERLAP", False):
        F.WideJointRNNT.OVERLAP = False
        F.PROFILE = prof0 = []
        for _ in range(3):
            step_resident()
        barrier()
        F.PROFILE = None
        F.WideJointRNNT.OVERLAP = True
        for name, e0, e1, _nk in prof0:
            alone.setdefault(name, []).append(e0.elapsed_time(e1))
        alone = {k: sum(v) / len(v) for k, v in alone.items()}
        step_resident()
        barrier()
    # (everything only one rank does goes IN FRONT of the barrier: a rank that enters the timed region late keeps the others
    # waiting in their first all-reduce, and the max over ranks then carries that wait -- 0.3 - 0.5 ms per step over 20 steps
    # when the sampler's start-up sat behind the barrier)
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    F.PROFILE = prof = []
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    F.PROFILE = None
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])

    # end-to-end through the public API with host buffers (H2D of the inputs + D2H of the loss every step)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])

    # per-kernel table from the events recorded inside the timed region
    per = {}
    launches = 0
    for name, e0, e1, nk in prof:
        per.setdefault(name, []).append(e0.elapsed_time(e1))
        launches += nk
    table = {k: {"calls": len(v), "avg_ms": sum(v) / len(v)} for k, v in per.items()}
    M = int((act_lens.long() * (label_lens.long() + 1)).sum())      # lattice cells actually present
    unit_flops = 2.0 * M * w["H"] * w["V"]           # one M x H x V contraction (F = 3 of these)
    pk = peaks()
    dom = max((k for k in table if k.startswith("ttx_joint_") or k.startswith("ttx_rows_") or k.startswith("ttx_wide_")),
              key=lambda k: table[k]["avg_ms"] * table[k]["calls"])
    dom_ms = alone.get(dom, table[dom]["avg_ms"])
    # algorithmic contractions (2*M*H*V each) one launch of the kernel accounts for: the fused forward+gradient
    # launch does the forward projection AND the dL/dA contraction; the chunked path's row kernels do none themselves
    alg_units = {"ttx_joint_fwd_grad": 2.0, "ttx_rows_lse": 0.0, "ttx_rows_grad": 0.0, "ttx_wide_sp": 1.0, "ttx_wide_pw": 1.0,
                 "ttx_wide_dw": 1.0}.get(dom, 1.0)
    unit_flops_dom = unit_flops * alg_units
    achieved = unit_flops_dom / (dom_ms * 1e-3) / 1e12
    traffic = tensor_pct = None
    tpath = os.path.join(ROOT, "profiles", "traffic_bytes.json")
    if os.path.exists(tpath):                    # ncu captures of an earlier run of this command (not measured live)
        tj = json.load(open(tpath))
        traffic = tj.get(args.workload, {}).get(dom)
        tensor_pct = tj.get(args.workload + "_tensor_pipe_active_pct", {}).get(dom)
    step_ms = ms / args.steps
    out = {
        "metric": "joint+RNN-T loss fwd+bwd utterances/s",
        "value": world * w["B"] * args.steps / (ms * 1e-3), "unit": "utt/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": ("bf16 tensor-core operands" if w.get("bf16") else "f16 tensor-core operands") +
                 ", f32 accumulate/softmax, 48-bit (float + float) lattice carrier (%s variant)" % ("bf16-input" if w.get("bf16") else "fp32"),
        "data": "synthetic",
        "config": workload_config(w0, world, args),
        "e2e": {"value": world * w["B"] * args.steps / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tflops"], "frac_burst": achieved / pk["burst"], "peak_burst": pk["burst"],
                     "traffic": traffic, "traffic_source": "profiles/traffic_bytes.json (ncu capture, see its _source)" if traffic else None,
                     "tensor_pipe_active_pct_ncu": tensor_pct, "kernel_ms": dom_ms,
                     "algorithmic_flops_per_launch": unit_flops_dom, "peak_source": pk["source"]},
        "roofline_step": {"algorithmic_flops": 3 * unit_flops, "achieved": 3 * unit_flops / (step_ms * 1e-3) / 1e12,
                          "frac": 3 * unit_flops / (step_ms * 1e-3) / 1e12 / pk["tflops"],
                          "frac_burst": 3 * unit_flops / (step_ms * 1e-3) / 1e12 / pk["burst"]},
        "kernels": table,
        "kernels_standalone_ms": alone or None,
    }
    if "ttx_lattice_fwd_bwd" in table:
        # lattice wavefront: 8 B read (2 fp32 log-probs) + 16 B written (fp64 alpha, beta) per cell; it is bound by
        # the T+U1-1 dependent steps, not by HBM -- reported against the HBM peak as the north-star asks
        lat_ms = alone.get("ttx_lattice_fwd_bwd", table["ttx_lattice_fwd_bwd"]["avg_ms"])      # stand-alone, not overlapped
        gbs = 24.0 * M / (lat_ms * 1e-3) / 1e9
        out["roofline_lattice"] = {"bound": "hbm (latency-bound in practice)", "achieved": gbs, "peak": pk["hbm"],
                                   "unit": "GB/s", "frac": gbs / pk["hbm"], "kernel_ms": lat_ms,
                                   "dependent_steps": int(w["T"] + w["U"])}
    if dist is not None:
        # DDP keeps the ranks in lockstep: the step time is the slowest GPU's.  Each rank's own product time (events around
        # its launches) shows which one that is.
        mine = {"rank": rank, "products_ms": sum(v["avg_ms"] for k, v in table.items() if k.startswith("ttx_wide_")
                                                 or k.startswith("ttx_joint_")),
                "kernels_ms": sum(v["avg_ms"] * v["calls"] for v in table.values()) / max(1, args.steps)}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        out["ranks"] = allr
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(w)
        if args.breakdown:
            for k, v in sorted(table.items(), key=lambda kv: -kv[1]["avg_ms"]):
                print("%-28s %8.3f ms x %d" % (k, v["avg_ms"], v["calls"]), file=sys.stderr)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
