#!/usr/bin/env python
"""bench.py -- throughput of the batched voice path on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload svf|biquad|delay|mfcc] [--impl reference]

A "step" is one block (1024 frames) of one bank: BASELINE.json configs[1] by default -- 1 Mi voices of
maxiOsc::saw -> maxiSVF (low-pass mix), per-voice fp64 output materialised time-major in HBM. With N > 1
(launched by torch.distributed.run, one rank per GPU) every rank owns its own 1 Mi voices (weak scaling);
voices are independent, so the headline block has no collective at any N.

  value      voice-samples/s of the whole job, inputs and state resident in HBM, CUDA-event timed, max over ranks
  mixdown    BASELINE.json configs[4] in "mix mode" (SURVEY.md 8(d) config 5), timed separately in the same run:
             1 Mi voices/GPU saw -> maxiBiquad -> maxiMix::stereo -> sum over voices, only the stereo bus is
             written; for N > 1 the bus is summed over the GPUs -- the path's only exchange step (--collective
             p2p: peer-memory exchange fused into the mix-reduce kernel; nccl: all_reduce). fp64-pipe bound,
             reported as such. `--mix 1` instead adds the bus to the headline block (output AND bus).
  e2e        the same block through the C ABI with HOST control data: per step one fp64 frequency array
             (8 MiB) is uploaded stream-ordered with mxb_bank_set_param_async from pinned memory and the stereo
             mix (16 KiB) is read back by mxb_bank_process(MXB_MEM_SPLIT); the voice signals stay on the device
  roofline   algorithmic bytes per launch / CUDA-event launch time against MEASURED_PEAKS.json
  cpu_baseline  the reference's own scalar code (oracle/_ref, or the C port) on all host cores, bounded sample

`--impl reference` times only that CPU leg with the same metric/config (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 1024
SR = 48000

WORKLOADS = {
    # name: (voices/GPU, osc, filt, env, delay_taps, algorithmic bytes per voice-sample, description)
    "svf": dict(voices=1 << 20, osc="saw", filt="svf", env=False, delay=0, bytes_per=8.0 + (88 + 32) / BLOCK,
                desc="configs[1]: 1Mi voices maxiOsc::saw -> maxiSVF lp, 1024-frame block, fp64 out[1024][V] materialised"),
    "biquad": dict(voices=1 << 20, osc="saw", filt="biquad", env=False, delay=0, bytes_per=8.0 + (80 + 24) / BLOCK,
                   desc="configs[4] shard: 1Mi voices/GPU maxiOsc::saw -> maxiBiquad lowpass, fp64 out materialised + stereo mix"),
    "delay": dict(voices=1 << 18, osc="saw", filt="none", env=True, delay=4096, bytes_per=24.0 + (108 + 60) / BLOCK,
                  desc="configs[2]: 256Ki voices saw -> maxiEnv::adsr -> maxiDelayline::dl(4096), fp64 out materialised"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel of workload `key`,
    from the committed `ncu --set full` capture of this same bench command (profiles/traffic.json names the summary
    file each figure was read from). None when no capture of that configuration is committed."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        return {"traffic": float(d["dram_bytes_per_launch"]), "traffic_unit": "bytes/launch", "traffic_source": d["source"]}
    except Exception:
        return {"traffic": None}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons), "samples": len(rows)}


# ------------------------------------------------------------------------------------------- CPU reference leg

class CpuFarm:
    """The reference on all host cores: `threads` workers, each owning its own slice of the voices as its own reference
    objects and its own output slab, all created INSIDE the worker (first touch = local NUMA node). One thread start per
    run; every worker runs all blocks of its voices (voices are independent: no barrier between blocks)."""

    def __init__(self, wl, voices, threads, kind):
        from maximilian_b200 import workloads as W
        from oracle import oracle_py as O
        self.wl, self.voices, self.kind = wl, voices, kind
        self.threads = max(1, min(threads, voices))
        self.bounds = np.linspace(0, voices, self.threads + 1).astype(int)
        p = W.voice_params(voices, seed=W.SEED, delay_size=wl["delay"] or 4096)
        self.parts = [None] * self.threads

        def make(i):
            lo, hi = int(self.bounds[i]), int(self.bounds[i + 1])
            b = O.Bank(hi - lo, osc=wl["osc"], filt=wl["filt"], env=wl["env"], delay=wl["delay"] > 0, sample_rate=SR,
                       delay_capacity=max(wl["delay"], 1), kind=kind)
            W.configure_bank(b, wl["filt"], {k: v[lo:hi] for k, v in p.items()}, wl["env"], wl["delay"] > 0)
            self.parts[i] = (b, np.zeros((BLOCK, hi - lo), dtype=np.float64))
        self._par(make)

    def _par(self, fn):
        ts = [threading.Thread(target=fn, args=(i,)) for i in range(self.threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]

    def run(self, nblocks, block_index0=0):
        """Seconds for `nblocks` consecutive blocks of the whole bank."""
        from maximilian_b200 import workloads as W
        from oracle import oracle_py as O
        gates = [(W.gate(self.voices, BLOCK, block_index0 + k) if self.wl["env"] else (None, None)) for k in range(nblocks)]

        def work(i):
            lo, hi = int(self.bounds[i]), int(self.bounds[i + 1])
            b, out = self.parts[i]
            for on, off in gates:
                on_i = np.ascontiguousarray(on[lo:hi]) if on is not None else None
                off_i = np.ascontiguousarray(off[lo:hi]) if off is not None else None
                rc = b.lib.mxo_bank_process(b.h, BLOCK, O._ip(on_i), O._ip(off_i), O._dp(out), None, 0, hi - lo)
                assert rc == 0, rc
        t0 = time.perf_counter()
        self._par(work)
        return time.perf_counter() - t0


def cpu_kind():
    from oracle import oracle_py as O
    if O.available("reference"):
        return "reference"
    O.build("port")
    return "port"


def cpu_sample_voices(wl):
    # each reference maxiDelayline is a 5.6 MB object (src/maximilian.h:273): keep the sample in RAM
    return 512 if wl["delay"] else 65536


def cpu_baseline(wl, budget_s=4.0):
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    voices = cpu_sample_voices(wl)
    farm = CpuFarm(wl, voices, cores, kind)
    farm.run(1)                                                      # warm-up block
    n, total = 0, 0.0
    while total < budget_s and n < 128:
        total += farm.run(8, 1 + n)                                  # 8 blocks per thread start
        n += 8
    v = voices * BLOCK * n / total
    return {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": f"{voices} voices x {BLOCK} frames x {n} blocks of the same chain; {farm.threads} host threads, each with its own "
                      "slice of the voices as reference objects (frame-outer / voice-inner like a reference play())"}


def reference_arm(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    voices = cpu_sample_voices(wl)
    farm = CpuFarm(wl, voices, cores, kind)
    farm.run(args.warmup)
    dt = farm.run(args.steps, args.warmup)
    v = voices * BLOCK * args.steps / dt
    line = {"impl": "reference", "metric": "voice_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "cpu_sample_voices": voices, "block": BLOCK, "sample_rate": SR},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                             "sample": f"each step = {voices} voices x {BLOCK} frames of the same chain on {cores} host threads"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU leg

def onbox_peaks(torch, out):
    """the streaming-WRITE and copy bandwidth of THIS box, measured with library kernels on the output
    buffer the bank just wrote (torch fill_ / copy_, best of 5, CUDA events), reported beside the driver-measured copy
    peak: the bank kernel is write-only traffic, which the copy figure does not isolate."""
    try:
        flat = out.view(-1)
        half = flat.numel() // 2
        a, b = flat[:half], flat[half:2 * half]
        best_fill = best_copy = 0.0
        for _ in range(5):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(); flat.fill_(0.0); e1.record(); b.copy_(a); e2.record()
            torch.cuda.synchronize()
            best_fill = max(best_fill, flat.numel() * flat.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
            best_copy = max(best_copy, 2 * half * flat.element_size() / (e1.elapsed_time(e2) * 1e-3) / 1e9)
        return {"fill_gbs": best_fill, "copy_gbs": best_copy,
                "how": "torch fill_ over the %.1f GB output buffer (write-only) and copy_ of one half onto the other (read+write bytes), best of 5" % (flat.numel() * flat.element_size() / 1e9)}
    except Exception as e:      # a missing number must not take the bench line down
        return {"error": str(e)[:200]}


def measure_mixdown(args, torch, dist, capi, W, ctx, dev, rank, world, mix, stream, barrier, sm_mhz):
    """BASELINE.json configs[4] in SURVEY.md 8(d)'s "mix mode", one shard per GPU: 1 Mi voices maxiOsc::saw ->
    maxiBiquad -> maxiMix::stereo -> sum over voices; no per-voice output is written, the stereo bus [1024][2] is the
    result. With N > 1 the bus is summed over the GPUs (the path's only exchange step). The kernel moves 0.1 B per
    voice-sample, so it is reported against the fp64 pipe, not against HBM; it is never
    mixed into the headline."""
    wl = WORKLOADS["biquad"]
    V = wl["voices"]
    p = W.voice_params(V, seed=W.SEED + 100 + rank)
    bank = capi.Bank(V, osc=wl["osc"], filt=wl["filt"], env=False, delay=False, delay_capacity=1, max_frames=BLOCK,
                     ctx=ctx, sample_rate=SR)
    W.configure_bank(bank, wl["filt"], p, False, False)
    p2p = world > 1 and args.collective == "p2p"
    if p2p:
        exch = capi.Exchange(ctx, rank, world, max_doubles=2 * BLOCK)
        exch.connect_with_torch_distributed()      # the IPC handles travel over the process group; the data never does
        exch.attach(bank)
    steps = max(3, min(args.steps, 50))

    def step():
        bank.process_device(BLOCK, out_ptr=None, mix_ptr=mix.data_ptr(), stream=stream.cuda_stream)
        if world > 1 and not p2p:
            dist.all_reduce(mix)

    for _ in range(3):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    # fp64-pipe instructions per voice-sample in the SASS of this instantiation: saw 3 (DSETP, 2 DADD), biquad 9
    # (5 DMUL, 4 DADD: -fmad=false keeps the reference's roundings), mix 2 (DMUL/DFMA per channel) + 1 (row sums)
    instr = 15.0
    rate = instr * V * BLOCK / (ms * 1e-3)                    # fp64 lane-instructions per second per GPU
    pipe_peak = 148 * 64 * (sm_mhz or 1920.0) * 1e6           # 64 fp64 lanes per SM (16 per sub-partition), one instr per lane-clock
    res = {"workload": "configs[4] shard, mix mode: 1Mi voices/GPU maxiOsc::saw -> maxiBiquad lowpass -> maxiMix::stereo -> sum over voices, "
                       "stereo bus [1024][2] fp64 per block, no per-voice output",
           "value": world * V * BLOCK / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "steps": steps,
           "bound": "fp64 pipe", "fp64_instr_per_voice_sample": instr, "fp64_pipe_frac_per_gpu": rate / pipe_peak,
           "fp64_pipe_peak": "148 SMs x 64 lanes x %.0f MHz (sampled SM clock)" % (sm_mhz or 1920.0),
           "collective": ("none (one GPU)" if world == 1 else
                          "peer-memory exchange of mix[1024][2] fp64 fused into the mix-reduce kernel (CUDA IPC over NVLink, rank-ordered sum)"
                          if p2p else "NCCL sum all-reduce of mix[1024][2] fp64 per block")}
    del bank
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="svf", choices=sorted(WORKLOADS) + ["mfcc"])
    ap.add_argument("--mix", type=int, default=-1,
                    help="1: the timed block also produces the stereo mix bus (+ the cross-GPU mix-down when gpus > 1). Default: the "
                         "headline block is the configured chain alone (same work at every N) and the mix-down configuration "
                         "(configs[4]) is timed separately and reported under 'mixdown'")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--f32-out", action="store_true", help="store the materialised output as fp32 (declared in the JSON)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1 mix-down: p2p = peer-memory exchange fused into the mix-reduce kernel (default); nccl = torch.distributed all_reduce")
    args = ap.parse_args()
    if args.workload == "mfcc":
        return main_mfcc(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, args.workload, wl)
        return
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"

    import torch
    import torch.distributed as dist
    from maximilian_b200 import capi
    from maximilian_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    want_mix = args.mix == 1

    V = wl["voices"]
    p = W.voice_params(V, seed=W.SEED + rank, delay_size=wl["delay"] or 4096)
    ctx = capi.Context(local, SR)
    bank = capi.Bank(V, osc=wl["osc"], filt=wl["filt"], env=wl["env"], delay=wl["delay"] > 0,
                     delay_capacity=max(wl["delay"], 1), max_frames=BLOCK, ctx=ctx, sample_rate=SR)
    W.configure_bank(bank, wl["filt"], p, wl["env"], wl["delay"] > 0)
    p2p = world > 1 and want_mix and args.collective == "p2p"
    if p2p:
        exch = capi.Exchange(ctx, rank, world, max_doubles=2 * BLOCK)
        exch.connect_with_torch_distributed()      # the IPC handles travel over the process group; the data never does
        exch.attach(bank)

    out_dtype = torch.float32 if args.f32_out else torch.float64
    out = torch.empty((BLOCK, V), dtype=out_dtype, device=dev)             # 8 GiB (fp64, 1 Mi voices) >> 126 MB L2
    mix = torch.zeros((BLOCK, 2), dtype=torch.float64, device=dev)
    gates = []
    if wl["env"]:   # gate arrays for 4 consecutive blocks, resident on the device (control data of the resident leg)
        for k in range(4):
            on, off = W.gate(V, BLOCK, k, seed=W.SEED + rank)
            gates.append((torch.from_numpy(on).to(dev), torch.from_numpy(off).to(dev)))
    stream = torch.cuda.current_stream()

    def step(k):
        on_p = off_p = None
        if gates:
            on_p, off_p = gates[k % 4][0].data_ptr(), gates[k % 4][1].data_ptr()
        bank.process_device(BLOCK, out_ptr=out.data_ptr(), mix_ptr=mix.data_ptr() if want_mix else None,
                            trig_on_ptr=on_p, trig_off_ptr=off_p, f32=args.f32_out, stream=stream.cuda_stream)
        if world > 1 and want_mix and not p2p:
            dist.all_reduce(mix)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.12)
    barrier()
    launches0 = bank.launches
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.perf_counter()
    evs[0].record(stream)
    for k in range(args.steps):
        step(args.warmup + k)
        evs[k + 1].record(stream)
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    launches = bank.launches - launches0
    ms_total = evs[0].elapsed_time(evs[-1])
    per_launch_ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    samples_per_step = V * BLOCK
    value = world * samples_per_step * args.steps / (ms_max * 1e-3)
    clocks = None
    if rank == 0:
        time.sleep(0.06)
        sampler.stop()
        clocks = sampler.summary(t_wall0, t_wall1)

    # ---- configs[4] shard: osc -> maxiBiquad + stereo mix bus (+ cross-GPU mix-down), timed on its own -------
    mixdown = None
    if args.mix < 0 and not wl["delay"]:
        mhz = torch.tensor([float((clocks or {}).get("sm_mhz") or 0.0)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.broadcast(mhz, 0)
        mixdown = measure_mixdown(args, torch, dist, capi, W, ctx, dev, rank, world, mix, stream, barrier, float(mhz.item()) or None)

    # ---- end to end through the C ABI with host control data ------------------------------------------------
    freq_host = torch.from_numpy(p["freq"].copy()).pin_memory()
    mix_host = torch.zeros((BLOCK, 2), dtype=torch.float64).pin_memory()
    fh, mh = freq_host.numpy(), mix_host.numpy()
    on_h = off_h = None
    if wl["env"]:
        on, off = W.gate(V, BLOCK, 0, seed=W.SEED + rank)
        on_t, off_t = torch.from_numpy(on).pin_memory(), torch.from_numpy(off).pin_memory()
        on_h, off_h = on_t.numpy(), off_t.numpy()
    e2e_steps = max(3, min(args.steps, 50))

    def e2e_step():
        bank.set_host_array("freq", fh, stream=stream.cuda_stream)        # H2D: this block's control data (stream-ordered)
        bank.process_split(BLOCK, out.data_ptr(), mh, on_h, off_h, f32=args.f32_out, stream=stream.cuda_stream)   # D2H: mix bus
        if world > 1:
            if not p2p:            # with the peer-memory exchange attached the bus that came back is already the global one
                m = mix_host.to(dev, non_blocking=True)
                dist.all_reduce(m)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * samples_per_step * e2e_steps / float(t.item())
    h2d = fh.nbytes + (on_h.nbytes + off_h.nbytes if on_h is not None else 0)
    d2h = mh.nbytes

    onbox = onbox_peaks(torch, out) if rank == 0 else None

    if rank == 0:
        peak, peak_src = load_peaks()
        med_ms = per_launch_ms[len(per_launch_ms) // 2]
        avg_ms = ms_total / args.steps
        bytes_per = (4.0 if args.f32_out else 8.0) + (wl["bytes_per"] - 8.0)
        achieved = bytes_per * samples_per_step / (avg_ms * 1e-3) / 1e9
        line = {
            "metric": "voice_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"] + (" + stereo mix bus" if want_mix else ""), "voices_per_gpu": V, "block": BLOCK,
                       "sample_rate": SR, "out_storage": "f32" if args.f32_out else "f64", "parallelism": f"voices sharded x{world}",
                       "collective": ("none" if not (world > 1 and want_mix) else
                                      "peer-memory exchange of mix[1024][2] fp64 fused into the mix-reduce kernel (CUDA IPC over NVLink, rank-ordered sum)" if p2p
                                      else "NCCL sum all-reduce of mix[1024][2] fp64 per block"),
                       "l2": "no flush needed: each step writes %.1f GB, inputs+outputs >> 126 MB L2" % (samples_per_step * (4 if args.f32_out else 8) / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         **load_traffic(args.workload + ("_mix" if want_mix else "") + ("_f32" if args.f32_out else "")),
                         "algorithmic_bytes_per_launch": bytes_per * samples_per_step,
                         "peak_source": peak_src, "kernel": "bank_kernel" if not wl["delay"] else "delay_bank_kernel",
                         "algorithmic_bytes_per_voice_sample": bytes_per, "launch_ms_avg": avg_ms, "launch_ms_median": med_ms},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "what": "mxb_bank_set_param(freq, pinned host) + mxb_bank_process(MXB_MEM_SPLIT): "
                                                 "host gates in, host mix out, voice signals materialised on the device"},
            "gpu_launches": launches, "clocks": clocks,
        }
        if onbox:
            line["roofline"]["onbox_peaks"] = onbox
            line["roofline"]["frac_of_onbox_fill"] = achieved / onbox["fill_gbs"]
        if mixdown is not None:
            line["mixdown"] = mixdown
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------- configs[3]: FFT + MFCC frames/s

MFCC_WL = dict(channels=65536, fft=1024, hop=512, filters=42, coeffs=40,
               # minimum traffic per (channel, hop) frame: 512 new fp32 samples in, 40 fp64 coefficients out
               bytes_per=512 * 4 + 40 * 8,
               desc="configs[3]: 64Ki channels maxiFFT(1024, hop 512) + maxiMFCC(42 filters, 40 coeffs), streaming, "
                    "one hop (65536 frames) per step, fused: spectra stay on chip, fp64 MFCCs written")


def cpu_mfcc_run(channels, hops, threads, kind):
    """threads x (maxiFFT + maxiMFCC per channel) over `hops` hops; returns seconds for the timed hops."""
    from maximilian_b200 import workloads as W
    from oracle import oracle_py as O
    wl = MFCC_WL
    threads = max(1, min(threads, channels // 8))      # at least 8 channels per host thread
    bounds = np.linspace(0, channels, threads + 1).astype(int)
    # 16 hops per library call: the time goes to the reference's C++ loops, not to Python call overhead
    calls, per_call = max(1, hops // 16), 16
    x = W.channel_streams(channels, per_call * wl["hop"], seed=5)
    objs = []
    for i in range(threads):
        c = int(bounds[i + 1] - bounds[i])
        if c <= 0:
            objs.append(None)
            continue
        objs.append((O.Stft(c, wl["fft"], wl["hop"], kind=kind), O.Mfcc(wl["fft"] // 2, wl["filters"], wl["coeffs"], 20.0, 20000.0, SR, kind=kind),
                     np.ascontiguousarray(x[bounds[i]:bounds[i + 1]])))

    def work(i, n):
        if objs[i] is None:
            return
        st, mf, xi = objs[i]
        for _ in range(n):
            r = st.process(xi, want=("mags",))
            mf.process(r["mags"])

    def run(n):
        ts = [threading.Thread(target=work, args=(i, n)) for i in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        return time.perf_counter() - t0
    run(1)
    dt = run(calls)
    return dt * hops / (calls * per_call)      # seconds per `hops` hops


def main_mfcc(args):
    wl = MFCC_WL
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        kind = cpu_kind(); cores = os.cpu_count() or 1
        ch = 4096
        dt = cpu_mfcc_run(ch, args.steps, cores, kind)
        v = ch * args.steps / dt
        print(json.dumps({"impl": "reference", "metric": "fft_mfcc_frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": 1, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"], "cpu_sample_channels": ch},
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                                           "sample": f"each step = one hop of {ch} channels on {cores} host threads"},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    import torch
    import torch.distributed as dist
    from maximilian_b200 import capi
    from maximilian_b200 import workloads as W
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)       # no collective on this path: channels shard, replicas only
    C, n, hop = wl["channels"], wl["fft"], wl["hop"]
    ctx = capi.Context(local, SR)
    st = capi.Stft(C, n, hop, ctx=ctx); mf = capi.Mfcc(n // 2, wl["filters"], wl["coeffs"], 20.0, 20000.0, ctx=ctx)
    base = W.channel_streams(1024, hop, seed=5 + rank)
    x_host = torch.from_numpy(np.tile(base, (C // 1024, 1))).pin_memory()           # [C][hop] planar fp32, 128 MiB
    x = x_host.to(dev)
    coeffs = torch.empty((C, 1, wl["coeffs"]), dtype=torch.float64, device=dev)
    co_host = torch.empty((C, 1, wl["coeffs"]), dtype=torch.float64).pin_memory()
    stream = torch.cuda.current_stream()

    def step():
        f = st.process_device(x.data_ptr(), hop, 1, hop, 1, mfcc=mf, coeffs=coeffs.data_ptr(), stream=stream.cuda_stream)
        assert f == 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.12)
    barrier()
    l0 = st.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    tw1 = time.perf_counter()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * C * args.steps / (ms_max * 1e-3)
    launches = st.launches - l0
    clocks = None
    if rank == 0:
        time.sleep(0.06); sampler.stop(); clocks = sampler.summary(tw0, tw1)
    # e2e: host samples in (pinned), host MFCCs out, through the C ABI (MXB_MEM_HOST)
    import ctypes as C_
    xh, ch_ = x_host.numpy(), co_host.numpy()
    nf = C_.c_int32(0)

    def e2e_step():
        capi.check(capi.lib().mxb_stft_process(st.h, xh.ctypes.data, hop, 1, hop, 1, None, None, None, None, mf.h, ch_.ctypes.data,
                                               C_.byref(nf), capi.MEM_HOST, C_.c_void_p(stream.cuda_stream)), "mxb_stft_process")
    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * C * e2e_steps / float(t.item())
    if rank == 0:
        peak, peak_src = load_peaks()
        avg_ms = ms / args.steps
        achieved = wl["bytes_per"] * C / (avg_ms * 1e-3) / 1e9
        line = {"metric": "fft_mfcc_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": wl["desc"], "channels_per_gpu": C, "parallelism": f"channels sharded x{world}", "collective": "none",
                           "l2": "per step 134 MB of samples in + 268 MB of assembly state r/w + 21 MB out: larger than the 126 MB L2"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, **load_traffic("mfcc"),
                             "algorithmic_bytes_per_launch": wl["bytes_per"] * C,
                             "peak_source": peak_src, "kernel": "stft_kernel", "algorithmic_bytes_per_frame": wl["bytes_per"],
                             "launch_ms_avg": avg_ms, "note": "ALU/shuffle bound by design of the reference transform (fp32 radix-2 + fp64 mel/DCT); "
                                                            "HBM fraction is reported as north_star asks, not expected to be high"},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(xh.nbytes), "d2h_bytes_per_step": int(ch_.nbytes),
                        "steps": e2e_steps, "what": "mxb_stft_process(MXB_MEM_HOST): pinned host samples in, fused STFT+MFCC, host MFCCs out"},
                "gpu_launches": launches, "clocks": clocks}
        if not args.no_cpu and world == 1:
            kind = cpu_kind(); cores = os.cpu_count() or 1
            ch = 4096
            dtc = cpu_mfcc_run(ch, 8, cores, kind)
            line["cpu_baseline"] = {"value": ch * 8 / dtc, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": f"{ch} channels x 8 hops, maxiFFT+maxiMFCC per channel, channels partitioned over {cores} host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
