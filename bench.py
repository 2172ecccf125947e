#!/usr/bin/env python3
"""bench.py -- headline benchmark of the zcordic engine (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic samples.  The default workload is
BASELINE.json configs[1]: rotation-mode CORDIC, 24-bit phase / 18-bit output, 20 stages, 2^30 samples
per GPU (weak scaling: every rank owns an independent 2^30-sample shard, no data-path collective).
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Gsamples/s CORDIC sin/cos (24b, 20 stages)"
UNIT = "Gsamples/s"

# name -> (kind, samples per GPU per step, algorithmic bytes per sample [SURVEY.md §8d])
WORKLOADS = {
    "rotate_cfg1": ("rotate_const", 1 << 30, 12),     # configs[1]  4 B phase in + 8 B (x,y) out
    "rotate_cfg1_noseed": ("rotate_const", 1 << 30, 12),
    "rotate_xy_cfg1": ("rotate", 1 << 29, 20),         # per-sample (x,y): 12 B in + 8 B out
    "topolar_cfg2": ("topolar", 1 << 28, 16),          # configs[2]  8 B in + 8 B out
    "sintable_p17": ("lut_sin", 1 << 30, 8),           # configs[3]
    "quarterwav_p18": ("lut_qwav", 1 << 30, 8),        # configs[3]
    "nco_cfg1": ("nco", 1 << 30, 8),                   # configs[4]  8 B out, no input stream
    "quadtbl_p18": ("lut_quad", 1 << 30, 8),           # rtl/quadtbl.v, PW18/OW13
}
CFG1 = dict(iw=18, ow=18, xtra=2, phase_bits=24, nstages=20)
X0, Y0 = (1 << 17) - 1, 0          # full-scale input of bench/cpp/cordic_tb.cpp:68-69
NCO_STEP = 0x01234567


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                power.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if mx and v > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx,
                "sm_mhz_min_under_load": min(busy) if busy else None, "power_w_max": max(power) if power else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def oracle():
    from tests import zo
    return zo


def cpu_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_run(kind, n, threads):
    """One pass of the oracle (CPU port of the reference datapath) over n synthetic samples."""
    zo = oracle()
    zo.NTHREADS = threads
    rng = np.random.default_rng(20261017)
    if kind in ("rotate_const", "nco"):
        rc, p = zo.derive_p2r(18, 18, 2, 24, 20)
        phase = (np.arange(n, dtype=np.uint32) & 0xFFFFFF)
        t0 = time.perf_counter()
        if kind == "nco":
            zo.nco(p, X0, Y0, 0, NCO_STEP, n)
        else:
            zo.rotate_const(p, X0, Y0, phase)
    elif kind == "rotate":
        rc, p = zo.derive_p2r(18, 18, 2, 24, 20)
        phase = (np.arange(n, dtype=np.uint32) & 0xFFFFFF)
        xy = rng.integers(-(1 << 17), 1 << 17, size=(n, 2), dtype=np.int64).astype(np.int32)
        t0 = time.perf_counter()
        zo.rotate(p, xy, phase)
    elif kind == "topolar":
        rc, p = zo.derive_r2p(16, 16, 2, 0, 0)
        xy = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int64).astype(np.int32)
        t0 = time.perf_counter()
        zo.topolar(p, xy)
    elif kind == "lut_quad":
        rc, q = zo.derive_qtbl(0, 13, 2, 18)
        phase = (np.arange(n, dtype=np.uint32) & 0x3FFFF)
        t0 = time.perf_counter()
        zo.quadtbl(q, phase)
    else:
        pw, ow = (17, 13) if kind == "lut_sin" else (18, 24)
        tbl = zo.sintable(pw, ow) if kind == "lut_sin" else zo.quarterwav(pw, ow)
        phase = (np.arange(n, dtype=np.uint64) * 4).astype(np.uint32)
        t0 = time.perf_counter()
        (zo.lut_sin if kind == "lut_sin" else zo.lut_qwav)(pw, ow, tbl, phase)
    return time.perf_counter() - t0


def verilator_standin():
    """The reference's own unmodified bench/cpp/cordic_tb.cpp over oracle/shim (no Verilator in this
    image), 24-bit/20-stage core, 2^24-sample sweep, single thread, VCD off: whole-program wall time."""
    exe = os.path.join(ROOT, "oracle", "_ref", "cordic_tb_cfg1")
    if not os.path.exists(exe):
        return None
    t0 = time.perf_counter()
    r = subprocess.run([exe], cwd="/tmp", env=dict(os.environ, ZC_SHIM_NOTRACE="1"), capture_output=True, text=True)
    dt = time.perf_counter() - t0
    return {"what": "unmodified bench/cpp/cordic_tb.cpp over oracle/shim cycle-accurate model (stand-in: no Verilator), "
                    "2^24 ticks + scoring + SFDR FFT, 1 thread, VCD off",
            "samples": 1 << 24, "seconds": round(dt, 3), "msamples_per_s": round((1 << 24) / dt / 1e6, 3),
            "passed": r.returncode == 0 and "SUCCESS" in r.stdout}


def run_reference(args, kind, nper):
    """--impl reference: the CPU implementation of the path on this box's host cores.  The reference's
    datapath only exists as Verilog (Verilator is not in the image), so this times the oracle port of
    it (oracle/zc_oracle.c), all host threads, on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    sample = min(nper, 1 << 26)
    for _ in range(args.warmup):
        cpu_run(kind, sample, threads)
    times = [cpu_run(kind, sample, threads) for _ in range(args.steps)]
    total = sum(times)
    value = sample * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": args.workload, "samples_per_step": sample,
                                        "note": "CPU oracle port of rtl/cordic.v, bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d samples/step x %d steps, pthreads over all host threads" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rotate_cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU per step")
    ap.add_argument("--phase", default="sweep", choices=["sweep", "random"])
    ap.add_argument("--seed-mode", default="auto", choices=["auto", "words", "packed", "regs"])
    ap.add_argument("--nco-step", type=lambda v: int(v, 0), default=NCO_STEP)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-dp2a", action="store_true", help="seeded word table: IMAD + negation instead of IDP.2A (A/B)")
    ap.add_argument("--no-tail", action="store_true", help="topolar: every stage in its full form (A/B of the short late stages)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-exchange", action="store_true", help="skip the NCCL scatter/gather-inclusive figure (N>1)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    kind, nper, bytes_per = WORKLOADS[args.workload]
    if args.samples:
        nper = args.samples
    if args.impl == "reference":
        return run_reference(args, kind, nper)

    import torch
    import torch.distributed as dist
    import cordic_b200 as zc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: libzcordic has no CPU path")
    torch.cuda.set_device(local)
    devname = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(devname))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    flags = zc.F_NO_SEED if args.workload.endswith("_noseed") else zc.F_DEFAULT
    flags |= {"auto": 0, "words": zc.F_SEED_WORDS, "packed": zc.F_SEED_PACKED, "regs": zc.F_SEED_REGS}[args.seed_mode]
    if args.no_dp2a:
        flags |= zc.F_NO_DP2A
    core = zc.Cordic(**CFG1)
    # ---- synthetic inputs, resident in HBM before the timed region (4-12 GiB: far larger than L2)
    g = torch.Generator(device=devname); g.manual_seed(20261017 + rank)
    first = rank * nper                       # this rank's shard of the global sample stream
    phase = xy = None
    if kind in ("rotate_const", "rotate"):
        if args.phase == "sweep":
            phase = (torch.arange(nper, dtype=torch.int64, device=devname) + first).bitwise_and_(0xFFFFFF).to(torch.int32)
        else:
            phase = torch.randint(0, 1 << 24, (nper,), dtype=torch.int32, device=devname, generator=g)
    if kind in ("rotate", "topolar"):
        lim = 1 << 17 if kind == "rotate" else 1 << 15
        xy = torch.randint(-lim, lim, (nper, 2), dtype=torch.int32, device=devname, generator=g)
    if kind in ("lut_sin", "lut_qwav", "lut_quad"):
        if args.phase == "sweep":
            phase = ((torch.arange(nper, dtype=torch.int64, device=devname) + first) * 4).bitwise_and_(0xFFFFFFFF).to(torch.int32)
        else:
            phase = torch.randint(-(1 << 31), 1 << 31, (nper,), dtype=torch.int64, device=devname, generator=g).to(torch.int32)
        lut = (zc.SinTable(phase_bits=17, ow=13) if kind == "lut_sin" else zc.QuarterWav(phase_bits=18, ow=24)
               if kind == "lut_qwav" else zc.QuadTbl(ow=13, phase_bits=18))
    if kind == "topolar":
        vcore = zc.Topolar(iw=16, ow=16, xtra=2)
        o_mag = torch.empty(nper, dtype=torch.int32, device=devname)
        o_ph = torch.empty(nper, dtype=torch.int32, device=devname)
    elif kind in ("lut_sin", "lut_qwav", "lut_quad"):
        o_val = torch.empty(nper, dtype=torch.int32, device=devname)
    else:
        o_xy = torch.empty((nper, 2), dtype=torch.int32, device=devname)

    def step():
        if kind == "rotate_const":
            core.rotate_const(X0, Y0, phase, out=o_xy, flags=flags)
        elif kind == "rotate":
            core.rotate(xy, phase, out=o_xy, flags=flags & zc.F_NO_DP2A)
        elif kind == "nco":
            core.nco(X0, Y0, 0, args.nco_step, nper, n0=first, out=o_xy, flags=flags)
        elif kind == "topolar":
            vcore.topolar(xy, mag=o_mag, phase=o_ph, flags=zc.F_NO_TAIL if args.no_tail else zc.F_DEFAULT)
        else:
            lut.lookup(phase, out=o_val)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = zc.launch_count()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = zc.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=devname)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * nper * args.steps / (ms_max * 1e-3) / 1e9

    # ---- self-check of what was just timed: the known full-sweep checksums --------
    ok = None
    if kind in ("rotate_const", "nco") and nper >= (1 << 24) and (kind == "nco" or args.phase == "sweep"):
        if kind == "rotate_const":
            sums = o_xy[:1 << 24].sum(dim=0, dtype=torch.int64).tolist()
            ok = sums == [-39316, -39316]
        else:   # an odd NCO step visits every 24-bit phase equally often over 2^32 samples; over 2^24 the
                # sum is not pinned, so only check the first sample: phase 0 -> (76313, 0)
            ok = (first != 0) or o_xy[0].tolist() == [76313, 0]

    # ---- end to end through the host-buffer ABI (pinned host memory, H2D + D2H inside the timing) ---
    e2e = None
    if not args.no_e2e:
        ne = nper if world == 1 else min(nper, 1 << 28)
        in_words = {"rotate_const": ne, "rotate": 3 * ne, "nco": 0, "topolar": 2 * ne, "lut_sin": ne, "lut_qwav": ne, "lut_quad": ne}[kind]
        out_words = {"rotate_const": 2 * ne, "rotate": 2 * ne, "nco": 2 * ne, "topolar": 2 * ne, "lut_sin": ne, "lut_qwav": ne, "lut_quad": ne}[kind]
        hin = zc.PinnedBuffer(max(in_words, 1), np.int32)
        hout = zc.PinnedBuffer(out_words, np.int32)
        if kind in ("rotate_const", "lut_sin", "lut_qwav", "lut_quad"):
            hin.array[:ne] = phase[:ne].cpu().numpy()
        elif kind == "rotate":
            hin.array[:ne] = phase[:ne].cpu().numpy(); hin.array[ne:] = xy[:ne].cpu().numpy().reshape(-1)
        elif kind == "topolar":
            hin.array[:] = xy[:ne].cpu().numpy().reshape(-1)

        def e2e_step():
            a, o = hin.array, hout.array
            if kind == "rotate_const":
                core.rotate_const_host(X0, Y0, a[:ne].view(np.uint32), o, device=local)
            elif kind == "rotate":
                core.rotate_host(a[ne:], a[:ne].view(np.uint32), o, device=local)
            elif kind == "nco":
                core.nco_host(X0, Y0, 0, args.nco_step, o, n0=first, device=local)
            elif kind == "topolar":
                vcore.topolar_host(a, o[:ne], o[ne:].view(np.uint32), device=local)
            else:
                lut.lookup_host(a[:ne].view(np.uint32), o, device=local)
        e2e_step()                                  # warm-up (also faults the pinned pages in)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()                              # returns when the outputs are in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=devname)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * ne * args.e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": 4 * in_words, "d2h_bytes_per_step": 4 * out_words,
               "samples_per_gpu_per_step": ne, "steps": args.e2e_steps,
               "api": "zc_%s_host (pinned host buffers, chunked H2D->kernel->D2H pipeline)" % kind}
        hin.free(); hout.free()

    # ---- N>1: the same stream held by rank 0, scattered and gathered over NCCL/NVLink each step --------------
    # north_star: "NCCL over NVLink only as a trivial scatter/gather of independent chunks".  Reported next to the
    # shard-resident figure above; rank 0's NVLink port (8 B/sample coming back) bounds it, not the kernels.
    exchange = None
    if world > 1 and kind == "rotate_const" and not args.no_exchange:
        try:
            nx = min(nper, 1 << 27)
            chunk_in = torch.empty(nx, dtype=torch.int32, device=devname)
            chunk_out = torch.empty((nx, 2), dtype=torch.int32, device=devname)
            if rank == 0:
                all_in = (torch.arange(world * nx, dtype=torch.int64, device=devname)).bitwise_and_(0xFFFFFF).to(torch.int32)
                all_out = torch.empty((world * nx, 2), dtype=torch.int32, device=devname)
                ins = list(all_in.split(nx)); outs = list(all_out.split(nx))
            else:
                ins = outs = None

            def xstep():
                dist.scatter(chunk_in, ins, src=0)
                core.rotate_const(X0, Y0, chunk_in, out=chunk_out, flags=flags)
                dist.gather(chunk_out, outs, dst=0)
            xstep()
            barrier()
            x0e, x1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xsteps = 5
            x0e.record(stream)
            for _ in range(xsteps):
                xstep()
            x1e.record(stream)
            barrier()
            tx = torch.tensor([x0e.elapsed_time(x1e)], dtype=torch.float64, device=devname)
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
            xok = None
            if rank == 0:
                xok = all_out[:1 << 24].sum(dim=0, dtype=torch.int64).tolist() == [-39316, -39316] and \
                    torch.equal(all_out[:nx], all_out[(world - 1) * nx:]) if nx % (1 << 24) == 0 else None
            exchange = {"value": world * nx * xsteps / (float(tx.item()) * 1e-3) / 1e9, "unit": UNIT,
                        "samples_per_gpu_per_step": nx, "steps": xsteps, "parity": xok,
                        "what": "rank 0 owns the whole phase stream: dist.scatter -> kernel -> dist.gather (NCCL) inside the timed region"}
            del chunk_in, chunk_out
            if rank == 0:
                del all_in, all_out, ins, outs
        except Exception as e:                                   # never lose the main line over the extra figure
            exchange = {"error": repr(e)[:200]}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = cpu_threads()
        sample = min(nper, 1 << 27)
        try:
            cpu_run(kind, 1 << 22, threads)
            dt_all = cpu_run(kind, sample, threads)
            dt_one = cpu_run(kind, sample >> 4, 1)
            cpu = {"value": sample / dt_all / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d samples of the same workload, oracle/zc_oracle.c (C port of rtl/cordic.v), %d pthreads"
                             % (sample, threads),
                   "single_thread_value": (sample >> 4) / dt_one / 1e9,
                   "verilator_path": verilator_standin() if kind == "rotate_const" else None}
        except OSError as e:
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": "oracle not built: %s" % e}

    if rank == 0:
        peak, peak_src = peaks()
        # one dominant kernel launch per step (the seeded path adds a ~2 us probe and a launch that returns at its
        # gate); its duration is the CUDA-event time of the timed region / steps
        per_step_s = ms * 1e-3 / args.steps
        achieved = bytes_per * nper / per_step_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": args.workload, "core": "p2r IW18 OW18 WW21 PW24 NSTAGES20" if kind in ("rotate_const", "rotate", "nco") else kind,
                       "samples_per_gpu_per_step": nper, "phase": args.phase, "seed_mode": args.seed_mode, "sharding": "independent shards, no data-path collective",
                       "l2": "inputs+outputs per step are %.1f GiB per GPU, far larger than the 126 MB L2" % (bytes_per * nper / 2**30)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": bytes_per, "kernel_ms": 1e3 * per_step_s,
                         "launches_per_step": launches / args.steps,
                         # `peak` is the 1:1 copy figure; a no-arithmetic kernel moving this workload's own mix (4 B read +
                         # 8 B written per sample, same access shape and launch geometry) measured 6100 GB/s on this pool
                         "traffic_mix_note": ("tools/membench2.cu moves 4 B in + 8 B out per sample with this kernel's access shape "
                                              "at 6100 GB/s (profiles/membench_r1.txt): frac of that = %.3f" % (achieved / 6100.0))
                         if args.workload == "rotate_cfg1" else None},
            "cpu_baseline": cpu, "e2e": e2e, "scatter_gather": exchange, "gpu_launches": int(launches), "clocks": clocks,
            "parity_spot_check": ok,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
