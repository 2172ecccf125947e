#!/usr/bin/env python3
"""bench.py -- headline benchmark of the zcordic engine (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic samples.  The headline workload is
BASELINE.json configs[1]: rotation-mode CORDIC, 24-bit phase / 18-bit output, 20 stages, 2^30 samples
per GPU (weak scaling: every rank owns an independent 2^30-sample shard, no data-path collective).
Rank 0 prints ONE JSON line.  After the headline's timed region the same run measures, each with its own CUDA
events and its own slice of the clock record,
  "configs":   every other BASELINE.json configuration (random-phase cfg1, per-sample vectors, cfg2 vectoring and
               its packed-word variant, cfg3 LUT cores at both table sizes on sweeps and random phases, cfg4 NCO
               sharded over the ranks in closed form),
  "sustained": the headline held for 400 steps (what the 1 kW power cap leaves of the burst figure).
"""
import argparse
import datetime
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

# name -> (kind, samples per GPU per step, algorithmic bytes per sample [SURVEY.md §8d], options)
WORKLOADS = {
    "rotate_cfg1": ("rotate_const", 1 << 30, 12, {}),     # configs[1]  4 B phase in + 8 B (x,y) out
    "rotate_cfg1_noseed": ("rotate_const", 1 << 30, 12, {}),
    "rotate_xy_cfg1": ("rotate", 1 << 29, 20, {}),         # per-sample (x,y): 12 B in + 8 B out
    "topolar_cfg2": ("topolar", 1 << 28, 16, {}),          # configs[2]  8 B in + 8 B out
    "topolar_i16_cfg2": ("topolar_i16", 1 << 28, 12, {}),  # configs[2], packed ports: int16 x 2 in, 8 B out
    "rotate_o16_cfg0": ("rotate_const_o16", 1 << 30, 8, {}),  # packed outputs for an OW<=16 core: 4 B in + 2 x int16 out
    "sintable_p17": ("lut_sin", 1 << 30, 8, {"pw": 17, "ow": 13}),     # configs[3], the shipped table
    "sintable_p23": ("lut_sin", 1 << 30, 8, {"pw": 23, "ow": 16}),     # configs[3], the largest sintable (sw/sintable.cpp:62)
    "sintable_o16_p17": ("lut_sin_o16", 1 << 30, 6, {"pw": 17, "ow": 13}),   # packed outputs: 4 B phase in + int16 out
    "quarterwav_p18": ("lut_qwav", 1 << 30, 8, {"pw": 18, "ow": 24}),  # configs[3], the shipped table
    "quarterwav_p25": ("lut_qwav", 1 << 30, 8, {"pw": 25, "ow": 16}),  # configs[3], the largest quarterwav (sw/sintable.cpp:190)
    "nco_cfg1": ("nco", 1 << 30, 8, {}),                   # configs[4]  8 B out, no input stream
    "nco_sintable_p17": ("nco_lut", 1 << 30, 4, {"pw": 17, "ow": 13, "q": False}),     # the same NCO through rtl/sintable.v: 4 B out
    "nco_quarterwav_p18": ("nco_lut", 1 << 30, 4, {"pw": 18, "ow": 24, "q": True}),    # ... through rtl/quarterwav.v
    "quadtbl_p18": ("lut_quad", 1 << 30, 8, {}),           # rtl/quadtbl.v, PW18/OW13
}
CFG1 = dict(iw=18, ow=18, xtra=2, phase_bits=24, nstages=20)
CFG0 = dict(iw=16, ow=16, xtra=2, phase_bits=16, nstages=0)
X0, Y0 = (1 << 17) - 1, 0          # full-scale input of bench/cpp/cordic_tb.cpp:68-69
NCO_STEP = 0x01234567

# (workload, phase pattern) measured after the headline, in this order (VERDICT r1 "Next" #1)
CONFIG_PASSES = [
    ("rotate_cfg1", "random"), ("rotate_xy_cfg1", "sweep"), ("rotate_xy_cfg1", "random"),
    ("topolar_cfg2", "random"), ("topolar_i16_cfg2", "random"), ("rotate_o16_cfg0", "sweep"),
    ("sintable_p17", "sweep"), ("sintable_p17", "random"), ("sintable_p23", "sweep"), ("sintable_p23", "random"),
    ("sintable_o16_p17", "sweep"), ("sintable_o16_p17", "random"),
    ("quarterwav_p18", "sweep"), ("quarterwav_p18", "random"), ("quarterwav_p25", "sweep"), ("quarterwav_p25", "random"),
    ("nco_cfg1", "nco"), ("nco_sintable_p17", "nco"), ("nco_quarterwav_p18", "nco"),
]


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for the whole run; window(t0, t1) summarises the samples that
    arrived inside one timed region (host clock)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def window(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, power = [], None, set(), []
        for ts, r in list(self.rows):
            if ts < t0 or ts > t1 + 0.03:
                continue
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


def core_name(kind, opts):
    if kind in ("rotate_const", "rotate", "nco"):
        return "p2r IW18 OW18 WW21 PW24 NSTAGES20"
    if kind == "rotate_const_o16":
        return "p2r IW16 OW16 WW19 PW16 NSTAGES13 (configs[0]), outputs packed as int16 x 2"
    if kind == "topolar":
        return "r2p IW16 OW16 WW24 PW24 NSTAGES21"
    if kind == "topolar_i16":
        return "r2p IW16 OW16 WW24 PW24 NSTAGES21, inputs packed as int16 x 2"
    if kind in ("lut_sin", "lut_qwav"):
        return "%s PW%d OW%d" % ("sintable" if kind == "lut_sin" else "quarterwav", opts["pw"], opts["ow"])
    if kind == "lut_sin_o16":
        return "sintable PW%d OW%d, outputs packed as int16" % (opts["pw"], opts["ow"])
    if kind == "nco_lut":
        return "%s PW%d OW%d fed by the NCO accumulator" % ("quarterwav" if opts["q"] else "sintable", opts["pw"], opts["ow"])
    return "quadtbl PW18 OW13"


def config_for(workload, kind, opts, nper, phase, bytes_per):
    """The workload-defining keys, identical for the `ours` and the `reference` arm."""
    return {"workload": workload, "core": core_name(kind, opts), "samples_per_gpu_per_step": nper,
            "phase": phase if kind not in ("nco", "nco_lut") else "nco step 0x%08x" % NCO_STEP,
            "sharding": "independent shards, no data-path collective",
            "l2": "inputs+outputs per step are %.1f GiB per GPU, far larger than the 126 MB L2" % (bytes_per * nper / 2**30)}


def cpu_run(kind, n, threads, opts=None):
    """One pass of the oracle (CPU port of the reference datapath) over n synthetic samples."""
    zo = oracle()
    zo.NTHREADS = threads
    opts = opts or {}
    rng = np.random.default_rng(20261017)
    if kind in ("rotate_const", "nco"):
        rc, p = zo.derive_p2r(18, 18, 2, 24, 20)
        phase = (np.arange(n, dtype=np.uint32) & 0xFFFFFF)
        t0 = time.perf_counter()
        if kind == "nco":
            zo.nco(p, X0, Y0, 0, NCO_STEP, n)
        else:
            zo.rotate_const(p, X0, Y0, phase)
    elif kind == "rotate_const_o16":
        rc, p = zo.derive_p2r(16, 16, 2, 16, 0)
        phase = (np.arange(n, dtype=np.uint32) & 0xFFFF)
        t0 = time.perf_counter()
        zo.rotate_const(p, 32767, 0, phase)
    elif kind == "rotate":
        rc, p = zo.derive_p2r(18, 18, 2, 24, 20)
        phase = (np.arange(n, dtype=np.uint32) & 0xFFFFFF)
        xy = rng.integers(-(1 << 17), 1 << 17, size=(n, 2), dtype=np.int64).astype(np.int32)
        t0 = time.perf_counter()
        zo.rotate(p, xy, phase)
    elif kind in ("topolar", "topolar_i16"):
        rc, p = zo.derive_r2p(16, 16, 2, 0, 0)
        xy = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int64).astype(np.int32)
        t0 = time.perf_counter()
        zo.topolar(p, xy)
    elif kind == "lut_quad":
        rc, q = zo.derive_qtbl(0, 13, 2, 18)
        phase = (np.arange(n, dtype=np.uint32) & 0x3FFFF)
        t0 = time.perf_counter()
        zo.quadtbl(q, phase)
    elif kind == "nco_lut":
        tbl = zo.quarterwav(opts["pw"], opts["ow"]) if opts["q"] else zo.sintable(opts["pw"], opts["ow"])
        phase = ((np.arange(n, dtype=np.uint64) * NCO_STEP) & 0xFFFFFFFF).astype(np.uint32)
        t0 = time.perf_counter()
        (zo.lut_qwav if opts["q"] else zo.lut_sin)(opts["pw"], opts["ow"], tbl, phase)
    else:
        sin = kind in ("lut_sin", "lut_sin_o16")
        pw, ow = opts.get("pw", 17 if sin else 18), opts.get("ow", 13 if sin else 24)
        tbl = zo.sintable(pw, ow) if sin else zo.quarterwav(pw, ow)
        phase = (np.arange(n, dtype=np.uint64) * 4).astype(np.uint32)
        t0 = time.perf_counter()
        (zo.lut_sin if sin else zo.lut_qwav)(pw, ow, tbl, phase)
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


def run_reference(args, kind, nper, bytes_per, opts):
    """--impl reference: the CPU implementation of the path on this box's host cores.  The reference's
    datapath only exists as Verilog (Verilator is not in the image), so this times the oracle port of
    it (oracle/zc_oracle.c), all host threads, over the SAME workload as the `ours` arm: every step is the
    full samples_per_gpu_per_step, walked in slices of 2^26 samples so that the host never holds more than
    one slice of input and output."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    piece = min(nper, 1 << 26)
    pieces = max(1, nper // piece)

    def one_step():
        return sum(cpu_run(kind, piece, threads, opts) for _ in range(pieces))
    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    total = sum(times)
    per_step = piece * pieces
    value = per_step * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": config_for(args.workload, kind, opts, per_step, args.phase, bytes_per),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d samples/step (the whole workload, in %d slices of %d) x %d steps, oracle/zc_oracle.c "
                                   "(C port of rtl/cordic.v), pthreads over all %d host threads; only the datapath is timed "
                                   "(input synthesis is outside the clock, as it is for the GPU arm)"
                                   % (per_step, pieces, piece, args.steps, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Bench:
    """Everything the `ours` arm shares between its passes."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import cordic_b200 as zc
        self.torch, self.dist, self.zc, self.args = torch, dist, zc, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device: libzcordic has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = "cuda:%d" % self.local
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
        assert self.world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
        self.sampler = ClockSampler(self.local)
        if self.rank == 0:
            self.sampler.start()
        self.peak, self.peak_src = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # ---- one workload: inputs resident in HBM, a step() closure, a cheap self-check ------------------------------
    def make(self, workload, phase_mode, nper, flags=0, no_tail=False, nco_step=NCO_STEP):
        torch, zc, devname, rank = self.torch, self.zc, self.dev, self.rank
        kind, _, bytes_per, opts = WORKLOADS[workload]
        g = torch.Generator(device=devname); g.manual_seed(20261017 + rank)
        first = rank * nper                       # this rank's shard of the global sample stream
        w = {"kind": kind, "nper": nper, "bytes_per": bytes_per, "opts": opts, "first": first}
        pmask = 0xFFFF if kind == "rotate_const_o16" else 0xFFFFFF
        if kind in ("rotate_const", "rotate", "rotate_const_o16"):
            if phase_mode == "sweep":
                phase = (torch.arange(nper, dtype=torch.int64, device=devname) + first).bitwise_and_(pmask).to(torch.int32)
            else:
                phase = torch.randint(0, pmask + 1, (nper,), dtype=torch.int32, device=devname, generator=g)
            w["phase"] = phase
        if kind in ("rotate", "topolar"):
            lim = 1 << 17 if kind == "rotate" else 1 << 15
            w["xy"] = torch.randint(-lim, lim, (nper, 2), dtype=torch.int32, device=devname, generator=g)
        if kind == "topolar_i16":
            w["iq"] = torch.randint(-(1 << 15), 1 << 15, (nper, 2), dtype=torch.int16, device=devname, generator=g)
        if kind in ("lut_sin", "lut_qwav", "lut_quad", "lut_sin_o16"):
            if phase_mode == "sweep":
                phase = ((torch.arange(nper, dtype=torch.int64, device=devname) + first) * 4).bitwise_and_(0xFFFFFFFF).to(torch.int32)
            else:
                phase = torch.randint(-(1 << 31), 1 << 31, (nper,), dtype=torch.int64, device=devname, generator=g).to(torch.int32)
            w["phase"] = phase
            w["lut"] = (zc.SinTable(phase_bits=opts["pw"], ow=opts["ow"]) if kind in ("lut_sin", "lut_sin_o16") else
                        zc.QuarterWav(phase_bits=opts["pw"], ow=opts["ow"]) if kind == "lut_qwav" else zc.QuadTbl(ow=13, phase_bits=18))
        if kind in ("topolar", "topolar_i16"):
            w["core"] = zc.Topolar(iw=16, ow=16, xtra=2)
            w["o_mag"] = torch.empty(nper, dtype=torch.int32, device=devname)
            w["o_ph"] = torch.empty(nper, dtype=torch.int32, device=devname)
        elif kind == "lut_sin_o16":
            w["o_val16"] = torch.empty(nper, dtype=torch.int16, device=devname)
        elif kind == "nco_lut":
            w["lut"] = (zc.QuarterWav if opts["q"] else zc.SinTable)(phase_bits=opts["pw"], ow=opts["ow"])
            w["o_val"] = torch.empty(nper, dtype=torch.int32, device=devname)
        elif kind in ("lut_sin", "lut_qwav", "lut_quad"):
            w["o_val"] = torch.empty(nper, dtype=torch.int32, device=devname)
        elif kind == "rotate_const_o16":
            w["core"] = zc.Cordic(**CFG0)
            w["o_xy16"] = torch.empty((nper, 2), dtype=torch.int16, device=devname)
        else:
            w["core"] = zc.Cordic(**CFG1)
            w["o_xy"] = torch.empty((nper, 2), dtype=torch.int32, device=devname)

        def step():
            if kind == "rotate_const":
                w["core"].rotate_const(X0, Y0, w["phase"], out=w["o_xy"], flags=flags)
            elif kind == "rotate_const_o16":
                w["core"].rotate_const_o16(32767, 0, w["phase"], out=w["o_xy16"])
            elif kind == "rotate":
                w["core"].rotate(w["xy"], w["phase"], out=w["o_xy"], flags=flags & zc.F_NO_DP2A)
            elif kind == "nco":
                w["core"].nco(X0, Y0, 0, nco_step, nper, n0=first, out=w["o_xy"], flags=flags)
            elif kind == "topolar":
                w["core"].topolar(w["xy"], mag=w["o_mag"], phase=w["o_ph"], flags=zc.F_NO_TAIL if no_tail else zc.F_DEFAULT)
            elif kind == "topolar_i16":
                w["core"].topolar_i16(w["iq"], mag=w["o_mag"], phase=w["o_ph"])
            elif kind == "lut_sin_o16":
                w["lut"].lookup_o16(w["phase"], out=w["o_val16"])
            elif kind == "nco_lut":
                w["lut"].nco(0, nco_step, nper, n0=first, out=w["o_val"])
            else:
                w["lut"].lookup(w["phase"], out=w["o_val"])
        w["step"] = step
        return w

    def spot_check(self, w, phase_mode):
        """A cheap device-side consistency check of what was just timed -- never the oracle (tests/ own parity):
        the known full-sweep checksums, or the same samples through a different kernel family."""
        torch, zc = self.torch, self.zc
        kind, nper = w["kind"], w["nper"]
        m = min(nper, 1 << 22)
        if kind == "rotate_const" and phase_mode == "sweep" and nper >= (1 << 24):
            return w["o_xy"][:1 << 24].sum(dim=0, dtype=torch.int64).tolist() == [-39316, -39316]
        if kind == "rotate_const":                     # table-seeded kernel vs every stage in registers
            ref = w["core"].rotate_const(X0, Y0, w["phase"][:m], flags=zc.F_NO_SEED)
            return bool(torch.equal(ref, w["o_xy"][:m]))
        if kind == "nco":                              # NCO kernel vs explicit phases through the plain kernel
            idx = torch.arange(m, dtype=torch.int64, device=self.dev) + w["first"]
            ph = ((idx * NCO_STEP) & 0xFFFFFFFF) >> 8
            ref = w["core"].rotate_const(X0, Y0, ph.to(torch.int32), flags=zc.F_NO_SEED)
            return bool(torch.equal(ref, w["o_xy"][:m])) and (w["first"] != 0 or w["o_xy"][0].tolist() == [76313, 0])
        if kind == "rotate":                           # table-directed kernel vs plain kernel
            ref = w["core"].rotate(w["xy"][:m], w["phase"][:m], flags=zc.F_NO_SEED)
            return bool(torch.equal(ref, w["o_xy"][:m]))
        if kind == "topolar_i16":                      # packed ports vs the 32-bit port words
            mag, ph = w["core"].topolar(w["iq"][:m].to(torch.int32).contiguous())
            return bool(torch.equal(mag, w["o_mag"][:m]) and torch.equal(ph, w["o_ph"][:m]))
        if kind == "rotate_const_o16":
            ref = w["core"].rotate_const(32767, 0, w["phase"][:m])
            return bool(torch.equal(ref.to(torch.int16), w["o_xy16"][:m]))
        if kind == "nco_lut":                          # phases generated in registers vs the same phases written out
            idx = torch.arange(m, dtype=torch.int64, device=self.dev) + w["first"]
            ph = ((idx * NCO_STEP) & 0xFFFFFFFF).to(torch.int32)
            return bool(torch.equal(w["lut"].lookup(ph), w["o_val"][:m]))
        if kind == "lut_sin_o16":                      # packed outputs vs the 32-bit entry point
            return bool(torch.equal(w["lut"].lookup(w["phase"][:m]).to(torch.int16), w["o_val16"][:m]))
        if kind in ("lut_sin", "lut_qwav"):            # the lookup rule of rtl/sintable.v / rtl/quarterwav.v in torch
            pw, ow = w["opts"]["pw"], w["opts"]["ow"]
            tbl = torch.from_numpy(w["lut"].table.astype(np.int64)).to(self.dev)
            ip = (w["phase"][:m].to(torch.int64) & 0xFFFFFFFF) >> (32 - pw)
            if kind == "lut_sin":
                v = tbl[ip]
            else:
                fold = (ip >> (pw - 2)) & 1
                idx = torch.where(fold == 1, ~ip, ip) & ((1 << (pw - 2)) - 1)
                v = tbl[idx]
                v = torch.where(((ip >> (pw - 1)) & 1) == 1, -v, v) & ((1 << ow) - 1)
            v = torch.where(v >= (1 << (ow - 1)), v - (1 << ow), v)
            return bool(torch.equal(v.to(torch.int32), w["o_val"][:m]))
        return None

    def packed_e2e(self, w, steps=3):
        """End to end through the packed-port host entry points (pinned buffers, copies inside the timing): what a
        narrower port word buys on the PCIe-bound host path (VERDICT r1 "Next" #8)."""
        zc, torch = self.zc, self.torch
        kind, ne = w["kind"], min(w["nper"], 1 << 28)
        if kind == "rotate_const_o16":
            hin, hout = zc.PinnedBuffer(ne, np.uint32), zc.PinnedBuffer(ne, np.int32)     # 4 B in, 2 x int16 out
            hin.array[:] = w["phase"][:ne].cpu().numpy().view(np.uint32)
            o16 = hout.array.view(np.int16).reshape(ne, 2)

            def call():
                w["core"].rotate_const_o16_host(32767, 0, hin.array, o16, device=self.local)
            bi, bo, api = 4 * ne, 4 * ne, "zc_rotate_const_o16_host"
        else:
            hin, hout = zc.PinnedBuffer(ne, np.int32), zc.PinnedBuffer(2 * ne, np.int32)  # int16 x 2 in, mag + phase out
            hin.array[:] = w["iq"][:ne].cpu().numpy().view(np.int32).reshape(-1)
            i16 = hin.array.view(np.int16).reshape(ne, 2)

            def call():
                w["core"].topolar_i16_host(i16, hout.array[:ne], hout.array[ne:].view(np.uint32), device=self.local)
            bi, bo, api = 4 * ne, 8 * ne, "zc_topolar_i16_host"
        call()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        hin.free(); hout.free()
        return {"value": ne * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo,
                "samples_per_gpu_per_step": ne, "steps": steps, "api": api + " (pinned host buffers)"}

    def timed(self, step, steps, warmup):
        """W untimed steps, then exactly K steps between CUDA events on the launching stream, barrier +
        synchronize on both sides; returns (ms on this rank, ms max over ranks, host-clock window, launches)."""
        torch, zc = self.torch, self.zc
        for _ in range(warmup):
            step()
        self.barrier()
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = zc.launch_count()
        self.barrier()
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        self.barrier()
        t1 = time.time()
        launches = zc.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        return ms, self.max_over_ranks(ms), (t0, t1), launches

    def roofline(self, bytes_per, nper, ms, steps):
        per_step_s = ms * 1e-3 / steps
        achieved = bytes_per * nper / per_step_s / 1e9
        return {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                "algorithmic_bytes_per_sample": bytes_per, "kernel_ms": 1e3 * per_step_s}

    def config_pass(self, workload, phase_mode, min_seconds=0.25):
        """One extra configuration: >= 5 steps (enough of them to cover min_seconds of device time, so that the clock
        record has samples inside the region), own events, own clock window."""
        torch = self.torch
        kind, nper, bytes_per, opts = WORKLOADS[workload]
        if self.args.samples:
            nper = min(nper, self.args.samples)
        w = self.make(workload, phase_mode, nper)
        try:
            ms3, _, _, _ = self.timed(w["step"], 3, 3)
            steps = int(max(5, min(400, min_seconds * 1e3 / max(ms3 / 3, 1e-3))))
            steps = int(self.max_over_ranks(steps))                 # every rank runs the same count
            ms, ms_max, win, launches = self.timed(w["step"], steps, 0)
            ok = self.spot_check(w, phase_mode)
            e2e = self.packed_e2e(w) if self.world == 1 and not self.args.no_e2e and kind in ("rotate_const_o16", "topolar_i16") else None
            out = {"workload": workload, "phase": phase_mode if kind not in ("nco", "nco_lut") else "nco step 0x%08x, n0 = rank * samples_per_gpu" % NCO_STEP,
                   "core": core_name(kind, opts), "samples_per_gpu_per_step": nper, "steps": steps, "warmup": 6,
                   "value": self.world * nper * steps / (ms_max * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_max / steps,
                   "roofline": self.roofline(bytes_per, nper, ms, steps), "launches_per_step": launches / steps,
                   "parity_spot_check": ok}
            if e2e:
                out["e2e"] = e2e
            if self.rank == 0:
                out["clocks"] = self.sampler.window(*win)
            return out
        finally:
            del w
            torch.cuda.empty_cache()


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
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: 2 at N=1, 5 at N>1")
    ap.add_argument("--no-dp2a", action="store_true", help="seeded word table: IMAD + negation instead of IDP.2A (A/B)")
    ap.add_argument("--no-tail", action="store_true", help="topolar: every stage in its full form (A/B of the short late stages)")
    ap.add_argument("--no-comb", action="store_true", help="NCO: keep the block mapping for every step (A/B of the comb mapping)")
    ap.add_argument("--no-merge", action="store_true", help="byte table: keep the TS lookup separate (A/B of the merged records)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-exchange", action="store_true", help="skip the NCCL scatter/gather-inclusive figure (N>1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE configurations after the headline")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-steps", type=int, default=400)
    ap.add_argument("--configs-seconds", type=float, default=0.25,
                    help="device time each extra configuration is held for (0: the minimum of 5 steps, e.g. under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    kind, nper, bytes_per, opts = WORKLOADS[args.workload]
    if args.samples:
        nper = args.samples
    if args.impl == "reference":
        return run_reference(args, kind, nper, bytes_per, opts)

    B = Bench(args)
    torch, dist, zc = B.torch, B.dist, B.zc
    world, rank, local, devname = B.world, B.rank, B.local, B.dev

    flags = zc.F_NO_SEED if args.workload.endswith("_noseed") else zc.F_DEFAULT
    flags |= {"auto": 0, "words": zc.F_SEED_WORDS, "packed": zc.F_SEED_PACKED, "regs": zc.F_SEED_REGS}[args.seed_mode]
    if args.no_dp2a:
        flags |= zc.F_NO_DP2A
    if args.no_comb:
        flags |= zc.F_NO_COMB
    if args.no_merge:
        flags |= zc.F_NO_MERGE
    # ---- headline: synthetic inputs resident in HBM before the timed region (4-12 GiB: far larger than L2) ---------
    w = B.make(args.workload, args.phase, nper, flags=flags, no_tail=args.no_tail, nco_step=args.nco_step)
    time.sleep(0.25)
    ms, ms_max, win, launches = B.timed(w["step"], args.steps, args.warmup)
    clocks = B.sampler.window(*win) if rank == 0 else None
    value = world * nper * args.steps / (ms_max * 1e-3) / 1e9
    ok = B.spot_check(w, args.phase) if args.nco_step == NCO_STEP else None

    # ---- the headline held for 400 steps: what the board's power cap leaves of the burst figure --------------------
    sustained = None
    if not args.no_sustained:
        sms, sms_max, swin, _ = B.timed(w["step"], args.sustained_steps, 0)
        sustained = {"steps": args.sustained_steps, "value": world * nper * args.sustained_steps / (sms_max * 1e-3) / 1e9,
                     "unit": UNIT, "ms_per_step": sms_max / args.sustained_steps,
                     "roofline": B.roofline(bytes_per, nper, sms, args.sustained_steps),
                     "clocks": B.sampler.window(*swin) if rank == 0 else None,
                     "what": "the headline workload, %d back-to-back steps right after the headline's timed region" % args.sustained_steps}

    # ---- end to end through the host-buffer ABI (pinned host memory, H2D + D2H inside the timing) ---
    # N=1: zc_*_host on this GPU.  N>1: rank 0 alone calls zc_*_host_multi over all N devices of the box (one host
    # thread + pipeline per device, pinned buffers bound to each device's NUMA node), the full samples_per_gpu on each;
    # the other ranks release their device memory and wait on the host (the rendezvous store), not on the GPU.
    e2e = None
    host_kinds = ("rotate_const", "rotate", "nco", "topolar", "lut_sin", "lut_qwav", "lut_quad")
    if not args.no_e2e and kind in host_kinds:
        ne = nper
        e2e_steps = args.e2e_steps or (2 if world == 1 else 5)
        in_words = {"rotate_const": ne, "rotate": 3 * ne, "nco": 0, "topolar": 2 * ne, "lut_sin": ne, "lut_qwav": ne, "lut_quad": ne}[kind]
        out_words = {"rotate_const": 2 * ne, "rotate": 2 * ne, "nco": 2 * ne, "topolar": 2 * ne, "lut_sin": ne, "lut_qwav": ne, "lut_quad": ne}[kind]
        if world == 1:
            hin = zc.PinnedBuffer(max(in_words, 1), np.int32)
            hout = zc.PinnedBuffer(out_words, np.int32)
            if kind in ("rotate_const", "lut_sin", "lut_qwav", "lut_quad"):
                hin.array[:ne] = w["phase"][:ne].cpu().numpy()
            elif kind == "rotate":
                hin.array[:ne] = w["phase"][:ne].cpu().numpy(); hin.array[ne:] = w["xy"][:ne].cpu().numpy().reshape(-1)
            elif kind == "topolar":
                hin.array[:] = w["xy"][:ne].cpu().numpy().reshape(-1)
            core, lut = w.get("core"), w.get("lut")

            def e2e_step():
                a, o = hin.array, hout.array
                if kind == "rotate_const":
                    core.rotate_const_host(X0, Y0, a[:ne].view(np.uint32), o, device=local)
                elif kind == "rotate":
                    core.rotate_host(a[ne:], a[:ne].view(np.uint32), o, device=local)
                elif kind == "nco":
                    core.nco_host(X0, Y0, 0, args.nco_step, o, n0=w["first"], device=local)
                elif kind == "topolar":
                    core.topolar_host(a, o[:ne], o[ne:].view(np.uint32), device=local)
                else:
                    lut.lookup_host(a[:ne].view(np.uint32), o, device=local)
            e2e_step()                                  # warm-up (also faults the pinned pages in)
            B.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()                              # returns when the outputs are in host memory
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e = {"value": ne * e2e_steps / dt / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": 4 * in_words, "d2h_bytes_per_step": 4 * out_words,
                   "samples_per_gpu_per_step": ne, "steps": e2e_steps,
                   "api": "zc_%s_host (pinned host buffers, chunked H2D->kernel->D2H pipeline)" % kind}
            hin.free(); hout.free()
        elif kind == "rotate_const":
            del w
            w = None
            torch.cuda.empty_cache()
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                def run_multi(ne_):
                    devices = list(range(world))
                    core = zc.Cordic(**CFG1)
                    total = world * ne_
                    hin = zc.ShardedPinnedBuffer(total, devices, np.uint32)      # shard g bound to device g's NUMA node
                    hout = zc.ShardedPinnedBuffer(2 * total, devices, np.int32)
                    ar = hin.array
                    period = 1 << 24                                              # the global sweep: i & 0xFFFFFF
                    ar[:min(period, total)] = np.arange(min(period, total), dtype=np.uint32)
                    for s0 in range(period, total, period):
                        ar[s0:s0 + period] = ar[:min(period, total - s0)]
                    core.rotate_const_host_multi(X0, Y0, ar, hout.array, devices)   # warm-up
                    t0 = time.perf_counter()
                    for _ in range(e2e_steps):
                        core.rotate_const_host_multi(X0, Y0, ar, hout.array, devices)
                    dt = time.perf_counter() - t0
                    o = hout.array.reshape(-1, 2)
                    sums = o[:period].sum(axis=0, dtype=np.int64).tolist()
                    last = o[total - period:].sum(axis=0, dtype=np.int64).tolist()
                    r = {"value": total * e2e_steps / dt / 1e9, "unit": UNIT,
                         "h2d_bytes_per_step": 4 * total, "d2h_bytes_per_step": 8 * total,
                         "samples_per_gpu_per_step": ne_, "steps": e2e_steps, "numa": hin.placement,
                         "parity_spot_check": sums == [-39316, -39316] and last == [-39316, -39316],
                         "api": "zc_rotate_const_host_multi: one process, one host thread + H2D->kernel->D2H pipeline per "
                                "device, %d devices, pinned host shards on each device's NUMA node" % world}
                    hin.free(); hout.free()
                    return r
                try:
                    e2e = run_multi(ne)
                except Exception as e:                              # e.g. the host cannot pin 12 GiB per GPU: a quarter each
                    try:
                        e2e = run_multi(ne >> 2)
                        e2e["note"] = "full size failed (%s); ran a quarter of samples_per_gpu_per_step" % repr(e)[:120]
                    except Exception as e2:                         # never lose the main line over it
                        e2e = {"error": repr(e2)[:300]}
                store.set("zc_e2e_done", "1")
            else:
                store.wait(["zc_e2e_done"], datetime.timedelta(seconds=3000))
            B.barrier()

    # ---- N>1: the same stream held by device 0, scattered and gathered over NCCL/NVLink each step --------------
    # north_star: "NCCL over NVLink only as a trivial scatter/gather of independent chunks".  Product code: the C++
    # client cordic_b200/zcordic_bench --scatter (libzcordic_nccl: ncclGroupStart/ncclSend/ncclRecv/ncclGroupEnd, chunked
    # so that scatter(k+1), kernel(k) and gather(k-1) overlap).  Rank 0 runs it as a child process over all N devices
    # while the other ranks wait on the host.
    exchange = None
    if world > 1 and kind == "rotate_const" and not args.no_exchange:
        if w is not None:
            del w
            w = None
        torch.cuda.empty_cache()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            exe = os.path.join(ROOT, "cordic_b200", "zcordic_bench")
            try:
                r = subprocess.run([exe, "-g", str(world), "--scatter", "-l", "28", "-s", "5", "--json"],
                                   capture_output=True, text=True, timeout=600)
                exchange = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
            except Exception as e:
                exchange = {"error": repr(e)[:300]}
            store.set("zc_xchg_done", "1")
        else:
            store.wait(["zc_xchg_done"], datetime.timedelta(seconds=3000))
        B.barrier()

    # ---- the other BASELINE configurations, each with its own timed region ---------------------------------------
    configs = None
    if not args.no_configs:
        if w is not None:
            del w
            w = None
        torch.cuda.empty_cache()
        configs = []
        for wl, pm in CONFIG_PASSES:
            if wl == args.workload and pm == args.phase:
                continue
            try:
                configs.append(B.config_pass(wl, pm, min_seconds=args.configs_seconds))
            except Exception as e:                                  # noqa: BLE001 - one pass must not cost the line
                configs.append({"workload": wl, "phase": pm, "error": repr(e)[:300]})

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = cpu_threads()
        sample = min(nper, 1 << 27)
        try:
            cpu_run(kind, 1 << 22, threads, opts)
            dt_all = cpu_run(kind, sample, threads, opts)
            dt_one = cpu_run(kind, sample >> 4, 1, opts)
            cpu = {"value": sample / dt_all / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d samples of the same workload, oracle/zc_oracle.c (C port of rtl/cordic.v), %d pthreads"
                             % (sample, threads),
                   "single_thread_value": (sample >> 4) / dt_one / 1e9,
                   "verilator_path": verilator_standin() if kind == "rotate_const" else None}
        except OSError as e:
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": "oracle not built: %s" % e}

    if rank == 0:
        B.sampler.stop()
        # one dominant kernel launch per step (the auto-selected path adds a launch that returns at its probe);
        # its duration is the CUDA-event time of the timed region / steps
        roof = B.roofline(bytes_per, nper, ms, args.steps)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except Exception:
            pass
        roof.update({
            "traffic": traffic,
            "traffic_source": "profiled constant: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed "
                              "`ncu --set full` capture of this kernel (profiles/traffic.json), not measured in this run",
            "peak_source": B.peak_src, "launches_per_step": launches / args.steps,
            # `peak` is the 1:1 copy figure; a no-arithmetic kernel moving this workload's own mix (4 B read +
            # 8 B written per sample, same access shape and launch geometry) measured 6100 GB/s on this pool
            "traffic_mix_note": ("tools/membench2.cu moves 4 B in + 8 B out per sample with this kernel's access shape "
                                 "at 6100 GB/s (profiles/membench_r1.txt): frac of that = %.3f" % (roof["achieved"] / 6100.0))
            if args.workload == "rotate_cfg1" else None,
            "scope_note": ("the table-seeded kernel serves a CONSTANT input vector (the sin/cos generator of cordic_tb.cpp:61-80); "
                           "this figure is for neighbouring phases (sweep); scattered phases, per-sample vectors and the all-"
                           "stages-in-registers kernel are under `configs`") if args.workload == "rotate_cfg1" and args.phase == "sweep" else None})
        cfg = config_for(args.workload, kind, opts, nper, args.phase, bytes_per)
        if args.seed_mode != "auto":
            cfg["seed_mode"] = args.seed_mode
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": cfg, "roofline": roof,
            "cpu_baseline": cpu, "e2e": e2e, "scatter_gather": exchange, "gpu_launches": int(launches), "clocks": clocks,
            "parity_spot_check": ok, "sustained": sustained, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
