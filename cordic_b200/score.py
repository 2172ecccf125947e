"""Quality scoring of a core's output, mirroring the reference test benches' acceptance maths, computed on the
GPU in float64 (torch; cuFFT for the spectrum -- scoring is not the hot path).

  score_rotation  bench/cpp/cordic_tb.cpp:221-373   avg/max error, gain alpha, CNR, SFDR, pass/fail
  score_topolar   bench/cpp/topolar_tb.cpp:221-256,303-315   max phase / magnitude error, pass/fail
  score_sine      bench/cpp/quadtbl_tb.cpp:146-212   max error and SFDR of a sine generator (LUT cores too)

These let the CORDIC and the table cores be compared on spur level as well as speed (tools/quality_report.py).
"""
import math


def _sfdr_db(re, im):
    """10*log10(|bin 1|^2 / max over the other bins |.|^2) of the forward FFT (cordic_tb.cpp:342-373)."""
    import torch
    spec = torch.fft.fft(torch.complex(re, im))
    power = spec.real * spec.real + spec.imag * spec.imag
    master = float(power[1])
    power[1] = 0.0
    spur = float(power.max())
    return 10.0 * math.log10(master / spur) if spur > 0 else float("inf")


def score_rotation(core, x0=None, y0=0, flags=0, sfdr=True):
    """Sweep all 2^PW phases with a constant input vector (cordic_tb.cpp:61-80,127-178) and score."""
    import torch
    IW, OW, PW, GAIN = core.IW, core.OW, core.PW, core.GAIN
    if x0 is None:
        x0 = (1 << (IW - 1)) - 1
    n = 1 << PW
    phase = torch.arange(n, dtype=torch.int32, device="cuda")
    out = core.rotate_const(x0, y0, phase, flags=flags).to(torch.float64)
    ph = torch.arange(n, dtype=torch.float64, device="cuda") * (2.0 * math.pi / float(1 << PW))
    dx = (torch.cos(ph) * x0 - torch.sin(ph) * y0) * GAIN
    dy = (torch.sin(ph) * x0 + torch.cos(ph) * y0) * GAIN
    shift = IW + 1 - OW                                  # :238-248 (the TB is only meaningful for OW <= IW+1)
    if shift > 0:
        dx, dy = dx / float(1 << shift), dy / float(1 << shift)
    xv, yv = out[:, 0], out[:, 1]
    err2 = (dx - xv) ** 2 + (dy - yv) ** 2
    averr = math.sqrt(float(err2.sum()) / n)
    mxerr = math.sqrt(float(err2.max()))
    mag = math.sqrt(float((xv * xv + yv * yv).sum()) / n)
    sumxy = float((dx * xv + dy * yv).sum())
    sumsq = float((xv * xv + yv * yv).sum())
    scale = math.sqrt(float(x0) * x0 + float(y0) * y0)
    expected_err = core.QUANTIZATION_VARIANCE + core.PHASE_VARIANCE_RAD * scale * scale * GAIN * GAIN   # :285-286
    alpha = sumxy / sumsq
    res = {"avg_err": averr, "max_err": mxerr, "expected_err": math.sqrt(expected_err), "mag": mag, "alpha": alpha,
           "cnr_db": 10.0 * math.log10((scale * GAIN) ** 2 / (averr * averr)), "best_cnr_db": core.BEST_POSSIBLE_CNR,
           "passed": averr <= 1.5 * math.sqrt(expected_err) and mxerr <= 5.2 * math.sqrt(expected_err)
                     and abs(alpha - 1.0) <= 0.01}
    if sfdr and PW < 26:
        res["sfdr_dbc"] = _sfdr_db(xv, yv)
    return res


def score_topolar(core, flags=0):
    """The full-scale circle of topolar_tb.cpp:133-147 (two revolutions) and its scoring (:221-256,303-315)."""
    import torch
    IW, OW, PW, GAIN = core.IW, core.OW, core.PW, core.GAIN
    n = 1 << PW
    i = torch.arange(n, dtype=torch.int64, device="cuda")
    ip = (i << 1).to(torch.int32).to(torch.float64)
    ph = ip * (math.pi / float(1 << (PW - 1)))
    mg = float((1 << (IW - 1)) - 1)
    ix = (mg * torch.cos(ph)).to(torch.int32)             # C (int) truncation
    iy = (mg * torch.sin(ph)).to(torch.int32)
    mag, phase = core.topolar(torch.stack([ix, iy], dim=1).contiguous(), flags=flags)
    maxphase = float(2 ** PW)
    rad_to_phase = maxphase / math.pi / 2.0
    ep = torch.atan2(iy.to(torch.float64), ix.to(torch.float64)) * rad_to_phase
    ep = torch.where(ep < 0, ep + maxphase, ep)
    # o_phase is read back sign-extended from PW bits by the TB (:175-181)
    oph = phase.to(torch.int64)
    oph = torch.where(oph >= (1 << (PW - 1)), oph - (1 << PW), oph).to(torch.float64) if PW < 32 else oph.to(torch.float64)
    d = oph - ep
    d = torch.where(d > maxphase / 2, d - maxphase, d)
    d = torch.where(d < -maxphase / 2, d + maxphase, d)
    d = torch.where(d > maxphase / 2, d - maxphase, d)
    d = torch.where(d < -maxphase / 2, d + maxphase, d)
    mxperr = float(d.abs().max())
    emag = mg * 2.0 ** (IW - 1 - OW)
    mxverr = float((mag.to(torch.float64) - emag * GAIN).abs().max())
    exp_ph = max(1.0, math.sqrt(core.PHASE_VARIANCE_RAD) * rad_to_phase)
    return {"max_phase_err": mxperr, "max_mag_err": mxverr, "avg_phase_err": math.sqrt(float((d * d).sum()) / n),
            "phase_limit": 3.4 * exp_ph, "mag_limit": 2.0 * math.sqrt(core.QUANTIZATION_VARIANCE),
            "passed": mxperr <= 3.4 * exp_ph and mxverr <= 2.0 * math.sqrt(core.QUANTIZATION_VARIANCE)}


def score_sine(lookup, PW, OW, sfdr=True):
    """A sine generator over all 2^PW phases (quadtbl_tb.cpp:96-212): max error against sin*(2^(OW-1)-1) and the
    SFDR of (sin shifted a quarter period) + j*sin.  ``lookup`` maps a CUDA tensor of 32-bit NCO words to o_val."""
    import torch
    n = 1 << PW
    words = (torch.arange(n, dtype=torch.int64, device="cuda") << (32 - PW)).to(torch.int32)
    s = lookup(words).to(torch.float64)
    ph = torch.arange(n, dtype=torch.float64, device="cuda") * (2.0 * math.pi / float(n))
    ideal = torch.sin(ph) * float((1 << (OW - 1)) - 1)
    res = {"max_err": float((ideal - s).abs().max()), "max": int(s.max()), "min": int(s.min())}
    if sfdr and PW < 26:
        res["sfdr_dbc"] = _sfdr_db(torch.roll(s, -(n // 4)), s)
    return res
