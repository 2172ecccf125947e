#!/usr/bin/env python3
"""tools/quality_report.py -- CORDIC vs table cores on quality (the reference test benches' own metrics, scored on
the GPU by cordic_b200.score) next to speed.  Run on a GPU box:  python tools/quality_report.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cordic_b200 as zc  # noqa: E402
from cordic_b200 import score  # noqa: E402


def main():
    rows = []
    for name, kw in [("cordic p2r shipped (IW13 OW13 PW20 N16)", dict(iw=13, ow=13, xtra=2)),
                     ("cordic p2r cfg0 (IW16 OW16 PW16 N13)", dict(iw=16, ow=16, xtra=2, phase_bits=16)),
                     ("cordic p2r cfg1 (IW18 OW18 PW24 N20)", dict(iw=18, ow=18, xtra=2, phase_bits=24, nstages=20))]:
        core = zc.Cordic(**kw)
        rows.append((name, score.score_rotation(core)))
    for name, kw in [("topolar r2p shipped (IW13 OW13)", dict(iw=13, ow=13, xtra=2)),
                     ("topolar r2p cfg2 (IW16 OW16)", dict(iw=16, ow=16, xtra=2))]:
        rows.append((name, score.score_topolar(zc.Topolar(**kw))))
    st = zc.SinTable(phase_bits=17, ow=13)
    rows.append(("sintable PW17 OW13", score.score_sine(st.lookup, 17, 13)))
    qw = zc.QuarterWav(phase_bits=18, ow=24)
    rows.append(("quarterwav PW18 OW24", score.score_sine(qw.lookup, 18, 24)))
    qt = zc.QuadTbl(ow=13, phase_bits=18)
    rows.append(("quadtbl PW18 OW13", score.score_sine(qt.lookup, 18, 13)))
    for name, r in rows:
        print(json.dumps({"core": name, **{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()}}))


if __name__ == "__main__":
    main()
