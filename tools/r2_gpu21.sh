#!/bin/bash
# round 2, last check of the committed state on 1 GPU: gpu tier + smoke
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu --no-configs --no-sustained --steps 20 --warmup 5 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["roofline"]["frac"],3), round(d["e2e"]["value"],2), d["parity_spot_check"])'
