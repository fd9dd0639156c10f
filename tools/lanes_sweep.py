"""Time the step under different lane settings (work sets side by side, blocks per SM of the cell kernels) on the GPU box:
    python tools/lanes_sweep.py [lanes:cell_blocks:lane_min ...]
Every line must show the same state hash (columns are independent: the cut into launches changes no bit)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
combos = [a for a in sys.argv[1:] if ":" in a] or ["1:0:131072", "2:0:131072:0", "2:0:131072", "2:3:131072", "4:0:131072", "4:3:131072", "4:2:131072", "8:0:131072", "8:3:131072", "8:2:131072", "8:3:65536"]
extra = [a for a in sys.argv[1:] if ":" not in a]
for c in combos:
    lanes, blocks, lmin, *rest = c.split(":")
    env = dict(os.environ, KIDMP_LANES=lanes, KIDMP_CELL_BLOCKS=blocks, KIDMP_LANE_MIN=lmin, KIDMP_STAGGER=(rest[0] if rest else "1"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "state_hash.py"), "--steps", "6", *extra], env=env,
                       capture_output=True, text=True)
    lines = [l for l in r.stdout.strip().splitlines() if "sha256" in l] or [r.stderr[-300:]]
    print("lanes %s cell_blocks %s lane_min %-7s stagger %s %s" % (lanes, blocks, lmin, env["KIDMP_STAGGER"], lines[-1][lines[-1].find("sha256"):]), flush=True)
