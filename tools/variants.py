"""Build tuning variants of the library into tools/_ab/ (git-ignored, they travel to the GPU box) and time them there:
    python tools/variants.py build [names]     (here, no GPU)
    python tools/variants.py run [names]       (on the GPU box: state hash + ms per step of every variant)
Each variant is a set of -D flags for kidmp_api.cu (launch shapes of the cell kernels)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
AB = os.path.join(ROOT, "tools", "_ab")
CLASSES = ("WARM", "ICE", "MIXNR", "FULL")


def shape(t, b, bars, classes=CLASSES):
    return ["-DKC_%s_T=%d" % (c, t) for c in classes] + ["-DKC_%s_B=%d" % (c, b) for c in classes] + \
           ["-DKC_%s_BARS=%d" % (c, bars) for c in classes]


def per(**kw):
    """per(WARM=(threads, blocks per SM, bars), ...)"""
    out = []
    for c, (t, b, bars) in kw.items():
        out += ["-DKC_%s_T=%d" % (c, t), "-DKC_%s_B=%d" % (c, b), "-DKC_%s_BARS=%d" % (c, bars)]
    return out


VARIANTS = {
    "base": [],
    "evict": ["-DKIDMP_EVICT_FIRST"],              # hand-off records stored and loaded with L1 no-allocate / L2 evict-first
    "aero443": ["-DKC_AERO_WARM_B=4", "-DKC_AERO_ICE_B=4", "-DKC_AERO_MIX_B=3"],     # launch shapes of the aerosol-aware cell kernels
    "aero442": ["-DKC_AERO_WARM_B=4", "-DKC_AERO_ICE_B=4", "-DKC_AERO_MIX_B=2"],
    "native32": ["-DKIDMP_NATIVE_F32"],          # f32 transcendentals on the SFU: what rule 2 of DESIGN.md section 4 costs
    "free": per(WARM=(256, 4, 0), ICE=(256, 4, 0), MIXNR=(256, 3, 0), FULL=(256, 3, 0)),
    "free128": per(WARM=(128, 8, 0), ICE=(128, 8, 0), MIXNR=(128, 6, 0), FULL=(128, 6, 0)),
    "ice5": per(ICE=(256, 5, 11), WARM=(256, 5, 11)),
    "ice3": per(ICE=(256, 3, 11)),
    "mix2": per(MIXNR=(256, 2, 11), FULL=(256, 2, 11)),
    "full2": per(FULL=(256, 2, 11)),
    "mix4": per(MIXNR=(256, 4, 11), FULL=(256, 4, 11)),
    "lock512": per(WARM=(512, 2, 11), ICE=(512, 2, 11), MIXNR=(384, 2, 11), FULL=(384, 2, 11)),
    "bars63": per(WARM=(256, 4, 63), ICE=(256, 4, 63), MIXNR=(256, 3, 63), FULL=(256, 3, 63)),
}

if __name__ == "__main__":
    what = sys.argv[1]
    names = [a for a in sys.argv[2:] if a in VARIANTS] or list(VARIANTS)
    if what == "build":
        from kid_b200 import build as b
        os.makedirs(AB, exist_ok=True)
        procs = []
        for n in names:
            out = os.path.join(AB, "libkidmp_%s.so" % n)
            cmd = [b.nvcc_path(), *b.flags(VARIANTS[n]), "-o", out, os.path.join(b.CSRC, "kidmp_api.cu")]
            procs.append((n, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for n, p in procs:
            o, _ = p.communicate()
            print(n, "rc", p.returncode, o[-400:] if p.returncode else "")
    else:
        for n in names:
            env = dict(os.environ, KIDMP_LIB=os.path.join(AB, "libkidmp_%s.so" % n))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "state_hash.py"), "--steps", "6"], env=env,
                               capture_output=True, text=True)
            lines = [l for l in r.stdout.strip().splitlines() if "sha256" in l] or [r.stderr[-300:]]
            print("%-16s %s" % (n, lines[-1][lines[-1].find("sha256"):]), flush=True)
