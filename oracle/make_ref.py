"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference's own files.

TEST / BASELINE INFRASTRUCTURE ONLY (nothing under theta_rrt_b200/ may touch it).

The reference (eshira/theta-rrt) is three pure-Python files; there is nothing to compile.  /root/reference exists only
in the build container, so to time the UNMODIFIED reference on the GPU box's host cores next to the CUDA path
(`bench.py`) its files are copied -- byte for byte, by this script, at build time -- into
oracle/_ref/, which is listed in .gitignore (never committed) but not in .gpurunignore (it travels with the snapshot,
like the built .so files).  `python oracle/make_ref.py` is run by __graft_entry__.build() when /root/reference is present.
"""
import hashlib
import os
import shutil
import sys

SRC = os.environ.get("THETA_RRT_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("main.py", "rrt.py", "search.py", "map1.png", "map2.png", "blank.png")


def make(verbose=False):
    if not os.path.isfile(os.path.join(SRC, "search.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    lines = []
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if not os.path.isfile(s):
            continue
        shutil.copyfile(s, d)
        lines.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(lines)} files copied from {SRC}", file=sys.stderr)
    return True


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) else 1)
