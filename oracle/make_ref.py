#!/usr/bin/env python
"""ORACLE (test infrastructure) -- recipe that stages the UNMODIFIED reference's Python package for the GPU box.

    python oracle/make_ref.py            (run in the build container, where /root/reference exists)

/root/reference does not exist on the GPU box, but `oracle/_ref/` travels there with the repo snapshot (it is
git-ignored, NOT gpurun-ignored).  This script copies the reference's `offpolicy_rnn/` package byte for byte from
/root/reference into `oracle/_ref/offpolicy_rnn/` -- nothing is edited, nothing enters the git history.  It is used
  * by `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg: the reference's OWN `train_one_batch` on the
    box's host cores (kind "reference");
  * by `tests/golden/make_golden_gpu.py` on the GPU box: the reference's cgpt decoder with flash-attn, and its in-tree
    Triton scans, to generate fixtures / same-box baselines that cannot be produced without a GPU.
Never imported by the product path (`recurrent-offpolicy-rl_b200/`).
"""
import filecmp
import os
import shutil
import sys

SRC = "/root/reference/offpolicy_rnn"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "offpolicy_rnn")


def build(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} not present (GPU box?) -- using the staged copy" if os.path.isdir(DST)
                  else "make_ref: neither /root/reference nor oracle/_ref present")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = 0
    for root, _, files in os.walk(DST):
        for f in files:
            n += 1
            rel = os.path.relpath(os.path.join(root, f), DST)
            assert filecmp.cmp(os.path.join(SRC, rel), os.path.join(root, f), shallow=False), rel
    if verbose:
        print(f"make_ref: staged {n} files of the unmodified reference under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
