#!/usr/bin/env python3
"""Stage the reference's three host-side Python files (env.py, utils.py, parameters.py) into the git-ignored
baseline/_ref/ so that the GPU box -- which has no /root/reference -- can run the UNMODIFIED reference on top of the
drop-in shims (tests/test_reference_on_dropin.py).  Runs where /root/reference exists
(__graft_entry__.build() calls it); elsewhere it keeps whatever is already staged.  Nothing staged is tracked by git and
the product never imports it."""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("F16_REFERENCE", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
FILES = ("env.py", "utils.py", "parameters.py")


def stage(verbose=True):
    if not all(os.path.exists(os.path.join(REF, f)) for f in FILES):
        if verbose:
            print(f"reference not present at {REF}: keeping baseline/_ref as it is")
        return all(os.path.exists(os.path.join(DST, f)) for f in FILES)
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    if verbose:
        print(f"staged {', '.join(FILES)} from {REF} into baseline/_ref (git-ignored)")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
