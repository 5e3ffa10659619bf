"""Recipe for oracle/_ref: a verbatim copy of the reference's Python package (the `diffsci/` tree of /root/reference, 1.5 MB) --
TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python (no build step); "building" it means placing its sources where `oracle/refload.py` can import
them.  /root/reference does not exist on the GPU box, oracle/_ref/ does: it is git-ignored (the reference's sources never enter
this repository's history) but travels with `gpurun` like the built `.so` files.  Used by `bench.py --impl reference` and by
`bench.py`'s `cpu_baseline` leg to time the UNMODIFIED reference (`KarrasModule.propagate_white_noise`,
karras/karrasmodule.py:867-931) on the box's host cores.  Nothing in the product path imports it.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/diffsci"
DST = os.path.join(HERE, "_ref", "diffsci")


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/build_ref.py: {SRC} not present (GPU box: the shipped oracle/_ref is used as it is)")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"oracle/build_ref.py: copied {SRC} -> {DST} ({n} files)")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
