"""Stage the UNMODIFIED reference package into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).

The reference (rballester/tntorch) is pure Python over torch.linalg, so there is nothing to compile: "building"
oracle/_ref is a verbatim copy of /root/reference/tntorch made by `__graft_entry__.build()` in the build container.
Nothing under oracle/_ref is committed, and nothing in the product path imports it: only bench.py's CPU legs
(`--impl reference`, `cpu_baseline`) do, to time the reference's own code on the box's host cores.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/tntorch"
DST = os.path.join(HERE, "_ref", "tntorch")


def stage(force: bool = False) -> str:
    """Copy the reference package if the source tree is here; returns the staged path ('' if unavailable)."""
    if os.path.isdir(SRC):
        stale = force or not os.path.isdir(DST) or any(
            not os.path.exists(os.path.join(DST, f)) or os.path.getmtime(os.path.join(SRC, f)) > os.path.getmtime(os.path.join(DST, f))
            for f in os.listdir(SRC) if f.endswith(".py"))
        if stale:
            shutil.rmtree(os.path.dirname(DST), ignore_errors=True)
            os.makedirs(os.path.dirname(DST), exist_ok=True)
            shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__"))
            for name in ("LICENSE",):
                if os.path.exists(os.path.join("/root/reference", name)):
                    shutil.copy(os.path.join("/root/reference", name), os.path.join(os.path.dirname(DST), name))
    return DST if os.path.isdir(DST) else ""


def load():
    """Import the staged reference as `tntorch` (None when it was never staged)."""
    import sys

    root = os.path.dirname(DST)
    if not os.path.isdir(DST):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import tntorch

    return tntorch


if __name__ == "__main__":
    print(stage(force=True) or "reference source tree not present; nothing staged")
