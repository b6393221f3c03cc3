"""Recipe for `baseline/_ref/`: the UNMODIFIED reference decoder modules, made available to the GPU box.

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE (same rule as the rest of `oracle/`): only `bench.py`'s
`--impl reference` / `cpu_baseline` legs and `__graft_entry__.build()` use it.

The reference is a pure-Python script tree without `setup.py` / `pyproject.toml`, so `pip install
/root/reference` is not applicable.  The three files the decoder path needs

    Captioning_models/attention.py
    Captioning_models/Depth_caption_model/depth_models.py       (imports only numpy, torch, attention)
    Captioning_models/Depth_caption_model/depth_train.py        (NOT copied: its loss lines :210-221 are
                                                                 restated by the caller in bench.py)

are copied byte for byte from the read-only checkout into `baseline/_ref/` -- a directory that is
git-ignored (nothing of the reference enters the history) but not gpurun-ignored, so it travels to the
GPU box next to the built `libdic.so`.  `/root/reference` does not exist there.  Run here, in the build
container:

    python oracle/make_ref.py            # idempotent; prints what it did

`load_reference()` returns the imported reference module namespace (or None when `baseline/_ref` is
absent), `/root/reference` first when it exists.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("DIC_REFERENCE", "/root/reference")
REF_DST = os.path.join(ROOT, "baseline", "_ref")
FILES = (
    "Captioning_models/attention.py",
    "Captioning_models/Depth_caption_model/depth_models.py",
)


def make(verbose: bool = True) -> bool:
    """Copy the reference decoder files into baseline/_ref (no-op when the checkout is absent)."""
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"[make_ref] {REF_SRC} not present: keeping whatever is in {REF_DST}")
        return os.path.isdir(REF_DST)
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print(f"[make_ref] {src} -> {dst}")
    return True


def load_reference():
    """-> (namespace with Soft_Attention, CD_RNNDecoderWithSoftAttention, ..., origin path) or (None, None)."""
    for base in (REF_SRC, REF_DST):
        if os.path.isfile(os.path.join(base, FILES[1])):
            if base not in sys.path:
                sys.path.insert(0, base)
            sys.dont_write_bytecode = True          # the checkout is read-only
            mod = importlib.import_module("Captioning_models.Depth_caption_model.depth_models")
            return mod, base
    return None, None


if __name__ == "__main__":
    ok = make()
    mod, origin = load_reference()
    print("reference importable from", origin if ok and mod is not None else None)
