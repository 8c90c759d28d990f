#!/usr/bin/env python
"""Recipe: compile the REFERENCE's own Cython sources into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``atropos_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

The three extension modules of the reference (``setup.py:92-99``) are cythonized
*from the sources where they lie* under ``/root/reference`` (read-only) with all
generated files (``.c``, ``.o``, ``.so``) written to ``oracle/_ref/`` which is
git-ignored but travels to the GPU box with ``gpurun``.  No reference source is
copied into the repository.

Outputs (module names are the reference's own, so ``oracle/ref_loader.py`` can
install them into ``sys.modules`` under ``atropos.align._align`` etc.):

    oracle/_ref/_align.<abi>.so      <- atropos/align/_align.pyx      (the hot path)
    oracle/_ref/_seqio.<abi>.so      <- atropos/io/_seqio.pyx         (Sequence; used by match_to goldens)
    oracle/_ref/_qualtrim.<abi>.so   <- atropos/commands/trim/_qualtrim.pyx (import-time dependency of the trim package)

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SOURCES = {
    "_align": "atropos/align/_align.pyx",
    "_seqio": "atropos/io/_seqio.pyx",
    "_qualtrim": "atropos/commands/trim/_qualtrim.pyx",
}


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def built(name):
    return os.path.exists(os.path.join(OUT, name + ext_suffix()))


def build(reference="/root/reference", force=False, verbose=False):
    """Build every module that is missing. Returns True if all are present."""
    if not os.path.isdir(reference):
        return all(built(n) for n in SOURCES)
    os.makedirs(OUT, exist_ok=True)
    from Cython.Compiler.Main import compile as cy_compile, CompilationOptions
    inc = sysconfig.get_paths()["include"]
    cc = os.environ.get("CC", "gcc")
    for name, rel in SOURCES.items():
        so = os.path.join(OUT, name + ext_suffix())
        src = os.path.join(reference, rel)
        if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(src):
            continue
        c_file = os.path.join(OUT, name + ".c")
        opts = CompilationOptions(output_file=c_file, language_level=3)
        res = cy_compile([src], opts)
        if res.num_errors:
            raise RuntimeError("cython failed for %s" % src)
        # Same optimisation level the reference's setup.py gets from distutils (-O2).
        cmd = [cc, "-O2", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing", "-w",
               "-I", inc, c_file, "-o", so]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return all(built(n) for n in SOURCES)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force, verbose=True)
    print("oracle/_ref:", "ok" if ok else "INCOMPLETE", sorted(os.listdir(OUT)) if os.path.isdir(OUT) else [])
    sys.exit(0 if ok else 1)
