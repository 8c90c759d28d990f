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

``stage_package()`` additionally lays the UNMODIFIED reference package out as an installed tree under
``baseline/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box, where ``/root/reference`` does not
exist): ``atropos/`` with the three compiled modules next to their ``.pyx`` sources, the reference's own ``tests/``
(+ ``tests/data``, ``tests/cut``), ``bin/atropos`` and ``pytest.ini``.  This is what the base contract's
``pip install --target baseline/_ref /root/reference`` would produce, plus the test-suite.  It serves
  * ``tests/test_gpu_reference_suite.py``: the reference's own tests and CLI goldens executed with
    ``atropos.align._align`` served by ``atropos_b200`` (the drop-in, dropped in);
  * ``bench.py --impl reference`` / ``cpu_baseline``: ``atropos trim -T N`` timed on the GPU box's host cores.
Nothing of it enters the git history.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
STAGE = os.path.join(os.path.dirname(HERE), "baseline", "_ref")

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


def staged():
    """True if baseline/_ref holds an importable reference package with its compiled modules and tests."""
    need = [os.path.join(STAGE, "atropos", "__init__.py"), os.path.join(STAGE, "tests", "utils.py"),
            os.path.join(STAGE, "bin", "atropos")]
    need += [os.path.join(STAGE, os.path.dirname(rel), name + ext_suffix()) for name, rel in SOURCES.items()]
    return all(os.path.exists(p) for p in need)


def stage_package(reference="/root/reference", force=False):
    """Copy the reference package, its tests and launcher into baseline/_ref/ and drop the compiled modules in."""
    import shutil
    if not os.path.isdir(reference):
        return staged()
    if staged() and not force:
        return True
    if not all(built(n) for n in SOURCES):
        build(reference)
    os.makedirs(STAGE, exist_ok=True)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.c", "*.o")
    for sub in ("atropos", "tests", "bin"):
        dst = os.path.join(STAGE, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(reference, sub), dst, ignore=ignore)
    for f in ("pytest.ini",):
        if os.path.exists(os.path.join(reference, f)):
            shutil.copy(os.path.join(reference, f), os.path.join(STAGE, f))
    for name, rel in SOURCES.items():
        shutil.copy(os.path.join(OUT, name + ext_suffix()), os.path.join(STAGE, os.path.dirname(rel), name + ext_suffix()))
    return staged()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force, verbose=True)
    ok = stage_package(a.reference, a.force) and ok
    print("baseline/_ref:", "staged" if staged() else "NOT staged")
    print("oracle/_ref:", "ok" if ok else "INCOMPLETE", sorted(os.listdir(OUT)) if os.path.isdir(OUT) else [])
    sys.exit(0 if ok else 1)
