"""Load the compiled reference modules from ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

Two levels:

* :func:`load_native` — only the reference's compiled aligner ``_align`` (the
  Cython module built by ``oracle/build_ref.py`` from
  ``/root/reference/atropos/align/_align.pyx``).  Works wherever ``oracle/_ref/``
  is present, including the GPU box (no ``/root/reference`` needed).  This is
  what ``bench.py`` times as ``cpu_baseline.kind == "reference"``.

* :func:`load_package` — the whole reference package ``atropos`` imported from
  ``/root/reference`` (read-only, this container only) with the three compiled
  extension modules injected through ``sys.modules``.  Used to validate the
  restatements in ``oracle/`` and to generate ``tests/golden/*``.

Never imported by the product package ``atropos_b200``.
"""
import importlib.machinery
import importlib.util
import os
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("ATROPOS_REFERENCE_ROOT", "/root/reference")

_EXT = {
    "atropos.align._align": "_align",
    "atropos.io._seqio": "_seqio",
    "atropos.commands.trim._qualtrim": "_qualtrim",
}


def _so_path(short):
    return os.path.join(REF_OUT, short + sysconfig.get_config_var("EXT_SUFFIX"))


def native_available():
    return os.path.exists(_so_path("_align"))


def package_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "atropos")) and all(
        os.path.exists(_so_path(s)) for s in _EXT.values())


class _RefExtFinder:
    """meta-path finder mapping the reference's three extension-module names to oracle/_ref/*.so"""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        short = _EXT.get(fullname)
        if short is None:
            return None
        so = _so_path(short)
        if not os.path.exists(so):
            return None
        loader = importlib.machinery.ExtensionFileLoader(fullname, so)
        return importlib.util.spec_from_file_location(fullname, so, loader=loader)


def _install_finder():
    if not any(f is _RefExtFinder for f in sys.meta_path):
        sys.meta_path.insert(0, _RefExtFinder)


def load_native():
    """Return the reference's compiled ``_align`` module (Aligner, MultiAligner,
    compare_prefixes, locate).  Does not need the ``atropos`` package: ``_align.pyx``
    imports nothing from it."""
    fullname = "atropos.align._align"
    if fullname in sys.modules:
        return sys.modules[fullname]
    so = _so_path("_align")
    if not os.path.exists(so):
        raise ImportError("%s not built: run `python oracle/build_ref.py` where /root/reference exists" % so)
    if package_available():
        # let the package import own the module so both views are the same object
        _install_finder()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import atropos.align._align as mod2
        return mod2
    spec = _RefExtFinder.find_spec(fullname)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[fullname] = mod
    return mod


def load_package():
    """Import the reference package from /root/reference; returns the ``atropos`` module."""
    if not package_available():
        raise ImportError("reference package not available (needs %s and oracle/_ref/*.so)" % REFERENCE_ROOT)
    _install_finder()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import atropos  # noqa: F401
    import atropos.align  # noqa: F401
    import atropos.adapters  # noqa: F401
    import atropos.util  # noqa: F401
    import atropos.io.seqio  # noqa: F401
    return atropos
