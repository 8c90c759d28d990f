#!/usr/bin/env python
"""Run the reference's OWN test-suite (baseline/_ref/tests, staged by oracle/build_ref.py) with
``atropos.align._align`` served by atropos_b200 (atropos_b200/integration.py).

    python tests/run_reference_suite.py --mode percall|batched [--sim] [pytest args ...]

--mode percall   only the module swap: every Aligner.locate / MultiAligner.locate / compare_prefixes of the reference
                 is one call through the C ABI;
--mode batched   additionally TrimPipeline.handle_records is the staged version (one GPU adapter stage per batch);
--sim            TEST-ONLY, for the GPU-less container: the engine's device functions run on the CPU (tests/simbackend.py).

Prints one line ``ATR_INTEGRATION {json}`` with the binding's counters and where `_align` came from, then exits with
pytest's return code. Test infrastructure -- the product never imports it.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGE = os.path.join(ROOT, "baseline", "_ref")


def main(argv):
    mode, sim, rest = "batched", False, []
    it = iter(argv)
    for a in it:
        if a == "--mode":
            mode = next(it)
        elif a == "--sim":
            sim = True
        else:
            rest.append(a)
    if not os.path.isdir(os.path.join(STAGE, "atropos")):
        print("ATR_INTEGRATION " + json.dumps({"error": "baseline/_ref is not staged (python oracle/build_ref.py)"}))
        return 3
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.path.insert(0, STAGE)                   # `import atropos` / `tests.utils` resolve to the staged reference
    os.chdir(STAGE)
    if sim:
        import simbackend
        simbackend.install()
    from atropos_b200 import integration
    integration.install(batched=(mode == "batched"))
    import pytest
    rc = pytest.main(["-p", "no:cacheprovider"] + rest)
    import atropos.align._align as al
    info = dict(integration.STATS)
    info.update(mode=mode, sim=sim, align_module=getattr(al, "__spec__", None) and al.__spec__.origin,
                shim=bool(getattr(al, "__atropos_b200__", False)), rc=int(rc))
    if not sim:
        from atropos_b200 import engine
        try:
            info["gpu_launches"] = engine.default_context(0).launch_count()
        except Exception as exc:
            info["gpu_launches"] = repr(exc)
    print("ATR_INTEGRATION " + json.dumps(info))
    return int(rc)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
