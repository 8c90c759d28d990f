"""atropos_b200 -- B200-native (sm_100a CUDA) adapter-alignment engine, a drop-in for the hot path of
jdidion/atropos: `atropos.align._align` (Aligner.locate, MultiAligner.locate, compare_prefixes),
`Adapter.match_to` and `InsertAligner.match_insert`. See DESIGN.md and INTEGRATION.md.

Importing this package does not touch the GPU; the first call that needs the engine loads
`libatropos_b200.so` (build: `python -m atropos_b200.build`) and fails loudly if it or a CUDA device
is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
