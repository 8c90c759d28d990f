import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from atropos_b200 import synth, fastq
from atropos_b200.adapters import Adapter, BACK
from atropos_b200.align import InsertAligner
from atropos_b200.util import RandomMatchProbability
n, L = 1_000_000, 150
r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3), device="cuda")
t1, t2 = synth.fastq_text(r1), synth.fastq_text(r2)
rmp = RandomMatchProbability()
kw = dict(max_error_rate=0.1, min_overlap=1, indel_cost=3, max_rmp=1e-6, match_probability=rmp)
tr = fastq.FastqPairTrimmer(Adapter(synth.TRUSEQ_R1, BACK, **kw), Adapter(synth.TRUSEQ_R2, BACK, **kw),
                            InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, match_probability=rmp, max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1),
                            max_len=L, merge_overlapping=True, merge_min_overlap=0.9, merge_error_rate=0.2)
o, st, _ = tr.trim(t1, t2)
print(st.merged, len(o[2]))
