"""ORACLE-ONLY tool: write the sampler lookup tables the reference compiles in (samplers.cpp:140-397
g_strata_permutation_sets; blue_noise_samplers/*_256spp.cpp sobol/scrambling/ranking tiles, Heitz et al. 2019)
as raw bytes to buas_pathtracer_b200/data/sampler_tables.bin.  They are DATA the device samplers must index
bit-identically; the caller of bpt_set_sampler_tables() supplies them (a drop-in binding passes the reference's own
arrays), and the standalone Python harness loads this file.
Layout: perm[16384] | sobol[65536] | scramble[131072] | rank[131072]  (u8).   Run from the repo root."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_oracle  # noqa: E402

perm, sobol, scr, rank = ref_oracle.sampler_tables()
blob = perm.tobytes() + sobol.tobytes() + scr.tobytes() + rank.tobytes()
out = os.path.join(os.path.dirname(ref_oracle.HERE), "buas_pathtracer_b200", "data", "sampler_tables.bin")
with open(out, "wb") as f:
    f.write(blob)
print(out, len(blob), hashlib.sha256(blob).hexdigest())
