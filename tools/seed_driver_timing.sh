#!/bin/bash
# where the wall clock of the seed-enabled driver goes on the configs[1] FASTA (MM2GB_VERBOSE timing of the glue)
cd /tmp && python - <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import __graft_entry__ as e
pkg = e.load_package()
from mm2gb_b200 import synth
import numpy as np
d = "/tmp/mm2gb_driver_ont"; os.makedirs(d, exist_ok=True)
ref = synth.simulate_reference(100_000_000, seed=1)
rds = synth.simulate_reads(ref, 10000, 10000, 100000, seed=2, err=0.10)
cl = 25_000_000
synth.write_fasta(d + "/ref.fa", [ref[i * cl:(i + 1) * cl] for i in range(4)], prefix="ref")
synth.write_fasta(d + "/reads.fa", rds, prefix="read")
PY
cd /tmp/mm2gb_driver_ont
for t in 4 8; do
S0=$(date +%s.%N); env MM2GB_GPU_SEED=1 MM2GB_VERBOSE=1 $GRAFT_REPO_ROOT/oracle/_ref/minimap2_b200_seed -t $t -x map-ont --max-chain-skip=2147483647 ref.fa reads.fa 2> err_$t.txt > out_$t.paf
echo "threads $t wall $(echo "$(date +%s.%N) - $S0" | bc) s"; grep -E "mm2gb|Real time|loaded/built|mapped" err_$t.txt | cut -c1-160 | head -30
done
