#!/bin/bash
# BASELINE config 3 at full size through the drop-in CLI: 50 000 SNPs x 500 individuals, all pairs (1 249 975 000 rows),
# TSV to /dev/null (≈ 90 GB of text), all visible GPUs.  Usage: bash scripts/full_config3_cli.sh   (GPU box)
set -e
D=/tmp/cfg3; mkdir -p $D
python - <<PY
import sys; sys.path.insert(0, "tests/golden")
import gen_synth
GL, pos = gen_synth.synth_fast(50000, 500, 11)
gen_synth.write("$D/c3.glf", GL, pos)
PY
ls -la $D
s=$(date +%s%N)
ngsld_b200/bin/ngsLD --geno $D/c3.glf --probs --n_ind 500 --n_sites 50000 --pos $D/c3.glf.pos --max_kb_dist 0 --verbose 0 --gpu_stats --out /dev/null
e=$(date +%s%N); echo "B200 CLI, config 3, process start -> output closed: $(( (e - s) / 1000000 )) ms"
