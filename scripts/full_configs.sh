#!/bin/bash
# BASELINE.json configs at their STATED sizes through the drop-in CLI (TSV to /dev/null unless OUT is set), with the
# CLI's own split of read / prepare / upload / scan times, per-GPU balance and writer throughput (--gpu_stats), and a
# fast-vs-strict sample of >= 1e5 pairs of the same input.  Usage (GPU box):
#   bash scripts/full_configs.sh cfg3 "8 4 2"      # 50 000 x 500 all pairs, for each GPU count
#   bash scripts/full_configs.sh cfg4 8            # 200 000 x 1000, --max_kb_dist 500
#   bash scripts/full_configs.sh cfg5 8            # 1 000 000 x 2000, --rnd_sample 0.01 --seed 1
set -u
CFG=$1; GPUS=${2:-$(nvidia-smi -L | wc -l)}
D=${DATA_DIR:-/dev/shm}; OUT=${OUT:-/dev/null}
CLI=ngsld_b200/bin/ngsLD
case $CFG in
  cfg3) NS=50000; NI=500; SEED=11; FLAGS="--max_kb_dist 0"; CHECK="--s1-hi 120" ;;
  cfg4) NS=200000; NI=1000; SEED=12; FLAGS="--max_kb_dist 500"; CHECK="--head 40000 --max-kb-dist 500 --s1-hi 150" ;;
  cfg5) NS=${NS5:-1000000}; NI=2000; SEED=13; FLAGS="--max_kb_dist 0 --rnd_sample 0.01 --seed 1"; CHECK="--head 60000 --rnd-sample 0.01 --seed 1 --s1-hi 8000" ;;
  *) echo "unknown config $CFG"; exit 2 ;;
esac
# host memory: the input file (tmpfs) + the CLI's in-place genotype matrix + expected genotypes (8 B per cell) + buffers;
# a box that cannot hold the stated size gets the largest prefix that fits (and says so) instead of an OOM kill
AVAIL_GB=$(awk '/MemAvailable/{print int($2/1048576)}' /proc/meminfo)
NEED_GB=$(( NS * NI * 56 / 1000000000 + 40 ))
if [ $NEED_GB -gt $AVAIL_GB ]; then
  NS_FIT=$(( (AVAIL_GB - 40) * 1000000000 / (NI * 56) / 1000 * 1000 ))
  echo "[$CFG] only $AVAIL_GB GB of host memory available, $NEED_GB GB needed for $NS sites: scaling down to $NS_FIT sites"
  NS=$NS_FIT
fi
G=$D/$CFG.glf
s=$(date +%s%N)
[ -f $G ] || python scripts/gen_big.py $NS $NI $SEED $G
e=$(date +%s%N); echo "[$CFG] input: $NS sites x $NI individuals, $(du -h $G | cut -f1), generated in $(( (e - s) / 1000000 )) ms"
free -g | head -2
for n in $GPUS; do
  s=$(date +%s%N)
  $CLI --geno $G --probs --n_ind $NI --n_sites $NS --pos $G.pos $FLAGS --gpu_n $n --gpu_stats --verbose 0 --out $OUT ${CLI_EXTRA:-} 2> $D/$CFG.err
  rc=$?; e=$(date +%s%N)
  echo "[$CFG] $n GPU(s): rc=$rc, process start -> exit $(( (e - s) / 1000000 )) ms, out=$OUT"
  grep -E "^\[(gpu|writer|time)" $D/$CFG.err
  [ "$OUT" != "/dev/null" ] && { ls -la $OUT* | head -3; du -ch $OUT* | tail -1; rm -f $OUT $OUT.part-*; }
done
if [ "${NO_CHECK:-0}" = "0" ]; then
  echo "[$CFG] fast vs bit-faithful kernel on a sample:"
  python scripts/contract_at_scale.py --geno $G --n-sites $NS --n-ind $NI $CHECK
fi
[ "${KEEP:-0}" = "1" ] || rm -f $G $G.pos
