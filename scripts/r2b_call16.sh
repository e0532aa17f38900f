#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
L=$PWD/howtotrainyourmamlpytorch_b200/lib/libmaml_b200_legacytrace.so
for r in 1 2; do
  timeout 300 python scripts/ab_inproc.py --steps 20 --rounds 2 "" > $O/ab16_new_$r.txt 2>&1
  MAML_B200_LIB=$L timeout 300 python scripts/ab_inproc.py --steps 20 --rounds 2 "" > $O/ab16_legacy_$r.txt 2>&1
done
for f in $O/ab16_new_1.txt $O/ab16_legacy_1.txt $O/ab16_new_2.txt $O/ab16_legacy_2.txt; do echo $f; tail -1 $f; done
