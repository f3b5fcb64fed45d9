#!/bin/bash
# run tools/kbench.py on every tuning variant of the library found in mps_born_machines_b200/_lib/
for so in mps_born_machines_b200/_lib/libbm*.so; do
  echo "== $so"; BM_LIB_PATH=$PWD/$so python tools/kbench.py "$@" 2>&1 | tail -1
done
