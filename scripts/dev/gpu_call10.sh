#!/bin/bash
for c in 0 1 2 3 4; do
  GSK_CFG_K32=$c scripts/dev/ab.sh "dev" "C3a C3b" 2097152 2>&1 | sed "s/^/cfg32=$c /"
done
