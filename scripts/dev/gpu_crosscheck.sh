#!/bin/bash
mkdir -p gpurun_out
{
GSKRIGE_LIB=$PWD/variants/dev.so GSK_NO_COMPACT_KEYS=1 python scripts/dev/ck_crosscheck.py write /tmp/exact
GSKRIGE_LIB=$PWD/variants/dev.so python scripts/dev/ck_crosscheck.py write /tmp/ck
python scripts/dev/ck_crosscheck.py compare /tmp/exact /tmp/ck
} 2>&1 | tee gpurun_out/ck_crosscheck.log
