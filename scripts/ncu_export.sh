#!/bin/bash
# ncu --set full capture of ONE k_step / BVC kernel launch, exported as CSV on the GPU box (the .ncu-rep embeds the
# whole cubin with source: ~75 MB, more than gpurun brings back).
#   scripts/ncu_export.sh <out name> <kernel regex> <skip launches> <driver args...>
set -u
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
rep=/tmp/$name
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $rep python scripts/prof_driver.py "$@" > gpurun_out/$name.ncu.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/$name.raw.csv.gz
ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$name.source.csv.gz
ncu -i $rep.ncu-rep --page details 2>/dev/null | gzip > gpurun_out/$name.details.txt.gz
rm -f $rep.ncu-rep
ls -la gpurun_out/$name.*
