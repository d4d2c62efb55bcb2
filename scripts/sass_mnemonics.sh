#!/bin/bash
# Regenerates profiles/rNN_sass_mnemonics.txt from the shipped library: usage scripts/sass_mnemonics.sh r02
set -e
cd "$(dirname "$0")/.."
out=profiles/${1:-r02}_sass_mnemonics.txt
so=distributedjets.jl_b200/libdjets_b200.so
{
  echo "# SASS mnemonics of $so (cuobjdump -sass, sm_100a), counted over all kernel instantiations"
  echo "# kernels: $(cuobjdump -sass $so | grep -c 'Function :') instantiations; arch: $(cuobjdump -lelf $so | grep -o 'sm_[0-9a]*' | sort -u | tr '\n' ' ')"
  cuobjdump -sass $so | grep -oE '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z][A-Z0-9_.]+' | sed -E 's/.*\s([A-Z][A-Z0-9_.]+)$/\1/' \
    | grep -E '^(FFMA|DFMA|LDG|STG|LDS|STS|SHFL|UBLKCP|UBLKPF|SYNCS|PREEXIT|ACQBULK|ATOMG|ATOM|RED|LDGSTS|LDGDEPBAR|DEPBAR|CCTL|MEMBAR|HMMA|UTC|LDTM|STTM|NANOSLEEP|BRX|FENCE|ERRBAR|UTMA)' \
    | sort | uniq -c | sort -rn
  echo "# no HMMA / UTC*MMA / LDTM / STTM above: the path is HBM-bound CUDA-core work (DESIGN.md section 4)"
} > $out
head -50 $out
