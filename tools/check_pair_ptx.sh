#!/bin/bash
# ptxas check of the CTA-pair (cta_group::2) PTX wrappers planned for the 256-row conv kernel: compiles the
# instantiation unit for sm_100a and lists the tcgen05 / TMA mnemonics found in its SASS.  No GPU needed.
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
SRC="$HERE/pytorch-segmentation_b200/csrc/experimental/pair_ptx_check.cu"
OUT="${TMPDIR:-/tmp}/pair_ptx_check.cubin"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -cubin -o "$OUT" "$SRC"
echo "ptxas accepted every wrapper; SASS mnemonics:"
/usr/local/cuda/bin/cuobjdump -sass "$OUT" | grep -oE "UTC[A-Z0-9_.]+|UTMA[A-Z0-9_.]+|UCGABAR[A-Z_.]*|SYNCS[A-Z0-9_.]+" | sort | uniq -c
