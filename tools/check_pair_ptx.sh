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

# the DRAFT pair kernel itself (compile only): resource usage from ptxas and the 2-CTA mnemonics in its SASS
SRC2="$HERE/pytorch-segmentation_b200/csrc/experimental/pair_kernel_check.cu"
OUT2="${TMPDIR:-/tmp}/pair_kernel_check.cubin"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 --expt-relaxed-constexpr -Xptxas -v -cubin -o "$OUT2" "$SRC2" 2>&1 | grep -A1 "conv_gemm_tc2_pair" | grep -E "registers|spill" || true
echo "draft pair kernel compiled; SASS mnemonics of conv_gemm_tc2_pair:"
/usr/local/cuda/bin/cuobjdump -sass "$OUT2" | awk '/Function : .*conv_gemm_tc2_pair/{f=1} /Function : /{if(!/conv_gemm_tc2_pair/)f=0} f' | grep -oE "UTCHMMA[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|UTMAREDG[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|UCGABAR[A-Z_.]*" | sort | uniq -c
